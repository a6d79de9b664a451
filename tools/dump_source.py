"""Writes the generated translation unit (model section + bdf_kernel.cuh) of a benchmark model to a file,
so it can be inspected / compiled offline:  nvcc -gencode arch=compute_100a,code=sm_100a -cubin -Xptxas -v -lineinfo --fmad=false <file>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
which = sys.argv[1] if len(sys.argv) > 1 else "mapk"
out = sys.argv[2] if len(sys.argv) > 2 else f"/tmp/{which}_tu.cu"
if which == "mapk":
    m, vary, nominal = H.mapk_scan_model()
    b = so.Batch(m, so.settings(1000.0, 100, 1e-9, 1e-6), vary, store=list(range(m.neq)))
else:
    raise SystemExit("unknown model")
open(out, "w").write(b.source())
print(out)
