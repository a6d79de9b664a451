"""Lane layout vs warp layout on models of growing size (chain models of tests/test_frontend_models.py and the
BASELINE models): where is the crossover?  usage: gpu_layout_crossover.py [nsets]"""
import os, sys, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
from test_frontend_models import chain_model_xml
nsets = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
def run(tag, m, opt, vary, vals):
    res = {}
    for warp in (0, 1):
        os.environ["ODES_WARP"] = str(warp)
        try:
            b = so.Batch(m, opt, vary, store=list(range(m.neq)))
            b.set_values(vals)
            ms = min(b.integrate(len(vals)) for _ in range(2))
            res[warp] = (len(vals) / ms * 1e3, b.kernel_info(), int((b.status(len(vals)) != 0).sum()))
            b.close()
        except Exception as e:
            res[warp] = (0.0, str(e)[:80], -1)
    os.environ.pop("ODES_WARP")
    print(f"{tag:14s} n={m.neq:3d} lane {res[0][0]:12.0f} traj/s  warp {res[1][0]:12.0f} traj/s   lane {res[0][1]}  warp {res[1][1]}", flush=True)
for n in (4, 8, 10, 12, 14, 16, 20, 24, 32):
    p = os.path.join(tempfile.mkdtemp(), f"chain{n}.xml"); open(p, "w").write(chain_model_xml(n))
    m = so.Model(p)
    vary = [m.index(f"v{i}") for i in range(1, n)] + [m.index("K")]
    rng = np.random.default_rng(n)
    vals = m.base_values()[vary][None, :] * np.exp2(rng.uniform(-1, 1, size=(nsets, len(vary))))
    run(f"chain{n}", m, so.settings(30.0, 30, 1e-9, 1e-6), vary, vals)
for name in ("mapk", "goodwin", "huang"):
    w = H.workload(name)
    run(name, w["model"], w["opt"], w["vary"], H.workload_values(w, nsets, seed=5))
