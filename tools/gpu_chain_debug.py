"""Where do the chain models (tests/test_frontend_models.py) differ from the oracle on the GPU?"""
import os, sys, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
from test_frontend_models import chain_model_xml
for n in [int(a) for a in sys.argv[1:]] or [3, 12, 17]:
    p = os.path.join(tempfile.mkdtemp(), f"chain{n}.xml"); open(p, "w").write(chain_model_xml(n))
    m = so.Model(p)
    vary = [m.index(f"v{i}") for i in range(1, n)] + [m.index("K")]
    rng = np.random.default_rng(n); ns = 600
    vals = m.base_values()[vary][None, :] * np.exp2(rng.uniform(-1, 1, size=(ns, len(vary))))
    opt = so.settings(30.0, 30, 1e-9, 1e-6)
    b = so.Batch(m, opt, vary); b.set_values(vals); b.integrate(ns)
    out, st, stats = b.results(ns), b.status(ns), b.stats(ns).T
    ref = H.oracle_run(m, opt, vary, vals, 30, compiled_so=H.compile_c_model(m, f"chain{n}"), threads=8)
    want = np.transpose(ref["out"], (1, 2, 0))
    print("chain", n, b.kernel_info(), "status ok", (st == 0).all(), "stats equal", np.array_equal(stats, ref["stats"]))
    d = out != want
    print(" differing entries", int(d.sum()), "of", d.size, "sets", int(d.any(axis=(0, 1)).sum()), "rows", np.nonzero(d.any(axis=(1, 2)))[0][:10],
          "values", np.nonzero(d.any(axis=(0, 2)))[0])
    if d.any():
        r, k, s = [a[0] for a in np.nonzero(d)]
        print(" first: row", r, "value", k, "set", s, "gpu", repr(out[r, k, s]), "oracle", repr(want[r, k, s]), "names", m.names() if hasattr(m, "names") else "")
        rel = np.abs(out - want)[d] / np.maximum(np.abs(want[d]), 1e-300)
        print(" max rel diff", rel.max(), "median", np.median(rel))
        print(" set", s, "gpu row", out[r, :, s], "\n oracle row", want[r, :, s])
    b.close()
