import os, sys, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
from test_frontend_models import TIME_MODEL
p = os.path.join(tempfile.mkdtemp(), "forced.xml"); open(p, "w").write(TIME_MODEL)
m = so.Model(p)
vary = [m.index("w"), m.index("k"), m.index("x")]
rng = np.random.default_rng(8)
vals = np.stack([rng.uniform(0.3, 2.0, 500), rng.uniform(0.2, 3.0, 500), rng.uniform(0.5, 3, 500)], axis=1)
opt = so.settings(7.0, 70, 1e-9, 1e-6)
b = so.Batch(m, opt, vary); b.set_values(vals); b.integrate(500)
out, stats = b.results(500), b.stats(500).T
ref = H.oracle_run(m, opt, vary, vals, 70, compiled_so=H.compile_c_model(m, "forced"), threads=8)
want = np.transpose(ref["out"], (1, 2, 0))
tol = np.maximum(1e-5 * np.abs(want), 1e-8)
r = np.abs(out - want) / tol
i = np.unravel_index(np.argmax(r), r.shape)
print("names", m.names, "worst", i, r[i], out[i], want[i], "stats gpu", stats[i[2]], "ref", ref["stats"][i[2]])
print("per value max ratio", r.max(axis=(0, 2)))
print("sets with different stats", int((stats != ref["stats"]).any(axis=1).sum()))
