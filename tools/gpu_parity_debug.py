import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
m, vary, nominal = H.mapk_scan_model()
rtol, atol, nout = 1e-6, 1e-9, 100
opt = so.settings(1000.0, nout, atol, rtol)
b = so.Batch(m, opt, vary, store=list(range(m.neq)))
vals = H.log_uniform_sets(nominal, 256, seed=7)
n = b.set_values(vals); b.integrate(n)
gpu = b.results(n); st = b.status(n); gs = b.stats(n).T
ref = H.oracle_run(m, opt, vary, vals, nout)
want = np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0))
tol = np.maximum(10*rtol*np.abs(want), 10*atol)
ratio = (np.abs(gpu-want)/tol).max(axis=(0,1))
print("sets with ratio>1:", int((ratio>1).sum()), "ratio>1e-3:", int((ratio>1e-3).sum()), "ratio>1e-6:", int((ratio>1e-6).sum()))
print("stats mismatching sets:", int((gs != ref["stats"]).any(axis=1).sum()))
w = int(np.argmax(ratio)); print("worst set", w, "ratio", ratio[w], "gpu stats", gs[w], "oracle stats", ref["stats"][w])
r_t = (np.abs(gpu-want)/tol)[:, :, w].max(axis=1); print("first row with ratio>1e-6:", int(np.argmax(r_t>1e-6)), "ratios by row (every 10th):", r_t[::10])
good = int(np.argmin(ratio)); print("best set", good, ratio[good], gs[good], ref["stats"][good])
print("median ratio", np.median(ratio))
