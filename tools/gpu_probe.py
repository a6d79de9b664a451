"""GPU probe: smoke + FP64 peak + MAPK scan timing at several batch sizes (development tool)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so
import helpers as H
import __graft_entry__ as g

sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["10000", "100000", "1000000"])]
L = so.lib()
print("devices", L.ODES_deviceCount())
try:
    g.smoke()
except AssertionError as e:
    print("SMOKE ASSERT", str(e)[:300])
print("fp64 peak TFLOP/s", L.ODES_measureFP64Peak(0, 5))
m, vary, nominal = H.mapk_scan_model()
opt = so.settings(1000.0, 100, 1e-9, 1e-6)
b = so.Batch(m, opt, vary, store=list(range(m.neq)))
print("kernel", b.kernel_info())
lo, hi = -np.ones(len(vary)), np.ones(len(vary))
for n in sizes:
    b.reserve(n)
    b.generate_values(n, 20261019, 0, nominal, lo, hi)
    ms = [b.integrate(n) for _ in range(3)]
    st = b.status(n); stats = b.stats(n)
    print(json.dumps(dict(n=n, ms=ms, traj_per_s=n / (min(ms) * 1e-3), failed=int((st != 0).sum()),
                          mean_stats=dict(zip("nst nfe nsetups nje nni ncfn netf".split(), [float(x) for x in stats.mean(axis=1)])),
                          max_nst=int(stats[0].max()))))
