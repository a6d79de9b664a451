"""Throughput of the sensitivity kernel (simultaneous corrector in the warp-per-trajectory layout).
usage: gpu_sens.py workload nsets ids|all [reps]   e.g.  gpu_sens.py mapk 100000 K1,Ki,MAPK,MAPK_P"""
import json, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
name = sys.argv[1]; n = int(sys.argv[2]); ids = None if sys.argv[3] == "all" else sys.argv[3].split(",")
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
w = H.workload(name, sensitivity=1, sens_ids=ids)
m = w["model"]
b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
b.set_values(H.workload_values(w, n, seed=77))
ms = [b.integrate(n) for _ in range(reps)]
st = b.stats(n)
print(json.dumps({"workload": name, "nsens": b.nsens, "n": n, "ms": min(ms), "traj_per_s": n / min(ms) * 1e3, "failed": int((b.status(n) != 0).sum()),
                  "nst": float(st[0].mean()), "nni": float(st[4].mean()), "cfg": {k: v for k, v in os.environ.items() if k.startswith("ODES_")}, "info": b.kernel_info()}))
