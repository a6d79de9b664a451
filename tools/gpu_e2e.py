"""End-to-end (host buffers) throughput of ODESBatch_runHost for several chunk sizes (development tool)."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
chunks = [int(x) for x in sys.argv[2:]] or [65536, 131072, 262144, 524288, 1048576]
L = so.lib()
w = H.workload("mapk"); m = w["model"]; nv = len(w["vary"])
b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
b.reserve(n)
b.generate_values(n, 20261019, 0, w["nominal"], -np.ones(nv), np.ones(nv))
rows = b.nrows * b.nstore
hv, hr, hs = L.ODES_hostAlloc(n * nv * 8), L.ODES_hostAlloc(rows * n * 8), L.ODES_hostAlloc(n * 4)
vals = np.ctypeslib.as_array(C.cast(hv, C.POINTER(C.c_double)), shape=(n, nv))
res = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_double)), shape=(n, b.nrows, b.nstore))
hst = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_int)), shape=(n,))
vals[:] = b.get_values(n)
for ch in chunks:
    b.run_host(vals, res, hst, chunk=ch)
    t0 = time.perf_counter()
    for _ in range(3):
        b.run_host(vals, res, hst, chunk=ch)
    dt = (time.perf_counter() - t0) / 3
    print(json.dumps(dict(chunk=ch, ms=dt * 1e3, traj_per_s=n / dt)), flush=True)
ms = [b.integrate(n) for _ in range(2)]
print(json.dumps(dict(kernel_ms=min(ms), traj_per_s=n / (min(ms) * 1e-3))))
