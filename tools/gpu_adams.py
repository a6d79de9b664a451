"""Adams-Moulton (CvodeMethod 1) in the warp kernel: throughput and parity of a sample against the vendored CVODE with CV_ADAMS.
usage: gpu_adams.py workload nsets [cpu_sample]"""
import json, os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
name = sys.argv[1]; n = int(sys.argv[2]); ncpu = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w = H.workload(name, method=1)
m = w["model"]
b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
vals = H.workload_values(w, n, seed=77)
b.set_values(vals)
ms = [b.integrate(n) for _ in range(2)]
st = b.stats(n)
line = {"workload": name, "method": "Adams-Moulton (orders 1..12)", "n": n, "neq": m.neq, "ms": min(ms), "traj_per_s": n / min(ms) * 1e3, "failed": int((b.status(n) != 0).sum()),
        "nst": float(st[0].mean()), "nfe": float(st[1].mean()), "info": b.kernel_info()}
if ncpu > 0 and H.have_ref():
    cores = os.cpu_count() or 1
    cso = H.compile_c_model(m, name)
    t0 = time.perf_counter()
    ref = H.oracle_run(m, w["opt"], w["vary"], vals[:ncpu], w["nout"], backend=H.REF, compiled_so=cso, threads=cores, adams=True)
    dt = time.perf_counter() - t0
    got = b.results(ncpu)
    line.update({"cpu_traj_per_s": ncpu / dt, "cores": cores, "cpu_kind": "reference: vendored CVODE 2.3.0 with CV_ADAMS + emitted C rhs / Jacobian", "cpu_sample": ncpu,
                 "bit_identical": bool(np.array_equal(got, np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), equal_nan=True)),
                 "same_step_sequence": bool(np.array_equal(st[:, :ncpu].T, ref["stats"]))})
print(json.dumps(line))
