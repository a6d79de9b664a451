"""Forward-sensitivity kernel (simultaneous corrector in the warp-per-trajectory layout): device-resident throughput,
parity of a sample against the reference's vendored CVODES 2.3.0 (oracle/_ref/libcvodes_ref.so, test infrastructure) and
that CPU path timed on all host threads beside it.
usage: gpu_sens.py workload nsets ids|all [reps] [cpu_sample]   e.g.  gpu_sens.py mapk 100000 K1,Ki,MAPK,MAPK_P 2 2048"""
import json, os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
name = sys.argv[1]; n = int(sys.argv[2]); ids = None if sys.argv[3] == "all" else sys.argv[3].split(",")
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ncpu = int(sys.argv[5]) if len(sys.argv) > 5 else 0
w = H.workload(name, sensitivity=1, sens_ids=ids)
m = w["model"]
b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
vals = H.workload_values(w, n, seed=77)
b.set_values(vals)
ms = [b.integrate(n) for _ in range(reps)]
st = b.stats(n)
line = {"workload": name, "nsens": b.nsens, "sens_ids": ids or "all constants", "n": n, "neq": m.neq, "nout": w["nout"], "ms": min(ms), "traj_per_s": n / min(ms) * 1e3,
        "failed": int((b.status(n) != 0).sum()), "nst": float(st[0].mean()), "nni": float(st[4].mean()),
        "cfg": {k: v for k, v in os.environ.items() if k.startswith("ODES_")}, "info": b.kernel_info()}
if ncpu > 0 and H.have_sens_ref():
    ncpu = min(ncpu, n)
    threads = os.cpu_count() or 1
    tp = H.tpoints(w["opt"], w["nout"])
    H.sens_ref_run(m, w["opt"], b, w["vary"], vals[:min(ncpu, 64)], tp, 10000, w["rtol"], w["atol"], threads=threads, tag=name + "_sens")   # builds the model
    t0 = time.perf_counter()
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals[:ncpu], tp, 10000, w["rtol"], w["atol"], threads=threads, tag=name + "_sens")
    dt = time.perf_counter() - t0
    out, sens = b.results(ncpu), b.sensitivities(ncpu)
    line.update({"cpu_traj_per_s": ncpu / dt, "cores": threads, "cpu_kind": "reference: vendored CVODES 2.3.0 + emitted C rhs / Jacobian / sense_f",
                 "cpu_sample": ncpu, "states_bit_identical": bool(np.array_equal(np.transpose(out, (2, 0, 1)), ref["out"], equal_nan=True)),
                 "sensitivities_bit_identical": bool(np.array_equal(sens, ref["sens"], equal_nan=True)),
                 "same_step_sequence": bool(np.array_equal(st[:, :ncpu].T, ref["stats"][:, :7])),
                 "max_ratio_sens": H.parity(sens, ref["sens"], w["rtol"], w["atol"])})
print(json.dumps(line))
