"""Bit-for-bit comparison of whole batches (not samples) with the vendored CVODE / CVODES on the host threads.
usage: gpu_parity_large.py [huang:20000 mapk:50000 sens:mapk:10000 ...]   (development / evidence tool)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H

threads = os.cpu_count() or 1
for spec in sys.argv[1:] or ["huang:20000", "mapk:50000", "sens:mapk:10000"]:
    parts = spec.split(":")
    sens = parts[0] == "sens"
    name, n = parts[-2], int(parts[-1])
    w = H.workload(name, sensitivity=1, sens_ids=["K1", "Ki", "V1"]) if sens else H.workload(name)
    m = w["model"]
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    vals = H.workload_values(w, n, seed=20261020)
    b.set_values(vals)
    ms = b.integrate(n)
    out, status, stats = b.results(n), b.status(n), b.stats(n)
    t0 = time.perf_counter()
    if sens:
        tp = H.tpoints(w["opt"], w["nout"])
        ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], threads=threads, tag=name + "_large_sens")
        same_out = np.array_equal(np.transpose(out, (2, 0, 1)), ref["out"], equal_nan=True)
        same_sens = bool(np.array_equal(b.sensitivities(n), ref["sens"], equal_nan=True))
        same_stats = bool(np.array_equal(stats.T, ref["stats"][:, :7]))
    else:
        ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF, compiled_so=H.compile_c_model(m, name), threads=threads)
        same_out = np.array_equal(out, np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), equal_nan=True)
        same_sens = None
        same_stats = bool(np.array_equal(stats.T, ref["stats"]))
    dt = time.perf_counter() - t0
    print(json.dumps(dict(workload=("sens:" if sens else "") + name, sets=n, gpu_ms=ms, traj_per_s=n / ms * 1e3, cpu_s=dt, cpu_threads=threads,
                          checker="vendored CVODES 2.3.0" if sens else "vendored CVODE 2.3.0", states_bit_identical=bool(same_out), sensitivities_bit_identical=same_sens,
                          counters_identical=same_stats, status_identical=bool(np.array_equal(status, ref["status"])), failed=int((status != 0).sum()))), flush=True)
    b.close()
