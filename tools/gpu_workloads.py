"""Times the BASELINE.json configurations other than the headline one (parity-test cases, not bench lines) and
checks a sample of each against the oracle.  usage: gpu_workloads.py [name:nsets ...]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H

specs = sys.argv[1:] or ["repressilator:100000", "events:100000", "goodwin:100000", "huang:20000"]
for spec in specs:
    name, n = spec.split(":"); n = int(n)
    w = H.workload(name)
    m = w["model"]
    store = list(range(m.neq))
    b = so.Batch(m, w["opt"], w["vary"], store=store)
    vals = H.workload_values(w, n, seed=77)
    b.set_values(vals)
    ms = [b.integrate(n) for _ in range(2)]
    st = b.status(n); stats = b.stats(n)
    ns = min(n, 256)
    ref = H.oracle_run(m, w["opt"], w["vary"], vals[:ns], w["nout"], compiled_so=H.compile_c_model(m, name), threads=os.cpu_count() or 1)
    gpu = b.results(ns)
    rep = H.parity_report(gpu, np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), stats[:, :ns].T, ref["stats"], w["rtol"], w["atol"])
    t0 = time.perf_counter()
    nref = min(n, 2048)
    H.oracle_run(m, w["opt"], w["vary"], vals[:nref], w["nout"], backend=H.REF if H.have_ref() else H.PORT, compiled_so=H.compile_c_model(m, name), threads=os.cpu_count() or 1)
    cpu = nref / (time.perf_counter() - t0)
    print(json.dumps(dict(workload=name, n=n, neq=m.neq, nout=w["nout"], ms=min(ms), traj_per_s=n / (min(ms) * 1e-3), failed=int((st != 0).sum()),
                          nst=float(stats[0].mean()), nfe=float(stats[1].mean()), same_path=rep["frac_same_path"], within=rep["frac_within"],
                          max_ratio=rep["max_ratio"], cpu_traj_per_s=cpu, cores=os.cpu_count(), info=b.kernel_info())), flush=True)
    b.close()
