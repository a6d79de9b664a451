"""Throughput of the event-detection modes on the events model (BASELINE configs[3]): reference semantics (output times),
StepResolution = 4 sub-steps, StepResolution < 0 (events located inside the steps, warp kernel).  usage: gpu_events_modes.py [nsets]"""
import ctypes as C, json, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = H.workload("events")
vals = H.workload_values(w, n, seed=77)
for label, k in (("output times (reference semantics)", 1), ("StepResolution 4", 4), ("located inside the steps (StepResolution -1)", -1)):
    opt = so.settings(w["end"], w["nout"], w["atol"], w["rtol"])
    so.lib().CvodeSettings_setStepResolution(opt, k)
    b = so.Batch(w["model"], opt, w["vary"])
    b.set_values(vals); b.integrate(n)
    ms = min(b.integrate(n) for _ in range(2))
    fired = b.event_log(n)
    print(json.dumps(dict(mode=label, n=n, ms=ms, traj_per_s=n / ms * 1e3, events_per_trajectory=float(np.mean([bin(int(x)).count("1") for x in fired[:, :2000].ravel()]) * fired.shape[0]),
                          failed=int((b.status(n) != 0).sum()), layout="warp" if "-DODES_WARP=1" in b.source().split("\n")[0] else "lane", info=b.kernel_info())))
    b.close()
