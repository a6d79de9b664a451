"""Small runs of every kernel variant for compute-sanitizer (memcheck / racecheck): lane kernel (MAPK), warp kernel (huang96),
sensitivity kernel with the entry DAG (MAPK + 3) and the entry shapes (repressilator + 3), StepResolution on the events model,
the one-shot entry with pageable arrays (staging ring) and the flux kernel.  usage: compute-sanitizer --tool memcheck python tools/gpu_sanitize.py"""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
L = so.lib()
def run(name, n, nout, **kw):
    w = H.workload(name, nout=nout, **kw)
    b = so.Batch(w["model"], w["opt"], w["vary"], store=list(range(w["model"].neq)))
    b.set_values(H.workload_values(w, n, seed=3)); b.integrate(n)
    print(name, kw, "ok", int((b.status(n) == 0).sum()), "/", n, b.kernel_info()["regs"], "regs")
    if b.nsens: b.sensitivities(n)
    b.close()
run("mapk", 300, 10)
run("huang", 40, 10)
run("mapk", 64, 10, sensitivity=1, sens_ids=["K1", "Ki", "MAPK"])
run("repressilator", 32, 20, sensitivity=1, sens_ids=["alpha", "beta", "x1"])
run("goodwin", 32, 10, sensitivity=1)
w = H.workload("events", nout=20)
so.lib().CvodeSettings_setStepResolution(w["opt"], 3)
b = so.Batch(w["model"], w["opt"], w["vary"]); b.set_values(H.workload_values(w, 100, seed=1)); b.integrate(100); b.results(100); b.close(); print("stepres ok")
w = H.workload("mapk", nout=10); m = w["model"]
ii = L.IntegratorInstance_create(m.om, w["opt"])
vis = (C.c_void_p * len(w["vary"]))(*[L.ODEModel_getVariableIndexByNum(m.om, i) for i in w["vary"]])
vals = H.workload_values(w, 500, seed=2); out = np.zeros((500, 11, m.nvalues)); st = np.zeros(500, dtype=np.int32)
print("entry failed:", L.IntegratorInstance_integrateBatch(ii, 500, len(w["vary"]), vis, vals.ctypes.data, out.ctypes.data, st.ctypes.data))
L.IntegratorInstance_free(ii)
