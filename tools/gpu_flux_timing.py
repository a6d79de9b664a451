"""Wall time of Model_odeSolverBatch / Model_odeSolverBatchFlat on N MAPK design points with the reaction fluxes mapped on
the device (odes_flux_kernel) and with the reference's host walk (ODES_HOST_FLUX=1; src/odeSolver.c:519-531).
usage: gpu_flux_timing.py [N]   -> JSON lines (profiles/r02_flux_mapping.jsonl)"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import sbml_odesolver_b200 as so, helpers as H
    L = so.lib()
    n, mode = int(sys.argv[2]), sys.argv[3]
    m = so.Model(H.model_path("mapk_cascade.xml"))
    rng = np.random.default_rng(5)
    pts = np.array([2.5, 9.0, 10.0])[None, :] * np.exp2(rng.uniform(-1, 1, size=(n, 3)))
    vs = L.VarySettings_allocate(3, n)
    for k in (b"V1", b"Ki", b"K1"):
        assert L.VarySettings_addParameter(vs, k, None)
    for row in pts:
        L.VarySettings_addDesignPoint(vs, np.ascontiguousarray(row).ctypes.data)
    opt = so.settings(1000.0, 100, 1e-9, 1e-6)
    out = {}
    for rep in range(2):                       # the second call shows the warm state (plans, buffers)
        t0 = time.perf_counter()
        if mode == "flat":
            r = L.Model_odeSolverBatchFlat(m.sbml, opt, vs)
            assert r, so.errors()
            dt = time.perf_counter() - t0
            j5 = L.BatchResults_getFlux(r, n - 1, b"J5", 100) if not os.environ.get("ODES_HOST_FLUX") else float("nan")
            L.BatchResults_free(r)
        else:
            r = L.Model_odeSolverBatch(m.sbml, opt, vs)
            assert r, so.errors()
            dt = time.perf_counter() - t0
            tc = L.SBMLResults_getTimeCourse(L.SBMLResultsArray_getResults(r, n - 1), b"J5")
            j5 = L.TimeCourse_getValue(tc, 100)
            L.SBMLResultsArray_free(r)
        out[f"call{rep + 1}_s"] = round(dt, 3)
    print(json.dumps(dict(entry="Model_odeSolverBatchFlat" if mode == "flat" else "Model_odeSolverBatch", design_points=n,
                          fluxes="host evaluateAST" if os.environ.get("ODES_HOST_FLUX") else "device odes_flux_kernel", last_J5=j5, **out)))
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "100000"
    for mode in ("array", "flat"):
        for host in ("1", ""):
            if mode == "flat" and host: continue
            env = dict(os.environ)
            if host: env["ODES_HOST_FLUX"] = "1"
            r = subprocess.run([sys.executable, __file__, "--one", n, mode], env=env, capture_output=True, text=True)
            print(r.stdout.strip().split("\n")[-1] if r.returncode == 0 else f"FAILED {mode} {host}: {r.stderr[-600:]}")
