"""CPU tier: the per-trajectory program of the CUDA kernel (generated model section + bdf_kernel.cuh),
compiled for the HOST by tests/emu/emu_harness.cpp with g++ -ffp-contract=off, must reproduce the oracle.

Against the oracle driven with the SAME emitted right-hand side / Jacobian (ODEModel_generateCSource, the
reference's compiled mode) every trajectory has to take the same steps, orders, Newton and Jacobian decisions
(identical work counters, flags and event logs) with the same arithmetic -- including pow(x, 1/n) of the step-size
controller, for which the kernel restates glibc's algorithm -- so every stored value is bit-identical.
Against the interpreted oracle (evaluateAST semantics) it has to
agree within max(10 rtol |x|, 10 atol) -- the tolerance the north star states.
The GPU tier (test_gpu_parity.py) checks the real kernel the same way."""
import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so


def _emulate(w, vals, store=None, tag=None):
    b = so.Batch(w["model"], w["opt"], w["vary"], store=store)
    emu = H.Emu(b, tag or w["name"])
    base, trig = b.base_values()
    r = emu.run(base, trig, vals, H.tpoints(w["opt"], w["nout"]), 10000, w["rtol"], w["atol"])
    b.close()
    return r


@pytest.mark.parametrize("name,nsets,nout", [("mapk", 48, None), ("repressilator", 8, 200), ("goodwin", 24, None), ("events", 32, None), ("huang", 24, None)])
def test_emulated_kernel_is_bit_identical_to_compiled_oracle(name, nsets, nout):
    w = H.workload(name, nout=nout)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=42)
    got = _emulate(w, vals)
    cso = H.compile_c_model(m, name)
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], compiled_so=cso)
    want = np.transpose(ref["out"], (1, 2, 0))                       # [row][value][set]
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["stats"].T, ref["stats"])
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert np.array_equal(got["out"], want, equal_nan=True)             # bit for bit
    # and within the stated tolerance of the interpreted oracle
    refi = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"])
    assert H.parity(got["out"], np.transpose(refi["out"], (1, 2, 0)), w["rtol"], w["atol"]) <= 1.0
    assert np.array_equal(got["fired"].T, refi["fired"])


@pytest.mark.parametrize("name,nsets,nout,layout", [("mapk", 24, None, "lane"), ("repressilator", 6, 200, "lane"), ("huang", 6, None, "lane"),
                                                    ("huang", 6, None, "warp"), ("mapk", 12, None, "warp"), ("events", 16, None, "warp")])
def test_emulated_difference_quotient_jacobian(monkeypatch, name, nsets, nout, layout):
    """UseJacobian 0 (src/cvodeSolver.c:620-623: no Jacobian function is registered): CVDENSE's CVDenseDQJac
    (cvdense.c:479-537) in both kernel layouts, against the vendored CVODE running its own difference quotients."""
    monkeypatch.setenv("ODES_WARP", "1" if layout == "warp" else "0")      # default: warp from 19 equations on, as with a Jacobian
    w = H.workload(name, nout=nout, use_jacobian=0)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=7)
    b = so.Batch(m, w["opt"], w["vary"])
    assert so.lib().ODESBatch_getDifferenceQuotients(b.h) == 2
    assert ("-DODES_WARP=1" in b.source().split("\n", 1)[0]) == (layout == "warp")
    b.close()
    got = _emulate(w, vals, tag="%s_dqjac_%s" % (name, layout))
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF if H.have_ref() else H.PORT, compiled_so=H.compile_c_model(m, name))
    assert np.array_equal(got["status"], ref["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(got["stats"].T, ref["stats"]) and (ref["stats"][:, 3] > 0).all()
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)


def test_emulated_store_selection_and_row0():
    """Only the requested value[] entries are stored; row 0 holds the varied entries (integratorInstance.c:1571-1572)."""
    w = H.workload("mapk", nout=10)
    m = w["model"]
    vals = H.workload_values(w, 4, seed=1)
    store = [0, 7, 8, m.index("V1")]
    got = _emulate(w, vals, store=store, tag="mapk_sel")
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], compiled_so=H.compile_c_model(m, "mapk"))
    want = np.transpose(ref["out"][:, :, store], (1, 2, 0))
    assert np.array_equal(got["out"], want)
    assert np.array_equal(got["out"][0, 3], vals[:, 0])              # V1 is the first varied entry


def test_emulated_failure_codes_and_nan_rows():
    w = H.workload("mapk", nout=4)
    m = w["model"]
    vals = H.workload_values(w, 3, seed=2)
    b = so.Batch(m, w["opt"], w["vary"])
    emu = H.Emu(b, "mapk_all")
    base, trig = b.base_values()
    r = emu.run(base, trig, vals, H.tpoints(w["opt"], 4), 15, w["rtol"], w["atol"])      # mxstep 15 per interval
    assert (r["status"] == -4).all()                                                    # CV_TOO_MUCH_WORK
    assert not np.isnan(r["out"][0]).any() and np.isnan(r["out"][1:]).all()
    b.close()


def test_emulated_events_halt_and_no_reset():
    m = so.Model(H.model_path("two_events_one_assignment.xml"))
    vary = [m.index("S1"), m.index("S2")]
    vals = np.array([[1.0, 0.0], [1.7, 0.3], [0.6, 0.1]])
    for kw in (dict(halt_on_event=1), dict()):
        opt = so.settings(10.0, 100, 1e-9, 1e-6, **kw)
        b = so.Batch(m, opt, vary)
        emu = H.Emu(b, "events_" + ("halt" if kw else "run"))
        base, trig = b.base_values()
        got = emu.run(base, trig, vals, H.tpoints(opt, 100), 10000, 1e-6, 1e-9)
        ref = H.oracle_run(m, opt, vary, vals, 100, compiled_so=H.compile_c_model(m, "events"))
        assert np.array_equal(got["fired"].T, ref["fired"])
        assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
        b.close()
    # nominal set reproduces the reference-semantics firing sequence of SURVEY appendix D
    fired = ref["fired"][0]
    assert [(e, i) for i in range(101) for e in (0, 1) if fired[i] >> e & 1][:4] == [(1, 7), (0, 24), (1, 25), (1, 34)]


@pytest.mark.parametrize("warp", [0, 1])
def test_emulated_step_resolution(monkeypatch, warp):
    """StepResolution = k: the kernel walks k sub-steps per print step with the reference's output-time event handling at
    each and keeps the rows of the print steps -- equal, bit for bit, to rows k, 2k, ... of the oracle run on the refined
    grid; the event mask of a kept row is the OR of the masks since the row before (tests/test_gpu_parity.py has the GPU twin)."""
    import ctypes as C
    monkeypatch.setenv("ODES_WARP", str(warp))
    k, nout = 4, 25
    m = so.Model(H.model_path("two_events_one_assignment.xml"))
    vary = [m.index("S1"), m.index("S2")]
    vals = np.array([[1.0, 0.0], [1.7, 0.3], [0.6, 0.1], [0.9, 0.25]])
    opt = so.settings(10.0, nout, 1e-9, 1e-6)
    so.lib().CvodeSettings_setStepResolution(opt, k)
    b = so.Batch(m, opt, vary)
    assert "-DODES_STEPRES=4" in b.source().split("\n")[0]
    tp = H.tpoints(opt, nout)
    grid = np.array([tp[i] if j == 0 else tp[i] + j * ((tp[i + 1] - tp[i]) / k) for i in range(nout) for j in range(k)] + [tp[-1]])
    emu = H.Emu(b, f"events_stepres{warp}")
    base, trig = b.base_values()
    got = emu.run(base, trig, vals, grid, 10000, 1e-6, 1e-9, stepres=k)
    opt_ref = so.settings(10.0, nout * k, 1e-9, 1e-6)
    arr = (C.c_double * (len(grid) - 1))(*grid[1:])
    assert so.lib().CvodeSettings_setTimeSeries(opt_ref, arr, len(grid) - 1) == 1
    ref = H.oracle_run(m, opt_ref, vary, vals, nout * k, compiled_so=H.compile_c_model(m, "events"))
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0))[::k], equal_nan=True)
    acc = np.zeros((nout + 1, len(vals)), dtype=np.uint32)
    for r in range(1, nout + 1):
        acc[r] = np.bitwise_or.reduce(ref["fired"][:, (r - 1) * k + 1:r * k + 1], axis=1)
    assert np.array_equal(got["fired"], acc) and acc.any()
    b.close()


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libcvode_ref.so not built")
@pytest.mark.parametrize("warp", [0, 1])
def test_emulated_event_location_inside_steps(monkeypatch, warp):
    """StepResolution < 0: events are LOCATED inside the internal steps (north star: "per-system event root-finding").  After
    every step the triggers are evaluated at the step's end, a trigger that became true is traced back by bisection on the
    dense output to CVODE's root tolerance 100 uround (|tn| + |hu|), the event is applied there with the reference's event
    semantics and the solver restarts -- in both layouts.  The checker drives the reference's vendored CVODE through the same procedure step by
    step (CV_ONE_STEP + CVodeGetDky, oracle/driver.c: run_rootfind): rows, event masks and work counters are bit-identical;
    and the decay S1' = -S1 with "S1 < 0.1 -> S1 = 1" now resets at t = ln 10 instead of at the next print step."""
    import ctypes as C
    monkeypatch.setenv("ODES_WARP", str(warp))
    nout = 100
    m = so.Model(H.model_path("two_events_one_assignment.xml"))
    vary = [m.index("S1"), m.index("S2")]
    vals = np.array([[1.0, 0.0], [1.7, 0.3], [0.6, 0.1], [0.9, 0.25]])
    opt = so.settings(10.0, nout, 1e-9, 1e-6)
    so.lib().CvodeSettings_setStepResolution(opt, -1)
    b = so.Batch(m, opt, vary)
    head = b.source().split("\n")[0]
    assert "-DODES_ROOTFIND=1" in head and ("-DODES_WARP=1" in head) == bool(warp)
    tp = H.tpoints(opt, nout)
    base, trig = b.base_values()
    got = H.Emu(b, f"events_rootfind{warp}").run(base, trig, vals, tp, 10000, 1e-6, 1e-9)
    ref = H.oracle_run(m, opt, vary, vals, nout, backend=H.REF, compiled_so=H.compile_c_model(m, "events"))
    b.close()
    assert (got["status"] == 0).all() and (ref["status"] == 0).all()
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)))
    assert np.array_equal(got["fired"].T, ref["fired"]) and ref["fired"].any()
    assert np.array_equal(got["stats"].T, ref["stats"])
    # set 0 starts at S1 = 1: the reset happens at ln 10 = 2.3026, so the row at t = 2.4 holds exp(-(2.4 - ln 10))
    k = int(np.searchsorted(tp, np.log(10.0)))
    assert abs(got["out"][k, 0, 0] - np.exp(-(tp[k] - np.log(10.0)))) < 1e-4
    # with the reference's semantics the same row holds the freshly reset value
    opt0 = so.settings(10.0, nout, 1e-9, 1e-6)
    plain = H.oracle_run(m, opt0, vary, vals[:1], nout, backend=H.REF, compiled_so=H.compile_c_model(m, "events"))
    assert plain["out"][0, k, 0] == 1.0
    # what the mode cannot be combined with is refused at plan creation
    opt_h = so.settings(10.0, nout, 1e-9, 1e-6, halt_on_event=1)
    so.lib().CvodeSettings_setStepResolution(opt_h, -1)
    with pytest.raises(so.OdesError):
        so.Batch(m, opt_h, vary)
    so.lib().SolverError_clear()


@pytest.mark.parametrize("name,nsets,nout", [("mapk", 8, None), ("events", 12, None), ("goodwin", 6, None)])
def test_emulated_f_sharing_the_words_of_y(monkeypatch, name, nsets, nout):
    """ODES_ALIAS_FY: with the analytic Jacobian the right-hand side can share the shared-memory words of its argument (one
    vector less per lane; the host switches it on where it buys a resident block, e.g. the six-equation repressilator).
    Forced on here for models that do not get it by default: bit-identical to the oracle."""
    monkeypatch.setenv("ODES_ALIAS_FY", "2")
    w = H.workload(name, nout=nout)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=41)
    b = so.Batch(m, w["opt"], w["vary"])
    assert "-DODES_ALIAS_FY=1" in b.source().split("\n")[0]
    base, trig = b.base_values()
    got = H.Emu(b, name + "_alias").run(base, trig, vals, H.tpoints(w["opt"], w["nout"]), 10000, w["rtol"], w["atol"])
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], compiled_so=H.compile_c_model(m, name))
    b.close()
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    assert np.array_equal(got["stats"].T, ref["stats"]) and np.array_equal(got["fired"].T, ref["fired"])


def test_generated_source_has_single_hot_instances():
    """The block pipeline keeps one instance of the model functions in the instruction stream of the hot loop."""
    w = H.workload("mapk", nout=10)
    b = so.Batch(w["model"], w["opt"], w["vary"], store=list(range(8)))
    src = b.source()
    assert "sm_100a" in src and "odes_bdf_kernel" in src and "__launch_bounds__" in src
    assert src.count("DEV void ode_f(") == 1 and src.count("DEV void jacobi_f(") == 1
    assert "#define NEQ 8" in src and "#define NNZ 26" in src and "#define NVARY 21" in src
    b.close()


@pytest.mark.parametrize("name,nsets,nout", [("mapk", 12, None), ("repressilator", 4, 100), ("goodwin", 8, None), ("events", 16, None)])
def test_emulated_warp_kernel_is_bit_identical_on_the_small_models(monkeypatch, name, nsets, nout):
    """The warp-per-trajectory layout (bdf_warp_kernel.cuh, chosen automatically for n > 12) forced onto the small
    models: events with solver restarts, pow in rate laws, long output grids -- same bits as the oracle."""
    monkeypatch.setenv("ODES_WARP", "1")
    w = H.workload(name, nout=nout)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=43)
    got = _emulate(w, vals, tag=name + "_warp")
    assert "-DODES_WARP=1" in so.Batch(m, w["opt"], w["vary"]).source().split("\n")[0]
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], compiled_so=H.compile_c_model(m, name))
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["stats"].T, ref["stats"])
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)


@pytest.mark.parametrize("warp", [0, 1])
def test_emulated_steady_state_halt(monkeypatch, warp):
    """CvodeSettings_setHaltOnSteadyState per trajectory (src/integratorInstance.c:1176-1182, :1442-1507): a run ends
    with the first output row at which mean + standard deviation of the rates falls below the threshold; later rows
    are not-a-number, the status stays 0.  huang96 settles; every set stops at the row the oracle stops at."""
    monkeypatch.setenv("ODES_WARP", str(warp))
    w = H.workload("huang", nout=40)
    m = w["model"]
    opt = so.settings(400.0, 40, 1e-9, 1e-6, steady_state=1, ss_threshold=1e-5)
    vals = H.workload_values(w, 6, seed=12)
    b = so.Batch(m, opt, w["vary"])
    emu = H.Emu(b, f"huang_steady{warp}")
    base, trig = b.base_values()
    got = emu.run(base, trig, vals, H.tpoints(opt, 40), 10000, 1e-6, 1e-9)
    b.close()
    ref = H.oracle_run(m, opt, w["vary"], vals, 40, compiled_so=H.compile_c_model(m, "huang"))
    want = np.transpose(ref["out"], (1, 2, 0))
    last = (~np.isnan(want[:, 0, :])).sum(axis=0)                      # rows stored per set
    assert ((last > 3) & (last < 41)).all(), last                       # every set halts somewhere inside the grid
    assert (got["status"] == 0).all() and np.array_equal(got["status"], ref["status"])
    assert np.array_equal(np.isnan(got["out"]), np.isnan(want))
    assert np.array_equal(got["out"], want, equal_nan=True)


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libcvode_ref.so not built")
@pytest.mark.parametrize("name,nsets,nout", [("mapk", 12, None), ("repressilator", 4, 100), ("goodwin", 8, None), ("events", 16, None), ("huang", 4, 40)])
def test_emulated_adams_moulton_is_bit_identical_to_vendored_cvode(name, nsets, nout):
    """cvodeSettings_t.CvodeMethod = 1 (CvodeSettings_setMethod(set, 1, 12), examples/simpleIntegratorInstance.c:72):
    Adams-Moulton, orders 1..12 (cvode.c:1763-1792 CVAdjustAdams, :1962-2066 CVSetAdams), in the warp-per-trajectory
    kernel.  The checker is the vendored CVODE itself with CV_ADAMS (the restated port holds BDF only)."""
    w = H.workload(name, nout=nout, method=1)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=21)
    b = so.Batch(m, w["opt"], w["vary"])
    assert "-DODES_ADAMS=1" in b.source().split("\n")[0]
    base, trig = b.base_values()
    got = H.Emu(b, name + "_adams").run(base, trig, vals, H.tpoints(w["opt"], w["nout"]), 10000, w["rtol"], w["atol"])
    b.close()
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF, compiled_so=H.compile_c_model(m, name), adams=True)
    assert np.array_equal(got["status"], ref["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(got["stats"].T, ref["stats"])
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    # Adams and BDF are different methods: the step counts differ from the BDF run of the same sets
    bdf = H.oracle_run(m, H.workload(name, nout=nout)["opt"], w["vary"], vals, w["nout"], backend=H.REF, compiled_so=H.compile_c_model(m, name))
    assert not np.array_equal(bdf["stats"], ref["stats"])


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libcvode_ref.so not built")
@pytest.mark.parametrize("name,method,nsets,nout", [("goodwin", 1, 8, None), ("goodwin", 0, 8, None), ("mapk", 1, 8, 30), ("repressilator", 1, 4, 60), ("events", 1, 16, None)])
def test_emulated_functional_iteration_is_bit_identical_to_vendored_cvode(name, method, nsets, nout):
    """cvodeSettings_t.IterMethod = 1: CVnlsFunctional (cvode.c:2180-2225), the fixed-point corrector without matrix, with BDF or
    (the classic non-stiff pairing) Adams-Moulton.  Checker: the vendored CVODE with CV_FUNCTIONAL."""
    w = H.workload(name, nout=nout, method=method, iter_method=1)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=22)
    b = so.Batch(m, w["opt"], w["vary"])
    assert "-DODES_FUNCTIONAL=1" in b.source().split("\n")[0]
    base, trig = b.base_values()
    got = H.Emu(b, name + "_func%d" % method).run(base, trig, vals, H.tpoints(w["opt"], w["nout"]), 10000, w["rtol"], w["atol"])
    b.close()
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF, compiled_so=H.compile_c_model(m, name), adams=bool(method), functional=True)
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["stats"].T, ref["stats"]) and (ref["stats"][:, 2:5] == 0).all()      # no set-ups, Jacobians, Newton iterations
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
