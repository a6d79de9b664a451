"""Forward sensitivity analysis through the batched warp-per-trajectory kernel.

The reference integrates the sensitivity equations with CVODES: simultaneous corrector, analytic right-hand side
built from the Jacobian and the parametric matrix, estimated tolerances, no error control on the sensitivity
variables (src/sensSolver.c:197-346), and reads them back after every output step (src/sensSolver.c:108-140).  The
checker is the reference's vendored CVODES 2.3.0 itself (oracle/sens_ref.c, compiled unmodified into
oracle/_ref/libcvodes_ref.so) driven with the emitted host-C model functions (the reference's compiled mode): the
kernel has to take the same steps (identical work counters) and produce BIT-IDENTICAL states and sensitivities --
far inside the north star's tolerance max(10 rtol |x|, 10 atol), which is asserted as well.
The reference's own unit test (unittest/test_sensSolver.c:14-31) runs MAPK.xml with the parameters MAPK, MAPK_P, K1, Ki;
it holds no numbers (smoke only), so the same selection is pinned here against the vendored CVODES."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so

REF_IDS = ["MAPK", "MAPK_P", "K1", "Ki"]                       # unittest/test_sensSolver.c:14-19
CASES = {
    "mapk": dict(ids=REF_IDS, nout=None),
    "repressilator": dict(ids=["alpha", "beta", "x1"], nout=100),
    "huang": dict(ids=["r_r1a_a1", "KKK", "r_r10b_k10"], nout=None),
    "goodwin": dict(ids=None, nout=20),                        # default: every constant of the model
    "events": dict(ids=["S1", "compartment", "S2"], nout=None),  # events re-initialise the solver from the current sensitivities
    "huang_all": dict(ids=None, nout=20, workload="huang"),    # 22 equations, all 31 constants: one vector per slab, 32 slabs
}
needs_ref = pytest.mark.skipif(not H.have_sens_ref(), reason="oracle/_ref/libcvodes_ref.so not built")


def _case(name, nout=None, method=0):
    c = CASES[name]
    w = H.workload(c.get("workload", name), nout=nout if nout is not None else c["nout"], sensitivity=1, sens_ids=c["ids"], sens_method=method)
    return w


def test_sense_structures_follow_the_reference():
    """ODESense_create (src/odeModel.c:2188-2278): index maps for variables and parameters, parametric matrix entries."""
    L = so.lib()
    m, _, _ = H.mapk_scan_model()
    L.ODEModel_constructJacobian(m.om)
    opt = so.settings(1000.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=REF_IDS)
    os_ = L.ODESense_create(m.om, opt)
    assert L.ODESense_getNsens(os_) == 4 and L.ODESense_getNeq(os_) == 8
    idx = []
    for k in range(4):
        vi = L.ODESense_getSensParamIndexByNum(os_, k)
        idx.append(L.ODEModel_getVariableName(m.om, vi).decode())
        L.VariableIndex_free(vi)
    assert idx == REF_IDS
    # variables have no column in the parametric matrix; d(dMKKK/dt)/dK1 is an expression, d(dMAPK/dt)/dK1 is zero
    assert L.ODESense_getSensIJEntry(os_, 0, 0) is None and L.ODESense_getSensIJEntry(os_, 0, 1) is None
    e = so.take_string(L.SBML_formulaToString(L.ODESense_getSensIJEntry(os_, 0, 2)))
    assert "K1" in e and "MKKK" in e
    assert so.take_string(L.SBML_formulaToString(L.ODESense_getSensIJEntry(os_, 5, 2))) == "0"
    L.ODESense_free(os_)
    # default selection: all constants, in value[] order
    os_ = L.ODEModel_constructSensitivity(m.om)
    assert L.ODESense_getNsens(os_) == m.nconst
    L.ODESense_free(os_)


def test_emitted_sense_function_follows_the_reference_generator():
    """src/odeModel.c:2994-3100: per row `= 0.0`, `+= (J_ij) * yS[j]` over the non-zero Jacobian entries, then the guarded
    parametric term."""
    L = so.lib()
    m, _, _ = H.mapk_scan_model()
    L.ODEModel_constructJacobian(m.om)
    opt = so.settings(1000.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=REF_IDS)
    os_ = L.ODESense_create(m.om, opt)
    src = so.take_string(L.ODEModel_generateCSourceSens(m.om, os_))
    L.ODESense_free(os_)
    body = src[src.index("void sense_f"):]
    assert body.count(" = 0.0;") == 8
    assert body.count(") * yS[") == len(m.jacobian_structure())
    assert "if ( 2 == iS ) ySdot[0] += " in body and "if ( 0 == iS )" not in body and "if ( 1 == iS )" not in body


@needs_ref
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("name,nsets", [("mapk", 24), ("repressilator", 6), ("huang", 6), ("goodwin", 6), ("events", 24), ("huang_all", 2)])
def test_emulated_sensitivities_are_bit_identical_to_vendored_cvodes(name, nsets, method):
    """method 0: CV_SIMULTANEOUS (the reference's default), 1: CV_STAGGERED, 2: CV_STAGGERED1 (src/sensSolver.c:286-290)."""
    w = _case(name, method=method)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=11)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    assert b.nsens == (len(CASES[name]["ids"]) if CASES[name]["ids"] else m.nconst)
    tp = H.tpoints(w["opt"], w["nout"])
    base, trig = b.base_values()
    got = H.Emu(b, name + "_sens%d" % method).run(base, trig, vals, tp, 10000, w["rtol"], w["atol"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], tag=name + "_sens")
    b.close()
    assert np.array_equal(got["status"], ref["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(got["fired"].T, ref["fired"]) and (name != "events" or ref["fired"].any())
    assert np.array_equal(got["stats"].T, ref["stats"][:, :7])                       # same steps, Newton and Jacobian decisions
    assert np.array_equal(np.transpose(got["out"], (2, 0, 1)), ref["out"])         # bit for bit
    assert np.array_equal(got["sens"], ref["sens"])
    # the stated gate, on the sensitivities as well
    assert H.parity(got["sens"], ref["sens"], w["rtol"], w["atol"]) <= 1.0
    # row 0: 1 for a variable with respect to its own initial value, else 0 (src/cvodeData.c:453-461)
    for k in range(b.nsens):
        i = so.lib().ODESBatch_getSensIndex(b.h, k) if b.h else None
    assert set(np.unique(got["sens"][:, 0])) <= {0.0, 1.0}


@needs_ref
def test_emulated_step_resolution_with_sensitivities():
    """StepResolution = k together with forward sensitivities (events re-initialise solver and sensitivities at every
    sub-step where they fire): kept rows of states AND sensitivities are rows k, 2k, ... of the vendored CVODES on the
    refined grid, bit for bit."""
    import ctypes as C
    k, nout, ids = 3, 20, ["S1", "compartment", "S2"]
    w = H.workload("events", nout=nout, sensitivity=1, sens_ids=ids)
    m = w["model"]
    vals = H.workload_values(w, 6, seed=4)
    so.lib().CvodeSettings_setStepResolution(w["opt"], k)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    assert "-DODES_STEPRES=3" in b.source().split("\n")[0] and b.nsens == 3
    tp = H.tpoints(w["opt"], nout)
    grid = np.array([tp[i] if j == 0 else tp[i] + j * ((tp[i + 1] - tp[i]) / k) for i in range(nout) for j in range(k)] + [tp[-1]])
    base, trig = b.base_values()
    got = H.Emu(b, "events_sens_stepres").run(base, trig, vals, grid, 10000, w["rtol"], w["atol"], stepres=k)
    w2 = H.workload("events", nout=nout * k, sensitivity=1, sens_ids=ids)
    arr = (C.c_double * (len(grid) - 1))(*grid[1:])
    assert so.lib().CvodeSettings_setTimeSeries(w2["opt"], arr, len(grid) - 1) == 1
    b2 = so.Batch(m, w2["opt"], w2["vary"], store=list(range(m.neq)))
    ref = H.sens_ref_run(m, w2["opt"], b2, w2["vary"], vals, grid, 10000, w["rtol"], w["atol"], tag="events_sens")
    b.close(); b2.close()
    assert ref["fired"].any()
    assert np.array_equal(np.transpose(got["out"], (2, 0, 1)), ref["out"][:, ::k])
    assert np.array_equal(got["sens"], ref["sens"][:, ::k])
    assert np.array_equal(got["stats"].T, ref["stats"][:, :7])


FORMS = {"rows": dict(ODES_WJE="0"), "shapes": dict(ODES_WJE="1", ODES_WDAG="0"), "dag": dict(ODES_WJE="1", ODES_WDAG="1")}


@needs_ref
@pytest.mark.parametrize("form", sorted(FORMS))
@pytest.mark.parametrize("name,nsets", [("mapk", 12), ("goodwin", 4), ("huang", 3)])
def test_emulated_jacobian_forms_agree_with_vendored_cvodes(monkeypatch, name, nsets, form):
    """The three ways the warp kernel evaluates the Jacobian and the parametric matrix for the sensitivity right-hand sides --
    row per lane, one entry per lane grouped by expression shape, common sub-expressions level by level (emit.c: entry
    tasks, entry DAG) -- perform the operations of the emitted expressions: each is bit-identical to the vendored CVODES.
    (The automatic choice is covered by the test above.)"""
    for k, v in FORMS[form].items():
        monkeypatch.setenv(k, v)
    w = _case(name)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=19)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    head = b.source().split("\n")[0]
    assert all(f"-D{k}={v}" in head for k, v in FORMS[form].items())
    if form != "rows":
        assert "#define W_JE_N 0" not in b.source().split("#if ODES_WARP")[1].split("#endif")[0] or name == "none"
    tp = H.tpoints(w["opt"], w["nout"])
    base, trig = b.base_values()
    got = H.Emu(b, f"{name}_sens_{form}").run(base, trig, vals, tp, 10000, w["rtol"], w["atol"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], tag=name + "_sens")
    b.close()
    assert np.array_equal(got["stats"].T, ref["stats"][:, :7])
    assert np.array_equal(np.transpose(got["out"], (2, 0, 1)), ref["out"])
    assert np.array_equal(got["sens"], ref["sens"])


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("form", sorted(FORMS))
@pytest.mark.parametrize("name,nsets", [("mapk", 256), ("goodwin", 64), ("repressilator", 64)])
def test_gpu_jacobian_forms_agree_with_vendored_cvodes(monkeypatch, name, nsets, form):
    """The same three forms on the device."""
    for k, v in FORMS[form].items():
        monkeypatch.setenv(k, v)
    w = _case(name)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=23)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    n = b.set_values(vals)
    b.integrate(n)
    out, sens, stats = b.results(n), b.sensitivities(n), b.stats(n).T
    tp = H.tpoints(w["opt"], w["nout"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], threads=8, tag=name + "_sens")
    b.close()
    assert np.array_equal(stats, ref["stats"][:, :7])
    assert np.array_equal(np.transpose(out, (2, 0, 1)), ref["out"], equal_nan=True)
    assert np.array_equal(sens, ref["sens"], equal_nan=True)


@needs_ref
def test_sensitivities_change_the_step_sequence_only_through_the_newton_test():
    """With errconS == FALSE the error test sees the state alone, but the corrector's convergence test sees every variable
    (cvodes.c:3907-3914): the reference's state trajectory with sensitivities is close to, not identical with, the plain
    run -- and the kernel reproduces whichever the reference produces."""
    w = _case("mapk", nout=20)
    wp = H.workload("mapk", nout=20)
    m = w["model"]
    vals = H.workload_values(w, 8, seed=5)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    tp = H.tpoints(w["opt"], 20)
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], tag="mapk_sens")
    b.close()
    plain = H.oracle_run(wp["model"], wp["opt"], wp["vary"], vals, 20, compiled_so=H.compile_c_model(wp["model"], "mapk"))
    assert H.parity(np.transpose(ref["out"], (1, 2, 0)), np.transpose(plain["out"][:, :, :m.neq], (1, 2, 0)), 1e-4, 1e-7) <= 1.0


@needs_ref
def test_sensitivities_agree_with_central_differences():
    """Size-independent property: dy/dp from the kernel against (y(p + d) - y(p - d)) / 2d of plain integrations at
    tight tolerances."""
    name = "mapk"
    w = H.workload(name, nout=10, sensitivity=1, sens_ids=["K1", "MKKK"])
    wp = H.workload(name, nout=10)
    m = w["model"]
    vals = H.workload_values(w, 2, seed=9)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    tp = H.tpoints(w["opt"], 10)
    base, trig = b.base_values()
    got = H.Emu(b, "mapk_sens_fd").run(base, trig, vals, tp, 10000, 1e-6, 1e-9)
    b.close()
    tight = so.settings(w["end"], 10, 1e-13, 1e-11)
    k1 = w["vary"].index(m.index("K1"))
    vary_fd = list(w["vary"]) + [m.index("MKKK")]
    for col, j in ((k1, 0), (len(vary_fd) - 1, 1)):
        v0 = np.concatenate([vals, np.full((2, 1), m.base_values()[m.index("MKKK")])], axis=1)
        d = 1e-5 * v0[:, col]
        vp, vm = v0.copy(), v0.copy()
        vp[:, col] += d
        vm[:, col] -= d
        yp = H.oracle_run(wp["model"], tight, vary_fd, vp, 10)["out"][:, :, :m.neq]
        ym = H.oracle_run(wp["model"], tight, vary_fd, vm, 10)["out"][:, :, :m.neq]
        fd = (yp - ym) / (2 * d)[:, None, None]
        s = got["sens"][:, :, j, :]
        assert np.allclose(s[:, 1:], fd[:, 1:], rtol=2e-3, atol=2e-4 * np.abs(fd).max()), (np.abs(s - fd).max(), np.abs(fd).max())


# difference quotients: parameters only (the reference's f freezes a VARIABLE listed as a sensitivity at its initial value in
# this mode, src/cvodeSolver.c:981-984 -- refused by the plan)
DQ_CASES = {
    "mapk": ["K1", "Ki", "V1"],
    "repressilator": ["alpha", "beta"],
    "huang": ["r_r1a_a1", "r_r10b_k10"],
    "events": ["compartment"],
}


def _dq_case(name, mode, method, nout=None):
    """mode 'nojac': UseJacobian 0 -> CVDenseDQJac and CVSensRhsDQ; 'nomatrix': the Jacobian is analytic, the parametric matrix
    is treated as not constructible (ODES_FORCE_SENS_DQ, a test knob for src/sensSolver.c:312-321 with om->jacobian set)."""
    return H.workload(name, nout=nout, sensitivity=1, sens_ids=DQ_CASES[name], sens_method=method, use_jacobian=0 if mode == "nojac" else 1)


@needs_ref
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("mode", ["nojac", "nomatrix"])
@pytest.mark.parametrize("name,nsets", [("mapk", 12), ("repressilator", 4), ("huang", 3), ("events", 12)])
def test_emulated_difference_quotient_sensitivities_are_bit_identical_to_vendored_cvodes(monkeypatch, name, nsets, mode, method):
    """src/sensSolver.c:312-321: without a parametric matrix (or without a Jacobian) the reference registers no sensitivity
    right-hand side -- CVODES' CVSensRhsDQ (centred, rhomax 0) over f with the parameters read from the perturbed array --
    and CVDENSE's CVDenseDQJac where there is no Jacobian.  The kernel's sens_dq / dq_jac against the vendored CVODES."""
    if mode == "nomatrix":
        monkeypatch.setenv("ODES_FORCE_SENS_DQ", "1")
    w = _dq_case(name, mode, method)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=13)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    assert b.nsens == len(DQ_CASES[name])
    assert so.lib().ODESBatch_getDifferenceQuotients(b.h) == (3 if mode == "nojac" else 1)
    tp = H.tpoints(w["opt"], w["nout"])
    base, trig = b.base_values()
    got = H.Emu(b, "%s_dq_%s%d" % (name, mode, method)).run(base, trig, vals, tp, 10000, w["rtol"], w["atol"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], tag="%s_dq_%s%d" % (name, mode, method))
    b.close()
    assert np.array_equal(got["status"], ref["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert np.array_equal(got["stats"].T, ref["stats"][:, :7])
    assert np.array_equal(np.transpose(got["out"], (2, 0, 1)), ref["out"])
    assert np.array_equal(got["sens"], ref["sens"])
    assert np.abs(ref["sens"][:, -1]).max() > 0


@needs_ref
def test_difference_quotient_sensitivities_are_close_to_the_analytic_ones():
    w = H.workload("mapk", nout=10, sensitivity=1, sens_ids=DQ_CASES["mapk"])
    wd = _dq_case("mapk", "nojac", 0, nout=10)
    vals = H.workload_values(w, 3, seed=2)
    tp = H.tpoints(w["opt"], w["nout"])
    res = []
    for ww in (w, wd):
        b = so.Batch(ww["model"], ww["opt"], ww["vary"], store=list(range(ww["model"].neq)))
        res.append(H.sens_ref_run(ww["model"], ww["opt"], b, ww["vary"], vals, tp, 10000, ww["rtol"], ww["atol"], tag="mapk_dqcmp%d" % len(res)))
        b.close()
    scale = np.abs(res[0]["sens"]).max()
    assert np.abs(res[0]["sens"] - res[1]["sens"]).max() < 1e-3 * scale


def test_unsupported_sensitivity_configurations_fail_loudly():
    m2, vary, _ = H.mapk_scan_model()
    opt2 = so.settings(10.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=["K1"])
    C.cast(opt2, C.POINTER(C.c_int * 64)).contents    # (SensMethod is range-checked by its setter: 0..2 only)
    opt3 = so.settings(10.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=["no_such_symbol"])
    with pytest.raises(so.OdesError):
        so.Batch(m2, opt3, vary)
    # difference quotients with respect to an initial value: the reference's f would freeze that variable
    with pytest.raises(so.OdesError, match="initial values"):
        so.Batch(m2, so.settings(10.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=["K1", "MAPK"], use_jacobian=0), vary)
    b = so.Batch(m2, so.settings(10.0, 10, 1e-9, 1e-6), vary)
    assert b.nsens == 0
    b.close()


# ------------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("name,nsets", [("mapk", 512), ("repressilator", 64), ("huang", 96), ("goodwin", 64), ("events", 512), ("huang_all", 32)])
def test_gpu_sensitivities_are_bit_identical_to_vendored_cvodes(name, nsets, method):
    w = _case(name, method=method)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=2027)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    n = b.set_values(vals)
    ms = b.integrate(n)
    out, sens, status, stats = b.results(n), b.sensitivities(n), b.status(n), b.stats(n).T
    fired = b.event_log(n).T
    tp = H.tpoints(w["opt"], w["nout"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], threads=8, tag=name + "_sens")
    print(name, "method", method, "nsens", b.nsens, f"{n / ms * 1e3:.0f} trajectories/s", b.kernel_info())
    b.close()
    assert np.array_equal(status, ref["status"])
    assert np.array_equal(fired, ref["fired"])
    assert np.array_equal(stats, ref["stats"][:, :7])
    assert H.parity(sens, ref["sens"], w["rtol"], w["atol"]) <= 1.0                  # the gate
    assert np.array_equal(np.transpose(out, (2, 0, 1)), ref["out"], equal_nan=True)  # and bit for bit
    assert np.array_equal(sens, ref["sens"], equal_nan=True)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("method", [0, 2])
@pytest.mark.parametrize("mode", ["nojac", "nomatrix"])
@pytest.mark.parametrize("name,nsets", [("mapk", 256), ("huang", 48), ("events", 256)])
def test_gpu_difference_quotient_sensitivities_are_bit_identical_to_vendored_cvodes(monkeypatch, name, nsets, mode, method):
    """CVSensRhsDQ / CVDenseDQJac of the reference's fallbacks (src/sensSolver.c:312-321) on the device."""
    if mode == "nomatrix":
        monkeypatch.setenv("ODES_FORCE_SENS_DQ", "1")
    w = _dq_case(name, mode, method)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=29)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    n = b.set_values(vals)
    ms = b.integrate(n)
    out, sens, status, stats = b.results(n), b.sensitivities(n), b.status(n), b.stats(n).T
    fired = b.event_log(n).T
    tp = H.tpoints(w["opt"], w["nout"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], threads=8, tag="%s_dq_%s%d" % (name, mode, method))
    print(name, mode, "method", method, f"{n / ms * 1e3:.0f} trajectories/s", b.kernel_info())
    b.close()
    assert np.array_equal(status, ref["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(fired, ref["fired"])
    assert np.array_equal(stats, ref["stats"][:, :7])
    assert np.array_equal(np.transpose(out, (2, 0, 1)), ref["out"], equal_nan=True)
    assert np.array_equal(sens, ref["sens"], equal_nan=True)


@pytest.mark.gpu
def test_gpu_sensitivities_respect_the_conservation_laws_at_scale():
    """Full-size property (no oracle): MKKK + MKKK_P is conserved, so the sensitivities of the pair sum to zero for a
    parameter and to one for the initial value of MKKK; failed rows are not-a-number in states and sensitivities alike."""
    w = H.workload("mapk", sensitivity=1, sens_ids=["K1", "MKKK", "r_J1_V2"])
    m = w["model"]
    nsets = 50_000
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    lo, hi = np.full(len(w["vary"]), -1.0), np.full(len(w["vary"]), 1.0)
    b.generate_values(nsets, 20261019, 0, w["nominal"], lo, hi)
    ms = b.integrate(nsets)
    sens, status = b.sensitivities(nsets), b.status(nsets)
    print(f"MAPK + 3 sensitivities: {nsets / ms * 1e3:.0f} trajectories/s", b.kernel_info())
    b.close()
    assert (status == 0).all()
    pair = sens[:, :, :, 0] + sens[:, :, :, 1]                    # [set][row][sens]
    scale = np.abs(sens[:, :, :, :2]).max(axis=3) + 1.0
    assert (np.abs(pair[:, :, 0]) <= 1e-6 * scale[:, :, 0]).all()
    assert (np.abs(pair[:, :, 2]) <= 1e-6 * scale[:, :, 2]).all()
    assert (np.abs(pair[:, :, 1] - 1.0) <= 1e-6 * scale[:, :, 1]).all()


@pytest.mark.gpu
@needs_ref
def test_host_entries_return_the_sensitivities():
    """ODESBatch_runHostSens (streamed host pipeline), IntegratorInstance_integrateBatchSens (one-shot entry, all values
    stored) and Model_odeSolverBatch (SBMLResults per design point) deliver the device-resident sensitivities."""
    L = so.lib()
    w = _case("mapk", nout=20)
    m = w["model"]
    vals = H.workload_values(w, 700, seed=4)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    n = b.set_values(vals)
    b.integrate(n)
    out, sens = b.results(n), b.sensitivities(n)
    res = np.zeros((n, b.nrows, b.nstore)); hs = np.zeros((n, b.nrows, b.nsens, m.neq)); st = np.zeros(n, dtype=np.int32)
    b.run_host(vals, res, st, chunk=256, sens=hs)
    assert np.array_equal(hs, sens) and np.array_equal(np.transpose(res, (1, 2, 0)), out) and (st == 0).all()
    b.close()
    # one-shot entry through an integrator instance
    ii = L.IntegratorInstance_create(m.om, w["opt"])
    assert ii and L.IntegratorInstance_getNsens(ii) == 4
    vis = (C.c_void_p * len(w["vary"]))(*[L.ODEModel_getVariableIndexByNum(m.om, i) for i in w["vary"]])
    full = np.zeros((n, 21, m.nvalues)); hs2 = np.zeros_like(hs); st2 = np.zeros(n, dtype=np.int32)
    v = np.ascontiguousarray(vals)
    failed = L.IntegratorInstance_integrateBatchSens(ii, n, len(w["vary"]), vis, v.ctypes.data, full.ctypes.data, hs2.ctypes.data, st2.ctypes.data)
    assert failed == 0 and np.array_equal(hs2, sens) and np.array_equal(full[:, :, :m.neq], np.transpose(out, (2, 0, 1)))
    for p in vis:
        L.VariableIndex_free(p)
    L.IntegratorInstance_free(ii)


@pytest.mark.gpu
def test_model_odesolverbatch_maps_sensitivities_per_design_point():
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(100.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=REF_IDS)
    vs = L.VarySettings_allocate(1, 3)
    L.VarySettings_addParameter(vs, b"K1", None)
    for x in (5.0, 10.0, 20.0):
        arr = (C.c_double * 1)(x)
        L.VarySettings_addDesignPoint(vs, arr)
    res = L.Model_odeSolverBatch(m.sbml, opt, vs)
    assert res and L.SBMLResultsArray_getNumResults(res) == 3
    # the same design points through a plan
    b = so.Batch(m, opt, [m.index("K1")], store=list(range(m.neq)))
    b.set_values(np.array([[5.0], [10.0], [20.0]]))
    b.integrate(3)
    sens = b.sensitivities(3)
    b.close()
    for k in range(3):
        r = L.SBMLResultsArray_getResults(res, k)
        assert L.SBMLResults_getNumSens(r) == 4 and L.SBMLResults_getSensParam(r, 2) == b"K1"
        tc = L.SBMLResults_getTimeCourse(r, b"MKKK_P")
        got = np.array([[L.TimeCourse_getSensitivity(tc, j, t) for t in range(11)] for j in range(4)])
        assert np.array_equal(got, sens[k, :, :, 1].T)
    L.SBMLResultsArray_free(res)
    L.VarySettings_free(vs)


@pytest.mark.gpu
def test_reference_entries_take_the_difference_quotient_fallback():
    """Model_odeSolverBatch and the single IntegratorInstance with UseJacobian 0 and sensitivity analysis: no refusal any more,
    the same numbers as a plan created with the same settings (CVSensRhsDQ + CVDenseDQJac, checked against CVODES above)."""
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    ids = ["K1", "Ki"]
    opt = so.settings(100.0, 10, 1e-9, 1e-6, sensitivity=1, sens_ids=ids, use_jacobian=0)
    vs = L.VarySettings_allocate(1, 2)
    L.VarySettings_addParameter(vs, b"K1", None)
    for x in (5.0, 20.0):
        L.VarySettings_addDesignPoint(vs, (C.c_double * 1)(x))
    res = L.Model_odeSolverBatch(m.sbml, opt, vs)
    assert res and L.SBMLResultsArray_getNumResults(res) == 2
    b = so.Batch(m, opt, [m.index("K1")], store=list(range(m.neq)))
    assert L.ODESBatch_getDifferenceQuotients(b.h) == 3
    b.set_values(np.array([[5.0], [20.0]]))
    b.integrate(2)
    sens = b.sensitivities(2)
    b.close()
    assert np.abs(sens[:, -1]).max() > 0
    for k in range(2):
        r = L.SBMLResultsArray_getResults(res, k)
        assert L.SBMLResults_getNumSens(r) == 2
        tc = L.SBMLResults_getTimeCourse(r, b"MKKK_P")
        got = np.array([[L.TimeCourse_getSensitivity(tc, j, t) for t in range(11)] for j in range(2)])
        assert np.array_equal(got, sens[k, :, :, 1].T)
    L.SBMLResultsArray_free(res)
    L.VarySettings_free(vs)
    # the single instance: integrate to the end, read the sensitivities of the last row
    ii = L.IntegratorInstance_create(m.om, opt)
    assert ii and L.IntegratorInstance_getNsens(ii) == 2
    assert L.IntegratorInstance_integrate(ii) == 1
    L.IntegratorInstance_free(ii)


@needs_ref
def test_failed_trajectories_leave_not_a_number_rows_in_states_and_sensitivities():
    """mxstep too small: CV_TOO_MUCH_WORK at the same output interval as the reference, the rows before it identical, the rows
    after it not-a-number in the states AND the sensitivities."""
    w = _case("mapk", nout=6)
    m = w["model"]
    vals = H.workload_values(w, 5, seed=8)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    tp = H.tpoints(w["opt"], 6)
    base, trig = b.base_values()
    got = H.Emu(b, "mapk_sens0").run(base, trig, vals, tp, 15, w["rtol"], w["atol"])
    ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 15, w["rtol"], w["atol"], tag="mapk_sens")
    b.close()
    assert (ref["status"] == -4).all() and np.array_equal(got["status"], ref["status"])
    assert np.isnan(ref["sens"]).any() and not np.isnan(ref["sens"][:, 0]).any()
    assert np.array_equal(np.transpose(got["out"], (2, 0, 1)), ref["out"], equal_nan=True)
    assert np.array_equal(got["sens"], ref["sens"], equal_nan=True)
