"""Result mapping onto SBML structures (include/sbmlsolver/sbmlResults.h; reference src/sbmlResults.c,
src/odeSolver.c:441-543): species / variable compartment / variable parameter time courses and reaction fluxes.

CPU tier: the mapping is host code.  It is fed with trajectories of the oracle here and pinned against the
reference's printed results, examples/testrun.out:28-129 (species) and :132-234 (fluxes J0..J9).  The GPU tier runs
the same calls end to end (Model_odeSolver, Model_odeSolverBatch)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so


class BatchResults(C.Structure):
    _fields_ = [("nsets", C.c_int), ("nout", C.c_int), ("nvalues", C.c_int), ("names", C.c_void_p),
                ("time", C.POINTER(C.c_double)), ("value", C.POINTER(C.c_double)), ("status", C.POINTER(C.c_int)),
                ("nsens", C.c_int), ("neq", C.c_int), ("sensParam", C.c_void_p), ("sens", C.POINTER(C.c_double)),
                ("nflux", C.c_int), ("fluxNames", C.c_void_p), ("flux", C.POINTER(C.c_double))]


def _flat(out, status, times):
    """batchResults_t over numpy buffers (out[nsets][nout][nvalues])."""
    out = np.ascontiguousarray(out, dtype=np.float64)
    status = np.ascontiguousarray(status, dtype=np.int32)
    times = np.ascontiguousarray(times, dtype=np.float64)
    b = BatchResults(out.shape[0], out.shape[1], out.shape[2], None, times.ctypes.data_as(C.POINTER(C.c_double)),
                     out.ctypes.data_as(C.POINTER(C.c_double)), status.ctypes.data_as(C.POINTER(C.c_int)), 0, 0, None, None, 0, None, None)
    b._keep = (out, status, times)
    return b


def _course(res, name):
    L = so.lib()
    tc = L.SBMLResults_getTimeCourse(res, name.encode())
    assert tc, name
    assert L.TimeCourse_getName(tc).decode() == name
    return np.array([L.TimeCourse_getValue(tc, i) for i in range(L.TimeCourse_getNumValues(tc))])


def _six(x):
    return np.array([float("%g" % v) for v in np.ravel(x)]).reshape(np.shape(x))


def _gold(name):
    return np.loadtxt(os.path.join(H.ROOT, "tests", "golden", name), comments="#")


def test_mapk_results_match_the_reference_printout():
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 100, 1e-9, 1e-4, mxstep=1000)
    r = H.oracle_run(m, opt, [], np.zeros((1, 0)), 100, backend=H.gpu_tier_backend())
    flat = _flat(r["out"], r["status"], np.arange(101) * 10.0)
    arr = L.SBMLResultsArray_fromBatch(m.sbml, m.om, C.byref(flat))
    assert arr and L.SBMLResultsArray_getNumResults(arr) == 1
    res = L.SBMLResultsArray_getResults(arr, 0)
    assert L.SBMLResults_getNout(res) == 101
    tt = L.SBMLResults_getTime(res)
    assert [L.TimeCourse_getValue(tt, i) for i in (0, 1, 100)] == [0.0, 10.0, 1000.0]
    sp, fl = _gold("mapk_testrun.txt"), _gold("mapk_testrun_fluxes.txt")
    for k, name in enumerate(["MKKK", "MKKK_P", "MKK", "MKK_P", "MKK_PP", "MAPK", "MAPK_P", "MAPK_PP"]):
        assert np.array_equal(_six(_course(res, name)), sp[:, 1 + k]), name
    # fluxes: every printed digit of examples/testrun.out:132-234
    for k in range(10):
        assert np.array_equal(_six(_course(res, f"J{k}")), fl[:, 1 + k]), f"J{k}"
    assert not L.SBMLResults_getTimeCourse(res, b"uVol")          # constant compartment: no time course
    assert not L.SBMLResults_getTimeCourse(res, b"V1")            # constant parameter: no time course
    L.SBMLResultsArray_free(arr)


def test_scan_fluxes_use_the_models_constants_like_the_reference():
    """src/odeSolver.c:463-471 substitutes the MODEL's parameter values into the kinetic laws, so in a scan the flux
    courses are evaluated with nominal constants on the varied trajectory; the assigned value stored with the flat
    results is the flux at the design point's own values."""
    L = so.lib()
    m, vary, nominal = H.mapk_scan_model()
    opt = so.settings(1000.0, 20, 1e-9, 1e-6)
    vals = H.log_uniform_sets(nominal, 3, seed=11)
    r = H.oracle_run(m, opt, vary, vals, 20, backend=H.gpu_tier_backend())
    flat = _flat(r["out"], r["status"], np.arange(21) * 50.0)
    arr = L.SBMLResultsArray_fromBatch(m.sbml, m.om, C.byref(flat))
    assert L.SBMLResultsArray_getNumResults(arr) == 3
    names = m.names
    for s in range(3):
        res = L.SBMLResultsArray_getResults(arr, s)
        x = r["out"][s]
        mkk_p, mapk_p = x[:, names.index("MKK_P")], x[:, names.index("MAPK_P")]
        assert np.array_equal(_course(res, "MAPK_P"), mapk_p)
        nom = dict(zip([names[i] for i in vary], nominal))
        var = dict(zip([names[i] for i in vary], vals[s]))
        j9 = _course(res, "J9")
        want_nominal = np.array([nom["r_J9_V10"] * v / (nom["r_J9_KK10"] + v) for v in mapk_p])
        want_varied = np.array([var["r_J9_V10"] * v / (var["r_J9_KK10"] + v) for v in mapk_p])
        assert np.array_equal(j9, want_nominal)
        # row 0 of an assigned value is the one computed at reset with the model's parameters (setVariableValue
        # patches only the varied entries of row 0, src/integratorInstance.c:1571-1572)
        assert np.array_equal(x[1:, names.index("J9")], want_varied[1:]) and x[0, names.index("J9")] == want_nominal[0]
        assert not np.array_equal(want_nominal, want_varied)
        del mkk_p
    L.SBMLResultsArray_free(arr)


def test_failed_design_point_ends_at_its_last_stored_row():
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    out = np.tile(m.base_values()[None, None, :], (2, 6, 1))
    out[1, 4:, :] = np.nan
    out = np.concatenate([out, out[1:2]])                              # a third point: halted (status 0), e.g. steady state
    flat = _flat(out, np.array([0, -4, 0]), np.arange(6.0))
    arr = L.SBMLResultsArray_fromBatch(m.sbml, m.om, C.byref(flat))
    assert L.SBMLResults_getNout(L.SBMLResultsArray_getResults(arr, 0)) == 6
    assert L.SBMLResults_getNout(L.SBMLResultsArray_getResults(arr, 1)) == 4
    assert L.SBMLResults_getNout(L.SBMLResultsArray_getResults(arr, 2)) == 4
    L.SBMLResultsArray_free(arr)


def test_variable_compartments_and_parameters_get_courses(tmp_path):
    """Rate-rule driven compartment and parameter: both are 'variable' and mapped; the constant ones are not.  A
    constant compartment listed BEFORE the variable one (the case the reference's indexing overruns)."""
    M = 'xmlns="http://www.w3.org/1998/Math/MathML"'
    xml = (f'<?xml version="1.0" encoding="UTF-8"?><sbml xmlns="http://www.sbml.org/sbml/level2" level="2" version="1"><model id="vc">'
           f'<listOfCompartments><compartment id="c0" size="1"/><compartment id="c1" size="2" constant="false"/></listOfCompartments>'
           f'<listOfSpecies><species id="A" compartment="c0" initialConcentration="1"/></listOfSpecies>'
           f'<listOfParameters><parameter id="k" value="0.5"/><parameter id="p" value="1" constant="false"/></listOfParameters>'
           f'<listOfRules><rateRule variable="c1"><math {M}><ci>k</ci></math></rateRule>'
           f'<rateRule variable="p"><math {M}><apply><times/><cn>-1</cn><ci>p</ci></apply></math></rateRule></listOfRules>'
           f'<listOfReactions><reaction id="R" reversible="false"><listOfReactants><speciesReference species="A"/></listOfReactants>'
           f'<kineticLaw><math {M}><apply><times/><ci>kd</ci><ci>A</ci><ci>p</ci></apply></math>'
           f'<listOfParameters><parameter id="kd" value="0.25"/></listOfParameters></kineticLaw></reaction></listOfReactions></model></sbml>')
    p = tmp_path / "vc.xml"
    p.write_text(xml)
    L = so.lib()
    m = so.Model(str(p))
    opt = so.settings(4.0, 8, 1e-12, 1e-10)
    r = H.oracle_run(m, opt, [], np.zeros((1, 0)), 8, backend=H.gpu_tier_backend())
    flat = _flat(r["out"], r["status"], np.arange(9) * 0.5)
    arr = L.SBMLResultsArray_fromBatch(m.sbml, m.om, C.byref(flat))
    res = L.SBMLResultsArray_getResults(arr, 0)
    t = np.arange(9) * 0.5
    np.testing.assert_allclose(_course(res, "c1"), 2 + 0.5 * t, rtol=1e-8)
    np.testing.assert_allclose(_course(res, "p"), np.exp(-t), rtol=1e-7)
    a, pp = _course(res, "A"), _course(res, "p")
    assert np.array_equal(_course(res, "R"), np.array([0.25 * x * y for x, y in zip(a, pp)]))
    assert not L.SBMLResults_getTimeCourse(res, b"c0") and not L.SBMLResults_getTimeCourse(res, b"k")
    L.SBMLResultsArray_free(arr)


@pytest.mark.gpu
def test_model_odesolver_on_gpu_matches_the_reference_printout():
    """SBML_odeSolver end to end on the device (integrator instance = batch of one): examples/testrun.out digits."""
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 100, 1e-9, 1e-4, mxstep=1000)
    res = L.SBML_odeSolver(m.doc, opt)
    assert res, so.errors()
    sp, fl = _gold("mapk_testrun.txt"), _gold("mapk_testrun_fluxes.txt")
    for k, name in enumerate(["MKKK", "MKKK_P", "MKK", "MKK_P", "MKK_PP", "MAPK", "MAPK_P", "MAPK_PP"]):
        assert np.array_equal(_six(_course(res, name)), sp[:, 1 + k]), name
    for k in range(10):
        assert np.array_equal(_six(_course(res, f"J{k}")), fl[:, 1 + k]), f"J{k}"
    L.SBMLResults_free(res)


@pytest.mark.gpu
def test_model_odesolverbatch_on_gpu_returns_results_per_design_point():
    """The reference's call shape: varySettings in, SBMLResultsArray_t out; values bit-identical to the oracle."""
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    nominal = {"V1": 2.5, "Ki": 9.0, "K1": 10.0}
    rng = np.random.default_rng(4)
    pts = np.array([[nominal[k] * 2 ** rng.uniform(-1, 1) for k in nominal] + [0.25 * 2 ** rng.uniform(-1, 1)] for _ in range(5)])
    vs = L.VarySettings_allocate(4, 5)
    for k in nominal:
        assert L.VarySettings_addParameter(vs, k.encode(), None)
    assert L.VarySettings_addParameter(vs, b"V2", b"J1")
    for row in pts:
        assert L.VarySettings_addDesignPoint(vs, np.ascontiguousarray(row).ctypes.data)
    opt = so.settings(1000.0, 50, 1e-9, 1e-6)
    arr = L.Model_odeSolverBatch(m.sbml, opt, vs)
    assert arr, so.errors()
    assert L.SBMLResultsArray_getNumResults(arr) == 5
    mg = so.Model(H.model_path("mapk_cascade.xml"), globalize=[("J1", "V2")])
    vary = [mg.index(k) for k in ("V1", "Ki", "K1", "r_J1_V2")]
    ref = H.oracle_run(mg, opt, vary, pts, 50, compiled_so=H.compile_c_model(mg, "mapk_g1"), backend=H.gpu_tier_backend())
    for s in range(5):
        res = L.SBMLResultsArray_getResults(arr, s)
        assert L.SBMLResults_getNout(res) == 51
        for name in ("MKKK", "MAPK_PP"):
            assert np.array_equal(_course(res, name), ref["out"][s, :, mg.names.index(name)]), (s, name)
    L.SBMLResultsArray_free(arr)
    L.VarySettings_free(vs)


@pytest.mark.gpu
def test_device_fluxes_equal_the_host_interpreter_and_the_reference_printout():
    """odes_flux_kernel over the stored rows == evaluateAST on the host (the reference's mapping, src/odeSolver.c:519-531)
    bit for bit, on the nominal MAPK run (examples/testrun.out:132-234 digits) and on a scanned batch."""
    L = so.lib()
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 100, 1e-9, 1e-4, mxstep=1000)
    b = so.Batch(m, opt, [])
    assert b.nflux == 10
    b.reserve(32)
    b.integrate(1)
    fl = b.fluxes(1)[0]                                                # [row][reaction]
    gold = _gold("mapk_testrun_fluxes.txt")
    for k in range(10):
        assert np.array_equal(_six(fl[:, k]), gold[:, 1 + k]), f"J{k}"
    out = np.transpose(b.results(1), (2, 0, 1))                        # [set][row][value]
    flat = _flat(out, np.array([0]), np.array([L.CvodeSettings_getTime(opt, i) for i in range(101)]))
    arr = L.SBMLResultsArray_fromBatch(m.sbml, m.om, C.byref(flat))     # no device fluxes in this record: host evaluateAST
    res = L.SBMLResultsArray_getResults(arr, 0)
    for k in range(10):
        assert np.array_equal(_course(res, f"J{k}"), fl[:, k]), f"J{k}"
    L.SBMLResultsArray_free(arr)
    b.close()


@pytest.mark.gpu
def test_model_odesolverbatch_fluxes_device_path_equals_host_path(monkeypatch):
    """Model_odeSolverBatch maps fluxes on the device; ODES_HOST_FLUX=1 keeps the reference's host walk: same bits, incl.
    the scanned kinetic-law parameter (globalised r_J1_V2) and row 0."""
    L = so.lib()
    rng = np.random.default_rng(11)
    pts = np.array([[2.5 * 2 ** rng.uniform(-1, 1), 0.25 * 2 ** rng.uniform(-1, 1)] for _ in range(40)])

    def run():
        m = so.Model(H.model_path("mapk_cascade.xml"))
        vs = L.VarySettings_allocate(2, len(pts))
        assert L.VarySettings_addParameter(vs, b"V1", None) and L.VarySettings_addParameter(vs, b"V2", b"J1")
        for row in pts:
            assert L.VarySettings_addDesignPoint(vs, np.ascontiguousarray(row).ctypes.data)
        arr = L.Model_odeSolverBatch(m.sbml, so.settings(1000.0, 50, 1e-9, 1e-6), vs)
        assert arr, so.errors()
        got = np.array([[_course(L.SBMLResultsArray_getResults(arr, s), f"J{k}") for k in range(10)] for s in range(len(pts))])
        L.SBMLResultsArray_free(arr)
        L.VarySettings_free(vs)
        return got
    dev = run()
    monkeypatch.setenv("ODES_HOST_FLUX", "1")
    host = run()
    assert np.array_equal(dev, host) and np.isfinite(dev).all() and (dev[:, 1, 1:] != dev[0, 1, 1:]).any()
