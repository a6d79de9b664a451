"""Pins the ORACLE (oracle/: restated CVODE 2.3.0 BDF + reference driver semantics) against the golden
vectors the reference itself ships, and against the reference's vendored CVODE built into oracle/_ref:

  examples/testrun.out:28-129   `integrate MAPK.xml 1000 100`, atol 1e-9, rtol 1e-4   (tests/golden/mapk_testrun.txt)
  examples/MAPK_10pt.dat, MAPK_100pt.dat    converged nominal MAPK on t in [0,100]       (tests/golden/mapk_*pt.dat.txt)
  SURVEY appendix D             event sequence of events-2-events-1-assignment-l2.xml (analytic: S1 decays as exp(-t))
"""
import os

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so


def _table(name):
    return np.loadtxt(os.path.join(H.ROOT, "tests", "golden", name), comments="#")


def _six_digits(x):
    """%g formatting as in the reference's printouts (6 significant digits)."""
    return np.array([[float("%g" % v) for v in row] for row in x])


def _nominal(model, opt, nout, **kw):
    return H.oracle_run(model, opt, [], np.zeros((1, 0)), nout, **kw)


@pytest.mark.parametrize("backend", ["port", "ref"])
def test_mapk_testrun_golden(backend):
    if backend == "ref" and not H.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    gold = _table("mapk_testrun.txt")
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 100, 1e-9, 1e-4, mxstep=1000)
    r = _nominal(m, opt, 100, backend=H.REF if backend == "ref" else H.PORT)
    assert r["status"][0] == 0
    got = _six_digits(r["out"][0, :, :8])
    assert np.array_equal(gold[:, 0], np.arange(101) * 10.0)
    # every printed digit of every row
    assert np.array_equal(got, gold[:, 1:]), np.abs(got - gold[:, 1:]).max()


def test_mapk_testrun_fluxes_are_assigned_values():
    """examples/testrun.out:132-234 prints J0..J9 = the assigned kinetic-law values stored with the results."""
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 100, 1e-9, 1e-4, mxstep=1000)
    r = _nominal(m, opt, 100)["out"][0]
    x = r[:, :8]
    uvol, v1, ki, k1 = r[0, 18:22]
    j0 = v1 * x[:, 0] / ((1 + (x[:, 7] / ki) ** 1) * (k1 + x[:, 0]))
    np.testing.assert_allclose(r[:, 8], j0, rtol=1e-14)
    j9 = 0.5 * x[:, 6] / (15 + x[:, 6])
    np.testing.assert_allclose(r[:, 17], j9, rtol=1e-14)


@pytest.mark.parametrize("name,nout", [("mapk_10pt.dat.txt", 10), ("mapk_100pt.dat.txt", 100)])
def test_mapk_converged_dat_files(name, nout):
    gold = _table(name)
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(100.0, nout, 1e-12, 1e-10)
    r = _nominal(m, opt, nout)
    got = _six_digits(r["out"][0, :, :8])
    # the .dat files were printed from a converged run: allow one unit in the 6th digit
    np.testing.assert_allclose(got, gold[:, 1:], rtol=2e-6)
    assert (got == gold[:, 1:]).mean() > 0.95


def test_port_equals_vendored_cvode_bit_for_bit():
    if not H.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    m, vary, nominal = H.mapk_scan_model()
    opt = so.settings(1000.0, 100, 1e-9, 1e-6)
    vals = H.log_uniform_sets(nominal, 24, seed=3)
    a = H.oracle_run(m, opt, vary, vals, 100, backend=H.PORT)
    b = H.oracle_run(m, opt, vary, vals, 100, backend=H.REF)
    assert np.array_equal(a["out"], b["out"])
    assert np.array_equal(a["stats"], b["stats"]) and np.array_equal(a["status"], b["status"])
    for name, kw in (("repressilator_ring.xml", dict(end=100.0, nout=100)), ("huang_ferrell_cascade.xml", dict(end=100.0, nout=20))):
        mm = so.Model(H.model_path(name))
        o = so.settings(kw["end"], kw["nout"], 1e-9, 1e-6)
        a = _nominal(mm, o, kw["nout"], backend=H.PORT)
        b = _nominal(mm, o, kw["nout"], backend=H.REF)
        assert np.array_equal(a["out"], b["out"]) and np.array_equal(a["stats"], b["stats"])


def test_no_jacobian_uses_difference_quotients():
    m = so.Model(H.model_path("repressilator_ring.xml"))
    a = _nominal(m, so.settings(10.0, 10, 1e-9, 1e-4, use_jacobian=0), 10)
    b = _nominal(m, so.settings(10.0, 10, 1e-9, 1e-4, use_jacobian=1), 10)
    assert a["stats"][0, 3] >= 1 and b["stats"][0, 3] >= 1     # nje: both evaluate Jacobians (CVODE 2.3.0 counts DQ f calls apart)
    np.testing.assert_allclose(a["out"], b["out"], rtol=1e-3)
    # SURVEY appendix D probe (DQ Jacobian, rtol 1e-4): t = 1 row
    want = [0.910799, 0.454976, 0.0882343, 0.796494, 0.20547, 0.0756264]
    np.testing.assert_allclose(a["out"][0, 1, :6], want, rtol=2e-5)


def test_event_sequence_nominal():
    """Events are examined at output times only, edge-triggered, assignments applied before the row is stored,
    solver re-initialised after each firing (src/integratorInstance.c:1029-1237)."""
    m = so.Model(H.model_path("two_events_one_assignment.xml"))
    opt = so.settings(10.0, 100, 1e-9, 1e-6)
    r = _nominal(m, opt, 100)
    fired = r["fired"][0]
    seq = [(e, i) for i in range(101) for e in (0, 1) if fired[i] >> e & 1]
    assert seq == [(1, 7), (0, 24), (1, 25), (1, 34), (0, 48), (1, 51), (1, 63), (0, 72), (1, 77), (1, 95), (0, 96)]
    np.testing.assert_allclose(r["out"][0, 100, :2], [0.670322, 0.339219], rtol=2e-6)
    # analytic check: between resets S1 decays as exp(-t); first reset of S1 at index 24 -> S1 = 1 stored in that row
    s1 = r["out"][0, :, 0]
    np.testing.assert_allclose(s1[1:24], np.exp(-0.1 * np.arange(1, 24)), rtol=2e-5)
    assert s1[24] == 1.0
    np.testing.assert_allclose(s1[25:48], np.exp(-0.1 * np.arange(1, 24)), rtol=2e-5)


def test_halt_on_event_marks_remaining_rows():
    m = so.Model(H.model_path("two_events_one_assignment.xml"))
    opt = so.settings(10.0, 100, 1e-9, 1e-6, halt_on_event=1)
    r = _nominal(m, opt, 100)
    assert r["fired"][0, 7] == 2 and not np.isnan(r["out"][0, 7]).any()
    assert np.isnan(r["out"][0, 8:]).all()


def test_too_much_work_flag():
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 2, 1e-12, 1e-10, mxstep=20)
    r = _nominal(m, opt, 2)
    assert r["status"][0] == -4 and np.isnan(r["out"][0, 1:]).all() and not np.isnan(r["out"][0, 0]).any()


def test_threads_do_not_change_results():
    m, vary, nominal = H.mapk_scan_model()
    opt = so.settings(1000.0, 20, 1e-9, 1e-6)
    vals = H.log_uniform_sets(nominal, 16, seed=5)
    a = H.oracle_run(m, opt, vary, vals, 20, threads=1)
    b = H.oracle_run(m, opt, vary, vals, 20, threads=4)
    assert np.array_equal(a["out"], b["out"])
