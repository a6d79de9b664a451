"""The C example callers (examples/*.c) are drop-in users of the kept API: they compile against include/ with a C99
compiler, and on a GPU box reproduce the reference's golden output (examples/testrun.out:27-129 is the reference's
`integrate MAPK.xml 1000 100`)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so

EX = os.path.join(H.ROOT, "examples")


def _build():
    subprocess.check_call(["make", "-s", "-C", EX])
    return os.path.join(EX, "_build")


def test_examples_compile_as_c99():
    d = _build()
    assert all(os.path.exists(os.path.join(d, x)) for x in ("integrate_timecourse", "parameter_scan", "batch_scan_results", "odesolver_cli"))


@pytest.mark.skipif(so.lib().ODES_deviceCount() > 0, reason="needs a box without a GPU")
def test_example_fails_loudly_without_gpu():
    d = _build()
    r = subprocess.run([os.path.join(d, "integrate_timecourse"), H.model_path("mapk_cascade.xml"), "1000", "100"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "CUDA" in r.stderr and "Integration not sucessful" in r.stderr


@pytest.mark.gpu
def test_integrate_timecourse_reproduces_testrun_golden():
    d = _build()
    r = subprocess.run([os.path.join(d, "integrate_timecourse"), H.model_path("mapk_cascade.xml"), "1000", "100"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().split("\n")
    assert lines[0] == "#time MKKK MKKK_P MKK MKK_P MKK_PP MAPK MAPK_P MAPK_PP"
    gold = [l.strip() for l in open(os.path.join(H.ROOT, "tests", "golden", "mapk_testrun.txt")) if not l.startswith("#")]
    assert lines[1:] == gold                                    # every printed character of the 101 rows


@pytest.mark.gpu
def test_parameter_scan_matches_oracle():
    d = _build()
    r = subprocess.run([os.path.join(d, "parameter_scan"), H.model_path("mapk_cascade.xml"), "1000", "50", "KK2", "J1", "4", "16", "7"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.strip().split("\n")]
    names = rows[0][1:-1]
    got = np.array([[float(x) for x in row[1:-1]] for row in rows[1:]])
    m = so.Model(H.model_path("mapk_cascade.xml"), globalize=[("J1", "KK2")])
    assert names == m.names
    opt = so.settings(1000.0, 50, 1e-9, 1e-6)
    vals = np.linspace(4.0, 16.0, 7)[:, None]
    ref = H.oracle_run(m, opt, [m.index("r_J1_KK2")], vals, 50)
    want = ref["out"][:, 50, :]
    tol = np.maximum(10 * 1e-6 * np.abs(want), 10 * 1e-9)
    assert (np.abs(got - want) <= tol + 1e-9 * np.abs(want)).all()      # printed with 10 digits
    assert all(row[-1] == "0" for row in rows[1:])


@pytest.mark.gpu
def test_batch_scan_results_prints_species_per_design_point():
    """2 x 3 grid over a global and a kinetic-law parameter; the printed %g columns equal the oracle's values."""
    d = _build()
    r = subprocess.run([os.path.join(d, "batch_scan_results"), H.model_path("mapk_cascade.xml"), "500", "5", "V1", "2", "4", "2", "KK2", "J1", "4", "16", "3"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    blocks = r.stdout.strip().split("### Parameters: ")[1:]
    assert len(blocks) == 6
    m = so.Model(H.model_path("mapk_cascade.xml"), globalize=[("J1", "KK2")])
    opt = so.settings(500.0, 5, 1e-9, 1e-6)
    pts = np.array([[2 + i * 1.0, 4 + j * 4.0] for i in range(2) for j in range(3)])
    ref = H.oracle_run(m, opt, [m.index("V1"), m.index("r_J1_KK2")], pts, 5)
    for s, blk in enumerate(blocks):
        lines = blk.strip().split("\n")
        assert lines[0] == "V1=%f, KK2=%f, " % (pts[s, 0], pts[s, 1])
        assert lines[1] == "## Printing Species time courses" and lines[2].split() == ["#time"] + m.names[:8]
        for n, l in enumerate(lines[3:9]):
            want = ["%g" % (100.0 * n)] + ["%g" % v for v in ref["out"][s, n, :8]]
            assert l.split() == want, (s, n)


def _cli(*args):
    d = _build()
    r = subprocess.run([os.path.join(d, "odesolver_cli"), H.model_path("mapk_cascade.xml")] + list(args), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout.split("\n")


@pytest.mark.gpu
def test_cli_concentration_block_reproduces_the_golden_dat_file():
    """examples/MAPK_10pt.dat: `#t names / ##CONCENTRATIONS / rows / ##CONCENTRATIONS / #t names` (odeSolver/printModel.c:915-944)."""
    names = "MKKK MKKK_P MKK MKK_P MKK_PP MAPK MAPK_P MAPK_PP"
    out = _cli("--time", "100", "--printstep", "10", "--error", "1e-12", "--rerror", "1e-10")
    assert out[0] == f"#t {names} " and out[1] == "##CONCENTRATIONS"
    assert out[13] == "##CONCENTRATIONS" and out[14] == f"#t {names} " and out[15:] == ["", "", ""]
    got = np.array([[float(x) for x in l.split()] for l in out[2:13]])
    gold = np.loadtxt(os.path.join(H.ROOT, "tests", "golden", "mapk_10pt.dat.txt"), comments="#")
    assert np.array_equal(got[:, 0], gold[:, 0])
    np.testing.assert_allclose(got[:, 1:], gold[:, 1:], rtol=2e-6)      # printed from a converged run: one unit in the 6th digit
    assert (got == gold).mean() > 0.95
    assert all(l.endswith(" ") for l in out[2:13])                      # "%g " per value, as the reference prints


@pytest.mark.gpu
def test_cli_reaction_rates_block_matches_testrun_fluxes():
    """-k with the CLI defaults (atol 1e-9, rtol 1e-4) prints the kinetic laws of examples/testrun.out:132-234."""
    out = _cli("--time", "1000", "--printstep", "100", "-k")
    assert out[0] == "#t J0 J1 J2 J3 J4 J5 J6 J7 J8 J9 " and out[1] == "##REACTION RATES" and out[103] == "##REACTION RATES"
    gold = [l.strip() for l in open(os.path.join(H.ROOT, "tests", "golden", "mapk_testrun_fluxes.txt")) if not l.startswith("#")]
    assert [l.strip() for l in out[2:103]] == gold                      # every printed character


@pytest.mark.gpu
def test_cli_all_values_and_ode_values():
    m = so.Model(H.model_path("mapk_cascade.xml"))
    out = _cli("--time", "50", "--printstep", "5", "-a")
    assert out[0].split() == ["#t"] + m.names and len(out[2].split()) == 1 + m.nvalues
    out = _cli("--time", "50", "--printstep", "5", "-y")
    assert out[1] == "##ODE VALUES" and out[0].split() == ["#t"] + m.names[:8]
    row = [float(x) for x in out[2].split()]
    assert abs(row[1] + row[2]) < 1e-12                                  # d[MKKK]/dt = -d[MKKK_P]/dt
