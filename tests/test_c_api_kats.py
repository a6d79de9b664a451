"""Builds and runs tests/c/test_api_kats.c: the reference's unit tests of VarySettings_*, SolverError_* and the
SBMLResults accessors and the AST_replace* helpers (unittest/test_odeSolver.c:111-290, test_solverError.c:18-148,
test_sbmlResults.c, test_modelSimplify.c) restated as a
plain C99 program against libodes_b200.so -- a drop-in caller's view of the kept API."""
import os
import subprocess

import pytest

import helpers as H


def _build(tmp_path, name):
    exe = tmp_path / name
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-I" + os.path.join(H.ROOT, "include"), os.path.join(H.ROOT, "tests", "c", name + ".c"),
                           "-o", str(exe), "-L" + os.path.join(H.ROOT, "sbml_odesolver_b200"), "-lodes_b200",
                           "-Wl,-rpath," + os.path.join(H.ROOT, "sbml_odesolver_b200"), "-lm"])
    return exe


def test_reference_api_unit_tests_in_c(tmp_path):
    exe = _build(tmp_path, "test_api_kats")
    r = subprocess.run([str(exe), H.MODELS], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_integrator_instance_unit_tests_compile(tmp_path):
    _build(tmp_path, "test_integrator_kats")


@pytest.mark.gpu
def test_integrator_instance_unit_tests_in_c(tmp_path):
    """unittest/test_integratorInstance.c:15-150 with the default settings, through the kept single-instance API."""
    exe = _build(tmp_path, "test_integrator_kats")
    r = subprocess.run([str(exe), H.model_path("mapk_cascade.xml")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_sensitivity_unit_tests_in_c_host_part(tmp_path):
    """unittest/test_cvodeData.c:131-215 (sensitivity result accessors, directional sensitivities) and the settings."""
    exe = _build(tmp_path, "test_sens_kats")
    r = subprocess.run([str(exe), H.model_path("mapk_cascade.xml"), "cpu"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


@pytest.mark.gpu
def test_sensitivity_unit_tests_in_c(tmp_path):
    """The fixture of unittest/test_sensSolver.c:14-62 (MAPK; MAPK, MAPK_P, K1, Ki; simultaneous) through
    IntegratorInstance_cvodeOneStep and SBML_odeSolver."""
    exe = _build(tmp_path, "test_sens_kats")
    r = subprocess.run([str(exe), H.model_path("mapk_cascade.xml")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_boundary_kats_in_c_host_part(tmp_path):
    """sizeof / offsets of struct cvodeSettings, DetectNegState refused, the callback-seam exports, determinantNAST,
    Model_setValues (tests/c/test_boundary_kats.c); without a device the getters fail loudly."""
    exe = _build(tmp_path, "test_boundary_kats")
    import sbml_odesolver_b200 as so
    args = [str(exe), H.model_path("mapk_cascade.xml")] + (["cpu"] if so.lib().ODES_deviceCount() == 0 else [])
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


@pytest.mark.gpu
def test_boundary_kats_in_c(tmp_path):
    """... and with a device: the seam getters return the cached kernel handle without compiling or loading again."""
    exe = _build(tmp_path, "test_boundary_kats")
    r = subprocess.run([str(exe), H.model_path("mapk_cascade.xml")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
