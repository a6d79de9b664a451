"""The reference's OWN unit tests (unittest/test_*.c, 5.6 k lines written for the `check` framework), compiled UNMODIFIED
against this library and run.

oracle/Makefile (target refut, part of build() where /root/reference exists) compiles one program per suite from the
sources where they lie, with tests/c/compat/check.h standing in for the absent framework (fork per test, assertion
macros, suite plumbing); the binaries travel in oracle/_ref/unittests like the vendored CVODE build.  Every test of a
suite has to pass except the ones listed in KNOWN with the reason.  Suites that integrate need the GPU."""
import os
import re
import subprocess

import pytest

import helpers as H
from test_reference_examples import MODELS

UT = os.path.join(H.ROOT, "oracle", "_ref", "unittests")
have = pytest.mark.skipif(not os.path.exists(os.path.join(UT, "odeModel")), reason="oracle/_ref/unittests not built (needs the reference tree at build time)")

# tests that cannot pass here, and why
KNOWN = {
    # the reference prints its determinant tree with empty operands (" - beta * ( - beta * ...", "" for singular systems); this
    # library builds a well-formed Laplace expansion instead (csrc/ode_model.c: ODEModel_constructDeterminant)
    "test_ODEModel_constructDeterminant_basic_model1_forward_l2", "test_ODEModel_constructDeterminant_basic",
    "test_ODEModel_constructDeterminant_events_1_event_1_assignment_l2", "test_ODEModel_constructDeterminant_events_2_events_1_assignment_l2",
    "test_ODEModel_constructDeterminant_repressilator",
    # compares the kinetic-law text of the reference's Level-1 file character for character ("V1*MKKK/(...)" without blanks);
    # the committed model files are re-serialised Level-2 documents (tests/golden/make_fixtures.py), their text is canonical
    "test_parseModel_MAPK",
}
CPU_SUITES = {"charBuffer": 5, "cvodeData": 11, "integratorSettings": 9, "modelSimplify": 8, "odeConstruct": 16, "odeModel": 35, "processAST": 11,
              "sbml": 7, "sbmlResults": 21, "solverError": 7, "util": 4}
GPU_SUITES = ["integratorInstance", "odeSolver"]


def _run_suite(tmp_path, suite):
    d = tmp_path / "refmodels"
    d.mkdir(exist_ok=True)
    for ref_name, ours in MODELS.items():
        if not (d / ref_name).exists():
            os.symlink(H.model_path(ours), d / ref_name)
    r = subprocess.run([os.path.join(UT, suite)], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    results = dict(re.findall(r"^[^:\n]+:[^:\n]+:(\w+): (ok|FAILED)$", r.stdout, re.M))
    return r, results


@pytest.mark.skipif(not os.path.isdir("/root/reference/unittest"), reason="needs the reference tree (build container only)")
def test_reference_unit_test_sources_compile_unmodified():
    subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "oracle"), "refut"])
    assert all(os.path.exists(os.path.join(UT, s)) for s in list(CPU_SUITES) + GPU_SUITES)


@have
@pytest.mark.parametrize("suite", sorted(CPU_SUITES))
def test_reference_unit_tests_pass_on_the_host(tmp_path, suite):
    r, results = _run_suite(tmp_path, suite)
    assert len(results) == CPU_SUITES[suite], (r.stdout[-2000:], r.stderr[-2000:])
    failed = {t for t, v in results.items() if v != "ok"}
    assert failed <= KNOWN, (failed - KNOWN, r.stderr[-3000:])


@have
@pytest.mark.gpu
@pytest.mark.parametrize("suite", GPU_SUITES)
def test_reference_unit_tests_pass_on_the_gpu(tmp_path, suite):
    r, results = _run_suite(tmp_path, suite)
    assert results, (r.stdout[-2000:], r.stderr[-2000:])
    failed = {t for t, v in results.items() if v != "ok"}
    print(suite, len(results), "tests,", len(failed), "failed")
    assert failed <= KNOWN, (failed - KNOWN, r.stderr[-3000:])
