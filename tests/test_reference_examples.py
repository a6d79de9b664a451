"""DROP-IN CHECK with the reference's own programs and the reference's own golden output.

oracle/Makefile (target refex, part of build() where /root/reference exists) compiles the reference's example sources
UNMODIFIED, from where they lie under /root/reference/examples, against this repository's headers (include/sbmlsolver, the
libSBML-named shims include/sbml/*.h, and a stand-in for the out-of-scope graph drawing header) and links them with
libodes_b200.so.  The binaries travel to the GPU box in oracle/_ref/examples (git-ignored like the vendored CVODE build);
nothing here reads /root/reference at run time.  examples/testrun.out is the printed output of these programs
(examples/testrun runs them in sequence); the sections are committed verbatim under tests/golden/testrun_*.txt
(tests/golden/make_fixtures.py).  A program run through this library on the GPU has to print the reference's text."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so

REFEX = os.path.join(H.ROOT, "oracle", "_ref", "examples")
GOLD = os.path.join(H.ROOT, "tests", "golden")
PROGRAMS = ["integrate", "batch_integrate", "ParameterScanner", "definedTimeSeries", "sensitivity", "senstest", "Sense", "simpleIntegratorInstance",
            "ChangingIntegratorInstance", "ChangingParameterIntegrator", "SharingIntegratorInstance", "TwinIntegratorInstance", "integrateODEs",
            "analyzeJacobian", "analyzeSensitivity", "printODEModel", "bistability", "testCompiler"]
# the reference's model file names -> this repository's re-serialised copies (tests/golden/models)
MODELS = {"MAPK.xml": "mapk_cascade.xml", "basic-model1-forward-l2.xml": "basic_forward.xml", "basic.xml": "basic_reversible.xml",
          "events-2-events-1-assignment-l2.xml": "two_events_one_assignment.xml", "events-1-event-1-assignment-l2.xml": "one_event_one_assignment.xml",
          "repressilator.xml": "repressilator_ring.xml", "huang96.xml": "huang_ferrell_cascade.xml"}
have = pytest.mark.skipif(not os.path.exists(os.path.join(REFEX, "integrate")), reason="oracle/_ref/examples not built (needs the reference tree at build time)")


def _gold(name):
    return [l.rstrip() for l in open(os.path.join(GOLD, name)).read().split("\n")[1:]]


def _run(tmp_path, prog, *args):
    for ref_name, ours in MODELS.items():
        dst = tmp_path / ref_name
        if not dst.exists():
            os.symlink(H.model_path(ours), dst)
    r = subprocess.run([os.path.join(REFEX, prog), *args], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    return r, [l.rstrip() for l in r.stdout.split("\n")]


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="needs the reference tree (build container only)")
def test_reference_example_sources_compile_unmodified_against_this_library():
    subprocess.check_call(["make", "-s", "-C", os.path.join(H.ROOT, "oracle"), "refex"])
    assert all(os.path.exists(os.path.join(REFEX, p)) for p in PROGRAMS)


@have
def test_printODEModel_prints_the_reference_text(tmp_path):
    """No device needed: model construction only (examples/testrun.out:6-21)."""
    r, out = _run(tmp_path, "printODEModel")
    assert r.returncode == 0, r.stderr
    gold = _gold("testrun_printODEs.txt")
    body = [l for l in out if l.strip()]
    assert body == [l for l in gold if l.strip()]


@have
@pytest.mark.gpu
def test_integrate_prints_the_reference_text(tmp_path):
    """./integrate MAPK.xml 1000 100: all species and flux time courses, every printed character (testrun.out:23-234)."""
    r, out = _run(tmp_path, "integrate", "MAPK.xml", "1000", "100")
    assert r.returncode == 0, r.stderr
    gold = _gold("testrun_integrate.txt")
    assert [l for l in out if l.strip()] == [l for l in gold if l.strip()]


@have
@pytest.mark.gpu
def test_definedTimeSeries_prints_the_reference_text(tmp_path):
    r, out = _run(tmp_path, "definedTimeSeries", "MAPK.xml")
    assert r.returncode == 0, r.stderr
    gold = [l for l in _gold("testrun_defSeries.txt") if l.strip()]
    body = [l for l in out if l.strip()]
    assert body[-len(gold):] == gold


@have
@pytest.mark.gpu
def test_sensitivity_prints_the_reference_text(tmp_path):
    """./sensitivity: forward sensitivities of MAPK.xml to all constants through IntegratorInstance_integrateOneStep,
    IntegratorInstance_dumpPSensitivities and CvodeResults_getSensitivity, then the re-run with V1 + 1
    (testrun.out:498-550) -- every printed digit."""
    r, out = _run(tmp_path, "sensitivity")
    assert r.returncode == 0, r.stderr
    gold = [l for l in _gold("testrun_sensitivity.txt") if l.strip()]
    assert [l for l in out if l.strip()] == gold


def _golden_sensitivity_rows():
    gold = _gold("testrun_sensitivity.txt")
    on_the_fly = [l.split() for l in gold[2:13]]                       # t + d y_i / d K1, i = 0..7
    i0 = gold.index("#time  MAPK_P  uVol V1 Ki K1")
    per_variable = [l.split() for l in gold[i0 + 1:i0 + 12]]           # t, MAPK_P, d MAPK_P / d (uVol V1 Ki K1)
    return on_the_fly, per_variable


def _check_against_golden(out, sens, tp):
    on_the_fly, per_variable = _golden_sensitivity_rows()
    for k in range(11):
        assert ["%g" % tp[k]] + ["%g" % v for v in sens[k, 3, :]] == on_the_fly[k]
        assert ["%g" % tp[k], "%g" % out[k, 6]] + ["%g" % sens[k, j, 6] for j in range(4)] == per_variable[k]


@pytest.mark.skipif(not H.have_sens_ref(), reason="oracle/_ref/libcvodes_ref.so not built")
def test_sensitivity_oracle_and_emulated_kernel_reproduce_the_golden_digits():
    """PIN of the sensitivity oracle: the vendored CVODES 2.3.0 driven like src/sensSolver.c:197-346 prints the digits of
    examples/testrun.out:499-526, and so does the kernel's per-trajectory program (host emulation)."""
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(300.0, 10, 1e-9, 1e-4, mxstep=10 ** 9, sensitivity=1)       # examples/sensitivity.c:62-69
    b = so.Batch(m, opt, [], store=list(range(m.neq)))
    tp = H.tpoints(opt, 10)
    none = np.zeros((1, 0))
    ref = H.sens_ref_run(m, opt, b, [], none, tp, 10 ** 9, 1e-4, 1e-9, tag="mapk_nominal_sens")
    _check_against_golden(ref["out"][0], ref["sens"][0], tp)
    base, trig = b.base_values()
    emu = H.Emu(b, "mapk_nominal_sens").run(base, trig, none, tp, 10 ** 9, 1e-4, 1e-9)
    b.close()
    _check_against_golden(np.transpose(emu["out"], (2, 0, 1))[0], emu["sens"][0], tp)


@pytest.mark.gpu
def test_sensitivity_kernel_reproduces_the_golden_digits():
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(300.0, 10, 1e-9, 1e-4, mxstep=10 ** 9, sensitivity=1)
    b = so.Batch(m, opt, [], store=list(range(m.neq)))
    b.reserve(1)
    b.integrate(1)
    out, sens = np.transpose(b.results(1), (2, 0, 1))[0], b.sensitivities(1)[0]
    b.close()
    _check_against_golden(out, sens, H.tpoints(opt, 10))


@have
@pytest.mark.gpu
def test_ParameterScanner_prints_the_reference_text(tmp_path):
    """./ParameterScanner MAPK.xml 200 50 Ki 0 15 5 MAPK_PP (testrun.out:585-1197): INDEFINITE integration, one step of
    length 50 per IntegratorInstance_integrateOneStep, 200 steps per value of Ki, reset + setVariableValue in between; Ki = 0
    makes the first step fail ("Parameter = 0") and the scan goes on.  Every printed digit of the 3 x 201 rows."""
    r, out = _run(tmp_path, "ParameterScanner", "MAPK.xml", "200", "50", "Ki", "0", "15", "5", "MAPK_PP")
    assert r.returncode == 0, r.stderr
    gold = _gold("testrun_ParameterScanner.txt")
    while gold and not gold[-1].strip():
        gold.pop()
    while out and not out[-1].strip():
        out.pop()
    assert out == gold


def _same_text(out, gold_name, skip_head=0):
    gold = [l for l in _gold(gold_name) if l.strip()]
    body = [l for l in out if l.strip()]
    assert body[skip_head:] == gold[skip_head:], "\n".join(f"{a!r} != {b!r}" for a, b in zip(body, gold) if a != b)[:2000]


@have
def test_analyzeSensitivity_prints_the_reference_text(tmp_path):
    """No device needed: ODESense_constructMatrix and the printed expressions of every d f_i / d p_j of MAPK.xml
    (examples/testrun.out:411-497) -- a known-answer test of the parametric matrix from the reference's own output."""
    r, out = _run(tmp_path, "analyzeSensitivity")
    assert r.returncode == 0, r.stderr
    _same_text(out, "testrun_analyzeSens.txt")


@have
@pytest.mark.gpu
def test_analyzeJacobian_prints_the_reference_text(tmp_path):
    """The symbolic Jacobian, its signs at the initial conditions and after integrating to t = 1000 (testrun.out:356-410)."""
    r, out = _run(tmp_path, "analyzeJacobian")
    assert r.returncode == 0, r.stderr
    _same_text(out, "testrun_analyzeJacobian.txt")


@have
@pytest.mark.gpu
def test_Sense_prints_the_reference_text(tmp_path):
    """./Sense MAPK.xml 200 7.5 p Ki MKKK_P MAPK_PP: sensitivity time courses as gnuplot data (testrun.out:1198-1620)."""
    r, out = _run(tmp_path, "Sense", "MAPK.xml", "200", "7.5", "p", "Ki", "MKKK_P", "MAPK_PP")
    assert r.returncode == 0, r.stderr
    _same_text(out, "testrun_Sense.txt")


@have
@pytest.mark.gpu
def test_ChangingIntegratorInstance_prints_the_reference_text(tmp_path):
    """Two instances of one model stepped side by side, values of the second one changed on the way (setVariableValue
    during a run re-initialises the solver at that time; testrun.out:320-350)."""
    r, out = _run(tmp_path, "ChangingIntegratorInstance")
    assert r.returncode == 0, r.stderr
    gold = [l for l in _gold("testrun_changeIntInst.txt") if l.strip()]
    body = [l for l in out if l.strip()]
    assert body[-len(gold):] == gold


@have
@pytest.mark.gpu
def test_senstest_prints_the_reference_text_for_the_analytic_cases(tmp_path):
    """./senstest up to the point where it frees the Jacobian and asks for CVODES' internal difference quotients (not
    built: the plan creation fails loudly there): repeated integrate / reset, sensitivities switched on in place, the
    default selection and a selection of parameters AND variables (initial-condition sensitivities), rtol 1e-10
    (testrun.out:1826-1863)."""
    r, out = _run(tmp_path, "senstest")
    gold = [l for l in _gold("testrun_senstest_analytic.txt") if l.strip()]
    body = [l for l in out if l.strip()]
    assert body[:len(gold)] == gold, "\n".join(f"{a!r} != {b!r}" for a, b in zip(body, gold) if a != b)[:2000]


@have
@pytest.mark.gpu
@pytest.mark.parametrize("prog,args", [("TwinIntegratorInstance", []), ("SharingIntegratorInstance", []), ("ChangingParameterIntegrator", []),
                                       ("integrateODEs", []),
                                       ("batch_integrate", ["MAPK.xml", "1000", "2", "1", "10", "5", "V1", "1", "10", "10", "k4", "J3"]),
                                       ("bistability", ["MAPK.xml", "MAPK_PP", "Ki", "5", "15", "5", "K1", "5", "15", "5", "100"])])
def test_further_reference_programs_run_to_completion(tmp_path, prog, args):
    """Programs without a golden section in examples/testrun.out: they have to run through this library and exit cleanly
    (two and shared integrator instances on the events models, parameters changed during a run, ODEs given as formulas
    without an SBML file, the reference's batch driver on a 5 x 10 grid with a kinetic-law parameter, a 2-D scan to steady
    states)."""
    r, out = _run(tmp_path, prog, *args)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert len([l for l in out if l.strip()]) > 3


@have
@pytest.mark.gpu
def test_simpleIntegratorInstance_prints_the_reference_text(tmp_path):
    """Settings dumps, an INDEFINITE ADAMS-MOULTON run (orders up to 12, atol 1e-20, rtol 1e-14) started at t = 0.05 with
    IntegratorInstance_setInitialTime, then IntegratorInstance_set with new settings and a finite BDF run
    (testrun.out:244-318) -- every printed digit."""
    r, out = _run(tmp_path, "simpleIntegratorInstance")
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    gold = [l for l in _gold("testrun_simpleIntInst.txt") if l.strip()]
    body = [l for l in out if l.strip()]
    assert len(body) == len(gold)
    for a, b in zip(body, gold):
        ta, tb = a.split(), b.split()
        if len(tb) == 6 and tb[0][0].isdigit():
            # a data row: time S1 S2 R1 R2 c.  With StoreResults == 0 the reference does not refresh the assignment rules at
            # an output time (src/integratorInstance.c:1127-1147), so its R1 column shows the rate of the solver's last
            # INTERNAL right-hand-side evaluation; this library prints the rate at the output time.  Everything else has to
            # agree to the last printed digit -- including the second run, which announces BDF and integrates with the
            # Adams-Moulton memory of the first (src/cvodeSolver.c:541-547)
            assert ta[:3] + ta[4:] == tb[:3] + tb[4:], (a, b)
        else:
            assert a == b
