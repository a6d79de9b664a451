#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the reference tree.

Run in the build container only (needs /root/reference); the GPU box and the
test-suite read the committed outputs, never /root/reference.

1. models/*.xml   -- the reference's example models, READ with this repo's SBML reader and
                     re-serialised by this repo's SBML writer (Level 2 Version 1 + MathML).
                     Content (ids, values, formulas) is the reference's test data; the text is ours.
2. mapk_testrun.txt, mapk_*pt.dat.txt -- the golden trajectory rows of examples/testrun.out:28-129 and
                     examples/MAPK_{10,100}pt.dat as plain number tables.
"""
import os
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import sbml_odesolver_b200 as so

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
MODELS = {
    "mapk_cascade.xml": "examples/MAPK.xml",
    "repressilator_ring.xml": "examples/repressilator.xml",
    "two_events_one_assignment.xml": "examples/events-2-events-1-assignment-l2.xml",
    "one_event_one_assignment.xml": "examples/events-1-event-1-assignment-l2.xml",
    "huang_ferrell_cascade.xml": "examples/huang96.xml",
    "basic_reversible.xml": "examples/basic.xml",
    "basic_forward.xml": "examples/basic-model1-forward-l2.xml",
    "goodwin_oscillator.xml": "tutorial/Goodwin_SomoS.xml",
}

L = so.lib()
for out, src in MODELS.items():
    d = L.readSBML(os.path.join(REF, src).encode())
    assert d, src
    assert L.writeSBML(d, os.path.join(OUT, "models", out).encode())
    L.SBMLDocument_free(d)
    print("wrote", out)

# golden trajectory of `integrate MAPK.xml 1000 100` (atol 1e-9, rtol 1e-4), species rows
lines = open(os.path.join(REF, "examples/testrun.out")).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("#time MKKK "))
rows = []
for l in lines[start + 1:]:
    if not re.match(r"^[0-9]", l):
        break
    rows.append(l.strip())
assert len(rows) == 101
open(os.path.join(OUT, "mapk_testrun.txt"), "w").write("# t MKKK MKKK_P MKK MKK_P MKK_PP MAPK MAPK_P MAPK_PP  (examples/testrun.out:28-129)\n" + "\n".join(rows) + "\n")
# reaction fluxes of the same run (examples/testrun.out:132-234) and the custom-grid rows (:235-243)
start = next(i for i, l in enumerate(lines) if l.startswith("#time J0 "))
rows = []
for l in lines[start + 1:]:
    if not re.match(r"^[0-9]", l):
        break
    rows.append(l.strip())
assert len(rows) == 101
open(os.path.join(OUT, "mapk_testrun_fluxes.txt"), "w").write("# t J0 .. J9  (examples/testrun.out:132-234)\n" + "\n".join(rows) + "\n")
start = [i for i, l in enumerate(lines) if l.startswith("#time MKKK ")][1]
rows = []
for l in lines[start + 1:]:
    if not re.match(r"^[0-9]", l):
        break
    rows.append(l.strip())
assert len(rows) == 7
open(os.path.join(OUT, "mapk_testrun_customgrid.txt"), "w").write("# t + 8 species on the grid 0,0.5,1,4,9,16,25  (examples/testrun.out:235-243)\n" + "\n".join(rows) + "\n")
for name in ("MAPK_10pt.dat", "MAPK_100pt.dat"):
    rows = [l.strip() for l in open(os.path.join(REF, "examples", name)) if re.match(r"^[0-9]", l)]
    open(os.path.join(OUT, name.lower() + ".txt"), "w").write(f"# rows of examples/{name}: t + 8 species\n" + "\n".join(rows) + "\n")
    print(name, len(rows))

# 5. further sections of examples/testrun.out: the printed output of the reference's own example programs
#    (examples/testrun runs them in sequence).  Kept verbatim: they are what the same programs, compiled unmodified
#    against this library (oracle/Makefile: refex), have to print.
def section(first, last, name, what):
    body = lines[first - 1:last]
    open(os.path.join(OUT, name), "w").write(f"# examples/testrun.out:{first}-{last} -- {what}\n" + "\n".join(body) + "\n")
    print(name, len(body))

section(6, 21, "testrun_printODEs.txt", "./printODEs (examples/printODEModel.c on basic-model1-forward-l2.xml)")
section(23, 234, "testrun_integrate.txt", "./integrate MAPK.xml 1000 100 (examples/integrate.c)")
section(235, 243, "testrun_defSeries.txt", "./defSeries MAPK.xml (examples/definedTimeSeries.c)")
section(498, 550, "testrun_sensitivity.txt", "./sensitivity (examples/sensitivity.c: forward sensitivities of MAPK.xml, all constants)")
section(585, 1197, "testrun_ParameterScanner.txt", "./ParameterScanner MAPK.xml 200 50 Ki 0 15 5 MAPK_PP (examples/ParameterScanner.c)")
section(353, 410, "testrun_analyzeJacobian.txt", "./analyzeJacobian (examples/analyzeJacobian.c: symbolic Jacobian of MAPK.xml, its signs at t = 0 and t = 1000)")
section(411, 497, "testrun_analyzeSens.txt", "./analyzeSens (examples/analyzeSensitivity.c: the parametric matrix of MAPK.xml as printed expressions)")
section(1198, 1608, "testrun_Sense.txt", "./Sense MAPK.xml 200 7.5 p Ki MKKK_P MAPK_PP (examples/Sense.c)")
section(320, 350, "testrun_changeIntInst.txt", "./changeIntInst (examples/ChangingIntegratorInstance.c on basic-model1-forward-l2.xml: two instances, the second one with values changed on the way)")
section(1826, 1863, "testrun_senstest_analytic.txt", "./senstest, first part (examples/senstest.c:45-150: repeated runs, default and selected sensitivities with analytic matrices; the rest uses CVODES' internal difference quotients)")
section(244, 318, "testrun_simpleIntInst.txt", "./simpleIntInst (examples/simpleIntegratorInstance.c on basic-model1-forward-l2.xml: indefinite Adams-Moulton run from t = 0.05, then a finite BDF run with other settings)")
