#!/usr/bin/env python
"""Extracts the known-answer tests the reference's own unit suite holds for the hot path into
tests/golden/kats.json (committed).  Run in the build container only (needs /root/reference).

Sources (reference file:line):
  unittest/test_odeModel.c:155-358    ODE structure: neq/nass/nconst, name order, equation strings
  unittest/test_odeModel.c:564-719    Jacobian sparsity: (i, j) of every structurally non-zero entry, in order
  unittest/test_processAST.c:135-188  evaluateAST values
  unittest/test_processAST.c:248-299  differentiateAST strings (d/dx)
  unittest/test_processAST.c:333-367  simplifyAST strings
  unittest/test_integratorSettings.c  defaults of CvodeSettings_create / createWithTime
  unittest/test_sbml.c:16-240          parseModel: object counts, parameter ids, reaction ids and kinetic-law formulas
  unittest/test_odeConstruct.c:34-91   Model_getValueById / Model_setValue
  unittest/test_odeConstruct.c:93-412  Model_reduceToOdes: object counts, parameter order, rule formulas
  unittest/test_odeConstruct.c:414-520 Species_odeFromReactions: the ODE of every species as a formula string
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kats.json")

# reference example file -> fixture name under tests/golden/models (written by make_fixtures.py)
MODEL_ALIAS = {
    "MAPK.xml": "mapk_cascade.xml", "repressilator.xml": "repressilator_ring.xml",
    "events-2-events-1-assignment-l2.xml": "two_events_one_assignment.xml",
    "events-1-event-1-assignment-l2.xml": "one_event_one_assignment.xml",
    "huang96.xml": "huang_ferrell_cascade.xml", "basic.xml": "basic_reversible.xml",
    "basic-model1-forward-l2.xml": "basic_forward.xml",
}


def tests_of(path):
    src = open(path).read()
    for m in re.finditer(r"START_TEST\((\w+)\)(.*?)END_TEST", src, re.S):
        yield m.group(1), m.group(2)


def cstr(s):
    return bytes(s, "utf-8").decode("unicode_escape")


kats = {"structure": [], "jacobian": [], "evaluate": [], "differentiate": [], "simplify": [], "reduce": [], "species_ode": [], "get_value": [], "set_value": [], "parse": []}

for name, body in tests_of(os.path.join(REF, "unittest/test_odeModel.c")):
    fm = re.search(r'EXAMPLES_FILENAME\("([^"]+)"\)', body)
    if not fm or fm.group(1) not in MODEL_ALIAS:
        continue
    model = MODEL_ALIAS[fm.group(1)]
    if name.startswith("test_ODEModel_createFromFile_"):
        e = {"test": name, "model": model, "counts": {}, "odes": [], "assignments": [], "constants": []}
        for fn, val in re.findall(r"ck_assert_int_eq\(ODEModel_(get\w+)\(model\), (\d+)\)", body):
            e["counts"][fn] = int(val)
        for idx, nm, eq in re.findall(r'CHECK_EQ\((\d+), "([^"]+)", "([^"]+)"\)', body):
            e["odes"].append([int(idx), nm, eq])
        for idx, k, nm, eq in re.findall(r'CHECK_ASSIGNMENT\((\d+), (\d+), "([^"]+)", "([^"]+)"\)', body):
            e["assignments"].append([int(idx), int(k), nm, eq])
        for idx, k, nm in re.findall(r'CHECK_CONSTANT\((\d+), (\d+), "([^"]+)"\)', body):
            e["constants"].append([int(idx), int(k), nm])
        kats["structure"].append(e)
    elif name.startswith("test_ODEModel_constructJacobian_"):
        nnz = re.search(r"ck_assert_int_eq\(ODEModel_getNumJacobiElements\(model\), (\d+)\)", body)
        elems = [[int(n), int(i), int(j)] for n, i, j in re.findall(r"CHECK_JACOBI_ELEMENT\((\d+), (\d+), (\d+)\)", body)]
        if nnz:
            kats["jacobian"].append({"test": name, "model": model, "nnz": int(nnz.group(1)), "elements": elems})

C_EXPR = {"M_PI": "math.pi", "log": "math.log", "sqrt": "math.sqrt", "cosh": "math.cosh", "sinh": "math.sinh"}
for name, body in tests_of(os.path.join(REF, "unittest/test_processAST.c")):
    if name == "test_evaluateAST":
        for inp, exp in re.findall(r'CHECK_EVAL\("([^"]+)", (.+)\);', body):
            kats["evaluate"].append([inp, exp.strip()])          # expected kept as the C expression text
    elif name == "test_differentiateAST":
        for inp, exp in re.findall(r'CHECK_DIFF\("([^"]+)", "([^"]+)"\)', body):
            kats["differentiate"].append([inp, exp])
    elif name == "test_simplifyAST":
        for inp, exp in re.findall(r'CHECK_SIMPLIFY\("([^"]+)", "([^"]+)"\)', body):
            kats["simplify"].append([inp, exp])

for name, body in tests_of(os.path.join(REF, "unittest/test_odeConstruct.c")):
    fm = re.search(r'EXAMPLES_FILENAME\("([^"]+)"\)', body)
    if not fm or fm.group(1) not in MODEL_ALIAS:
        continue
    model = MODEL_ALIAS[fm.group(1)]
    if name.startswith("test_Model_reduceToOdes_"):
        counts = {fn: int(v) for fn, v in re.findall(r"ck_assert\(Model_(getNum\w+)\(omodel\) == (\d+)\)", body)}
        params = [[int(i), nm] for i, nm in re.findall(r'CHECK_PARAMETER\(omodel, (\d+), "([^"]+)"\)', body)]
        rules = [[int(i), f] for i, f in re.findall(r'CHECK_RULE\(omodel, (\d+), "([^"]+)"\)', body)]
        kats["reduce"].append({"test": name, "model": model, "counts": counts, "parameters": params, "rules": rules})
    elif name.startswith("test_Species_odeFromReactions_"):
        kats["species_ode"].append({"test": name, "model": model, "odes": [[int(i), f] for i, f in re.findall(r'check_species\((\d+), "([^"]+)"\)', body)]})
    elif name == "test_Model_getValueById":
        pairs = re.findall(r'Model_getValueById\(model, "([^"]+)"\);\s*ck_assert\(v == ([^)]+)\)', body)
        kats["get_value"].append({"model": model, "values": [[k, float(v)] for k, v in pairs]})
    elif name == "test_Model_setValue":
        calls = re.findall(r'Model_setValue\(model, "([^"]+)", (NULL|"[^"]+"), ([0-9.]+)\);\s*ck_assert\(r == (\d)\)', body)
        kats["set_value"].append({"model": model, "calls": [[i, None if r == "NULL" else r.strip('"'), float(v), int(ok)] for i, r, v, ok in calls]})

for name, body in tests_of(os.path.join(REF, "unittest/test_sbml.c")):
    fm = re.search(r'EXAMPLES_FILENAME\("([^"]+)"\)', body)
    if not fm or fm.group(1) not in MODEL_ALIAS or not name.startswith("test_parseModel_"):
        continue
    counts = {fn: int(v) for fn, v in re.findall(r"ck_assert\(Model_(getNum\w+)\(model\) == (\d+)\)", body)}
    params = [[int(i), nm] for i, nm in re.findall(r'CHECK_PARAMETER\(model, (\d+), "([^"]+)"\)', body)]
    reactions = [[int(i), rid, f] for i, rid, f in re.findall(r'CHECK_REACTION\(model, (\d+), "([^"]+)", "([^"]+)"\)', body)]
    kats["parse"].append({"test": name, "model": MODEL_ALIAS[fm.group(1)], "counts": counts, "parameters": params, "reactions": reactions})

json.dump(kats, open(OUT, "w"), indent=1)
print({k: len(v) for k, v in kats.items()})
