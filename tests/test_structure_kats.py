"""Host front-end against the reference's own known-answer tests (tests/golden/kats.json, extracted by
tests/golden/make_kats.py from unittest/test_odeModel.c:155-358, :564-719 and unittest/test_processAST.c:135-367):
bit-exact ODE / assignment structure, species indexing and Jacobian sparsity order."""
import ctypes as C
import json
import math
import os

import pytest

import helpers as H
import sbml_odesolver_b200 as so

KATS = json.load(open(os.path.join(H.ROOT, "tests", "golden", "kats.json")))


@pytest.mark.parametrize("kat", KATS["structure"], ids=lambda k: k["model"])
def test_ode_structure(kat):
    L = so.lib()
    m = so.Model(H.model_path(kat["model"]))
    for fn, want in kat["counts"].items():
        assert getattr(L, "ODEModel_" + fn)(m.om) == want, fn
    assert L.ODEModel_hasVariable(m.om, b"no_such_variable") == 0
    assert not L.ODEModel_hasCycle(m.om)
    for idx, name, eq in kat["odes"]:
        assert m.names[idx] == name
        assert m.index(name) == idx
        assert m.equation(idx) == eq
    for idx, _k, name, eq in kat["assignments"]:
        assert m.names[idx] == name
        assert m.equation(idx) == eq
    for idx, _k, name in kat["constants"]:
        assert m.names[idx] == name


@pytest.mark.parametrize("kat", KATS["jacobian"], ids=lambda k: k["model"])
def test_jacobian_sparsity(kat):
    m = so.Model(H.model_path(kat["model"]))
    js = m.jacobian_structure()
    assert len(js) == kat["nnz"]
    for n, i, j in kat["elements"]:
        assert (js[n][0], js[n][1]) == (i, j)
    # row-major order of the non-zero list (src/odeModel.c:1782-1800)
    assert [(i, j) for i, j, _ in js] == sorted((i, j) for i, j, _ in js)


def _parse(s):
    n = so.lib().SBML_parseFormula(s.encode())
    assert n, s
    return n


def _str(node):
    return so.take_string(so.lib().SBML_formulaToString(node))


@pytest.mark.parametrize("inp,want", KATS["evaluate"])
def test_evaluate_ast(inp, want):
    L = so.lib()
    n = _parse(inp)
    got = L.evaluateAST(n, None)
    L.ASTNode_free(n)
    expect = eval(want, {"M_PI": math.pi, "log": math.log, "sqrt": math.sqrt, "cosh": math.cosh, "sinh": math.sinh})
    assert got == pytest.approx(expect, rel=1e-12, abs=1e-12)


@pytest.mark.parametrize("inp,want", KATS["differentiate"])
def test_differentiate_ast(inp, want):
    L = so.lib()
    n = _parse(inp)
    d = L.differentiateAST(n, b"x")
    assert _str(d) == want
    L.ASTNode_free(d)
    L.ASTNode_free(n)


@pytest.mark.parametrize("inp,want", KATS["simplify"])
def test_simplify_ast(inp, want):
    L = so.lib()
    n = _parse(inp)
    s = L.simplifyAST(n)
    assert _str(s) == want
    L.ASTNode_free(s)
    L.ASTNode_free(n)


def test_copy_and_index_ast():
    L = so.lib()
    for expr in ["0", "1.5", "foo", "(x + 1) * exp(y) / (pi * log(y - exponentiale))", "lambda(x, y, x + y)"]:
        n = _parse(expr)
        c = L.copyAST(n)
        assert _str(c) == expr
        L.ASTNode_free(c)
        L.ASTNode_free(n)
    names = (C.c_char_p * 5)(b"a", b"b", b"x", b"y", b"z")
    n = _parse("x + tan(y) / (2 * z)")
    idx = L.indexAST(n, 5, names)
    assert _str(idx) == "x + tan(y) / (2 * z)"
    L.ASTNode_free(idx)
    L.ASTNode_free(n)


def test_settings_defaults_and_time_grid():
    """unittest/test_integratorSettings.c:29-87: defaults BDF / NEWTON, grid i*Time/PrintStep (integratorSettings.c:416-456)."""
    L = so.lib()
    cs = L.CvodeSettings_create()
    assert L.CvodeSettings_getMethod(cs) == b"BDF"
    assert L.CvodeSettings_getIterMethod(cs) == b"NEWTON"
    assert L.CvodeSettings_getMaxOrder(cs) == 5
    assert L.CvodeSettings_getResetCvodeOnEvent(cs) == 1
    assert L.CvodeSettings_getHaltOnEvent(cs) == 0
    assert L.CvodeSettings_getError(cs) == 1e-18 and L.CvodeSettings_getRError(cs) == 1e-10
    assert L.CvodeSettings_getMxstep(cs) == 10000
    L.CvodeSettings_free(cs)
    cs = L.CvodeSettings_createWithTime(10.0, 100)
    assert L.CvodeSettings_getEndTime(cs) == 10.0 and L.CvodeSettings_getPrintsteps(cs) == 100
    assert [L.CvodeSettings_getTime(cs, i) for i in (0, 1, 50, 100)] == [0.0, 1 * 10.0 / 100, 50 * 10.0 / 100, 10.0]
    L.CvodeSettings_free(cs)
    cs = L.CvodeSettings_createWithTime(0.7, 29)
    assert L.CvodeSettings_getEndTime(cs) == 0.7 and L.CvodeSettings_getPrintsteps(cs) == 29
    L.CvodeSettings_free(cs)


def test_mapk_globalised_parameters_are_appended():
    """src/odeSolver.c:354-391: globalised kinetic-law locals get ids r_<reaction>_<param> after the existing constants."""
    m, vary, nominal = H.mapk_scan_model()
    assert m.neq == 8 and m.nass == 10 and m.nconst == 4 + 18
    assert m.names[18:22] == ["uVol", "V1", "Ki", "K1"]
    assert m.names[22:] == [f"r_{r}_{p}" for r, p in so.MAPK_LOCALS]
    assert len(vary) == 21 and (nominal > 0).all()


# ---- ODE assembly (src/odeConstruct.c:85-584) against unittest/test_odeConstruct.c ----
class _Rule(C.Structure):
    _fields_ = [("type", C.c_int), ("variable", C.c_char_p), ("math", C.c_void_p)]


class _Param(C.Structure):
    _fields_ = [("id", C.c_char_p), ("value", C.c_double), ("isSetValue", C.c_int), ("constant", C.c_int)]


def _model_api():
    L = so.lib()
    for fn in ("Model_getNumCompartments", "Model_getNumSpecies", "Model_getNumParameters", "Model_getNumReactions", "Model_getNumRules",
               "Model_getNumEvents", "Model_getNumInitialAssignments", "Model_getNumFunctionDefinitions"):
        getattr(L, fn).restype, getattr(L, fn).argtypes = C.c_uint, [C.c_void_p]
    for fn in ("Model_getRule", "Model_getParameter", "Model_getSpecies"):
        getattr(L, fn).restype, getattr(L, fn).argtypes = C.c_void_p, [C.c_void_p, C.c_uint]
    L.Model_reduceToOdes.restype, L.Model_reduceToOdes.argtypes = C.c_void_p, [C.c_void_p]
    L.Species_odeFromReactions.restype, L.Species_odeFromReactions.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p]
    L.Model_getValueById.restype, L.Model_getValueById.argtypes = C.c_double, [C.c_void_p, C.c_char_p]
    L.Model_setValue.restype, L.Model_setValue.argtypes = C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double]
    L.Model_free.restype, L.Model_free.argtypes = None, [C.c_void_p]
    return L


def _doc(name):
    L = so.lib()
    d = L.readSBML(H.model_path(name).encode())
    assert d
    return d, L.SBMLDocument_getModel(d)


@pytest.mark.parametrize("kat", KATS["reduce"], ids=lambda k: k["model"])
def test_model_reduce_to_odes(kat):
    """unittest/test_odeConstruct.c:93-412: the reduced model (reactions replaced by rate rules + one assignment rule
    per reaction): object counts, parameter order, and the formula of every rule, in order."""
    L = _model_api()
    d, m = _doc(kat["model"])
    om = L.Model_reduceToOdes(m)
    assert om and om != m
    have = {"getNumCompartments", "getNumSpecies", "getNumParameters", "getNumReactions", "getNumRules", "getNumEvents",
            "getNumInitialAssignments", "getNumFunctionDefinitions"}
    for fn, want in kat["counts"].items():
        if fn in have:
            assert getattr(L, "Model_" + fn)(om) == want, fn
        else:
            assert want == 0, fn            # unit definitions, compartment / species types, constraints: not modelled, reference expects none
    for i, name in kat["parameters"]:
        assert C.cast(L.Model_getParameter(om, i), C.POINTER(_Param)).contents.id.decode() == name
    for i, formula in kat["rules"]:
        r = C.cast(L.Model_getRule(om, i), C.POINTER(_Rule)).contents
        assert _str(r.math) == formula, (i, r.variable)
    L.Model_free(om)
    L.SBMLDocument_free(d)


@pytest.mark.parametrize("kat", KATS["species_ode"], ids=lambda k: k["model"])
def test_species_ode_from_reactions(kat):
    """unittest/test_odeConstruct.c:414-520: the ODE of every species, built from the stoichiometry, as a formula string."""
    L = _model_api()
    d, m = _doc(kat["model"])
    for i, formula in kat["odes"]:
        n = L.Species_odeFromReactions(L.Model_getSpecies(m, i), m)
        assert n, i
        assert _str(n) == formula, i
        L.ASTNode_free(n)
    L.SBMLDocument_free(d)


def test_model_get_value_and_set_value():
    """unittest/test_odeConstruct.c:34-91"""
    L = _model_api()
    for kat in KATS["get_value"]:
        d, m = _doc(kat["model"])
        for ident, want in kat["values"]:
            assert L.Model_getValueById(m, ident.encode()) == want, ident
        L.SBMLDocument_free(d)
    for kat in KATS["set_value"]:
        d, m = _doc(kat["model"])
        for ident, rid, value, ok in kat["calls"]:
            assert L.Model_setValue(m, ident.encode(), rid.encode() if rid else None, value) == ok, (ident, rid)
        L.SBMLDocument_free(d)


class _KineticLaw(C.Structure):
    _fields_ = [("math", C.c_void_p)]


class _Reaction(C.Structure):
    _fields_ = [("id", C.c_char_p), ("reversible", C.c_int), ("fast", C.c_int), ("reactants", C.c_byte * 16), ("products", C.c_byte * 16),
                ("modifiers", C.c_byte * 16), ("kineticLaw", C.POINTER(_KineticLaw))]


@pytest.mark.parametrize("kat", KATS["parse"], ids=lambda k: k["model"])
def test_parse_model(kat):
    """unittest/test_sbml.c:16-240: what the SBML reader delivers -- object counts, parameter ids, reaction ids and
    kinetic-law formulas (compared without white space: the Level 1 files keep their formula text in the reference,
    here every formula is rendered from the tree)."""
    L = _model_api()
    L.Model_getReaction.restype, L.Model_getReaction.argtypes = C.c_void_p, [C.c_void_p, C.c_uint]
    d, m = _doc(kat["model"])
    have = {"getNumCompartments", "getNumSpecies", "getNumParameters", "getNumReactions", "getNumRules", "getNumEvents",
            "getNumInitialAssignments", "getNumFunctionDefinitions"}
    for fn, want in kat["counts"].items():
        if fn in have:
            assert getattr(L, "Model_" + fn)(m) == want, fn
        else:
            assert want == 0, fn
    for i, name in kat["parameters"]:
        assert C.cast(L.Model_getParameter(m, i), C.POINTER(_Param)).contents.id.decode() == name
    for i, rid, formula in kat["reactions"]:
        r = C.cast(L.Model_getReaction(m, i), C.POINTER(_Reaction)).contents
        assert r.id.decode() == rid
        assert _str(r.kineticLaw.contents.math).replace(" ", "") == formula.replace(" ", ""), rid
    L.SBMLDocument_free(d)
