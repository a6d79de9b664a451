"""Front-end breadth: synthetic SBML Level 2 models with function definitions, initial
assignments, rate and assignment rules, piecewise terms, stoichiometryMath, relational/logical operators and
events go through the whole host chain -- reader, ODE construction, symbolic Jacobian, emitter -- and the emulated
kernel must reproduce the oracle (same emitted right-hand side: bit for bit; interpreter: within tolerance)."""
import math

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so

M = 'xmlns="http://www.w3.org/1998/Math/MathML"'

FEATURES = f'''<?xml version="1.0" encoding="UTF-8"?>
<sbml xmlns="http://www.sbml.org/sbml/level2" level="2" version="1">
 <model id="features">
  <listOfFunctionDefinitions>
   <functionDefinition id="mm">
    <math {M}><lambda><bvar><ci>v</ci></bvar><bvar><ci>s</ci></bvar><bvar><ci>k</ci></bvar>
     <apply><divide/><apply><times/><ci>v</ci><ci>s</ci></apply><apply><plus/><ci>k</ci><ci>s</ci></apply></apply></lambda></math>
   </functionDefinition>
  </listOfFunctionDefinitions>
  <listOfCompartments><compartment id="c" size="2"/></listOfCompartments>
  <listOfSpecies>
   <species id="A" compartment="c" initialConcentration="4"/>
   <species id="B" compartment="c" initialConcentration="0.5"/>
   <species id="C" compartment="c" initialConcentration="0"/>
   <species id="E" compartment="c" initialConcentration="1" boundaryCondition="true"/>
  </listOfSpecies>
  <listOfParameters>
   <parameter id="vmax" value="1.5"/><parameter id="km" value="0.8"/><parameter id="kd" value="0.3"/>
   <parameter id="gain" value="0" constant="false"/><parameter id="x" value="0.1" constant="false"/>
   <parameter id="thr" value="2.5"/><parameter id="x0" value="0"/>
  </listOfParameters>
  <listOfInitialAssignments>
   <initialAssignment symbol="x0"><math {M}><apply><times/><cn>2</cn><ci>km</ci></apply></math></initialAssignment>
  </listOfInitialAssignments>
  <listOfRules>
   <assignmentRule variable="gain"><math {M}><piecewise><piece><cn>2</cn><apply><gt/><ci>A</ci><ci>thr</ci></apply></piece><otherwise><cn>0.5</cn></otherwise></piecewise></math></assignmentRule>
   <rateRule variable="x"><math {M}><apply><minus/><apply><times/><ci>gain</ci><ci>B</ci></apply><apply><times/><ci>kd</ci><ci>x</ci></apply></apply></math></rateRule>
  </listOfRules>
  <listOfReactions>
   <reaction id="r1" reversible="false">
    <listOfReactants><speciesReference species="A"/></listOfReactants>
    <listOfProducts><speciesReference species="B"><stoichiometryMath><math {M}><apply><plus/><cn>1</cn><cn>1</cn></apply></math></stoichiometryMath></speciesReference></listOfProducts>
    <listOfModifiers><modifierSpeciesReference species="E"/></listOfModifiers>
    <kineticLaw><math {M}><apply><times/><ci>c</ci><ci>E</ci><apply><ci>mm</ci><ci>vmax</ci><ci>A</ci><ci>km</ci></apply></apply></math></kineticLaw>
   </reaction>
   <reaction id="r2" reversible="false">
    <listOfReactants><speciesReference species="B"/></listOfReactants>
    <listOfProducts><speciesReference species="C"/></listOfProducts>
    <kineticLaw><math {M}><apply><times/><ci>c</ci><ci>k2</ci><ci>B</ci><apply><power/><apply><plus/><cn>1</cn><ci>x</ci></apply><cn>1.5</cn></apply></apply></math>
     <listOfParameters><parameter id="k2" value="0.4"/></listOfParameters></kineticLaw>
   </reaction>
  </listOfReactions>
  <listOfEvents>
   <event id="refill"><trigger><math {M}><apply><and/><apply><lt/><ci>A</ci><cn>0.5</cn></apply><apply><geq/><ci>C</ci><cn>1</cn></apply></apply></math></trigger>
    <listOfEventAssignments><eventAssignment variable="A"><math {M}><apply><plus/><ci>A</ci><cn>3</cn></apply></math></eventAssignment></listOfEventAssignments>
   </event>
  </listOfEvents>
 </model>
</sbml>
'''


@pytest.fixture(scope="module")
def features_model(tmp_path_factory):
    p = tmp_path_factory.mktemp("m") / "features.xml"
    p.write_text(FEATURES)
    return so.Model(str(p))


def test_structure_of_the_feature_model(features_model):
    m = features_model
    assert m.neq == 4 and m.names[:4] == ["x", "A", "B", "C"]            # predefined rate rule first, then species
    assert m.nass == 3                                                   # gain + one per reaction
    eq = {m.names[i]: m.equation(i) for i in range(m.neq + m.nass)}
    assert eq["x"] == "gain * B - kd * x"
    assert eq["A"] == "-(r1 / c)" and eq["C"] == "r2 / c"          # unary minus like unittest/test_odeModel.c:247
    assert eq["B"] == "((1 + 1) * r1 - r2) / c"                            # stoichiometryMath multiplies the rate
    assert eq["r1"] == "c * E * (vmax * A / (km + A))"                   # function definition inlined
    assert eq["r2"] == "c * 0.4 * B * (1 + x)^1.5"                        # local parameter replaced by its value
    assert "piecewise" in eq["gain"]
    assert m.base_values()[m.index("x0")] == 0.0                         # initial assignment applies at reset, not in the file values


@pytest.mark.parametrize("halt", [0, 1])
def test_feature_model_emulated_kernel_vs_oracle(features_model, halt):
    m = features_model
    vary = [m.index(n) for n in ("vmax", "km", "kd", "A")]
    rng = np.random.default_rng(3)
    vals = np.stack([rng.uniform(0.5, 3, 24), rng.uniform(0.2, 2, 24), rng.uniform(0.05, 1, 24), rng.uniform(1, 6, 24)], axis=1)
    opt = so.settings(40.0, 80, 1e-9, 1e-6, halt_on_event=halt)
    b = so.Batch(m, opt, vary)
    base, trig = b.base_values()
    assert base[m.index("x0")] == pytest.approx(1.6)                     # 2*km with the model's own km
    got = H.Emu(b, f"features{halt}").run(base, trig, vals, H.tpoints(opt, 80), 10000, 1e-6, 1e-9)
    ref = H.oracle_run(m, opt, vary, vals, 80, compiled_so=H.compile_c_model(m, "features"))
    assert np.array_equal(got["fired"].T, ref["fired"]) and ref["fired"].any()
    assert np.array_equal(got["status"], ref["status"]) and np.array_equal(got["stats"].T, ref["stats"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    refi = H.oracle_run(m, opt, vary, vals, 80, want_margin=True)
    safe = (refi["margin"] > 1e-4).all(axis=1)
    assert safe.mean() > 0.8
    assert np.array_equal(got["fired"].T[safe], refi["fired"][safe])
    assert H.parity(got["out"][:, :, safe], np.transpose(refi["out"], (1, 2, 0))[:, :, safe], 1e-6, 1e-9) <= 1.0
    b.close()


def chain_model_xml(n):
    """A cascade S1 -> S2 -> ... -> Sn with saturable steps and feedback of the last species on the first step."""
    sp = "".join(f'<species id="S{i}" compartment="c" initialConcentration="{1.0 if i == 1 else 0.1}"/>' for i in range(1, n + 1))
    rx = []
    for i in range(1, n):
        law = (f"<apply><divide/><apply><times/><ci>v{i}</ci><ci>S{i}</ci></apply><apply><plus/><ci>K</ci><ci>S{i}</ci></apply></apply>")
        if i == 1:
            law = f"<apply><divide/>{law}<apply><plus/><cn>1</cn><apply><power/><ci>S{n}</ci><cn>2</cn></apply></apply></apply>"
        rx.append(f'<reaction id="r{i}" reversible="false"><listOfReactants><speciesReference species="S{i}"/></listOfReactants>'
                  f'<listOfProducts><speciesReference species="S{i + 1}"/></listOfProducts><kineticLaw><math {M}>{law}</math></kineticLaw></reaction>')
    rx.append(f'<reaction id="r{n}" reversible="false"><listOfReactants><speciesReference species="S{n}"/></listOfReactants>'
              f'<listOfProducts><speciesReference species="S1"/></listOfProducts><kineticLaw><math {M}><apply><times/><ci>kback</ci><ci>S{n}</ci></apply></math></kineticLaw></reaction>')
    pars = "".join(f'<parameter id="v{i}" value="{0.5 + 0.1 * i}"/>' for i in range(1, n)) + '<parameter id="K" value="0.3"/><parameter id="kback" value="0.2"/>'
    return (f'<?xml version="1.0" encoding="UTF-8"?><sbml xmlns="http://www.sbml.org/sbml/level2" level="2" version="1"><model id="chain{n}">'
            f'<listOfCompartments><compartment id="c" size="1"/></listOfCompartments><listOfSpecies>{sp}</listOfSpecies>'
            f'<listOfParameters>{pars}</listOfParameters><listOfReactions>{"".join(rx)}</listOfReactions></model></sbml>')


def amplifier_model_xml(n):
    """A cascade S1 -> 3 S2 -> 9 S3 ... with product stoichiometry 3 and first-order decay of every species: below the
    diagonal the Newton matrix I - gamma J holds -3 gamma k against 1 + gamma (k + d) on it, so as soon as the step size
    is large the partial pivoting of gefa exchanges rows (smalldense.c:74-96) -- the case the reaction networks of the
    benchmark never meet."""
    sp = "".join(f'<species id="S{i}" compartment="c" initialConcentration="{2.0 if i == 1 else 0.05}"/>' for i in range(1, n + 1))
    rx = []
    for i in range(1, n):
        rx.append(f'<reaction id="r{i}" reversible="false"><listOfReactants><speciesReference species="S{i}"/></listOfReactants>'
                  f'<listOfProducts><speciesReference species="S{i + 1}" stoichiometry="3"/></listOfProducts><kineticLaw><math {M}>'
                  f'<apply><times/><ci>k{i}</ci><ci>S{i}</ci></apply></math></kineticLaw></reaction>')
    for i in range(1, n + 1):
        rx.append(f'<reaction id="d{i}" reversible="false"><listOfReactants><speciesReference species="S{i}"/></listOfReactants>'
                  f'<kineticLaw><math {M}><apply><times/><ci>dec</ci><ci>S{i}</ci></apply></math></kineticLaw></reaction>')
    rx.append(f'<reaction id="src" reversible="false"><listOfProducts><speciesReference species="S1"/></listOfProducts>'
              f'<kineticLaw><math {M}><apply><divide/><ci>vin</ci><apply><plus/><cn>1</cn><ci>S{n}</ci></apply></apply></math></kineticLaw></reaction>')
    pars = "".join(f'<parameter id="k{i}" value="{4.0 + i}"/>' for i in range(1, n)) + '<parameter id="dec" value="0.8"/><parameter id="vin" value="1.5"/>'
    return (f'<?xml version="1.0" encoding="UTF-8"?><sbml xmlns="http://www.sbml.org/sbml/level2" level="2" version="1"><model id="amp{n}">'
            f'<listOfCompartments><compartment id="c" size="1"/></listOfCompartments><listOfSpecies>{sp}</listOfSpecies>'
            f'<listOfParameters>{pars}</listOfParameters><listOfReactions>{"".join(rx)}</listOfReactions></model></sbml>')


def amplifier_case(tmp_path, n, nsets):
    p = tmp_path / f"amp{n}.xml"
    p.write_text(amplifier_model_xml(n))
    m = so.Model(str(p))
    vary = [m.index(f"k{i}") for i in range(1, n)] + [m.index("dec")]
    rng = np.random.default_rng(100 + n)
    vals = m.base_values()[vary][None, :] * np.exp2(rng.uniform(-1, 1, size=(nsets, len(vary))))
    return m, vary, vals, so.settings(20.0, 20, 1e-9, 1e-6)


@pytest.mark.parametrize("n", [4, 8])
def test_left_looking_lu_with_row_exchanges_emulated(tmp_path, n):
    """LU form 2 of bdf_kernel.cuh (left-looking, LAPACK row layout, what the GPU build uses with tensor memory) on a model
    whose factorisations exchange rows: bit-identical to the vendored CVODE; the exchanges are counted."""
    import ctypes as C
    m, vary, vals, opt = amplifier_case(tmp_path, n, 12)
    b = so.Batch(m, opt, vary)
    base, trig = b.base_values()
    emu = H.Emu(b, ["-DODES_LU_FORM=2", "-DODES_EMU_COUNT_SWAPS"])
    got = emu.run(base, trig, vals, H.tpoints(opt, 20), 10000, 1e-6, 1e-9)
    ref = H.oracle_run(m, opt, vary, vals, 20, backend=H.gpu_tier_backend(), compiled_so=H.compile_c_model(m, f"amp{n}"))
    assert C.c_long.in_dll(emu.lib, "odes_emu_swaps").value > 100
    assert np.array_equal(got["stats"].T, ref["stats"]) and (got["status"] == 0).all()
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)))
    b.close()


@pytest.mark.parametrize("n", [3, 12, 17, 20, 32])
def test_chain_models_emulated_kernel_vs_oracle(tmp_path, n):
    """State dimensions around the storage thresholds: one and two 8-row chunks per matrix column, packed and
    array pivots (n > 16); n = 20 and n = 32 (every lane owns a row) take the warp-per-trajectory layout.  n = 3 runs enough sets to meet arguments where libm's pow(x, 2) is not x * x rounded
    (about 3e-4 of the calls): stored assignment-rule values follow the interpreter (pow), the ODEs the compiled x * x."""
    p = tmp_path / f"chain{n}.xml"
    p.write_text(chain_model_xml(n))
    m = so.Model(str(p))
    assert m.neq == n and m.nass == n
    vary = [m.index(f"v{i}") for i in range(1, n)] + [m.index("K")]
    rng = np.random.default_rng(n)
    vals = m.base_values()[vary][None, :] * np.exp2(rng.uniform(-1, 1, size=(400 if n == 3 else 6, len(vary))))
    opt = so.settings(30.0, 30, 1e-9, 1e-6)
    b = so.Batch(m, opt, vary)
    base, trig = b.base_values()
    got = H.Emu(b, f"chain{n}").run(base, trig, vals, H.tpoints(opt, 30), 10000, 1e-6, 1e-9)
    ref = H.oracle_run(m, opt, vary, vals, 30, compiled_so=H.compile_c_model(m, f"chain{n}"))
    assert np.array_equal(got["stats"].T, ref["stats"]) and (got["status"] == 0).all()
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)))
    if n == 3:
        import math
        y3 = got["out"][1:, 2, :].ravel()
        assert (np.array([math.pow(v, 2.0) for v in y3]) != y3 * y3).any()
    b.close()


TIME_MODEL = f'''<?xml version="1.0" encoding="UTF-8"?>
<sbml xmlns="http://www.sbml.org/sbml/level2" level="2" version="1"><model id="forced">
 <listOfCompartments><compartment id="c" size="1"/></listOfCompartments>
 <listOfParameters><parameter id="x" value="0.5" constant="false"/><parameter id="w" value="0.7"/><parameter id="k" value="0.9"/>
  <parameter id="drive" value="0" constant="false"/><parameter id="flag" value="0" constant="false"/></listOfParameters>
 <listOfRules>
  <assignmentRule variable="drive"><math {M}><apply><exp/><apply><times/><apply><minus/><ci>w</ci></apply><csymbol encoding="text" definitionURL="http://www.sbml.org/sbml/symbols/time">time</csymbol></apply></apply></math></assignmentRule>
  <rateRule variable="x"><math {M}><apply><minus/><apply><plus/><cn>2</cn><ci>drive</ci></apply><apply><times/><ci>k</ci><ci>x</ci><apply><ln/><apply><plus/><cn>2</cn><ci>x</ci></apply></apply></apply></apply></math></rateRule>
 </listOfRules>
 <listOfEvents><event id="late"><trigger><math {M}><apply><gt/><csymbol encoding="text" definitionURL="http://www.sbml.org/sbml/symbols/time">time</csymbol><cn>3.3</cn></apply></math></trigger>
  <listOfEventAssignments><eventAssignment variable="flag"><math {M}><cn>1</cn></math></eventAssignment></listOfEventAssignments></event></listOfEvents>
</model></sbml>
'''


def test_time_dependent_model_sees_time_as_float(tmp_path):
    """The reference keeps the current time in a float (src/sbmlsolver/cvodeData.h:116): formulas that mention time see
    t rounded to 24 bits, in the right-hand side, in the rules of the stored rows and in event triggers (SURVEY
    appendix A.3).  The emitted code does the same, so the kernel is bit-identical to the oracle driven with the emitted
    C flavour and within tolerance of the interpreted oracle, with the event at the same output row."""
    p = tmp_path / "forced.xml"
    p.write_text(TIME_MODEL)
    m = so.Model(str(p))
    assert m.neq == 1 and m.names[0] == "x" and m.equation(0) == "2 + drive - k * x * log(2 + x)"
    vary = [m.index("w"), m.index("k"), m.index("x")]
    rng = np.random.default_rng(8)
    vals = np.stack([rng.uniform(0.3, 2.0, 10), rng.uniform(0.2, 3.0, 10), rng.uniform(0.5, 3, 10)], axis=1)
    opt = so.settings(7.0, 70, 1e-10, 1e-8)
    b = so.Batch(m, opt, vary)
    assert "(float)t" in b.source()
    base, trig = b.base_values()
    got = H.Emu(b, "forced").run(base, trig, vals, H.tpoints(opt, 70), 10000, 1e-8, 1e-10)
    ref = H.oracle_run(m, opt, vary, vals, 70, compiled_so=H.compile_c_model(m, "forced"))
    assert np.array_equal(got["stats"].T, ref["stats"]) and (got["status"] == 0).all()
    assert np.array_equal(got["fired"].T, ref["fired"])
    assert [int(np.nonzero(f)[0][0]) for f in ref["fired"]] == [34] * 10          # time > 3.3 first holds at t = 3.4
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)))
    refi = H.oracle_run(m, opt, vary, vals, 70)
    assert np.array_equal(got["fired"].T, refi["fired"])
    assert H.parity(got["out"], np.transpose(refi["out"], (1, 2, 0)), 1e-8, 1e-10) <= 1.0
    # the stored rule value is exp(-w * float(t)), not exp(-w * t)
    k = m.index("drive")
    t = np.float32(0.1 * 33)
    assert np.array_equal(got["out"][33, k, :], np.array([math.exp(-w * float(t)) for w in vals[:, 0]]))
    b.close()
