"""Test helpers: oracle binding (ctypes), kernel host-emulation builder, synthetic parameter sets.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

import sbml_odesolver_b200 as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = os.path.join(ROOT, "tests", "golden", "models")
ORACLE = os.path.join(ROOT, "oracle", "liboracle.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libcvode_ref.so")
PORT, REF = 0, 1

_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        so.lib()
        _oracle = C.CDLL(ORACLE)
        _oracle.oracle_integrate_batch.restype = C.c_int
        _oracle.oracle_integrate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_char_p, C.c_int]
        _oracle.oracle_last_error.restype = C.c_char_p
    return _oracle


def have_ref():
    return os.path.exists(REF_LIB)


def gpu_tier_backend():
    """Checker of the GPU tier: the reference's own vendored CVODE (oracle/_ref, travels to the GPU box) when it was built,
    else the restated port (bit-identical to it in the CPU tier, tests/test_oracle_golden.py)."""
    return REF if have_ref() else PORT


def model_path(name):
    return os.path.join(MODELS, name)


def oracle_run(model, opt, vary, values, nout, backend=PORT, threads=1, compiled_so=None, want_margin=False, adams=False, functional=False):
    """Returns dict(out[nsets][nout+1][nvalues], status, stats[nsets][7], fired[nsets][nout+1]).
    adams: Adams-Moulton instead of BDF -- only the vendored CVODE (backend=REF) has it."""
    if adams or functional:
        assert backend == REF, "the restated port holds BDF with Newton iteration only"
        C.c_int.in_dll(C.CDLL(REF_LIB), "refcv_adams").value = int(adams)
        C.c_int.in_dll(C.CDLL(REF_LIB), "refcv_functional").value = int(functional)
        try:
            return oracle_run(model, opt, vary, values, nout, backend=backend, threads=threads, compiled_so=compiled_so, want_margin=want_margin)
        finally:
            C.c_int.in_dll(C.CDLL(REF_LIB), "refcv_adams").value = 0
            C.c_int.in_dll(C.CDLL(REF_LIB), "refcv_functional").value = 0
    values = np.ascontiguousarray(values, dtype=np.float64)
    nsets = values.shape[0]
    vary = np.ascontiguousarray(vary, dtype=np.int32)
    out = np.zeros((nsets, nout + 1, model.nvalues))
    status = np.zeros(nsets, dtype=np.int32)
    stats = np.zeros((nsets, 7), dtype=np.int32)
    fired = np.zeros((nsets, nout + 1), dtype=np.uint32)
    margin = np.zeros((nsets, nout + 1)) if want_margin else None
    r = oracle().oracle_integrate_batch(model.om, opt, nsets, len(vary), vary.ctypes.data, values.ctypes.data, out.ctypes.data,
                                        status.ctypes.data, stats.ctypes.data, fired.ctypes.data,
                                        None if margin is None else margin.ctypes.data, backend,
                                        1 if compiled_so else 0, compiled_so.encode() if compiled_so else None, threads)
    if r < 0:
        raise RuntimeError("oracle failed: " + oracle().oracle_last_error().decode())
    return dict(out=out, status=status, stats=stats, fired=fired, margin=margin)


def compile_c_model(model, tag, opt=None):
    """gcc-compiles ODEModel_generateCSource (the reference's compiled mode) into a shared object; with a settings object
    that asks for sensitivities the source also holds sense_f for its parameters."""
    if opt is not None and so.lib().CvodeSettings_getSensitivity(opt):
        if not so.lib().ODEModel_getNumJacobiElements(model.om):
            so.lib().ODEModel_constructJacobian(model.om)
        os_ = so.lib().ODESense_create(model.om, opt)
        src = so.take_string(so.lib().ODEModel_generateCSourceSens(model.om, os_))
        so.lib().ODESense_free(os_)
    else:
        src = so.take_string(so.lib().ODEModel_generateCSource(model.om))
    h = hashlib.sha1(src.encode()).hexdigest()[:12]
    d = os.path.join(ROOT, "oracle", "_ref", "models")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"{tag}_{h}.so")
    if not os.path.exists(path):
        cpath = path[:-3] + ".c"
        open(cpath, "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", cpath, "-o", path, "-lm"])
    return path


REFS_LIB = os.path.join(ROOT, "oracle", "_ref", "libcvodes_ref.so")
_refs = None


def have_sens_ref():
    return os.path.exists(REFS_LIB)


def sens_ref_run(model, opt, batch, vary, values, tpoints, mxstep, rtol, atol, threads=1, tag="sens"):
    """Forward sensitivities from the reference's vendored CVODES 2.3.0 (oracle/sens_ref.c) driven with the emitted host-C
    model functions.  Returns dict(out[nsets][nout+1][neq], sens[nsets][nout+1][nsens][neq], status, stats[nsets][8])."""
    global _refs
    if _refs is None:
        oracle()                                          # liboracle.so (AST interpreter for the events) first
        _refs = C.CDLL(REFS_LIB)
        _refs.refcvs_integrate_batch.restype = C.c_int
    path = compile_c_model(model, tag, opt)
    values = np.ascontiguousarray(values, dtype=np.float64)
    nsets, nout, neq, ns = values.shape[0], len(tpoints) - 1, model.neq, batch.nsens
    vary = np.ascontiguousarray(vary, dtype=np.int32)
    base, _ = batch.base_values()
    base = np.ascontiguousarray(base, dtype=np.float64)
    sens_idx = [so.lib().ODESBatch_getSensIndex(batch.h, k) for k in range(ns)]
    sens_var = np.array([i if i < neq else -1 for i in sens_idx], dtype=np.int32)
    tp = np.ascontiguousarray(tpoints, dtype=np.float64)
    out = np.zeros((nsets, nout + 1, neq))
    sens = np.zeros((nsets, nout + 1, ns, neq))
    status = np.zeros(nsets, dtype=np.int32)
    stats = np.zeros((nsets, 8), dtype=np.int32)
    fired = np.zeros((nsets, nout + 1), dtype=np.uint32)
    _, trig0 = batch.base_values()
    trig0 = np.ascontiguousarray(trig0, dtype=np.int32)
    L = so.lib()
    _refs.refcvs_integrate_batch_events.restype = C.c_int

    class _Opt(C.Structure):                     # head of cvodeSettings_t up to SensMethod (include/sbmlsolver/integratorSettings.h)
        _fields_ = [("Time", C.c_double), ("PrintStep", C.c_int), ("StepResolution", C.c_int), ("TimePoints", C.c_void_p), ("Indefinitely", C.c_int),
                    ("Error", C.c_double), ("RError", C.c_double), ("Mxstep", C.c_int), ("DetectNegState", C.c_int),
                    ("CvodeMethod", C.c_int), ("IterMethod", C.c_int), ("MaxOrder", C.c_int), ("ResetCvodeOnEvent", C.c_int), ("SetTStop", C.c_int),
                    ("Sensitivity", C.c_int), ("sensIDs", C.c_void_p), ("nsens", C.c_int), ("SensMethod", C.c_int)]
    C.c_int.in_dll(_refs, "refcvs_sens_method").value = C.cast(opt, C.POINTER(_Opt)).contents.SensMethod
    dq = L.ODESBatch_getDifferenceQuotients(batch.h)          # which of the reference's difference-quotient fallbacks apply
    C.c_int.in_dll(_refs, "refcvs_sens_dq").value = dq & 1
    C.c_int.in_dll(_refs, "refcvs_jac_dq").value = (dq >> 1) & 1
    arr = (C.c_int * 64).in_dll(_refs, "refcvs_sens_index")
    for k in range(ns):
        arr[k] = sens_idx[k]
    r = _refs.refcvs_integrate_batch_events(path.encode(), C.c_int(neq), C.c_int(model.nvalues), C.c_int(ns), C.c_void_p(sens_var.ctypes.data),
                                            C.c_void_p(base.ctypes.data), C.c_int(nsets), C.c_int(len(vary)), C.c_void_p(vary.ctypes.data),
                                            C.c_void_p(values.ctypes.data), C.c_void_p(tp.ctypes.data), C.c_int(nout), C.c_double(rtol), C.c_double(atol),
                                            C.c_long(mxstep), C.c_void_p(out.ctypes.data), C.c_void_p(sens.ctypes.data), C.c_void_p(status.ctypes.data),
                                            C.c_void_p(stats.ctypes.data), C.c_int(threads), C.c_void_p(model.om),
                                            C.c_int(L.CvodeSettings_getResetCvodeOnEvent(opt)), C.c_int(L.CvodeSettings_getHaltOnEvent(opt)),
                                            C.c_void_p(trig0.ctypes.data), C.c_void_p(fired.ctypes.data))
    if r < 0:
        raise RuntimeError("sensitivity reference failed to set up")
    return dict(out=out, sens=sens, status=status, stats=stats, fired=fired)


class Emu:
    """Host build of the generated kernel translation unit (tests/emu/emu_harness.cpp)."""

    def __init__(self, batch, tag):
        src = batch.source()
        # extra defines for the host build (e.g. ODES_EMU_DEFS="-DODES_LU_FORM=2": the left-looking linear algebra the
        # GPU build selects with tensor memory, run here on the shared-memory word space)
        extra = (tag if isinstance(tag, (list, tuple)) else []) or os.environ.get("ODES_EMU_DEFS", "").split()
        tag = tag if isinstance(tag, str) else "emu"
        h = hashlib.sha1((src + " ".join(extra)).encode()).hexdigest()[:12]
        d = os.path.join(ROOT, "tests", "emu", "_build")
        os.makedirs(d, exist_ok=True)
        cu = os.path.join(d, f"{tag}_{h}.cu")
        lib = os.path.join(d, f"{tag}_{h}.so")
        if not os.path.exists(lib):
            open(cu, "w").write(src)
            subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-w"] + extra + (["-DODES_ROOT_POW_LIBM"] if os.environ.get("ODES_EMU_LIBM_POW") else []) +
                                  [t for t in src.split("\n")[0].split() if t.startswith("-DODES_STEPRES") or t.startswith("-DODES_ROOTFIND") or t.startswith("-DODES_LA_PACKED") or t.startswith("-DODES_ALIAS_FY")] +
                                  ([t for t in src.split("\n")[0].split() if t.startswith("-DODES_W") or t.startswith("-DODES_ADAMS") or t.startswith("-DODES_FUNCTIONAL")] if "-DODES_WARP=1" in src.split("\n")[0] else []) + [
                                   f'-DODES_MODEL_SOURCE="{cu}"', os.path.join(ROOT, "tests", "emu", "emu_harness.cpp"), "-o", lib])
        self.lib = C.CDLL(lib)
        self.batch = batch

    def run(self, base, trig0, values, tpoints, mxstep, rtol, atol, stepres=1):
        """stepres: the ODES_STEPRES the unit was built with (tpoints is then the refined grid; rows are kept every stepres)"""
        b = self.batch
        values = np.ascontiguousarray(values, dtype=np.float64)
        nsets = values.shape[0]
        nout = len(tpoints) - 1
        vary = np.ascontiguousarray(values.T)                      # SoA [nvary][nsets]
        raw = np.zeros((nsets, nout // stepres + 1, b.nstore))     # the library's [set][row][value] layout
        out = np.transpose(raw, (1, 2, 0))                         # viewed as [row][value][set]
        status = np.zeros(nsets, dtype=np.int32)
        stats = np.zeros((7, nsets), dtype=np.int32)
        fired = np.zeros((nout // stepres + 1, nsets), dtype=np.uint32)
        base = np.ascontiguousarray(base, dtype=np.float64)
        trig0 = np.ascontiguousarray(trig0, dtype=np.int32)
        tp = np.ascontiguousarray(tpoints, dtype=np.float64)
        sens = np.zeros((nsets, nout // stepres + 1, max(b.nsens, 1), b.model.neq))
        self.lib.emu_run(C.c_void_p(base.ctypes.data), C.c_void_p(trig0.ctypes.data), C.c_void_p(vary.ctypes.data), C.c_void_p(raw.ctypes.data),
                         C.c_void_p(status.ctypes.data), C.c_void_p(stats.ctypes.data), C.c_void_p(fired.ctypes.data), C.c_void_p(tp.ctypes.data),
                         C.c_ulonglong(nsets), C.c_int(nsets), C.c_int(nout), C.c_int(mxstep), C.c_double(rtol), C.c_double(atol),
                         C.c_void_p(sens.ctypes.data))
        return dict(out=out, status=status, stats=stats, fired=fired, sens=sens)


def mapk_scan_model():
    """MAPK with the 18 kinetic-law locals globalised; returns (model, vary indices, nominal values)."""
    m = so.Model(model_path("mapk_cascade.xml"), globalize=so.MAPK_LOCALS)
    names = ["V1", "Ki", "K1"] + m.globalized
    vary = [m.index(n) for n in names]
    nominal = m.base_values()[vary]
    return m, vary, nominal


def log_uniform_sets(nominal, nsets, seed, span=1.0):
    """nominal * 2^u, u ~ U(-span, span)  (numpy generator for the CPU-side tests)."""
    rng = np.random.default_rng(seed)
    u = rng.uniform(-span, span, size=(nsets, len(nominal)))
    return nominal[None, :] * np.exp2(u)


def parity(gpu, ref, rtol, atol):
    """max over points of |gpu-ref| / max(10*rtol*|ref|, 10*atol); NaN rows must coincide."""
    nan_g, nan_r = np.isnan(gpu), np.isnan(ref)
    if not np.array_equal(nan_g, nan_r):
        return np.inf
    tol = np.maximum(10 * rtol * np.abs(ref), 10 * atol)
    with np.errstate(invalid="ignore"):
        ratio = np.where(nan_r, 0.0, np.abs(gpu - ref) / tol)
    return float(ratio.max()) if ratio.size else 0.0


def parity_report(gpu, ref, gstats, rstats, rtol, atol):
    """Per-set comparison of a device run with the oracle.

    gpu/ref: [nrows][nstore][nsets]; gstats/rstats: [nsets][7] CVODE work counters.
    Returns the per-set violation ratio |gpu-ref| / max(10 rtol |ref|, 10 atol), which sets took exactly the
    same steps and decisions as the reference CVODE run (identical counters), and summary numbers.
    """
    nan_g, nan_r = np.isnan(gpu), np.isnan(ref)
    tol = np.maximum(10 * rtol * np.abs(ref), 10 * atol)
    with np.errstate(invalid="ignore"):
        r = np.where(nan_r & nan_g, 0.0, np.where(nan_r | nan_g, np.inf, np.abs(gpu - ref) / tol))
    ratio = r.max(axis=(0, 1))
    same = (np.asarray(gstats) == np.asarray(rstats)).all(axis=1)
    return dict(ratio=ratio, same_path=same, frac_same_path=float(same.mean()),
                max_ratio_same_path=float(ratio[same].max()) if same.any() else 0.0,
                max_ratio=float(ratio.max()), frac_within=float((ratio <= 1.0).mean()))


# ---------------------------------------------------------------------------------------------
# The benchmark configurations of BASELINE.json / SURVEY section 8(d) as (model, settings, varied
# indices, value generator).  Shared by the emulation tests, the GPU parity tests and bench.py.
# huang96: the 30 mass-action constants a/d/k are kinetic-law locals, (reaction, parameter) in model order
HUANG_LOCALS = [(f"r{i}a", f"{c}{i}") for i in range(1, 11) for c in ("a", "d")]
HUANG_LOCALS = sorted(HUANG_LOCALS + [("r1b", "k2")] + [(f"r{i}b", f"k{i}") for i in range(2, 11)],
                      key=lambda rp: (int(rp[0][1:-1]), rp[0][-1], rp[1][0] != "a"))


def workload(name, nout=None, **settings_kw):
    """Returns dict(model, opt, vary, nominal, span, log10, rtol, atol, nout, end, store) for a config name."""
    rtol, atol = 1e-6, 1e-9
    if name == "mapk":
        m, vary, nominal = mapk_scan_model()
        end, n_out, span, log10, kw = 1000.0, 100, (-1.0, 1.0), False, {}
    elif name == "repressilator":
        m = so.Model(model_path("repressilator_ring.xml"))
        vary = [m.index(n) for n in ("x1", "x2", "x3", "y1", "y2", "y3")]
        nominal = np.ones(6)                                  # x_i, y_i = 10^u, u ~ U(-2, 0.5)
        end, n_out, span, log10, kw = 1000.0, 1000, (-2.0, 0.5), True, {}
    elif name == "events":
        m = so.Model(model_path("two_events_one_assignment.xml"))
        vary = [m.index("S1"), m.index("S2")]
        nominal = None                                        # S1 ~ U(0.5, 2), S2 ~ U(0, 0.4): linear ranges
        end, n_out, span, log10, kw = 10.0, 100, None, False, {}
    elif name == "huang":
        m = so.Model(model_path("huang_ferrell_cascade.xml"), globalize=HUANG_LOCALS)
        vary = [m.index(n) for n in m.globalized]
        nominal = m.base_values()[vary]
        end, n_out, span, log10, kw = 100.0, 100, (-1.0, 1.0), False, {}
    elif name == "goodwin":
        m = so.Model(model_path("goodwin_oscillator.xml"))
        vary = [m.index(f"k{i}") for i in range(1, 11)]
        nominal = m.base_values()[vary]
        end, n_out, span, log10, kw = 100.0, 100, (-1.0, 1.0), False, {}
    else:
        raise KeyError(name)
    if nout is not None:
        end, n_out = end * nout / n_out, nout
    kw.update(settings_kw)
    opt = so.settings(end, n_out, atol, rtol, **kw)
    return dict(name=name, model=m, opt=opt, vary=vary, nominal=nominal, span=span, log10=log10,
                rtol=rtol, atol=atol, nout=n_out, end=end)


def workload_values(w, nsets, seed):
    """Host-side synthetic sets for a workload (numpy generator; the device-side Philox generator is used by bench.py)."""
    rng = np.random.default_rng(seed)
    if w["name"] == "events":
        return np.stack([rng.uniform(0.5, 2.0, nsets), rng.uniform(0.0, 0.4, nsets)], axis=1)
    lo, hi = w["span"]
    u = rng.uniform(lo, hi, size=(nsets, len(w["vary"])))
    return w["nominal"][None, :] * (np.power(10.0, u) if w["log10"] else np.exp2(u))


def tpoints(opt, nout):
    return np.array([so.lib().CvodeSettings_getTime(opt, i) for i in range(nout + 1)])
