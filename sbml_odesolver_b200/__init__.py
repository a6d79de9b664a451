"""sbml_odesolver_b200 -- Python loader for the C-ABI library libodes_b200.so.

The product is the shared library (ISO-C host side + CUDA kernels for sm_100a);
this module only binds it with ctypes so that tests, bench.py and
__graft_entry__ can call the same entry points a C caller of the reference's
API would use (names, argument meaning and error behaviour are the
reference's: ODEModel_*, CvodeSettings_*, IntegratorInstance_*, VarySettings_*,
plus the new ODESBatch_* / IntegratorInstance_integrateBatch batched entry).

There is no Python or CPU implementation of the integrator behind it: if the
library is missing, importing raises; if no CUDA device is present, every
launch entry fails with SOLVER_ERROR_CUDA_UNAVAILABLE (150000).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libodes_b200.so")

c_p, c_i, c_d, c_s = C.c_void_p, C.c_int, C.c_double, C.c_char_p

# every symbol the headers in include/sbmlsolver declare and Python uses: (restype, argtypes)
_SIGS = {
    "readSBML": (c_p, [c_s]), "readSBMLFromString": (c_p, [c_s]), "writeSBMLToString": (c_p, [c_p]),
    "writeSBML": (c_i, [c_p, c_s]), "SBMLDocument_free": (None, [c_p]), "SBMLDocument_getModel": (c_p, [c_p]),
    "SBML_parseFormula": (c_p, [c_s]), "SBML_formulaToString": (c_p, [c_p]), "ASTNode_free": (None, [c_p]),
    "readMathMLFromString": (c_p, [c_s]), "writeMathMLToString": (c_p, [c_p]),
    "differentiateAST": (c_p, [c_p, c_s]), "simplifyAST": (c_p, [c_p]), "copyAST": (c_p, [c_p]),
    "indexAST": (c_p, [c_p, c_i, c_p]), "evaluateAST": (c_d, [c_p, c_p]),
    "Model_globalizeParameter": (c_i, [c_p, c_s, c_s]), "Model_localizeParameter": (c_i, [c_p, c_s, c_s]),
    "ODEModel_create": (c_p, [c_p]), "ODEModel_createFromFile": (c_p, [c_s]), "ODEModel_createFromSBML2": (c_p, [c_p]),
    "ODEModel_free": (None, [c_p]),
    "ODEModel_getNeq": (c_i, [c_p]), "ODEModel_getNumAssignments": (c_i, [c_p]), "ODEModel_getNumConstants": (c_i, [c_p]),
    "ODEModel_getNumValues": (c_i, [c_p]), "ODEModel_getNalg": (c_i, [c_p]), "ODEModel_hasCycle": (c_i, [c_p]),
    "ODEModel_hasVariable": (c_i, [c_p, c_s]),
    "ODEModel_getNumAssignmentsBeforeODEs": (c_i, [c_p]), "ODEModel_getNumAssignmentsBeforeEvents": (c_i, [c_p]),
    "ODEModel_getNumJacobiElements": (c_i, [c_p]), "ODEModel_getVariableIndexFields": (c_i, [c_p, c_s]),
    "ODEModel_getVariableIndexByNum": (c_p, [c_p, c_i]), "ODEModel_getVariableIndex": (c_p, [c_p, c_s]),
    "ODEModel_getVariableName": (c_s, [c_p, c_p]), "VariableIndex_free": (None, [c_p]),
    "ODEModel_getOde": (c_p, [c_p, c_p]), "ODEModel_getAssignment": (c_p, [c_p, c_p]),
    "ODEModel_constructJacobian": (c_i, [c_p]), "ODEModel_getJacobiElement": (c_p, [c_p, c_i]),
    "ODEModel_getJacobianIJEntry": (c_p, [c_p, c_i, c_i]),
    "ODEModel_generateCSource": (c_p, [c_p]), "ODEModel_generateCSourceSens": (c_p, [c_p, c_p]),
    "ODESense_create": (c_p, [c_p, c_p]), "ODESense_free": (None, [c_p]), "ODESense_getNsens": (c_i, [c_p]), "ODESense_getNeq": (c_i, [c_p]),
    "ODESense_getSensIJEntry": (c_p, [c_p, c_i, c_i]), "ODESense_getSensParamIndexByNum": (c_p, [c_p, c_i]),
    "ODEModel_constructSensitivity": (c_p, [c_p]),
    "CvodeSettings_setSensParams": (c_i, [c_p, c_p, c_i]), "CvodeSettings_unsetSensParams": (None, [c_p]),
    "CvodeSettings_setSensitivity": (None, [c_p, c_i]), "CvodeSettings_setSensMethod": (None, [c_p, c_i]),
    "CvodeSettings_getSensitivity": (c_i, [c_p]),
    "ODESBatch_getNsens": (c_i, [c_p]), "ODESBatch_getSensIndex": (c_i, [c_p, c_i]), "ODESBatch_getDifferenceQuotients": (c_i, [c_p]), "ODESBatch_getSensitivities": (c_i, [c_p, c_i, c_p]),
    "ODESBatch_deviceSensitivities": (c_p, [c_p]),
    "ODESBatch_runHostSens": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_i]),
    "ODESBatch_runHostAllDevicesSens": (c_i, [c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_i]),
    "IntegratorInstance_integrateBatchSens": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "IntegratorInstance_getNsens": (c_i, [c_p]), "IntegratorInstance_getSensitivityByNum": (c_d, [c_p, c_i, c_i]),
    "SBMLResults_getNumSens": (c_i, [c_p]), "SBMLResults_getSensParam": (c_s, [c_p, c_i]), "TimeCourse_getSensitivity": (c_d, [c_p, c_i, c_i]),
    "ODEModel_generateCUDASource": (c_p, [c_p, c_p, c_i, c_p, c_i, c_p, c_p]),
    "CvodeSettings_create": (c_p, []), "CvodeSettings_createWithTime": (c_p, [c_d, c_i]),
    "CvodeSettings_createWith": (c_p, [c_d, c_i, c_d, c_d] + [c_i] * 10), "CvodeSettings_free": (None, [c_p]),
    "CvodeSettings_setTime": (c_i, [c_p, c_d, c_i]), "CvodeSettings_setTimeSeries": (c_i, [c_p, c_p, c_i]),
    "CvodeSettings_getTime": (c_d, [c_p, c_i]), "CvodeSettings_getError": (c_d, [c_p]), "CvodeSettings_getRError": (c_d, [c_p]),
    "CvodeSettings_getMxstep": (c_i, [c_p]), "CvodeSettings_getMethod": (c_s, [c_p]), "CvodeSettings_getIterMethod": (c_s, [c_p]),
    "CvodeSettings_getMaxOrder": (c_i, [c_p]), "CvodeSettings_getJacobian": (c_i, [c_p]), "CvodeSettings_getStoreResults": (c_i, [c_p]),
    "CvodeSettings_getHaltOnEvent": (c_i, [c_p]), "CvodeSettings_getResetCvodeOnEvent": (c_i, [c_p]),
    "CvodeSettings_getPrintsteps": (c_i, [c_p]), "CvodeSettings_getEndTime": (c_d, [c_p]),
    "CvodeSettings_setHaltOnEvent": (None, [c_p, c_i]), "CvodeSettings_setResetCvodeOnEvent": (None, [c_p, c_i]),
    "CvodeSettings_setTStop": (None, [c_p, c_i]), "CvodeSettings_setJacobian": (None, [c_p, c_i]),
    "CvodeSettings_setErrors": (None, [c_p, c_d, c_d, c_i]),
    "CvodeSettings_setHaltOnSteadyState": (None, [c_p, c_i]), "CvodeSettings_setSteadyStateThreshold": (None, [c_p, c_d]),
    "CvodeSettings_setStepResolution": (None, [c_p, c_i]), "CvodeSettings_getStepResolution": (c_i, [c_p]),
    "IntegratorInstance_create": (c_p, [c_p, c_p]), "IntegratorInstance_free": (None, [c_p]),
    "IntegratorInstance_integrate": (c_i, [c_p]), "IntegratorInstance_integrateOneStep": (c_i, [c_p]),
    "IntegratorInstance_timeCourseCompleted": (c_i, [c_p]), "IntegratorInstance_getTime": (c_d, [c_p]),
    "IntegratorInstance_getResults": (c_p, [c_p]), "IntegratorInstance_reset": (c_i, [c_p]),
    "IntegratorInstance_setVariableValue": (None, [c_p, c_p, c_d]), "IntegratorInstance_getVariableValue": (c_d, [c_p, c_p]),
    "IntegratorInstance_integrateBatch": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "CvodeResults_getNout": (c_i, [c_p]), "CvodeResults_getTime": (c_d, [c_p, c_i]), "CvodeResults_getValue": (c_d, [c_p, c_p, c_i]),
    "VarySettings_allocate": (c_p, [c_i, c_i]), "VarySettings_addParameter": (c_i, [c_p, c_s, c_s]),
    "VarySettings_addDesignPoint": (c_i, [c_p, c_p]), "VarySettings_free": (None, [c_p]),
    "Model_odeSolverBatch": (c_p, [c_p, c_p, c_p]), "Model_odeSolverBatchFlat": (c_p, [c_p, c_p, c_p]),
    "Model_odeSolver": (c_p, [c_p, c_p]), "SBML_odeSolver": (c_p, [c_p, c_p]), "SBMLResults_fromIntegrator": (c_p, [c_p, c_p]),
    "SBMLResultsArray_fromBatch": (c_p, [c_p, c_p, c_p]), "SBMLResultsArray_free": (None, [c_p]),
    "SBMLResultsArray_getNumResults": (c_i, [c_p]), "SBMLResultsArray_getResults": (c_p, [c_p, c_i]),
    "SBMLResults_free": (None, [c_p]), "SBMLResults_getNout": (c_i, [c_p]), "SBMLResults_getTime": (c_p, [c_p]),
    "SBMLResults_getTimeCourse": (c_p, [c_p, c_s]), "SBMLResults_dump": (None, [c_p]),
    "TimeCourse_getName": (c_s, [c_p]), "TimeCourse_getNumValues": (c_i, [c_p]), "TimeCourse_getValue": (c_d, [c_p, c_i]),
    "BatchResults_getValue": (c_d, [c_p, c_i, c_s, c_i]), "BatchResults_free": (None, [c_p]),
    "ODESBatch_create": (c_p, [c_p, c_p, c_i, c_p, c_i, c_p, c_i]), "ODESBatch_free": (None, [c_p]),
    "ODESBatch_reserve": (c_i, [c_p, c_i]), "ODESBatch_getCapacity": (c_i, [c_p]), "ODESBatch_getNout": (c_i, [c_p]),
    "ODESBatch_getNstore": (c_i, [c_p]), "ODESBatch_getSource": (c_s, [c_p]), "ODESBatch_getCompileLog": (c_s, [c_p]),
    "ODESBatch_getKernelInfo": (c_i, [c_p] + [c_p] * 5), "ODESBatch_getBaseValues": (None, [c_p, c_p, c_p]),
    "ODESBatch_setValues": (c_i, [c_p, c_i, c_p]), "ODESBatch_getValues": (c_i, [c_p, c_i, c_p]),
    "ODESBatch_generateValues": (c_i, [c_p, c_i, C.c_ulonglong, C.c_longlong, c_p, c_p, c_p, c_i]),
    "ODESBatch_deviceValues": (c_p, [c_p]), "ODESBatch_deviceResults": (c_p, [c_p]),
    "ODESBatch_integrate": (c_i, [c_p, c_i]), "ODESBatch_synchronize": (c_i, [c_p]),
    "ODESBatch_timerStart": (c_i, [c_p]), "ODESBatch_timerStop": (C.c_float, [c_p]),
    "ODESBatch_getLastKernelMs": (C.c_float, [c_p]), "ODESBatch_getNumLaunches": (c_i, [c_p]),
    "ODESBatch_getResults": (c_i, [c_p, c_i, c_p]), "ODESBatch_getTrajectory": (c_i, [c_p, c_i, c_p]),
    "ODESBatch_getStatus": (c_i, [c_p, c_i, c_p]), "ODESBatch_getStats": (c_i, [c_p, c_i, c_p]),
    "ODESBatch_getEventLog": (c_i, [c_p, c_i, c_p]), "ODESBatch_runHost": (c_i, [c_p, c_i, c_p, c_p, c_p, c_i]),
    "ODESBatch_runHostAllDevices": (c_i, [c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_i]),
    "ODESBatch_getNflux": (c_i, [c_p]), "ODESBatch_computeFluxes": (c_i, [c_p, c_i]), "ODESBatch_getFluxes": (c_i, [c_p, c_i, c_p]),
    "ODESBatch_getLastFluxMs": (C.c_float, [c_p]), "ODESBatch_runHostEx": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i]),
    "BatchResults_getFlux": (c_d, [c_p, c_i, c_s, c_i]),
    "ODES_getCounters": (None, [c_p]), "ODEModel_freeKernelCache": (None, [c_p]),
    "ODES_deviceCount": (c_i, []), "ODES_deviceName": (c_i, [c_i, c_p, c_i]), "ODES_measureFP64Peak": (c_d, [c_i, c_i]),
    "ODES_hostAlloc": (c_p, [C.c_size_t]), "ODES_hostFree": (None, [c_p]),
    "SolverError_getNum": (c_i, [c_i]), "SolverError_getLastCode": (c_i, [c_i]), "SolverError_clear": (None, []),
    "SolverError_dumpToString": (c_p, []), "SolverError_freeDumpString": (None, [c_p]),
}

_lib = None


def lib():
    """Loads libodes_b200.so (RTLD_GLOBAL so the oracle test library can resolve against it)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C sbml_odesolver_b200/csrc).  There is no Python/CPU fallback.")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


COUNTERS = ("nvrtc_compiles", "cubin_cache_loads", "module_loads", "device_allocations", "plans_created", "plan_cache_hits",
            "kernel_launches", "host_registrations")


def counters():
    """The library's work counters (ODES_getCounters) as a dict."""
    v = (C.c_longlong * 8)()
    lib().ODES_getCounters(v)
    return dict(zip(COUNTERS, list(v)))


def take_string(ptr):
    """Copies and frees a malloc'ed C string returned by the library."""
    if not ptr:
        return None
    s = C.string_at(ptr).decode()
    C.CDLL(None).free(C.c_void_p(ptr))
    return s


def errors():
    L = lib()
    p = L.SolverError_dumpToString()
    s = C.string_at(p).decode() if p else ""
    L.SolverError_freeDumpString(p)
    return s


class OdesError(RuntimeError):
    pass


def _check(ok, what):
    if not ok:
        msg = errors()
        lib().SolverError_clear()
        raise OdesError(f"{what} failed:\n{msg}")


class Model:
    """An SBML document + the odeModel_t derived from it."""

    def __init__(self, path, globalize=()):
        L = lib()
        self.doc = L.readSBML(os.fsencode(path))
        if not self.doc:
            raise OdesError(f"cannot read SBML file {path}")
        self.sbml = L.SBMLDocument_getModel(self.doc)
        self.globalized = []
        for rid, pid in globalize:          # local kinetic-law parameter -> r_<rid>_<pid>
            _check(L.Model_globalizeParameter(self.sbml, pid.encode(), rid.encode()), f"globalize {rid}.{pid}")
            self.globalized.append(f"r_{rid}_{pid}")
        self.om = L.ODEModel_create(self.sbml)
        _check(self.om, "ODEModel_create")
        self.neq = L.ODEModel_getNeq(self.om)
        self.nass = L.ODEModel_getNumAssignments(self.om)
        self.nconst = L.ODEModel_getNumConstants(self.om)
        self.nvalues = L.ODEModel_getNumValues(self.om)
        self.names = []
        for i in range(self.nvalues):
            vi = L.ODEModel_getVariableIndexByNum(self.om, i)
            self.names.append(L.ODEModel_getVariableName(self.om, vi).decode())
            L.VariableIndex_free(vi)

    def index(self, name):
        return lib().ODEModel_getVariableIndexFields(self.om, name.encode())

    def base_values(self):
        """value[] as the model file gives it (before initial assignments)."""
        class _OM(C.Structure):
            _fields_ = [("d", c_p), ("m", c_p), ("simple", c_p), ("values", C.POINTER(c_d))]
        om = C.cast(self.om, C.POINTER(_OM)).contents
        return np.array([om.values[i] for i in range(self.nvalues)])

    def equation(self, i):
        L = lib()
        vi = L.ODEModel_getVariableIndexByNum(self.om, i)
        node = L.ODEModel_getOde(self.om, vi) if i < self.neq else (L.ODEModel_getAssignment(self.om, vi) if i < self.neq + self.nass else None)
        L.VariableIndex_free(vi)
        return take_string(L.SBML_formulaToString(node)) if node else None

    def jacobian_structure(self):
        L = lib()
        L.ODEModel_constructJacobian(self.om)

        class _NZ(C.Structure):
            _fields_ = [("i", c_i), ("j", c_i), ("ij", c_p)]
        out = []
        for k in range(L.ODEModel_getNumJacobiElements(self.om)):
            e = C.cast(L.ODEModel_getJacobiElement(self.om, k), C.POINTER(_NZ)).contents
            out.append((e.i, e.j, take_string(L.SBML_formulaToString(e.ij))))
        return out


def settings(end_time, print_step, atol, rtol, mxstep=10000, use_jacobian=1, halt_on_event=0, store_results=1, steady_state=0, ss_threshold=None,
             sensitivity=0, sens_ids=None, sens_method=0, method=0, iter_method=0, step_resolution=None):
    """CvodeSettings_createWith(Time, PrintStep, Error, RError, Mxstep, Method (0 BDF, 1 Adams-Moulton), Iter=Newton, ...).
    sensitivity=1 switches forward sensitivity analysis on (sens_method 0: simultaneous corrector, 1: staggered); sens_ids: ids of the parameters /
    variables (CvodeSettings_setSensParams), None = all constants of the model."""
    opt = lib().CvodeSettings_createWith(end_time, print_step, atol, rtol, mxstep, method, iter_method, use_jacobian, 0, halt_on_event, steady_state, store_results, sensitivity, sens_method)
    _check(opt, "CvodeSettings_createWith")
    if sens_ids is not None:
        arr = (c_s * len(sens_ids))(*[x.encode() for x in sens_ids])
        lib().CvodeSettings_setSensParams(opt, arr, len(sens_ids))
    if ss_threshold is not None:
        lib().CvodeSettings_setSteadyStateThreshold(opt, ss_threshold)
    if step_resolution is not None:
        lib().CvodeSettings_setStepResolution(opt, step_resolution)      # k > 1: sub-steps per print step; k < 0: events located inside the steps
    return opt


class Batch:
    """ODESBatch_* plan: one model, one settings object, one set of varied value[] indices."""

    def __init__(self, model, opt, vary, store=None, device=0):
        L = lib()
        self.model, self.opt = model, opt
        self.vary = np.asarray(vary, dtype=np.int32)
        self.store = None if store is None else np.asarray(store, dtype=np.int32)
        self.h = L.ODESBatch_create(model.om, opt, len(self.vary), self.vary.ctypes.data,
                                    0 if self.store is None else len(self.store),
                                    None if self.store is None else self.store.ctypes.data, device)
        _check(self.h, "ODESBatch_create")
        self.nrows = L.ODESBatch_getNout(self.h)
        self.nstore = L.ODESBatch_getNstore(self.h)
        self.nsens = L.ODESBatch_getNsens(self.h)

    def sensitivities(self, nsets):
        """Forward sensitivities [set][row][sensitivity parameter][equation]."""
        out = np.empty((nsets, self.nrows, self.nsens, self.model.neq))
        _check(lib().ODESBatch_getSensitivities(self.h, nsets, out.ctypes.data), "ODESBatch_getSensitivities")
        return out

    @property
    def nflux(self):
        return lib().ODESBatch_getNflux(self.h)

    def fluxes(self, nsets):
        """Reaction fluxes [set][row][reaction] of the stored rows, evaluated on the device (odes_flux_kernel)."""
        _check(lib().ODESBatch_computeFluxes(self.h, nsets), "ODESBatch_computeFluxes")
        out = np.empty((nsets, self.nrows, self.nflux))
        _check(lib().ODESBatch_getFluxes(self.h, nsets, out.ctypes.data), "ODESBatch_getFluxes")
        return out

    def source(self):
        return lib().ODESBatch_getSource(self.h).decode()

    def base_values(self):
        """(value[] after reset with the model's own parameters, initial trigger flags)."""
        v = np.zeros(max(self.model.nvalues, 1))
        trig = np.zeros(32, dtype=np.int32)
        lib().ODESBatch_getBaseValues(self.h, v.ctypes.data, trig.ctypes.data)
        return v[:self.model.nvalues], trig

    def reserve(self, n):
        _check(lib().ODESBatch_reserve(self.h, n), "ODESBatch_reserve")

    def set_values(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        _check(lib().ODESBatch_setValues(self.h, v.shape[0], v.ctypes.data), "ODESBatch_setValues")
        return v.shape[0]

    def get_values(self, nsets):
        v = np.empty((nsets, len(self.vary)))
        _check(lib().ODESBatch_getValues(self.h, nsets, v.ctypes.data), "ODESBatch_getValues")
        return v

    def generate_values(self, nsets, seed, first_set, base, lo, hi, log10=False):
        base, lo, hi = (np.ascontiguousarray(x, dtype=np.float64) for x in (base, lo, hi))
        _check(lib().ODESBatch_generateValues(self.h, nsets, seed, first_set, base.ctypes.data, lo.ctypes.data, hi.ctypes.data, int(log10)),
               "ODESBatch_generateValues")

    def integrate(self, nsets, sync=True):
        L = lib()
        _check(L.ODESBatch_integrate(self.h, nsets), "ODESBatch_integrate")
        if sync:
            _check(L.ODESBatch_synchronize(self.h), "ODESBatch_synchronize")
        return L.ODESBatch_getLastKernelMs(self.h)

    def synchronize(self):
        _check(lib().ODESBatch_synchronize(self.h), "ODESBatch_synchronize")
        return lib().ODESBatch_getLastKernelMs(self.h)

    def results(self, nsets):
        """Stored trajectories as a [row][value][set] VIEW of the library's [set][row][value] block."""
        out = np.empty((nsets, self.nrows, self.nstore))
        _check(lib().ODESBatch_getResults(self.h, nsets, out.ctypes.data), "ODESBatch_getResults")
        return np.transpose(out, (1, 2, 0))

    def status(self, nsets):
        st = np.empty(nsets, dtype=np.int32)
        _check(lib().ODESBatch_getStatus(self.h, nsets, st.ctypes.data), "ODESBatch_getStatus")
        return st

    def stats(self, nsets):
        st = np.empty((7, nsets), dtype=np.int32)
        _check(lib().ODESBatch_getStats(self.h, nsets, st.ctypes.data), "ODESBatch_getStats")
        return st

    def event_log(self, nsets):
        ev = np.empty((self.nrows, nsets), dtype=np.uint32)
        _check(lib().ODESBatch_getEventLog(self.h, nsets, ev.ctypes.data), "ODESBatch_getEventLog")
        return ev

    def kernel_info(self):
        v = [c_i() for _ in range(5)]
        _check(lib().ODESBatch_getKernelInfo(self.h, *[C.byref(x) for x in v]), "ODESBatch_getKernelInfo")
        return dict(zip(("regs", "smem_bytes_per_block", "local_bytes_per_thread", "threads_per_block", "blocks_per_sm"), (x.value for x in v)))

    def run_host(self, values, results=None, status=None, chunk=0, sens=None, flux=None):
        v = np.ascontiguousarray(values, dtype=np.float64)
        n = v.shape[0]
        _check(lib().ODESBatch_runHostEx(self.h, n, v.ctypes.data, None if results is None else results.ctypes.data,
                                         None if sens is None else sens.ctypes.data, None if flux is None else flux.ctypes.data,
                                         None if status is None else status.ctypes.data, chunk), "ODESBatch_runHost")

    def close(self):
        if self.h:
            lib().ODESBatch_free(self.h)
            self.h = None


# MAPK (Kholodenko 2000) kinetic-law local parameters, in reaction order: the 18 locals that the
# headline parameter scan globalises (Hill exponent n of J0 stays fixed), SURVEY section 8(d)
MAPK_LOCALS = [("J1", "V2"), ("J1", "KK2"), ("J2", "k3"), ("J2", "KK3"), ("J3", "k4"), ("J3", "KK4"),
               ("J4", "V5"), ("J4", "KK5"), ("J5", "V6"), ("J5", "KK6"), ("J6", "k7"), ("J6", "KK7"),
               ("J7", "k8"), ("J7", "KK8"), ("J8", "V9"), ("J8", "KK9"), ("J9", "V10"), ("J9", "KK10")]
