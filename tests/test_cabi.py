"""The drop-in boundary: libodes_b200.so loads, exports every function include/sbmlsolver/*.h declares,
and fails loudly (SOLVER_ERROR_CUDA_UNAVAILABLE, no CPU fallback) when there is no CUDA device."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so

DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?[\s\*]([A-Za-z_]\w*)\s*\(([^;{}]*)\)\s*;", re.M)


def declared_functions():
    names = set()
    for path in sorted(glob.glob(os.path.join(H.ROOT, "include", "sbmlsolver", "*.h"))):
        text = re.sub(r"/\*.*?\*/", " ", open(path).read(), flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        text = re.sub(r"typedef[^;]*\(\s*\*\s*\w+\s*\)[^;]*;", " ", text)      # function-pointer typedefs
        for m in DECL.finditer(text):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(so.LIB_PATH)
    names = declared_functions()
    assert len(names) > 120, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    for must in ("IntegratorInstance_integrateBatch", "ODESBatch_create", "ODESBatch_runHost", "Model_odeSolverBatch",
                 "Compiler_compile", "CompiledCode_getFunction", "CompiledCode_free", "IntegratorInstance_integrateOneStep",
                 "CvodeSettings_createWith", "ODEModel_createFromFile", "VarySettings_allocate"):
        assert must in names, must


def test_python_binding_signatures_resolve():
    L = so.lib()
    for name in so._SIGS:
        assert getattr(L, name) is not None


def test_no_torch_types_in_headers():
    for path in glob.glob(os.path.join(H.ROOT, "include", "**", "*.h"), recursive=True):
        text = re.sub(r"/\*.*?\*/", " ", open(path).read(), flags=re.S)
        assert "torch" not in text and "at::" not in text and "#include <cuda" not in text, path


@pytest.mark.skipif(so.lib().ODES_deviceCount() > 0, reason="needs a box without a GPU")
def test_launch_entries_fail_loudly_without_a_device():
    L = so.lib()
    L.SolverError_clear()
    w = H.workload("mapk", nout=4)
    b = so.Batch(w["model"], w["opt"], w["vary"])          # code generation + NVRTC compile need no device
    assert "odes_bdf_kernel" in b.source()
    with pytest.raises(so.OdesError) as e:
        b.reserve(32)
    assert "no CPU fallback" in str(e.value) or "150000" in str(e.value) or "CUDA" in str(e.value)
    vals = H.workload_values(w, 4, seed=0)
    with pytest.raises(so.OdesError):
        b.set_values(vals)
    b.close()
    # the reference-shaped single-instance entry has no host integrator behind it either
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(10.0, 2, 1e-9, 1e-6)
    ii = L.IntegratorInstance_create(m.om, opt)
    if ii:
        assert L.IntegratorInstance_integrate(ii) == 0
        L.IntegratorInstance_free(ii)
    assert L.SolverError_getNum(1) + L.SolverError_getNum(0) > 0
    L.SolverError_clear()


def test_compiler_seam_three_call_shape():
    """Compiler_compile / CompiledCode_getFunction / CompiledCode_free (src/compiler.h:89-107) on a tiny CUDA unit:
    NVRTC targets sm_100a without a device; loading the module needs one."""
    L = C.CDLL(so.LIB_PATH)
    L.Compiler_compile.restype = C.c_void_p
    L.Compiler_compile.argtypes = [C.c_char_p]
    L.CompiledCode_free.argtypes = [C.c_void_p]
    L.CompiledCode_getFunction.restype = C.c_void_p
    L.CompiledCode_getFunction.argtypes = [C.c_void_p, C.c_char_p]
    code = L.Compiler_compile(b'extern "C" __global__ void k(double *x) { x[threadIdx.x] = 2.0 * x[threadIdx.x]; }\n')
    assert code
    if so.lib().ODES_deviceCount() == 0:
        assert not L.CompiledCode_getFunction(code, b"k")
        so.lib().SolverError_clear()
    L.CompiledCode_free(code)
    bad = L.Compiler_compile(b"this is not CUDA\n")
    assert not bad
    assert so.lib().SolverError_getNum(1) > 0               # ERROR_ERROR_TYPE
    so.lib().SolverError_clear()


def test_vary_settings_and_design_points():
    L = so.lib()
    vs = L.VarySettings_allocate(2, 3)
    assert vs
    assert L.VarySettings_addParameter(vs, b"V1", None) == 1
    assert L.VarySettings_addParameter(vs, b"KK2", b"J1") == 2
    for row in ([2.5, 8.0], [1.0, 4.0], [5.0, 16.0]):
        arr = (C.c_double * 2)(*row)
        assert L.VarySettings_addDesignPoint(vs, arr) > 0
    L.VarySettings_free(vs)
