"""The kernel evaluates pow exactly as the host libm does (csrc/glibc_pow.inc + glibc_pow_tables.h): this is what
makes GPU trajectories bit-identical to the reference's CVODE, whose step-size controller calls libm's pow.  The
restatement is checked here against THIS machine's libm, so a libm with a different pow (another glibc build, a
non-FMA CPU) is noticed in the CPU tier instead of as a parity drift on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np

import helpers as H

SRC = r'''
#include <math.h>
#include <string.h>
#define ODES_HOST_EMULATION 1
#define DEV static inline
#include "%(root)s/sbml_odesolver_b200/csrc/glibc_pow_tables.h"
#include "%(root)s/sbml_odesolver_b200/csrc/glibc_pow.inc"
long check(const double *x, const double *y, long n) { long bad = 0; for (long i = 0; i < n; i++) if (odes_pow(x[i], y[i]) != pow(x[i], y[i])) bad++; return bad; }
static int same(double a, double b) { return a == b || (a != a && b != b); }
long check_exp(const double *x, long n) { long bad = 0; for (long i = 0; i < n; i++) if (!same(odes_exp(x[i]), exp(x[i]))) bad++; return bad; }
long check_log(const double *x, long n) { long bad = 0; for (long i = 0; i < n; i++) if (!same(odes_log(x[i]), log(x[i]))) bad++; return bad; }
'''


def _lib(tmp_path):
    c = tmp_path / "p.c"
    c.write_text(SRC % {"root": H.ROOT})
    so_ = tmp_path / "p.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", str(c), "-o", str(so_), "-lm"])
    lib = C.CDLL(str(so_))
    lib.check.restype = C.c_long
    lib.check.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    for f in (lib.check_exp, lib.check_log):
        f.restype, f.argtypes = C.c_long, [C.c_void_p, C.c_long]
    return lib


def test_restated_pow_is_bit_identical_to_libm(tmp_path):
    lib = _lib(tmp_path)
    rng = np.random.default_rng(5)
    n = 2_000_000
    # the controller's calls: (BIAS * error ratio) ^ (1/n), n = 1..7
    x = np.power(10.0, rng.uniform(-25, 25, n))
    y = 1.0 / rng.integers(1, 8, n).astype(np.float64)
    assert lib.check(x.ctypes.data, y.ctypes.data, n) == 0
    # model expressions: Hill-type exponents, integer powers, negative exponents
    x = np.power(10.0, rng.uniform(-8, 8, n))
    y = np.concatenate([rng.uniform(-6, 6, n // 2), rng.integers(-10, 11, n - n // 2).astype(np.float64)])
    assert lib.check(x.ctypes.data, y.ctypes.data, n) == 0
    # negative bases with integral exponents (sign handled on top of pow(|x|, y), as glibc does)
    x = -np.power(10.0, rng.uniform(-8, 8, n))
    y = rng.integers(-10, 11, n).astype(np.float64)
    assert lib.check(x.ctypes.data, y.ctypes.data, n) == 0
    # edge arguments take the libm fallback inside odes_pow: still identical on the host
    x = np.array([0.0, -2.0, 1.0, 1e-320, 1e305, 2.0, 2.0, np.inf, 3.0])
    y = np.array([0.5, 3.0, 1e300, 0.5, 2.0, 0.0, 1e-20, 0.5, np.nan])
    assert lib.check(x.ctypes.data, y.ctypes.data, len(x)) <= 1          # the NaN exponent compares unequal to itself


def test_restated_exp_and_log_are_bit_identical_to_libm(tmp_path):
    """exp and ln of model expressions (odes_exp, odes_log): the host libm's values on the main paths, the device
    library outside them (on the host both sides are libm then)."""
    lib = _lib(tmp_path)
    rng = np.random.default_rng(6)
    n = 2_000_000
    for x in (rng.uniform(-700, 700, n), rng.uniform(-5, 5, n), rng.uniform(-1e-3, 1e-3, n), rng.uniform(-1e-17, 1e-17, n),
              np.array([0.0, -0.0, 1e-320, 709.0, 710.0, -745.0, -746.0, np.inf, -np.inf, np.nan, 511.9999, 512.0])):
        x = np.ascontiguousarray(x)
        assert lib.check_exp(x.ctypes.data, len(x)) == 0
    for x in (np.power(10.0, rng.uniform(-300, 300, n)), rng.uniform(0.9, 1.1, n), rng.uniform(0.9375, 1.0647, n), 1 + rng.uniform(-1e-9, 1e-9, n),
              np.array([1.0, 0.0, -0.0, -1.0, 5e-324, 1e-310, np.inf, np.nan, 0.9375, 1.064697265625, 2.0, 0.5])):
        x = np.ascontiguousarray(x)
        assert lib.check_log(x.ctypes.data, len(x)) == 0


def test_tables_header_matches_this_libm():
    """tools/gen_pow_tables.py regenerates the header from the local libm; the committed one must be identical."""
    hdr = os.path.join(H.ROOT, "sbml_odesolver_b200", "csrc", "glibc_pow_tables.h")
    before = open(hdr).read()
    r = subprocess.run(["python", os.path.join(H.ROOT, "tools", "gen_pow_tables.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches against libm pow on 6e7 samples: 0" in r.stdout
    assert open(hdr).read() == before
