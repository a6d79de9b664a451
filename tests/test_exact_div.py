"""csrc/exact_div.inc: the quotient from a precomputed correctly rounded reciprocal (five fused operations) must be
the IEEE quotient bit for bit -- the warp kernel's back substitution relies on it to stay bit-identical to the
reference's gesl, which divides (smalldense.c:150-158)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H

SRC = r'''
#include <math.h>
#include <stdbool.h>
#define DEV static inline
#define ODES_HOST_EMULATION 1        /* the device-only sequence of the file is checked on the GPU (below) */
#include "%(root)s/sbml_odesolver_b200/csrc/exact_div.inc"
long check(const double *a, const double *d, long n)
{
  long bad = 0;
  for (long i = 0; i < n; i++) {
    const double y = 1.0 / d[i];
    const double q = (odes_div_safe(a[i]) && odes_div_safe(d[i])) ? odes_div_by_rcp(a[i], d[i], y) : a[i] / d[i];
    const double want = a[i] / d[i];
    if (!(q == want) && !(q != q && want != want)) bad++;
    if (q == 0.0 && want == 0.0 && signbit(q) != signbit(want)) bad++;
  }
  return bad;
}
'''


def _lib(tmp_path):
    c = tmp_path / "d.c"
    c.write_text(SRC % {"root": H.ROOT})
    so_ = tmp_path / "d.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", str(c), "-o", str(so_), "-lm"])
    lib = C.CDLL(str(so_))
    lib.check.restype = C.c_long
    lib.check.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    return lib


def _bad(lib, a, d):
    a = np.ascontiguousarray(a, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    return lib.check(a.ctypes.data, d.ctypes.data, len(a))


def test_quotient_from_reciprocal_is_the_ieee_quotient(tmp_path):
    lib = _lib(tmp_path)
    rng = np.random.default_rng(11)
    n = 4_000_000
    # random mantissas and signs over a wide exponent range (inside and outside the guards)
    for span in (3, 40, 200):
        a = rng.uniform(1, 2, n) * np.exp2(rng.integers(-span, span, n)) * rng.choice([-1.0, 1.0], n)
        d = rng.uniform(1, 2, n) * np.exp2(rng.integers(-span, span, n)) * rng.choice([-1.0, 1.0], n)
        assert _bad(lib, a, d) == 0
    # structured: divisors next to powers of two and with all-ones mantissas, quotients next to rounding boundaries
    k = np.arange(1, 200_001, dtype=np.float64)
    eps = np.finfo(float).eps
    near_one = np.concatenate([1 + k * eps, 2 - k * eps / 1, 1 + k * eps * 2 ** 20])
    a = rng.uniform(1, 2, len(near_one))
    assert _bad(lib, a, near_one) == 0
    assert _bad(lib, near_one, rng.uniform(1, 2, len(near_one))) == 0
    assert _bad(lib, near_one, near_one[::-1].copy()) == 0
    # exact quotients, integers, and (a = q d rounded) +- 1 ulp
    q = rng.integers(1, 2 ** 26, n).astype(np.float64)
    d = rng.integers(1, 2 ** 26, n).astype(np.float64)
    assert _bad(lib, q * d, d) == 0
    assert _bad(lib, np.nextafter(q * d, np.inf), d) == 0
    assert _bad(lib, np.nextafter(q * d, -np.inf), d) == 0
    d = rng.uniform(1, 2, n)
    qd = rng.uniform(1, 2, n) * d
    assert _bad(lib, np.nextafter(qd, 4.0), d) == 0
    # zeros, infinities, not-a-numbers, denormals: the guards send them to the division
    sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1e308, -1e-200, 3.0])
    aa, dd = np.meshgrid(sp, sp)
    assert _bad(lib, aa.ravel(), dd.ravel()) == 0


@pytest.mark.gpu
def test_division_sequence_with_kept_reciprocal_is_the_division_on_the_gpu(tmp_path):
    """odes_div_seq / odes_rcp_seq (the default of the warp kernel's back substitution): the compiler's own division
    sequence with its reciprocal part kept by the factorisation, against a / d on 6e8 operand pairs on the device
    (tests/cuda/div_seq_check.cu: random and structured mantissas, range tests, zeros, infinities, NaNs, denormals)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "div_seq_check"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false", "-O2", "-o", str(exe),
                           os.path.join(H.ROOT, "tests", "cuda", "div_seq_check.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([str(exe), "2048"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
