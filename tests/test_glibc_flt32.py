"""crazyara_b200/csrc/glibc_flt32.cuh restates glibc's powf / logf (the functions std::pow / std::log resolve to in
apply_temperature, util/blazeutil.h:78-88, and in libstdc++'s gamma sampler behind get_dirichlet_noise, :113-124) so
that the device produces the reference's bits.  Pinned here against the live libm of the box:
  * CPU (host build of the same header): strided sweeps over EVERY binade of x for the exponents the engine uses
    (1/1.7 default, 1/alpha of the Dirichlet sampler, others), random (x, y) pairs, the special-case branches;
    the exhaustive sweep (stride 1, 2^31 inputs per exponent, ~1 min each) was run once when the port was written:
    0 mismatches for logf and for powf with y in {1/1.7, 1/0.3, 5, 2, 0.5, 1/1.3, 10, 0.01, 100}.
  * GPU: the device build of the header through ara_debug_powf_logf on the same kind of inputs."""
import ctypes

import numpy as np
import pytest


def _he():
    from tests import hostemu
    L = hostemu.lib()
    L.he_powf_sweep.restype = ctypes.c_ulonglong
    L.he_powf_sweep.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_float, ctypes.POINTER(ctypes.c_uint)]
    L.he_logf_sweep.restype = ctypes.c_ulonglong
    L.he_logf_sweep.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint)]
    L.he_powf.restype = ctypes.c_float
    L.he_powf.argtypes = [ctypes.c_float, ctypes.c_float]
    L.he_logf.restype = ctypes.c_float
    L.he_logf.argtypes = [ctypes.c_float]
    return L


_libm = ctypes.CDLL("libm.so.6")
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]

EXPONENTS = [1 / 1.7, 1 / 0.3, 1 / 0.2, 1 / 0.6, 2.0, 0.5, 1 / 1.3, 1 / 0.8, 1 / 2.5, 10.0, 0.01, 100.0]


def test_logf_equals_libm_on_a_sweep_of_all_positive_floats():
    bad = ctypes.c_uint(0)
    # 0 .. +inf inclusive, every 61st bit pattern (35 M inputs), plus every pattern of the first and last binades
    assert _he().he_logf_sweep(0, 0x7F800001, 61, ctypes.byref(bad)) == 0, hex(bad.value)
    assert _he().he_logf_sweep(0, 0x01000000, 1, ctypes.byref(bad)) == 0, hex(bad.value)
    assert _he().he_logf_sweep(0x3F000000, 0x40000000, 1, ctypes.byref(bad)) == 0, hex(bad.value)  # [0.5, 2)


@pytest.mark.parametrize("y", EXPONENTS)
def test_powf_equals_libm_on_a_sweep_of_all_positive_floats(y):
    bad = ctypes.c_uint(0)
    assert _he().he_powf_sweep(0, 0x7F800001, 211, y, ctypes.byref(bad)) == 0, hex(bad.value)
    # the priors live in (0, 1]: every bit pattern of [2^-10, 1] for the default exponent, a denser sweep for the others
    step = 1 if y == EXPONENTS[0] else 7
    assert _he().he_powf_sweep(0x3A800000, 0x3F800001, step, y, ctypes.byref(bad)) == 0, hex(bad.value)


def test_powf_logf_special_cases_and_random_pairs():
    L = _he()
    bits = lambda f: np.float32(f).view(np.uint32)  # noqa: E731
    same = lambda a, b: bits(a) == bits(b) or (np.isnan(a) and np.isnan(b))  # noqa: E731
    tiny, huge = np.float32(1e-45), np.float32(3e38)
    for x in (0.0, 1.0, tiny, 1e-40, 1.17549435e-38, huge, np.inf, 0.99999994, 1.0000001):
        assert same(L.he_logf(x), _libm.logf(x)), x
        for y in (0.0, 1.0, 0.5, 1 / 1.7, 3.0, 200.0, 1e-3, np.inf, 127.9, 150.0):
            assert same(L.he_powf(x, y), _libm.powf(x, y)), (x, y)
    rng = np.random.default_rng(5)
    xs = np.exp(rng.uniform(-100, 80, 200000)).astype(np.float32)
    ys = np.exp(rng.uniform(-6, 6, 200000)).astype(np.float32)
    for x, y in zip(xs[:50000], ys[:50000]):
        assert same(L.he_powf(float(x), float(y)), _libm.powf(float(x), float(y))), (x, y)


@pytest.mark.gpu
def test_device_powf_logf_equal_libm():
    """The DEVICE build (__fma_rn / __dmul_rn / __dadd_rn, tables in device memory) against the box's libm: 4 x 2 M
    (x, y) pairs -- priors with the default exponent, priors with the engine's other exponents, wide-range pairs."""
    import crazyara_b200
    L = crazyara_b200.lib()
    L.ara_debug_powf_logf.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 2
    he = _he()
    he.he_libm_powf_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    he.he_libm_logf_array.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rng = np.random.default_rng(11)
    n = 1 << 21
    x_prior = rng.integers(0x00000001, 0x3F800001, n, dtype=np.uint32).view(np.float32)
    x_wide = np.exp(rng.uniform(-103, 88, n)).astype(np.float32)
    for x, y in ((x_prior, np.full(n, 1 / 1.7, np.float32)),
                 (x_prior, rng.choice(np.array(EXPONENTS, np.float32), n)),
                 (x_wide, np.exp(rng.uniform(-6, 6, n)).astype(np.float32)),
                 (x_wide, np.full(n, 1 / 0.3, np.float32))):
        x = np.ascontiguousarray(x)
        y = np.ascontiguousarray(y)
        p, lg, ref_p, ref_l = (np.zeros(n, np.float32) for _ in range(4))
        assert L.ara_debug_powf_logf(x.ctypes.data, y.ctypes.data, n, p.ctypes.data, lg.ctypes.data) == 0
        he.he_libm_powf_array(x.ctypes.data, y.ctypes.data, n, ref_p.ctypes.data)
        he.he_libm_logf_array(x.ctypes.data, n, ref_l.ctypes.data)
        assert np.array_equal(p.view(np.uint32), ref_p.view(np.uint32))
        assert np.array_equal(lg.view(np.uint32), ref_l.view(np.uint32))
