"""rt_device.cuh's powfGlibc — the restatement of glibc 2.39's powf (a third-party routine the reference calls through
std::pow, scene.cpp:824 etc.) — pinned against the host C library's own powf, bit for bit, on the CPU box.  The same
source runs on the device (tests/test_gpu_parity.py then requires bit-exact frames on the specular scenes)."""
import ctypes as C

import numpy as np

from helpers import shim



def _call(name, x, y):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    out = np.empty_like(x)
    fn = getattr(shim(), name)
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    fn(x.ctypes.data, y.ctypes.data, len(x), out.ctypes.data)
    return out


def host_powf(x, y):
    return _call("shim_libm_powf", x, y)     # glibc powf, called from C one pair at a time


def ours(x, y):
    return _call("shim_powf", x, y)


def assert_same(x, y):
    a, b = ours(x, y), host_powf(x, y)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    bad = np.flatnonzero(~same)
    assert bad.size == 0, [(x[i], y[i], a[i], b[i]) for i in bad[:10]]


def test_specular_domain():
    # what castRay evaluates: pow(max(0, R.-D), nSpecular) with the base in [0, 1 + eps]
    rng = np.random.default_rng(1)
    n = 4_000_000
    x = rng.random(n, dtype=np.float32)
    x[: n // 8] = np.float32(1.0) - rng.random(n // 8, dtype=np.float32) * np.float32(1e-3)    # highlights
    x[n // 8: n // 4] = rng.random(n // 8, dtype=np.float32) * np.float32(1e-3)                 # grazing
    y = rng.choice(np.array([5.0, 10.0, 2.0, 1.0, 0.5, 32.0, 100.0, 3.7, 17.25], np.float32), n)
    assert_same(x, y)


def test_sobel_domain_squares():
    # launchSSAA: powf(|G|, 2) on unclamped gradient magnitudes (scene.cpp:565)
    rng = np.random.default_rng(2)
    x = np.abs(rng.normal(size=2_000_000).astype(np.float32)) * np.float32(3.0)
    assert_same(x, np.full_like(x, 2.0))


def test_wide_random_bits_and_special_values():
    rng = np.random.default_rng(3)
    n = 2_000_000
    x = rng.integers(0, 2**32, n, dtype=np.uint32).view(np.float32)
    y = (rng.normal(size=n) * 8).astype(np.float32)
    assert_same(x, y)
    y_int = rng.integers(-12, 13, n).astype(np.float32)
    assert_same(x, y_int)
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-40, 3.4e38, 0.5, 2.0, -2.0, 3.0, -3.0,
                   1.1754944e-38, 0.99999994, 1.0000001, 126.0, -150.0, -149.5, 127.99999, 128.0], np.float32)
    xs, ys = np.meshgrid(sp, sp)
    assert_same(xs.ravel(), ys.ravel())
    # results near the overflow / underflow boundaries
    xb = np.full(200_000, 2.0, np.float32)
    yb = np.concatenate([np.linspace(125, 129, 100_000), np.linspace(-152, -124, 100_000)]).astype(np.float32)
    assert_same(xb, yb)
