"""Device big-integer primitives against Python integers (bit-exact)."""
import ctypes as C
import random

import pytest

pytestmark = pytest.mark.gpu


def words(v, W):
    v &= (1 << (64 * W)) - 1
    return (C.c_uint64 * W)(*[(v >> (64 * k)) & (2**64 - 1) for k in range(W)])


def to_int(buf, n, signed=False):
    v = 0
    for k in range(n - 1, -1, -1):
        v = (v << 64) | int(buf[k])
    if signed and v >> (64 * n - 1):
        v -= 1 << (64 * n)
    return v


def signed_rand(rng, bits):
    v = rng.getrandbits(rng.randint(1, bits))
    return -v if rng.random() < 0.5 else v


@pytest.mark.parametrize("W", [1, 2, 3, 4, 5, 9, 17])
def test_primitives(W):
    from relp_b200 import _lib
    lib = _lib.load()
    rng = random.Random(W)
    M = 1 << (64 * W)
    out = (C.c_uint64 * (2 * W))()
    for _ in range(40):
        a, b, c, d = (signed_rand(rng, 64 * W - 1) for _ in range(4))
        assert lib.rg_selftest(0, W, words(a, W), words(b, W), words(c, W), words(d, W), 0, out) == 0
        assert to_int(out, W) == (a * b) % M
        assert lib.rg_selftest(1, W, words(a, W), words(b, W), words(c, W), words(d, W), 0, out) == 0
        assert to_int(out, W) == (a * b + c * d) % M
        assert lib.rg_selftest(6, W, words(a, W), words(b, W), words(c, W), words(d, W), 0, out) == 0
        assert to_int(out, W) == (a * b + c * d) % M
        s = signed_rand(rng, 63)
        assert lib.rg_selftest(2, W, words(a, W), words(b, W), words(c, W), words(d, W), s, out) == 0
        n = min(W + 2, 2 * W)
        assert to_int(out, n) == (c + a * s) % (1 << (64 * n))
        odd = abs(a) | 1
        assert lib.rg_selftest(3, W, words(odd, W), words(b, W), words(c, W), words(d, W), 0, out) == 0
        assert (to_int(out, W) * odd) % M == 1
        ua, ub = abs(a), abs(b)
        assert lib.rg_selftest(4, W, words(ua, W), words(ub, W), words(c, W), words(d, W), 0, out) == 0
        assert to_int(out, 2 * W) == ua * ub
        assert lib.rg_selftest(5, W, words(a, W), words(b, W), words(c, W), words(d, W), 0, out) == 0
        x = a * b - c * d
        assert to_int(out, 1, signed=True) == (x > 0) - (x < 0)
