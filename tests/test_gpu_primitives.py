"""GPU unit tests of the device field / group arithmetic against Python integers and the oracle."""
import random
import pytest
import oracle_lib

pytestmark = pytest.mark.gpu

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


@pytest.fixture(scope="module")
def dbg():
    import kzgb200
    return kzgb200.Debug()


def _edge(mod):
    return [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 2**32 - 1, 2**32, mod >> 1]


@pytest.mark.parametrize("field,mod", [("fp", P), ("fr", R)])
def test_field_ops(dbg, field, mod):
    rng = random.Random(1)
    a = _edge(mod) + [rng.randrange(mod) for _ in range(500)]
    b = list(reversed(_edge(mod))) + [rng.randrange(mod) for _ in range(500)]
    assert dbg.field_op(field, a, b, 0) == [x * y % mod for x, y in zip(a, b)]
    assert dbg.field_op(field, a, b, 1) == [(x + y) % mod for x, y in zip(a, b)]
    assert dbg.field_op(field, a, b, 2) == [(x - y) % mod for x, y in zip(a, b)]
    assert dbg.field_op(field, a, b, 4) == [x * x % mod for x in a]        # dedicated squaring


def test_fp_inverse(dbg):
    rng = random.Random(2)
    a = [1, 2, P - 1] + [rng.randrange(1, P) for _ in range(61)]
    assert dbg.field_op("fp", a, a, 3) == [pow(x, -1, P) for x in a]


def _msm(points, scalars):
    import ctypes
    L = oracle_lib.lib()
    out = ctypes.create_string_buffer(48)
    rc = L.ko_g1_msm(b"".join(points), b"".join(s.to_bytes(32, "big") for s in scalars), ctypes.c_size_t(len(points)), out)
    assert rc == 0
    return out.raw


def test_g1_ops_against_oracle(dbg):
    m, l, _ = oracle_lib.load_setup()
    pts = [m[48 * i:48 * i + 48] for i in range(40)]
    inf = bytes([0xc0]) + bytes(47)
    a = pts[:20] + [pts[3], inf, pts[5], inf, pts[7]]
    # b: distinct points, the same point (doubling path), infinity, and the negation (-> infinity)
    neg7 = _msm([pts[7]], [R - 1])
    b = pts[20:40] + [pts[3], pts[4], inf, inf, neg7]
    exp_add = [_msm([x, y], [1, 1]) for x, y in zip(a, b)]
    assert dbg.g1_op(a, b, 0) == exp_add
    assert dbg.g1_op(a, b, 1) == exp_add
    assert dbg.g1_op(a, b, 2) == [_msm([x], [2]) for x in a]
