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
    if field == "fp":       # sum of two products under one Montgomery reduction (needs 3p < 2^384: Fp only): x*y + (x+y)*(x-y)
        assert dbg.field_op(field, a, b, 5) == [(x * y + (x + y) * (x - y)) % mod for x, y in zip(a, b)]


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


X2 = 0xd201000000010000 ** 2          # x^2 for the BLS parameter x; lambda = -x^2 is a cube root of unity mod r


def test_glv_split_and_recoding(dbg):
    """s = k1 - k2 x^2 (mod r) with |k1|, |k2| < 2^127, digits in [-8, 8] (vmsm.cuh: glv_split, recode16_signed)"""
    rng = random.Random(11)
    half = X2 // 2
    s = [0, 1, 2, R - 1, R - 2, X2, X2 - 1, X2 + 1, half, half + 1, half - 1, half * X2, (half + 1) * X2, (half + 1) * X2 + half + 1,
         (X2 - 1) * X2, (X2 - 1) * X2 + X2 - 1, R - X2, 2**255 % R, 2**128, 2**127, 2**127 - 1]
    s = [v % R for v in s] + [rng.randrange(R) for _ in range(1000)]
    for v, (d1, d2) in zip(s, dbg.glv_digits(s)):
        assert all(-8 <= d <= 8 for d in d1 + d2)
        k1 = sum(d * 16**i for i, d in enumerate(d1)); k2 = sum(d * 16**i for i, d in enumerate(d2))
        assert abs(k1) < 2**127 and abs(k2) < 2**127
        assert (k1 - k2 * X2 - v) % R == 0, hex(v)


def test_bucket_msm_against_oracle(dbg):
    """the verifiers' variable-base MSM kernels on their own: 1 point, one work item, several work items;
    repeated points and infinity in the input; edge scalars"""
    import kzgb200
    ctx = kzgb200.Context(commit_window=8, fk20_window=8)
    try:
        m, l, _ = oracle_lib.load_setup()
        rng = random.Random(12)
        inf = bytes([0xc0]) + bytes(47)
        for n in (1, 2, 100, 128, 129, 300):
            pts = [l[48 * i:48 * i + 48] for i in range(n)]
            sc = [rng.randrange(R) for _ in range(n)]
            if n >= 100:
                pts[5] = pts[4]; pts[9] = inf; sc[7] = 0; sc[8] = R - 1; sc[10] = 1; sc[11] = X2; pts[20] = pts[21] = pts[22] = pts[4]
            assert dbg.vmsm(ctx, pts, sc) == _msm(pts, sc), n
        # all scalars equal and all points equal: every addition into a bucket is a doubling
        pts = [m[:48]] * 64
        assert dbg.vmsm(ctx, pts, [3] * 64) == _msm(pts, [3] * 64)
        # P and -P with the same scalar cancel to infinity
        neg = _msm([m[:48]], [R - 1])
        assert dbg.vmsm(ctx, [m[:48], neg], [12345, 12345]) == inf
    finally:
        ctx.close()
