"""csrc/bingcd.cuh (binary-GCD modular inversion) compiled for the HOST and checked against Python integers."""
import ctypes, os, random, subprocess, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _lib(tmp):
    so = os.path.join(tmp, "bingcd_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "bingcd_host.cpp"), "-o", so])
    return ctypes.CDLL(so)


def _limbs(vals, n):
    arr = (ctypes.c_uint32 * (len(vals) * n))()
    for i, v in enumerate(vals):
        for k in range(n):
            arr[i * n + k] = (v >> (32 * k)) & 0xffffffff
    return arr


def test_inverse_matches_python():
    rng = random.Random(3)
    with tempfile.TemporaryDirectory() as tmp:
        L = _lib(tmp)
        for mod, n in ((P, 12), (R, 8)):
            ys = [1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, 2**31, 2**31 - 1, 2**32, 2**33 + 1, 2**64 - 1, 2**64, 2**65 + 5,
                  2**(32 * n - 4) % mod, pow(2, -1, mod), pow(3, -1, mod)]
            ys += [rng.randrange(1, mod) for _ in range(20000)]
            ys += [rng.randrange(1, 2**k) for k in (8, 31, 32, 33, 63, 64, 65, 96, 128, 200) for _ in range(200)]
            ninv = (-pow(mod, -1, 2**31)) % 2**31
            out = (ctypes.c_uint32 * (len(ys) * n))()
            worst = L.bingcd_inverse(n, _limbs(ys, n), _limbs([mod], n), ninv, out, len(ys))
            got = [sum(out[i * n + k] << (32 * k) for k in range(n)) for i in range(len(ys))]
            bad = [(hex(y), hex(g)) for y, g in zip(ys, got) if g != pow(y, -1, mod)]
            assert not bad, bad[:3]
            assert worst <= (2 * 32 * n) // 31 + 2, worst          # rounds actually needed stay within the loop bound
            print(n, "limbs: worst outer rounds", worst)
