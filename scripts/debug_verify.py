import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200, oracle_lib
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
c = kzgb200.Context(commit_window=8, fk20_window=8)
L = c.L; OL = oracle_lib.lib()
m, l, g2 = oracle_lib.load_setup()
# 1. scalar mul
pts = [m[48 * i:48 * i + 48] for i in range(1, 5)]
scal = [(12345).to_bytes(32, "big"), (R - 1).to_bytes(32, "big"), (2**255 - 19).to_bytes(32, "big"), bytes(32)]
out = ctypes.create_string_buffer(48 * 4)
print("g1_mul rc", L.kzgb200_dbg_g1_mul(b"".join(pts), b"".join(scal), out, 4))
for i in range(4):
    e = ctypes.create_string_buffer(48); OL.ko_g1_msm(pts[i], scal[i], ctypes.c_size_t(1), e)
    print(" mul", i, out.raw[48 * i:48 * i + 48] == e.raw)
# 2. dump constants
d = (ctypes.c_uint32 * 96)()
print("dump rc", L.kzgb200_dbg_dump_pairing(c.ctx, d))
od = ctypes.create_string_buffer(4 * 96)
print("oracle lines", OL.ko_dbg_dump_pairing(g2[:96], od))
names = ["gamma1", "A0", "B0", "A67"]
for s in range(4):
    for h in range(2):
        dev = sum(d[s * 24 + h * 12 + k] << (32 * k) for k in range(12))
        ora = int.from_bytes(od.raw[s * 96 + h * 48: s * 96 + h * 48 + 48], "big")
        print(" ", names[s], "c%d" % h, dev == ora)
# 3. pairings
G = m[:48]; sG = m[48:96]
def neg(p48):
    e = ctypes.create_string_buffer(48); OL.ko_g1_msm(p48, (R - 1).to_bytes(32, "big"), ctypes.c_size_t(1), e); return e.raw
inf = bytes([0xc0]) + bytes(47)
tests = [(sG, 0, neg(G), 1, 1), (G, 0, neg(G), 0, 1), (G, 0, neg(G), 1, 0), (G, 1, neg(sG), 0, 1), (inf, 0, inf, 1, 1), (G, 0, inf, 1, 0), (m[64*48:65*48], 0, neg(G), 2, 1)]
n = len(tests)
qa = (ctypes.c_int * n)(*[t[1] for t in tests]); qb = (ctypes.c_int * n)(*[t[3] for t in tests]); res = (ctypes.c_int * n)()
print("pairing rc", L.kzgb200_dbg_pairing(c.ctx, b"".join(t[0] for t in tests), qa, b"".join(t[2] for t in tests), qb, res, n))
for i, t in enumerate(tests):
    g2s = [g2[:96], g2[96:192], g2[64 * 96:65 * 96]]
    ora = OL.ko_pairing_check(t[0] + t[2], g2s[t[1]] + g2s[t[3]], ctypes.c_size_t(2))
    print(" pairing", i, "gpu", res[i], "oracle", ora, "expected", t[4])
