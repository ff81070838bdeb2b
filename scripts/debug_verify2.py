import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200, oracle_lib
from golden_util import cases, resolve
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
c = kzgb200.Context(commit_window=8, fk20_window=8)
L = c.L; OL = oracle_lib.lib()
m, l, g2 = oracle_lib.load_setup()
G = m[:48]
def msm(pts, sc):
    e = ctypes.create_string_buffer(48); OL.ko_g1_msm(b"".join(pts), b"".join(x.to_bytes(32, "big") for x in sc), ctypes.c_size_t(len(pts)), e); return e.raw
for cs in [x for x in cases("verify_kzg_proof") if "correct_proof" in x["name"]][:8]:
    i = cs["input"]; C = resolve(i["commitment"]); z = resolve(i["z"]); y = resolve(i["y"]); pi = resolve(i["proof"])
    out = ctypes.create_string_buffer(336)
    L.kzgb200_dbg_verify_steps(c.ctx, C, pi, z, y, out)
    zi = int.from_bytes(z, "big"); yi = int.from_bytes(y, "big")
    e_yG = msm([G], [yi]); e_zpi = msm([pi], [zi]); e_A = msm([C, G, pi], [1, (R - yi) % R, zi])
    o = out.raw
    print(cs["name"][-16:], "yG", o[:48] == e_yG, "zPi", o[48:96] == e_zpi, "A", o[96:144] == e_A, "-yG", o[144:192] == msm([G], [(R - yi) % R]), "C-yG", o[192:240] == msm([C, G], [1, (R - yi) % R]), "swapped", o[240:288] == msm([C, G], [1, (R - yi) % R]), "ool", o[288:336] == msm([C, G], [1, (R - yi) % R]),
          "status", c.verify_kzg_proof(C, z, y, pi))
