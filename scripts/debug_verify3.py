import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200
from golden_util import cases, resolve
c = kzgb200.Context(commit_window=4, fk20_window=4)
for cs in [x for x in cases("verify_kzg_proof") if "correct_proof" in x["name"]][:4]:
    i = cs["input"]; C = resolve(i["commitment"]); z = resolve(i["z"]); y = resolve(i["y"]); pi = resolve(i["proof"])
    import oracle_lib
    OL = oracle_lib.lib(); m_, l_, g2_ = oracle_lib.load_setup(); G = m_[:48]
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    def msm(pts, sc):
        e = ctypes.create_string_buffer(48); OL.ko_g1_msm(b"".join(pts), b"".join(x.to_bytes(32, "big") for x in sc), ctypes.c_size_t(len(pts)), e); return e.raw
    c.L.kzgb200_dbg_set_verify_dump(1, None)
    st = c.verify_kzg_proof(C, z, y, pi)
    out = ctypes.create_string_buffer(192)
    c.L.kzgb200_dbg_set_verify_dump(0, out)
    yi = int.from_bytes(y, "big"); zi = int.from_bytes(z, "big")
    cands = {"C-yG+zpi": msm([C, G, pi], [1, (R - yi) % R, zi]), "C+yG+zpi": msm([C, G, pi], [1, yi, zi]), "C": C, "C+zpi": msm([C, pi], [1, zi]),
             "-yG": msm([G], [(R - yi) % R]), "yG": msm([G], [yi]), "-yG+zpi": msm([G, pi], [(R - yi) % R, zi])}
    print(cs["name"][-16:], "status", st, "A=", [k for k, v in cands.items() if v == out.raw[:48]], "negyG=", [k for k, v in cands.items() if v == out.raw[48:96]], "A0=", [k for k, v in cands.items() if v == out.raw[96:144]], "A1=", [k for k, v in cands.items() if v == out.raw[144:192]], flush=True)
