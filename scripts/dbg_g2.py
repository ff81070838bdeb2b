import ctypes, sys
sys.path.insert(0, "go-eth-kzg_b200"); sys.path.insert(0, "tests")
import kzgb200, oracle_lib
L = kzgb200.load_library()
g2 = oracle_lib.load_setup()[2]
for i in (0, 1, 64):
    m = ctypes.c_int()
    print(i, L.kzgb200_dbg_g2_selftest(g2[96 * i:96 * i + 96], ctypes.byref(m)), bin(m.value))
