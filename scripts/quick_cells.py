import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200, oracle_lib
cw = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fw = int(sys.argv[2]) if len(sys.argv) > 2 else 12
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 256
t = time.time(); c = kzgb200.Context(commit_window=cw, fk20_window=fw); print("init %.2fs" % (time.time() - t), c.info())
blobs = [oracle_lib.rand_blob(b << 20) for b in range(4)]
o = oracle_lib.get_oracle()
got = c.compute_cells_and_kzg_proofs_batch(blobs[:2])
print("parity:", all(g == o.compute_cells_and_kzg_proofs(b) for g, b in zip(got, blobs)))
big = (blobs * ((nb + 3) // 4))[:nb]
for rep in range(3):
    t = time.time(); c.compute_cells_and_kzg_proofs_batch(big); dt = time.time() - t
    print("n=%d wall %.1f ms, device %.1f ms -> %.0f blobs/s (device)" % (nb, dt * 1e3, c.last_device_ms(), nb / c.last_device_ms() * 1e3))
