import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kzgb200, oracle_lib
cw = int(sys.argv[1]) if len(sys.argv) > 1 else 13
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = kzgb200.Debug()
print("imad peak: %.3f T/s" % (d.imad_peak() / 1e12))
t = time.time(); c = kzgb200.Context(commit_window=cw); print("init %.2fs" % (time.time() - t), c.info())
blobs = [oracle_lib.rand_blob(b << 20) for b in range(8)]
o = oracle_lib.get_oracle()
got = c.blob_to_kzg_commitment_batch(blobs)
print("parity:", all(g == (0, o.blob_to_kzg_commitment(b)[1]) for g, b in zip(got, blobs)))
big = (blobs * ((nb + 7) // 8))[:nb]
for rep in range(3):
    t = time.time(); c.blob_to_kzg_commitment_batch(big); dt = time.time() - t
    print("n=%d wall %.1f ms, device %.1f ms -> %.0f blobs/s (device)" % (nb, dt * 1e3, c.last_device_ms(), nb / c.last_device_ms() * 1e3))
