"""Index maps of the multi-level forms of the FK20 G1 transform (csrc/g1fft.cuh: g1lvl_term, g1lvl8_term) against the definition
P = FFT_128(pad(IFFT_128(S)[0..63])) (toeplitz.go:113-125, fk20.go:82-90), evaluated over a small prime field that has a 128th root
of unity.  No GPU: the maps are host + device functions reached through a debug hook."""
import ctypes, os, random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, W = 257, pow(3, 2, 257)          # 3 generates F_257^*, so 3^2 has order 128


def brp7(x):
    return int(format(x, "07b")[::-1], 2)


def run_form(L, form, levels, sums):
    cur = sums
    src, e = ctypes.c_int(), ctypes.c_int()
    for lvl, (n_out, R) in enumerate(levels):
        out = [0] * 128
        for o in range(n_out):
            acc = 0
            for j in range(R):
                assert L.kzgb200_dbg_g1_level_term(form, lvl, o, j, ctypes.byref(src), ctypes.byref(e)) == 0
                assert 0 <= src.value < 128 and 0 <= e.value < 128
                acc = (acc + cur[src.value] * pow(W, e.value, P)) % P
            out[o] = acc
        cur = out
    return cur


def test_level_maps_compute_the_fk20_g1_transform():
    assert pow(W, 64, P) == P - 1
    L = ctypes.CDLL(os.path.join(ROOT, "go-eth-kzg_b200", "libkzgb200.so"))
    rng = random.Random(3)
    for _ in range(3):
        S = [rng.randrange(P) for _ in range(128)]
        sums = [0] * 128
        for f in range(128):
            sums[brp7(f)] = S[f]                                   # the MSM sums arrive with frequency f at position brp7(f)
        h = [sum(S[f] * pow(W, (-f * m) % 128, P) for f in range(128)) % P for m in range(64)]
        want = [sum(h[m] * pow(W, (m * k) % 128, P) for m in range(64)) % P for k in range(128)]
        assert run_form(L, 0, [(128, 16), (64, 8), (128, 8), (128, 8)], sums) == want
        assert run_form(L, 1, [(128, 4), (128, 4), (128, 4), (64, 2), (128, 2), (128, 4), (128, 4), (128, 2)], sums) == want
