// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU oracle, part 2: a restatement of go-eth-kzg's hot path with the REFERENCE'S algorithmic
// structure (generic Pippenger without precompute, 128 x MSM-64 per blob, G1 FFTs with one full
// scalar multiplication per twiddle, sequential per-blob verify loop).  Each function cites the
// reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product (libkzgb200.so)
// never does.
//
// Parity status: PINNED -- tests/test_oracle_vectors.py runs all 311 consensus-spec vectors
// (tests/golden/, packed from /root/reference/tests by tests/golden/make_golden.py) and the
// reference's KATs (fiatshamir_test.go:14-26, api_test.go:18-27) through this file.
#include "bls12_381.hpp"
#include "sha256.hpp"
#include <algorithm>
#include <thread>
#include <atomic>
#include <cstdio>

using namespace ko;

namespace {

enum Status {
    ST_OK = 0, ST_VERIFY_FAILED = 1, ST_NON_CANONICAL_SCALAR = 2, ST_BAD_G1_ENCODING = 3,
    ST_NOT_ON_CURVE = 4, ST_NOT_IN_SUBGROUP = 5, ST_LENGTH_MISMATCH = 6, ST_BAD_CELL_INDEX = 7,
    ST_CELL_IDS_NOT_ASCENDING = 8, ST_NOT_ENOUGH_CELLS = 9, ST_ROW_INDEX = 10, ST_BAD_ARGS = 11
};

const int N_BLOB = 4096, N_EXT = 8192, CELL = 64, N_CELLS = 128;

// ---- internal/domain/domain.go:127-149 -------------------------------------------------
static inline uint64_t brp_int(uint64_t k, uint64_t n) {
    int bits = __builtin_ctzll(n);
    uint64_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((k >> i) & 1) << (bits - 1 - i);
    return r;
}
template <class T> static void bit_reverse(T *a, size_t n) {
    for (size_t i = 0; i < n; ++i) { size_t j = brp_int(i, n); if (j > i) std::swap(a[i], a[j]); }
}

// ---- internal/domain/domain.go:21-98 ---------------------------------------------------
struct Domain {
    uint64_t n;
    Fr n_inv, gen, gen_inv;
    std::vector<Fr> roots, inv_roots;
    void init(uint64_t x) {
        n = x;
        // 2^32-th root of unity 10238227357739495823651030575849232062558860180284477541189508159991286009131
        static const u64 w32[4] = {0x3829971f439f0d2bULL, 0xb63683508c2280b9ULL, 0xd09b681922c813b4ULL, 0x16a2a19edfe81f20ULL};
        Fr root = Fr::from_limbs(w32);
        int logx = __builtin_ctzll(x);
        gen = root.pow_u64(1ULL << (32 - logx));
        gen_inv = gen.inv();
        n_inv = Fr::from_u64(x).inv();
        roots.resize(x);
        Fr cur = Fr::one();
        for (uint64_t i = 0; i < x; ++i) { roots[i] = cur; cur = cur * gen; }
        inv_roots = roots;
        batch_invert(inv_roots.data(), x);
    }
    void reverse_roots() { bit_reverse(roots.data(), n); bit_reverse(inv_roots.data(), n); }
};

// ---- internal/domain/fft.go:109-144 (radix-2 DIF + final bit reversal) -----------------
static void fft_fr(Fr *v, size_t n, const Fr &root) {
    if (n <= 1) return;
    for (size_t size = n; size >= 2; size >>= 1) {
        size_t half = size >> 1;
        Fr wstep = root.pow_u64(n / size);
        for (size_t start = 0; start < n; start += size) {
            Fr w = Fr::one();
            for (size_t k = 0; k < half; ++k) {
                Fr a = v[start + k], b = v[start + k + half];
                v[start + k] = a + b;
                v[start + k + half] = (a - b) * w;
                w = w * wstep;
            }
        }
    }
    bit_reverse(v, n);
}
static void ifft_fr(Fr *v, size_t n, const Domain &d) {   // fft.go:100-107
    fft_fr(v, n, d.gen_inv);
    for (size_t i = 0; i < n; ++i) v[i] = v[i] * d.n_inv;
}
// ---- internal/domain/coset_fft.go:41-70 ------------------------------------------------
static void coset_fft_fr(Fr *v, size_t n, const Domain &d, const Fr &coset_gen) {
    Fr s = Fr::one();
    for (size_t i = 0; i < n; ++i) { v[i] = v[i] * s; s = s * coset_gen; }
    fft_fr(v, n, d.gen);
}
static void coset_ifft_fr(Fr *v, size_t n, const Domain &d, const Fr &inv_coset_gen) {
    ifft_fr(v, n, d);
    Fr s = Fr::one();
    for (size_t i = 0; i < n; ++i) { v[i] = v[i] * s; s = s * inv_coset_gen; }
}

// ---- internal/domain/fft.go:49-92: recursive G1 FFT, one scalar-mul per non-trivial twiddle
static std::vector<G1Jac> fft_g1_rec(const std::vector<G1Jac> &values, const Fr &root) {
    size_t n = values.size();
    if (n == 1) return values;
    Fr root2 = root.sqr();
    std::vector<G1Jac> even(n / 2), odd(n / 2);
    for (size_t i = 0; i < n / 2; ++i) { even[i] = values[2 * i]; odd[i] = values[2 * i + 1]; }
    std::vector<G1Jac> fe = fft_g1_rec(even, root2), fo = fft_g1_rec(odd, root2);
    std::vector<G1Jac> out(n, G1Jac::infinity());
    Fr w = Fr::one();
    for (size_t k = 0; k < n / 2; ++k) {
        G1Jac t = (w == Fr::one()) ? fo[k] : g1_mul_fr(fo[k], w);
        out[k] = fe[k].add(t);
        out[k + n / 2] = fe[k].add(t.neg());
        w = w * root;
    }
    return out;
}
static void fft_g1(std::vector<G1Jac> &v, const Domain &d) { v = fft_g1_rec(v, d.gen); }
static void ifft_g1(std::vector<G1Jac> &v, const Domain &d) {   // fft.go:32-42
    v = fft_g1_rec(v, d.gen_inv);
    for (auto &p : v) p = g1_mul_fr(p, d.n_inv);
}

// ---- internal/multiexp/multiexp.go:20-26 -> gnark MultiExp: bucket method, signed digits,
// no precomputation.
static G1Jac msm(const G1Affine *pts, const Fr *scalars, size_t n) {
    if (n == 0) return G1Jac::infinity();
    int c = 2;
    while ((1u << (c + 3)) <= n && c < 14) ++c;      // ~log2(n) - 2
    int nwin = (255 + c) / c + 1;
    std::vector<int32_t> digits(n * nwin);
    for (size_t i = 0; i < n; ++i) {
        u64 k[4]; scalars[i].to_limbs(k);
        int carry = 0;
        for (int w = 0; w < nwin; ++w) {
            int bit = w * c;
            int64_t d = carry;
            if (bit < 256) {
                u64 lo = k[bit / 64] >> (bit % 64);
                if (bit % 64 + c > 64 && bit / 64 + 1 < 4) lo |= k[bit / 64 + 1] << (64 - bit % 64);
                d += (int64_t)(lo & ((1ULL << c) - 1));
            }
            if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
            digits[i * nwin + w] = (int32_t)d;
        }
    }
    G1Jac acc = G1Jac::infinity();
    std::vector<G1Jac> buckets(1u << (c - 1));
    for (int w = nwin - 1; w >= 0; --w) {
        for (int s = 0; s < c; ++s) acc = acc.dbl();
        for (auto &b : buckets) b = G1Jac::infinity();
        for (size_t i = 0; i < n; ++i) {
            int32_t d = digits[i * nwin + w];
            if (d > 0) buckets[d - 1] = buckets[d - 1].add_affine(pts[i]);
            else if (d < 0) buckets[-d - 1] = buckets[-d - 1].add_affine(pts[i].neg());
        }
        G1Jac run = G1Jac::infinity(), sum = G1Jac::infinity();
        for (size_t b = buckets.size(); b-- > 0;) { run = run.add(buckets[b]); sum = sum.add(run); }
        acc = acc.add(sum);
    }
    return acc;
}

// ---- the Context: api.go:17-28, 90-149 -------------------------------------------------
struct Ctx {
    Domain dom4096, dom8192, dom128, dom64;
    std::vector<G1Affine> g1_monomial;       // natural order
    std::vector<G1Affine> g1_lagrange_brp;   // bit-reversed (api.go:131)
    G2Affine g2_gen, g2_alpha, g2_s64;
    G1Affine g1_gen;
    // kzg_multi OpeningKey (internal/kzg_multi/srs.go:60-103)
    std::vector<Fr> coset_shift, coset_shift_inv, coset_shift_pow64;
    // FK20 (internal/kzg_multi/fk20/toeplitz.go:50-93): table[i][j], i<128, j<64
    std::vector<G1Affine> fk_table;
    // erasure_code.go:46-72
    Fr rec_coset_gen, rec_coset_gen_inv;
};

static void batch_to_affine(const std::vector<G1Jac> &in, G1Affine *out) {
    for (size_t i = 0; i < in.size(); ++i) out[i] = in[i].to_affine();
}

// fk20.go:23-52 + toeplitz.go:50-93
static void build_fk20_table(Ctx &c) {
    std::vector<G1Affine> srs(c.g1_monomial.rbegin(), c.g1_monomial.rend());   // slices.Reverse
    // srsTruncated = srs[64:], takeEveryNth(.., 64): vector j = {srsTruncated[j + 64*m]}, 63 points
    c.fk_table.assign(N_CELLS * CELL, G1Affine::infinity());
    std::vector<std::vector<G1Jac>> cols(CELL);
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&] {
        for (;;) {
            int j = next.fetch_add(1);
            if (j >= CELL) break;
            std::vector<G1Jac> v;
            for (size_t m = j + CELL; m < srs.size(); m += CELL) v.push_back(G1Jac::from_affine(srs[m]));
            while (v.size() < 128) v.push_back(G1Jac::infinity());   // pad 63 -> 64 -> 128 (fk20.go:141-167)
            fft_g1(v, c.dom128);
            cols[j] = v;
        }
    });
    for (auto &t : th) t.join();
    for (int j = 0; j < CELL; ++j)
        for (int i = 0; i < 128; ++i) c.fk_table[i * CELL + j] = cols[j][i].to_affine();   // transposeVectors
}

// ---- serialization.go:134-146 ----------------------------------------------------------
static int deserialize_scalars(Fr *out, const uint8_t *b, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!Fr::from_bytes_be(out[i], b + 32 * i)) return ST_NON_CANONICAL_SCALAR;
    return ST_OK;
}
static int deserialize_g1(G1Affine &out, const uint8_t *b) {   // serialization.go:108-115
    return g1_decompress(out, b, true);
}

// ---- fiatshamir.go:22-40 ---------------------------------------------------------------
static Fr compute_challenge(const uint8_t *blob, const uint8_t *commitment) {
    Sha256 h;
    h.update((const uint8_t *)"FSBLOBVERIFY_V1_", 16);
    uint8_t deg[16] = {0};
    uint64_t n = N_BLOB;
    for (int i = 0; i < 8; ++i) deg[8 + i] = (uint8_t)(n >> (56 - 8 * i));
    h.update(deg, 16);
    h.update(blob, N_BLOB * 32);
    h.update(commitment, 48);
    uint8_t d[32]; h.final(d);
    return Fr::from_bytes_be_reduce(d);
}

// ---- internal/domain/domain.go:163-235 -------------------------------------------------
static Fr eval_lagrange(const Ctx &c, const Fr *poly, const Fr &z, int64_t &index) {
    const Domain &d = c.dom4096;
    index = -1;
    for (uint64_t i = 0; i < d.n; ++i) if (d.roots[i] == z) { index = (int64_t)i; return poly[i]; }
    std::vector<Fr> den(d.n);
    for (uint64_t i = 0; i < d.n; ++i) den[i] = z - d.roots[i];
    batch_invert(den.data(), d.n);
    Fr res = Fr::zero();
    for (uint64_t i = 0; i < d.n; ++i) res = res + poly[i] * d.roots[i] * den[i];
    Fr t = (z.pow_u64(d.n) - Fr::one()) * d.n_inv;
    return t * res;
}

// ---- internal/kzg/kzg_prove.go:14-180 --------------------------------------------------
static void kzg_open(const Ctx &c, const Fr *poly, const Fr &z, G1Affine &proof, Fr &y) {
    const Domain &d = c.dom4096;
    int64_t idx;
    y = eval_lagrange(c, poly, z, idx);
    std::vector<Fr> q(d.n);
    if (idx < 0) {   // outside domain (:81-111)
        for (uint64_t i = 0; i < d.n; ++i) q[i] = d.roots[i] - z;
        batch_invert(q.data(), d.n);
        for (uint64_t i = 0; i < d.n; ++i) q[i] = q[i] * (poly[i] - y);
    } else {         // on domain (:118-180)
        Fr invz = d.inv_roots[idx];
        std::vector<Fr> den(d.n);
        for (uint64_t i = 0; i < d.n; ++i) den[i] = d.roots[i] - z;
        den[idx] = Fr::one();
        batch_invert(den.data(), d.n);
        q[idx] = Fr::zero();
        for (uint64_t j = 0; j < d.n; ++j) {
            if ((int64_t)j == idx) continue;
            Fr qj = (poly[j] - y) * den[j];
            q[j] = qj;
            q[idx] = q[idx] + qj.neg() * d.roots[j] * invz;
        }
    }
    proof = msm(c.g1_lagrange_brp.data(), q.data(), d.n).to_affine();
}

// ---- internal/kzg/kzg_verify.go:35-100 -------------------------------------------------
static int kzg_verify(const Ctx &c, const G1Affine &commitment, const G1Affine &proof, const Fr &z, const Fr &y) {
    G2Jac zg2 = g2_mul_fr(G2Jac::from_affine(c.g2_gen), z);
    G2Affine alpha_minus_z = G2Jac::from_affine(c.g2_alpha).add(zg2.neg()).to_affine();
    G1Jac yg1 = g1_mul_fr(G1Jac::from_affine(c.g1_gen), y);
    G1Affine f_minus_y = G1Jac::from_affine(commitment).add(yg1.neg()).to_affine();
    G1Affine P[2] = {f_minus_y, proof};
    G2Affine Q[2] = {c.g2_gen.neg(), alpha_minus_z};
    return pairing_check(P, Q, 2) ? ST_OK : ST_VERIFY_FAILED;
}

static Fr rlc_challenge(const void *seed, size_t n) {
    // The reference draws r from crypto/rand (internal/kzg/kzg_verify.go:136-137); any r works.
    Sha256 h; h.update((const uint8_t *)"ORACLE_RLC", 10); h.update((const uint8_t *)seed, n);
    uint8_t d[32]; h.final(d);
    return Fr::from_bytes_be_reduce(d);
}
static std::vector<Fr> compute_powers(const Fr &x, size_t n) {   // internal/utils/utils.go:22-34
    std::vector<Fr> p(n);
    Fr cur = Fr::one();
    for (size_t i = 0; i < n; ++i) { p[i] = cur; cur = cur * x; }
    return p;
}

// ---- internal/kzg/kzg_verify.go:111-231 ------------------------------------------------
static int kzg_batch_verify(const Ctx &c, const std::vector<G1Affine> &commitments, const std::vector<G1Affine> &proofs,
                            const std::vector<Fr> &zs, const std::vector<Fr> &ys) {
    size_t n = commitments.size();
    if (n == 0) return ST_OK;
    if (n == 1) return kzg_verify(c, commitments[0], proofs[0], zs[0], ys[0]);
    Fr r = rlc_challenge(ys.data(), n * sizeof(Fr));
    std::vector<Fr> rp = compute_powers(r, n);
    G1Jac folded_q = msm(proofs.data(), rp.data(), n);
    Fr folded_y = Fr::zero();
    for (size_t i = 0; i < n; ++i) folded_y = folded_y + ys[i] * rp[i];
    G1Jac folded_c = msm(commitments.data(), rp.data(), n);
    G1Jac yg = g1_mul_fr(G1Jac::from_affine(c.g1_gen), folded_y);
    folded_c = folded_c.add(yg.neg());
    std::vector<Fr> rz(n);
    for (size_t i = 0; i < n; ++i) rz[i] = rp[i] * zs[i];
    G1Jac folded_zq = msm(proofs.data(), rz.data(), n);
    folded_c = folded_c.add(folded_zq);
    G1Affine P[2] = {folded_c.to_affine(), folded_q.neg().to_affine()};
    G2Affine Q[2] = {c.g2_gen, c.g2_alpha};
    return pairing_check(P, Q, 2) ? ST_OK : ST_VERIFY_FAILED;
}

// ---- api_eip7594.go:29-39 : blob (Lagrange, brp order) -> monomial coefficients ---------
static void blob_to_coeffs(const Ctx &c, std::vector<Fr> &poly) {
    bit_reverse(poly.data(), poly.size());
    ifft_fr(poly.data(), poly.size(), c.dom4096);
}

// ---- fk20.go:58-74 + api_eip7594.go:79-91 ----------------------------------------------
static void compute_cells_from_coeffs(const Ctx &c, const std::vector<Fr> &coeffs, uint8_t *cells_out) {
    std::vector<Fr> ext(coeffs);
    ext.resize(N_EXT, Fr::zero());
    fft_fr(ext.data(), N_EXT, c.dom8192.gen);
    bit_reverse(ext.data(), N_EXT);
    for (int i = 0; i < N_EXT; ++i) ext[i].to_bytes_be(cells_out + 32 * i);   // partition(…,64) is contiguous
}

// ---- fk20.go:76-124 + toeplitz.go:17-29,95-125 -----------------------------------------
static void compute_proofs_from_coeffs(const Ctx &c, const std::vector<Fr> &coeffs, uint8_t *proofs_out) {
    std::vector<Fr> rev(coeffs.rbegin(), coeffs.rend());
    // toeplitzRows[i] = {rev[i + 64*m]}; column = [row[0],0,...]; circulant = col || [0,row[63],...,row[1]]
    std::vector<std::vector<Fr>> circ(CELL, std::vector<Fr>(128, Fr::zero()));
    for (int i = 0; i < CELL; ++i) {
        Fr row[CELL];
        for (int m = 0; m < CELL; ++m) row[m] = rev[i + CELL * m];
        circ[i][0] = row[0];
        for (int k = 1; k < CELL; ++k) circ[i][CELL + k] = row[CELL - k];
        fft_fr(circ[i].data(), 128, c.dom128.gen);
    }
    std::vector<G1Jac> res(128);
    for (int k = 0; k < 128; ++k) {           // 128 MSMs of 64 (toeplitz.go:113-119)
        Fr sc[CELL];
        for (int i = 0; i < CELL; ++i) sc[i] = circ[i][k];
        res[k] = msm(&c.fk_table[k * CELL], sc, CELL);
    }
    ifft_g1(res, c.dom128);
    res.resize(CELL);                          // first half (toeplitz.go:124)
    res.resize(128, G1Jac::infinity());        // pad with identity (fk20.go:82-85)
    fft_g1(res, c.dom128);
    bit_reverse(res.data(), 128);
    for (int k = 0; k < 128; ++k) g1_compress(proofs_out + 48 * k, res[k].to_affine());
}

// ---- erasure_code.go:75-90,110-164 -----------------------------------------------------
static void recover_coeffs(const Ctx &c, std::vector<Fr> &data /*8192, natural order, 0 where missing*/,
                           const std::vector<uint64_t> &missing /*brp_128 of absent ids*/, std::vector<Fr> &coeffs) {
    // vanishingPolyCoeff: prod (x - w128^m), O(k^2)
    std::vector<Fr> shortz(1, Fr::one());
    for (uint64_t m : missing) {
        Fr negx = c.dom128.roots[m].neg();
        std::vector<Fr> nz(shortz.size() + 1, Fr::zero());
        for (size_t i = 0; i < shortz.size(); ++i) { nz[i] = nz[i] + shortz[i] * negx; nz[i + 1] = nz[i + 1] + shortz[i]; }
        shortz.swap(nz);
    }
    std::vector<Fr> zx(N_EXT, Fr::zero());
    for (size_t i = 0; i < shortz.size(); ++i) zx[i * CELL] = shortz[i];
    std::vector<Fr> zeval(zx);
    fft_fr(zeval.data(), N_EXT, c.dom8192.gen);
    std::vector<Fr> ez(N_EXT);
    for (int i = 0; i < N_EXT; ++i) ez[i] = data[i] * zeval[i];
    ifft_fr(ez.data(), N_EXT, c.dom8192);
    coset_fft_fr(zx.data(), N_EXT, c.dom8192, c.rec_coset_gen);
    coset_fft_fr(ez.data(), N_EXT, c.dom8192, c.rec_coset_gen);
    batch_invert(zx.data(), N_EXT);
    for (int i = 0; i < N_EXT; ++i) ez[i] = ez[i] * zx[i];
    coset_ifft_fr(ez.data(), N_EXT, c.dom8192, c.rec_coset_gen_inv);
    coeffs.assign(ez.begin(), ez.begin() + N_BLOB);
}

// ---- api_eip7594.go:93-142 -------------------------------------------------------------
static int recover_polynomial_coeffs(const Ctx &c, const uint64_t *ids, size_t n_ids, const uint8_t *cells, size_t n_cells,
                                     std::vector<Fr> &coeffs) {
    if (n_ids != n_cells) return ST_LENGTH_MISMATCH;
    for (size_t i = 1; i < n_ids; ++i) if (ids[i] <= ids[i - 1]) return ST_CELL_IDS_NOT_ASCENDING;
    for (size_t i = 0; i < n_ids; ++i) if (ids[i] >= (uint64_t)N_CELLS) return ST_BAD_CELL_INDEX;
    if (n_ids < (size_t)(N_BLOB / CELL)) return ST_NOT_ENOUGH_CELLS;
    std::vector<uint64_t> missing;
    for (uint64_t id = 0; id < (uint64_t)N_CELLS; ++id)
        if (!std::binary_search(ids, ids + n_ids, id)) missing.push_back(brp_int(id, N_CELLS));
    std::vector<Fr> ext(N_EXT, Fr::zero());
    for (size_t i = 0; i < n_ids; ++i) {
        int st = deserialize_scalars(&ext[ids[i] * CELL], cells + 2048 * i, CELL);
        if (st) return st;
    }
    bit_reverse(ext.data(), N_EXT);
    recover_coeffs(c, ext, missing, coeffs);
    return ST_OK;
}

// ---- internal/kzg_multi/kzg_verify.go:16-105 -------------------------------------------
static int verify_multi_batch(const Ctx &c, const std::vector<G1Affine> &comms, const std::vector<uint64_t> &row_idx,
                              const uint64_t *cell_idx, const std::vector<G1Affine> &proofs, std::vector<std::vector<Fr>> &evals) {
    size_t n = row_idx.size();
    Fr r = rlc_challenge(evals[0].data(), CELL * sizeof(Fr));
    std::vector<Fr> rp = compute_powers(r, n);
    G1Jac sum_proofs = msm(proofs.data(), rp.data(), n);
    std::vector<Fr> weights(comms.size(), Fr::zero());
    for (size_t k = 0; k < n; ++k) weights[row_idx[k]] = weights[row_idx[k]] + rp[k];
    G1Jac sum_comms = msm(comms.data(), weights.data(), comms.size());
    std::vector<Fr> interp(CELL, Fr::zero());
    for (size_t k = 0; k < n; ++k) {
        std::vector<Fr> &e = evals[k];
        bit_reverse(e.data(), CELL);
        coset_ifft_fr(e.data(), CELL, c.dom64, c.coset_shift_inv[cell_idx[k]]);
        for (int i = 0; i < CELL; ++i) interp[i] = interp[i] + e[i] * rp[k];
    }
    G1Jac sum_interp = msm(c.g1_monomial.data(), interp.data(), CELL);   // srs.go:143-149
    std::vector<Fr> wr(n);
    for (size_t k = 0; k < n; ++k) wr[k] = c.coset_shift_pow64[cell_idx[k]] * rp[k];
    G1Jac weighted_proofs = msm(proofs.data(), wr.data(), n);
    G1Jac rl = sum_comms.add(sum_interp.neg()).add(weighted_proofs);
    G1Affine P[2] = {sum_proofs.to_affine(), rl.to_affine()};
    G2Affine Q[2] = {c.g2_s64, c.g2_gen.neg()};
    return pairing_check(P, Q, 2) ? ST_OK : ST_VERIFY_FAILED;
}

}  // namespace

// =========================================================================================
// C API (ctypes).  Return value: Status above.
// =========================================================================================
extern "C" {

// setup = 4096*48 monomial G1 | 4096*48 Lagrange G1 | n_g2*96 G2 (compressed), as packed by
// tests/golden/make_golden.py from trusted_setup.json.   (trusted_setup.go:90-134, api.go:90-149)
void *ko_ctx_new(const uint8_t *g1_monomial, const uint8_t *g1_lagrange, const uint8_t *g2, size_t n_g2) {
    init_all();
    if (n_g2 < 65) return nullptr;
    Ctx *c = new Ctx();
    c->g1_monomial.resize(N_BLOB);
    c->g1_lagrange_brp.resize(N_BLOB);
    for (int i = 0; i < N_BLOB; ++i) {
        if (g1_decompress(c->g1_monomial[i], g1_monomial + 48 * i, false)) { delete c; return nullptr; }
        if (g1_decompress(c->g1_lagrange_brp[i], g1_lagrange + 48 * i, false)) { delete c; return nullptr; }
    }
    bit_reverse(c->g1_lagrange_brp.data(), N_BLOB);
    if (g2_decompress(c->g2_gen, g2) || g2_decompress(c->g2_alpha, g2 + 96) || g2_decompress(c->g2_s64, g2 + 96 * 64)) { delete c; return nullptr; }
    c->g1_gen = g1_generator();
    c->dom4096.init(N_BLOB); c->dom4096.reverse_roots();
    c->dom8192.init(N_EXT);
    c->dom128.init(128);
    c->dom64.init(64);
    // srs.go:60-103
    std::vector<Fr> ext_roots = c->dom8192.roots;
    bit_reverse(ext_roots.data(), N_EXT);
    c->coset_shift.resize(N_CELLS); c->coset_shift_inv.resize(N_CELLS); c->coset_shift_pow64.resize(N_CELLS);
    for (int k = 0; k < N_CELLS; ++k) {
        c->coset_shift[k] = ext_roots[k * CELL];
        c->coset_shift_inv[k] = c->coset_shift[k].inv();
        c->coset_shift_pow64[k] = c->coset_shift[k].pow_u64(CELL);
    }
    c->rec_coset_gen = Fr::from_u64(7);            // erasure_code.go:58
    c->rec_coset_gen_inv = c->rec_coset_gen.inv();
    build_fk20_table(*c);
    return c;
}
void ko_ctx_free(void *ctx) { delete (Ctx *)ctx; }

// prove.go:13-34
int ko_blob_to_kzg_commitment(void *ctx, const uint8_t *blob, uint8_t *out48) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> poly(N_BLOB);
    int st = deserialize_scalars(poly.data(), blob, N_BLOB);
    if (st) return st;
    g1_compress(out48, msm(c.g1_lagrange_brp.data(), poly.data(), N_BLOB).to_affine());
    return ST_OK;
}
// prove.go:85-111
int ko_compute_kzg_proof(void *ctx, const uint8_t *blob, const uint8_t *z32, uint8_t *proof48, uint8_t *y32) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> poly(N_BLOB);
    int st = deserialize_scalars(poly.data(), blob, N_BLOB);
    if (st) return st;
    Fr z, y;
    if (!Fr::from_bytes_be(z, z32)) return ST_NON_CANONICAL_SCALAR;
    G1Affine proof;
    kzg_open(c, poly.data(), z, proof, y);
    g1_compress(proof48, proof);
    y.to_bytes_be(y32);
    return ST_OK;
}
// prove.go:46-77
int ko_compute_blob_kzg_proof(void *ctx, const uint8_t *blob, const uint8_t *commitment48, uint8_t *proof48) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> poly(N_BLOB);
    int st = deserialize_scalars(poly.data(), blob, N_BLOB);
    if (st) return st;
    G1Affine cm;
    if ((st = deserialize_g1(cm, commitment48))) return st;
    Fr z = compute_challenge(blob, commitment48), y;
    G1Affine proof;
    kzg_open(c, poly.data(), z, proof, y);
    g1_compress(proof48, proof);
    return ST_OK;
}
// verify.go:12-41
int ko_verify_kzg_proof(void *ctx, const uint8_t *commitment48, const uint8_t *z32, const uint8_t *y32, const uint8_t *proof48) {
    Ctx &c = *(Ctx *)ctx;
    Fr z, y; G1Affine cm, pf; int st;
    if (!Fr::from_bytes_be(y, y32)) return ST_NON_CANONICAL_SCALAR;
    if (!Fr::from_bytes_be(z, z32)) return ST_NON_CANONICAL_SCALAR;
    if ((st = deserialize_g1(cm, commitment48))) return st;
    if ((st = deserialize_g1(pf, proof48))) return st;
    return kzg_verify(c, cm, pf, z, y);
}
// verify.go:48-82
int ko_verify_blob_kzg_proof(void *ctx, const uint8_t *blob, const uint8_t *commitment48, const uint8_t *proof48) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> poly(N_BLOB);
    int st = deserialize_scalars(poly.data(), blob, N_BLOB);
    if (st) return st;
    G1Affine cm, pf;
    if ((st = deserialize_g1(cm, commitment48))) return st;
    if ((st = deserialize_g1(pf, proof48))) return st;
    Fr z = compute_challenge(blob, commitment48);
    int64_t idx;
    Fr y = eval_lagrange(c, poly.data(), z, idx);
    return kzg_verify(c, cm, pf, z, y);
}
// verify.go:88-145  (sequential per-blob loop, then one RLC check)
int ko_verify_blob_kzg_proof_batch(void *ctx, const uint8_t *blobs, size_t n_blobs, const uint8_t *commitments, size_t n_comm,
                                   const uint8_t *proofs, size_t n_proofs) {
    Ctx &c = *(Ctx *)ctx;
    if (n_blobs != n_comm || n_blobs != n_proofs) return ST_LENGTH_MISMATCH;
    size_t n = n_blobs;
    std::vector<G1Affine> cms(n), pfs(n);
    std::vector<Fr> zs(n), ys(n), poly(N_BLOB);
    for (size_t i = 0; i < n; ++i) {
        int st;
        if ((st = deserialize_g1(cms[i], commitments + 48 * i))) return st;
        if ((st = deserialize_g1(pfs[i], proofs + 48 * i))) return st;
        if ((st = deserialize_scalars(poly.data(), blobs + (size_t)N_BLOB * 32 * i, N_BLOB))) return st;
        zs[i] = compute_challenge(blobs + (size_t)N_BLOB * 32 * i, commitments + 48 * i);
        int64_t idx;
        ys[i] = eval_lagrange(c, poly.data(), zs[i], idx);
    }
    return kzg_batch_verify(c, cms, pfs, zs, ys);
}
// api_eip7594.go:12-26
int ko_compute_cells(void *ctx, const uint8_t *blob, uint8_t *cells_out) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> poly(N_BLOB);
    int st = deserialize_scalars(poly.data(), blob, N_BLOB);
    if (st) return st;
    blob_to_coeffs(c, poly);
    compute_cells_from_coeffs(c, poly, cells_out);
    return ST_OK;
}
// api_eip7594.go:28-52
int ko_compute_cells_and_kzg_proofs(void *ctx, const uint8_t *blob, uint8_t *cells_out, uint8_t *proofs_out) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> poly(N_BLOB);
    int st = deserialize_scalars(poly.data(), blob, N_BLOB);
    if (st) return st;
    blob_to_coeffs(c, poly);
    compute_cells_from_coeffs(c, poly, cells_out);
    compute_proofs_from_coeffs(c, poly, proofs_out);
    return ST_OK;
}
// api_eip.go:8-15
int ko_recover_cells(void *ctx, const uint64_t *ids, size_t n_ids, const uint8_t *cells, size_t n_cells, uint8_t *cells_out) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> coeffs;
    int st = recover_polynomial_coeffs(c, ids, n_ids, cells, n_cells, coeffs);
    if (st) return st;
    compute_cells_from_coeffs(c, coeffs, cells_out);
    return ST_OK;
}
// api_eip7594.go:144-161
int ko_recover_cells_and_kzg_proofs(void *ctx, const uint64_t *ids, size_t n_ids, const uint8_t *cells, size_t n_cells,
                                    uint8_t *cells_out, uint8_t *proofs_out) {
    Ctx &c = *(Ctx *)ctx;
    std::vector<Fr> coeffs;
    int st = recover_polynomial_coeffs(c, ids, n_ids, cells, n_cells, coeffs);
    if (st) return st;
    compute_cells_from_coeffs(c, coeffs, cells_out);
    compute_proofs_from_coeffs(c, coeffs, proofs_out);
    return ST_OK;
}
// api_eip7594.go:163-265
int ko_verify_cell_kzg_proof_batch(void *ctx, const uint8_t *commitments, size_t n_comm, const uint64_t *cell_idx, size_t n_idx,
                                   const uint8_t *cells, size_t n_cells, const uint8_t *proofs, size_t n_proofs) {
    Ctx &c = *(Ctx *)ctx;
    // deduplicateKZGCommitments (:238-265): first-seen order on raw bytes
    std::vector<const uint8_t *> uniq;
    std::vector<uint64_t> row_idx(n_comm);
    for (size_t i = 0; i < n_comm; ++i) {
        size_t j = 0;
        for (; j < uniq.size(); ++j) if (!memcmp(uniq[j], commitments + 48 * i, 48)) break;
        if (j == uniq.size()) uniq.push_back(commitments + 48 * i);
        row_idx[i] = j;
    }
    size_t n = n_comm;
    if (n != n_idx || n != n_cells || n != n_proofs) return ST_LENGTH_MISMATCH;
    if (n == 0) return ST_OK;
    for (size_t i = 0; i < n; ++i) if (cell_idx[i] >= (uint64_t)N_CELLS) return ST_BAD_CELL_INDEX;
    std::vector<G1Affine> cms(uniq.size()), pfs(n);
    int st;
    for (size_t i = 0; i < uniq.size(); ++i) if ((st = deserialize_g1(cms[i], uniq[i]))) return st;
    for (size_t i = 0; i < n; ++i) if ((st = deserialize_g1(pfs[i], proofs + 48 * i))) return st;
    std::vector<std::vector<Fr>> evals(n, std::vector<Fr>(CELL));
    for (size_t i = 0; i < n; ++i) if ((st = deserialize_scalars(evals[i].data(), cells + 2048 * i, CELL))) return st;
    return verify_multi_batch(c, cms, row_idx, cell_idx, pfs, evals);
}

// bench_test.go:17-46: GetRandBlob(seed): scalar j = SHA-256(be_int64(seed + 32 j)) mod r
void ko_rand_blob(int64_t seed, uint8_t *blob) {
    init_all();
    for (int j = 0; j < N_BLOB; ++j) {
        int64_t s = seed + 32 * (int64_t)j;
        uint8_t b[8]; for (int i = 0; i < 8; ++i) b[i] = (uint8_t)((uint64_t)s >> (56 - 8 * i));
        Sha256 h; h.update(b, 8);
        uint8_t d[32]; h.final(d);
        Fr::from_bytes_be_reduce(d).to_bytes_be(blob + 32 * j);
    }
}
void ko_compute_challenge(const uint8_t *blob, const uint8_t *commitment48, uint8_t *out32) {
    init_all();
    compute_challenge(blob, commitment48).to_bytes_be(out32);
}
void ko_sha256(const uint8_t *p, size_t n, uint8_t *out32) { Sha256 h; h.update(p, n); h.final(out32); }

// CPU baseline helper: run fn over n independent blobs on `threads` host threads (blob-parallel).
// kind: 0 = blob_to_kzg_commitment, 1 = compute_cells_and_kzg_proofs
int ko_parallel_blobs(void *ctx, int kind, const uint8_t *blobs, size_t n, int threads, uint8_t *out_a, uint8_t *out_b) {
    std::atomic<size_t> next(0);
    std::atomic<int> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&] {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            int st = kind == 0 ? ko_blob_to_kzg_commitment(ctx, blobs + i * 131072, out_a + 48 * i)
                               : ko_compute_cells_and_kzg_proofs(ctx, blobs + i * 131072, out_a + i * 262144, out_b + i * 6144);
            if (st) bad = st;
        }
    });
    for (auto &t : th) t.join();
    return bad;
}

// low-level hooks for unit tests of the CUDA field/curve kernels
int ko_g1_decompress(const uint8_t *in48, uint8_t *xy96, int subgroup_check) {
    init_all();
    G1Affine a; int st = g1_decompress(a, in48, subgroup_check != 0);
    if (st) return st;
    if (a.inf) { memset(xy96, 0, 96); return 0; }
    a.x.to_bytes_be(xy96); a.y.to_bytes_be(xy96 + 48);
    return 0;
}
// G2Affine.SetBytes semantics (gnark-crypto, as used by trusted_setup.go:45-83): decode and, if asked, test [r]Q == O
int ko_g2_decompress(const uint8_t *in96, int subgroup_check) {
    init_all();
    if ((in96[0] >> 5) == 6) {          // infinity: every other bit must be clear
        if (in96[0] & 0x1f) return DEC_BAD_ENCODING;
        for (int i = 1; i < 96; ++i) if (in96[i]) return DEC_BAD_ENCODING;
        return DEC_OK;
    }
    G2Affine q; int st = g2_decompress(q, in96);
    if (st) return st;
    if (subgroup_check) {
        if (!G2Jac::from_affine(q).mul(FR_MOD, 4).is_inf()) return DEC_NOT_IN_SUBGROUP;      // the definition, as g1_in_subgroup
    }
    return DEC_OK;
}
// sum_i scalars[i] * points[i]  (compressed points in, compressed point out, BE scalars)
int ko_g1_msm(const uint8_t *points48, const uint8_t *scalars32, size_t n, uint8_t *out48) {
    init_all();
    std::vector<G1Affine> p(n); std::vector<Fr> s(n);
    for (size_t i = 0; i < n; ++i) {
        if (g1_decompress(p[i], points48 + 48 * i, false)) return ST_BAD_G1_ENCODING;
        s[i] = Fr::from_bytes_be_reduce(scalars32 + 32 * i);
    }
    g1_compress(out48, msm(p.data(), s.data(), n).to_affine());
    return 0;
}
// out48[i] = [scalars[i]] P and out96[i] = [scalars[i]] Q for ONE base point each (compressed in/out, BE scalars):
// what srs_insecure.go:27-58 and kzg_multi/srs.go:113-141 do with a known secret to build a test SRS
int ko_g1_mul_many(const uint8_t *p48, const uint8_t *scalars32, size_t n, uint8_t *out48) {
    init_all();
    G1Affine p;
    if (g1_decompress(p, p48, false)) return ST_BAD_G1_ENCODING;
    G1Jac pj = G1Jac::from_affine(p);
    for (size_t i = 0; i < n; ++i) g1_compress(out48 + 48 * i, g1_mul_fr(pj, Fr::from_bytes_be_reduce(scalars32 + 32 * i)).to_affine());
    return 0;
}
int ko_g2_mul_many(const uint8_t *q96, const uint8_t *scalars32, size_t n, uint8_t *out96) {
    init_all();
    G2Affine q;
    if (g2_decompress(q, q96)) return ST_BAD_G1_ENCODING;
    G2Jac qj = G2Jac::from_affine(q);
    for (size_t i = 0; i < n; ++i) {
        G2Affine a = g2_mul_fr(qj, Fr::from_bytes_be_reduce(scalars32 + 32 * i)).to_affine();
        uint8_t *b = out96 + 96 * i;
        if (a.inf) { memset(b, 0, 96); b[0] = 0xc0; continue; }
        a.x.c1.to_bytes_be(b); a.x.c0.to_bytes_be(b + 48);
        b[0] |= 0x80;
        if (a.y.lex_largest()) b[0] |= 0x20;
    }
    return 0;
}
// pairing product check over n (G1 compressed, G2 compressed) pairs: 1 if product == 1
int ko_pairing_check(const uint8_t *g1s, const uint8_t *g2s, size_t n) {
    init_all();
    std::vector<G1Affine> p(n); std::vector<G2Affine> q(n);
    for (size_t i = 0; i < n; ++i) {
        if (g1_decompress(p[i], g1s + 48 * i, false)) return -1;
        if (g2_decompress(q[i], g2s + 96 * i)) return -1;
    }
    return pairing_check(p.data(), q.data(), (int)n) ? 1 : 0;
}

}  // extern "C"

// debug: gamma = (1+u)^((p-1)/6) and the first/last Miller line coefficients of a G2 point,
// as 4 x (c0,c1) x 48 big-endian bytes, for comparison with the device precomputation
extern "C" int ko_dbg_dump_pairing(const uint8_t *g2_96, uint8_t *out) {
    init_all();
    G2Affine Q;
    if (g2_decompress(Q, g2_96)) return -1;
    Fp2 vals[4];
    vals[0] = frob_tab().g1[1];
    G2Affine T = Q;
    int li = 0;
    for (int i = 62; i >= 0; --i) {
        Fp2 lambda = (T.x.sqr().dbl() + T.x.sqr()) * T.y.dbl().inv();
        Fp2 A = lambda * T.x - T.y, B = lambda.neg();
        if (li == 0) { vals[1] = A; vals[2] = B; }
        if (li == 67) vals[3] = A;
        ++li;
        Fp2 x3 = lambda.sqr() - T.x.dbl();
        Fp2 y3 = lambda * (T.x - x3) - T.y;
        T = {x3, y3, false};
        if ((BLS_X_ABS >> i) & 1) {
            Fp2 lam2 = (Q.y - T.y) * (Q.x - T.x).inv();
            Fp2 A2 = lam2 * T.x - T.y;
            if (li == 67) vals[3] = A2;
            ++li;
            Fp2 x4 = lam2.sqr() - T.x - Q.x;
            Fp2 y4 = lam2 * (T.x - x4) - T.y;
            T = {x4, y4, false};
        }
    }
    for (int s = 0; s < 4; ++s) { vals[s].c0.to_bytes_be(out + s * 96); vals[s].c1.to_bytes_be(out + s * 96 + 48); }
    return li;
}

// debug: Granger-Scott cyclotomic squaring vs generic squaring on an element of the cyclotomic
// subgroup (used to validate the formula the CUDA pairing uses).  Returns 1 if equal.
extern "C" int ko_dbg_cyclotomic_sqr_check(uint64_t seed) {
    init_all();
    auto rnd = [&]() { seed = seed * 6364136223846793005ULL + 1442695040888963407ULL; return Fp::from_u64(seed >> 11) * Fp::from_u64(seed | 1); };
    Fp12 a;
    Fp2 *c[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
    for (auto p : c) { p->c0 = rnd(); p->c1 = rnd(); }
    // easy part -> cyclotomic subgroup
    Fp12 f = a.conj() * a.inv();
    f = frobenius(frobenius(f)) * f;
    Fp12 ref = f.sqr();
    // Fp4 view: A0 = g0 + h1 s, A1 = h0 + g2 s, A2 = g1 + h2 s   (s = v w, s^2 = xi)
    Fp2 x0 = f.c0.c0, y0 = f.c1.c1, x1 = f.c1.c0, y1 = f.c0.c2, x2 = f.c0.c1, y2 = f.c1.c2;
    auto sq4 = [](const Fp2 &x, const Fp2 &y, Fp2 &rx, Fp2 &ry) {   // (x + y s)^2
        Fp2 xx = x.sqr(), yy = y.sqr();
        rx = xx + yy.mul_xi();
        ry = (x + y).sqr() - xx - yy;
    };
    Fp2 t0x, t0y, t1x, t1y, t2x, t2y;
    sq4(x0, y0, t0x, t0y); sq4(x1, y1, t1x, t1y); sq4(x2, y2, t2x, t2y);
    auto three = [](const Fp2 &z) { return z.dbl() + z; };
    // A0' = 3 A0^2 - 2 conj(A0);  A1' = 3 s A2^2 + 2 conj(A1);  A2' = 3 A1^2 - 2 conj(A2)
    Fp2 n0x = three(t0x) - x0.dbl(), n0y = three(t0y) + y0.dbl();
    Fp2 sx = t2y.mul_xi(), sy = t2x;                 // s * (t2x + t2y s) = xi t2y + t2x s
    Fp2 n1x = three(sx) + x1.dbl(), n1y = three(sy) - y1.dbl();
    Fp2 n2x = three(t1x) - x2.dbl(), n2y = three(t1y) + y2.dbl();
    Fp12 r;
    r.c0.c0 = n0x; r.c1.c1 = n0y; r.c1.c0 = n1x; r.c0.c2 = n1y; r.c0.c1 = n2x; r.c1.c2 = n2y;
    return r == ref ? 1 : 0;
}
