// libkzgb200.so -- trusted-setup ingestion: the reference's JSON format
// (trusted_setup.go:23-27) and CheckTrustedSetupIsWellFormed (trusted_setup.go:45-83) as batched kernels.
#include "ctx.cuh"
#include "g2.cuh"

#define CUS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(KZGB200_ERR_CUDA, #call, e_); } while (0)

namespace {
// ---- a parser for exactly the shape of JSONTrustedSetup: an object whose keys g1_monomial, g1_lagrange,
// g2_monomial hold arrays of hex strings (0x prefix optional, trusted_setup.go:194-200 trims it) --------
struct Cursor { const char *p, *end; };
static void skip_ws(Cursor &c) { while (c.p < c.end && (*c.p == ' ' || *c.p == '\n' || *c.p == '\r' || *c.p == '\t')) ++c.p; }
static int hexval(char ch) { return ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : -1; }
// parses "…" at the cursor into bytes; returns false on malformed string / odd or non-hex digits
static bool parse_hex_string(Cursor &c, std::vector<uint8_t> &out) {
    skip_ws(c);
    if (c.p >= c.end || *c.p != '"') return false;
    const char *s = ++c.p;
    while (c.p < c.end && *c.p != '"') { if (*c.p == '\\') return false; ++c.p; }
    if (c.p >= c.end) return false;
    const char *e = c.p++;
    if (e - s >= 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
    if ((e - s) & 1) return false;
    out.clear();
    for (; s < e; s += 2) {
        int a = hexval(s[0]), b = hexval(s[1]);
        if (a < 0 || b < 0) return false;
        out.push_back((uint8_t)(a * 16 + b));
    }
    return true;
}
// array of hex strings, each exactly `width` bytes, REPLACING the contents of out: a key that appears twice keeps
// its last value, as Go's encoding/json does for JSONTrustedSetup (trusted_setup.go:23-27)
static bool parse_point_array(Cursor &c, size_t width, std::vector<uint8_t> &out, size_t &count) {
    skip_ws(c);
    if (c.p >= c.end || *c.p != '[') return false;
    ++c.p; count = 0; out.clear();
    skip_ws(c);
    if (c.p < c.end && *c.p == ']') { ++c.p; return true; }
    std::vector<uint8_t> one;
    for (;;) {
        if (!parse_hex_string(c, one) || one.size() != width) return false;
        out.insert(out.end(), one.begin(), one.end()); ++count;
        skip_ws(c);
        if (c.p >= c.end) return false;
        if (*c.p == ',') { ++c.p; continue; }
        if (*c.p == ']') { ++c.p; return true; }
        return false;
    }
}
static bool skip_value(Cursor &c) {      // unknown keys: skip a string / array / object / scalar
    skip_ws(c);
    if (c.p >= c.end) return false;
    if (*c.p == '"') { ++c.p; while (c.p < c.end && *c.p != '"') { if (*c.p == '\\') ++c.p; ++c.p; } if (c.p >= c.end) return false; ++c.p; return true; }
    if (*c.p == '[' || *c.p == '{') {
        int depth = 0; bool in_str = false;
        for (; c.p < c.end; ++c.p) {
            char ch = *c.p;
            if (in_str) { if (ch == '\\') ++c.p; else if (ch == '"') in_str = false; continue; }
            if (ch == '"') in_str = true;
            else if (ch == '[' || ch == '{') ++depth;
            else if (ch == ']' || ch == '}') { if (--depth == 0) { ++c.p; return true; } }
        }
        return false;
    }
    while (c.p < c.end && *c.p != ',' && *c.p != '}' && *c.p != ']') ++c.p;
    return true;
}
}  // namespace

namespace kzg {
// canonical check only: status[i / per_item] = NON_CANONICAL if scalar i >= r
static __global__ void k_blob_to_scalars_any(const uint8_t *__restrict__ in, int32_t *__restrict__ status, size_t n_scalars, size_t per_item) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_scalars) return;
    uint32_t l[8];
    load_be32(l, in + i * 32);
    if (!fr_is_canonical(l)) atomicMax(&status[i / per_item], (int32_t)ST_NON_CANONICAL_SCALAR);
}
}  // namespace kzg

extern "C" int kzgb200_parse_trusted_setup_json(const char *json, size_t len, uint8_t *g1_monomial, uint8_t *g1_lagrange,
                                                 uint8_t *g2_monomial, size_t g2_capacity, size_t *n_g2) {
    if (!json || !g1_monomial || !g1_lagrange || !g2_monomial || !n_g2) return set_err(KZGB200_ERR_ARGS, "null argument");
    Cursor c{json, json + len};
    std::vector<uint8_t> m, l, g2;
    size_t nm = 0, nl = 0, ng = 0;
    bool have_m = false, have_l = false, have_g = false;
    skip_ws(c);
    if (c.p >= c.end || *c.p != '{') return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: expected an object");
    ++c.p;
    for (;;) {
        skip_ws(c);
        if (c.p < c.end && *c.p == '}') break;
        if (c.p >= c.end || *c.p != '"') return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: expected a key");
        const char *ks = ++c.p;
        while (c.p < c.end && *c.p != '"') ++c.p;
        if (c.p >= c.end) return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: unterminated key");
        std::string key(ks, c.p++);
        skip_ws(c);
        if (c.p >= c.end || *c.p != ':') return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: expected ':'");
        ++c.p;
        bool ok;
        if (key == "g1_monomial") { ok = parse_point_array(c, 48, m, nm); have_m = true; }
        else if (key == "g1_lagrange") { ok = parse_point_array(c, 48, l, nl); have_l = true; }
        else if (key == "g2_monomial") { ok = parse_point_array(c, 96, g2, ng); have_g = true; }
        else ok = skip_value(c);
        if (!ok) return set_err(KZGB200_ERR_SETUP, ("trusted setup JSON: malformed value of \"" + key + "\" (hex strings of 48 / 96 bytes expected)").c_str());
        skip_ws(c);
        if (c.p < c.end && *c.p == ',') { ++c.p; continue; }
        if (c.p < c.end && *c.p == '}') break;
        return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: expected ',' or '}'");
    }
    if (!have_m || !have_l || !have_g) return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: g1_monomial, g1_lagrange and g2_monomial are required");
    if (nm != (size_t)N_BLOB || nl != (size_t)N_BLOB) return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: need exactly 4096 G1 points in each basis (trusted_setup.go:23-27)");
    if (ng > g2_capacity) return set_err(KZGB200_ERR_ARGS, "g2_monomial buffer too small");
    if (m.size() != (size_t)N_BLOB * 48 || l.size() != (size_t)N_BLOB * 48 || g2.size() != ng * 96) return set_err(KZGB200_ERR_SETUP, "trusted setup JSON: inconsistent point count");
    memcpy(g1_monomial, m.data(), (size_t)N_BLOB * 48); memcpy(g1_lagrange, l.data(), (size_t)N_BLOB * 48);
    if (ng) memcpy(g2_monomial, g2.data(), ng * 96);
    *n_g2 = ng;
    return KZGB200_OK;
}

extern "C" int kzgb200_ctx_new_from_json(const char *json, size_t len, const kzgb200_opts *opts, kzgb200_ctx **out) {
    if (!out) return set_err(KZGB200_ERR_ARGS, "null argument");
    *out = nullptr;
    std::vector<uint8_t> m(N_BLOB * 48), l(N_BLOB * 48), g2(4096 * 96);
    size_t ng = 0;
    int rc = kzgb200_parse_trusted_setup_json(json, len, m.data(), l.data(), g2.data(), 4096, &ng);
    if (rc) return rc;
    return kzgb200_ctx_new(m.data(), l.data(), g2.data(), ng, opts, out);
}

// CheckTrustedSetupIsWellFormed (trusted_setup.go:45-83): Lagrange G1, then monomial G1, then G2; every point must
// decode and lie in the prime-order subgroup.  *result = OK or the status of the first failing point in that order.
extern "C" int kzgb200_check_trusted_setup(int device, const uint8_t *g1_lagrange, size_t n_lagrange, const uint8_t *g1_monomial, size_t n_monomial,
                                           const uint8_t *g2_monomial, size_t n_g2, int32_t *result, size_t *bad_index) {
    if (!result || (n_lagrange && !g1_lagrange) || (n_monomial && !g1_monomial) || (n_g2 && !g2_monomial)) return set_err(KZGB200_ERR_ARGS, "null argument");
    CUS(cudaSetDevice(device));
    const size_t n1 = n_lagrange + n_monomial, n = n1 + n_g2;
    *result = KZGB200_OK;
    if (bad_index) *bad_index = 0;
    if (!n) return KZGB200_OK;
    uint8_t *d_in = nullptr; int32_t *d_st = nullptr;
    CUS(cudaMalloc(&d_in, n1 * 48 + n_g2 * 96 + 16)); CUS(cudaMalloc(&d_st, n * sizeof(int32_t)));
    CUS(cudaMemset(d_st, 0, n * sizeof(int32_t)));
    if (n_lagrange) CUS(cudaMemcpy(d_in, g1_lagrange, n_lagrange * 48, cudaMemcpyHostToDevice));
    if (n_monomial) CUS(cudaMemcpy(d_in + n_lagrange * 48, g1_monomial, n_monomial * 48, cudaMemcpyHostToDevice));
    if (n_g2) CUS(cudaMemcpy(d_in + n1 * 48, g2_monomial, n_g2 * 96, cudaMemcpyHostToDevice));
    if (n1) k_g1_check<MulCall, 4><<<(unsigned)((n1 + 63) / 64), 64>>>(d_in, nullptr, d_st, n1, 1, 1);
    if (n_g2) k_g2_check<<<(unsigned)((n_g2 + 31) / 32), 32>>>(d_in + n1 * 48, d_st + n1, n_g2, 1);
    CUS(cudaGetLastError());
    std::vector<int32_t> st(n);
    CUS(cudaMemcpy(st.data(), d_st, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_st);
    for (size_t i = 0; i < n; ++i) if (st[i] != KZGB200_OK) { *result = st[i]; if (bad_index) *bad_index = i; break; }
    return KZGB200_OK;
}

// ---- codec validation entry points ---------------------------------------
// DeserializeKZGCommitment / DeserializeKZGProof (serialization.go:108-131): status[i] of n compressed G1 points
// (decode + subgroup check, exactly what every API call applies to its G1 inputs).
extern "C" int lane_check_g1_points(kzg_lane *c, const uint8_t *points48, size_t n, int32_t *status) {
    if (!c || (n && (!points48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    CUS(cudaSetDevice(c->device));
    if (!n) return KZGB200_OK;
    const bool st_dev = is_device_ptr(status);
    int rc;
    const void *d_in = nullptr;
    if ((rc = stage_in(c, points48, n * 48, c->in_small, &d_in))) return rc;
    if (!st_dev && (rc = c->status.ensure(n * sizeof(int32_t)))) return rc;
    int32_t *d_st = st_dev ? status : (int32_t *)c->status.p;
    CUS(cudaMemsetAsync(d_st, 0, n * sizeof(int32_t), c->stream));
    k_g1_check<MulCall, 4><<<(unsigned)((n + 63) / 64), 64, 0, c->stream>>>((const uint8_t *)d_in, nullptr, d_st, n, 1, 1);
    c->launches += 1;
    CUS(cudaGetLastError());
    if (!st_dev) CUS(cudaMemcpyAsync(status, d_st, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUS(cudaStreamSynchronize(c->stream));
    return KZGB200_OK;
}
// DeserializeScalar / DeserializeBlob (serialization.go:134-159): status[i] = OK or NON_CANONICAL_SCALAR for n items of
// `scalars_per_item` big-endian 32-byte scalars each (1 for a Scalar, 4096 for a Blob, 64 for a Cell)
extern "C" int lane_check_scalars(kzg_lane *c, const uint8_t *scalars32, size_t n_items, size_t scalars_per_item, int32_t *status) {
    if (!c || !scalars_per_item || (n_items && (!scalars32 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    CUS(cudaSetDevice(c->device));
    if (!n_items) return KZGB200_OK;
    const bool st_dev = is_device_ptr(status);
    const size_t ns = n_items * scalars_per_item;
    int rc;
    const void *d_in = nullptr;
    if ((rc = stage_in(c, scalars32, ns * 32, c->in_bytes, &d_in))) return rc;
    if (!st_dev && (rc = c->status.ensure(n_items * sizeof(int32_t)))) return rc;
    int32_t *d_st = st_dev ? status : (int32_t *)c->status.p;
    CUS(cudaMemsetAsync(d_st, 0, n_items * sizeof(int32_t), c->stream));
    k_blob_to_scalars_any<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>((const uint8_t *)d_in, d_st, ns, scalars_per_item);
    c->launches += 1;
    CUS(cudaGetLastError());
    if (!st_dev) CUS(cudaMemcpyAsync(status, d_st, n_items * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUS(cudaStreamSynchronize(c->stream));
    return KZGB200_OK;
}

// ---- unit-test hook (include/kzgb200_debug.h): self-consistency of the G2 group law on one point ----
namespace kzg {
__device__ __forceinline__ bool g2j_on_curve(const G2J &p) {      // Y^2 == X^3 + 4(1+u) Z^6
    Fp four = Fp::dbl(Fp::dbl(Fp::one()));
    Fp2 b2; b2.c0 = four; b2.c1 = four;
    Fp2 z2 = fp2_sqr(p.Z), z6 = fp2_mul(fp2_sqr(z2), z2);
    return fp2_eq(fp2_sqr(p.Y), fp2_add(fp2_mul(fp2_sqr(p.X), p.X), fp2_mul(b2, z6)));
}
static __global__ void k_dbg_g2_selftest(const uint8_t *in96, int *mask) {
    if (threadIdx.x || blockIdx.x) return;
    G2Dec d = g2_decompress(in96);
    int m = 0;
    if (d.st == ST_OK && !d.a.inf) {
        m |= 1;
        G2J q; q.X = d.a.x; q.Y = d.a.y; q.Z = fp2_one();
        if (g2j_on_curve(q)) m |= 2;
        G2J q2 = g2j_dbl(q);
        if (g2j_on_curve(q2)) m |= 4;
        G2J q3 = g2j_add_affine(q2, d.a.x, d.a.y);
        if (g2j_on_curve(q3)) m |= 8;
        G2J q4a = g2j_dbl(q2), q4b = g2j_add_affine(q3, d.a.x, d.a.y);      // 4Q two ways: X_a Z_b^2 == X_b Z_a^2, Y_a Z_b^3 == Y_b Z_a^3
        Fp2 za2 = fp2_sqr(q4a.Z), zb2 = fp2_sqr(q4b.Z);
        if (fp2_eq(fp2_mul(q4a.X, zb2), fp2_mul(q4b.X, za2)) && fp2_eq(fp2_mul(q4a.Y, fp2_mul(zb2, q4b.Z)), fp2_mul(q4b.Y, fp2_mul(za2, q4a.Z)))) m |= 16;
        G2J same = g2j_add_affine(q, d.a.x, d.a.y);                             // Q + Q through the addition's doubling branch
        Fp2 zs2 = fp2_sqr(same.Z), z22 = fp2_sqr(q2.Z);
        if (fp2_eq(fp2_mul(same.X, z22), fp2_mul(q2.X, zs2))) m |= 32;
        G2J neg = g2j_add_affine(q, d.a.x, fp2_neg(d.a.y));                     // Q + (-Q) = O
        if (fp2_is_zero(neg.Z)) m |= 64;
        if (g2_in_subgroup(d.a)) m |= 128;
        {   // the same ladder inline
            G2J acc = q;
#pragma unroll 1
            for (int bit = 253; bit >= 0; --bit) {
                acc = g2j_dbl(acc);
                if ((FR_MOD[bit >> 5] >> (bit & 31)) & 1) acc = g2j_add_affine(acc, d.a.x, d.a.y);
            }
            if (fp2_is_zero(acc.Z)) m |= 256;
        }
    }
    *mask = m;
}
}  // namespace kzg
extern "C" int kzgb200_dbg_g2_selftest(const uint8_t *in96, int *mask) {
    uint8_t *d; int *dm;
    CUS(cudaMalloc(&d, 96)); CUS(cudaMalloc(&dm, 4));
    CUS(cudaMemcpy(d, in96, 96, cudaMemcpyHostToDevice));
    k_dbg_g2_selftest<<<1, 32>>>(d, dm);
    CUS(cudaGetLastError());
    CUS(cudaMemcpy(mask, dm, 4, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(dm);
    return 0;
}
