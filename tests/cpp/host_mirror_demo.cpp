// Exercises include/kzgb200.hpp (the C++ mirror of the reference's Context) end to end:
//   usage: host_mirror_demo <trusted_setup.bin> <blob.bin> <out.bin>
// writes commitment(48) | blob proof(48) | cells(262144) | cell proofs(6144) | 4 status bytes
// (verify blob proof, verify cell batch, verify with a corrupted proof, recover == compute).
#include "kzgb200.hpp"
#include <cstdio>
#include <vector>

static std::vector<uint8_t> slurp(const char *p) {
    FILE *f = fopen(p, "rb"); if (!f) { perror(p); exit(2); }
    std::vector<uint8_t> v; uint8_t buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
    fclose(f); return v;
}

int main(int argc, char **argv) {
    if (argc != 4) return 2;
    auto ts = slurp(argv[1]), blob = slurp(argv[2]);
    if (ts.size() != 4096 * 48 * 2 + 65 * 96 || blob.size() != kzgb200::BytesPerBlob) return 3;
    kzgb200_opts opts{}; opts.commit_window = 8; opts.fk20_window = 8;
    kzgb200::Context ctx(ts.data(), ts.data() + 4096 * 48, ts.data() + 2 * 4096 * 48, 65, &opts);
    std::vector<uint8_t> out(48 + 48 + 262144 + 6144 + 4);
    uint8_t *cm = out.data(), *pf = cm + 48, *cells = pf + 48, *cproofs = cells + 262144, *st = cproofs + 6144;
    using kzgb200::Error;
    { Error e0 = ctx.BlobToKZGCommitment(blob.data(), cm); if (e0 != Error::Ok) { fprintf(stderr, "BlobToKZGCommitment -> %d (%s)\n", (int)e0, kzgb200_last_error()); return 10; } }
    if (ctx.ComputeBlobKZGProof(blob.data(), cm, pf) != Error::Ok) return 11;
    if (ctx.ComputeCellsAndKZGProofs(blob.data(), cells, cproofs) != Error::Ok) return 12;
    st[0] = (uint8_t)ctx.VerifyBlobKZGProof(blob.data(), cm, pf);
    std::vector<uint8_t> cms(128 * 48); std::vector<uint64_t> idx(128);
    for (int i = 0; i < 128; ++i) { memcpy(&cms[48 * i], cm, 48); idx[i] = i; }
    st[1] = (uint8_t)ctx.VerifyCellKZGProofBatch(cms.data(), idx.data(), cells, cproofs, 128);
    std::vector<uint8_t> bad(cproofs, cproofs + 6144); memcpy(&bad[0], &bad[48], 48);
    st[2] = (uint8_t)ctx.VerifyCellKZGProofBatch(cms.data(), idx.data(), cells, bad.data(), 128);
    // recover from the odd cells
    std::vector<uint64_t> ids; std::vector<uint8_t> half;
    for (int i = 1; i < 128; i += 2) { ids.push_back(i); half.insert(half.end(), cells + 2048 * i, cells + 2048 * (i + 1)); }
    std::vector<uint8_t> rc(262144), rp(6144);
    Error e = ctx.RecoverCellsAndComputeKZGProofs(ids.data(), half.data(), ids.size(), rc.data(), rp.data());
    st[3] = (e == Error::Ok && !memcmp(rc.data(), cells, 262144) && !memcmp(rp.data(), cproofs, 6144)) ? 0 : 1;
    FILE *f = fopen(argv[3], "wb"); fwrite(out.data(), 1, out.size(), f); fclose(f);
    return 0;
}
