// Host-side sharding of one batched call over the GPUs of a context (pure C++, no CUDA): contiguous
// ranges of independent units, one per GPU, outputs written to disjoint slices of the caller's buffers, no collective.
// Unit-tested without a GPU through kzgb200_dbg_shard_plan_json (tests/test_sharding.py), and at world size 2 over gloo.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

struct ShardRange { size_t lo, hi; };        // units [lo, hi)

// n independent units over at most n_dev devices, at least min_per_dev units per device (so that small calls stay
// on one GPU: a split costs a host thread hand-off and leaves each GPU a partial wave).  Never returns empty ranges;
// returns one range for n < 2 * min_per_dev.
static inline std::vector<ShardRange> shard_units(size_t n, size_t n_dev, size_t min_per_dev) {
    std::vector<ShardRange> r;
    if (n == 0) return r;
    size_t parts = std::max<size_t>(1, std::min(n_dev, n / std::max<size_t>(1, min_per_dev)));
    for (size_t i = 0; i < parts; ++i) {
        size_t lo = n * i / parts, hi = n * (i + 1) / parts;
        if (hi > lo) r.push_back({lo, hi});
    }
    return r;
}

// Verdicts of VerifyCellKZGProofBatch over the devices: verdict b covers cells [off[b], off[b+1]).  Ranges are over
// VERDICTS, balanced by cell count (greedy prefix cut at multiples of total/parts), each holding >= 1 verdict.
static inline std::vector<ShardRange> shard_verdicts(const uint64_t *off, size_t nb, size_t n_dev, size_t min_cells_per_dev) {
    std::vector<ShardRange> r;
    if (nb == 0) return r;
    const uint64_t total = off[nb] - off[0];
    size_t parts = std::max<size_t>(1, std::min<size_t>(std::min(n_dev, nb), (size_t)(total / std::max<size_t>(1, min_cells_per_dev))));
    size_t b = 0;
    for (size_t i = 0; i < parts && b < nb; ++i) {
        const uint64_t target = off[0] + total * (i + 1) / parts;
        size_t e = b + 1;                                            // at least one verdict
        while (e < nb && off[e] < target && (nb - e) > (parts - 1 - i)) ++e;      // leave one verdict for every later part
        if (i + 1 == parts) e = nb;
        r.push_back({b, e});
        b = e;
    }
    return r;
}

// Three-way merge of the sub-verdicts of ONE logical verdict computed in pieces (index order): the first error wins
// (the reference reports the first failing element before any pairing, verify.go:102-119, api_eip7594.go:190-213),
// else VERIFY_FAILED if any piece failed, else OK.
static inline int32_t merge_sub_verdicts(const int32_t *sub, size_t n) {
    int32_t out = 0;
    for (size_t i = 0; i < n; ++i) {
        if (sub[i] > 1) return sub[i];
        if (sub[i] == 1) out = 1;
    }
    return out;
}
