// Host bookkeeping of VerifyCellKZGProofBatch (pure C++, no CUDA): what the reference does in
// api_eip7594.go:163-265 before any arithmetic -- per verdict: de-duplication of the commitments on their raw bytes in
// first-seen order (deduplicateKZGCommitments, :238-265), the cell-index range check (:184-188), the grouping of
// the cells by unique commitment ("row") -- plus the work-item lists of the GPU kernels.  Unit-tested without a GPU
// against a Python model (tests/test_cell_plan.py) through kzgb200_dbg_plan_cell_batches_json.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

struct CellPlan {
    // per verdict b
    std::vector<int32_t> bstatus;               // OK or BAD_CELL_INDEX
    std::vector<uint64_t> batch_start;          // first cell of verdict b
    std::vector<uint64_t> batch_row_off;        // rows [batch_row_off[b], batch_row_off[b+1]) belong to verdict b
    std::vector<uint64_t> batch_item_off;       // interpolation items of verdict b
    std::vector<int32_t> large_of;              // index among the large verdicts, or -1
    // per cell k
    std::vector<uint32_t> batch_of;
    // rows = unique commitments, in first-seen order inside each verdict
    std::vector<uint8_t> uniq_bytes;            // 48 bytes per row
    std::vector<uint64_t> row_off;              // cells of row r: row_cells[row_off[r] .. row_off[r+1])
    std::vector<uint32_t> row_cells, row_batch;
    // work items (runs of <= item cells of one verdict): interpolation (all verdicts) and 4-bit bucket MSM (small verdicts only)
    std::vector<uint64_t> item_start, item_end, vs_item_start, vs_item_end, vs_batch_item_off;
    // large verdicts (>= large cells): cells grouped by cell index; slot = large index * 128 + column
    std::vector<uint32_t> large_ids, l_order;   // l_order: cell indices k grouped by (large verdict, column)
    std::vector<uint64_t> l_item_start, l_item_end, l_slot_item_off;     // positions in l_order
    std::vector<uint64_t> r_item_start, r_item_end, r_slot_item_off;     // runs of rows of each large verdict
};

// false: batch_offsets not monotone / out of range / not covering exactly [0, N) (batch_offsets[0] == 0 and
// batch_offsets[nb] == N are required: a cell outside every verdict would still be decoded and its errors folded
// into verdict 0 -- ADVICE r1)
static inline bool plan_cell_batches(const uint8_t *h_cm, const uint64_t *cell_indices, size_t N, const uint64_t *batch_offsets, size_t nb,
                                     uint64_t item, uint64_t large, uint64_t row_item, int32_t st_bad_cell_index, CellPlan &P, uint64_t l_item = 0) {
    if (!l_item) l_item = item;               // run length of the large verdicts' column items (default: the interpolation's)
    P = CellPlan();
    P.bstatus.assign(nb, 0); P.batch_start.assign(nb, 0); P.batch_row_off.assign(nb + 1, 0); P.batch_item_off.assign(nb + 1, 0);
    P.large_of.assign(nb, -1); P.batch_of.assign(N, 0); P.row_cells.assign(N, 0); P.row_off.assign(1, 0);
    P.vs_batch_item_off.assign(nb + 1, 0); P.l_slot_item_off.assign(1, 0); P.r_slot_item_off.assign(1, 0);
    std::vector<uint32_t> row_of(N), row_count;
    std::unordered_map<std::string, uint32_t> seen;
    if (nb == 0 ? N != 0 : (batch_offsets[0] != 0 || batch_offsets[nb] != N)) return false;
    for (size_t b = 0; b < nb; ++b) {
        const uint64_t lo = batch_offsets[b], hi = batch_offsets[b + 1];
        if (hi < lo || hi > N) return false;
        P.batch_start[b] = lo;
        // de-duplicate on raw bytes, first-seen order.  Runs of equal commitments (the usual layout: all cells of a
        // blob together) skip the hash lookup.
        seen.clear(); row_count.clear();
        uint32_t prev_row = 0;
        for (uint64_t k = lo; k < hi; ++k) {
            P.batch_of[k] = (uint32_t)b;
            uint32_t row;
            if (k > lo && memcmp(h_cm + k * 48, h_cm + (k - 1) * 48, 48) == 0) row = prev_row;
            else {
                std::string key((const char *)h_cm + k * 48, 48);
                auto it = seen.find(key);
                if (it == seen.end()) {
                    row = (uint32_t)row_count.size(); seen.emplace(key, row); row_count.push_back(0);
                    P.uniq_bytes.insert(P.uniq_bytes.end(), key.begin(), key.end());
                } else row = it->second;
            }
            prev_row = row; row_of[k] = row; ++row_count[row];
            if (cell_indices[k] >= 128) P.bstatus[b] = st_bad_cell_index;
        }
        // counting sort of the verdict's cells by row
        const size_t row_base = P.row_off.size() - 1;
        for (uint32_t cnt : row_count) P.row_off.push_back(P.row_off.back() + cnt);
        std::vector<uint64_t> fill(P.row_off.begin() + row_base, P.row_off.begin() + row_base + row_count.size());
        for (uint64_t k = lo; k < hi; ++k) P.row_cells[fill[row_of[k]]++] = (uint32_t)k;
        P.batch_row_off[b + 1] = P.row_off.size() - 1;
        for (uint64_t s = lo; s < hi; s += item) { P.item_start.push_back(s); P.item_end.push_back(std::min(hi, s + item)); }
        P.batch_item_off[b + 1] = P.item_start.size();
        if (hi - lo >= large) {
            P.large_of[b] = (int32_t)P.large_ids.size();
            P.large_ids.push_back((uint32_t)b);
            for (uint64_t s = P.batch_row_off[b]; s < P.batch_row_off[b + 1]; s += row_item) {
                P.r_item_start.push_back(s); P.r_item_end.push_back(std::min<uint64_t>(P.batch_row_off[b + 1], s + row_item));
            }
            P.r_slot_item_off.push_back(P.r_item_start.size());
            uint64_t cnt[129] = {0};
            for (uint64_t k = lo; k < hi; ++k) ++cnt[(cell_indices[k] & 127) + 1];
            const uint64_t base = P.l_order.size();
            for (int q = 0; q < 128; ++q) cnt[q + 1] += cnt[q];
            P.l_order.resize(base + (hi - lo));
            uint64_t fillc[128];
            for (int q = 0; q < 128; ++q) fillc[q] = base + cnt[q];
            for (uint64_t k = lo; k < hi; ++k) P.l_order[fillc[cell_indices[k] & 127]++] = (uint32_t)k;
            for (int q = 0; q < 128; ++q) {
                for (uint64_t s = base + cnt[q]; s < base + cnt[q + 1]; s += l_item) { P.l_item_start.push_back(s); P.l_item_end.push_back(std::min(base + cnt[q + 1], s + l_item)); }
                P.l_slot_item_off.push_back(P.l_item_start.size());
            }
        } else {
            for (uint64_t s = lo; s < hi; s += item) { P.vs_item_start.push_back(s); P.vs_item_end.push_back(std::min(hi, s + item)); }
        }
        P.vs_batch_item_off[b + 1] = P.vs_item_start.size();
    }
    const size_t U = P.row_off.size() - 1;
    P.row_batch.assign(U, 0);
    for (size_t b = 0; b < nb; ++b) for (uint64_t rw = P.batch_row_off[b]; rw < P.batch_row_off[b + 1]; ++rw) P.row_batch[rw] = (uint32_t)b;
    return true;
}
