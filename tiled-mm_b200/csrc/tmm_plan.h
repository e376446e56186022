// Pure host logic of the tile scheduler: given a problem and a device-memory budget, decide the regime,
// the phase-1 column block and k-chunks, the phase-2 column blocks, or the streaming ring geometry.
// No CUDA calls: unit-tested on CPU through tmm_plan_describe (tests/test_plan.py).
//
// Replaces the reference's tiling decisions: optimal_tile_size (mm_handle.cpp:89-110), get_num_tiles /
// get_tile_sizes (tiled_mm.cpp:126-165) and the C-tile visiting order of round_robin (tiled_mm.cpp:292-301).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace tmm {

struct PlanInput {
    int dtype = 1;
    char ta = 'N', tb = 'N';
    int64_t m = 0, n = 0, k = 0;
    bool beta_nonzero = false, copy_c_back = true;
    size_t budget = 0;  // device bytes available for A/B/C storage of this call (full device C excluded when copy_c_back = false)
    int n_streams = 2, tile_m = 5000, tile_n = 5000, tile_k = 5000;  // user hints
    int sm_count = 148;
    double flops = 0;  // GEMM rate the schedule is sized for (flop/s; 0 = the default of the element type) - timing model only
    double h2d_bw = 0, d2h_bw = 0;  // host-link rates in bytes/s the schedule is sized for (0 = one GPU alone on its link: 52 GB/s) - timing model only
    int parts_a = 1, parts_b = 1;  // GPU grid: this rank uploads 1/parts_a of every A panel and 1/parts_b of every B panel (timing model only)
};

enum Regime : int { REGIME_RESIDENT = 0, REGIME_STREAMING = 1 };

struct Plan {
    int regime = REGIME_RESIDENT;
    size_t es = 8;
    int64_t a_rows = 0, a_cols = 0, b_rows = 0, b_cols = 0;  // stored shapes (reference tiled_mm.cpp:507-514)
    int64_t pitch_a = 0, pitch_b = 0, pitch_c = 0;           // device pitches in elements (128-byte multiples)
    size_t bytes_a = 0, bytes_b = 0, bytes_c = 0;            // device bytes (bytes_c = 0 when C lives in the context's full C)
    size_t bytes_c_stage = 0;                                // resident regime, beta != 0: staging copy of the caller's C (added at the end of each block's accumulation)
    // resident regime
    int64_t n1 = 0;                  // phase-1 column block [0, n1)
    std::vector<int64_t> chunks;     // phase-1 k-chunk widths, sum = k
    std::vector<int64_t> blocks;     // phase-2 column block widths, sum = n - n1
    // streaming regime
    int64_t MB = 0, NB = 0, kc = 0;  // C super-block and k-chunk
    int slots = 3, n_cbuf = 1;
    bool c_is_full = false;          // C super-blocks are windows of the context's full device C
    size_t a_slot_bytes = 0, b_slot_bytes = 0;
    int64_t pa_slot = 0, pb_slot = 0, pc_blk = 0;
    // bookkeeping the tests check
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches = 0;
    std::string error;
};

int optimal_tile_size(int dim, int max_tile);
Plan make_plan(const PlanInput& in);
std::string plan_to_json(const PlanInput& in, const Plan& p);

}  // namespace tmm
