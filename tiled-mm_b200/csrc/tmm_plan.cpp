#include "tmm_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <sstream>

namespace tmm {

namespace {
int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }
size_t elem_size(int dt) { return dt == 0 ? 4 : (dt == 1 ? 8 : (dt == 2 ? 8 : 16)); }

// Throughput model used only to size chunks (never a correctness input).  Measured on B200 (profiles/r1_probe_b200.txt):
// FP64 DMMA GEMM 35.6 TFLOP/s, pinned H2D 55.6 GB/s (48 GB/s while D2H runs).
constexpr double kFlopsF64 = 35e12;   // DGEMM / ZGEMM (8mnk counted for complex): DMMA pipe
constexpr double kFlopsF32 = 140e12;  // SGEMM, FP32-accurate 3xTF32 on tcgen05 (profiles/r1_tc_sgemm_v2.txt); with the FP64 figure the planner
                                      // chose a first block 2 - 5x too narrow for float and left the SMs idle while A streamed in
constexpr double kH2D_alone = 52e9;   // one GPU alone on its link; a GPU grid passes what it measured with all its links busy (PlanInput::h2d_bw)
constexpr double kD2H_alone = 52e9;
constexpr int64_t BM = 128, BN = 64;  // CTA tile of the FP64 kernels

// developer knobs for schedule experiments (never needed for correctness)
double env_or(const char* name, double dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atof(v) : dflt;
}

// fraction of the last wave of CTAs that is filled (2 CTAs per SM resident)
double wave_fill(int64_t tiles, int sm_count) {
    const int64_t slots = 2 * (int64_t)sm_count;
    return (double)tiles / (double)round_up(tiles, slots);
}

// Column-block width: a multiple of 64 near `target` whose CTA count fills whole waves (2 CTAs per SM).
int64_t pick_block_cols(int64_t m, int64_t target, int64_t remaining, int sm_count) {
    if (remaining <= target) return remaining;
    const int64_t tiles_m = (m + BM - 1) / BM;
    const int64_t slots = 2 * (int64_t)sm_count;
    int64_t best = std::min(remaining, round_up(target, BN));
    double best_eff = 0.0;
    for (int64_t cols = std::max<int64_t>(BN, round_up(target * 3 / 4, BN)); cols <= target * 5 / 4 && cols <= remaining; cols += BN) {
        const int64_t tiles = tiles_m * (cols / BN);
        const double eff = (double)tiles / (double)round_up(tiles, slots);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = cols; }
    }
    return best;
}
}  // namespace

// Same function of (dim, max) as the reference heuristic (mm_handle.cpp:89-110): dim if it fits, else the largest
// divisor of dim that is <= max when that divisor is at least half of max, else max.
int optimal_tile_size(int dim, int max_tile) {
    if (dim <= max_tile) return dim;
    int best = 1;
    for (int d = max_tile; d >= 1; --d)
        if (dim % d == 0) { best = d; break; }
    return (max_tile - best <= max_tile / 2) ? best : max_tile;
}

Plan make_plan(const PlanInput& in) {
    Plan p;
    p.es = elem_size(in.dtype);
    const int64_t es = (int64_t)p.es;
    const int64_t align = 128 / es;
    const int64_t m = in.m, n = in.n, k = in.k;
    p.a_rows = in.ta == 'N' ? m : k; p.a_cols = in.ta == 'N' ? k : m;
    p.b_rows = in.tb == 'N' ? k : n; p.b_cols = in.tb == 'N' ? n : k;
    p.pitch_a = round_up(p.a_rows, align);
    p.pitch_b = round_up(p.b_rows, align);
    p.pitch_c = in.copy_c_back ? round_up(m, align) : m;  // device-resident C is compact, ld = m (reference README.md:102-103)
    const size_t full_a = (size_t)p.pitch_a * p.a_cols * es, full_b = (size_t)p.pitch_b * p.b_cols * es;
    const size_t full_c = in.copy_c_back ? (size_t)p.pitch_c * n * es : 0;
    const double F = (in.dtype >= 2) ? 8.0 : 2.0;
    const double kFlops = in.flops > 0 ? in.flops : (in.dtype == 0 ? env_or("TMM_PLAN_F32_FLOPS", kFlopsF32) : env_or("TMM_PLAN_F64_FLOPS", kFlopsF64));  // complex<float>: SIMT unless the caller says otherwise
    const double kH2D = in.h2d_bw > 0 ? in.h2d_bw : kH2D_alone, kD2H = in.d2h_bw > 0 ? in.d2h_bw : kD2H_alone;
    const int64_t kc_cap = std::max<int64_t>(256, std::min<int64_t>(2048, round_up(std::max(64, in.tile_k), 64)));

    // beta != 0 in the resident regime: the caller's C is uploaded into a staging copy and added when a block's accumulation is complete, so
    // that its upload does not gate the first GEMM (make_plan's consumer: run_resident).  TMM_PLAN_DEFER_C=0: upload it up front as in round 1.
    const bool defer_c = in.beta_nonzero && env_or("TMM_PLAN_DEFER_C", 1.0) != 0.0;
    const size_t stage_c = defer_c ? (size_t)p.pitch_c * n * es : 0;
    if (full_a + full_b + full_c + stage_c <= in.budget) {
        // ---------------- resident ----------------
        p.regime = REGIME_RESIDENT;
        p.bytes_a = full_a; p.bytes_b = full_b; p.bytes_c = full_c; p.bytes_c_stage = stage_c;
        // phase-1 column block: wide enough that a k-chunk's GEMM outlasts its upload (with 20 % margin)
        // (m too small for that => the call is PCIe-bound whatever we do: bring A in behind a narrow block and let
        //  phase 2 overlap the D2H of finished C blocks with the H2D of later B blocks)
        const double margin = env_or("TMM_PLAN_MARGIN", 1.3);
        int64_t n1 = std::min<int64_t>(n, 1024);
        bool d2h_bound = false;
        const double sa = 1.0 / std::max(1, in.parts_a), sb = 1.0 / std::max(1, in.parts_b);  // upload shares on a GPU grid
        const double denom = F * (double)m / kFlops - margin * (double)es * sb / kH2D;
        if (denom > 0) {
            const double need = margin * (double)es * sa * (double)m / kH2D / denom;
            n1 = (int64_t)std::min<double>((double)n, std::max(512.0, need));
        }
        {
            // The first block's C can only leave once every k-chunk of A has arrived, i.e. at the end of phase 1; its D2H is hidden
            // only if the remaining columns still have at least that much GEMM work: d2h(n1) <= gemm(n - n1), which bounds n1 by
            // n * rho / (1 + rho), rho = (F k / P) / (es / BW_d2h).  Without the bound, mid-size products (~4000 - 8000 square in FP64)
            // ran as ONE phase and ended with the D2H of the whole C exposed.  (Timeline model tools/model_resident.py, calibrated on
            // the measured 10000^3, whose n1 = 5504 is below its bound of 6600 and unchanged: 9 - 20 % shorter calls predicted for
            // n = 4000 ... 7000; to be measured.  TMM_PLAN_D2H_BOUND=0 switches the bound off.)
            // (per column of C: the D2H costs es m / BW; the window that hides it is the longer of the column's GEMM and its own upload)
            const double hide = std::max(F * (double)m * (double)k / kFlops, (double)es * ((double)k * sb + (in.beta_nonzero ? (double)m : 0.0)) / kH2D);
            const double rho = hide / ((double)es * (double)m / kD2H);
            const int64_t cap = (int64_t)(0.85 * (double)n * rho / (1.0 + rho));  // 15 % slack: the D2H competes with H2D for host memory
            if (in.copy_c_back && env_or("TMM_PLAN_D2H_BOUND", 1.0) != 0.0) {
                if (n1 > cap) { n1 = std::max<int64_t>(std::min<int64_t>(n, 1024), cap); d2h_bound = true; }
                // PCIe-bound whatever we do (denom <= 0): the widest first block whose D2H still hides is also the one that leaves the
                // least GEMM work for after the last byte of B has arrived
                else if (denom <= 0 && cap > n1) { n1 = std::min<int64_t>(n, cap); d2h_bound = true; }
            }
        }
        n1 = std::min<int64_t>(n, round_up(n1, BN));
        if (n1 < n && !d2h_bound) {
            // nudge n1 upwards (at most 12 %) to the width whose CTA count fills whole waves best
            const int64_t tiles_m = (m + BM - 1) / BM;
            int64_t best = n1;
            double best_fill = wave_fill(tiles_m * (n1 / BN), in.sm_count);
            for (int64_t c = n1 + BN; c <= std::min<int64_t>(n, n1 + n1 / 8); c += BN) {
                const double f = wave_fill(tiles_m * (c / BN), in.sm_count);
                if (f > best_fill + 0.01) { best_fill = f; best = c; }
            }
            n1 = best;
        }
        if (n - n1 < 256) n1 = n;  // not worth a second phase
        p.n1 = n1;
        // phase-1 k-chunks: small first chunk (short prologue); a chunk may grow only as fast as the previous chunk's
        // GEMM can hide its upload (ratio r of GEMM time to upload time per unit of k), up to the cap
        {
            const double r = (F * (double)m * (double)n1 / kFlops) / ((double)es * (sa * (double)m + sb * (double)n1) / kH2D);
            // (round-2 sweep at 10000^3, r = 1.32: growth 1.25 -> 58.57 ms, 1.5 -> 58.16 ms, 2.0 -> 60.08 ms; profiles/r2_sweep_plan.txt)
            const double growth = env_or("TMM_PLAN_GROWTH", std::max(1.25, std::min(2.0, 1.15 * r)));
            // When phase 1 as a whole is upload-bound (beta != 0 adds the C block to its uploads; narrow m; slow links of a GPU grid), the GEMMs keep
            // up with the arriving chunks and what remains after the LAST byte has landed is the last chunk's GEMM - exposed in full.  Then the
            // chunks taper off again: never more than ~1/3 of the k range that is still to come.  (dgemm 10000^3 beta = 1: the 2896-wide last chunk
            // of the growth-only schedule left 8.9 ms of GEMM behind the last upload, profiles/r2_beta1_trace_before.txt.)
            const double t_up1 = (double)es * ((double)k * (sa * (double)m + sb * (double)n1) + ((in.beta_nonzero && !defer_c) ? (double)m * (double)n1 : 0.0)) / kH2D;
            const double t_gemm1 = F * (double)m * (double)n1 * (double)k / kFlops;
            const bool taper = env_or("TMM_PLAN_TAPER", t_gemm1 < 1.15 * t_up1 ? 1.0 : 0.0) != 0.0;
            int64_t done = 0;
            int64_t kc = (int64_t)env_or("TMM_PLAN_KC0", 192);  // (first chunk 64 ... 512 swept in round 2: 192 is the shortest call by 0.14 ms, profiles/r2_sweep_plan_fine*.txt)
            const int64_t cap = (int64_t)env_or("TMM_PLAN_KCMAX", (double)kc_cap);
            while (done < k) {
                int64_t c = std::min(kc, k - done);
                if (taper) {
                    const int64_t third = std::max<int64_t>(512, (int64_t)(0.35 * (double)(k - done)) / 64 * 64);
                    c = std::min(c, third);
                    if (k - done - c < 384) c = k - done;
                } else if (k - done - c < kc / 2) c = k - done;  // fold a small remainder into this chunk
                p.chunks.push_back(c);
                done += c;
                kc = std::min<int64_t>(cap, std::max<int64_t>(kc + 64, (int64_t)((double)kc * growth) / 64 * 64));
            }
        }
        // phase-2 column blocks, shrinking towards the end so the last D2H is short
        const int64_t target = std::max<int64_t>(512, std::min<int64_t>((int64_t)env_or("TMM_PLAN_NB", 2048), round_up(std::max(64, in.tile_n), 64)));
        int64_t j0 = n1;
        while (j0 < n) {
            const int64_t remaining = n - j0;
            int64_t nb;
            // (carving a final 128- or 64-column block off the last one, so that the exposed last D2H is shorter, was measured in round 2: 58.18 vs
            //  58.20 ms at 10000^3 - nothing, profiles/r2_final_single_gpu.txt)
            if (remaining <= 512) nb = remaining;
            else if (remaining <= target + 512) nb = std::max<int64_t>(BN, std::min((remaining - 256) / BN * BN, pick_block_cols(m, target, remaining, in.sm_count)));
            else nb = pick_block_cols(m, target, remaining, in.sm_count);
            nb = std::min(nb, remaining);
            p.blocks.push_back(nb);
            j0 += nb;
        }
        p.launches = (int)(p.chunks.size() + p.blocks.size());
        p.h2d_bytes = (uint64_t)es * ((uint64_t)m * k + (uint64_t)k * n + (in.beta_nonzero ? (uint64_t)m * n : 0));
        p.d2h_bytes = in.copy_c_back ? (uint64_t)es * m * n : 0;
        return p;
    }

    // ---------------- streaming (out-of-core) ----------------
    p.regime = REGIME_STREAMING;
    p.slots = 3;
    p.c_is_full = !in.copy_c_back;
    int64_t kc = std::min(kc_cap, round_up(std::max<int64_t>(k, 1), 64));
    int64_t MB = m, NB = n;
    if (!p.c_is_full) {
        const double cbytes = (double)round_up(m, align) * (double)n * (double)es;
        if (cbytes > 0.6 * (double)in.budget) {
            // two C buffers of MB x NB must fit in 60 % of the budget
            const double side = std::sqrt(0.3 * (double)in.budget / (double)es);
            MB = std::min<int64_t>(m, std::max<int64_t>(BM, (int64_t)side / BM * BM));
            NB = std::min<int64_t>(n, std::max<int64_t>(BN, (int64_t)(0.3 * (double)in.budget / (double)es / (double)round_up(MB, align)) / BN * BN));
        }
    }
    p.n_cbuf = (MB == m && NB == n) ? 1 : 2;
    auto c_need = [&]() { return p.c_is_full ? (size_t)0 : (size_t)p.n_cbuf * (size_t)round_up(MB, align) * (size_t)NB * (size_t)es; };
    auto slot_a = [&](int64_t kcc) {
        const int64_t pitch = in.ta == 'N' ? round_up(MB, align) : round_up(kcc, align);
        return (size_t)pitch * (size_t)(in.ta == 'N' ? kcc : MB) * (size_t)es;
    };
    auto slot_b = [&](int64_t kcc) {
        const int64_t pitch = in.tb == 'N' ? round_up(kcc, align) : round_up(NB, align);
        return (size_t)pitch * (size_t)(in.tb == 'N' ? NB : kcc) * (size_t)es;
    };
    while (kc > 64 && c_need() + (size_t)p.slots * (slot_a(kc) + slot_b(kc)) > in.budget) kc = std::max<int64_t>(64, kc / 2 / 64 * 64);
    while (!p.c_is_full && (MB > BM || NB > BN) && c_need() + (size_t)p.slots * (slot_a(kc) + slot_b(kc)) > in.budget) {
        if (MB >= 2 * NB && MB > BM) MB = std::max<int64_t>(BM, MB / 2 / BM * BM);
        else if (NB > BN) NB = std::max<int64_t>(BN, NB / 2 / BN * BN);
        else MB = std::max<int64_t>(BM, MB / 2 / BM * BM);
        p.n_cbuf = 2;
    }
    if (c_need() + (size_t)p.slots * (slot_a(kc) + slot_b(kc)) > in.budget) {
        p.error = "device budget too small for the streaming regime";
        return p;
    }
    p.MB = MB; p.NB = NB; p.kc = kc;
    p.pa_slot = in.ta == 'N' ? round_up(MB, align) : round_up(kc, align);
    p.pb_slot = in.tb == 'N' ? round_up(kc, align) : round_up(NB, align);
    p.a_slot_bytes = slot_a(kc); p.b_slot_bytes = slot_b(kc);
    p.pc_blk = round_up(MB, align);
    p.bytes_a = p.a_slot_bytes * p.slots; p.bytes_b = p.b_slot_bytes * p.slots; p.bytes_c = c_need();
    const int64_t bm = (m + MB - 1) / MB, bn = (n + NB - 1) / NB, nchunks = (k + kc - 1) / kc;
    p.launches = (int)(bm * bn * nchunks);
    // A row-panels are re-sent once per column of super-blocks, B column-panels once per row of super-blocks
    p.h2d_bytes = (uint64_t)es * ((uint64_t)m * k * bn + (uint64_t)k * n * bm + (in.beta_nonzero ? (uint64_t)m * n : 0));
    p.d2h_bytes = in.copy_c_back ? (uint64_t)es * m * n : 0;
    return p;
}

std::string plan_to_json(const PlanInput& in, const Plan& p) {
    std::ostringstream o;
    auto vec = [&](const std::vector<int64_t>& v) {
        o << "[";
        for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
        o << "]";
    };
    o << "{\"regime\":" << p.regime << ",\"elem_size\":" << p.es << ",\"m\":" << in.m << ",\"n\":" << in.n << ",\"k\":" << in.k
      << ",\"pitch_a\":" << p.pitch_a << ",\"pitch_b\":" << p.pitch_b << ",\"pitch_c\":" << p.pitch_c << ",\"bytes_a\":" << p.bytes_a
      << ",\"bytes_b\":" << p.bytes_b << ",\"bytes_c\":" << p.bytes_c << ",\"n1\":" << p.n1 << ",\"chunks\":";
    vec(p.chunks);
    o << ",\"blocks\":";
    vec(p.blocks);
    o << ",\"MB\":" << p.MB << ",\"NB\":" << p.NB << ",\"kc\":" << p.kc << ",\"slots\":" << p.slots << ",\"n_cbuf\":" << p.n_cbuf
      << ",\"c_is_full\":" << (p.c_is_full ? "true" : "false") << ",\"h2d_bytes\":" << p.h2d_bytes << ",\"d2h_bytes\":" << p.d2h_bytes
      << ",\"launches\":" << p.launches << ",\"error\":\"" << p.error << "\"}";
    return o.str();
}

}  // namespace tmm
