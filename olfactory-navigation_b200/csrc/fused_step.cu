// Fused belief update + value-function evaluation: ONE pass over the belief rows per simulation step.
//
// The reference's run_test loop (olfactory_navigation/simulation.py:1452-1497) calls, per step,
//     agent.update_state(...)      -> BeliefSet.update          (agents/pbvi_agent.py:783-787)
//     agent.choose_action()        -> ValueFunction.evaluate_at (agents/pbvi_agent.py:741)   [next iteration]
// As two kernels that is 12·S bytes of HBM traffic per agent-step (update: read b, write b'; evaluate: read b').  Here the
// update IS the A-operand producer of the evaluation GEMM, so b' is written once and never read back: 8·S bytes.
//
//   b'(s') = O(o|s',a) · Σ_s T(s'|s,a) · b(s) · scale          (same arithmetic, same operation order as belief.cu)
//   action' = actions[argmax_g b' · alpha_g]                    (same 3xTF32 two-level accumulation as umma_gemm.cu)
//
// Warp roles (640 threads, one CTA per SM, persistent over tiles of 128 agents):
//   warps 0, 3  source producers: per agent row, 1-D bulk copies (cp.async.bulk, mbarrier complete_tx) of the 128-float window of b
//               that feeds the current 128 destination states (shifted by the row's own move: the y-shift is a row-aligned offset
//               of the copy, wrap-around windows are two copies), +4 floats of halo on both sides for the x-shift, plus the chunk's
//               slice of every likelihood table.  ONE elected lane per warp issues the copies straight-line (~21 cycles each).
//   warp 1      MMA issuer (tcgen05.mma.kind::tf32, A from tensor memory, B = alpha hi / lo planes from shared memory)
//   warp 2      TMEM allocator, then TMA producer of the alpha tiles (2-D, 128B swizzle)
//   warps 4-11  update + transform, two passes per chunk (see the role's own comment)
//   warps 12-19 epilogue: partial accumulators -> FP32 register sums every chunk, first argmax per column half, action lookup
//
// Status (r1, measured on B200, C2 = 100 000 agents x 30 793 states, G = 75): bit-identical rows, 7.1 ms per step against
// 6.45 ms for the two separate kernels, so run_test still uses the two kernels (OLFNAV_FUSED_STEP=1 opts in).  What the clock64
// timeline (OLFNAV_FUSED_DBG=1024) and the ablation bits showed on the way:
//   * 128 threads arriving on one mbarrier = 128 serialised shared-memory atomics (~500 cycles): arrive once per warp;
//   * a cp.async.bulk issued by 32 lanes with different operands is an ELECT / R2UR / UBLKCP loop of ~100 cycles per copy;
//     one elected lane issues a copy per ~21 cycles, and the SM's copy engine retires ~one 512-byte copy per 22 cycles whatever
//     the number of issuing threads (loads and stores together): b' therefore leaves through coalesced st.global, not bulk stores;
//   * with 227 KB of shared memory the L1 is too small for local memory: every spill or dynamically indexed array is an L2 round trip;
//   * the update pass is instruction-issue bound (~95 SASS instructions per 512-byte row visit, 3x the stand-alone belief kernel);
//   * TMA tensor copies cannot do the x-shift: the innermost start coordinate must be 16-byte aligned (scripts/tma_row_test.cu:
//     columns -24 and 4 work, -25 and 1 raise "illegal instruction").
#include "common.cuh"
#include "stencil.cuh"
#include "umma_gemm.cuh"
#include "umma_ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace {

constexpr int FT_THREADS = 640;          // 4 service warps + 8 update/transform warps + 8 epilogue warps: 102 registers each, no spills
constexpr int FKC = 128;                       // destination states per chunk = 4 k-blocks of 32
constexpr int FKB = FKC / BK;                  // 4 k-blocks per chunk = one accumulator lifetime = the 4 TMEM A slots
constexpr int FHALO = 4;
constexpr int FWIN = FKC + 2 * FHALO;          // floats copied per row and chunk
constexpr int FPITCH = FWIN + 4;               // 140 floats: 35 16-byte pieces per row (odd) => conflict-free row-owner accesses
constexpr int FA_STAGE_BYTES = BM * FPITCH * 4;
constexpr int FSA = 2;                         // source stages (2 x 70 KB)
constexpr int F_A_BASE = 256;                  // TMEM: [0, 2 BN) accumulators, [256 + 64 k, +32) hi and (+32, +64) lo of k-block k
constexpr int FUSED_SMEM_TABLES = 8;           // likelihood tables (A_eff x O) staged in shared memory per chunk; more: per-row loads
constexpr int FUSED_MAX_BN = 80;               // the epilogue keeps BN FP32 sums per thread next to 8 transform warps (128 registers)

template <int BN>
struct FusedLayout {
    static constexpr int B_TILE_BYTES = BN * BK * 4;
    static constexpr int B_STAGE_BYTES = 2 * B_TILE_BYTES;
    static constexpr int SB = 3;
    static constexpr int A_OFFSET = 0;
    static constexpr int B_OFFSET = FSA * FA_STAGE_BYTES;          // 143 360: a multiple of 1024 (swizzle atom alignment)
    static constexpr int BAR_OFFSET = B_OFFSET + SB * B_STAGE_BYTES;
    static constexpr int NUM_BARS = 2 * FSA + 2 * SB + 2 * FKB + 4;
    static constexpr int Z_OFFSET = BAR_OFFSET + 256;             // 128 floats: partial normalisers of the second transform group
    static constexpr int RTAB_OFFSET = Z_OFFSET + 512;            // 128 x 8 B: source row address | move class, per tile
    static constexpr int RP_OFFSET = RTAB_OFFSET + 1024;          // per-tile row parameters of the update: scale, move, table offset, source row
    static constexpr int E_OFFSET = RP_OFFSET + 128 * (4 + 4 + 8);  // 128 x (value, index): argmax of the upper column half
    static constexpr int TAB_OFFSET = E_OFFSET + 128 * 8;           // per source stage: the chunk's slice of up to 8 likelihood tables
    static constexpr int TOTAL = TAB_OFFSET + FSA * FUSED_SMEM_TABLES * FKC * 4 + 1024;
    static_assert(B_OFFSET % 1024 == 0 && B_STAGE_BYTES % 1024 == 0, "alpha tiles must stay 1024-byte aligned");
    static_assert(TOTAL <= 227 * 1024 && 8 * NUM_BARS + 8 <= 256, "shared memory budget");
    static_assert(2 * BN <= F_A_BASE && BN <= FUSED_MAX_BN, "tensor memory / register budget");
};

struct FusedParams {
    const float* b_in; float* b_out; long long ld;
    const int* act; const int* obs; const int* src_rows;
    const float* scale_in; float* scale_out; unsigned char* ok;
    const float* obs_tab;
    const int* alpha_actions;
    float* out_val; int* out_arg; int* out_action;
    int n; const int* n_dev;
    int G;
    int num_chunks;            // ceil(Sp / 128)
    int dbg;                   // experiments (OLFNAV_FUSED_DBG): 1 no b' stores, 2 no update math, 4 no MMAs, 8 no source loads, 16 no likelihood loads, 32 no TMEM stores, 64 no stage writes
};

// timeline experiment (OLFNAV_FUSED_DBG & 1024): clock64 stamps of block 0, chunks TS_C0 .. TS_C0 + 3 of its first tile
constexpr int TS_C0 = 100;
__device__ long long g_ts[8][4][8];
#define TS(role, slot) do { if ((p.dbg & 1024) && blockIdx.x == 0 && tile == 0 && kc >= TS_C0 && kc < TS_C0 + 4 && (threadIdx.x & 31) == 0) g_ts[role][kc - TS_C0][slot] = clock64(); } while (0)

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One lane polls, the warp follows: 400 threads spinning on try_wait keep the shared-memory pipe busy for nothing.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void transform_group_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int BN, bool WRAPY, bool WRAPX>
__global__ void __launch_bounds__(FT_THREADS, 1)
fused_step_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                  const FusedParams p, const __grid_constant__ Geom g) {
    using L = FusedLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

    auto a_stage = [&](int s) { return base + L::A_OFFSET + s * FA_STAGE_BYTES; };
    auto b_hi = [&](int s) { return base + L::B_OFFSET + s * L::B_STAGE_BYTES; };
    auto b_lo = [&](int s) { return base + L::B_OFFSET + s * L::B_STAGE_BYTES + L::B_TILE_BYTES; };
    const uint32_t bars = base + L::BAR_OFFSET;
    // Arrival counts are WARPS, not threads: 128 threads arriving on one mbarrier are 128 serialised shared-memory atomics
    // (~500 cycles, measured with the clock64 timeline of OLFNAV_FUSED_DBG=1024); one lane arrives after __syncwarp().
    auto bar_sfull = [&](int s) { return bars + 8 * s; };                                     // source windows landed (2 producers + tx)
    auto bar_sempty = [&](int s) { return bars + 8 * (FSA + s); };                            // 8 transform warps: b' stored, stage read
    constexpr int B0 = 2 * FSA;
    auto bar_bfull = [&](int s) { return bars + 8 * (B0 + s); };
    auto bar_bempty = [&](int s) { return bars + 8 * (B0 + L::SB + s); };
    auto bar_aready = [&](int k) { return bars + 8 * (B0 + 2 * L::SB + k); };                 // hi / lo of k-block k are in TMEM (4 warps)
    auto bar_aempty = [&](int k) { return bars + 8 * (B0 + 2 * L::SB + FKB + k); };           // ... and have been consumed by the MMAs
    auto bar_tfull = [&](int a) { return bars + 8 * (B0 + 2 * L::SB + 2 * FKB + a); };
    auto bar_tempty = [&](int a) { return bars + 8 * (B0 + 2 * L::SB + 2 * FKB + 2 + a); };   // 8 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + L::BAR_OFFSET + 8 * L::NUM_BARS);
    float* zbuf = reinterpret_cast<float*>(gbase + L::Z_OFFSET);
    unsigned long long* rtab = reinterpret_cast<unsigned long long*>(gbase + L::RTAB_OFFSET);
    long long* rp_src = reinterpret_cast<long long*>(gbase + L::RP_OFFSET);
    float* rp_sc = reinterpret_cast<float*>(gbase + L::RP_OFFSET + 128 * 8);
    int* rp_move = reinterpret_cast<int*>(gbase + L::RP_OFFSET + 128 * 12);
    float* ebuf_val = reinterpret_cast<float*>(gbase + L::E_OFFSET);
    int* ebuf_idx = reinterpret_cast<int*>(gbase + L::E_OFFSET + 128 * 4);
    auto tab_stage = [&](int s) { return base + L::TAB_OFFSET + s * (FUSED_SMEM_TABLES * FKC * 4); };
    const int NT = (g.obs_per_action ? g.A : 1) * g.O;              // likelihood tables
    const bool smem_tables = NT <= FUSED_SMEM_TABLES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 2 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmBlo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < FSA; ++s) { mbar_init(bar_sfull(s), 2); mbar_init(bar_sempty(s), 8); }
        for (int s = 0; s < L::SB; ++s) { mbar_init(bar_bfull(s), 1); mbar_init(bar_bempty(s), 1); }
        for (int k = 0; k < FKB; ++k) { mbar_init(bar_aready(k), 4); mbar_init(bar_aempty(k), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n = p.n_dev ? min(p.n, *p.n_dev) : p.n;
    const int num_tiles = (n + BM - 1) / BM;
    const int NC = p.num_chunks;
    const int Sp = g.Sp, Wp = g.Wp;

    if (warp == 0 || warp == 3) {
        // ===================== source producers: per-row bulk copies of the shifted window =====================
        // A cp.async.bulk issued by 32 lanes with different operands compiles to an ELECT / R2UR / UBLKCP / BRA.U.ANY loop of
        // ~100 cycles per copy (measured: 128 copies per chunk took 6 us).  Here the lanes only build a per-tile row table
        // (window origin | move class); ONE elected lane per producer warp then issues the copies of its 64 rows straight-line
        // (~21 cycles per copy).  Warp 0 serves rows 0-63 of the tile, warp 3 rows 64-127.
        const int half = warp == 0 ? 0 : 1;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int cnt[3] = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = 64 * half + lane + 32 * j;
                const int row = tile * BM + r;
                unsigned long long e = 3ull;                     // class 3: no copy (row beyond n)
                int cls = 3;
                if (row < n) {
                    const long long src = p.src_rows ? p.src_rows[row] : row;
                    int a = p.act[row];
                    if (a < 0 || a >= g.A) a = 0;
                    cls = g.moves[a][0] + 1;
                    // byte address of the row's window origin at chunk 0 (state -dy*Wp - FHALO: may lie before the row) | class
                    e = (unsigned long long)((long long)(uintptr_t)(p.b_in + src * p.ld) - 4ll * ((cls - 1) * Wp + FHALO)) | (unsigned long long)cls;
                }
                rtab[r] = e;
#pragma unroll
                for (int c = 0; c < 3; ++c) cnt[c] += __popc(__ballot_sync(0xffffffffu, cls == c));
            }
            __syncwarp();
            if (elect_one_sync()) {
                const bool full_tile = (cnt[0] + cnt[1] + cnt[2]) == 64;
                const unsigned long long* tab = rtab + 64 * half;
                for (int kc = 0; kc < NC; ++kc, ++it) {
                    const int s = it % FSA;
                    const uint32_t ph = (it / FSA) & 1;
                    mbar_wait(bar_sempty(s), ph ^ 1);
                    TS(0, 0);
                    const uint32_t stage = a_stage(s) + (uint32_t)(64 * half) * (FPITCH * 4);
                    const uint32_t bar = bar_sfull(s);
                    // interior chunk: the windows of all three move classes lie inside the row => one 544-byte copy per row from
                    // (window origin + 512 kc), nothing else depends on the class
                    const bool interior = full_tile && (kc * FKC - Wp - FHALO >= 0) && (kc * FKC + Wp - FHALO + FWIN <= Sp);
                    // the chunk's slice of every likelihood table rides on the same barrier (warp 0)
                    const uint32_t tab_bytes = (half == 0 && smem_tables && !(p.dbg & 8)) ? (uint32_t)min(FKC, Sp - kc * FKC) * 4u : 0u;
                    auto issue_tables = [&]() {
                        if (tab_bytes)
                            for (int k = 0; k < NT; ++k)
                                bulk_g2s(tab_stage(s) + (uint32_t)k * (FKC * 4), p.obs_tab + (size_t)k * Sp + kc * FKC, tab_bytes, bar);
                    };
                    if (interior && (p.dbg & 8)) { mbar_expect_tx(bar, 0); continue; }
                    if (interior) {
                        mbar_expect_tx(bar, 64 * FWIN * 4 + NT * tab_bytes);
                        issue_tables();
                        const unsigned long long koff = (unsigned long long)kc * (FKC * 4);
#pragma unroll 1
                        for (int j0 = 0; j0 < 64; j0 += 16) {       // table reads batched ahead of the copies (one thread: latency bound)
                            unsigned long long e[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) e[u] = tab[j0 + u];
#pragma unroll
                            for (int u = 0; u < 16; ++u)
                                bulk_g2s(stage + (uint32_t)(j0 + u) * (FPITCH * 4), reinterpret_cast<const float*>((uintptr_t)((e[u] & ~3ull) + koff)), FWIN * 4, bar);
                        }
                        TS(0, 1);
                        continue;
                    }
                    // boundary chunks (the first and last few of a row): per move class, window [u0, u0 + FWIN) in unwrapped source
                    // coordinates; the part inside [0, Sp) is copy A, the part below 0 or above Sp wraps around (copy B, wrap_y) or
                    // has no source (clip: pass 1 zeroes it)
                    uint32_t offA[3], bytA[3], offB[3], bytB[3];
                    long long srcB[3];
                    uint32_t total = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int u0 = kc * FKC - (c - 1) * Wp - FHALO;
                        const int c0 = max(u0, 0), c1 = min(u0 + FWIN, Sp);
                        offA[c] = (uint32_t)(c0 - u0) * 4u; bytA[c] = c1 > c0 ? (uint32_t)(c1 - c0) * 4u : 0u;
                        int w0 = 0, w1 = 0, ws = 0;
                        if ((p.dbg & 8)) bytA[c] = 0;
                        else if (WRAPY) {
                            if (u0 < 0) { w0 = u0; w1 = min(u0 + FWIN, 0); ws = u0 + Sp; }
                            else if (u0 + FWIN > Sp) { w0 = max(u0, Sp); w1 = u0 + FWIN; ws = w0 - Sp; }
                        }
                        offB[c] = (uint32_t)(w0 - u0) * 4u; bytB[c] = w1 > w0 ? (uint32_t)(w1 - w0) * 4u : 0u;
                        srcB[c] = 4ll * (ws - u0);                // source of copy B relative to the window origin of THIS chunk
                        total += (uint32_t)cnt[c] * (bytA[c] + bytB[c]);
                    }
                    mbar_expect_tx(bar, total + NT * tab_bytes);
                    issue_tables();
                    const long long koff = (long long)kc * (FKC * 4);
                    auto issue = [&](int j, const char* origin, int c) {
                        const uint32_t dst = stage + (uint32_t)j * (FPITCH * 4);
                        if (bytA[c]) bulk_g2s(dst + offA[c], reinterpret_cast<const float*>(origin + offA[c]), bytA[c], bar);
                        if (bytB[c]) bulk_g2s(dst + offB[c], reinterpret_cast<const float*>(origin + srcB[c]), bytB[c], bar);
                    };
#pragma unroll 4
                    for (int j = 0; j < 64; ++j) {
                        const unsigned long long e = tab[j];
                        const int c = (int)(e & 3ull);
                        const char* origin = reinterpret_cast<const char*>((uintptr_t)(e & ~3ull)) + koff;
                        if (c == 0) issue(j, origin, 0);
                        else if (c == 1) issue(j, origin, 1);
                        else if (c == 2) issue(j, origin, 2);
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_tf32<BN>();
        uint32_t it = 0, itb = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < NC; ++kc, ++it) {
                const int acc = it & 1;
                mbar_wait(bar_tempty(acc), ((it >> 1) & 1) ^ 1);
                TS(1, 0);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll 1
                for (int ko = 0; ko < FKB; ++ko, ++itb) {
                    const int kb = ((ko & 1) << 1) | (ko >> 1);        // 0, 2, 1, 3: the order in which the two transform groups deliver
                    const int sb = itb % L::SB;
                    mbar_wait(bar_bfull(sb), (itb / L::SB) & 1);
                    mbar_wait(bar_aready(kb), it & 1);
                    TS(1, 1 + ko);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t ah = tmem_base + F_A_BASE + 64 * kb, al = ah + 32;
                        const uint64_t dbh = make_sw128_desc(b_hi(sb)), dbl = make_sw128_desc(b_lo(sb));
#pragma unroll
                        for (int kk = 0; kk < ((p.dbg & 4) ? 0 : BK / UMMA_K); ++kk) {
                            const uint64_t adv = (uint64_t)(kk * UMMA_K * 4 / 16);
                            tc_mma_tf32_ts(d_tmem, ah + kk * UMMA_K, dbh + adv, idesc, (ko > 0 || kk > 0) ? 1u : 0u);
                            tc_mma_tf32_ts(d_tmem, al + kk * UMMA_K, dbh + adv, idesc, 1u);
                            tc_mma_tf32_ts(d_tmem, ah + kk * UMMA_K, dbl + adv, idesc, 1u);
                        }
                        tc_commit(bar_bempty(sb));
                        tc_commit(bar_aempty(kb));
                        if (ko == FKB - 1) tc_commit(bar_tfull(acc));
                    }
                    __syncwarp();
                }
                TS(1, 5);
            }
        }
    } else if (warp == 2) {
        // ===================== alpha tiles (TMA) =====================
        if (elect_one_sync()) {
            uint32_t itb = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kbo = 0; kbo < NC * FKB; ++kbo, ++itb) {
                    const int ko = kbo & 3;
                    const int kbg = (kbo & ~3) | ((ko & 1) << 1) | (ko >> 1);      // same k-block order as the MMA issuer
                    const int sb = itb % L::SB;
                    mbar_wait(bar_bempty(sb), ((itb / L::SB) & 1) ^ 1);
                    if (p.dbg & 128) { mbar_expect_tx(bar_bfull(sb), 0); continue; }
                    mbar_expect_tx(bar_bfull(sb), 2 * L::B_TILE_BYTES);
                    tma_load_2d(b_hi(sb), &tmBhi, kbg * BK, 0, bar_bfull(sb));
                    tma_load_2d(b_lo(sb), &tmBlo, kbg * BK, 0, bar_bfull(sb));
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== update + transform (8 warps, two passes per chunk) =====================
        // Pass 1, coalesced: warp w updates rows w, w + 8, ... of the stage, lane q owning float4 q of the row's window.  The move,
        // scale and likelihood table are uniform across the warp (one row at a time): no divergence, x-shift by warp shuffles,
        // likelihoods of the chunk prefetched into registers (one float4 per table), b' written back in place.
        // b' is written back in place (for pass 2) and streamed to HBM with coalesced 128-bit stores.  (Bulk stores from the stage
        // were measured first: the SM's bulk-copy engine retires one ~512-byte copy per ~22 cycles, loads and stores together, which
        // is the whole HBM-time budget of a chunk; plain coalesced stores leave the engine to the 128 source copies.)
        // Pass 2, row owners: thread t reads b' of agent row t (= TMEM lane t) and writes its TF32 hi / lo split to tensor memory;
        // group gi (warps 4-7 / 8-11) owns k-blocks 2 gi and 2 gi + 1 of the chunk.
        const int tw = warp - 4;                                   // 0..7
        const int gi = tw >> 2;
        const int t = (threadIdx.x - 128) & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int H = g.H, W = g.W, W4 = g.W4;
        const int e = W - 1, e4 = e >> 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            transform_group_sync();                                // the row table of the previous tile is no longer read
            if (gi == 0) {
                const int row = tile * BM + t;
                int a = 0, o = 0;
                float sc = 0.f;
                long long src = 0;
                if (row < n) {
                    src = p.src_rows ? p.src_rows[row] : row;
                    a = p.act[row]; o = p.obs[row];
                    sc = p.scale_in ? p.scale_in[src] : 1.f;
                    if (a < 0 || a >= g.A || o < 0 || o >= g.O) { a = 0; o = 0; sc = 0.f; }     // invalid ids => failed update
                }
                rp_sc[t] = sc;
                // bit 0 valid, bit 1 dx > 0, bit 2 dx < 0, bits 3-4 dy + 1, bits 8.. likelihood table index x 128 (floats)
                rp_move[t] = (row < n ? 1 : 0) | (g.moves[a][1] > 0 ? 2 : 0) | (g.moves[a][1] < 0 ? 4 : 0) | ((g.moves[a][0] + 1) << 3) |
                             ((((g.obs_per_action ? a : 0) * g.O + o) * FKC) << 8);
                rp_src[t] = src;
            }
            transform_group_sync();
            float zacc = 0.f, zcomp = 0.f;                        // this row owner's share of Z (pass 2), compensated across chunks
            int y = lane / W4, c4 = lane - y * W4;                 // grid position of this lane's float4 (same for every row)
            float* out_tile = p.b_out + ((long long)tile * BM + tw) * p.ld;     // row tw of the tile; this warp's rows are 8 apart

            for (int kc = 0; kc < NC; ++kc, ++it) {
                const int s = it % FSA;
                const int d = kc * FKC + 4 * lane;                 // this lane's first destination state
                const bool in_row = d < Sp;
                TS(3 + gi, 0);
                mbar_wait_warp(bar_sfull(s), (it / FSA) & 1);
                TS(3 + gi, 1);
                float* stg = reinterpret_cast<float*>(gbase + L::A_OFFSET + s * FA_STAGE_BYTES);
                // Fast path (C2 / C5: cyclic y, clipped x, every row of the tile valid, whole chunk inside the row, likelihood slices
                // staged): BRANCH-FREE row code, so that the four rows of an iteration form one basic block and their dependency
                // chains interleave (with uniform branches per row the compiler serialises the rows: ~380 cycles per row visit,
                // measured).  The boundary columns become chunk-invariant per-lane masks.
                const bool fast = WRAPY && !WRAPX && smem_tables && (p.dbg & ~1024) == 0 && (tile + 1) * BM <= n && (kc + 1) * FKC <= Sp;
                if (fast) {
                    const int ek = e & 3;
                    const float* tabs = reinterpret_cast<const float*>(gbase + L::TAB_OFFSET + s * (FUSED_SMEM_TABLES * FKC * 4)) + 4 * lane;
                    float* wrow = stg + tw * FPITCH + FHALO;
                    float* orow = out_tile + d;
                    const bool first_col = c4 == 0, last_col = c4 == W4 - 1, lane0 = lane == 0, lane31 = lane == 31;
                    // dx > 0: the clipped cell W - 1 keeps its own mass (one component of the last float4 of a grid row)
                    const float kx = (c4 == e4 && ek == 0) ? 1.f : 0.f, ky = (c4 == e4 && ek == 1) ? 1.f : 0.f;
                    const float kz = (c4 == e4 && ek == 2) ? 1.f : 0.f, kw = (c4 == e4 && ek == 3) ? 1.f : 0.f;
                    const float k0 = first_col ? 1.f : 0.f;      // dx < 0: cell 0 keeps its own mass
#pragma unroll 1
                    for (int i0 = 0; i0 < 16; i0 += 4) {
                        float4 cur[4], lk[4], rr[4];
                        float hl[4], hr[4], sc[4];
                        int mv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float* w = wrow + u * (8 * FPITCH);
                            cur[u] = *reinterpret_cast<const float4*>(w + 4 * lane);
                            hl[u] = w[-1]; hr[u] = w[FKC];
                            mv[u] = rp_move[tw + 8 * (i0 + u)]; sc[u] = rp_sc[tw + 8 * (i0 + u)];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) lk[u] = *reinterpret_cast<const float4*>(tabs + (mv[u] >> 8));
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 c = cur[u];
                            const bool dxp = (mv[u] & 2) != 0, dxn = (mv[u] & 4) != 0;
                            float left = __shfl_up_sync(0xffffffffu, c.w, 1);
                            float right = __shfl_down_sync(0xffffffffu, c.x, 1);
                            left = lane0 ? hl[u] : left;
                            right = lane31 ? hr[u] : right;
                            left = first_col ? 0.f : left;
                            right = last_col ? 0.f : right;
                            const float fp = dxp ? 1.f : 0.f, fn = dxn ? 1.f : 0.f;
                            float4 out;
                            out.x = dxp ? left : (dxn ? c.y : c.x);
                            out.y = dxp ? c.x : (dxn ? c.z : c.y);
                            out.z = dxp ? c.y : (dxn ? c.w : c.z);
                            out.w = dxp ? c.z : (dxn ? right : c.w);
                            out.x += (fp * kx + fn * k0) * c.x;
                            out.y += (fp * ky) * c.y;
                            out.z += (fp * kz) * c.z;
                            out.w += (fp * kw) * c.w;
                            const float s1 = sc[u];
                            rr[u].x = out.x * s1 * lk[u].x; rr[u].y = out.y * s1 * lk[u].y; rr[u].z = out.z * s1 * lk[u].z; rr[u].w = out.w * s1 * lk[u].w;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            *reinterpret_cast<float4*>(wrow + u * (8 * FPITCH) + 4 * lane) = rr[u];
                            st_stream4(orow + (long long)(8 * u) * p.ld, rr[u]);
                        }
                        wrow += 32 * FPITCH;
                        orow += 32 * p.ld;
                    }
                } else if (!(p.dbg & 2)) {
                    // Four rows per iteration: loads first, then four independent dependency chains, then the stores.  The kernel is
                    // instruction-issue bound here (measured), so the body is kept lean: moves are warp-uniform branches, boundary
                    // columns are real (rare) branches, likelihoods come from the staged table slices.
                    const int ek = e & 3;
                    const float* tabs = reinterpret_cast<const float*>(gbase + L::TAB_OFFSET + s * (FUSED_SMEM_TABLES * FKC * 4)) + 4 * lane;
                    float* wrow = stg + tw * FPITCH + FHALO;
                    float* orow = out_tile + d;
#pragma unroll 1
                    for (int i0 = 0; i0 < 16; i0 += 4) {
                        float4 cur[4], rr[4];
                        float hl[4], hr[4], sc[4];
                        int mv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float* w = wrow + u * (8 * FPITCH);
                            cur[u] = *reinterpret_cast<const float4*>(w + 4 * lane);
                            hl[u] = w[-1]; hr[u] = w[FKC];
                            mv[u] = rp_move[tw + 8 * (i0 + u)]; sc[u] = rp_sc[tw + 8 * (i0 + u)];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = mv[u];
                            float4 c = cur[u];
                            bool src_ok = true;
                            int dy = 0;
                            if (!WRAPY || WRAPX) dy = ((m >> 3) & 3) - 1;
                            if (!WRAPY) {                           // clip: the upstream edge row has no source
                                const int ys = y - dy;
                                src_ok = ys >= 0 && ys < H;
                                if (!src_ok) c = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            float4 out = c;
                            if (m & 2) {                            // dx > 0 (uniform across the warp: one row)
                                float left = __shfl_up_sync(0xffffffffu, c.w, 1);
                                if (lane == 0) left = src_ok ? hl[u] : 0.f;
                                if (c4 == 0) left = 0.f;
                                out = make_float4(left, c.x, c.y, c.z);
                                if (!WRAPX && c4 == e4) add4(out, ek, comp4(c, ek));             // clipped: the edge cell keeps its mass
                            } else if (m & 4) {                     // dx < 0
                                float right = __shfl_down_sync(0xffffffffu, c.x, 1);
                                if (lane == 31) right = src_ok ? hr[u] : 0.f;
                                if (c4 == W4 - 1) right = 0.f;
                                out = make_float4(c.y, c.z, c.w, right);
                                if (!WRAPX && c4 == 0) out.x += c.x;
                            }
                            if (WRAPX) {
                                if ((m & 6) && (c4 == 0 || c4 == e4) && src_ok && (m & 1)) {      // cyclic x: the far end of the source grid row
                                    int ys = y - dy;
                                    if (WRAPY) ys = ys < 0 ? ys + H : (ys >= H ? ys - H : ys);
                                    const float* srow = p.b_in + rp_src[tw + 8 * (i0 + u)] * p.ld + (size_t)ys * Wp;
                                    if ((m & 2) && c4 == 0) out.x = srow[e];
                                    if ((m & 4) && c4 == e4) set4(out, ek, srow[0]);
                                }
                            }
                            if (!WRAPY && dy != 0 && (m & 1) && in_row && y == (dy > 0 ? H - 1 : 0)) {   // clipped edge row keeps its own mass
                                const int dx = (m & 2) ? 1 : ((m & 4) ? -1 : 0);
                                const float4 own = xshift_row(p.b_in + rp_src[tw + 8 * (i0 + u)] * p.ld + (size_t)y * Wp, c4, dx, W, W4, WRAPX ? 1 : 0);
                                out.x += own.x; out.y += own.y; out.z += own.z; out.w += own.w;
                            }
                            rr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (in_row) {
                                float4 lk;
                                if (p.dbg & 16) lk = make_float4(1.f, 1.f, 1.f, 1.f);
                                else if (smem_tables) lk = *reinterpret_cast<const float4*>(tabs + (m >> 8));
                                else lk = __ldg(reinterpret_cast<const float4*>(p.obs_tab + (size_t)((m >> 8) / FKC) * Sp + d));
                                const float s1 = sc[u];
                                rr[u].x = out.x * s1 * lk.x; rr[u].y = out.y * s1 * lk.y; rr[u].z = out.z * s1 * lk.z; rr[u].w = out.w * s1 * lk.w;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            *reinterpret_cast<float4*>(wrow + u * (8 * FPITCH) + 4 * lane) = rr[u];     // b' replaces b in the stage (read back by pass 2)
                            if (in_row && (mv[u] & 1) && !(p.dbg & 1))                                    // ... and goes to HBM: 512 contiguous bytes per row visit
                                st_stream4(orow + (long long)(8 * u) * p.ld, rr[u]);
                        }
                        wrow += 32 * FPITCH;
                        orow += 32 * p.ld;
                    }
                }
                TS(3 + gi, 2);
                transform_group_sync();
                TS(3 + gi, 3);

                // pass 2: row owner t, float4 16 gi .. 16 gi + 15 of its row
                const float* rowp = stg + t * FPITCH + FHALO + 64 * gi;
                float zchunk = 0.f;
#pragma unroll 1
                for (int kk = 0; kk < 2; ++kk) {
                    const int kb = 2 * gi + kk;
                    mbar_wait_warp(bar_aempty(kb), (it & 1) ^ 1);
                    TS(3 + gi, 4 + 2 * kk);
                    tc_fence_after();
                    const uint32_t th = tmem_base + lane_base + F_A_BASE + 64 * kb;
                    if (!(p.dbg & 32)) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float4 v[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const float4*>(rowp + 4 * (kk * 8 + half * 4 + c));
                        zchunk += ((v[0].x + v[0].y) + (v[0].z + v[0].w)) + ((v[1].x + v[1].y) + (v[1].z + v[1].w)) +
                                  (((v[2].x + v[2].y) + (v[2].z + v[2].w)) + ((v[3].x + v[3].y) + (v[3].z + v[3].w)));
                        uint32_t h[16], l[16];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float x4[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t hb = __float_as_uint(x4[k]) & 0xffffe000u;
                                h[4 * c + k] = hb;
                                l[4 * c + k] = __float_as_uint(x4[k] - __uint_as_float(hb)) & 0xffffe000u;
                            }
                        }
                        tc_st16(th + half * 16, h);
                        tc_st16(th + 32 + half * 16, l);
                    }
                    tc_wait_st();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_aready(kb));
                    TS(3 + gi, 5 + 2 * kk);
                }
                {   // compensated add of the chunk's share
                    const float yv = zchunk - zcomp, tv = zacc + yv;
                    zcomp = (tv - zacc) - yv;
                    zacc = tv;
                }
                // stage release: pass 2 of this warp has read the stage
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_sempty(s));
                c4 += FKC / 4;
                while (c4 >= W4) { c4 -= W4; ++y; }
            }
            // Z of row t = the two row owners' shares
            if (gi == 1) zbuf[t] = zacc;
            transform_group_sync();
            if (gi == 0) {
                const int row = tile * BM + t;
                if (row < n) {
                    const float z = zacc + zbuf[t];
                    const bool good = z > 0.f && isfinite(z);
                    p.scale_out[row] = good ? 1.f / z : 0.f;
                    p.ok[row] = good ? 1 : 0;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 12-19) =====================
        // Warp pair (q, ch): TMEM lane quarter q = warp & 3, column half ch: columns [0, C0) or [C0, BN).  Partial accumulators ->
        // FP32 register sums every chunk (two-level accumulation), first argmax per half, halves combined through shared memory.
        constexpr int C0 = ((BN + 31) / 32) * 16;
        constexpr int NCOL = C0;                                   // register sums per thread (second half: BN - C0 <= C0)
        const int q = warp & 3;
        const int ch = (warp - 12) >> 2;
        const int cbeg = ch == 0 ? 0 : C0, ncol = ch == 0 ? C0 : BN - C0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int trow = q * 32 + lane;
            const int row = tile * BM + trow;
            float sum[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) sum[j] = 0.f;
            for (int kc = 0; kc < NC; ++kc, ++it) {
                const int acc = it & 1;
                TS(5, 0);
                mbar_wait_warp(bar_tfull(acc), (it >> 1) & 1);
                TS(5, 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < NCOL; c0 += 16) {
                    if (c0 < ncol) {
                        uint32_t v[16];
                        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cbeg + c0), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty(acc));
                TS(5, 2);
            }
            float best = -INFINITY;
            int best_idx = 0;
#pragma unroll
            for (int j = 0; j < NCOL; ++j)
                if (j < ncol && cbeg + j < p.G && sum[j] > best) { best = sum[j]; best_idx = cbeg + j; }     // strict '>' keeps the first maximum
            if (ch == 1) { ebuf_val[trow] = best; ebuf_idx[trow] = best_idx; }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (ch == 0) {
                const float v1 = ebuf_val[trow];
                if (v1 > best) { best = v1; best_idx = ebuf_idx[trow]; }          // the lower half wins ties (lower indices)
                if (row < n) {
                    if (p.out_val) p.out_val[row] = best;
                    if (p.out_arg) p.out_arg[row] = best_idx;
                    if (p.out_action) p.out_action[row] = p.alpha_actions[best_idx];
                }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

__global__ void fused_value_scale_kernel(int n, const int* __restrict__ n_dev, float* __restrict__ val, const float* __restrict__ scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (n_dev && i >= *n_dev)) return;
    val[i] *= scale[i];
}

template <int BN, bool WRAPY, bool WRAPX>
int launch_fused2(const CUtensorMap& tBh, const CUtensorMap& tBl, const FusedParams& p, const Geom& g, int grid, cudaStream_t st) {
    using L = FusedLayout<BN>;
    static bool configured = false;
    if (!configured) {
        OLF_CUDA(cudaFuncSetAttribute(fused_step_kernel<BN, WRAPY, WRAPX>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        configured = true;
    }
    fused_step_kernel<BN, WRAPY, WRAPX><<<grid, FT_THREADS, L::TOTAL, st>>>(tBh, tBl, p, g);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
template <int BN>
int launch_fused(const CUtensorMap& tBh, const CUtensorMap& tBl, const FusedParams& p, const Geom& g, int grid, cudaStream_t st) {
    if (g.wrap_y) return g.wrap_x ? launch_fused2<BN, true, true>(tBh, tBl, p, g, grid, st) : launch_fused2<BN, true, false>(tBh, tBl, p, g, grid, st);
    return g.wrap_x ? launch_fused2<BN, false, true>(tBh, tBl, p, g, grid, st) : launch_fused2<BN, false, false>(tBh, tBl, p, g, grid, st);
}

}  // namespace

extern "C" int olfnav_can_fuse_policy(const olfnav_model* m, const olfnav_alphas* a) {
    return (m && a && !m->dense && a->Gp <= FUSED_MAX_BN && m->g.Sp >= m->g.Wp + FWIN + 8) ? 1 : 0;
}

// n = upper bound of the rows (grid size), n_dev (optional) = exact device-side count.
int olf_fused_step_launch(const olfnav_model* m, const olfnav_alphas* a, const float* b_in, float* b_out, int64_t ld, int64_t n,
                          const int32_t* act, const int32_t* obs, const int32_t* src_rows, const float* scale_in, float* scale_out,
                          uint8_t* ok, float* value, int32_t* arg, int32_t* action, const int* n_dev, cudaStream_t st) {
    OLF_REQUIRE(m && a && b_in && b_out && act && obs && scale_out && ok, "belief_step_evaluate: null argument");
    OLF_REQUIRE(b_in != b_out, "belief_step_evaluate: the update cannot run in place (use two buffers)");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0, "belief_step_evaluate: ld must be >= Sp and a multiple of 4");
    OLF_REQUIRE(n >= 0 && n <= 0x7fffffffLL, "belief_step_evaluate: n out of range");
    OLF_REQUIRE((((uintptr_t)b_in | (uintptr_t)b_out) & 15) == 0, "belief_step_evaluate: belief buffers must be 16-byte aligned");
    if (!olfnav_can_fuse_policy(m, a)) {
        olfnav_set_error("belief_step_evaluate: needs <= 80 alpha vectors and a grid of >= %d padded states beyond one row (use belief_step + evaluate)", FWIN + 8);
        return OLFNAV_ERR_UNSUPPORTED;
    }
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.b_in = b_in; p.b_out = b_out; p.ld = ld; p.act = act; p.obs = obs; p.src_rows = src_rows;
    p.scale_in = scale_in; p.scale_out = scale_out; p.ok = ok; p.obs_tab = m->obs; p.alpha_actions = a->actions;
    p.out_val = value; p.out_arg = arg; p.out_action = action;
    p.n = (int)n; p.n_dev = n_dev; p.G = a->G;
    p.num_chunks = olf_div_up(m->g.Sp, FKC);
    { const char* e = getenv("OLFNAV_FUSED_DBG"); p.dbg = e ? atoi(e) : 0; }

    const int BN = a->Gp;                       // ceil16(G) <= 128: one n tile
    CUtensorMap tBh, tBl;
    int rc = olf_make_map_2d(&tBh, a->hi, (uint64_t)a->ld, (uint64_t)a->Gp, (uint64_t)a->ld * 4, (uint32_t)BN);
    if (rc == OLFNAV_OK) rc = olf_make_map_2d(&tBl, a->lo, (uint64_t)a->ld, (uint64_t)a->Gp, (uint64_t)a->ld * 4, (uint32_t)BN);
    if (rc != OLFNAV_OK) return rc;
    const int tiles = olf_div_up(n, BM);
    const int grid = tiles < m->sm_count ? tiles : m->sm_count;
    switch (BN) {
        case 16:  rc = launch_fused<16>(tBh, tBl, p, m->g, grid, st); break;
        case 32:  rc = launch_fused<32>(tBh, tBl, p, m->g, grid, st); break;
        case 48:  rc = launch_fused<48>(tBh, tBl, p, m->g, grid, st); break;
        case 64:  rc = launch_fused<64>(tBh, tBl, p, m->g, grid, st); break;
        default:  rc = launch_fused<80>(tBh, tBl, p, m->g, grid, st); break;
    }
    if (rc != OLFNAV_OK) return rc;
    if (p.dbg & 1024) {
        static long long h[8][4][8];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_ts, sizeof(h));
        const long long t0 = h[3][0][0];
        const char* names[6] = {"producer", "mma", "store", "xform0", "xform1", "epilogue"};
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 6; ++r) {
                fprintf(stderr, "chunk+%d %-9s", c, names[r]);
                for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", h[r][c][k] ? h[r][c][k] - t0 : -1);
                fprintf(stderr, "\n");
            }
    }
    if (value) {
        fused_value_scale_kernel<<<olf_div_up(n, 256), 256, 0, st>>>((int)n, n_dev, value, scale_out);
        OLF_LAUNCH_CHECK();
    }
    return OLFNAV_OK;
}

extern "C" int olfnav_belief_step_evaluate(const olfnav_model* m, const olfnav_alphas* a, const float* b_in, float* b_out, int64_t ld,
                                           int64_t n, const int32_t* act, const int32_t* obs, const int32_t* src_rows,
                                           const float* scale_in, float* scale_out, uint8_t* ok, float* value, int32_t* arg,
                                           int32_t* action, void* stream) {
    return olf_fused_step_launch(m, a, b_in, b_out, ld, n, act, obs, src_rows, scale_in, scale_out, ok, value, arg, action, nullptr,
                                 (cudaStream_t)stream);
}
