// One-pass belief update + policy (olfnav_belief_step_evaluate; the update of olfnav_sim_step when the value function fits one
// tensor-core tile):   BeliefSet.update (reference agents/pbvi_agent.py:783-787)  followed by  ValueFunction.evaluate_at of the
// UPDATED beliefs (agents/pbvi_agent.py:741 = the choose_action of the next run_test iteration, simulation.py:1454)
// with 8·S bytes of HBM traffic per agent instead of 12·S: the updated row is produced in shared memory, leaves for HBM from there
// and is, at the same time, the A operand of the tcgen05 GEMM against the alpha planes.
//
// What the round-1 kernel (one CTA = 128 agents that all moved differently, per-row window copies) got wrong, and what was
// measured for this one (scripts/gather_stream_bench.cu, scripts/write_pattern_bench.cu, profiles/r2_stream_design.md):
//   * tiles are homogeneous in the ACTION: the survivors of a step are bucketed by action (stable counting sort, fused2_plan_*),
//     a tile = 128 rows that make the same move.  The y part of the move is then one shift of the tile's load coordinate
//     (b'(s') = O(s')·b(s' - dy·Wp - dx): the tile is built in the OUTPUT frame, so the alpha planes, the likelihood slices and the
//     stores are unshifted), the x part a register rename in the thread that owns the row, and every table slice is tile uniform.
//   * the rows of a tile are wherever the previous step left them: they are gathered with TMA tile::gather4 (4 rows x 128 B per
//     instruction, 128B swizzle) — measured 6.5-6.7 TB/s read-only on rows picked at random, the same as box loads of
//     consecutive rows.  The updated rows go out to 128 CONSECUTIVE rows of the other buffer (TMA tensor stores); where the
//     agent's belief now lives is handed back to the caller (row_out).
//   * HBM takes 256-byte row segments at 5.9 TB/s only when they are line aligned (4.8 TB/s at a 16-byte offset): grid rows are
//     padded to 64 floats (olfnav_padded_width_for), which also makes the state space a whole number of 64-state K blocks.
//   * work = whole tiles, the CTAs of a wave walking their K blocks in lockstep (they touch the same K range of the tables together), and
//     the tiles of the last, partial wave cut along K into up to 8 parts so that the wave covers all SMs (785 tiles on 148 SMs cost
//     5.3 waves, not 6; a stream-K split of the whole (tile, K block) sequence measured the same and was dropped).  Every unit writes
//     FP32 partial sums; fused2_finish_kernel adds them in unit order (deterministic), takes the first argmax and the normaliser, and
//     writes scale / ok / action per agent.
//
// Numerics: the row is b' = (w·scale_in)·O with w = cA·b(s'-δ) + cB·b(s'-dy·Wp) (cA, cB 0/1 masks of the x move: bit-identical
// to belief_step_tma_kernel).  The GEMM is the 2-way FP16 split of umma_f16_kernel with ONE change: the power-of-two scale of the
// A operand is chosen per (row, K block) from the block's maximum (the normaliser of the row is not known before its last block)
// and undone when the block's partial sum is added to the FP32 register sums — block floating point with FP32 accumulation.
//
// Warp groups (512 threads, one CTA per SM; setmaxnreg moves registers from WG0 to the others — the pool is what the CTA was
// launched with, 512 x 128):
//   WG0  warp 0  producer: rows 0-63 of the tile (TMA gather4) + the table slices of the block
//        warp 1  producer: rows 64-127 of the tile
//        warp 2  TMEM allocation, MMA issuer, loader of the alpha planes (two-deep ring, one block ahead)
//        warp 3  store issuer (TMA tensor stores of the updated block; see the store modes below)
//   WG1  warps 4-7    update + A transform of the even K blocks (thread = row = TMEM lane)
//   WG2  warps 8-11   ... of the odd K blocks
//   WG3  warps 12-15  epilogue: tcgen05.ld of each block's accumulator, FP32 register sums, partial sums out per unit
//
// Store modes (the hand-over of a shared-memory stage from the TMA stores that read it to the TMA gathers that refill it
// is where this kernel has an open fault).  Default: the store thread releases the stage after cp.async.bulk.wait_group.read — fast,
// faults rarely under host contention.  OLFNAV_F2_STORE_SYNC=1: full completion wait after every block; =2: the stage is released
// one block later; OLFNAV_F2_STORE_MODE=1: the transform warps store their own 32 rows and wait.  The three safe variants write
// the same rows and cost 11-15 %.
#include "common.cuh"
#include "umma_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

int olf_make_map_2d(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_rows);

namespace {

constexpr int F2_BK = 64;                  // states per K block
constexpr int F2_STAGES = 5;                // A ring: gathered rows (updated in place) + table slices of a K block
constexpr int F2_BSTAGES = 2;               // B ring: the two FP16 alpha-plane tiles of a K block (L2 resident, consumed by the MMAs only)
constexpr int F2_TA = 4;                   // TMEM slots of the A operand
constexpr int F2_THREADS = 512;
constexpr int F2_BOX = BM * 128;           // bytes of one 32-float box of 128 rows
constexpr int F2_A_BYTES = 2 * F2_BOX;     // raw / updated tile of a K block: 128 rows x 64 floats
constexpr int F2_MAX_O = 6;                // likelihood slices staged per K block
constexpr int F2_TAB_BYTES = (F2_MAX_O + 2) * F2_BK * 4;
constexpr int F2_STORE_MODE = 0;           // see F2Params::store_mode (run_test sets OLFNAV_F2_STORE_MODE=1 when it is forced onto a busy host)
constexpr int F2_STORE_SYNC = 0;           // 1: the store thread waits for the FULL completion of its TMA stores after every K block (safe, -15 %)
constexpr int F2_RING = 8;                 // per-block A scales in flight between the transform and the epilogue
constexpr int F2_MAX_BN = 80;
constexpr int F2_PLAN_ITEMS = 4096;        // agents per CTA of the plan kernels

struct F2Params {
    const float* b_in; long long ld;
    int KB, Sp, Wp, W;
    int store_mode;                        // 1: every transform warp stores its own 32 rows of a block and waits for the FULL completion of that
                                           // store before it releases the stage (OLFNAV_F2_STORE_MODE; 0: one store thread, release after the
                                           // shared-memory read: the open fault of DESIGN §6a)
    int store_sync;                        // OLFNAV_F2_STORE_SYNC: 0 release after the .read wait (fast, the open fault), 1 full completion wait after every
                                           // block, 2 release one block later, after the full completion of the stores (deferred)
    int dbg;                               // OLFNAV_F2_DBG ablation bits: 1 no stores, 2 no MMAs, 4 no tensor-memory writes, 8 no update arithmetic,
                                           // 32 no alpha-plane loads (8 + 32 + 2 or 4: the mode that reproduced the store fault in every run)
    int n_tiles_max; const int* n_tiles_dev;
    const int* tile_act;                   // [n_tiles_max]
    const int* slot_src;                   // [n_tiles_max * 128] belief row in b_in, -1 = empty slot
    const int* slot_obs;                   // [n_tiles_max * 128]
    const float* scale_in;
    const float* obs_tab;                  // [A_eff][O][Sp]
    const float* coef;                     // [A][2][Sp]
    int A, O, obs_per_action;
    int dy[OLFNAV_MAX_ACTIONS], dx[OLFNAV_MAX_ACTIONS];
    float* ws_sum;                         // [segments][128][BN]
    float* ws_z;                           // [segments][2][128]
    const float* binv;                     // [Gp] inverse plane scales
};

template <int BN>
struct F2Layout {
    static constexpr int B_TILE = BN * 128;                     // one FP16 plane tile: BN rows x 64 halfs
    static constexpr int STAGE = F2_A_BYTES + F2_TAB_BYTES;     // A ring stage
    static constexpr int OFF_TAB = F2_A_BYTES;
    static constexpr int BSTAGE = 2 * B_TILE;                   // B ring stage: plane 1, plane 2
    static constexpr int OFF_B = F2_STAGES * STAGE;
    static constexpr int OFF_SINV = OFF_B + F2_BSTAGES * BSTAGE;   // [F2_RING][128] floats
    static constexpr int OFF_EDGE = OFF_SINV + F2_RING * BM * 4;      // [F2_RING][128] floats: raw boundary element of a block (x moves)
    static constexpr int OFF_BARS = OFF_EDGE + F2_RING * BM * 4;
    static constexpr int NUM_BARS = 3 * F2_STAGES + 2 * F2_BSTAGES + F2_TA + 4 + F2_RING;
    static constexpr int TOTAL = OFF_BARS + 8 * NUM_BARS + 16 + 1024;
    static constexpr int A_BASE = 256;                          // TMEM: [0, 2 BN) accumulators, [256 + 64 t, +64) A slot t
    static_assert((STAGE % 1024) == 0 && (BSTAGE % 1024) == 0 && (B_TILE % 1024) == 0, "swizzled tiles start on 1024-byte boundaries");
    static_assert(2 * BN <= A_BASE && A_BASE + 64 * F2_TA <= 512, "tensor memory budget");
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// L2 policies: the belief rows stream through once (evict first), the alpha planes and table slices are re-read by every tile
// (evict last).  Without them the 25 GB of rows that pass through the 126 MB L2 per step push the 10 MB of planes out between two
// uses: measured 15.9 GB of DRAM reads per launch for 12.7 GB of rows (ncu r2j).
__device__ __forceinline__ uint64_t f2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t f2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void f2_gather4(uint32_t dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void f2_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void f2_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" :: "l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void f2_bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// Work units.  Every CTA walks the K blocks of a tile from 0 upwards, and the CTAs of a wave do so together: the streams of
// the 148 tiles in flight then touch the same column range of their rows at the same time, which is what HBM wants for this pattern
// (scripts/gather_stream_bench.cu: 5.4-5.6 TB/s in lockstep against 4.8-5.1 TB/s with staggered K positions; a stream-K split of
// the (tile, K block) sequence staggers them and cost 12-27 %).  Tiles [0, full_units) are whole units; the tiles of the last,
// partial wave are cut along K into `split` parts each so that the wave is spread over all SMs (unit full_units + t * split + q).
// Every unit writes its FP32 partial sums and its share of the normalisers to slot `unit` of the workspaces.
struct F2Work {
    int n_tiles, full_units, split, per, num_units, KB;
    __device__ __forceinline__ void init(const F2Params& p) {
        const int nt = *p.n_tiles_dev;
        n_tiles = nt < p.n_tiles_max ? nt : p.n_tiles_max;
        KB = p.KB;
        const int grid = (int)gridDim.x;
        full_units = (n_tiles / grid) * grid;
        const int tail = n_tiles - full_units;
        split = tail ? (grid / tail < 8 ? grid / tail : 8) : 1;
        if (split < 1) split = 1;
        per = (KB + split - 1) / split;
        num_units = full_units + tail * split;
    }
    __device__ __forceinline__ void unit(int u, int& tile, int& kb0, int& nblk) const {
        if (u < full_units) { tile = u; kb0 = 0; nblk = KB; return; }
        const int t = (u - full_units) / split, q = (u - full_units) - t * split;
        tile = full_units + t;
        kb0 = q * per < KB ? q * per : KB;
        nblk = per < KB - kb0 ? per : KB - kb0;
    }
};

// The update of one row segment (64 states of one agent, thread-private): raw values from the 128B-swizzled tile, updated values
// written back in place and kept in bp[].  DX = x component of the tile's move (compile time: the one-cell shift is a register
// rename); MASKED = the block touches a clipped edge of the grid row or its padding (0 / 1 masks cA, cB from the table slices),
// otherwise every cell simply takes its upstream neighbour.  (v * sc) * O in this order: bit-identical to belief_step_tma_kernel.
template <int DX, bool MASKED>
__device__ __forceinline__ void f2_update(uint8_t* rowp, int sw, const float* lik, const float* cA, const float* cB, float sc, float halo,
                                          float (&bp)[F2_BK], float& zacc, float& mx) {
    if constexpr (DX == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float4* q = reinterpret_cast<float4*>(rowp + (c >> 3) * F2_BOX + (((c & 7) ^ sw) << 4));
            const float4 v = *q;
            const float4 l = *reinterpret_cast<const float4*>(lik + 4 * c);
            float4 r;
            r.x = v.x * sc * l.x; r.y = v.y * sc * l.y; r.z = v.z * sc * l.z; r.w = v.w * sc * l.w;
            *q = r;
            zacc += (r.x + r.y) + (r.z + r.w);
            mx = fmaxf(fmaxf(mx, fmaxf(r.x, r.y)), fmaxf(r.z, r.w));
            bp[4 * c] = r.x; bp[4 * c + 1] = r.y; bp[4 * c + 2] = r.z; bp[4 * c + 3] = r.w;
        }
    } else {
        float v[F2_BK];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(rowp + (c >> 3) * F2_BOX + (((c & 7) ^ sw) << 4));
            v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 l = *reinterpret_cast<const float4*>(lik + 4 * c);
            const float la[4] = {l.x, l.y, l.z, l.w};
            float ca[4] = {1.f, 1.f, 1.f, 1.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (MASKED) {
                const float4 ma = *reinterpret_cast<const float4*>(cA + 4 * c);
                const float4 mb = *reinterpret_cast<const float4*>(cB + 4 * c);
                ca[0] = ma.x; ca[1] = ma.y; ca[2] = ma.z; ca[3] = ma.w;
                cb[0] = mb.x; cb[1] = mb.y; cb[2] = mb.z; cb[3] = mb.w;
            }
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = 4 * c + e;
                // value arriving from one cell back along the move (k - DX), from the halo across the block boundary
                const float nb = DX > 0 ? (k == 0 ? halo : v[k > 0 ? k - 1 : 0]) : (k == F2_BK - 1 ? halo : v[k < F2_BK - 1 ? k + 1 : k]);
                const float w = MASKED ? fmaf(ca[e], nb, cb[e] * v[k]) : nb;     // masks are 0 / 1: exact, one rounding for the merged cell
                r[e] = w * sc * la[e];
            }
            *reinterpret_cast<float4*>(rowp + (c >> 3) * F2_BOX + (((c & 7) ^ sw) << 4)) = make_float4(r[0], r[1], r[2], r[3]);
            zacc += (r[0] + r[1]) + (r[2] + r[3]);
            mx = fmaxf(fmaxf(mx, fmaxf(r[0], r[1])), fmaxf(r[2], r[3]));
            bp[4 * c] = r[0]; bp[4 * c + 1] = r[1]; bp[4 * c + 2] = r[2]; bp[4 * c + 3] = r[3];
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(F2_THREADS, 1)
fused2_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut32,
              const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ F2Params p) {
    using L = F2Layout<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + L::OFF_BARS;
    auto bar_fullA = [&](int s) { return bars + 8 * s; };                       // gathered rows + table slices landed (2 producers)
    auto bar_ready = [&](int s) { return bars + 8 * (F2_STAGES + s); };         // row updated in place, A operand in tensor memory
    auto bar_emptyA = [&](int s) { return bars + 8 * (2 * F2_STAGES + s); };    // the stores of the tile have read it
    auto bar_fullB = [&](int s) { return bars + 8 * (3 * F2_STAGES + s); };     // alpha planes landed
    auto bar_emptyB = [&](int s) { return bars + 8 * (3 * F2_STAGES + F2_BSTAGES + s); };    // the MMAs of the block have retired
    auto bar_aempty = [&](int t) { return bars + 8 * (3 * F2_STAGES + 2 * F2_BSTAGES + t); };    // TMEM A slot t read by its MMAs
    auto bar_tfull = [&](int a) { return bars + 8 * (3 * F2_STAGES + 2 * F2_BSTAGES + F2_TA + a); };
    auto bar_tempty = [&](int a) { return bars + 8 * (3 * F2_STAGES + 2 * F2_BSTAGES + F2_TA + 2 + a); };
    auto bar_edge = [&](int e) { return bars + 8 * (3 * F2_STAGES + 2 * F2_BSTAGES + F2_TA + 4 + e); };    // boundary element of block e published
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + L::OFF_BARS + 8 * L::NUM_BARS);
    float* sinv_ring = reinterpret_cast<float*>(gbase + L::OFF_SINV);
    float* edge_ring = reinterpret_cast<float*>(gbase + L::OFF_EDGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 3 && lane == 0) {
        for (int s = 0; s < F2_STAGES; ++s) { mbar_init(bar_fullA(s), 2); mbar_init(bar_ready(s), 4); mbar_init(bar_emptyA(s), p.store_mode ? 4 : 1); }
        for (int s = 0; s < F2_BSTAGES; ++s) { mbar_init(bar_fullB(s), 1); mbar_init(bar_emptyB(s), 1); }
        for (int t = 0; t < F2_TA; ++t) mbar_init(bar_aempty(t), 1);
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 4); }
        for (int e = 0; e < F2_RING; ++e) mbar_init(bar_edge(e), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmIn) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmOut) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmOut32) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB1) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB2) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    F2Work Wk;
    Wk.init(p);
    const int KB = p.KB;
    // K block visited at position j of a unit: upwards, except for tiles that move west (dx < 0) — the neighbour a row segment needs
    // across its block boundary then always belongs to the block visited just before (see the transform warps)
    auto kb_at = [&](int dx, int kb0, int nblk, int j) { return dx < 0 ? kb0 + nblk - 1 - j : kb0 + j; };

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp < 2) {
            // ===================== producers: gathered rows (64 each), table slices (warp 0), alpha planes (warp 1) =====================
            if (elect_one_sync()) {
                const uint64_t pol_stream = f2_policy_evict_first(), pol_keep = f2_policy_evict_last();
                uint32_t it = 0;
                for (int u = blockIdx.x; u < Wk.num_units; u += gridDim.x) {
                    int tile, kb0, nblk;
                    Wk.unit(u, tile, kb0, nblk);
                    const int a = p.tile_act[tile];
                    const int shift = p.dy[a] * p.Wp;                   // source frame = output frame - dy grid rows (wrap_y: modulo Sp)
                    const bool coef = p.dx[a] != 0;
                    const int4* sl = reinterpret_cast<const int4*>(p.slot_src + (long long)tile * BM + warp * 64);
                    const float* obs_a = p.obs_tab + (size_t)((p.obs_per_action ? a : 0) * p.O) * p.Sp;
                    const float* coef_a = p.coef + (size_t)a * 2 * p.Sp;
                    for (int j = 0; j < nblk; ++j, ++it) {
                        const int s = it % F2_STAGES;
                        const uint32_t ph = (it / F2_STAGES) & 1;
                        const int k0 = kb_at(p.dx[a], kb0, nblk, j) * F2_BK;
                        int c = k0 - shift;
                        if (c < 0) c += p.Sp;
                        if (c >= p.Sp) c -= p.Sp;
                        const uint32_t st = base + s * L::STAGE;
                        mbar_wait(bar_emptyA(s), ph ^ 1);
                        const bool tabs = warp == 0;
                        mbar_expect_tx(bar_fullA(s), F2_A_BYTES / 2 + (tabs ? (uint32_t)(p.O + (coef ? 2 : 0)) * F2_BK * 4 : 0u));
                        const uint32_t dst = st + warp * 64 * 128;
#pragma unroll 4
                        for (int q = 0; q < 16; ++q) {
                            int4 r = __ldg(sl + q);                       // rows of 4 consecutive slots (-1: empty slot, any valid row will do)
                            r.x = max(r.x, 0); r.y = max(r.y, 0); r.z = max(r.z, 0); r.w = max(r.w, 0);
                            f2_gather4(dst + q * 512, &tmIn, c, r.x, r.y, r.z, r.w, bar_fullA(s), pol_stream);
                            f2_gather4(dst + F2_BOX + q * 512, &tmIn, c + 32, r.x, r.y, r.z, r.w, bar_fullA(s), pol_stream);
                        }
                        if (tabs) {
                            for (int o = 0; o < p.O; ++o)
                                f2_bulk_1d(st + L::OFF_TAB + o * F2_BK * 4, obs_a + (size_t)o * p.Sp + k0, F2_BK * 4, bar_fullA(s), pol_keep);
                            if (coef) {
                                f2_bulk_1d(st + L::OFF_TAB + F2_MAX_O * F2_BK * 4, coef_a + k0, F2_BK * 4, bar_fullA(s), pol_keep);
                                f2_bulk_1d(st + L::OFF_TAB + (F2_MAX_O + 1) * F2_BK * 4, coef_a + p.Sp + k0, F2_BK * 4, bar_fullA(s), pol_keep);
                            }
                        }
                    }
                }
            }
        } else if (warp == 2) {
            // ===================== MMA issuer (also fetches the alpha planes, one block ahead: they are L2 resident and only the MMAs
            // read them, so their ring is two deep and independent of how far the row gather runs ahead) =====================
            constexpr uint32_t idesc = make_idesc_f16<BN>();
            const uint64_t pol_keep = f2_policy_evict_last();
            // cursor of the block whose planes are fetched next (one block ahead of the MMAs)
            int pu = blockIdx.x, pj = 0, ptile = 0, pkb0 = 0, pnblk = 0;
            uint32_t pit = 0;
            auto cursor_settle = [&]() {          // skip empty units; returns false at the end of this CTA's work
                while (pu < Wk.num_units) {
                    Wk.unit(pu, ptile, pkb0, pnblk);
                    if (pj < pnblk) return true;
                    pu += gridDim.x; pj = 0;
                }
                return false;
            };
            auto load_b = [&]() {
                if (!cursor_settle()) return;
                const int sb = pit % F2_BSTAGES;
                const uint32_t bst = base + L::OFF_B + sb * L::BSTAGE;
                const int k0 = kb_at(p.dx[p.tile_act[ptile]], pkb0, pnblk, pj) * F2_BK;
                mbar_wait(bar_emptyB(sb), ((pit / F2_BSTAGES) & 1) ^ 1);
                if (elect_one_sync()) {
                    if (p.dbg & 32) mbar_arrive(bar_fullB(sb));
                    else {
                        mbar_expect_tx(bar_fullB(sb), 2 * L::B_TILE);
                        f2_load_2d(bst, &tmB1, k0, 0, bar_fullB(sb), pol_keep);
                        f2_load_2d(bst + L::B_TILE, &tmB2, k0, 0, bar_fullB(sb), pol_keep);
                    }
                }
                __syncwarp();
                ++pit; ++pj;
            };
            load_b();
            uint32_t it = 0;
            for (int u = blockIdx.x; u < Wk.num_units; u += gridDim.x) {
              int tile, kb0, nblk;
              Wk.unit(u, tile, kb0, nblk);
              for (int j = 0; j < nblk; ++j, ++it) {
                load_b();
                const int s = it % F2_STAGES;
                const uint32_t ph = (it / F2_STAGES) & 1;
                const int acc = it & 1;
                mbar_wait(bar_tempty(acc), ((it >> 1) & 1) ^ 1);
                const int sb = it % F2_BSTAGES;
                mbar_wait(bar_ready(s), ph);
                mbar_wait(bar_fullB(sb), (it / F2_BSTAGES) & 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const int ta = it % F2_TA;
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    const uint32_t a1 = tmem_base + L::A_BASE + 64 * ta, a2 = a1 + 32;
                    const uint64_t db1 = make_sw128_desc(base + L::OFF_B + sb * L::BSTAGE), db2 = make_sw128_desc(base + L::OFF_B + sb * L::BSTAGE + L::B_TILE);
                    // cross terms a1.b2 + a2.b1 first, then D = a1.b1 + D * 2^-11 and the remaining main terms
                    if (!(p.dbg & 2)) {
#pragma unroll
                        for (int kk = 0; kk < F2_BK / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 / 16);
                            tc_mma_f16_ts(d_tmem, a1 + kk * 8, db2 + adv, idesc, kk > 0 ? 1u : 0u);
                            tc_mma_f16_ts(d_tmem, a2 + kk * 8, db1 + adv, idesc, 1u);
                        }
#pragma unroll
                        for (int kk = 0; kk < F2_BK / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 / 16);
                            if (kk == 0) tc_mma_f16_ts_scale11(d_tmem, a1, db1, idesc);
                            else tc_mma_f16_ts(d_tmem, a1 + kk * 8, db1 + adv, idesc, 1u);
                        }
                    }
                    tc_commit(bar_emptyB(sb));
                    tc_commit(bar_aempty(ta));
                    tc_commit(bar_tfull(acc));
                }
                __syncwarp();
              }
            }
        } else {
            // ===================== store issuer: updated tile -> 128 consecutive rows of b_out (store_mode 0) =====================
            if (!p.store_mode && elect_one_sync()) {
                // store_sync = 2: a stage is released one block LATER, when the next block is ready to be stored and the stores of this one
                // have had a whole block period to complete — the full-completion wait then finds nothing to wait for, at most one
                // block's stores are ever outstanding, and the price is one stage of ring depth instead of a stall
                const bool deferred = p.store_sync == 2;
                int prev_s = -1;
                uint32_t it = 0;
                for (int u = blockIdx.x; u < Wk.num_units; u += gridDim.x) {
                    int tile, kb0, nblk;
                    Wk.unit(u, tile, kb0, nblk);
                    const int sdx = p.dx[p.tile_act[tile]];
                    for (int j = 0; j < nblk; ++j, ++it) {
                        const int s = it % F2_STAGES;
                        mbar_wait(bar_ready(s), (it / F2_STAGES) & 1);
                        if (deferred && prev_s >= 0) {
                            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                            mbar_arrive(bar_emptyA(prev_s));
                        }
                        const uint32_t src = base + s * L::STAGE;
                        if (!(p.dbg & 1)) {
                            const int k0 = kb_at(sdx, kb0, nblk, j) * F2_BK;
                            f2_store_2d(&tmOut, src, k0, tile * BM);
                            f2_store_2d(&tmOut, src + F2_BOX, k0 + 32, tile * BM);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        // the stage goes back to the producers as soon as the stores have read it (a store read takes a fraction of
                        // a block period; waiting for the next block's commit first would hold every stage one period longer)
                        if (deferred) { prev_s = s; continue; }
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_arrive(bar_emptyA(s));
                        // OPEN FAULT.  Releasing the stage after the .read wait alone ends, rarely, in a hardware exception
                        // ("unspecified launch failure"): once in a few hundred run_test calls when all eight GPUs of a box run this
                        // kernel or the host is saturated, never in 3 000+ calls on a quiet single-GPU box, and in EVERY run of the
                        // ablation mode that removes the arithmetic and the plane loads (OLFNAV_F2_DBG=46).  Without the stores
                        // (OLFNAV_F2_DBG bit 1) it never happens.  Waiting for the FULL completion of the stores before the stage is
                        // released removes it: store_sync = 1 (this thread waits after every block: 0 faults in 27 ablation runs,
                        // 5.85 ms instead of 5.1) or store_mode = 1 (the transform warps store their own rows and wait: 0 faults,
                        // 5.3-5.5 ms, rows still bit-identical); a full wait every 4th or 8th block does not help.  run_test keeps
                        // this fast path for quiet single-rank hosts (simulation.py) and takes store_mode 1 when it is forced elsewhere.
                        if (p.store_sync == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    }
                }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                if (deferred && prev_s >= 0) mbar_arrive(bar_emptyA(prev_s));
            }
        }
    } else if (warp < 12) {
        // ===================== update + A transform: thread = row of the tile = TMEM lane =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        const int grp = (warp - 4) >> 2;                          // K blocks with (it & 1) == grp
        const int row = (threadIdx.x - 128) & 127;
        const int sw = row & 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t it = 0;
        for (int u = blockIdx.x; u < Wk.num_units; u += gridDim.x) {
            int tile, kb0, nblk;
            Wk.unit(u, tile, kb0, nblk);
            const int a = p.tile_act[tile];
            const int dx = p.dx[a];
            const int shift = p.dy[a] * p.Wp;
            const long long slot = (long long)tile * BM + row;
            const int src = p.slot_src[slot];
            int o = p.slot_obs[slot];
            float sc = (src >= 0 && p.scale_in) ? p.scale_in[src] : (src >= 0 ? 1.f : 0.f);
            if (o < 0 || o >= p.O) { o = 0; sc = 0.f; }             // invalid ids fail the row (normaliser 0), like belief_step
            const float* brow = p.b_in + (long long)(src < 0 ? 0 : src) * p.ld;
            float zacc = 0.f;
            // The neighbour a row segment needs across its block boundary under an x move (the last element of the block before it for
            // a move east, the first element of the block after it for a move west) belongs to the block this CTA visited just before
            // (west tiles walk K downwards), which the OTHER transform group owns: every group publishes the raw boundary element of its
            // block in shared memory before it rewrites the tile in place.  (Fetching it from global memory instead — one scattered
            // 4-byte load per row and block — cost 0.7-0.9 ms per step: the streamed rows are evict-first in L2.)  Only the first
            // block of a unit that starts inside a tile (K split of the last wave) reads its neighbour from global memory.
            const int j_first = (int)((grp - (int)(it & 1u)) & 1);
            const uint32_t it_seg = it;
            for (int j = j_first; j < nblk; j += 2) {
                it = it_seg + j;
                const int kb = kb_at(dx, kb0, nblk, j);
                const int s = it % F2_STAGES;
                const uint32_t ph = (it / F2_STAGES) & 1;
                const int ta = it % F2_TA;
                float halo = 0.f;
                if (dx != 0 && j == 0) {
                    int c = kb * F2_BK - shift;
                    if (c < 0) c += p.Sp;
                    if (c >= p.Sp) c -= p.Sp;
                    int hc = dx > 0 ? c - 1 : c + F2_BK;
                    hc = hc < 0 ? 0 : (hc >= p.Sp ? p.Sp - 1 : hc);       // out of range only where the mask is 0: any finite value will do
                    halo = ld_stream1(brow + hc);
                }
                mbar_wait(bar_fullA(s), ph);
                mbar_wait(bar_aempty(ta), ((it / F2_TA) & 1) ^ 1);
                tc_fence_after();
                uint8_t* stage = gbase + s * L::STAGE;
                uint8_t* rowp = stage + row * 128;
                if (dx != 0) {
                    // raw boundary element of this block: element 63 (move east) / element 0 (move west) of the row segment
                    const float mine = dx > 0 ? *reinterpret_cast<const float*>(rowp + F2_BOX + ((7 ^ sw) << 4) + 12)
                                              : *reinterpret_cast<const float*>(rowp + ((0 ^ sw) << 4));
                    edge_ring[(it % F2_RING) * BM + row] = mine;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_edge(it % F2_RING));       // every block arrives (the phase of the barrier counts blocks)
                if (dx != 0) {
                    if (j > 0) {
                        const uint32_t ip = it - 1;
                        mbar_wait(bar_edge(ip % F2_RING), (ip / F2_RING) & 1);
                        halo = edge_ring[(ip % F2_RING) * BM + row];
                    }
                }
                const float* tab = reinterpret_cast<const float*>(stage + L::OFF_TAB);
                const float* lik = tab + o * F2_BK;
                const float* cA = tab + F2_MAX_O * F2_BK;
                const float* cB = cA + F2_BK;
                float bp[F2_BK];                                    // the updated row segment
                float mx = 0.f;
                if (p.dbg & 8) {
#pragma unroll
                    for (int k = 0; k < F2_BK; ++k) bp[k] = 0.f;
                } else if (dx == 0) {
                    f2_update<0, false>(rowp, sw, lik, cA, cB, sc, halo, bp, zacc, mx);
                } else {
                    // a block is 64 consecutive columns of one grid row (Wp % 64 == 0): only the first block of a row (column 0) and the
                    // one that holds the clipped edge / the padding need the masks
                    const int x0 = (kb * F2_BK) % p.Wp;
                    const bool masked = x0 == 0 || x0 + F2_BK - 1 > p.W - 2;
                    if (dx > 0) {
                        if (masked) f2_update<1, true>(rowp, sw, lik, cA, cB, sc, halo, bp, zacc, mx);
                        else        f2_update<1, false>(rowp, sw, lik, cA, cB, sc, halo, bp, zacc, mx);
                    } else {
                        if (masked) f2_update<-1, true>(rowp, sw, lik, cA, cB, sc, halo, bp, zacc, mx);
                        else        f2_update<-1, false>(rowp, sw, lik, cA, cB, sc, halo, bp, zacc, mx);
                    }
                }
                // block scale: the power of two that brings the block's maximum into [2^13, 2^14)
                const uint32_t em = (__float_as_uint(mx) >> 23) & 0xffu;
                const uint32_t es = (em >= 40u && em <= 230u) ? 267u - em : 127u;      // zero / denormal-range blocks: scale 1
                const float scl = __uint_as_float(es << 23);
                sinv_ring[(it % F2_RING) * BM + row] = __uint_as_float((254u - es) << 23);
                const uint32_t th = tmem_base + lane_base + L::A_BASE + 64 * ta;
                if (!(p.dbg & 4))
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int k = 32 * half + 4 * c;
                        const float x0 = bp[k] * scl, x1 = bp[k + 1] * scl, x2 = bp[k + 2] * scl, x3 = bp[k + 3] * scl;
                        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn((x0 - f01.x) * 2048.f, (x1 - f01.y) * 2048.f);
                        const __half2 l23 = __floats2half2_rn((x2 - f23.x) * 2048.f, (x3 - f23.y) * 2048.f);
                        h[2 * c] = *reinterpret_cast<const uint32_t*>(&h01);
                        h[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                        l[2 * c] = *reinterpret_cast<const uint32_t*>(&l01);
                        l[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                    }
                    tc_st16(th + half * 16, h);
                    tc_st16(th + 32 + half * 16, l);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the in-place row is read by the TMA store
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready(s));
                if (p.store_mode) {
                    // this warp's 32 rows of the block go out now; the stage is released when the store has COMPLETED (not merely read
                    // shared memory: DESIGN §6a) — the wait falls into the half block period this warp would idle anyway
                    if (elect_one_sync()) {
                        if (!(p.dbg & 1)) {
                            const uint32_t srcw = base + s * L::STAGE + (uint32_t)(warp & 3) * 32 * 128;
                            f2_store_2d(&tmOut32, srcw, kb * F2_BK, tile * BM + (warp & 3) * 32);
                            f2_store_2d(&tmOut32, srcw + F2_BOX, kb * F2_BK + 32, tile * BM + (warp & 3) * 32);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                        mbar_arrive(bar_emptyA(s));
                    }
                    __syncwarp();
                }
            }
            it = it_seg + nblk;
            // normaliser share of this segment (both groups write theirs; the finishing kernel adds them in a fixed order)
            p.ws_z[((long long)u * 2 + grp) * BM + row] = zacc;
        }
    } else if (warp < 16) {
        // ===================== epilogue: block accumulators -> FP32 register sums -> partial sums of the segment =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t it = 0;
        for (int u = blockIdx.x; u < Wk.num_units; u += gridDim.x) {
            int tile, kb0, nblk;
            Wk.unit(u, tile, kb0, nblk);
            float sum[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] = 0.f;
            for (int j = 0; j < nblk; ++j, ++it) {
                const int acc = it & 1;
                // (no wait on bar_ready here: the epilogue may run more than one phase of that barrier behind the transform, and a parity
                // wait that is lapped never returns.  The block's A scale was written before the transform arrived on bar_ready; the MMA
                // thread acquired that barrier before it issued the MMAs whose commit this wait acquires: visible by causality.)
                mbar_wait(bar_tfull(acc), (it >> 1) & 1);
                tc_fence_after();
                const float sinv = sinv_ring[(it % F2_RING) * BM + row];
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
#pragma unroll
                    for (int e = 0; e < 16; ++e) sum[c0 + e] = fmaf(__uint_as_float(v[e]), sinv, sum[c0 + e]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty(acc));
            }
            float* wsrow = p.ws_sum + ((long long)u * BM + row) * BN;
#pragma unroll
            for (int j = 0; j < BN; j += 4)
                *reinterpret_cast<float4*>(wsrow + j) = make_float4(sum[j] * __ldg(p.binv + j), sum[j + 1] * __ldg(p.binv + j + 1),
                                                                    sum[j + 2] * __ldg(p.binv + j + 2), sum[j + 3] * __ldg(p.binv + j + 3));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- plan: survivors bucketed by action, 128 per tile ----------------------------------------------------------------------
// counts[c][a] = survivors with action a among the items of CTA c (F2_PLAN_ITEMS each)
__global__ void __launch_bounds__(1024) fused2_plan_count_kernel(int n, const int* __restrict__ n_dev, const int* __restrict__ act, int A,
                                                                  int* __restrict__ counts) {
    __shared__ int cnt[OLFNAV_MAX_ACTIONS];
    if (threadIdx.x < OLFNAV_MAX_ACTIONS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int m = n_dev ? min(n, *n_dev) : n;
    const int i0 = blockIdx.x * F2_PLAN_ITEMS + 4 * threadIdx.x;
    int local[OLFNAV_MAX_ACTIONS];
#pragma unroll
    for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) local[a] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < m) {
            int a = act[i0 + k];
            if (a < 0 || a >= A) a = 0;
#pragma unroll
            for (int b = 0; b < OLFNAV_MAX_ACTIONS; ++b) local[b] += (a == b) ? 1 : 0;
        }
#pragma unroll
    for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) {
        int v = local[a];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[a], v);
    }
    __syncthreads();
    if (threadIdx.x < OLFNAV_MAX_ACTIONS) counts[blockIdx.x * OLFNAV_MAX_ACTIONS + threadIdx.x] = cnt[threadIdx.x];
}

// Stable placement: survivor i with action a goes to slot 128 * (first tile of a) + (number of survivors j < i with action a).
__global__ void __launch_bounds__(1024) fused2_plan_place_kernel(int n, const int* __restrict__ n_dev, const int* __restrict__ act,
                                                                  const int* __restrict__ obs, const int* __restrict__ src_rows,
                                                                  const int* __restrict__ in_rows, int A, int n_ctas,
                                                                  const int* __restrict__ counts, int* __restrict__ slot_src,
                                                                  int* __restrict__ slot_obs, int* __restrict__ slot_id,
                                                                  int* __restrict__ tile_act, int* __restrict__ n_tiles_dev, int n_tiles_max) {
    __shared__ int before[OLFNAV_MAX_ACTIONS], total[OLFNAV_MAX_ACTIONS], first_tile[OLFNAV_MAX_ACTIONS + 1];
    __shared__ int wsum[32][OLFNAV_MAX_ACTIONS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < OLFNAV_MAX_ACTIONS) {
        int b = 0, t = 0;
        for (int c = 0; c < n_ctas; ++c) {
            const int v = counts[c * OLFNAV_MAX_ACTIONS + threadIdx.x];
            if (c < (int)blockIdx.x) b += v;
            t += v;
        }
        before[threadIdx.x] = b;
        total[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int ft = 0;
        for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) { first_tile[a] = ft; ft += (total[a] + BM - 1) / BM; }
        first_tile[OLFNAV_MAX_ACTIONS] = ft;
        if (blockIdx.x == 0) *n_tiles_dev = ft;
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a)
            for (int t = first_tile[a] + threadIdx.x; t < first_tile[a + 1] && t < n_tiles_max; t += blockDim.x) tile_act[t] = a;

    const int m = n_dev ? min(n, *n_dev) : n;
    const int i0 = blockIdx.x * F2_PLAN_ITEMS + 4 * threadIdx.x;
    int mya[4];
    int local[OLFNAV_MAX_ACTIONS];
#pragma unroll
    for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) local[a] = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mya[k] = -1;
        if (i0 + k < m) {
            int a = act[i0 + k];
            mya[k] = (a < 0 || a >= A) ? (0 | 0x100) : a;          // invalid action: bucket 0, flagged (the row fails)
#pragma unroll
            for (int b = 0; b < OLFNAV_MAX_ACTIONS; ++b) local[b] += ((mya[k] & 0xff) == b) ? 1 : 0;
        }
    }
    // exclusive prefix of the per-thread counts over the CTA, per action
    int pre[OLFNAV_MAX_ACTIONS];
#pragma unroll
    for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) {
        int v = local[a];
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        pre[a] = v - local[a];
        if (lane == 31) wsum[w][a] = v;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) {
            int v = wsum[lane][a];
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            wsum[lane][a] = v;
        }
    }
    __syncthreads();
    int run[OLFNAV_MAX_ACTIONS];
#pragma unroll
    for (int a = 0; a < OLFNAV_MAX_ACTIONS; ++a) run[a] = before[a] + (w > 0 ? wsum[w - 1][a] : 0) + pre[a];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (mya[k] < 0) continue;
        const int a = mya[k] & 0xff;
        int rank = 0;
#pragma unroll
        for (int b = 0; b < OLFNAV_MAX_ACTIONS; ++b) if (a == b) { rank = run[b]; run[b] += 1; }
        const long long slot = (long long)first_tile[a] * BM + rank;
        const int i = i0 + k;
        const int r = src_rows ? src_rows[i] : i;
        slot_src[slot] = in_rows ? in_rows[r] : r;
        slot_obs[slot] = (mya[k] & 0x100) ? -1 : obs[i];
        slot_id[slot] = i;
    }
}

// ---- finish: partial sums of the segments of a tile, in CTA order -> first argmax, normaliser, per-agent outputs ---------------
__global__ void __launch_bounds__(128) fused2_finish_kernel(int BN, int G, int KB, int grid_main, int n_tiles_max, const int* __restrict__ n_tiles_dev,
                                                            const int* __restrict__ slot_src, const int* __restrict__ slot_id,
                                                            const float* __restrict__ ws_sum, const float* __restrict__ ws_z,
                                                            const int* __restrict__ alpha_actions, float* __restrict__ scale_out,
                                                            unsigned char* __restrict__ ok, float* __restrict__ value, int* __restrict__ arg,
                                                            int* __restrict__ action, int* __restrict__ row_out) {
    const int tile = blockIdx.x, l = threadIdx.x;
    const int nt = min(*n_tiles_dev, n_tiles_max);
    if (tile >= nt) return;
    const long long slot = (long long)tile * BM + l;
    const int src = slot_src[slot];
    if (src < 0) { scale_out[slot] = 0.f; return; }
    // the units that hold this tile's partial sums (same arithmetic as F2Work)
    const int full_units = (nt / grid_main) * grid_main;
    const int tail = nt - full_units;
    int split = tail ? (grid_main / tail < 8 ? grid_main / tail : 8) : 1;
    if (split < 1) split = 1;
    const int u0 = tile < full_units ? tile : full_units + (tile - full_units) * split;
    const int nu = tile < full_units ? 1 : split;
    float z = 0.f;
    for (int c = 0; c < nu; ++c) {
        const float* zs = ws_z + (long long)(u0 + c) * 2 * BM;
        z += zs[l];
        z += zs[BM + l];
    }
    float best = -INFINITY;
    int best_idx = 0;
    for (int j0 = 0; j0 < G; j0 += 4) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < nu; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(ws_sum + ((long long)(u0 + c) * BM + l) * BN + j0);
            x.x += v.x; x.y += v.y; x.z += v.z; x.w += v.w;
        }
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (j0 + e < G && xs[e] > best) { best = xs[e]; best_idx = j0 + e; }
    }
    const bool good = z > 0.f && isfinite(z);
    const float inv = good ? 1.f / z : 0.f;
    scale_out[slot] = inv;
    const int i = slot_id[slot];
    ok[i] = good ? 1 : 0;
    if (value) value[i] = best * inv;
    if (arg) arg[i] = best_idx;
    if (action) action[i] = alpha_actions[best_idx];
    if (row_out) row_out[i] = (int)slot;
}

typedef CUresult (*F2EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int f2_encode(CUtensorMap* map, CUtensorMapDataType dt, const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride_bytes, uint32_t b0,
              uint32_t b1, CUtensorMapL2promotion promo) {
    static F2EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (F2EncodeTiledFn)q;
    }
    if (!fn) { olfnav_set_error("fused step: cuTensorMapEncodeTiled not available"); return OLFNAV_ERR_CUDA; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {stride_bytes};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { olfnav_set_error("fused step: cuTensorMapEncodeTiled failed (%d)", (int)r); return OLFNAV_ERR_CUDA; }
    return OLFNAV_OK;
}

template <int BN>
int f2_launch(const CUtensorMap& tIn, const CUtensorMap& tOut, const CUtensorMap& tOut32, const CUtensorMap& tB1, const CUtensorMap& tB2,
              const F2Params& p, int grid, cudaStream_t st) {
    using L = F2Layout<BN>;
    static OlfPerDeviceOnce configured;
    if (configured.first())
        OLF_CUDA(cudaFuncSetAttribute(fused2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    fused2_kernel<BN><<<grid, F2_THREADS, L::TOTAL, st>>>(tIn, tOut, tOut32, tB1, tB2, p);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

}  // namespace

extern "C" int olfnav_can_fuse_policy(const olfnav_model* m, const olfnav_alphas* a) {
    if (!m || !a || m->dense || !a->h1 || a->Gp > F2_MAX_BN || !m->coef) return 0;
    const Geom& g = m->g;
    if ((g.Sp % F2_BK) != 0 || g.O > F2_MAX_O || g.A > OLFNAV_MAX_ACTIONS) return 0;
    for (int k = 0; k < g.A; ++k) {
        const int dy = g.moves[k][0], dx = g.moves[k][1];
        if (dy < -1 || dy > 1 || dx < -1 || dx > 1) return 0;
        if (dy != 0 && (!g.wrap_y || (g.Wp % F2_BK) != 0)) return 0;     // a clipped vertical move merges two grid rows: two-kernel path
        if (dx != 0 && g.wrap_x) return 0;                               // a wrapped horizontal move reads a far cell: two-kernel path
    }
    return 1;
}

// rows the output buffer of olfnav_belief_step_evaluate must hold for n input rows: every action's last tile is padded to 128
extern "C" int64_t olfnav_fused_out_rows(const olfnav_model* m, int64_t n) {
    if (!m) return 0;
    return ((n + BM - 1) / BM + m->g.A) * BM;
}

// n = upper bound of the rows, n_dev (optional) = exact device-side count.
int olf_fused_step_launch(const olfnav_model* m, const olfnav_alphas* a, const float* b_in, int64_t rows_in, float* b_out, int64_t rows_out,
                          int64_t ld, int64_t n, const int32_t* act, const int32_t* obs, const int32_t* src_rows, const int32_t* in_rows,
                          const float* scale_in, float* scale_out, uint8_t* ok, float* value, int32_t* arg, int32_t* action,
                          int32_t* row_out, const int* n_dev, cudaStream_t st) {
    OLF_REQUIRE(m && a && b_in && b_out && act && obs && scale_out && ok, "belief_step_evaluate: null argument");
    OLF_REQUIRE(b_in != b_out, "belief_step_evaluate: the update cannot run in place (use two buffers)");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0, "belief_step_evaluate: ld must be >= Sp and a multiple of 4");
    OLF_REQUIRE(n >= 0 && n <= 0x7fffffffLL / 2, "belief_step_evaluate: n out of range");
    OLF_REQUIRE((((uintptr_t)b_in | (uintptr_t)b_out) & 15) == 0, "belief_step_evaluate: belief buffers must be 16-byte aligned");
    if (n == 0) return OLFNAV_OK;
    if (!olfnav_can_fuse_policy(m, a)) {
        olfnav_set_error("belief_step_evaluate: needs <= %d alpha vectors, grid rows padded to 64 floats, wrapped vertical and clipped horizontal moves "
                         "(use belief_step + evaluate)", F2_MAX_BN);
        return OLFNAV_ERR_UNSUPPORTED;
    }
    OLF_REQUIRE(rows_out >= olfnav_fused_out_rows(m, n), "belief_step_evaluate: b_out / scale_out need olfnav_fused_out_rows(m, n) rows");
    OLF_REQUIRE(rows_in >= 1, "belief_step_evaluate: rows_in");
    OLF_CUDA(cudaSetDevice(m->device));
    const Geom& g = m->g;
    const int BN = a->Gp;
    const int n_tiles_max = (int)(olfnav_fused_out_rows(m, n) / BM);
    const int grid = m->sm_count;
    const int n_seg = n_tiles_max + grid;
    const int plan_ctas = olf_div_up(n, F2_PLAN_ITEMS);

    OlfScratch s_counts(st), s_tile(st), s_slots(st), s_sum(st), s_z(st);
    OLF_CUDA(s_counts.alloc((size_t)plan_ctas * OLFNAV_MAX_ACTIONS * sizeof(int)));
    OLF_CUDA(s_tile.alloc((size_t)(n_tiles_max + 1) * sizeof(int)));
    OLF_CUDA(s_slots.alloc((size_t)3 * n_tiles_max * BM * sizeof(int)));
    OLF_CUDA(s_sum.alloc((size_t)n_seg * BM * BN * sizeof(float)));
    OLF_CUDA(s_z.alloc((size_t)n_seg * 2 * BM * sizeof(float)));
    int* tile_act = s_tile.as<int>();
    int* n_tiles_dev = tile_act + n_tiles_max;
    int* slot_src = s_slots.as<int>();
    int* slot_obs = slot_src + (size_t)n_tiles_max * BM;
    int* slot_id = slot_obs + (size_t)n_tiles_max * BM;

    OLF_CUDA(cudaMemsetAsync(slot_src, 0xff, (size_t)n_tiles_max * BM * sizeof(int), st));        // -1: empty slot
    fused2_plan_count_kernel<<<plan_ctas, 1024, 0, st>>>((int)n, n_dev, act, g.A, s_counts.as<int>());
    fused2_plan_place_kernel<<<plan_ctas, 1024, 0, st>>>((int)n, n_dev, act, obs, src_rows, in_rows, g.A, plan_ctas, s_counts.as<int>(), slot_src,
                                                         slot_obs, slot_id, tile_act, n_tiles_dev, n_tiles_max);
    olf_count_launch(1);
    OLF_LAUNCH_CHECK();

    // diagnostics: a host that stalls between the plan kernels and the main kernel
    static const int dbg_stall_us = [] { const char* e = getenv("OLFNAV_DEBUG_STALL_US"); return e ? atoi(e) : 0; }();
    if (dbg_stall_us > 0) { static unsigned ctr = 0; if ((++ctr % 3) == 0) usleep((useconds_t)dbg_stall_us); }
    F2Params p;
    memset(&p, 0, sizeof(p));
    p.b_in = b_in; p.ld = ld; p.KB = g.Sp / F2_BK; p.Sp = g.Sp; p.Wp = g.Wp; p.W = g.W;
    { const char* e = getenv("OLFNAV_F2_DBG"); p.dbg = e ? atoi(e) : 0; }
    { const char* e = getenv("OLFNAV_F2_STORE_MODE"); p.store_mode = e ? (atoi(e) != 0) : F2_STORE_MODE; }
    { const char* e = getenv("OLFNAV_F2_STORE_SYNC"); p.store_sync = e ? atoi(e) : F2_STORE_SYNC; if (p.store_sync < 0) p.store_sync = 0; }
    p.n_tiles_max = n_tiles_max; p.n_tiles_dev = n_tiles_dev; p.tile_act = tile_act; p.slot_src = slot_src; p.slot_obs = slot_obs;
    p.scale_in = scale_in; p.obs_tab = m->obs; p.coef = m->coef; p.A = g.A; p.O = g.O; p.obs_per_action = g.obs_per_action;
    for (int k = 0; k < g.A; ++k) { p.dy[k] = g.moves[k][0]; p.dx[k] = g.moves[k][1]; }
    p.ws_sum = s_sum.as<float>(); p.ws_z = s_z.as<float>(); p.binv = a->binv;

    CUtensorMap tIn, tOut, tOut32, tB1, tB2;
    int rc = f2_encode(&tIn, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, b_in, (uint64_t)g.Sp, (uint64_t)rows_in, (uint64_t)ld * 4, 32, 1, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rc == OLFNAV_OK) rc = f2_encode(&tOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, b_out, (uint64_t)g.Sp, (uint64_t)rows_out, (uint64_t)ld * 4, 32, BM, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc == OLFNAV_OK) rc = f2_encode(&tOut32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, b_out, (uint64_t)g.Sp, (uint64_t)rows_out, (uint64_t)ld * 4, 32, 32, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc == OLFNAV_OK) rc = f2_encode(&tB1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a->h1, (uint64_t)a->ldh, (uint64_t)a->Gp, (uint64_t)a->ldh * 2, F2_BK, (uint32_t)BN, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc == OLFNAV_OK) rc = f2_encode(&tB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a->h2, (uint64_t)a->ldh, (uint64_t)a->Gp, (uint64_t)a->ldh * 2, F2_BK, (uint32_t)BN, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (rc != OLFNAV_OK) return rc;
    switch (BN) {
        case 16: rc = f2_launch<16>(tIn, tOut, tOut32, tB1, tB2, p, grid, st); break;
        case 32: rc = f2_launch<32>(tIn, tOut, tOut32, tB1, tB2, p, grid, st); break;
        case 48: rc = f2_launch<48>(tIn, tOut, tOut32, tB1, tB2, p, grid, st); break;
        case 64: rc = f2_launch<64>(tIn, tOut, tOut32, tB1, tB2, p, grid, st); break;
        default: rc = f2_launch<80>(tIn, tOut, tOut32, tB1, tB2, p, grid, st); break;
    }
    if (rc != OLFNAV_OK) return rc;
    fused2_finish_kernel<<<n_tiles_max, BM, 0, st>>>(BN, a->G, p.KB, grid, n_tiles_max, n_tiles_dev, slot_src, slot_id, p.ws_sum, p.ws_z, a->actions,
                                                     scale_out, ok, value, arg, action, row_out);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_belief_step_evaluate(const olfnav_model* m, const olfnav_alphas* a, const float* b_in, int64_t rows_in, float* b_out,
                                           int64_t rows_out, int64_t ld, int64_t n, const int32_t* act, const int32_t* obs,
                                           const int32_t* src_rows, const float* scale_in, float* scale_out, uint8_t* ok, float* value,
                                           int32_t* arg, int32_t* action, int32_t* row_out, void* stream) {
    return olf_fused_step_launch(m, a, b_in, rows_in, b_out, rows_out, ld, n, act, obs, src_rows, nullptr, scale_in, scale_out, ok, value, arg,
                                 action, row_out, nullptr, (cudaStream_t)stream);
}
