// tcgen05 / TMEM / TMA GEMM with FP32-accurate split precision (3xTF32) and a fused per-row segmented
// max/argmax epilogue.  Serves
//   * ValueFunction.evaluate_at   (reference agents/pbvi_agent.py:741):  beliefs (N x S) x alphas (S x G)
//   * the PBVI backup contraction (inside solvers.*.solve, reference agents/fsvi_agent.py:201-224):
//     beliefs (B x S) x Gamma_{a,o} (S x G*A*O), argmax over g inside every (a,o) segment
//   * the infotaxis sums (reference agents/infotaxis_agent.py:271-274): beliefs x pull tables, plain store epilogue.
//
// D (128 x BN, FP32, in TMEM) += A_hi·B_hi + A_lo·B_hi + A_hi·B_lo per K block of 32, where x_hi is x with the
// low 13 mantissa bits cleared (exactly TF32) and x_lo = trunc_tf32(x - x_hi).  The dropped terms are
// below 2^-21 |a||b|, which protects the argmax the same way an FP32 FMA chain does.
//
// Two-level accumulation: the tensor core adds into its FP32 accumulator with truncation (measured on B200:
// a relative bias of about -1.5e-8 per accumulate, -2e-4 after the 11 580 accumulates of K = 30 876).  The
// accumulator therefore only lives for KCHUNK k-blocks (48 MMAs); the epilogue warps then add the partial
// sums into FP32 registers (round to nearest) while the MMA warp fills the other TMEM buffer.
//
// Warp roles (384 threads, one CTA per SM, persistent over (m tile, n tile) work units):
//   warp 0      TMA producer: A tile (raw FP32) + B_hi + B_lo tiles, 128B swizzle, mbarrier complete_tx
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (A from TMEM, B from shared memory)
//   warp 2      TMEM allocator
//   warps 4-7   A transform: raw FP32 row (shared memory) -> hi / lo in tensor memory (tcgen05.st)
//   warps 8-11  epilogue: tcgen05.ld each K-chunk's partial accumulator, FP32 register sums, segmented max/argmax
// Every mbarrier wait carries a clock watchdog that traps instead of hanging the GPU.
//
// umma_f16_kernel (second half of this file) is the same pipeline with a 2-way FP16 split and power-of-two block scales instead of
// 3xTF32 (three kind::f16 MMAs per K = 16, one accumulator through scale-input-d): half the tensor work per K.  It serves the backup
// selection from G = 128 up and evaluate_at / the run_test policy from G = 64 up; OLFNAV_GEMM_F16=0 falls back to the TF32 kernel.
#include "common.cuh"
#include "umma_gemm.cuh"
#include "umma_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int NTHREADS = 384;
constexpr int KCHUNK = 4;        // k-blocks accumulated inside TMEM before the partial sum is flushed to registers
constexpr int A_TILE_BYTES = BM * BK * 4;   // 16 KB

struct GemmParams {
    int n_rows, K, rows_b;
    int n_seg, seg_len, seg_stride;
    float* out_val;
    int* out_arg;
    unsigned long long* ws;      // packed (value, first index) keys combined with atomicMax when a row's columns span several units
    int num_m_tiles, num_n_tiles, num_k_blocks;
    int num_units, group_w;      // work units = (m tile, n tile); n tiles are visited in groups of group_w so that the CTAs of one
                                 // wave share A and B tiles in L2 (a wave of 148 units = 148 / group_w m tiles x group_w n tiles)
    float* out_store;            // != NULL: store mode, out_store[row * out_ld + col] = product (col < rows_b); no max / argmax
    long long out_ld;
    // Tail split: the units of the last, partial wave are cut along K into tail_split parts each, so that the
    // wave is spread over all SMs instead of costing a whole wave on a few (measured: evaluate of 100 000 rows = 5.28 waves ran
    // as long as 6 full waves).  Units [0, full_units) are whole tiles; unit full_units + t * tail_split + q is part q of tail tile t.
    // The parts' partial sums go to tail_ws[t][q][row][col] and are added in a fixed order by tail_finish_kernel (deterministic).
    int full_units, tail_split;
    float* tail_ws;
};

// unit id -> k-block range [kb0, kb1) and tail part (-1: whole tile); the tile itself comes from unit_to_tile
__device__ __forceinline__ int unit_tile_index(const GemmParams& p, int u) {
    return (p.tail_split > 1 && u >= p.full_units) ? p.full_units + (u - p.full_units) / p.tail_split : u;
}
__device__ __forceinline__ void unit_k_range(const GemmParams& p, int u, int& kb0, int& kb1, int& part) {
    kb0 = 0; kb1 = p.num_k_blocks; part = -1;
    if (p.tail_split > 1 && u >= p.full_units) {
        part = (u - p.full_units) % p.tail_split;
        const int per = (p.num_k_blocks + p.tail_split - 1) / p.tail_split;
        kb0 = min(part * per, p.num_k_blocks);
        kb1 = min(kb0 + per, p.num_k_blocks);
    }
}

// unit id -> (m tile, n tile)
__device__ __forceinline__ void unit_to_tile(const GemmParams& p, int u, int& mt, int& nt) {
    u = unit_tile_index(p, u);
    const int per_group = p.num_m_tiles * p.group_w;
    int g = u / per_group;
    const int n_groups = (p.num_n_tiles + p.group_w - 1) / p.group_w;
    if (g >= n_groups) g = n_groups - 1;
    const int rem = u - g * per_group;
    const int wg = min(p.group_w, p.num_n_tiles - g * p.group_w);
    mt = rem / wg;
    nt = g * p.group_w + (rem - mt * wg);
}

// order-preserving float -> uint32, packed with the inverted column index: atomicMax keeps the largest value and, among equal
// values, the smallest index (np.argmax's first maximum)
__device__ __forceinline__ unsigned long long pack_key(float v, int idx) {
    unsigned u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);
}


// ============================================================================================================
// v3: A operand in TENSOR MEMORY.  Measured on B200 (scripts/umma_rate_bench.cu, OLFNAV_GEMM_DBG experiments): with both
// operands in shared memory the kernel is bound by the 128 B/clk shared-memory port — TMA writes, the hi/lo transform's
// loads and stores AND the tensor core's own operand reads (4 KB of A per K = 8 step for FP32 operands) all go through it.
// Here the transform warps read the raw FP32 A tile once (conflict-free row reads of the 128B-swizzled tile), split it in
// registers and write hi and lo straight into tensor memory (tcgen05.st); the MMAs take A from TMEM and only B from shared
// memory.  Shared memory per stage shrinks to the raw A tile + the two B planes, so more stages fit.
//   TMEM columns: [0, 2*BN) two accumulator buffers, [A_BASE + 64*s, +32) hi and (+32, +64) lo of stage s.
template <int BN, int STAGES, int TA>
struct SmemLayoutTS {
    static constexpr int B_TILE_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static constexpr int A_BASE = BN <= 64 ? 128 : 256;
    static constexpr int NUM_BARS = 3 * STAGES + 4 + TA;
    static_assert(A_BASE + 64 * TA <= 512 && 2 * BN <= A_BASE, "tensor memory budget");
    static_assert(TOTAL <= 227 * 1024 && 8 * NUM_BARS + 8 <= 256, "shared memory budget");
};

template <int BN, int STAGES, int TA>
__global__ void __launch_bounds__(NTHREADS, 1)
umma_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const GemmParams p) {
    using L = SmemLayoutTS<BN, STAGES, TA>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

    auto a_raw = [&](int s) { return base + s * L::STAGE_BYTES; };
    auto b_hi = [&](int s) { return base + s * L::STAGE_BYTES + A_TILE_BYTES; };
    auto b_lo = [&](int s) { return base + s * L::STAGE_BYTES + A_TILE_BYTES + L::B_TILE_BYTES; };
    const uint32_t bars = base + L::BAR_OFFSET;
    auto bar_full = [&](int s) { return bars + 8 * s; };
    auto bar_ready = [&](int s) { return bars + 8 * (STAGES + s); };
    auto bar_empty = [&](int s) { return bars + 8 * (2 * STAGES + s); };
    auto bar_tfull = [&](int a) { return bars + 8 * (3 * STAGES + a); };
    auto bar_tempty = [&](int a) { return bars + 8 * (3 * STAGES + 2 + a); };
    auto bar_aempty = [&](int a) { return bars + 8 * (3 * STAGES + 4 + a); };     // TMEM A slot a has been read by its MMAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + L::BAR_OFFSET + 8 * L::NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmBlo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_ready(s), 128); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 128); }
        for (int a = 0; a < TA; ++a) mbar_init(bar_aempty(a), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int KB = p.num_k_blocks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            uint32_t it = 0;
            for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
                int mt, nt, kbeg, kend, part;
                unit_to_tile(p, u, mt, nt);
                unit_k_range(p, u, kbeg, kend, part);
                for (int kb = kbeg; kb < kend; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_full(s), A_TILE_BYTES + 2 * L::B_TILE_BYTES);
                    tma_load_2d(a_raw(s), &tmA, kb * BK, mt * BM, bar_full(s));
                    tma_load_2d(b_hi(s), &tmBhi, kb * BK, nt * BN, bar_full(s));
                    tma_load_2d(b_lo(s), &tmBlo, kb * BK, nt * BN, bar_full(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_tf32<BN>();
        uint32_t it = 0, ccount = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int kbeg, kend, part;
            unit_k_range(p, u, kbeg, kend, part);
            for (int kb0 = kbeg; kb0 < kend; kb0 += KCHUNK, ++ccount) {
                const int acc = ccount & 1;
                const uint32_t aph = (ccount >> 1) & 1;
                mbar_wait(bar_tempty(acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int kb1 = min(kend, kb0 + KCHUNK);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(bar_ready(s), ph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const int ta = it % TA;
                        const uint32_t ah = tmem_base + L::A_BASE + 64 * ta, al = ah + 32;
                        const uint64_t dbh = make_sw128_desc(b_hi(s)), dbl = make_sw128_desc(b_lo(s));
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * UMMA_K * 4 / 16);   // +32 B per K step inside the swizzle row
                            tc_mma_tf32_ts(d_tmem, ah + kk * UMMA_K, dbh + adv, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                            tc_mma_tf32_ts(d_tmem, al + kk * UMMA_K, dbh + adv, idesc, 1u);
                            tc_mma_tf32_ts(d_tmem, ah + kk * UMMA_K, dbl + adv, idesc, 1u);
                        }
                        tc_commit(bar_empty(s));                 // frees the shared-memory stage once these MMAs retire
                        tc_commit(bar_aempty(ta));               // ... and the TMEM slot of A
                        if (kb == kb1 - 1) tc_commit(bar_tfull(acc));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== A transform: raw FP32 row (shared memory) -> hi / lo (tensor memory) =====================
        // Thread t owns row t of the tile (= TMEM lane t; warp w may touch lanes 32 (w % 4) .. +31).  The tile is 128B-swizzled:
        // the 16-byte piece c of row r sits at piece c ^ (r & 7), which makes the row-wise reads of a quarter warp conflict free.
        const int t = threadIdx.x - 128;
        const uint8_t* rowp = gbase + t * 128;
        const int sw = t & 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int kbeg, kend, part;
            unit_k_range(p, u, kbeg, kend, part);
            for (int kb = kbeg; kb < kend; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int ta = it % TA;
                mbar_wait(bar_full(s), ph);
                mbar_wait(bar_aempty(ta), ((it / TA) & 1) ^ 1);
                tc_fence_after();
                const uint8_t* src = rowp + s * L::STAGE_BYTES;
                const uint32_t th = tmem_base + lane_base + L::A_BASE + 64 * ta;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const float4*>(src + (((half * 4 + c) ^ sw) << 4));
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float x[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t hb = __float_as_uint(x[e]) & 0xffffe000u;
                            h[4 * c + e] = hb;
                            l[4 * c + e] = __float_as_uint(x[e] - __uint_as_float(hb)) & 0xffffe000u;
                        }
                    }
                    tc_st16(th + half * 16, h);
                    tc_st16(th + 32 + half * 16, l);
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(bar_ready(s));
            }
        }
    } else if (warp >= 8) {
        // ===================== epilogue: FP32 register accumulation + segmented max / argmax =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const bool direct = p.ws == nullptr;         // one n tile covers all columns: write the results directly
        uint32_t ccount = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int mt, nt;
            unit_to_tile(p, u, mt, nt);
            const long long row = (long long)mt * BM + q * 32 + lane;
            const bool row_ok = row < p.n_rows;
            const int col0 = nt * BN;
            int cur_seg = col0 / p.seg_stride, r = col0 - cur_seg * p.seg_stride;
            float best = -INFINITY;
            int best_idx = 0;
            bool dirty = false;
            float sum[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] = 0.f;
            int kbeg, kend, part;
            unit_k_range(p, u, kbeg, kend, part);
            for (int kb0 = kbeg; kb0 < kend; kb0 += KCHUNK, ++ccount) {
                const int acc = ccount & 1;
                const uint32_t aph = (ccount >> 1) & 1;
                mbar_wait(bar_tfull(acc), aph);
                tc_fence_after();
                if constexpr (BN >= 112 || (BN % 32) != 0) {
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 16) {
                        uint32_t v[16];
                        tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(v[j]);
                    }
                } else {
#pragma unroll
                    for (int c0 = 0; c0 < BN; c0 += 32) {
                        uint32_t v[32];
                        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_tempty(acc));
            }
            if (part >= 0) {          // tail part: partial sums of this K range, combined by tail_finish_kernel
                float* wsrow = p.tail_ws + ((((long long)(unit_tile_index(p, u) - p.full_units) * p.tail_split + part) * BM) + q * 32 + lane) * BN;
#pragma unroll
                for (int j = 0; j < BN; j += 4) *reinterpret_cast<float4*>(wsrow + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
                continue;
            }
            if (p.out_store) {
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < BN; ++j)
                        if (col0 + j < p.rows_b) p.out_store[row * p.out_ld + col0 + j] = sum[j];
                }
                continue;
            }
            auto flush = [&]() {
                if (row_ok && cur_seg < p.n_seg) {
                    if (direct) {
                        p.out_val[row * p.n_seg + cur_seg] = best;
                        p.out_arg[row * p.n_seg + cur_seg] = best_idx;
                    } else if (dirty) {
                        atomicMax(p.ws + row * p.n_seg + cur_seg, pack_key(best, best_idx));
                    }
                }
            };
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                const float x = sum[j];
                if (r < p.seg_len && cur_seg < p.n_seg) {
                    dirty = true;
                    if (x > best) { best = x; best_idx = r; }
                }
                if (++r == p.seg_stride) {
                    flush();
                    r = 0; ++cur_seg; best = -INFINITY; best_idx = 0; dirty = false;
                }
            }
            if (r != 0 && !direct) flush();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ============================================================================================================
// FP16 2-way split variant (segmented max / argmax epilogue only; DESIGN §8 item 0, scripts/f16_ts_probe.cu,
// scripts/fp16_split_study.py).  x = s a with s = 2^k per row, a1 = fp16(x), a2 = fp16((x - a1) 2^11); the B planes b1, b2 are
// built the same way by the caller with a power-of-two scale t_j per row of B (binv[j] = 1 / t_j).
//     a.b = (a1.b1 + 2^-11 (a1.b2 + a2.b1)) / (s t_j)
// Three tcgen05.mma.kind::f16 per K = 16 instead of three kind::tf32 per K = 8: half the tensor time, same 2^-22 accuracy.
// One accumulator: the cross terms of a K chunk are issued first, then the first main MMA reads the accumulator scaled by 2^-11
// (scale-input-d), then the remaining main MMAs.  K block = 64: raw FP32 A tile of 32 KB (two TMA boxes), two FP16 B planes of
// BN x 128 B.  A stage is released after the main pass of its chunk, so STAGES >= 2 * KCHUNK_H and TA >= KCHUNK_H.
constexpr int BKH = 64;          // K per block in the FP16 kernel
constexpr int A_TILE_BYTES_H = 2 * A_TILE_BYTES;
constexpr int NTHREADS_H = 512;  // warps 0-3 TMA / MMA / TMEM alloc / idle, warps 4-11 A transform (two threads per row), warps 12-15 epilogue

// KC = k-blocks per accumulator lifetime (2: 24 accumulates, as in the TF32 kernel's 4 x 32)
template <int BN, int STAGES, int TA, int KC>
struct SmemLayoutH {
    static constexpr int B_TILE_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_TILE_BYTES_H + 2 * B_TILE_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static constexpr int A_BASE = BN <= 64 ? 128 : 256;
    static constexpr int NUM_BARS = 3 * STAGES + 4 + TA;
    static_assert(A_BASE + 64 * TA <= 512 && 2 * BN <= A_BASE, "tensor memory budget");
    static_assert(TOTAL <= 227 * 1024 && 8 * NUM_BARS + 8 <= 256, "shared memory budget");
    static_assert(STAGES >= KC + 1 && TA >= KC, "a chunk's stages stay resident until its main pass");
};

// power-of-two row scale: 2^14 * 2^floor(log2 f) (f = the row's normaliser 1 / sum, so that every entry * s <= 2^14)
__device__ __forceinline__ float row_scale_pow2(const float* a_scale, long long row, bool ok) {
    float f = (a_scale && ok) ? a_scale[row] : 1.f;
    if (!(f > 1e-30f && f < 1e30f)) f = 1.f;          // dead rows (sum 0: normaliser inf / nan) hold zeros; any scale will do
    return __uint_as_float(__float_as_uint(f) & 0x7f800000u) * 16384.f;
}

template <int BN, int STAGES, int TA, int KC>
__global__ void __launch_bounds__(NTHREADS_H, 1)
umma_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                const __grid_constant__ CUtensorMap tmB2, const GemmParams p, const float* __restrict__ a_scale,
                const float* __restrict__ binv) {
    using L = SmemLayoutH<BN, STAGES, TA, KC>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

    auto a_raw = [&](int s) { return base + s * L::STAGE_BYTES; };
    auto b_1 = [&](int s) { return base + s * L::STAGE_BYTES + A_TILE_BYTES_H; };
    auto b_2 = [&](int s) { return base + s * L::STAGE_BYTES + A_TILE_BYTES_H + L::B_TILE_BYTES; };
    const uint32_t bars = base + L::BAR_OFFSET;
    auto bar_full = [&](int s) { return bars + 8 * s; };
    auto bar_ready = [&](int s) { return bars + 8 * (STAGES + s); };
    auto bar_empty = [&](int s) { return bars + 8 * (2 * STAGES + s); };
    auto bar_tfull = [&](int a) { return bars + 8 * (3 * STAGES + a); };
    auto bar_tempty = [&](int a) { return bars + 8 * (3 * STAGES + 2 + a); };
    auto bar_aempty = [&](int a) { return bars + 8 * (3 * STAGES + 4 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + L::BAR_OFFSET + 8 * L::NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB1) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB2) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_ready(s), 8); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 128); }
        for (int a = 0; a < TA; ++a) mbar_init(bar_aempty(a), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // 512 threads start with 128 registers each; the epilogue keeps BN partial sums in registers and takes what the others give up
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            uint32_t it = 0;
            for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
                int mt, nt, kbeg, kend, part;
                unit_to_tile(p, u, mt, nt);
                unit_k_range(p, u, kbeg, kend, part);
                for (int kb = kbeg; kb < kend; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_full(s), A_TILE_BYTES_H + 2 * L::B_TILE_BYTES);
                    tma_load_2d(a_raw(s), &tmA, kb * BKH, mt * BM, bar_full(s));
                    tma_load_2d(a_raw(s) + A_TILE_BYTES, &tmA, kb * BKH + BK, mt * BM, bar_full(s));
                    tma_load_2d(b_1(s), &tmB1, kb * BKH, nt * BN, bar_full(s));
                    tma_load_2d(b_2(s), &tmB2, kb * BKH, nt * BN, bar_full(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_f16<BN>();
        uint32_t it = 0, ccount = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int kbeg, kend, part;
            unit_k_range(p, u, kbeg, kend, part);
            for (int kb0 = kbeg; kb0 < kend; kb0 += KC, ++ccount) {
                const int acc = ccount & 1;
                const uint32_t aph = (ccount >> 1) & 1;
                mbar_wait(bar_tempty(acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int nkb = min(kend - kb0, KC);
                // cross terms a1.b2 + a2.b1 of the whole chunk
                for (int j = 0; j < nkb; ++j) {
                    const uint32_t itj = it + j;
                    const int s = itj % STAGES;
                    mbar_wait(bar_ready(s), (itj / STAGES) & 1);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t a1 = tmem_base + L::A_BASE + 64 * (itj % TA), a2 = a1 + 32;
                        const uint64_t db1 = make_sw128_desc(b_1(s)), db2 = make_sw128_desc(b_2(s));
#pragma unroll
                        for (int kk = 0; kk < BKH / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 / 16);       // +32 B per K = 16 step inside the swizzle row
                            tc_mma_f16_ts(d_tmem, a1 + kk * 8, db2 + adv, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                            tc_mma_f16_ts(d_tmem, a2 + kk * 8, db1 + adv, idesc, 1u);
                        }
                    }
                    __syncwarp();
                }
                // main terms: D = a1.b1 + D * 2^-11 first, plain accumulation afterwards
                for (int j = 0; j < nkb; ++j) {
                    const uint32_t itj = it + j;
                    const int s = itj % STAGES;
                    if (elect_one_sync()) {
                        const int ta = itj % TA;
                        const uint32_t a1 = tmem_base + L::A_BASE + 64 * ta;
                        const uint64_t db1 = make_sw128_desc(b_1(s));
#pragma unroll
                        for (int kk = 0; kk < BKH / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 / 16);
                            if (j == 0 && kk == 0) tc_mma_f16_ts_scale11(d_tmem, a1, db1, idesc);
                            else tc_mma_f16_ts(d_tmem, a1 + kk * 8, db1 + adv, idesc, 1u);
                        }
                        tc_commit(bar_empty(s));
                        tc_commit(bar_aempty(ta));
                        if (j == nkb - 1) tc_commit(bar_tfull(acc));
                    }
                    __syncwarp();
                }
                it += nkb;
            }
        }
    }
    } else if (warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        // ===================== A transform: raw FP32 row (shared memory) -> a1 / a2 FP16 pairs (tensor memory) =====================
        // thread (t, sub): row t of the tile (TMEM lane t: warps 4-7 and 8-11 both map warp % 4 to the lane quarter), 32-float half sub
        const int t = (threadIdx.x - 128) & 127;
        const int sub = (threadIdx.x - 128) >> 7;
        const uint8_t* rowp = gbase + t * 128 + sub * A_TILE_BYTES;
        const int sw = t & 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int mt, nt, kbeg, kend, part;
            unit_to_tile(p, u, mt, nt);
            unit_k_range(p, u, kbeg, kend, part);
            const long long row = (long long)mt * BM + t;
            const float sc = row_scale_pow2(a_scale, row, row < p.n_rows);
            for (int kb = kbeg; kb < kend; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int ta = it % TA;
                mbar_wait(bar_full(s), ph);
                mbar_wait(bar_aempty(ta), ((it / TA) & 1) ^ 1);
                tc_fence_after();
                const uint32_t th = tmem_base + lane_base + L::A_BASE + 64 * ta;
                {
                    const uint8_t* src = rowp + s * L::STAGE_BYTES;
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(src + ((c ^ sw) << 4));
                        const float x0 = v.x * sc, x1 = v.y * sc, x2 = v.z * sc, x3 = v.w * sc;
                        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn((x0 - f01.x) * 2048.f, (x1 - f01.y) * 2048.f);
                        const __half2 l23 = __floats2half2_rn((x2 - f23.x) * 2048.f, (x3 - f23.y) * 2048.f);
                        h[2 * c] = *reinterpret_cast<const uint32_t*>(&h01);
                        h[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                        l[2 * c] = *reinterpret_cast<const uint32_t*>(&l01);
                        l[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                    }
                    tc_st16(th + sub * 16, h);
                    tc_st16(th + 32 + sub * 16, l);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready(s));
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ===================== epilogue: FP32 register accumulation + segmented max / argmax =====================
        const int q = warp & 3;
        const bool direct = p.ws == nullptr;
        uint32_t ccount = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
            int mt, nt;
            unit_to_tile(p, u, mt, nt);
            const long long row = (long long)mt * BM + q * 32 + lane;
            const bool row_ok = row < p.n_rows;
            const int col0 = nt * BN;
            int cur_seg = col0 / p.seg_stride, r = col0 - cur_seg * p.seg_stride;
            float best = -INFINITY;
            int best_idx = 0;
            bool dirty = false;
            float sum[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] = 0.f;
            int kbeg, kend, part;
            unit_k_range(p, u, kbeg, kend, part);
            for (int kb0 = kbeg; kb0 < kend; kb0 += KC, ++ccount) {
                const int acc = ccount & 1;
                const uint32_t aph = (ccount >> 1) & 1;
                mbar_wait(bar_tfull(acc), aph);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                mbar_arrive(bar_tempty(acc));
            }
            // undo the block scales (powers of two: exact)
            const float inv_s = 1.f / row_scale_pow2(a_scale, row, row_ok);
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] *= inv_s * ((col0 + j < p.rows_b) ? __ldg(binv + col0 + j) : 0.f);
            if (part >= 0) {
                float* wsrow = p.tail_ws + ((((long long)(unit_tile_index(p, u) - p.full_units) * p.tail_split + part) * BM) + q * 32 + lane) * BN;
#pragma unroll
                for (int j = 0; j < BN; j += 4) *reinterpret_cast<float4*>(wsrow + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
                continue;
            }
            auto flush = [&]() {
                if (row_ok && cur_seg < p.n_seg) {
                    if (direct) {
                        p.out_val[row * p.n_seg + cur_seg] = best;
                        p.out_arg[row * p.n_seg + cur_seg] = best_idx;
                    } else if (dirty) {
                        atomicMax(p.ws + row * p.n_seg + cur_seg, pack_key(best, best_idx));
                    }
                }
            };
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                const float x = sum[j];
                if (r < p.seg_len && cur_seg < p.n_seg) {
                    dirty = true;
                    if (x > best) { best = x; best_idx = r; }
                }
                if (++r == p.seg_stride) {
                    flush();
                    r = 0; ++cur_seg; best = -INFINITY; best_idx = 0; dirty = false;
                }
            }
            if (r != 0 && !direct) flush();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// ============================================================================================================
// CTA-PAIR variant of the FP16 split kernel (cta_group::2) for the large backup GEMMs.  umma_f16_kernel<128,...> pulls 64 KB through
// L2 -> shared memory per K block and SM for 768 tensor-core cycles: at full tensor rate that is 24.8 TB/s over the chip, twice what
// the L2 delivers (~6 300 B/clk = 12.4 TB/s: ncu showed 12.6 TB/s at 53 % tensor-pipe activity — the kernel sits on that ceiling).
// Two CTAs of a cluster (the two SMs of a TPC) share one tile of B: the work unit is 256 rows x BN columns, each CTA loads and
// transforms its own 128 rows of A (raw FP32 -> FP16 pairs in ITS tensor memory), loads HALF of the two B planes (BN / 2 rows) into
// ITS shared memory, the leader issues ONE tcgen05.mma.cta_group::2 of M = 256 per step, and each CTA's accumulator (its 128 rows x
// BN columns) lands in its own tensor memory: 48 KB per K block and SM instead of 64 (-25 % L2 traffic per flop).
// Barriers: fullA / empty / aempty / tfull are per CTA (tcgen05.commit multicasts to both); fullB / ready / tempty live in the
// leader and the peer's TMA engine, transform warps and epilogue warps arrive on them through the cluster address space.
// Two K blocks per accumulator lifetime (24 MMAs, as in the narrow single-CTA configurations), no K split of a last partial wave (the caller only takes this kernel when there
// are many waves).
template <int BN, int STAGES, int TA, int KC>
struct SmemLayoutP {
    static constexpr int B_HALF_BYTES = (BN / 2) * 128;
    static constexpr int STAGE_BYTES = A_TILE_BYTES_H + 2 * B_HALF_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
    static constexpr int A_BASE = 256;
    static constexpr int NUM_BARS = 4 * STAGES + 4 + TA;
    static_assert(A_BASE + 64 * TA <= 512 && 2 * BN <= A_BASE, "tensor memory budget");
    static_assert(TOTAL <= 227 * 1024 && 8 * NUM_BARS + 8 <= 256, "shared memory budget");
    static_assert(STAGES >= KC + 1 && TA >= KC, "a chunk's stages stay resident until its main pass");
};

template <int BN, int STAGES, int TA, int KC>
__global__ void __launch_bounds__(NTHREADS_H, 1)
umma_f16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                     const __grid_constant__ CUtensorMap tmB2, const GemmParams p, const float* __restrict__ a_scale,
                     const float* __restrict__ binv) {
    using L = SmemLayoutP<BN, STAGES, TA, KC>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));

    auto a_raw = [&](int s) { return base + s * L::STAGE_BYTES; };
    auto b_1 = [&](int s) { return base + s * L::STAGE_BYTES + A_TILE_BYTES_H; };
    auto b_2 = [&](int s) { return base + s * L::STAGE_BYTES + A_TILE_BYTES_H + L::B_HALF_BYTES; };
    const uint32_t bars = base + L::BAR_OFFSET;
    auto bar_fullA = [&](int s) { return bars + 8 * s; };
    auto bar_fullB = [&](int s) { return bars + 8 * (STAGES + s); };                 // leader's copy is the one in use
    auto bar_ready = [&](int s) { return bars + 8 * (2 * STAGES + s); };             // leader's copy
    auto bar_empty = [&](int s) { return bars + 8 * (3 * STAGES + s); };
    auto bar_tfull = [&](int a) { return bars + 8 * (4 * STAGES + a); };
    auto bar_tempty = [&](int a) { return bars + 8 * (4 * STAGES + 2 + a); };        // leader's copy
    auto bar_aempty = [&](int a) { return bars + 8 * (4 * STAGES + 4 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + L::BAR_OFFSET + 8 * L::NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB1) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmB2) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_fullA(s), 1); mbar_init(bar_fullB(s), 1); mbar_init(bar_ready(s), 16); mbar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), 8); }
        for (int a = 0; a < TA; ++a) mbar_init(bar_aempty(a), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // the peer's barriers are initialised and its tensor memory allocated before anything crosses
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own rows of A, own half of the B planes =====================
        if (elect_one_sync()) {
            const uint32_t leader_fullB0 = mapa_rank(bar_fullB(0), 0);
            uint32_t it = 0;
            for (int u = cluster_id; u < p.num_units; u += n_clusters) {
                int mt, nt;
                unit_to_tile(p, u, mt, nt);
                for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(bar_empty(s), ph ^ 1);
                    mbar_expect_tx(bar_fullA(s), A_TILE_BYTES_H);
                    tma_load_2d(a_raw(s), &tmA, kb * BKH, mt * 2 * BM + (int)rank * BM, bar_fullA(s));
                    tma_load_2d(a_raw(s) + A_TILE_BYTES, &tmA, kb * BKH + BK, mt * 2 * BM + (int)rank * BM, bar_fullA(s));
                    if (leader) mbar_expect_tx(bar_fullB(s), 4 * L::B_HALF_BYTES);          // two planes x two CTAs
                    tma_load_2d_pair(b_1(s), &tmB1, kb * BKH, nt * BN + (int)rank * (BN / 2), leader_fullB0 + 8 * s);
                    tma_load_2d_pair(b_2(s), &tmB2, kb * BKH, nt * BN + (int)rank * (BN / 2), leader_fullB0 + 8 * s);
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ===================== MMA issuer (leader only): M = 256 over both CTAs =====================
        // KC K blocks per accumulator lifetime: the hand-over of an accumulator crosses the pair twice (commit -> peer epilogue ->
        // remote arrive), about as long as the MMAs of one K block take; with one block per lifetime the pair ran at 1.1 us per
        // block against 0.73 us for the single-CTA kernel.
        constexpr uint32_t idesc = make_idesc_f16_pair<BN>();
        uint32_t it = 0, ccount = 0;
        for (int u = cluster_id; u < p.num_units; u += n_clusters) {
            for (int kb0 = 0; kb0 < p.num_k_blocks; kb0 += KC, ++ccount) {
                const int acc = ccount & 1;
                const int nkb = min(p.num_k_blocks - kb0, KC);
                mbar_wait_cluster(bar_tempty(acc), ((ccount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                // cross terms a1.b2 + a2.b1 of the whole chunk
                for (int j = 0; j < nkb; ++j) {
                    const uint32_t itj = it + j;
                    const int s = itj % STAGES;
                    mbar_wait_cluster(bar_ready(s), (itj / STAGES) & 1);
                    mbar_wait_cluster(bar_fullB(s), (itj / STAGES) & 1);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t a1 = tmem_base + L::A_BASE + 64 * (itj % TA), a2 = a1 + 32;
                        const uint64_t db1 = make_sw128_desc(b_1(s)), db2 = make_sw128_desc(b_2(s));
#pragma unroll
                        for (int kk = 0; kk < BKH / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 / 16);
                            tc_mma_f16_ts_pair(d_tmem, a1 + kk * 8, db2 + adv, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                            tc_mma_f16_ts_pair(d_tmem, a2 + kk * 8, db1 + adv, idesc, 1u);
                        }
                    }
                    __syncwarp();
                }
                // main terms: D = a1.b1 + D * 2^-11 first, plain accumulation afterwards
                for (int j = 0; j < nkb; ++j) {
                    const uint32_t itj = it + j;
                    const int s = itj % STAGES;
                    if (elect_one_sync()) {
                        const int ta = itj % TA;
                        const uint32_t a1 = tmem_base + L::A_BASE + 64 * ta;
                        const uint64_t db1 = make_sw128_desc(b_1(s));
#pragma unroll
                        for (int kk = 0; kk < BKH / 16; ++kk) {
                            const uint64_t adv = (uint64_t)(kk * 32 / 16);
                            if (j == 0 && kk == 0) tc_mma_f16_ts_pair_scale11(d_tmem, a1, db1, idesc);
                            else tc_mma_f16_ts_pair(d_tmem, a1 + kk * 8, db1 + adv, idesc, 1u);
                        }
                        tc_commit_pair(bar_empty(s));
                        tc_commit_pair(bar_aempty(ta));
                        if (j == nkb - 1) tc_commit_pair(bar_tfull(acc));
                    }
                    __syncwarp();
                }
                it += nkb;
            }
        }
    }
    } else if (warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        // ===================== A transform (both CTAs): own rows -> FP16 pairs in own tensor memory, arrival on the leader =====================
        const int t = (threadIdx.x - 128) & 127;
        const int sub = (threadIdx.x - 128) >> 7;
        const uint8_t* rowp = gbase + t * 128 + sub * A_TILE_BYTES;
        const int sw = t & 7;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t leader_ready0 = mapa_rank(bar_ready(0), 0);
        uint32_t it = 0;
        for (int u = cluster_id; u < p.num_units; u += n_clusters) {
            int mt, nt;
            unit_to_tile(p, u, mt, nt);
            const long long row = (long long)mt * 2 * BM + rank * BM + t;
            const float sc = row_scale_pow2(a_scale, row, row < p.n_rows);
            for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int ta = it % TA;
                mbar_wait(bar_fullA(s), ph);
                mbar_wait(bar_aempty(ta), ((it / TA) & 1) ^ 1);
                tc_fence_after();
                const uint32_t th = tmem_base + lane_base + L::A_BASE + 64 * ta;
                {
                    const uint8_t* src = rowp + s * L::STAGE_BYTES;
                    uint32_t h[16], l[16];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4*>(src + ((c ^ sw) << 4));
                        const float x0 = v.x * sc, x1 = v.y * sc, x2 = v.z * sc, x3 = v.w * sc;
                        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn((x0 - f01.x) * 2048.f, (x1 - f01.y) * 2048.f);
                        const __half2 l23 = __floats2half2_rn((x2 - f23.x) * 2048.f, (x3 - f23.y) * 2048.f);
                        h[2 * c] = *reinterpret_cast<const uint32_t*>(&h01);
                        h[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                        l[2 * c] = *reinterpret_cast<const uint32_t*>(&l01);
                        l[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                    }
                    tc_st16(th + sub * 16, h);
                    tc_st16(th + 32 + sub * 16, l);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_ready0 + 8 * s);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ===================== epilogue (both CTAs): own 128 rows of the accumulator =====================
        const int q = warp & 3;
        const bool direct = p.ws == nullptr;
        const uint32_t leader_tempty0 = mapa_rank(bar_tempty(0), 0);
        uint32_t it = 0;
        for (int u = cluster_id; u < p.num_units; u += n_clusters) {
            int mt, nt;
            unit_to_tile(p, u, mt, nt);
            const long long row = (long long)mt * 2 * BM + rank * BM + q * 32 + lane;
            const bool row_ok = row < p.n_rows;
            const int col0 = nt * BN;
            int cur_seg = col0 / p.seg_stride, r = col0 - cur_seg * p.seg_stride;
            float best = -INFINITY;
            int best_idx = 0;
            bool dirty = false;
            float sum[BN];
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] = 0.f;
            for (int kb0 = 0; kb0 < p.num_k_blocks; kb0 += KC, ++it) {
                const int acc = it & 1;
                mbar_wait(bar_tfull(acc), (it >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 16) {
                    uint32_t v[16];
                    tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8 * acc);
            }
            const float inv_s = 1.f / row_scale_pow2(a_scale, row, row_ok);
#pragma unroll
            for (int j = 0; j < BN; ++j) sum[j] *= inv_s * ((col0 + j < p.rows_b) ? __ldg(binv + col0 + j) : 0.f);
            auto flush = [&]() {
                if (row_ok && cur_seg < p.n_seg) {
                    if (direct) {
                        p.out_val[row * p.n_seg + cur_seg] = best;
                        p.out_arg[row * p.n_seg + cur_seg] = best_idx;
                    } else if (dirty) {
                        atomicMax(p.ws + row * p.n_seg + cur_seg, pack_key(best, best_idx));
                    }
                }
            };
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                const float x = sum[j];
                if (r < p.seg_len && cur_seg < p.n_seg) {
                    dirty = true;
                    if (x > best) { best = x; best_idx = r; }
                }
                if (++r == p.seg_stride) {
                    flush();
                    r = 0; ++cur_seg; best = -INFINITY; best_idx = 0; dirty = false;
                }
            }
            if (r != 0 && !direct) flush();
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // neither CTA leaves (or frees its tensor memory) while the other may still touch it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}



// ws -> (value, first index); entries no unit touched stay (-inf, 0)
__global__ void segmax_decode_kernel(long long n, const unsigned long long* __restrict__ ws, float* __restrict__ out_val,
                                     int* __restrict__ out_arg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = ws[i];
    if (k == 0ull) { out_val[i] = -INFINITY; out_arg[i] = 0; return; }
    unsigned u = (unsigned)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    out_val[i] = __uint_as_float(u);
    out_arg[i] = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
}

// Tail units: add the K parts in a fixed order, then the same epilogue as the kernel — plain store, one direct segment, or (several n
// tiles) the running segmented max combined with atomicMax on the packed keys, exactly like the kernel's own flush.
__global__ void tail_finish_kernel(const GemmParams p, int BN, int tail_units) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // (tail unit, row of the tile)
    if (i >= (long long)tail_units * BM) return;
    const int t = (int)(i / BM), rr = (int)(i - (long long)t * BM);
    int mt, nt;
    GemmParams q = p;
    q.tail_split = 1;                                                             // decode a plain unit index
    unit_to_tile(q, p.full_units + t, mt, nt);
    const long long row = (long long)mt * BM + rr;
    if (row >= p.n_rows) return;
    const float* w = p.tail_ws + ((long long)t * p.tail_split * BM + rr) * BN;
    const int col0 = nt * BN;
    const bool direct = p.ws == nullptr;
    int cur_seg = col0 / p.seg_stride, r = col0 - cur_seg * p.seg_stride;
    float best = -INFINITY;
    int best_idx = 0;
    bool dirty = false;
    for (int j = 0; j < BN; ++j) {
        float x = 0.f;
        for (int k = 0; k < p.tail_split; ++k) x += w[(long long)k * BM * BN + j];
        if (p.out_store) { if (col0 + j < p.rows_b) p.out_store[row * p.out_ld + col0 + j] = x; continue; }
        if (r < p.seg_len && cur_seg < p.n_seg) {
            dirty = true;
            if (x > best) { best = x; best_idx = r; }
        }
        if (++r == p.seg_stride) {
            if (cur_seg < p.n_seg) {
                if (direct) { p.out_val[row * p.n_seg + cur_seg] = best; p.out_arg[row * p.n_seg + cur_seg] = best_idx; }
                else if (dirty) atomicMax(p.ws + row * p.n_seg + cur_seg, pack_key(best, best_idx));
            }
            r = 0; ++cur_seg; best = -INFINITY; best_idx = 0; dirty = false;
        }
    }
    if (!p.out_store && r != 0 && !direct && dirty && cur_seg < p.n_seg) atomicMax(p.ws + row * p.n_seg + cur_seg, pack_key(best, best_idx));
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map_2d(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_rows,
                CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { olfnav_set_error("umma: cuTensorMapEncodeTiled not available"); return OLFNAV_ERR_CUDA; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { olfnav_set_error("umma: cuTensorMapEncodeTiled failed (%d)", (int)r); return OLFNAV_ERR_CUDA; }
    return OLFNAV_OK;
}

template <int BN, int STAGES, int TA>
int launch_ts(const CUtensorMap& tA, const CUtensorMap& tBh, const CUtensorMap& tBl, const GemmParams& p, int grid, cudaStream_t st) {
    using L = SmemLayoutTS<BN, STAGES, TA>;
    static OlfPerDeviceOnce configured;
    if (configured.first())
        OLF_CUDA(cudaFuncSetAttribute(umma_ts_kernel<BN, STAGES, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    umma_ts_kernel<BN, STAGES, TA><<<grid, NTHREADS, L::TOTAL, st>>>(tA, tBh, tBl, p);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

}  // namespace

int olf_make_map_2d(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_rows) {
    return make_map_2d(map, ptr, inner, outer, row_stride_bytes, box_rows);
}

static int umma_run(const float* A, int64_t lda, int64_t n_rows, int K, const float* B_hi, const float* B_lo, int64_t ldb,
                    int rows_b_stored, GemmParams p, int sm_count, cudaStream_t st) {
    // Tile width: exactly ceil16(rows_b) when it fits one tile (no MMA work on padding columns), else 128-wide tiles.
    // <= 128 because the epilogue keeps BN FP32 partial sums in registers.
    const int BN = rows_b_stored <= 128 ? ((rows_b_stored + 15) & ~15) : 128;
    p.n_rows = (int)n_rows; p.K = K;
    p.num_m_tiles = olf_div_up(n_rows, BM);
    p.num_n_tiles = olf_div_up(rows_b_stored, BN);
    p.num_k_blocks = olf_div_up(K, BK);
    p.num_units = p.num_m_tiles * p.num_n_tiles;
    p.group_w = p.num_n_tiles < 4 ? p.num_n_tiles : 4;
    p.ws = nullptr;
    const size_t n_out = (size_t)n_rows * p.n_seg;
    OlfScratch ws(st), tail_ws(st);
    if (!p.out_store && p.num_n_tiles > 1) {
        OLF_CUDA(ws.alloc(n_out * sizeof(unsigned long long)));
        p.ws = ws.as<unsigned long long>();
        OLF_CUDA(cudaMemsetAsync(p.ws, 0, n_out * sizeof(unsigned long long), st));
    }

    CUtensorMap tA, tBh, tBl;
    // L2 promotion of the A-operand requests: 256 B (the next k-block's 128 bytes of every row arrive with this one's) streams 3 %
    // faster than 128 B at G <= 36 (92 % vs 89 % of HBM, measured); OLFNAV_GEMM_PROMO=0|64|128|256 overrides
    static const CUtensorMapL2promotion a_promo = [] {
        const char* e = getenv("OLFNAV_GEMM_PROMO");
        const int v = e ? atoi(e) : 256;
        return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
    }();
    int rc = make_map_2d(&tA, A, (uint64_t)K, (uint64_t)n_rows, (uint64_t)lda * 4, BM, a_promo);
    if (rc == OLFNAV_OK) rc = make_map_2d(&tBh, B_hi, (uint64_t)ldb, (uint64_t)rows_b_stored, (uint64_t)ldb * 4, BN);
    if (rc == OLFNAV_OK) rc = make_map_2d(&tBl, B_lo, (uint64_t)ldb, (uint64_t)rows_b_stored, (uint64_t)ldb * 4, BN);
    if (rc != OLFNAV_OK) return rc;

    int grid = p.num_units < sm_count ? p.num_units : sm_count;
    // tail split: more than one wave of units and a last wave that is at most half full
    p.full_units = p.num_units; p.tail_split = 1; p.tail_ws = nullptr;
    int tail_tiles = 0;
    static const bool tail_off = [] { const char* e = getenv("OLFNAV_GEMM_TAIL"); return e && e[0] == '0'; }();
    if (!tail_off && p.num_units > sm_count && p.num_k_blocks >= 64) {
        const int tail = p.num_units % sm_count;
        const int split = tail ? (sm_count / tail < 4 ? sm_count / tail : 4) : 1;
        if (tail && split > 1) {
            tail_tiles = tail;
            p.full_units = p.num_units - tail;
            p.tail_split = split;
            p.num_units = p.full_units + tail * split;
            OLF_CUDA(tail_ws.alloc((size_t)tail * split * BM * BN * sizeof(float)));
            p.tail_ws = tail_ws.as<float>();
        }
    }
    // A in tensor memory: shared-memory stage = 16 KB raw A + 2 B planes of BN x 128 B (as many as fit ~216 KB);
    // TMEM holds 64 columns (hi, lo) per A slot next to the two accumulator buffers (512 columns in all)
    switch (BN) {
        case 16:  rc = launch_ts<16, 6, 6>(tA, tBh, tBl, p, grid, st); break;
        case 32:  rc = launch_ts<32, 6, 6>(tA, tBh, tBl, p, grid, st); break;
        case 48:  rc = launch_ts<48, 6, 6>(tA, tBh, tBl, p, grid, st); break;
        case 64:  rc = launch_ts<64, 6, 6>(tA, tBh, tBl, p, grid, st); break;
        case 80:  rc = launch_ts<80, 6, 4>(tA, tBh, tBl, p, grid, st); break;
        case 96:  rc = launch_ts<96, 5, 4>(tA, tBh, tBl, p, grid, st); break;
        case 112: rc = launch_ts<112, 4, 4>(tA, tBh, tBl, p, grid, st); break;
        default:  rc = launch_ts<128, 4, 4>(tA, tBh, tBl, p, grid, st); break;
    }
    if (p.tail_ws) {
        if (rc == OLFNAV_OK) {
            tail_finish_kernel<<<olf_div_up((int64_t)tail_tiles * BM, 128), 128, 0, st>>>(p, BN, tail_tiles);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("umma: tail finish kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
    }
    if (p.ws) {
        if (rc == OLFNAV_OK) {
            segmax_decode_kernel<<<olf_div_up((int64_t)n_out, 256), 256, 0, st>>>((long long)n_out, p.ws, p.out_val, p.out_arg);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("umma_segmax: decode kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
    }
    return rc;
}

int olf_umma_segmax(const float* A, int64_t lda, int64_t n_rows, int K, const float* B_hi, const float* B_lo, int64_t ldb,
                    int rows_b, int n_seg, int seg_len, int seg_stride, float* out_val, int* out_arg, int sm_count,
                    cudaStream_t st) {
    OLF_REQUIRE(A && B_hi && B_lo && out_val && out_arg, "umma_segmax: null argument");
    OLF_REQUIRE((lda % 4) == 0 && (((uintptr_t)A) & 15) == 0, "umma_segmax: A must be 16-byte aligned with lda %% 4 == 0");
    OLF_REQUIRE((ldb % 32) == 0 && ldb >= K, "umma_segmax: ldb must be a multiple of 32 and >= K");
    OLF_REQUIRE((seg_stride % 16) == 0 && seg_len <= seg_stride && rows_b == n_seg * seg_stride, "umma_segmax: bad segment layout");
    OLF_REQUIRE(n_rows > 0 && K > 0, "umma_segmax: empty problem");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.rows_b = rows_b;
    p.n_seg = n_seg; p.seg_len = seg_len; p.seg_stride = seg_stride;
    p.out_val = out_val; p.out_arg = out_arg;
    return umma_run(A, lda, n_rows, K, B_hi, B_lo, ldb, rows_b, p, sm_count, st);
}

int olf_umma_store(const float* A, int64_t lda, int64_t n_rows, int K, const float* B_hi, const float* B_lo, int64_t ldb,
                   int rows_b, float* out, int64_t out_ld, int sm_count, cudaStream_t st) {
    OLF_REQUIRE(A && B_hi && B_lo && out, "umma_store: null argument");
    OLF_REQUIRE((lda % 4) == 0 && (((uintptr_t)A) & 15) == 0, "umma_store: A must be 16-byte aligned with lda %% 4 == 0");
    OLF_REQUIRE((ldb % 32) == 0 && ldb >= K, "umma_store: ldb must be a multiple of 32 and >= K");
    OLF_REQUIRE(rows_b >= 1 && rows_b <= 128 && out_ld >= rows_b, "umma_store: 1 <= rows_b <= 128 and out_ld >= rows_b");
    OLF_REQUIRE(n_rows > 0 && K > 0, "umma_store: empty problem");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.rows_b = rows_b;
    p.n_seg = 1; p.seg_len = rows_b; p.seg_stride = (rows_b + 15) & ~15;
    p.out_store = out; p.out_ld = out_ld;
    return umma_run(A, lda, n_rows, K, B_hi, B_lo, ldb, (rows_b + 15) & ~15, p, sm_count, st);
}

namespace {
// FP16 2-way split planes with a power-of-two scale per row (umma_gemm.cu, umma_f16_kernel): one block per destination row.
//   t = 2^(13 - floor(log2 max|v|)) so that |t v| < 2^14;  b1 = fp16(t v), b2 = fp16((t v - b1) 2^11), binv = 1 / t
__global__ void split_planes_f16_kernel(const float* __restrict__ src, long long ld_src, long long rows, int K,
                                        __half* __restrict__ b1, __half* __restrict__ b2, float* __restrict__ binv, long long ld) {
    __shared__ float red[8];
    const long long r = blockIdx.y;
    const long long dst = (long long)blockIdx.z * gridDim.y + r;
    const float* srow = src + ((long long)blockIdx.z * rows + r) * ld_src;
    const bool live = r < rows;
    float mx = 0.f;
    if (live)
        for (int c = threadIdx.x; c < K; c += blockDim.x) mx = fmaxf(mx, fabsf(srow[c]));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    const float p2 = __uint_as_float(__float_as_uint(mx) & 0x7f800000u);          // 2^floor(log2 max) (0 for zero / subnormal rows)
    const float t = p2 > 0.f ? 8192.f / p2 : 1.f;
    if (threadIdx.x == 0) binv[dst] = 1.f / t;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
        const float y = (live && c < K) ? srow[c] * t : 0.f;
        const __half h = __float2half_rn(y);
        b1[dst * ld + c] = h;
        b2[dst * ld + c] = __float2half_rn((y - __half2float(h)) * 2048.f);
    }
}

}  // namespace

// rows of every segment padded to rows_padded destination rows; nseg segments of `rows` source rows each
int olf_split_planes_f16(const float* src, int64_t ld_src, int64_t rows, int K, void* b1, void* b2, float* binv, int64_t ldh,
                         int rows_padded, int nseg, cudaStream_t st) {
    OLF_REQUIRE(src && b1 && b2 && binv && (ldh % 64) == 0 && ldh >= K && rows_padded >= rows, "split_planes_f16: bad argument");
    dim3 grid(1, rows_padded, nseg);
    split_planes_f16_kernel<<<grid, 256, 0, st>>>(src, ld_src, rows, K, (__half*)b1, (__half*)b2, binv, ldh);
    olf_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("split_planes_f16: launch failed"); return OLFNAV_ERR_CUDA; }
    return OLFNAV_OK;
}

// FP16 split variant (see umma_f16_kernel): B1 / B2 are FP16 planes [rows_b x ldb] (ldb in halfs, multiple of 64, zero beyond K),
// binv[j] = 1 / t_j the inverse power-of-two scale of row j of B, a_scale (may be NULL) the per-row normaliser of A (1 / row sum).
int olf_umma_segmax_f16(const float* A, int64_t lda, int64_t n_rows, int K, const float* a_scale,
                        const void* B1, const void* B2, const float* binv, int64_t ldb, int rows_b,
                        int n_seg, int seg_len, int seg_stride, float* out_val, int* out_arg, int sm_count, cudaStream_t st) {
    OLF_REQUIRE(A && B1 && B2 && binv && out_val && out_arg, "umma_segmax_f16: null argument");
    OLF_REQUIRE((lda % 4) == 0 && (((uintptr_t)A) & 15) == 0, "umma_segmax_f16: A must be 16-byte aligned with lda %% 4 == 0");
    OLF_REQUIRE((ldb % 64) == 0 && ldb >= K, "umma_segmax_f16: ldb must be a multiple of 64 and >= K");
    OLF_REQUIRE((seg_stride % 16) == 0 && seg_len <= seg_stride && rows_b == n_seg * seg_stride, "umma_segmax_f16: bad segment layout");
    OLF_REQUIRE(n_rows > 0 && K > 0, "umma_segmax_f16: empty problem");
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.rows_b = rows_b;
    p.n_seg = n_seg; p.seg_len = seg_len; p.seg_stride = seg_stride;
    p.out_val = out_val; p.out_arg = out_arg;
    static const int wide_bn = [] { const char* e = getenv("OLFNAV_GEMM_F16_BN"); const int v = e ? atoi(e) : 128; return (v == 80 || v == 96) ? v : 128; }();
    const int BN = rows_b <= 80 ? ((rows_b + 15) & ~15) : wide_bn;
    p.n_rows = (int)n_rows; p.K = K;
    // CTA pairs (umma_f16_pair_kernel: 256-row tiles, half of B per SM) where the problem is many waves of 128-column tiles, i.e. the
    // backup selection of a large belief set
    // OFF by default: measured 16.1 ms against 11.9 ms for the backup selection at B = 10 000, G = 512 (tensor pipe 38 % against 59 %,
    // the pair halves B traffic but keeps the same 192 KB in flight per SM over a longer load path — latency bound).
    // OLFNAV_GEMM_PAIR=1 takes it where it applies, =2 makes the call fail when it does not (tests use it to know which kernel ran).
    const char* pair_env = getenv("OLFNAV_GEMM_PAIR");
    const int pair_mode = pair_env ? atoi(pair_env) : 0;
    const bool pair = pair_mode != 0 && BN == 128 && (sm_count % 2) == 0 && olf_div_up(K, BKH) >= 32 &&
                      (int64_t)olf_div_up(n_rows, 2 * BM) * olf_div_up(rows_b, BN) >= 4 * (sm_count / 2);
    OLF_REQUIRE(pair || pair_mode != 2, "umma_segmax_f16: OLFNAV_GEMM_PAIR=2 but the CTA-pair kernel does not apply to this problem");
    p.num_m_tiles = olf_div_up(n_rows, pair ? 2 * BM : BM);
    p.num_n_tiles = olf_div_up(rows_b, BN);
    p.num_k_blocks = olf_div_up(K, BKH);
    p.num_units = p.num_m_tiles * p.num_n_tiles;
    p.group_w = p.num_n_tiles < 4 ? p.num_n_tiles : 4;
    const size_t n_out = (size_t)n_rows * p.n_seg;
    OlfScratch ws(st), tail_ws(st);
    if (p.num_n_tiles > 1) {
        OLF_CUDA(ws.alloc(n_out * sizeof(unsigned long long)));
        p.ws = ws.as<unsigned long long>();
        OLF_CUDA(cudaMemsetAsync(p.ws, 0, n_out * sizeof(unsigned long long), st));
    }
    CUtensorMap tA, tB1, tB2;
    int rc = make_map_2d(&tA, A, (uint64_t)K, (uint64_t)n_rows, (uint64_t)lda * 4, BM, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    EncodeTiledFn enc = get_encode();
    for (int k = 0; k < 2 && rc == OLFNAV_OK; ++k) {
        cuuint64_t dims[2] = {(cuuint64_t)ldb, (cuuint64_t)rows_b};
        cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
        cuuint32_t box[2] = {(cuuint32_t)BKH, (cuuint32_t)(pair ? BN / 2 : BN)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(k ? &tB2 : &tB1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(k ? B2 : B1), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { olfnav_set_error("umma_segmax_f16: cuTensorMapEncodeTiled failed (%d)", (int)r); rc = OLFNAV_ERR_CUDA; }
    }
    if (rc != OLFNAV_OK) return rc;

    int grid = p.num_units < sm_count ? p.num_units : sm_count;
    p.full_units = p.num_units; p.tail_split = 1; p.tail_ws = nullptr;
    int tail_tiles = 0;
    if (pair) {
        using L = SmemLayoutP<128, 4, 4, 2>;
        static OlfPerDeviceOnce configured;
        if (configured.first() &&
            cudaFuncSetAttribute(umma_f16_pair_kernel<128, 4, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) {
            olfnav_set_error("umma_segmax_f16: cudaFuncSetAttribute (pair kernel) failed");
            return OLFNAV_ERR_CUDA;
        }
        const int clusters = p.num_units < sm_count / 2 ? p.num_units : sm_count / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * clusters, 1, 1);
        cfg.blockDim = dim3(NTHREADS_H, 1, 1);
        cfg.dynamicSmemBytes = L::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, umma_f16_pair_kernel<128, 4, 4, 2>, tA, tB1, tB2, p, a_scale, binv) != cudaSuccess) {
            olfnav_set_error("umma_segmax_f16: pair kernel launch failed (%s)", cudaGetErrorString(cudaGetLastError()));
            return OLFNAV_ERR_CUDA;
        }
        olf_count_launch(1);
        if (p.ws) {
            segmax_decode_kernel<<<olf_div_up((int64_t)n_out, 256), 256, 0, st>>>((long long)n_out, p.ws, p.out_val, p.out_arg);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("umma_segmax_f16: decode kernel launch failed"); return OLFNAV_ERR_CUDA; }
        }
        return OLFNAV_OK;
    }
    if (p.num_units > sm_count && p.num_k_blocks >= 32) {
        const int tail = p.num_units % sm_count;
        const int split = tail ? (sm_count / tail < 4 ? sm_count / tail : 4) : 1;
        if (tail && split > 1) {
            tail_tiles = tail;
            p.full_units = p.num_units - tail;
            p.tail_split = split;
            p.num_units = p.full_units + tail * split;
            OLF_CUDA(tail_ws.alloc((size_t)tail * split * BM * BN * sizeof(float)));
            p.tail_ws = tail_ws.as<float>();
        }
    }
#define OLF_LAUNCH_H(BN_, ST_, TA_, KC_)                                                                                             \
    do {                                                                                                                         \
        using L = SmemLayoutH<BN_, ST_, TA_, KC_>;                                                                                   \
        static OlfPerDeviceOnce configured;                                                                                      \
        if (configured.first() &&                                                                                                \
            cudaFuncSetAttribute(umma_f16_kernel<BN_, ST_, TA_, KC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) { \
            olfnav_set_error("umma_segmax_f16: cudaFuncSetAttribute failed"); rc = OLFNAV_ERR_CUDA; break;                       \
        }                                                                                                                        \
        umma_f16_kernel<BN_, ST_, TA_, KC_><<<grid, NTHREADS_H, L::TOTAL, st>>>(tA, tB1, tB2, p, a_scale, binv);                        \
        olf_count_launch(1);                                                                                                     \
        if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("umma_segmax_f16: launch failed"); rc = OLFNAV_ERR_CUDA; }     \
    } while (0)
    switch (BN) {
        case 16: OLF_LAUNCH_H(16, 5, 5, 2); break;
        case 32: OLF_LAUNCH_H(32, 5, 5, 2); break;
        case 48: OLF_LAUNCH_H(48, 5, 5, 2); break;
        case 64: OLF_LAUNCH_H(64, 4, 4, 2); break;
        case 80: OLF_LAUNCH_H(80, 4, 4, 2); break;
        case 96: OLF_LAUNCH_H(96, 4, 4, 2); break;
        default: OLF_LAUNCH_H(128, 3, 3, 1); break;
    }
#undef OLF_LAUNCH_H
    if (p.tail_ws) {
        if (rc == OLFNAV_OK) {
            tail_finish_kernel<<<olf_div_up((int64_t)tail_tiles * BM, 128), 128, 0, st>>>(p, BN, tail_tiles);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("umma_segmax_f16: tail finish kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
    }
    if (p.ws) {
        if (rc == OLFNAV_OK) {
            segmax_decode_kernel<<<olf_div_up((int64_t)n_out, 256), 256, 0, st>>>((long long)n_out, p.ws, p.out_val, p.out_arg);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("umma_segmax_f16: decode kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
    }
    return rc;
}
