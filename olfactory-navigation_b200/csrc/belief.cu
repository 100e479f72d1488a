// Belief filter kernels: init, fused belief step, pack/unpack, entropies.
//
// belief_step replaces pomdp_toolkit's BeliefSet.update as called by the reference at
// olfactory_navigation/agents/pbvi_agent.py:783-787 and agents/infotaxis_agent.py:335-339:
//     b'(s') ∝ O(o|s',a) · Σ_s T(s'|s,a) · b(s)
// For the grid model of exact_converter (agents/model_based_util/environment_converter.py:111-122)
// T is the deterministic move R_a(y,x) = (B_y(y+dy), B_x(x+dx)) with B = clip or wrap per axis, so the
// sum is a one-cell stencil shift with edge accumulation ("push"): under clip the downstream edge
// cells receive two sources and the upstream edge cells none; under wrap the shift is cyclic.
//
// One CTA per agent.  The belief row is streamed once (read 4 B + write 4 B per state, 128-bit accesses) in chunks of whole
// grid rows.  Production kernel: belief_step_tma_kernel — the source rows of the next three chunks are always in flight as 1-D
// bulk copies (cp.async.bulk + mbarrier) into a shared-memory ring, the y-shift is a row-aligned offset of the copy (padded row
// layout), the x-shift a one-element offset of the shared-memory read, the likelihood a read of the (L2-resident) observation
// table, the per-agent normaliser a shuffle + shared-memory block reduction.  Normalisation is deferred: the row is stored
// unnormalised and 1/Z goes to scale[].  CTAs start their traversal at different chunks and alternate direction (see the
// kernel); in place (b_out == b_in, needed for 10^6 agents on 83x371: 123 GB of beliefs on a 180 GB part) the chunk order is the
// one that reads every source before it is overwritten.  belief_step_kernel (register streaming, x-shift by warp shuffles) is
// the fallback for grid rows that do not fit a ring stage, and the dense-model kernel lives in dense.cu.
#include "common.cuh"
#include "stencil.cuh"

#include <cstdlib>

namespace {

constexpr int BS_THREADS = 256;
constexpr int BS_UNROLL = 4;

struct StepArgs {
    const float* b_in;
    float* b_out;
    long long ld;
    const int* act;
    const int* obs;
    const int* src_rows;
    const int* dst_rows;
    const float* scale_in;
    float* scale_out;
    unsigned char* ok;
    const float* obs_tab;
    int n;
    int stagger;
    int rot;               // > 0: CTA i starts its traversal at chunk (i * rot) % nchunks (where the order is free)
    int rot_inplace;       // ... and in place with dy != 0 (cyclic order + one stashed source row): 1 always, 0 never, -1 on wide grids
    int inplace;
    const int* n_dev;      // optional device-side row count: CTAs with blockIdx.x >= *n_dev exit (grid sized by an upper bound)
};

template <bool INPLACE>
__global__ void __launch_bounds__(BS_THREADS, 2)
belief_step_kernel(const StepArgs p, const __grid_constant__ Geom g) {
    extern __shared__ __align__(128) float smem[];
    float* halo = smem;                 // Wp floats: the wrapped source row
    float* red = smem + g.Wp;           // 33 floats

    const int tid = threadIdx.x, lane = tid & 31;
    const int i = blockIdx.x;
    if (p.n_dev && i >= *p.n_dev) return;
    const long long src = p.src_rows ? p.src_rows[i] : i;
    const long long dst = p.dst_rows ? p.dst_rows[i] : i;
    int a = p.act[i], o = p.obs[i];
    float sc = p.scale_in ? p.scale_in[src] : 1.f;
    if (a < 0 || a >= g.A || o < 0 || o >= g.O) { a = 0; o = 0; sc = 0.f; }   // invalid ids => failed update
    const int dy = g.moves[a][0], dx = g.moves[a][1];
    const int H = g.H, W = g.W, W4 = g.W4, wrap_x = g.wrap_x;

    const float* bin = p.b_in + src * p.ld;
    float* bout = p.b_out + dst * p.ld;
    const float* ot = p.obs_tab + (size_t)((g.obs_per_action ? a : 0) * g.O + o) * g.Sp;

    float zacc = 0.f;

    // ---- prologue ---------------------------------------------------------------------------------------
    int lo_row = 0, hi_row = H;      // destination rows left to the main loop
    if (dy != 0) {
        if (g.wrap_y) {
            const int hrow = dy > 0 ? H - 1 : 0;
            for (int c = tid; c < W4; c += BS_THREADS)
                reinterpret_cast<float4*>(halo)[c] = ld_stream4(bin + 4 * (hrow * W4 + c));
            __syncthreads();
        } else {
            // clip: the downstream edge row receives its neighbour AND itself
            const int edge = dy > 0 ? H - 1 : 0;
            const int prim = edge - dy;
            float4 u[BS_UNROLL];
#pragma unroll
            for (int j = 0; j < BS_UNROLL; ++j) {
                const int c = tid + j * BS_THREADS;
                u[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < W4) {
                    u[j] = xshift_row(bin + (size_t)edge * g.Wp, c, dx, W, W4, wrap_x);
                    if (prim >= 0 && prim < H) {
                        const float4 q = xshift_row(bin + (size_t)prim * g.Wp, c, dx, W, W4, wrap_x);
                        u[j].x += q.x; u[j].y += q.y; u[j].z += q.z; u[j].w += q.w;
                    }
                }
            }
            if (INPLACE) __syncthreads();
#pragma unroll
            for (int j = 0; j < BS_UNROLL; ++j) {
                const int c = tid + j * BS_THREADS;
                if (c < W4) {
                    const int idx = edge * W4 + c;
                    const float4 l = __ldg(reinterpret_cast<const float4*>(ot) + idx);
                    float4 r;
                    r.x = u[j].x * sc * l.x; r.y = u[j].y * sc * l.y; r.z = u[j].z * sc * l.z; r.w = u[j].w * sc * l.w;
                    zacc += (r.x + r.y) + (r.z + r.w);
                    st_stream4(bout + 4 * (size_t)idx, r);
                }
            }
            if (dy > 0) hi_row = H - 1; else lo_row = 1;
        }
    }

    // ---- main loop over chunks of whole grid rows ----------------------------------------------------------
    const int rpc = max(1, (BS_THREADS * BS_UNROLL) / W4);      // grid rows per chunk
    const int nrows = hi_row - lo_row;
    const int nchunks = (nrows + rpc - 1) / rpc;
    const int row_cap = BS_THREADS * BS_UNROLL;                  // float4 slots per pass
    const int total4 = H * W4;

    for (int k = 0; k < nchunks; ++k) {
        int y0, y1;
        if (dy > 0) { y1 = hi_row - k * rpc; y0 = max(lo_row, y1 - rpc); }     // descending
        else        { y0 = lo_row + k * rpc; y1 = min(hi_row, y0 + rpc); }     // ascending
        const int base4 = y0 * W4;
        const int n4 = (y1 - y0) * W4;
        // W4 > row_cap only for grids wider than 4096 cells: several passes per row (host refuses INPLACE there)
        for (int pass0 = 0; pass0 < n4; pass0 += row_cap) {
            float4 u[BS_UNROLL];
#pragma unroll
            for (int j = 0; j < BS_UNROLL; ++j) {
                const int f = pass0 + tid + j * BS_THREADS;
                const bool valid = f < n4;
                const int idx = base4 + f;                 // destination float4 index
                int sidx = idx - dy * W4;                  // primary source float4 index
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                bool from_halo = false, has_src = valid;
                if (sidx < 0 || sidx >= total4) {
                    if (g.wrap_y) { from_halo = valid; sidx += (sidx < 0) ? total4 : -total4; }
                    else has_src = false;                  // clip: upstream edge row receives nothing
                }
                if (dx == 0) {
                    if (has_src) v = from_halo ? reinterpret_cast<const float4*>(halo)[sidx - (dy > 0 ? (H - 1) * W4 : 0)]
                                               : ld_stream4(bin + 4 * (size_t)sidx);
                } else {
                    // c4 = f mod W4 (position of this float4 inside its grid row)
                    const int q = (W4 == 1) ? f : (int)__umulhi((unsigned)f, g.div_w4);
                    const int c4 = f - q * W4;
                    const float* srow;                     // start of the source grid row
                    if (from_halo) srow = halo;
                    else srow = bin + 4 * (size_t)(sidx - c4);
                    if (has_src) v = from_halo ? reinterpret_cast<const float4*>(halo)[c4] : ld_stream4(bin + 4 * (size_t)sidx);
                    const int e = W - 1;
                    if (dx > 0) {
                        const float up = __shfl_up_sync(0xffffffffu, v.w, 1);
                        float left;
                        if (c4 == 0) left = (wrap_x && has_src) ? srow[e] : 0.f;
                        else if (lane == 0) left = has_src ? srow[4 * c4 - 1] : 0.f;
                        else left = up;
                        float4 out = make_float4(left, v.x, v.y, v.z);
                        if (!wrap_x && (e >> 2) == c4) add4(out, e & 3, comp4(v, e & 3));
                        v = out;
                    } else {
                        const float dn = __shfl_down_sync(0xffffffffu, v.x, 1);
                        float right;
                        if (c4 == W4 - 1) right = 0.f;
                        else if (lane == 31) right = has_src ? srow[4 * c4 + 4] : 0.f;
                        else right = dn;
                        float4 out = make_float4(v.y, v.z, v.w, right);
                        if (wrap_x && (e >> 2) == c4) set4(out, e & 3, has_src ? srow[0] : 0.f);
                        if (!wrap_x && c4 == 0) out.x += v.x;
                        v = out;
                    }
                }
                u[j] = v;
            }
            if (INPLACE) __syncthreads();
#pragma unroll
            for (int j = 0; j < BS_UNROLL; ++j) {
                const int f = pass0 + tid + j * BS_THREADS;
                if (f < n4) {
                    const int idx = base4 + f;
                    const float4 l = __ldg(reinterpret_cast<const float4*>(ot) + idx);
                    float4 r;
                    r.x = u[j].x * sc * l.x; r.y = u[j].y * sc * l.y; r.z = u[j].z * sc * l.z; r.w = u[j].w * sc * l.w;
                    zacc += (r.x + r.y) + (r.z + r.w);
                    st_stream4(bout + 4 * (size_t)idx, r);
                }
            }
        }
    }

    // ---- per-agent normaliser ---------------------------------------------------------------------------------
    const float z = block_sum<BS_THREADS>(zacc, red);
    if (tid == 0) {
        const bool good = z > 0.f && isfinite(z);
        p.scale_out[dst] = good ? 1.f / z : 0.f;
        p.ok[i] = good ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// v2: TMA-staged ring.  Same contract and chunk order as belief_step_kernel, but the source rows of the next
// BS2_STAGES chunks are always in flight as 1-D bulk copies (cp.async.bulk, mbarrier complete_tx) into shared
// memory, so bytes in flight are bounded by shared memory (4 CTAs x 48 KB per SM) instead of registers, and the
// E/W shift is a one-element offset of the shared-memory read.  In-place safety is unchanged: a chunk's bulk load
// completes (mbarrier) before any thread stores that chunk, and later chunks never read rows that earlier chunks
// wrote (same ordering argument), so prefetching them early is safe.
// Two shapes of the same kernel: 256 threads x 16 KB stages x 4 CTAs per SM (grid rows up to ~128 float4: C2's 93), and
// 512 threads x 32 KB stages x 2 CTAs per SM for wider grids (C5's 186 float4 per row: 11 rows per chunk instead of 5, so the
// per-CTA stagger below still leaves chunks of >= 8 rows).  Bytes in flight per SM are the same (192 KB).
constexpr int BS2_UNROLL = 4;
constexpr int BS2_STAGES = 3;

__device__ __forceinline__ uint32_t bs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bs_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) {       // watchdog: trap instead of hanging the GPU
            printf("olfnav belief_step: mbarrier watchdog (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

template <int BS2_THREADS>
__global__ void __launch_bounds__(BS2_THREADS, 1024 / BS2_THREADS)
belief_step_tma_kernel(const StepArgs p, const __grid_constant__ Geom g) {
    constexpr int BS2_STAGE_FLOATS = BS2_THREADS * BS2_UNROLL * 4;     // 16 KB (256 threads) or 32 KB (512 threads)
    extern __shared__ __align__(128) float smem[];
    float* stage_base = smem;                                        // BS2_STAGES x BS2_STAGE_FLOATS
    float* halo = smem + BS2_STAGES * BS2_STAGE_FLOATS;              // Wp floats: the wrapped source row
    float* stash = halo + g.Wp;                                      // Wp floats: the one source row a rotated in-place traversal overwrites early
    float* red = stash + g.Wp;                                       // 40 floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 40);          // BS2_STAGES mbarriers (8-byte aligned: Wp % 4 == 0)

    const int tid = threadIdx.x;
    const int i = blockIdx.x;
    if (p.n_dev && i >= *p.n_dev) return;
    const long long src = p.src_rows ? p.src_rows[i] : i;
    const long long dst = p.dst_rows ? p.dst_rows[i] : i;
    int a = p.act[i], o = p.obs[i];
    float sc = p.scale_in ? p.scale_in[src] : 1.f;
    if (a < 0 || a >= g.A || o < 0 || o >= g.O) { a = 0; o = 0; sc = 0.f; }
    const int dy = g.moves[a][0], dx = g.moves[a][1];
    const int H = g.H, W = g.W, W4 = g.W4, Wp = g.Wp, wrap_x = g.wrap_x;

    const float* bin = p.b_in + src * p.ld;
    float* bout = p.b_out + dst * p.ld;
    const float* ot = p.obs_tab + (size_t)((g.obs_per_action ? a : 0) * g.O + o) * g.Sp;

    // destination rows of the main loop and its chunking (identical to v1)
    int lo_row = 0, hi_row = H;
    if (dy != 0 && !g.wrap_y) { if (dy > 0) hi_row = H - 1; else lo_row = 1; }
    // Traversal order.  100 000 agents that all take the same action and all walk their rows in the same order run in
    // lockstep, and the read and write streams then collide in DRAM: 64-78 % of the HBM bandwidth on uniform actions (measured,
    // C2 and C5).  Where the order is free (out of place, or dy == 0) CTA i therefore starts at chunk (i * rot) % nchunks and
    // neighbouring CTAs alternate direction: 97-100 % on every action pattern (profiles/r1j_rot_sweep.jsonl).  In place with
    // dy != 0 the direction is dictated by the move (sources are read before they are overwritten: descending for dy > 0,
    // ascending for dy < 0) but the start is rotated too (see `rot` below); with rot = 0 the chunk SIZE varies with the CTA
    // index instead, which dephases the CTAs over time (C2 98.6 %, C5 87 %).
    const int rpc_max = max(1, (BS2_THREADS * BS2_UNROLL) / W4);
    const bool free_dir = !p.inplace || dy == 0;
    // in place with dy != 0: rotate where the per-CTA chunk-size stagger would leave small chunks (wide grids: C5 89 % -> 98 %);
    // on C2 (11 rows per chunk) the stagger is as good (98 % vs 96 %), so it stays
    const bool may_rotate = p.rot > 0 && (free_dir || p.rot_inplace == 1 || (p.rot_inplace < 0 && rpc_max < 8));
    const int rpc = max(1, rpc_max - ((!free_dir && !may_rotate && p.stagger > 1) ? (int)(blockIdx.x % (unsigned)p.stagger) : 0));
    const int nrows = hi_row - lo_row;
    const int nchunks = (nrows + rpc - 1) / rpc;
    const bool desc = free_dir ? ((blockIdx.x & 1u) != 0u) : (dy > 0);
    // The start is rotated in place with dy != 0 as well: the chunks still follow the dictated direction cyclically, and the only source
    // row that is overwritten before its reader runs — the first row written by the first chunk, needed by the last chunk of the
    // cycle — is stashed in shared memory by the prologue.
    const int rot = (may_rotate && nchunks > 1) ? (int)(((unsigned)blockIdx.x * (unsigned)p.rot) % (unsigned)nchunks) : 0;
    const bool use_stash = !free_dir && rot > 0;
    auto chunk_rows = [&](int k, int& y0, int& y1) {
        k += rot;
        if (k >= nchunks) k -= nchunks;
        if (desc) { y1 = hi_row - k * rpc; y0 = max(lo_row, y1 - rpc); }
        else      { y0 = lo_row + k * rpc; y1 = min(hi_row, y0 + rpc); }
    };
    // bulk copy of the in-range source rows of chunk k into stage k % STAGES (row r of the stage = source of dest row y0 + r)
    auto issue = [&](int k) {
        int y0, y1;
        chunk_rows(k, y0, y1);
        const int s0 = max(0, y0 - dy), s1 = min(H, y1 - dy);           // valid source rows [s0, s1)
        if (s1 <= s0) return;
        const uint32_t bytes = (uint32_t)(s1 - s0) * Wp * 4u;
        const int st = k % BS2_STAGES;
        const uint32_t bar = bs_smem_u32(&bars[st]);
        const uint32_t dstp = bs_smem_u32(stage_base + st * BS2_STAGE_FLOATS + (size_t)(s0 + dy - y0) * Wp);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dstp), "l"(bin + (size_t)s0 * Wp), "r"(bytes), "r"(bar) : "memory");
    };

    if (tid == 0) {
        for (int st = 0; st < BS2_STAGES; ++st)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bs_smem_u32(&bars[st])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int k = 0; k < BS2_STAGES && k < nchunks; ++k) issue(k);

    float zacc = 0.f;

    // ---- prologue (the bulk loads above are already in flight) ------------------------------------------
    int stash_src = -1;
    if (use_stash) {
        int y0r, y1r;
        chunk_rows(0, y0r, y1r);                        // the first chunk of the rotated cycle
        stash_src = desc ? y1r - 1 : y0r;               // its first-written row is the source of the row just beyond it
        for (int c = tid; c < W4; c += BS2_THREADS)
            reinterpret_cast<float4*>(stash)[c] = ld_stream4(bin + 4 * (stash_src * W4 + c));
    }
    if (dy != 0) {
        if (g.wrap_y) {
            const int hrow = dy > 0 ? H - 1 : 0;
            for (int c = tid; c < W4; c += BS2_THREADS)
                reinterpret_cast<float4*>(halo)[c] = ld_stream4(bin + 4 * (hrow * W4 + c));
            __syncthreads();
        } else {
            const int edge = dy > 0 ? H - 1 : 0;
            const int prim = edge - dy;
            float4 u[BS2_UNROLL];
#pragma unroll
            for (int j = 0; j < BS2_UNROLL; ++j) {
                const int c = tid + j * BS2_THREADS;
                u[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < W4) {
                    u[j] = xshift_row(bin + (size_t)edge * Wp, c, dx, W, W4, wrap_x);
                    if (prim >= 0 && prim < H) {
                        const float4 q = xshift_row(bin + (size_t)prim * Wp, c, dx, W, W4, wrap_x);
                        u[j].x += q.x; u[j].y += q.y; u[j].z += q.z; u[j].w += q.w;
                    }
                }
            }
            __syncthreads();            // in place: every read of the edge row precedes its overwrite
#pragma unroll
            for (int j = 0; j < BS2_UNROLL; ++j) {
                const int c = tid + j * BS2_THREADS;
                if (c < W4) {
                    const int idx = edge * W4 + c;
                    const float4 l = __ldg(reinterpret_cast<const float4*>(ot) + idx);
                    float4 r;
                    r.x = u[j].x * sc * l.x; r.y = u[j].y * sc * l.y; r.z = u[j].z * sc * l.z; r.w = u[j].w * sc * l.w;
                    zacc += (r.x + r.y) + (r.z + r.w);
                    st_stream4(bout + 4 * (size_t)idx, r);
                }
            }
        }
    }

    // ---- main loop: consume the ring -----------------------------------------------------------------------
    uint32_t phase_bits = 0;            // per-stage mbarrier parity: a stage's phase only advances when a copy was issued into it
    for (int k = 0; k < nchunks; ++k) {
        int y0, y1;
        chunk_rows(k, y0, y1);
        const int st = k % BS2_STAGES;
        const float* stg = stage_base + st * BS2_STAGE_FLOATS;
        const int s0 = max(0, y0 - dy), s1 = min(H, y1 - dy);
        if (s1 > s0) {
            bs_mbar_wait(bs_smem_u32(&bars[st]), (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
        }
        const int base4 = y0 * W4;
        const int n4 = (y1 - y0) * W4;
        const bool stash_chunk = use_stash && k == nchunks - 1;        // the last chunk of a rotated in-place cycle reads the stashed row
        const bool special = (y0 - dy < 0) || (y1 - 1 - dy >= H) || stash_chunk;   // chunk holds a destination row without a source in its stage
#pragma unroll
        for (int j = 0; j < BS2_UNROLL; ++j) {
            const int f = tid + j * BS2_THREADS;
            if (f < n4) {
                float4 v;
                if (dx == 0 && !special) {
                    v = reinterpret_cast<const float4*>(stg)[f];
                } else {
                    const int r = (W4 == 1) ? f : (int)__umulhi((unsigned)f, g.div_w4);
                    const int c4 = f - r * W4;
                    const int ys = y0 + r - dy;
                    if (stash_chunk && ys == stash_src) v = xshift_row(stash, c4, dx, W, W4, wrap_x);
                    else if (ys >= 0 && ys < H) v = xshift_row(stg + (size_t)r * Wp, c4, dx, W, W4, wrap_x);
                    else if (g.wrap_y)     v = xshift_row(halo, c4, dx, W, W4, wrap_x);
                    else                   v = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int idx = base4 + f;
                const float4 l = __ldg(reinterpret_cast<const float4*>(ot) + idx);
                float4 rr;
                rr.x = v.x * sc * l.x; rr.y = v.y * sc * l.y; rr.z = v.z * sc * l.z; rr.w = v.w * sc * l.w;
                zacc += (rr.x + rr.y) + (rr.z + rr.w);
                st_stream4(bout + 4 * (size_t)idx, rr);
            }
        }
        __syncthreads();                                               // stage fully consumed
        if (tid == 0 && k + BS2_STAGES < nchunks) issue(k + BS2_STAGES);
    }

    const float z = block_sum<BS2_THREADS>(zacc, red);
    if (tid == 0) {
        const bool good = z > 0.f && isfinite(z);
        p.scale_out[dst] = good ? 1.f / z : 0.f;
        p.ok[i] = good ? 1 : 0;
    }
}

__global__ void belief_init_kernel(const float* __restrict__ start, float* __restrict__ b, long long ld, int Sp4,
                                   float* __restrict__ scale) {
    const long long row = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < Sp4) st_stream4(b + row * ld + 4 * (size_t)c, __ldg(reinterpret_cast<const float4*>(start) + c));
    if (c == 0 && scale) scale[row] = 1.f;
}

template <typename T>
__global__ void pack_rows_kernel(const T* __restrict__ dense, float* __restrict__ padded, long long ld, int H, int W, int Wp) {
    const long long row = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // padded index
    if (c >= H * Wp) return;
    const int y = c / Wp, x = c - y * Wp;
    padded[row * ld + c] = (x < W) ? (float)dense[row * (long long)(H * W) + y * W + x] : 0.f;
}

template <typename T>
__global__ void unpack_rows_kernel(const float* __restrict__ padded, long long ld, const float* __restrict__ scale,
                                   T* __restrict__ dense, int H, int W, int Wp) {
    const long long row = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;   // dense index
    if (c >= H * W) return;
    const int y = c / W, x = c - y * W;
    const float s = scale ? scale[row] : 1.f;
    dense[row * (long long)(H * W) + c] = (T)(padded[row * ld + y * Wp + x] * s);
}

__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long ld, int len4,
                                 const int* __restrict__ src_rows, const float* __restrict__ scale_src, float* __restrict__ scale_dst) {
    const long long i = blockIdx.y;
    const long long r = src_rows[i];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < len4) st_stream4(dst + i * ld + 4 * (size_t)c, ld_stream4(src + r * ld + 4 * (size_t)c));
    if (c == 0 && scale_src && scale_dst) scale_dst[i] = scale_src[r];
}

// H(b) = -Σ p log p of the normalised row p = b * scale   (agents/infotaxis_agent.py:257)
__global__ void __launch_bounds__(256) entropies_kernel(const float* __restrict__ b, long long ld, int Sp4,
                                                        const float* __restrict__ scale, float* __restrict__ Hout) {
    __shared__ float red[33];
    const long long row = blockIdx.x;
    const float s = scale ? scale[row] : 1.f;
    const float4* r4 = reinterpret_cast<const float4*>(b + row * ld);
    float acc = 0.f;
    for (int c = threadIdx.x; c < Sp4; c += 256) {
        const float4 v = r4[c];
        const float px = v.x * s, py = v.y * s, pz = v.z * s, pw = v.w * s;
        acc += (px > 0.f ? px * __logf(px) : 0.f) + (py > 0.f ? py * __logf(py) : 0.f)
             + (pz > 0.f ? pz * __logf(pz) : 0.f) + (pw > 0.f ? pw * __logf(pw) : 0.f);
    }
    const float t = block_sum<256>(acc, red);
    if (threadIdx.x == 0) Hout[row] = -t;
}

}  // namespace

// ======================================================================================================
extern "C" int olfnav_belief_init(const olfnav_model* m, float* b, int64_t ld, int64_t n, float* scale, void* stream) {
    OLF_REQUIRE(m && b && n >= 0, "belief_init: bad arguments");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0, "belief_init: ld must be >= Sp and a multiple of 4");
    if (n == 0) return OLFNAV_OK;
    OLF_REQUIRE(n <= 65535LL * 65535LL, "belief_init: n too large");
    OLF_CUDA(cudaSetDevice(m->device));
    const int Sp4 = m->g.Sp / 4;
    // grid.y is limited to 65535: loop over slabs of rows
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int rows = (int)((n - r0 < 65535) ? (n - r0) : 65535);
        dim3 grid(olf_div_up(Sp4, 256), rows);
        belief_init_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m->start, b + r0 * ld, ld, Sp4, scale ? scale + r0 : nullptr);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_belief_step(const olfnav_model* m, const float* b_in, float* b_out, int64_t ld, int64_t n,
                                  const int32_t* act, const int32_t* obs, const int32_t* src_rows, const int32_t* dst_rows,
                                  const float* scale_in, float* scale_out, uint8_t* ok, void* stream) {
    return olf_belief_step_launch(m, b_in, b_out, ld, n, act, obs, src_rows, dst_rows, scale_in, scale_out, ok, nullptr, stream);
}

// n = upper bound of the rows to update (grid size); n_dev (optional) = device-side exact count.
int olf_belief_step_launch(const olfnav_model* m, const float* b_in, float* b_out, int64_t ld, int64_t n,
                           const int32_t* act, const int32_t* obs, const int32_t* src_rows, const int32_t* dst_rows,
                           const float* scale_in, float* scale_out, uint8_t* ok, const int* n_dev, void* stream) {
    OLF_REQUIRE(m && b_in && b_out && act && obs && scale_out && ok, "belief_step: null argument");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0, "belief_step: ld must be >= Sp and a multiple of 4");
    OLF_REQUIRE(n >= 0 && n <= 0x7fffffffLL, "belief_step: n out of range");
    OLF_REQUIRE((((uintptr_t)b_in | (uintptr_t)b_out) & 15) == 0, "belief_step: belief buffers must be 16-byte aligned");
    if (n == 0) return OLFNAV_OK;
    if (m->dense) {
        OLF_CUDA(cudaSetDevice(m->device));
        return olf_dense_belief_step(m, b_in, b_out, ld, n, act, obs, src_rows, dst_rows, scale_in, scale_out, ok, n_dev, (cudaStream_t)stream);
    }
    const bool inplace = (b_in == b_out);
    if (inplace && m->g.W4 > BS_THREADS * BS_UNROLL) {
        olfnav_set_error("belief_step: in-place update needs W <= %d", 4 * BS_THREADS * BS_UNROLL);
        return OLFNAV_ERR_UNSUPPORTED;
    }
    OLF_CUDA(cudaSetDevice(m->device));
    StepArgs p;
    p.b_in = b_in; p.b_out = b_out; p.ld = ld; p.act = act; p.obs = obs;
    p.src_rows = src_rows; p.dst_rows = dst_rows; p.scale_in = scale_in; p.scale_out = scale_out;
    p.ok = ok; p.obs_tab = m->obs; p.n = (int)n;
    static const int stagger = [] { const char* e = getenv("OLFNAV_BELIEF_STAGGER"); return e ? atoi(e) : 4; }();
    p.stagger = stagger;
    static const int rot = [] { const char* e = getenv("OLFNAV_BELIEF_ROT"); return e ? atoi(e) : 7; }();
    p.rot = rot;
    { const char* e = getenv("OLFNAV_BELIEF_ROT_INPLACE"); p.rot_inplace = e ? (e[0] == '1' ? 1 : 0) : -1; }
    p.inplace = inplace ? 1 : 0;
    p.n_dev = n_dev;
    // v2 (TMA ring) whenever a grid row fits one stage; OLFNAV_BELIEF_V1=1 forces the register-streaming v1 kernel.
    // 256-thread shape (16 KB stages, 4 CTAs per SM) while a grid row fits a stage, 512-thread shape (32 KB stages) for grids wider
    // than 4096 cells; OLFNAV_BELIEF_THREADS=256|512 overrides the choice (read per call: tests and experiments toggle it).
    static const bool force_v1 = [] { const char* e = getenv("OLFNAV_BELIEF_V1"); return e && e[0] == '1'; }();
    const char* ft = getenv("OLFNAV_BELIEF_THREADS");
    const int force_threads = ft ? atoi(ft) : 0;
    auto ring_smem = [&](int threads) {
        return (size_t)(BS2_STAGES * threads * BS2_UNROLL * 4 + 2 * m->g.Wp + 40) * sizeof(float) + BS2_STAGES * sizeof(uint64_t);
    };
    const bool fits256 = m->g.W4 <= 256 * BS2_UNROLL && ring_smem(256) <= 56 * 1024;       // 4 CTAs per SM
    const bool fits512 = m->g.W4 <= 512 * BS2_UNROLL && ring_smem(512) <= 112 * 1024;      // 2 CTAs per SM
    if (!force_v1 && (fits256 || fits512)) {
        const int threads = (fits256 && !(force_threads == 512 && fits512)) ? 256 : 512;
        const size_t smem2 = ring_smem(threads);
        static OlfPerDeviceOnce configured;
        if (configured.first()) {
            OLF_CUDA(cudaFuncSetAttribute(belief_step_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024));
            OLF_CUDA(cudaFuncSetAttribute(belief_step_tma_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        }
        if (threads == 256) belief_step_tma_kernel<256><<<(unsigned)n, 256, smem2, (cudaStream_t)stream>>>(p, m->g);
        else                belief_step_tma_kernel<512><<<(unsigned)n, 512, smem2, (cudaStream_t)stream>>>(p, m->g);
    } else {
        const size_t smem = (size_t)(m->g.Wp + 40) * sizeof(float);
        if (inplace) belief_step_kernel<true><<<(unsigned)n, BS_THREADS, smem, (cudaStream_t)stream>>>(p, m->g);
        else         belief_step_kernel<false><<<(unsigned)n, BS_THREADS, smem, (cudaStream_t)stream>>>(p, m->g);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_pack_rows(const olfnav_model* m, const double* dense_f64, const float* dense_f32,
                                float* padded, int64_t ld, int64_t n, void* stream) {
    OLF_REQUIRE(m && padded && ((dense_f64 != nullptr) != (dense_f32 != nullptr)), "pack_rows: exactly one dense source");
    OLF_REQUIRE(ld >= m->g.Sp, "pack_rows: ld too small");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int rows = (int)((n - r0 < 65535) ? (n - r0) : 65535);
        dim3 grid(olf_div_up(m->g.Sp, 256), rows);
        if (dense_f64) pack_rows_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(dense_f64 + r0 * m->g.S, padded + r0 * ld, ld, m->g.H, m->g.W, m->g.Wp);
        else           pack_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(dense_f32 + r0 * m->g.S, padded + r0 * ld, ld, m->g.H, m->g.W, m->g.Wp);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_unpack_rows(const olfnav_model* m, const float* padded, int64_t ld, int64_t n, const float* scale,
                                  double* dense_f64, float* dense_f32, void* stream) {
    OLF_REQUIRE(m && padded && ((dense_f64 != nullptr) != (dense_f32 != nullptr)), "unpack_rows: exactly one dense destination");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int rows = (int)((n - r0 < 65535) ? (n - r0) : 65535);
        dim3 grid(olf_div_up(m->g.S, 256), rows);
        if (dense_f64) unpack_rows_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(padded + r0 * ld, ld, scale ? scale + r0 : nullptr, dense_f64 + r0 * m->g.S, m->g.H, m->g.W, m->g.Wp);
        else           unpack_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(padded + r0 * ld, ld, scale ? scale + r0 : nullptr, dense_f32 + r0 * m->g.S, m->g.H, m->g.W, m->g.Wp);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_entropies(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                                float* H_out, void* stream) {
    OLF_REQUIRE(m && b && H_out, "entropies: null argument");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    entropies_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(b, ld, m->g.Sp / 4, scale, H_out);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_copy_rows(const float* src, float* dst, int64_t ld, int64_t row_len, const int32_t* src_rows,
                                const float* scale_src, float* scale_dst, int64_t n, void* stream) {
    OLF_REQUIRE(src && dst && src_rows && src != dst, "copy_rows: bad pointers (src and dst must differ)");
    OLF_REQUIRE((ld % 4) == 0 && (row_len % 4) == 0 && row_len <= ld, "copy_rows: ld and row_len must be multiples of 4");
    if (n == 0) return OLFNAV_OK;
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int rows = (int)((n - r0 < 65535) ? (n - r0) : 65535);
        dim3 grid(olf_div_up(row_len / 4, 256), rows);
        copy_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst + r0 * ld, ld, (int)(row_len / 4), src_rows + r0,
                                                                 scale_src, scale_dst ? scale_dst + r0 : nullptr);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
