// Infotaxis action choice as one fused reduction per agent — nothing is materialised.
//
// Replaces Infotaxis_Agent.choose_action (reference olfactory_navigation/agents/infotaxis_agent.py:244-301),
// which materialises every successor belief b_{a,o} (BeliefSet.generate_all_successors, :260-262), their
// entropies (:270), P(o|b,a) through an (n,S,R) einsum (:271-274) and then loops in Python (:279-293).
//
// With w = O_o(s') b_a(s'), p_{a,o} = Σ_{s'} w and Σ_o O_o = 1 (asserted by the reference at
// agents/model_based_util/environment_converter.py:97):
//     Σ_o p_{a,o} H(b_{a,o}) = Σ_o p log p − Σ_{s'} b_a(s') C(s') − Σ_{s'} b_a(s') log b_a(s'),
//     C(s') = Σ_o O_o(s') log O_o(s')  (precomputed table)
// so   ΔH_a = H(b) − Σ_o p_{a,o} log p_{a,o} + Σ b_a C + Σ b_a log b_a ,   a* = first argmax_a ΔH_a.
//
// v2 (default): the pushes are permutations of the grid except at clipped edges, so Σ b_a log b_a = Σ b log b + corr_a with
//     corr_a = Σ_{merge cells} [ xlogx(v1 + v2) − xlogx(v1) − xlogx(v2) ]      (edge cell v1 and its upstream neighbour v2)
// and H(b) = −Σ b log b cancels:   ΔH_a = −Σ_o p_{a,o} log p_{a,o} + Σ_s b(s) C(R_a(s)) + corr_a ,
// with p_{a,o} = Σ_s b(s) O_o(R_a(s)) in pull form.  The A·(O+1) sums are ONE pass of the tcgen05 split-precision GEMM
// (beliefs x pull tables, umma_gemm.cu store mode: 4·S bytes read per decision, no logarithm per state); a small finish
// kernel adds the O(H + W) edge corrections and takes the first argmax.
// v1 (OLFNAV_INFOTAXIS_V1=1, or more than 128 pull rows): one CTA per agent, one pass over the pushed belief per action.
#include "common.cuh"
#include "stencil.cuh"
#include "umma_gemm.cuh"

#include <cstdlib>

namespace {

constexpr int IT_THREADS = 256;
constexpr int IT_MAXO = 8;

__device__ __forceinline__ float xlogx(float p) { return p > 0.f ? p * logf(p) : 0.f; }

template <int NO>
__global__ void __launch_bounds__(IT_THREADS)
infotaxis_kernel(const __grid_constant__ Geom g, const float* __restrict__ b, long long ld, const float* __restrict__ scale,
                 const float* __restrict__ obs, const float* __restrict__ obs_ent, int* __restrict__ action, float* __restrict__ dH) {
    __shared__ float red[33];
    __shared__ float s_dh[OLFNAV_MAX_ACTIONS];
    const int tid = threadIdx.x;
    const long long row = blockIdx.x;
    const float sc = scale ? scale[row] : 1.f;
    const float* bin = b + row * ld;
    const int total4 = g.H * g.W4;

    // H(b)
    float hacc = 0.f;
    for (int f = tid; f < total4; f += IT_THREADS) {
        const float4 v = *reinterpret_cast<const float4*>(bin + 4 * (size_t)f);
        hacc += xlogx(v.x * sc) + xlogx(v.y * sc) + xlogx(v.z * sc) + xlogx(v.w * sc);
    }
    const float Hb = -block_sum<IT_THREADS>(hacc, red);

    for (int a = 0; a < g.A; ++a) {
        const int dy = g.moves[a][0], dx = g.moves[a][1];
        const float* ot = obs + (size_t)(g.obs_per_action ? a : 0) * g.O * g.Sp;
        const float4* ce = reinterpret_cast<const float4*>(obs_ent + (size_t)(g.obs_per_action ? a : 0) * g.Sp);
        float P[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) P[o] = 0.f;
        float csum = 0.f, lsum = 0.f;
        for (int f = tid; f < total4; f += IT_THREADS) {
            const int y = (g.W4 == 1) ? f : (int)__umulhi((unsigned)f, g.div_w4);
            const int c4 = f - y * g.W4;
            float4 v = pushed4(bin, y, c4, dy, dx, g);
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            // padding columns: the pushed value can be non-zero there (x' = W gets b[W-1] for dx = +1); mask them
            const int x0 = 4 * c4;
            if (x0 + 3 >= g.W) {
                if (x0 + 0 >= g.W) v.x = 0.f;
                if (x0 + 1 >= g.W) v.y = 0.f;
                if (x0 + 2 >= g.W) v.z = 0.f;
                v.w = 0.f;
            }
            const float4 c = __ldg(ce + f);
            csum += v.x * c.x + v.y * c.y + v.z * c.z + v.w * c.w;
            lsum += xlogx(v.x) + xlogx(v.y) + xlogx(v.z) + xlogx(v.w);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (o < g.O) {
                    const float4 l = __ldg(reinterpret_cast<const float4*>(ot + (size_t)o * g.Sp) + f);
                    P[o] += v.x * l.x + v.y * l.y + v.z * l.z + v.w * l.w;
                }
            }
        }
        float plogp = 0.f;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            if (o < g.O) plogp += xlogx(block_sum<IT_THREADS>(P[o], red));
        }
        const float cs = block_sum<IT_THREADS>(csum, red);
        const float ls = block_sum<IT_THREADS>(lsum, red);
        if (tid == 0) s_dh[a] = Hb - plogp + cs + ls;
    }
    __syncthreads();
    if (tid == 0) {
        float best = -INFINITY;
        int ba = -1;
        for (int a = 0; a < g.A; ++a) {
            const float d = s_dh[a];
            if (dH) dH[row * g.A + a] = d;
            if (best < d) { best = d; ba = a; }     // strict '<' : first maximum (infotaxis_agent.py:289)
        }
        action[row] = ba;
    }
}

// One warp per agent: edge corrections + combination of the GEMM sums + first argmax.
constexpr int IF_WARPS = 8;
__global__ void __launch_bounds__(IF_WARPS * 32)
infotaxis_finish_kernel(const __grid_constant__ Geom g, int n, const float* __restrict__ b, long long ld, const float* __restrict__ scale,
                        const float* __restrict__ sums, int ld_sums, int* __restrict__ action, float* __restrict__ dH) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * IF_WARPS + (threadIdx.x >> 5);
    if (row >= n) return;
    const float sc = scale ? scale[row] : 1.f;
    const float* bin = b + row * ld;
    const float* sm = sums + row * ld_sums;
    float best = -INFINITY;
    int ba = -1;
    for (int a = 0; a < g.A; ++a) {
        const int dy = g.moves[a][0], dx = g.moves[a][1];
        // merge cells: a clipped axis with a non-zero move component; every cell of the downstream edge line receives itself
        // and its upstream neighbour.  (Both axes clipped and a diagonal move would need the corner's three sources: the
        // reference's action sets are axis moves; olfnav_model_create rejects nothing here, so handle the general case.)
        float corr = 0.f;
        const bool my = dy != 0 && !g.wrap_y && g.H > 1, mx = dx != 0 && !g.wrap_x && g.W > 1;
        if (my || mx) {
            // destination cells on the downstream edge line(s); for each, gather its sources by pulling R_a^{-1}
            const int ey = dy > 0 ? g.H - 1 : 0, ex = dx > 0 ? g.W - 1 : 0;
            const int n_line_y = my ? g.W : 0;                     // cells (ey, x)
            const int n_line_x = mx ? g.H : 0;                     // cells (y, ex)
            for (int k = lane; k < n_line_y + n_line_x; k += 32) {
                int y, x;
                if (k < n_line_y) { y = ey; x = k; } else { y = k - n_line_y; x = ex; if (my && y == ey) continue; }   // corner once
                // sources of (y, x): every (ys, xs) with R_a(ys, xs) = (y, x)
                float tot = 0.f, parts = 0.f;
                for (int sy = -1; sy <= 0; ++sy)
                    for (int sx = -1; sx <= 0; ++sx) {
                        // candidate source = destination minus the move (sy/sx = -1) or the destination itself along that axis (0)
                        const int ys = (sy == 0) ? y : y - dy, xs = (sx == 0) ? x : x - dx;
                        if (sy == 0 && !(my && y == ey) && dy != 0) continue;      // staying along y only happens on the clipped edge
                        if (sx == 0 && !(mx && x == ex) && dx != 0) continue;
                        if (sy == -1 && dy == 0) continue;                           // no move along y: only the "stay" candidate
                        if (sx == -1 && dx == 0) continue;
                        int yy = ys, xx = xs;
                        if (g.wrap_y) { if (yy < 0) yy += g.H; if (yy >= g.H) yy -= g.H; } else if (yy < 0 || yy >= g.H) continue;
                        if (g.wrap_x) { if (xx < 0) xx += g.W; if (xx >= g.W) xx -= g.W; } else if (xx < 0 || xx >= g.W) continue;
                        const float v = bin[(size_t)yy * g.Wp + xx] * sc;
                        tot += v;
                        parts += xlogx(v);
                    }
                corr += xlogx(tot) - parts;
            }
            corr = warp_sum(corr);
        }
        float plogp = 0.f;
        for (int o = 0; o < g.O; ++o) plogp += xlogx(sm[a * (g.O + 1) + o] * sc);
        const float d = -plogp + sm[a * (g.O + 1) + g.O] * sc + corr;
        if (lane == 0 && dH) dH[row * g.A + a] = d;
        if (best < d) { best = d; ba = a; }          // strict '<' : first maximum (infotaxis_agent.py:289)
    }
    if (lane == 0) action[row] = ba;
}

}  // namespace

extern "C" int olfnav_infotaxis_choose(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                                       int32_t* action, float* dH, void* stream) {
    OLF_REQUIRE(m && b && action, "infotaxis_choose: null argument");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0 && n >= 0 && n <= 0x7fffffffLL, "infotaxis_choose: bad ld or n");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (m->dense) return olf_dense_infotaxis(m, b, ld, n, scale, action, dH, st);
    static const bool force_v1 = [] { const char* e = getenv("OLFNAV_INFOTAXIS_V1"); return e && e[0] == '1'; }();
    if (!force_v1 && m->pull_rows <= 128) {
        const int ld_sums = m->pull_rows;
        OlfScratch s_sums(st);
        OLF_CUDA(s_sums.alloc((size_t)n * ld_sums * sizeof(float)));
        float* sums = s_sums.as<float>();
        int rc = olf_umma_store(b, ld, n, m->g.Sp, m->pull_hi, m->pull_lo, m->pull_ld, m->pull_rows, sums, ld_sums, m->sm_count, st);
        if (rc == OLFNAV_OK) {
            infotaxis_finish_kernel<<<olf_div_up(n, IF_WARPS), IF_WARPS * 32, 0, st>>>(m->g, (int)n, b, ld, scale, sums, ld_sums, action, dH);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("infotaxis_choose: finish kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
        return rc;
    }
    if (m->g.O > IT_MAXO) {
        olfnav_set_error("infotaxis_choose: the per-agent kernel supports at most %d observations (model has %d)", IT_MAXO, m->g.O);
        return OLFNAV_ERR_UNSUPPORTED;
    }
    if (m->g.O <= 3)      infotaxis_kernel<3><<<(unsigned)n, IT_THREADS, 0, st>>>(m->g, b, ld, scale, m->obs, m->obs_ent, action, dH);
    else if (m->g.O <= 5) infotaxis_kernel<5><<<(unsigned)n, IT_THREADS, 0, st>>>(m->g, b, ld, scale, m->obs, m->obs_ent, action, dH);
    else                  infotaxis_kernel<8><<<(unsigned)n, IT_THREADS, 0, st>>>(m->g, b, ld, scale, m->obs, m->obs_ent, action, dH);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
