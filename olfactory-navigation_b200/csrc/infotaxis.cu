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
//
// Accuracy.  ΔH_a is a difference of terms of order one (the entropy of the observation and its conditional entropy) while the gaps
// between actions that decide the argmax can be 1e-6 and less on flat beliefs: FP32 sums (1e-6 relative) cannot settle those.
// The finish kernel therefore combines the sums in float64, bounds its own error per action (E_a = kappa (1 + H_o + |Σ b C| + |corr|)) and
// lists the agents for which another action lies within the bounds of the best one, together with the mask of those actions;
// infotaxis_exact_kernel re-derives the listed (agent, action) values with every sum in double (absolute error ~1e-14 on terms of order
// one, both normalised by Σ_o p_o = Σ b like the fast path) and infotaxis_pick_kernel takes the first maximum among them.  A caller that
// asks for the ΔH values themselves (dH != NULL, diagnostics and tests) gets the float64 kernels for every agent and action.  Decisions
// therefore agree with a float64 evaluation of the reference's formula (agents/infotaxis_agent.py:257-293) except at ties below 1e-9 relative.
#include "common.cuh"
#include "stencil.cuh"
#include "umma_gemm.cuh"

#include <cstdlib>
#include <algorithm>

namespace {

constexpr int IT_THREADS = 256;
constexpr int IT_MAXO = 8;

__device__ __forceinline__ float xlogx(float p) { return p > 0.f ? p * logf(p) : 0.f; }
__device__ __forceinline__ double xlogx_d(double p) { return p > 0.0 ? p * log(p) : 0.0; }

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
                        const float* __restrict__ sums, int ld_sums, int* __restrict__ action, float* __restrict__ dH,
                        float kappa, int* __restrict__ unsure_rows, int* __restrict__ unsure_count, int* __restrict__ unsure_mask) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * IF_WARPS + (threadIdx.x >> 5);
    if (row >= n) return;
    const float sc = scale ? scale[row] : 1.f;
    const float* bin = b + row * ld;
    const float* sm = sums + row * ld_sums;
    double best = -INFINITY, best_err = 0.0;
    int ba = -1;
    double d_all[OLFNAV_MAX_ACTIONS], e_all[OLFNAV_MAX_ACTIONS];
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
        // float64 combination of the FP32 sums.  Σ_o p_o = Σ_s b(s) exactly (Σ_o O_o = 1 everywhere), so the observation
        // probabilities are renormalised by their own sum: the common relative bias of the tensor-core accumulation drops out.
        double psum = 0.0;
        for (int o = 0; o < g.O; ++o) psum += (double)sm[a * (g.O + 1) + o];
        const double pinv = psum > 0.0 ? 1.0 / psum : 0.0;
        double plogp = 0.0, habs = 0.0;
        for (int o = 0; o < g.O; ++o) {
            const double t = xlogx_d((double)sm[a * (g.O + 1) + o] * pinv);
            plogp += t;
            habs -= t;
        }
        const double bc = (double)sm[a * (g.O + 1) + g.O] * pinv;       // same bias as the p sums: divided out too
        const double d = -plogp + bc + (double)corr;
        d_all[a] = d;
        e_all[a] = (double)kappa * (1.0 + habs + fabs(bc) + fabs((double)corr));
        if (lane == 0 && dH) dH[row * g.A + a] = (float)d;
        if (best < d) { best = d; best_err = e_all[a]; ba = a; }          // strict '<' : first maximum (infotaxis_agent.py:289)
    }
    if (lane == 0) {
        action[row] = ba;
        // another action within the error bounds of the best one: float64 re-evaluation (infotaxis_exact_kernel)
        // (only the actions inside the bounds take part: the others are lower for certain)
        unsigned mask = 0;
        for (int a = 0; a < g.A; ++a)
            if (a == ba || best - d_all[a] <= best_err + e_all[a]) mask |= 1u << a;
        if ((mask & (mask - 1)) && unsure_rows) {
            const int k = atomicAdd(unsure_count, 1);
            unsure_rows[k] = (int)row;
            unsure_mask[k] = (int)mask;
        }
    }
}


__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum of doubles, result in every thread; red holds >= 33 doubles
template <int THREADS>
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum_d(v);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = (lane < THREADS / 32) ? red[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
}
__device__ __forceinline__ int bmove_d(int v, int d, int n, int wrap) {
    v += d;
    if (wrap) { if (v < 0) v += n; if (v >= n) v -= n; } else { v = max(0, min(n - 1, v)); }
    return v;
}

// float64 evaluation of ΔH_a for the agents listed in `rows` (count on the device; masks[j] = the actions of entry j that take part)
// or for every agent and action (rows == NULL).  ΔH_a = -Σ_o p_o log p_o + Σ_s b(s) C(R_a(s)) + corr_a with p_o = Σ_s b(s) O_o(R_a(s)),
// C = Σ_o O_o log O_o (obs_ent64, double) and corr_a the entropy change of the push at clipped edges: the same three terms as the FP32
// path, every sum in double (absolute error ~1e-14 on terms of order one; the gaps this path is asked to settle are 1e-9 and up).
// A CTA takes IX_ROWS list entries and ONE action, one pass over the rows: the likelihood and C tables of a cell (20 bytes per state and
// action against 4 bytes of belief) are loaded once for the group — the path is bound by L2 traffic, and a thousand unsure agents must
// not cost as much as the hundred thousand sure ones.  ΔH_a goes to dh64[entry * A + a]; infotaxis_pick_kernel takes the first maximum.
constexpr int IX_ROWS = 4;
template <int NO>
__global__ void __launch_bounds__(IT_THREADS)
infotaxis_exact_kernel(const __grid_constant__ Geom g, int n, const int* __restrict__ rows, const int* __restrict__ n_rows, const int* __restrict__ masks,
                       const float* __restrict__ b, long long ld, const float* __restrict__ scale, const float* __restrict__ obs,
                       const double* __restrict__ obs_ent64, double* __restrict__ dh64) {
    __shared__ double red[33];
    const int a = blockIdx.y;
    const int tid = threadIdx.x;
    const int entries = rows ? min(*n_rows, n) : n;
    const int dy = g.moves[a][0], dx = g.moves[a][1];
    const int ae = g.obs_per_action ? a : 0;
    const float* ot = obs + (size_t)ae * g.O * g.Sp;
    const double* ce = obs_ent64 + (size_t)ae * g.Sp;
    const bool my = dy != 0 && !g.wrap_y && g.H > 1, mx = dx != 0 && !g.wrap_x && g.W > 1;
    // the grid is a fixed number of CTAs per action (the list is usually a percent of the agents), each walks the groups with the grid's stride
    for (int j0 = blockIdx.x * IX_ROWS; j0 < entries; j0 += gridDim.x * IX_ROWS) {
        const float* bin[IX_ROWS];
        bool on[IX_ROWS];
        bool any = false;
#pragma unroll
        for (int r = 0; r < IX_ROWS; ++r) {
            const int j = j0 + r;
            on[r] = j < entries && (!rows || ((masks[j] >> a) & 1));
            const long long row = on[r] ? (rows ? rows[j] : j) : 0;
            bin[r] = b + row * ld;
            any |= on[r];
        }
        if (!any) continue;
        double P[IX_ROWS][NO], C[IX_ROWS];
#pragma unroll
        for (int r = 0; r < IX_ROWS; ++r) {
            C[r] = 0.0;
#pragma unroll
            for (int o = 0; o < NO; ++o) P[r][o] = 0.0;
        }
        for (int y = 0; y < g.H; ++y) {
            const int y2 = bmove_d(y, dy, g.H, g.wrap_y) * g.Wp;
            for (int x = tid; x < g.W; x += IT_THREADS) {
                const int s = y * g.Wp + x, s2 = y2 + bmove_d(x, dx, g.W, g.wrap_x);
                double v[IX_ROWS];
#pragma unroll
                for (int r = 0; r < IX_ROWS; ++r) v[r] = on[r] ? (double)bin[r][s] : 0.0;
                const double c = ce[s2];
                double O[NO];
#pragma unroll
                for (int o = 0; o < NO; ++o) O[o] = o < g.O ? (double)__ldg(ot + (size_t)o * g.Sp + s2) : 0.0;
                // (the FP64 pipe is what bounds this kernel: no multiply-add for observations the model does not have — NO is the
                // model's count rounded up to 3 / 5 / 8 — nor for the empty rows of a group)
#pragma unroll
                for (int r = 0; r < IX_ROWS; ++r) {
                    if (!on[r]) continue;
                    C[r] += v[r] * c;
#pragma unroll
                    for (int o = 0; o < NO; ++o) P[r][o] += v[r] * O[o];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < IX_ROWS; ++r) {
            if (!on[r]) continue;                                   // uniform over the CTA
            // entropy change of the push: cells of the clipped downstream edge lines receive themselves and their upstream neighbours
            double corr = 0.0;
            if (my || mx) {
                const int ey = dy > 0 ? g.H - 1 : 0, ex = dx > 0 ? g.W - 1 : 0;
                const int n_line_y = my ? g.W : 0, n_line_x = mx ? g.H : 0;
                for (int k = tid; k < n_line_y + n_line_x; k += IT_THREADS) {
                    int y, x;
                    if (k < n_line_y) { y = ey; x = k; } else { y = k - n_line_y; x = ex; if (my && y == ey) continue; }
                    double tot = 0.0, parts = 0.0;
                    for (int sy = -1; sy <= 0; ++sy)
                        for (int sx = -1; sx <= 0; ++sx) {
                            const int ys = (sy == 0) ? y : y - dy, xs = (sx == 0) ? x : x - dx;
                            if (sy == 0 && !(my && y == ey) && dy != 0) continue;
                            if (sx == 0 && !(mx && x == ex) && dx != 0) continue;
                            if (sy == -1 && dy == 0) continue;
                            if (sx == -1 && dx == 0) continue;
                            int yy = ys, xx = xs;
                            if (g.wrap_y) { if (yy < 0) yy += g.H; if (yy >= g.H) yy -= g.H; } else if (yy < 0 || yy >= g.H) continue;
                            if (g.wrap_x) { if (xx < 0) xx += g.W; if (xx >= g.W) xx -= g.W; } else if (xx < 0 || xx >= g.W) continue;
                            const double v = (double)bin[r][(size_t)yy * g.Wp + xx];
                            tot += v;
                            parts += xlogx_d(v);
                        }
                    corr += xlogx_d(tot) - parts;          // of the UNNORMALISED row: k (x log x)(tot) - Σ k (x log x)(v) = k [same of the raw values], tot = Σ v
                }
            }
            // Σ_o p_o = Σ_s b(s) exactly (Σ_o O_o = 1 everywhere): everything is normalised by that sum, like the fast path — the row then
            // sums to one to the last bit, whatever the few ulps between its carried FP32 scale and 1 / Σ b
            const double lin = block_sum_d<IT_THREADS>(C[r] + corr, red);
            double p[NO], psum = 0.0;
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                p[o] = o < g.O ? block_sum_d<IT_THREADS>(P[r][o], red) : 0.0;
                psum += p[o];
            }
            const double pinv = psum > 0.0 ? 1.0 / psum : 0.0;
            double d = lin * pinv;
#pragma unroll
            for (int o = 0; o < NO; ++o)
                if (o < g.O) d -= xlogx_d(p[o] * pinv);
            if (tid == 0) dh64[(size_t)(j0 + r) * g.A + a] = d;
        }
    }
}

// first maximum of the float64 values of an entry's participating actions (strict '<': infotaxis_agent.py:289)
__global__ void infotaxis_pick_kernel(int A, int n, const int* __restrict__ rows, const int* __restrict__ n_rows, const int* __restrict__ masks,
                                      const double* __restrict__ dh64, int* __restrict__ action, float* __restrict__ dH) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || (rows && j >= *n_rows)) return;
    const long long row = rows ? rows[j] : j;
    const unsigned mask = rows ? (unsigned)masks[j] : 0xffffffffu;
    double best = -INFINITY;
    int ba = -1;
    for (int a = 0; a < A; ++a) {
        if (!((mask >> a) & 1)) continue;
        const double d = dh64[(size_t)j * A + a];
        if (dH) dH[row * A + a] = (float)d;
        if (best < d) { best = d; ba = a; }
    }
    action[row] = ba;
}

void launch_exact(const olfnav_model* m, int64_t n, const int* rows, const int* n_rows, const int* masks, const float* b, int64_t ld,
                  const float* scale, double* dh64, cudaStream_t st) {
    const dim3 grid((unsigned)std::min<int64_t>(olf_div_up(n, IX_ROWS), (int64_t)m->sm_count * 8), m->g.A);
    if (m->g.O <= 3)      infotaxis_exact_kernel<3><<<grid, IT_THREADS, 0, st>>>(m->g, (int)n, rows, n_rows, masks, b, ld, scale, m->obs, m->obs_ent64, dh64);
    else if (m->g.O <= 5) infotaxis_exact_kernel<5><<<grid, IT_THREADS, 0, st>>>(m->g, (int)n, rows, n_rows, masks, b, ld, scale, m->obs, m->obs_ent64, dh64);
    else                  infotaxis_exact_kernel<8><<<grid, IT_THREADS, 0, st>>>(m->g, (int)n, rows, n_rows, masks, b, ld, scale, m->obs, m->obs_ent64, dh64);
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
    const bool exact_ok = m->obs_ent64 && m->g.O <= IT_MAXO;
    if (dH && exact_ok && !force_v1) {
        // the values themselves are asked for (diagnostics, tests): float64 for every agent and action
        OlfScratch s_dh(st);
        OLF_CUDA(s_dh.alloc((size_t)n * m->g.A * sizeof(double)));
        launch_exact(m, n, nullptr, nullptr, nullptr, b, ld, scale, s_dh.as<double>(), st);
        OLF_LAUNCH_CHECK();
        infotaxis_pick_kernel<<<olf_div_up(n, 256), 256, 0, st>>>(m->g.A, (int)n, nullptr, nullptr, nullptr, s_dh.as<double>(), action, dH);
        OLF_LAUNCH_CHECK();
        return OLFNAV_OK;
    }
    if (!force_v1 && m->pull_rows <= 128) {
        const int ld_sums = m->pull_rows;
        OlfScratch s_sums(st), s_list(st), s_dh(st);
        OLF_CUDA(s_sums.alloc((size_t)n * ld_sums * sizeof(float)));
        float* sums = s_sums.as<float>();
        int* list = nullptr;
        if (exact_ok) {
            // rows [0, n) | count [n] | action masks [n + 1, 2n + 1); float64 values of the listed (entry, action) pairs
            OLF_CUDA(s_list.alloc((2 * (size_t)n + 1) * sizeof(int)));
            OLF_CUDA(s_dh.alloc((size_t)n * m->g.A * sizeof(double)));
            list = s_list.as<int>();
            OLF_CUDA(cudaMemsetAsync(list + n, 0, sizeof(int), st));
        }
        // error bound of the FP32 path per unit of (1 + H_o + |Σ b C|): the split-precision GEMM is good to 1.2e-6 of Σ|a||b|
        // (measured); OLFNAV_INFOTAXIS_KAPPA overrides (0 = never re-evaluate)
        static const float kappa = [] { const char* e = getenv("OLFNAV_INFOTAXIS_KAPPA"); return e ? (float)atof(e) : 4e-6f; }();
        int rc = olf_umma_store(b, ld, n, m->g.Sp, m->pull_hi, m->pull_lo, m->pull_ld, m->pull_rows, sums, ld_sums, m->sm_count, st);
        if (rc == OLFNAV_OK) {
            infotaxis_finish_kernel<<<olf_div_up(n, IF_WARPS), IF_WARPS * 32, 0, st>>>(m->g, (int)n, b, ld, scale, sums, ld_sums, action, dH, kappa,
                                                                                      kappa > 0.f ? list : nullptr, list ? list + n : nullptr, list ? list + n + 1 : nullptr);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("infotaxis_choose: finish kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
        if (rc == OLFNAV_OK && list && kappa > 0.f) {
            launch_exact(m, n, list, list + n, list + n + 1, b, ld, scale, s_dh.as<double>(), st);
            infotaxis_pick_kernel<<<olf_div_up(n, 256), 256, 0, st>>>(m->g.A, (int)n, list, list + n, list + n + 1, s_dh.as<double>(), action, nullptr);
            olf_count_launch(2);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("infotaxis_choose: exact kernel launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
        if (rc == OLFNAV_OK && list && m->it_unsure) OLF_CUDA(cudaMemcpyAsync(m->it_unsure, list + n, sizeof(int), cudaMemcpyDeviceToDevice, st));
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

// How many agents of the last olfnav_infotaxis_choose call on this model went through the float64 re-evaluation (synchronises the stream).
extern "C" int olfnav_infotaxis_last_unsure(const olfnav_model* m, int32_t* count, void* stream) {
    OLF_REQUIRE(m && count, "infotaxis_last_unsure: null argument");
    *count = 0;
    if (!m->it_unsure) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    OLF_CUDA(cudaMemcpyAsync(count, m->it_unsure, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OLF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return OLFNAV_OK;
}
