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
// One CTA per agent; per action one pass over the pushed belief b_a (re-reads hit L1/L2).
#include "common.cuh"
#include "stencil.cuh"

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

}  // namespace

extern "C" int olfnav_infotaxis_choose(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                                       int32_t* action, float* dH, void* stream) {
    OLF_REQUIRE(m && b && action, "infotaxis_choose: null argument");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0 && n >= 0 && n <= 0x7fffffffLL, "infotaxis_choose: bad ld or n");
    if (m->g.O > IT_MAXO) {
        olfnav_set_error("infotaxis_choose: at most %d observations supported (model has %d)", IT_MAXO, m->g.O);
        return OLFNAV_ERR_UNSUPPORTED;
    }
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (m->g.O <= 3)      infotaxis_kernel<3><<<(unsigned)n, IT_THREADS, 0, st>>>(m->g, b, ld, scale, m->obs, m->obs_ent, action, dH);
    else if (m->g.O <= 5) infotaxis_kernel<5><<<(unsigned)n, IT_THREADS, 0, st>>>(m->g, b, ld, scale, m->obs, m->obs_ent, action, dH);
    else                  infotaxis_kernel<8><<<(unsigned)n, IT_THREADS, 0, st>>>(m->g, b, ld, scale, m->obs, m->obs_ent, action, dH);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
