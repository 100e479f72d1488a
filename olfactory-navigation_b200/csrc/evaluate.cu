// Value-function handle and evaluation dispatch.
//
// evaluate replaces pomdp_toolkit's ValueFunction.evaluate_at as called by the reference at
// olfactory_navigation/agents/pbvi_agent.py:741:  V = b·Γᵀ, max / first-argmax over the alpha vectors,
// action = actions[argmax].
//
// Two hand-written paths:
//   * segmax_simt_kernel  — FP32 FMA on CUDA cores, one warp per belief row, alphas from L1/L2.
//     HBM-bound for a handful of alpha vectors (QMDP: G = A = 4) and the exact-FP32 cross-check
//     for the tensor-core path.
//   * umma_segmax         — tcgen05 / TMEM / TMA 3xTF32 split-precision GEMM with a fused per-row
//     (segmented) max/argmax epilogue (umma_gemm.cu).
#include "common.cuh"
#include "umma_gemm.cuh"

#include <cstdlib>
#include <cstring>
#include <new>

namespace {

// hi = v with the low 13 mantissa bits cleared (exactly representable in TF32);
// lo = (v - hi) truncated the same way.  v - hi - lo is below 2^-22 |v|.
__global__ void alpha_split_kernel(const float* __restrict__ alpha, long long ld_alpha, int G, int Sp,
                                   float* __restrict__ full, float* __restrict__ hi, float* __restrict__ lo, long long ld) {
    const int g = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ld) return;
    const float v = (g < G && c < Sp) ? alpha[(long long)g * ld_alpha + c] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const float l = __uint_as_float(__float_as_uint(v - h) & 0xffffe000u);
    full[(long long)g * ld + c] = v;
    hi[(long long)g * ld + c] = h;
    lo[(long long)g * ld + c] = l;
}

constexpr int SM_WARPS = 8;   // belief rows per CTA

// out_val[i][seg] = max_{g in seg} b_i · M_g ; out_arg = first argmax (index inside the segment).
template <int GT>
__global__ void __launch_bounds__(SM_WARPS * 32)
segmax_simt_kernel(const float* __restrict__ b, long long ld, int n, int K4,
                   const float* __restrict__ M, long long ldM, int n_seg, int seg_len,
                   float* __restrict__ out_val, int* __restrict__ out_arg) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * SM_WARPS + w;
    if (i >= n) return;
    const float* row = b + i * ld;
    for (int seg = 0; seg < n_seg; ++seg) {
        float best = -INFINITY;
        int best_g = 0;
        for (int g0 = 0; g0 < seg_len; g0 += GT) {
            float acc[GT];
#pragma unroll
            for (int t = 0; t < GT; ++t) acc[t] = 0.f;
            const float* Mg = M + (long long)(seg * seg_len + g0) * ldM;
            const int gt = min(GT, seg_len - g0);
            for (int q = lane; q < K4; q += 32) {
                const float4 v = ld_stream4(row + 4 * (size_t)q);
#pragma unroll
                for (int t = 0; t < GT; ++t) {
                    if (t < gt) {
                        const float4 m = __ldg(reinterpret_cast<const float4*>(Mg + (long long)t * ldM) + q);
                        acc[t] = fmaf(v.x, m.x, acc[t]);
                        acc[t] = fmaf(v.y, m.y, acc[t]);
                        acc[t] = fmaf(v.z, m.z, acc[t]);
                        acc[t] = fmaf(v.w, m.w, acc[t]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < GT; ++t) {
                const float s = warp_sum(acc[t]);
                if (t < gt && s > best) { best = s; best_g = g0 + t; }   // strict '>' keeps the first maximum
            }
        }
        if (lane == 0) {
            out_val[i * n_seg + seg] = best;
            out_arg[i * n_seg + seg] = best_g;
        }
    }
}

__global__ void eval_finish_kernel(int n, const float* __restrict__ val, const int* __restrict__ arg,
                                   const float* __restrict__ scale, const int* __restrict__ actions,
                                   float* __restrict__ value, int* __restrict__ arg_out, int* __restrict__ action) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = arg[i];
    if (value) value[i] = val[i] * (scale ? scale[i] : 1.f);
    if (arg_out) arg_out[i] = g;
    if (action) action[i] = actions[g];
}

}  // namespace

// Shared by evaluate and backup_select.
int olf_segmax_simt(const float* b, int64_t ld, int64_t n, int K4, const float* M, int64_t ldM, int n_seg, int seg_len,
                    float* out_val, int* out_arg, cudaStream_t st) {
    const unsigned grid = (unsigned)olf_div_up(n, SM_WARPS);
    if (seg_len <= 4) segmax_simt_kernel<4><<<grid, SM_WARPS * 32, 0, st>>>(b, ld, (int)n, K4, M, ldM, n_seg, seg_len, out_val, out_arg);
    else              segmax_simt_kernel<8><<<grid, SM_WARPS * 32, 0, st>>>(b, ld, (int)n, K4, M, ldM, n_seg, seg_len, out_val, out_arg);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_alphas_create(olfnav_alphas** out, const olfnav_model* m, const float* alpha, int64_t ld_alpha,
                                    const int32_t* actions, int G, void* stream) {
    OLF_REQUIRE(out && m && alpha && actions, "alphas_create: null argument");
    OLF_REQUIRE(G >= 1 && ld_alpha >= m->g.Sp, "alphas_create: bad G or ld_alpha");
    OLF_CUDA(cudaSetDevice(m->device));
    olfnav_alphas* a = new (std::nothrow) olfnav_alphas();
    if (!a) { olfnav_set_error("alphas_create: out of host memory"); return OLFNAV_ERR_NOMEM; }
    memset(a, 0, sizeof(*a));
    a->device = m->device;
    a->G = G;
    a->Gp = (G + 15) & ~15;
    a->ld = (m->g.Sp + 31) & ~31;
    const size_t plane = (size_t)a->Gp * a->ld * sizeof(float);
    cudaError_t e = cudaMalloc((void**)&a->full, plane);
    if (e == cudaSuccess) e = cudaMalloc((void**)&a->hi, plane);
    if (e == cudaSuccess) e = cudaMalloc((void**)&a->lo, plane);
    if (e == cudaSuccess) e = cudaMalloc((void**)&a->actions, (size_t)G * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(a->actions, actions, (size_t)G * sizeof(int32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        olfnav_set_error("alphas_create: %s", cudaGetErrorString(e));
        olfnav_alphas_destroy(a);
        return OLFNAV_ERR_CUDA;
    }
    dim3 grid(olf_div_up(a->ld, 256), a->Gp);
    alpha_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(alpha, ld_alpha, G, m->g.Sp, a->full, a->hi, a->lo, a->ld);
    olf_count_launch(1);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        olfnav_set_error("alphas_create: %s", cudaGetErrorString(le));
        olfnav_alphas_destroy(a);
        return OLFNAV_ERR_CUDA;
    }
    a->f16_eval = G >= 64 ? 1 : 0;
    {
        // FP16 split planes (always built: the one-pass update + policy kernel uses them at every G; evaluate_at from G = 64 up)
        // FP16 split planes for the tensor path (half the tensor time of 3xTF32, same accuracy); below G = 64 the product is HBM bound
        // with either kernel (measured: G = 64 2.02 -> 1.97 ms, G = 75 inside the C2 step 2.17 -> 1.93 ms, G = 512 12.9 -> 8.67 ms)
        a->ldh = (m->g.Sp + 63) & ~63;
        e = cudaMalloc(&a->h1, (size_t)a->Gp * a->ldh * 2);
        if (e == cudaSuccess) e = cudaMalloc(&a->h2, (size_t)a->Gp * a->ldh * 2);
        if (e == cudaSuccess) e = cudaMalloc((void**)&a->binv, (size_t)a->Gp * sizeof(float));
        int rc = OLFNAV_OK;
        if (e != cudaSuccess) { olfnav_set_error("alphas_create: %s", cudaGetErrorString(e)); rc = OLFNAV_ERR_CUDA; }
        if (rc == OLFNAV_OK) rc = olf_split_planes_f16(alpha, ld_alpha, G, m->g.Sp, a->h1, a->h2, a->binv, a->ldh, a->Gp, 1, (cudaStream_t)stream);
        if (rc != OLFNAV_OK) { olfnav_alphas_destroy(a); return rc; }
    }
    *out = a;
    return OLFNAV_OK;
}

extern "C" void olfnav_alphas_destroy(olfnav_alphas* a) {
    if (!a) return;
    cudaSetDevice(a->device);
    cudaFree(a->full); cudaFree(a->hi); cudaFree(a->lo); cudaFree(a->actions);
    cudaFree(a->h1); cudaFree(a->h2); cudaFree(a->binv);
    delete a;
}

int olf_evaluate_into(const olfnav_model* m, const olfnav_alphas* a, const float* b, int64_t ld, int64_t n, const float* scale,
                      float* val, int* idx, float* value, int32_t* arg, int32_t* action, int path, cudaStream_t st) {
    // auto: FP32 SIMT for a handful of alpha vectors on small batches (lower latency), tcgen05 otherwise — at 100 000 rows the
    // tensor path streams at 89 % of HBM even for G = 4 (SIMT: 79 %, scripts/bench_rows.py)
    if (path == 0) path = (a->G <= 8 && n < 8192) ? 1 : 2;
    int rc;
    if (path == 1) rc = olf_segmax_simt(b, ld, n, m->g.Sp / 4, a->full, a->ld, 1, a->G, val, idx, st);
    else {
        const char* f16_env = getenv("OLFNAV_GEMM_F16");           // read per call: "0" selects the 3xTF32 kernel
        if (a->h1 && a->f16_eval && !(f16_env && f16_env[0] == '0'))
            rc = olf_umma_segmax_f16(b, ld, n, m->g.Sp, scale, a->h1, a->h2, a->binv, a->ldh, a->Gp, 1, a->G, a->Gp, val, idx, m->sm_count, st);
        else
            rc = olf_umma_segmax(b, ld, n, m->g.Sp, a->hi, a->lo, a->ld, a->Gp, 1, a->G, a->Gp, val, idx, m->sm_count, st);
    }
    if (rc != OLFNAV_OK) return rc;
    eval_finish_kernel<<<olf_div_up(n, 256), 256, 0, st>>>((int)n, val, idx, scale, a->actions, value, arg, action);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_evaluate(const olfnav_model* m, const olfnav_alphas* a, const float* b, int64_t ld, int64_t n,
                               const float* scale, float* value, int32_t* arg, int32_t* action, int path, void* stream) {
    OLF_REQUIRE(m && a && b, "evaluate: null argument");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0 && n >= 0 && n <= 0x7fffffffLL, "evaluate: bad ld or n");
    OLF_REQUIRE(path >= 0 && path <= 2, "evaluate: path must be 0, 1 or 2");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    OlfScratch val(st), idx(st);
    OLF_CUDA(val.alloc((size_t)n * sizeof(float)));
    OLF_CUDA(idx.alloc((size_t)n * sizeof(int)));
    return olf_evaluate_into(m, a, b, ld, n, scale, val.as<float>(), idx.as<int>(), value, arg, action, path, st);
}
