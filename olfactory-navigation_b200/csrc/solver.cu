// Environment stepping, MDP value iteration and the streaming halves of the PBVI backup.
#include "common.cuh"
#include "umma_gemm.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

int olf_segmax_simt(const float* b, int64_t ld, int64_t n, int K4, const float* M, int64_t ldM, int n_seg, int seg_len,
                    float* out_val, int* out_arg, cudaStream_t st);

namespace {

__device__ __forceinline__ int bmove_dev(int v, int d, int n, int wrap) {
    v += d;
    if (wrap) { if (v < 0) v += n; if (v >= n) v -= n; }
    else { v = max(0, min(n - 1, v)); }
    return v;
}

// simulation.py:1457-1466 + agent.py:401-439, one thread per agent.
__global__ void env_step_kernel(const olfnav_env e, int n, int* __restrict__ pos, const int* __restrict__ act,
                                const int* __restrict__ moves, const long long* __restrict__ time_shift, long long step,
                                const double* __restrict__ thr, int n_thr,
                                double* __restrict__ odor, int* __restrict__ obs_id, unsigned char* __restrict__ reached) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = act[i];
    // environment.py:711-769
    const int y = bmove_dev(pos[2 * i], moves[2 * a], e.H, e.wrap_y);
    const int x = bmove_dev(pos[2 * i + 1], moves[2 * a + 1], e.W, e.wrap_x);
    pos[2 * i] = y;
    pos[2 * i + 1] = x;
    // environment.py:591 (time loops), :638-651 (0.0 outside the data zone)
    long long t = (time_shift[i] + step) % e.T;
    if (t < 0) t += e.T;
    const int dyp = y - e.my, dxp = x - e.mx;
    double v = 0.0;
    if (dyp >= 0 && dyp < e.h && dxp >= 0 && dxp < e.w) v = e.data[((size_t)t * e.h + dyp) * e.w + dxp];
    odor[i] = v;
    // environment.py:655-680
    const double ddy = (double)y - e.sy, ddx = (double)x - e.sx;
    const bool hit = (ddy * ddy + ddx * ddx) <= e.r2;
    reached[i] = hit ? 1 : 0;
    // agent.py:420 bins [t_k, t_{k+1}); :437 goal id = number of bins
    int id = -1;
    for (int k = 0; k + 1 < n_thr; ++k)
        if (v >= thr[k] && v < thr[k + 1]) { id = k; break; }
    obs_id[i] = hit ? (n_thr - 1) : id;
}

// solvers.VI: Q_out[a][s] = r_a(s) + discount * max_a' Q_in[a'][R_a(s)]   (float64: V reaches 1/(1-gamma))
__global__ void vi_sweep_kernel(const Geom g, const float* __restrict__ reward, const double* __restrict__ Qin,
                                double* __restrict__ Qout, long long ldq, double discount, double* __restrict__ max_change) {
    const int sp = blockIdx.x * blockDim.x + threadIdx.x;
    double change = 0.0;
    if (sp < g.Sp) {
        const int y = sp / g.Wp, x = sp - y * g.Wp;
        if (x < g.W) {
            double vin = -1e300, vout = -1e300;
            for (int a = 0; a < g.A; ++a) vin = fmax(vin, Qin[a * ldq + sp]);
            for (int a = 0; a < g.A; ++a) {
                const int d = bmove_dev(y, g.moves[a][0], g.H, g.wrap_y) * g.Wp + bmove_dev(x, g.moves[a][1], g.W, g.wrap_x);
                double vd = -1e300;
                for (int a2 = 0; a2 < g.A; ++a2) vd = fmax(vd, Qin[a2 * ldq + d]);
                const double q = (double)reward[(size_t)a * g.Sp + sp] + discount * vd;
                Qout[a * ldq + sp] = q;
                vout = fmax(vout, q);
            }
            change = fabs(vout - vin);
        } else {
            for (int a = 0; a < g.A; ++a) Qout[a * ldq + sp] = 0.0;
        }
    }
    // non-negative doubles order like their bit patterns
    for (int o = 16; o > 0; o >>= 1) change = fmax(change, __shfl_xor_sync(0xffffffffu, change, o));
    if ((threadIdx.x & 31) == 0 && change > 0.0)
        atomicMax(reinterpret_cast<unsigned long long*>(max_change), (unsigned long long)__double_as_longlong(change));
}

// Gamma[a][o][g][s] = discount * O_o(R_a(s)) * alpha_g(R_a(s))      (SURVEY App. B "Backup (pull)")
__global__ void backup_gamma_kernel(const Geom g, const float* __restrict__ obs, const float* __restrict__ alpha, long long ld_alpha,
                                    int G, float discount, float* __restrict__ out, long long ld_out) {
    const int sp = blockIdx.x * blockDim.x + threadIdx.x;
    const int gi = blockIdx.y, a = blockIdx.z;
    if (sp >= ld_out) return;
    const int y = sp / g.Wp, x = sp - y * g.Wp;
    const bool real = sp < g.Sp && x < g.W;
    int d = 0;
    float av = 0.f;
    if (real) {
        d = bmove_dev(y, g.moves[a][0], g.H, g.wrap_y) * g.Wp + bmove_dev(x, g.moves[a][1], g.W, g.wrap_x);
        av = discount * alpha[(long long)gi * ld_alpha + d];
    }
    const float* ob = obs + (size_t)(g.obs_per_action ? a : 0) * g.O * g.Sp;
    for (int o = 0; o < g.O; ++o) {
        const float v = real ? av * ob[(size_t)o * g.Sp + d] : 0.f;
        out[((long long)(a * g.O + o) * G + gi) * ld_out + sp] = v;
    }
}

// best_a[b] = first argmax_a ( b.r_a + sum_o maxval[b,a,o] ), all in *scaled* (true belief) units.
__global__ void backup_pick_kernel(int n, int A, int O, const float* __restrict__ scale, const float* __restrict__ seg_val,
                                   const float* __restrict__ rew_val, int* __restrict__ best_a, float* __restrict__ best_value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float best = -INFINITY;
    int ba = 0;
    for (int a = 0; a < A; ++a) {
        float q = rew_val[i * A + a];
        for (int o = 0; o < O; ++o) q += seg_val[(long long)i * A * O + a * O + o];
        if (q > best) { best = q; ba = a; }
    }
    best_a[i] = ba;
    if (best_value) best_value[i] = best * (scale ? scale[i] : 1.f);
}

// alpha_new[b][:] = r_{a*} + sum_o Gamma[a*][o][g*(b,a*,o)][:]
__global__ void backup_assemble_kernel(const Geom g, const float* __restrict__ reward, const float* __restrict__ gamma_ao,
                                       long long ld_gamma, int G, const int* __restrict__ best_g, const int* __restrict__ best_a,
                                       float* __restrict__ out, long long ld_out) {
    const long long b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
    if (4 * c >= ld_out) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < g.Sp / 4) {
        const int a = best_a[b];
        acc = __ldg(reinterpret_cast<const float4*>(reward + (size_t)a * g.Sp) + c);
        for (int o = 0; o < g.O; ++o) {
            const int gi = best_g[(b * g.A + a) * g.O + o];
            const float4 v = __ldg(reinterpret_cast<const float4*>(gamma_ao + ((long long)(a * g.O + o) * G + gi) * ld_gamma) + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    *reinterpret_cast<float4*>(out + b * ld_out + 4 * (size_t)c) = acc;
}

// The FP16 split planes of Gamma built straight from the alpha vectors — Gamma itself is never written (at G = 512 it is 783 MB
// of FP32 that the old path wrote, read back, and split: 1.24 of the 12.9 ms of a backup).  One CTA per Gamma row (a, o, g): pass 1
// finds the row's maximum for the power-of-two scale, pass 2 recomputes the row (the alpha vector is L2 / L1 resident by then) and
// writes b1 = fp16(t v), b2 = fp16((t v - b1) 2^11), binv = 1 / t (same scheme as split_planes_f16_kernel, same values: the row is
// computed with the same operations as backup_gamma_kernel).  grid = (1, Gs, A * O), rows >= G of a segment are zero.
__global__ void __launch_bounds__(256) gamma_planes_f16_kernel(const Geom g, const float* __restrict__ obs, const float* __restrict__ alpha,
                                                               long long ld_alpha, int G, float discount, __half* __restrict__ b1,
                                                               __half* __restrict__ b2, float* __restrict__ binv, long long ldh) {
    __shared__ float red[8];
    const int gi = blockIdx.y, ao = blockIdx.z;
    const int a = ao / g.O, o = ao - a * g.O;
    const long long dst = (long long)ao * gridDim.y + gi;
    const bool live = gi < G;
    const float* arow = alpha + (long long)gi * ld_alpha;
    const float* ob = obs + ((size_t)(g.obs_per_action ? a : 0) * g.O + o) * g.Sp;
    const int dy = g.moves[a][0], dx = g.moves[a][1];
    // Gamma value at grid cell (y, x): (discount alpha(R_a(y, x))) O_o(R_a(y, x)), the operations of backup_gamma_kernel
    auto value = [&](int y, int x) -> float {
        const int d = bmove_dev(y, dy, g.H, g.wrap_y) * g.Wp + bmove_dev(x, dx, g.W, g.wrap_x);
        return __fmul_rn(__fmul_rn(discount, arow[d]), __ldg(ob + d));
    };
    float mx = 0.f;
    if (live)
        for (int y = 0; y < g.H; ++y)
            for (int x = threadIdx.x; x < g.W; x += blockDim.x) mx = fmaxf(mx, fabsf(value(y, x)));
    for (int w = 16; w > 0; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    const float p2 = __uint_as_float(__float_as_uint(mx) & 0x7f800000u);
    const float t = p2 > 0.f ? 8192.f / p2 : 1.f;
    if (threadIdx.x == 0) binv[dst] = 1.f / t;
    // two columns per thread: 32-bit stores of half pairs (Wp and ldh are even)
    __half2* r1 = reinterpret_cast<__half2*>(b1 + dst * ldh);
    __half2* r2 = reinterpret_cast<__half2*>(b2 + dst * ldh);
    for (int y = 0; y < g.H; ++y)
        for (int x = 2 * threadIdx.x; x < g.Wp; x += 2 * blockDim.x) {
            const float v0 = (live && x < g.W) ? value(y, x) * t : 0.f;
            const float v1 = (live && x + 1 < g.W) ? value(y, x + 1) * t : 0.f;
            const __half2 h = __floats2half2_rn(v0, v1);
            const float2 f = __half22float2(h);
            r1[(y * g.Wp + x) >> 1] = h;
            r2[(y * g.Wp + x) >> 1] = __floats2half2_rn((v0 - f.x) * 2048.f, (v1 - f.y) * 2048.f);
        }
    for (int c = g.Sp + 2 * threadIdx.x; c < ldh; c += 2 * blockDim.x) {          // columns beyond the padded row: zero
        r1[c >> 1] = __floats2half2_rn(0.f, 0.f);
        r2[c >> 1] = __floats2half2_rn(0.f, 0.f);
    }
}

// alpha_new[b][:] = r_{a*} + sum_o Gamma[a*][o][g*(b,a*,o)][:] with the Gamma rows recomputed from the alpha vectors (same
// operations, same order as backup_gamma_kernel + backup_assemble_kernel: bit-identical vectors)
__global__ void backup_assemble_alphas_kernel(const Geom g, const float* __restrict__ reward, const float* __restrict__ obs,
                                              const float* __restrict__ alpha, long long ld_alpha, float discount,
                                              const int* __restrict__ best_g, const int* __restrict__ best_a, float* __restrict__ out,
                                              long long ld_out) {
    const long long b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
    if (4 * c >= ld_out) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < g.Sp / 4) {
        const int a = best_a[b];
        acc = __ldg(reinterpret_cast<const float4*>(reward + (size_t)a * g.Sp) + c);
        const int y = (g.W4 == 1) ? c : (int)__umulhi((unsigned)c, g.div_w4);
        const int c4 = c - y * g.W4;
        const int dy = g.moves[a][0], dx = g.moves[a][1];
        const int y2 = bmove_dev(y, dy, g.H, g.wrap_y);
        const float* ob = obs + (size_t)(g.obs_per_action ? a : 0) * g.O * g.Sp;
        // products are rounded before they are added, like the Gamma rows the three-call path stores (no FMA contraction)
        if (dx == 0) {
            const int d4 = y2 * g.W4 + c4;                     // the destination cells of these four columns: one aligned float4
            for (int o = 0; o < g.O; ++o) {
                const int gi = best_g[(b * g.A + a) * g.O + o];
                const float4 av = __ldg(reinterpret_cast<const float4*>(alpha + (long long)gi * ld_alpha) + d4);
                const float4 l = __ldg(reinterpret_cast<const float4*>(ob + (size_t)o * g.Sp) + d4);
                const bool px = 4 * c4 + 0 < g.W, py = 4 * c4 + 1 < g.W, pz = 4 * c4 + 2 < g.W, pw = 4 * c4 + 3 < g.W;
                acc.x = __fadd_rn(acc.x, px ? __fmul_rn(__fmul_rn(discount, av.x), l.x) : 0.f);
                acc.y = __fadd_rn(acc.y, py ? __fmul_rn(__fmul_rn(discount, av.y), l.y) : 0.f);
                acc.z = __fadd_rn(acc.z, pz ? __fmul_rn(__fmul_rn(discount, av.z), l.z) : 0.f);
                acc.w = __fadd_rn(acc.w, pw ? __fmul_rn(__fmul_rn(discount, av.w), l.w) : 0.f);
            }
        } else {
            int d[4];
            bool real[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int x = 4 * c4 + e;
                real[e] = x < g.W;
                d[e] = y2 * g.Wp + bmove_dev(real[e] ? x : 0, dx, g.W, g.wrap_x);
            }
            for (int o = 0; o < g.O; ++o) {
                const int gi = best_g[(b * g.A + a) * g.O + o];
                const float* ar = alpha + (long long)gi * ld_alpha;
                const float* lr = ob + (size_t)o * g.Sp;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = real[e] ? __fmul_rn(__fmul_rn(discount, __ldg(ar + d[e])), __ldg(lr + d[e])) : 0.f;
                acc.x = __fadd_rn(acc.x, v[0]); acc.y = __fadd_rn(acc.y, v[1]); acc.z = __fadd_rn(acc.z, v[2]); acc.w = __fadd_rn(acc.w, v[3]);
            }
        }
    }
    *reinterpret_cast<float4*>(out + b * ld_out + 4 * (size_t)c) = acc;
}

// blockIdx.z = segment: source rows seg * rows .. + rows, destination rows seg * gridDim.y .. (segments padded to gridDim.y rows)
__global__ void split_planes_kernel(const float* __restrict__ src, long long ld_src, long long rows, int K,
                                    float* __restrict__ hi, float* __restrict__ lo, long long ld) {
    const long long r = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ld) return;
    src += (long long)blockIdx.z * rows * ld_src;
    hi += (long long)blockIdx.z * gridDim.y * ld;
    lo += (long long)blockIdx.z * gridDim.y * ld;
    const float v = (r < rows && c < K) ? src[r * ld_src + c] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[r * ld + c] = h;
    lo[r * ld + c] = __uint_as_float(__float_as_uint(v - h) & 0xffffe000u);
}

}  // namespace

// ======================================================================================================
extern "C" int olfnav_env_step(const olfnav_env* e, int64_t n, int32_t* pos, const int32_t* act, const int32_t* moves,
                               const int64_t* time_shift, int64_t step, const double* thresholds, int n_thr,
                               double* odor, int32_t* obs_id, uint8_t* reached, void* stream) {
    OLF_REQUIRE(e && pos && act && moves && time_shift && thresholds && odor && obs_id && reached, "env_step: null argument");
    OLF_REQUIRE(n >= 0 && n <= 0x7fffffffLL && n_thr >= 2, "env_step: bad n or n_thr");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(e->device));
    env_step_kernel<<<olf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(*e, (int)n, pos, act, moves, (const long long*)time_shift,
                                                                         (long long)step, thresholds, n_thr, odor, obs_id, reached);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_vi_sweep(const olfnav_model* m, const double* Q_in, double* Q_out, int64_t ld_q, double discount,
                               double* max_change, void* stream) {
    OLF_REQUIRE(m && Q_in && Q_out && max_change && ld_q >= m->g.Sp, "vi_sweep: bad argument");
    OLF_CUDA(cudaSetDevice(m->device));
    if (m->dense) return olf_dense_vi_sweep(m, Q_in, Q_out, ld_q, discount, max_change, (cudaStream_t)stream);
    vi_sweep_kernel<<<olf_div_up(m->g.Sp, 256), 256, 0, (cudaStream_t)stream>>>(m->g, m->reward, Q_in, Q_out, ld_q, discount, max_change);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_backup_gamma(const olfnav_model* m, const float* alpha, int64_t ld_alpha, int G, float discount,
                                   float* gamma_ao, int64_t ld_gamma, void* stream) {
    OLF_REQUIRE(m && alpha && gamma_ao && G >= 1, "backup_gamma: bad argument");
    OLF_REQUIRE(ld_alpha >= m->g.Sp && ld_gamma >= m->g.Sp && (ld_gamma % 4) == 0, "backup_gamma: bad leading dimension");
    OLF_REQUIRE(G <= 65535, "backup_gamma: G too large");
    OLF_CUDA(cudaSetDevice(m->device));
    if (m->dense) return olf_dense_backup_gamma(m, alpha, ld_alpha, G, discount, gamma_ao, ld_gamma, (cudaStream_t)stream);
    dim3 grid(olf_div_up(ld_gamma, 256), G, m->g.A);
    backup_gamma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m->g, m->obs, alpha, ld_alpha, G, discount, gamma_ao, ld_gamma);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

// Selection with the Gamma planes built straight from the alpha vectors (gamma_planes_f16_kernel): no FP32 Gamma in memory.
extern "C" int olfnav_backup_select_alphas(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                                           const float* alpha, int64_t ld_alpha, int G, float discount,
                                           int32_t* best_g, int32_t* best_a, float* best_value, void* stream) {
    OLF_REQUIRE(m && b && alpha && best_g && best_a, "backup_select_alphas: null argument");
    OLF_REQUIRE(!m->dense, "backup_select_alphas: grid models only (dense models: olfnav_backup_gamma + olfnav_backup_select)");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0 && ld_alpha >= m->g.Sp, "backup_select_alphas: bad leading dimension");
    OLF_REQUIRE(n >= 0 && n <= 0x7fffffffLL && G >= 1 && G <= 65535, "backup_select_alphas: bad n or G");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int A = m->g.A, O = m->g.O, nseg = A * O;
    const int Gs = (G + 15) & ~15;
    const long long ldh = (m->g.Sp + 63) & ~63;
    OlfScratch s_seg(st), s_rew(st), s_arg(st), s_p1(st), s_p2(st), s_inv(st);
    OLF_CUDA(s_seg.alloc((size_t)n * nseg * sizeof(float)));
    OLF_CUDA(s_rew.alloc((size_t)n * A * sizeof(float)));
    OLF_CUDA(s_arg.alloc((size_t)n * A * sizeof(int)));
    OLF_CUDA(s_p1.alloc((size_t)nseg * Gs * ldh * sizeof(__half)));
    OLF_CUDA(s_p2.alloc((size_t)nseg * Gs * ldh * sizeof(__half)));
    OLF_CUDA(s_inv.alloc((size_t)nseg * Gs * sizeof(float)));
    dim3 grid(1, Gs, nseg);
    gamma_planes_f16_kernel<<<grid, 256, 0, st>>>(m->g, m->obs, alpha, ld_alpha, G, discount, s_p1.as<__half>(), s_p2.as<__half>(), s_inv.as<float>(), ldh);
    OLF_LAUNCH_CHECK();
    int rc = olf_umma_segmax_f16(b, ld, n, m->g.Sp, scale, s_p1.as<__half>(), s_p2.as<__half>(), s_inv.as<float>(), ldh, nseg * Gs, nseg, G, Gs,
                                 s_seg.as<float>(), best_g, m->sm_count, st);
    if (rc == OLFNAV_OK) rc = olf_segmax_simt(b, ld, n, m->g.Sp / 4, m->reward, m->g.Sp, A, 1, s_rew.as<float>(), s_arg.as<int>(), st);
    if (rc == OLFNAV_OK) {
        backup_pick_kernel<<<olf_div_up(n, 256), 256, 0, st>>>((int)n, A, O, scale, s_seg.as<float>(), s_rew.as<float>(), best_a, best_value);
        OLF_LAUNCH_CHECK();
    }
    return rc;
}

extern "C" int olfnav_backup_assemble_alphas(const olfnav_model* m, const float* alpha, int64_t ld_alpha, int G, float discount,
                                             const int32_t* best_g, const int32_t* best_a, int64_t n, float* alpha_new, int64_t ld_new,
                                             void* stream) {
    OLF_REQUIRE(m && alpha && best_g && best_a && alpha_new, "backup_assemble_alphas: null argument");
    OLF_REQUIRE(!m->dense && ld_alpha >= m->g.Sp && (ld_alpha % 4) == 0 && ld_new >= m->g.Sp && (ld_new % 4) == 0 && G >= 1, "backup_assemble_alphas: bad argument");
    OLF_REQUIRE((((uintptr_t)alpha | (uintptr_t)alpha_new) & 15) == 0, "backup_assemble_alphas: 16-byte aligned vectors");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int rows = (int)((n - r0 < 65535) ? (n - r0) : 65535);
        dim3 grid(olf_div_up(ld_new / 4, 256), rows);
        backup_assemble_alphas_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m->g, m->reward, m->obs, alpha, ld_alpha, discount,
                                                                              best_g + r0 * m->g.A * m->g.O, best_a + r0,
                                                                              alpha_new + r0 * ld_new, ld_new);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

extern "C" int olfnav_backup_select(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                                    const float* gamma_ao, int64_t ld_gamma, int G,
                                    int32_t* best_g, int32_t* best_a, float* best_value, int path, void* stream) {
    OLF_REQUIRE(m && b && gamma_ao && best_g && best_a, "backup_select: null argument");
    OLF_REQUIRE(ld >= m->g.Sp && (ld % 4) == 0 && ld_gamma >= m->g.Sp && (ld_gamma % 4) == 0, "backup_select: bad leading dimension");
    OLF_REQUIRE(n >= 0 && n <= 0x7fffffffLL && G >= 1 && path >= 0 && path <= 2, "backup_select: bad n, G or path");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int A = m->g.A, O = m->g.O, nseg = A * O;
    if (path == 0) path = ((long long)G * nseg <= 64) ? 1 : 2;

    // scratch is returned to the pool on every exit path, the failing ones included (OlfScratch)
    OlfScratch s_seg(st), s_rew(st), s_arg(st), s_p1(st), s_p2(st), s_inv(st);
    OLF_CUDA(s_seg.alloc((size_t)n * nseg * sizeof(float)));
    OLF_CUDA(s_rew.alloc((size_t)n * A * sizeof(float)));
    OLF_CUDA(s_arg.alloc((size_t)n * A * sizeof(int)));
    float *seg_val = s_seg.as<float>(), *rew_val = s_rew.as<float>();
    int* rew_arg = s_arg.as<int>();
    int rc = OLFNAV_OK;
    if (path == 1) {
        rc = olf_segmax_simt(b, ld, n, m->g.Sp / 4, gamma_ao, ld_gamma, nseg, G, seg_val, best_g, st);
    } else {
        // split the Gamma rows into TF32 hi/lo planes with each (a,o) segment padded to a multiple of 16 rows
        const int Gs = (G + 15) & ~15;
        const char* f16_env = getenv("OLFNAV_GEMM_F16");           // read per call: "0" selects the 3xTF32 kernel
        // (below G = 128 the 3xTF32 kernel with its narrower tiles is as fast: measured 3.6 vs 4.1 ms at G = 64, B = 10 000)
        if (!(f16_env && f16_env[0] == '0') && (G >= 128 || (f16_env && f16_env[0] == '1'))) {
            // 2-way FP16 split planes with a power-of-two scale per Gamma row (half the tensor time of 3xTF32, same accuracy)
            const long long ldh = (m->g.Sp + 63) & ~63;
            OLF_CUDA(s_p1.alloc((size_t)nseg * Gs * ldh * sizeof(__half)));
            OLF_CUDA(s_p2.alloc((size_t)nseg * Gs * ldh * sizeof(__half)));
            OLF_CUDA(s_inv.alloc((size_t)nseg * Gs * sizeof(float)));
            __half *b1 = s_p1.as<__half>(), *b2 = s_p2.as<__half>();
            float* binv = s_inv.as<float>();
            rc = olf_split_planes_f16(gamma_ao, ld_gamma, G, m->g.Sp, b1, b2, binv, ldh, Gs, nseg, st);
            if (rc == OLFNAV_OK)
                rc = olf_umma_segmax_f16(b, ld, n, m->g.Sp, scale, b1, b2, binv, ldh, nseg * Gs, nseg, G, Gs, seg_val, best_g, m->sm_count, st);
        } else {
        const long long ldp = (m->g.Sp + 31) & ~31;
        OLF_CUDA(s_p1.alloc((size_t)nseg * Gs * ldp * sizeof(float)));
        OLF_CUDA(s_p2.alloc((size_t)nseg * Gs * ldp * sizeof(float)));
        float *hi = s_p1.as<float>(), *lo = s_p2.as<float>();
        {
            dim3 grid(olf_div_up(ldp, 256), Gs, nseg);           // one launch for all (a, o) segments
            split_planes_kernel<<<grid, 256, 0, st>>>(gamma_ao, ld_gamma, G, m->g.Sp, hi, lo, ldp);
            olf_count_launch(1);
            if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("backup_select: split launch failed"); rc = OLFNAV_ERR_CUDA; }
        }
        if (rc == OLFNAV_OK)
            rc = olf_umma_segmax(b, ld, n, m->g.Sp, hi, lo, ldp, nseg * Gs, nseg, G, Gs, seg_val, best_g, m->sm_count, st);
        }
    }
    // b . r_a : A one-row segments (exact FP32)
    if (rc == OLFNAV_OK) rc = olf_segmax_simt(b, ld, n, m->g.Sp / 4, m->reward, m->g.Sp, A, 1, rew_val, rew_arg, st);
    if (rc == OLFNAV_OK) {
        backup_pick_kernel<<<olf_div_up(n, 256), 256, 0, st>>>((int)n, A, O, scale, seg_val, rew_val, best_a, best_value);
        olf_count_launch(1);
        if (cudaGetLastError() != cudaSuccess) { olfnav_set_error("backup_select: pick launch failed"); rc = OLFNAV_ERR_CUDA; }
    }
    return rc;
}

extern "C" int olfnav_backup_assemble(const olfnav_model* m, const float* gamma_ao, int64_t ld_gamma, int G,
                                      const int32_t* best_g, const int32_t* best_a, int64_t n, float* alpha_new, int64_t ld_new,
                                      void* stream) {
    OLF_REQUIRE(m && gamma_ao && best_g && best_a && alpha_new, "backup_assemble: null argument");
    OLF_REQUIRE(ld_gamma >= m->g.Sp && (ld_gamma % 4) == 0 && ld_new >= m->g.Sp && (ld_new % 4) == 0, "backup_assemble: bad leading dimension");
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    for (int64_t r0 = 0; r0 < n; r0 += 65535) {
        const int rows = (int)((n - r0 < 65535) ? (n - r0) : 65535);
        dim3 grid(olf_div_up(ld_new / 4, 256), rows);
        backup_assemble_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(m->g, m->reward, gamma_ao, ld_gamma, G,
                                                                       best_g + r0 * m->g.A * m->g.O, best_a + r0,
                                                                       alpha_new + r0 * ld_new, ld_new);
    }
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}


// ------------------------------------------------------------------------------------------------------
// Row-pair reductions (see olfnav_pair_reduce in the header).  One CTA per (x_i, y_j) pair; both rows stream from L2 / HBM.
namespace {
template <int MODE>
__global__ void __launch_bounds__(256) pair_reduce_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ sx,
                                                          const float* __restrict__ Y, long long ldy, const float* __restrict__ sy, int ny,
                                                          const float* __restrict__ Wt, long long ldw, float vmax, float vmin, int len4,
                                                          float* __restrict__ out) {
    __shared__ float red[33];
    const long long i = blockIdx.y, j = blockIdx.x;
    const float a = sx ? sx[i] : 1.f, b = sy ? sy[j] : 1.f;
    const float4* x4 = reinterpret_cast<const float4*>(X + i * ldx);
    const float4* y4 = reinterpret_cast<const float4*>(Y + j * ldy);
    const float4* w4 = MODE == OLFNAV_PAIR_GER ? reinterpret_cast<const float4*>(Wt + j * ldw) : nullptr;
    float acc = 0.f, acc2 = 0.f;
    for (int c = threadIdx.x; c < len4; c += 256) {
        const float4 xv = x4[c], yv = y4[c];
        const float xs[4] = {xv.x * a, xv.y * a, xv.z * a, xv.w * a};
        const float ys[4] = {yv.x * b, yv.y * b, yv.z * b, yv.w * b};
        if (MODE == OLFNAV_PAIR_L1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc += fabsf(xs[e] - ys[e]);
        } else if (MODE == OLFNAV_PAIR_GER) {
            const float4 wv = w4[c];
            const float ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = xs[e] - ys[e];
                acc += d * ((d >= 0.f ? vmax : vmin) - ws[e]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc += (ys[e] < xs[e]) ? 1.f : 0.f; acc2 += (ys[e] > xs[e]) ? 1.f : 0.f; }
        }
    }
    const float t = block_sum<256>(acc, red);
    if (MODE == OLFNAV_PAIR_DOMINATED) {
        const float t2 = block_sum<256>(acc2, red);
        if (threadIdx.x == 0) out[i * ny + j] = (t == 0.f && t2 > 0.f) ? 1.f : 0.f;
    } else if (threadIdx.x == 0) {
        out[i * ny + j] = t;
    }
}
}  // namespace

extern "C" int olfnav_pair_reduce(int mode, const float* X, int64_t ldx, const float* sx, int64_t nx, const float* Y, int64_t ldy,
                                  const float* sy, int64_t ny, const float* W, int64_t ldw, float vmax, float vmin, int64_t row_len,
                                  float* out, void* stream) {
    OLF_REQUIRE(X && Y && out, "pair_reduce: null argument");
    OLF_REQUIRE(mode >= 0 && mode <= 2 && (mode != OLFNAV_PAIR_GER || W), "pair_reduce: bad mode or missing W");
    OLF_REQUIRE((row_len % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0 && (ldw % 4) == 0, "pair_reduce: rows must be float4 aligned");
    OLF_REQUIRE(nx >= 0 && ny >= 0 && nx <= 65535 && ny <= 0x7fffffffLL, "pair_reduce: nx <= 65535 rows per call");
    if (nx == 0 || ny == 0) return OLFNAV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ny, (unsigned)nx);
    const int len4 = (int)(row_len / 4);
    if (mode == OLFNAV_PAIR_L1)       pair_reduce_kernel<OLFNAV_PAIR_L1><<<grid, 256, 0, st>>>(X, ldx, sx, Y, ldy, sy, (int)ny, W, ldw, vmax, vmin, len4, out);
    else if (mode == OLFNAV_PAIR_GER) pair_reduce_kernel<OLFNAV_PAIR_GER><<<grid, 256, 0, st>>>(X, ldx, sx, Y, ldy, sy, (int)ny, W, ldw, vmax, vmin, len4, out);
    else                              pair_reduce_kernel<OLFNAV_PAIR_DOMINATED><<<grid, 256, 0, st>>>(X, ldx, sx, Y, ldy, sy, (int)ny, W, ldw, vmax, vmin, len4, out);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
