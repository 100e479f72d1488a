// Dense stochastic-transition models — what the reference's minimal_converter builds
// (olfactory_navigation/agents/model_based_util/environment_converter.py:137-292): a few dozen partition cells + one goal
// state, POMDP(transition_table = T[s, a, s'], observation_table = O[s', a, o]).  No grid, no stencil: the transition is a small
// dense matrix per action, kept in L2 / shared memory.  These kernels are the generic counterparts of the grid kernels
//   belief step   b'(s') = O[s', a, o] * sum_s T[s, a, s'] b(s)                      (BeliefSet.update, agents/pbvi_agent.py:783-787)
//   VI sweep      Q_a(s) = r_a(s) + gamma * sum_s' T[s, a, s'] max_a' Q_a'(s')       (solvers.VI, agents/qmdp_agent.py:145-153)
//   backup Gamma  Gamma_{a,o}^g(s) = gamma * sum_s' T[s, a, s'] O[s', a, o] alpha_g(s')
//   infotaxis     Delta H_a = H(b) - sum_o p(o | b, a) H(b_{a,o})                     (agents/infotaxis_agent.py:257-301)
// behind the same C-ABI entry points (dispatch on olfnav_model::dense).  The models have S <= a few hundred states: the work is
// latency bound whatever is done, so the kernels are kept simple (one CTA per agent / per output row).
#include "common.cuh"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace {

constexpr int DN_THREADS = 128;

// one CTA per agent: b (scaled) in shared memory, thread s' owns destination states s', s' + 128, ...
__global__ void __launch_bounds__(DN_THREADS)
dense_belief_step_kernel(const Geom g, const float* __restrict__ Tpush, const float* __restrict__ obs,
                         const float* __restrict__ b_in, float* __restrict__ b_out, long long ld,
                         const int* __restrict__ act, const int* __restrict__ ob, const int* __restrict__ src_rows,
                         const int* __restrict__ dst_rows, const float* __restrict__ scale_in, float* __restrict__ scale_out,
                         unsigned char* __restrict__ ok, const int* __restrict__ n_dev) {
    extern __shared__ float sh[];          // Sp floats + 40
    float* red = sh + g.Sp;
    const int i = blockIdx.x;
    if (n_dev && i >= *n_dev) return;
    const long long src = src_rows ? src_rows[i] : i, dst = dst_rows ? dst_rows[i] : i;
    int a = act[i], o = ob[i];
    float sc = scale_in ? scale_in[src] : 1.f;
    if (a < 0 || a >= g.A || o < 0 || o >= g.O) { a = 0; o = 0; sc = 0.f; }
    for (int s = threadIdx.x; s < g.Sp; s += DN_THREADS) sh[s] = (s < g.S) ? b_in[src * ld + s] * sc : 0.f;
    __syncthreads();
    float z = 0.f;
    for (int sp = threadIdx.x; sp < g.Sp; sp += DN_THREADS) {
        float u = 0.f;
        if (sp < g.S) {
            const float* trow = Tpush + ((size_t)a * g.S + sp) * g.Sp;      // T[s, a, s'] for fixed (a, s'), contiguous in s
            for (int s = 0; s < g.S; ++s) u = fmaf(trow[s], sh[s], u);
            u *= obs[((size_t)a * g.O + o) * g.Sp + sp];
        }
        z += u;
        b_out[dst * ld + sp] = u;
    }
    z = block_sum<DN_THREADS>(z, red);
    if (threadIdx.x == 0) {
        const bool good = z > 0.f && isfinite(z);
        scale_out[dst] = good ? 1.f / z : 0.f;
        ok[i] = good ? 1 : 0;
    }
}

__global__ void dense_vi_sweep_kernel(const Geom g, const float* __restrict__ Tpull, const float* __restrict__ reward,
                                      const double* __restrict__ Qin, double* __restrict__ Qout, long long ldq, double discount,
                                      double* __restrict__ max_change) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    double change = 0.0;
    if (s < g.Sp) {
        if (s < g.S) {
            double vin = -1e300, vout = -1e300;
            for (int a = 0; a < g.A; ++a) vin = fmax(vin, Qin[a * ldq + s]);
            for (int a = 0; a < g.A; ++a) {
                double acc = 0.0;
                const float* trow = Tpull + ((size_t)a * g.S + s) * g.Sp;   // T[s, a, :]
                for (int sp = 0; sp < g.S; ++sp) {
                    const double t = trow[sp];
                    if (t != 0.0) {
                        double vd = -1e300;
                        for (int a2 = 0; a2 < g.A; ++a2) vd = fmax(vd, Qin[a2 * ldq + sp]);
                        acc += t * vd;
                    }
                }
                const double q = (double)reward[(size_t)a * g.Sp + s] + discount * acc;
                Qout[a * ldq + s] = q;
                vout = fmax(vout, q);
            }
            change = fabs(vout - vin);
        } else {
            for (int a = 0; a < g.A; ++a) Qout[a * ldq + s] = 0.0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) change = fmax(change, __shfl_xor_sync(0xffffffffu, change, o));
    if ((threadIdx.x & 31) == 0 && change > 0.0)
        atomicMax(reinterpret_cast<unsigned long long*>(max_change), (unsigned long long)__double_as_longlong(change));
}

// out[(a, o, g)][s] = discount * sum_s' T[s, a, s'] O[s', a, o] alpha_g(s')
__global__ void dense_backup_gamma_kernel(const Geom g, const float* __restrict__ Tpull, const float* __restrict__ obs,
                                          const float* __restrict__ alpha, long long ld_alpha, int G, float discount,
                                          float* __restrict__ out, long long ld_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int gi = blockIdx.y, a = blockIdx.z;
    if (s >= ld_out) return;
    for (int o = 0; o < g.O; ++o) {
        float acc = 0.f;
        if (s < g.S) {
            const float* trow = Tpull + ((size_t)a * g.S + s) * g.Sp;
            const float* orow = obs + ((size_t)a * g.O + o) * g.Sp;
            const float* al = alpha + (long long)gi * ld_alpha;
            for (int sp = 0; sp < g.S; ++sp) acc = fmaf(trow[sp] * orow[sp], al[sp], acc);
            acc *= discount;
        }
        out[((long long)(a * g.O + o) * G + gi) * ld_out + s] = acc;
    }
}

// one CTA per agent; natural logarithm, 0 log 0 = 0; first maximum wins (strict '<' at agents/infotaxis_agent.py:289)
__global__ void __launch_bounds__(DN_THREADS)
dense_infotaxis_kernel(const Geom g, const float* __restrict__ Tpush, const float* __restrict__ obs, const float* __restrict__ b,
                       long long ld, const float* __restrict__ scale, int* __restrict__ action, float* __restrict__ dH) {
    extern __shared__ float sh[];          // b (Sp) | u_a (Sp) | red (40)
    float* ua = sh + g.Sp;
    float* red = ua + g.Sp;
    const long long i = blockIdx.x;
    const float sc = scale ? scale[i] : 1.f;
    float hb = 0.f;
    for (int s = threadIdx.x; s < g.Sp; s += DN_THREADS) {
        const float v = (s < g.S) ? b[i * ld + s] * sc : 0.f;
        sh[s] = v;
        if (v > 0.f) hb -= v * logf(v);
    }
    __syncthreads();
    hb = block_sum<DN_THREADS>(hb, red);
    float best = -INFINITY;
    int best_a = 0;
    for (int a = 0; a < g.A; ++a) {
        for (int sp = threadIdx.x; sp < g.Sp; sp += DN_THREADS) {
            float u = 0.f;
            if (sp < g.S) {
                const float* trow = Tpush + ((size_t)a * g.S + sp) * g.Sp;
                for (int s = 0; s < g.S; ++s) u = fmaf(trow[s], sh[s], u);
            }
            ua[sp] = u;
        }
        __syncthreads();
        float expected = 0.f;
        for (int o = 0; o < g.O; ++o) {
            const float* orow = obs + ((size_t)a * g.O + o) * g.Sp;
            float p = 0.f, vlogv = 0.f;
            for (int sp = threadIdx.x; sp < g.S; sp += DN_THREADS) {
                const float v = ua[sp] * orow[sp];
                p += v;
                if (v > 0.f) vlogv += v * logf(v);
            }
            p = block_sum<DN_THREADS>(p, red);
            vlogv = block_sum<DN_THREADS>(vlogv, red);
            if (p > 0.f) expected += p * logf(p) - vlogv;          // p * H(b_ao) = p log p - sum v log v
        }
        const float d = hb - expected;
        if (dH && threadIdx.x == 0) dH[i * g.A + a] = d;
        if (d > best) { best = d; best_a = a; }
        __syncthreads();
    }
    if (threadIdx.x == 0) action[i] = best_a;
}

}  // namespace

// ======================================================================================================
extern "C" int olfnav_model_create_dense(olfnav_model** out, int device, int S, int A, int O,
                                         const double* transition_table, const double* obs_table,
                                         const uint8_t* end_mask, const double* start) {
    OLF_REQUIRE(out && transition_table && obs_table && end_mask && start, "model_create_dense: null argument");
    OLF_REQUIRE(S >= 1 && S <= 4096 && A >= 1 && A <= OLFNAV_MAX_ACTIONS && O >= 1, "model_create_dense: S <= 4096, A <= %d", OLFNAV_MAX_ACTIONS);
    int ndev = 0;
    OLF_CUDA(cudaGetDeviceCount(&ndev));
    OLF_REQUIRE(device >= 0 && device < ndev, "model_create_dense: device %d of %d", device, ndev);
    OLF_CUDA(cudaSetDevice(device));
    olfnav_model* m = new (std::nothrow) olfnav_model();
    if (!m) { olfnav_set_error("model_create_dense: out of host memory"); return OLFNAV_ERR_NOMEM; }
    memset(m, 0, sizeof(*m));
    m->device = device;
    m->dense = 1;
    Geom& g = m->g;
    g.H = 1; g.W = S; g.Wp = (S + 3) & ~3; g.W4 = g.Wp / 4; g.S = S; g.Sp = g.Wp;
    g.A = A; g.O = O; g.wrap_y = 0; g.wrap_x = 0; g.obs_per_action = 1;
    g.div_w4 = (g.W4 > 1) ? (unsigned)(((1ULL << 32) + g.W4 - 1) / g.W4) : 0u;
    const int Sp = g.Sp;
    std::vector<float> h_pull((size_t)A * S * Sp, 0.f), h_push((size_t)A * S * Sp, 0.f), h_obs((size_t)A * O * Sp, 0.f),
        h_ent((size_t)A * Sp, 0.f), h_rew((size_t)A * Sp, 0.f), h_start(Sp, 0.f);
    std::vector<uint8_t> h_end(Sp, 0);
    for (int s = 0; s < S; ++s) {
        h_start[s] = (float)start[s];
        h_end[s] = end_mask[s] ? 1 : 0;
        for (int a = 0; a < A; ++a) {
            double rew = 0.0, c = 0.0;
            for (int sp = 0; sp < S; ++sp) {
                const double t = transition_table[((size_t)s * A + a) * S + sp];
                h_pull[((size_t)a * S + s) * Sp + sp] = (float)t;
                h_push[((size_t)a * S + sp) * Sp + s] = (float)t;
                if (end_mask[sp]) rew += t;
            }
            h_rew[(size_t)a * Sp + s] = (float)rew;
            for (int o = 0; o < O; ++o) {
                const double pr = obs_table[((size_t)s * A + a) * O + o];
                h_obs[((size_t)a * O + o) * Sp + s] = (float)pr;
                if (pr > 0.0) c += pr * log(pr);
            }
            h_ent[(size_t)a * Sp + s] = (float)c;
        }
    }
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dptr, const void* src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dptr, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
    };
    up((void**)&m->obs, h_obs.data(), h_obs.size() * sizeof(float));
    up((void**)&m->obs_ent, h_ent.data(), h_ent.size() * sizeof(float));
    up((void**)&m->reward, h_rew.data(), h_rew.size() * sizeof(float));
    up((void**)&m->start, h_start.data(), h_start.size() * sizeof(float));
    up((void**)&m->end_mask, h_end.data(), h_end.size());
    up((void**)&m->t_pull, h_pull.data(), h_pull.size() * sizeof(float));
    up((void**)&m->t_push, h_push.data(), h_push.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        olfnav_set_error("model_create_dense: %s", cudaGetErrorString(e));
        olfnav_model_destroy(m);
        return OLFNAV_ERR_CUDA;
    }
    *out = m;
    return OLFNAV_OK;
}

int olf_dense_belief_step(const olfnav_model* m, const float* b_in, float* b_out, int64_t ld, int64_t n, const int32_t* act,
                          const int32_t* obs, const int32_t* src_rows, const int32_t* dst_rows, const float* scale_in,
                          float* scale_out, uint8_t* ok, const int* n_dev, cudaStream_t st) {
    // the row is staged in shared memory before anything is written: in place (b_out == b_in, same rows) is safe
    const size_t smem = (size_t)(m->g.Sp + 40) * sizeof(float);
    dense_belief_step_kernel<<<(unsigned)n, DN_THREADS, smem, st>>>(m->g, m->t_push, m->obs, b_in, b_out, ld, act, obs, src_rows, dst_rows,
                                                                    scale_in, scale_out, ok, n_dev);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

int olf_dense_vi_sweep(const olfnav_model* m, const double* Q_in, double* Q_out, int64_t ld_q, double discount, double* max_change,
                       cudaStream_t st) {
    dense_vi_sweep_kernel<<<olf_div_up(m->g.Sp, 128), 128, 0, st>>>(m->g, m->t_pull, m->reward, Q_in, Q_out, ld_q, discount, max_change);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

int olf_dense_backup_gamma(const olfnav_model* m, const float* alpha, int64_t ld_alpha, int G, float discount, float* gamma_ao,
                           int64_t ld_gamma, cudaStream_t st) {
    dim3 grid(olf_div_up(ld_gamma, 128), G, m->g.A);
    dense_backup_gamma_kernel<<<grid, 128, 0, st>>>(m->g, m->t_pull, m->obs, alpha, ld_alpha, G, discount, gamma_ao, ld_gamma);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}

int olf_dense_infotaxis(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale, int32_t* action, float* dH,
                        cudaStream_t st) {
    const size_t smem = (size_t)(2 * m->g.Sp + 40) * sizeof(float);
    dense_infotaxis_kernel<<<(unsigned)n, DN_THREADS, smem, st>>>(m->g, m->t_push, m->obs, b, ld, scale, action, dH);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
