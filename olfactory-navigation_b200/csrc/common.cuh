// Shared device/host helpers for the olfnav_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/olfnav_b200.h"

#ifndef OLFNAV_MAX_ACTIONS
#error "header not found"
#endif

// ---- error plumbing (never throw across the C ABI) ------------------------------------------------
void olfnav_set_error(const char* fmt, ...);
void olf_count_launch(int n);

#define OLF_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            olfnav_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OLFNAV_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)

#define OLF_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            olfnav_set_error(__VA_ARGS__);                                                     \
            return OLFNAV_ERR_INVALID;                                                         \
        }                                                                                      \
    } while (0)

#define OLF_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        olf_count_launch(1);                                                                   \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            olfnav_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OLFNAV_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)

// Stream-ordered scratch that goes back to the pool on EVERY exit path (run_test catches out-of-memory and retries with more
// batches, reference simulation.py:1346-1367: a failed attempt must not leak what it had already taken).
struct OlfScratch {
    cudaStream_t st;
    void* p = nullptr;
    explicit OlfScratch(cudaStream_t s) : st(s) {}
    OlfScratch(const OlfScratch&) = delete;
    OlfScratch& operator=(const OlfScratch&) = delete;
    ~OlfScratch() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, st); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// Function attributes (opt-in dynamic shared memory) belong to a device context: one flag per device, not per process.
constexpr int OLF_MAX_DEVICES = 64;
struct OlfPerDeviceOnce {
    bool done[OLF_MAX_DEVICES] = {};
    // true the first time it is asked on the current device
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= OLF_MAX_DEVICES) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

// ---- grid geometry shared by every kernel ------------------------------------------------------------
struct Geom {
    int H, W, Wp, W4;      // Wp = W rounded up to 4, W4 = Wp / 4 (float4 per grid row)
    int S, Sp;             // S = H*W (reference states), Sp = H*Wp (padded)
    int A, O;
    int wrap_y, wrap_x;    // boundary per axis: 1 = wrap, 0 = clip ("stop")
    int obs_per_action;    // 1: observation table differs per action (layered); 0: shared
    unsigned div_w4;       // ceil(2^32 / W4): q = umulhi(n, div_w4) for n < 2^16
    int moves[OLFNAV_MAX_ACTIONS][2];   // (dy, dx)
};

struct olfnav_model {
    int device;
    Geom g;
    float* obs;        // [A_eff][O][Sp]   P(o | arrive s'), zero in padding columns
    float* obs_ent;    // [A_eff][Sp]      C(s') = sum_o O_o(s') log O_o(s')   (infotaxis)
    double* obs_ent64; // the same in double (float64 re-evaluation of near-tied infotaxis choices)
    int* it_unsure;    // [1] agents re-evaluated by the last infotaxis_choose call
    float* reward;     // [A][Sp]          r_a(s) = 1[R_a(s) in end states]
    float* start;      // [Sp]
    uint8_t* end_mask; // [Sp]
    int sm_count;
    // infotaxis "pull" tables as GEMM operand planes (TF32 hi / lo, see evaluate.cu): row a*(O+1)+o = O_o(R_a(s)) for o < O,
    // row a*(O+1)+O = C(R_a(s)); pull_rows = A*(O+1), stored rows rounded up to 16, row stride pull_ld (multiple of 32)
    float* pull_hi;
    float* pull_lo;
    int pull_rows, pull_rows_p;
    int64_t pull_ld;
    // one-pass update + policy kernel (fused_step2.cu): 0 / 1 masks per action, [A][2][Sp]: coef[a][0][s'] = 1 when the cell one x-move
    // before s' exists (its mass arrives at s'), coef[a][1][s'] = 1 when s' is the clipped edge cell of the move (it keeps its own mass)
    float* coef;
    // dense stochastic-transition models (dense.cu; reference minimal_converter): T[s, a, s'] as [A][S][Sp] planes,
    // t_pull rows = source state s (contiguous in s'), t_push rows = destination state s' (contiguous in s)
    int dense;
    float* t_pull;
    float* t_push;
};

struct olfnav_alphas {
    int device;
    int G, Gp;           // Gp = G rounded up to the tensor-core tile granularity
    int64_t ld;          // row stride (floats) of the planes, multiple of 32
    float* hi;           // [Gp][ld]  tf32-representable high part (low 13 mantissa bits zero)
    float* lo;           // [Gp][ld]  alpha - hi, also tf32-representable
    float* full;         // [Gp][ld]  fp32 copy (SIMT path)
    int32_t* actions;    // [G]
    // 2-way FP16 split planes with a power-of-two scale per alpha vector (umma_f16_kernel), built when G >= 64
    void* h1;            // [Gp][ldh] fp16(t_g alpha)
    void* h2;            // [Gp][ldh] fp16((t_g alpha - h1) 2^11)
    float* binv;         // [Gp] 1 / t_g
    int64_t ldh;         // row stride (halfs), multiple of 64
    int f16_eval;        // evaluate_at takes the FP16 split kernel (G >= 64; below that both kernels are HBM bound)
};

struct olfnav_env {
    int device;
    int H, W, wrap_y, wrap_x;
    int T, h, w, my, mx;
    int L;             // layers (1 without layers)
    double sy, sx, r2;
    double* data;      // [L][T][h][w]
};

// ---- small device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_stream1(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in every thread.  `red` holds >= 33 floats.
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = (lane < THREADS / 32) ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    float r = red[32];
    __syncthreads();
    return r;
}

static inline int olf_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- internal entry points shared between translation units ---------------------------------------------
int olf_belief_step_launch(const olfnav_model* m, const float* b_in, float* b_out, int64_t ld, int64_t n,
                           const int32_t* act, const int32_t* obs, const int32_t* src_rows, const int32_t* dst_rows,
                           const float* scale_in, float* scale_out, uint8_t* ok, const int* n_dev, void* stream);
// evaluate with caller-provided scratch (val, idx: n entries each); any of value / arg / action may be NULL
int olf_evaluate_into(const olfnav_model* m, const olfnav_alphas* a, const float* b, int64_t ld, int64_t n, const float* scale,
                      float* val, int* idx, float* value, int32_t* arg, int32_t* action, int path, cudaStream_t st);
// one-pass belief update + evaluation (fused_step2.cu); n = upper bound of the rows, n_dev (optional) = exact device-side count;
// src_rows[i] = agent slot of input i (NULL = i), in_rows[slot] = its row in b_in (NULL = slot)
int olf_fused_step_launch(const olfnav_model* m, const olfnav_alphas* a, const float* b_in, int64_t rows_in, float* b_out, int64_t rows_out,
                          int64_t ld, int64_t n, const int32_t* act, const int32_t* obs, const int32_t* src_rows, const int32_t* in_rows,
                          const float* scale_in, float* scale_out, uint8_t* ok, float* value, int32_t* arg, int32_t* action,
                          int32_t* row_out, const int* n_dev, cudaStream_t st);
// dense-model kernels (dense.cu)
int olf_dense_belief_step(const olfnav_model* m, const float* b_in, float* b_out, int64_t ld, int64_t n, const int32_t* act,
                          const int32_t* obs, const int32_t* src_rows, const int32_t* dst_rows, const float* scale_in,
                          float* scale_out, uint8_t* ok, const int* n_dev, cudaStream_t st);
int olf_dense_vi_sweep(const olfnav_model* m, const double* Q_in, double* Q_out, int64_t ld_q, double discount, double* max_change,
                       cudaStream_t st);
int olf_dense_backup_gamma(const olfnav_model* m, const float* alpha, int64_t ld_alpha, int G, float discount, float* gamma_ao,
                           int64_t ld_gamma, cudaStream_t st);
int olf_dense_infotaxis(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale, int32_t* action, float* dH,
                        cudaStream_t st);
