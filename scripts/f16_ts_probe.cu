// Probe for a round-2 idea: the policy / backup GEMM with a 2-way FP16 split instead of 3xTF32
// (3 tcgen05.mma.kind::f16 per K = 16 instead of 3 kind::tf32 per K = 8: half the tensor time; accuracy study in
// scripts/fp16_split_study.py).  One CTA, one 128 x 64 x 64 tile.  Answers, on the real hardware:
//   1. how 16-bit A elements are packed in tensor memory for the TS form (element 2c in the low or the high half of column c),
//   2. that a 128B-swizzled K-major FP16 B tile takes the same shared-memory descriptor and the same +32 B K-step as the FP32 one,
//   3. whether `scale-input-d` (D = A.B + D * 2^-11) is accepted and exact, which would fold the two accumulators into one.
// Build: nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -I olfactory-navigation_b200/csrc scripts/f16_ts_probe.cu -o scripts/f16_ts_probe.bin
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma_ptx.cuh"

constexpr int PBN = 64, PK = 64;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
#ifdef PROBE_SCALE_D
__device__ __forceinline__ void mma_f16_ts_scaled(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p, 11;\n\t}"
                 :: "r"(d), "r"(a), "l"(bdesc), "r"(idesc) : "memory");
}
#endif

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, float a_scale, const __half* __restrict__ B1, const __half* __restrict__ B2,
             float* __restrict__ out0, float* __restrict__ out1, float* __restrict__ out2, int high_first) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sb1 = base, sb2 = base + PBN * 128, bar = base + 2 * PBN * 128;
    uint32_t* slot = reinterpret_cast<uint32_t*>(g + 2 * PBN * 128 + 16);
    const int t = threadIdx.x, warp = t >> 5;

    if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // B planes -> shared memory, K-major rows of 64 halfs = 128 B, 16-byte piece c of row r at piece c ^ (r & 7)
    for (int i = t; i < PBN * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(g + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B1 + r * PK + c * 8);
        *reinterpret_cast<uint4*>(g + PBN * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B2 + r * PK + c * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    constexpr uint32_t D0 = 0, D1 = 64, D2 = 128, A1 = 256, A2 = 288;

    // A row t -> a1 = fp16(x), a2 = fp16((x - a1) * 2^11), packed two per 32-bit column
    {
        uint32_t h[32], l[32];
        for (int c = 0; c < 32; ++c) {
            const float x0 = A[t * PK + 2 * c] * a_scale, x1 = A[t * PK + 2 * c + 1] * a_scale;
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            const __half l0 = __float2half_rn((x0 - __half2float(h0)) * 2048.f), l1 = __float2half_rn((x1 - __half2float(h1)) * 2048.f);
            const uint32_t uh0 = __half_as_ushort(h0), uh1 = __half_as_ushort(h1), ul0 = __half_as_ushort(l0), ul1 = __half_as_ushort(l1);
            h[c] = high_first ? ((uh0 << 16) | uh1) : ((uh1 << 16) | uh0);
            l[c] = high_first ? ((ul0 << 16) | ul1) : ((ul1 << 16) | ul0);
        }
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        uint32_t v[16];
        for (int part = 0; part < 2; ++part) {
            for (int j = 0; j < 16; ++j) v[j] = h[part * 16 + j];
            tc_st16(tm + lane_base + A1 + part * 16, v);
            for (int j = 0; j < 16; ++j) v[j] = l[part * 16 + j];
            tc_st16(tm + lane_base + A2 + part * 16, v);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 1) {
        if (elect_one_sync()) {
            // c_format F32 (1) @4, a_format F16 (0) @7, b_format F16 (0) @10, K-major, N>>3 @17, M>>4 @24
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(PBN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t d1 = make_sw128_desc(sb1), d2 = make_sw128_desc(sb2);
            for (int kk = 0; kk < PK / 16; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 32 / 16);
                mma_f16_ts(tm + D0, tm + A1 + kk * 8, d1 + adv, idesc, kk > 0);
                mma_f16_ts(tm + D1, tm + A1 + kk * 8, d2 + adv, idesc, kk > 0);
                mma_f16_ts(tm + D1, tm + A2 + kk * 8, d1 + adv, idesc, 1);
            }
#ifdef PROBE_SCALE_D
            // single accumulator: cross terms first, then D2 = a1.b1 + D2 * 2^-11, then the remaining main terms
            for (int kk = 0; kk < PK / 16; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 32 / 16);
                mma_f16_ts(tm + D2, tm + A1 + kk * 8, d2 + adv, idesc, kk > 0);
                mma_f16_ts(tm + D2, tm + A2 + kk * 8, d1 + adv, idesc, 1);
            }
            mma_f16_ts_scaled(tm + D2, tm + A1, d1, idesc);
            for (int kk = 1; kk < PK / 16; ++kk) mma_f16_ts(tm + D2, tm + A1 + kk * 8, d1 + (uint64_t)(kk * 32 / 16), idesc, 1);
#endif
            tc_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    {
        uint32_t v[32];
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int part = 0; part < 2; ++part) {
            tc_ld32(tm + lane_base + D0 + part * 32, v);
            for (int j = 0; j < 32; ++j) out0[t * PBN + part * 32 + j] = __uint_as_float(v[j]);
            tc_ld32(tm + lane_base + D1 + part * 32, v);
            for (int j = 0; j < 32; ++j) out1[t * PBN + part * 32 + j] = __uint_as_float(v[j]);
            tc_ld32(tm + lane_base + D2 + part * 32, v);
            for (int j = 0; j < 32; ++j) out2[t * PBN + part * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
    }
}

int main() {
    std::vector<float> A(128 * PK), B(PBN * PK);
    srand(3);
    for (auto& v : A) { const double u = rand() / (double)RAND_MAX; v = (float)(exp(-30.0 * u) * (rand() % 7 != 0)); }   // wide dynamic range, some zeros
    for (auto& v : B) v = (float)(100.0 * pow(rand() / (double)RAND_MAX, 3.0));
    const float a_scale = 16384.f, b_scale = 128.f;
    std::vector<__half> B1(PBN * PK), B2(PBN * PK);
    for (int i = 0; i < PBN * PK; ++i) {
        const float y = B[i] * b_scale;
        B1[i] = __float2half_rn(y);
        B2[i] = __float2half_rn((y - __half2float(B1[i])) * 2048.f);
    }
    float *dA, *o0, *o1, *o2; __half *dB1, *dB2;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB1, B1.size() * 2); cudaMalloc(&dB2, B2.size() * 2);
    cudaMalloc(&o0, 128 * PBN * 4); cudaMalloc(&o1, 128 * PBN * 4); cudaMalloc(&o2, 128 * PBN * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB1, B1.data(), B1.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB2, B2.data(), B2.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 2 * PBN * 128 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int high_first = 0; high_first < 2; ++high_first) {
        cudaMemset(o2, 0, 128 * PBN * 4);
        probe_kernel<<<1, 128, smem>>>(dA, a_scale, dB1, dB2, o0, o1, o2, high_first);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("high_first=%d: %s\n", high_first, cudaGetErrorString(e)); return 1; }
        std::vector<float> h0(128 * PBN), h1(128 * PBN), h2(128 * PBN);
        cudaMemcpy(h0.data(), o0, h0.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost);
        double e_main = 0, e_cross = 0, e_total = 0, e_single = 0;
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < PBN; ++c) {
                double m = 0, x = 0, exact = 0, norm = 0;
                for (int k = 0; k < PK; ++k) {
                    const float xa = A[r * PK + k] * a_scale;
                    const __half a1 = __float2half_rn(xa);
                    const __half a2 = __float2half_rn((xa - __half2float(a1)) * 2048.f);
                    m += (double)__half2float(a1) * __half2float(B1[c * PK + k]);
                    x += (double)__half2float(a1) * __half2float(B2[c * PK + k]) + (double)__half2float(a2) * __half2float(B1[c * PK + k]);
                    exact += (double)A[r * PK + k] * B[c * PK + k];
                    norm += fabs((double)A[r * PK + k] * B[c * PK + k]);
                }
                const double sc = (double)a_scale * b_scale;
                e_main = fmax(e_main, fabs(h0[r * PBN + c] - m) / (fabs(m) + 1e-30));
                e_cross = fmax(e_cross, fabs(h1[r * PBN + c] - x) / (fabs(x) + 1e-30));
                e_total = fmax(e_total, fabs((h0[r * PBN + c] + h1[r * PBN + c] / 2048.0) / sc - exact) / (norm + 1e-30));
                e_single = fmax(e_single, fabs(h2[r * PBN + c] / sc - exact) / (norm + 1e-30));
            }
        printf("A packing %s: main rel err %.2e, cross rel err %.2e, combined |err|/sum|a||b| %.2e, single-accumulator (scale-input-d) %.2e\n",
               high_first ? "element 2c in the HIGH half" : "element 2c in the LOW half", e_main, e_cross, e_total, e_single);
    }
    return 0;
}
