// Micro-benchmark: issue-to-completion rate of tcgen05.mma.kind::tf32 (M = 128, K = 8) for small N, operands in shared
// memory (128B swizzle, K-major), one CTA per SM.  Prints cycles per MMA for chains of dependent MMAs on one accumulator.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a scripts/umma_rate_bench.cu -o scripts/umma_rate_bench.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000LL) { printf("watchdog\n"); __trap(); } }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tc_mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) { return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }

// mode 0: `reps` dependent MMAs, one commit at the end.  mode 1: commit after every `per` MMAs (waiting only at the end).
// mode 2: two accumulators alternating every MMA.
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int reps, int mode, int per, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t* g = smem + (base - smem_u32(smem));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar_store[2];
    for (int i = threadIdx.x; i < (16384 * 2 + 32768 * 2) / 4; i += blockDim.x) reinterpret_cast<float*>(g)[i] = 1.0f;
    const uint32_t bar = smem_u32(&bar_store[0]);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x < 32 && elect_one_sync()) {
        const uint64_t da = make_sw128_desc(base), da2 = make_sw128_desc(base + 16384);
        const uint64_t db = make_sw128_desc(base + 32768), db2 = make_sw128_desc(base + 65536);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint64_t adv = (uint64_t)((r & 3) * 2);
            const uint32_t d = tmem + ((mode == 2) ? (r & 1) * 256 : 0);
            tc_mma_tf32(d, ((r % 3) == 1 ? da2 : da) + adv, ((r % 3) == 2 ? db2 : db) + adv, idesc, r > 1 ? 1u : 0u);
            if (mode == 1 && (r % per) == per - 1 && r != reps - 1) tc_commit(smem_u32(&bar_store[1]));   // nobody waits on this one
        }
        const long long t1 = clock64();
        tc_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* out; CK(cudaMalloc(&out, 16));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int mode = 0; mode < 3; ++mode)
        for (int N : Ns)
            for (int reps : {48, 480}) {
                // bar_store[1] is used uninitialised as a sink in mode 1: initialise through a first launch in mode 0 is not needed,
                // the commit only arrives on it
                rate_kernel<<<sms, 128, 99 * 1024>>>(N, reps, mode, 12, out);
                CK(cudaDeviceSynchronize());
                long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
                printf("mode %d N %3d reps %3d: issue %6lld cyc (%.1f / MMA), complete %6lld cyc (%.1f / MMA)\n", mode, N, reps, h[0], (double)h[0] / reps, h[1], (double)h[1] / reps);
            }
    return 0;
}
