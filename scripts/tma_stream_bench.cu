// Micro-benchmark: how fast can ONE CTA per SM stream a [128 rows x K] fp32 tile whose rows are ~123 KB apart in HBM
// (the A operand of the evaluate / infotaxis GEMM and of a fused update+evaluate kernel)?
// Variants: TMA 2D tiles of different box widths / swizzles / stage counts, and per-row 1D bulk copies.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a scripts/tma_stream_bench.cu -o gpurun_out/tma_stream_bench -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000LL) { printf("watchdog block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); } }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct P {
    const float* A; long long ld; int n_tiles; int K;
    int mode;         // 0: TMA 2D box (bk x rows_per_box), 1: per-row 1D bulk copies of bk floats
    int bk;           // floats per row per stage
    int stages;
    int split;        // TMA mode: number of row-boxes per stage (128 / split rows each), issued by `split` different warps
    float* sink;
};

// warp 0..3: producers (lane 0 of warp w issues box w when split > w; in 1D mode all 128 threads issue one row each)
// warp 4: consumer (reads one float per row to keep the data "used", frees the stage)
__global__ void __launch_bounds__(160, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int stage_bytes = 128 * p.bk * 4;
    const uint32_t bars = base + p.stages * stage_bytes;
    auto full = [&](int s) { return bars + 8 * s; };
    auto empty = [&](int s) { return bars + 8 * (p.stages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), p.mode == 0 ? p.split : 128); mbar_init(empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int KB = (p.K + p.bk - 1) / p.bk;
    if (warp < 4) {
        uint32_t it = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                if (p.mode == 0) {
                    if (warp < p.split && lane == 0) {
                        mbar_wait(empty(s), ph ^ 1);
                        const int rows = 128 / p.split;
                        mbar_expect_tx(full(s), rows * p.bk * 4);
                        tma_load_2d(base + s * stage_bytes + warp * rows * p.bk * 4, &tm, kb * p.bk, t * 128 + warp * rows, full(s));
                    }
                } else {
                    mbar_wait(empty(s), ph ^ 1);
                    const int r = threadIdx.x;      // 0..127
                    int nb = p.bk;
                    if (kb * p.bk + nb > p.K) nb = p.K - kb * p.bk;
                    nb &= ~3;
                    mbar_expect_tx(full(s), nb * 4);
                    if (nb > 0) bulk_load_1d(base + s * stage_bytes + r * p.bk * 4, p.A + (long long)(t * 128 + r) * p.ld + (long long)kb * p.bk, nb * 4, full(s));
                }
            }
        }
    } else {
        uint32_t it = 0;
        float acc = 0.f;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                mbar_wait(full(s), ph);
                acc += *reinterpret_cast<const float*>(smem + (base - smem_u32(smem)) + s * stage_bytes + lane * 16);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty(s));
            }
        }
        if (acc == 123.456f) p.sink[0] = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int Sp = 83 * 372;                  // C2 padded states
    const long long ld = Sp;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int waves = 4;
    const int n_tiles = sms * waves;
    const long long n_rows = (long long)n_tiles * 128;
    float* A;
    CK(cudaMalloc(&A, n_rows * ld * 4));
    CK(cudaMemset(A, 0, n_rows * ld * 4));
    float* sink;
    CK(cudaMalloc(&sink, 4));
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));

    struct V { const char* name; int mode, bk, stages, split; CUtensorMapSwizzle sw; CUtensorMapL2promotion promo; };
    std::vector<V> vs = {
        {"tma2d sw128 bk32 x128rows st4", 0, 32, 4, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"tma2d sw128 bk32 x128rows st8", 0, 32, 8, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"tma2d sw128 bk32 x128rows st12", 0, 32, 12, 1, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"tma2d sw128 bk32 4x32rows st8", 0, 32, 8, 4, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"tma2d none  bk32 x128rows st8", 0, 32, 8, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"tma2d none  bk64 x128rows st4", 0, 64, 4, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"tma2d none  bk64 x128rows st6", 0, 64, 6, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"tma2d none  bk128 x128rows st3", 0, 128, 3, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"tma2d none  bk256 x128rows st1", 0, 256, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {"bulk1d bk32 st8", 1, 32, 8, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"bulk1d bk64 st6", 1, 64, 6, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"bulk1d bk128 st3", 1, 128, 3, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {"bulk1d bk372 st1 (one grid row)", 1, 372, 1, 1, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
    };
    for (const V& v : vs) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)Sp, (cuuint64_t)n_rows};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {(cuuint32_t)v.bk, (cuuint32_t)(128 / v.split)};
        cuuint32_t estr[2] = {1, 1};
        if (v.mode == 0) {
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw, v.promo,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("%-36s encode failed %d\n", v.name, (int)r); continue; }
        }
        P p{A, ld, n_tiles, Sp, v.mode, v.bk, v.stages, v.split, sink};
        const int smem = v.stages * 128 * v.bk * 4 + 1024 + 16 * v.stages + 64;
        if (smem > 220 * 1024) { printf("%-36s smem %d too large\n", v.name, smem); continue; }
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        stream_kernel<<<sms, 160, smem>>>(tm, p);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; ++i) stream_kernel<<<sms, 160, smem>>>(tm, p);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 3;
        const double bytes = (double)n_rows * Sp * 4;
        printf("%-36s %8.3f ms  %7.1f GB/s  (%.1f GB/s per SM)\n", v.name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms);
    }
    return 0;
}
