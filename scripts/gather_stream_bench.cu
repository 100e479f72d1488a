// Micro-benchmark for the one-pass update + policy kernel (csrc/fused_step2.cu): can ONE persistent CTA per SM
//   * gather 128 arbitrary belief rows (123 KB apart, picked by an index list) K-block by K-block (64 floats = 256 B per row) into a
//     128B-swizzled shared-memory tile,
//   * rewrite the tile in place (thread = row, the A-operand transform of the tcgen05 GEMM does exactly this read),
//   * stream it back out to 128 consecutive rows of another buffer with TMA tensor stores at a shifted column (the N/S move),
//   * while also pulling the two FP16 alpha planes of the block (L2 resident) through shared memory,
// at the HBM bound (read 4·S + write 4·S bytes per row)?  Load variants: 0 = TMA box of consecutive rows (no gather: upper bound),
// 1 = cp.async (LDGSTS) 16 B per lane with the swizzle applied by the lane, 2 = TMA tile::gather4.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a scripts/gather_stream_bench.cu -o scripts/gather_stream_bench.bin -lcuda
// Run:   gather_stream_bench <mode> <stages> <nprod warps> <shift floats> <bload 0|1> <transform 0|1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <random>

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
    while (!mbar_try_wait(bar, parity)) { if (clock64() - t0 > 2000000000LL) { printf("watchdog block %d thread %d bar %u\n", blockIdx.x, threadIdx.x, bar); __trap(); } }
}
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

struct P {
    const float* A; long long ld; int n_tiles; int K;
    const int* rows;      // [n_tiles * 128] source row of every tile slot
    int mode, stages, nprod, shift, bload, transform, rot, noload, nostore, nissue, storemode, bmode;
    float* out;
    const uint16_t* planes; long long ldh;
    float* sink;
};

constexpr int BKF = 64;                     // floats per row per K block
constexpr int BOX_BYTES = 128 * 128;        // one 32-float box of 128 rows
constexpr int A_BYTES = 2 * BOX_BYTES;      // 32 KB
constexpr int B_ROWS = 80;
constexpr int B_BYTES = 2 * B_ROWS * 128;   // two FP16 planes of 80 x 64 halfs
constexpr int MAX_STAGES = 6;
constexpr int NTHREADS = 32 * 13;           // warps 0-3 producers, 4-7 transform, 8-11 store (8 alone issues TMA stores in storemode 0), 12 plane loader
// K block visited at position j of tile t's loop: rotated start, neighbouring tiles walk in opposite directions (rot > 0)
__device__ __forceinline__ int kb_of(int t, int j, int KB, int rot) {
    if (rot <= 0) return j;
    int k = j + (int)(((unsigned)t * (unsigned)rot) % (unsigned)KB);
    if (k >= KB) k -= KB;
    return (t & 1) ? KB - 1 - k : k;
}           // warps 0-3 producers, 4-7 transform, 8 store issuer, 9 plane loader

__global__ void __launch_bounds__(NTHREADS, 1)
gather_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmIn1, const __grid_constant__ CUtensorMap tmOut,
              const __grid_constant__ CUtensorMap tmB, const P p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int stage_bytes = A_BYTES + (p.bload ? B_BYTES : 0);
    const uint32_t bars = base + p.stages * stage_bytes;
    auto full = [&](int s) { return bars + 8 * s; };
    auto ready = [&](int s) { return bars + 8 * (MAX_STAGES + s); };
    auto empty = [&](int s) { return bars + 8 * (2 * MAX_STAGES + s); };
    auto bfull = [&](int s) { return bars + 8 * (3 * MAX_STAGES + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full(s), (p.mode == 0 || p.noload) ? 1 : p.nissue);
            mbar_init(ready(s), 4);
            mbar_init(empty(s), p.storemode ? 4 : 1);
            mbar_init(bfull(s), p.bmode ? 32 : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int KB = p.K / BKF;                 // whole blocks only (the kernel proper handles the tail with an overlapping block)

    if (warp < 4) {
        // ===================== producers =====================
        if (warp < p.nissue && !p.noload) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                const int* rl = p.rows + (long long)t * 128;
                // this warp's share of the tile's rows: 128 / nissue rows, 4 per instruction
                const int per = 128 / p.nissue;
                int r4[8][4];
#pragma unroll
                for (int q = 0; q < 8; ++q)
#pragma unroll
                    for (int e = 0; e < 4; ++e) r4[q][e] = (4 * q < per) ? rl[warp * per + 4 * q + e] : 0;
                if (elect_one_sync()) {
                    for (int kj = 0; kj < KB; ++kj, ++it) {
                        const int kb = kb_of(t, kj, KB, p.rot);
                        const int s = it % p.stages;
                        const uint32_t ph = (it / p.stages) & 1;
                        mbar_wait(empty(s), ph ^ 1);
                        const uint32_t dst = base + s * stage_bytes;
                        if (p.mode == 0) {
                            if (warp == 0) {
                                mbar_expect_tx(full(s), A_BYTES);
                                tma_load_2d(dst, &tmIn, kb * BKF, rl[0], full(s));
                                tma_load_2d(dst + BOX_BYTES, &tmIn, kb * BKF + 32, rl[0], full(s));
                            }
                        } else {
                            mbar_expect_tx(full(s), A_BYTES / p.nissue);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (4 * q < per) {
                                    const int r = warp * per + 4 * q;
                                    tma_gather4(dst + r * 128, &tmIn1, kb * BKF, r4[q][0], r4[q][1], r4[q][2], r4[q][3], full(s));
                                    tma_gather4(dst + BOX_BYTES + r * 128, &tmIn1, kb * BKF + 32, r4[q][0], r4[q][1], r4[q][2], r4[q][3], full(s));
                                }
                            }
                        }
                    }
                } else {
                    it += KB;
                }
                __syncwarp();
            }
        }
    } else if (warp < 8) {
        // ===================== transform: thread = row, in place =====================
        const int t = threadIdx.x - 128;
        const int sw = t & 7;
        uint32_t it = 0;
        float acc = 0.f;
        for (int tl = blockIdx.x; tl < p.n_tiles; tl += gridDim.x) {
            for (int kj = 0; kj < KB; ++kj, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                if (!p.noload) mbar_wait(full(s), ph); else mbar_wait(empty(s), ph ^ 1);
                if (p.bload) mbar_wait(bfull(s), ph);
                uint8_t* rowp = gbase + s * stage_bytes + t * 128;
                if (p.transform) {
#pragma unroll
                    for (int box = 0; box < 2; ++box) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float4* q = reinterpret_cast<float4*>(rowp + box * BOX_BYTES + ((c ^ sw) << 4));
                            float4 v = *q;
                            v.x *= 2.f; v.y *= 2.f; v.z *= 2.f; v.w *= 2.f;
                            acc += v.x;
                            *q = v;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(ready(s));
            }
        }
        if (acc == 123.456f) p.sink[0] = acc;
    } else if (warp < 12 && p.storemode == 1) {
        // ===================== store warps: swizzled tile -> coalesced 256-byte row segments (st.global.v4) =====================
        const int w = warp - 8;
        const int r_in = lane >> 4, c16 = lane & 15;
        const int box = c16 >> 3, piece = c16 & 7;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            for (int kj = 0; kj < KB; ++kj, ++it) {
                const int kb = kb_of(t, kj, KB, p.rot);
                const int s = it % p.stages;
                const uint32_t ph = (it / p.stages) & 1;
                mbar_wait(ready(s), ph);
                const uint8_t* src = gbase + s * stage_bytes + box * BOX_BYTES;
                float* dst = p.out + ((long long)t * 128 + w * 32) * p.ld + (long long)kb * BKF + p.shift + c16 * 4;
                const bool col_ok = kb * BKF + p.shift + c16 * 4 + 3 < p.K;
                if (!p.nostore) {
#pragma unroll 4
                    for (int r = 0; r < 32; r += 2) {
                        const int row = w * 32 + r + r_in;
                        const float4 v = *reinterpret_cast<const float4*>(src + row * 128 + ((piece ^ (row & 7)) << 4));
                        if (col_ok)
                            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                                         :: "l"(dst + (long long)(r + r_in) * p.ld), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(empty(s));
            }
        }
    } else if (warp == 8) {
        // ===================== store issuer =====================
        if (elect_one_sync()) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                for (int kj = 0; kj < KB; ++kj, ++it) {
                        const int kb = kb_of(t, kj, KB, p.rot); (void)kb;
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(ready(s), ph);
                    const uint32_t src = base + s * stage_bytes;
                    if (!p.nostore) {
                        tma_store_2d(&tmOut, src, kb * BKF + p.shift, t * 128);
                        tma_store_2d(&tmOut, src + BOX_BYTES, kb * BKF + 32 + p.shift, t * 128);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (it > 0) {
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        mbar_arrive(empty((it - 1) % p.stages));
                    }
                }
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (it > 0) mbar_arrive(empty((it - 1) % p.stages));
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 12) {
        // ===================== plane loader (L2-resident B operand) =====================
        if (p.bload && p.bmode == 0 && elect_one_sync()) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                for (int kj = 0; kj < KB; ++kj, ++it) {
                    const int kb = kb_of(t, kj, KB, p.rot);
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(empty(s), ph ^ 1);
                    mbar_expect_tx(bfull(s), B_BYTES);
                    const uint32_t dst = base + s * stage_bytes + A_BYTES;
                    tma_load_2d(dst, &tmB, kb * BKF, 0, bfull(s));
                    tma_load_2d(dst + B_ROWS * 128, &tmB, kb * BKF, B_ROWS, bfull(s));
                }
            }
        } else if (p.bload && p.bmode == 1) {
            // cp.async: 2 planes x 80 rows x 128 B = 1280 16-byte pieces, 40 per lane; the swizzle is applied by the lane
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
                for (int kj = 0; kj < KB; ++kj, ++it) {
                    const int kb = kb_of(t, kj, KB, p.rot);
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(empty(s), ph ^ 1);
                    const uint32_t dst = base + s * stage_bytes + A_BYTES;
#pragma unroll 8
                    for (int j = 0; j < 40; ++j) {
                        const int idx = j * 32 + lane;          // piece index: row = idx >> 3 (0..159 over both planes), piece = idx & 7
                        const int row = idx >> 3, piece = idx & 7;
                        cp_async16(dst + row * 128 + ((piece ^ (row & 7)) << 4), p.planes + (long long)row * p.ldh + (long long)kb * BKF + piece * 8);
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(bfull(s)) : "memory");
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int stages = argc > 2 ? atoi(argv[2]) : 4;
    const int nprod = argc > 3 ? atoi(argv[3]) : 4;
    const int shift = argc > 4 ? atoi(argv[4]) : 0;
    const int bload = argc > 5 ? atoi(argv[5]) : 0;
    const int transform = argc > 6 ? atoi(argv[6]) : 1;
    const int g4rows = 1;
    const int rot = argc > 7 ? atoi(argv[7]) : 0;
    const int ctas = argc > 8 ? atoi(argv[8]) : 1;
    const int noload = argc > 9 ? atoi(argv[9]) : 0;
    const int nostore = argc > 10 ? atoi(argv[10]) : 0;
    const int nissue = argc > 11 ? atoi(argv[11]) : 1;
    const int promo = argc > 12 ? atoi(argv[12]) : 256;
    const int storemode = argc > 13 ? atoi(argv[13]) : 0;
    const int bmode = argc > 14 ? atoi(argv[14]) : 0;
    const int Wp = argc > 15 ? atoi(argv[15]) : 372;
    const int Sp = 83 * Wp;
    const long long ld = argc > 16 ? atoll(argv[16]) : Sp;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int waves = 4;
    const int n_tiles = sms * waves;
    const long long n_rows = (long long)n_tiles * 128;
    float *A, *Bo;
    CK(cudaMalloc(&A, n_rows * ld * 4));
    CK(cudaMalloc(&Bo, n_rows * ld * 4));
    {   // recognisable contents: A[r][c] = r + c * 1e-5 (approximately), filled on the host for a few rows, the rest by memset pattern
        std::vector<float> h((size_t)ld);
        CK(cudaMemset(A, 0, n_rows * ld * 4));
        for (long long r = 0; r < n_rows; r += 997) {
            for (long long c = 0; c < ld; ++c) h[c] = (float)(r % 4096) + (float)c * 1e-4f;
            CK(cudaMemcpy(A + r * ld, h.data(), ld * 4, cudaMemcpyHostToDevice));
        }
        CK(cudaMemset(Bo, 0xff, n_rows * ld * 4));
    }
    // index list: identity for mode 0, a random permutation otherwise
    std::vector<int> rows((size_t)n_rows);
    for (long long i = 0; i < n_rows; ++i) rows[i] = (int)i;
    if (mode != 0) { std::mt19937 g(1234); std::shuffle(rows.begin(), rows.end(), g); }
    int* drows;
    CK(cudaMalloc(&drows, n_rows * 4));
    CK(cudaMemcpy(drows, rows.data(), n_rows * 4, cudaMemcpyHostToDevice));
    uint16_t* planes = nullptr;
    const long long ldh = (Sp + 63) / 64 * 64;
    CK(cudaMalloc(&planes, 2 * B_ROWS * ldh * 2));
    CK(cudaMemset(planes, 0, 2 * B_ROWS * ldh * 2));
    float* sink;
    CK(cudaMalloc(&sink, 4));

    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    auto make = [&](CUtensorMap* tm, void* ptr, CUtensorMapDataType dt, int esz, uint64_t inner, uint64_t outer, uint64_t stride_bytes, uint32_t b0, uint32_t b1) {
        cuuint64_t dims[2] = {inner, outer};
        cuuint64_t strides[1] = {stride_bytes};
        cuuint32_t box[2] = {b0, b1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(tm, dt, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
        (void)esz;
    };
    CUtensorMap tmIn, tmIn1, tmOut, tmB;
    make(&tmIn, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Sp, n_rows, ld * 4, 32, 128);
    make(&tmIn1, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Sp, n_rows, ld * 4, 32, g4rows);
    make(&tmOut, Bo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Sp, n_rows, ld * 4, 32, 128);
    make(&tmB, planes, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ldh, 2 * B_ROWS, ldh * 2, 64, B_ROWS);

    P p{A, ld, n_tiles, Sp, drows, mode, stages, nprod, shift, bload, transform, rot, noload, nostore, nissue, storemode, bmode, Bo, planes, ldh, sink};
    const int stage_bytes = A_BYTES + (bload ? B_BYTES : 0);
    const int smem = stages * stage_bytes + 1024 + 8 * 4 * MAX_STAGES + 64;
    if (smem > 227 * 1024 || stages > MAX_STAGES) { printf("smem %d too large\n", smem); return 3; }
    auto launch = [&]() { gather_kernel<<<sms * ctas, NTHREADS, smem>>>(tmIn, tmIn1, tmOut, tmB, p); };
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    launch();
    CK(cudaDeviceSynchronize());
    if (!noload && !nostore)
    // verify: output row t*128+l, column c + shift == 2^(transform) * A[rows[t*128+l]][c] for the rows that were filled
    {
        std::vector<float> h((size_t)ld);
        int bad = 0, checked = 0;
        for (long long i = 0; i < n_rows && checked < 64; ++i) {
            const long long r = rows[i];
            if (r % 997 != 0) continue;
            CK(cudaMemcpy(h.data(), Bo + i * ld, ld * 4, cudaMemcpyDeviceToHost));
            const int KBf = (Sp / BKF) * BKF;
            for (long long c = 0; c < KBf; ++c) {
                if (c + shift < 0 || c + shift >= Sp) continue;
                const float want = ((float)(r % 4096) + (float)c * 1e-4f) * (transform ? 2.f : 1.f);
                if (h[c + shift] != want) { if (bad < 5) printf("mismatch out row %lld col %lld: %g want %g\n", i, c + shift, h[c + shift], want); ++bad; }
            }
            ++checked;
        }
        printf("verify: %d rows checked, %d mismatches\n", checked, bad);
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 3;
    const double bytes = 2.0 * (double)n_rows * (Sp / BKF) * BKF * 4;
    printf("Wp %d ld %lld mode %d stages %d nprod %d shift %d bload %d transform %d rot %d ctas %d noload %d nostore %d nissue %d promo %d storemode %d bmode %d: %8.3f ms  %7.1f GB/s (read + write)\n", Wp, ld, mode, stages, nprod, shift, bload,
           transform, rot, ctas, noload, nostore, nissue, promo, storemode, bmode, ms, bytes / ms / 1e6 / ((noload || nostore) ? 2 : 1));
    return 0;
}
