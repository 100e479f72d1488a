// tcgen05 split-precision GEMM with fused segmented max/argmax epilogue (see umma_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// out_val[i][s] = max_{g < seg_len} A[i,:] . B[s*seg_stride + g,:]   (3xTF32, FP32 accumulate)
// out_arg[i][s] = first argmax g.
//   A      [n_rows x K] fp32, row stride lda (multiple of 4 floats, 16-byte aligned base)
//   B_hi/B_lo [rows_b x ldb] fp32 planes whose values are exactly TF32-representable
//             (low 13 mantissa bits zero), ldb a multiple of 32, >= K, zero beyond K.
//   rows_b = n_seg * seg_stride, seg_stride a multiple of 16.
int olf_umma_segmax(const float* A, int64_t lda, int64_t n_rows, int K,
                    const float* B_hi, const float* B_lo, int64_t ldb, int rows_b,
                    int n_seg, int seg_len, int seg_stride,
                    float* out_val, int* out_arg, int sm_count, cudaStream_t st);

// Plain product with the same split precision: out[i*out_ld + c] = A[i,:] . B[c,:] for c < rows_b (rows_b <= 128, stored
// rows of B rounded up to 16).  Used by the infotaxis choice (belief x pull tables).
int olf_umma_store(const float* A, int64_t lda, int64_t n_rows, int K,
                   const float* B_hi, const float* B_lo, int64_t ldb, int rows_b,
                   float* out, int64_t out_ld, int sm_count, cudaStream_t st);

// 2-D tensor map over a row-major FP32 matrix, box = 32 floats x box_rows, 128B swizzle (K-major UMMA operand tiles).
int olf_make_map_2d(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_rows);

// Same contraction with a 2-way FP16 split (three kind::f16 MMAs per K = 16: half the tensor time of 3xTF32, same accuracy):
//   B1 / B2 [rows_b x ldb] FP16 planes (ldb in halfs, multiple of 64, >= K, zero beyond K): b1 = fp16(t_j v), b2 = fp16((t_j v - b1) 2^11)
//   with t_j a power of two per row; binv[j] = 1 / t_j; a_scale (may be NULL) = per-row normaliser of A (entries * a_scale <= 1).
int olf_umma_segmax_f16(const float* A, int64_t lda, int64_t n_rows, int K, const float* a_scale,
                        const void* B1, const void* B2, const float* binv, int64_t ldb, int rows_b,
                        int n_seg, int seg_len, int seg_stride, float* out_val, int* out_arg, int sm_count, cudaStream_t st);

// FP16 split planes of a row-major FP32 matrix (one power-of-two scale per row, see olf_umma_segmax_f16): nseg segments of `rows`
// source rows, each padded to rows_padded destination rows; ldh in halfs (multiple of 64, >= K).
int olf_split_planes_f16(const float* src, int64_t ld_src, int64_t rows, int K, void* b1, void* b2, float* binv, int64_t ldh,
                         int rows_padded, int nseg, cudaStream_t st);
