// Register-tiled fp32 GEMM on shared-memory operands, shared by the FFMA
// (PVS_MATH_FP32) kernels.
//
//   C[r][n] += sum_kk A[r][kk] * Wt[kk][n]
//
// 256 threads: rg = tid >> 4 (16 row groups), cg = tid & 15 (16 col groups).
// A thread owns rows rg + 16*i (i < NI) and columns 4*cg + 64*j + {0..3}
// (j < NJ4).  A is row-major with pitch lda (multiple of 4, not of 32), Wt is
// [KI][64*NJ4] (weights transposed while loading, zero padded).  Per 4 k-steps
// a thread issues NI + 4*NJ4 LDS.128 for 64*NI*NJ4/4 ... FFMA, so the loop is
// FFMA-bound; the two row groups of a warp hit different banks because
// lda % 32 == 4.
#pragma once

namespace pvs {

template <int NI, int NJ4>
__device__ __forceinline__ void tile_gemm(const float *__restrict__ A, int lda,
                                          const float *__restrict__ Wt, int KI,
                                          float (&acc)[NI][NJ4][4]) {
    constexpr int LDW = 64 * NJ4;
    const int rg = threadIdx.x >> 4, cg = threadIdx.x & 15;
    const float *a_base = A + rg * lda;
    const float *w_base = Wt + 4 * cg;
#pragma unroll 2
    for (int kk = 0; kk < KI; kk += 4) {
        float4 a[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i)
            a[i] = *reinterpret_cast<const float4 *>(a_base + 16 * i * lda + kk);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            float4 w[NJ4];
#pragma unroll
            for (int j = 0; j < NJ4; ++j)
                w[j] = *reinterpret_cast<const float4 *>(w_base + (kk + s) * LDW + 64 * j);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const float av = s == 0 ? a[i].x : s == 1 ? a[i].y : s == 2 ? a[i].z : a[i].w;
#pragma unroll
                for (int j = 0; j < NJ4; ++j) {
                    acc[i][j][0] = fmaf(av, w[j].x, acc[i][j][0]);
                    acc[i][j][1] = fmaf(av, w[j].y, acc[i][j][1]);
                    acc[i][j][2] = fmaf(av, w[j].z, acc[i][j][2]);
                    acc[i][j][3] = fmaf(av, w[j].w, acc[i][j][3]);
                }
            }
        }
    }
}

// Load W [ko][ki] (row-major, pitch ld_w: nn.Linear layout) transposed into
// Wt [KIP][LDW], zero padding rows >= ki and columns >= ko.
__device__ __forceinline__ void load_wt(float *__restrict__ Wt, int KIP, int LDW,
                                        const float *__restrict__ W, int ld_w,
                                        int ki, int ko) {
    for (int idx = threadIdx.x; idx < KIP * LDW; idx += blockDim.x) {
        int kk = idx / LDW, n = idx - kk * LDW;
        Wt[idx] = (kk < ki && n < ko) ? W[(size_t)n * ld_w + kk] : 0.0f;
    }
}

// sum over the 16 lanes that share a row group (cg = lane & 15)
__device__ __forceinline__ float rowgroup_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    return v;
}

}  // namespace pvs
