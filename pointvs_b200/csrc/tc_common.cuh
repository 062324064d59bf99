// tcgen05 / TMEM / mbarrier PTX wrappers and the swizzled bf16 hi/lo tile
// helpers shared by the tensor-core kernels (egnn_edge_tc.cu, egnn_node_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "pvs_common.cuh"

namespace pvs {

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: a descriptor bug must fail the launch, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
          "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ... and the store (the thread's own lane; completes before it returns)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
          "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
          "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 128-byte rows:
// start address [0,14) (>>4), LBO [16,30) = 1 (unused for swizzled K-major),
// SBO [32,46) = 1024 B between 8-row groups, version [46,48) = 1,
// layout type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// Instruction descriptor (kind::f16): D fp32 [4,6)=1, A bf16 [7,10)=1,
// B bf16 [10,13)=1, both K-major, N>>3 at [17,23), M>>4 at [24,29).
// f16 = true: A and B are IEEE half (format code 0) instead of bf16 (1).
__host__ __device__ constexpr uint32_t tc_idesc(uint32_t n, bool f16 = false) {   // M = 128, N = n
    return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((n >> 3) << 17) |
           ((128u >> 4) << 24);
}
constexpr uint32_t TC_IDESC = tc_idesc(64);
constexpr uint32_t TC_IDESC_F16 = tc_idesc(64, true);

// byte offset of 16-byte chunk `c` (8 bf16 = channels 8c..8c+7) of row `r`
__device__ __forceinline__ uint32_t swz(int r, int c) {
    return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
}

// split 8 fp32 into bf16 hi (round-to-nearest) and bf16 lo = bf16(x - hi)
template <bool WITH_LO>
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        h[i] = *reinterpret_cast<uint32_t *>(&hb);
        if (WITH_LO) {
            float r0 = v[2 * i] - __uint_as_float(h[i] << 16);
            float r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
            __nv_bfloat162 lb = __floats2bfloat162_rn(r0, r1);
            l[i] = *reinterpret_cast<uint32_t *>(&lb);
        } else {
            l[i] = 0u;
        }
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// W block [n_valid][k_valid] fp32 with pitch ld (nn.Linear [out][in]) ->
// swizzled bf16 hi / lo B tiles of n_rows x 64 (zero padded); whole CTA.
template <bool WITH_LO>
__device__ void load_weight_tiles(uint8_t *hi_tile, uint8_t *lo_tile,
                                  const float *__restrict__ W, int ld, int n_valid,
                                  int k_valid, int n_rows = 64) {
    for (int idx = threadIdx.x; idx < n_rows * 8; idx += blockDim.x) {
        const int n = idx >> 3, c = idx & 7;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = 8 * c + i;
            v[i] = (n < n_valid && kk < k_valid) ? W[(size_t)n * ld + kk] : 0.0f;
        }
        uint4 hi, lo;
        split8<WITH_LO>(v, hi, lo);
        *reinterpret_cast<uint4 *>(hi_tile + swz(n, c)) = hi;
        if (WITH_LO) *reinterpret_cast<uint4 *>(lo_tile + swz(n, c)) = lo;
    }
}


// W^T block -> swizzled bf16 hi / lo tile: tile row n (= input channel of W),
// K index = output channel:  tile[n][kk] = W[kk][n]
__device__ inline void load_weight_tiles_T(uint8_t *hi_tile, uint8_t *lo_tile,
                                    const float *__restrict__ W, int ld, int n_valid,
                                    int k_valid) {
    for (int idx = threadIdx.x; idx < 64 * 8; idx += blockDim.x) {
        const int n = idx >> 3, c = idx & 7;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kk = 8 * c + i;
            v[i] = (n < n_valid && kk < k_valid) ? W[(size_t)kk * ld + n] : 0.0f;
        }
        uint4 hi, lo;
        split8<true>(v, hi, lo);
        *reinterpret_cast<uint4 *>(hi_tile + swz(n, c)) = hi;
        *reinterpret_cast<uint4 *>(lo_tile + swz(n, c)) = lo;
    }
}

template <bool X3>
__device__ __forceinline__ float silu_mode(float v) {
    return X3 ? siluf_(v) : siluf_fast_(v);
}
template <bool X3>
__device__ __forceinline__ float2 silu2_mode(float2 v) {
    return X3 ? silu2_(v) : silu2_fast_(v);
}
// in-place SiLU of two packed pairs (shared reciprocal in the fp32-class mode)
template <bool X3>
__device__ __forceinline__ void silu4_mode(float2 &u, float2 &v) {
    if (X3) {
        silu4_(u, v);
    } else {
        u = silu2_fast_(u);
        v = silu2_fast_(v);
    }
}

// split 4 fp32 pairs into bf16 hi (round-to-nearest) and bf16 lo = bf16(x - hi),
// packed arithmetic: per pair 1 cvt + 2 ALU + 1 FFMA2 + 1 cvt
template <bool WITH_LO>
__device__ __forceinline__ void split8p(const float2 (&v)[4], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 hb = __floats2bfloat162_rn(v[i].x, v[i].y);
        h[i] = *reinterpret_cast<uint32_t *>(&hb);
        if (WITH_LO) {
            const float2 hf = make_float2(__uint_as_float(h[i] << 16),
                                          __uint_as_float(h[i] & 0xffff0000u));
            const float2 r = ffma2(hf, make_float2(-1.0f, -1.0f), v[i]);
            __nv_bfloat162 lb = __floats2bfloat162_rn(r.x, r.y);
            l[i] = *reinterpret_cast<uint32_t *>(&lb);
        } else {
            l[i] = 0u;
        }
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- fp16 operand helpers (PVS_MATH_FP16X2: activations as ONE fp16 tile) ----
// two fp32 -> packed half2, round to nearest, saturating at +-65504
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t w) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}
__device__ __forceinline__ uint4 pack8_f16(const float2 (&v)[4]) {
    return make_uint4(pack_f16x2(v[0].x, v[0].y), pack_f16x2(v[1].x, v[1].y),
                      pack_f16x2(v[2].x, v[2].y), pack_f16x2(v[3].x, v[3].y));
}
// W block -> fp16 hi / lo B tiles (w = hi + lo to 2^-22), same layout as
// load_weight_tiles
__device__ __forceinline__ void load_weight_tiles_f16(uint8_t *hi_tile, uint8_t *lo_tile,
                                                      const float *__restrict__ W, int ld,
                                                      int n_valid, int k_valid,
                                                      int n_rows = 64) {
    for (int idx = threadIdx.x; idx < n_rows * 8; idx += blockDim.x) {
        const int n = idx >> 3, c = idx & 7;
        float2 v[4], r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = 8 * c + 2 * i;
            v[i].x = (n < n_valid && kk < k_valid) ? W[(size_t)n * ld + kk] : 0.0f;
            v[i].y = (n < n_valid && kk + 1 < k_valid) ? W[(size_t)n * ld + kk + 1] : 0.0f;
        }
        const uint4 hi = pack8_f16(v);
        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 h = unpack_f16x2(hw[i]);
            r[i] = make_float2(v[i].x - h.x, v[i].y - h.y);
        }
        *reinterpret_cast<uint4 *>(hi_tile + swz(n, c)) = hi;
        *reinterpret_cast<uint4 *>(lo_tile + swz(n, c)) = pack8_f16(r);
    }
}

// MMAs of one 64-wide K block: D[128 x N] (+)= A[128 x 64] . B[N x 64]^T with the
// bf16 hi/lo operand tiles (X3: Ahi.Bhi + Alo.Bhi + Ahi.Blo).  `acc` = 0 makes the
// first MMA overwrite D.  Issued by ONE thread; follow with umma_commit().
template <bool X3>
__device__ __forceinline__ void issue_kblock(uint32_t tmem_d, uint32_t idesc,
                                             const uint8_t *a_hi, const uint8_t *a_lo,
                                             const uint8_t *b_hi, const uint8_t *b_lo,
                                             uint32_t acc) {
    const uint64_t ah = make_desc(smem_u32(a_hi)), al = make_desc(smem_u32(a_lo));
    const uint64_t bh = make_desc(smem_u32(b_hi)), bl = make_desc(smem_u32(b_lo));
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);   // 16 bf16 = 32 B = 2 x 16 B
        umma_bf16(tmem_d, ah + adv, bh + adv, idesc, acc);
        acc = 1;
        if (X3) {
            umma_bf16(tmem_d, al + adv, bh + adv, idesc, 1);
            umma_bf16(tmem_d, ah + adv, bl + adv, idesc, 1);
        }
    }
}

}  // namespace pvs
