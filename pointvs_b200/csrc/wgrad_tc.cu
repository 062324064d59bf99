// Node-level weight gradients on the 5th-generation tensor cores.
//
// d_w[ko][ki] += A[rows, ko]^T . B[rows, ki] for the six or seven products a
// layer's backward needs over its N node rows (node MLP, node attention and the
// factorised first edge layer: reference egnn_satorras.py:86-121, autograd of
// nn.Linear), plus the column sums of A for the bias gradients.  Same
// contraction as the edge kernel's G3 / G5 (egnn_edge_bwd_tc.cu): 128 rows of A
// and of B become bf16 hi + lo tiles with 128-byte rows (SWIZZLE_128B), read
// MN-major with the row index as K; D[64 x 64] (UMMA M = 64, fp32) accumulates in
// tensor memory over the CTA's whole row range and is written once into the
// per-CTA partial block that wgrad_group_reduce_kernel sums in a fixed order.
//
// One CTA = one (job, row range); 64 KB of tiles and 64 TMEM columns, so three
// CTAs share an SM and one CTA's loads overlap another's MMAs.
#include "egnn_bwd_common.cuh"
#include "tc_common.cuh"

namespace pvs {

namespace {

constexpr int WT = 256;
constexpr int WROWS = 128;        // rows per MMA step (8 K-steps of 16)
// M = 64, N = 64, bf16, fp32 accumulate, A and B MN-major (bits 15, 16)
constexpr uint32_t IDESC_WG = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                              ((64u >> 3) << 17) | ((64u >> 4) << 24);

struct WgTcSmem {
    uint8_t A[2][WROWS * 128];    // hi, lo
    uint8_t B[2][WROWS * 128];
    uint64_t mbar;
    uint32_t tmem_base;
};

// 8 channels [8 c, 8 c + 8) of row r of X (pitch ld, kx valid channels); zeros
// outside.  Vector loads when the pitch and base allow.
__device__ __forceinline__ void load8(const float *__restrict__ X, int ld, int kx, bool vec,
                                      int r, bool valid, int c, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    if (!valid || 8 * c >= kx) return;
    const float *src = X + (size_t)r * ld + 8 * c;
    if (vec && 8 * c + 8 <= kx) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (8 * c + i < kx) v[i] = __ldg(src + i);
    }
}

__global__ void __launch_bounds__(WT)
wgrad_group_tc_kernel(const __grid_constant__ WgradGroup G, float *__restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    WgTcSmem &S = *reinterpret_cast<WgTcSmem *>(
        (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j_id = blockIdx.x % G.n_jobs, chunk = blockIdx.x / G.n_jobs;
    const WgradJob &J = G.job[j_id];
    const bool has_b = J.B != nullptr;
    const int r_lo = chunk * G.rows_per, r_hi = min(G.rows, r_lo + G.rows_per);
    const bool vec_a = (J.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(J.A) & 15) == 0;
    const bool vec_b = has_b && (J.ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(J.B) & 15) == 0;
    const int c = tid & 7, slot = tid >> 3;       // 16-byte chunk, row slot (32 rows per pass)

    if (tid == 0) mbar_init(&S.mbar, 1);
    if (warp == 0) tmem_alloc<64>(&S.tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    pdl_wait();                  // chain kernel: see pvs_common.cuh
    pdl_launch_dependents();

    float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t phase = 0, acc = 0;
    for (int r0 = r_lo; r0 < r_hi; r0 += WROWS) {
        if (acc && has_b) {            // the previous step's MMAs still read the tiles
            mbar_wait(&S.mbar, phase);
            phase ^= 1;
            tc_fence_after();
        }
#pragma unroll
        for (int p = 0; p < WROWS / 32; ++p) {
            const int rl = slot + 32 * p, r = r0 + rl;
            float v[8];
            load8(J.A, J.lda, J.ko, vec_a, r, r < r_hi, c, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) csum[i] += v[i];
            if (has_b) {
                uint4 hi, lo;
                split8<true>(v, hi, lo);
                *reinterpret_cast<uint4 *>(S.A[0] + swz(rl, c)) = hi;
                *reinterpret_cast<uint4 *>(S.A[1] + swz(rl, c)) = lo;
                load8(J.B, J.ldb, J.ki, vec_b, r, r < r_hi, c, v);
                split8<true>(v, hi, lo);
                *reinterpret_cast<uint4 *>(S.B[0] + swz(rl, c)) = hi;
                *reinterpret_cast<uint4 *>(S.B[1] + swz(rl, c)) = lo;
            }
        }
        if (!has_b) continue;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t xh = make_desc(smem_u32(S.A[0])), xl = make_desc(smem_u32(S.A[1]));
            const uint64_t yh = make_desc(smem_u32(S.B[0])), yl = make_desc(smem_u32(S.B[1]));
            uint32_t a_ = acc;
#pragma unroll
            for (int ks = 0; ks < WROWS / 16; ++ks) {
                const uint64_t adv = (uint64_t)(ks * (2048 >> 4));
                umma_bf16(tmem, xh + adv, yh + adv, IDESC_WG, a_);
                a_ = 1;
                umma_bf16(tmem, xl + adv, yh + adv, IDESC_WG, 1);
                umma_bf16(tmem, xh + adv, yl + adv, IDESC_WG, 1);
            }
            umma_commit(&S.mbar);
        }
        acc = 1;
    }

    float *out = partial + ((size_t)j_id * G.chunks + chunk) * WG_PART;
    // column sums: 32 row slots -> one value per channel, in slot order
    // (through the A tile, free once the last MMAs are done)
    if (acc && has_b) {
        mbar_wait(&S.mbar, phase);
        tc_fence_after();
    }
    float (*red)[64] = reinterpret_cast<float (*)[64]>(S.A[0]);       // [32][64]
#pragma unroll
    for (int i = 0; i < 8; ++i) red[slot][8 * c + i] = csum[i];
    __syncthreads();
    if (tid < 64) {
        float s = 0.0f;
        for (int q = 0; q < 32; ++q) s += red[q][tid];
        out[64 * 128 + tid] = s;
    }
    // D: UMMA M = 64 puts row n on TMEM lane 32 (n / 16) + n % 16
    if (has_b && warp < 4) {
        const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float v[16];
            tmem_ld16(tl + 16 * q, v);
            if (lane < 16) {
                const int n = 16 * warp + lane;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4 *>(&out[n * 128 + 16 * q + 4 * i]) =
                        acc ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3])
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

}  // namespace

int launch_wgrad_group_tc(WgradGroup &G, int rows, int max_ctas, float *partial,
                          cudaStream_t st) {
    for (int j = 0; j < G.n_jobs; ++j)
        if (G.job[j].ko > 64 || (G.job[j].B && G.job[j].ki > 64)) return PVS_ERR_INVALID_ARG;
    int cap = max_ctas / G.n_jobs;
    if (cap < 1) cap = 1;
    const int steps = (rows + WROWS - 1) / WROWS;           // 128-row steps in all
    const int per = (steps + cap - 1) / cap;                // steps per CTA
    G.rows = rows;
    G.rows_per = per * WROWS;
    G.chunks = (rows + G.rows_per - 1) / G.rows_per;
    const size_t smem = sizeof(WgTcSmem) + 1024;
    const int rc = ensure_smem(wgrad_group_tc_kernel, smem);
    if (rc) return rc;
    launch_chained(wgrad_group_tc_kernel, dim3(G.chunks * G.n_jobs), dim3(WT), smem, st, G, partial);
    return PVS_OK;
}

}  // namespace pvs
