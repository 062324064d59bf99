// K3 edge stage on the 5th-generation tensor cores.
//
// Backward of the per-edge part of EGNNLayer.forward (reference
// egnn_satorras.py:123-176; dataflow in SURVEY.md 9.2), same mathematics as
// egnn_edge_bwd_kernel (egnn_bwd.cu, fp32 FFMA) with its six 128x64x64 tile
// contractions moved to tcgen05.mma in the error-compensated bf16x3 arithmetic
// of the forward kernel (fp32-class: operands as bf16 hi + lo, fp32
// accumulation in tensor memory):
//
//   G1  t2  = s1 . W2^T        recompute     A = s1  tile, K-major   B = W2
//   G2  p   = m  . Wc1^T       recompute     A = m   tile            B = Wc1
//   G3  dWc1 += dp^T . m       weight grad   A = dp  tile, MN-major  B = m  tile, MN-major
//   G4  dmc = dp . Wc1         data grad     A = dp  tile, K-major   B = Wc1^T
//   G5  dW2 += dt2^T . s1      weight grad   A = dt2 tile, MN-major  B = s1 tile, MN-major
//   G6  ds1 = dt2 . W2         data grad     A = dt2 tile, K-major   B = W2^T
//
// Every activation tile is ONE shared-memory image (128 edge rows x 64 channels,
// bf16 hi and lo, 128-byte rows, SWIZZLE_128B).  Read with the edge index as M
// it is the K-major A operand of a data GEMM; read with the edge index as K it
// is the MN-major operand of a weight-gradient GEMM (M = N = 64, K = 128 edges:
// the contraction over edges the FFMA kernel did with a register-tiled loop).
// The two weight-gradient accumulators (64 x 64 fp32, UMMA M = 64) stay in
// tensor memory for the whole life of the CTA and are written out once.
//
// Elementwise work runs in the forward kernel's epilogue style: a thread owns
// one edge row and 32 of the 64 accumulator columns (two warps per TMEM lane
// quarter).  t2 and p stay in tensor memory and are re-read instead of keeping
// silu'(.) in registers or shared memory.  Column sums over edges (bias and
// 64 -> 1 head gradients) are a register transpose-reduce inside each warp
// (31 shuffles per 32-channel vector), so a thread carries four scalar
// accumulators, summed over warps in a fixed order at the end: like the FFMA
// kernel, no floating-point atomics, bitwise reproducible.
//
// One CTA of 8 warps per SM; work unit = the node-aligned tiles of
// pvs_build_tiles (every destination segment lives in one tile), as in the FFMA
// kernel.  Configurations with edge residual, softmax attention or incoming
// message gradients stay on the FFMA kernel (edge_bwd_tc_supported).
#include "egnn_bwd_common.cuh"
#include "tc_common.cuh"

namespace pvs {

namespace {

// Column split of the row-per-thread epilogues: CQ warps share a TMEM lane
// quarter, each thread owns one edge row and CW = 64 / CQ accumulator columns.
// CQ = 2: 8 warps, 221 registers.  CQ = 4 (16 warps at 128 registers, built with
// -DPVS_BWD_CQ=4) passes the same tests and is 1.7 % SLOWER on the training step:
// the per-row dot products then need a four-way exchange and every barrier
// waits for twice the warps, which costs more than the shorter per-thread
// chains save (profiles/r02_experiments/README.md).
#ifndef PVS_BWD_CQ
#define PVS_BWD_CQ 2
#endif
constexpr int CQ = PVS_BWD_CQ;
constexpr int CW = 64 / CQ;
constexpr int BT = 128 * CQ;
constexpr int WG_ISSUER = 128;    // thread that issues the weight-gradient MMAs
constexpr int KB = 64;
// TMEM columns
constexpr uint32_t C_D1 = 0, C_D2 = 64, C_D3 = 128, C_D4 = 192, C_DW2 = 256, C_DWC1 = 320;
// m = silu(t2 + b2) and silu'(t2 + b2), parked by E1 for E3 (fp32, the thread's
// own lane and columns): E3 reads them back instead of evaluating 32 SiLU pairs
// per thread again
constexpr uint32_t C_M = 384, C_G2 = 448;

struct BwdTcSmem {
    // activation tiles, bf16 hi [0] / lo [1], 1024-byte aligned
    uint8_t S1[2][TE * 128];    // s1 = silu(t1)
    uint8_t Mt[2][TE * 128];    // m  = silu(t2)
    uint8_t Xt[2][TE * 128];    // dp, then dt2
    // fp32, two half tiles of 32 channels (128-byte rows, 16-byte chunks
    // XOR-swizzled by row): silu'(t1), then dt1
    float SG[2][TE * 32];
    // weights: [n][k] K-major (G1, G2) and transposed (G4, G6), hi / lo
    uint8_t W2[2][64 * 128], Wc1[2][64 * 128], W2T[2][64 * 128], Wc1T[2][64 * 128];
    float b2[64], bc1[64], wc2[64], wa[64], wr[64];
    float T[PVS_MAX_EDGE_CLASSES][64];
    float acc_vec[5][64];                       // db2, dbc1, dwc2, dwa, dwr
    float acc_T[PVS_MAX_EDGE_CLASSES][64];
    float acc_s[2];                             // dba, dgate (unused here)
    float e_rad[TE], e_dx[TE], e_dy[TE], e_dz[TE];      // normalised diff
    float e_rx[TE], e_ry[TE], e_rz[TE], e_invn[TE];     // raw diff, 1/(sqrt r + eps)
    float e_tx[TE], e_ty[TE], e_tz[TE];                 // d trans
    float e_z[TE], e_alpha[TE], e_dza[TE], e_c[TE], e_dcraw[TE], e_dr[TE];
    float e_ddx[TE], e_ddy[TE], e_ddz[TE];
    // column-split partials of the per-row dot products.  Two arrays serve four
    // uses: z (E1) / a (E3) and c (E2) / r (E4) are separated by CTA barriers,
    // z and c are not (only an mbarrier wait lies between them).
    float p_za[CQ][TE], p_cr[CQ][TE];
    int e_rowl[TE], e_col[TE], e_attr[TE];
    int rp[TN + 1];
    float xsum[TN][3];
    float xn[TN][3], dxo[TN][3];                // x_in, d_x_out of the tile's nodes
    uint64_t mbar;                              // data-gradient GEMMs
    uint64_t mbar_wg;                           // weight-gradient GEMMs (waited late)
    uint32_t tmem_base;
};

static_assert(sizeof(BwdTcSmem) + 1024 <= 232448, "BwdTcSmem exceeds the 227 KB opt-in limit");

// instruction descriptor of the weight-gradient MMAs: M = 64, N = 64, bf16,
// fp32 accumulate, A and B MN-major (bits 15, 16)
constexpr uint32_t IDESC_WGRAD = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                 ((64u >> 3) << 17) | ((64u >> 4) << 24);

// D[128 x 64] = A . B^T, K = 64, both K-major, bf16x3
__device__ __forceinline__ void mma_data(uint32_t d, const uint8_t (*A)[TE * 128],
                                         const uint8_t (*B)[64 * 128]) {
    const uint64_t ah = make_desc(smem_u32(A[0])), al = make_desc(smem_u32(A[1]));
    const uint64_t bh = make_desc(smem_u32(B[0])), bl = make_desc(smem_u32(B[1]));
    uint32_t acc = 0;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        umma_bf16(d, ah + adv, bh + adv, TC_IDESC, acc);
        acc = 1;
        umma_bf16(d, al + adv, bh + adv, TC_IDESC, 1);
        umma_bf16(d, ah + adv, bl + adv, TC_IDESC, 1);
    }
}
// D[64 x 64] (+)= X^T . Y over the 128 edge rows: both tiles MN-major (the
// 64 channels of a 128-byte row are the M / N index, rows are K; one K step =
// 16 rows = 2048 bytes), bf16x3
__device__ __forceinline__ void mma_wgrad(uint32_t d, const uint8_t (*X)[TE * 128],
                                          const uint8_t (*Y)[TE * 128], uint32_t acc) {
    const uint64_t xh = make_desc(smem_u32(X[0])), xl = make_desc(smem_u32(X[1]));
    const uint64_t yh = make_desc(smem_u32(Y[0])), yl = make_desc(smem_u32(Y[1]));
#pragma unroll
    for (int ks = 0; ks < TE / 16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * (2048 >> 4));
        umma_bf16(d, xh + adv, yh + adv, IDESC_WGRAD, acc);
        acc = 1;
        umma_bf16(d, xl + adv, yh + adv, IDESC_WGRAD, 1);
        umma_bf16(d, xh + adv, yl + adv, IDESC_WGRAD, 1);
    }
}

// sigma(t), silu(t), silu'(t) with the raw MUFU forms of the forward kernel
__device__ __forceinline__ void silu_pair(float t, float &s, float &g) {
    const float sg = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * t));
    s = t * sg;
    g = sg * (1.0f + t * (1.0f - sg));
}

// Column sums over the 32 rows of a warp of N (= 32 or 16) values per row: lane
// l returns sum_rows v_row[l % N].  Recursive halving over the low lane bits,
// then (N = 16) one exchange across lane bit 4; the order of the additions is
// fixed.
template <int N>
__device__ __forceinline__ float warp_colsum(float (&v)[N], int lane) {
#pragma unroll
    for (int o = N / 2; o >= 1; o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    float r = v[0];
#pragma unroll
    for (int o = N; o < 32; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

// byte offset of 16-byte chunk `ch` (4 floats) of row r in an SG half tile
// (swizzled by the GLOBAL edge index, rot = first edge of the chunk & 7, so that
// the tile is the image of its rows in the DT1 planes: dt1_at)
__device__ __forceinline__ uint32_t sg_off(int r, int ch, int rot) {
    return (uint32_t)(r * 128 + ((ch ^ ((r + rot) & 7)) << 4));
}
constexpr int DT1_ISSUER = 64;    // thread that issues the bulk stores of the dt1 tile
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() {      // sources may be overwritten
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() {       // writes complete
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

#ifdef PVS_PHASE_PROF
// Debug build only (scripts/phase_prof.py --bwd): cycles of thread 0 between
// the phase boundaries of a tile, summed over all CTAs and tiles.
__device__ unsigned long long g_bwd_phase_cycles[16];
#define BPH(i)                                                                \
    do {                                                                      \
        if (tid == 0) {                                                       \
            const long long now_ = clock64();                                 \
            atomicAdd(&g_bwd_phase_cycles[i], (unsigned long long)(now_ - t_prev_)); \
            t_prev_ = now_;                                                   \
        }                                                                     \
    } while (0)
#else
#define BPH(i) do { } while (0)
#endif

__global__ void __launch_bounds__(BT, 1)
egnn_edge_bwd_tc_kernel(const EdgeBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    BwdTcSmem &S = *reinterpret_cast<BwdTcSmem *>(smem_dyn);
    if ((smem_u32(smem_dyn) & 1023u) != 0u) __trap();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int erow = 32 * (warp & 3) + lane;      // edge row of this thread (epilogues)
    const int cq = warp >> 2;                     // column block: columns [CW cq, CW cq + CW)
    const int k = a.k;
    const bool f_att = a.flags & PVS_F_EDGE_ATTENTION;
    const bool f_coords = a.flags & PVS_F_UPDATE_COORDS;
    const bool f_norm = a.flags & PVS_F_NORMALIZE;

    load_weight_tiles<true>(S.W2[0], S.W2[1], a.edge_w2, k, k, k);
    load_weight_tiles<true>(S.Wc1[0], S.Wc1[1], a.coord_w1, k, k, k);
    load_weight_tiles_T(S.W2T[0], S.W2T[1], a.edge_w2, k, k, k);
    load_weight_tiles_T(S.Wc1T[0], S.Wc1T[1], a.coord_w1, k, k, k);
    const int col_r = (a.flags & PVS_F_PERM_INVARIANT) ? k : 2 * k;
    for (int n = tid; n < 64; n += BT) {
        const bool ok = n < k;
        S.b2[n] = ok ? a.edge_b2[n] : 0.0f;
        S.bc1[n] = ok ? a.coord_b1[n] : 0.0f;
        S.wc2[n] = ok ? a.coord_w2[n] : 0.0f;
        S.wa[n] = (ok && a.att_w) ? a.att_w[n] : 0.0f;
        S.wr[n] = ok ? a.edge_w1[(size_t)n * a.in_e + col_r] : 0.0f;
        for (int c = 0; c < PVS_MAX_EDGE_CLASSES; ++c) {
            S.T[c][n] = (ok && c < a.n_classes)
                            ? a.edge_w1[(size_t)n * a.in_e + col_r + 1 + c] : 0.0f;
            S.acc_T[c][n] = 0.0f;
        }
        for (int v = 0; v < 5; ++v) S.acc_vec[v][n] = 0.0f;
    }
    if (tid < 2) S.acc_s[tid] = 0.0f;
    if (tid == 0) { mbar_init(&S.mbar, 1); mbar_init(&S.mbar_wg, 1); }
    if (tid < 32) tmem_alloc<512>(&S.tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(CW * cq);
    uint32_t phase = 0, phase_wg = 0;
    bool wg_pending = false;               // a weight-gradient GEMM still reads S1 / X / M
    const float att_b = (f_att && a.att_b) ? a.att_b[0] : 0.0f;
    // per-thread accumulators that live across tiles: column CW cq + lane % CW
    float gb2 = 0.f, gbc1 = 0.f, gwc2 = 0.f, gwa = 0.f, gba = 0.f;
    uint32_t acc_w2 = 0, acc_wc1 = 0;      // 0 until the first weight-gradient MMA
    // d w_r and d T[class] of channels 4 (lane & 15) .. + 3 over the nodes this
    // half warp sums in S6c
    constexpr int FAST_CLASSES = 4;
    const bool fast_classes = a.n_classes <= FAST_CLASSES;
    float4 awr = make_float4(0.f, 0.f, 0.f, 0.f), aT[FAST_CLASSES];
#pragma unroll
    for (int c = 0; c < FAST_CLASSES; ++c) aT[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n_tiles = *a.n_tiles;

    auto ld32 = [&](uint32_t col, float (&v)[CW]) {      // this thread's CW columns
#pragma unroll
        for (int b = 0; b < CW / 16; ++b) {
            float t16[16];
            tmem_ld16(tmem_lane + col + 16 * b, t16);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 * b + i] = t16[i];
        }
    };
    auto st32 = [&](uint32_t col, const float (&v)[CW]) {
#pragma unroll
        for (int b = 0; b < CW / 16; ++b) {
            float t16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t16[i] = v[16 * b + i];
            tmem_st16(tmem_lane + col + 16 * b, t16);
        }
    };
    // CW channels of this thread's row -> bf16 hi / lo tile (16-byte chunks
    // (CW / 8) cq ...)
    auto store_row = [&](uint8_t (*tile)[TE * 128], const float (&v)[CW]) {
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
            float2 p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = make_float2(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
            uint4 hi, lo;
            split8p<true>(p, hi, lo);
            const uint32_t off = swz(erow, (CW / 8) * cq + j);
            *reinterpret_cast<uint4 *>(tile[0] + off) = hi;
            *reinterpret_cast<uint4 *>(tile[1] + off) = lo;
        }
    };
    auto commit_wait = [&]() {
        if (tid == 0) umma_commit(&S.mbar);
        mbar_wait(&S.mbar, phase);
        phase ^= 1;
        tc_fence_after();
    };
    // The weight-gradient GEMMs are committed to their own barrier, AFTER the
    // data GEMM of the same step: the epilogue that needs the data GEMM starts
    // as soon as that one is done, and the tiles the weight-gradient GEMM reads
    // are only waited for right before they are overwritten.
    auto wait_wgrad = [&]() {
        if (wg_pending) {
            mbar_wait(&S.mbar_wg, phase_wg);
            phase_wg ^= 1;
            tc_fence_after();
            wg_pending = false;
        }
    };
    // make this thread's shared-memory writes visible to the tensor core, then barrier
    auto publish = [&]() {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    };

#ifdef PVS_PHASE_PROF
    long long t_prev_ = clock64();
#endif
    // The next tile's node window, first 128 edges and their coordinates are
    // requested one tile ahead, into registers, in four steps placed where the
    // current tile waits on the tensor core anyway: each step needs the result
    // of the one before (tile_ptr -> row_ptr -> col -> x), and taken at the top
    // of the tile those four dependent global round trips were 16 % of the
    // kernel.  (TN + 1 <= BT: one row_ptr entry per thread.)
    static_assert(TN + 1 <= BT && TE <= BT, "one prefetched entry per thread");
    int nx_n0 = 0, nx_nn = -1, nx_e0 = 0, nx_e1 = 0, nx_rp = 0, nx_col = 0, nx_attr = 0;
    float nx_xj[3] = {0.f, 0.f, 0.f}, nx_xi[3] = {0.f, 0.f, 0.f}, nx_dxo[3] = {0.f, 0.f, 0.f};
    auto pf_tile = [&](int tn) {
        nx_nn = -1;
        if (tn < n_tiles) {
            nx_n0 = a.tile_ptr[tn];
            nx_nn = a.tile_ptr[tn + 1] - nx_n0;
        }
    };
    auto pf_rows = [&]() {
        if (nx_nn < 0) return;
        nx_e0 = a.row_ptr[nx_n0];
        nx_e1 = a.row_ptr[nx_n0 + nx_nn];
        if (tid <= nx_nn) nx_rp = a.row_ptr[nx_n0 + tid];
        if (tid < nx_nn) {
            const int i = nx_n0 + tid;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                nx_xi[c] = a.x_in[3 * i + c];
                nx_dxo[c] = (f_coords && a.d_x_out) ? a.d_x_out[3 * i + c] : 0.0f;
            }
        }
    };
    auto pf_edges = [&]() {
        if (nx_nn < 0 || tid >= TE || nx_e0 + tid >= nx_e1) return;
        nx_col = a.col[nx_e0 + tid];
        nx_attr = a.attr ? a.attr[nx_e0 + tid] : 0;
    };
    auto pf_coords = [&]() {
        if (nx_nn < 0 || tid >= TE || nx_e0 + tid >= nx_e1) return;
#pragma unroll
        for (int c = 0; c < 3; ++c) nx_xj[c] = a.x_in[3 * nx_col + c];
    };
    // Everything above reads only what is static over the backward pass
    // (weights, graph, forward activations); d_x_out and dM come from kernels
    // earlier in the chain.
    pdl_wait();
    pdl_launch_dependents();
    pf_tile(blockIdx.x);
    pf_rows();
    pf_edges();
    pf_coords();
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int n0 = nx_n0, nn = nx_nn;
        const int my_col = nx_col, my_attr = nx_attr;
        const float my_xj[3] = {nx_xj[0], nx_xj[1], nx_xj[2]};
        __syncthreads();
        if (tid <= nn) S.rp[tid] = nx_rp;
        if (tid < nn) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                S.xn[tid][c] = nx_xi[c];
                S.dxo[tid][c] = nx_dxo[c];
                S.xsum[tid][c] = 0.0f;
            }
        }
        __syncthreads();
        pf_tile(t + gridDim.x);
        const int e0 = S.rp[0], e1 = S.rp[nn];
        const int n_chunks = max(1, (e1 - e0 + TE - 1) / TE);
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int c0 = e0 + ch * TE;
            const int ne = min(TE, e1 - c0);
            const int rot = c0 & 7;
            if (ne > 0) {
            // ---- S0: geometry and d(trans) ----
            BPH(0);
            if (tid == DT1_ISSUER) bulk_wait_read();   // previous dt1 tile has left SG
            if (tid < TE) {
                if (tid < ne) {
                    const int e = c0 + tid;
                    int lo = 0, hi = nn;
                    while (hi - lo > 1) {
                        int mid = (lo + hi) >> 1;
                        if (S.rp[mid] <= e) lo = mid; else hi = mid;
                    }
                    const int j = ch == 0 ? my_col : a.col[e];
                    float xj[3];
                    if (ch == 0) { xj[0] = my_xj[0]; xj[1] = my_xj[1]; xj[2] = my_xj[2]; }
                    else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) xj[c] = a.x_in[3 * j + c];
                    }
                    const float rx = S.xn[lo][0] - xj[0];
                    const float ry = S.xn[lo][1] - xj[1];
                    const float rz = S.xn[lo][2] - xj[2];
                    const float r = rx * rx + ry * ry + rz * rz;
                    const float invn = f_norm ? 1.0f / (sqrtf(r) + 1e-8f) : 1.0f;
                    S.e_rowl[tid] = lo; S.e_col[tid] = j;
                    S.e_attr[tid] = ch == 0 ? my_attr : (a.attr ? a.attr[e] : 0);
                    S.e_rad[tid] = r;
                    S.e_rx[tid] = rx; S.e_ry[tid] = ry; S.e_rz[tid] = rz;
                    S.e_invn[tid] = invn;
                    S.e_dx[tid] = rx * invn; S.e_dy[tid] = ry * invn; S.e_dz[tid] = rz * invn;
                    float tx = 0.f, ty = 0.f, tz = 0.f;
                    if (f_coords && a.d_x_out) {
                        const int cnt = S.rp[lo + 1] - S.rp[lo];
                        const float ic = 1.0f / (float)(cnt > 0 ? cnt : 1);
                        tx = S.dxo[lo][0] * ic;
                        ty = S.dxo[lo][1] * ic;
                        tz = S.dxo[lo][2] * ic;
                    }
                    S.e_tx[tid] = tx; S.e_ty[tid] = ty; S.e_tz[tid] = tz;
                } else {
                    // rows past the chunk end repeat its last edge in the gather
                    // (finite values) and carry zero gradients
                    S.e_rowl[tid] = 0; S.e_col[tid] = 0; S.e_attr[tid] = 0;
                    S.e_rad[tid] = 0.f;
                    S.e_rx[tid] = S.e_ry[tid] = S.e_rz[tid] = 0.f; S.e_invn[tid] = 0.f;
                    S.e_dx[tid] = S.e_dy[tid] = S.e_dz[tid] = 0.f;
                    S.e_tx[tid] = S.e_ty[tid] = S.e_tz[tid] = 0.f;
                    S.e_z[tid] = 0.f; S.e_alpha[tid] = 0.f; S.e_dza[tid] = 0.f;
                    S.e_c[tid] = 0.f; S.e_dcraw[tid] = 0.f; S.e_dr[tid] = 0.f;
                }
            }
            __syncthreads();
            // ---- S1: t1 -> s1 (S1 tile), silu'(t1) (SG); 8 lanes per edge row ----
            BPH(1);
            wait_wgrad();                  // dW2 of the previous tile read S1 and X
            {
                const int c = tid & 7, slot = tid >> 3;
                constexpr int SLOTS = BT / 8, PASSES = TE / SLOTS;   // rows per pass, passes
                float4 buf[PASSES][4];     // all passes in flight
                auto issue = [&](int p, float4 (&bq)[4]) {
                    const int r = min(p * SLOTS + slot, ne - 1);
                    const float4 *pp = reinterpret_cast<const float4 *>(
                        a.P + (size_t)(n0 + S.e_rowl[r]) * KB + 8 * c);
                    const float4 *qq = reinterpret_cast<const float4 *>(
                        a.Q + (size_t)S.e_col[r] * KB + 8 * c);
                    bq[0] = __ldg(pp); bq[1] = __ldg(pp + 1);
                    bq[2] = __ldg(qq); bq[3] = __ldg(qq + 1);
                };
#pragma unroll
                for (int p = 0; p < PASSES; ++p) issue(p, buf[p]);
#pragma unroll
                for (int p = 0; p < PASSES; ++p) {
                    const float4 (&bq)[4] = buf[p];
                    const int r = p * SLOTS + slot, re = min(r, ne - 1);
                    const float rad = S.e_rad[re];
                    const int at = S.e_attr[re];
                    const float pv[8] = {bq[0].x, bq[0].y, bq[0].z, bq[0].w,
                                         bq[1].x, bq[1].y, bq[1].z, bq[1].w};
                    const float qv[8] = {bq[2].x, bq[2].y, bq[2].z, bq[2].w,
                                         bq[3].x, bq[3].y, bq[3].z, bq[3].w};
                    float s1v[8], sgv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int n = 8 * c + i;
                        const float t1 = fmaf(S.wr[n], rad, pv[i] + qv[i]) + S.T[at][n];
                        silu_pair(t1, s1v[i], sgv[i]);
                    }
                    uint4 hi, lo;
                    split8<true>(s1v, hi, lo);
                    *reinterpret_cast<uint4 *>(S.S1[0] + swz(r, c)) = hi;
                    *reinterpret_cast<uint4 *>(S.S1[1] + swz(r, c)) = lo;
                    uint8_t *sg = reinterpret_cast<uint8_t *>(S.SG[c >> 2]);
                    *reinterpret_cast<float4 *>(sg + sg_off(r, (2 * c) & 7, rot)) =
                        make_float4(sgv[0], sgv[1], sgv[2], sgv[3]);
                    *reinterpret_cast<float4 *>(sg + sg_off(r, (2 * c + 1) & 7, rot)) =
                        make_float4(sgv[4], sgv[5], sgv[6], sgv[7]);
                }
            }
            publish();
            // ---- G1: t2 = s1 . W2^T -> D1 (kept until dt2 is formed) ----
            BPH(2);
            if (tid == 0) {
                tc_fence_after();
                mma_data(tmem_base + C_D1, S.S1, S.W2);
            }
            if (ch == 0) pf_rows();
            commit_wait();
            // ---- E1: m = silu(t2 + b2) -> M tile; attention logit partial ----
            BPH(3);
            {
                float v[CW], g2[CW];
                ld32(C_D1, v);
                float dot = 0.0f;
#pragma unroll
                for (int i = 0; i < CW; ++i) {
                    const int n = CW * cq + i;
                    float s;
                    silu_pair(v[i] + S.b2[n], s, g2[i]);
                    v[i] = s;
                    dot = fmaf(S.wa[n], s, dot);
                }
                store_row(S.Mt, v);
                st32(C_M, v);
                st32(C_G2, g2);
                S.p_za[cq][erow] = dot;
            }
            publish();
            // ---- G2: p = m . Wc1^T -> D2 ----
            BPH(4);
            if (f_coords && tid == 0) {
                tc_fence_after();
                mma_data(tmem_base + C_D2, S.Mt, S.Wc1);
            }
            if (tid < TE) {
                float z = 0.0f, al = 1.0f;
                if (f_att) {
                    z = att_b;
#pragma unroll
                    for (int b = 0; b < CQ; ++b) z += S.p_za[b][tid];
                    al = apply_act(z, a.att_act);
                }
                S.e_z[tid] = z;
                S.e_alpha[tid] = al;
            }
            // dM of this thread's row (needed in E3; requested before the GEMM wait)
            float dMv[CW];
            auto load_dM = [&]() {
                const float4 *src = reinterpret_cast<const float4 *>(
                    a.dM + (size_t)(n0 + S.e_rowl[erow]) * KB + CW * cq);
#pragma unroll
                for (int i = 0; i < CW / 4; ++i) {
                    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (erow < ne) d4 = __ldg(src + i);
                    dMv[4 * i] = d4.x; dMv[4 * i + 1] = d4.y;
                    dMv[4 * i + 2] = d4.z; dMv[4 * i + 3] = d4.w;
                }
            };
            if (ch == 0) pf_edges();
            if (f_coords) {
                commit_wait();
                // ---- E2: q = silu(p + bc1), craw = wc2 . q; then dp -> X tile ----
            BPH(5);
                float q[CW], sgp[CW];
                {
                    float v[CW];
                    ld32(C_D2, v);
                    float dot = 0.0f;
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        const int n = CW * cq + i;
                        silu_pair(v[i] + S.bc1[n], q[i], sgp[i]);
                        dot = fmaf(S.wc2[n], q[i], dot);
                    }
                    S.p_cr[cq][erow] = dot;
                }
                tc_fence_before();
                __syncthreads();
                BPH(6);
                if (tid < TE) {
                    float c = 0.0f, dcraw = 0.0f;
                    if (tid < ne) {
                        float craw = 0.0f;
#pragma unroll
                        for (int b = 0; b < CQ; ++b) craw += S.p_cr[b][tid];
                        c = (a.flags & PVS_F_TANH) ? tanhf(craw) : craw;
                        const float dc = S.e_tx[tid] * S.e_dx[tid] + S.e_ty[tid] * S.e_dy[tid] +
                                         S.e_tz[tid] * S.e_dz[tid];
                        dcraw = (a.flags & PVS_F_TANH) ? dc * (1.0f - c * c) : dc;
                    }
                    S.e_c[tid] = c;
                    S.e_dcraw[tid] = dcraw;
                }
                __syncthreads();
                {
                    const float dcr = S.e_dcraw[erow];
                    float dp[CW];
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        dp[i] = dcr * S.wc2[CW * cq + i] * sgp[i];
                        q[i] *= dcr;                   // d wc2 contribution
                    }
                    store_row(S.Xt, dp);
                    gbc1 += warp_colsum<CW>(dp, lane);
                    gwc2 += warp_colsum<CW>(q, lane);
                }
                publish();
                // ---- G3: dWc1 += dp^T . m ; G4: dmc = dp . Wc1 -> D3 ----
            BPH(7);
                if (tid == 0) {
                    tc_fence_after();
                    mma_data(tmem_base + C_D3, S.Xt, S.Wc1T);
                    umma_commit(&S.mbar);
                } else if (tid == WG_ISSUER) {
                    // 24 more instructions: issued from another warp, so that
                    // neither issuer is late for its own epilogue rows
                    tc_fence_after();
                    mma_wgrad(tmem_base + C_DWC1, S.Xt, S.Mt, acc_wc1);
                    umma_commit(&S.mbar_wg);
                }
                acc_wc1 = 1;
                wg_pending = true;
                load_dM();                 // in flight during the GEMM
                mbar_wait(&S.mbar, phase);
                phase ^= 1;
                tc_fence_after();
                BPH(8);
            } else {
                load_dM();
            }
            // ---- E3: attention backward, total dm, dt2 -> X tile ----
            {
                float m[CW], sg2[CW];
                {
                    ld32(C_M, m);
                    ld32(C_G2, sg2);
                    float dot = 0.0f;
#pragma unroll
                    for (int i = 0; i < CW; ++i) dot = fmaf(dMv[i], m[i], dot);
                    S.p_za[cq][erow] = dot;
                }
                __syncthreads();
                if (f_att && tid < TE) {
                    float dza = 0.0f;
                    if (tid < ne) {
                        float dot = 0.0f;
#pragma unroll
                        for (int b = 0; b < CQ; ++b) dot += S.p_za[b][tid];
                        dza = dot * act_grad(S.e_z[tid], S.e_alpha[tid], a.att_act);
                        gba += dza;
                    }
                    S.e_dza[tid] = dza;
                }
                __syncthreads();
                const float al = S.e_alpha[erow];
                const float dz = f_att ? S.e_dza[erow] : 0.0f;
                float dt2[CW];
                if (f_coords) ld32(C_D3, dt2);       // dm of the coordinate branch
                else {
#pragma unroll
                    for (int i = 0; i < CW; ++i) dt2[i] = 0.0f;
                }
#pragma unroll
                for (int i = 0; i < CW; ++i) {
                    const int n = CW * cq + i;
                    const float dm = dt2[i] + al * dMv[i] + dz * S.wa[n];
                    dt2[i] = (erow < ne && n < k) ? dm * sg2[i] : 0.0f;
                    m[i] *= dz;                        // d wa contribution
                }
                wait_wgrad();              // dWc1 read the X tile (dp) and the M tile
                store_row(S.Xt, dt2);
                gb2 += warp_colsum<CW>(dt2, lane);
                if (f_att) gwa += warp_colsum<CW>(m, lane);
            }
            publish();
            // ---- G5: dW2 += dt2^T . s1 ; G6: ds1 = dt2 . W2 -> D4 ----
            BPH(9);
            if (tid == 0) {
                tc_fence_after();
                mma_data(tmem_base + C_D4, S.Xt, S.W2T);
                umma_commit(&S.mbar);
            } else if (tid == WG_ISSUER) {
                tc_fence_after();
                mma_wgrad(tmem_base + C_DW2, S.Xt, S.S1, acc_w2);
                umma_commit(&S.mbar_wg);
            }
            acc_w2 = 1;
            wg_pending = true;
            if (ch == 0) pf_coords();
            mbar_wait(&S.mbar, phase);
            phase ^= 1;
            tc_fence_after();
            BPH(10);
            // ---- E4: dt1 = ds1 * silu'(t1) -> SG (in place), DT1; d radial partial ----
            {
                float v[CW];
                ld32(C_D4, v);
                // columns CW cq ..: half tile (CW cq) / 32, chunks ((CW cq) % 32) / 4 ...
                uint8_t *sg = reinterpret_cast<uint8_t *>(S.SG[(CW * cq) >> 5]);
                const int ch0 = ((CW * cq) & 31) >> 2;
                float dot = 0.0f;
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                    float4 *cell = reinterpret_cast<float4 *>(sg + sg_off(erow, ch0 + j, rot));
                    const float4 g4 = *cell;
                    const float4 d4 = make_float4(v[4 * j] * g4.x, v[4 * j + 1] * g4.y,
                                                  v[4 * j + 2] * g4.z, v[4 * j + 3] * g4.w);
                    *cell = d4;
                    const float4 w4 = *reinterpret_cast<const float4 *>(&S.wr[CW * cq + 4 * j]);
                    dot += w4.x * d4.x + w4.y * d4.y + w4.z * d4.z + w4.w * d4.w;
                }
                S.p_cr[cq][erow] = dot;
                tc_fence_before();
                fence_proxy_async();       // the dt1 tile is read by the bulk copies below
            }
            __syncthreads();
            // the dt1 tile -> its rows of the two DT1 planes (dt1_at): ne x 128 bytes
            // each, contiguous in HBM and in shared memory
            if (tid == DT1_ISSUER) {
                bulk_store(a.DT1 + (size_t)c0 * 32, S.SG[0], (uint32_t)ne * 128u);
                bulk_store(a.DT1 + (size_t)a.dt1_rows * 32 + (size_t)c0 * 32, S.SG[1],
                           (uint32_t)ne * 128u);
                bulk_commit();
            }
            // ---- S6a: per-edge d(diff): dd = 2 d dr + d(d_hat) / norm ----
            BPH(11);
            if (tid < TE) {
                float ddx = 0.f, ddy = 0.f, ddz = 0.f;
                if (tid < ne) {
                    const float c = f_coords ? S.e_c[tid] : 0.0f;
                    float drs = 0.0f;
#pragma unroll
                    for (int b = 0; b < CQ; ++b) drs += S.p_cr[b][tid];
                    const float dr2 = 2.0f * drs;
                    const float invn = S.e_invn[tid];   // norm is detached (:184)
                    ddx = fmaf(dr2, S.e_rx[tid], S.e_tx[tid] * c * invn);
                    ddy = fmaf(dr2, S.e_ry[tid], S.e_ty[tid] * c * invn);
                    ddz = fmaf(dr2, S.e_rz[tid], S.e_tz[tid] * c * invn);
                    float *dst = a.DD + (size_t)(c0 + tid) * 3;
                    dst[0] = ddx; dst[1] = ddy; dst[2] = ddz;
                }
                S.e_ddx[tid] = ddx; S.e_ddy[tid] = ddy; S.e_ddz[tid] = ddz;
            }
            // ---- S6b: d w_r and d T[class] from the dt1 tile.  Up to FAST_CLASSES
            // edge classes these sums ride along with S6c below (per-lane register
            // accumulators over all tiles, reduced once at the end); with more
            // classes one owner thread per channel walks the tile.  Both are
            // deterministic (fixed order, no atomics). ----
            if (!fast_classes && tid >= TE && tid < TE + 64) {
                const int n = tid - TE;
                const uint8_t *sg = reinterpret_cast<const uint8_t *>(S.SG[n >> 5]);
                const int chn = (n & 31) >> 2, sub = n & 3;
                float swr = 0.0f, sT[PVS_MAX_EDGE_CLASSES] = {};
                for (int el = 0; el < ne; ++el) {
                    const float d = reinterpret_cast<const float *>(sg + sg_off(el, chn, rot))[sub];
                    swr = fmaf(d, S.e_rad[el], swr);
                    const int at = S.e_attr[el];
#pragma unroll
                    for (int c = 0; c < PVS_MAX_EDGE_CLASSES; ++c) sT[c] += (at == c) ? d : 0.0f;
                }
                S.acc_vec[4][n] += swr;
                for (int c = 0; c < a.n_classes; ++c) S.acc_T[c][n] += sT[c];
            }
            // ---- S6c: dP_i = sum over the node's edges of dt1; half a warp per
            // node (16 lanes x 4 channels), 16 nodes at a time ----
            for (int nl = 2 * warp + (lane >> 4); nl < nn; nl += BT / 16) {
                const int lo = max(S.rp[nl], c0) - c0;
                const int hi = min(S.rp[nl + 1], c0 + TE) - c0;
                const int l16 = lane & 15;
                const uint8_t *sg = reinterpret_cast<const uint8_t *>(S.SG[l16 >> 3]);
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fast_classes) {
                    for (int el = lo; el < hi; ++el) {
                        const float4 d = *reinterpret_cast<const float4 *>(sg + sg_off(el, l16 & 7, rot));
                        sum.x += d.x; sum.y += d.y; sum.z += d.z; sum.w += d.w;
                        const float rad = S.e_rad[el];
                        awr.x = fmaf(d.x, rad, awr.x); awr.y = fmaf(d.y, rad, awr.y);
                        awr.z = fmaf(d.z, rad, awr.z); awr.w = fmaf(d.w, rad, awr.w);
                        const int at = S.e_attr[el];     // uniform over the half warp
#pragma unroll
                        for (int c = 0; c < FAST_CLASSES; ++c) {
                            const float w = (at == c || (c == FAST_CLASSES - 1 && at > c)) ? 1.0f : 0.0f;
                            aT[c].x = fmaf(d.x, w, aT[c].x); aT[c].y = fmaf(d.y, w, aT[c].y);
                            aT[c].z = fmaf(d.z, w, aT[c].z); aT[c].w = fmaf(d.w, w, aT[c].w);
                        }
                    }
                } else {
                    for (int el = lo; el < hi; ++el) {
                        const float4 d = *reinterpret_cast<const float4 *>(sg + sg_off(el, l16 & 7, rot));
                        sum.x += d.x; sum.y += d.y; sum.z += d.z; sum.w += d.w;
                    }
                }
                float4 *dst = reinterpret_cast<float4 *>(a.dP + (size_t)(n0 + nl) * KB + 4 * l16);
                if (ch == 0) *dst = sum;
                else if (hi > lo) {
                    const float4 o = *dst;
                    *dst = make_float4(o.x + sum.x, o.y + sum.y, o.z + sum.z, o.w + sum.w);
                }
            }
            __syncthreads();
            // ---- S6d: row part of dx ----
            BPH(12);
            if (tid < nn) {
                const int lo = max(S.rp[tid], c0) - c0;
                const int hi = min(S.rp[tid + 1], c0 + TE) - c0;
                float sx = 0.f, sy = 0.f, sz = 0.f;
                for (int el = lo; el < hi; ++el) {
                    sx += S.e_ddx[el]; sy += S.e_ddy[el]; sz += S.e_ddz[el];
                }
                S.xsum[tid][0] += sx; S.xsum[tid][1] += sy; S.xsum[tid][2] += sz;
            }
            } else {
                // edgeless tile: dP = 0 (and the whole look-ahead in one go)
                pf_rows();
                pf_edges();
                pf_coords();
                for (int nl = warp; nl < nn; nl += BT / 32)
                    *reinterpret_cast<float2 *>(a.dP + (size_t)(n0 + nl) * KB + 2 * lane) =
                        make_float2(0.f, 0.f);
            }
            __syncthreads();
            BPH(13);
        }
        if (tid < nn) {
            const int i = n0 + tid;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                a.d_x_in[3 * i + c] = (a.d_x_out ? a.d_x_out[3 * i + c] : 0.0f) + S.xsum[tid][c];
            // (d_x_out itself, not S.dxo: that copy is zero when coordinates are frozen)
        }
    }
    // ---- per-CTA partials (fixed-order reductions: bitwise reproducible) ----
    if (tid == DT1_ISSUER) bulk_wait_all();
    wait_wgrad();
    __syncthreads();
    // per-warp column sums, [BT / 32][4][32] floats in the X tile (free by now)
    float (*red)[4][32] = reinterpret_cast<float (*)[4][32]>(S.Xt[0]);
    red[warp][0][lane] = gb2;
    red[warp][1][lane] = gbc1;
    red[warp][2][lane] = gwc2;
    red[warp][3][lane] = gwa;
    gba = warp_sum(gba);
    if (lane == 0) S.p_za[0][warp] = gba;
    __syncthreads();
    if (tid < 256) {
        const int v = tid >> 6, n = tid & 63;   // 4 vectors x 64 channels
        const int b = n / CW;                   // column block: warps 4 b .. 4 b + 3
        float s = 0.0f;
        for (int w = 0; w < 4; ++w) s += red[4 * b + w][v][n % CW];
        S.acc_vec[v][n] = s;
    }
    if (tid == 0) {
        float s0 = 0.0f;
        for (int wi = 0; wi < 4; ++wi) s0 += S.p_za[0][wi];   // warps 0..3 hold edges
        S.acc_s[0] = s0;
        S.acc_s[1] = 0.0f;
    }
    __syncthreads();
    if (fast_classes) {
        // the S6c riders: warp partials -> channel totals, in warp order (the SG
        // tile is free by now)
        // [BT / 16][1 + FAST_CLASSES][64] floats in the activation tiles (free by now)
        static_assert(sizeof(float) * (BT / 16) * (1 + FAST_CLASSES) * 64 <= 6 * TE * 128,
                      "rider scratch fits the activation tiles");
        float *wp = reinterpret_cast<float *>(S.S1[0]);
        const int slot = 2 * warp + (lane >> 4), l16 = lane & 15;
        *reinterpret_cast<float4 *>(&wp[(slot * (1 + FAST_CLASSES)) * 64 + 4 * l16]) = awr;
#pragma unroll
        for (int c = 0; c < FAST_CLASSES; ++c)
            *reinterpret_cast<float4 *>(
                &wp[(slot * (1 + FAST_CLASSES) + 1 + c) * 64 + 4 * l16]) = aT[c];
        __syncthreads();
        for (int idx = tid; idx < (1 + FAST_CLASSES) * 64; idx += BT) {
            const int v = idx >> 6, n = idx & 63;
            float s_ = 0.0f;
            for (int w = 0; w < BT / 16; ++w) s_ += wp[(w * (1 + FAST_CLASSES) + v) * 64 + n];
            if (v == 0) S.acc_vec[4][n] += s_;
            else S.acc_T[v - 1][n] += s_;
        }
        __syncthreads();
    }
    float *out = a.partial + (size_t)blockIdx.x * EP_STRIDE;
    // weight-gradient accumulators: UMMA M = 64 puts row n on TMEM lane
    // 32 (n / 16) + n % 16, so lanes 0..15 of warps 0..3 read 16 rows each
    if (warp < 4) {
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float w2v[16], wc1v[16];
            tmem_ld16(tl + C_DW2 + 16 * q, w2v);
            tmem_ld16(tl + C_DWC1 + 16 * q, wc1v);
            if (lane < 16) {
                const int n = 16 * warp + lane;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    out[EP_W2 + n * 64 + 16 * q + i] = acc_w2 ? w2v[i] : 0.0f;
                    out[EP_WC1 + n * 64 + 16 * q + i] = acc_wc1 ? wc1v[i] : 0.0f;
                }
            }
        }
        tc_fence_before();
    }
    for (int n = tid; n < 64; n += BT) {
        out[EP_B2 + n] = S.acc_vec[0][n];
        out[EP_BC1 + n] = S.acc_vec[1][n];
        out[EP_WC2 + n] = S.acc_vec[2][n];
        out[EP_WA + n] = S.acc_vec[3][n];
        out[EP_WR + n] = S.acc_vec[4][n];
        for (int c = 0; c < PVS_MAX_EDGE_CLASSES; ++c) out[EP_T + c * 64 + n] = S.acc_T[c][n];
    }
    if (tid == 0) { out[EP_BA] = S.acc_s[0]; out[EP_GATE] = S.acc_s[1]; }
    __syncthreads();
    if (tid < 32) tmem_dealloc<512>(tmem_base);
}

}  // namespace

#ifdef PVS_PHASE_PROF
extern "C" int pvs_debug_bwd_phase_cycles(unsigned long long *out, int reset) {
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_bwd_phase_cycles, sizeof(g_bwd_phase_cycles));
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_bwd_phase_cycles, z, sizeof(z));
    }
    return 0;
}
#endif

bool edge_bwd_tc_supported(const EdgeBwdArgs &a) {
    const bool eres = (a.flags & PVS_F_EDGE_RESIDUAL) && a.m_prev != nullptr;
    const bool softmax = (a.flags & PVS_F_EDGE_ATTENTION) &&
                         (a.flags & PVS_F_SOFTMAX_ATTENTION);
    return !eres && !softmax && a.d_m_out == nullptr && a.d_m_prev == nullptr &&
           a.alpha_in == nullptr;
}

int launch_edge_bwd_tc(const EdgeBwdArgs &a, int grid, cudaStream_t st) {
    const size_t smem = sizeof(BwdTcSmem) + 1024;
    const int rc = ensure_smem(egnn_edge_bwd_tc_kernel, smem);
    if (rc) return rc;
    launch_chained(egnn_edge_bwd_tc_kernel, dim3(grid), dim3(BT), smem, st, a);
    return PVS_OK;
}

}  // namespace pvs
