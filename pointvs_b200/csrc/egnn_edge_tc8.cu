// K2 edge stage, tcgen05 bf16x3, variant "tc8": 8-warp groups and the message
// segment-reduce on the tensor core.
//
// Same tiles, operands and arithmetic as egnn_edge_tc_kernel<TCM_BF16X3>
// (egnn_edge_tc.cu; reference egnn_satorras.py:123-176).  Two changes:
//
//  * A group is 8 warps instead of 4: two warps share each TMEM lane quarter and
//    take 32 of the 64 accumulator columns each, so the per-thread epilogue
//    chain is half as long.  Four groups per CTA (1024 threads, 64 registers).
//
//  * M_i = sum_e alpha_e m_e is no longer summed by threads reading the m tile
//    back from shared memory (the largest single phase of the round-1 kernel:
//    -11 % when removed, profiles/README.md).  It is ONE more small GEMM per
//    tile, contracted over the tile's 128 edges:
//        M^T [64 ch x 16 nodes] = m^T [64 ch x 128 edges] . (alpha S)^T
//    A = the m tile as it already sits in shared memory, read MN-major (a
//    128-byte row of 64 channels per edge = K index); B = alpha S [16 node rows
//    x 128 edges], K-major, one non-zero per edge (alpha_e in the row of the
//    edge's destination node), bf16 hi + lo like every other operand
//    (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo); D = 16 TMEM columns in the UMMA
//    M = 64 layout.  Each thread writes (and afterwards clears) two 2-byte
//    entries of B; lanes 0..15 of four warps read the 16 columns back and store
//    the node rows.  Tiles with more than 16 destination nodes take further
//    16-node passes through the same B tile.  The additions happen inside the
//    tensor core in a fixed order: still no atomics, bitwise reproducible.
//
// Selected at run time by PVS_EDGE_TC8 (see launch_edge_tc).
#include "egnn_common.cuh"
#include "tc_common.cuh"

namespace pvs {

namespace {

constexpr int G8 = 4;                 // groups per CTA
constexpr int GT8 = 256;              // threads per group
constexpr int T8_THREADS = G8 * GT8;
constexpr int T8_K = 64;
constexpr uint32_t T8_GROUP_COLS = 128;   // D (64) + D_M (16), power-of-two stride
constexpr uint32_t T8_DM_COL = 64;
constexpr int NCH = 16;               // destination nodes per reduce pass

struct alignas(16) Tc8Misc {
    float e_rad[TE], e_dx[TE], e_dy[TE], e_dz[TE];
    float e_zp[2][TE];        // logit partials of the two column halves, then
                              // (same storage) the coordinate-head partials
    int e_col[TE];
    uint8_t e_rowl[TE], e_attr[TE];
    int rp[TN + 1];
    float xsum[TN][3];
    int split_lo, split_hi;
    uint64_t mbar;
};

struct __align__(1024) Tc8Smem {
    uint8_t A_hi[G8][TE * 128];
    uint8_t A_lo[G8][TE * 128];
    // alpha S, K-major: [hi / lo][K atom of 64 edges][16 node rows x 128 bytes]
    uint8_t B[G8][2][2][NCH * 128];
    uint8_t W2_hi[T8_K * 128], W2_lo[T8_K * 128];
    uint8_t Wc1_hi[T8_K * 128], Wc1_lo[T8_K * 128];
    float b2[64], bc1[64], wc2[64], wa[64], wr[64];
    float T[PVS_MAX_EDGE_CLASSES][64];
    Tc8Misc grp[G8];
    uint32_t tmem_base;
};

__device__ __forceinline__ void gsync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT8) : "memory");
}

// M = 64, N = 16, bf16, fp32 accumulate, A MN-major (bit 15), B K-major
constexpr uint32_t IDESC_REDUCE = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) |
                                  ((16u >> 3) << 17) | ((64u >> 4) << 24);

__device__ __forceinline__ void gemm_x3(uint32_t d, const uint8_t *a_hi, const uint8_t *a_lo,
                                        const uint8_t *b_hi, const uint8_t *b_lo) {
    const uint64_t ah = make_desc(smem_u32(a_hi)), al = make_desc(smem_u32(a_lo));
    const uint64_t bh = make_desc(smem_u32(b_hi)), bl = make_desc(smem_u32(b_lo));
    uint32_t acc = 0;
#pragma unroll
    for (int ks = 0; ks < T8_K / 16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        umma_bf16(d, ah + adv, bh + adv, TC_IDESC, acc);
        acc = 1;
        umma_bf16(d, al + adv, bh + adv, TC_IDESC, 1);
        umma_bf16(d, ah + adv, bl + adv, TC_IDESC, 1);
    }
}
// D_M [64 ch x 16 nodes] = m^T . (alpha S)^T over the 128 edge rows
__device__ __forceinline__ void gemm_reduce(uint32_t d, const uint8_t *a_hi, const uint8_t *a_lo,
                                            const uint8_t (*b)[2][NCH * 128]) {
    const uint64_t ah = make_desc(smem_u32(a_hi)), al = make_desc(smem_u32(a_lo));
    uint32_t acc = 0;
#pragma unroll
    for (int ks = 0; ks < TE / 16; ++ks) {
        const uint64_t aadv = (uint64_t)(ks * (2048 >> 4));     // 16 edge rows
        const uint64_t bh = make_desc(smem_u32(b[0][ks >> 2])) + (uint64_t)((ks & 3) * 2);
        const uint64_t bl = make_desc(smem_u32(b[1][ks >> 2])) + (uint64_t)((ks & 3) * 2);
        umma_bf16(d, ah + aadv, bh, IDESC_REDUCE, acc);
        acc = 1;
        umma_bf16(d, al + aadv, bh, IDESC_REDUCE, 1);
        umma_bf16(d, ah + aadv, bl, IDESC_REDUCE, 1);
    }
}

__global__ void __launch_bounds__(T8_THREADS, 1)
egnn_edge_tc8_kernel(const EdgeArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    Tc8Smem &S = *reinterpret_cast<Tc8Smem *>(smem_dyn);
    if ((smem_u32(smem_dyn) & 1023u) != 0u) __trap();
    const int g = threadIdx.x / GT8, tid = threadIdx.x % GT8;
    const int lane = tid & 31, warp = tid >> 5;
    const int erow = 32 * (warp & 3) + lane;      // edge row of this thread (epilogues)
    const int hf = warp >> 2;                     // column half
    Tc8Misc &Gm = S.grp[g];
    uint8_t *A_hi = S.A_hi[g], *A_lo = S.A_lo[g];
    const int k = a.k;
    const bool f_att = a.flags & PVS_F_EDGE_ATTENTION;
    const bool f_softmax = f_att && (a.flags & PVS_F_SOFTMAX_ATTENTION);
    const bool f_coords = (a.flags & PVS_F_UPDATE_COORDS) && a.x_out != nullptr;
    const bool f_eres = (a.flags & PVS_F_EDGE_RESIDUAL) && a.m_prev != nullptr;

    // ---- one-time setup (whole CTA) ----
    load_weight_tiles<true>(S.W2_hi, S.W2_lo, a.edge_w2, k, k, k);
    load_weight_tiles<true>(S.Wc1_hi, S.Wc1_lo, a.coord_w1, k, k, k);
    const int col_r = (a.flags & PVS_F_PERM_INVARIANT) ? k : 2 * k;
    for (int n = threadIdx.x; n < 64; n += T8_THREADS) {
        const bool ok = n < k;
        S.b2[n] = ok ? a.edge_b2[n] : 0.0f;
        S.bc1[n] = ok ? a.coord_b1[n] : 0.0f;
        S.wc2[n] = ok ? a.coord_w2[n] : 0.0f;
        S.wa[n] = (ok && a.att_w) ? a.att_w[n] : 0.0f;
        S.wr[n] = ok ? a.edge_w1[(size_t)n * a.in_e + col_r] : 0.0f;
        for (int c = 0; c < PVS_MAX_EDGE_CLASSES; ++c)
            S.T[c][n] = (ok && c < a.n_classes)
                            ? a.edge_w1[(size_t)n * a.in_e + col_r + 1 + c] : 0.0f;
    }
    for (int i = threadIdx.x; i < (int)(sizeof(S.B) / 16); i += T8_THREADS)
        reinterpret_cast<uint4 *>(&S.B[0][0][0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) mbar_init(&Gm.mbar, 1);
    if (threadIdx.x < 32) tmem_alloc<512>(&S.tmem_base);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t tmem_grp = tmem_base + (uint32_t)g * T8_GROUP_COLS;
    const uint32_t tmem_q = tmem_grp + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t phase = 0;
    const float att_b = (f_att && a.att_b) ? a.att_b[0] : 0.0f;
    float gate = 1.0f;
    if (f_eres && a.edge_gate) gate = a.edge_gate[0];
    const int n_tiles = *a.n_ptiles;
    const int E_total = a.row_ptr[a.n_nodes];

    for (int t = blockIdx.x * G8 + g; t < n_tiles; t += gridDim.x * G8) {
        const int E0 = t * TE, E1 = min(E0 + TE, E_total);
        int cover_lo = 0;
        if (t > 0) {
            const int na = a.ptile_last[t - 1];
            cover_lo = a.row_ptr[na + 1] > E0 ? na : na + 1;
        }
        const int cover_hi = (t == n_tiles - 1) ? a.n_nodes - 1 : a.ptile_last[t];
        for (int n0 = cover_lo; n0 <= cover_hi; n0 += TN) {
        const int nn = min(TN, cover_hi - n0 + 1);
        gsync(g);
        for (int i = tid; i <= nn; i += GT8) {
            const int v = a.row_ptr[n0 + i];
            Gm.rp[i] = min(max(v, E0), E1);
            if (i == 0) Gm.split_lo = v < E0;
            if (i == nn) Gm.split_hi = v > E1;
        }
        for (int i = tid; i < nn * 3; i += GT8) (&Gm.xsum[0][0])[i] = 0.0f;
        gsync(g);
        const int e0 = Gm.rp[0], e1 = Gm.rp[nn];
        const int c0 = e0;
        const int ne = e1 - e0;      // <= 128: one MMA tile
        // row of M (or of the partial slots) that receives node nl of the window
        auto m_row = [&](int nl) -> float * {
            if (nl == 0 && Gm.split_lo) return a.Mpart + ((size_t)t * 2) * T8_K;
            if (nl == nn - 1 && Gm.split_hi) return a.Mpart + ((size_t)t * 2 + 1) * T8_K;
            return a.M + (size_t)(n0 + nl) * T8_K;
        };
        if (ne > 0) {
            // ---- stage 0: geometry, one thread per edge ----
            if (tid < ne) {
                const int e = c0 + tid;
                int lo = 0, hi = nn;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (Gm.rp[mid] <= e) lo = mid; else hi = mid;
                }
                const int i = n0 + lo, j = a.col[e];
                float dx = a.x_in[3 * i] - a.x_in[3 * j];
                float dy = a.x_in[3 * i + 1] - a.x_in[3 * j + 1];
                float dz = a.x_in[3 * i + 2] - a.x_in[3 * j + 2];
                float r = dx * dx + dy * dy + dz * dz;
                if (a.flags & PVS_F_NORMALIZE) {
                    float inv = 1.0f / (sqrtf(r) + 1e-8f);
                    dx *= inv; dy *= inv; dz *= inv;
                }
                Gm.e_rowl[tid] = (uint8_t)lo;
                Gm.e_col[tid] = j;
                Gm.e_attr[tid] = a.attr ? a.attr[e] : 0;
                Gm.e_rad[tid] = r;
                Gm.e_dx[tid] = dx; Gm.e_dy[tid] = dy; Gm.e_dz[tid] = dz;
            }
            gsync(g);
            // ---- stage 1: s1 = silu(P_i + Q_j + w_r r + T[a]) -> A tile; 8 lanes
            // per edge row, 32 rows a pass, loads two passes ahead ----
            {
                const int c = tid & 7, slot = tid >> 3;
                float2 wr2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    wr2[i] = *reinterpret_cast<const float2 *>(&S.wr[8 * c + 2 * i]);
                constexpr int ROWS = GT8 / 8, PASSES = TE / ROWS, DEPTH = 2;
                float4 buf[DEPTH][4];
                auto issue = [&](int p, float4 (&bq)[4]) {
                    const int r = min(p * ROWS + slot, ne - 1);
                    const float4 *pp = reinterpret_cast<const float4 *>(
                        a.P + (size_t)(n0 + Gm.e_rowl[r]) * T8_K + 8 * c);
                    const float4 *qq = reinterpret_cast<const float4 *>(
                        a.Q + (size_t)Gm.e_col[r] * T8_K + 8 * c);
                    bq[0] = __ldg(pp); bq[1] = __ldg(pp + 1);
                    bq[2] = __ldg(qq); bq[3] = __ldg(qq + 1);
                };
                auto finish = [&](int p, const float4 (&bq)[4]) {
                    const int r = p * ROWS + slot, re = min(r, ne - 1);
                    const float rad = Gm.e_rad[re];
                    const int at = Gm.e_attr[re];
                    const float2 rad2 = make_float2(rad, rad);
                    const float4 *t4 = reinterpret_cast<const float4 *>(&S.T[at][8 * c]);
                    const float4 ta = t4[0], tb = t4[1];
                    const float2 pv[4] = {make_float2(bq[0].x, bq[0].y), make_float2(bq[0].z, bq[0].w),
                                          make_float2(bq[1].x, bq[1].y), make_float2(bq[1].z, bq[1].w)};
                    const float2 qv[4] = {make_float2(bq[2].x, bq[2].y), make_float2(bq[2].z, bq[2].w),
                                          make_float2(bq[3].x, bq[3].y), make_float2(bq[3].z, bq[3].w)};
                    const float2 tv[4] = {make_float2(ta.x, ta.y), make_float2(ta.z, ta.w),
                                          make_float2(tb.x, tb.y), make_float2(tb.z, tb.w)};
                    float2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        v[i] = ffma2(wr2[i], rad2, fadd2(fadd2(pv[i], qv[i]), tv[i]));
                    silu4_(v[0], v[1]);
                    silu4_(v[2], v[3]);
                    uint4 hi, lo;
                    split8p<true>(v, hi, lo);
                    *reinterpret_cast<uint4 *>(A_hi + swz(r, c)) = hi;
                    *reinterpret_cast<uint4 *>(A_lo + swz(r, c)) = lo;
                };
#pragma unroll
                for (int p = 0; p < DEPTH; ++p) issue(p, buf[p]);
#pragma unroll
                for (int p = 0; p < PASSES; ++p) {
                    finish(p, buf[p % DEPTH]);
                    if (p + DEPTH < PASSES) issue(p + DEPTH, buf[p % DEPTH]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            gsync(g);
            // ---- GEMM 1: t2 = s1 . W2^T ----
            if (tid == 0) {
                tc_fence_after();
                gemm_x3(tmem_grp, A_hi, A_lo, S.W2_hi, S.W2_lo);
                umma_commit(&Gm.mbar);
            }
            mbar_wait(&Gm.mbar, phase);
            phase ^= 1;
            tc_fence_after();
            // ---- epilogue 1: m = silu(t2 + b2) (+ edge residual) -> A tile; the
            // attention-logit partial of this column half ----
            {
                const int r = erow;
                float2 dot2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
                for (int qq = 0; qq < 2; ++qq) {
                    const int q = 2 * hf + qq;
                    float acc[16];
                    tmem_ld16(tmem_q + 16 * q, acc);
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        const int nb = 16 * q + 8 * hlf;
                        const float4 *b4 = reinterpret_cast<const float4 *>(&S.b2[nb]);
                        const float4 *w4 = reinterpret_cast<const float4 *>(&S.wa[nb]);
                        const float4 ba = b4[0], bb = b4[1], wa0 = w4[0], wa1 = w4[1];
                        const float2 bias[4] = {make_float2(ba.x, ba.y), make_float2(ba.z, ba.w),
                                                make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};
                        const float2 wat[4] = {make_float2(wa0.x, wa0.y), make_float2(wa0.z, wa0.w),
                                               make_float2(wa1.x, wa1.y), make_float2(wa1.z, wa1.w)};
                        float2 mv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            mv[i] = fadd2(
                                make_float2(acc[8 * hlf + 2 * i], acc[8 * hlf + 2 * i + 1]),
                                bias[i]);
                        silu4_(mv[0], mv[1]);
                        silu4_(mv[2], mv[3]);
                        if (f_eres && r < ne) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int n = nb + i;
                                if (n < k) {
                                    float &mref = (i & 1) ? mv[i >> 1].y : mv[i >> 1].x;
                                    float mp = a.m_prev[(size_t)(c0 + r) * k + n];
                                    if (a.flags & PVS_F_REZERO) mref = mp + gate * mref;
                                    else if (a.flags & PVS_F_GATED_RESIDUAL) {
                                        float gg = fmaxf(gate, 0.0f);
                                        mref = gg * mref + (1.0f - gg) * mp;
                                    } else mref = mref + mp;
                                }
                            }
                        }
                        if (r >= ne) {
                            // rows past the tile's edges enter the reduce GEMM as
                            // K rows: they must be exactly zero
#pragma unroll
                            for (int i = 0; i < 4; ++i) mv[i] = make_float2(0.0f, 0.0f);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) dot2 = ffma2(wat[i], mv[i], dot2);
                        const int c = 2 * q + hlf;
                        uint4 hi, lo;
                        split8p<true>(mv, hi, lo);
                        *reinterpret_cast<uint4 *>(A_hi + swz(r, c)) = hi;
                        *reinterpret_cast<uint4 *>(A_lo + swz(r, c)) = lo;
                    }
                }
                Gm.e_zp[hf][r] = dot2.x + dot2.y;
            }
            fence_proxy_async();
            tc_fence_before();
            gsync(g);
            // ---- GEMM 2 (coordinate MLP; reuses the D columns) is issued now and
            // committed together with the reduce GEMM below ----
            if (f_coords && tid == 0) {
                tc_fence_after();
                gemm_x3(tmem_grp, A_hi, A_lo, S.Wc1_hi, S.Wc1_lo);
                if (f_softmax) umma_commit(&Gm.mbar);
            }
            // attention value per edge (alpha, or the raw logit when a softmax
            // pass follows)
            float al = 1.0f;
            if (tid < TE && f_att) {
                const float z = Gm.e_zp[0][tid] + Gm.e_zp[1][tid] + att_b;
                al = f_softmax ? z : apply_act(z, a.att_act);
                if (a.att_out && tid < ne) a.att_out[c0 + tid] = al;
            }
            if (f_softmax) gsync(g);     // e_zp is rewritten by epilogue 2
            if (!f_softmax) {
                // ---- M_i = sum_e alpha_e m_e on the tensor core, 16 nodes a pass ----
                const int my_rowl = tid < ne ? Gm.e_rowl[tid] : -1;
                uint32_t ah = 0, alo = 0;
                if (tid < ne) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(al);
                    const __nv_bfloat16 l = __float2bfloat16_rn(al - __bfloat162float(h));
                    ah = *reinterpret_cast<const uint16_t *>(&h);
                    alo = *reinterpret_cast<const uint16_t *>(&l);
                }
                const int n_pass = (nn + NCH - 1) / NCH;
                for (int ps = 0; ps < n_pass; ++ps) {
                    const bool mine = my_rowl >= 0 && (my_rowl / NCH) == ps;
                    uint32_t off = 0;
                    if (mine) {
                        off = (uint32_t)((tid >> 6) * (NCH * 128)) +
                              swz(my_rowl % NCH, (tid & 63) >> 3) + 2u * (tid & 7);
                        *reinterpret_cast<uint16_t *>(&S.B[g][0][0][0] + off) = (uint16_t)ah;
                        *reinterpret_cast<uint16_t *>(&S.B[g][1][0][0] + off) = (uint16_t)alo;
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    gsync(g);
                    if (tid == 0) {
                        tc_fence_after();
                        gemm_reduce(tmem_grp + T8_DM_COL, A_hi, A_lo, S.B[g]);
                        umma_commit(&Gm.mbar);
                    }
                    mbar_wait(&Gm.mbar, phase);
                    phase ^= 1;
                    tc_fence_after();
                    // read back: UMMA M = 64 puts channel c on TMEM lane
                    // 32 (c / 16) + c % 16; columns = the pass's 16 nodes
                    if (warp < 4) {
                        float mv[16];
                        tmem_ld16(tmem_q + T8_DM_COL, mv);
                        if (lane < 16) {
                            const int ch = 16 * warp + lane;
#pragma unroll
                            for (int j = 0; j < NCH; ++j) {
                                const int nl = ps * NCH + j;
                                if (nl < nn) m_row(nl)[ch] = mv[j];
                            }
                        }
                        tc_fence_before();
                    }
                    if (mine) {     // leave the B tile all-zero again
                        *reinterpret_cast<uint16_t *>(&S.B[g][0][0][0] + off) = 0;
                        *reinterpret_cast<uint16_t *>(&S.B[g][1][0][0] + off) = 0;
                    }
                    if (ps + 1 < n_pass) gsync(g);   // D_M is read before the next pass
                }
            }
        } else if (!f_softmax) {
            // window without edges: its nodes receive no message
            for (int idx = tid; idx < nn * 16; idx += GT8)
                reinterpret_cast<float4 *>(m_row(idx >> 4))[idx & 15] =
                    make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- messages out (edge residual of the next layer / softmax) ----
        if (a.m_out != nullptr) {
            for (int idx = tid; idx < ne * (T8_K / 2); idx += GT8) {
                const int el = idx >> 5, w = idx & 31;
                const uint32_t off = swz(el, w >> 2) + ((w & 3) << 2);
                const uint32_t h = *reinterpret_cast<const uint32_t *>(A_hi + off);
                const uint32_t l = *reinterpret_cast<const uint32_t *>(A_lo + off);
                const float m0 = __uint_as_float(h << 16) + __uint_as_float(l << 16);
                const float m1 = __uint_as_float(h & 0xffff0000u) +
                                 __uint_as_float(l & 0xffff0000u);
                float *dst = a.m_out + (size_t)(c0 + el) * a.ld_m;
                if (2 * w < a.ld_m) dst[2 * w] = m0;
                if (2 * w + 1 < a.ld_m) dst[2 * w + 1] = m1;
            }
        }
        if (ne > 0 && f_coords) {
            // ---- epilogue 2: c = [tanh](wc2 . silu(Wc1 m + bc1)) ----
            if (f_softmax) {        // not yet waited for (no reduce pass)
                mbar_wait(&Gm.mbar, phase);
                phase ^= 1;
                tc_fence_after();
            }
            float2 d2 = make_float2(0.0f, 0.0f);
#pragma unroll 1
            for (int qq = 0; qq < 2; ++qq) {
                const int q = 2 * hf + qq;
                float acc[16];
                tmem_ld16(tmem_q + 16 * q, acc);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(&S.bc1[16 * q + 4 * v4]);
                    const float4 ww = *reinterpret_cast<const float4 *>(&S.wc2[16 * q + 4 * v4]);
                    float2 s0 = fadd2(make_float2(acc[4 * v4], acc[4 * v4 + 1]),
                                      make_float2(bb.x, bb.y));
                    float2 s1 = fadd2(make_float2(acc[4 * v4 + 2], acc[4 * v4 + 3]),
                                      make_float2(bb.z, bb.w));
                    silu4_(s0, s1);
                    d2 = ffma2(make_float2(ww.x, ww.y), s0, d2);
                    d2 = ffma2(make_float2(ww.z, ww.w), s1, d2);
                }
            }
            Gm.e_zp[hf][erow] = d2.x + d2.y;
            tc_fence_before();
        }
        gsync(g);
        // ---- coordinate messages d_e c_e (one thread per edge; the reference
        // rounds the product before its scatter_add too: egnn_satorras.py:172-173),
        // then summed per node ----
        if (f_coords && ne > 0) {
            if (tid < TE) {
                float c = Gm.e_zp[0][tid] + Gm.e_zp[1][tid];
                if (a.flags & PVS_F_TANH) c = tanhf(c);
                Gm.e_dx[tid] *= c;
                Gm.e_dy[tid] *= c;
                Gm.e_dz[tid] *= c;
            }
            gsync(g);
        }
        if (f_coords && tid < nn) {
            const int lo = max(Gm.rp[tid], c0) - c0;
            const int hi = min(Gm.rp[tid + 1], c0 + TE) - c0;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (int el = lo; el < hi; ++el) {
                sx += Gm.e_dx[el];
                sy += Gm.e_dy[el];
                sz += Gm.e_dz[el];
            }
            Gm.xsum[tid][0] += sx;
            Gm.xsum[tid][1] += sy;
            Gm.xsum[tid][2] += sz;
        }
        if (a.x_out != nullptr && tid < nn) {
            const int i = n0 + tid;
            const bool sp_lo = tid == 0 && Gm.split_lo;
            const bool sp_hi = tid == nn - 1 && Gm.split_hi;
            if (sp_lo || sp_hi) {
                float *xp = a.xpart + ((size_t)t * 2 + (sp_lo ? 0 : 1)) * 4;
                xp[0] = Gm.xsum[tid][0];
                xp[1] = Gm.xsum[tid][1];
                xp[2] = Gm.xsum[tid][2];
            } else {
                const int cnt = Gm.rp[tid + 1] - Gm.rp[tid];
                const float inv = 1.0f / (float)(cnt > 0 ? cnt : 1);
                float ax = 0.f, ay = 0.f, az = 0.f;
                if (f_coords) {
                    ax = Gm.xsum[tid][0] * inv;
                    ay = Gm.xsum[tid][1] * inv;
                    az = Gm.xsum[tid][2] * inv;
                }
                a.x_out[3 * i] = a.x_in[3 * i] + ax;
                a.x_out[3 * i + 1] = a.x_in[3 * i + 1] + ay;
                a.x_out[3 * i + 2] = a.x_in[3 * i + 2] + az;
            }
        }
        }   // node windows of the tile
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
}

}  // namespace

int launch_edge_tc8(const EdgeArgs &a, int n_ptiles_cap, cudaStream_t st) {
    const size_t smem = sizeof(Tc8Smem);
    int grid = num_sms();   // one persistent CTA per SM
    const int need = (n_ptiles_cap + G8 - 1) / G8;
    if (need < grid) grid = need;
    if (grid < 1) grid = 1;
    const int rc = ensure_smem(egnn_edge_tc8_kernel, smem);
    if (rc) return rc;
    egnn_edge_tc8_kernel<<<grid, T8_THREADS, smem, st>>>(a);
    return PVS_OK;
}

}  // namespace pvs
