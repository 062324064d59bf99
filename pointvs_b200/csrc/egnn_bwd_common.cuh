// Argument block and per-CTA partial layout shared by the two edge-backward
// kernels (egnn_bwd.cu: fp32 FFMA; egnn_edge_bwd_tc.cu: tcgen05).
#pragma once
#include "egnn_common.cuh"

namespace pvs {

constexpr int EP_W2 = 0, EP_WC1 = 4096, EP_B2 = 8192, EP_BC1 = 8256, EP_WC2 = 8320,
              EP_WA = 8384, EP_WR = 8448, EP_T = 8512, EP_BA = 9024, EP_GATE = 9025,
              EP_STRIDE = 9088;

struct EdgeBwdArgs {
    const int32_t *row_ptr, *col, *tile_ptr, *n_tiles;
    const uint8_t *attr;
    const float *P, *Q, *x_in, *m_prev;
    const float *dM;        // [N][64]
    const float *d_x_out;   // [N][3] or null
    const float *d_m_out;   // [E][k] or null
    float *dP;              // [N][64]
    float *DT1;             // two planes [2][dt1_rows][32] (dt1_at below)
    int64_t dt1_rows;       // rows per plane (the workspace's edge capacity)
    float *DD;              // [E][3]
    float *d_x_in;          // [N][3]
    float *d_m_prev;        // [E][k] or null
    float *partial;         // [grid][EP_STRIDE]
    const float *alpha_in;  // [E] softmax attention values (softmax mode) or null
    const float *seg_s;     // [N] sum over the dst segment of alpha * d(alpha)
    const float *edge_w1, *edge_w2, *edge_b2, *coord_w1, *coord_b1, *coord_w2;
    const float *att_w, *att_b, *edge_gate;
    int k, in_e, n_classes;
    uint32_t flags;
    int att_act;
};

// dt1 (gradient of the first edge layer's pre-activation, [E][64]) round-trips
// HBM between the edge backward and the CSC gather.  It is stored as two planes
// of 32 channels, each row 128 bytes with its 16-byte chunks XOR-swizzled by the
// low bits of the GLOBAL edge index: that is byte for byte the shared-memory
// image of the tcgen05 kernel's dt1 tile, so a tile leaves the SM as two bulk
// asynchronous copies (TMA engine) instead of 2048 per-thread 16-byte stores.
__host__ __device__ __forceinline__ int64_t dt1_at(int64_t dt1_rows, int64_t e, int n) {
    const int ch = (n & 31) >> 2;
    return (int64_t)(n >> 5) * dt1_rows * 32 + e * 32 + (((ch ^ (int)(e & 7)) << 2) | (n & 3));
}

// tcgen05 edge backward (egnn_edge_bwd_tc.cu).  Supports the configurations
// without edge residual / softmax attention / incoming message gradients; the
// caller falls back to the FFMA kernel otherwise.
bool edge_bwd_tc_supported(const EdgeBwdArgs &a);
int launch_edge_bwd_tc(const EdgeBwdArgs &a, int grid, cudaStream_t st);

struct NodeBwdArgs {
    const float *h_in;      // [N][k]
    const float *M;         // [N][64]
    const float *d_h_out;   // [N][k]
    float *d_h_in;          // [N][k]   residual part + Wn1 h-part
    float *dM;              // [N][64]
    float *DO, *U, *DV, *O; // [N][64]  factors of the node weight gradients
    float *dzn, *gdot;      // [N]
    const float *node_w1, *node_b1, *node_w2, *node_b2, *natt_w, *natt_b, *node_gate;
    int n_nodes, k;
    uint32_t flags;
    int att_act;
    // GraphNorm (phase 1: stop at dy = dL/d(gn output); phase 2: resume from dv)
    int phase;                     // 0 = no GraphNorm
    const float *V;                // [N][64] pre-activation (phase 1, 2)
    const float *gn_a, *gn_b;      // y = a v + b
    const float *gn_shift, *gn_invstd;   // c_hat = (v - shift) * invstd
    float *DY, *DYC;               // [N][64] dy and dy * c_hat (phase 1 out, 2 in)
    const float *coef;             // [3][64]: dv = c0 dy + c1 c_hat + c2 (phase 2)
};
// tcgen05 node backward (egnn_node_tc.cu), phase 0 (no GraphNorm) only.
int launch_node_bwd_tc(const NodeBwdArgs &a, cudaStream_t st);

struct BwdSide {
    cudaStream_t stream;
    cudaEvent_t ev_a, ev_b;     // fork points on the main stream
};
extern thread_local const BwdSide *g_bwd_side;

// Grouped weight gradients: up to WG_MAX_JOBS products d_w += A^T B (+ column
// sums of A into d_b) over the same rows, one launch + one reduce (egnn_bwd.cu:
// FFMA; wgrad_tc.cu: tcgen05).  Per-CTA partial block: [64][128] products, then
// 64 column sums.
constexpr int WG_MAX_JOBS = 8;
constexpr int WG_PART = 64 * 128 + 64;
struct WgradJob {
    const float *A; const float *B;   // B == nullptr: column sums of A only
    float *d_w; float *d_b;
    int lda, ko, ldb, ki, ld_dw;
};
struct WgradGroup {
    WgradJob job[WG_MAX_JOBS];
    int n_jobs, rows, chunks, rows_per;   // chunks CTAs per job, rows_per rows each
};
// Fills G.chunks / G.rows_per (multiples of 128 rows, at most max_ctas CTAs) and
// launches the tcgen05 kernel; every job needs ko, ki <= 64.
int launch_wgrad_group_tc(WgradGroup &G, int rows, int max_ctas, float *partial,
                          cudaStream_t st);

}  // namespace pvs
