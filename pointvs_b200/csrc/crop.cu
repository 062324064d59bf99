// K0: receptor crop + atom typing on the device, in front of K1 (row N2).
//
// The reference builds each complex on the host: concat [ligand; receptor],
// keep the receptor atoms closer than `radius` to ANY ligand atom
// (make_box, preprocessing/preprocessing.py:165-195: fp64 scipy cdist
// `sqrt(dx^2+dy^2+dz^2) < radius`), drop hydrogens, one-hot the atom types
// (make_bit_vector, :214-239).  For screening, the same receptor is paired
// with thousands of ligand poses; here the receptor stays resident in HBM and
// one CTA per pose does the distance test, so the host only ships the ligand.
//
//   pvs_crop_count : keep-bit per (pose, receptor atom) + atoms per complex
//   (exclusive scan of the counts by the caller: pvs_exclusive_scan)
//   pvs_crop_fill  : packed coords (fp64), bp, one-hot features for K1 / K2
//
// Bit-exact with the host path: the distance uses the same operation order in
// round-to-nearest fp64 without FMA contraction, and IEEE sqrt.
#include "pvs_common.cuh"

namespace pvs {

constexpr int CROP_THREADS = 256;
constexpr int CROP_LIG_CHUNK = 256;   // ligand atoms staged in shared memory at a time

__global__ void __launch_bounds__(CROP_THREADS)
crop_count_kernel(const double *__restrict__ lig_xyz, const uint8_t *__restrict__ lig_emit,
                  const int32_t *__restrict__ lig_ptr, const double *__restrict__ rec_xyz,
                  const uint8_t *__restrict__ rec_emit, const int32_t *__restrict__ rec_ptr,
                  const int32_t *__restrict__ rec_of_pose, int words, double radius,
                  uint32_t *__restrict__ mask, int32_t *__restrict__ counts) {
    __shared__ double lx[CROP_LIG_CHUNK], ly[CROP_LIG_CHUNK], lz[CROP_LIG_CHUNK];
    __shared__ int s_count;
    const int b = blockIdx.x;
    const int l0 = lig_ptr[b], l1 = lig_ptr[b + 1];
    const int rid = rec_of_pose ? rec_of_pose[b] : 0;
    const int r0 = rec_ptr[rid], r1 = rec_ptr[rid + 1];
    const int nr = min(r1 - r0, words * 32);
    // volatile: lane 0 of the owning warp updates a word that the whole warp
    // re-reads in the next ligand chunk
    volatile uint32_t *mrow = mask + (size_t)b * words;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_count = 0;
    for (int w = threadIdx.x; w < words; w += CROP_THREADS) mrow[w] = 0u;
    int mine = 0;   // emitted ligand atoms seen by this thread
    for (int i = l0 + threadIdx.x; i < l1; i += CROP_THREADS) mine += lig_emit[i] ? 1 : 0;
    const int n_blocks = (nr + CROP_THREADS - 1) / CROP_THREADS;
    for (int c0 = l0; c0 < l1; c0 += CROP_LIG_CHUNK) {
        const int nc = min(CROP_LIG_CHUNK, l1 - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc; i += CROP_THREADS) {
            lx[i] = lig_xyz[3 * (size_t)(c0 + i)];
            ly[i] = lig_xyz[3 * (size_t)(c0 + i) + 1];
            lz[i] = lig_xyz[3 * (size_t)(c0 + i) + 2];
        }
        __syncthreads();
        for (int blk = 0; blk < n_blocks; ++blk) {
            const int j = blk * CROP_THREADS + threadIdx.x;
            const int w = j >> 5;
            // each mask word is owned by one warp, so this read sees its own writes
            const bool todo = j < nr && rec_emit[r0 + j] && !((mrow[w] >> lane) & 1u);
            bool near = false;
            if (todo) {
                const double x = rec_xyz[3 * (size_t)(r0 + j)];
                const double y = rec_xyz[3 * (size_t)(r0 + j) + 1];
                const double z = rec_xyz[3 * (size_t)(r0 + j) + 2];
                for (int i = 0; i < nc; ++i) {
                    const double dx = __dsub_rn(lx[i], x), dy = __dsub_rn(ly[i], y),
                                 dz = __dsub_rn(lz[i], z);
                    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)),
                                                __dmul_rn(dz, dz));
                    if (__dsqrt_rn(d2) < radius) { near = true; break; }
                }
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, near);
            if (lane == 0 && bits) mrow[w] |= bits;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < words; w += CROP_THREADS) mine += __popc(mrow[w]);
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) counts[b] = s_count;
}

__device__ __forceinline__ void write_atom(int out, double x, double y, double z, int bpv,
                                           int code, int n_types, int compact, int n_feat,
                                           double *coords, int32_t *bp, float *feats) {
    coords[3 * (size_t)out] = x;
    coords[3 * (size_t)out + 1] = y;
    coords[3 * (size_t)out + 2] = z;
    bp[out] = bpv;
    float *f = feats + (size_t)out * n_feat;
    for (int k = 0; k < n_feat; ++k) f[k] = 0.0f;
    if (compact) {
        // one-hot of code % n over n + 1 columns; last column := code / n
        // (preprocessing.py:231-234, including its wrap for unmapped elements)
        const int idx = code % n_types;
        f[idx] = 1.0f;
        f[n_feat - 1] = (float)(code / n_types);
    } else if (code >= 0 && code < n_feat) {
        f[code] = 1.0f;
    }
}

__global__ void __launch_bounds__(CROP_THREADS)
crop_fill_kernel(const double *__restrict__ lig_xyz, const uint8_t *__restrict__ lig_emit,
                 const int16_t *__restrict__ lig_code, const int32_t *__restrict__ lig_ptr,
                 const double *__restrict__ rec_xyz, const int16_t *__restrict__ rec_code,
                 const int32_t *__restrict__ rec_ptr, const int32_t *__restrict__ rec_of_pose,
                 const uint32_t *__restrict__ mask, int words,
                 const int32_t *__restrict__ complex_ptr, int n_types, int compact, int n_feat,
                 double *__restrict__ coords, int32_t *__restrict__ bp, float *__restrict__ feats) {
    __shared__ int s_warp[CROP_THREADS / 32];
    __shared__ int s_base;
    const int b = blockIdx.x;
    const int l0 = lig_ptr[b], l1 = lig_ptr[b + 1];
    const int rid = rec_of_pose ? rec_of_pose[b] : 0;
    const int r0 = rec_ptr[rid];
    const uint32_t *mrow = mask + (size_t)b * words;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = complex_ptr[b];
    __syncthreads();
    // ordered compaction, CROP_THREADS candidates a round: ligand, then receptor
    const int n_lig = l1 - l0, n_rec_slots = words * 32;
    const int total = n_lig + n_rec_slots;
    for (int c0 = 0; c0 < total; c0 += CROP_THREADS) {
        const int idx = c0 + threadIdx.x;
        bool keep = false;
        bool is_lig = false;
        int src = 0;
        if (idx < n_lig) {
            is_lig = true;
            src = l0 + idx;
            keep = lig_emit[src] != 0;
        } else if (idx < total) {
            const int j = idx - n_lig;
            src = r0 + j;
            keep = (mrow[j >> 5] >> (j & 31)) & 1u;
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bits);
        __syncthreads();
        int before = 0, round_total = 0;
        for (int w = 0; w < CROP_THREADS / 32; ++w) {
            const int cnt = s_warp[w];
            if (w < warp) before += cnt;
            round_total += cnt;
        }
        const int base = s_base;
        if (keep) {
            const int out = base + before + __popc(bits & ((1u << lane) - 1u));
            if (is_lig)
                write_atom(out, lig_xyz[3 * (size_t)src], lig_xyz[3 * (size_t)src + 1],
                           lig_xyz[3 * (size_t)src + 2], 0, lig_code[src], n_types, compact,
                           n_feat, coords, bp, feats);
            else
                write_atom(out, rec_xyz[3 * (size_t)src], rec_xyz[3 * (size_t)src + 1],
                           rec_xyz[3 * (size_t)src + 2], 1, rec_code[src], n_types, compact,
                           n_feat, coords, bp, feats);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + round_total;
        __syncthreads();
    }
}

}  // namespace pvs

using namespace pvs;

extern "C" {

int pvs_crop_count(const double *lig_xyz, const uint8_t *lig_emit, const int32_t *lig_ptr,
                   int32_t n_poses, const double *rec_xyz, const uint8_t *rec_emit,
                   const int32_t *rec_ptr, const int32_t *rec_of_pose, int32_t mask_words,
                   double radius, uint32_t *mask, int32_t *counts, void *stream) {
    if (n_poses < 0 || mask_words < 0 || !(radius >= 0.0)) return PVS_ERR_INVALID_ARG;
    if (n_poses == 0) return PVS_OK;
    if (!lig_xyz || !lig_emit || !lig_ptr || !rec_xyz || !rec_emit || !rec_ptr || !mask ||
        !counts)
        return PVS_ERR_INVALID_ARG;
    crop_count_kernel<<<n_poses, CROP_THREADS, 0, (cudaStream_t)stream>>>(
        lig_xyz, lig_emit, lig_ptr, rec_xyz, rec_emit, rec_ptr, rec_of_pose, mask_words, radius,
        mask, counts);
    return check_launch();
}

int pvs_crop_fill(const double *lig_xyz, const uint8_t *lig_emit, const int16_t *lig_code,
                  const int32_t *lig_ptr, int32_t n_poses, const double *rec_xyz,
                  const int16_t *rec_code, const int32_t *rec_ptr, const int32_t *rec_of_pose,
                  const uint32_t *mask, int32_t mask_words, const int32_t *complex_ptr,
                  int32_t n_atom_types, int32_t compact, double *coords, int32_t *bp,
                  float *feats, void *stream) {
    if (n_poses < 0 || mask_words < 0 || n_atom_types <= 0) return PVS_ERR_INVALID_ARG;
    if (n_poses == 0) return PVS_OK;
    if (!lig_xyz || !lig_emit || !lig_code || !lig_ptr || !rec_xyz || !rec_code || !rec_ptr ||
        !mask || !complex_ptr || !coords || !bp || !feats)
        return PVS_ERR_INVALID_ARG;
    const int n_feat = compact ? n_atom_types + 1 : 2 * n_atom_types;
    crop_fill_kernel<<<n_poses, CROP_THREADS, 0, (cudaStream_t)stream>>>(
        lig_xyz, lig_emit, lig_code, lig_ptr, rec_xyz, rec_code, rec_ptr, rec_of_pose, mask,
        mask_words, complex_ptr, n_atom_types, compact ? 1 : 0, n_feat, coords, bp, feats);
    return check_launch();
}

}  // extern "C"
