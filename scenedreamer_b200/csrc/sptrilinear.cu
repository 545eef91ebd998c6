// voxlib.sp_trilinear_worldcoord / sp_trilinear_worldcoord_backward: sparse tri-linear interpolation of per-corner feature
// vectors at world coordinates, corner ids looked up on the fly in a dense int32 volume (GANcraft's block features).
// Behavioural contract: imaginaire/model_utils/gancraft/voxlib/sp_trilinear_worldcoord_kernel.cu:48-338 (kernels),
// :351-437, :453-520 (host), bound at voxlib.cpp:15,17,27-28.  SceneDreamer itself never calls it (only
// generators/gancraft_base.py:442 does); it is here so that the `voxlib` drop-in has the reference's whole surface.
//
// Semantics restated: cell = floor(p), local = p - floor(p); weight of corner (a,b,c) = prod of local / 1-local in the
// order (x, y, z) with z the fastest corner bit; corner cells are CLAMPED to the volume; a NaN coordinate selects nothing;
// with ign_zero the stored ids are 1-based and 0 means "no feature"; out[c] = sum_j fmaf(feature[id_j][c], w_j, acc) in
// corner order j = 0..7 (bit-identical to the reference's accumulation).  Backward: feature_grad[id_j][c] += g[c] * w_j.
//
// Mapping: one warp per entry, lanes over the channels (coalesced feature rows; the 8 ids / weights are computed once per
// lane -- 30 flops -- instead of being exchanged).  HBM/L2-gather bound: 8 rows of C floats per entry.
#include "common.cuh"

namespace {

struct SpParams {
    long long E;
    int C;
    long long M;
    long long dims[3], strides[3];
    int ign_zero;
};

__device__ __forceinline__ void sp_corners(const SpParams &p, const int32_t *__restrict__ lut, const float *__restrict__ wc,
                                           long long e, int (&idx)[8], float (&w)[8])
{
    const float x = wc[e * 3 + 0], y = wc[e * 3 + 1], z = wc[e * 3 + 2];
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const float lx = x - fx, ly = y - fy, lz = z - fz;
    const float ax[2] = {1.0f - lx, lx}, ay[2] = {1.0f - ly, ly}, az[2] = {1.0f - lz, lz};
#pragma unroll
    for (int j = 0; j < 8; j++) w[j] = __fmul_rn(__fmul_rn(ax[(j >> 2) & 1], ay[(j >> 1) & 1]), az[j & 1]);
    if (isnan(x) || isnan(y) || isnan(z)) {
#pragma unroll
        for (int j = 0; j < 8; j++) idx[j] = -1;
    } else {
        const int v[3] = {(int)fx, (int)fy, (int)fz};
        long long o[3][2];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int hi = (int)p.dims[d] - 1;
            o[d][0] = p.strides[d] * (long long)min(max(v[d], 0), hi);
            o[d][1] = p.strides[d] * (long long)min(max(v[d] + 1, 0), hi);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) idx[j] = __ldg(lut + o[0][(j >> 2) & 1] + o[1][(j >> 1) & 1] + o[2][j & 1]);
    }
    if (p.ign_zero) {
#pragma unroll
        for (int j = 0; j < 8; j++) idx[j] -= 1;
    }
}

__global__ void __launch_bounds__(256)
sp_trilinear_forward_kernel(const float *__restrict__ feat, const int32_t *__restrict__ lut, const float *__restrict__ wc,
                            float *__restrict__ out, const SpParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
    for (long long e = warp0; e < p.E; e += nwarps) {
        int idx[8];
        float w[8];
        sp_corners(p, lut, wc, e, idx, w);
        for (int c = lane; c < p.C; c += 32) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (idx[j] >= 0) acc = __fmaf_rn(__ldg(feat + (long long)idx[j] * p.C + c), w[j], acc);
            out[e * p.C + c] = acc;
        }
    }
}

__global__ void __launch_bounds__(256)
sp_trilinear_backward_kernel(const float *__restrict__ gout, const int32_t *__restrict__ lut, const float *__restrict__ wc,
                             float *__restrict__ gfeat, const SpParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (long long)blockDim.x) >> 5;
    for (long long e = warp0; e < p.E; e += nwarps) {
        int idx[8];
        float w[8];
        sp_corners(p, lut, wc, e, idx, w);
        for (int c = lane; c < p.C; c += 32) {
            const float g = gout[e * p.C + c];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (idx[j] >= 0) atomicAdd(gfeat + (long long)idx[j] * p.C + c, __fmul_rn(g, w[j]));
        }
    }
}

int sp_fill(SpParams &p, int64_t M, int32_t C, const int64_t *dims, const int64_t *strides, int64_t E, int ign_zero) {
    if (!dims || !strides || M < 0 || C < 1 || E < 0) return SDB_EINVAL;
    p.E = E; p.C = C; p.M = M; p.ign_zero = ign_zero ? 1 : 0;
    for (int d = 0; d < 3; d++) {
        if (dims[d] < 1) return SDB_EINVAL;
        p.dims[d] = dims[d];
        p.strides[d] = strides[d];
    }
    return SDB_OK;
}

int sp_grid(long long E) {
    const long long want = (E + 7) / 8;                       // 8 warps per CTA
    const long long cap = (long long)sdb_num_sms() * 16;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}
}  // namespace

extern "C" int sdb_sp_trilinear_worldcoord(const float *d_feature, int64_t M, int32_t C, const int32_t *d_corner_lut,
                                           const int64_t lut_dims[3], const int64_t lut_strides[3], const float *d_worldcoord,
                                           int64_t E, int ign_zero, float *d_out, void *stream)
{
    SpParams p;
    const int rc = sp_fill(p, M, C, lut_dims, lut_strides, E, ign_zero);
    if (rc != SDB_OK) return rc;
    if (E == 0) return SDB_OK;
    if (!d_feature || !d_corner_lut || !d_worldcoord || !d_out) return SDB_EINVAL;
    sp_trilinear_forward_kernel<<<sp_grid(E), 256, 0, (cudaStream_t)stream>>>(d_feature, d_corner_lut, d_worldcoord, d_out, p);
    SDB_CHECK_LAUNCH();
    return SDB_OK;
}

extern "C" int sdb_sp_trilinear_worldcoord_backward(const float *d_out_grad, int64_t M, int32_t C, const int32_t *d_corner_lut,
                                                    const int64_t lut_dims[3], const int64_t lut_strides[3],
                                                    const float *d_worldcoord, int64_t E, int ign_zero, float *d_feature_grad,
                                                    void *stream)
{
    SpParams p;
    const int rc = sp_fill(p, M, C, lut_dims, lut_strides, E, ign_zero);
    if (rc != SDB_OK) return rc;
    if (!d_feature_grad && M > 0) return SDB_EINVAL;
    SDB_CUDA(cudaMemsetAsync(d_feature_grad, 0, (size_t)M * C * sizeof(float), (cudaStream_t)stream));
    if (E == 0) return SDB_OK;
    if (!d_out_grad || !d_corner_lut || !d_worldcoord) return SDB_EINVAL;
    sp_trilinear_backward_kernel<<<sp_grid(E), 256, 0, (cudaStream_t)stream>>>(d_out_grad, d_corner_lut, d_worldcoord, d_feature_grad, p);
    SDB_CHECK_LAUNCH();
    return SDB_OK;
}
