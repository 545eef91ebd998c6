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
// Mapping: a group of G lanes per entry, each lane owning float4 chunks of the channels (G = the power of two >= C/4, at most
// 32: 16 lanes for the 64-channel block features, two entries per warp), so a feature row is fetched with 16-byte loads and
// the gradient scattered with 16-byte vector reductions; the 8 ids / weights are computed once per lane -- 30 flops --
// instead of being exchanged.  C not a multiple of 4 (or unaligned rows): one warp per entry, scalar lanes.
// HBM/L2-gather bound: 8 rows of C floats per entry.
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

template <int G>
__global__ void __launch_bounds__(256)
sp_trilinear_forward_vec4_kernel(const float *__restrict__ feat, const int32_t *__restrict__ lut, const float *__restrict__ wc,
                                 float *__restrict__ out, const SpParams p)
{
    const int sub = threadIdx.x % G, C4 = p.C >> 2;
    const long long g0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G, ng = (gridDim.x * (long long)blockDim.x) / G;
    for (long long e = g0; e < p.E; e += ng) {
        int idx[8];
        float w[8];
        sp_corners(p, lut, wc, e, idx, w);
        for (int c4 = sub; c4 < C4; c4 += G) {
            float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (idx[j] >= 0) {
                    const float4 f = __ldg(reinterpret_cast<const float4 *>(feat + (long long)idx[j] * p.C) + c4);
                    acc.x = __fmaf_rn(f.x, w[j], acc.x);
                    acc.y = __fmaf_rn(f.y, w[j], acc.y);
                    acc.z = __fmaf_rn(f.z, w[j], acc.z);
                    acc.w = __fmaf_rn(f.w, w[j], acc.w);
                }
            reinterpret_cast<float4 *>(out + e * p.C)[c4] = acc;
        }
    }
}

template <int G>
__global__ void __launch_bounds__(256)
sp_trilinear_backward_vec4_kernel(const float *__restrict__ gout, const int32_t *__restrict__ lut, const float *__restrict__ wc,
                                  float *__restrict__ gfeat, const SpParams p)
{
    const int sub = threadIdx.x % G, C4 = p.C >> 2;
    const long long g0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G, ng = (gridDim.x * (long long)blockDim.x) / G;
    for (long long e = g0; e < p.E; e += ng) {
        int idx[8];
        float w[8];
        sp_corners(p, lut, wc, e, idx, w);
        for (int c4 = sub; c4 < C4; c4 += G) {
            const float4 g = reinterpret_cast<const float4 *>(gout + e * p.C)[c4];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (idx[j] >= 0)
                    atomicAdd(reinterpret_cast<float4 *>(gfeat + (long long)idx[j] * p.C) + c4,
                              make_float4(__fmul_rn(g.x, w[j]), __fmul_rn(g.y, w[j]), __fmul_rn(g.z, w[j]), __fmul_rn(g.w, w[j])));
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

int sp_grid(long long E, int lanes_per_entry = 32) {
    const long long per_cta = 256 / lanes_per_entry;          // entries one CTA pass covers
    const long long want = (E + per_cta - 1) / per_cta;
    const long long cap = (long long)sdb_num_sms() * 16;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// lanes per entry of the float4 kernels, 0 = scalar kernel (C % 4 != 0 or rows not 16-byte aligned)
int sp_group(int C, const void *a, const void *b) {
    if (C % 4 != 0 || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return 0;
    int g = 1;
    while (g < 32 && g * 4 < C) g <<= 1;
    return g;
}

#define SP_DISPATCH_G(KERNEL, G, ...)                                                                     \
    switch (G) {                                                                                          \
        case 1: KERNEL<1><<<sp_grid(E, 1), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;           \
        case 2: KERNEL<2><<<sp_grid(E, 2), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;           \
        case 4: KERNEL<4><<<sp_grid(E, 4), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;           \
        case 8: KERNEL<8><<<sp_grid(E, 8), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;           \
        case 16: KERNEL<16><<<sp_grid(E, 16), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;        \
        default: KERNEL<32><<<sp_grid(E, 32), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__); break;        \
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
    const int G = sp_group(C, d_feature, d_out);
    if (G == 0) {
        sp_trilinear_forward_kernel<<<sp_grid(E), 256, 0, (cudaStream_t)stream>>>(d_feature, d_corner_lut, d_worldcoord, d_out, p);
    } else {
        SP_DISPATCH_G(sp_trilinear_forward_vec4_kernel, G, d_feature, d_corner_lut, d_worldcoord, d_out, p)
    }
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
    const int G = sp_group(C, d_out_grad, d_feature_grad);
    if (G == 0) {
        sp_trilinear_backward_kernel<<<sp_grid(E), 256, 0, (cudaStream_t)stream>>>(d_out_grad, d_corner_lut, d_worldcoord, d_feature_grad, p);
    } else {
        SP_DISPATCH_G(sp_trilinear_backward_vec4_kernel, G, d_out_grad, d_corner_lut, d_worldcoord, d_feature_grad, p)
    }
    SDB_CHECK_LAUNCH();
    return SDB_OK;
}
