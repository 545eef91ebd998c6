// a8 (training side): the style modulation of the five ModLinear layers of LightningMLP folded into plain weights for ONE
// style code, forward and backward, in a handful of launches.
// Behavioural contract: imaginaire/model_utils/layers.py:241-271 (ModLinear with bias=False, mod_bias=True, output_mode=True) as
// LightningMLP uses it (layers.py:92-126):   alpha = weight_alpha z + bias_alpha  [I]     beta = weight_beta z + bias_beta  [O]
//                                            W'[o][i] = W[o][i] * alpha[i]                 (the layer's bias is beta)
// The fused renderer consumes W' and beta (sdb_pack_mlp) and returns dL/dW', dL/dbeta; this file carries them on to the raw
// parameters and to z:   dW = dW' * alpha     dalpha[i] = sum_o dW'[o][i] W[o][i]     d weight_alpha = dalpha (x) z
//                        d bias_alpha = dalpha     d weight_beta = dbeta (x) z     d bias_beta = dbeta
//                        dz = sum_layers weight_alpha^T dalpha + weight_beta^T dbeta
// Under torch autograd the same algebra is ~110 launches of a few microseconds per training view (addmm / mul / cat and their
// backward nodes), which makes the backward host-bound on a slow host; here: 2 launches forward, 3 backward.  All float32;
// sums are sequential per output (deterministic), so results differ from cuBLAS's gemv by rounding only.
// HBM-bound, ~6.6 MB forward / ~13 MB backward for the 256-wide network: a few microseconds each.
#include "common.cuh"

namespace {

constexpr int kModLayers = 5;

struct ModPtrs {            // per layer: weight [O, I], weight_alpha [I, Cz], bias_alpha [I], weight_beta [O, Cz], bias_beta [O]
    const float *W[kModLayers], *wa[kModLayers], *ba[kModLayers], *wb[kModLayers], *bb[kModLayers];
};
struct ModGrads {
    float *dW[kModLayers], *dwa[kModLayers], *dba[kModLayers], *dwb[kModLayers], *dbb[kModLayers];
};

// one warp per coefficient: alpha[l][i] (rows 0..I-1 of a layer) or beta[l][o] (rows I..I+O-1)
__global__ void __launch_bounds__(256)
mod_coeff_kernel(const ModPtrs P, const float *__restrict__ z, int O, int I, int Cz, float *__restrict__ alpha,
                 float *__restrict__ beta)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per = I + O;
    if (w >= kModLayers * per) return;
    const int l = w / per, r = w - l * per;
    const bool is_alpha = r < I;
    const int row = is_alpha ? r : r - I;
    const float *m = (is_alpha ? P.wa[l] : P.wb[l]) + (size_t)row * Cz;
    float acc = 0.0f;
    for (int c = lane; c < Cz; c += 32) acc = fmaf(__ldg(m + c), __ldg(z + c), acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        if (is_alpha) alpha[l * I + row] = acc + __ldg(P.ba[l] + row);
        else beta[l * O + row] = acc + __ldg(P.bb[l] + row);
    }
}

__global__ void __launch_bounds__(256)
mod_scale_kernel(const ModPtrs P, const float *__restrict__ alpha, int O, int I, float *__restrict__ wh)
{
    const long long n = (long long)kModLayers * O * I;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(t / ((long long)O * I));
        const long long e = t - (long long)l * O * I;
        const int i = (int)(e % I);
        wh[t] = __ldg(P.W[l] + e) * alpha[l * I + i];
    }
}

// block = 32 columns x 8 row groups of one layer: dW = dW' * alpha, dalpha = column sums of dW' * W (fixed order: deterministic)
__global__ void __launch_bounds__(256)
mod_bwd_weight_kernel(const ModPtrs P, const ModGrads G, const float *__restrict__ alpha, const float *__restrict__ g_wh, int O, int I,
                      float *__restrict__ dalpha)
{
    __shared__ float part[8][33];
    const int tiles = (I + 31) / 32;
    const int l = blockIdx.x / tiles, i = (blockIdx.x % tiles) * 32 + (threadIdx.x & 31), rg = threadIdx.x >> 5;
    float acc = 0.0f;
    if (i < I) {
        const float a = alpha[l * I + i];
        const float *g = g_wh + (size_t)l * O * I, *Wl = P.W[l];
        float *dW = G.dW[l];
        for (int o = rg; o < O; o += 8) {
            const float gv = __ldg(g + (size_t)o * I + i);
            acc = fmaf(gv, __ldg(Wl + (size_t)o * I + i), acc);
            dW[(size_t)o * I + i] = gv * a;
        }
    }
    part[rg][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rg == 0 && i < I) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) s += part[k][threadIdx.x & 31];
        dalpha[l * I + i] = s;
        G.dba[l][i] = s;
    }
}

// outer products with z: d weight_alpha[l][i][:] = dalpha[l][i] z, d weight_beta[l][o][:] = dbeta[l][o] z; d bias_beta = dbeta
__global__ void __launch_bounds__(256)
mod_bwd_outer_kernel(const ModGrads G, const float *__restrict__ z, const float *__restrict__ dalpha, const float *__restrict__ g_bh,
                     int O, int I, int Cz)
{
    const int per = I + O;
    const long long n = (long long)kModLayers * per * Cz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % Cz);
        const long long rr = t / Cz;
        const int l = (int)(rr / per), r = (int)(rr - (long long)l * per);
        if (r < I) {
            G.dwa[l][(size_t)r * Cz + c] = dalpha[l * I + r] * __ldg(z + c);
        } else {
            const int o = r - I;
            const float gb = __ldg(g_bh + l * O + o);
            G.dwb[l][(size_t)o * Cz + c] = gb * __ldg(z + c);
            if (c == 0) G.dbb[l][o] = gb;
        }
    }
}

// dz[c] = sum over layers and rows of weight_alpha[l][i][c] dalpha[l][i] + weight_beta[l][o][c] dbeta[l][o]; block = 32 columns
__global__ void __launch_bounds__(256)
mod_bwd_z_kernel(const ModPtrs P, const float *__restrict__ dalpha, const float *__restrict__ g_bh, int O, int I, int Cz,
                 float *__restrict__ dz)
{
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rg = threadIdx.x >> 5;
    float acc = 0.0f;
    if (c < Cz) {
        for (int l = 0; l < kModLayers; l++) {
            for (int i = rg; i < I; i += 8) acc = fmaf(__ldg(P.wa[l] + (size_t)i * Cz + c), dalpha[l * I + i], acc);
            for (int o = rg; o < O; o += 8) acc = fmaf(__ldg(P.wb[l] + (size_t)o * Cz + c), __ldg(g_bh + l * O + o), acc);
        }
    }
    part[rg][threadIdx.x & 31] = acc;
    __syncthreads();
    if (rg == 0 && c < Cz) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++) s += part[k][threadIdx.x & 31];
        dz[c] = s;
    }
}

int fill(ModPtrs &P, const void *const *ptrs) {
    for (int k = 0; k < 5 * kModLayers; k++)
        if (ptrs[k] == nullptr) return SDB_EINVAL;
    for (int l = 0; l < kModLayers; l++) {
        P.W[l] = (const float *)ptrs[l];
        P.wa[l] = (const float *)ptrs[kModLayers + l];
        P.ba[l] = (const float *)ptrs[2 * kModLayers + l];
        P.wb[l] = (const float *)ptrs[3 * kModLayers + l];
        P.bb[l] = (const float *)ptrs[4 * kModLayers + l];
    }
    return SDB_OK;
}
}  // namespace

extern "C" int sdb_modulate_forward(const void *const d_params[25], const float *d_z, int32_t O, int32_t I, int32_t Cz,
                                    float *d_alpha, float *d_wh, float *d_bh, void *stream)
{
    if (!d_params || !d_z || !d_alpha || !d_wh || !d_bh || O < 1 || I < 1 || Cz < 1) return SDB_EINVAL;
    ModPtrs P;
    const int rc = fill(P, d_params);
    if (rc != SDB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int warps = kModLayers * (I + O);
    mod_coeff_kernel<<<sdb_div_up(warps, 8), 256, 0, st>>>(P, d_z, O, I, Cz, d_alpha, d_bh);
    SDB_CHECK_LAUNCH();
    const long long n = (long long)kModLayers * O * I;
    const long long blocks = (n + 255) / 256, cap = (long long)sdb_num_sms() * 8;
    mod_scale_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(P, d_alpha, O, I, d_wh);
    SDB_CHECK_LAUNCH();
    return SDB_OK;
}

extern "C" int sdb_modulate_backward(const void *const d_params[25], void *const d_grads[25], const float *d_z, const float *d_alpha,
                                     const float *d_g_wh, const float *d_g_bh, int32_t O, int32_t I, int32_t Cz, float *d_dalpha,
                                     float *d_dz, void *stream)
{
    if (!d_params || !d_grads || !d_z || !d_alpha || !d_g_wh || !d_g_bh || !d_dalpha || !d_dz || O < 1 || I < 1 || Cz < 1)
        return SDB_EINVAL;
    ModPtrs P;
    const int rc = fill(P, d_params);
    if (rc != SDB_OK) return rc;
    ModGrads G;
    for (int k = 0; k < 5 * kModLayers; k++)
        if (d_grads[k] == nullptr) return SDB_EINVAL;
    for (int l = 0; l < kModLayers; l++) {
        G.dW[l] = (float *)d_grads[l];
        G.dwa[l] = (float *)d_grads[kModLayers + l];
        G.dba[l] = (float *)d_grads[2 * kModLayers + l];
        G.dwb[l] = (float *)d_grads[3 * kModLayers + l];
        G.dbb[l] = (float *)d_grads[4 * kModLayers + l];
    }
    cudaStream_t st = (cudaStream_t)stream;
    mod_bwd_weight_kernel<<<kModLayers * sdb_div_up(I, 32), 256, 0, st>>>(P, G, d_alpha, d_g_wh, O, I, d_dalpha);
    SDB_CHECK_LAUNCH();
    const long long n = (long long)kModLayers * (I + O) * Cz;
    const long long blocks = (n + 255) / 256, cap = (long long)sdb_num_sms() * 8;
    mod_bwd_outer_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(G, d_z, d_dalpha, d_g_bh, O, I, Cz);
    SDB_CHECK_LAUNCH();
    mod_bwd_z_kernel<<<sdb_div_up(Cz, 32), 256, 0, st>>>(P, d_dalpha, d_g_bh, O, I, Cz, d_dz);
    SDB_CHECK_LAUNCH();
    return SDB_OK;
}
