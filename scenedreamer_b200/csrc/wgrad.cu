// Weight gradients of the per-pixel networks on the tcgen05 tensor cores (sm_100a):  dW[n_out, k_in] = sum over samples of
// dZ[sample, n_out] * A[sample, k_in]  -- the backward of nn.Linear / ModLinear (imaginaire/model_utils/layers.py:247-269,
// :92-126) and of SKYMLP's layers (generators/gancraft_base.py:150-169) under torch.autograd in the reference.
//
// The SAMPLES are the reduction dimension.  The forward / gradient-chain kernels leave their bf16 records as MMA-ready
// tiles (rf_common.cuh: rec_chunk):  [work item][8-column chunk][128 sample rows][8 columns]  = 16 B per (row, chunk),
// 2 KB per chunk, so one item of an array is ONE contiguous range that a single 1-D bulk copy (TMA engine) drops into shared
// memory, where it IS a canonical MN-major, no-swizzle tcgen05 operand with K = samples: 8 columns contiguous, the 8
// samples of a core matrix 16 B apart, K-adjacent core matrices +128 B (LBO), MN-adjacent ones +2048 B (SBO)
// (operand form verified by sdb_tc_selftest_mn).  No transpose, no zero-filled padding, no library GEMM.
//
// Work decomposition: the output of every layer is cut into M-tiles of 128 rows of n_out ("jobs"); a job's samples are
// split over several CTAs (interleaved items, so the CTAs of the two M-tiles of a layer stream the same A tiles at the same
// time and the second read hits L2).  One CTA = one job x one split: fp32 accumulators [128 x k_in] stay in TMEM for the
// whole reduction, then are added to the (zeroed) gradient with red.global.add.  A CTA runs 4 warps: lane 0 of warp 0
// issues the bulk copies (2-stage ring, 100 KB per stage), lane 0 of warp 1 the MMAs, all four read the accumulators out.
// Bound: HBM -- per sample and M-tile 800 B in, 17.8 MFLOP per 128-sample item against ~2.3 us of load time per SM.
#include "rf_common.cuh"
#include "wgrad.cuh"

namespace rf {

namespace {
constexpr int kWgStages = 2;
constexpr int kWgThreads = 128;
constexpr uint32_t kChunkBytes = kRows * 16;                  // 2048: one 8-column chunk of a 128-sample item
constexpr uint32_t kWgABytes = (kActCols / 8) * kChunkBytes;  // 69,632: the widest A tile (272 columns)
constexpr uint32_t kWgZBytes = 16 * kChunkBytes;              // 32,768: 128 columns of dZ
constexpr uint32_t kWgStageBytes = kWgABytes + kWgZBytes;
constexpr uint32_t kWgSmem = kWgStages * kWgStageBytes + 256;

struct WgParams {
    WgJob job[kWgMaxJobs];
    int n_jobs;
    const int32_t *n_live;       // device: live ray tiles of the recorded pass (nullptr: every item up to cap_items exists)
    int items_per_live;
    long long cap_items;
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const WgParams prm)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kWgStages * kWgStageBytes);      // full[2], empty[2], done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // which (job, split) is this CTA?
    int j = 0, split = blockIdx.x;
    while (j < prm.n_jobs && split >= prm.job[j].nsplit) { split -= prm.job[j].nsplit; j++; }
    if (j >= prm.n_jobs) return;
    const WgJob jb = prm.job[j];
    const long long n_items = prm.n_live ? (long long)__ldg(prm.n_live) * prm.items_per_live : prm.cap_items;
    const long long my_items = split < n_items ? (n_items - split + jb.nsplit - 1) / jb.nsplit : 0;

    // dZ chunks the job does not own stay zero for the whole kernel (M-tiles narrower than 128 rows: fc_out_c, fc_sigma)
    for (int s = 0; s < kWgStages; s++) {
        uint4 *z = reinterpret_cast<uint4 *>(smem + s * kWgStageBytes + kWgABytes);
        for (uint32_t i = tid; i < kWgZBytes / 16; i += kWgThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int s = 0; s < kWgStages; s++) { tc05::mbar_init(&bars[s], 1); tc05::mbar_init(&bars[kWgStages + s], 1); }
        tc05::mbar_init(&bars[2 * kWgStages], 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc(tmem_slot, 512);
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t a_bytes = (uint32_t)jb.a_chunks * kChunkBytes, z_bytes = (uint32_t)jb.z_chunks * kChunkBytes;

    if (my_items > 0) {
        // role lanes run their loops inside `if (lane == 0)`; the other 31 lanes of those warps park at __syncwarp (a hardware
        // barrier), they do not spin next to the working lane
        if (warp == 0) {
            if (lane == 0) {
                // ---- loader: one bulk copy per operand and item ----
                for (long long n = 0; n < my_items; n++) {
                    const int s = (int)(n % kWgStages);
                    if (n >= kWgStages) tc05::mbar_wait_backoff(&bars[kWgStages + s], (uint32_t)((n / kWgStages - 1) & 1), 32);
                    const long long item = split + n * jb.nsplit;
                    uint8_t *sA = smem + s * kWgStageBytes, *sZ = sA + kWgABytes;
                    tc05::mbar_arrive_expect_tx(&bars[s], a_bytes + z_bytes);
                    tc05::bulk_g2s(sA, reinterpret_cast<const uint8_t *>(jb.A) + (size_t)item * a_bytes, a_bytes, &bars[s]);
                    tc05::bulk_g2s(sZ, reinterpret_cast<const uint8_t *>(jb.Z) + ((size_t)item * jb.z_chunks_total + jb.z_chunk0) * kChunkBytes,
                                   z_bytes, &bars[s]);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            if (lane == 0) {
                // ---- MMA issuer: D[128 x N] += Zpart^T [128 x 128 samples] * A [128 samples x N], 16 samples per instruction ----
                const int N = jb.a_chunks * 8;
                const int n_main = N > 256 ? 256 : N;                   // k_in = 272 = 256 + 16: two instructions per 16 samples
                const uint32_t idesc_main = tc05::make_idesc(128, n_main, true) | (1u << 15) | (1u << 16);     // A and B MN-major
                const uint32_t idesc_tail = tc05::make_idesc(128, N - n_main > 0 ? N - n_main : 16, true) | (1u << 15) | (1u << 16);
                for (long long n = 0; n < my_items; n++) {
                    const int s = (int)(n % kWgStages);
                    tc05::mbar_wait(&bars[s], (uint32_t)((n / kWgStages) & 1));
                    tc05::fence_after_thread_sync();
                    const uint32_t sA = tc05::smem_u32(smem + s * kWgStageBytes), sZ = sA + kWgABytes;
#pragma unroll 1
                    for (int kk = 0; kk < kRows / 16; kk++) {
                        const uint32_t acc = (n > 0 || kk > 0) ? 1u : 0u;
                        const uint64_t dz = tc05::make_smem_desc(sZ + kk * 256, 128, kChunkBytes);
                        tc05::mma_f16_ss(tmem, dz, tc05::make_smem_desc(sA + kk * 256, 128, kChunkBytes), idesc_main, acc);
                        if (N > 256)
                            tc05::mma_f16_ss(tmem + 256, dz, tc05::make_smem_desc(sA + 32 * kChunkBytes + kk * 256, 128, kChunkBytes),
                                             idesc_tail, acc);
                    }
                    tc05::mma_commit(&bars[kWgStages + s]);            // the stage is free once these MMAs have read it
                }
                tc05::mma_commit(&bars[2 * kWgStages]);
            }
            __syncwarp();
        }
        // ---- accumulators -> gradient (thread = output row) ----
        tc05::mbar_wait_backoff(&bars[2 * kWgStages], 0, 128);
        __syncwarp();
        tc05::fence_after_thread_sync();
        const int N = jb.a_chunks * 8;
        const int row = warp * 32 + lane;
        float *dst = jb.out + (size_t)(jb.row0 + row) * jb.ld_out;
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc05::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tc05::tmem_ld_wait();
            if (row < jb.rows) {
#pragma unroll
                for (int q = 0; q < 16; q++) atomicAdd(dst + c0 + q, v[q]);
            }
        }
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc(tmem, 512);
}
}  // namespace

// Splits: CTAs are handed out in proportion to the bytes a job streams per item, one CTA per SM over the whole list.
int launch_wgrad(WgJob *jobs, int n_jobs, const int32_t *d_n_live, int items_per_live, long long cap_items, cudaStream_t st)
{
    if (n_jobs < 1 || n_jobs > kWgMaxJobs) return SDB_EINVAL;
    if (cap_items <= 0) return SDB_OK;
    const long long n_items = cap_items;
    WgParams prm{};
    prm.n_jobs = n_jobs;
    prm.n_live = d_n_live;
    prm.items_per_live = items_per_live;
    prm.cap_items = cap_items;
    double total = 0.0;
    for (int i = 0; i < n_jobs; i++) {
        const WgJob &jb = jobs[i];
        if (!jb.A || !jb.Z || !jb.out || jb.a_chunks < 1 || jb.a_chunks > kActCols / 8 || (jb.a_chunks * 8) % 16 != 0 ||
            jb.z_chunks < 1 || jb.z_chunks > 16 || jb.z_chunk0 + jb.z_chunks > jb.z_chunks_total || jb.rows < 1 || jb.rows > 128)
            return SDB_EINVAL;
        total += jb.a_chunks + jb.z_chunks;
    }
    const int sms = sdb_num_sms();
    int used = 0;
    for (int i = 0; i < n_jobs; i++) {
        int ns = (int)((double)sms * (jobs[i].a_chunks + jobs[i].z_chunks) / total);
        if (ns < 1) ns = 1;
        if ((long long)ns > n_items) ns = (int)n_items;
        jobs[i].nsplit = ns;
        used += ns;
        prm.job[i] = jobs[i];
    }
    SDB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem));
    wgrad_kernel<<<used, kWgThreads, kWgSmem, st>>>(prm);
    SDB_CHECK_LAUNCH();
    return SDB_OK;
}

}  // namespace rf
