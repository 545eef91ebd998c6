// Jobs of the tensor-core weight-gradient kernel (wgrad.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rf {

constexpr int kWgMaxJobs = 16;

// One output M-tile: out[row0 + r][0 .. 8*a_chunks) += sum_samples Z[sample][8*z_chunk0 + r] * A[sample][.], r < rows.
// A and Z are tiled bf16 records ([item][chunk][128 rows][8]); a_chunks * 8 = k_in (a multiple of 16, <= 272),
// z_chunks <= 16 chunks of the z_chunks_total chunks of Z's items.  nsplit is filled in by launch_wgrad.
struct WgJob {
    const uint16_t *A;
    const uint16_t *Z;
    float *out;
    int a_chunks;
    int z_chunks_total, z_chunk0, z_chunks;
    int ld_out, row0, rows;
    int nsplit;
};

// dW += over the first n_items work items (128 samples each); `out` buffers must have been zeroed on the stream.
// n_items = *d_n_live * items_per_live read ON THE DEVICE (no host round trip), or cap_items when d_n_live is null;
// cap_items (the record's capacity) only sizes the split of the jobs over the CTAs.
int launch_wgrad(WgJob *jobs, int n_jobs, const int32_t *d_n_live, int items_per_live, long long cap_items, cudaStream_t st);

}  // namespace rf
