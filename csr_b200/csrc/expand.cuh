// expand.cuh -- two small building blocks shared by the stable transpose (transpose.cu) and the
// SpMV slab-plan builder (spmv_slab.cu): the owning row of every nnz position, and segment
// bounds read off a sorted key array.
#pragma once

#include "common.cuh"

namespace csrk {

// rows[i] = the row that owns nnz position i (last r with rp[r] <= i).
// One CTA per tile of EXP_TILE consecutive positions: the rows that start inside the tile mark
// their first position in shared memory (empty rows collide on one position; the largest wins,
// which is the owner), a running maximum over the tile fills the gaps.  Two binary searches per
// tile instead of one per entry.
constexpr int EXP_TILE = 4096;
template <typename RPT>
__global__ void __launch_bounds__(256)
k_expand_rows(const RPT *__restrict__ rp, int32_t nrows, int64_t nnz, int32_t *__restrict__ rows)
{
    constexpr int PER = EXP_TILE / 256;
    __shared__ int s_mark[EXP_TILE];
    __shared__ int s_wmax[8];
    __shared__ int64_t s_r[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t tbase = (int64_t)blockIdx.x * EXP_TILE;
    const int cnt = (int)min((int64_t)EXP_TILE, nnz - tbase);
    for (int j = tid; j < EXP_TILE; j += 256)
        s_mark[j] = -1;
    if (tid < 2) {
        const int64_t pos = tid == 0 ? tbase : tbase + cnt - 1;
        s_r[tid] = lower_bound_rp(rp, 0, (int64_t)nrows + 1, pos + 1) - 1;
    }
    __syncthreads();
    const int64_t r0 = s_r[0], r1 = s_r[1];
    if (tid == 0)
        s_mark[0] = (int)r0;
    for (int64_t r = r0 + 1 + tid; r <= r1; r += 256)
        atomicMax(&s_mark[(int)((int64_t)rp[r] - tbase)], (int)r);  // in (0, cnt-1] by the choice of r0, r1
    __syncthreads();
    // running maximum: PER consecutive marks per thread, then across the threads
    int v[PER];
    int run = -1;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        run = max(run, s_mark[tid * PER + k]);
        v[k] = run;
    }
    int inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d)
            inc = max(inc, o);
    }
    if (lane == 31)
        s_wmax[w] = inc;
    __syncthreads();
    int before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0)
        before = -1;
    for (int k = 0; k < w; k++)
        before = max(before, s_wmax[k]);
#pragma unroll
    for (int k = 0; k < PER; k++)
        s_mark[tid * PER + k] = max(v[k], before);
    __syncthreads();
    for (int j = tid; j < cnt; j += 256)
        rows[tbase + j] = s_mark[j];
}

// Output rowptrs from the column-sorted keys: rp[c] = first position whose key is >= c.
// Position i writes rp for every column in (keys[i-1], keys[i]] -- each column exactly once.
// Position nnz closes the tail with keys[nnz] := ncols.  Four positions per thread (one 16-byte
// load; the workspace hands out 256-byte aligned blocks).
template <typename RPT>
__global__ void k_key_bounds(const int32_t *__restrict__ keys, int64_t nnz, int32_t ncols, RPT *__restrict__ rp)
{
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 > nnz)
        return;
    int32_t k[5];
    k[0] = i0 ? keys[i0 - 1] : -1;
    if (i0 + 4 <= nnz) {
        const int4 q = *reinterpret_cast<const int4 *>(keys + i0);
        k[1] = q.x, k[2] = q.y, k[3] = q.z, k[4] = q.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            k[j + 1] = i0 + j < nnz ? keys[i0 + j] : ncols;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (i0 + j > nnz)
            break;
        for (int32_t c = k[j] + 1; c <= k[j + 1]; c++)
            rp[c] = (RPT)(i0 + j);
    }
}

}  // namespace csrk
