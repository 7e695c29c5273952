// scan.cuh -- hand-written device-wide exclusive prefix sum (reduce / scan-of-sums /
// apply).  Input comes through a loader functor so callers can scan row counts,
// keep-flags, histogram cells ... without materialising them; the output has
// n+1 entries (out[n] = total), which is exactly a rowptrs array.
//
// HBM traffic: input read twice + output written once; the middle kernel is a
// single CTA over n/4096 partials.
#pragma once

#include "common.cuh"

namespace csrk {

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <typename T> __device__ __forceinline__ T block_exclusive_scan(T v, T *warp_tot /*[32]*/, T &block_total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    if (lane == 31)
        warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T w = lane < nw ? warp_tot[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o)
                winc += t;
        }
        if (lane < nw)
            warp_tot[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31)
            warp_tot[32] = winc;  // block total (slot 32)
    }
    __syncthreads();
    block_total = warp_tot[32];
    T res = warp_tot[wid] + inc - v;
    __syncthreads();  // warp_tot reusable after return
    return res;
}

template <typename OutT, typename Loader>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_reduce(Loader load, int64_t n, OutT *partials)
{
    __shared__ OutT wt[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    OutT s = 0;
#pragma unroll 4
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + (int64_t)k * SCAN_BLOCK + threadIdx.x;
        if (i < n)
            s += (OutT)load(i);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0)
        wt[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        OutT t = 0;
        for (int w = 0; w < SCAN_BLOCK / 32; w++)
            t += wt[w];
        partials[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of the partials in place; partials[ntiles] = grand total
template <typename OutT> __global__ void __launch_bounds__(1024) k_scan_partials(OutT *partials, int64_t ntiles)
{
    __shared__ OutT wt[33];
    __shared__ OutT carry_s;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        OutT v = i < ntiles ? partials[i] : OutT(0);
        OutT tot;
        OutT ex = block_exclusive_scan(v, wt, tot);
        OutT c = carry_s;
        if (i < ntiles)
            partials[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0)
            carry_s = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        partials[ntiles] = carry_s;
}

template <typename OutT, typename Loader>
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(Loader load, int64_t n, const OutT *partials, OutT *out)
{
    __shared__ OutT wt[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    OutT v[SCAN_ITEMS];
    OutT s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = i < n ? (OutT)load(i) : OutT(0);
        s += v[k];
    }
    OutT tot;
    OutT ex = block_exclusive_scan(s, wt, tot) + partials[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n)
            out[i] = ex;
        ex += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
        out[n] = partials[gridDim.x];
}

template <typename OutT> __global__ void k_scan_empty(OutT *out) { out[0] = 0; }

// out[0..n] = exclusive scan of load(0..n-1); out[n] = total.  Enqueues on s.
template <typename OutT, typename Loader> int exclusive_scan(Loader load, int64_t n, OutT *out, cudaStream_t s)
{
    if (n <= 0) {
        CSRK_LAUNCH(k_scan_empty<OutT>, 1, 1, 0, s, out);
        return CSRK_OK;
    }
    const int64_t ntiles = div_up(n, SCAN_TILE);
    DevBuf partials;
    CSRK_TRY(partials.alloc(sizeof(OutT) * (size_t)(ntiles + 1), s));
    CSRK_LAUNCH((k_scan_reduce<OutT, Loader>), (unsigned)ntiles, SCAN_BLOCK, 0, s, load, n, partials.as<OutT>());
    CSRK_LAUNCH((k_scan_partials<OutT>), 1, 1024, 0, s, partials.as<OutT>(), ntiles);
    CSRK_LAUNCH((k_scan_apply<OutT, Loader>), (unsigned)ntiles, SCAN_BLOCK, 0, s, load, n, partials.as<OutT>(), out);
    return CSRK_OK;
}

template <typename T> struct ArrayLoader {
    const T *p;
    __device__ __forceinline__ T operator()(int64_t i) const { return p[i]; }
};

}  // namespace csrk
