// radix.cuh -- hand-written stable LSD radix-sort pass (8-bit digits) over int32 keys with
// an int32 payload and an optional value payload.  Used by the stable transpose /
// order_columns (transpose.cu) and by the SpMV slab-format builder (spmv_psf.cu).
//
//   k_radix_hist     per-tile digit histogram, written digit-major [digit][tile]
//   exclusive_scan   over the flattened histogram -> global offset of every (digit, tile)
//   k_radix_scatter  stable rank inside the tile + scatter
// Stability inside a tile: each warp walks 32 consecutive entries per step, equal digits
// are ranked by lane with __match_any_sync against per-warp digit counters, and an
// exclusive prefix over the warps of the CTA orders the warps.
#pragma once

#include <type_traits>

#include "common.cuh"
#include "scan.cuh"

namespace csrk {

constexpr int RS_BLOCK = 256;
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int RS_STEPS = 16;
constexpr int RS_TILE = RS_BLOCK * RS_STEPS;  // 4096 entries per CTA
constexpr int RS_WARP_ITEMS = 32 * RS_STEPS;

struct NoPayload {};

static __global__ void __launch_bounds__(RS_BLOCK)
k_radix_hist(const int32_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ tile_hist, int64_t ntiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_STEPS; k++) {
        int64_t i = base + (int64_t)k * RS_BLOCK + threadIdx.x;
        if (i < n)
            atomicAdd(&h[(keys[i] >> shift) & 255], 1u);
    }
    __syncthreads();
    tile_hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <typename VT>
__global__ void __launch_bounds__(RS_BLOCK)
k_radix_scatter(const int32_t *__restrict__ keys_in, const int32_t *__restrict__ rows_in, const VT *__restrict__ vals_in,
                int32_t *__restrict__ keys_out, int32_t *__restrict__ rows_out, VT *__restrict__ vals_out, int64_t n,
                int shift, const int64_t *__restrict__ tile_off, int64_t ntiles)
{
    __shared__ uint32_t wc[RS_WARPS][256];
    __shared__ int64_t toff[256];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++)
        wc[k][tid] = 0;
    toff[tid] = tile_off[(int64_t)tid * ntiles + blockIdx.x];
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * RS_WARP_ITEMS;
    int32_t key[RS_STEPS];
    uint32_t rank[RS_STEPS];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < RS_STEPS; k++) {
        const int64_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? keys_in[i] : 0;
        // invalid lanes get a unique pseudo-digit so they match nobody
        const unsigned d = valid ? (unsigned)((key[k] >> shift) & 255) : 256u + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & lt);
        const uint32_t cur = valid ? wc[w][d] : 0u;
        __syncwarp();
        if (valid && before == 0)
            wc[w][d] = cur + __popc(peers);
        __syncwarp();
        rank[k] = cur + before;
    }
    __syncthreads();
    {
        // exclusive prefix over the warps for digit `tid`
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < RS_WARPS; k++) {
            uint32_t c = wc[k][tid];
            wc[k][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_STEPS; k++) {
        const int64_t i = wbase + k * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)((key[k] >> shift) & 255);
            const int64_t pos = toff[d] + wc[w][d] + rank[k];
            keys_out[pos] = key[k];
            rows_out[pos] = rows_in[i];
            if constexpr (!std::is_same<VT, NoPayload>::value)
                vals_out[pos] = vals_in[i];
        }
    }
}

// One stable pass: (kin, rin, vin) -> (kout, rout, vout) ordered by digit `shift/8` of the key.
// hist: uint32[256*ntiles], offs: int64[256*ntiles+1] scratch.
template <typename VT>
static int radix_pass(const int32_t *kin, const int32_t *rin, const VT *vin, int32_t *kout, int32_t *rout, VT *vout,
                      int64_t n, int shift, uint32_t *hist, int64_t *offs, cudaStream_t s)
{
    const int64_t ntiles = div_up(n, RS_TILE);
    CSRK_LAUNCH(k_radix_hist, (unsigned)ntiles, RS_BLOCK, 0, s, kin, n, shift, hist, ntiles);
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<uint32_t>{hist}, 256 * ntiles, offs, s)));
    CSRK_LAUNCH((k_radix_scatter<VT>), (unsigned)ntiles, RS_BLOCK, 0, s, kin, rin, vin, kout, rout, vout, n, shift, offs,
                ntiles);
    return CSRK_OK;
}

// Stable LSD sort of n (key, row, value) triples by the low `key_bits` bits of the key.
// Inputs are read-only; the sorted payloads land in out_rows / out_vals (the sorted keys
// are not kept).  Enqueues on s; temporaries are stream-ordered.
template <typename VT>
static int radix_sort_by_key(const int32_t *keys, const int32_t *rows, const VT *vals, int64_t n, int key_bits,
                             int32_t *out_rows, VT *out_vals, cudaStream_t s)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    if (n <= 0)
        return CSRK_OK;
    const int npass = key_bits <= 8 ? 1 : (key_bits + 7) / 8;
    const int64_t ntiles = div_up(n, RS_TILE);
    DevBuf keysA, keysB, rowsT, valsT, hist, offs;
    CSRK_TRY(keysA.alloc(sizeof(int32_t) * (size_t)n, s));
    if (npass > 1) {
        CSRK_TRY(keysB.alloc(sizeof(int32_t) * (size_t)n, s));
        CSRK_TRY(rowsT.alloc(sizeof(int32_t) * (size_t)n, s));
        if (HASV)
            CSRK_TRY(valsT.alloc(sizeof(VT) * (size_t)n, s));
    }
    CSRK_TRY(hist.alloc(sizeof(uint32_t) * 256 * (size_t)ntiles, s));
    CSRK_TRY(offs.alloc(sizeof(int64_t) * (256 * (size_t)ntiles + 1), s));
    CSRK_TRACE_MARK("  sort: buffers allocated", s);
    const int32_t *kin = keys;
    const int32_t *rin = rows;
    const VT *vin = vals;
    for (int pass = 0; pass < npass; pass++) {
        // alternate destinations so that the final pass writes the outputs
        const bool to_out = ((npass - 1 - pass) % 2) == 0;
        int32_t *kout = (pass % 2 == 0) ? keysA.as<int32_t>() : keysB.as<int32_t>();
        int32_t *rout = to_out ? out_rows : rowsT.as<int32_t>();
        VT *vout = HASV ? (to_out ? out_vals : valsT.as<VT>()) : nullptr;
        CSRK_TRY((radix_pass<VT>(kin, rin, vin, kout, rout, vout, n, 8 * pass, hist.as<uint32_t>(), offs.as<int64_t>(), s)));
        kin = kout;
        rin = rout;
        vin = vout;
        CSRK_TRACE_MARK("  sort: pass", s);
    }
    return CSRK_OK;
}

}  // namespace csrk
