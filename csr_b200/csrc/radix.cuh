// radix.cuh -- hand-written stable LSD radix-sort pass (8- or 9-bit digits) over int32 keys with
// an int32 payload and an optional value payload.  Used by the stable transpose /
// order_columns (transpose.cu) and by the SpMV slab-plan builder (spmv_slab.cu).
//
//   k_radix_hist     per-tile digit histogram, written digit-major [digit][tile]
//   exclusive_scan   over the flattened histogram -> global offset of every (digit, tile)
//   k_radix_scatter  stable rank inside the tile, tile sorted in shared memory, coalesced write-out
// Stability inside a tile: each warp walks 32 consecutive entries per step, equal digits
// are ranked by lane with __match_any_sync against per-warp digit counters, and an
// exclusive prefix over the warps of the CTA orders the warps.
// The digit width is picked per sort: 9 bits when that saves a pass (17-18 bit keys take two
// passes instead of three), else 8.
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

template <int DB>
static __global__ void __launch_bounds__(RS_BLOCK)
k_radix_hist(const int32_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ tile_hist, int64_t ntiles)
{
    constexpr int ND = 1 << DB;
    __shared__ uint32_t h[ND];
    for (int d = threadIdx.x; d < ND; d += RS_BLOCK)
        h[d] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_STEPS; k++) {
        // (warp-aggregating these with __match_any_sync was measured 2.5x slower than the plain atomics)
        const int64_t i = base + (int64_t)k * RS_BLOCK + threadIdx.x;
        if (i < n)
            atomicAdd(&h[(keys[i] >> shift) & (ND - 1)], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < ND; d += RS_BLOCK)
        tile_hist[(int64_t)d * ntiles + blockIdx.x] = h[d];
}

// global -> shared without a register stop (LDGSTS): the destination is the entry's sorted slot
template <int BYTES> __device__ __forceinline__ void cp_async_small(void *smem_dst, const void *gsrc)
{
    static_assert(BYTES == 4 || BYTES == 8, "cp.async.ca moves 4, 8 or 16 bytes");
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Staging buffer: the tile's (key, row) pairs and its values go through the same bytes one
// after the other, which keeps four CTAs resident per SM.
template <typename VT> struct RadixSmem {
    static constexpr size_t vbytes = std::is_same<VT, NoPayload>::value ? 0 : sizeof(VT);
    static constexpr size_t bytes = (size_t)RS_TILE * (vbytes > 8 ? vbytes : 8);
};

// Stable rank + scatter of one tile.  The tile is put in its sorted order in shared memory
// (digit-major, stable) and written out with consecutive threads on consecutive addresses:
// every digit's run of the tile is one contiguous burst in HBM instead of 4-8 byte pieces.
// (9-bit digits with a float64 payload need 74 registers: at 4 CTAs per SM -- 64 registers -- the kernel spilled
// 92 bytes per thread; 3 CTAs per SM are enough to cover its latency)
template <typename VT, int DB>
__global__ void __launch_bounds__(RS_BLOCK, (DB == 9 && sizeof(VT) == 8) ? 3 : 4)
k_radix_scatter(const int32_t *__restrict__ keys_in, const int32_t *__restrict__ rows_in, const VT *__restrict__ vals_in,
                int32_t *__restrict__ keys_out, int32_t *__restrict__ rows_out, VT *__restrict__ vals_out, int64_t n,
                int shift, const int64_t *__restrict__ tile_off, int64_t ntiles)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    constexpr int ND = 1 << DB;
    constexpr int DPT = ND / RS_BLOCK;  // digits per thread in the prefix step
    using DT = typename std::conditional<DB <= 8, uint8_t, uint16_t>::type;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    int32_t *s_keys = reinterpret_cast<int32_t *>(rs_raw);
    int32_t *s_rows = s_keys + RS_TILE;
    VT *s_vals = reinterpret_cast<VT *>(rs_raw);
    __shared__ uint16_t wc[RS_WARPS][ND];
    __shared__ int64_t delta[ND];  // global offset of the digit's run minus its start in the tile
    __shared__ uint16_t dstart[ND];
    __shared__ DT s_dig[RS_TILE];
    __shared__ int s_wt[33];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int d = tid; d < ND; d += RS_BLOCK) {
#pragma unroll
        for (int k = 0; k < RS_WARPS; k++)
            wc[k][d] = 0;
    }
    __syncthreads();

    const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
    const int64_t wbase = tbase + (int64_t)w * RS_WARP_ITEMS;
    int32_t key[RS_STEPS];
    int lp[RS_STEPS];  // rank among the warp's equal digits, later the position in the sorted tile
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < RS_STEPS; k++) {
        const int64_t i = wbase + k * 32 + lane;
        key[k] = i < n ? keys_in[i] : 0;
    }
#pragma unroll
    for (int k = 0; k < RS_STEPS; k++) {
        const int64_t i = wbase + k * 32 + lane;
        const bool valid = i < n;
        // invalid lanes get a unique pseudo-digit so they match nobody
        const unsigned d = valid ? (unsigned)((key[k] >> shift) & (ND - 1)) : (unsigned)ND + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & lt);
        const unsigned cur = valid ? wc[w][d] : 0u;
        __syncwarp();
        if (valid && before == 0)
            wc[w][d] = (uint16_t)(cur + __popc(peers));
        __syncwarp();
        lp[k] = (int)(cur + before);
    }
    __syncthreads();
    {
        // exclusive prefix over the warps for this thread's digits, then over the digits of the tile
        int tot[DPT], mine = 0;
#pragma unroll
        for (int q = 0; q < DPT; q++) {
            const int d = tid * DPT + q;
            int run = 0;
#pragma unroll
            for (int k = 0; k < RS_WARPS; k++) {
                const int c = wc[k][d];
                wc[k][d] = (uint16_t)run;
                run += c;
            }
            tot[q] = run;
            mine += run;
        }
        int all;
        int start = block_exclusive_scan<int>(mine, s_wt, all);
#pragma unroll
        for (int q = 0; q < DPT; q++) {
            const int d = tid * DPT + q;
            dstart[d] = (uint16_t)start;
            delta[d] = tile_off[(int64_t)d * ntiles + blockIdx.x] - start;
            start += tot[q];
        }
    }
    __syncthreads();
    // round 1: keys and rows in sorted order
#pragma unroll
    for (int k = 0; k < RS_STEPS; k++) {
        const int64_t i = wbase + k * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)((key[k] >> shift) & (ND - 1));
            lp[k] += (int)dstart[d] + (int)wc[w][d];
            cp_async_small<4>(&s_rows[lp[k]], &rows_in[i]);
            s_keys[lp[k]] = key[k];
            s_dig[lp[k]] = (DT)d;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    // element j of the sorted tile goes to delta[digit] + j
    const int cnt = (int)min((int64_t)RS_TILE, n - tbase);
    for (int j = tid; j < cnt; j += RS_BLOCK) {
        const int64_t pos = delta[s_dig[j]] + j;
        if (keys_out)
            keys_out[pos] = s_keys[j];  // not needed after the last pass
        rows_out[pos] = s_rows[j];
    }
    if constexpr (HASV) {
        // round 2: the values through the same buffer
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RS_STEPS; k++) {
            const int64_t i = wbase + k * 32 + lane;
            if (i < n)
                cp_async_small<(int)sizeof(VT)>(&s_vals[lp[k]], &vals_in[i]);
        }
        cp_async_wait_all();
        __syncthreads();
        for (int j = tid; j < cnt; j += RS_BLOCK)
            vals_out[delta[s_dig[j]] + j] = s_vals[j];
    }
}

// One stable pass: (kin, rin, vin) -> (kout, rout, vout) ordered by the DB-bit digit at `shift`.
// hist: uint32[ND*ntiles], offs: int64[ND*ntiles+1] scratch.
template <typename VT, int DB>
static int radix_pass(const int32_t *kin, const int32_t *rin, const VT *vin, int32_t *kout, int32_t *rout, VT *vout,
                      int64_t n, int shift, uint32_t *hist, int64_t *offs, cudaStream_t s)
{
    constexpr int ND = 1 << DB;
    const int64_t ntiles = div_up(n, RS_TILE);
    CSRK_LAUNCH((k_radix_hist<DB>), (unsigned)ntiles, RS_BLOCK, 0, s, kin, n, shift, hist, ntiles);
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<uint32_t>{hist}, ND * ntiles, offs, s)));
    auto k = k_radix_scatter<VT, DB>;
    static bool optin = false;  // per instantiation
    if (!optin) {
        CSRK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RadixSmem<VT>::bytes));
        optin = true;
    }
    CSRK_LAUNCH(k, (unsigned)ntiles, RS_BLOCK, RadixSmem<VT>::bytes, s, kin, rin, vin, kout, rout, vout, n, shift, offs,
                ntiles);
    return CSRK_OK;
}

template <typename VT, int DB>
static int radix_sort_passes(const int32_t *keys, const int32_t *rows, const VT *vals, int64_t n, int npass,
                             int32_t *out_rows, VT *out_vals, int32_t *out_keys, cudaStream_t s)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    constexpr int ND = 1 << DB;
    const int64_t ntiles = div_up(n, RS_TILE);
    DevBuf keysA, keysB, rowsT, valsT, hist, offs;
    if (npass > 1) {
        CSRK_TRY(keysA.alloc(sizeof(int32_t) * (size_t)n, s));
        if (npass > 2)
            CSRK_TRY(keysB.alloc(sizeof(int32_t) * (size_t)n, s));
        CSRK_TRY(rowsT.alloc(sizeof(int32_t) * (size_t)n, s));
        if (HASV)
            CSRK_TRY(valsT.alloc(sizeof(VT) * (size_t)n, s));
    }
    CSRK_TRY(hist.alloc(sizeof(uint32_t) * ND * (size_t)ntiles, s));
    CSRK_TRY(offs.alloc(sizeof(int64_t) * (ND * (size_t)ntiles + 1), s));
    CSRK_TRACE_MARK("  sort: buffers allocated", s);
    const int32_t *kin = keys;
    const int32_t *rin = rows;
    const VT *vin = vals;
    for (int pass = 0; pass < npass; pass++) {
        // alternate destinations so that the final pass writes the outputs
        const bool to_out = ((npass - 1 - pass) % 2) == 0;
        int32_t *kout = pass == npass - 1 ? out_keys : (pass % 2 == 0) ? keysA.as<int32_t>() : keysB.as<int32_t>();
        int32_t *rout = to_out ? out_rows : rowsT.as<int32_t>();
        VT *vout = HASV ? (to_out ? out_vals : valsT.as<VT>()) : nullptr;
        CSRK_TRY((radix_pass<VT, DB>(kin, rin, vin, kout, rout, vout, n, DB * pass, hist.as<uint32_t>(), offs.as<int64_t>(),
                                     s)));
        kin = kout;
        rin = rout;
        vin = vout;
        CSRK_TRACE_MARK("  sort: pass", s);
    }
    return CSRK_OK;
}

// Stable LSD sort of n (key, row, value) triples by the low `key_bits` bits of the key.
// Inputs are read-only; the sorted payloads land in out_rows / out_vals, the sorted keys in
// out_keys when that is not null.  Enqueues on s; temporaries come from the operation's workspace.
template <typename VT>
static int radix_sort_by_key(const int32_t *keys, const int32_t *rows, const VT *vals, int64_t n, int key_bits,
                             int32_t *out_rows, VT *out_vals, cudaStream_t s, int32_t *out_keys = nullptr)
{
    if (n <= 0)
        return CSRK_OK;
    const int np8 = key_bits <= 8 ? 1 : (key_bits + 7) / 8;
    const int np9 = key_bits <= 9 ? 1 : (key_bits + 8) / 9;
    const int64_t force = options().radix_bits.load();
    if (force == 9 || (force == 0 && np9 < np8))
        return radix_sort_passes<VT, 9>(keys, rows, vals, n, np9, out_rows, out_vals, out_keys, s);
    return radix_sort_passes<VT, 8>(keys, rows, vals, n, np8, out_rows, out_vals, out_keys, s);
}

}  // namespace csrk
