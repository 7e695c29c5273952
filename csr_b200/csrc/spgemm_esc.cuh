// spgemm_esc.cuh -- expand / sort / compress path of mult_ab for WIDE results (included by spgemm.cu).
//
// The row-wise accumulators of spgemm.cu are shared-memory hash tables (rows of up to 8192 distinct columns) and
// dense column windows (any row, when ncols fits a few windows of shared memory).  A wide product -- configs[3]:
// 5M columns, rows of 10^3 .. 10^7 products that almost never collide -- has rows beyond the hash capacity and
// 188 windows per row: the round-1 fallback (a dense accumulator + bitmap per CTA in global scratch, swept in
// full for every row) ran at 6 G products/s.  Here such rows are cut by COLUMN RANGE into pseudo-rows of about
// ESC_TARGET products, the products are written out once, grouped by pseudo-row, and every pseudo-row is then a
// small, contiguous, hash-sized problem:
//   1. k_esc_count    products per (row, column range): range = col * R_i / ncols, R_i = ceil(P_i / target)
//   2. scan           offsets of the pseudo-rows in the expansion (int64)
//   3. k_esc_scatter  (column, a*b) of every product into its pseudo-row's segment
//   4. k_esc_sortmerge per pseudo-row: bucket sort by column in shared memory, equal columns summed, the sorted
//                     distinct (column, sum) pairs written back IN PLACE at the head of the segment -> exact nnz
//                     (k_esc_reduce, a hash accumulator, takes the pseudo-rows whose columns pile up)
//   5. after the rowptr scan: k_esc_emit copies the segments into C.
// One phase: the symbolic count of these rows is a by-product of the numeric work (the reference's two passes,
// multiply.py:60-129, give the same rowptrs: distinct columns per row).  The walk over A and B happens in small
// CTAs (many per SM) whose lanes fetch the extents of 32 B rows at once, so the dependent-load chain
// A.colind -> B.rowptr -> B.colind is hidden by occupancy rather than paid per row by one big CTA.
// Rows whose products are so skewed that a pseudo-row exceeds the hash capacity are handed back to the caller
// (old path); if the expansion does not fit in memory the whole path declines.
#pragma once

namespace csrk {

constexpr int ESC_CAP = 8192;          // products per pseudo-row the largest reduce kernel takes
constexpr int ESC_ITEM = 128;          // A entries per work item of the count / scatter kernels
constexpr int ESC_WALK_THREADS = 128;

struct EscState {
    bool active = false;
    int n_rows = 0;                 // rows on this path (bad ones included)
    const int32_t *rows = nullptr;  // their ids
    int64_t np = 0;                 // pseudo-rows
    int64_t n_exp = 0;              // expanded products
    DevBuf ecol, eval;              // expansion, then the compacted sorted results (pool allocations: returned after the call)
    DevBuf pbase;                   // int[n_rows+1]: first pseudo-row of each row
    DevBuf prow;                    // int[np]: row index (into rows[]) of each pseudo-row
    DevBuf poff;                    // int64[np+1]: segment offsets
    DevBuf pnnz;                    // int[np]: distinct columns
    DevBuf pg;                      // int64[np+1]: exclusive scan of pnnz
    DevBuf bad;                     // int[n_rows]: row handed back
    DevBuf old_list;                // int32[n_rows]: ids of the rows handed back
    int n_old = 0;
};

// ranges and work items per row
// A row with hundreds of times more products than the result has columns is dense many times over: expanding it
// would write hundreds of products per output entry.  Such rows are handed back at once (bad = 1) to the dense
// accumulators of spgemm.cu.  (A factor of 4 was tried at configs[4] -- rows of 10^7 entries, up to 100 products per
// output entry -- and lost: the global-scratch accumulator works one row per CTA, 1.5 s for a few dozen rows, while
// their pseudo-rows -- some twenty columns hit 1536 times -- cost the hash kernel 60 ms in all.)
constexpr int ESC_DENSE_FACTOR = 256;

__global__ void __launch_bounds__(256) k_esc_rows(MatView A, const int32_t *__restrict__ rows, int n, const int64_t *__restrict__ prod,
                                                   int64_t target, int64_t ncols_out, int *__restrict__ nrange,
                                                   int *__restrict__ nitem, int *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int32_t row = rows[i];
    const int64_t len = ld_rp(A.rp, A.rp64, (int64_t)row + 1) - ld_rp(A.rp, A.rp64, row);
    if (prod[row] > ESC_DENSE_FACTOR * ncols_out) {
        nrange[i] = 0;
        nitem[i] = 0;
        bad[i] = 1;
        return;
    }
    nrange[i] = (int)((prod[row] + target - 1) / target);
    nitem[i] = (int)((len + ESC_ITEM - 1) / ESC_ITEM);
}

__global__ void __launch_bounds__(256) k_esc_prow(const int *__restrict__ pbase, int n, int32_t *__restrict__ prow)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n)
        return;
    for (int p = pbase[i] + lane; p < pbase[i + 1]; p += 32)
        prow[p] = i;
}

// The walk of one work item (ESC_ITEM consecutive A entries of one row): the item's entries are dealt evenly to
// the CTA's warps; a warp loads the B-row extents of 32 entries at once and then streams those B rows in chunks of
// 32*U entries (lane l takes entries l, l+32, ..).  ld(act, kk) fetches a lane's payload, use(act[], av, pay[])
// consumes a chunk -- both are called by all 32 lanes together -- and the payload of the NEXT chunk is requested
// before the current one is consumed, so a warp always has loads in flight while it goes through its shuffles
// and atomics.
template <int U, typename Pay, typename LD, typename USE>
__device__ __forceinline__ void esc_walk(const MatView &A, const MatView &B, int64_t as, int64_t ae, LD &&ld, USE &&use)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t per = (ae - as + nw - 1) / nw;
    const int64_t s0 = as + (int64_t)w * per, s1 = min(ae, s0 + per);
    for (int64_t base = s0; base < s1; base += 32) {
        const int64_t jj = base + lane;
        int64_t bs = 0, be = 0;
        double av = 0.0;
        if (jj < s1) {
            const int32_t j = A.ci[jj];
            av = ld_val(A.vs, A.vk, jj);
            bs = ld_rp(B.rp, B.rp64, j);
            be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
        }
        const int ne = (int)min((int64_t)32, s1 - base);
        // chunk cursor (warp-uniform): entry e, chunk start k0 inside [bs_e, be_e)
        int e = -1;
        int64_t k0 = 0, be_e = 0;
        double av_e = 0.0;
        auto advance = [&]() -> bool {
            if (e >= 0 && k0 + 32 * U < be_e) {
                k0 += 32 * U;
                return true;
            }
            while (++e < ne) {
                k0 = __shfl_sync(0xffffffffu, bs, e);
                be_e = __shfl_sync(0xffffffffu, be, e);
                av_e = __shfl_sync(0xffffffffu, av, e);
                if (k0 < be_e)
                    return true;
            }
            return false;
        };
        if (!advance())
            continue;
        bool act[U];
        Pay pay[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            act[u] = k0 + lane + 32 * u < be_e;
            pay[u] = ld(act[u], k0 + lane + 32 * u);
        }
        double cav = av_e;
        while (true) {
            const bool have = advance();
            bool nact[U];
            Pay npay[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                nact[u] = have && (k0 + lane + 32 * u < be_e);
                npay[u] = ld(nact[u], k0 + lane + 32 * u);
            }
            use(act, cav, pay);
            if (!have)
                break;
#pragma unroll
            for (int u = 0; u < U; u++) {
                act[u] = nact[u];
                pay[u] = npay[u];
            }
            cav = av_e;
        }
    }
}

// Lanes with equal range index are contiguous (B rows are sorted; if they are not, the runs are just shorter):
// the first lane of every run adds the run length to the pseudo-row's counter.  esc_claim_issue returns the
// counter's old value in the run's first lane (the atomic is in flight: nothing waits for it here) and that
// lane's index in `hl`; esc_claim_slot turns both into the slot of this lane's product.
__device__ __forceinline__ unsigned esc_claim_issue(bool act, unsigned r, unsigned *__restrict__ counters, int &hl)
{
    const int lane = threadIdx.x & 31;
    const unsigned actm = __ballot_sync(0xffffffffu, act);
    const unsigned prev = __shfl_up_sync(0xffffffffu, r, 1);
    const bool head = act && (lane == 0 || r != prev);
    const unsigned h = __ballot_sync(0xffffffffu, head);
    const unsigned le = lanemask_lt() | (1u << lane);
    hl = (31 - __clz((int)(h & le))) & 31;   // my run's first lane (garbage when !act)
    unsigned old = 0;
    if (head) {
        const unsigned next = h & ~le;
        const int nl = next ? __ffs((int)next) - 1 : __popc(actm);   // active lanes are lanes 0 .. popc-1
        old = atomicAdd(&counters[r], (unsigned)(nl - lane));
    }
    return old;
}
__device__ __forceinline__ unsigned esc_claim_slot(unsigned old, int hl)
{
    return __shfl_sync(0xffffffffu, old, hl) + (unsigned)((int)(threadIdx.x & 31) - hl);
}

struct EscItem {
    int idx;          // row index into rows[]
    int64_t as, ae;   // A entries of this item
};
__device__ __forceinline__ EscItem esc_item(const MatView &A, const int32_t *__restrict__ rows, int n,
                                            const int *__restrict__ item_off, int *s_idx, unsigned stride)
{
    // stride > 1 (count kernel): consecutive CTAs take items `stride` apart (coprime to the item count), so that the
    // items of one heavy row, which all add to the same pseudo-row counters, are spread over the kernel's lifetime
    // (21 -> 17 ms at configs[3]).  The scatter kernel keeps them together (stride 1): their partial-sector writes
    // into the same segments then merge in L2 (41 ms; 68 ms when spread)
    const int item = (int)(((uint64_t)blockIdx.x * stride) % gridDim.x);
    if (threadIdx.x == 0) {
        int lo = 0, hi = n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (item_off[mid] <= item)
                lo = mid;
            else
                hi = mid;
        }
        *s_idx = lo;
    }
    __syncthreads();
    EscItem it;
    it.idx = *s_idx;
    const int32_t row = rows[it.idx];
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    it.as = as + (int64_t)(item - item_off[it.idx]) * ESC_ITEM;
    it.ae = min(ae, it.as + ESC_ITEM);
    return it;
}

__global__ void __launch_bounds__(ESC_WALK_THREADS)
k_esc_count(MatView A, MatView B, const int32_t *__restrict__ rows, int n, const int *__restrict__ item_off,
            const int *__restrict__ pbase, unsigned *__restrict__ pcount, unsigned stride)
{
    __shared__ int s_idx;
    const EscItem it = esc_item(A, rows, n, item_off, &s_idx, stride);
    const int pb = pbase[it.idx];
    const uint64_t scale = ((uint64_t)(pbase[it.idx + 1] - pb) << 32) / (uint64_t)B.ncols;
    unsigned *cnt = pcount + pb;
    esc_walk<4, int32_t>(
        A, B, it.as, it.ae, [&](bool act, int64_t kk) -> int32_t { return act ? ld_stream_i32(B.ci + kk) : 0; },
        [&](const bool (&act)[4], double, const int32_t (&k)[4]) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int hl;
                const unsigned r = act[u] ? (unsigned)(((uint64_t)k[u] * scale) >> 32) : 0xffffffffu;
                (void)esc_claim_issue(act[u], r, cnt, hl);
            }
        });
}

// Pseudo-rows above the capacity.  If the row's column ranges are at most ESC_CAP columns wide, a pseudo-row has at
// most ESC_CAP DISTINCT columns however many products pile up on them (A*A^T on power-law data: a popular column
// collects a product from almost every entry of a heavy row): the hash kernel takes it as it is (fail = 1).
// Otherwise the row goes back to the caller.
__global__ void __launch_bounds__(256) k_esc_check(const unsigned *__restrict__ pcount, const int32_t *__restrict__ prow, int64_t np,
                                                    const int *__restrict__ pbase, int64_t ncols_out, int *__restrict__ bad,
                                                    int *__restrict__ fail, int *__restrict__ any_fail)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < np && pcount[p] > (unsigned)ESC_CAP) {
        const int i = prow[p];
        const int64_t R = pbase[i + 1] - pbase[i];
        if (ncols_out / R + 2 <= ESC_CAP) {
            fail[p] = 1;
            *any_fail = 1;
        } else
            bad[i] = 1;
    }
}
__global__ void __launch_bounds__(256) k_esc_mask(unsigned *__restrict__ pcount, const int32_t *__restrict__ prow, int64_t np,
                                                   const int *__restrict__ bad)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < np && bad[prow[p]])
        pcount[p] = 0;
}
__global__ void __launch_bounds__(256) k_esc_badlist(const int32_t *__restrict__ rows, int n, const int *__restrict__ bad,
                                                      int32_t *__restrict__ old_list, int *__restrict__ n_old)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && bad[i])
        old_list[atomicAdd(n_old, 1)] = rows[i];
}

__global__ void __launch_bounds__(ESC_WALK_THREADS)
k_esc_scatter(MatView A, MatView B, const int32_t *__restrict__ rows, int n, const int *__restrict__ item_off,
              const int *__restrict__ pbase, const int *__restrict__ bad, const int64_t *__restrict__ poff,
              unsigned *__restrict__ pcur, int32_t *__restrict__ ecol, double *__restrict__ eval, int both_f32,
              unsigned stride)
{
    __shared__ int s_idx;
    const EscItem it = esc_item(A, rows, n, item_off, &s_idx, stride);
    if (bad[it.idx])
        return;
    const int pb = pbase[it.idx];
    const uint64_t scale = ((uint64_t)(pbase[it.idx + 1] - pb) << 32) / (uint64_t)B.ncols;
    unsigned *cur = pcur + pb;
    const int64_t *off = poff + pb;
    struct Pay {
        int32_t k;
        double v;
    };
    constexpr int U = 2;
    esc_walk<U, Pay>(
        A, B, it.as, it.ae,
        [&](bool act, int64_t kk) -> Pay {
            Pay q{0, 0.0};
            if (act) {
                q.k = ld_stream_i32(B.ci + kk);
                q.v = ld_val(B.vs, B.vk, kk);
            }
            return q;
        },
        [&](const bool (&act)[U], double av, const Pay (&q)[U]) {
            unsigned r[U], old[U];
            int hl[U];
#pragma unroll
            for (int u = 0; u < U; u++) {   // all the chunk's atomics go out before anything waits for one
                r[u] = act[u] ? (unsigned)(((uint64_t)q[u].k * scale) >> 32) : 0xffffffffu;
                old[u] = esc_claim_issue(act[u], r[u], cur, hl[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned slot = esc_claim_slot(old[u], hl[u]);
                if (act[u]) {
                    const int64_t o = off[r[u]] + slot;
                    ecol[o] = q[u].k;
                    eval[o] = product(av, q[u].v, both_f32);
                }
            }
        });
}

// One warp per pseudo-row of at most SLOTS/2 products (the tiny ones); columns ranked by counting.
template <int SLOTS>
__global__ void __launch_bounds__(256) k_esc_reduce_warp(const int32_t *__restrict__ plist, int nbin, const int64_t *__restrict__ poff,
                                                         int32_t *ecol, double *eval, int32_t *__restrict__ pnnz)
{
    __shared__ int32_t keys[8][SLOTS];
    __shared__ double vals[8][SLOTS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int idx = blockIdx.x * 8 + w;
    for (int i = lane; i < SLOTS; i += 32) {
        keys[w][i] = EMPTY_KEY;
        vals[w][i] = 0.0;
    }
    __syncwarp();
    if (idx >= nbin)
        return;
    const int32_t p = plist[idx];
    const int64_t off = poff[p];
    const int n = (int)(poff[p + 1] - off);
    int count = 0;
    for (int i = lane; i < n; i += 32) {
        const int32_t k = ecol[off + i];
        const double v = eval[off + i];
        unsigned h = hash_col(k, SLOTS - 1);
        while (true) {
            const int32_t old = atomicCAS(&keys[w][h], EMPTY_KEY, k);
            if (old == EMPTY_KEY || old == k) {
                count += old == EMPTY_KEY;
                atomicAdd(&vals[w][h], v);
                break;
            }
            h = (h + 1) & (SLOTS - 1);
        }
    }
    __syncwarp();   // every product has been read: the segment's head can be overwritten
    for (int i = lane; i < SLOTS; i += 32) {
        const int32_t k = keys[w][i];
        if (k != EMPTY_KEY) {
            int rank = 0;
            for (int j = 0; j < SLOTS; j++)
                rank += keys[w][j] < k;
            ecol[off + rank] = k;
            eval[off + rank] = vals[w][i];
        }
    }
    count = warp_sum(count);
    if (lane == 0)
        pnnz[p] = count;
}

// Pseudo-rows of at most THREADS*PER products: bucket sort + merge, everything in shared memory and on NATIVE
// 32-bit atomics (the hash accumulators of spgemm.cu pay a CAS for the key and a CAS loop for the float64 add).
//   a. products -> registers; [kmin, kmax] of the columns
//   b. bucket = (k - kmin) * NB / range (monotone in k; NB = 2 buckets per product of capacity, so most products
//      are alone in theirs): one atomicAdd per product counts the bucket and returns the product's rank in it
//      (16-bit counters, two per word)
//   c. scan of the bucket counts; products -> staging at start[bucket] + rank
//   d. one thread per 2*PER consecutive buckets = one contiguous piece of the staging area: a single pass appends
//      the products whose column exceeds everything before (the common case: buckets are ordered among
//      themselves) and inserts the others by (column, value bits); equal columns are then summed in that order,
//      so the sum does not depend on the order the atomics happened in
//   e. scan of the distinct counts; the pieces go back to the head of the segment, in column order.
// A pseudo-row with a bucket of more than ESC_BUCKET_MAX products (one column hit very often, clustered
// columns) is left untouched and flagged: k_esc_reduce below takes it.
constexpr int ESC_BUCKET_MAX = 32;
constexpr int ESC_BPP = 2;   // buckets per product of capacity
constexpr size_t esc_smem(int cap) { return (size_t)cap * (12 + 2 * ESC_BPP) + 16; }   // staging + 16-bit counters

template <int THREADS, int PER, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_esc_sortmerge(const int32_t *__restrict__ plist, int nbin,
                                                                 const int64_t *__restrict__ poff, int32_t *ecol, double *eval,
                                                                 int32_t *__restrict__ pnnz, int *__restrict__ fail,
                                                                 int *__restrict__ any_fail)
{
    constexpr int CAP = THREADS * PER, NB = ESC_BPP * CAP, WPT = ESC_BPP * PER / 2;   // WPT counter words per thread
    extern __shared__ __align__(16) unsigned char s_raw[];
    double *sval = reinterpret_cast<double *>(s_raw);
    int32_t *skey = reinterpret_cast<int32_t *>(s_raw + sizeof(double) * CAP);
    unsigned *cntw = reinterpret_cast<unsigned *>(s_raw + (sizeof(double) + sizeof(int32_t)) * CAP);   // NB/2 + 1 words
    const unsigned short *cnt16 = reinterpret_cast<const unsigned short *>(cntw);                     // NB + 1 halves
    __shared__ int s_min, s_max, s_big;
    __shared__ int s_wt[33];
    const int tid = threadIdx.x;
    for (int it = blockIdx.x; it < nbin; it += gridDim.x) {
        const int32_t p = plist[it];
        const int64_t off = poff[p];
        const int n = (int)(poff[p + 1] - off);
        int32_t *oc = ecol + off;
        double *ov = eval + off;
#pragma unroll
        for (int q = 0; q < WPT; q++)
            cntw[tid + q * THREADS] = 0;
        if (tid == 0) {
            s_min = INT32_MAX;
            s_max = -1;
            s_big = 0;
        }
        int32_t k[PER];
        double v[PER];
        int kmin = INT32_MAX, kmax = -1;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int i = tid + q * THREADS;
            k[q] = 0;
            v[q] = 0.0;
            if (i < n) {
                k[q] = oc[i];
                v[q] = ov[i];
                kmin = min(kmin, k[q]);
                kmax = max(kmax, k[q]);
            }
        }
        __syncthreads();
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if ((tid & 31) == 0 && kmax >= 0) {
            atomicMin(&s_min, kmin);
            atomicMax(&s_max, kmax);
        }
        __syncthreads();
        kmin = s_min;
        // float arithmetic: int -> float, a multiplication by a positive constant and the truncation are all
        // monotone, which is all the bucket map has to be
        const float fscale = (float)NB / ((float)(s_max - kmin) + 1.0f);
        auto bucket = [&](int32_t key) -> unsigned {
            return (unsigned)min(NB - 1, __float2int_rz(__int2float_rz(key - kmin) * fscale));
        };
        unsigned rk[(PER + 3) / 4];   // ranks, 8 bits each (a rank above ESC_BUCKET_MAX flags the pseudo-row anyway)
#pragma unroll
        for (int q = 0; q < (PER + 3) / 4; q++)
            rk[q] = 0;
#pragma unroll
        for (int q = 0; q < PER; q++)
            if (tid + q * THREADS < n) {
                const unsigned b = bucket(k[q]), sh = 16u * (b & 1u);
                const unsigned old = (atomicAdd(&cntw[b >> 1], 1u << sh) >> sh) & 0xffffu;
                rk[q >> 2] |= min(old, 255u) << (8 * (q & 3));
            }
        __syncthreads();
        {
            unsigned loc[WPT];
            int sum = 0, big = 0;
#pragma unroll
            for (int q = 0; q < WPT; q++) {
                loc[q] = cntw[tid * WPT + q];
                const int lo = (int)(loc[q] & 0xffffu), hi = (int)(loc[q] >> 16);
                sum += lo + hi;
                big |= (lo > ESC_BUCKET_MAX) | (hi > ESC_BUCKET_MAX);
            }
            if (big)
                s_big = 1;
            int tot;
            int ex = block_exclusive_scan<int>(sum, s_wt, tot);
#pragma unroll
            for (int q = 0; q < WPT; q++) {
                const int lo = (int)(loc[q] & 0xffffu), hi = (int)(loc[q] >> 16);
                cntw[tid * WPT + q] = (unsigned)ex | (unsigned)(ex + lo) << 16;
                ex += lo + hi;
            }
            if (tid == THREADS - 1)
                cntw[NB / 2] = (unsigned)ex;
        }
        __syncthreads();
        if (s_big) {   // (uniform) left as it is for the hash kernel
            if (tid == 0) {
                fail[p] = 1;
                *any_fail = 1;
            }
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int q = 0; q < PER; q++)
            if (tid + q * THREADS < n) {
                const unsigned pos = cnt16[bucket(k[q])] + ((rk[q >> 2] >> (8 * (q & 3))) & 0xffu);
                skey[pos] = k[q];
                sval[pos] = v[q];
            }
        __syncthreads();
        const int s0 = (int)cnt16[tid * 2 * WPT], s1 = (int)cnt16[(tid + 1) * 2 * WPT];
        int w = s0, ndup = 0;   // the piece's sorted entries live in [s0, w); ndup of them repeat a column
        {
            int32_t last = -1;
            for (int i = s0; i < s1; i++) {
                const int32_t ki = skey[i];
                if (ki > last) {
                    if (w != i) {
                        skey[w] = ki;
                        sval[w] = sval[i];
                    }
                    last = ki;
                } else {
                    const double vi = sval[i];
                    const long long bi = __double_as_longlong(vi);
                    int j = w - 1;
                    bool eq = false;
                    while (j >= s0) {
                        const int32_t kj = skey[j];
                        if (kj < ki)
                            break;
                        const double vj = sval[j];
                        if (kj == ki) {
                            eq = true;
                            if (__double_as_longlong(vj) <= bi)
                                break;
                        }
                        skey[j + 1] = kj;
                        sval[j + 1] = vj;
                        j--;
                    }
                    skey[j + 1] = ki;
                    sval[j + 1] = vi;
                    ndup += eq;
                }
                w++;
            }
        }
        int tot;
        int out = block_exclusive_scan<int>(w - s0 - ndup, s_wt, tot);
        for (int i = s0; i < w; out++) {   // equal columns are adjacent and ordered by value: summed in that order
            const int32_t ki = skey[i];
            double acc = sval[i];
            for (i++; i < w && skey[i] == ki; i++)
                acc += sval[i];
            oc[out] = ki;
            ov[out] = acc;
        }
        if (tid == 0)
            pnnz[p] = tot;
        __syncthreads();
    }
}

// The general kernel for a pseudo-row of at most SLOTS/2 products: k_num_cta's hash accumulation + distribution
// sort (spgemm.cu), reading the expansion instead of walking A and B, writing back into the segment.  Persistent
// CTAs over the list; only pseudo-rows flagged by k_esc_sortmerge are taken.
// AGG: the pseudo-rows of k_esc_check -- far more products than columns, millions on the most popular column: the
// lanes of a warp that hold the same column add their values up first (match_any + 32 shuffles) and one of them
// updates the table, so that a hot slot sees one CAS loop per warp and step instead of up to 32.
template <int SLOTS, int THREADS, int NBUCK, bool AGG = false>
__global__ void __launch_bounds__(THREADS) k_esc_reduce(const int32_t *__restrict__ plist, int nbin, const int64_t *__restrict__ poff,
                                                        int32_t *ecol, double *eval, int32_t *__restrict__ pnnz,
                                                        const int *__restrict__ fail, const int *__restrict__ any_fail)
{
    static_assert(NBUCK % THREADS == 0, "one scan pass");
    if (!*any_fail)
        return;
    extern __shared__ __align__(16) unsigned char s_raw[];
    double *vals = reinterpret_cast<double *>(s_raw);
    int32_t *keys = reinterpret_cast<int32_t *>(s_raw + sizeof(double) * SLOTS);
    unsigned *cnt = reinterpret_cast<unsigned *>(s_raw + (sizeof(double) + sizeof(int32_t)) * SLOTS);
    __shared__ int s_min, s_max, s_big, s_nz;
    __shared__ int s_wt[33];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int it = blockIdx.x; it < nbin; it += gridDim.x) {
        const int32_t p = plist[it];
        if (!fail[p])
            continue;
        __syncthreads();   // the previous pseudo-row is finished with the table
        for (int i = tid; i < SLOTS; i += THREADS) {
            keys[i] = EMPTY_KEY;
            vals[i] = 0.0;
        }
        for (int i = tid; i < NBUCK; i += THREADS)
            cnt[i] = 0;
        if (tid == 0) {
            s_min = INT32_MAX;
            s_max = -1;
            s_big = 0;
            s_nz = 0;
        }
        __syncthreads();
        const int64_t off = poff[p];
        const int n = (int)(poff[p + 1] - off);
        int32_t *oc = ecol + off;
        double *ov = eval + off;
        int kmin = INT32_MAX, kmax = -1, count = 0;
        for (int i0 = 0; i0 < n; i0 += THREADS) {   // (uniform trip count: the AGG variant votes)
            const int i = i0 + tid;
            bool act = i < n;
            int32_t k = act ? oc[i] : -1 - lane;     // (idle lanes: distinct negative keys, they match nobody)
            double v = act ? ov[i] : 0.0;
            if (AGG) {
                const unsigned m = __match_any_sync(0xffffffffu, k);
                double acc = 0.0;
#pragma unroll 8
                for (int l = 0; l < 32; l++) {
                    const double vl = __shfl_sync(0xffffffffu, v, l);
                    if ((m >> l) & 1u)
                        acc += vl;                   // the group's values in lane order
                }
                act = act && lane == __ffs((int)m) - 1;
                v = acc;
            }
            if (!act)
                continue;
            kmin = min(kmin, k);
            kmax = max(kmax, k);
            unsigned h = hash_col(k, SLOTS - 1);
            while (true) {
                const int32_t old = atomicCAS(&keys[h], EMPTY_KEY, k);
                if (old == EMPTY_KEY || old == k) {
                    count += old == EMPTY_KEY;
                    atomicAdd(&vals[h], v);
                    break;
                }
                h = (h + 1) & (SLOTS - 1);
            }
        }
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        count = warp_sum(count);
        if (lane == 0 && kmax >= 0) {
            atomicMin(&s_min, kmin);
            atomicMax(&s_max, kmax);
            atomicAdd(&s_nz, count);
        }
        __syncthreads();   // every product has been read: the segment's head can be overwritten
        const int nz = s_nz;
        if (tid == 0)
            pnnz[p] = nz;
        if (nz == 0)
            continue;
        kmin = s_min;
        const uint64_t range = (uint64_t)(s_max - kmin) + 1;
        const uint64_t scale = ((uint64_t)NBUCK << 32) / range;   // bucket = (k - kmin) * NBUCK / range, monotone in k
        for (int i = tid; i < SLOTS; i += THREADS) {
            const int32_t k = keys[i];
            if (k != EMPTY_KEY)
                atomicAdd(&cnt[(unsigned)(((uint64_t)(k - kmin) * scale) >> 32)], 1u);
        }
        __syncthreads();
        {
            constexpr int PER = NBUCK / THREADS;
            unsigned loc[PER];
            int sum = 0, big = 0;
#pragma unroll
            for (int q = 0; q < PER; q++) {
                loc[q] = cnt[tid * PER + q];
                sum += loc[q];
                big |= loc[q] > NUM_BUCKET_MAX;
            }
            if (big)
                s_big = 1;
            int tot;
            int ex = block_exclusive_scan<int>(sum, s_wt, tot);
#pragma unroll
            for (int q = 0; q < PER; q++) {
                cnt[tid * PER + q] = ex;
                ex += loc[q];
            }
        }
        __syncthreads();
        if (!s_big) {
            for (int i = tid; i < SLOTS; i += THREADS) {
                const int32_t k = keys[i];
                if (k != EMPTY_KEY) {
                    const unsigned b = (unsigned)(((uint64_t)(k - kmin) * scale) >> 32);
                    const unsigned pos = atomicAdd(&cnt[b], 1u);
                    oc[pos] = k;
                    ov[pos] = vals[i];
                }
            }
            __syncthreads();
            for (int b = tid; b < NBUCK; b += THREADS) {
                const int s0 = b ? (int)cnt[b - 1] : 0, e0 = (int)cnt[b];
                for (int i = s0 + 1; i < e0; i++) {
                    const int32_t k = oc[i];
                    const double v = ov[i];
                    int j = i - 1;
                    while (j >= s0 && oc[j] > k) {
                        oc[j + 1] = oc[j];
                        ov[j + 1] = ov[j];
                        j--;
                    }
                    oc[j + 1] = k;
                    ov[j + 1] = v;
                }
            }
            continue;
        }
        // clustered columns: bitonic sort of the whole table (empty slots, INT32_MAX, go last)
        for (int k = 2; k <= SLOTS; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < SLOTS / 2; t += THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int q = i | j;
                    const bool up = (i & k) == 0;
                    const int32_t ki = keys[i], kq = keys[q];
                    if ((ki > kq) == up) {
                        keys[i] = kq;
                        keys[q] = ki;
                        const double vi = vals[i];
                        vals[i] = vals[q];
                        vals[q] = vi;
                    }
                }
                __syncthreads();
            }
        }
        for (int i = tid; i < nz; i += THREADS) {
            oc[i] = keys[i];
            ov[i] = vals[i];
        }
    }
}

__global__ void __launch_bounds__(256) k_esc_rownnz(const int32_t *__restrict__ rows, int n, const int *__restrict__ pbase,
                                                     const int64_t *__restrict__ pg, const int *__restrict__ bad,
                                                     int32_t *__restrict__ row_nnz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !bad[i])
        row_nnz[rows[i]] = (int32_t)(pg[pbase[i + 1]] - pg[pbase[i]]);
}

// segment p -> C, at rowptr + (distinct columns of the row's earlier ranges)
__global__ void __launch_bounds__(128) k_esc_emit(const int32_t *__restrict__ rows, const int32_t *__restrict__ prow,
                                                   const int *__restrict__ pbase, const int64_t *__restrict__ poff,
                                                   const int64_t *__restrict__ pg, const int32_t *__restrict__ ecol,
                                                   const double *__restrict__ eval, const int64_t *__restrict__ c_rp,
                                                   int32_t *__restrict__ c_ci, double *__restrict__ c_vs)
{
    const int64_t p = blockIdx.x;
    const int n = (int)(pg[p + 1] - pg[p]);
    if (n == 0)
        return;
    const int i = prow[p];
    const int64_t dst = c_rp[rows[i]] + (pg[p] - pg[pbase[i]]), src = poff[p];
    for (int k = threadIdx.x; k < n; k += 128) {
        c_ci[dst + k] = ecol[src + k];
        __stcs(&c_vs[dst + k], eval[src + k]);
    }
}

// Steps 1-4 for the `n` rows of `rows`.  Returns CSRK_OK with st.active = false when the path declines (nothing
// has been written then), st.active = true when row_nnz of the accepted rows is in place; st.n_old rows
// (st.old_list) were handed back.
static int esc_symbolic(const MatView &A, const MatView &B, const int32_t *rows, int n, const int64_t *prod, int both_f32,
                        int32_t *row_nnz, EscState &st, cudaStream_t s)
{
    st.active = false;
    if (n <= 0)
        return CSRK_OK;
    const int64_t target = std::max<int64_t>(16, options().esc_target.load());
    DevBuf nrange, nitem, item_off;
    CSRK_TRY(nrange.alloc(sizeof(int) * (size_t)n, s));
    CSRK_TRY(nitem.alloc(sizeof(int) * (size_t)n, s));
    CSRK_TRY(item_off.alloc(sizeof(int) * ((size_t)n + 1), s));
    CSRK_TRY(st.pbase.alloc(sizeof(int) * ((size_t)n + 1), s));
    CSRK_TRY(st.bad.alloc_zero(sizeof(int) * (size_t)n, s));
    CSRK_LAUNCH(k_esc_rows, (unsigned)div_up(n, 256), 256, 0, s, A, rows, n, prod, target, (int64_t)B.ncols, nrange.as<int>(),
                nitem.as<int>(), st.bad.as<int>());
    CSRK_TRY((exclusive_scan<int>(ArrayLoader<int>{nrange.as<int>()}, (int64_t)n, st.pbase.as<int>(), s)));
    CSRK_TRY((exclusive_scan<int>(ArrayLoader<int>{nitem.as<int>()}, (int64_t)n, item_off.as<int>(), s)));
    int tot[2] = {0, 0};
    CSRK_CUDA(cudaMemcpyAsync(&tot[0], st.pbase.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&tot[1], item_off.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    const int64_t np = tot[0];
    const int nitems = tot[1];
    if (np <= 0 || nitems <= 0)
        return CSRK_OK;
    DevBuf pcount, pcur, n_old_d;
    CSRK_TRY(pcount.alloc_zero(sizeof(unsigned) * (size_t)np, s));
    CSRK_TRY(st.prow.alloc(sizeof(int32_t) * (size_t)np, s));
    CSRK_TRY(st.poff.alloc(sizeof(int64_t) * ((size_t)np + 1), s));
    CSRK_TRY(st.old_list.alloc(sizeof(int32_t) * (size_t)n, s));
    CSRK_TRY(n_old_d.alloc_zero(sizeof(int), s));
    CSRK_LAUNCH(k_esc_prow, (unsigned)div_up((int64_t)n * 32, 256), 256, 0, s, st.pbase.as<int>(), n, st.prow.as<int32_t>());
    unsigned stride = 1;
    if (options().esc_stride.load())
        for (unsigned c : {7919u, 7907u, 7901u, 7883u, 7879u})   // a prime that does not divide the item count
            if ((unsigned)nitems % c) {
                stride = c;
                break;
            }
    CSRK_LAUNCH(k_esc_count, (unsigned)nitems, ESC_WALK_THREADS, 0, s, A, B, rows, n, item_off.as<int>(), st.pbase.as<int>(),
                pcount.as<unsigned>(), stride);
    const unsigned pgrid = (unsigned)div_up(np, 256);
    DevBuf fail;   // pseudo-rows for the hash kernel (set here for the oversized ones, later by the sort/merge kernels)
    CSRK_TRY(fail.alloc_zero(sizeof(int) * ((size_t)np + 1), s));
    int *fl = fail.as<int>(), *af = fl + np;
    CSRK_LAUNCH(k_esc_check, pgrid, 256, 0, s, pcount.as<unsigned>(), st.prow.as<int32_t>(), np, st.pbase.as<int>(),
                (int64_t)B.ncols, st.bad.as<int>(), fl, af);
    CSRK_LAUNCH(k_esc_mask, pgrid, 256, 0, s, pcount.as<unsigned>(), st.prow.as<int32_t>(), np, st.bad.as<int>());
    CSRK_LAUNCH(k_esc_badlist, (unsigned)div_up(n, 256), 256, 0, s, rows, n, st.bad.as<int>(), st.old_list.as<int32_t>(),
                n_old_d.as<int>());
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<unsigned>{pcount.as<unsigned>()}, np, st.poff.as<int64_t>(), s)));
    int64_t n_exp = 0;
    int n_old = 0;
    CSRK_CUDA(cudaMemcpyAsync(&n_exp, st.poff.as<int64_t>() + np, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&n_old, n_old_d.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    CSRK_TRACE_MARK("spgemm esc: ranges + count", s);
    // the expansion: from the pool, so that it goes back when the call ends; decline if it does not fit
    {
        size_t free_b = 0, total_b = 0;
        CSRK_CUDA(cudaMemGetInfo(&free_b, &total_b));
        {   // blocks the stream-ordered pool holds in reserve are free for this purpose
            cudaMemPool_t pool;
            uint64_t reserved = 0, used = 0;
            if (cudaDeviceGetDefaultMemPool(&pool, ctx().device) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
                free_b += (size_t)(reserved - used);
            (void)cudaGetLastError();
        }
        const int64_t budget = options().esc_budget.load();
        const size_t need = (size_t)n_exp * 12 + (1u << 20);
        // what is left must still hold the result (at most 12 B per expanded product) beside the expansion
        if ((budget > 0 && need > (size_t)budget) || need * 2 > free_b)
            return CSRK_OK;
        if (st.ecol.alloc_owned(sizeof(int32_t) * (size_t)n_exp, s) != CSRK_OK ||
            st.eval.alloc_owned(sizeof(double) * (size_t)n_exp, s) != CSRK_OK) {
            st.ecol.reset();
            st.eval.reset();
            return CSRK_OK;
        }
    }
    CSRK_TRY(pcur.alloc_zero(sizeof(unsigned) * (size_t)np, s));
    CSRK_LAUNCH(k_esc_scatter, (unsigned)nitems, ESC_WALK_THREADS, 0, s, A, B, rows, n, item_off.as<int>(), st.pbase.as<int>(),
                st.bad.as<int>(), st.poff.as<int64_t>(), pcur.as<unsigned>(), st.ecol.as<int32_t>(), st.eval.as<double>(),
                both_f32, 1u);
    CSRK_TRACE_MARK("spgemm esc: scatter", s);
    // pseudo-rows by size
    BinSpec spec{{0, 64, 512, 2048, ESC_CAP, INT64_MAX}};
    int cnt[NBINS], off[NBINS + 1];
    DevBuf plist;
    CSRK_TRY(bin_rows(reinterpret_cast<const int32_t *>(pcount.p), np, spec, cnt, off, plist, s));
    CSRK_TRY(st.pnnz.alloc_zero(sizeof(int32_t) * (size_t)np, s));
    const int32_t *PL = plist.as<int32_t>();
    const int64_t *po = st.poff.as<int64_t>();
    int32_t *ec = st.ecol.as<int32_t>();
    double *ev = st.eval.as<double>();
    int32_t *pz = st.pnnz.as<int32_t>();
    if (cnt[1])
        CSRK_LAUNCH((k_esc_reduce_warp<128>), (unsigned)div_up(cnt[1], 8), 256, 0, s, PL + off[1], cnt[1], po, ec, ev, pz);
    const int sms = ctx().sm_count;
    // bucket sort + merge, then the hash kernel for what it flagged (usually nothing: its CTAs only read the flags)
    if (cnt[2]) {
        auto k = k_esc_sortmerge<128, 4, 8>;
        CSRK_LAUNCH(k, (unsigned)std::min(cnt[2], sms * 32), 128, esc_smem(512), s, PL + off[2], cnt[2], po, ec, ev, pz, fl, af);
        auto h = k_esc_reduce<1024, 128, 512>;
        CSRK_LAUNCH(h, (unsigned)std::min(cnt[2], sms * 8), 128, 1024 * 12 + 512 * 4, s, PL + off[2], cnt[2], po, ec, ev, pz, fl, af);
    }
    if (cnt[3]) {
        auto k = k_esc_sortmerge<256, 8, 5>;
        CSRK_LAUNCH(k, (unsigned)std::min(cnt[3], sms * 16), 256, esc_smem(2048), s, PL + off[3], cnt[3], po, ec, ev, pz, fl, af);
        auto h = k_esc_reduce<4096, 256, 2048>;
        CSRK_TRY(optin_smem(h, 4096 * 12 + 2048 * 4));
        CSRK_LAUNCH(h, (unsigned)std::min(cnt[3], sms * 4), 256, 4096 * 12 + 2048 * 4, s, PL + off[3], cnt[3], po, ec, ev, pz, fl, af);
    }
    if (cnt[4]) {
        auto k = k_esc_sortmerge<512, 16, 1>;
        CSRK_TRY(optin_smem(k, esc_smem(8192)));
        CSRK_LAUNCH(k, (unsigned)std::min(cnt[4], sms), 512, esc_smem(8192), s, PL + off[4], cnt[4], po, ec, ev, pz, fl, af);
        auto h = k_esc_reduce<16384, 512, 4096>;
        CSRK_TRY(optin_smem(h, 16384 * 12 + 4096 * 4));
        CSRK_LAUNCH(h, (unsigned)std::min(cnt[4], sms), 512, 16384 * 12 + 4096 * 4, s, PL + off[4], cnt[4], po, ec, ev, pz, fl, af);
    }
    if (cnt[5]) {   // more products than ESC_CAP on at most ESC_CAP columns (k_esc_check): straight to the hash kernel
        auto h = k_esc_reduce<16384, 512, 4096, true>;
        CSRK_TRY(optin_smem(h, 16384 * 12 + 4096 * 4));
        CSRK_LAUNCH(h, (unsigned)std::min(cnt[5], sms), 512, 16384 * 12 + 4096 * 4, s, PL + off[5], cnt[5], po, ec, ev, pz, fl, af);
    }
    CSRK_TRY(st.pg.alloc(sizeof(int64_t) * ((size_t)np + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<int32_t>{pz}, np, st.pg.as<int64_t>(), s)));
    CSRK_LAUNCH(k_esc_rownnz, (unsigned)div_up(n, 256), 256, 0, s, rows, n, st.pbase.as<int>(), st.pg.as<int64_t>(),
                st.bad.as<int>(), row_nnz);
    CSRK_TRACE_MARK("spgemm esc: reduce", s);
    st.active = true;
    st.n_rows = n;
    st.rows = rows;
    st.np = np;
    st.n_exp = n_exp;
    st.n_old = n_old;
    return CSRK_OK;
}

static int esc_emit(const EscState &st, const int64_t *c_rp, int32_t *c_ci, double *c_vs, cudaStream_t s)
{
    if (!st.active || st.np == 0)
        return CSRK_OK;
    CSRK_LAUNCH(k_esc_emit, (unsigned)st.np, 128, 0, s, st.rows, st.prow.as<int32_t>(), st.pbase.as<int>(), st.poff.as<int64_t>(),
                st.pg.as<int64_t>(), st.ecol.as<int32_t>(), st.eval.as<double>(), c_rp, c_ci, c_vs);
    return CSRK_OK;
}

}  // namespace csrk
