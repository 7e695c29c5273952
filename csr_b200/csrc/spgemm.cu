// spgemm.cu -- mult_ab (csr/kernels/numba/multiply.py:13-38) for sm_100a:
// two-phase row-wise SpGEMM, C = A*B, with SORTED column output.
//
// The reference is SMMP: a symbolic pass (_sym_mm, multiply.py:60-100) that
// counts and lists the distinct columns of every output row, then a numeric
// pass (_num_mm, :103-129) that accumulates a*b products into a dense float64
// work array.  Here:
//   0. k_row_products   P_i = sum over A's row i of len(B row j): the upper bound
//                       that sizes the symbolic accumulator of row i
//   1. symbolic         rows binned by P_i: warp-per-row / CTA-per-row hash sets
//                       in shared memory, or a column BITMAP (shared memory when
//                       ncols fits, else per-CTA global scratch) for heavy rows
//                       -> exact row nnz (value independent => rowptrs bit-exact)
//   2. exclusive scan   -> rowptrs (int64 internally; int32 out when Z <= INT32_MAX,
//                       the dtype rule of csr/csr.py:90-93)
//   3. numeric          rows binned by their exact nnz: shared-memory hash
//                       accumulators (float64 values) sorted by column before
//                       they are written, or -- heavy rows -- a dense accumulator
//                       over column windows in shared memory whose sweep emits the
//                       columns in order: 64-bit fixed point on native atomics over
//                       power-of-two equilibrated operands with an exact side list
//                       (k_num_fixed), or float64 owner-computes without atomics
//                       (k_num_owner).
// Wide results (more columns than four windows) take the one-phase expand / sort /
// compress path of spgemm_esc.cuh for every row above the small hash bins.
// Products are formed in numba's promoted type (f4*f4 -> f4, else f8) and summed
// in float64, like multiply.py:120.  Only the summation ORDER differs from the
// reference (hence rtol instead of bit-exact values).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "radix.cuh"

namespace csrk {

struct MatView {
    int32_t nrows, ncols;
    int64_t nnz;
    const void *rp;
    int rp64;
    const int32_t *ci;
    const void *vs;
    int vk;
};

static MatView view(const csrk_matrix *m)
{
    return MatView{m->nrows, m->ncols, m->nnz, m->rp, m->rp_is64, m->ci, m->vs, m->val_kind};
}

constexpr int NBINS = 6;
struct BinSpec {
    int64_t upper[NBINS];  // value <= upper[b] -> bin b (first match); last is INT64_MAX
};

constexpr int32_t EMPTY_KEY = INT32_MAX;  // ncols <= INT32_MAX so no column equals it

__device__ __forceinline__ unsigned hash_col(int32_t k, unsigned mask)
{
    return ((unsigned)k * 0x9E3779B1u >> 7) & mask;
}

__device__ __forceinline__ double product(double av, double bv, int both_f32)
{
    return both_f32 ? (double)__fmul_rn((float)av, (float)bv) : __dmul_rn(av, bv);
}

// ------------------------------------------------------------ step 0: products
// One warp per row; rows longer than ROWP_LONG entries (the 10^7-entry rows of configs[4]: one warp needed 171 ms for
// the longest) are cut into pieces of ROWP_LONG entries that k_row_products_long's CTAs sum and add atomically.
constexpr int ROWP_LONG = 1 << 16;

__global__ void __launch_bounds__(256) k_row_products(MatView A, MatView B, int64_t *__restrict__ prod,
                                                      unsigned long long *__restrict__ total, int64_t *__restrict__ long_items,
                                                      int max_items)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t p = 0;
    if (row < A.nrows) {
        const int64_t s = ld_rp(A.rp, A.rp64, row), e = ld_rp(A.rp, A.rp64, row + 1);
        if (e - s > ROWP_LONG) {
            const int n = (int)((e - s + ROWP_LONG - 1) / ROWP_LONG);
            if (lane == 0) {
                prod[row] = 0;
                atomicMax(total + 2, (unsigned long long)(e - s));
            }
            int base = 0;
            if (lane == 0)
                base = (int)atomicAdd(total + 3, (unsigned long long)n);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (int c = lane; c < n && base + c < max_items; c += 32)
                long_items[base + c] = (row << 20) | c;   // (fewer than 2^20 pieces per row: rows have < 2^31 entries)
            return;
        }
        for (int64_t jj = s + lane; jj < e; jj += 32) {
            const int32_t j = A.ci[jj];
            p += ld_rp(B.rp, B.rp64, (int64_t)j + 1) - ld_rp(B.rp, B.rp64, j);
        }
        p = warp_sum(p);
        if (lane == 0) {
            prod[row] = p;
            if (p) {
                atomicAdd(total, (unsigned long long)p);
                atomicMax(total + 1, (unsigned long long)p);  // the heaviest row decides whether rows get chunked
                atomicMax(total + 2, (unsigned long long)(e - s));  // longest row with products: sizes the fixed-point headroom
            }
        }
    }
}

// the pieces of the long rows: persistent CTAs over the item list (its length is still on the device)
__global__ void __launch_bounds__(256) k_row_products_long(MatView A, MatView B, int64_t *__restrict__ prod,
                                                           unsigned long long *__restrict__ total,
                                                           const int64_t *__restrict__ long_items, int max_items)
{
    __shared__ long long s_sum[8];
    const int n = (int)min((unsigned long long)max_items, total[3]);
    for (int it = blockIdx.x; it < n; it += gridDim.x) {
        const int64_t row = long_items[it] >> 20, c = long_items[it] & ((1 << 20) - 1);
        const int64_t s = ld_rp(A.rp, A.rp64, row) + c * ROWP_LONG, e = min(ld_rp(A.rp, A.rp64, row + 1), s + ROWP_LONG);
        long long p = 0;
        for (int64_t jj = s + threadIdx.x; jj < e; jj += 256) {
            const int32_t j = A.ci[jj];
            p += ld_rp(B.rp, B.rp64, (int64_t)j + 1) - ld_rp(B.rp, B.rp64, j);
        }
        p = warp_sum(p);
        if ((threadIdx.x & 31) == 0)
            s_sum[threadIdx.x >> 5] = p;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0;
            for (int w = 0; w < 8; w++)
                t += s_sum[w];
            if (t) {
                atomicAdd(reinterpret_cast<unsigned long long *>(prod) + row, (unsigned long long)t);
                atomicAdd(total, (unsigned long long)t);
            }
        }
        __syncthreads();
    }
}
// the heaviest row may be one of the long ones
__global__ void __launch_bounds__(256) k_row_products_max(const int64_t *__restrict__ prod, unsigned long long *__restrict__ total,
                                                          const int64_t *__restrict__ long_items, int max_items)
{
    const int n = (int)min((unsigned long long)max_items, total[3]);
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x)
        if ((long_items[it] & ((1 << 20) - 1)) == 0)
            atomicMax(total + 1, (unsigned long long)prod[long_items[it] >> 20]);
}

// ------------------------------------------------------------------- binning
template <typename T>
__global__ void __launch_bounds__(256) k_bin_count(const T *__restrict__ val, int64_t n, BinSpec spec, int *__restrict__ counts)
{
    __shared__ int local[NBINS];
    if (threadIdx.x < NBINS)
        local[threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t v = (int64_t)val[i];
        int b = 0;
        while (b < NBINS - 1 && v > spec.upper[b])
            b++;
        atomicAdd(&local[b], 1);
    }
    __syncthreads();
    if (threadIdx.x < NBINS && local[threadIdx.x])
        atomicAdd(&counts[threadIdx.x], local[threadIdx.x]);
}

template <typename T>
__global__ void __launch_bounds__(256) k_bin_fill(const T *__restrict__ val, int64_t n, BinSpec spec, int *__restrict__ cursors,
                                                  int32_t *__restrict__ list)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t v = (int64_t)val[i];
        int b = 0;
        while (b < NBINS - 1 && v > spec.upper[b])
            b++;
        list[atomicAdd(&cursors[b], 1)] = (int32_t)i;
    }
}

// ----------------------------------------------------------- symbolic: hashes
// One warp per row; SLOTS-entry hash set per warp.  Requires P_i <= SLOTS/2.
template <int SLOTS>
__global__ void __launch_bounds__(256) k_sym_warp(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin,
                                                  int32_t *__restrict__ row_nnz)
{
    __shared__ int32_t tab[8][SLOTS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int idx = blockIdx.x * 8 + w;
    for (int i = lane; i < SLOTS; i += 32)
        tab[w][i] = EMPTY_KEY;
    __syncwarp();
    if (idx >= nbin)
        return;
    const int32_t row = rows[idx];
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    int count = 0;
    for (int64_t jj = as; jj < ae; jj++) {
        const int32_t j = A.ci[jj];
        const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
        for (int64_t kk = bs + lane; kk < be; kk += 32) {
            const int32_t k = B.ci[kk];
            unsigned h = hash_col(k, SLOTS - 1);
            while (true) {
                const int32_t old = atomicCAS(&tab[w][h], EMPTY_KEY, k);
                if (old == EMPTY_KEY) {
                    count++;
                    break;
                }
                if (old == k)
                    break;
                h = (h + 1) & (SLOTS - 1);
            }
        }
    }
    count = warp_sum(count);
    if (lane == 0)
        row_nnz[row] = count;
}

// One CTA per row; SLOTS-entry hash set in dynamic shared memory.  P_i <= SLOTS/2.
template <int SLOTS, int THREADS>
__global__ void __launch_bounds__(THREADS) k_sym_cta(MatView A, MatView B, const int32_t *__restrict__ rows,
                                                     int32_t *__restrict__ row_nnz)
{
    extern __shared__ int32_t s_tab[];
    __shared__ int s_count;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < SLOTS; i += THREADS)
        s_tab[i] = EMPTY_KEY;
    if (tid == 0)
        s_count = 0;
    __syncthreads();
    const int32_t row = rows[blockIdx.x];
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    int count = 0;
    for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
        const int32_t j = A.ci[jj];
        const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
        for (int64_t kk = bs + lane; kk < be; kk += 32) {
            const int32_t k = B.ci[kk];
            unsigned h = hash_col(k, SLOTS - 1);
            while (true) {
                const int32_t old = atomicCAS(&s_tab[h], EMPTY_KEY, k);
                if (old == EMPTY_KEY) {
                    count++;
                    break;
                }
                if (old == k)
                    break;
                h = (h + 1) & (SLOTS - 1);
            }
        }
    }
    count = warp_sum(count);
    if (lane == 0 && count)
        atomicAdd(&s_count, count);
    __syncthreads();
    if (tid == 0)
        row_nnz[row] = s_count;
}

// ------------------------------------------------- symbolic: bitmap (heavy rows)
// Persistent CTAs pull rows from a counter.  The column bitmap lives in shared
// memory when SMEM_BM, else in this CTA's slice of a zero-initialised global
// scratch; either way it is restored to zero while it is counted.  When `keep` is
// given, the row's bitmap is also stored there (slot = position in the bin) so the
// numeric phase does not have to rebuild it: keep_slot[row] = slot.
// BYTES: the marks are one BYTE per column in shared memory, set with plain stores (idempotent: no atomics;
// stores of several lanes into one 32-bit word merge in the same wavefront, where atomicOr on the hot words of
// popular columns serialises), and packed into bitmap words while they are counted.
template <bool SMEM_BM, int THREADS, bool BYTES = false>
__global__ void __launch_bounds__(THREADS) k_sym_bitmap(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin,
                                                        int32_t *__restrict__ row_nnz, unsigned *__restrict__ gbm,
                                                        int n_words, int *__restrict__ work_counter,
                                                        unsigned *__restrict__ keep, int32_t *__restrict__ keep_slot,
                                                        const int *__restrict__ item_off)
{
    extern __shared__ __align__(16) unsigned s_bm[];
    __shared__ int s_idx, s_item, s_count;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    static_assert(!BYTES || SMEM_BM, "the byte map lives in shared memory");
    unsigned *bm = SMEM_BM ? s_bm : gbm + (size_t)blockIdx.x * n_words;
    unsigned char *by = reinterpret_cast<unsigned char *>(s_bm);   // BYTES: 32 bytes per bitmap word
    if (SMEM_BM) {
        for (int i = tid; i < n_words * (BYTES ? 8 : 1); i += THREADS)
            s_bm[i] = 0;
    }
    // the 32 marks of bitmap word i -> the word, and the marks back to zero
    auto take_word = [&](int i) -> unsigned {
        if constexpr (BYTES) {
            uint4 *p = reinterpret_cast<uint4 *>(by + 32 * (size_t)i);
            const uint4 q0 = p[0], q1 = p[1];
            const unsigned w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            unsigned bits = 0;
#pragma unroll
            for (int j = 0; j < 8; j++)
                bits |= ((w[j] & 1u) | ((w[j] >> 7) & 2u) | ((w[j] >> 14) & 4u) | ((w[j] >> 21) & 8u)) << (4 * j);
            if (bits) {
                p[0] = make_uint4(0u, 0u, 0u, 0u);
                p[1] = make_uint4(0u, 0u, 0u, 0u);
            }
            return bits;
        } else {
            const unsigned bits = SMEM_BM ? bm[i] : __ldcg(&bm[i]);
            if (bits) {
                if (SMEM_BM)
                    bm[i] = 0;
                else
                    __stcg(&bm[i], 0u);
            }
            return bits;
        }
    };
    while (true) {
        __syncthreads();
        if (tid == 0) {
            // item_off (optional): heavy rows are cut into chunks of A entries, one work item each
            const int item = atomicAdd(work_counter, 1);
            int ri = item < nbin ? item : -1;
            if (item_off) {
                ri = -1;
                if (item < item_off[nbin]) {
                    int lo = 0, hi = nbin;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (item_off[mid] <= item)
                            lo = mid;
                        else
                            hi = mid;
                    }
                    ri = lo;
                }
            }
            s_idx = ri;
            s_item = item;
            s_count = 0;
        }
        __syncthreads();
        const int idx = s_idx;
        if (idx < 0)
            break;
        const int32_t row = rows[idx];
        int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
        int nch = 1;
        if (item_off) {
            nch = item_off[idx + 1] - item_off[idx];
            if (nch > 1) {
                const int chunk = s_item - item_off[idx];
                const int64_t len = ae - as;
                ae = as + len * (chunk + 1) / nch;
                as = as + len * chunk / nch;
            }
        }
        for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
            const int32_t j = A.ci[jj];
            const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
            for (int64_t kk = bs + lane; kk < be; kk += 32) {
                const int32_t k = B.ci[kk];
                if constexpr (BYTES)
                    by[k] = 1;
                else
                    atomicOr(&bm[k >> 5], 1u << (k & 31));
            }
        }
        __syncthreads();
        int count = 0;
        unsigned *dst = keep ? keep + (size_t)idx * n_words : nullptr;
        if (nch > 1) {
            // OR this chunk's columns into the row's (zero-initialised) kept bitmap; k_sym_finish counts it
            for (int i = tid; i < n_words; i += THREADS) {
                const unsigned bits = take_word(i);
                if (bits)
                    atomicOr(&dst[i], bits);
            }
            continue;
        }
        for (int i = tid; i < n_words; i += THREADS) {
            const unsigned bits = take_word(i);
            if (dst)
                dst[i] = bits;
            count += __popc(bits);
        }
        count = warp_sum(count);
        if (lane == 0 && count)
            atomicAdd(&s_count, count);
        __syncthreads();
        if (tid == 0) {
            row_nnz[row] = s_count;
            if (keep_slot)
                keep_slot[row] = idx;
        }
    }
}

// rows that were symbolically processed in chunks: nnz = popcount of the OR-ed bitmap
__global__ void __launch_bounds__(128)
k_sym_finish(const int32_t *__restrict__ rows, const int *__restrict__ item_off, int n_words,
             const unsigned *__restrict__ keep, int32_t *__restrict__ row_nnz, int32_t *__restrict__ keep_slot)
{
    __shared__ int s_count;
    const int idx = blockIdx.x;
    if (item_off[idx + 1] - item_off[idx] <= 1)
        return;
    if (threadIdx.x == 0)
        s_count = 0;
    __syncthreads();
    const unsigned *bm = keep + (size_t)idx * n_words;
    int count = 0;
    for (int i = threadIdx.x; i < n_words; i += 128)
        count += __popc(bm[i]);
    count = warp_sum(count);
    if ((threadIdx.x & 31) == 0 && count)
        atomicAdd(&s_count, count);
    __syncthreads();
    if (threadIdx.x == 0) {
        row_nnz[rows[idx]] = s_count;
        keep_slot[rows[idx]] = idx;
    }
}

// ------------------------------------------------------------ numeric: hashes
// One warp per row, nnz_i <= SLOTS/2.  Columns are ranked by counting (tiny rows).
template <int SLOTS>
__global__ void __launch_bounds__(256) k_num_warp(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin,
                                                  const int64_t *__restrict__ c_rp, int32_t *__restrict__ c_ci,
                                                  double *__restrict__ c_vs, int both_f32)
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
    const int32_t row = rows[idx];
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    for (int64_t jj = as; jj < ae; jj++) {
        const int32_t j = A.ci[jj];
        const double av = ld_val(A.vs, A.vk, jj);
        const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
        for (int64_t kk = bs + lane; kk < be; kk += 32) {
            const int32_t k = B.ci[kk];
            const double p = product(av, ld_val(B.vs, B.vk, kk), both_f32);
            unsigned h = hash_col(k, SLOTS - 1);
            while (true) {
                const int32_t old = atomicCAS(&keys[w][h], EMPTY_KEY, k);
                if (old == EMPTY_KEY || old == k) {
                    atomicAdd(&vals[w][h], p);
                    break;
                }
                h = (h + 1) & (SLOTS - 1);
            }
        }
    }
    __syncwarp();
    const int64_t out = c_rp[row];
    for (int i = lane; i < SLOTS; i += 32) {
        const int32_t k = keys[w][i];
        if (k != EMPTY_KEY) {
            int rank = 0;
            for (int j = 0; j < SLOTS; j++)
                rank += keys[w][j] < k;
            c_ci[out + rank] = k;
            c_vs[out + rank] = vals[w][i];
        }
    }
}

// One CTA per row, nnz_i <= SLOTS/2.  After the hash accumulation the occupied slots are
// DISTRIBUTION-sorted: the (distinct) keys are dealt into NBUCK buckets by a monotone
// linear map of [kmin, kmax], a block scan turns bucket counts into offsets, the entries
// are scattered straight into the output row, and every bucket (about one entry on
// average) is finished by an insertion sort -- O(n) instead of a bitonic sort of the whole
// table.  Badly clustered keys (a bucket of more than NUM_BUCKET_MAX entries) fall back to
// the bitonic sort in shared memory.
constexpr int NUM_BUCKET_MAX = 24;

template <int SLOTS, int THREADS, int NBUCK>
__global__ void __launch_bounds__(THREADS) k_num_cta(MatView A, MatView B, const int32_t *__restrict__ rows,
                                                     const int64_t *__restrict__ c_rp, int32_t *__restrict__ c_ci,
                                                     double *__restrict__ c_vs, int both_f32)
{
    static_assert(NBUCK % THREADS == 0, "one scan pass");
    extern __shared__ __align__(16) unsigned char s_raw[];
    double *vals = reinterpret_cast<double *>(s_raw);
    int32_t *keys = reinterpret_cast<int32_t *>(s_raw + sizeof(double) * SLOTS);
    unsigned *cnt = reinterpret_cast<unsigned *>(s_raw + (sizeof(double) + sizeof(int32_t)) * SLOTS);
    __shared__ int s_min, s_max, s_big;
    __shared__ int s_wt[33];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
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
    }
    __syncthreads();
    const int32_t row = rows[blockIdx.x];
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    int kmin = INT32_MAX, kmax = -1;
    for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
        const int32_t j = A.ci[jj];
        const double av = ld_val(A.vs, A.vk, jj);
        const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
        for (int64_t kk = bs + lane; kk < be; kk += 32) {
            const int32_t k = B.ci[kk];
            const double p = product(av, ld_val(B.vs, B.vk, kk), both_f32);
            kmin = min(kmin, k);
            kmax = max(kmax, k);
            unsigned h = hash_col(k, SLOTS - 1);
            while (true) {
                const int32_t old = atomicCAS(&keys[h], EMPTY_KEY, k);
                if (old == EMPTY_KEY || old == k) {
                    atomicAdd(&vals[h], p);
                    break;
                }
                h = (h + 1) & (SLOTS - 1);
            }
        }
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0 && kmax >= 0) {
        atomicMin(&s_min, kmin);
        atomicMax(&s_max, kmax);
    }
    __syncthreads();
    const int64_t out = c_rp[row];
    const int nz = (int)(c_rp[row + 1] - out);
    kmin = s_min;
    const uint64_t range = (uint64_t)(s_max - kmin) + 1;
    const uint64_t scale = ((uint64_t)NBUCK << 32) / range;  // bucket = (k - kmin) * NBUCK / range, monotone in k
    // bucket counts
    for (int i = tid; i < SLOTS; i += THREADS) {
        const int32_t k = keys[i];
        if (k != EMPTY_KEY)
            atomicAdd(&cnt[(unsigned)(((uint64_t)(k - kmin) * scale) >> 32)], 1u);
    }
    __syncthreads();
    // exclusive scan of the counts (NBUCK/THREADS consecutive buckets per thread)
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
        // scatter into the output row; cnt[b] becomes the END of bucket b
        for (int i = tid; i < SLOTS; i += THREADS) {
            const int32_t k = keys[i];
            if (k != EMPTY_KEY) {
                const unsigned b = (unsigned)(((uint64_t)(k - kmin) * scale) >> 32);
                const unsigned pos = atomicAdd(&cnt[b], 1u);
                c_ci[out + pos] = k;
                c_vs[out + pos] = vals[i];
            }
        }
        __syncthreads();
        // finish every bucket with an insertion sort (columns are distinct)
        for (int b = tid; b < NBUCK; b += THREADS) {
            const int s0 = b ? (int)cnt[b - 1] : 0, e0 = (int)cnt[b];
            for (int i = s0 + 1; i < e0; i++) {
                const int32_t k = c_ci[out + i];
                const double v = c_vs[out + i];
                int j = i - 1;
                while (j >= s0 && c_ci[out + j] > k) {
                    c_ci[out + j + 1] = c_ci[out + j];
                    c_vs[out + j + 1] = c_vs[out + j];
                    j--;
                }
                c_ci[out + j + 1] = k;
                c_vs[out + j + 1] = v;
            }
        }
        return;
    }
    // fallback: bitonic sort of (key, val) over all SLOTS
    for (int k = 2; k <= SLOTS; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < SLOTS / 2; t += THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const int32_t ki = keys[i], kp = keys[p];
                if ((ki > kp) == up) {
                    keys[i] = kp;
                    keys[p] = ki;
                    const double vi = vals[i];
                    vals[i] = vals[p];
                    vals[p] = vi;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < nz; i += THREADS) {
        c_ci[out + i] = keys[i];
        c_vs[out + i] = vals[i];
    }
}

// ---------------------------------------------------- numeric: dense (heavy rows)
// Dense float64 accumulator + column bitmap per persistent CTA; the sweep walks the
// bitmap in column order, so the row comes out sorted, and zeroes what it reads.
//   SMEM_ACC   the accumulator covers a WINDOW of `win` columns in shared memory and the
//              row is done in ceil(ncols/win) passes over its products (one pass when the
//              whole row fits); shared-memory float64 CAS adds were measured ~2x faster
//              than L2 RED.F64 here.
//   otherwise  accumulator and bitmap are this CTA's slice of zeroed global scratch (stays
//              L2-resident on B200); used when ncols would need too many passes.
// If the symbolic phase kept the row's bitmap (`keep`), the numeric phase issues ONE atomic
// per product (the add) instead of two (add + bitmap OR).
template <bool SMEM_ACC, int THREADS>
__global__ void __launch_bounds__(THREADS) k_num_dense(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin,
                                                       const int64_t *__restrict__ c_rp, int32_t *__restrict__ c_ci,
                                                       double *__restrict__ c_vs, int both_f32, double *__restrict__ gacc,
                                                       unsigned *__restrict__ gbm, int n_cols, int win,
                                                       int *__restrict__ work_counter, const unsigned *__restrict__ keep,
                                                       const int32_t *__restrict__ keep_slot)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_idx;
    __shared__ int s_wt[33];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int win_words = (win + 31) >> 5, n_words = (n_cols + 31) >> 5;
    double *acc = SMEM_ACC ? reinterpret_cast<double *>(s_raw) : gacc + (size_t)blockIdx.x * n_cols;
    unsigned *bm = SMEM_ACC ? reinterpret_cast<unsigned *>(s_raw + sizeof(double) * (size_t)win)
                            : gbm + (size_t)blockIdx.x * n_words;
    if (SMEM_ACC) {
        for (int i = tid; i < win; i += THREADS)
            acc[i] = 0.0;
        for (int i = tid; i < win_words; i += THREADS)
            bm[i] = 0;
    }
    while (true) {
        __syncthreads();
        if (tid == 0)
            s_idx = atomicAdd(work_counter, 1);
        __syncthreads();
        const int idx = s_idx;
        if (idx >= nbin)
            break;
        const int32_t row = rows[idx];
        const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
        const unsigned *kept = keep ? keep + (size_t)keep_slot[row] * n_words : nullptr;
        int64_t out = c_rp[row];
        for (int c0 = 0; c0 < n_cols; c0 += win) {  // win is a multiple of 32; one pass when !SMEM_ACC
            const int c1 = min(c0 + win, n_cols);
            for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
                const int32_t j = A.ci[jj];
                const double av = ld_val(A.vs, A.vk, jj);
                const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
                for (int64_t kk = bs + lane; kk < be; kk += 32) {
                    const int32_t k = B.ci[kk];
                    if (k >= c0 && k < c1) {
                        const double p = product(av, ld_val(B.vs, B.vk, kk), both_f32);
                        const int kl = k - c0;
                        if (!kept)
                            atomicOr(&bm[kl >> 5], 1u << (kl & 31));
                        atomicAdd(&acc[kl], p);
                    }
                }
            }
            __syncthreads();
            const int nw = (c1 - c0 + 31) >> 5;
            for (int base = 0; base < nw; base += THREADS) {
                const int i = base + tid;
                unsigned bits = 0;
                if (i < nw)
                    bits = kept ? kept[(c0 >> 5) + i] : (SMEM_ACC ? bm[i] : __ldcg(&bm[i]));
                int tot;
                const int off = block_exclusive_scan<int>(__popc(bits), s_wt, tot);
                if (bits) {
                    if (!kept) {
                        if (SMEM_ACC)
                            bm[i] = 0;
                        else
                            __stcg(&bm[i], 0u);
                    }
                    int64_t o = out + off;
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int kl = i * 32 + b;
                        c_ci[o] = c0 + kl;
                        if (SMEM_ACC) {
                            c_vs[o] = acc[kl];
                            acc[kl] = 0.0;
                        } else {
                            c_vs[o] = __ldcg(&acc[kl]);
                            __stcg(&acc[kl], 0.0);
                        }
                        o++;
                    }
                }
                out += tot;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------- numeric: dense, owner-computes (no atomics)
// For heavy rows when B's rows are strictly increasing in column (always true for the
// transposed operand of A*B^T) and the symbolic bitmaps were kept.  The accumulator window
// is cut into OWN_NW column ranges of about equal B-entry mass, one per warp, and every B row
// is pre-split at those boundaries (k_own_split), so warp w only ever touches its own range:
// plain LDS / DMUL / DADD / STS, no atomics, and every output element is summed in A-row
// order -- the reference's own order (multiply.py:111-120), so values are bit-identical.
// Lanes fetch the metadata of 32 A entries at once; the B pieces of OWN_DEPTH entries are in
// flight before the first is consumed.
constexpr int OWN_NW = 16, OWN_DEPTH = 16;
constexpr int OWN_MAX_CHUNKS = 64;      // slices of one heavy row (A entries) handed to different CTAs
constexpr int DENSE_THREADS = 512;
constexpr int DENSE_WIN = 26624;                      // max columns per shared-memory accumulator window (208 KB)
constexpr int DENSE_MAX_PASSES = 4;                   // beyond that: global scratch, one pass

__global__ void k_col_hist(const int32_t *__restrict__ ci, int64_t nnz, int *__restrict__ cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz)
        atomicAdd(&cnt[ci[i]], 1);
}

// is some row of B not strictly increasing in column?
__global__ void k_rows_not_strict(MatView B, int *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= B.nnz)
        return;
    if (B.ci[i - 1] >= B.ci[i]) {
        // is i a row start?
        int64_t lo = 0, hi = (int64_t)B.nrows + 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ld_rp(B.rp, B.rp64, mid) < i)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (!(lo <= B.nrows && ld_rp(B.rp, B.rp64, lo) == i))
            *flag = 1;
    }
}

// bounds[q*nwarps + w]: first column of warp w's range in window q (mass-balanced inside the window)
__global__ void k_own_bounds(const int64_t *__restrict__ cum, int n, int win, int passes, int nwarps,
                             int32_t *__restrict__ bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > passes * nwarps)
        return;
    if (i == passes * nwarps) {
        bounds[i] = n;
        return;
    }
    const int q = i / nwarps, w = i % nwarps;
    const int c0 = q * win, c1 = min(c0 + win, n);
    if (w == 0) {
        bounds[i] = c0;
        return;
    }
    const int64_t target = cum[c0] + (cum[c1] - cum[c0]) * w / nwarps;
    bounds[i] = (int32_t)lower_bound_rp(cum, (int64_t)c0, (int64_t)c1, target);
}

// split[j*(nb+1) + k]: offset inside B row j of its first entry with column >= bounds[k]
__global__ void k_own_split(MatView B, const int32_t *__restrict__ bounds, int nb, int32_t *__restrict__ split)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B.nrows * (nb + 1))
        return;
    const int64_t j = i / (nb + 1);
    const int k = (int)(i % (nb + 1));
    const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, j + 1);
    const int32_t key = bounds[k];
    int64_t lo = bs, hi = be;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (B.ci[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    split[i] = (int32_t)(lo - bs);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
k_num_owner(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin, const int64_t *__restrict__ c_rp,
            int32_t *__restrict__ c_ci, double *__restrict__ c_vs, int both_f32, int n_cols, int win, int passes,
            int *__restrict__ work_counter, const unsigned *__restrict__ keep, const int32_t *__restrict__ keep_slot,
            const int32_t *__restrict__ split, const int *__restrict__ item_off, const int *__restrict__ chunk_base,
            int nitems, double *__restrict__ partial)
{
    constexpr int THREADS = NW * 32;
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_idx, s_ri;
    __shared__ int s_wt[33];
    double *acc = reinterpret_cast<double *>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n_words = (n_cols + 31) >> 5, nb = passes * NW;
    for (int i = tid; i < win; i += THREADS)
        acc[i] = 0.0;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            // work item -> (row of the list, chunk of its A entries): last ri with item_off[ri] <= item
            const int item = atomicAdd(work_counter, 1);
            int lo = 0, hi = nbin;
            if (item < nitems) {
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (item_off[mid] <= item)
                        lo = mid;
                    else
                        hi = mid;
                }
            }
            s_idx = item;
            s_ri = lo;
        }
        __syncthreads();
        if (s_idx >= nitems)
            break;
        const int ri = s_ri;
        const int chunk = s_idx - item_off[ri], nch = item_off[ri + 1] - item_off[ri];
        const int32_t row = rows[ri];
        int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
        if (nch > 1) {
            // a row with too many products for one CTA: this item takes one slice of its A entries and
            // leaves its accumulator windows in `partial`; k_own_combine adds the slices in order
            const int64_t len = ae - as;
            ae = as + len * (chunk + 1) / nch;
            as = as + len * chunk / nch;
        }
        const unsigned *kept = keep + (size_t)keep_slot[row] * n_words;
        int64_t out = c_rp[row];
        for (int q = 0; q < passes; q++) {
            const int c0 = q * win, c1 = min(c0 + win, n_cols);
            const int kidx = q * NW + w;
            // metadata of the next 32 A entries is fetched while the current 32 are consumed
            int32_t jn = 0;
            double avn = 0.0;
            if (as + lane < ae) {
                jn = A.ci[as + lane];
                avn = ld_val(A.vs, A.vk, as + lane);
            }
            for (int64_t base = as; base < ae; base += 32) {
                const bool valid = base + lane < ae;
                const int32_t j = jn;
                const double av = avn;
                if (base + 32 + lane < ae) {
                    jn = A.ci[base + 32 + lane];
                    avn = ld_val(A.vs, A.vk, base + 32 + lane);
                }
                int64_t start = 0;
                int len = 0;
                if (valid) {
                    const int32_t *sp = split + (size_t)j * (nb + 1) + kidx;
                    const int s_off = sp[0];
                    len = sp[1] - s_off;
                    start = ld_rp(B.rp, B.rp64, j) + s_off;
                }
                // The work of these 32 pieces is cut into UNITS of <= 32 consecutive B entries,
                // enumerated in piece order (an inclusive warp scan of the unit counts locates the
                // piece of unit i), so long pieces pipeline like short ones and every accumulator
                // still sees its contributions in A-row order.
                const int cu = (len + 31) >> 5;
                int incl = cu;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o)
                        incl += t;
                }
                const int excl = incl - cu;
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                for (int i0 = 0; i0 < total; i0 += OWN_DEPTH) {
                    int col[OWN_DEPTH];
                    double val[OWN_DEPTH];
                    int meta[OWN_DEPTH];  // (piece lane << 8) | entries in the unit
#pragma unroll
                    for (int d = 0; d < OWN_DEPTH; d++) {
                        const int i = i0 + d;
                        col[d] = c0;
                        val[d] = 0.0;
                        meta[d] = 0;
                        if (i < total) {
                            const int u = __popc(__ballot_sync(0xffffffffu, excl <= i)) - 1;
                            const int off = (i - __shfl_sync(0xffffffffu, excl, u)) << 5;
                            const int64_t st = __shfl_sync(0xffffffffu, start, u) + off;
                            const int ln = min(32, __shfl_sync(0xffffffffu, len, u) - off);
                            meta[d] = (u << 8) | ln;
                            if (lane < ln) {
                                col[d] = B.ci[st + lane];
                                val[d] = ld_val(B.vs, B.vk, st + lane);
                            }
                        }
                    }
#pragma unroll
                    for (int d = 0; d < OWN_DEPTH; d++) {
                        if (i0 + d < total) {
                            const double avu = __shfl_sync(0xffffffffu, av, meta[d] >> 8);
                            if (lane < (meta[d] & 255)) {
                                const int kl = col[d] - c0;
                                acc[kl] = __dadd_rn(acc[kl], product(avu, val[d], both_f32));
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            __syncthreads();
            if (nch > 1) {
                double *dst = partial + ((size_t)(chunk_base[ri] + chunk) * passes + q) * win;
                for (int i = tid; i < win; i += THREADS) {
                    dst[i] = acc[i];
                    acc[i] = 0.0;
                }
                __syncthreads();
                continue;
            }
            // sweep the window in column order with the bitmap kept by the symbolic phase
            const int nw = (c1 - c0 + 31) >> 5;
            for (int wb = 0; wb < nw; wb += THREADS) {
                const int i = wb + tid;
                unsigned bits = i < nw ? kept[(c0 >> 5) + i] : 0u;
                int tot;
                const int off = block_exclusive_scan<int>(__popc(bits), s_wt, tot);
                int64_t o = out + off;
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int kl = i * 32 + b;
                    c_ci[o] = c0 + kl;
                    c_vs[o] = acc[kl];
                    acc[kl] = 0.0;
                    o++;
                }
                out += tot;
            }
            __syncthreads();
        }
    }
}

// chunks per heavy row: rows with more than `chunk_prod` products are cut into slices of A entries
__global__ void k_own_items(MatView A, const int32_t *__restrict__ rows, int n, const int64_t *__restrict__ prod,
                            int64_t chunk_prod, int *__restrict__ nchunk, int *__restrict__ nsplit, int *__restrict__ last_split)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int32_t row = rows[i];
    const int64_t len = ld_rp(A.rp, A.rp64, (int64_t)row + 1) - ld_rp(A.rp, A.rp64, row);
    int64_t c = (prod[row] + chunk_prod - 1) / chunk_prod;
    c = max((int64_t)1, min(min(c, len), (int64_t)OWN_MAX_CHUNKS));
    nchunk[i] = (int)c;
    nsplit[i] = c > 1 ? (int)c : 0;
    if (c > 1)
        atomicMax(last_split, i + 1);  // the list is LPT-ordered, so this stays near the number of chunked rows
}

// One CTA per (row of the list, window): rows that were cut into chunks get their partial windows
// added in chunk order (deterministic) and are emitted through the kept bitmap like any other row.
__global__ void __launch_bounds__(256)
k_own_combine(const int32_t *__restrict__ rows, int nbin, const int *__restrict__ item_off,
              const int *__restrict__ chunk_base, const double *__restrict__ partial, int n_cols, int win, int passes,
              const unsigned *__restrict__ keep, const int32_t *__restrict__ keep_slot, const int64_t *__restrict__ c_rp,
              int32_t *__restrict__ c_ci, double *__restrict__ c_vs)
{
    __shared__ int s_wt[33];
    const int ri = blockIdx.x / passes, q = blockIdx.x % passes;
    const int nch = item_off[ri + 1] - item_off[ri];
    if (nch <= 1)
        return;
    const int tid = threadIdx.x;
    const int32_t row = rows[ri];
    const int n_words = (n_cols + 31) >> 5;
    const unsigned *kept = keep + (size_t)keep_slot[row] * n_words;
    const int c0 = q * win, c1 = min(c0 + win, n_cols);
    // output position of the window's first entry: the row's kept columns below c0
    int below = 0;
    for (int i = tid; i < (c0 >> 5); i += 256)
        below += __popc(kept[i]);
    int tot;
    block_exclusive_scan<int>(below, s_wt, tot);
    int64_t out = c_rp[row] + tot;
    __syncthreads();
    const double *p0 = partial + ((size_t)chunk_base[ri] * passes + q) * win;
    const size_t cstride = (size_t)passes * win;
    const int nw = (c1 - c0 + 31) >> 5;
    for (int wb = 0; wb < nw; wb += 256) {
        const int i = wb + tid;
        unsigned bits = i < nw ? kept[(c0 >> 5) + i] : 0u;
        const int off = block_exclusive_scan<int>(__popc(bits), s_wt, tot);
        int64_t o = out + off;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int kl = i * 32 + b;
            double v = 0.0;
            for (int ch = 0; ch < nch; ch++)
                v = __dadd_rn(v, p0[ch * cstride + kl]);
            c_ci[o] = c0 + kl;
            c_vs[o] = v;
            o++;
        }
        out += tot;
        __syncthreads();
    }
}

// ------------------------------------------- numeric: dense, 64-bit fixed point (native atomics)
// Shared-memory float64 atomicAdd is a CAS loop (ATOMS.CAST.SPIN: 2.0 updates/clk/SM on random
// columns, 0.5 with the hot columns of real rating data), the owner-computes kernel above avoids
// atomics but leaves most lanes idle (0.24 products/clk/SM).  Native 32-bit ATOMS.ADD runs at ~10
// updates/clk/SM (tools/micro/atoms.cu), so when the data allow it the accumulator is a 64-bit
// two's-complement FIXED-POINT number per column, kept as two 32-bit words: the low word is added
// with an atomicAdd that returns the old value (carry-out = unsigned overflow), the high word gets
// its part plus the carry (4.3-5.1 updates/clk/SM).  Integer addition is associative: the result
// does not depend on the order of the products, on how rows are chunked or on the number of GPUs.
//
// Scale of row i: 2^(e0 - hb_i) with hb_i = ceil(log2(len_i + 1)) bits of headroom for the at most
// len_i terms of one output element and 2^e0 * max|a*b| <= 2^62.  Every term is rounded once to a
// multiple of 2^-(e0-hb_i), i.e. with an absolute error of at most 2^(hb-63) * max|a*b|.  spgemm_run only
// takes this path when all values are finite and max|a*b| / min|a*b| <= 2^(26-hb): then every term's rounding
// error is below 2^-37 of that term's own magnitude, and the error of an output element is below
// 2^-35 * sum_k |a_ik||b_kj| = 2.9e-11 of the magnitude that was summed (rtol 1e-10 of the parity contract,
// stated the way a re-ordered floating-point sum is bounded).  For non-negative operands that is the relative
// error of the element itself; signs are welcome (two's-complement words), a wide dynamic range is not --
// mean-centred ratings have values arbitrarily close to 0, so min|a*b| is tiny -- and takes the owner kernel.
// Equilibration (round 2).  A fixed scale per row only covers operands whose magnitudes span a few bits, and the
// interesting inputs of A*B^T do not: unit-normalised rows (cosine similarity) differ by the rows' norms,
// mean-centred ratings come arbitrarily close to 0.  So the operands are first brought to a common magnitude
// with exact power-of-two factors: ea[i] = frexp exponent of the largest |a_ik| of A's row i, eb[j] the same for
// B's COLUMN j; the kernel multiplies a_ik * 2^-ea[i] (folded into the row's scale) with b'_kj = b_kj * 2^-eb[j]
// (a scaled copy of B's values), so every product is below 1 in magnitude, accumulates T = rn(p' * 2^(62 - hb_i))
// and the sweep multiplies the sum by 2^(ea[i] + eb[j] - 62 + hb_i).  A product with |T| < 2^35 -- one whose
// rounding error would exceed 2^-36 of its own magnitude -- is NOT added: it goes, exactly, to a side list
// (row, column, p') and k_fix_tiny adds it to the finished element in float64.  Every other term carries an error
// below 2^-36 of its magnitude, so every output element is within 2^-35 * sum_k |a_ik||b_kj| whatever the values
// are; the gate only asks for finite operands whose exponents stay within +-400.
__global__ void __launch_bounds__(256) k_row_expo(MatView A, int *__restrict__ ea)
{
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= A.nrows)
        return;
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    int e = INT32_MIN;
    for (int64_t k = as + lane; k < ae; k += 32) {
        const double v = ld_val(A.vs, A.vk, k);
        if (v != 0.0)
            e = max(e, (int)((__double_as_longlong(v) >> 52) & 0x7ff) - 1022);   // |v| < 2^e (Inf/NaN: 1025)
    }
    e = __reduce_max_sync(0xffffffffu, e);
    if (lane == 0)
        ea[row] = e;
}
__global__ void __launch_bounds__(256) k_col_expo(MatView B, int *__restrict__ eb)
{
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < B.nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const double v = ld_val(B.vs, B.vk, k);
        if (v != 0.0) {
            const int e = (int)((__double_as_longlong(v) >> 52) & 0x7ff) - 1022;
            int *p = &eb[B.ci[k]];
            if (__ldcg(p) < e)   // (a plain read first: after a column's first few entries almost no atomic is left --
                atomicMax(p, e);  //  the popular columns of rating data otherwise serialise 10^5 atomics on one address)
        }
    }
}
// bit patterns of the smallest non-zero and the largest finite |v| (mm[0], mm[1]): a narrow range lets the
// fixed-point kernel run without its side-list test
__global__ void __launch_bounds__(256) k_val_range(const void *__restrict__ vs, int vk, int64_t nnz, unsigned long long *__restrict__ mm)
{
    unsigned long long lo = ~0ull, hi = 0ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs(ld_val(vs, vk, i));
        if (v != 0.0) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(v);
            lo = min(lo, b);
            hi = max(hi, b);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && hi) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}
__global__ void __launch_bounds__(256) k_fill_i32(int *__restrict__ p, int64_t n, int v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}
// smallest and largest exponent of the non-empty rows / columns
__global__ void __launch_bounds__(256) k_expo_range(const int *__restrict__ e, int n, int *__restrict__ mm)
{
    int lo = INT32_MAX, hi = INT32_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = e[i];
        if (v != INT32_MIN) {
            lo = min(lo, v);
            hi = max(hi, v);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && hi != INT32_MIN) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }   // |e| < 1023
template <typename T>
__global__ void __launch_bounds__(256) k_scale_cols(const int32_t *__restrict__ ci, const void *__restrict__ vs, int vk, int64_t nnz,
                                                     const int *__restrict__ eb, T *__restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) {
        const int e = eb[ci[k]];
        out[k] = (T)(ld_val(vs, vk, k) * (e == INT32_MIN ? 1.0 : pow2(-e)));   // exact: a power of two
    }
}

// side list of the fixed-point kernel: warps reserve blocks of 32 entries (one global atomic per block)
struct TinyEnt {
    int32_t row, col;
    double p;   // the equilibrated product a' * b' (numba's product, scaled by an exact power of two)
};
struct TinyList {
    TinyEnt *buf;
    unsigned long long *count;   // entries reserved (blocks of 32; may run past cap: the host checks)
    unsigned long long cap;
};
// all 32 lanes call this together; `has` lanes append one entry each
__device__ __forceinline__ void tiny_append(const TinyList &t, bool has, int32_t row, int32_t col, double p,
                                            unsigned long long *s_base, int *s_used, int warp, int lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (!m)
        return;
    const int k = __popc(m);
    int used = s_used[warp];
    if (used + k > 32) {   // a new block; what is left of the old one stays marked empty
        if (lane == 0) {
            s_base[warp] = atomicAdd(t.count, 32ull);
            s_used[warp] = 0;
        }
        __syncwarp();
        const unsigned long long b = s_base[warp];
        if (b + lane < t.cap)
            t.buf[b + lane].row = -1;
        __syncwarp();
        used = 0;
    }
    const unsigned long long idx = s_base[warp] + used + __popc(m & lanemask_lt());
    if (has && idx < t.cap)
        t.buf[idx] = TinyEnt{row, col, p};
    __syncwarp();
    if (lane == 0)
        s_used[warp] = used + k;
    __syncwarp();
}

__device__ __forceinline__ int headroom_bits(int64_t len)
{
    return 64 - __clzll((unsigned long long)len);  // ceil(log2(len + 1))
}

constexpr int FIX_MAX_PASSES = 4;

// SIDE = false: the caller has checked that no product can fall below the grid (max|a*b| / min|a*b| <= 2^(25-hb)):
// the side-list test is compiled out
template <int THREADS, bool SIDE>
__global__ void __launch_bounds__(THREADS, 1)
k_num_fixed(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin, const int64_t *__restrict__ c_rp,
            int32_t *__restrict__ c_ci, double *__restrict__ c_vs, int both_f32, int n_cols, int win, int passes,
            int *__restrict__ work_counter, const unsigned *__restrict__ keep, const int32_t *__restrict__ keep_slot,
            const int32_t *__restrict__ psplit, const int *__restrict__ item_off, const int *__restrict__ chunk_base,
            int nitems, long long *__restrict__ partial, const int *__restrict__ ea, const int *__restrict__ eb, int eb_u,
            TinyList tiny)
{
    static_assert(THREADS * 2 * 32 >= DENSE_WIN, "two bitmap words per thread must cover a window");
    static_assert(THREADS % 32 == 0 && THREADS <= 1024, "whole warps");
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_idx, s_ri;
    __shared__ int s_next[FIX_MAX_PASSES];
    __shared__ int s_wt[33];
    __shared__ unsigned s_bits[2 * THREADS];
    __shared__ int s_wpre[2 * THREADS];
    __shared__ unsigned long long s_tbase[THREADS / 32];
    __shared__ int s_tused[THREADS / 32];
    unsigned *slo = reinterpret_cast<unsigned *>(s_raw), *shi = slo + win;
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int NWARP = THREADS / 32;
    if (lane == 0)
        s_tused[tid >> 5] = 32;   // no block yet
    const int n_words = (n_cols + 31) >> 5;
    for (int i = tid; i < 2 * win; i += THREADS)
        slo[i] = 0u;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            const int item = atomicAdd(work_counter, 1);
            int lo = 0, hi = nbin;
            if (item < nitems) {
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (item_off[mid] <= item)
                        lo = mid;
                    else
                        hi = mid;
                }
            }
            s_idx = item;
            s_ri = lo;
        }
        if (tid < FIX_MAX_PASSES)
            s_next[tid] = 0;
        __syncthreads();
        if (s_idx >= nitems)
            break;
        const int ri = s_ri;
        const int chunk = s_idx - item_off[ri], nch = item_off[ri + 1] - item_off[ri];
        const int32_t row = rows[ri];
        int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
        const int hbits = headroom_bits(ae - as);  // the WHOLE row's length: all chunks share one scale
        const int ea_i = ea[row];                  // (a heavy row has products, but all its values may be 0)
        // eb == nullptr: all of B's columns share the exponent eb_u (raw or centred ratings): B is read as it is
        // and 2^-eb_u is folded into the row's factor
        const int ea_f = (ea_i == INT32_MIN ? 0 : ea_i) + (eb ? 0 : eb_u);
        const int sh = 62 - hbits - ea_f;
        const double scale = pow2(sh);             // a_ik * 2^sh: equilibrated and scaled at once (float64 operands)
        const double scale_f = pow2(62 - hbits), eq_a = pow2(-ea_f);   // float32 operands: product first
        const double inv_scale = pow2(-sh);
        if (nch > 1) {
            const int64_t len = ae - as;
            ae = as + len * (chunk + 1) / nch;
            as = as + len * chunk / nch;
        }
        const unsigned *kept = keep + (size_t)keep_slot[row] * n_words;
        int64_t out = c_rp[row];
        // batches of A entries are handed to the warps dynamically (pieces differ a lot in length)
        const int bsz = (int)max((int64_t)1, min((int64_t)32, (ae - as + 4 * NWARP - 1) / (4 * NWARP)));
        const int nbatch = (int)((ae - as + bsz - 1) / bsz);
        for (int q = 0; q < passes; q++) {
            const int c0 = q * win, c1 = min(c0 + win, n_cols);
            const int nw = (c1 - c0 + 31) >> 5;
            // the window's kept-bitmap words are fetched now and used after the accumulation
            const unsigned kb0 = tid < nw ? kept[(c0 >> 5) + tid] : 0u;
            const unsigned kb1 = THREADS + tid < nw ? kept[(c0 >> 5) + THREADS + tid] : 0u;
            while (true) {
                int bi = 0;
                if (lane == 0)
                    bi = atomicAdd(&s_next[q], 1);
                bi = __shfl_sync(0xffffffffu, bi, 0);
                if (bi >= nbatch)
                    break;
                const int64_t base = as + (int64_t)bi * bsz;
                int64_t bs = 0;
                int len = 0;
                double av = 0.0;
                if (lane < bsz && base + lane < ae) {
                    const int32_t j = A.ci[base + lane];
                    av = ld_val(A.vs, A.vk, base + lane);
                    const int32_t *sp = psplit + (size_t)j * (passes + 1) + q;
                    const int o0 = sp[0];
                    len = sp[1] - o0;
                    bs = ld_rp(B.rp, B.rp64, j) + o0;
                }
                const int cnt = (int)min((int64_t)bsz, ae - base);
                // One piece after the other, 128 entries at a time: eight independent loads per lane, then the
                // atomics.  No pipelining across pieces: 32 warps per SM cover the load latency, and a piece
                // costs ~15 instructions of bookkeeping (the register-pipelined version this replaces spent ~300
                // per piece: 12.7 G of the kernel's 19.9 G warp instructions at configs[2]).
                for (int t = 0; t < cnt; t++) {
                    const int64_t pbs = __shfl_sync(0xffffffffu, bs, t);
                    const int plen = __shfl_sync(0xffffffffu, len, t);
                    const double pav = __shfl_sync(0xffffffffu, av, t) * (both_f32 ? eq_a : scale);
                    for (int k0 = 0; k0 < plen; k0 += 128) {
                        int col[4];
                        double val[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int k = k0 + u * 32 + lane;
                            col[u] = -1;
                            val[u] = 0.0;
                            if (k < plen) {
                                col[u] = B.ci[pbs + k];
                                val[u] = ld_val(B.vs, B.vk, pbs + k);
                            }
                        }
                        unsigned small = 0;   // bit u: product u is below the accumulator's grid
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            // (float64 operands: the row's scale is a power of two folded into a, exactly; a lane
                            // past the end of the piece has val = 0: T = 0, nothing happens)
                            const double pr = both_f32 ? (double)__fmul_rn((float)pav, (float)val[u]) * scale_f
                                                       : __dmul_rn(pav, val[u]);
                            const long long T = __double2ll_rn(pr);
                            const bool big = !SIDE || fabs(pr) >= 34359738368.0;   // 2^35: rounding error <= 2^-36 |term|
                            if (SIDE) {
                                val[u] = pr;
                                small |= (unsigned)(!big && pr != 0.0) << u;
                            }
                            if (SIDE ? big : col[u] >= 0) {   // (kept this short: two predicated atomics, the four chains overlap)
                                const unsigned tlo = (unsigned)T, thi = (unsigned)((unsigned long long)T >> 32);
                                const unsigned old = atomicAdd(&slo[col[u] - c0], tlo);
                                atomicAdd(&shi[col[u] - c0], thi + ((old + tlo) < old ? 1u : 0u));
                            }
                        }
                        // rounded too coarsely for the accumulator: exact side list (rare; one vote per 128 entries)
                        if (SIDE && __any_sync(0xffffffffu, small != 0)) {
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                tiny_append(tiny, (small >> u) & 1u, row, col[u], val[u], s_tbase, s_tused, tid >> 5, lane);
                        }
                    }
                }
            }
            __syncthreads();
            if (nch > 1) {
                long long *dst = partial + ((size_t)(chunk_base[ri] + chunk) * passes + q) * win;
                for (int i = tid; i < win; i += THREADS) {
                    dst[i] = (long long)(((unsigned long long)shi[i] << 32) | slo[i]);
                    slo[i] = 0u;
                    shi[i] = 0u;
                }
                __syncthreads();
                continue;
            }
            // sweep: one thread per COLUMN, so that a warp writes consecutive output slots; the slot of
            // a column = kept bits below it (per-word prefix from a block scan + popcount inside the word)
            int tot0, tot1 = 0;
            const int ex0 = block_exclusive_scan<int>(__popc(kb0), s_wt, tot0);
            s_bits[tid] = kb0;
            s_wpre[tid] = ex0;
            if (THREADS * 32 < DENSE_WIN) {  // a second word per thread only for the smaller CTAs
                __syncthreads();
                const int ex1 = block_exclusive_scan<int>(__popc(kb1), s_wt, tot1);
                s_bits[THREADS + tid] = kb1;
                s_wpre[THREADS + tid] = tot0 + ex1;
            }
            __syncthreads();
            // (four columns per thread and step: the four loads of the columns' exponents are in flight together)
            for (int col0 = tid; col0 < c1 - c0; col0 += 4 * THREADS) {
                int ecol[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int col = col0 + u * THREADS;
                    ecol[u] = 0;
                    if (eb && col < c1 - c0)
                        ecol[u] = eb[c0 + col];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int col = col0 + u * THREADS;
                    if (col >= c1 - c0)
                        break;
                    const unsigned bits = s_bits[col >> 5];
                    const unsigned bit = 1u << (col & 31);
                    if (bits & bit) {
                        const int64_t o = out + s_wpre[col >> 5] + __popc(bits & (bit - 1u));
                        const long long v = (long long)(((unsigned long long)shi[col] << 32) | slo[col]);
                        __stcs(&c_ci[o], c0 + col);  // streaming: the 12 B/entry result must not evict B from L2
                        double r = (double)v * inv_scale;
                        if (eb)
                            r *= pow2(ecol[u] == INT32_MIN ? 0 : ecol[u]);
                        __stcs(&c_vs[o], r);
                        slo[col] = 0u;
                        shi[col] = 0u;
                    }
                }
            }
            out += tot0 + tot1;
            __syncthreads();
        }
    }
}

// chunked rows of the fixed-point kernel: exact integer sum of the chunks' windows, then the scale
__global__ void __launch_bounds__(256)
k_fix_combine(MatView A, const int32_t *__restrict__ rows, int nbin, const int *__restrict__ item_off,
              const int *__restrict__ chunk_base, const long long *__restrict__ partial, int n_cols, int win, int passes,
              const unsigned *__restrict__ keep, const int32_t *__restrict__ keep_slot, const int64_t *__restrict__ c_rp,
              int32_t *__restrict__ c_ci, double *__restrict__ c_vs, const int *__restrict__ ea, const int *__restrict__ eb, int eb_u)
{
    __shared__ int s_wt[33];
    const int ri = blockIdx.x / passes, q = blockIdx.x % passes;
    const int nch = item_off[ri + 1] - item_off[ri];
    if (nch <= 1)
        return;
    const int tid = threadIdx.x;
    const int32_t row = rows[ri];
    const int ea_i = ea[row];
    const int sh = 62 - headroom_bits(ld_rp(A.rp, A.rp64, (int64_t)row + 1) - ld_rp(A.rp, A.rp64, row)) -
                   (ea_i == INT32_MIN ? 0 : ea_i) - (eb ? 0 : eb_u);
    const double inv_scale = pow2(-sh);
    const int n_words = (n_cols + 31) >> 5;
    const unsigned *kept = keep + (size_t)keep_slot[row] * n_words;
    const int c0 = q * win, c1 = min(c0 + win, n_cols);
    int below = 0;
    for (int i = tid; i < (c0 >> 5); i += 256)
        below += __popc(kept[i]);
    int tot;
    block_exclusive_scan<int>(below, s_wt, tot);
    int64_t out = c_rp[row] + tot;
    __syncthreads();
    const long long *p0 = partial + ((size_t)chunk_base[ri] * passes + q) * win;
    const size_t cstride = (size_t)passes * win;
    const int nw = (c1 - c0 + 31) >> 5;
    for (int wb = 0; wb < nw; wb += 256) {
        const int i = wb + tid;
        unsigned bits = i < nw ? kept[(c0 >> 5) + i] : 0u;
        const int off = block_exclusive_scan<int>(__popc(bits), s_wt, tot);
        int64_t o = out + off;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int kl = i * 32 + b;
            long long v = 0;
            for (int ch = 0; ch < nch; ch++)
                v += p0[ch * cstride + kl];
            c_ci[o] = c0 + kl;
            const int e = eb ? eb[c0 + kl] : 0;
            c_vs[o] = (double)v * inv_scale * pow2(e == INT32_MIN ? 0 : e);
            o++;
        }
        out += tot;
        __syncthreads();
    }
}

// the side list: every entry is added, in float64, to its (finished) output element
__global__ void __launch_bounds__(256)
k_fix_tiny(const TinyEnt *__restrict__ buf, unsigned long long n, MatView A, const int *__restrict__ ea, const int *__restrict__ eb,
           int eb_u, const void *__restrict__ c_rp, int c_rp64, const int32_t *__restrict__ c_ci, double *__restrict__ c_vs)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const TinyEnt t = buf[i];
    if (t.row < 0)
        return;
    const int ea_i = ea[t.row], eb_j = eb ? eb[t.col] : 0;
    const int sh = 62 - headroom_bits(ld_rp(A.rp, A.rp64, (int64_t)t.row + 1) - ld_rp(A.rp, A.rp64, t.row)) -
                   (ea_i == INT32_MIN ? 0 : ea_i) - (eb ? 0 : eb_u);
    int64_t lo = ld_rp(c_rp, c_rp64, t.row), hi = ld_rp(c_rp, c_rp64, (int64_t)t.row + 1);
    while (lo < hi) {   // the column exists: the symbolic pass saw this product
        const int64_t mid = (lo + hi) >> 1;
        if (c_ci[mid] < t.col)
            lo = mid + 1;
        else
            hi = mid;
    }
    atomicAdd(&c_vs[lo], t.p * pow2(-sh) * pow2(eb_j == INT32_MIN ? 0 : eb_j));
}

__global__ void k_fix_bounds(int n, int win, int passes, int32_t *__restrict__ bounds)
{
    const int i = threadIdx.x;
    if (i <= passes)
        bounds[i] = i == passes ? n : min(i * win, n);
}

// sort key for longest-processing-time-first scheduling of the heavy rows: descending products
__global__ void k_lpt_keys(const int32_t *__restrict__ rows, int n, const int64_t *__restrict__ prod, int32_t *__restrict__ keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t p = prod[rows[i]] >> 6;
        keys[i] = 0x7FFFFFFF - (int32_t)(p > 0x7FFFFFFF ? 0x7FFFFFFF : p);
    }
}

template <typename T> __global__ void k_narrow_rp(const int64_t *__restrict__ in, T *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = (T)in[i];
}

// --------------------------------------------------------------- host driver
template <typename T>
static int bin_rows(const T *val, int64_t n, const BinSpec &spec, int counts[NBINS], int offs[NBINS + 1], DevBuf &list,
                    cudaStream_t s)
{
    DevBuf dcnt;
    CSRK_TRY(dcnt.alloc_zero(sizeof(int) * NBINS, s));
    CSRK_LAUNCH((k_bin_count<T>), (unsigned)div_up(n, 256), 256, 0, s, val, n, spec, dcnt.as<int>());
    CSRK_CUDA(cudaMemcpyAsync(counts, dcnt.p, sizeof(int) * NBINS, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    offs[0] = 0;
    for (int b = 0; b < NBINS; b++)
        offs[b + 1] = offs[b] + counts[b];
    CSRK_CUDA(cudaMemcpyAsync(dcnt.p, offs, sizeof(int) * NBINS, cudaMemcpyHostToDevice, s));
    CSRK_TRY(list.alloc(sizeof(int32_t) * (size_t)n, s));
    CSRK_LAUNCH((k_bin_fill<T>), (unsigned)div_up(n, 256), 256, 0, s, val, n, spec, dcnt.as<int>(), list.as<int32_t>());
    return CSRK_OK;
}

template <typename K> static int optin_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        CSRK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return CSRK_OK;
}

// symbolic bins on P_i (products): hash sets of 2x the bound, then the bitmap;
// numeric bins on nnz_i: hash accumulators of 2x the bound, then the dense accumulator.
// The bitmap/dense boundaries coincide (8192) so that every dense-numeric row went
// through the symbolic bitmap kernel and can reuse its stored bitmap.
constexpr int SYM_WARP_SLOTS = 256, SYM_C1 = 2048, SYM_C2 = 8192, SYM_C3 = 16384;
constexpr int NUM_WARP_SLOTS = 128, NUM_C1 = 1024, NUM_C2 = 4096, NUM_C3 = 16384;
constexpr size_t KEEP_BITMAP_BUDGET = (size_t)8 << 30;  // bytes of symbolic bitmaps kept for the numeric phase

static int make_empty_result(csrk_matrix *a, csrk_matrix *b, csrk_matrix **c, cudaStream_t s)
{
    csrk_matrix *m = nullptr;
    CSRK_TRY(matrix_alloc(&m, a->nrows, b->ncols, 0, 0, 8, s));
    cudaError_t e = cudaMemsetAsync(m->rp, 0, ((size_t)a->nrows + 1) * 4, s);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        matrix_destroy(m, s);
        return cuda_fail(e, "empty product", __FILE__, __LINE__);
    }
    m->stat_products = 0;
    m->stat_out_nnz = 0;
    *c = m;
    return CSRK_OK;
}

}  // namespace csrk
#include "spgemm_esc.cuh"
namespace csrk {

__global__ void __launch_bounds__(256) k_esc_zero_rows(const int32_t *__restrict__ rows, int n, const int *__restrict__ bad,
                                                        int32_t *__restrict__ row_nnz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !bad[i])
        row_nnz[rows[i]] = 0;
}

int spgemm_run(csrk_matrix *a, csrk_matrix *b, csrk_matrix **c, cudaStream_t s)
{
    *c = nullptr;
    const int32_t m = a->nrows, n = b->ncols;
    if (m == 0 || a->nnz == 0 || b->nnz == 0)
        return make_empty_result(a, b, c, s);
    const MatView A = view(a), B = view(b);
    const int both_f32 = (a->val_kind == 4 && b->val_kind == 4) ? 1 : 0;
    const int sms = ctx().sm_count;
    const size_t smem_max = ctx().smem_optin;
    const int n_words = (int)div_up(n, 32);

    CSRK_TRACE_MARK("spgemm: enter", s);
    // ---- step 0: products per row
    DevBuf prod, total;
    CSRK_TRY(prod.alloc(sizeof(int64_t) * (size_t)m, s));
    CSRK_TRY(total.alloc_zero(sizeof(unsigned long long) * 4, s));
    DevBuf long_items;   // pieces of rows longer than ROWP_LONG entries: at most nnz / ROWP_LONG + one per such row
    const int max_items = (int)std::min<int64_t>(a->nnz / ROWP_LONG * 2 + 1, INT32_MAX);
    CSRK_TRY(long_items.alloc(sizeof(int64_t) * (size_t)max_items, s));
    CSRK_LAUNCH(k_row_products, (unsigned)div_up((int64_t)m * 32, 256), 256, 0, s, A, B, prod.as<int64_t>(),
                total.as<unsigned long long>(), long_items.as<int64_t>(), max_items);
    if (a->nnz > ROWP_LONG) {
        CSRK_LAUNCH(k_row_products_long, (unsigned)std::min(max_items, sms * 8), 256, 0, s, A, B, prod.as<int64_t>(),
                    total.as<unsigned long long>(), long_items.as<int64_t>(), max_items);
        CSRK_LAUNCH(k_row_products_max, (unsigned)std::min((int)div_up(max_items, 256), sms), 256, 0, s, prod.as<int64_t>(),
                    total.as<unsigned long long>(), long_items.as<int64_t>(), max_items);
    }
    unsigned long long PM[3] = {0, 0, 0};  // total products, products of the heaviest row, longest row (landed by bin_rows' sync)
    CSRK_CUDA(cudaMemcpyAsync(PM, total.p, sizeof PM, cudaMemcpyDeviceToHost, s));

    // ---- step 1: symbolic
    BinSpec sspec{{0, SYM_WARP_SLOTS / 2, SYM_C1 / 2, SYM_C2 / 2, SYM_C3 / 2, INT64_MAX}};
    int cnt[NBINS], off[NBINS + 1];
    DevBuf list;
    CSRK_TRY(bin_rows(prod.as<int64_t>(), (int64_t)m, sspec, cnt, off, list, s));
    const unsigned long long P = PM[0];
    // Heavy rows are cut into chunks of A entries when the heaviest row alone would take more than half
    // of an SM's fair share of the products (the top item of configs[2] is 20 M products = 51 ms on one
    // CTA: the whole critical path of a multi-GPU run, where a rank's block is ~1/N of the products);
    // chunks are 1/8 of the fair share.  "own_chunk_prod": 0 = that rule, > 0 = products per chunk
    // (always chunk rows above it), < 0 = never chunk.
    const int64_t opt_chunk = options().own_chunk_prod.load();
    const int64_t fair = (int64_t)(P / (unsigned long long)sms) + 1;
    const int64_t chunk_prod = opt_chunk > 0 ? opt_chunk : std::max<int64_t>(fair / 8, (int64_t)1 << 18);
    const bool chunking = opt_chunk > 0 ? (int64_t)PM[1] > chunk_prod
                                        : (opt_chunk == 0 && (int64_t)PM[1] > std::max<int64_t>(fair / 2, chunk_prod));
    DevBuf row_nnz;
    CSRK_TRY(row_nnz.alloc_zero(sizeof(int32_t) * (size_t)m, s));
    DevBuf counter;
    CSRK_TRY(counter.alloc_zero(sizeof(int) * 2, s));
    const int32_t *L = list.as<int32_t>();
    // Wide results (more column windows than the dense kernels take): rows above the small hash bins go
    // through the expand / sort / compress path (spgemm_esc.cuh), which also does their numeric work; rows it
    // hands back (skewed beyond a pseudo-row's capacity) and everything else follow the bins below.
    EscState esc;
    const int64_t esc_opt = options().spgemm_esc.load();
    if ((esc_opt == 2 || (esc_opt == 1 && div_up((int64_t)n, DENSE_WIN) > DENSE_MAX_PASSES)) &&
        P / 16 < (1ull << 30)) {   // (pseudo-row and work-item counts are 32-bit)
        CSRK_TRY(esc_symbolic(A, B, L + off[3], cnt[3] + cnt[4] + cnt[5], prod.as<int64_t>(), both_f32, row_nnz.as<int32_t>(),
                              esc, s));
        if (esc.active)
            cnt[3] = cnt[4] = 0;
    }
    const int32_t *L5 = esc.active ? esc.old_list.as<int32_t>() : L + off[5];
    const int cnt5 = esc.active ? esc.n_old : cnt[5];
    if (cnt[1])
        CSRK_LAUNCH((k_sym_warp<SYM_WARP_SLOTS>), (unsigned)div_up(cnt[1], 8), 256, 0, s, A, B, L + off[1], cnt[1],
                    row_nnz.as<int32_t>());
    if (cnt[2]) {
        auto k = k_sym_cta<SYM_C1, 128>;
        CSRK_LAUNCH(k, (unsigned)cnt[2], 128, SYM_C1 * 4, s, A, B, L + off[2], row_nnz.as<int32_t>());
    }
    if (cnt[3]) {
        auto k = k_sym_cta<SYM_C2, 128>;
        CSRK_LAUNCH(k, (unsigned)cnt[3], 128, SYM_C2 * 4, s, A, B, L + off[3], row_nnz.as<int32_t>());
    }
    if (cnt[4]) {
        auto k = k_sym_cta<SYM_C3, 256>;
        CSRK_TRY(optin_smem(k, SYM_C3 * 4));
        CSRK_LAUNCH(k, (unsigned)cnt[4], 256, SYM_C3 * 4, s, A, B, L + off[4], row_nnz.as<int32_t>());
    }
    DevBuf gbm, keep, keep_slot;
    if (cnt5) {
        const size_t bm_bytes = (size_t)n_words * 4;
        if (bm_bytes * (size_t)cnt5 <= KEEP_BITMAP_BUDGET) {
            CSRK_TRY(keep.alloc(bm_bytes * (size_t)cnt5, s));
            CSRK_TRY(keep_slot.alloc(sizeof(int32_t) * (size_t)m, s));
        }
        // chunked symbolic rows OR into their kept bitmap, so chunking needs the kept bitmaps (zeroed)
        DevBuf s_nchunk, s_nsplit, s_last, s_item_off;
        const int *sio = nullptr;
        if (chunking && keep.p) {
            CSRK_TRY(s_nchunk.alloc(sizeof(int) * (size_t)cnt5, s));
            CSRK_TRY(s_nsplit.alloc(sizeof(int) * (size_t)cnt5, s));
            CSRK_TRY(s_last.alloc_zero(sizeof(int), s));
            CSRK_TRY(s_item_off.alloc(sizeof(int) * ((size_t)cnt5 + 1), s));
            CSRK_LAUNCH(k_own_items, (unsigned)div_up(cnt5, 256), 256, 0, s, A, L5, cnt5, prod.as<int64_t>(),
                        chunk_prod, s_nchunk.as<int>(), s_nsplit.as<int>(), s_last.as<int>());
            CSRK_TRY((exclusive_scan<int>(ArrayLoader<int>{s_nchunk.as<int>()}, (int64_t)cnt5, s_item_off.as<int>(), s)));
            CSRK_CUDA(cudaMemsetAsync(keep.p, 0, bm_bytes * (size_t)cnt5, s));
            sio = s_item_off.as<int>();
        }
        const int64_t max_items = (int64_t)cnt5 + (sio ? (int64_t)sms * 8 + 1 : 0);
        if (options().sym_bytes.load() && bm_bytes * 8 <= 100 * 1024) {
            // one byte per column, plain stores (up to 102 400 columns: two CTAs per SM; four below 51 200)
            auto k = k_sym_bitmap<true, DENSE_THREADS, true>;
            CSRK_TRY(optin_smem(k, bm_bytes * 8));
            const int grid = (int)std::min(max_items, (int64_t)sms * (bm_bytes * 8 > 50 * 1024 ? 2 : 4));
            CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, bm_bytes * 8, s, A, B, L5, cnt5, row_nnz.as<int32_t>(),
                        (unsigned *)nullptr, n_words, counter.as<int>(), keep.as<unsigned>(), keep_slot.as<int32_t>(), sio);
        } else if (bm_bytes + 1024 <= smem_max - 8 * 1024) {
            auto k = k_sym_bitmap<true, DENSE_THREADS>;
            CSRK_TRY(optin_smem(k, bm_bytes));
            const int grid = (int)std::min(max_items, (int64_t)sms * (bm_bytes > 100 * 1024 ? 1 : 2));
            CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, bm_bytes, s, A, B, L5, cnt5, row_nnz.as<int32_t>(),
                        (unsigned *)nullptr, n_words, counter.as<int>(), keep.as<unsigned>(), keep_slot.as<int32_t>(), sio);
        } else {
            auto k = k_sym_bitmap<false, DENSE_THREADS>;
            const int grid = (int)std::min(max_items, (int64_t)sms * 2);
            CSRK_TRY(gbm.alloc_zero(bm_bytes * (size_t)grid, s));
            CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, 0, s, A, B, L5, cnt5, row_nnz.as<int32_t>(),
                        gbm.as<unsigned>(), n_words, counter.as<int>(), keep.as<unsigned>(), keep_slot.as<int32_t>(), sio);
        }
        if (sio)
            CSRK_LAUNCH(k_sym_finish, (unsigned)cnt5, 128, 0, s, L5, sio, n_words, keep.as<unsigned>(),
                        row_nnz.as<int32_t>(), keep_slot.as<int32_t>());
    }

    CSRK_TRACE_MARK("spgemm: products + symbolic", s);
    // ---- step 2: rowptrs
    DevBuf rp64;
    CSRK_TRY(rp64.alloc(sizeof(int64_t) * ((size_t)m + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<int32_t>{row_nnz.as<int32_t>()}, (int64_t)m, rp64.as<int64_t>(), s)));
    int64_t Z = 0;
    CSRK_CUDA(cudaMemcpyAsync(&Z, rp64.as<int64_t>() + m, sizeof Z, cudaMemcpyDeviceToHost, s));
    // numeric binning (its sync also lands Z and P)
    BinSpec nspec{{0, NUM_WARP_SLOTS / 2, NUM_C1 / 2, NUM_C2 / 2, NUM_C3 / 2, INT64_MAX}};
    int ncnt[NBINS], noff[NBINS + 1];
    DevBuf nlist;
    if (esc.active)   // their values exist already (esc_emit copies them): keep them out of the numeric bins
        CSRK_LAUNCH(k_esc_zero_rows, (unsigned)div_up(esc.n_rows, 256), 256, 0, s, esc.rows, esc.n_rows, esc.bad.as<int>(),
                    row_nnz.as<int32_t>());
    CSRK_TRY(bin_rows(row_nnz.as<int32_t>(), (int64_t)m, nspec, ncnt, noff, nlist, s));
    gbm.reset();

    CSRK_TRACE_MARK("spgemm: rowptr scan + numeric bins", s);
    csrk_matrix *out = nullptr;
    const int rp_is64 = Z > (int64_t)INT32_MAX ? 1 : 0;
    CSRK_TRY(matrix_alloc(&out, m, n, Z, rp_is64, 8, s));
    CSRK_TRACE_MARK("spgemm: result allocated", s);
    int rc = CSRK_OK;
    auto fail = [&](int code) {
        matrix_destroy(out, s);
        return code;
    };
    {
        const unsigned grid = (unsigned)div_up((int64_t)m + 1, 256);
        if (rp_is64)
            k_narrow_rp<int64_t><<<grid, 256, 0, s>>>(rp64.as<int64_t>(), (int64_t *)out->rp, (int64_t)m + 1);
        else
            k_narrow_rp<int32_t><<<grid, 256, 0, s>>>(rp64.as<int64_t>(), (int32_t *)out->rp, (int64_t)m + 1);
        g_launches.fetch_add(1);
    }

    // heavy rows are handed to the persistent CTAs most expensive first (LPT): with dynamic
    // scheduling the tail of the kernel is then made of cheap rows
    DevBuf lpt_keys, lpt_rows;
    if (ncnt[5] > 1) {
        rc = lpt_keys.alloc(sizeof(int32_t) * (size_t)ncnt[5], s);
        if (rc == CSRK_OK)
            rc = lpt_rows.alloc(sizeof(int32_t) * (size_t)ncnt[5], s);
        if (rc != CSRK_OK)
            return fail(rc);
        k_lpt_keys<<<(unsigned)div_up(ncnt[5], 256), 256, 0, s>>>(nlist.as<int32_t>() + noff[5], ncnt[5], prod.as<int64_t>(),
                                                                   lpt_keys.as<int32_t>());
        g_launches.fetch_add(1);
        rc = radix_sort_by_key<NoPayload>(lpt_keys.as<int32_t>(), nlist.as<int32_t>() + noff[5], (const NoPayload *)nullptr,
                                          (int64_t)ncnt[5], 31, lpt_rows.as<int32_t>(), (NoPayload *)nullptr, s);
        if (rc != CSRK_OK)
            return fail(rc);
        cudaError_t ce = cudaMemcpyAsync(nlist.as<int32_t>() + noff[5], lpt_rows.p, sizeof(int32_t) * (size_t)ncnt[5],
                                         cudaMemcpyDeviceToDevice, s);
        if (ce != cudaSuccess)
            return fail(cuda_fail(ce, "lpt copy", __FILE__, __LINE__));
    }

    CSRK_TRACE_MARK("spgemm: LPT order", s);
    // ---- step 3: numeric
    const int32_t *NL = nlist.as<int32_t>();
    const int64_t *crp = rp64.as<int64_t>();
    double *cvs = (double *)out->vs;
    auto numeric = [&]() -> int {
        CSRK_TRY(esc_emit(esc, crp, out->ci, cvs, s));
        if (ncnt[1])
            CSRK_LAUNCH((k_num_warp<NUM_WARP_SLOTS>), (unsigned)div_up(ncnt[1], 8), 256, 0, s, A, B, NL + noff[1], ncnt[1],
                        crp, out->ci, cvs, both_f32);
        if (ncnt[2]) {
            auto k = k_num_cta<NUM_C1, 128, 512>;
            CSRK_LAUNCH(k, (unsigned)ncnt[2], 128, NUM_C1 * 12 + 512 * 4, s, A, B, NL + noff[2], crp, out->ci, cvs, both_f32);
        }
        if (ncnt[3]) {
            auto k = k_num_cta<NUM_C2, 256, 2048>;
            CSRK_TRY(optin_smem(k, NUM_C2 * 12 + 2048 * 4));
            CSRK_LAUNCH(k, (unsigned)ncnt[3], 256, NUM_C2 * 12 + 2048 * 4, s, A, B, NL + noff[3], crp, out->ci, cvs,
                        both_f32);
        }
        if (ncnt[4]) {
            auto k = k_num_cta<NUM_C3, 512, 4096>;
            CSRK_TRY(optin_smem(k, NUM_C3 * 12 + 4096 * 4));
            CSRK_LAUNCH(k, (unsigned)ncnt[4], 512, NUM_C3 * 12 + 4096 * 4, s, A, B, NL + noff[4], crp, out->ci, cvs,
                        both_f32);
        }
        // heavy rows; a second attempt (without the fixed-point kernel) only if its side list overflowed
        for (int attempt = 0; ncnt[5] && attempt < 2; attempt++) {
            int *wc = counter.as<int>() + 1;
            if (attempt)
                CSRK_CUDA(cudaMemsetAsync(wc, 0, sizeof(int), s));
            const unsigned *kp = keep.as<unsigned>();
            const int32_t *ks = keep_slot.as<int32_t>();
            const int passes = (int)div_up((int64_t)n, DENSE_WIN);
            bool owner = passes <= DENSE_MAX_PASSES && kp != nullptr && n >= 1024;
            bool fixed = false;
            DevBuf ea, eb, bscaled, tiny_buf, tiny_cnt;
            TinyList tiny{nullptr, nullptr, 0};
            MatView Bs = B;   // B with its values equilibrated by column (fixed-point kernel)
            bool eb_uniform = false, fix_side = true;
            int eb_u = 0;
            if (owner) {
                // both dense kernels walk B's rows by column range: rows strictly increasing in column
                DevBuf flag, erange;
                CSRK_TRY(flag.alloc_zero(sizeof(int), s));
                if (b->nnz > 1)
                    CSRK_LAUNCH(k_rows_not_strict, (unsigned)div_up(b->nnz - 1, 256), 256, 0, s, B, flag.as<int>());
                // fixed-point accumulator: needs finite operands of sane exponents (see the equilibration note)
                int hb = 0;
                for (unsigned long long l = PM[2]; l; l >>= 1)
                    hb++;  // ceil(log2(longest row + 1)): terms per output element
                const bool want_fixed = attempt == 0 && options().spgemm_fixed.load() != 0 && hb <= 24 && passes <= FIX_MAX_PASSES;
                int er[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};   // min / max exponent: A's rows, B's columns
                unsigned long long vr[4] = {~0ull, 0ull, ~0ull, 0ull};       // smallest / largest |v| of A, of B (bits)
                DevBuf vrange;
                if (want_fixed) {
                    CSRK_TRY(vrange.alloc(sizeof vr, s));
                    CSRK_CUDA(cudaMemcpyAsync(vrange.p, vr, sizeof vr, cudaMemcpyHostToDevice, s));
                    CSRK_LAUNCH(k_val_range, (unsigned)std::min<int64_t>(div_up(a->nnz, 256), (int64_t)sms * 8), 256, 0, s, a->vs,
                                a->val_kind, a->nnz, vrange.as<unsigned long long>());
                    CSRK_LAUNCH(k_val_range, (unsigned)std::min<int64_t>(div_up(b->nnz, 256), (int64_t)sms * 8), 256, 0, s, b->vs,
                                b->val_kind, b->nnz, vrange.as<unsigned long long>() + 2);
                    CSRK_CUDA(cudaMemcpyAsync(vr, vrange.p, sizeof vr, cudaMemcpyDeviceToHost, s));
                    CSRK_TRY(ea.alloc(sizeof(int) * (size_t)m, s));
                    CSRK_TRY(eb.alloc(sizeof(int) * (size_t)n, s));
                    CSRK_TRY(erange.alloc(sizeof er, s));
                    CSRK_CUDA(cudaMemcpyAsync(erange.p, er, sizeof er, cudaMemcpyHostToDevice, s));
                    CSRK_LAUNCH(k_fill_i32, (unsigned)div_up((int64_t)n, 256), 256, 0, s, eb.as<int>(), (int64_t)n, INT32_MIN);
                    CSRK_LAUNCH(k_row_expo, (unsigned)div_up((int64_t)m * 32, 256), 256, 0, s, A, ea.as<int>());
                    CSRK_LAUNCH(k_col_expo, (unsigned)std::min<int64_t>(div_up(b->nnz, 256), (int64_t)sms * 16), 256, 0, s, B,
                                eb.as<int>());
                    CSRK_LAUNCH(k_expo_range, (unsigned)std::min<int64_t>(div_up((int64_t)m, 256), (int64_t)sms * 4), 256, 0, s,
                                ea.as<int>(), (int)m, erange.as<int>());
                    CSRK_LAUNCH(k_expo_range, (unsigned)std::min<int64_t>(div_up((int64_t)n, 256), (int64_t)sms * 4), 256, 0, s,
                                eb.as<int>(), (int)n, erange.as<int>() + 2);
                    CSRK_CUDA(cudaMemcpyAsync(er, erange.p, sizeof er, cudaMemcpyDeviceToHost, s));
                }
                int bad = 0;
                CSRK_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
                CSRK_CUDA(cudaStreamSynchronize(s));
                owner = !bad;
                if (owner && want_fixed && er[1] != INT32_MIN && er[3] != INT32_MIN && er[0] >= -400 && er[1] <= 400 &&
                    er[2] >= -400 && er[3] <= 400) {
                    // the scaled copy of B's values (its own dtype: a power-of-two factor is exact in either) and
                    // the side list: room for 1/32 of the products, from the pool (it goes back after the call)
                    const int bvk = b->val_kind == 4 ? 4 : 8;
                    const unsigned long long cap =
                        (unsigned long long)std::min<int64_t>(std::max<int64_t>((int64_t)(P / 32), (int64_t)1 << 20), (int64_t)1 << 28);
                    const int64_t opt_cap = options().fix_tiny_cap.load();
                    tiny.cap = opt_cap > 0 ? (unsigned long long)opt_cap : cap;
                    eb_uniform = er[2] == er[3];   // every column of B peaks at the same exponent: no scaled copy
                    eb_u = er[2];
                    if (vr[1] && vr[3]) {   // products spanning at most 2^(25-hb): none can fall below the grid
                        double lim[4];
                        memcpy(lim, vr, sizeof lim);
                        fix_side = !(lim[1] * lim[3] <= ldexp(lim[0] * lim[2], 25 - hb));
                    }
                    if ((eb_uniform || bscaled.alloc((size_t)b->nnz * bvk, s) == CSRK_OK) &&
                        tiny_buf.alloc_owned(sizeof(TinyEnt) * (size_t)tiny.cap, s) == CSRK_OK &&
                        tiny_cnt.alloc_zero(sizeof(unsigned long long), s) == CSRK_OK) {
                        if (eb_uniform)
                            ;
                        else if (bvk == 4)
                            CSRK_LAUNCH(k_scale_cols<float>, (unsigned)div_up(b->nnz, 256), 256, 0, s, b->ci, b->vs, b->val_kind,
                                        b->nnz, eb.as<int>(), bscaled.as<float>());
                        else
                            CSRK_LAUNCH(k_scale_cols<double>, (unsigned)div_up(b->nnz, 256), 256, 0, s, b->ci, b->vs, b->val_kind,
                                        b->nnz, eb.as<int>(), bscaled.as<double>());
                        if (!eb_uniform) {
                            Bs.vs = bscaled.p;
                            Bs.vk = bvk;
                        }
                        tiny.buf = tiny_buf.as<TinyEnt>();
                        tiny.count = tiny_cnt.as<unsigned long long>();
                        fixed = true;
                    }
                }
            }
            if (owner) {
                const int win = (int)(div_up(div_up((int64_t)n, passes), 32) * 32);
                const int own_nw = options().own_nw.load() == 8 ? 8 : 16;
                const int nb = fixed ? passes : passes * own_nw;
                DevBuf hist, cum, bounds, split;
                CSRK_TRY(bounds.alloc(sizeof(int32_t) * ((size_t)nb + 1), s));
                CSRK_TRY(split.alloc(sizeof(int32_t) * (size_t)b->nrows * (nb + 1), s));
                if (fixed) {
                    CSRK_LAUNCH(k_fix_bounds, 1u, 32, 0, s, (int)n, win, passes, bounds.as<int32_t>());
                } else {
                    CSRK_TRY(hist.alloc_zero(sizeof(int) * ((size_t)n + 1), s));
                    CSRK_TRY(cum.alloc(sizeof(int64_t) * ((size_t)n + 1), s));
                    CSRK_LAUNCH(k_col_hist, (unsigned)div_up(b->nnz, 256), 256, 0, s, b->ci, b->nnz, hist.as<int>());
                    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<int>{hist.as<int>()}, (int64_t)n, cum.as<int64_t>(), s)));
                    CSRK_LAUNCH(k_own_bounds, (unsigned)div_up(nb + 1, 64), 64, 0, s, cum.as<int64_t>(), (int)n, win, passes,
                                own_nw, bounds.as<int32_t>());
                }
                CSRK_LAUNCH(k_own_split, (unsigned)div_up((int64_t)b->nrows * (nb + 1), 256), 256, 0, s, B,
                            bounds.as<int32_t>(), nb, split.as<int32_t>());
                const size_t bytes = (size_t)win * 8;
                // work items: a row, or one chunk of a heavy row's A entries (see chunk_prod above)
                DevBuf nchunk, nsplit, item_off, chunk_base, partial, last_split;
                CSRK_TRY(last_split.alloc_zero(sizeof(int), s));
                CSRK_TRY(nchunk.alloc(sizeof(int) * (size_t)ncnt[5], s));
                CSRK_TRY(nsplit.alloc(sizeof(int) * (size_t)ncnt[5], s));
                CSRK_TRY(item_off.alloc(sizeof(int) * ((size_t)ncnt[5] + 1), s));
                CSRK_TRY(chunk_base.alloc(sizeof(int) * ((size_t)ncnt[5] + 1), s));
                CSRK_LAUNCH(k_own_items, (unsigned)div_up(ncnt[5], 256), 256, 0, s, A, NL + noff[5], ncnt[5], prod.as<int64_t>(),
                            chunking ? chunk_prod : INT64_MAX / 2, nchunk.as<int>(), nsplit.as<int>(), last_split.as<int>());
                CSRK_TRY((exclusive_scan<int>(ArrayLoader<int>{nchunk.as<int>()}, (int64_t)ncnt[5], item_off.as<int>(), s)));
                CSRK_TRY((exclusive_scan<int>(ArrayLoader<int>{nsplit.as<int>()}, (int64_t)ncnt[5], chunk_base.as<int>(), s)));
                int tots[3] = {0, 0, 0};
                CSRK_CUDA(cudaMemcpyAsync(&tots[2], last_split.p, sizeof(int), cudaMemcpyDeviceToHost, s));
                CSRK_CUDA(cudaMemcpyAsync(&tots[0], item_off.as<int>() + ncnt[5], sizeof(int), cudaMemcpyDeviceToHost, s));
                CSRK_CUDA(cudaMemcpyAsync(&tots[1], chunk_base.as<int>() + ncnt[5], sizeof(int), cudaMemcpyDeviceToHost, s));
                CSRK_CUDA(cudaStreamSynchronize(s));
                const int nitems = tots[0], nparts = tots[1];
                if (nparts)
                    CSRK_TRY(partial.alloc(sizeof(double) * (size_t)nparts * passes * win, s));
                const int grid = (int)std::min((int64_t)nitems, (int64_t)sms);
                const int64_t ncomb = tots[2];  // chunked rows sit at the head of the LPT-ordered list
                CSRK_TRACE_MARK("spgemm: light bins + dense prep (value range, bounds, split, items)", s);
                if (fixed) {
                    const int *ebp = eb_uniform ? nullptr : eb.as<int>();
                    const int64_t ft = options().fix_threads.load();
                    if (ft == 1024) {
                        auto k = fix_side ? k_num_fixed<1024, true> : k_num_fixed<1024, false>;
                        CSRK_TRY(optin_smem(k_num_fixed<1024, true>, bytes));
                        CSRK_TRY(optin_smem(k_num_fixed<1024, false>, bytes));
                        CSRK_LAUNCH(k, (unsigned)grid, 1024, bytes, s, A, Bs, NL + noff[5], ncnt[5], crp, out->ci, cvs, both_f32,
                                    (int)n, win, passes, wc, kp, ks, split.as<int32_t>(), item_off.as<int>(),
                                    chunk_base.as<int>(), nitems, partial.as<long long>(), ea.as<int>(), ebp, eb_u, tiny);
                    } else if (ft == 768) {
                        auto k = fix_side ? k_num_fixed<768, true> : k_num_fixed<768, false>;
                        CSRK_TRY(optin_smem(k_num_fixed<768, true>, bytes));
                        CSRK_TRY(optin_smem(k_num_fixed<768, false>, bytes));
                        CSRK_LAUNCH(k, (unsigned)grid, 768, bytes, s, A, Bs, NL + noff[5], ncnt[5], crp, out->ci, cvs, both_f32,
                                    (int)n, win, passes, wc, kp, ks, split.as<int32_t>(), item_off.as<int>(),
                                    chunk_base.as<int>(), nitems, partial.as<long long>(), ea.as<int>(), ebp, eb_u, tiny);
                    } else {
                        auto k = fix_side ? k_num_fixed<512, true> : k_num_fixed<512, false>;
                        CSRK_TRY(optin_smem(k_num_fixed<512, true>, bytes));
                        CSRK_TRY(optin_smem(k_num_fixed<512, false>, bytes));
                        CSRK_LAUNCH(k, (unsigned)grid, 512, bytes, s, A, Bs, NL + noff[5], ncnt[5], crp, out->ci, cvs, both_f32,
                                    (int)n, win, passes, wc, kp, ks, split.as<int32_t>(), item_off.as<int>(),
                                    chunk_base.as<int>(), nitems, partial.as<long long>(), ea.as<int>(), ebp, eb_u, tiny);
                    }
                    if (nparts) {
                        CSRK_TRACE_MARK("spgemm: numeric (fixed-point kernel)", s);
                        CSRK_LAUNCH(k_fix_combine, (unsigned)(ncomb * passes), 256, 0, s, A, NL + noff[5], ncnt[5],
                                    item_off.as<int>(), chunk_base.as<int>(), partial.as<long long>(), (int)n, win, passes, kp,
                                    ks, crp, out->ci, cvs, ea.as<int>(), ebp, eb_u);
                    }
                    // the side list: how many entries, and did they fit?
                    unsigned long long n_tiny = 0;
                    CSRK_CUDA(cudaMemcpyAsync(&n_tiny, tiny.count, sizeof n_tiny, cudaMemcpyDeviceToHost, s));
                    CSRK_CUDA(cudaStreamSynchronize(s));
                    if (n_tiny > tiny.cap)
                        continue;   // too many coarse products for the list: the owner kernel redoes the heavy rows
                    if (n_tiny)
                        CSRK_LAUNCH(k_fix_tiny, (unsigned)div_up((int64_t)n_tiny, 256), 256, 0, s, tiny.buf, n_tiny, A, ea.as<int>(),
                                    ebp, eb_u, (const void *)crp, 1, out->ci, cvs);
                    out->stat_tiny = (int64_t)n_tiny;
                    out->stat_path = 2;
                } else {
                    if (own_nw == 8) {
                        auto k = k_num_owner<8>;
                        CSRK_TRY(optin_smem(k, bytes));
                        CSRK_LAUNCH(k, (unsigned)grid, 8 * 32, bytes, s, A, B, NL + noff[5], ncnt[5], crp, out->ci, cvs, both_f32,
                                    (int)n, win, passes, wc, kp, ks, split.as<int32_t>(), item_off.as<int>(),
                                    chunk_base.as<int>(), nitems, partial.as<double>());
                    } else {
                        auto k = k_num_owner<16>;
                        CSRK_TRY(optin_smem(k, bytes));
                        CSRK_LAUNCH(k, (unsigned)grid, 16 * 32, bytes, s, A, B, NL + noff[5], ncnt[5], crp, out->ci, cvs, both_f32,
                                    (int)n, win, passes, wc, kp, ks, split.as<int32_t>(), item_off.as<int>(),
                                    chunk_base.as<int>(), nitems, partial.as<double>());
                    }
                    if (nparts) {
                        CSRK_TRACE_MARK("spgemm: numeric (owner kernel)", s);
                        CSRK_LAUNCH(k_own_combine, (unsigned)(ncomb * passes), 256, 0, s, NL + noff[5], ncnt[5], item_off.as<int>(),
                                    chunk_base.as<int>(), partial.as<double>(), (int)n, win, passes, kp, ks, crp, out->ci, cvs);
                    }
                    out->stat_path = 1;
                }
            } else if (passes <= DENSE_MAX_PASSES) {
                const int win = (int)(div_up(div_up((int64_t)n, passes), 32) * 32);  // balanced windows
                const size_t bytes = (size_t)win * 8 + (size_t)(win / 32 + 1) * 4;
                auto k = k_num_dense<true, DENSE_THREADS>;
                CSRK_TRY(optin_smem(k, bytes));
                const int grid = (int)std::min((int64_t)ncnt[5], (int64_t)sms * (bytes > 100 * 1024 ? 1 : 2));
                CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, bytes, s, A, B, NL + noff[5], ncnt[5], crp, out->ci, cvs,
                            both_f32, (double *)nullptr, (unsigned *)nullptr, (int)n, win, wc, kp, ks);
            } else {
                auto k = k_num_dense<false, DENSE_THREADS>;
                const int grid = (int)std::min((int64_t)ncnt[5], (int64_t)sms * 2);
                DevBuf gacc, gbm2;
                CSRK_TRY(gacc.alloc_zero((size_t)n * 8 * (size_t)grid, s));
                CSRK_TRY(gbm2.alloc_zero((size_t)n_words * 4 * (size_t)grid, s));
                const int win = n_words * 32;  // one pass over all columns
                CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, 0, s, A, B, NL + noff[5], ncnt[5], crp, out->ci, cvs, both_f32,
                            gacc.as<double>(), gbm2.as<unsigned>(), (int)n, win, wc, kp, ks);
            }
            break;
        }
        return CSRK_OK;
    };
    rc = numeric();
    if (rc != CSRK_OK)
        return fail(rc);
    CSRK_TRACE_MARK("spgemm: numeric", s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return fail(cuda_fail(e, "spgemm", __FILE__, __LINE__));
    out->stat_products = (int64_t)P;
    out->stat_out_nnz = Z;
    *c = out;
    return CSRK_OK;
}

}  // namespace csrk
