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
//                       they are written, or a dense float64 accumulator + bitmap
//                       whose sweep emits columns already in order.
// Products are formed in numba's promoted type (f4*f4 -> f4, else f8) and summed
// in float64, like multiply.py:120.  Only the summation ORDER differs from the
// reference (hence rtol instead of bit-exact values).
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "scan.cuh"

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

constexpr int NBINS = 5;
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
    return both_f32 ? (double)((float)av * (float)bv) : av * bv;
}

// ------------------------------------------------------------ step 0: products
__global__ void __launch_bounds__(256) k_row_products(MatView A, MatView B, int64_t *__restrict__ prod,
                                                      unsigned long long *__restrict__ total)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t p = 0;
    if (row < A.nrows) {
        const int64_t s = ld_rp(A.rp, A.rp64, row), e = ld_rp(A.rp, A.rp64, row + 1);
        for (int64_t jj = s + lane; jj < e; jj += 32) {
            const int32_t j = A.ci[jj];
            p += ld_rp(B.rp, B.rp64, (int64_t)j + 1) - ld_rp(B.rp, B.rp64, j);
        }
        p = warp_sum(p);
        if (lane == 0) {
            prod[row] = p;
            if (p)
                atomicAdd(total, (unsigned long long)p);
        }
    }
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
// scratch; either way it is restored to zero while it is counted.
template <bool SMEM_BM, int THREADS>
__global__ void __launch_bounds__(THREADS) k_sym_bitmap(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin,
                                                        int32_t *__restrict__ row_nnz, unsigned *__restrict__ gbm,
                                                        int n_words, int *__restrict__ work_counter)
{
    extern __shared__ unsigned s_bm[];
    __shared__ int s_idx, s_count;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned *bm = SMEM_BM ? s_bm : gbm + (size_t)blockIdx.x * n_words;
    if (SMEM_BM) {
        for (int i = tid; i < n_words; i += THREADS)
            s_bm[i] = 0;
    }
    while (true) {
        __syncthreads();
        if (tid == 0) {
            s_idx = atomicAdd(work_counter, 1);
            s_count = 0;
        }
        __syncthreads();
        const int idx = s_idx;
        if (idx >= nbin)
            break;
        const int32_t row = rows[idx];
        const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
        for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
            const int32_t j = A.ci[jj];
            const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
            for (int64_t kk = bs + lane; kk < be; kk += 32) {
                const int32_t k = B.ci[kk];
                atomicOr(&bm[k >> 5], 1u << (k & 31));
            }
        }
        __syncthreads();
        int count = 0;
        for (int i = tid; i < n_words; i += THREADS) {
            const unsigned bits = SMEM_BM ? bm[i] : __ldcg(&bm[i]);
            if (bits) {
                count += __popc(bits);
                if (SMEM_BM)
                    bm[i] = 0;
                else
                    __stcg(&bm[i], 0u);
            }
        }
        count = warp_sum(count);
        if (lane == 0 && count)
            atomicAdd(&s_count, count);
        __syncthreads();
        if (tid == 0)
            row_nnz[row] = s_count;
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

// One CTA per row, nnz_i <= SLOTS/2; the whole table is bitonic-sorted by key in
// shared memory (empty slots carry INT32_MAX and sink to the end).
template <int SLOTS, int THREADS>
__global__ void __launch_bounds__(THREADS) k_num_cta(MatView A, MatView B, const int32_t *__restrict__ rows,
                                                     const int64_t *__restrict__ c_rp, int32_t *__restrict__ c_ci,
                                                     double *__restrict__ c_vs, int both_f32)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    double *vals = reinterpret_cast<double *>(s_raw);
    int32_t *keys = reinterpret_cast<int32_t *>(s_raw + sizeof(double) * SLOTS);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < SLOTS; i += THREADS) {
        keys[i] = EMPTY_KEY;
        vals[i] = 0.0;
    }
    __syncthreads();
    const int32_t row = rows[blockIdx.x];
    const int64_t as = ld_rp(A.rp, A.rp64, row), ae = ld_rp(A.rp, A.rp64, (int64_t)row + 1);
    for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
        const int32_t j = A.ci[jj];
        const double av = ld_val(A.vs, A.vk, jj);
        const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
        for (int64_t kk = bs + lane; kk < be; kk += 32) {
            const int32_t k = B.ci[kk];
            const double p = product(av, ld_val(B.vs, B.vk, kk), both_f32);
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
    __syncthreads();
    // bitonic sort of (key, val) over all SLOTS
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
    const int64_t out = c_rp[row];
    const int nz = (int)(c_rp[row + 1] - out);
    for (int i = tid; i < nz; i += THREADS) {
        c_ci[out + i] = keys[i];
        c_vs[out + i] = vals[i];
    }
}

// ---------------------------------------------------- numeric: dense (heavy rows)
// Dense float64 accumulator of B.ncols entries + bitmap per persistent CTA
// (shared memory when SMEM_ACC, else this CTA's slice of zeroed global scratch,
// which on B200 stays L2-resident).  The sweep walks the bitmap in column order,
// so the row comes out sorted; accumulator and bitmap are zeroed as they are read.
template <bool SMEM_ACC, int THREADS>
__global__ void __launch_bounds__(THREADS) k_num_dense(MatView A, MatView B, const int32_t *__restrict__ rows, int nbin,
                                                       const int64_t *__restrict__ c_rp, int32_t *__restrict__ c_ci,
                                                       double *__restrict__ c_vs, int both_f32, double *__restrict__ gacc,
                                                       unsigned *__restrict__ gbm, int n_cols, int n_words,
                                                       int *__restrict__ work_counter)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_idx;
    __shared__ int s_wt[33];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double *acc = SMEM_ACC ? reinterpret_cast<double *>(s_raw) : gacc + (size_t)blockIdx.x * n_cols;
    unsigned *bm = SMEM_ACC ? reinterpret_cast<unsigned *>(s_raw + sizeof(double) * (size_t)n_cols)
                            : gbm + (size_t)blockIdx.x * n_words;
    if (SMEM_ACC) {
        for (int i = tid; i < n_cols; i += THREADS)
            acc[i] = 0.0;
        for (int i = tid; i < n_words; i += THREADS)
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
        for (int64_t jj = as + w; jj < ae; jj += THREADS / 32) {
            const int32_t j = A.ci[jj];
            const double av = ld_val(A.vs, A.vk, jj);
            const int64_t bs = ld_rp(B.rp, B.rp64, j), be = ld_rp(B.rp, B.rp64, (int64_t)j + 1);
            for (int64_t kk = bs + lane; kk < be; kk += 32) {
                const int32_t k = B.ci[kk];
                const double p = product(av, ld_val(B.vs, B.vk, kk), both_f32);
                atomicOr(&bm[k >> 5], 1u << (k & 31));
                atomicAdd(&acc[k], p);
            }
        }
        __syncthreads();
        int64_t out = c_rp[row];
        for (int base = 0; base < n_words; base += THREADS) {
            const int i = base + tid;
            unsigned bits = 0;
            if (i < n_words)
                bits = SMEM_ACC ? bm[i] : __ldcg(&bm[i]);
            int tot;
            int off = block_exclusive_scan<int>(__popc(bits), s_wt, tot);
            if (bits) {
                if (SMEM_ACC)
                    bm[i] = 0;
                else
                    __stcg(&bm[i], 0u);
                int64_t o = out + off;
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int k = i * 32 + b;
                    c_ci[o] = k;
                    if (SMEM_ACC) {
                        c_vs[o] = acc[k];
                        acc[k] = 0.0;
                    } else {
                        c_vs[o] = __ldcg(&acc[k]);
                        __stcg(&acc[k], 0.0);
                    }
                    o++;
                }
            }
            out += tot;
        }
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

// symbolic thresholds on P_i (products) and numeric thresholds on nnz_i
constexpr int SYM_WARP_SLOTS = 256, SYM_CTA1_SLOTS = 4096, SYM_CTA2_SLOTS = 32768;
constexpr int NUM_WARP_SLOTS = 128, NUM_CTA1_SLOTS = 2048, NUM_CTA2_SLOTS = 16384;
constexpr int DENSE_THREADS = 512;

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

    // ---- step 0: products per row
    DevBuf prod, total;
    CSRK_TRY(prod.alloc(sizeof(int64_t) * (size_t)m, s));
    CSRK_TRY(total.alloc_zero(sizeof(unsigned long long), s));
    CSRK_LAUNCH(k_row_products, (unsigned)div_up((int64_t)m * 32, 256), 256, 0, s, A, B, prod.as<int64_t>(),
                total.as<unsigned long long>());

    // ---- step 1: symbolic
    BinSpec sspec{{0, SYM_WARP_SLOTS / 2, SYM_CTA1_SLOTS / 2, SYM_CTA2_SLOTS / 2, INT64_MAX}};
    int cnt[NBINS], off[NBINS + 1];
    DevBuf list;
    CSRK_TRY(bin_rows(prod.as<int64_t>(), (int64_t)m, sspec, cnt, off, list, s));
    unsigned long long P = 0;
    CSRK_CUDA(cudaMemcpyAsync(&P, total.p, sizeof P, cudaMemcpyDeviceToHost, s));
    DevBuf row_nnz;
    CSRK_TRY(row_nnz.alloc_zero(sizeof(int32_t) * (size_t)m, s));
    DevBuf counter;
    CSRK_TRY(counter.alloc_zero(sizeof(int) * 2, s));
    const int32_t *L = list.as<int32_t>();
    if (cnt[1])
        CSRK_LAUNCH((k_sym_warp<SYM_WARP_SLOTS>), (unsigned)div_up(cnt[1], 8), 256, 0, s, A, B, L + off[1], cnt[1],
                    row_nnz.as<int32_t>());
    if (cnt[2]) {
        auto k = k_sym_cta<SYM_CTA1_SLOTS, 128>;
        CSRK_LAUNCH(k, (unsigned)cnt[2], 128, SYM_CTA1_SLOTS * 4, s, A, B, L + off[2], row_nnz.as<int32_t>());
    }
    if (cnt[3]) {
        auto k = k_sym_cta<SYM_CTA2_SLOTS, 256>;
        CSRK_TRY(optin_smem(k, SYM_CTA2_SLOTS * 4));
        CSRK_LAUNCH(k, (unsigned)cnt[3], 256, SYM_CTA2_SLOTS * 4, s, A, B, L + off[3], row_nnz.as<int32_t>());
    }
    DevBuf gbm;
    if (cnt[4]) {
        const size_t bm_bytes = (size_t)n_words * 4;
        if (bm_bytes + 1024 <= smem_max - 8 * 1024) {
            auto k = k_sym_bitmap<true, DENSE_THREADS>;
            CSRK_TRY(optin_smem(k, bm_bytes));
            const int grid = (int)std::min((int64_t)cnt[4], (int64_t)sms * (bm_bytes > 100 * 1024 ? 1 : 2));
            CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, bm_bytes, s, A, B, L + off[4], cnt[4], row_nnz.as<int32_t>(),
                        (unsigned *)nullptr, n_words, counter.as<int>());
        } else {
            auto k = k_sym_bitmap<false, DENSE_THREADS>;
            const int grid = (int)std::min((int64_t)cnt[4], (int64_t)sms * 2);
            CSRK_TRY(gbm.alloc_zero(bm_bytes * (size_t)grid, s));
            CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, 0, s, A, B, L + off[4], cnt[4], row_nnz.as<int32_t>(),
                        gbm.as<unsigned>(), n_words, counter.as<int>());
        }
    }

    // ---- step 2: rowptrs
    DevBuf rp64;
    CSRK_TRY(rp64.alloc(sizeof(int64_t) * ((size_t)m + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<int32_t>{row_nnz.as<int32_t>()}, (int64_t)m, rp64.as<int64_t>(), s)));
    int64_t Z = 0;
    CSRK_CUDA(cudaMemcpyAsync(&Z, rp64.as<int64_t>() + m, sizeof Z, cudaMemcpyDeviceToHost, s));
    // numeric binning (its sync also lands Z and P)
    BinSpec nspec{{0, NUM_WARP_SLOTS / 2, NUM_CTA1_SLOTS / 2, NUM_CTA2_SLOTS / 2, INT64_MAX}};
    int ncnt[NBINS], noff[NBINS + 1];
    DevBuf nlist;
    CSRK_TRY(bin_rows(row_nnz.as<int32_t>(), (int64_t)m, nspec, ncnt, noff, nlist, s));
    gbm.reset();

    csrk_matrix *out = nullptr;
    const int rp_is64 = Z > (int64_t)INT32_MAX ? 1 : 0;
    CSRK_TRY(matrix_alloc(&out, m, n, Z, rp_is64, 8, s));
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

    // ---- step 3: numeric
    const int32_t *NL = nlist.as<int32_t>();
    const int64_t *crp = rp64.as<int64_t>();
    double *cvs = (double *)out->vs;
    auto numeric = [&]() -> int {
        if (ncnt[1])
            CSRK_LAUNCH((k_num_warp<NUM_WARP_SLOTS>), (unsigned)div_up(ncnt[1], 8), 256, 0, s, A, B, NL + noff[1], ncnt[1],
                        crp, out->ci, cvs, both_f32);
        if (ncnt[2]) {
            auto k = k_num_cta<NUM_CTA1_SLOTS, 128>;
            CSRK_LAUNCH(k, (unsigned)ncnt[2], 128, NUM_CTA1_SLOTS * 12, s, A, B, NL + noff[2], crp, out->ci, cvs, both_f32);
        }
        if (ncnt[3]) {
            auto k = k_num_cta<NUM_CTA2_SLOTS, 512>;
            CSRK_TRY(optin_smem(k, NUM_CTA2_SLOTS * 12));
            CSRK_LAUNCH(k, (unsigned)ncnt[3], 512, NUM_CTA2_SLOTS * 12, s, A, B, NL + noff[3], crp, out->ci, cvs, both_f32);
        }
        if (ncnt[4]) {
            const size_t acc_bytes = (size_t)n * 8 + (size_t)n_words * 4;
            int *wc = counter.as<int>() + 1;
            if (acc_bytes + 1024 <= smem_max - 8 * 1024) {
                auto k = k_num_dense<true, DENSE_THREADS>;
                CSRK_TRY(optin_smem(k, acc_bytes));
                const int grid = (int)std::min((int64_t)ncnt[4], (int64_t)sms * (acc_bytes > 100 * 1024 ? 1 : 2));
                CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, acc_bytes, s, A, B, NL + noff[4], ncnt[4], crp, out->ci, cvs,
                            both_f32, (double *)nullptr, (unsigned *)nullptr, (int)n, n_words, wc);
            } else {
                auto k = k_num_dense<false, DENSE_THREADS>;
                const int grid = (int)std::min((int64_t)ncnt[4], (int64_t)sms * 2);
                DevBuf gacc, gbm2;
                CSRK_TRY(gacc.alloc_zero((size_t)n * 8 * (size_t)grid, s));
                CSRK_TRY(gbm2.alloc_zero((size_t)n_words * 4 * (size_t)grid, s));
                CSRK_LAUNCH(k, (unsigned)grid, DENSE_THREADS, 0, s, A, B, NL + noff[4], ncnt[4], crp, out->ci, cvs, both_f32,
                            gacc.as<double>(), gbm2.as<unsigned>(), (int)n, n_words, wc);
            }
        }
        return CSRK_OK;
    };
    rc = numeric();
    if (rc != CSRK_OK)
        return fail(rc);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return fail(cuda_fail(e, "spgemm", __FILE__, __LINE__));
    out->stat_products = (int64_t)P;
    out->stat_out_nnz = Z;
    *c = out;
    return CSRK_OK;
}

}  // namespace csrk
