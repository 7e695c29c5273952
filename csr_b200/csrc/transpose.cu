// transpose.cu -- stable CSR->CSC (csr/structure.py:172-247) and order_columns
// (csr/kernels/numba/__init__.py:47-52 -> structure.py:156-169) for sm_100a.
//
// The reference transpose is a counting sort by column whose scatter is STABLE
// (structure.py:191-197: entries of an output row appear in source order).  An
// atomic-cursor scatter is not, so the device version is a hand-written stable
// LSD radix sort of (column key, source row, value) with 8-bit digits:
//   per pass   k_radix_hist     per-tile digit histogram        (reads 4 B/nnz)
//              exclusive_scan   digit-major (digit, tile) scan  (tiny)
//              k_radix_scatter  stable rank + scatter           (reads+writes 8+V B/nnz)
// Stability inside a tile comes from processing 32 consecutive entries per warp
// step, ranking equal digits by lane with __match_any_sync, per-warp digit
// counters, and an exclusive prefix over the warps of the CTA.
//
// order_columns is the same machinery applied twice: (A^T)^T with both counting
// sorts stable leaves every row sorted by column with equal columns in their
// original relative order -- exactly what the reference's bubble sort produces.
#include "radix.cuh"
#include "expand.cuh"

namespace csrk {

__global__ void k_f32_to_f64(const float *__restrict__ in, double *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = (double)in[i];
}

template <typename RPT>
__global__ void k_rows_unsorted(const RPT *__restrict__ rp, const int32_t *__restrict__ ci, int32_t nrows, int64_t nnz,
                                int *__restrict__ flag)
{
    // entry i is out of order if it is not the first of its row and ci[i-1] > ci[i]
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= nnz)
        return;
    if (ci[i - 1] > ci[i]) {
        // is i a row start?  (some r with rp[r] == i)
        int64_t r = lower_bound_rp(rp, 0, (int64_t)nrows + 1, i);
        if (!(r <= nrows && (int64_t)rp[r] == i))
            *flag = 1;
    }
}

struct CoreOut {
    DevBuf rp;  // RPT[ncols+1]
    DevBuf ci;  // int32[nnz]  (source row of every entry, i.e. the transpose's colinds)
    DevBuf vs;  // VT[nnz]
};

static int key_bits(int32_t ncols)
{
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) < (int64_t)ncols)
        b++;
    return b;
}

// Stable counting sort of the nnz entries by column.  VT is the payload type moved
// with each entry (NoPayload / float / double).
template <typename RPT, typename VT>
static int transpose_core(int32_t nrows, int32_t ncols, int64_t nnz, const RPT *rp, const int32_t *ci, const VT *vs,
                          CoreOut &out, cudaStream_t s)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    CSRK_TRACE_MARK("transpose: enter", s);
    CSRK_TRY(out.rp.alloc_owned(sizeof(RPT) * ((size_t)ncols + 1), s));
    CSRK_TRY(out.ci.alloc_owned(sizeof(int32_t) * (size_t)nnz, s));
    if (HASV)
        CSRK_TRY(out.vs.alloc_owned(sizeof(VT) * (size_t)nnz, s));
    if (nnz == 0) {
        CSRK_CUDA(cudaMemsetAsync(out.rp.p, 0, sizeof(RPT) * ((size_t)ncols + 1), s));
        return CSRK_OK;
    }

    // 1. source row of every entry
    DevBuf rows0;
    CSRK_TRY(rows0.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_LAUNCH((k_expand_rows<RPT>), (unsigned)div_up(nnz, EXP_TILE), 256, 0, s, rp, nrows, nnz, rows0.as<int32_t>());

    CSRK_TRACE_MARK("transpose: expand rows", s);
    // 2. stable LSD radix sort by column; the payloads land in out.ci / out.vs
    DevBuf skeys;
    CSRK_TRY(skeys.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY((radix_sort_by_key<VT>(ci, rows0.as<int32_t>(), vs, nnz, key_bits(ncols), out.ci.as<int32_t>(),
                                    HASV ? out.vs.as<VT>() : nullptr, s, skeys.as<int32_t>())));
    CSRK_TRACE_MARK("transpose: radix sort", s);
    // 3. output rowptrs = where each column starts in the sorted keys
    CSRK_LAUNCH((k_key_bounds<RPT>), (unsigned)div_up(div_up(nnz + 1, 4), 256), 256, 0, s, skeys.as<int32_t>(), nnz, ncols,
                out.rp.as<RPT>());
    CSRK_TRACE_MARK("transpose: rowptrs", s);
    return CSRK_OK;
}

template <typename RPT>
static int transpose_typed(csrk_matrix *a, int vk, CoreOut &out, cudaStream_t s)
{
    const RPT *rp = (const RPT *)a->rp;
    if (vk == 4)
        return transpose_core<RPT, float>(a->nrows, a->ncols, a->nnz, rp, a->ci, (const float *)a->vs, out, s);
    if (vk == 8)
        return transpose_core<RPT, double>(a->nrows, a->ncols, a->nnz, rp, a->ci, (const double *)a->vs, out, s);
    return transpose_core<RPT, NoPayload>(a->nrows, a->ncols, a->nnz, rp, a->ci, (const NoPayload *)nullptr, out, s);
}

int transpose_run(csrk_matrix *a, int with_values, csrk_matrix **result, cudaStream_t s)
{
    *result = nullptr;
    const int vk = (with_values && a->val_kind) ? a->val_kind : 0;
    CoreOut out;
    if (a->rp_is64)
        CSRK_TRY(transpose_typed<int64_t>(a, vk, out, s));
    else
        CSRK_TRY(transpose_typed<int32_t>(a, vk, out, s));
    csrk_matrix *m = new (std::nothrow) csrk_matrix();
    if (!m) {
        set_error("host allocation failed");
        return CSRK_ENOMEM;
    }
    m->nrows = a->ncols;
    m->ncols = a->nrows;
    m->nnz = a->nnz;
    m->rp_is64 = a->rp_is64;
    m->val_kind = vk ? 8 : 0;  // structure.py:177: transposed values are always float64
    if (vk == 4) {
        DevBuf v64;
        int rc = v64.alloc_owned(sizeof(double) * (size_t)a->nnz, s);
        if (rc != CSRK_OK) {
            delete m;
            return rc;
        }
        if (a->nnz) {
            k_f32_to_f64<<<(unsigned)div_up(a->nnz, 256), 256, 0, s>>>(out.vs.as<float>(), v64.as<double>(), a->nnz);
            g_launches.fetch_add(1);
        }
        m->vs = v64.release();
    } else if (vk == 8) {
        m->vs = out.vs.release();
    }
    m->rp = out.rp.release();
    m->ci = (int32_t *)out.ci.release();
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        matrix_destroy(m, s);
        return cuda_fail(e, "transpose", __FILE__, __LINE__);
    }
    *result = m;
    return CSRK_OK;
}

template <typename RPT, typename VT>
static int order_typed(csrk_matrix *h, cudaStream_t s)
{
    CoreOut t1, t2;
    CSRK_TRY((transpose_core<RPT, VT>(h->nrows, h->ncols, h->nnz, (const RPT *)h->rp, h->ci, (const VT *)h->vs, t1, s)));
    CSRK_TRY((transpose_core<RPT, VT>(h->ncols, h->nrows, h->nnz, t1.rp.as<RPT>(), t1.ci.as<int32_t>(), t1.vs.as<VT>(),
                                      t2, s)));
    CSRK_CUDA(cudaStreamSynchronize(s));
    // rowptrs are unchanged by a within-row permutation; swap in the sorted arrays
    dev_free(h->ci, s);
    h->ci = (int32_t *)t2.ci.release();
    if (!std::is_same<VT, NoPayload>::value) {
        dev_free(h->vs, s);
        h->vs = t2.vs.release();
    }
    return CSRK_OK;
}

int order_columns_run(csrk_matrix *h, cudaStream_t s)
{
    if (h->nnz < 2)
        return CSRK_OK;
    // cheap exit: already sorted (the reference's bubble sort is O(n) then, too)
    DevBuf flag;
    CSRK_TRY(flag.alloc_zero(sizeof(int), s));
    if (h->rp_is64)
        CSRK_LAUNCH((k_rows_unsorted<int64_t>), (unsigned)div_up(h->nnz - 1, 256), 256, 0, s, (const int64_t *)h->rp,
                    h->ci, h->nrows, h->nnz, flag.as<int>());
    else
        CSRK_LAUNCH((k_rows_unsorted<int32_t>), (unsigned)div_up(h->nnz - 1, 256), 256, 0, s, (const int32_t *)h->rp,
                    h->ci, h->nrows, h->nnz, flag.as<int>());
    int unsorted = 0;
    CSRK_CUDA(cudaMemcpyAsync(&unsorted, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    if (!unsorted)
        return CSRK_OK;
    if (h->rp_is64) {
        if (h->val_kind == 4) return order_typed<int64_t, float>(h, s);
        if (h->val_kind == 8) return order_typed<int64_t, double>(h, s);
        return order_typed<int64_t, NoPayload>(h, s);
    }
    if (h->val_kind == 4) return order_typed<int32_t, float>(h, s);
    if (h->val_kind == 8) return order_typed<int32_t, double>(h, s);
    return order_typed<int32_t, NoPayload>(h, s);
}

// ---- from_coo (csr/structure.py:11-58): stable sort of the triples by row -----------------------
// The reference counts per row and scatters with a per-row cursor, i.e. entries keep their COO
// order inside a row.  Same machinery as the transpose with the roles swapped: key = row,
// payloads = column and value; the rowptrs are read off the sorted keys.
__global__ void k_range_flag(const int32_t *__restrict__ idx, int64_t n, int32_t bound, int *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (idx[i] < 0 || idx[i] >= bound))
        *flag = 1;
}

template <typename RPT, typename VT>
static int from_coo_typed(csrk_matrix *m, const int32_t *d_rows, const int32_t *d_cols, const VT *d_vals, cudaStream_t s)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    DevBuf skeys;
    CSRK_TRY(skeys.alloc(sizeof(int32_t) * (size_t)m->nnz, s));
    CSRK_TRY((radix_sort_by_key<VT>(d_rows, d_cols, d_vals, m->nnz, key_bits(m->nrows), m->ci,
                                    HASV ? (VT *)m->vs : nullptr, s, skeys.as<int32_t>())));
    CSRK_LAUNCH((k_key_bounds<RPT>), (unsigned)div_up(div_up(m->nnz + 1, 4), 256), 256, 0, s, skeys.as<int32_t>(), m->nnz,
                m->nrows, (RPT *)m->rp);
    return CSRK_OK;
}

// d_rows/d_cols: int32[nnz] on the device, d_vals: val_kind-typed or null.  Indices outside the shape
// are an argument error (the reference asserts them on the host, csr/csr.py:152-159).
int from_coo_run(int32_t nrows, int32_t ncols, int64_t nnz, const int32_t *d_rows, const int32_t *d_cols,
                 const void *d_vals, int val_kind, csrk_matrix **out, cudaStream_t s)
{
    *out = nullptr;
    const int rp_is64 = nnz > (int64_t)INT32_MAX ? 1 : 0;  // csr/csr.py:90-93
    csrk_matrix *m = nullptr;
    CSRK_TRY(matrix_alloc(&m, nrows, ncols, nnz, rp_is64, val_kind, s));
    auto fail = [&](int rc) {
        matrix_destroy(m, s);
        return rc;
    };
    int rc = CSRK_OK;
    if (nnz == 0) {
        cudaError_t e = cudaMemsetAsync(m->rp, 0, ((size_t)nrows + 1) * (rp_is64 ? 8 : 4), s);
        if (e != cudaSuccess)
            return fail(cuda_fail(e, "from_coo", __FILE__, __LINE__));
    } else {
        DevBuf flag;
        rc = flag.alloc_zero(sizeof(int), s);
        if (rc != CSRK_OK)
            return fail(rc);
        k_range_flag<<<(unsigned)div_up(nnz, 256), 256, 0, s>>>(d_rows, nnz, nrows, flag.as<int>());
        k_range_flag<<<(unsigned)div_up(nnz, 256), 256, 0, s>>>(d_cols, nnz, ncols, flag.as<int>());
        g_launches.fetch_add(2);
        int bad = 0;
        cudaError_t e = cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(s);
        if (e != cudaSuccess)
            return fail(cuda_fail(e, "from_coo", __FILE__, __LINE__));
        if (bad) {
            set_error("from_coo: a row or column index is outside the shape %d x %d", nrows, ncols);
            return fail(CSRK_EARG);
        }
        if (rp_is64) {
            if (val_kind == 4) rc = from_coo_typed<int64_t, float>(m, d_rows, d_cols, (const float *)d_vals, s);
            else if (val_kind == 8) rc = from_coo_typed<int64_t, double>(m, d_rows, d_cols, (const double *)d_vals, s);
            else rc = from_coo_typed<int64_t, NoPayload>(m, d_rows, d_cols, (const NoPayload *)nullptr, s);
        } else {
            if (val_kind == 4) rc = from_coo_typed<int32_t, float>(m, d_rows, d_cols, (const float *)d_vals, s);
            else if (val_kind == 8) rc = from_coo_typed<int32_t, double>(m, d_rows, d_cols, (const double *)d_vals, s);
            else rc = from_coo_typed<int32_t, NoPayload>(m, d_rows, d_cols, (const NoPayload *)nullptr, s);
        }
        if (rc != CSRK_OK)
            return fail(rc);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return fail(cuda_fail(e, "from_coo", __FILE__, __LINE__));
    *out = m;
    return CSRK_OK;
}

}  // namespace csrk
