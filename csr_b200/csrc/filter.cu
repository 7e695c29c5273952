// filter.cu -- _filter_zeros (csr/_struct.py:61-79) on the device: drop stored
// zeros in place.  The reference compacts serially on the CPU after every
// multiply (csr/csr.py:555); here it is a keep-flag scan + scatter so the result
// of mult_ab can be filtered before it ever crosses PCIe.
#include "common.cuh"
#include "scan.cuh"

namespace csrk {

template <typename VT> struct KeepFlag {
    const VT *vs;
    __device__ __forceinline__ int operator()(int64_t i) const { return vs[i] != (VT)0 ? 1 : 0; }
};

template <typename VT>
__global__ void k_filter_scatter(const int32_t *__restrict__ ci, const VT *__restrict__ vs, int64_t nnz,
                                 const int64_t *__restrict__ pos, int32_t *__restrict__ ci_out, VT *__restrict__ vs_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz && vs[i] != (VT)0) {
        ci_out[pos[i]] = ci[i];
        vs_out[pos[i]] = vs[i];
    }
}

// new rowptrs[r] = pos[old rowptrs[r]]   (pos has nnz+1 entries; pos[nnz] = kept)
template <typename RPT>
__global__ void k_filter_rowptrs(RPT *__restrict__ rp, int64_t n, const int64_t *__restrict__ pos)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n)
        rp[r] = (RPT)pos[(int64_t)rp[r]];
}

template <typename VT> static int filter_typed(csrk_matrix *h, cudaStream_t s)
{
    const int64_t nnz = h->nnz;
    DevBuf pos;
    CSRK_TRY(pos.alloc(sizeof(int64_t) * ((size_t)nnz + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(KeepFlag<VT>{(const VT *)h->vs}, nnz, pos.as<int64_t>(), s)));
    int64_t kept = 0;
    CSRK_CUDA(cudaMemcpyAsync(&kept, pos.as<int64_t>() + nnz, sizeof kept, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    if (kept == nnz)
        return CSRK_OK;  // nothing stored is zero
    DevBuf ci2, vs2;
    CSRK_TRY(ci2.alloc_owned(sizeof(int32_t) * (size_t)kept, s));
    CSRK_TRY(vs2.alloc_owned(sizeof(VT) * (size_t)kept, s));
    CSRK_LAUNCH((k_filter_scatter<VT>), (unsigned)div_up(nnz, 256), 256, 0, s, h->ci, (const VT *)h->vs, nnz,
                pos.as<int64_t>(), ci2.as<int32_t>(), vs2.as<VT>());
    const unsigned grid = (unsigned)div_up((int64_t)h->nrows + 1, 256);
    // the rowptr width is kept (the reference rewrites the same array in place: _struct.py:66,75)
    if (h->rp_is64)
        CSRK_LAUNCH((k_filter_rowptrs<int64_t>), grid, 256, 0, s, (int64_t *)h->rp, (int64_t)h->nrows + 1, pos.as<int64_t>());
    else
        CSRK_LAUNCH((k_filter_rowptrs<int32_t>), grid, 256, 0, s, (int32_t *)h->rp, (int64_t)h->nrows + 1, pos.as<int64_t>());
    CSRK_CUDA(cudaStreamSynchronize(s));
    dev_free(h->ci, s);
    dev_free(h->vs, s);
    h->ci = (int32_t *)ci2.release();
    h->vs = vs2.release();
    h->nnz = kept;
    plan_invalidate(h, s);  // the SpMV tile map depends on rowptrs
    return CSRK_OK;
}

int filter_zeros_run(csrk_matrix *h, cudaStream_t s)
{
    if (h->val_kind == 0 || h->nnz == 0)
        return CSRK_OK;  // _filter_zeros only acts when values are present (csr.py:595-597)
    if (h->val_kind == 4)
        return filter_typed<float>(h, s);
    return filter_typed<double>(h, s);
}

}  // namespace csrk
