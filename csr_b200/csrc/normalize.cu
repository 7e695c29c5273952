// normalize.cu -- row normalisation on the device: center_rows / unit_rows of
// csr/transform.py:13-66, reached through CSR.normalize_rows (csr/csr.py:443-469).  It is the step
// right before mult_abt in item-item similarity; doing it on the handle keeps the matrix in HBM.
//
// One warp per row, three sweeps over the row's values (the second and third hit L1/L2 for all
// but very long rows).  Sums are float64 with a lane-strided order + shuffle tree: the reference
// accumulates sequentially in the values' dtype, so results agree to rounding, not bit for bit
// (tests: rtol 1e-12 for float64, 1e-5 for float32, the reference's own test tolerances are looser).
#include "common.cuh"

namespace csrk {

template <typename RPT, typename VT>
__global__ void __launch_bounds__(256)
k_center_rows(int32_t nrows, const RPT *__restrict__ rp, VT *__restrict__ vs, VT *__restrict__ means)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows)
        return;
    const int64_t sp = (int64_t)rp[row], ep = (int64_t)rp[row + 1];
    if (sp == ep) {
        if (lane == 0)
            means[row] = (VT)0;  // transform.py:19-20: empty rows keep mean 0
        return;
    }
    double s = 0.0;
    for (int64_t k = sp + lane; k < ep; k += 32)
        s += (double)vs[k];
    s = warp_sum(s);
    s = __shfl_sync(0xffffffffu, s, 0);
    const double m = s / (double)(ep - sp);
    if (lane == 0)
        means[row] = (VT)m;
    for (int64_t k = sp + lane; k < ep; k += 32)
        vs[k] = (VT)((double)vs[k] - m);  // transform.py:24: evaluated in float64, stored in the array's dtype
}

template <typename RPT, typename VT>
__global__ void __launch_bounds__(256)
k_unit_rows(int32_t nrows, const RPT *__restrict__ rp, VT *__restrict__ vs, VT *__restrict__ norms)
{
    constexpr int MAXEXP = sizeof(VT) == 4 ? 128 : 1024, MINEXP = sizeof(VT) == 4 ? -126 : -1022;  // np.finfo
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows)
        return;
    const int64_t sp = (int64_t)rp[row], ep = (int64_t)rp[row + 1];
    if (sp == ep) {
        if (lane == 0)
            norms[row] = (VT)0;
        return;
    }
    // transform.py:48-55: largest magnitude -> power-of-two pre-normalisation
    double vmax = 0.0;
    for (int64_t k = sp + lane; k < ep; k += 32) {
        const double a = fabs((double)vs[k]);
        if (a > vmax || a != a)
            vmax = a;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double b = __shfl_xor_sync(0xffffffffu, vmax, o);
        if (b > vmax || b != b)
            vmax = b;
    }
    int ve = 0;
    (void)frexp(vmax, &ve);
    int pnexp = min(-ve, MAXEXP - 1);
    pnexp = max(pnexp, MINEXP);
    const double prenorm = ldexp(1.0, pnexp);
    // transform.py:56-60: norm of the pre-normalised values (as stored in the array's dtype)
    double ss = 0.0;
    for (int64_t k = sp + lane; k < ep; k += 32) {
        const VT w = (VT)((double)vs[k] * prenorm);
        ss += (double)w * (double)w;
    }
    ss = warp_sum(ss);
    ss = __shfl_sync(0xffffffffu, ss, 0);
    const VT inorm = (VT)sqrt(ss);
    if (lane == 0)
        norms[row] = (VT)((double)inorm / prenorm);
    for (int64_t k = sp + lane; k < ep; k += 32) {
        const VT w = (VT)((double)vs[k] * prenorm);
        vs[k] = w / inorm;  // 0/0 = NaN for an all-zero row, as the reference (test_transform.py:146)
    }
}

template <typename RPT, typename VT>
static int normalize_typed(csrk_matrix *h, int kind, void *d_vec, cudaStream_t s)
{
    const unsigned grid = (unsigned)div_up((int64_t)h->nrows * 32, 256);
    if (kind == 0)
        CSRK_LAUNCH((k_center_rows<RPT, VT>), grid, 256, 0, s, h->nrows, (const RPT *)h->rp, (VT *)h->vs, (VT *)d_vec);
    else
        CSRK_LAUNCH((k_unit_rows<RPT, VT>), grid, 256, 0, s, h->nrows, (const RPT *)h->rp, (VT *)h->vs, (VT *)d_vec);
    return CSRK_OK;
}

// kind: 0 = center, 1 = unit.  d_vec: nrows values of the matrix's value type (device).
int normalize_rows_run(csrk_matrix *h, int kind, void *d_vec, cudaStream_t s)
{
    if (h->nrows == 0)
        return CSRK_OK;
    if (h->rp_is64) {
        if (h->val_kind == 4)
            CSRK_TRY((normalize_typed<int64_t, float>(h, kind, d_vec, s)));
        else
            CSRK_TRY((normalize_typed<int64_t, double>(h, kind, d_vec, s)));
    } else {
        if (h->val_kind == 4)
            CSRK_TRY((normalize_typed<int32_t, float>(h, kind, d_vec, s)));
        else
            CSRK_TRY((normalize_typed<int32_t, double>(h, kind, d_vec, s)));
    }
    plan_invalidate(h, s);  // slab plans hold re-laid-out copies of the values
    return CSRK_OK;
}

}  // namespace csrk
