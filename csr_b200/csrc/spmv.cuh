// spmv.cuh -- pieces shared by the two SpMV kernels (spmv.cu: CSR tile kernel, spmv_slab.cu: slab-stream
// kernel): the value-less marker type, numba's product promotion and the (multi-destination) y store.
#pragma once

#include "common.cuh"

namespace csrk {

struct NoVal {};

// Where a finished row goes: p[0] is this GPU's y; p[1..n) are the same segment inside the gather
// buffers of the peer GPUs (NVLink peer memory), written by the same kernel so that the y
// all-gather of the row-partitioned SpMV needs no separate collective.
constexpr int SPMV_MAX_OUT = 8;
struct YOut {
    double *p[SPMV_MAX_OUT];
    int n;
    int mc;  // p[1] is an NVLink multicast (NVLS) address: ONE store lands in every GPU of the group
};
// `final` = the value is the row's result (not the head piece of a row that continues in later
// tiles and gets its carries added by k_spmv_fixup): only final values leave the GPU.
template <bool MULTI> __device__ __forceinline__ void store_y(const YOut &y, int64_t r, double v, bool final)
{
    y.p[0][r] = v;
    if (!MULTI || !final)
        return;
    if (y.mc) {
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(y.p[1] + r), "d"(v) : "memory");
        return;
    }
#pragma unroll
    for (int k = 1; k < SPMV_MAX_OUT; k++)
        if (k < y.n)
            y.p[k][r] = v;
}

template <typename VT, typename XT> struct Prod {
    using type = double;
};
template <> struct Prod<float, float> {
    using type = float;
};

}  // namespace csrk
