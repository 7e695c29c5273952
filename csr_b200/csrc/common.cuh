// common.cuh -- shared infrastructure of libcsr_cuda.so (sm_100a only).
//
// Context (device, stream, memory pool), per-thread error string, RAII device
// buffers on the stream-ordered allocator, launch accounting and small device
// helpers used by every kernel file.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "csrk.h"

namespace csrk {

// ------------------------------------------------------------------ errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CSRK_CUDA(call)                                                          \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess)                                                   \
            return ::csrk::cuda_fail(_e, #call, __FILE__, __LINE__);             \
    } while (0)

#define CSRK_TRY(call)                                                           \
    do {                                                                         \
        int _rc = (call);                                                        \
        if (_rc != CSRK_OK)                                                      \
            return _rc;                                                          \
    } while (0)

#define CSRK_ARG(cond, ...)                                                      \
    do {                                                                         \
        if (!(cond)) {                                                           \
            ::csrk::set_error(__VA_ARGS__);                                      \
            return CSRK_EARG;                                                    \
        }                                                                        \
    } while (0)

// ----------------------------------------------------------------- context
struct Context {
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;  // max dynamic shared memory per CTA (227 KB on B200)
    cudaStream_t stream = nullptr;
    std::mutex mu;
};
Context &ctx();
int ensure_init();  // CSRK_OK once csrk_init has run (auto-inits device 0 / $LOCAL_RANK)

extern std::atomic<int64_t> g_launches;

// CSRK_TRACE=1 in the environment prints "[csrk] <label> <ms>" lines to stderr at the phase marks
// below (each mark synchronises the stream, so traced runs are slower); the analogue of the
// reference shim's compile-time LK_TRACE (csr/kernels/mkl/mkl_ops.c:57-59).
bool trace_enabled();
void trace_mark(const char *label, cudaStream_t s);
#define CSRK_TRACE_MARK(label, s)                                                \
    do {                                                                         \
        if (::csrk::trace_enabled())                                             \
            ::csrk::trace_mark(label, s);                                        \
    } while (0)

// every kernel launch goes through this so csrk_launch_count() is exact
#define CSRK_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
    do {                                                                         \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
        ::csrk::g_launches.fetch_add(1, std::memory_order_relaxed);              \
        CSRK_CUDA(cudaGetLastError());                                           \
    } while (0)

// ------------------------------------------------------------ workspace arena
// Temporaries of the multi-kernel operations (transpose, SpGEMM, filter, plan builds) are bump-
// allocated from device chunks that are kept for the life of the library and rewound when the
// operation ends.  Churning multi-hundred-MB temporaries of varying sizes through the
// stream-ordered pool made it re-map physical memory at random (a 100M-nnz transpose took
// 15 ms or 170 ms); outputs owned by a handle still come from cudaMallocAsync.
// A WsScope serialises the operations that use the arena (they share the library stream anyway).
struct WsScope {
    WsScope();
    ~WsScope();
    WsScope(const WsScope &) = delete;
    WsScope &operator=(const WsScope &) = delete;
    std::vector<size_t> marks;
};
void ws_release_all();         // csrk_shutdown
void *ws_alloc(size_t bytes);  // nullptr if no scope is active on this thread or the device is out of memory

// ------------------------------------------------- stream-ordered device buffer
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    bool from_ws = false;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { reset(); }
    int alloc(size_t nbytes, cudaStream_t stream);       // temporary: workspace arena when a WsScope is active
    int alloc_zero(size_t nbytes, cudaStream_t stream);
    int alloc_owned(size_t nbytes, cudaStream_t stream);  // from the pool: may be release()d into a handle
    void reset();
    void *release()
    {
        void *q = p;
        p = nullptr;
        bytes = 0;
        return q;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

int dev_alloc(void **p, size_t bytes, cudaStream_t s);
void dev_free(void *p, cudaStream_t s);

// ------------------------------------------------------------------ matrix
struct SpmvPlan;  // spmv.cu: tile map of the CSR kernel
struct StreamPlan;  // spmv_slab.cu: slab-major re-layout of the entries for the slab kernel
struct YOut;        // spmv.cuh

// tunables settable through csrk_set_option (tests force the slab path on small inputs)
struct Options {
    std::atomic<int64_t> spmv_mode{0};                  // 0 auto, 1 CSR tile kernel, 2 slab kernel
    std::atomic<int64_t> stream_min_nnz{4 * 1000 * 1000};  // auto: smallest nnz worth a stream plan
    std::atomic<int64_t> stream_slab_bytes{0};          // > 0: cap on the x slab size (tests force many slabs)
    std::atomic<int64_t> stream_ctas{0};                // > 0: CTAs (row groups) of the stream kernel instead of one per SM
    std::atomic<int64_t> stream_warps{16};              // consumer warps per CTA of the stream kernel (1..31)
    std::atomic<int64_t> stream_piece{1024};             // longest pseudo-row: longer rows are cut into interleaved pieces
    std::atomic<int64_t> stream_ring_bytes{4096};       // per-warp prefetch ring of the entry stream (4096 | 8192)
    std::atomic<int64_t> stream_ring_chunks{2};         // chunks (bulk copies) per ring: 2 | 4
    std::atomic<int64_t> stream_xbufs{2};               // x slabs resident per CTA: 2 | 3
    std::atomic<int64_t> radix_bits{0};                 // digit width of the stable sort: 0 = pick (9 when it saves a pass), 8, 9
    std::atomic<int64_t> spmv_zero_copy_y{1};           // csrk_spmv: store rows straight into pinned host y
    std::atomic<int64_t> fix_threads{1024};             // threads per CTA of the fixed-point SpGEMM kernel (512 | 768 | 1024)
    std::atomic<int64_t> sym_bytes{1};                  // SpGEMM symbolic pass of heavy rows: byte marks with plain stores when the columns fit
    std::atomic<int64_t> spgemm_fixed{1};               // SpGEMM heavy rows: fixed-point atomics when the value range allows
    std::atomic<int64_t> own_chunk_prod{0};             // SpGEMM heavy-row chunking: 0 auto, > 0 products per chunk, < 0 off
    std::atomic<int64_t> fix_tiny_cap{0};               // > 0: capacity (entries) of the fixed-point kernel's side list (tests force an overflow)
    std::atomic<int64_t> spgemm_esc{1};                 // SpGEMM expand/sort/compress path: 0 off, 1 for wide results, 2 always (tests)
    std::atomic<int64_t> esc_target{1536};              // products per pseudo-row (row x column range) of that path
    std::atomic<int64_t> esc_stride{1};                 // its walk kernels take work items a prime stride apart (0: in order)
    std::atomic<int64_t> esc_budget{0};                 // > 0: cap in bytes on its expansion (else: half of the free memory)
    std::atomic<int64_t> own_nw{16};                    // warps (column ranges) per CTA in the owner-computes SpGEMM
};
Options &options();

}  // namespace csrk

// The opaque handle of csrk.h.  All pointers are device pointers owned here.
struct csrk_matrix {
    int32_t nrows = 0, ncols = 0;
    int64_t nnz = 0;
    void *rp = nullptr;     // int32[nrows+1] or int64[nrows+1]
    int rp_is64 = 0;
    int32_t *ci = nullptr;  // int32[nnz]
    void *vs = nullptr;     // float[nnz] / double[nnz] / nullptr
    int val_kind = 0;       // 0, 4, 8
    int64_t stat_products = -1, stat_out_nnz = -1, stat_tiny = 0;   // stat_tiny: products that went through the fixed-point kernel's side list
    int stat_path = 0;  // dense numeric path of the product that made this matrix: 0 none, 1 owner-computes, 2 fixed point
    csrk::SpmvPlan *plan = nullptr;  // lazily built SpMV tile map
    csrk::StreamPlan *stream[2] = {nullptr, nullptr};  // lazily built stream plans for float32 / float64 x
    bool stream_failed[2] = {false, false};
    std::atomic<int> spmv_calls{0};  // auto mode builds the slab plan on the SECOND mult_vec: one-shot handles never pay for it
    std::mutex mu;
};

namespace csrk {

int matrix_alloc(csrk_matrix **out, int32_t nrows, int32_t ncols, int64_t nnz, int rp_is64, int val_kind,
                 cudaStream_t s);
void matrix_destroy(csrk_matrix *m, cudaStream_t s);
void plan_destroy(SpmvPlan *p, cudaStream_t s);
void plan_invalidate(csrk_matrix *m, cudaStream_t s);
int normalize_rows_run(csrk_matrix *h, int kind, void *d_vec, cudaStream_t s);
int from_coo_run(int32_t nrows, int32_t ncols, int64_t nnz, const int32_t *d_rows, const int32_t *d_cols,
                 const void *d_vals, int val_kind, csrk_matrix **out, cudaStream_t s);  // syncs internally
void stream_destroy(StreamPlan *p, cudaStream_t s);
int stream_build(csrk_matrix *h, int x_kind, StreamPlan **out, cudaStream_t s);  // syncs internally; CSRK_EOVERFLOW = not representable
int stream_run(csrk_matrix *h, StreamPlan *p, const void *d_x, const YOut &y, cudaStream_t s);
void stream_info(const StreamPlan *p, int64_t *out /*[11]: G, NW, nslab, S, P, Q, n_split, smem, stream bytes, piece, ring*/);

// ops implemented across the .cu files (all enqueue on `s`, no sync unless stated)
int spmv_run(csrk_matrix *h, const void *d_x, int x_kind, double *d_y, cudaStream_t s);
bool spmv_uses_slab(const csrk_matrix *h, int x_kind, const void *d_x);  // the slab kernel (spmv_slab.cu) will serve this call
int spmv_run_multi(csrk_matrix *h, const void *d_x, int x_kind, double *const *d_ys, int n_out, cudaStream_t s,
                   bool multicast = false);
int mc_broadcast_run(void *mc_dst, const void *src, int64_t nbytes, cudaStream_t s);
int transpose_run(csrk_matrix *a, int with_values, csrk_matrix **out, cudaStream_t s);  // syncs internally
int order_columns_run(csrk_matrix *h, cudaStream_t s);                                  // syncs internally
int spgemm_run(csrk_matrix *a, csrk_matrix *b, csrk_matrix **c, cudaStream_t s);        // syncs internally
int filter_zeros_run(csrk_matrix *h, cudaStream_t s);                                   // syncs internally

// ---------------------------------------------------------- device helpers
__device__ __forceinline__ int64_t ld_rp(const void *rp, int is64, int64_t i)
{
    return is64 ? reinterpret_cast<const int64_t *>(rp)[i] : (int64_t) reinterpret_cast<const int32_t *>(rp)[i];
}

__device__ __forceinline__ double ld_val(const void *vs, int kind, int64_t i)
{
    if (kind == 8)
        return reinterpret_cast<const double *>(vs)[i];
    if (kind == 4)
        return (double)reinterpret_cast<const float *>(vs)[i];
    return 1.0;
}

// streaming 128-bit loads: read-only path, do not allocate in L1 (keep L1 for gathers)
__device__ __forceinline__ int4 ld_stream_int4(const void *ptr)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr));
    return r;
}
__device__ __forceinline__ float4 ld_stream_float4(const void *ptr)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(ptr));
    return r;
}
__device__ __forceinline__ double2 ld_stream_double2(const void *ptr)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(ptr));
    return r;
}

__device__ __forceinline__ int ld_stream_i32(const int *ptr)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(ptr));
    return r;
}
template <typename T> __device__ __forceinline__ T ld_stream(const T *ptr);
template <> __device__ __forceinline__ float ld_stream<float>(const float *ptr)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(ptr));
    return r;
}
template <> __device__ __forceinline__ double ld_stream<double>(const double *ptr)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(ptr));
    return r;
}

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// first index i in [lo, hi) with rp[i] >= key (hi if none)
template <typename RPT> __device__ __forceinline__ int64_t lower_bound_rp(const RPT *rp, int64_t lo, int64_t hi, int64_t key)
{
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)rp[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace csrk
