// context.cu -- library context, error reporting and the handle lifecycle
// (to_handle / from_handle / release_handle of the kernel contract:
// csr/kernels/numba/__init__.py:16-44, precedent csr/kernels/mkl/handle.py:46-148).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <thread>

#include "common.cuh"

static void (*g_out_pool_release)() = nullptr;   // set once csrk_export has created its pinned slots

namespace csrk {

static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation)
        return CSRK_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return CSRK_ENODEV;
    return CSRK_ECUDA;
}

Context &ctx()
{
    static Context c;
    return c;
}

bool trace_enabled()
{
    static const bool on = [] {
        const char *e = getenv("CSRK_TRACE");
        return e && *e && *e != '0';
    }();
    return on;
}

void trace_mark(const char *label, cudaStream_t s)
{
    static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[csrk] %-28s %9.3f ms\n", label, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

Options &options()
{
    static Options o;
    return o;
}

static int init_locked(Context &c, int device)
{
    if (c.inited) {
        CSRK_ARG(device < 0 || device == c.device, "csrk_init(%d): library already bound to device %d", device, c.device);
        return CSRK_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no usable CUDA device (%s); libcsr_cuda has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        (void)cudaGetLastError();
        return CSRK_ENODEV;
    }
    if (device < 0) {
        const char *lr = getenv("LOCAL_RANK");
        device = lr ? atoi(lr) % ndev : 0;
    }
    CSRK_ARG(device < ndev, "csrk_init(%d): only %d device(s)", device, ndev);
    CSRK_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    CSRK_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        set_error("device %d is sm_%d%d; libcsr_cuda is built for sm_100a only", device, p.major, p.minor);
        return CSRK_ENODEV;
    }
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    c.cc_major = p.major;
    c.cc_minor = p.minor;
    c.smem_optin = p.sharedMemPerBlockOptin;
    CSRK_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    // keep freed blocks in the pool: handle churn (to_handle per call, csr.py:582) must not hit cudaMalloc
    cudaMemPool_t pool;
    CSRK_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    CSRK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    c.inited = true;
    return CSRK_OK;
}

int ensure_init()
{
    Context &c = ctx();
    std::lock_guard<std::mutex> g(c.mu);
    CSRK_TRY(init_locked(c, -1));
    // the runtime's current device is per host thread
    CSRK_CUDA(cudaSetDevice(c.device));
    return CSRK_OK;
}

int dev_alloc(void **p, size_t bytes, cudaStream_t s)
{
    *p = nullptr;
    if (bytes == 0)
        bytes = 16;  // never hand out NULL for an empty array: kernels may form (unused) pointers
    cudaError_t e = cudaMallocAsync(p, bytes, s);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? CSRK_ENOMEM : CSRK_ECUDA;
    }
    return CSRK_OK;
}

void dev_free(void *p, cudaStream_t s)
{
    if (p)
        (void)cudaFreeAsync(p, s);
}

// ---- workspace arena
namespace {
struct WsChunk {
    char *p;
    size_t cap, top;
};
struct Workspace {
    std::vector<WsChunk> chunks;
    std::recursive_mutex mu;
};
Workspace &ws()
{
    static Workspace w;
    return w;
}
thread_local int t_ws_depth = 0;
}  // namespace

WsScope::WsScope()
{
    ws().mu.lock();
    t_ws_depth++;
    for (auto &c : ws().chunks)
        marks.push_back(c.top);
}

WsScope::~WsScope()
{
    auto &ch = ws().chunks;
    for (size_t i = 0; i < ch.size(); i++)
        ch[i].top = i < marks.size() ? marks[i] : 0;
    t_ws_depth--;
    ws().mu.unlock();
}

void ws_release_all()
{
    std::lock_guard<std::recursive_mutex> g(ws().mu);
    for (auto &c : ws().chunks)
        (void)cudaFree(c.p);
    ws().chunks.clear();
}

void *ws_alloc(size_t bytes)
{
    if (t_ws_depth == 0)
        return nullptr;
    bytes = (bytes + 511) & ~(size_t)511;
    if (bytes == 0)
        bytes = 512;
    auto &ch = ws().chunks;
    for (auto &c : ch) {
        if (c.cap - c.top >= bytes) {
            void *p = c.p + c.top;
            c.top += bytes;
            return p;
        }
    }
    // new chunk: at least 256 MB, or the request rounded up to 64 MB
    size_t cap = std::max<size_t>((size_t)256 << 20, (bytes + ((size_t)64 << 20) - 1) & ~(((size_t)64 << 20) - 1));
    void *p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) {
        (void)cudaGetLastError();
        cap = bytes;
        if (cudaMalloc(&p, cap) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
    }
    ch.push_back(WsChunk{(char *)p, cap, bytes});
    return p;
}

int DevBuf::alloc(size_t nbytes, cudaStream_t stream)
{
    reset();
    s = stream;
    if (void *q = ws_alloc(nbytes)) {
        p = q;
        from_ws = true;
        bytes = nbytes;
        return CSRK_OK;
    }
    CSRK_TRY(dev_alloc(&p, nbytes, stream));
    bytes = nbytes;
    return CSRK_OK;
}

int DevBuf::alloc_owned(size_t nbytes, cudaStream_t stream)
{
    reset();
    s = stream;
    CSRK_TRY(dev_alloc(&p, nbytes, stream));
    bytes = nbytes;
    return CSRK_OK;
}

int DevBuf::alloc_zero(size_t nbytes, cudaStream_t stream)
{
    CSRK_TRY(alloc(nbytes, stream));
    CSRK_CUDA(cudaMemsetAsync(p, 0, nbytes ? nbytes : 16, stream));
    return CSRK_OK;
}

void DevBuf::reset()
{
    if (p && !from_ws)
        dev_free(p, s);
    p = nullptr;
    bytes = 0;
    from_ws = false;
}

int matrix_alloc(csrk_matrix **out, int32_t nrows, int32_t ncols, int64_t nnz, int rp_is64, int val_kind, cudaStream_t s)
{
    *out = nullptr;
    csrk_matrix *m = new (std::nothrow) csrk_matrix();
    if (!m) {
        set_error("host allocation failed");
        return CSRK_ENOMEM;
    }
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->rp_is64 = rp_is64;
    m->val_kind = val_kind;
    int rc = dev_alloc(&m->rp, ((size_t)nrows + 1) * (rp_is64 ? 8 : 4), s);
    if (rc == CSRK_OK)
        rc = dev_alloc((void **)&m->ci, (size_t)nnz * 4, s);
    if (rc == CSRK_OK && val_kind)
        rc = dev_alloc(&m->vs, (size_t)nnz * val_kind, s);
    if (rc != CSRK_OK) {
        matrix_destroy(m, s);
        return rc;
    }
    *out = m;
    return CSRK_OK;
}

void matrix_destroy(csrk_matrix *m, cudaStream_t s)
{
    if (!m)
        return;
    dev_free(m->rp, s);
    dev_free(m->ci, s);
    dev_free(m->vs, s);
    plan_destroy(m->plan, s);
    stream_destroy(m->stream[0], s);
    stream_destroy(m->stream[1], s);
    delete m;
}

void plan_invalidate(csrk_matrix *m, cudaStream_t s)
{
    std::lock_guard<std::mutex> g(m->mu);
    plan_destroy(m->plan, s);
    m->plan = nullptr;
    for (int k = 0; k < 2; k++) {
        stream_destroy(m->stream[k], s);
        m->stream[k] = nullptr;
        m->stream_failed[k] = false;
    }
}

static int check_shape(int32_t nrows, int32_t ncols, int64_t nnz, int rp_is64, int val_kind)
{
    CSRK_ARG(nrows >= 0 && ncols >= 0 && nnz >= 0, "negative dimension (nrows=%d ncols=%d nnz=%lld)", nrows, ncols,
             (long long)nnz);
    CSRK_ARG(val_kind == 0 || val_kind == 4 || val_kind == 8, "val_kind must be 0, 4 or 8 (got %d)", val_kind);
    CSRK_ARG(rp_is64 || nnz <= INT32_MAX, "int32 rowptrs cannot address %lld entries", (long long)nnz);
    return CSRK_OK;
}

template <typename RPT> __global__ void k_rebase(const RPT *src, RPT *dst, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = src[i] - src[0];
}

}  // namespace csrk

using namespace csrk;

extern "C" {

int csrk_version(void) { return 100; }

const char *csrk_last_error(void) { return t_error.c_str(); }

int csrk_init(int device)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> g(c.mu);
    return init_locked(c, device);
}

int csrk_shutdown(void)
{
    Context &c = ctx();
    std::lock_guard<std::mutex> g(c.mu);
    if (!c.inited)
        return CSRK_OK;
    CSRK_CUDA(cudaSetDevice(c.device));
    CSRK_CUDA(cudaStreamSynchronize(c.stream));
    ws_release_all();
    if (g_out_pool_release)
        g_out_pool_release();
    CSRK_CUDA(cudaStreamDestroy(c.stream));
    c.stream = nullptr;
    c.inited = false;
    return CSRK_OK;
}

int csrk_device_info(int *sm_count, int64_t *mem_total, int64_t *mem_free, int *cc_major, int *cc_minor)
{
    CSRK_TRY(ensure_init());
    size_t f = 0, t = 0;
    CSRK_CUDA(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx().sm_count;
    if (mem_total) *mem_total = (int64_t)t;
    if (mem_free) *mem_free = (int64_t)f;
    if (cc_major) *cc_major = ctx().cc_major;
    if (cc_minor) *cc_minor = ctx().cc_minor;
    return CSRK_OK;
}

int64_t csrk_launch_count(void) { return g_launches.load(); }

int csrk_set_option(const char *name, int64_t value)
{
    CSRK_ARG(name != nullptr, "option name is NULL");
    if (!strcmp(name, "spmv_mode")) {
        CSRK_ARG(value >= 0 && value <= 2, "spmv_mode must be 0 (auto), 1 (CSR tile kernel) or 2 (slab kernel)");
        options().spmv_mode = value;
    } else if (!strcmp(name, "stream_min_nnz")) {
        options().stream_min_nnz = value;
    } else if (!strcmp(name, "stream_slab_bytes")) {
        CSRK_ARG(value >= 0, "stream_slab_bytes must be >= 0");
        options().stream_slab_bytes = value;
    } else if (!strcmp(name, "stream_ctas")) {
        CSRK_ARG(value >= 0, "stream_ctas must be >= 0");
        options().stream_ctas = value;
    } else if (!strcmp(name, "stream_warps")) {
        CSRK_ARG(value >= 1 && value <= 31, "stream_warps must be 1..31");
        options().stream_warps = value;
    } else if (!strcmp(name, "stream_piece")) {
        CSRK_ARG(value >= 8 && value <= 4096, "stream_piece must be 8..4096");
        options().stream_piece = value;
    } else if (!strcmp(name, "stream_ring_bytes")) {
        CSRK_ARG(value == 4096 || value == 8192, "stream_ring_bytes must be 4096 or 8192");
        options().stream_ring_bytes = value;
    } else if (!strcmp(name, "stream_ring_chunks")) {
        CSRK_ARG(value == 2 || value == 4, "stream_ring_chunks must be 2 or 4");
        options().stream_ring_chunks = value;
    } else if (!strcmp(name, "sym_bytes")) {
        options().sym_bytes = value ? 1 : 0;
    } else if (!strcmp(name, "stream_xbufs")) {
        CSRK_ARG(value == 2 || value == 3, "stream_xbufs must be 2 or 3");
        options().stream_xbufs = value;
    } else if (!strcmp(name, "radix_bits")) {
        CSRK_ARG(value == 0 || value == 8 || value == 9, "radix_bits must be 0, 8 or 9");
        options().radix_bits = value;
    } else if (!strcmp(name, "spmv_zero_copy_y")) {
        options().spmv_zero_copy_y = value ? 1 : 0;
    } else if (!strcmp(name, "fix_threads")) {
        CSRK_ARG(value == 512 || value == 768 || value == 1024, "fix_threads must be 512, 768 or 1024");
        options().fix_threads = value;
    } else if (!strcmp(name, "spgemm_fixed")) {
        options().spgemm_fixed = value ? 1 : 0;
    } else if (!strcmp(name, "own_chunk_prod")) {
        options().own_chunk_prod = value;
    } else if (!strcmp(name, "fix_tiny_cap")) {
        CSRK_ARG(value >= 0, "fix_tiny_cap is an entry count (0 = 1/32 of the products)");
        options().fix_tiny_cap = value;
    } else if (!strcmp(name, "spgemm_esc")) {
        CSRK_ARG(value >= 0 && value <= 2, "spgemm_esc must be 0 (off), 1 (wide results) or 2 (always)");
        options().spgemm_esc = value;
    } else if (!strcmp(name, "esc_target")) {
        CSRK_ARG(value >= 16 && value <= 4096, "esc_target must be in 16..4096");
        options().esc_target = value;
    } else if (!strcmp(name, "esc_stride")) {
        options().esc_stride = value ? 1 : 0;
    } else if (!strcmp(name, "esc_budget")) {
        CSRK_ARG(value >= 0, "esc_budget is a byte count (0 = what the device has free)");
        options().esc_budget = value;
    } else if (!strcmp(name, "own_nw")) {
        CSRK_ARG(value == 8 || value == 16, "own_nw must be 8 or 16");
        options().own_nw = value;
    } else {
        set_error("unknown option `%s`", name);
        return CSRK_EARG;
    }
    return CSRK_OK;
}

int csrk_get_stream(void **stream)
{
    CSRK_ARG(stream != nullptr, "stream pointer is NULL");
    CSRK_TRY(ensure_init());
    *stream = (void *)ctx().stream;
    return CSRK_OK;
}

int csrk_synchronize(void)
{
    CSRK_TRY(ensure_init());
    CSRK_CUDA(cudaStreamSynchronize(ctx().stream));
    return CSRK_OK;
}

// `s` is the stream the copies run on.  The handle's memory always comes from the LIBRARY stream's pool
// allocation point (every later operation on the handle, and its cudaFreeAsync, run there), so for a
// caller-owned stream (csrk_create_dev) the two streams are ordered explicitly: s waits for the allocation,
// the library stream waits for the copies.
static int create_impl(int32_t nrows, int32_t ncols, int64_t nnz, const void *rowptrs, int rp_is64,
                       const int32_t *colinds, const void *values, int val_kind, cudaStream_t s, cudaMemcpyKind kind,
                       csrk_h *out)
{
    CSRK_ARG(out != nullptr, "out handle pointer is NULL");
    *out = nullptr;
    CSRK_TRY(check_shape(nrows, ncols, nnz, rp_is64, val_kind));
    CSRK_ARG(rowptrs != nullptr, "rowptrs is NULL");
    CSRK_ARG(nnz == 0 || colinds != nullptr, "colinds is NULL");
    CSRK_ARG(val_kind == 0 || nnz == 0 || values != nullptr, "values is NULL but val_kind=%d", val_kind);
    cudaStream_t ls = ctx().stream;
    csrk_matrix *m = nullptr;
    CSRK_TRY(matrix_alloc(&m, nrows, ncols, nnz, rp_is64, val_kind, ls));
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaSuccess;
    if (s != ls) {
        e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaEventRecord(ev, ls);
        if (e == cudaSuccess)
            e = cudaStreamWaitEvent(s, ev, 0);
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->rp, rowptrs, ((size_t)nrows + 1) * (rp_is64 ? 8 : 4), kind, s);
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(m->ci, colinds, (size_t)nnz * 4, kind, s);
    if (e == cudaSuccess && nnz && val_kind)
        e = cudaMemcpyAsync(m->vs, values, (size_t)nnz * val_kind, kind, s);
    if (e == cudaSuccess && s != ls) {
        e = cudaEventRecord(ev, s);
        if (e == cudaSuccess)
            e = cudaStreamWaitEvent(ls, ev, 0);
    }
    if (ev)
        (void)cudaEventDestroy(ev);
    if (e == cudaSuccess && kind == cudaMemcpyHostToDevice)
        e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        (void)cudaStreamSynchronize(s);
        matrix_destroy(m, ls);
        return cuda_fail(e, "upload", __FILE__, __LINE__);
    }
    *out = m;
    return CSRK_OK;
}

int csrk_create(int32_t nrows, int32_t ncols, int64_t nnz, const void *rowptrs, int rp_is64, const int32_t *colinds,
                const void *values, int val_kind, csrk_h *out)
{
    CSRK_TRY(ensure_init());
    return create_impl(nrows, ncols, nnz, rowptrs, rp_is64, colinds, values, val_kind, ctx().stream,
                       cudaMemcpyHostToDevice, out);
}

int csrk_create_dev(int32_t nrows, int32_t ncols, int64_t nnz, const void *d_rowptrs, int rp_is64,
                    const int32_t *d_colinds, const void *d_values, int val_kind, void *stream, csrk_h *out)
{
    CSRK_TRY(ensure_init());
    cudaStream_t s = (cudaStream_t)stream;
    return create_impl(nrows, ncols, nnz, d_rowptrs, rp_is64, d_colinds, d_values, val_kind, s,
                       cudaMemcpyDeviceToDevice, out);
}

int csrk_free(csrk_h h)
{
    if (!h)
        return CSRK_OK;
    CSRK_TRY(ensure_init());
    matrix_destroy(h, ctx().stream);
    return CSRK_OK;
}

int csrk_dims(csrk_h h, int32_t *nrows, int32_t *ncols, int64_t *nnz, int *rp_is64, int *val_kind)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    if (nrows) *nrows = h->nrows;
    if (ncols) *ncols = h->ncols;
    if (nnz) *nnz = h->nnz;
    if (rp_is64) *rp_is64 = h->rp_is64;
    if (val_kind) *val_kind = h->val_kind;
    return CSRK_OK;
}

// Large device -> PAGEABLE host copies.  cudaMemcpy into pageable memory runs at ~4 GB/s (the driver stages through
// its own buffer and one thread takes the page faults of a freshly allocated destination): the 19.6 GB result of
// configs[2] took 4.9 s to reach its NumPy arrays.  Here a few host threads each own a pinned slot and a stream: DMA
// a chunk into the slot, copy it into place, next chunk -- while some threads copy (and fault pages in) the
// others' DMAs run.  Pinned destinations and small copies take the plain path.
namespace {
constexpr size_t OUT_SLOT = (size_t)16 << 20;
constexpr int OUT_THREADS = 8;
struct OutPoolState {
    char *pin = nullptr;
    cudaStream_t st[OUT_THREADS] = {};
    bool ok = false, tried = false;
};
struct OutPool : OutPoolState {
    std::mutex mu;
    OutPool &operator=(const OutPoolState &o)
    {
        static_cast<OutPoolState &>(*this) = o;
        return *this;
    }
};
OutPool &out_pool()
{
    static OutPool p;
    return p;
}
}  // namespace

static void out_pool_release()
{
    OutPool &P = out_pool();
    std::lock_guard<std::mutex> g(P.mu);
    if (P.pin)
        (void)cudaFreeHost(P.pin);
    for (int k = 0; k < OUT_THREADS; k++)
        if (P.st[k])
            (void)cudaStreamDestroy(P.st[k]);
    P = OutPoolState();
}

static int copy_to_host(void *dst, const void *src, size_t bytes, cudaStream_t s)
{
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    OutPool &P = out_pool();
    if (pinned || bytes < 8 * OUT_SLOT) {
        CSRK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        return CSRK_OK;
    }
    std::lock_guard<std::mutex> g(P.mu);
    if (!P.tried) {
        P.tried = true;
        g_out_pool_release = out_pool_release;
        P.ok = cudaHostAlloc((void **)&P.pin, OUT_SLOT * OUT_THREADS, cudaHostAllocDefault) == cudaSuccess;
        for (int k = 0; P.ok && k < OUT_THREADS; k++)
            P.ok = cudaStreamCreateWithFlags(&P.st[k], cudaStreamNonBlocking) == cudaSuccess;
        (void)cudaGetLastError();
    }
    if (!P.ok) {
        CSRK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        return CSRK_OK;
    }
    CSRK_CUDA(cudaStreamSynchronize(s));   // what is copied has been produced on the library stream
    const size_t nchunk = (bytes + OUT_SLOT - 1) / OUT_SLOT;
    std::atomic<int> failed{0};
    const int device = ctx().device;
    auto work = [&](int k) {
        if (cudaSetDevice(device) != cudaSuccess) {
            failed = 1;
            return;
        }
        char *slot = P.pin + (size_t)k * OUT_SLOT;
        for (size_t c = (size_t)k; c < nchunk && !failed; c += OUT_THREADS) {
            const size_t off = c * OUT_SLOT, n = std::min(OUT_SLOT, bytes - off);
            if (cudaMemcpyAsync(slot, (const char *)src + off, n, cudaMemcpyDeviceToHost, P.st[k]) != cudaSuccess ||
                cudaStreamSynchronize(P.st[k]) != cudaSuccess) {
                failed = 1;
                return;
            }
            memcpy((char *)dst + off, slot, n);
        }
    };
    std::thread th[OUT_THREADS];
    int started = 0;
    try {
        for (; started < OUT_THREADS; started++)
            th[started] = std::thread(work, started);
    } catch (...) {   // no more threads to be had: the chunks of the missing ones are copied here, afterwards
    }
    for (int k = 0; k < started; k++)
        th[k].join();
    for (int k = started; k < OUT_THREADS; k++)
        work(k);
    if (failed) {
        set_error("device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return CSRK_ECUDA;
    }
    return CSRK_OK;
}

int csrk_export(csrk_h h, void *rowptrs, int32_t *colinds, void *values)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_ARG(rowptrs != nullptr, "rowptrs is NULL");
    CSRK_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    CSRK_CUDA(cudaMemcpyAsync(rowptrs, h->rp, ((size_t)h->nrows + 1) * (h->rp_is64 ? 8 : 4), cudaMemcpyDeviceToHost, s));
    if (h->nnz) {
        CSRK_ARG(colinds != nullptr, "colinds is NULL");
        CSRK_TRY(copy_to_host(colinds, h->ci, (size_t)h->nnz * 4, s));
        if (h->val_kind) {
            CSRK_ARG(values != nullptr, "values is NULL");
            CSRK_TRY(copy_to_host(values, h->vs, (size_t)h->nnz * h->val_kind, s));
        }
    }
    CSRK_CUDA(cudaStreamSynchronize(s));
    return CSRK_OK;
}

int csrk_device_ptrs(csrk_h h, void **d_rowptrs, int32_t **d_colinds, void **d_values)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    if (d_rowptrs) *d_rowptrs = h->rp;
    if (d_colinds) *d_colinds = h->ci;
    if (d_values) *d_values = h->vs;
    return CSRK_OK;
}

int csrk_subset_rows(csrk_h h, int32_t begin, int32_t end, csrk_h *out)
{
    CSRK_ARG(h != nullptr && out != nullptr, "NULL handle");
    CSRK_ARG(0 <= begin && begin <= end && end <= h->nrows, "row range [%d,%d) outside [0,%d]", begin, end, h->nrows);
    CSRK_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    *out = nullptr;
    // the two bounding row pointers decide the slice of colinds/values
    int64_t st = 0, ed = 0;
    const size_t w = h->rp_is64 ? 8 : 4;
    int64_t tmp[2] = {0, 0};
    CSRK_CUDA(cudaMemcpyAsync(&tmp[0], (const char *)h->rp + (size_t)begin * w, w, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&tmp[1], (const char *)h->rp + (size_t)end * w, w, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    if (h->rp_is64) {
        st = tmp[0];
        ed = tmp[1];
    } else {
        int32_t a32, b32;
        memcpy(&a32, &tmp[0], 4);
        memcpy(&b32, &tmp[1], 4);
        st = a32;
        ed = b32;
    }
    const int64_t nnz = ed - st;
    const int32_t nr = end - begin;
    csrk_matrix *m = nullptr;
    // keep the parent's rowptr width (subset_rows slices the same array: structure.py:72-74)
    CSRK_TRY(matrix_alloc(&m, nr, h->ncols, nnz, h->rp_is64, h->val_kind, s));
    const unsigned grid = (unsigned)div_up((int64_t)nr + 1, 256);
    cudaError_t e = cudaSuccess;
    if (h->rp_is64)
        k_rebase<int64_t><<<grid, 256, 0, s>>>((const int64_t *)h->rp + begin, (int64_t *)m->rp, (int64_t)nr + 1);
    else
        k_rebase<int32_t><<<grid, 256, 0, s>>>((const int32_t *)h->rp + begin, (int32_t *)m->rp, (int64_t)nr + 1);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
    if (e == cudaSuccess && nnz)
        e = cudaMemcpyAsync(m->ci, h->ci + st, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && nnz && h->val_kind)
        e = cudaMemcpyAsync(m->vs, (const char *)h->vs + (size_t)st * h->val_kind, (size_t)nnz * h->val_kind,
                            cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        matrix_destroy(m, s);
        return cuda_fail(e, "subset_rows", __FILE__, __LINE__);
    }
    *out = m;
    return CSRK_OK;
}

int csrk_spmv_dev(csrk_h h, const void *d_x, int x_kind, double *d_y, void *stream)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_ARG(x_kind == 4 || x_kind == 8, "x_kind must be 4 or 8 (got %d)", x_kind);
    CSRK_ARG(h->ncols == 0 || d_x != nullptr, "x is NULL");
    CSRK_ARG(h->nrows == 0 || d_y != nullptr, "y is NULL");
    CSRK_TRY(ensure_init());
    cudaStream_t s = (cudaStream_t)stream;
    return spmv_run(h, d_x, x_kind, d_y, s);
}

int csrk_spmv_plan_info(csrk_h h, int x_kind, int64_t info[12])
{
    CSRK_ARG(h != nullptr && info != nullptr, "NULL argument");
    CSRK_ARG(x_kind == 4 || x_kind == 8, "x_kind must be 4 or 8 (got %d)", x_kind);
    for (int i = 0; i < 12; i++)
        info[i] = 0;
    std::lock_guard<std::mutex> g(h->mu);
    if (StreamPlan *p = h->stream[x_kind == 4 ? 0 : 1]) {
        info[0] = 1;
        stream_info(p, info + 1);
    }
    return CSRK_OK;
}

int csrk_spmv_dev_multi(csrk_h h, const void *d_x, int x_kind, double *const *d_ys, int n_out, void *stream)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_ARG(x_kind == 4 || x_kind == 8, "x_kind must be 4 or 8 (got %d)", x_kind);
    CSRK_ARG(n_out >= 1 && n_out <= 8 && d_ys != nullptr, "n_out must be 1..8 (got %d)", n_out);
    for (int k = 0; k < n_out; k++)
        CSRK_ARG(h->nrows == 0 || d_ys[k] != nullptr, "y[%d] is NULL", k);
    CSRK_ARG(h->ncols == 0 || d_x != nullptr, "x is NULL");
    CSRK_TRY(ensure_init());
    return spmv_run_multi(h, d_x, x_kind, d_ys, n_out, (cudaStream_t)stream);
}

int csrk_spmv_dev_mc(csrk_h h, const void *d_x, int x_kind, double *d_y, double *d_y_mc, void *stream)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_ARG(x_kind == 4 || x_kind == 8, "x_kind must be 4 or 8 (got %d)", x_kind);
    CSRK_ARG(h->nrows == 0 || (d_y != nullptr && d_y_mc != nullptr), "y / multicast y is NULL");
    CSRK_ARG(((uintptr_t)d_y_mc & 7) == 0, "multicast y must be 8-byte aligned");
    CSRK_ARG(h->ncols == 0 || d_x != nullptr, "x is NULL");
    CSRK_TRY(ensure_init());
    double *ys[2] = {d_y, d_y_mc};
    return spmv_run_multi(h, d_x, x_kind, ys, 2, (cudaStream_t)stream, true);
}

int csrk_mc_broadcast(void *mc_dst, const void *d_src, int64_t nbytes, void *stream)
{
    CSRK_ARG(nbytes >= 0 && nbytes % 4 == 0, "nbytes must be a non-negative multiple of 4 (got %lld)", (long long)nbytes);
    CSRK_ARG(nbytes == 0 || (mc_dst != nullptr && d_src != nullptr), "NULL pointer");
    CSRK_ARG((((uintptr_t)mc_dst | (uintptr_t)d_src) & 15) == 0, "multicast copies need 16-byte aligned pointers");
    CSRK_TRY(ensure_init());
    return mc_broadcast_run(mc_dst, d_src, nbytes, (cudaStream_t)stream);
}

int csrk_spmv(csrk_h h, const void *x, int x_kind, double *y)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_ARG(x_kind == 4 || x_kind == 8, "x_kind must be 4 or 8 (got %d)", x_kind);
    CSRK_ARG(h->ncols == 0 || x != nullptr, "x is NULL");
    CSRK_ARG(h->nrows == 0 || y != nullptr, "y is NULL");
    CSRK_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    DevBuf dx, dy;
    CSRK_TRY(dx.alloc((size_t)h->ncols * x_kind, s));
    CSRK_TRY(dy.alloc((size_t)h->nrows * 8, s));
    if (h->ncols)
        CSRK_CUDA(cudaMemcpyAsync(dx.p, x, (size_t)h->ncols * x_kind, cudaMemcpyHostToDevice, s));
    // y in pinned (mapped) host memory: the tile kernel stores every finished row straight into it over PCIe
    // as a second output, so the 8 B/row device-to-host copy overlaps the compute instead of following it.
    // (The slab kernel finishes all rows at its very end, in bin order: a plain copy after it is faster.)
    double *y_mapped = nullptr;
    if (h->nrows && h->nnz && options().spmv_zero_copy_y.load() && !spmv_uses_slab(h, x_kind, dx.p)) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, y) == cudaSuccess) {
            if (at.type == cudaMemoryTypeHost && at.devicePointer != nullptr)
                y_mapped = (double *)at.devicePointer;
        } else {
            (void)cudaGetLastError();
        }
    }
    if (y_mapped) {
        double *ys[2] = {dy.as<double>(), y_mapped};
        CSRK_TRY(spmv_run_multi(h, dx.p, x_kind, ys, 2, s, false));
    } else {
        CSRK_TRY(spmv_run(h, dx.p, x_kind, dy.as<double>(), s));
        if (h->nrows)
            CSRK_CUDA(cudaMemcpyAsync(y, dy.p, (size_t)h->nrows * 8, cudaMemcpyDeviceToHost, s));
    }
    CSRK_CUDA(cudaStreamSynchronize(s));
    return CSRK_OK;
}

int csrk_spgemm(csrk_h a, csrk_h b, csrk_h *c)
{
    CSRK_ARG(a && b && c, "NULL handle");
    CSRK_ARG(a->ncols == b->nrows, "mult_ab: a.ncols (%d) != b.nrows (%d)", a->ncols, b->nrows);
    CSRK_TRY(ensure_init());
    WsScope scope;
    return spgemm_run(a, b, c, ctx().stream);
}

int csrk_spgemm_abt(csrk_h a, csrk_h b, csrk_h *c)
{
    CSRK_ARG(a && b && c, "NULL handle");
    CSRK_ARG(a->ncols == b->ncols, "mult_abt: a.ncols (%d) != b.ncols (%d)", a->ncols, b->ncols);
    CSRK_TRY(ensure_init());
    cudaStream_t s = ctx().stream;
    WsScope scope;
    // multiply.py:56-57: bt = b.transpose(); mult_ab(a, bt)
    csrk_matrix *bt = nullptr;
    CSRK_TRY(transpose_run(b, 1, &bt, s));
    int rc = spgemm_run(a, bt, c, s);
    matrix_destroy(bt, s);
    return rc;
}

int csrk_spgemm_stats(csrk_h c, int64_t *products, int64_t *out_nnz)
{
    CSRK_ARG(c != nullptr, "NULL handle");
    if (products) *products = c->stat_products;
    if (out_nnz) *out_nnz = c->stat_out_nnz;
    return CSRK_OK;
}

int csrk_spgemm_path(csrk_h c, int *path)
{
    CSRK_ARG(c != nullptr && path != nullptr, "NULL argument");
    *path = c->stat_path;
    return CSRK_OK;
}

int csrk_spgemm_side_list(csrk_h c, int64_t *entries)
{
    CSRK_ARG(c != nullptr && entries != nullptr, "NULL argument");
    *entries = c->stat_tiny;
    return CSRK_OK;
}

int csrk_transpose(csrk_h a, int with_values, csrk_h *at)
{
    CSRK_ARG(a && at, "NULL handle");
    CSRK_TRY(ensure_init());
    WsScope scope;
    return transpose_run(a, with_values, at, ctx().stream);
}

int csrk_order_columns(csrk_h h)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_TRY(ensure_init());
    WsScope scope;
    return order_columns_run(h, ctx().stream);
}

int csrk_from_coo(int32_t nrows, int32_t ncols, int64_t nnz, const int32_t *rows, const int32_t *cols, const void *values,
                  int val_kind, csrk_h *out)
{
    CSRK_ARG(out != nullptr, "out is NULL");
    *out = nullptr;
    CSRK_ARG(nrows >= 0 && ncols >= 0 && nnz >= 0, "negative dimension");
    CSRK_ARG(val_kind == 0 || val_kind == 4 || val_kind == 8, "val_kind must be 0, 4 or 8 (got %d)", val_kind);
    CSRK_ARG(nnz == 0 || (rows != nullptr && cols != nullptr), "rows / cols is NULL");
    CSRK_ARG(nnz == 0 || val_kind == 0 || values != nullptr, "values is NULL");
    CSRK_ARG(nnz == 0 || (nrows > 0 && ncols > 0), "entries in an empty shape");
    CSRK_TRY(ensure_init());
    WsScope scope;
    cudaStream_t s = ctx().stream;
    DevBuf dr, dc, dv;
    if (nnz) {
        CSRK_TRY(dr.alloc((size_t)nnz * 4, s));
        CSRK_TRY(dc.alloc((size_t)nnz * 4, s));
        CSRK_CUDA(cudaMemcpyAsync(dr.p, rows, (size_t)nnz * 4, cudaMemcpyHostToDevice, s));
        CSRK_CUDA(cudaMemcpyAsync(dc.p, cols, (size_t)nnz * 4, cudaMemcpyHostToDevice, s));
        if (val_kind) {
            CSRK_TRY(dv.alloc((size_t)nnz * val_kind, s));
            CSRK_CUDA(cudaMemcpyAsync(dv.p, values, (size_t)nnz * val_kind, cudaMemcpyHostToDevice, s));
        }
    }
    csrk_matrix *m = nullptr;
    CSRK_TRY(from_coo_run(nrows, ncols, nnz, dr.as<int32_t>(), dc.as<int32_t>(), dv.p, val_kind, &m, s));
    *out = m;
    return CSRK_OK;
}

int csrk_normalize_rows(csrk_h h, int kind, void *vec, void *values_out)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_ARG(kind == 0 || kind == 1, "kind must be 0 (center) or 1 (unit), got %d", kind);
    CSRK_ARG(h->val_kind == 4 || h->val_kind == 8, "normalize_rows needs a matrix with values");
    CSRK_ARG(h->nrows == 0 || vec != nullptr, "output vector is NULL");
    CSRK_TRY(ensure_init());
    WsScope scope;
    cudaStream_t s = ctx().stream;
    DevBuf dvec;
    const size_t bytes = (size_t)h->nrows * (size_t)h->val_kind;
    CSRK_TRY(dvec.alloc(bytes ? bytes : 8, s));
    CSRK_TRY(normalize_rows_run(h, kind, dvec.p, s));
    if (bytes)
        CSRK_CUDA(cudaMemcpyAsync(vec, dvec.p, bytes, cudaMemcpyDeviceToHost, s));
    if (values_out && h->nnz)
        CSRK_CUDA(cudaMemcpyAsync(values_out, h->vs, (size_t)h->nnz * h->val_kind, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    return CSRK_OK;
}

int csrk_filter_zeros(csrk_h h)
{
    CSRK_ARG(h != nullptr, "NULL handle");
    CSRK_TRY(ensure_init());
    WsScope scope;
    return filter_zeros_run(h, ctx().stream);
}

}  // extern "C"
