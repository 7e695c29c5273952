// spmv.cu -- mult_vec (csr/kernels/numba/__init__.py:55-67) for sm_100a.
//
//   y[r] = sum_{i in row r} x[colinds[i]] * (values[i] or 1),   y float64.
//
// Work decomposition ("nnz-balanced row binning"): the nnz axis is cut into
// tiles of SPMV_TILE entries; a tile map built once per handle stores, for each
// tile, the first row that STARTS inside it (binary search on rowptrs).  One CTA
// owns one tile, so every CTA streams the same number of bytes no matter how
// skewed the row lengths are (a 1M-entry row is simply 244 tiles).
//
// Inside a tile
//   phase 1  all threads stream colinds/values with 128-bit no-L1-allocate loads,
//            gather x through the read-only path and park the products in shared
//            memory (product type = numba's promotion: f4*f4 -> f4, else f8);
//   phase 2  rows that start in the tile are reduced from shared memory in
//            float64: scalar-per-row for short rows, vector(warp)-per-row with a
//            shuffle reduction for long ones; the head of the tile that belongs
//            to a row begun in an earlier tile is block-reduced into carry[tile].
// A tiny fix-up kernel adds the carries of multi-tile rows in tile order, so the
// result is deterministic (no float atomics anywhere).
//
// Algorithmic HBM bytes per nnz (f4 values, f4 x, int32 rowptrs):
//   4 (colind) + 4 (value) + (nrows+1)*4/nnz + ncols*4/nnz + nrows*8/nnz  ~ 8.16 B at cfg2.
#include <type_traits>

#include "common.cuh"
#include <cstdio>

#include "spmv.cuh"

namespace csrk {

constexpr int SPMV_BLOCK = 256;
constexpr int SPMV_ITEMS = 16;
constexpr int SPMV_TILE = SPMV_BLOCK * SPMV_ITEMS;  // 4096 nnz per CTA
constexpr int SPMV_LONG = 48;                       // rows longer than this inside a tile go to a warp
constexpr int SPMV_QCAP = SPMV_TILE / SPMV_LONG + 2;

struct SpmvPlan {
    int64_t ntiles = 0;
    int32_t *tile_row = nullptr;  // [ntiles+1] first row with rowptr >= tile base; [ntiles] = nrows
};

void plan_destroy(SpmvPlan *p, cudaStream_t s)
{
    if (!p)
        return;
    dev_free(p->tile_row, s);
    delete p;
}

template <typename RPT>
__global__ void k_spmv_plan(const RPT *__restrict__ rp, int32_t nrows, int64_t ntiles, int32_t *__restrict__ tile_row)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles)
        return;
    if (t == ntiles) {
        tile_row[t] = nrows;
        return;
    }
    tile_row[t] = (int32_t)lower_bound_rp(rp, 0, (int64_t)nrows + 1, t * SPMV_TILE);
}

__global__ void k_zero_f64(double *y, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        y[i] = 0.0;
}

template <typename RPT, typename VT, typename XT, bool MULTI>
__global__ void __launch_bounds__(SPMV_BLOCK)
k_spmv_tile(int32_t nrows, int64_t nnz, const RPT *__restrict__ rp, const int32_t *__restrict__ ci,
            const VT *__restrict__ vs, const XT *__restrict__ x, YOut y,
            const int32_t *__restrict__ tile_row, double *__restrict__ carry)
{
    constexpr bool HASV = !std::is_same<VT, NoVal>::value;
    using RV = typename std::conditional<HASV, VT, double>::type;  // real value type
    using PT = typename Prod<RV, XT>::type;
    __shared__ __align__(16) PT prod[SPMV_TILE];
    __shared__ double wsum[SPMV_BLOCK / 32];
    __shared__ int q_row[SPMV_QCAP];
    __shared__ int q_n;

    const int tid = threadIdx.x;
    const int64_t t = blockIdx.x;
    const int64_t base = t * SPMV_TILE;
    const int cnt = (int)min((int64_t)SPMV_TILE, nnz - base);
    if (tid == 0)
        q_n = 0;

    // ---- phase 1: stream + gather + multiply
    // Lane i of a warp takes entry 32k+i: ADJACENT LANES HOLD ADJACENT ENTRIES.  The x gather is
    // bound by the L1 tag stage (one 128-byte line per cycle per SM, profiles/r01_spmv_v1_ncu.md),
    // and in a long row consecutive entries are a few columns apart, so a warp-wide gather then
    // touches C/L lines instead of 32.  The index/value streams are fully coalesced 128-byte
    // warp loads (no L1 allocation); all of a thread's loads are issued before the first gather.
    if (cnt == SPMV_TILE) {
        int c[SPMV_ITEMS];
        RV v[SPMV_ITEMS];
#pragma unroll
        for (int k = 0; k < SPMV_ITEMS; k++) {
            const int64_t e = base + k * SPMV_BLOCK + tid;
            c[k] = ld_stream_i32(ci + e);
            if constexpr (HASV)
                v[k] = ld_stream<RV>(reinterpret_cast<const RV *>(vs) + e);
        }
#pragma unroll
        for (int k = 0; k < SPMV_ITEMS; k++) {
            const XT xv = __ldg(x + c[k]);
            if constexpr (HASV)
                prod[k * SPMV_BLOCK + tid] = (PT)xv * (PT)v[k];
            else
                prod[k * SPMV_BLOCK + tid] = (PT)xv;
        }
    } else {
        for (int i = tid; i < cnt; i += SPMV_BLOCK) {
            const int64_t e = base + i;
            XT xv = __ldg(x + ci[e]);
            if constexpr (HASV)
                prod[i] = (PT)xv * (PT) reinterpret_cast<const RV *>(vs)[e];
            else
                prod[i] = (PT)xv;
        }
    }
    __syncthreads();

    // ---- phase 2: segmented reduction in float64
    const int32_t rlo = tile_row[t], rhi = tile_row[t + 1];
    const int64_t tend = base + cnt;
    // head of the tile that continues a row begun earlier (rp[rlo] is valid: rlo <= nrows)
    const int lead = (int)(min((int64_t)rp[rlo], tend) - base);
    if (lead > 0) {
        double s = 0.0;
        for (int i = tid; i < lead; i += SPMV_BLOCK)
            s += (double)prod[i];
        s = warp_sum(s);
        if ((tid & 31) == 0)
            wsum[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < SPMV_BLOCK / 32; w++)
                tot += wsum[w];
            carry[t] = tot;
        }
    } else if (tid == 0) {
        carry[t] = 0.0;
    }

    // scalar-per-row for short rows; long rows are queued for the warps
    for (int32_t r = rlo + tid; r < rhi; r += SPMV_BLOCK) {
        const int s0 = (int)((int64_t)rp[r] - base);
        const int64_t re = (int64_t)rp[r + 1];
        const int e0 = (int)(min(re, tend) - base);
        if (e0 - s0 > SPMV_LONG) {
            int slot = atomicAdd(&q_n, 1);
            q_row[slot] = r;
        } else {
            double s = 0.0;
            for (int i = s0; i < e0; i++)
                s += (double)prod[i];
            store_y<MULTI>(y, r, s, re <= tend);  // complete row, or the head piece of a row that continues (carries are added later)
        }
    }
    __syncthreads();
    const int nq = q_n;
    for (int q = tid >> 5; q < nq; q += SPMV_BLOCK / 32) {
        const int32_t r = q_row[q];
        const int s0 = (int)((int64_t)rp[r] - base);
        const int64_t re = (int64_t)rp[r + 1];
        const int e0 = (int)(min(re, tend) - base);
        double s = 0.0;
        for (int i = s0 + (tid & 31); i < e0; i += 32)
            s += (double)prod[i];
        s = warp_sum(s);
        if ((tid & 31) == 0)
            store_y<MULTI>(y, r, s, re <= tend);
    }
}

// One warp per tile t >= 1.  If tile t is the FIRST continuation tile of the row that spills into it, add
// that row's carries: lanes take every 32nd tile, then a shuffle tree -- a fixed order, so deterministic
// (a 1M-entry row has 244 carries: one thread adding them in sequence took 16 us of a 0.36 ms step).
template <typename RPT, bool MULTI>
__global__ void __launch_bounds__(256)
k_spmv_fixup(int64_t ntiles, int64_t nnz, const RPT *__restrict__ rp,
             const int32_t *__restrict__ tile_row, const double *__restrict__ carry, YOut y)
{
    const int64_t t = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) + 1;
    const int lane = threadIdx.x & 31;
    if (t >= ntiles)
        return;
    const int64_t base = t * SPMV_TILE;
    const int32_t rlo = tile_row[t];
    if ((int64_t)rp[rlo] <= base)
        return;                    // no head piece
    const int32_t row = rlo - 1;   // the row spanning `base`
    const int64_t rs = (int64_t)rp[row];
    if (rs / SPMV_TILE != t - 1)
        return;                    // an earlier tile is this row's first continuation
    const int64_t re = (int64_t)rp[row + 1];
    const int64_t tlast = (re - 1) / SPMV_TILE;
    double tot = 0.0;
    for (int64_t u = t + lane; u <= tlast; u += 32)
        tot += carry[u];
    tot = warp_sum(tot);
    if (lane == 0)
        store_y<MULTI>(y, row, y.p[0][row] + tot, true);
}

static int ensure_plan(csrk_matrix *h, cudaStream_t s, SpmvPlan **out)
{
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->plan) {
        SpmvPlan *p = new (std::nothrow) SpmvPlan();
        if (!p) {
            set_error("host allocation failed");
            return CSRK_ENOMEM;
        }
        p->ntiles = div_up(h->nnz, SPMV_TILE);
        int rc = dev_alloc((void **)&p->tile_row, sizeof(int32_t) * (size_t)(p->ntiles + 1), s);
        if (rc != CSRK_OK) {
            delete p;
            return rc;
        }
        const unsigned grid = (unsigned)div_up(p->ntiles + 1, 256);
        if (h->rp_is64)
            k_spmv_plan<int64_t><<<grid, 256, 0, s>>>((const int64_t *)h->rp, h->nrows, p->ntiles, p->tile_row);
        else
            k_spmv_plan<int32_t><<<grid, 256, 0, s>>>((const int32_t *)h->rp, h->nrows, p->ntiles, p->tile_row);
        g_launches.fetch_add(1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            plan_destroy(p, s);
            return cuda_fail(e, "k_spmv_plan", __FILE__, __LINE__);
        }
        // the plan may be consumed on another stream (csrk_spmv_dev): make it visible first
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            plan_destroy(p, s);
            return cuda_fail(e, "plan sync", __FILE__, __LINE__);
        }
        h->plan = p;
    }
    *out = h->plan;
    return CSRK_OK;
}

template <typename RPT, typename VT, typename XT>
static int launch_spmv(csrk_matrix *h, SpmvPlan *p, const void *d_x, YOut d_y, double *carry, cudaStream_t s)
{
    if (d_y.n > 1) {
        CSRK_LAUNCH((k_spmv_tile<RPT, VT, XT, true>), (unsigned)p->ntiles, SPMV_BLOCK, 0, s, h->nrows, h->nnz,
                    (const RPT *)h->rp, h->ci, (const VT *)h->vs, (const XT *)d_x, d_y, p->tile_row, carry);
        if (p->ntiles > 1)
            CSRK_LAUNCH((k_spmv_fixup<RPT, true>), (unsigned)div_up((p->ntiles - 1) * 32, 256), 256, 0, s, p->ntiles, h->nnz,
                        (const RPT *)h->rp, p->tile_row, carry, d_y);
        return CSRK_OK;
    }
    CSRK_LAUNCH((k_spmv_tile<RPT, VT, XT, false>), (unsigned)p->ntiles, SPMV_BLOCK, 0, s, h->nrows, h->nnz,
                (const RPT *)h->rp, h->ci, (const VT *)h->vs, (const XT *)d_x, d_y, p->tile_row, carry);
    if (p->ntiles > 1)
        CSRK_LAUNCH((k_spmv_fixup<RPT, false>), (unsigned)div_up((p->ntiles - 1) * 32, 256), 256, 0, s, p->ntiles, h->nnz,
                    (const RPT *)h->rp, p->tile_row, carry, d_y);
    return CSRK_OK;
}

template <typename RPT, typename VT>
static int launch_spmv_x(csrk_matrix *h, SpmvPlan *p, const void *d_x, int x_kind, YOut d_y, double *carry,
                         cudaStream_t s)
{
    if (x_kind == 4)
        return launch_spmv<RPT, VT, float>(h, p, d_x, d_y, carry, s);
    return launch_spmv<RPT, VT, double>(h, p, d_x, d_y, carry, s);
}

template <typename RPT>
static int launch_spmv_v(csrk_matrix *h, SpmvPlan *p, const void *d_x, int x_kind, YOut d_y, double *carry,
                         cudaStream_t s)
{
    switch (h->val_kind) {
    case 4: return launch_spmv_x<RPT, float>(h, p, d_x, x_kind, d_y, carry, s);
    case 8: return launch_spmv_x<RPT, double>(h, p, d_x, x_kind, d_y, carry, s);
    default: return launch_spmv_x<RPT, NoVal>(h, p, d_x, x_kind, d_y, carry, s);
    }
}

// The slab kernel (spmv_slab.cu) stages x in shared memory and reads a re-laid-out copy of the
// entries; it pays G copies of x over the L2->SM fabric for never gathering x through L1/L2.  Auto mode
// takes it when that trade wins: a matrix large enough to be worth a plan (one stable sort of the entries,
// built on the first mult_vec of the handle) whose x, re-read once per SM, is smaller than its entry stream.
static bool stream_wanted(const csrk_matrix *h, int x_kind, const void *d_x)
{
    const int64_t mode = options().spmv_mode.load();
    if (mode == 1 || ((uintptr_t)d_x & 15) != 0)  // TMA bulk copies need a 16-byte aligned x
        return false;
    if (mode == 2)
        return true;
    if (h->nnz < options().stream_min_nnz.load())
        return false;
    // CSR.mult_vec of the reference makes a handle per call (csr/csr.py:582): a plan (10-60 ms) only pays for
    // handles that are used again, so the first call of a handle stays on the tile kernel
    if (h->spmv_calls.load(std::memory_order_relaxed) < 1 && !h->stream[x_kind == 4 ? 0 : 1])
        return false;
    // measured on 1M-row blocks of 100M nnz: x of 4 MB (reload = 0.7 x the stream) 0.193 ms against 0.362 ms for the
    // tile kernel, 8 MB (1.5 x) 0.260 against 0.366; 16 MB and more stay on the tile kernel
    const double reload = (double)ctx().sm_count * (double)h->ncols * x_kind;
    return reload <= 1.6 * (double)h->nnz * (4 + h->val_kind);
}

bool spmv_uses_slab(const csrk_matrix *h, int x_kind, const void *d_x)
{
    if (h->nrows == 0 || h->nnz == 0 || !stream_wanted(h, x_kind, d_x))
        return false;
    return !h->stream_failed[x_kind == 4 ? 0 : 1];
}

static int ensure_stream(csrk_matrix *h, int x_kind, StreamPlan **out)
{
    const int k = x_kind == 4 ? 0 : 1;
    *out = nullptr;
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->stream[k] && !h->stream_failed[k]) {
        StreamPlan *p = nullptr;
        const int rc = stream_build(h, x_kind, &p, ctx().stream);
        if (rc == CSRK_EOVERFLOW) {
            if (trace_enabled())
                fprintf(stderr, "[csrk] %s: CSR tile kernel instead\n", csrk_last_error());
            h->stream_failed[k] = true;  // not representable (too many rows per SM): stay on the CSR kernel
            return CSRK_OK;
        }
        if (rc != CSRK_OK)
            return rc;
        h->stream[k] = p;
    }
    *out = h->stream[k];
    return CSRK_OK;
}

int spmv_run_multi(csrk_matrix *h, const void *d_x, int x_kind, double *const *d_ys, int n_out, cudaStream_t s,
                   bool multicast)
{
    YOut yo;
    yo.n = n_out;
    yo.mc = multicast ? 1 : 0;
    for (int k = 0; k < SPMV_MAX_OUT; k++)
        yo.p[k] = k < n_out ? d_ys[k] : nullptr;
    if (h->nrows == 0)
        return CSRK_OK;
    if (h->nnz == 0) {
        // (a multicast address takes plain stores too: multimem.st is an ordinary STG in SASS)
        for (int k = 0; k < n_out; k++)
            CSRK_LAUNCH(k_zero_f64, (unsigned)div_up(h->nrows, 256), 256, 0, s, d_ys[k], (int64_t)h->nrows);
        return CSRK_OK;
    }
    const bool slab = stream_wanted(h, x_kind, d_x);
    h->spmv_calls.fetch_add(1, std::memory_order_relaxed);
    if (slab) {
        StreamPlan *sp = nullptr;
        CSRK_TRY(ensure_stream(h, x_kind, &sp));
        if (sp)
            return stream_run(h, sp, d_x, yo, s);
    }
    SpmvPlan *p = nullptr;
    CSRK_TRY(ensure_plan(h, ctx().stream, &p));
    DevBuf carry;
    CSRK_TRY(carry.alloc(sizeof(double) * (size_t)p->ntiles, s));
    if (h->rp_is64)
        return launch_spmv_v<int64_t>(h, p, d_x, x_kind, yo, carry.as<double>(), s);
    return launch_spmv_v<int32_t>(h, p, d_x, x_kind, yo, carry.as<double>(), s);
}

int spmv_run(csrk_matrix *h, const void *d_x, int x_kind, double *d_y, cudaStream_t s)
{
    double *ys[1] = {d_y};
    return spmv_run_multi(h, d_x, x_kind, ys, 1, s, false);
}

// ---- NVLS broadcast: copy `nbytes` from local memory to a multicast address -----------------
// (the x broadcast of the row-partitioned SpMV: the root stores once, the NVSwitch replicates)
// Measured (tools/exp_mc.py, 32 MB): 60 us = 530 GB/s at 2 GPUs, 84 us at 8 -- the same for weak or
// relaxed.sys stores, 1x/4x/8x unrolling and for the copy engine (cudaMemcpyAsync to the multicast
// address, 75 us): the multicast write rate of the fabric, not this kernel, sets it.
__global__ void __launch_bounds__(256) k_mc_copy(float *__restrict__ mc_dst, const float *__restrict__ src, int64_t n16,
                                                 int64_t nwords)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = i0; i < n16; i += stride) {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(src) + i);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + 4 * i), "f"(q.x),
                     "f"(q.y), "f"(q.z), "f"(q.w)
                     : "memory");
    }
    for (int64_t i = 4 * n16 + i0; i < nwords; i += stride)  // tail: < 4 words
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_dst + i), "f"(src[i]) : "memory");
}

int mc_broadcast_run(void *mc_dst, const void *src, int64_t nbytes, cudaStream_t s)
{
    if (nbytes == 0)
        return CSRK_OK;
    const int64_t nwords = nbytes / 4, n16 = nbytes / 16;
    const unsigned grid = (unsigned)std::min<int64_t>(div_up(std::max<int64_t>(n16, 1), 256), (int64_t)ctx().sm_count * 8);
    CSRK_LAUNCH(k_mc_copy, grid, 256, 0, s, (float *)mc_dst, (const float *)src, n16, nwords);
    return CSRK_OK;
}

}  // namespace csrk
