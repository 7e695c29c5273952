// spmv_psf.cu -- mult_vec for large matrices: the "panel/slab" SpMV.
//
// Why: with x gathered through L1, a CSR SpMV on B200 is bound by the L1 tag stage
// (one 128-byte line per cycle per SM ~ 1 gather/cycle/SM; measured 33 % of the HBM
// roofline at 1M x 1M / 100M nnz, profiles/r01_*).  Shared memory serves ~9 random
// 4-byte reads per cycle, so x is staged there slab by slab:
//
//   * columns are cut into SLABS of 64 KB of x (16384 f32 / 8192 f64 columns);
//   * rows are grouped into PANELS, one CTA at a time; the CTA walks the slabs in
//     order, a producer warp streaming x[slab] into a 2-deep shared-memory ring with
//     cp.async.bulk (TMA) + mbarriers while NW consumer warps work on the panel's
//     entries that fall in the resident slab;
//   * the panel's float64 row accumulators live in shared memory for the whole walk,
//     so y is written once per row and x is read from L2 once per panel.
//
// The entries are re-laid out ONCE per handle (lazily, at the first mult_vec) by a
// stable radix sort on (panel, slab); each entry is (row-slot << logW | column-in-slab,
// value): 4 + V bytes per nnz, the same stream volume as CSR.
//
// Two kinds of panel keep every warp busy without atomics and keep the result
// deterministic:
//   light panels  rows of <= TH entries (the bulk of the rows): up to 7168 rows; every
//                 warp OWNS a contiguous range of rows, so its accumulators are private;
//   heavy panels  rows of > TH entries (most of the nnz in a power-law matrix), split
//                 into chunks of <= CH entries: up to 7168/NW chunks; the entries of a
//                 (panel, slab) cell are split EVENLY over the warps, each warp has its
//                 own copy of the accumulators, and the copies are summed in warp order.
// Inside a 128-entry block (4 consecutive entries per lane, one 128-bit load each for
// the indices and the values) equal row-slots are combined by an in-lane pass and a
// warp-shuffle segmented scan; only the last lane of a run touches shared memory.
#include <algorithm>
#include <vector>

#include "radix.cuh"

namespace csrk {

constexpr int PSF_NW = 24;                        // consumer warps per CTA
constexpr int PSF_THREADS = (PSF_NW + 1) * 32;    // + one producer warp
constexpr int PSF_SLAB_BYTES = 64 * 1024;         // x bytes per slab
constexpr int PSF_PR = 7168;                      // float64 accumulator slots per CTA
constexpr int PSF_RE = PSF_PR / PSF_NW;           // row-chunks in a heavy (even-split) panel
constexpr size_t PSF_SMEM = 2 * (size_t)PSF_SLAB_BYTES + (size_t)PSF_PR * 8 + 64;
constexpr uint32_t PSF_SENT = 0xFFFFFFFFu;        // "no row-slot"
constexpr int32_t PSF_PARTIAL = (int32_t)0x80000000;  // chunk_row flag: one of several chunks of its row

struct PsfCfg {
    int logw, nslabs;
    int64_t PN, TH, CH, DL, DH, wfull;
};

struct PsfPlan {
    int x_kind = 0;
    PsfCfg cfg{};
    int npanels = 0, n_hpanels = 0;
    int64_t nchunks = 0, n_hchunks = 0;
    int n_split = 0;
    uint32_t *ent_idx = nullptr;   // [nnz+4] (row-slot << logw) | column-in-slab, sorted by (panel, slab)
    void *ent_val = nullptr;       // [nnz+4] values in the matrix's dtype (absent for structure-only)
    int64_t *woff = nullptr;       // [ncells*NW+1] start of every (cell, warp) sub-range
    int32_t *panel_first = nullptr;  // [npanels+1] first chunk of each panel
    int32_t *wchunk = nullptr;       // [npanels*(NW+1)] light panels: first row-slot of each warp
    int32_t *chunk_row = nullptr;    // [nchunks] row of each chunk (| PSF_PARTIAL)
    int32_t *split = nullptr;        // [3*n_split] (row, first chunk, chunk count) of multi-chunk rows
};

void psf_destroy(PsfPlan *p, cudaStream_t s)
{
    if (!p)
        return;
    dev_free(p->ent_idx, s);
    dev_free(p->ent_val, s);
    dev_free(p->woff, s);
    dev_free(p->panel_first, s);
    dev_free(p->wchunk, s);
    dev_free(p->chunk_row, s);
    dev_free(p->split, s);
    delete p;
}

// ------------------------------------------------------------------ builder
// per-row quantities scanned over the rows: 0 heavy chunk count, 1 light chunk count,
// 2 heavy panel weight, 3 light panel weight
template <typename RPT, int WHICH> struct PsfRowLoader {
    const RPT *rp;
    PsfCfg c;
    __device__ __forceinline__ int64_t operator()(int64_t r) const
    {
        const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
        const bool heavy = len > c.TH;
        const int64_t nch = (len + c.CH - 1) / c.CH;
        if (WHICH == 0) return heavy ? nch : 0;
        if (WHICH == 1) return heavy ? 0 : 1;
        if (WHICH == 2) return heavy ? len * PSF_RE + nch * c.PN : 0;
        return heavy ? 0 : len * PSF_PR + c.PN;
    }
};

template <typename RPT>
__global__ void k_psf_chunks(const RPT *__restrict__ rp, int32_t nrows, PsfCfg c, const int64_t *__restrict__ hcb,
                             const int64_t *__restrict__ lcb, const int64_t *__restrict__ hwc,
                             const int64_t *__restrict__ lwc, int64_t n_hc, int64_t n_lc, int n_hp,
                             int32_t *__restrict__ chunk_row, int32_t *__restrict__ chunk_panel,
                             int64_t *__restrict__ chunk_lnnz, int32_t *__restrict__ split, int *__restrict__ split_cnt)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0)
        chunk_lnnz[n_lc] = (lwc[nrows] - n_lc * c.PN) / PSF_PR;  // total light nnz
    if (r >= nrows)
        return;
    const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
    if (len > c.TH) {
        const int64_t nch = (len + c.CH - 1) / c.CH, cb = hcb[r];
        for (int64_t j = 0; j < nch; j++) {
            chunk_row[cb + j] = (int32_t)r | (nch > 1 ? PSF_PARTIAL : 0);
            chunk_panel[cb + j] = (int32_t)((hwc[r] + j * c.wfull) / c.DH);
        }
        if (nch > 1) {
            const int k = atomicAdd(split_cnt, 1);
            split[3 * k] = (int32_t)r;
            split[3 * k + 1] = (int32_t)cb;
            split[3 * k + 2] = (int32_t)nch;
        }
    } else {
        const int64_t lc = lcb[r];
        chunk_row[n_hc + lc] = (int32_t)r;
        chunk_panel[n_hc + lc] = n_hp + (int32_t)(lwc[r] / c.DL);
        chunk_lnnz[lc] = (lwc[r] - lc * c.PN) / PSF_PR;  // light nnz before this row
    }
}

__global__ void k_psf_panel_first(const int32_t *__restrict__ chunk_panel, int64_t nchunks, int npanels,
                                  int32_t *__restrict__ panel_first)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0)
        panel_first[npanels] = (int32_t)nchunks;
    if (c >= nchunks)
        return;
    if (c == 0 || chunk_panel[c] != chunk_panel[c - 1])
        panel_first[chunk_panel[c]] = (int32_t)c;
}

// light panels: warp w owns row-slots [wchunk[w], wchunk[w+1]) holding ~1/NW of the panel's nnz
__global__ void k_psf_wchunk(const int32_t *__restrict__ panel_first, const int64_t *__restrict__ chunk_lnnz, int npanels,
                             int n_hp, int64_t n_hc, int32_t *__restrict__ wchunk)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npanels * (PSF_NW + 1))
        return;
    const int p = i / (PSF_NW + 1), w = i % (PSF_NW + 1);
    if (p < n_hp) {
        wchunk[i] = 0;
        return;
    }
    const int64_t lo = panel_first[p] - n_hc, hi = panel_first[p + 1] - n_hc;
    const int64_t n0 = chunk_lnnz[lo], n1 = chunk_lnnz[hi];
    int64_t res;
    if (w == 0)
        res = lo;
    else if (w == PSF_NW)
        res = hi;
    else
        res = lower_bound_rp(chunk_lnnz, lo, hi, n0 + (n1 - n0) * w / PSF_NW);
    wchunk[i] = (int32_t)(res - lo);
}

template <typename RPT>
__global__ void k_psf_keys(const RPT *__restrict__ rp, const int32_t *__restrict__ ci, int32_t nrows, int64_t nnz, PsfCfg c,
                           const int64_t *__restrict__ hcb, const int64_t *__restrict__ lcb, int64_t n_hc, int n_hp,
                           const int32_t *__restrict__ chunk_panel, const int32_t *__restrict__ panel_first,
                           const int32_t *__restrict__ wchunk, int32_t *__restrict__ key, int32_t *__restrict__ packed,
                           uint32_t *__restrict__ wcnt)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz)
        return;
    const int64_t r = lower_bound_rp(rp, 0, (int64_t)nrows + 1, e + 1) - 1;
    const int64_t rs = (int64_t)rp[r], len = (int64_t)rp[r + 1] - rs;
    const int64_t ch = len > c.TH ? hcb[r] + (e - rs) / c.CH : n_hc + lcb[r];
    const int p = chunk_panel[ch];
    const int cl = (int)(ch - panel_first[p]);
    const int col = ci[e];
    const int cell = p * c.nslabs + (col >> c.logw);
    key[e] = cell;
    packed[e] = (int32_t)(((uint32_t)cl << c.logw) | ((uint32_t)col & ((1u << c.logw) - 1)));
    int w = 0;
    if (p >= n_hp) {
        const int32_t *wc = wchunk + (size_t)p * (PSF_NW + 1);
        int lo = 0, hi = PSF_NW;  // last w with wc[w] <= cl
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (wc[mid] <= cl)
                lo = mid;
            else
                hi = mid;
        }
        w = lo;
    }
    atomicAdd(&wcnt[(size_t)cell * PSF_NW + w], 1u);
}

// heavy panels: cut every cell evenly (at multiples of 4 entries) over the warps
__global__ void k_psf_even(int64_t *__restrict__ woff, int n_hcells)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hcells * (PSF_NW - 1))
        return;
    const int cell = i / (PSF_NW - 1), w = i % (PSF_NW - 1) + 1;
    const int64_t cb = woff[(size_t)cell * PSF_NW], ce = woff[(size_t)(cell + 1) * PSF_NW];
    int64_t v = (cb + (ce - cb) * w / PSF_NW) & ~(int64_t)3;
    woff[(size_t)cell * PSF_NW + w] = v < cb ? cb : v;
}

__global__ void k_psf_fixup(const int32_t *__restrict__ split, int n_split, const double *__restrict__ chunk_sums,
                            double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_split)
        return;
    const int row = split[3 * i], cb = split[3 * i + 1], n = split[3 * i + 2];
    double s = 0.0;
    for (int j = 0; j < n; j++)
        s += chunk_sums[cb + j];
    y[row] = s;
}

static int bits_for(int64_t n)
{
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) < n)
        b++;
    return b;
}

template <typename RPT, typename VT>
static int psf_build_typed(csrk_matrix *h, PsfPlan *P, cudaStream_t s)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    const PsfCfg c = P->cfg;
    const int32_t nrows = h->nrows;
    const int64_t nnz = h->nnz;
    const RPT *rp = (const RPT *)h->rp;
    // ---- rows -> chunks -> panels
    DevBuf hcb, lcb, hwc, lwc;
    const size_t rb = sizeof(int64_t) * ((size_t)nrows + 1);
    CSRK_TRY(hcb.alloc(rb, s));
    CSRK_TRY(lcb.alloc(rb, s));
    CSRK_TRY(hwc.alloc(rb, s));
    CSRK_TRY(lwc.alloc(rb, s));
    CSRK_TRY((exclusive_scan<int64_t>(PsfRowLoader<RPT, 0>{rp, c}, (int64_t)nrows, hcb.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(PsfRowLoader<RPT, 1>{rp, c}, (int64_t)nrows, lcb.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(PsfRowLoader<RPT, 2>{rp, c}, (int64_t)nrows, hwc.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(PsfRowLoader<RPT, 3>{rp, c}, (int64_t)nrows, lwc.as<int64_t>(), s)));
    int64_t tot[4];
    CSRK_CUDA(cudaMemcpyAsync(&tot[0], hcb.as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&tot[1], lcb.as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&tot[2], hwc.as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&tot[3], lwc.as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    const int64_t n_hc = tot[0], n_lc = tot[1];
    const int n_hp = n_hc ? (int)(tot[2] / c.DH) + 1 : 0;
    const int n_lp = n_lc ? (int)(tot[3] / c.DL) + 1 : 0;
    const int npanels = n_hp + n_lp;
    const int64_t nchunks = n_hc + n_lc;
    if (nchunks >= INT32_MAX || (int64_t)npanels * c.nslabs >= (1 << 30)) {
        set_error("matrix too large for the slab SpMV plan");
        return CSRK_EOVERFLOW;
    }
    P->npanels = npanels;
    P->n_hpanels = n_hp;
    P->nchunks = nchunks;
    P->n_hchunks = n_hc;
    const int64_t max_split = nnz / c.CH + 1;
    DevBuf chunk_panel, chunk_lnnz, split_cnt;
    CSRK_TRY(dev_alloc((void **)&P->chunk_row, sizeof(int32_t) * (size_t)nchunks, s));
    CSRK_TRY(dev_alloc((void **)&P->split, sizeof(int32_t) * 3 * (size_t)max_split, s));
    CSRK_TRY(chunk_panel.alloc(sizeof(int32_t) * (size_t)nchunks, s));
    CSRK_TRY(chunk_lnnz.alloc(sizeof(int64_t) * ((size_t)n_lc + 1), s));
    CSRK_TRY(split_cnt.alloc_zero(sizeof(int), s));
    CSRK_LAUNCH((k_psf_chunks<RPT>), (unsigned)div_up((int64_t)nrows, 256), 256, 0, s, rp, nrows, c, hcb.as<int64_t>(),
                lcb.as<int64_t>(), hwc.as<int64_t>(), lwc.as<int64_t>(), n_hc, n_lc, n_hp, P->chunk_row,
                chunk_panel.as<int32_t>(), chunk_lnnz.as<int64_t>(), P->split, split_cnt.as<int>());
    CSRK_TRY(dev_alloc((void **)&P->panel_first, sizeof(int32_t) * ((size_t)npanels + 1), s));
    CSRK_CUDA(cudaMemsetAsync(P->panel_first, 0xFF, sizeof(int32_t) * ((size_t)npanels + 1), s));
    CSRK_LAUNCH(k_psf_panel_first, (unsigned)div_up(nchunks, 256), 256, 0, s, chunk_panel.as<int32_t>(), nchunks, npanels,
                P->panel_first);
    {
        // panels that received no chunk (the estimate of the panel count can overshoot by one)
        // take the start of their successor, i.e. they are empty
        std::vector<int32_t> pf((size_t)npanels + 1);
        CSRK_CUDA(cudaMemcpyAsync(pf.data(), P->panel_first, sizeof(int32_t) * pf.size(), cudaMemcpyDeviceToHost, s));
        CSRK_CUDA(cudaMemcpyAsync(&P->n_split, split_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CSRK_CUDA(cudaStreamSynchronize(s));
        for (int p = npanels - 1; p >= 0; p--)
            if (pf[p] < 0)
                pf[p] = pf[p + 1];
        CSRK_CUDA(cudaMemcpyAsync(P->panel_first, pf.data(), sizeof(int32_t) * pf.size(), cudaMemcpyHostToDevice, s));
        CSRK_CUDA(cudaStreamSynchronize(s));
    }
    CSRK_TRY(dev_alloc((void **)&P->wchunk, sizeof(int32_t) * (size_t)npanels * (PSF_NW + 1), s));
    CSRK_LAUNCH(k_psf_wchunk, (unsigned)div_up((int64_t)npanels * (PSF_NW + 1), 256), 256, 0, s, P->panel_first,
                chunk_lnnz.as<int64_t>(), npanels, n_hp, n_hc, P->wchunk);

    // ---- entries -> (cell key, packed index); counts per (cell, warp)
    const int64_t ncells = (int64_t)npanels * c.nslabs;
    DevBuf key, packed, wcnt;
    CSRK_TRY(key.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(packed.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(wcnt.alloc_zero(sizeof(uint32_t) * (size_t)ncells * PSF_NW, s));
    CSRK_LAUNCH((k_psf_keys<RPT>), (unsigned)div_up(nnz, 256), 256, 0, s, rp, h->ci, nrows, nnz, c, hcb.as<int64_t>(),
                lcb.as<int64_t>(), n_hc, n_hp, chunk_panel.as<int32_t>(), P->panel_first, P->wchunk, key.as<int32_t>(),
                packed.as<int32_t>(), wcnt.as<uint32_t>());
    hcb.reset();
    lcb.reset();
    hwc.reset();
    lwc.reset();
    // ---- stable sort by cell; payloads land in the plan (4 spare entries: aligned 128-bit reads)
    CSRK_TRY(dev_alloc((void **)&P->ent_idx, sizeof(uint32_t) * ((size_t)nnz + 4), s));
    CSRK_CUDA(cudaMemsetAsync(P->ent_idx + nnz, 0, sizeof(uint32_t) * 4, s));
    if (HASV) {
        CSRK_TRY(dev_alloc(&P->ent_val, sizeof(VT) * ((size_t)nnz + 4), s));
        CSRK_CUDA(cudaMemsetAsync((char *)P->ent_val + sizeof(VT) * (size_t)nnz, 0, sizeof(VT) * 4, s));
    }
    CSRK_TRY((radix_sort_by_key<VT>(key.as<int32_t>(), packed.as<int32_t>(), (const VT *)h->vs, nnz, bits_for(ncells),
                                    (int32_t *)P->ent_idx, (VT *)P->ent_val, s)));
    // ---- sub-range offsets
    CSRK_TRY(dev_alloc((void **)&P->woff, sizeof(int64_t) * ((size_t)ncells * PSF_NW + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<uint32_t>{wcnt.as<uint32_t>()}, ncells * PSF_NW, P->woff, s)));
    const int64_t n_hcells = (int64_t)n_hp * c.nslabs;
    if (n_hcells)
        CSRK_LAUNCH(k_psf_even, (unsigned)div_up(n_hcells * (PSF_NW - 1), 256), 256, 0, s, P->woff, (int)n_hcells);
    CSRK_CUDA(cudaStreamSynchronize(s));
    return CSRK_OK;
}

// ------------------------------------------------------------------ kernel
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 ld_stream_uint4(const void *ptr)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr));
    return r;
}

struct PsfArgs {
    const uint32_t *ent_idx;
    const void *ent_val;
    const int64_t *woff;
    const int32_t *panel_first, *wchunk, *chunk_row;
    int npanels, n_hpanels, nslabs, logw;
    int32_t ncols;
};

template <typename VT> struct PsfVal4 {
    __device__ static __forceinline__ void load(const void *vs, int64_t e, double (&v)[4]);
};
template <> __device__ __forceinline__ void PsfVal4<float>::load(const void *vs, int64_t e, double (&v)[4])
{
    const float4 f = ld_stream_float4((const float *)vs + e);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
template <> __device__ __forceinline__ void PsfVal4<double>::load(const void *vs, int64_t e, double (&v)[4])
{
    const double2 a = ld_stream_double2((const double *)vs + e), b = ld_stream_double2((const double *)vs + e + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <> __device__ __forceinline__ void PsfVal4<NoPayload>::load(const void *, int64_t, double (&v)[4])
{
    v[0] = v[1] = v[2] = v[3] = 1.0;
}

// numba's promotion: f4 * f4 is rounded to f4 before the float64 accumulation
template <typename VT, typename XT> __device__ __forceinline__ double psf_prod(XT xv, double v)
{
    if (std::is_same<VT, float>::value && std::is_same<XT, float>::value)
        return (double)((float)xv * (float)v);
    return (double)xv * v;
}

template <typename VT, typename XT>
__global__ void __launch_bounds__(PSF_THREADS, 1)
k_psf_spmv(PsfArgs a, const XT *__restrict__ x, double *__restrict__ y, double *__restrict__ chunk_sums,
           int *__restrict__ counter)
{
    extern __shared__ __align__(128) unsigned char psf_smem[];
    XT *xbuf = reinterpret_cast<XT *>(psf_smem);
    double *yacc = reinterpret_cast<double *>(psf_smem + 2 * PSF_SLAB_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(psf_smem + 2 * PSF_SLAB_BYTES + PSF_PR * 8);
    uint64_t *empty = full + 2;
    __shared__ int s_panel;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = 1 << a.logw;
    const uint32_t cmask = (uint32_t)W - 1u;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], PSF_NW);
        mbar_init(&empty[1], PSF_NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t it = 0;  // slab iterations this CTA has gone through (ring position)
    while (true) {
        __syncthreads();
        if (tid == 0)
            s_panel = atomicAdd(counter, 1);
        __syncthreads();
        const int p = s_panel;
        if (p >= a.npanels)
            break;
        const int c0 = a.panel_first[p], nc = a.panel_first[p + 1] - c0;
        const bool even = p < a.n_hpanels;
        if (warp == PSF_NW) {
            // ---------------- producer: stream x slab by slab into the ring
            if (lane == 0) {
                for (int s = 0; s < a.nslabs; s++) {
                    const uint32_t i = it + s;
                    const int buf = i & 1;
                    mbar_wait(&empty[buf], ((i >> 1) & 1) ^ 1);
                    const int64_t col0 = (int64_t)s << a.logw;
                    const int ncol = (int)min((int64_t)W, (int64_t)a.ncols - col0);
                    const uint32_t bytes = (uint32_t)ncol * sizeof(XT), bulk = bytes & ~15u;
                    XT *dst = xbuf + (size_t)buf * W;
                    for (int k = bulk / sizeof(XT); k < ncol; k++)
                        dst[k] = x[col0 + k];  // <16 B tail that a bulk copy cannot carry
                    mbar_expect_tx(&full[buf], bulk);
                    if (bulk)
                        bulk_g2s(dst, x + col0, bulk, &full[buf]);
                }
            }
        } else {
            // ---------------- consumers
            const int accbase = even ? warp * nc : 0;
            if (even) {
                for (int i = lane; i < nc; i += 32)
                    yacc[accbase + i] = 0.0;
            } else {
                const int w0 = a.wchunk[(size_t)p * (PSF_NW + 1) + warp], w1 = a.wchunk[(size_t)p * (PSF_NW + 1) + warp + 1];
                for (int i = w0 + lane; i < w1; i += 32)
                    yacc[i] = 0.0;
            }
            __syncwarp();
            const int64_t *wo = a.woff + ((size_t)p * a.nslabs) * PSF_NW + warp;
            int64_t b = wo[0], e = wo[1];
            for (int s = 0; s < a.nslabs; s++) {
                int64_t bn = 0, en = 0;
                if (s + 1 < a.nslabs) {
                    bn = wo[(size_t)(s + 1) * PSF_NW];
                    en = wo[(size_t)(s + 1) * PSF_NW + 1];
                }
                const uint32_t i = it + s;
                const int buf = i & 1;
                const int64_t g0 = b >> 2, g1 = (e + 3) >> 2;
                // first block's loads are issued before waiting for the slab
                int64_t g = g0 + lane;
                bool lv = g < g1 && b < e;
                uint4 pk = make_uint4(0, 0, 0, 0);
                double v[4] = {0.0, 0.0, 0.0, 0.0};
                if (lv) {
                    pk = ld_stream_uint4(a.ent_idx + 4 * g);
                    PsfVal4<VT>::load(a.ent_val, 4 * g, v);
                }
                mbar_wait(&full[buf], (i >> 1) & 1);
                const XT *xs = xbuf + (size_t)buf * W;
                for (int64_t gb = g0; gb < g1 && b < e; gb += 32) {
                    // prefetch the next block of this cell
                    const int64_t gn = gb + 32 + lane;
                    const bool lvn = gn < g1;
                    uint4 pkn = make_uint4(0, 0, 0, 0);
                    double vn[4] = {0.0, 0.0, 0.0, 0.0};
                    if (lvn) {
                        pkn = ld_stream_uint4(a.ent_idx + 4 * gn);
                        PsfVal4<VT>::load(a.ent_val, 4 * gn, vn);
                    }
                    // ---- products; entries outside [b, e) contribute +0 and borrow a neighbour's slot
                    const int64_t e0 = 4 * g;
                    const bool m0 = lv && e0 >= b && e0 < e, m1 = lv && e0 + 1 >= b && e0 + 1 < e;
                    const bool m2 = lv && e0 + 2 >= b && e0 + 2 < e, m3 = lv && e0 + 3 >= b && e0 + 3 < e;
                    const double p0 = m0 ? psf_prod<VT, XT>(xs[pk.x & cmask], v[0]) : 0.0;
                    const double p1 = m1 ? psf_prod<VT, XT>(xs[pk.y & cmask], v[1]) : 0.0;
                    const double p2 = m2 ? psf_prod<VT, XT>(xs[pk.z & cmask], v[2]) : 0.0;
                    const double p3 = m3 ? psf_prod<VT, XT>(xs[pk.w & cmask], v[3]) : 0.0;
                    uint32_t k0 = m0 ? pk.x >> a.logw : PSF_SENT;
                    uint32_t k1 = m1 ? pk.y >> a.logw : k0;
                    uint32_t k2 = m2 ? pk.z >> a.logw : k1;
                    uint32_t k3 = m3 ? pk.w >> a.logw : k2;
                    k2 = k2 == PSF_SENT ? k3 : k2;
                    k1 = k1 == PSF_SENT ? k2 : k1;
                    k0 = k0 == PSF_SENT ? k1 : k0;
                    // ---- runs inside the lane
                    const bool b1 = k1 != k0, b2 = k2 != k1, b3 = k3 != k2;
                    const bool uni = !(b1 | b2 | b3);
                    const double H = p0 + (b1 ? 0.0 : p1 + (b2 ? 0.0 : p2 + (b3 ? 0.0 : p3)));
                    const double T = p3 + (b3 ? 0.0 : p2 + (b2 ? 0.0 : p1 + (b1 ? 0.0 : p0)));
                    const bool i1 = b1 && (b2 || b3), i2 = b2 && b3;
                    const double I1 = p1 + (b2 ? 0.0 : p2);
                    // ---- runs across lanes: segmented inclusive scan of the tail sums
                    uint32_t pk3 = __shfl_up_sync(0xffffffffu, k3, 1);
                    uint32_t nk0 = __shfl_down_sync(0xffffffffu, k0, 1);
                    if (lane == 0) pk3 = PSF_SENT - 1;
                    if (lane == 31) nk0 = PSF_SENT - 1;
                    const bool cin = k0 == pk3 && k0 != PSF_SENT;
                    const bool cout = k3 == nk0 && k3 != PSF_SENT;
                    const bool head = !(uni && cin);  // this lane's tail run starts here
                    const unsigned hm = __ballot_sync(0xffffffffu, head);
                    const int start = 31 - __clz(hm & (0xffffffffu >> (31 - lane)));  // lane 0 is always a head
                    const int span = lane - start;
                    const int maxspan = __reduce_max_sync(0xffffffffu, span);
                    double S = T;
                    for (int d = 1; d <= maxspan; d <<= 1) {
                        const double t = __shfl_up_sync(0xffffffffu, S, d);
                        if (span >= d)
                            S += t;
                    }
                    const double Sprev = __shfl_up_sync(0xffffffffu, S, 1);
                    // ---- flush: every run is written by exactly one lane; slots within one
                    // statement are distinct, and a slot belongs to this warp only
                    double *acc = yacc + accbase;
                    if (!uni && k0 != PSF_SENT)
                        acc[k0] += H + (cin ? Sprev : 0.0);
                    if (i1)
                        acc[k1] += I1;
                    if (i2)
                        acc[k2] += p2;
                    if (!cout && k3 != PSF_SENT)
                        acc[k3] += S;
                    __syncwarp();
                    pk = pkn;
                    v[0] = vn[0]; v[1] = vn[1]; v[2] = vn[2]; v[3] = vn[3];
                    lv = lvn;
                    g = gn;
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&empty[buf]);
                b = bn;
                e = en;
            }
            if (!even) {
                // light panel: this warp owns its rows outright
                const int w0 = a.wchunk[(size_t)p * (PSF_NW + 1) + warp], w1 = a.wchunk[(size_t)p * (PSF_NW + 1) + warp + 1];
                for (int i = w0 + lane; i < w1; i += 32)
                    y[a.chunk_row[c0 + i]] = yacc[i];
            }
        }
        it += a.nslabs;
        __syncthreads();
        if (even) {
            // heavy panel: sum the per-warp copies in warp order
            for (int i = tid; i < nc; i += PSF_THREADS) {
                double sum = 0.0;
#pragma unroll 4
                for (int w = 0; w < PSF_NW; w++)
                    sum += yacc[w * nc + i];
                const int32_t row = a.chunk_row[c0 + i];
                if (row < 0)
                    chunk_sums[c0 + i] = sum;
                else
                    y[row] = sum;
            }
        }
    }
}

template <typename VT, typename XT>
static int psf_launch(csrk_matrix *h, PsfPlan *P, const void *d_x, double *d_y, cudaStream_t s)
{
    auto k = k_psf_spmv<VT, XT>;
    static bool optin = false;  // per instantiation
    if (!optin) {
        CSRK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PSF_SMEM));
        optin = true;
    }
    DevBuf counter, sums;
    CSRK_TRY(counter.alloc_zero(sizeof(int), s));
    CSRK_TRY(sums.alloc(sizeof(double) * (size_t)std::max<int64_t>(P->n_hchunks, 1), s));
    PsfArgs a{P->ent_idx, P->ent_val, P->woff, P->panel_first, P->wchunk, P->chunk_row,
              P->npanels, P->n_hpanels, P->cfg.nslabs, P->cfg.logw, h->ncols};
    const int grid = std::min(P->npanels, ctx().sm_count);
    CSRK_LAUNCH(k, (unsigned)grid, PSF_THREADS, PSF_SMEM, s, a, (const XT *)d_x, d_y, sums.as<double>(), counter.as<int>());
    if (P->n_split)
        CSRK_LAUNCH(k_psf_fixup, (unsigned)div_up(P->n_split, 128), 128, 0, s, P->split, P->n_split, sums.as<double>(), d_y);
    return CSRK_OK;
}

// ------------------------------------------------------------------ entry points used by spmv.cu
int psf_build(csrk_matrix *h, int x_kind, PsfPlan **out, cudaStream_t s)
{
    *out = nullptr;
    PsfPlan *P = new (std::nothrow) PsfPlan();
    if (!P) {
        set_error("host allocation failed");
        return CSRK_ENOMEM;
    }
    WsScope scope;  // build temporaries come from the workspace arena
    P->x_kind = x_kind;
    PsfCfg &c = P->cfg;
    c.logw = x_kind == 4 ? 14 : 13;
    c.nslabs = (int)div_up((int64_t)h->ncols, (int64_t)1 << c.logw);
    const int sms = std::max(ctx().sm_count, 1);
    c.PN = std::min<int64_t>(std::max<int64_t>(h->nnz / (2 * sms), 32768), 1 << 20);
    c.TH = std::max<int64_t>(256, c.PN / (2 * PSF_RE));
    c.CH = std::max<int64_t>((c.PN / 2) & ~(int64_t)3, c.TH + 1);
    c.DL = c.PN * PSF_PR;
    c.DH = c.PN * PSF_RE;
    c.wfull = c.CH * PSF_RE + c.PN;
    int rc;
    if (h->rp_is64) {
        if (h->val_kind == 4) rc = psf_build_typed<int64_t, float>(h, P, s);
        else if (h->val_kind == 8) rc = psf_build_typed<int64_t, double>(h, P, s);
        else rc = psf_build_typed<int64_t, NoPayload>(h, P, s);
    } else {
        if (h->val_kind == 4) rc = psf_build_typed<int32_t, float>(h, P, s);
        else if (h->val_kind == 8) rc = psf_build_typed<int32_t, double>(h, P, s);
        else rc = psf_build_typed<int32_t, NoPayload>(h, P, s);
    }
    if (rc != CSRK_OK) {
        psf_destroy(P, s);
        return rc;
    }
    *out = P;
    return CSRK_OK;
}

int psf_run(csrk_matrix *h, PsfPlan *P, const void *d_x, double *d_y, cudaStream_t s)
{
    if (P->x_kind == 4) {
        if (h->val_kind == 4) return psf_launch<float, float>(h, P, d_x, d_y, s);
        if (h->val_kind == 8) return psf_launch<double, float>(h, P, d_x, d_y, s);
        return psf_launch<NoPayload, float>(h, P, d_x, d_y, s);
    }
    if (h->val_kind == 4) return psf_launch<float, double>(h, P, d_x, d_y, s);
    if (h->val_kind == 8) return psf_launch<double, double>(h, P, d_x, d_y, s);
    return psf_launch<NoPayload, double>(h, P, d_x, d_y, s);
}

int psf_panels(const PsfPlan *P) { return P->npanels; }

}  // namespace csrk
