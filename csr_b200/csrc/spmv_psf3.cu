// spmv_psf3.cu -- "cell-tile" SpMV: x in shared memory, entries streamed by TMA.  spmv_mode = 3.
//
// Third design of the slab SpMV (see spmv_psf.cu for the idea and DESIGN.md 4.2 for the history).
// It keeps the CHEAP structure of the CSR tile kernel -- products parked in shared memory, then
// every row segment summed by one owning thread (or warp) -- and removes what made v1 slow:
//
//   * rows are grouped into PANELS (<= 7168 row-chunks, ~nnz/#SM entries); a panel's entries are
//     sorted by column SLAB (32 KB of x); one (panel, slab) cell is cut into WORK ITEMS of at most
//     NMAX entries.  A work item is three small arrays, contiguous in HBM:
//         col16[n]   column inside the slab            (2 B / entry)
//         val[n]     matrix value                      (V B / entry)
//         seg[m+1]   one descriptor per row segment: row-slot | start << 13   (4 B / segment)
//     i.e. 6-7 B per nnz for f32 instead of CSR's 8: no per-entry row information at all;
//   * a producer warp streams the work items through a 3-stage shared-memory ring and the x slabs
//     through a 2-slab ring with cp.async.bulk (TMA) + mbarriers, up to three items ahead of the
//     consumers, so HBM latency is never exposed and no registers are spent on prefetching;
//   * 16 consumer warps per CTA: phase 1 multiplies in place (x gathered from shared memory),
//     phase 2 gives every segment to one thread (short) or one warp (long) which adds its sum into
//     the panel's float64 accumulators in shared memory -- plain read-modify-write, no atomics,
//     no shuffles on the common path; the order of additions is fixed => deterministic.
// Supported value/x types: (f32, f32), (f64, f32), (f64, f64): the product overwrites the value
// slot.  Other combinations use the CSR tile kernel.
#include <algorithm>
#include <vector>

#include "radix.cuh"

namespace csrk {

constexpr int P3_NC = 16;                         // consumer warps
constexpr int P3_CTHREADS = P3_NC * 32;
constexpr int P3_THREADS = P3_CTHREADS + 32;      // + producer warp
constexpr int P3_STAGES = 3;
constexpr int P3_STAGE_BYTES = 32 * 1024;
constexpr int P3_SLAB_BYTES = 32 * 1024;
constexpr int P3_PR = 7168;                       // float64 accumulator slots (row-chunks per panel)
constexpr int P3_LONG = 64;                       // segments longer than this go to a warp
constexpr int P3_QCAP = 64;                       // long segments per work item (NMAX / LONG rounded up)
constexpr size_t P3_SMEM = 2 * (size_t)P3_SLAB_BYTES + P3_STAGES * (size_t)P3_STAGE_BYTES + (size_t)P3_PR * 8 + 256;
constexpr int32_t P3_PARTIAL = (int32_t)0x80000000;

struct P3Item {          // 16 bytes
    uint32_t e8;         // entry offset / 8   (col16 and val arrays)
    uint32_t g4;         // descriptor offset / 4
    uint16_t n, m;       // entries, segments
    uint16_t slab;
    uint16_t flags;      // 1: first item of its slab in the panel, 2: last
};

struct P3Cfg {
    int logw, nslabs, nmax;
    int64_t PN, CH, D, wfull;
};

struct Psf3Plan {
    int x_kind = 0, val_kind = 0;
    P3Cfg cfg{};
    int npanels = 0;
    int64_t nchunks = 0, nitems = 0;
    int n_split = 0;
    uint16_t *col16 = nullptr;
    void *val = nullptr;
    uint32_t *seg = nullptr;
    P3Item *items = nullptr;
    int32_t *item_first = nullptr;   // [npanels+1]
    int32_t *panel_first = nullptr;  // [npanels+1] first chunk of each panel
    int32_t *chunk_row = nullptr;    // [nchunks]
    int32_t *split = nullptr;        // [3*n_split]
};

void psf3_destroy(Psf3Plan *p, cudaStream_t s)
{
    if (!p)
        return;
    dev_free(p->col16, s);
    dev_free(p->val, s);
    dev_free(p->seg, s);
    dev_free(p->items, s);
    dev_free(p->item_first, s);
    dev_free(p->panel_first, s);
    dev_free(p->chunk_row, s);
    dev_free(p->split, s);
    delete p;
}

// ------------------------------------------------------------------ builder
// per-row scans: 0 = chunk count, 1 = panel weight (max(len*PR, PN) per chunk)
template <typename RPT, int WHICH> struct P3RowLoader {
    const RPT *rp;
    P3Cfg c;
    __device__ __forceinline__ int64_t operator()(int64_t r) const
    {
        const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
        const int64_t nch = len > c.CH ? (len + c.CH - 1) / c.CH : 1;
        if (WHICH == 0)
            return nch;
        const int64_t last = len - (nch - 1) * c.CH;
        const int64_t wl = last * P3_PR > c.PN ? last * P3_PR : c.PN;
        return (nch - 1) * c.wfull + wl;
    }
};

template <typename RPT>
__global__ void k3_chunks(const RPT *__restrict__ rp, int32_t nrows, P3Cfg c, const int64_t *__restrict__ cb,
                          const int64_t *__restrict__ wc, int32_t *__restrict__ chunk_row, int32_t *__restrict__ chunk_panel,
                          int32_t *__restrict__ split, int *__restrict__ split_cnt)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows)
        return;
    const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
    const int64_t nch = len > c.CH ? (len + c.CH - 1) / c.CH : 1, base = cb[r];
    for (int64_t j = 0; j < nch; j++) {
        chunk_row[base + j] = (int32_t)r | (nch > 1 ? P3_PARTIAL : 0);
        chunk_panel[base + j] = (int32_t)((wc[r] + j * c.wfull) / c.D);
    }
    if (nch > 1) {
        const int k = atomicAdd(split_cnt, 1);
        split[3 * k] = (int32_t)r;
        split[3 * k + 1] = (int32_t)base;
        split[3 * k + 2] = (int32_t)nch;
    }
}

__global__ void k3_panel_first(const int32_t *__restrict__ chunk_panel, int64_t nchunks, int npanels,
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

// key = panel*nslabs + slab;  packed = row-slot << 13 | column inside the slab
template <typename RPT>
__global__ void k3_keys(const RPT *__restrict__ rp, const int32_t *__restrict__ ci, int32_t nrows, int64_t nnz, P3Cfg c,
                        const int64_t *__restrict__ cb, const int32_t *__restrict__ chunk_panel,
                        const int32_t *__restrict__ panel_first, int32_t *__restrict__ key, int32_t *__restrict__ packed,
                        uint32_t *__restrict__ cell_cnt)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz)
        return;
    const int64_t r = lower_bound_rp(rp, 0, (int64_t)nrows + 1, e + 1) - 1;
    const int64_t rs = (int64_t)rp[r], len = (int64_t)rp[r + 1] - rs;
    const int64_t ch = cb[r] + (len > c.CH ? (e - rs) / c.CH : 0);
    const int p = chunk_panel[ch];
    const int cl = (int)(ch - panel_first[p]);
    const int col = ci[e];
    const int k = p * c.nslabs + (col >> c.logw);
    key[e] = k;
    packed[e] = (int32_t)(((uint32_t)cl << 13) | ((uint32_t)col & ((1u << c.logw) - 1)));
    atomicAdd(&cell_cnt[k], 1u);
}

// a segment starts where the row-slot changes, at the start of a cell and at every NMAX-th entry of a cell
struct P3FlagLoader {
    const int32_t *skey, *spacked;
    const int64_t *cell_off;
    int nmax;
    __device__ __forceinline__ int64_t operator()(int64_t i) const
    {
        const int64_t li = i - cell_off[skey[i]];
        if (li % nmax == 0)
            return 1;
        return (spacked[i] >> 13) != (spacked[i - 1] >> 13) ? 1 : 0;
    }
};

struct P3ItemCount {
    const uint32_t *cell_cnt;
    int nmax;
    __device__ __forceinline__ int64_t operator()(int64_t c) const { return (cell_cnt[c] + nmax - 1) / nmax; }
};

// one thread per work item: its extent, its padded sizes
__global__ void k3_items_a(const int64_t *__restrict__ item_base, int64_t ncells, int64_t nitems, const int64_t *__restrict__ cell_off,
                           const int64_t *__restrict__ fpre, int nmax, int nslabs, int64_t *__restrict__ it_i0,
                           int32_t *__restrict__ it_n, int32_t *__restrict__ it_m, int32_t *__restrict__ it_meta,
                           int32_t *__restrict__ epad, int32_t *__restrict__ gpad)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nitems)
        return;
    // cell = last c with item_base[c] <= w
    int64_t lo = 0, hi = ncells;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (item_base[mid] <= w)
            lo = mid;
        else
            hi = mid;
    }
    const int64_t cell = lo, j = w - item_base[cell], cnt = item_base[cell + 1] - item_base[cell];
    const int64_t i0 = cell_off[cell] + j * nmax;
    const int n = (int)min((int64_t)nmax, cell_off[cell + 1] - i0);
    const int m = (int)(fpre[i0 + n] - fpre[i0]);
    it_i0[w] = i0;
    it_n[w] = n;
    it_m[w] = m;
    it_meta[w] = (int)(cell % nslabs) | ((j == 0 ? 1 : 0) << 16) | ((j == cnt - 1 ? 2 : 0) << 16);
    epad[w] = (n + 7) & ~7;
    gpad[w] = (m + 1 + 3) & ~3;
}

__global__ void k3_items_b(int64_t nitems, const int64_t *__restrict__ e_off, const int64_t *__restrict__ g_off,
                           const int32_t *__restrict__ it_n, const int32_t *__restrict__ it_m,
                           const int32_t *__restrict__ it_meta, P3Item *__restrict__ items, uint32_t *__restrict__ seg)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nitems)
        return;
    P3Item it;
    it.e8 = (uint32_t)(e_off[w] >> 3);
    it.g4 = (uint32_t)(g_off[w] >> 2);
    it.n = (uint16_t)it_n[w];
    it.m = (uint16_t)it_m[w];
    it.slab = (uint16_t)(it_meta[w] & 0xFFFF);
    it.flags = (uint16_t)(it_meta[w] >> 16);
    items[w] = it;
    seg[g_off[w] + it_m[w]] = (uint32_t)it_n[w] << 13;  // sentinel: end of the last segment
}

template <typename VT>
__global__ void k3_place(const int32_t *__restrict__ skey, const int32_t *__restrict__ spacked, const VT *__restrict__ sval,
                         int64_t nnz, const int64_t *__restrict__ cell_off, const int64_t *__restrict__ item_base,
                         const int64_t *__restrict__ fpre, const int64_t *__restrict__ e_off, const int64_t *__restrict__ g_off,
                         int nmax, uint16_t *__restrict__ col16, VT *__restrict__ val, uint32_t *__restrict__ seg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz)
        return;
    const int k = skey[i];
    const int64_t li = i - cell_off[k];
    const int64_t w = item_base[k] + li / nmax;
    const int pos = (int)(li % nmax);
    const uint32_t pk = (uint32_t)spacked[i];
    col16[e_off[w] + pos] = (uint16_t)(pk & 0x1FFF);
    val[e_off[w] + pos] = sval[i];
    const bool flag = pos == 0 || (pk >> 13) != ((uint32_t)spacked[i - 1] >> 13);
    if (flag) {
        const int64_t rank = fpre[i] - fpre[i - pos];  // segments of this item before this one
        seg[g_off[w] + rank] = (pk >> 13) | ((uint32_t)pos << 13);
    }
}

__global__ void k3_item_first(const int64_t *__restrict__ item_base, int npanels, int nslabs, int32_t *__restrict__ item_first)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p <= npanels)
        item_first[p] = (int32_t)item_base[(int64_t)p * nslabs];
}

__global__ void k3_fixup(const int32_t *__restrict__ split, int n_split, const double *__restrict__ chunk_sums,
                         double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_split)
        return;
    const int row = split[3 * i], cbase = split[3 * i + 1], n = split[3 * i + 2];
    double s = 0.0;
    for (int j = 0; j < n; j++)
        s += chunk_sums[cbase + j];
    y[row] = s;
}

static int p3_bits_for(int64_t n)
{
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) < n)
        b++;
    return b;
}

template <typename RPT, typename VT>
static int psf3_build_typed(csrk_matrix *h, Psf3Plan *P, cudaStream_t s)
{
    const P3Cfg c = P->cfg;
    const int32_t nrows = h->nrows;
    const int64_t nnz = h->nnz;
    const RPT *rp = (const RPT *)h->rp;
    // ---- rows -> chunks -> panels
    DevBuf cb, wc;
    const size_t rb = sizeof(int64_t) * ((size_t)nrows + 1);
    CSRK_TRY(cb.alloc(rb, s));
    CSRK_TRY(wc.alloc(rb, s));
    CSRK_TRY((exclusive_scan<int64_t>(P3RowLoader<RPT, 0>{rp, c}, (int64_t)nrows, cb.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(P3RowLoader<RPT, 1>{rp, c}, (int64_t)nrows, wc.as<int64_t>(), s)));
    int64_t tot[2];
    CSRK_CUDA(cudaMemcpyAsync(&tot[0], cb.as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&tot[1], wc.as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    const int64_t nchunks = tot[0];
    const int npanels = (int)(tot[1] / c.D) + 1;
    const int64_t ncells = (int64_t)npanels * c.nslabs;
    if (nchunks >= INT32_MAX || ncells >= ((int64_t)1 << 30)) {
        set_error("matrix too large for the slab SpMV plan");
        return CSRK_EOVERFLOW;
    }
    P->npanels = npanels;
    P->nchunks = nchunks;
    DevBuf chunk_panel, split_cnt;
    CSRK_TRY(dev_alloc((void **)&P->chunk_row, sizeof(int32_t) * (size_t)nchunks, s));
    CSRK_TRY(dev_alloc((void **)&P->split, sizeof(int32_t) * 3 * (size_t)(nnz / c.CH + 1), s));
    CSRK_TRY(chunk_panel.alloc(sizeof(int32_t) * (size_t)nchunks, s));
    CSRK_TRY(split_cnt.alloc_zero(sizeof(int), s));
    CSRK_LAUNCH((k3_chunks<RPT>), (unsigned)div_up((int64_t)nrows, 256), 256, 0, s, rp, nrows, c, cb.as<int64_t>(),
                wc.as<int64_t>(), P->chunk_row, chunk_panel.as<int32_t>(), P->split, split_cnt.as<int>());
    CSRK_TRY(dev_alloc((void **)&P->panel_first, sizeof(int32_t) * ((size_t)npanels + 1), s));
    CSRK_CUDA(cudaMemsetAsync(P->panel_first, 0xFF, sizeof(int32_t) * ((size_t)npanels + 1), s));
    CSRK_LAUNCH(k3_panel_first, (unsigned)div_up(nchunks, 256), 256, 0, s, chunk_panel.as<int32_t>(), nchunks, npanels,
                P->panel_first);
    {
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
    // ---- entries: keys, cell counts, stable sort by (panel, slab)
    DevBuf key, packed, cell_cnt, skey, spacked, sval;
    CSRK_TRY(key.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(packed.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(cell_cnt.alloc_zero(sizeof(uint32_t) * ((size_t)ncells + 1), s));
    CSRK_LAUNCH((k3_keys<RPT>), (unsigned)div_up(nnz, 256), 256, 0, s, rp, h->ci, nrows, nnz, c, cb.as<int64_t>(),
                chunk_panel.as<int32_t>(), P->panel_first, key.as<int32_t>(), packed.as<int32_t>(), cell_cnt.as<uint32_t>());
    cb.reset();
    wc.reset();
    CSRK_TRY(skey.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(spacked.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(sval.alloc(sizeof(VT) * (size_t)nnz, s));
    const int kb = p3_bits_for(ncells);
    CSRK_TRY((radix_sort_by_key<VT>(key.as<int32_t>(), packed.as<int32_t>(), (const VT *)h->vs, nnz, kb,
                                    spacked.as<int32_t>(), sval.as<VT>(), s)));
    CSRK_TRY((radix_sort_by_key<NoPayload>(key.as<int32_t>(), key.as<int32_t>(), (const NoPayload *)nullptr, nnz, kb,
                                           skey.as<int32_t>(), (NoPayload *)nullptr, s)));
    key.reset();
    packed.reset();
    // ---- cells -> work items
    DevBuf cell_off, item_base, fpre;
    CSRK_TRY(cell_off.alloc(sizeof(int64_t) * ((size_t)ncells + 1), s));
    CSRK_TRY(item_base.alloc(sizeof(int64_t) * ((size_t)ncells + 1), s));
    CSRK_TRY(fpre.alloc(sizeof(int64_t) * ((size_t)nnz + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<uint32_t>{cell_cnt.as<uint32_t>()}, ncells, cell_off.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(P3ItemCount{cell_cnt.as<uint32_t>(), c.nmax}, ncells, item_base.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(P3FlagLoader{skey.as<int32_t>(), spacked.as<int32_t>(), cell_off.as<int64_t>(), c.nmax},
                                      nnz, fpre.as<int64_t>(), s)));
    int64_t nitems = 0;
    CSRK_CUDA(cudaMemcpyAsync(&nitems, item_base.as<int64_t>() + ncells, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    if (nitems >= INT32_MAX) {
        set_error("matrix too large for the slab SpMV plan");
        return CSRK_EOVERFLOW;
    }
    P->nitems = nitems;
    const size_t ni = (size_t)std::max<int64_t>(nitems, 1);
    DevBuf it_i0, it_n, it_m, it_meta, epad, gpad, e_off, g_off;
    CSRK_TRY(it_i0.alloc(sizeof(int64_t) * ni, s));
    CSRK_TRY(it_n.alloc(sizeof(int32_t) * ni, s));
    CSRK_TRY(it_m.alloc(sizeof(int32_t) * ni, s));
    CSRK_TRY(it_meta.alloc(sizeof(int32_t) * ni, s));
    CSRK_TRY(epad.alloc(sizeof(int32_t) * ni, s));
    CSRK_TRY(gpad.alloc(sizeof(int32_t) * ni, s));
    CSRK_TRY(e_off.alloc(sizeof(int64_t) * (ni + 1), s));
    CSRK_TRY(g_off.alloc(sizeof(int64_t) * (ni + 1), s));
    if (nitems)
        CSRK_LAUNCH(k3_items_a, (unsigned)div_up(nitems, 256), 256, 0, s, item_base.as<int64_t>(), ncells, nitems,
                    cell_off.as<int64_t>(), fpre.as<int64_t>(), c.nmax, c.nslabs, it_i0.as<int64_t>(), it_n.as<int32_t>(),
                    it_m.as<int32_t>(), it_meta.as<int32_t>(), epad.as<int32_t>(), gpad.as<int32_t>());
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<int32_t>{epad.as<int32_t>()}, nitems, e_off.as<int64_t>(), s)));
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<int32_t>{gpad.as<int32_t>()}, nitems, g_off.as<int64_t>(), s)));
    int64_t etot = 0, gtot = 0;
    CSRK_CUDA(cudaMemcpyAsync(&etot, e_off.as<int64_t>() + nitems, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&gtot, g_off.as<int64_t>() + nitems, 8, cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    if ((etot >> 3) >= ((int64_t)1 << 32) || (gtot >> 2) >= ((int64_t)1 << 32)) {
        set_error("matrix too large for the slab SpMV plan");
        return CSRK_EOVERFLOW;
    }
    CSRK_TRY(dev_alloc((void **)&P->col16, sizeof(uint16_t) * (size_t)(etot + 8), s));
    CSRK_TRY(dev_alloc(&P->val, sizeof(VT) * (size_t)(etot + 8), s));
    CSRK_TRY(dev_alloc((void **)&P->seg, sizeof(uint32_t) * (size_t)(gtot + 4), s));
    CSRK_TRY(dev_alloc((void **)&P->items, sizeof(P3Item) * ni, s));
    CSRK_TRY(dev_alloc((void **)&P->item_first, sizeof(int32_t) * ((size_t)npanels + 1), s));
    CSRK_CUDA(cudaMemsetAsync(P->col16, 0, sizeof(uint16_t) * (size_t)(etot + 8), s));
    CSRK_CUDA(cudaMemsetAsync(P->val, 0, sizeof(VT) * (size_t)(etot + 8), s));
    CSRK_CUDA(cudaMemsetAsync(P->seg, 0, sizeof(uint32_t) * (size_t)(gtot + 4), s));
    if (nitems)
        CSRK_LAUNCH(k3_items_b, (unsigned)div_up(nitems, 256), 256, 0, s, nitems, e_off.as<int64_t>(), g_off.as<int64_t>(),
                    it_n.as<int32_t>(), it_m.as<int32_t>(), it_meta.as<int32_t>(), P->items, P->seg);
    CSRK_LAUNCH((k3_place<VT>), (unsigned)div_up(nnz, 256), 256, 0, s, skey.as<int32_t>(), spacked.as<int32_t>(),
                sval.as<VT>(), nnz, cell_off.as<int64_t>(), item_base.as<int64_t>(), fpre.as<int64_t>(), e_off.as<int64_t>(),
                g_off.as<int64_t>(), c.nmax, P->col16, (VT *)P->val, P->seg);
    CSRK_LAUNCH(k3_item_first, (unsigned)div_up((int64_t)npanels + 1, 256), 256, 0, s, item_base.as<int64_t>(), npanels,
                c.nslabs, P->item_first);
    CSRK_CUDA(cudaStreamSynchronize(s));
    return CSRK_OK;
}

// ------------------------------------------------------------------ kernel
__device__ __forceinline__ uint32_t p3_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p3_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p3_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void p3_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p3_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p3_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p3_u32(bar)) : "memory");
}
__device__ __forceinline__ void p3_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(p3_u32(bar)), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void p3_bulk(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     p3_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(p3_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void p3_csync()  // barrier among the consumer warps only
{
    asm volatile("bar.sync 1, %0;" ::"n"(P3_CTHREADS) : "memory");
}

struct P3Args {
    const uint16_t *col16;
    const void *val;
    const uint32_t *seg;
    const P3Item *items;
    const int32_t *item_first, *panel_first, *chunk_row;
    int npanels, logw, nmax;
    int32_t ncols;
};

// VT: value type (float/double); XT: x type.  The product (numba's promotion) overwrites the value.
template <typename VT, typename XT>
__global__ void __launch_bounds__(P3_THREADS, 1)
k_psf3_spmv(P3Args a, const XT *__restrict__ x, double *__restrict__ y, double *__restrict__ chunk_sums,
            int *__restrict__ counter)
{
    extern __shared__ __align__(128) unsigned char p3_smem[];
    XT *xbuf = reinterpret_cast<XT *>(p3_smem);                                   // [2][W]
    unsigned char *stage0 = p3_smem + 2 * P3_SLAB_BYTES;                          // [STAGES][STAGE_BYTES]
    double *yacc = reinterpret_cast<double *>(stage0 + P3_STAGES * P3_STAGE_BYTES);  // [PR]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(yacc) + (size_t)P3_PR * 8);
    uint64_t *sfull = bars, *sempty = bars + P3_STAGES, *xfull = bars + 2 * P3_STAGES, *xempty = xfull + 2;
    __shared__ int s_panel;
    __shared__ int q_seg[P3_QCAP];
    __shared__ int q_cnt[3];   // long-segment queue length, rotating over items (see the reset below)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = 1 << a.logw;
    // fixed sub-buffers of a stage: col16 | val | seg
    const int off_val = (a.nmax * 2 + 15) & ~15;
    const int off_seg = off_val + ((a.nmax * (int)sizeof(VT) + 15) & ~15);
    if (tid == 0) {
        for (int i = 0; i < P3_STAGES; i++) {
            p3_init(&sfull[i], 1);
            p3_init(&sempty[i], P3_NC);
        }
        for (int i = 0; i < 2; i++) {
            p3_init(&xfull[i], 1);
            p3_init(&xempty[i], P3_NC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t q = 0;     // work items this CTA has gone through (stage ring position)
    uint32_t xs = 0;    // x slabs this CTA has gone through (slab ring position)
    while (true) {
        __syncthreads();
        if (tid == 0)
            s_panel = atomicAdd(counter, 1);
        __syncthreads();
        const int p = s_panel;
        if (p >= a.npanels)
            break;
        const int c0 = a.panel_first[p], nc = a.panel_first[p + 1] - c0;
        const int i0 = a.item_first[p], i1 = a.item_first[p + 1];
        if (warp == P3_NC) {
            // ---------------- producer
            if (lane == 0) {
                for (int it = i0; it < i1; it++) {
                    const P3Item w = a.items[it];
                    if (w.flags & 1) {
                        const int xb = xs & 1;
                        p3_wait(&xempty[xb], ((xs >> 1) & 1) ^ 1);
                        const int64_t col0 = (int64_t)w.slab << a.logw;
                        const int ncol = (int)min((int64_t)W, (int64_t)a.ncols - col0);
                        const uint32_t bytes = (uint32_t)ncol * sizeof(XT), bulk = bytes & ~15u;
                        XT *dst = xbuf + (size_t)xb * W;
                        for (int k = bulk / sizeof(XT); k < ncol; k++)
                            dst[k] = x[col0 + k];
                        p3_expect(&xfull[xb], bulk);
                        if (bulk)
                            p3_bulk(dst, x + col0, bulk, &xfull[xb]);
                        xs++;
                    }
                    const int st = q % P3_STAGES;
                    p3_wait(&sempty[st], ((q / P3_STAGES) & 1) ^ 1);
                    unsigned char *sb = stage0 + (size_t)st * P3_STAGE_BYTES;
                    const uint32_t nb_c = (((uint32_t)w.n + 7) & ~7u) * 2;
                    const uint32_t nb_v = (((uint32_t)w.n + 7) & ~7u) * (uint32_t)sizeof(VT);
                    const uint32_t nb_s = (((uint32_t)w.m + 1 + 3) & ~3u) * 4;
                    p3_expect(&sfull[st], nb_c + nb_v + nb_s);
                    p3_bulk(sb, a.col16 + (size_t)w.e8 * 8, nb_c, &sfull[st]);
                    p3_bulk(sb + off_val, reinterpret_cast<const VT *>(a.val) + (size_t)w.e8 * 8, nb_v, &sfull[st]);
                    p3_bulk(sb + off_seg, a.seg + (size_t)w.g4 * 4, nb_s, &sfull[st]);
                    q++;
                }
            } else {
                // keep the ring counters of all producer lanes in step (only lane 0 uses them)
                for (int it = i0; it < i1; it++) {
                    if (a.items[it].flags & 1)
                        xs++;
                    q++;
                }
            }
        } else {
            // ---------------- consumers
            for (int i = tid; i < nc; i += P3_CTHREADS)
                yacc[i] = 0.0;
            if (tid == 0)
                q_cnt[q % 3] = 0;
            int xb = (xs - 1) & 1;
            for (int it = i0; it < i1; it++) {
                const P3Item w = a.items[it];
                const int st = q % P3_STAGES;
                unsigned char *sb = stage0 + (size_t)st * P3_STAGE_BYTES;
                const uint16_t *cols = reinterpret_cast<const uint16_t *>(sb);
                VT *vals = reinterpret_cast<VT *>(sb + off_val);
                const uint32_t *segs = reinterpret_cast<const uint32_t *>(sb + off_seg);
                if (w.flags & 1) {
                    xb = xs & 1;
                    p3_wait(&xfull[xb], (xs >> 1) & 1);
                    xs++;
                }
                p3_wait(&sfull[st], (q / P3_STAGES) & 1);
                const XT *xsl = xbuf + (size_t)xb * W;
                const int n = w.n, m = w.m;
                // the counter of the NEXT item is cleared now: stragglers can only still be reading
                // the previous item's counter (a third slot), never this one
                int *q_n = &q_cnt[q % 3];
                if (tid == 0)
                    q_cnt[(q + 1) % 3] = 0;
                // phase 1: product in place
                for (int i = tid; i < n; i += P3_CTHREADS) {
                    if (std::is_same<VT, float>::value && std::is_same<XT, float>::value)
                        vals[i] = (VT)((float)xsl[cols[i]] * (float)vals[i]);
                    else
                        vals[i] = (VT)((double)xsl[cols[i]] * (double)vals[i]);
                }
                if ((w.flags & 2) && lane == 0)
                    p3_arrive(&xempty[xb]);   // this warp is done with the slab (x is only read in phase 1)
                p3_csync();
                // phase 2: one thread per short segment, long ones queued for the warps
                for (int g = tid; g < m; g += P3_CTHREADS) {
                    const uint32_t d = segs[g];
                    const int s0 = (int)(d >> 13), e0 = (int)(segs[g + 1] >> 13);
                    if (e0 - s0 > P3_LONG) {
                        const int slot = atomicAdd(q_n, 1);
                        if (slot < P3_QCAP)
                            q_seg[slot] = g;
                        else {   // queue full (cannot happen for NMAX/LONG <= QCAP): do it here
                            double sum = 0.0;
                            for (int i = s0; i < e0; i++)
                                sum += (double)vals[i];
                            yacc[d & 0x1FFF] += sum;
                        }
                    } else {
                        double sum = 0.0;
                        for (int i = s0; i < e0; i++)
                            sum += (double)vals[i];
                        yacc[d & 0x1FFF] += sum;
                    }
                }
                p3_csync();
                const int nq = min(*q_n, P3_QCAP);
                for (int k = warp; k < nq; k += P3_NC) {
                    const int g = q_seg[k];
                    const uint32_t d = segs[g];
                    const int s0 = (int)(d >> 13), e0 = (int)(segs[g + 1] >> 13);
                    double sum = 0.0;
                    for (int i = s0 + lane; i < e0; i += 32)
                        sum += (double)vals[i];
                    sum = warp_sum(sum);
                    if (lane == 0)
                        yacc[d & 0x1FFF] += sum;
                }
                __syncwarp();
                if (lane == 0)
                    p3_arrive(&sempty[st]);   // this warp no longer reads the stage
                q++;
                // no barrier here: the next item's phase-1 barrier keeps its phase 2 (same accumulators,
                // same queue array) behind everybody's phase 2 of this item
            }
        }
        __syncthreads();
        // epilogue: rows of the panel
        for (int i = tid; i < nc; i += P3_THREADS) {
            const int32_t row = a.chunk_row[c0 + i];
            if (row < 0)
                chunk_sums[c0 + i] = yacc[i];
            else
                y[row] = yacc[i];
        }
    }
}

template <typename VT, typename XT>
static int psf3_launch(csrk_matrix *h, Psf3Plan *P, const void *d_x, double *d_y, cudaStream_t s)
{
    auto k = k_psf3_spmv<VT, XT>;
    static bool optin = false;
    if (!optin) {
        CSRK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P3_SMEM));
        optin = true;
    }
    DevBuf counter, sums;
    CSRK_TRY(counter.alloc_zero(sizeof(int), s));
    CSRK_TRY(sums.alloc(sizeof(double) * (size_t)std::max<int64_t>(P->nchunks, 1), s));
    P3Args a{P->col16, P->val, P->seg, P->items, P->item_first, P->panel_first, P->chunk_row,
             P->npanels, P->cfg.logw, P->cfg.nmax, h->ncols};
    const int grid = std::min(P->npanels, ctx().sm_count);
    CSRK_LAUNCH(k, (unsigned)grid, P3_THREADS, P3_SMEM, s, a, (const XT *)d_x, d_y, sums.as<double>(), counter.as<int>());
    if (P->n_split)
        CSRK_LAUNCH(k3_fixup, (unsigned)div_up(P->n_split, 128), 128, 0, s, P->split, P->n_split, sums.as<double>(), d_y);
    return CSRK_OK;
}

bool psf3_supported(const csrk_matrix *h, int x_kind)
{
    return (h->val_kind == 4 && x_kind == 4) || h->val_kind == 8;
}

int psf3_build(csrk_matrix *h, int x_kind, Psf3Plan **out, cudaStream_t s)
{
    *out = nullptr;
    Psf3Plan *P = new (std::nothrow) Psf3Plan();
    if (!P) {
        set_error("host allocation failed");
        return CSRK_ENOMEM;
    }
    WsScope scope;  // build temporaries come from the workspace arena
    P->x_kind = x_kind;
    P->val_kind = h->val_kind;
    P3Cfg &c = P->cfg;
    c.logw = x_kind == 4 ? 13 : 12;
    c.nslabs = (int)div_up((int64_t)h->ncols, (int64_t)1 << c.logw);
    // stage = 2n + V n + 4 (n + 1) bytes (+ alignment slack) <= 32 KB
    c.nmax = h->val_kind == 4 ? 3072 : 2048;
    const int sms = std::max(ctx().sm_count, 1);
    c.PN = std::min<int64_t>(std::max<int64_t>(h->nnz / sms, 65536), 1 << 21);
    c.CH = std::max<int64_t>((c.PN / 2) & ~(int64_t)7, 1024);
    c.D = c.PN * P3_PR;
    c.wfull = std::max<int64_t>(c.CH * P3_PR, c.PN);
    int rc;
    if (h->rp_is64)
        rc = h->val_kind == 4 ? psf3_build_typed<int64_t, float>(h, P, s) : psf3_build_typed<int64_t, double>(h, P, s);
    else
        rc = h->val_kind == 4 ? psf3_build_typed<int32_t, float>(h, P, s) : psf3_build_typed<int32_t, double>(h, P, s);
    if (rc != CSRK_OK) {
        psf3_destroy(P, s);
        return rc;
    }
    *out = P;
    return CSRK_OK;
}

int psf3_run(csrk_matrix *h, Psf3Plan *P, const void *d_x, double *d_y, cudaStream_t s)
{
    if (h->val_kind == 4)
        return psf3_launch<float, float>(h, P, d_x, d_y, s);
    if (P->x_kind == 4)
        return psf3_launch<double, float>(h, P, d_x, d_y, s);
    return psf3_launch<double, double>(h, P, d_x, d_y, s);
}

}  // namespace csrk
