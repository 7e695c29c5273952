// spmv_stream.cu -- mult_vec (csr/kernels/numba/__init__.py:55-67) for matrices whose x does not fit L1:
// the slab-stream kernel.
//
// Why.  The CSR tile kernel (spmv.cu) gathers x through L1/L2: with columns spread over an x of several
// MB every 4-byte gather costs one 32-byte L2 sector, 5x the bytes of the (colind, value) stream itself,
// and the kernel sits on the L2->SM fabric at a third of the HBM roofline (profiles/r01_spmv_tile_ncu.md).
// Here x is staged in SHARED memory, slab by slab, and the entry stream is re-laid out ONCE per handle so
// that every byte of it is still read exactly once, fully coalesced:
//
//   * x is cut into slabs of S columns (~80 KB); one persistent CTA per SM keeps TWO slabs in shared
//     memory, a producer warp fetching slab s+1 with cp.async.bulk (TMA) + mbarriers while the consumer
//     warps work on slab s.  The gathers become LDS.
//   * rows longer than 4096 entries are cut into interleaved pieces (piece j = entries j, j+n, j+2n, ... so
//     that every piece spans the row's whole column range); the pieces ("pseudo-rows") are sorted by length and
//     dealt to G*NW bins (one bin per consumer WARP) in snake order, so every warp owns ~Q/(G*NW) pseudo-rows
//     with the same number of entries and the same mix of long and short ones: no dynamic scheduling, no
//     inter-warp communication, no CTA-wide barrier in the whole kernel.
//   * a warp's entries are stored as ONE contiguous stream ordered by slab ("cells"), 8 bytes per entry as
//     in CSR: a packed word (local row << 16 | column inside the slab) and the value.  Inside a cell the
//     entries keep CSR order, so the entries of a pseudo-row are adjacent.
//   * a lane takes 8 consecutive entries (128-bit loads), sums runs of equal rows in registers (float64)
//     and adds every finished run to the warp's private float64 accumulators in shared memory with a plain
//     read-modify-write.  Runs that cross lanes: when no row fills a whole lane a row can only sit in the
//     tail of one lane and the head of the next, so all tails are flushed first and all heads second
//     (never the same address in one instruction); blocks with longer runs join them with one segmented
//     shuffle scan.  No atomics anywhere: results are deterministic.
//   * cells are padded to multiples of 8 entries with INERT entries (value 0, column S = a slot of the
//     slab buffer that always holds 0.0, row = the row before them), so the kernel never tests entries.
//   * pieces of split rows go to a carry array and a tiny fix-up kernel adds them in piece order.
//
// Cost model: HBM bytes are the stream (nnz*(4+V)) + x once + y; the L2->SM fabric carries
// the stream plus G copies of x, which is why auto mode only picks this kernel when G*ncols*X is below the
// stream size.
#include <type_traits>

#include "expand.cuh"
#include "radix.cuh"
#include "spmv.cuh"

namespace csrk {

constexpr int ST_E = 8;                    // entries per lane per block (two 128-bit loads of packed words)
constexpr int ST_BLK = 32 * ST_E;          // entries per warp block
constexpr int ST_PIECE = 4096;             // longest pseudo-row
constexpr uint32_t ST_NOROW = 0xffffffffu;
constexpr int ST_MAX_WARPS = 31;           // consumer warps (+1 producer warp = 1024 threads)

struct StreamPlan {
    int G = 0, NW = 0, nslab = 0, S = 0, P = 0, x_kind = 0;
    int slab_bytes = 0;
    size_t smem_bytes = 0;
    int64_t Q = 0;
    int n_split = 0;
    uint32_t *idx = nullptr;   // [npad]  local row << 16 | column - slab*S
    void *val = nullptr;       // [npad]  VT (absent for structure-only matrices)
    int64_t *binbase = nullptr;  // [G*NW] start of bin b's stream
    uint32_t *cstart = nullptr;  // [G*NW][nslab+1] start of cell s relative to the bin's stream ([nslab] = its end); multiples of 8
    int32_t *rowmap = nullptr; // [G*NW][P]: >= 0 row of y, -1 unused, <= -2 carry slot -(v+2)
    int32_t *split = nullptr;  // [3*n_split]: row, first carry slot, number of pieces
};

void stream_destroy(StreamPlan *p, cudaStream_t s)
{
    if (!p)
        return;
    dev_free(p->idx, s);
    dev_free(p->val, s);
    dev_free(p->binbase, s);
    dev_free(p->cstart, s);
    dev_free(p->rowmap, s);
    dev_free(p->split, s);
    delete p;
}

// ------------------------------------------------------------------ plan builder
template <typename RPT> struct StPieceLoader {
    const RPT *rp;
    __device__ __forceinline__ int64_t operator()(int64_t r) const
    {
        const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
        return len > ST_PIECE ? (len + ST_PIECE - 1) / ST_PIECE : 1;
    }
};

// one thread per row: sort key (ST_PIECE - length: longest first), id and y destination of each piece
template <typename RPT>
__global__ void k_st_pieces(const RPT *__restrict__ rp, int32_t nrows, const int64_t *__restrict__ qbase,
                            int32_t *__restrict__ qkey, int32_t *__restrict__ qid, int32_t *__restrict__ qdest,
                            int32_t *__restrict__ split, int *__restrict__ split_cnt)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows)
        return;
    const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
    const int64_t q0 = qbase[r], nq = qbase[r + 1] - q0;
    for (int64_t j = 0; j < nq; j++) {
        const int64_t l = (len - j + nq - 1) / nq;   // piece j = entries j, j+nq, j+2nq, ... of the row
        qkey[q0 + j] = (int32_t)(ST_PIECE - l);
        qid[q0 + j] = (int32_t)(q0 + j);
        qdest[q0 + j] = nq == 1 ? (int32_t)r : -(int32_t)(q0 + j) - 2;
    }
    if (nq > 1) {
        const int k = atomicAdd(split_cnt, 1);
        split[3 * k] = (int32_t)r;
        split[3 * k + 1] = (int32_t)q0;
        split[3 * k + 2] = (int32_t)nq;
    }
}

// sorted position j -> bin (snake order over the B bins) and local row j / B
__global__ void k_st_deal(const int32_t *__restrict__ order, const int32_t *__restrict__ qdest, int64_t Q, int B, int P,
                          int32_t *__restrict__ qbl, int32_t *__restrict__ rowmap)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Q)
        return;
    const int q = order[j];
    const int round = (int)(j / B), k = (int)(j % B);
    const int bin = (round & 1) ? B - 1 - k : k;
    qbl[q] = bin << 16 | round;
    rowmap[(int64_t)bin * P + round] = qdest[q];
}

__global__ void k_st_fill_i32(int32_t *p, int64_t n, int32_t v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}

// Sort input, in PIECE-MAJOR order: position e' of row r's range is entry i of piece j (the pieces of a row
// one after the other), which is entry t = i*nq + j of the row.  A long row's pieces are interleaved so that
// every piece samples the row's whole column range -- a contiguous piece of a dense row would put its 4096
// entries into one or two slabs, and since the two-deep x ring keeps the warps of a CTA within one slab of
// each other, the warp holding it would stall the other thirty.  Piece-major order keeps the entries of a
// pseudo-row adjacent inside every cell after the stable sort.
// key = bin*nslab + slab, packed = local row << 16 | column inside the slab, pval = the entry's value.
template <typename RPT, typename VT>
__global__ void k_st_keys(const RPT *__restrict__ rp, const int32_t *__restrict__ ci, const VT *__restrict__ vs,
                          const int32_t *__restrict__ rows, int64_t nnz, const int64_t *__restrict__ qbase,
                          const int32_t *__restrict__ qbl, int S, int nslab, int32_t *__restrict__ key,
                          int32_t *__restrict__ packed, VT *__restrict__ pval)
{
    const int64_t ep = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ep >= nnz)
        return;
    const int32_t r = rows[ep];
    const int64_t r0 = (int64_t)rp[r], len = (int64_t)rp[r + 1] - r0;
    const int64_t u = ep - r0;
    const int64_t q0 = qbase[r], nq = qbase[r + 1] - q0;
    int64_t j = 0, t = u;
    if (nq > 1) {
        const int64_t f = len / nq, m = len % nq;   // the first m pieces hold f+1 entries, the others f
        int64_t i;
        if (u < m * (f + 1)) {
            j = u / (f + 1);
            i = u % (f + 1);
        } else {
            const int64_t v = u - m * (f + 1);
            j = m + v / f;
            i = v % f;
        }
        t = i * nq + j;
    }
    const int64_t e = r0 + t;
    const int32_t bl = qbl[q0 + j];
    const int32_t c = ci[e];
    const int slab = c / S;
    key[ep] = (bl >> 16) * nslab + slab;
    packed[ep] = (int32_t)(((uint32_t)(bl & 0xffff) << 16) | (uint32_t)(c - slab * S));
    if constexpr (!std::is_same<VT, NoPayload>::value)
        pval[ep] = vs[e];
}

struct StPadLoader {
    const int64_t *cs;
    __device__ __forceinline__ int64_t operator()(int64_t k) const { return (cs[k + 1] - cs[k] + ST_E - 1) & ~(int64_t)(ST_E - 1); }
};

__global__ void k_st_cells(const int64_t *__restrict__ pstart, int B, int nslab, int64_t *__restrict__ binbase,
                           uint32_t *__restrict__ cstart)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * (nslab + 1))
        return;
    const int64_t b = i / (nslab + 1);
    const int s = (int)(i % (nslab + 1));
    const int64_t base = pstart[b * nslab];
    if (s == 0)
        binbase[b] = base;
    cstart[i] = (uint32_t)(pstart[b * nslab + s] - base);
}

template <typename VT>
__global__ void k_st_place(const int32_t *__restrict__ skeys, const int32_t *__restrict__ spacked, const VT *__restrict__ svals,
                           int64_t nnz, const int64_t *__restrict__ cs, const int64_t *__restrict__ pstart,
                           uint32_t *__restrict__ idx, VT *__restrict__ val)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz)
        return;
    const int32_t k = skeys[i];
    const int64_t dst = pstart[k] + (i - cs[k]);
    idx[dst] = (uint32_t)spacked[i];
    if constexpr (!std::is_same<VT, NoPayload>::value)
        val[dst] = svals[i];
}

// inert entries behind the last real entry of every cell: same row, column S (always 0.0), value 0
template <typename VT>
__global__ void k_st_pad(const int32_t *__restrict__ spacked, const int64_t *__restrict__ cs, const int64_t *__restrict__ pstart,
                         int64_t ncells, int S, uint32_t *__restrict__ idx, VT *__restrict__ val)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncells)
        return;
    const int64_t cnt = cs[k + 1] - cs[k];
    if (cnt == 0 || (cnt & (ST_E - 1)) == 0)
        return;
    const uint32_t word = ((uint32_t)spacked[cs[k + 1] - 1] & 0xffff0000u) | (uint32_t)S;
    for (int64_t p = pstart[k] + cnt; p < pstart[k + 1]; p++) {
        idx[p] = word;
        if constexpr (!std::is_same<VT, NoPayload>::value)
            val[p] = VT(0);
    }
}

static int st_bits(int64_t n)
{
    int b = 1;
    while (((int64_t)1 << b) < n)
        b++;
    return b;
}

template <typename RPT, typename VT>
static int stream_build_typed(csrk_matrix *h, StreamPlan *P, cudaStream_t s)
{
    constexpr bool HASV = !std::is_same<VT, NoPayload>::value;
    const RPT *rp = (const RPT *)h->rp;
    const int32_t nrows = h->nrows;
    const int64_t nnz = h->nnz;
    const int B = P->G * P->NW;

    // 1. pseudo-rows
    DevBuf qbase;
    CSRK_TRY(qbase.alloc(sizeof(int64_t) * ((size_t)nrows + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(StPieceLoader<RPT>{rp}, (int64_t)nrows, qbase.as<int64_t>(), s)));
    int64_t Q = 0;
    CSRK_CUDA(cudaMemcpyAsync(&Q, qbase.as<int64_t>() + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    P->Q = Q;
    P->P = (int)std::max<int64_t>(div_up(Q, B), 1);
    // shared memory: two x slabs + NW*P float64 accumulators + 4 mbarriers
    const size_t acc_bytes = (size_t)P->NW * P->P * 8;
    const size_t smem_max = ctx().smem_optin;
    if (P->P > 65534 || acc_bytes + 64 + 2 * 8192 > smem_max)
        return CSRK_EOVERFLOW;  // too many rows for shared-memory accumulators: stay on the tile kernel
    // a buffer is one slab + 128 bytes (the always-zero slot the inert entries point at)
    int64_t slab = (int64_t)((smem_max - acc_bytes - 64) / 2 - 128) & ~(int64_t)127;
    slab = std::min<int64_t>(slab, (int64_t)65408 * P->x_kind);   // columns inside a slab (and the zero slot) fit 16 bits
    slab = std::min<int64_t>(slab, (((int64_t)h->ncols * P->x_kind) + 127) & ~(int64_t)127);
    const int64_t cap = options().stream_slab_bytes.load();
    if (cap > 0)
        slab = std::min<int64_t>(slab, std::max<int64_t>(cap & ~(int64_t)127, 128));
    slab = std::max<int64_t>(slab, 128);
    P->slab_bytes = (int)slab;
    P->S = (int)(slab / P->x_kind);
    P->nslab = (int)std::max<int64_t>(div_up((int64_t)h->ncols, P->S), 1);
    P->smem_bytes = 2 * ((size_t)slab + 128) + acc_bytes + 64;
    const int64_t ncells = (int64_t)B * P->nslab;
    if (ncells >= ((int64_t)1 << 30) || nnz / B + 2 * ST_PIECE + (int64_t)ST_E * P->nslab >= ((int64_t)1 << 31))
        return CSRK_EOVERFLOW;   // cell offsets inside a bin's stream are 32-bit

    DevBuf qkey, qid, qdest, order, qbl, splitcnt;
    CSRK_TRY(qkey.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(qid.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(qdest.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(order.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(qbl.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(splitcnt.alloc_zero(sizeof(int), s));
    const int64_t max_split = nnz / ST_PIECE + 1;
    CSRK_TRY(dev_alloc((void **)&P->split, sizeof(int32_t) * 3 * (size_t)max_split, s));
    CSRK_LAUNCH((k_st_pieces<RPT>), (unsigned)div_up((int64_t)nrows, 256), 256, 0, s, rp, nrows, qbase.as<int64_t>(),
                qkey.as<int32_t>(), qid.as<int32_t>(), qdest.as<int32_t>(), P->split, splitcnt.as<int>());
    // 2. longest first (stable), dealt to the bins in snake order
    CSRK_TRY((radix_sort_by_key<NoPayload>(qkey.as<int32_t>(), qid.as<int32_t>(), (const NoPayload *)nullptr, Q, 13,
                                           order.as<int32_t>(), (NoPayload *)nullptr, s)));
    CSRK_TRY(dev_alloc((void **)&P->rowmap, sizeof(int32_t) * (size_t)B * P->P, s));
    CSRK_LAUNCH(k_st_fill_i32, (unsigned)div_up((int64_t)B * P->P, 256), 256, 0, s, P->rowmap, (int64_t)B * P->P, -1);
    CSRK_LAUNCH(k_st_deal, (unsigned)div_up(Q, 256), 256, 0, s, order.as<int32_t>(), qdest.as<int32_t>(), Q, B, P->P,
                qbl.as<int32_t>(), P->rowmap);
    CSRK_TRACE_MARK("stream plan: pieces dealt", s);

    // 3. entries: key = (bin, slab), stable sort keeps CSR order inside a cell
    DevBuf rows, key, packed, pval, skeys, spacked, svals;
    CSRK_TRY(rows.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(key.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(packed.alloc(sizeof(int32_t) * (size_t)nnz, s));
    if (HASV)
        CSRK_TRY(pval.alloc(sizeof(VT) * (size_t)nnz, s));
    CSRK_LAUNCH((k_expand_rows<RPT>), (unsigned)div_up(nnz, EXP_TILE), 256, 0, s, rp, nrows, nnz, rows.as<int32_t>());
    CSRK_LAUNCH((k_st_keys<RPT, VT>), (unsigned)div_up(nnz, 256), 256, 0, s, rp, h->ci, (const VT *)h->vs,
                rows.as<int32_t>(), nnz, qbase.as<int64_t>(), qbl.as<int32_t>(), P->S, P->nslab, key.as<int32_t>(),
                packed.as<int32_t>(), HASV ? pval.as<VT>() : nullptr);
    CSRK_TRY(skeys.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(spacked.alloc(sizeof(int32_t) * (size_t)nnz, s));
    if (HASV)
        CSRK_TRY(svals.alloc(sizeof(VT) * (size_t)nnz, s));
    CSRK_TRY((radix_sort_by_key<VT>(key.as<int32_t>(), packed.as<int32_t>(), HASV ? pval.as<VT>() : nullptr, nnz, st_bits(ncells),
                                    spacked.as<int32_t>(), HASV ? svals.as<VT>() : nullptr, s, skeys.as<int32_t>())));
    CSRK_TRACE_MARK("stream plan: entries sorted", s);

    // 4. cells padded to multiples of 8 entries (a lane's share), bin streams contiguous
    DevBuf cs, pstart;
    CSRK_TRY(cs.alloc(sizeof(int64_t) * ((size_t)ncells + 1), s));
    CSRK_TRY(pstart.alloc(sizeof(int64_t) * ((size_t)ncells + 1), s));
    CSRK_LAUNCH((k_key_bounds<int64_t>), (unsigned)div_up(div_up(nnz + 1, 4), 256), 256, 0, s, skeys.as<int32_t>(), nnz,
                (int32_t)ncells, cs.as<int64_t>());
    CSRK_TRY((exclusive_scan<int64_t>(StPadLoader{cs.as<int64_t>()}, ncells, pstart.as<int64_t>(), s)));
    const size_t npad = (size_t)nnz + (ST_E - 1) * (size_t)ncells + ST_E;
    CSRK_TRY(dev_alloc((void **)&P->idx, sizeof(uint32_t) * npad, s));
    if (HASV)
        CSRK_TRY(dev_alloc(&P->val, sizeof(VT) * npad, s));
    CSRK_TRY(dev_alloc((void **)&P->binbase, sizeof(int64_t) * (size_t)B, s));
    CSRK_TRY(dev_alloc((void **)&P->cstart, sizeof(uint32_t) * (size_t)B * (P->nslab + 1), s));
    CSRK_LAUNCH(k_st_cells, (unsigned)div_up((int64_t)B * (P->nslab + 1), 256), 256, 0, s, pstart.as<int64_t>(), B,
                P->nslab, P->binbase, P->cstart);
    CSRK_LAUNCH((k_st_place<VT>), (unsigned)div_up(nnz, 256), 256, 0, s, skeys.as<int32_t>(), spacked.as<int32_t>(),
                HASV ? svals.as<VT>() : nullptr, nnz, cs.as<int64_t>(), pstart.as<int64_t>(), P->idx, (VT *)P->val);
    CSRK_LAUNCH((k_st_pad<VT>), (unsigned)div_up(ncells, 256), 256, 0, s, spacked.as<int32_t>(), cs.as<int64_t>(),
                pstart.as<int64_t>(), ncells, P->S, P->idx, (VT *)P->val);
    CSRK_CUDA(cudaMemcpyAsync(&P->n_split, splitcnt.as<int>(), sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    CSRK_TRACE_MARK("stream plan: placed", s);
    return CSRK_OK;
}

int stream_build(csrk_matrix *h, int x_kind, StreamPlan **out, cudaStream_t s)
{
    *out = nullptr;
    StreamPlan *P = new (std::nothrow) StreamPlan();
    if (!P) {
        set_error("host allocation failed");
        return CSRK_ENOMEM;
    }
    P->x_kind = x_kind;
    const int64_t g = options().stream_ctas.load(), nw = options().stream_warps.load();
    P->G = (int)(g > 0 ? std::min<int64_t>(g, 4 * (int64_t)ctx().sm_count) : ctx().sm_count);
    P->NW = (int)std::min<int64_t>(std::max<int64_t>(nw, 1), ST_MAX_WARPS);
    int rc;
    {
        WsScope scope;
        if (h->rp_is64)
            rc = h->val_kind == 4   ? stream_build_typed<int64_t, float>(h, P, s)
                 : h->val_kind == 8 ? stream_build_typed<int64_t, double>(h, P, s)
                                    : stream_build_typed<int64_t, NoPayload>(h, P, s);
        else
            rc = h->val_kind == 4   ? stream_build_typed<int32_t, float>(h, P, s)
                 : h->val_kind == 8 ? stream_build_typed<int32_t, double>(h, P, s)
                                    : stream_build_typed<int32_t, NoPayload>(h, P, s);
        if (rc != CSRK_OK)
            (void)cudaStreamSynchronize(s);  // nothing may still read the workspace when the scope rewinds it
    }
    if (rc != CSRK_OK) {
        stream_destroy(P, s);
        return rc;
    }
    *out = P;
    return CSRK_OK;
}

// ------------------------------------------------------------------ kernel
__device__ __forceinline__ uint32_t st_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_bar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void st_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(st_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(st_u32(bar)), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void st_bulk(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     st_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(st_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned lanemask_le()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint32_t *p)
{
    const int4 q = ld_stream_int4(p);
    return make_uint4((uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w);
}

// the 8 values of a lane (nothing for a structure-only matrix)
template <typename VT> struct StVals {
    VT v[ST_E];
    __device__ __forceinline__ void load(const VT *p)
    {
        if constexpr (sizeof(VT) == 4) {
            const float4 a = ld_stream_float4(p), b = ld_stream_float4(p + 4);
            v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < ST_E; k += 2) {
                const double2 a = ld_stream_double2(p + k);
                v[k] = a.x, v[k + 1] = a.y;
            }
        }
    }
};
template <> struct StVals<NoVal> {
    __device__ __forceinline__ void load(const NoVal *) {}
};

template <typename VT, typename XT> __device__ __forceinline__ double st_prod(const StVals<VT> &v, int k, XT xv)
{
    if constexpr (std::is_same<VT, NoVal>::value) {
        return (double)xv;
    } else {
        using PT = typename Prod<VT, XT>::type;   // numba's promotion: f4*f4 -> f4, else f8; the product is
        if constexpr (std::is_same<PT, float>::value)  // rounded before it is added (no FMA contraction)
            return (double)__fmul_rn((float)xv, (float)v.v[k]);
        else
            return __dmul_rn((double)xv, (double)v.v[k]);
    }
}

struct StArgs {
    const uint32_t *idx;
    const void *val;
    const int64_t *binbase;
    const uint32_t *cstart;
    const int32_t *rowmap;
    int nslab, S, P, NW, slab_bytes;
    int32_t ncols;
};

// One 256-entry block of a cell.  This lane's 8 entries are the packed words w[] / values vv (all real or
// inert when `lv`; the lane lies beyond the cell's end otherwise).  (carry_row, carry_sum) is an open run
// handed from a scan block to the next block (all lanes hold the same copy).
template <typename VT, typename XT>
__device__ __forceinline__ void st_block(const uint32_t (&w)[ST_E], const StVals<VT> &vv, const bool lv,
                                         const XT *__restrict__ xs, double *__restrict__ acc_w, const int lane,
                                         uint32_t &carry_row, double &carry_sum)
{
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t hr = ST_NOROW, tr = ST_NOROW;
    double H = 0.0, V = 0.0;
    bool single = true;
    if (lv) {
        // Branch-free: all gathers, products and row comparisons are independent; d[k] becomes the running
        // sum of the run entry k lies in; a run that ends at k < 7 and did not start at entry 0 is complete
        // inside this lane and is added to its accumulator -- all loads first, then all stores (the rows of
        // distinct runs are distinct, in this lane and across the warp).
        uint32_t r[ST_E];
        double d[ST_E];
#pragma unroll
        for (int k = 0; k < ST_E; k++) {
            r[k] = w[k] >> 16;
            d[k] = st_prod<VT, XT>(vv, k, xs[w[k] & 0xffffu]);
        }
        bool bnd[ST_E];        // bnd[k]: entry k starts a new run
        bool inner[ST_E];      // inner[k]: a boundary at or before k
        bnd[0] = false;
        inner[0] = false;
        H = d[0];
#pragma unroll
        for (int k = 1; k < ST_E; k++) {
            bnd[k] = r[k] != r[k - 1];
            inner[k] = inner[k - 1] || bnd[k];
            if (!bnd[k])
                d[k] += d[k - 1];
            if (!inner[k])
                H = d[k];      // still inside the first run
        }
        single = !inner[ST_E - 1];
        hr = r[0];
        tr = r[ST_E - 1];
        V = d[ST_E - 1];
        double old[ST_E - 1];
#pragma unroll
        for (int k = 1; k < ST_E - 1; k++)          // a run ending at k: bnd[k+1]; not the first run: inner[k]
            if (bnd[k + 1] && inner[k])
                old[k] = acc_w[r[k]];
#pragma unroll
        for (int k = 1; k < ST_E - 1; k++)
            if (bnd[k + 1] && inner[k])
                acc_w[r[k]] = old[k] + d[k];
    }
    if (!__any_sync(FULL, lv && single)) {
        // No row fills a whole lane, so a row occupies at most the last run of one lane and the first run of
        // the next: flush all last runs, then all first runs -- never one address twice in an instruction.
        if (lane == 0 && carry_row != ST_NOROW)
            acc_w[carry_row] += carry_sum;   // the run a scan block left open
        carry_row = ST_NOROW;
        carry_sum = 0.0;
        if (lv)
            acc_w[tr] += V;
        __syncwarp();
        if (lv)
            acc_w[hr] += H;
        __syncwarp();
        return;
    }
    // Long runs: join the pieces of a run that spans lanes with a segmented inclusive scan over the lanes.
    uint32_t prev_tr = __shfl_up_sync(FULL, tr, 1);
    if (lane == 0)
        prev_tr = carry_row;
    const bool match = lv && hr == prev_tr;   // my first run continues the chain of the lane before me
    const bool cont = single && match;        // ... and I am nothing but that run: the chain passes through
    const unsigned starts = __ballot_sync(FULL, !cont);
    const unsigned below = starts & lanemask_le();
    const int dist = below ? lane - (31 - __clz(below)) : lane;   // lanes between me and the head of my chain
    double out = V;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(FULL, out, o);
        if (dist >= o)
            out += t;
    }
    if (below == 0)
        out += carry_sum;   // the chain started in an earlier block
    double prev_out = __shfl_up_sync(FULL, out, 1);
    if (lane == 0)
        prev_out = carry_sum;
    if (lane == 0 && !match && carry_row != ST_NOROW)
        acc_w[carry_row] += carry_sum;   // the carried run ended with the previous block
    if (lv && !single)
        acc_w[hr] += match ? H + prev_out : H;   // a chain ends in my first run
    const int next_match = __shfl_down_sync(FULL, match ? 1 : 0, 1);
    if (lv && lane < 31 && !next_match)
        acc_w[tr] += out;   // nobody continues my last run
    carry_row = __shfl_sync(FULL, tr, 31);
    carry_sum = __shfl_sync(FULL, out, 31);
    __syncwarp();
}

template <typename VT, typename XT, bool MULTI>
__global__ void __launch_bounds__(1024, 1)
k_spmv_stream(StArgs a, const XT *__restrict__ x, YOut y, double *__restrict__ carry)
{
    extern __shared__ __align__(128) unsigned char st_smem[];
    const size_t xstride = (size_t)a.slab_bytes + 128;                                 // slab + the zero slot
    unsigned char *xbuf = st_smem;                                                     // [2][xstride]
    double *acc = reinterpret_cast<double *>(st_smem + 2 * xstride);                   // [NW][P]
    uint64_t *bars = reinterpret_cast<uint64_t *>(acc + (size_t)a.NW * a.P);           // full[2], empty[2]
    uint64_t *full = bars, *empty = bars + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        st_bar_init(&full[0], 1);
        st_bar_init(&full[1], 1);
        st_bar_init(&empty[0], a.NW);
        st_bar_init(&empty[1], a.NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < 2)
        reinterpret_cast<XT *>(xbuf + tid * xstride)[a.S] = XT(0);   // what the inert entries multiply
    __syncthreads();   // the only CTA-wide barrier
    // CTA g walks the slabs starting at slab g*nslab/G and wraps around, so that at any moment the CTAs pull
    // different parts of x out of L2
    const int slab0 = (int)(((int64_t)blockIdx.x * a.nslab) / gridDim.x);

    if (warp == a.NW) {
        // ---------------- producer: the k-th slab of this CTA's walk into buffer k & 1
        if (lane == 0) {
            for (int k = 0; k < a.nslab; k++) {
                const int st = k & 1;
                if (k >= 2)
                    st_wait(&empty[st], (uint32_t)(((k >> 1) - 1) & 1));
                int s = k + slab0;
                if (s >= a.nslab)
                    s -= a.nslab;
                const int64_t c0 = (int64_t)s * a.S;
                const int n = (int)min((int64_t)a.S, (int64_t)a.ncols - c0);
                const uint32_t bytes = (uint32_t)n * (uint32_t)sizeof(XT), b16 = bytes & ~15u;
                XT *dst = reinterpret_cast<XT *>(xbuf + st * xstride);
                for (int i = (int)(b16 / sizeof(XT)); i < n; i++)   // < 16 bytes that a bulk copy cannot move
                    dst[i] = x[c0 + i];
                if (b16) {
                    st_expect(&full[st], b16);
                    for (uint32_t o = 0; o < b16; o += 16384)
                        st_bulk(reinterpret_cast<unsigned char *>(dst) + o, reinterpret_cast<const unsigned char *>(x + c0) + o,
                                min(16384u, b16 - o), &full[st]);
                } else {
                    st_arrive(&full[st]);
                }
            }
        }
        return;
    }

    // ---------------- consumers: warp `warp` owns bin (blockIdx.x, warp)
    const int64_t bin = (int64_t)blockIdx.x * a.NW + warp;
    double *acc_w = acc + (size_t)warp * a.P;
    for (int i = lane; i < a.P; i += 32)
        acc_w[i] = 0.0;
    __syncwarp();
    const int64_t base = a.binbase[bin];
    const uint32_t *idx = a.idx + base;
    const VT *val = reinterpret_cast<const VT *>(a.val) + base;
    const uint32_t *cst = a.cstart + bin * (a.nslab + 1);
    // bounds of the first cell; the next cell's bounds are fetched one slab ahead
    uint32_t nx0 = cst[slab0], nx1 = cst[slab0 + 1];
    for (int k = 0; k < a.nslab; k++) {
        const uint32_t c0 = nx0, c1 = nx1;
        {
            int sn = k + 1 + slab0;
            if (sn >= a.nslab)
                sn -= a.nslab;   // (k + 1 == nslab wraps to slab0: a harmless extra load)
            nx0 = cst[sn];
            nx1 = cst[sn + 1];
        }
        const int st = k & 1;
        st_wait(&full[st], (uint32_t)((k >> 1) & 1));
        const XT *xs = reinterpret_cast<const XT *>(xbuf + st * xstride);
        uint32_t carry_row = ST_NOROW;
        double carry_sum = 0.0;
        for (uint32_t b = c0; b < c1; b += ST_BLK) {
            const uint32_t p = b + lane * ST_E;
            const bool lv = p < c1;
            uint32_t w[ST_E] = {0, 0, 0, 0, 0, 0, 0, 0};
            StVals<VT> vv;
            if (lv) {
                const uint4 q0 = ld_stream_u4(idx + p), q1 = ld_stream_u4(idx + p + 4);
                w[0] = q0.x, w[1] = q0.y, w[2] = q0.z, w[3] = q0.w, w[4] = q1.x, w[5] = q1.y, w[6] = q1.z, w[7] = q1.w;
                vv.load(val + p);
            }
            st_block<VT, XT>(w, vv, lv, xs, acc_w, lane, carry_row, carry_sum);
        }
        if (lane == 0 && carry_row != ST_NOROW)
            acc_w[carry_row] += carry_sum;
        __syncwarp();
        if (lane == 0)
            st_arrive(&empty[st]);   // this warp is done with the slab
    }
    // ---------------- results: rows straight to y, pieces of split rows to their carry slots
    const int32_t *rm = a.rowmap + bin * a.P;
    for (int i = lane; i < a.P; i += 32) {
        const int32_t r = rm[i];
        if (r >= 0)
            store_y<MULTI>(y, r, acc_w[i], true);
        else if (r <= -2)
            carry[-(r + 2)] = acc_w[i];
    }
}

// one thread per split row: add its pieces in piece order (deterministic)
template <bool MULTI>
__global__ void k_stream_fixup(const int32_t *__restrict__ split, int n_split, const double *__restrict__ carry, YOut y)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_split)
        return;
    const int32_t row = split[3 * k], q0 = split[3 * k + 1], nq = split[3 * k + 2];
    double tot = 0.0;
    for (int j = 0; j < nq; j++)
        tot += carry[q0 + j];
    store_y<MULTI>(y, row, tot, true);
}

template <typename VT, typename XT, bool MULTI>
static int stream_launch(StreamPlan *P, const StArgs &a, const void *d_x, const YOut &y, double *carry, cudaStream_t s)
{
    auto k = k_spmv_stream<VT, XT, MULTI>;
    static size_t optin = 0;   // per instantiation
    if (optin < P->smem_bytes) {
        CSRK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx().smem_optin));
        optin = ctx().smem_optin;
    }
    CSRK_LAUNCH(k, (unsigned)P->G, (unsigned)(P->NW + 1) * 32, P->smem_bytes, s, a, (const XT *)d_x, y, carry);
    if (P->n_split)
        CSRK_LAUNCH((k_stream_fixup<MULTI>), (unsigned)div_up(P->n_split, 128), 128, 0, s, P->split, P->n_split, carry, y);
    return CSRK_OK;
}

template <typename VT, typename XT>
static int stream_launch_m(StreamPlan *P, const StArgs &a, const void *d_x, const YOut &y, double *carry, cudaStream_t s)
{
    if (y.n > 1)
        return stream_launch<VT, XT, true>(P, a, d_x, y, carry, s);
    return stream_launch<VT, XT, false>(P, a, d_x, y, carry, s);
}

template <typename VT>
static int stream_launch_x(StreamPlan *P, const StArgs &a, const void *d_x, const YOut &y, double *carry, cudaStream_t s)
{
    if (P->x_kind == 4)
        return stream_launch_m<VT, float>(P, a, d_x, y, carry, s);
    return stream_launch_m<VT, double>(P, a, d_x, y, carry, s);
}

int stream_run(csrk_matrix *h, StreamPlan *P, const void *d_x, const YOut &y, cudaStream_t s)
{
    StArgs a;
    a.idx = P->idx;
    a.val = P->val;
    a.binbase = P->binbase;
    a.cstart = P->cstart;
    a.rowmap = P->rowmap;
    a.nslab = P->nslab;
    a.S = P->S;
    a.P = P->P;
    a.NW = P->NW;
    a.slab_bytes = P->slab_bytes;
    a.ncols = h->ncols;
    // carry slots of the split rows: per call (concurrent calls on one handle must not share them)
    DevBuf carry;
    if (P->n_split)
        CSRK_TRY(carry.alloc(sizeof(double) * (size_t)P->Q, s));
    switch (h->val_kind) {
    case 4: return stream_launch_x<float>(P, a, d_x, y, carry.as<double>(), s);
    case 8: return stream_launch_x<double>(P, a, d_x, y, carry.as<double>(), s);
    default: return stream_launch_x<NoVal>(P, a, d_x, y, carry.as<double>(), s);
    }
}

void stream_info(const StreamPlan *P, int64_t *out /*[8]*/)
{
    out[0] = P->G, out[1] = P->NW, out[2] = P->nslab, out[3] = P->S, out[4] = P->P, out[5] = P->Q, out[6] = P->n_split,
    out[7] = (int64_t)P->smem_bytes;
}

}  // namespace csrk
