// spmv_slab.cu -- mult_vec (csr/kernels/numba/__init__.py:55-67) for matrices whose x does not fit L1:
// the slab kernel with run-packed entry streams.
//
// Why.  The CSR tile kernel (spmv.cu) gathers x through L1/L2: with columns spread over an x of several
// MB every 4-byte gather costs one 32-byte L2 sector, 5x the bytes of the (colind, value) stream itself,
// and the kernel sits on the L2->SM fabric at a third of the HBM roofline (profiles/r01_spmv_tile_ncu.md).
// Here x is staged in SHARED memory, slab by slab, and the entries are re-laid out ONCE per handle:
//
//   * x is cut into slabs of S columns; one persistent CTA per SM keeps two slabs in shared memory, a
//     producer warp fetching the next slab with cp.async.bulk (TMA) + mbarriers while the consumer warps
//     work on the current one.  The gathers become LDS.
//   * rows longer than `piece` entries are cut into interleaved pieces (piece j = entries j, j+n, j+2n, ...
//     so that every piece spans the row's whole column range); the pieces ("pseudo-rows") are sorted by
//     length and dealt to G*NW bins (one bin per consumer WARP) in snake order: every warp owns the same
//     number of pseudo-rows with the same mix of lengths.  No dynamic scheduling, no inter-warp
//     communication, no CTA-wide barrier in the whole kernel; every pseudo-row has one float64 accumulator
//     in its warp's private part of shared memory.
//   * a warp's entries in one slab form a CELL.  Inside a cell the entries of one pseudo-row are a RUN.
//     Runs of two or more entries are sorted by length and packed 32 at a time into J GROUPS: lane l owns
//     run l, step t of the group holds entry t of every run that is longer than t, stored contiguously
//     (jagged-diagonal order).  A lane sums its run in a register -- no row tracking, no cross-lane
//     reduction -- and adds it to the accumulator with one plain read-modify-write (the 32 runs of a
//     group belong to 32 different pseudo-rows).  Groups are cut every 8 steps so that a group is at most
//     a few KB.  Runs of exactly one entry go, 32 at a time, into L GROUPS: one read-modify-write per
//     entry, again 32 distinct rows per instruction.  No atomics anywhere: results are deterministic.
//   * J entries cost 2 + V bytes (16-bit column inside the slab, value; the row lives in the group header),
//     L entries 4 + V (row << 16 | column): the stream is smaller than the CSR arrays it replaces.
//   * every warp's stream is ONE contiguous byte sequence in the order the warp consumes it (cell header,
//     J groups, L groups; cells in the CTA's slab walk order), and the warp itself prefetches it into a
//     private shared-memory ring with cp.async.bulk, several chunks ahead: HBM latency is hidden by the
//     ring, not by occupancy, and the loop never waits for a global load.
//   * pieces of split rows go to a carry array and a tiny fix-up kernel adds them in piece order.
//
// Cost model: HBM bytes are the stream + x once + y; the L2->SM fabric carries the stream
// plus G copies of x, which is why auto mode only picks this kernel when G*ncols*X is below 1.6 x the stream size.
#include <type_traits>
#include <algorithm>
#include <utility>

#include "expand.cuh"
#include "radix.cuh"
#include "spmv.cuh"

namespace csrk {

constexpr int SL_TB = 8;                   // steps per J group
constexpr int SL_NP = SL_TB / 2;           // pair-steps per J group: a lane reads two entries of its run at a time
constexpr int SL_HDR = 144;                // J group header: 32 x (row << 16 | pair-steps << 8 | my pairs), 8 x u16 byte offsets
constexpr int SL_NST_MAX = 4;              // chunks per ring: 2 or 4
constexpr int SL_MAX_WARPS = 31;           // consumer warps (+1 producer warp = 1024 threads)

struct StreamPlan {
    int G = 0, NW = 0, nslab = 0, S = 0, P = 0, x_kind = 0, piece = 0, ring = 0, nst = 4, nxb = 2;
    int slab_bytes = 0;
    size_t smem_bytes = 0;
    int64_t Q = 0, stream_bytes = 0;
    int n_split = 0;
    unsigned char *stream = nullptr;  // all bins' byte streams, bin after bin
    int64_t *binbase = nullptr;       // [G*NW] start of bin b's stream
    uint32_t *binlen = nullptr;       // [G*NW] its length (bytes, multiple of 16)
    int32_t *rowmap = nullptr;        // [G*NW][P]: >= 0 row of y, -1 unused, <= -2 carry slot -(v+2)
    int32_t *split = nullptr;         // [3*n_split]: row, first carry slot, number of pieces
};

void stream_destroy(StreamPlan *p, cudaStream_t s)
{
    if (!p)
        return;
    dev_free(p->stream, s);
    dev_free(p->binbase, s);
    dev_free(p->binlen, s);
    dev_free(p->rowmap, s);
    dev_free(p->split, s);
    delete p;
}

__host__ __device__ __forceinline__ uint32_t sl_pad16(uint32_t b) { return (b + 15u) & ~15u; }

// ------------------------------------------------------------------ plan builder
template <typename RPT> struct SlPieceLoader {
    const RPT *rp;
    int piece;
    __device__ __forceinline__ int64_t operator()(int64_t r) const
    {
        const int64_t len = (int64_t)rp[r + 1] - (int64_t)rp[r];
        return len > piece ? (len + piece - 1) / piece : 1;
    }
};

// one thread per row: sort key (piece - length: longest first), id and y destination of each piece
template <typename RPT>
__global__ void k_sl_pieces(const RPT *__restrict__ rp, int32_t nrows, int piece, const int64_t *__restrict__ qbase,
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
        qkey[q0 + j] = (int32_t)(piece - l);
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
__global__ void k_sl_deal(const int32_t *__restrict__ order, const int32_t *__restrict__ qdest, int64_t Q, int B, int P,
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

__global__ void k_sl_fill_i32(int32_t *p, int64_t n, int32_t v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        p[i] = v;
}

// Sort input, in PIECE-MAJOR order: position e' of row r's range is entry i of piece j (the pieces of a row
// one after the other), which is entry t = i*nq + j of the row.  Piece-major order keeps the entries of a
// pseudo-row adjacent inside every cell after the stable sort, whatever the column order inside the row.
// key = bin*nslab + (position of the slab in the CTA's walk), packed = local row << 16 | column inside the
// slab, pval = the entry's value.
template <typename RPT, typename VT>
__global__ void k_sl_keys(const RPT *__restrict__ rp, const int32_t *__restrict__ ci, const VT *__restrict__ vs,
                          const int32_t *__restrict__ rows, int64_t nnz, const int64_t *__restrict__ qbase,
                          const int32_t *__restrict__ qbl, int S, int nslab, int G, int NW, int32_t *__restrict__ key,
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
    const int bin = bl >> 16;
    int walk = slab - (int)(((int64_t)(bin / NW) * nslab) / G);   // the CTA starts its walk at slab g*nslab/G
    if (walk < 0)
        walk += nslab;
    key[ep] = bin * nslab + walk;
    packed[ep] = (int32_t)(((uint32_t)(bl & 0xffff) << 16) | (uint32_t)(c - slab * S));
    if constexpr (!std::is_same<VT, NoPayload>::value)
        pval[ep] = vs[e];
}

// 1 where a run starts: first entry, new cell, or new pseudo-row
struct SlRunFlag {
    const int32_t *k, *p;
    __device__ __forceinline__ int32_t operator()(int64_t i) const
    {
        return (i == 0 || k[i] != k[i - 1] || ((uint32_t)p[i] >> 16) != ((uint32_t)p[i - 1] >> 16)) ? 1 : 0;
    }
};

__global__ void k_sl_run_starts(SlRunFlag f, const int32_t *__restrict__ ex, int64_t nnz, int32_t *__restrict__ rs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz)
        return;
    if (i == nnz)
        rs[ex[nnz]] = (int32_t)nnz;
    else if (f(i))
        rs[ex[i]] = (int32_t)i;
}

// run -> sort key: cell, then longest first (runs of one entry end up last in their cell)
__global__ void k_sl_run_keys(const int32_t *__restrict__ rs, const int32_t *__restrict__ skeys, int32_t NR, int LB,
                              int32_t *__restrict__ rkey, int32_t *__restrict__ rid)
{
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= NR)
        return;
    const int32_t len = rs[r + 1] - rs[r];
    const int32_t lmax = (1 << LB) - 1;
    rkey[r] = (skeys[rs[r]] << LB) | (lmax - len);
    rid[r] = r;
}

// sorted runs -> (cell*2 + (len == 1)) for the segment bounds, and the length itself
__global__ void k_sl_run_class(const int32_t *__restrict__ srkey, int32_t NR, int LB, int32_t *__restrict__ k3,
                               int32_t *__restrict__ slen)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NR)
        return;
    const int32_t lmax = (1 << LB) - 1;
    const int32_t len = lmax - (srkey[j] & lmax);
    slen[j] = len;
    k3[j] = ((srkey[j] >> LB) << 1) | (len == 1 ? 1 : 0);
}

// items of a cell: 1 header + J groups (32 runs each, all their steps) + L groups (32 one-entry runs each)
struct SlItemCount {
    const int32_t *b;   // [2*ncells+1]
    __device__ __forceinline__ int64_t operator()(int64_t c) const
    {
        const int32_t nj = b[2 * c + 1] - b[2 * c], nl = b[2 * c + 2] - b[2 * c + 1];
        return 1 + (nj + 31) / 32 + (nl + 31) / 32;
    }
};

struct SlItem {
    int64_t cell;
    int kind;        // 0 header, 1 J group, 2 L group
    int32_t j0, j1;  // sorted-run range of the group
};
__device__ __forceinline__ SlItem sl_decode(int64_t item, const int64_t *__restrict__ itembase, int64_t ncells,
                                            const int32_t *__restrict__ b)
{
    int64_t lo = 0, hi = ncells;   // largest c with itembase[c] <= item
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (itembase[mid] <= item)
            lo = mid;
        else
            hi = mid;
    }
    SlItem it;
    it.cell = lo;
    const int local = (int)(item - itembase[lo]);
    const int32_t bj = b[2 * lo], bl = b[2 * lo + 1], be = b[2 * lo + 2];
    const int njg = (bl - bj + 31) / 32;
    if (local == 0) {
        it.kind = 0;
        it.j0 = it.j1 = 0;
    } else if (local <= njg) {
        it.kind = 1;
        it.j0 = bj + 32 * (local - 1);
        it.j1 = min(it.j0 + 32, bl);
    } else {
        it.kind = 2;
        it.j0 = bl + 32 * (local - 1 - njg);
        it.j1 = min(it.j0 + 32, be);
    }
    return it;
}

// one warp per item: its size in the stream; J groups also count their 8-step sub-groups into the cell
template <int VB>
__global__ void __launch_bounds__(256)
k_sl_item_bytes(int64_t NI, const int64_t *__restrict__ itembase, int64_t ncells, const int32_t *__restrict__ b,
                const int32_t *__restrict__ slen, uint32_t *__restrict__ itembytes, int32_t *__restrict__ cellsub)
{
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= NI)
        return;
    const SlItem it = sl_decode(item, itembase, ncells, b);
    uint32_t bytes;
    if (it.kind == 0) {
        bytes = 16;
    } else if (it.kind == 2) {
        bytes = 128 + 32 * VB;
    } else {
        const int32_t len = it.j0 + lane < it.j1 ? slen[it.j0 + lane] : 0;
        const int32_t T = __shfl_sync(0xffffffffu, len, 0);
        bytes = 0;
        int nsub = 0;
        for (int t0 = 0; t0 < T; t0 += SL_TB, nsub++) {
            const int n = 2 * warp_sum((min(max(len - t0, 0), SL_TB) + 1) >> 1);   // odd runs end in a null entry
            bytes += SL_HDR + sl_pad16(2u * n) + sl_pad16((uint32_t)VB * n);
        }
        if (lane == 0)
            atomicAdd(&cellsub[it.cell], nsub);
    }
    if (lane == 0)
        itembytes[item] = bytes;
}

// one warp per item: write it
template <typename VT>
__global__ void __launch_bounds__(256)
k_sl_fill(int64_t NI, const int64_t *__restrict__ itembase, int64_t ncells, const int32_t *__restrict__ b,
          const int32_t *__restrict__ slen, const int32_t *__restrict__ srid, const int32_t *__restrict__ rs,
          const int32_t *__restrict__ spacked, const VT *__restrict__ svals, const int64_t *__restrict__ itemoff,
          const int32_t *__restrict__ cellsub, unsigned char *__restrict__ stream, int P, int S)
{
    constexpr int VB = std::is_same<VT, NoPayload>::value ? 0 : (int)sizeof(VT);
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= NI)
        return;
    const SlItem it = sl_decode(item, itembase, ncells, b);
    unsigned char *o = stream + itemoff[item];
    if (it.kind == 0) {
        if (lane == 0) {
            const int32_t bj = b[2 * it.cell], bl = b[2 * it.cell + 1], be = b[2 * it.cell + 2];
            (void)bj;
            *reinterpret_cast<uint4 *>(o) = make_uint4((uint32_t)cellsub[it.cell], (uint32_t)((be - bl + 31) / 32), 0u, 0u);
        }
        return;
    }
    const bool have = it.j0 + lane < it.j1;
    const int32_t start = have ? rs[srid[it.j0 + lane]] : 0;
    if (it.kind == 2) {
        // one-entry runs; unused lanes are inert: dummy accumulator P, the always-zero x slot S, value 0
        reinterpret_cast<uint32_t *>(o)[lane] = have ? (uint32_t)spacked[start] : ((uint32_t)P << 16 | (uint32_t)S);
        if constexpr (VB > 0)
            reinterpret_cast<VT *>(o + 128)[lane] = have ? svals[start] : VT(0);
        return;
    }
    const int32_t len = have ? slen[it.j0 + lane] : 0;
    const uint32_t row = have ? ((uint32_t)spacked[start] >> 16) : (uint32_t)P;
    const int32_t T = __shfl_sync(0xffffffffu, len, 0);
    for (int t0 = 0; t0 < T; t0 += SL_TB) {
        const int lk = min(max(len - t0, 0), SL_TB);       // my entries in this group
        const int pk = (lk + 1) >> 1;                      // ... as pairs; an odd run ends in a null entry
        const int Tp = (min(T - t0, SL_TB) + 1) >> 1;      // pair-steps of the group
        const int n = 2 * warp_sum(pk);
        reinterpret_cast<uint32_t *>(o)[lane] = row << 16 | (uint32_t)Tp << 8 | (uint32_t)pk;
        // byte offsets of pair-steps 1..3 inside the column section, n, the same inside the value section, 0
        uint16_t *offs = reinterpret_cast<uint16_t *>(o + 128);
        uint32_t *cols = reinterpret_cast<uint32_t *>(o + SL_HDR);     // two 16-bit columns per pair
        unsigned char *vals = o + SL_HDR + sl_pad16(2u * n);
        int off = 0;   // pairs before this pair-step
        for (int t = 0; t < SL_NP; t++) {
            const bool act = t < pk;
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (act) {
                const int32_t src = start + t0 + 2 * t;
                const bool two = 2 * t + 1 < lk;
                const uint32_t c0 = (uint32_t)spacked[src] & 0xffffu;
                const uint32_t c1 = two ? ((uint32_t)spacked[src + 1] & 0xffffu) : (uint32_t)S;   // null: the zero x slot
                cols[off + lane] = c0 | c1 << 16;
                if constexpr (VB > 0) {
                    reinterpret_cast<VT *>(vals)[2 * (off + lane)] = svals[src];
                    reinterpret_cast<VT *>(vals)[2 * (off + lane) + 1] = two ? svals[src + 1] : VT(0);
                }
            }
            off += __popc(bal);
            if (lane == 0) {
                offs[t] = (uint16_t)(t < SL_NP - 1 ? 4 * off : n);            // offs[3] = n
                offs[4 + t] = (uint16_t)(t < SL_NP - 1 ? 2 * VB * off : 0);
            }
        }
        o += SL_HDR + sl_pad16(2u * n) + sl_pad16((uint32_t)VB * n);
    }
}

__global__ void k_sl_bins(const int64_t *__restrict__ itembase, const int64_t *__restrict__ itemoff, int B, int nslab,
                          int64_t *__restrict__ binbase, uint32_t *__restrict__ binlen, int ring, int *__restrict__ too_long)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B)
        return;
    const int64_t o0 = itemoff[itembase[(int64_t)b * nslab]], o1 = itemoff[itembase[(int64_t)(b + 1) * nslab]];
    binbase[b] = o0;
    binlen[b] = (uint32_t)(o1 - o0);
    if (o1 - o0 + ring >= ((int64_t)1 << 31))
        *too_long = 1;   // positions inside a warp's stream are 32-bit
}

static int sl_bits(int64_t n)
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
    constexpr int VB = HASV ? (int)sizeof(VT) : 0;
    const RPT *rp = (const RPT *)h->rp;
    const int32_t nrows = h->nrows;
    const int64_t nnz = h->nnz;
    const int B = P->G * P->NW;
    if (nnz >= ((int64_t)1 << 31) - 64)
    {
        set_error("slab plan: nnz %lld needs 64-bit run ids", (long long)nnz);
        return CSRK_EOVERFLOW;   // run starts and run ids are 32-bit
    }

    // 1. pseudo-rows
    DevBuf qbase;
    CSRK_TRY(qbase.alloc(sizeof(int64_t) * ((size_t)nrows + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(SlPieceLoader<RPT>{rp, P->piece}, (int64_t)nrows, qbase.as<int64_t>(), s)));
    int64_t Q = 0;
    CSRK_CUDA(cudaMemcpyAsync(&Q, qbase.as<int64_t>() + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    P->Q = Q;
    P->P = (int)std::max<int64_t>(div_up(Q, B), 1);
    // shared memory: mbarriers + alignment of the rings (up to one ring) + NW rings + NW*(P+1) float64
    // accumulators + nxb x slabs
    const size_t acc_bytes = (size_t)P->NW * (P->P + 1) * 8;
    const size_t ring_bytes = (size_t)P->NW * P->ring;
    const size_t bar_bytes = (size_t)(2 * 3 + P->NW * SL_NST_MAX) * 8 + 48 + (size_t)P->ring;
    const size_t smem_max = ctx().smem_optin;
    if (P->P > 65534 || Q >= ((int64_t)1 << 31) || acc_bytes + ring_bytes + bar_bytes + (size_t)P->nxb * 4096 > smem_max)
    {
        set_error("slab plan: %lld pseudo-rows per warp do not fit shared memory", (long long)P->P);
        return CSRK_EOVERFLOW;  // too many rows for shared-memory accumulators: stay on the tile kernel
    }
    // a buffer is one slab + 128 bytes (the always-zero slot the inert entries point at)
    int64_t slab = (int64_t)((smem_max - acc_bytes - ring_bytes - bar_bytes) / P->nxb - 128) & ~(int64_t)127;
    slab = std::min<int64_t>(slab, (int64_t)65408 * P->x_kind);   // columns inside a slab (and the zero slot) fit 16 bits
    slab = std::min<int64_t>(slab, (((int64_t)h->ncols * P->x_kind) + 127) & ~(int64_t)127);
    const int64_t cap = options().stream_slab_bytes.load();
    if (cap > 0)
        slab = std::min<int64_t>(slab, std::max<int64_t>(cap & ~(int64_t)127, 128));
    slab = std::max<int64_t>(slab, 128);
    P->slab_bytes = (int)slab;
    P->S = (int)(slab / P->x_kind);
    P->nslab = (int)std::max<int64_t>(div_up((int64_t)h->ncols, P->S), 1);
    P->smem_bytes = ring_bytes + (size_t)P->nxb * ((size_t)slab + 128) + acc_bytes + bar_bytes;
    const int64_t ncells = (int64_t)B * P->nslab;
    const int LB = sl_bits((int64_t)P->piece + 2);   // run lengths 1..piece (duplicate columns may exceed S)
    if (ncells >= ((int64_t)1 << (30 - LB)))
    {
        set_error("slab plan: %lld cells x %d length bits exceed the 31-bit run keys", (long long)ncells, LB);
        return CSRK_EOVERFLOW;   // (cell, length) run keys are 31-bit
    }

    DevBuf qkey, qid, qdest, order, qbl, splitcnt;
    CSRK_TRY(qkey.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(qid.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(qdest.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(order.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(qbl.alloc(sizeof(int32_t) * (size_t)Q, s));
    CSRK_TRY(splitcnt.alloc_zero(sizeof(int), s));
    const int64_t max_split = nnz / P->piece + 1;
    CSRK_TRY(dev_alloc((void **)&P->split, sizeof(int32_t) * 3 * (size_t)max_split, s));
    CSRK_LAUNCH((k_sl_pieces<RPT>), (unsigned)div_up((int64_t)nrows, 256), 256, 0, s, rp, nrows, P->piece,
                qbase.as<int64_t>(), qkey.as<int32_t>(), qid.as<int32_t>(), qdest.as<int32_t>(), P->split,
                splitcnt.as<int>());
    // 2. longest first (stable), dealt to the bins in snake order
    CSRK_TRY((radix_sort_by_key<NoPayload>(qkey.as<int32_t>(), qid.as<int32_t>(), (const NoPayload *)nullptr, Q,
                                           sl_bits((int64_t)P->piece + 1), order.as<int32_t>(), (NoPayload *)nullptr, s)));
    CSRK_TRY(dev_alloc((void **)&P->rowmap, sizeof(int32_t) * (size_t)B * P->P, s));
    CSRK_LAUNCH(k_sl_fill_i32, (unsigned)div_up((int64_t)B * P->P, 256), 256, 0, s, P->rowmap, (int64_t)B * P->P, -1);
    CSRK_LAUNCH(k_sl_deal, (unsigned)div_up(Q, 256), 256, 0, s, order.as<int32_t>(), qdest.as<int32_t>(), Q, B, P->P,
                qbl.as<int32_t>(), P->rowmap);
    CSRK_TRACE_MARK("slab plan: pieces dealt", s);

    // 3. entries: key = (bin, slab in walk order), stable sort keeps the runs together
    DevBuf rows, key, packed, pval, skeys, spacked, svals;
    CSRK_TRY(rows.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(key.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(packed.alloc(sizeof(int32_t) * (size_t)nnz, s));
    if (HASV)
        CSRK_TRY(pval.alloc(sizeof(VT) * (size_t)nnz, s));
    CSRK_LAUNCH((k_expand_rows<RPT>), (unsigned)div_up(nnz, EXP_TILE), 256, 0, s, rp, nrows, nnz, rows.as<int32_t>());
    CSRK_LAUNCH((k_sl_keys<RPT, VT>), (unsigned)div_up(nnz, 256), 256, 0, s, rp, h->ci, (const VT *)h->vs,
                rows.as<int32_t>(), nnz, qbase.as<int64_t>(), qbl.as<int32_t>(), P->S, P->nslab, P->G, P->NW,
                key.as<int32_t>(), packed.as<int32_t>(), HASV ? pval.as<VT>() : nullptr);
    CSRK_TRY(skeys.alloc(sizeof(int32_t) * (size_t)nnz, s));
    CSRK_TRY(spacked.alloc(sizeof(int32_t) * (size_t)nnz, s));
    if (HASV)
        CSRK_TRY(svals.alloc(sizeof(VT) * (size_t)nnz, s));
    CSRK_TRY((radix_sort_by_key<VT>(key.as<int32_t>(), packed.as<int32_t>(), HASV ? pval.as<VT>() : nullptr, nnz,
                                    sl_bits(ncells), spacked.as<int32_t>(), HASV ? svals.as<VT>() : nullptr, s,
                                    skeys.as<int32_t>())));
    CSRK_TRACE_MARK("slab plan: entries sorted", s);

    // 4. runs: starts, lengths, sorted by (cell, longest first)
    DevBuf ex, rs;
    CSRK_TRY(ex.alloc(sizeof(int32_t) * ((size_t)nnz + 1), s));
    const SlRunFlag flag{skeys.as<int32_t>(), spacked.as<int32_t>()};
    CSRK_TRY((exclusive_scan<int32_t>(flag, nnz, ex.as<int32_t>(), s)));
    int32_t NR = 0;
    CSRK_CUDA(cudaMemcpyAsync(&NR, ex.as<int32_t>() + nnz, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    CSRK_TRY(rs.alloc(sizeof(int32_t) * ((size_t)NR + 1), s));
    CSRK_LAUNCH(k_sl_run_starts, (unsigned)div_up(nnz + 1, 256), 256, 0, s, flag, ex.as<int32_t>(), nnz, rs.as<int32_t>());
    DevBuf rkey, rid, srkey, srid, k3, slen, bnd;
    CSRK_TRY(rkey.alloc(sizeof(int32_t) * (size_t)NR, s));
    CSRK_TRY(rid.alloc(sizeof(int32_t) * (size_t)NR, s));
    CSRK_TRY(srkey.alloc(sizeof(int32_t) * (size_t)NR, s));
    CSRK_TRY(srid.alloc(sizeof(int32_t) * (size_t)NR, s));
    CSRK_LAUNCH(k_sl_run_keys, (unsigned)div_up((int64_t)NR, 256), 256, 0, s, rs.as<int32_t>(), skeys.as<int32_t>(), NR, LB,
                rkey.as<int32_t>(), rid.as<int32_t>());
    CSRK_TRY((radix_sort_by_key<NoPayload>(rkey.as<int32_t>(), rid.as<int32_t>(), (const NoPayload *)nullptr, (int64_t)NR,
                                           sl_bits(ncells) + LB, srid.as<int32_t>(), (NoPayload *)nullptr, s,
                                           srkey.as<int32_t>())));
    CSRK_TRY(k3.alloc(sizeof(int32_t) * (size_t)NR, s));
    CSRK_TRY(slen.alloc(sizeof(int32_t) * (size_t)NR, s));
    CSRK_LAUNCH(k_sl_run_class, (unsigned)div_up((int64_t)NR, 256), 256, 0, s, srkey.as<int32_t>(), NR, LB, k3.as<int32_t>(),
                slen.as<int32_t>());
    CSRK_TRY(bnd.alloc(sizeof(int32_t) * (2 * (size_t)ncells + 1), s));
    CSRK_LAUNCH((k_key_bounds<int32_t>), (unsigned)div_up(div_up((int64_t)NR + 1, 4), 256), 256, 0, s, k3.as<int32_t>(),
                (int64_t)NR, (int32_t)(2 * ncells), bnd.as<int32_t>());
    CSRK_TRACE_MARK("slab plan: runs sorted", s);

    // 5. items (cell headers and groups): sizes, offsets, contents
    DevBuf itembase, itembytes, itemoff, cellsub;
    CSRK_TRY(itembase.alloc(sizeof(int64_t) * ((size_t)ncells + 1), s));
    CSRK_TRY((exclusive_scan<int64_t>(SlItemCount{bnd.as<int32_t>()}, ncells, itembase.as<int64_t>(), s)));
    int64_t NI = 0;
    CSRK_CUDA(cudaMemcpyAsync(&NI, itembase.as<int64_t>() + ncells, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    CSRK_TRY(itembytes.alloc(sizeof(uint32_t) * (size_t)NI, s));
    CSRK_TRY(itemoff.alloc(sizeof(int64_t) * ((size_t)NI + 1), s));
    CSRK_TRY(cellsub.alloc_zero(sizeof(int32_t) * (size_t)ncells, s));
    CSRK_LAUNCH((k_sl_item_bytes<VB>), (unsigned)div_up(NI * 32, 256), 256, 0, s, NI, itembase.as<int64_t>(), ncells,
                bnd.as<int32_t>(), slen.as<int32_t>(), itembytes.as<uint32_t>(), cellsub.as<int32_t>());
    CSRK_TRY((exclusive_scan<int64_t>(ArrayLoader<uint32_t>{itembytes.as<uint32_t>()}, NI, itemoff.as<int64_t>(), s)));
    int64_t total = 0;
    CSRK_CUDA(cudaMemcpyAsync(&total, itemoff.as<int64_t>() + NI, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    P->stream_bytes = total;
    // (the ring prefetch reads whole chunks: the last bin's final chunk may run past its stream)
    CSRK_TRY(dev_alloc((void **)&P->stream, (size_t)total + (size_t)P->ring, s));
    CSRK_TRY(dev_alloc((void **)&P->binbase, sizeof(int64_t) * (size_t)B, s));
    CSRK_TRY(dev_alloc((void **)&P->binlen, sizeof(uint32_t) * (size_t)B, s));
    CSRK_LAUNCH((k_sl_fill<VT>), (unsigned)div_up(NI * 32, 256), 256, 0, s, NI, itembase.as<int64_t>(), ncells,
                bnd.as<int32_t>(), slen.as<int32_t>(), srid.as<int32_t>(), rs.as<int32_t>(), spacked.as<int32_t>(),
                HASV ? svals.as<VT>() : nullptr, itemoff.as<int64_t>(), cellsub.as<int32_t>(), P->stream, P->P, P->S);
    DevBuf toolong;
    CSRK_TRY(toolong.alloc_zero(sizeof(int), s));
    CSRK_LAUNCH(k_sl_bins, (unsigned)div_up(B, 256), 256, 0, s, itembase.as<int64_t>(), itemoff.as<int64_t>(), B, P->nslab,
                P->binbase, P->binlen, P->ring, toolong.as<int>());
    int too_long = 0;
    CSRK_CUDA(cudaMemcpyAsync(&too_long, toolong.as<int>(), sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaMemcpyAsync(&P->n_split, splitcnt.as<int>(), sizeof(int), cudaMemcpyDeviceToHost, s));
    CSRK_CUDA(cudaStreamSynchronize(s));
    if (too_long)
    {
        set_error("slab plan: a warp's stream of %lld bytes needs 64-bit positions", (long long)(total / B));
        return CSRK_EOVERFLOW;   // positions inside a bin's stream are 32-bit
    }
    CSRK_TRACE_MARK("slab plan: streams written", s);
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
    P->NW = (int)std::min<int64_t>(std::max<int64_t>(nw, 1), SL_MAX_WARPS);
    P->piece = (int)std::min<int64_t>(std::max<int64_t>(options().stream_piece.load(), 8), 4096);
    P->ring = options().stream_ring_bytes.load() >= 8192 ? 8192 : 4096;
    P->nst = options().stream_ring_chunks.load() == 2 ? 2 : 4;
    P->nxb = options().stream_xbufs.load() == 3 ? 3 : 2;
    if (h->val_kind == 8)
        P->ring = 8192;   // a J group of float64 values is up to 2.7 KB and must fit three chunks
    P->NW = (int)std::min<int64_t>(P->NW, (int64_t)(ctx().smem_optin / 2) / P->ring);   // rings take at most half
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
__device__ __forceinline__ uint32_t sl_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sl_bar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sl_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sl_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sl_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sl_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sl_u32(bar)) : "memory");
}
// `backoff` > 0: sleep that many ns between polls (coarse waits: an x slab; a spinning warp steals issue slots)
__device__ __forceinline__ void sl_wait(uint64_t *bar, uint32_t parity, unsigned backoff = 0)
{
    uint32_t ok = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(sl_u32(bar)), "r"(parity)
                     : "memory");
        if (ok)
            break;
        if (backoff)
            __nanosleep(backoff);
    }
}
__device__ __forceinline__ void sl_bulk(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sl_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(sl_u32(bar))
                 : "memory");
}

template <typename VT, typename XT> __device__ __forceinline__ double sl_prod(VT v, XT xv)
{
    if constexpr (std::is_same<VT, NoVal>::value) {
        return (double)xv;
    } else {
        using PT = typename Prod<VT, XT>::type;   // numba's promotion: f4*f4 -> f4, else f8; the product is
        if constexpr (std::is_same<PT, float>::value)  // rounded before it is added (no FMA contraction)
            return (double)__fmul_rn((float)xv, (float)v);
        else
            return __dmul_rn((double)xv, (double)v);
    }
}

struct SlArgs {
    const unsigned char *stream;
    const int64_t *binbase;
    const uint32_t *binlen;
    const int32_t *rowmap;
    int nslab, S, P, NW, slab_bytes, ring, nst, nxb;
    int32_t ncols;
    unsigned long long *dbg;   // CSRK_SLAB_TIMING=1: per consumer warp, globaltimer at the end of its walk / of its epilogue
};

// shared-memory accesses by 32-bit shared address (the ring offset arithmetic below is done on addresses)
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
template <typename T> __device__ __forceinline__ T lds_val(uint32_t a);
template <> __device__ __forceinline__ float lds_val<float>(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
template <> __device__ __forceinline__ double lds_val<double>(uint32_t a) { return lds_f64(a); }

// A warp's view of its ring: bytes [0, avail) of the stream have landed, chunks [0, tail) were requested.
// The ring's shared-memory ADDRESS is a multiple of its size, so stream offset o lives at address
// base | (o & mask): one LOP3.
struct SlRing {
    uint32_t base;               // shared address of this warp's ring
    uint32_t full;               // shared address of its nst mbarriers
    const unsigned char *src;
    uint32_t len, mask, chunk, csh, nst, nsh;
    uint32_t avail, tail;
    int lane;

    __device__ __forceinline__ void request(uint32_t i)
    {
        const uint32_t st = i & (nst - 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * st), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         base + st * chunk),
                     "l"(src + (size_t)i * chunk), "r"(chunk), "r"(full + 8 * st)
                     : "memory");
    }
    __device__ __forceinline__ void start()
    {
        avail = 0;
        tail = nst;
        if (lane == 0)
            for (uint32_t i = 0; i < nst; i++)
                if (i * chunk < len)
                    request(i);
    }
    // make bytes [.., end) readable
    __device__ __forceinline__ void ensure(uint32_t end)
    {
#pragma unroll 1
        while (avail < end) {
            const uint32_t i = avail >> csh;
            const uint32_t bar = full + 8 * (i & (nst - 1)), parity = (i >> nsh) & 1u;
            uint32_t ok = 0;
            while (!ok)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(bar), "r"(parity)
                    : "memory");
            avail += chunk;
        }
    }
    // everything below `pos` has been consumed by all lanes (call after __syncwarp): refill the freed chunks
    __device__ __forceinline__ void release(uint32_t pos)
    {
        const uint32_t want = (pos >> csh) + nst;
        if (tail < want) {
            if (lane == 0) {
#pragma unroll 1
                for (uint32_t i = tail; i < want; i++)
                    if (i * chunk < len)
                        request(i);
            }
            tail = want;
        }
    }
    __device__ __forceinline__ uint32_t addr(uint32_t off) const { return base | (off & mask); }
};

// the term of one entry in the type numba computes it in (f4*f4 -> f4, else f8; no value: x itself), rounded
// before it is added (no FMA contraction)
template <typename VT, typename XT> struct SlTerm {
    using type = typename Prod<VT, XT>::type;
};
template <typename XT> struct SlTerm<NoVal, XT> {
    using type = XT;
};
template <typename VT, typename XT>
__device__ __forceinline__ typename SlTerm<VT, XT>::type sl_term(uint32_t vaddr, XT xv)
{
    using PT = typename SlTerm<VT, XT>::type;
    if constexpr (std::is_same<VT, NoVal>::value)
        return xv;
    else if constexpr (std::is_same<PT, float>::value)
        return __fmul_rn((float)xv, (float)lds_val<VT>(vaddr));
    else
        return __dmul_rn((double)xv, (double)lds_val<VT>(vaddr));
}
// v if keep else +0, chosen on the bits (what a masked lane computed is arbitrary, NaN included)
__device__ __forceinline__ float sl_keep(float v, bool keep) { return __uint_as_float(keep ? __float_as_uint(v) : 0u); }
__device__ __forceinline__ double sl_keep(double v, bool keep)
{
    return __longlong_as_double(keep ? __double_as_longlong(v) : 0ll);
}

// two values of a lane's pair
template <typename T> struct SlPair {
    T a, b;
};
template <typename T> __device__ __forceinline__ SlPair<T> lds_pair(uint32_t addr);
template <> __device__ __forceinline__ SlPair<float> lds_pair<float>(uint32_t addr)
{
    SlPair<float> p;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(p.a), "=f"(p.b) : "r"(addr));
    return p;
}
template <> __device__ __forceinline__ SlPair<double> lds_pair<double>(uint32_t addr)
{
    SlPair<double> p;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(p.a), "=d"(p.b) : "r"(addr));
    return p;
}
template <typename VT, typename XT>
__device__ __forceinline__ typename SlTerm<VT, XT>::type sl_mul(VT v, XT xv)
{
    using PT = typename SlTerm<VT, XT>::type;
    if constexpr (std::is_same<PT, float>::value)
        return __fmul_rn((float)xv, (float)v);
    else
        return __dmul_rn((double)xv, (double)v);
}

// pair-steps T0..T1-1 of a J group, fully predicated (no branches): the loads of all steps are independent.
// co[t] / vo[t] = byte offset of pair-step t inside the column / value section; lane l's pair is the l-th there.
// A lane past the end of its run reads the ZERO PAD instead (columns S = the always-zero x slot, values 0): its
// terms are exactly 0 whatever the other lanes' entries hold.  The null entry that ends an odd run is the same.
template <int T0, int T1, typename VT, typename XT>
__device__ __forceinline__ void sl_steps(const SlRing &R, const int pk, const uint32_t cl, const uint32_t vl,
                                         const uint32_t (&co)[SL_NP], const uint32_t (&vo)[SL_NP], const uint32_t xs,
                                         const uint32_t zpad, double &sum)
{
#pragma unroll
    for (int t = T0; t < T1; t++) {
        const bool act = t < pk;
        const uint32_t cc = lds_u32(act ? R.addr(cl + co[t]) : zpad);
        const XT x0 = lds_val<XT>(xs + (cc & 0xffffu) * (uint32_t)sizeof(XT));
        const XT x1 = lds_val<XT>(xs + (cc >> 16) * (uint32_t)sizeof(XT));
        if constexpr (std::is_same<VT, NoVal>::value) {
            sum += (double)x0;
            sum += (double)x1;
        } else {
            const SlPair<VT> v = lds_pair<VT>(act ? R.addr(vl + vo[t]) : zpad + 16u);
            sum += (double)sl_mul<VT, XT>(v.a, x0);
            sum += (double)sl_mul<VT, XT>(v.b, x1);
        }
    }
}

template <typename VT, typename XT, bool MULTI>
__global__ void __launch_bounds__(1024, 1)
k_spmv_slab(SlArgs a, const XT *__restrict__ x, YOut y, double *__restrict__ carry)
{
    constexpr uint32_t VB = std::is_same<VT, NoVal>::value ? 0u : (uint32_t)sizeof(VT);
    extern __shared__ __align__(1024) unsigned char sl_smem[];
    // layout: mbarriers | (pad to a multiple of the ring size in the shared ADDRESS space) rings[NW] | x[nxb] | acc[NW][P+1]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sl_smem);   // full[nxb], empty[nxb], ring[NW][nst], zero pad
    uint64_t *full = bars, *empty = bars + a.nxb, *rbar = bars + 2 * a.nxb;
    const uint32_t smem0 = sl_u32(sl_smem);
    // zero pad, 32 bytes at a 16-byte boundary: a column pair (S, S), then 16 zero bytes (a value pair)
    const uint32_t zpad = (smem0 + (uint32_t)(2 * a.nxb + a.NW * a.nst) * 8u + 15u) & ~15u;
    const uint32_t ring0 = (zpad + 32u + (uint32_t)a.ring - 1u) & ~((uint32_t)a.ring - 1u);
    const uint32_t xstride = (uint32_t)a.slab_bytes + 128u;                            // slab + the zero slot
    const uint32_t xbuf0 = ring0 + (uint32_t)a.NW * (uint32_t)a.ring;
    unsigned char *xbuf = sl_smem + (xbuf0 - smem0);
    const uint32_t acc0 = xbuf0 + (uint32_t)a.nxb * xstride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < a.nxb; i++) {
            sl_bar_init(&full[i], 1);
            sl_bar_init(&empty[i], a.NW);
        }
        for (int i = 0; i < a.NW * a.nst; i++)
            sl_bar_init(&rbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < a.nxb)
        reinterpret_cast<XT *>(xbuf + tid * xstride)[a.S] = XT(0);   // what the inert entries multiply
    if (tid == 32)
        *reinterpret_cast<uint4 *>(sl_smem + (zpad - smem0)) = make_uint4((uint32_t)a.S | (uint32_t)a.S << 16, 0u, 0u, 0u);
    if (tid == 33)
        *reinterpret_cast<uint4 *>(sl_smem + (zpad + 16u - smem0)) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();   // the only CTA-wide barrier
    // programmatic dependent launch: k_slab_fixup may be scheduled now (its CTAs fit beside this one and wait in
    // griddepcontrol.wait until this grid has completed and its carries are visible): no launch gap after the kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // CTA g walks the slabs starting at slab g*nslab/G and wraps around, so that at any moment the CTAs pull
    // different parts of x out of L2
    const int slab0 = (int)(((int64_t)blockIdx.x * a.nslab) / gridDim.x);

    if (warp == a.NW) {
        // ---------------- producer: the k-th slab of this CTA's walk into buffer k % nxb
        if (lane == 0) {
            int st = 0;
            uint32_t lap = 0;
            for (int k = 0; k < a.nslab; k++) {
                if (lap)
                    sl_wait(&empty[st], (lap - 1) & 1u, 1000);
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
                    sl_expect(&full[st], b16);
                    for (uint32_t o = 0; o < b16; o += 16384)
                        sl_bulk(reinterpret_cast<unsigned char *>(dst) + o, reinterpret_cast<const unsigned char *>(x + c0) + o,
                                min(16384u, b16 - o), &full[st]);
                } else {
                    sl_arrive(&full[st]);
                }
                if (++st == a.nxb) {
                    st = 0;
                    lap++;
                }
            }
        }
        return;
    }

    // ---------------- consumers: warp `warp` owns bin (blockIdx.x, warp)
    const int64_t bin = (int64_t)blockIdx.x * a.NW + warp;
    const uint32_t acc_w = acc0 + (uint32_t)warp * (uint32_t)(a.P + 1) * 8u;
    SlRing R;
    R.base = ring0 + (uint32_t)warp * (uint32_t)a.ring;
    R.nst = (uint32_t)a.nst;
    R.nsh = a.nst == 2 ? 1u : 2u;
    R.full = sl_u32(rbar + warp * a.nst);
    R.src = a.stream + a.binbase[bin];
    R.len = a.binlen[bin];
    R.mask = (uint32_t)a.ring - 1u;
    R.chunk = (uint32_t)a.ring / R.nst;
    R.csh = 31 - __clz((int)R.chunk);
    R.lane = lane;
    R.start();
    for (int i = lane; i <= a.P; i += 32)
        sts_f64(acc_w + 8u * i, 0.0);
    __syncwarp();
    uint32_t pos = 0;
    int st = 0;
    uint32_t lap = 0;
    const uint32_t inert = (uint32_t)a.P << 16 | (uint32_t)a.S;
    for (int k = 0; k < a.nslab; k++) {
        R.ensure(pos + 16);
        const uint2 hdr = lds_u32x2(R.addr(pos));   // J groups, 32-entry L blocks
        pos += 16;
        sl_wait(&full[st], lap & 1u, 250);
        const uint32_t xs = xbuf0 + (uint32_t)st * xstride;
        // J groups: lane l sums run l over the group's steps
#pragma unroll 1
        for (uint32_t j = 0; j < hdr.x; j++) {
            R.ensure(pos + SL_HDR);
            const uint32_t m = lds_u32(R.addr(pos + 4u * lane));   // row << 16 | pair-steps << 8 | my pairs
            const uint4 o = lds_u32x4(R.addr(pos + 128));
            const uint32_t co[SL_NP] = {0u, o.x & 0xffffu, o.x >> 16, o.y & 0xffffu};
            const uint32_t vo[SL_NP] = {0u, o.z & 0xffffu, o.z >> 16, o.w & 0xffffu};
            const uint32_t n = o.y >> 16;
            const uint32_t vsec = pos + SL_HDR + sl_pad16(2u * n);
            const uint32_t cl = pos + SL_HDR + 4u * lane, vl = vsec + 2u * VB * lane;
            const uint32_t end = vsec + sl_pad16(VB * n);
            R.ensure(end);
            const int lk = (int)(m & 0xffu);
            double sum = 0.0;
            sl_steps<0, 3, VT, XT>(R, lk, cl, vl, co, vo, xs, zpad, sum);
            if (((m >> 8) & 0xffu) > 3)
                sl_steps<3, SL_NP, VT, XT>(R, lk, cl, vl, co, vo, xs, zpad, sum);
            if (lk) {
                const uint32_t aa = acc_w + 8u * (m >> 16);
                sts_f64(aa, lds_f64(aa) + sum);
            }
            __syncwarp();
            pos = end;
            R.release(pos);
        }
        // L blocks: one entry per lane, all rows of a cell's L blocks distinct; up to four blocks at a time
#pragma unroll 1
        for (uint32_t j = 0; j < hdr.y; j += 4) {
            const uint32_t nb = min(hdr.y - j, 4u);
            const uint32_t end = pos + nb * (128 + 32 * VB);
            R.ensure(end);
            uint32_t w[4];
            typename SlTerm<VT, XT>::type p[4];
            double old[4];
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                // (blocks past nb: the ring offset is masked, so the read is harmless; its result is unused)
                const uint32_t bo = pos + i * (128 + 32 * VB);
                w[i] = lds_u32(R.addr(bo + 4u * lane));
                w[i] = i < nb ? w[i] : inert;
                p[i] = sl_term<VT, XT>(R.addr(bo + 128 + VB * lane), lds_val<XT>(xs + (w[i] & 0xffffu) * (uint32_t)sizeof(XT)));
            }
#pragma unroll
            for (uint32_t i = 0; i < 4; i++)
                old[i] = lds_f64(acc_w + 8u * (w[i] >> 16));
#pragma unroll
            for (uint32_t i = 0; i < 4; i++)
                if (i < nb)
                    sts_f64(acc_w + 8u * (w[i] >> 16), old[i] + (double)p[i]);
            __syncwarp();
            pos = end;
            R.release(pos);
        }
        if (lane == 0)
            sl_arrive(&empty[st]);   // this warp is done with the slab
        if (++st == a.nxb) {
            st = 0;
            lap++;
        }
    }
    if (a.dbg && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.dbg[2 * bin] = t;
    }
    // ---------------- results: rows straight to y, pieces of split rows to their carry slots
    // (eight row ids per lane are fetched together: one load latency per 256 rows instead of one per 32 -- the
    // dependent load -> store chain of the plain loop kept the last warps busy for 8 us after their walk)
    const int32_t *rm = a.rowmap + bin * a.P;
    for (int i0 = 0; i0 < a.P; i0 += 256) {
        int32_t r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = i0 + 32 * u + lane;
            r[u] = i < a.P ? ld_stream_i32(rm + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = i0 + 32 * u + lane;
            if (r[u] >= 0)
                store_y<MULTI>(y, r[u], lds_f64(acc_w + 8u * i), true);
            else if (r[u] <= -2)
                carry[-(r[u] + 2)] = lds_f64(acc_w + 8u * i);
        }
    }
    if (a.dbg && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.dbg[2 * bin + 1] = t;
    }
}

// one warp per split row: lanes add every 32nd piece, then a shuffle tree -- a fixed order: deterministic
template <bool MULTI>
__global__ void __launch_bounds__(256)
k_slab_fixup(const int32_t *__restrict__ split, int n_split, const double *__restrict__ carry, YOut y)
{
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int32_t row = 0, q0 = 0, nq = 0;
    if (k < n_split)   // (the plan's arrays do not depend on the slab kernel: fetched while it still runs)
        row = split[3 * k], q0 = split[3 * k + 1], nq = split[3 * k + 2];
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the slab kernel's carries (no-op without the launch attribute)
    if (k >= n_split)
        return;
    double tot = 0.0;
    for (int j = lane; j < nq; j += 32)
        tot += __ldcg(&carry[q0 + j]);   // written by the grid this one overlaps: from L2
    tot = warp_sum(tot);
    if (lane == 0)
        store_y<MULTI>(y, row, tot, true);
}

template <typename VT, typename XT, bool MULTI>
static int slab_launch(StreamPlan *P, const SlArgs &a, const void *d_x, const YOut &y, double *carry, cudaStream_t s)
{
    auto k = k_spmv_slab<VT, XT, MULTI>;
    static size_t optin = 0;   // per instantiation
    if (optin < P->smem_bytes) {
        CSRK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx().smem_optin));
        optin = ctx().smem_optin;
    }
    CSRK_LAUNCH(k, (unsigned)P->G, (unsigned)(P->NW + 1) * 32, P->smem_bytes, s, a, (const XT *)d_x, y, carry);
    if (P->n_split) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)div_up((int64_t)P->n_split * 32, 256));
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute at = {};
        at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at.val.programmaticStreamSerializationAllowed = 1;
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        (void)cudaStreamIsCapturing(s, &cap);   // (inside a graph capture -- the multi-GPU step -- a plain edge)
        cfg.attrs = &at;
        cfg.numAttrs = cap == cudaStreamCaptureStatusNone ? 1 : 0;
        const int32_t *split = P->split;
        const int n_split = P->n_split;
        const double *cr = carry;
        CSRK_CUDA(cudaLaunchKernelEx(&cfg, k_slab_fixup<MULTI>, split, n_split, cr, y));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return CSRK_OK;
}

// Rows leave the slab kernel at its very end and in bin order (scattered 8-byte stores): with several
// destinations (the gather buffers of peer GPUs / an NVLink multicast address) the kernel writes the local y
// only and the finished segment is copied out in one coalesced pass (measured on 2 B200: 8-byte multimem.st
// from the kernel's epilogue cost 0.23 ms, the copy 0.02 ms).
template <typename VT, typename XT>
static int slab_launch_m(StreamPlan *P, const SlArgs &a, const void *d_x, const YOut &y, double *carry, cudaStream_t s,
                         int32_t nrows)
{
    YOut local = y;
    local.n = 1;
    local.mc = 0;
    CSRK_TRY((slab_launch<VT, XT, false>(P, a, d_x, local, carry, s)));
    if (y.n > 1) {
        const size_t bytes = (size_t)nrows * 8;
        if (y.mc)
            return mc_broadcast_run(y.p[1], y.p[0], (int64_t)bytes, s);
        for (int k = 1; k < y.n; k++)
            CSRK_CUDA(cudaMemcpyAsync(y.p[k], y.p[0], bytes, cudaMemcpyDeviceToDevice, s));
    }
    return CSRK_OK;
}

template <typename VT>
static int slab_launch_x(StreamPlan *P, const SlArgs &a, const void *d_x, const YOut &y, double *carry, cudaStream_t s,
                         int32_t nrows)
{
    if (P->x_kind == 4)
        return slab_launch_m<VT, float>(P, a, d_x, y, carry, s, nrows);
    return slab_launch_m<VT, double>(P, a, d_x, y, carry, s, nrows);
}

int stream_run(csrk_matrix *h, StreamPlan *P, const void *d_x, const YOut &y, cudaStream_t s)
{
    SlArgs a;
    a.stream = P->stream;
    a.binbase = P->binbase;
    a.binlen = P->binlen;
    a.rowmap = P->rowmap;
    a.nslab = P->nslab;
    a.S = P->S;
    a.P = P->P;
    a.NW = P->NW;
    a.slab_bytes = P->slab_bytes;
    a.ring = P->ring;
    a.nst = P->nst;
    a.nxb = P->nxb;
    a.ncols = h->ncols;
    a.dbg = nullptr;
    static const bool timing = getenv("CSRK_SLAB_TIMING") != nullptr;
    DevBuf dbg;
    if (timing) {
        CSRK_TRY(dbg.alloc_zero(sizeof(unsigned long long) * 2 * (size_t)P->G * P->NW, s));
        a.dbg = dbg.as<unsigned long long>();
    }
    // carry slots of the split rows: per call (concurrent calls on one handle must not share them)
    DevBuf carry;
    if (P->n_split)
        CSRK_TRY(carry.alloc(sizeof(double) * (size_t)P->Q, s));
    int rc;
    switch (h->val_kind) {
    case 4: rc = slab_launch_x<float>(P, a, d_x, y, carry.as<double>(), s, h->nrows); break;
    case 8: rc = slab_launch_x<double>(P, a, d_x, y, carry.as<double>(), s, h->nrows); break;
    default: rc = slab_launch_x<NoVal>(P, a, d_x, y, carry.as<double>(), s, h->nrows); break;
    }
    if (timing && rc == CSRK_OK) {   // diagnostic: how evenly do the warps finish?
        std::vector<unsigned long long> t(2 * (size_t)P->G * P->NW);
        CSRK_CUDA(cudaMemcpyAsync(t.data(), a.dbg, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CSRK_CUDA(cudaStreamSynchronize(s));
        unsigned long long lo = ~0ull, hi = 0, hi2 = 0;
        for (size_t i = 0; i < t.size(); i += 2) {
            lo = std::min(lo, t[i]);
            hi = std::max(hi, t[i]);
            hi2 = std::max(hi2, t[i + 1]);
        }
        double sum = 0;
        for (size_t i = 0; i < t.size(); i += 2)
            sum += (double)(t[i] - lo);
        const double mean = sum / (t.size() / 2);
        unsigned long long cmin = ~0ull, cmax = 0;
        std::vector<std::pair<unsigned long long, int>> ce;
        for (int g = 0; g < P->G; g++) {
            unsigned long long c = 0;
            for (int w = 0; w < P->NW; w++)
                c = std::max(c, t[2 * ((size_t)g * P->NW + w)]);
            cmin = std::min(cmin, c), cmax = std::max(cmax, c);
            ce.push_back({c, g});
        }
        std::sort(ce.begin(), ce.end());
        fprintf(stderr, "[csrk] slab CTAs by end time (us after the first): median %.1f; last six:", (double)(ce[ce.size() / 2].first - cmin) / 1e3);
        for (size_t i = ce.size() >= 6 ? ce.size() - 6 : 0; i < ce.size(); i++)
            fprintf(stderr, " #%d %.1f", ce[i].second, (double)(ce[i].first - cmin) / 1e3);
        fprintf(stderr, "; first three:");
        for (size_t i = 0; i < 3 && i < ce.size(); i++)
            fprintf(stderr, " #%d", ce[i].second);
        fprintf(stderr, "\n");
        fprintf(stderr, "[csrk] slab warps: the first walk ends %.1f us before the mean, the last %.1f us after it; epilogues end "
                        "%.1f us after the last walk; last warp of a CTA: %.1f us between the CTAs\n",
                mean / 1e3, ((double)(hi - lo) - mean) / 1e3, (double)(hi2 - hi) / 1e3, (double)(cmax - cmin) / 1e3);
    }
    return rc;
}

void stream_info(const StreamPlan *P, int64_t *out /*[11]*/)
{
    out[0] = P->G, out[1] = P->NW, out[2] = P->nslab, out[3] = P->S, out[4] = P->P, out[5] = P->Q, out[6] = P->n_split,
    out[7] = (int64_t)P->smem_bytes, out[8] = P->stream_bytes, out[9] = P->piece, out[10] = P->ring;
}

}  // namespace csrk
