/*
 * csrk.h -- C ABI of libcsr_cuda.so, the B200 (sm_100a) kernel backend for
 * lenskit/csr's kernel contract.
 *
 * The reference has no C ABI for this path: its boundary is the Python module
 * protocol `csr.kernels.<name>` (docs/kernels.rst:61-104, csr/kernel.py:9-16).
 * Its one native precedent is the MKL shim (csr/kernels/mkl/mkl_ops.h:11-31);
 * the entry points below are what a `csr.kernels.cuda` module binds through
 * ctypes, and each cites the reference function it stands in for.
 *
 * Conventions
 *   - Plain pointers and sizes only.  `csrk_h` is an opaque handle owning device
 *     memory; host callers never see device pointers except through the *_dev
 *     entry points (used by the multi-GPU layer and the benchmark, which own
 *     device buffers of their own).
 *   - Every function returns an int status: 0 = ok, else a CSRK_E* code;
 *     csrk_last_error() gives the message for the calling thread.  Nothing
 *     aborts the process (unlike mkl_ops.c:18-45).
 *   - rowptrs are int32 or int64 (`rp_is64`), colinds int32, values float32,
 *     float64 or absent (`val_kind` = 4, 8, 0) -- the dtype rules of
 *     csr/csr.py:79-100.
 *   - Host-pointer entry points run on the library's own stream and synchronise
 *     before returning.  *_dev entry points enqueue on the cudaStream_t passed as
 *     `stream` -- used verbatim, so NULL is CUDA's default stream, as in every
 *     CUDA API; csrk_get_stream() returns the library stream -- and return
 *     without synchronising.
 *   - Re-entrant: concurrent calls from several threads on distinct or shared
 *     read-only handles are legal (the numba kernels are nogil:
 *     numba/__init__.py:55, multiply.py:13,41).  csrk_order_columns,
 *     csrk_filter_zeros and csrk_free mutate their handle.
 */
#ifndef CSRK_H
#define CSRK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct csrk_matrix *csrk_h;

enum {
    CSRK_OK = 0,
    CSRK_EARG = 1,      /* bad argument / shape mismatch -> ValueError            */
    CSRK_ENOMEM = 2,    /* device or host allocation failed -> MemoryError        */
    CSRK_ECUDA = 3,     /* CUDA runtime error -> RuntimeError                     */
    CSRK_ENODEV = 4,    /* no usable CUDA device -> RuntimeError                  */
    CSRK_EOVERFLOW = 5  /* a size exceeds what the output dtype can address       */
};

/* ---- library / device context ------------------------------------------- */
int csrk_version(void);
const char *csrk_last_error(void);
/* Select the device for the calling process (one process per GPU) and create
 * the library stream + memory pool.  Idempotent for the same device. */
int csrk_init(int device);
int csrk_shutdown(void);
/* sm count, HBM bytes (total, free) of the active device */
int csrk_device_info(int *sm_count, int64_t *mem_total, int64_t *mem_free, int *cc_major, int *cc_minor);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t csrk_launch_count(void);
int csrk_synchronize(void);
/* Tunables: "spmv_mode" 0 auto | 1 CSR tile kernel | 2 slab kernel (x staged in shared memory);
 * "stream_min_nnz" smallest nnz for which auto mode builds a stream plan;
 * "stream_slab_bytes" > 0 caps the x slab size, "stream_ctas" > 0 sets the number of CTAs (row groups),
 *               "stream_warps" 1..31 the consumer warps per CTA of the slab kernel, "stream_piece" 8..4096 the
 *               longest pseudo-row (longer rows are cut), "stream_ring_bytes" 4096 | 8192 the per-warp prefetch
 *               ring, "stream_ring_chunks" 2 | 4 the bulk copies per ring, "stream_xbufs" 2 | 3 the x slabs resident
 *               per CTA (tests, tuning);
 * "sym_bytes"   1 | 0: SpGEMM symbolic pass of heavy rows marks columns with byte stores in shared memory when they
 *               fit (else bitmap words with atomicOr);
 * "own_nw"      8 | 16 column ranges (warps) per CTA in the owner-computes SpGEMM numeric kernel;
 * "spmv_zero_copy_y" 1 | 0: csrk_spmv stores finished rows straight into y when y is pinned host memory;
 * "fix_threads" 512 | 768 | 1024 threads per CTA of that kernel (default 1024);
 * "spgemm_fixed" 1 | 0: allow the fixed-point numeric kernel for heavy rows (see csrk_spgemm_path);
 * "own_chunk_prod" SpGEMM: rows with more products than this are cut into chunks of A entries handled by
 *               different CTAs and summed in chunk order (0 = 1/8 of an SM's fair share, < 0 = never:
 *               every output element is then summed in the reference's own order, bit-identical values);
 * "radix_bits"  0 | 8 | 9 digit width of the stable sort behind transpose/order_columns
 *               (0 picks 9 when that saves a pass);
 * "fix_tiny_cap" > 0: capacity (entries) of the fixed-point kernel's side list (0 = 1/32 of the products; when it
 *               overflows the owner-computes kernel redoes the heavy rows);
 * "spgemm_esc"  SpGEMM expand/sort/compress path: 0 off | 1 for wide results (more than four shared-memory column
 *               windows: ncols > 106 496) | 2 always; "esc_target" 16..4096 products per pseudo-row (row x column
 *               range) of that path; "esc_stride" 1 | 0: its count kernel takes the work items a prime stride apart;
 *               "esc_budget" > 0 caps the bytes of its expansion (default: half of the free
 *               device memory; over budget the path declines and the per-row kernels run). */
int csrk_set_option(const char *name, int64_t value);
/* the library's own (non-blocking) stream, as a cudaStream_t */
int csrk_get_stream(void **stream);

/* ---- handle lifecycle: to_handle / from_handle / release_handle ----------
 * numba/__init__.py:16-44 (identity there); precedent lk_mkl_spcreate /
 * lk_mkl_spexport_p / lk_mkl_spfree, mkl_ops.h:11-21. */
/* H2D copy of the three arrays; the matrix stays resident until csrk_free. */
int csrk_create(int32_t nrows, int32_t ncols, int64_t nnz,
                const void *rowptrs, int rp_is64,
                const int32_t *colinds,
                const void *values, int val_kind,
                csrk_h *out);
/* Same, from device pointers.  The D2D copies are enqueued on `stream` (the source arrays may be reused
 * once work enqueued on `stream` after this call has run) and the library stream is made to wait for
 * them, so the handle may be used by any entry point right away; the call does not synchronise. */
int csrk_create_dev(int32_t nrows, int32_t ncols, int64_t nnz,
                    const void *d_rowptrs, int rp_is64,
                    const int32_t *d_colinds,
                    const void *d_values, int val_kind,
                    void *stream, csrk_h *out);
int csrk_free(csrk_h h);
int csrk_dims(csrk_h h, int32_t *nrows, int32_t *ncols, int64_t *nnz, int *rp_is64, int *val_kind);
/* D2H copy-out into caller-owned buffers sized from csrk_dims (values may be
 * NULL when val_kind == 0). */
int csrk_export(csrk_h h, void *rowptrs, int32_t *colinds, void *values);
/* Borrow the device arrays (valid until csrk_free / a mutating call). */
int csrk_device_ptrs(csrk_h h, void **d_rowptrs, int32_t **d_colinds, void **d_values);
/* Rows [begin, end) as a new handle with rebased rowptrs: the device form of
 * subset_rows, csr/structure.py:70-81 (used for sharding, csr.py:599-621). */
int csrk_subset_rows(csrk_h h, int32_t begin, int32_t end, csrk_h *out);

/* ---- mult_vec: numba/__init__.py:55-67; lk_mkl_spmv, mkl_ops.h:29 --------
 * y[r] = sum_i x[colinds[i]] * (values[i] or 1), float64 accumulate, float64 y.
 * x_kind = 4 (float32) or 8 (float64).  Host pointers (pinned or pageable).  When y is pinned
 * (cudaHostAlloc / cudaHostRegister, e.g. a torch pin_memory() tensor) the kernel writes each finished
 * row directly into it over PCIe, overlapping the device-to-host transfer with the compute. */
int csrk_spmv(csrk_h h, const void *x, int x_kind, double *y);
/* Which SpMV kernel serves this handle for x of x_kind, and the shape of its plan:
 * info[0] = 1 when a slab plan exists (built by the first mult_vec that selected it), else 0 (CSR tile
 * kernel); info[1..11] = CTAs, consumer warps per CTA, x slabs, columns per slab, pseudo-rows per warp,
 * pseudo-rows, split rows, shared-memory bytes per CTA, bytes of the re-laid-out entry stream, longest
 * pseudo-row, prefetch ring bytes per warp. */
int csrk_spmv_plan_info(csrk_h h, int x_kind, int64_t info[12]);
/* Device pointers; y has nrows doubles. */
int csrk_spmv_dev(csrk_h h, const void *d_x, int x_kind, double *d_y, void *stream);
/* Fused SpMV + gather for the row-partitioned multi-GPU SpMV: every finished row is stored to
 * d_ys[0] (this GPU) AND to d_ys[1..n_out) (the same y segment inside the peers' gather buffers,
 * NVLink peer / symmetric memory), n_out <= 8.  d_ys is a HOST array of device pointers. */
int csrk_spmv_dev_multi(csrk_h h, const void *d_x, int x_kind, double *const *d_ys, int n_out, void *stream);
/* The same over NVLink multicast (NVLS): d_y_mc is the MULTICAST address of this rank's y segment
 * inside a symmetric gather buffer (cuMulticast* / torch symmetric memory `multicast_ptr` + offset),
 * d_y the segment's ordinary local address.  Every finished row is stored once with multimem.st and
 * the NVSwitch replicates it into every GPU of the group; the caller adds the barrier. */
int csrk_spmv_dev_mc(csrk_h h, const void *d_x, int x_kind, double *d_y, double *d_y_mc, void *stream);
/* NVLS broadcast: copy nbytes (multiple of 4, 16-byte aligned pointers) from local device memory to a
 * multicast address -- the root's half of "x is broadcast" (BASELINE north_star), one NVLink egress
 * instead of world-1.  The caller orders it with barriers on both sides. */
int csrk_mc_broadcast(void *mc_dst, const void *d_src, int64_t nbytes, void *stream);

/* ---- mult_ab / mult_abt: multiply.py:13-57; lk_mkl_spmab/spmabt ----------
 * C = A*B (a.ncols == b.nrows) and C = A*B^T (a.ncols == b.ncols) as a NEW
 * handle the caller frees.  Output: float64 values, int32 colinds sorted
 * ascending within each row, rowptrs int32 (int64 when out-nnz > INT32_MAX,
 * the rule of csr.py:90-93). */
int csrk_spgemm(csrk_h a, csrk_h b, csrk_h *c);
int csrk_spgemm_abt(csrk_h a, csrk_h b, csrk_h *c);
/* Work counters of the last product that produced `c`: products P (sum over
 * A's entries of the referenced B row length) and out-nnz Z. */
int csrk_spgemm_stats(csrk_h c, int64_t *products, int64_t *out_nnz);
/* Which numeric kernel handled the heavy (dense-accumulator) rows of the product that made c:
 * 0 none / the general dense kernels, 1 owner-computes (float64, reference summation order),
 * 2 64-bit fixed point on native shared-memory atomics over power-of-two equilibrated operands (taken for
 * finite values with exponents within +-400; products too small for the accumulator's grid go, exactly, through a
 * side list and are added in float64, so every output element is within 2^-35 of sum_k |a_ik||b_kj|;
 * "spgemm_fixed" = 0 disables it). */
int csrk_spgemm_path(csrk_h c, int *path);
/* how many products of that multiplication went through the fixed-point kernel's side list */
int csrk_spgemm_side_list(csrk_h c, int64_t *entries);

/* ---- transpose: csr/structure.py:172-247 ---------------------------------
 * Stable CSR->CSC.  rowptr dtype follows the input; values become float64
 * (structure.py:177); with_values == 0 (or a value-less input) gives a
 * structure-only result. */
int csrk_transpose(csrk_h a, int with_values, csrk_h *at);

/* ---- order_columns: numba/__init__.py:47-52 -> structure.py:156-169 ------
 * In-place ascending column sort inside each row, values carried along; equal
 * columns keep their relative order (the bubble sort is stable). */
int csrk_order_columns(csrk_h h);

/* ---- from_coo: csr/csr.py:140-169 -> csr/structure.py:11-58 -----------------
 * Build a handle straight from COO triples on the host (int32 rows/cols, values of val_kind 0|4|8):
 * entries keep their COO order inside a row (stable), values keep their dtype, rowptrs are int32
 * unless nnz > INT32_MAX.  An index outside the shape is CSRK_EARG. */
int csrk_from_coo(int32_t nrows, int32_t ncols, int64_t nnz, const int32_t *rows, const int32_t *cols, const void *values,
                  int val_kind, csrk_h *out);

/* ---- normalize_rows: csr/csr.py:443-469 -> csr/transform.py:13-66 ----------
 * In place on the handle's values.  kind 0 = 'center' (subtract each non-empty row's mean, returns the
 * means), kind 1 = 'unit' (power-of-two pre-normalisation, divide by the Euclidean norm, returns the
 * norms; an all-zero row becomes NaN as in the reference).  vec: HOST array of nrows elements of the
 * matrix's value type.  values_out (optional): HOST array of nnz elements that receives the normalised
 * values (what the reference leaves in csr.values).  A matrix without values is an argument error. */
int csrk_normalize_rows(csrk_h h, int kind, void *vec, void *values_out);

/* ---- _filter_zeros: csr/_struct.py:61-79 ----------------------------------
 * In-place removal of stored zeros (no-op for value-less matrices). */
int csrk_filter_zeros(csrk_h h);

#ifdef __cplusplus
}
#endif
#endif /* CSRK_H */
