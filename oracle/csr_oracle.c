/*
 * csr_oracle.c -- CPU restatement of the lenskit/csr *numba* kernel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under csr_b200/ may import, link or call
 * this file.  It is used by tests/ (as the parity checker), by
 * __graft_entry__.smoke() (as the checker) and by bench.py's cpu_baseline /
 * --impl reference legs (as the CPU implementation that is timed).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
 * bit-for-bit (structure AND values) against outputs of the unmodified reference
 * (numba kernel, imported from /root/reference in the build container) stored
 * in tests/golden/ by tests/golden/make_golden.py.
 *
 * Every loop below follows the reference statement by statement, including the
 * arithmetic types: numba computes `a * b` in the promoted type of its operands
 * (f4*f4 -> f4, anything else -> f8) and then adds into a float64 accumulator.
 * Compile with -ffp-contract=off so no FMA is formed (numba does not fuse).
 *
 * Reference (relative to /root/reference/):
 *   mult_vec            csr/kernels/numba/__init__.py:55-67
 *   mult_ab             csr/kernels/numba/multiply.py:13-38
 *   mult_abt            csr/kernels/numba/multiply.py:41-57
 *   _sym_mm             csr/kernels/numba/multiply.py:60-100
 *   _num_mm             csr/kernels/numba/multiply.py:103-129
 *   _transpose_values   csr/structure.py:172-204
 *   _transpose_structure csr/structure.py:207-237
 *   sort_rows           csr/structure.py:156-169
 *   _filter_zeros       csr/_struct.py:61-79
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef int32_t i32;
typedef int64_t i64;

#define ORC_OK 0
#define ORC_EARG 1
#define ORC_ENOMEM 2
#define ORC_EOVERFLOW 3

/* ---------------------------------------------------------------- mult_vec */
/* numba/__init__.py:55-67: one serial pass over nnz, row cursor advanced by a
 * while loop (:62-63), res[row] += v[col] * value (:65), res is float64 zeros
 * (:57).  A missing value array means the value 1.0 (float64) -- _wiring.py:78-89. */
#define DEF_MULT_VEC(NAME, RPT, VT, XT, PT)                                        \
    static void NAME(i32 nrows, i64 nnz, const RPT *rp, const i32 *ci,            \
                     const VT *vs, const XT *x, double *res)                      \
    {                                                                              \
        i64 row = 0;                                                               \
        memset(res, 0, sizeof(double) * (size_t)nrows);                            \
        for (i64 i = 0; i < nnz; i++) {                                            \
            while (i == (i64)rp[row + 1])                                          \
                row++;                                                             \
            i32 col = ci[i];                                                       \
            PT p = (PT)x[col] * (PT)vs[i];                                         \
            res[row] += (double)p;                                                 \
        }                                                                          \
    }

#define DEF_MULT_VEC_NV(NAME, RPT, XT)                                             \
    static void NAME(i32 nrows, i64 nnz, const RPT *rp, const i32 *ci,            \
                     const XT *x, double *res)                                    \
    {                                                                              \
        i64 row = 0;                                                               \
        memset(res, 0, sizeof(double) * (size_t)nrows);                            \
        for (i64 i = 0; i < nnz; i++) {                                            \
            while (i == (i64)rp[row + 1])                                          \
                row++;                                                             \
            res[row] += (double)x[ci[i]] * 1.0;                                    \
        }                                                                          \
    }

DEF_MULT_VEC(mv_r4_v4_x4, i32, float, float, float)
DEF_MULT_VEC(mv_r4_v4_x8, i32, float, double, double)
DEF_MULT_VEC(mv_r4_v8_x4, i32, double, float, double)
DEF_MULT_VEC(mv_r4_v8_x8, i32, double, double, double)
DEF_MULT_VEC(mv_r8_v4_x4, i64, float, float, float)
DEF_MULT_VEC(mv_r8_v4_x8, i64, float, double, double)
DEF_MULT_VEC(mv_r8_v8_x4, i64, double, float, double)
DEF_MULT_VEC(mv_r8_v8_x8, i64, double, double, double)
DEF_MULT_VEC_NV(mv_r4_nv_x4, i32, float)
DEF_MULT_VEC_NV(mv_r4_nv_x8, i32, double)
DEF_MULT_VEC_NV(mv_r8_nv_x4, i64, float)
DEF_MULT_VEC_NV(mv_r8_nv_x8, i64, double)

int orc_mult_vec(i32 nrows, i32 ncols, i64 nnz, const void *rp, int rp_is64,
                 const i32 *ci, const void *vs, int val_kind, const void *x,
                 int x_kind, double *y)
{
    (void)ncols;
    if (x_kind != 4 && x_kind != 8)
        return ORC_EARG;
#define MV_CALL(F, RPT, VT, XT) F(nrows, nnz, (const RPT *)rp, ci, (const VT *)vs, (const XT *)x, y)
#define MV_CALL_NV(F, RPT, XT) F(nrows, nnz, (const RPT *)rp, ci, (const XT *)x, y)
    if (!rp_is64) {
        if (val_kind == 4) {
            if (x_kind == 4) MV_CALL(mv_r4_v4_x4, i32, float, float);
            else MV_CALL(mv_r4_v4_x8, i32, float, double);
        } else if (val_kind == 8) {
            if (x_kind == 4) MV_CALL(mv_r4_v8_x4, i32, double, float);
            else MV_CALL(mv_r4_v8_x8, i32, double, double);
        } else if (val_kind == 0) {
            if (x_kind == 4) MV_CALL_NV(mv_r4_nv_x4, i32, float);
            else MV_CALL_NV(mv_r4_nv_x8, i32, double);
        } else
            return ORC_EARG;
    } else {
        if (val_kind == 4) {
            if (x_kind == 4) MV_CALL(mv_r8_v4_x4, i64, float, float);
            else MV_CALL(mv_r8_v4_x8, i64, float, double);
        } else if (val_kind == 8) {
            if (x_kind == 4) MV_CALL(mv_r8_v8_x4, i64, double, float);
            else MV_CALL(mv_r8_v8_x8, i64, double, double);
        } else if (val_kind == 0) {
            if (x_kind == 4) MV_CALL_NV(mv_r8_nv_x4, i64, float);
            else MV_CALL_NV(mv_r8_nv_x8, i64, double);
        } else
            return ORC_EARG;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ SpGEMM */
typedef struct {
    i32 nrows, ncols;
    i64 nnz;
    const void *rp;
    int rp_is64;
    const i32 *ci;
    const void *vs;
    int val_kind; /* 4 or 8; the reference rejects value-less inputs here */
} orc_mat;

static inline i64 rp_at(const orc_mat *m, i64 i)
{
    return m->rp_is64 ? ((const i64 *)m->rp)[i] : (i64)((const i32 *)m->rp)[i];
}

static inline i32 imax3(i32 a, i32 b, i32 c)
{
    i32 m = a > b ? a : b;
    return m > c ? m : c;
}

/* _sym_mm, multiply.py:60-100.  Fills c_rp (int32, length nrows+1, c_rp[0]==0
 * on entry) and returns the column indices in the reference's order: each row
 * is the SMMP linked list popped from its head, i.e. reverse first-touch order
 * (:79-82 push, :94-97 pop).  Row-pointer arithmetic is done in int32 exactly
 * as np.intc does (:28,93), so out-nnz > INT32_MAX wraps like the reference;
 * we report that instead of silently continuing. */
static int sym_mm(const orc_mat *a, const orc_mat *b, i32 *c_rp, i32 **out_ci, i64 *out_len)
{
    i32 wlen = imax3(a->nrows, a->ncols, b->ncols); /* :62 */
    i32 *index = (i32 *)malloc(sizeof(i32) * (size_t)(wlen > 0 ? wlen : 1));
    i64 c_len = a->nnz > b->nnz ? a->nnz : b->nnz; /* :64 */
    if (c_len < 16)
        c_len = 16; /* the reference would loop forever growing 0 by 0//2; nnz==0 never reaches growth */
    i32 *c_ci = (i32 *)malloc(sizeof(i32) * (size_t)c_len);
    i64 c_pos = 0;
    if (!index || !c_ci) {
        free(index);
        free(c_ci);
        return ORC_ENOMEM;
    }
    for (i32 k = 0; k < wlen; k++)
        index[k] = -1; /* :63 */

    for (i32 i = 0; i < a->nrows; i++) {
        i32 istart = wlen; /* :69 */
        i64 length = 0;
        i64 a_rs = rp_at(a, i), a_re = rp_at(a, (i64)i + 1);
        for (i64 jj = a_rs; jj < a_re; jj++) { /* :74-82 */
            i32 j = a->ci[jj];
            i64 b_rs = rp_at(b, j), b_re = rp_at(b, (i64)j + 1);
            for (i64 kk = b_rs; kk < b_re; kk++) {
                i32 k = b->ci[kk];
                if (index[k] < 0) {
                    index[k] = istart;
                    istart = k;
                    length++;
                }
            }
        }
        while (c_pos + length > c_len) { /* :85-90, growth by half */
            i64 ncl = c_len + c_len / 2;
            i32 *ci2 = (i32 *)realloc(c_ci, sizeof(i32) * (size_t)ncl);
            if (!ci2) {
                free(index);
                free(c_ci);
                return ORC_ENOMEM;
            }
            c_ci = ci2;
            c_len = ncl;
        }
        if ((i64)c_rp[i] + length > (i64)INT32_MAX) { /* np.intc would wrap here (:93) */
            free(index);
            free(c_ci);
            return ORC_EOVERFLOW;
        }
        c_rp[i + 1] = c_rp[i] + (i32)length; /* :93 */
        for (i64 j = c_rp[i]; j < c_rp[i + 1]; j++) { /* :94-97 */
            c_ci[j] = istart;
            istart = index[istart];
            index[c_ci[j]] = -1;
        }
        c_pos += length;
    }
    free(index);
    *out_ci = c_ci;
    *out_len = c_pos;
    return ORC_OK;
}

/* _num_mm, multiply.py:103-129.  Dense float64 accumulator work[wlen] (:107);
 * work[k] += av * bv with the product in the operands' promoted type (:120);
 * gather through c_ci and reset (:124-127). */
#define DEF_NUM_MM(NAME, AVT, BVT, PT)                                             \
    static int NAME(const orc_mat *a, const orc_mat *b, const i32 *c_rp,          \
                    const i32 *c_ci, i64 c_nnz, double *c_vs)                      \
    {                                                                              \
        i32 wlen = imax3(a->nrows, a->ncols, b->ncols);                            \
        double *work = (double *)calloc((size_t)(wlen > 0 ? wlen : 1), sizeof(double)); \
        const AVT *avs = (const AVT *)a->vs;                                       \
        const BVT *bvs = (const BVT *)b->vs;                                       \
        if (!work)                                                                 \
            return ORC_ENOMEM;                                                     \
        (void)c_nnz;                                                               \
        for (i32 i = 0; i < a->nrows; i++) {                                       \
            i64 a_rs = rp_at(a, i), a_re = rp_at(a, (i64)i + 1);                   \
            for (i64 jj = a_rs; jj < a_re; jj++) {                                 \
                i32 j = a->ci[jj];                                                 \
                AVT av = avs[jj];                                                  \
                i64 b_rs = rp_at(b, j), b_re = rp_at(b, (i64)j + 1);               \
                for (i64 kk = b_rs; kk < b_re; kk++) {                             \
                    i32 k = b->ci[kk];                                             \
                    PT p = (PT)av * (PT)bvs[kk];                                   \
                    work[k] += (double)p;                                          \
                }                                                                  \
            }                                                                      \
            for (i64 jj = c_rp[i]; jj < c_rp[i + 1]; jj++) {                       \
                i32 j = c_ci[jj];                                                  \
                c_vs[jj] = work[j];                                                \
                work[j] = 0;                                                       \
            }                                                                      \
        }                                                                          \
        free(work);                                                                \
        return ORC_OK;                                                             \
    }

DEF_NUM_MM(num_mm_44, float, float, float)
DEF_NUM_MM(num_mm_48, float, double, double)
DEF_NUM_MM(num_mm_84, double, float, double)
DEF_NUM_MM(num_mm_88, double, double, double)

static int mat_init(orc_mat *m, i32 nrows, i32 ncols, i64 nnz, const void *rp,
                    int rp_is64, const i32 *ci, const void *vs, int val_kind)
{
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->rp = rp;
    m->rp_is64 = rp_is64;
    m->ci = ci;
    m->vs = vs;
    m->val_kind = val_kind;
    return (val_kind == 4 || val_kind == 8) ? ORC_OK : ORC_EARG;
}

/* mult_ab, multiply.py:13-38.  c_rp: caller buffer of nrows+1 int32.  *c_ci and
 * *c_vs are malloc'd here (free with orc_free); *c_nnz = c_rp[nrows]. */
int orc_mult_ab(i32 a_nrows, i32 a_ncols, i64 a_nnz, const void *a_rp, int a_rp64,
                const i32 *a_ci, const void *a_vs, int a_vk,
                i32 b_nrows, i32 b_ncols, i64 b_nnz, const void *b_rp, int b_rp64,
                const i32 *b_ci, const void *b_vs, int b_vk,
                i32 *c_rp, i32 **c_ci, double **c_vs, i64 *c_nnz)
{
    orc_mat a, b;
    int rc;
    if (a_ncols != b_nrows) /* :26 */
        return ORC_EARG;
    if (mat_init(&a, a_nrows, a_ncols, a_nnz, a_rp, a_rp64, a_ci, a_vs, a_vk) ||
        mat_init(&b, b_nrows, b_ncols, b_nnz, b_rp, b_rp64, b_ci, b_vs, b_vk))
        return ORC_EARG;
    memset(c_rp, 0, sizeof(i32) * ((size_t)a_nrows + 1)); /* :28 */
    i32 *ci = NULL;
    i64 len = 0;
    rc = sym_mm(&a, &b, c_rp, &ci, &len); /* :31 */
    if (rc)
        return rc;
    double *vs = (double *)calloc((size_t)(len > 0 ? len : 1), sizeof(double)); /* :109 */
    if (!vs) {
        free(ci);
        return ORC_ENOMEM;
    }
    if (a_vk == 4 && b_vk == 4) rc = num_mm_44(&a, &b, c_rp, ci, len, vs);
    else if (a_vk == 4) rc = num_mm_48(&a, &b, c_rp, ci, len, vs);
    else if (b_vk == 4) rc = num_mm_84(&a, &b, c_rp, ci, len, vs);
    else rc = num_mm_88(&a, &b, c_rp, ci, len, vs);
    if (rc) {
        free(ci);
        free(vs);
        return rc;
    }
    *c_ci = ci;
    *c_vs = vs;
    *c_nnz = len;
    return ORC_OK;
}

/* symbolic phase alone (used by test_symb-style invariants and size probes) */
int orc_sym_mm(i32 a_nrows, i32 a_ncols, i64 a_nnz, const void *a_rp, int a_rp64, const i32 *a_ci,
               i32 b_nrows, i32 b_ncols, i64 b_nnz, const void *b_rp, int b_rp64, const i32 *b_ci,
               i32 *c_rp, i32 **c_ci, i64 *c_nnz)
{
    orc_mat a, b;
    if (a_ncols != b_nrows)
        return ORC_EARG;
    mat_init(&a, a_nrows, a_ncols, a_nnz, a_rp, a_rp64, a_ci, NULL, 8);
    mat_init(&b, b_nrows, b_ncols, b_nnz, b_rp, b_rp64, b_ci, NULL, 8);
    memset(c_rp, 0, sizeof(i32) * ((size_t)a_nrows + 1));
    return sym_mm(&a, &b, c_rp, c_ci, c_nnz);
}

void orc_free(void *p) { free(p); }

/* --------------------------------------------------------------- transpose */
/* structure.py:172-204 / :207-237.  Count per column into brp[j+1] (:180-184),
 * running sum (:187-188), stable scatter using brp[j] as the cursor (:191-197),
 * shift the pointers back (:200-202).  brp has the input's rowptr dtype (:175),
 * bci is int32 (:176), bvs is ALWAYS float64 (:177). */
#define DEF_TRANSPOSE(NAME, RPT)                                                   \
    static void NAME(i32 nrows, i32 ncols, i64 nnz, const RPT *rp, const i32 *ci, \
                     const void *vs, int val_kind, RPT *brp, i32 *bci, double *bvs) \
    {                                                                              \
        (void)nnz;                                                                 \
        memset(brp, 0, sizeof(RPT) * ((size_t)ncols + 1));                         \
        for (i32 i = 0; i < nrows; i++)                                            \
            for (i64 jj = rp[i]; jj < (i64)rp[i + 1]; jj++)                        \
                brp[ci[jj] + 1] += 1;                                              \
        for (i32 j = 0; j < ncols; j++)                                            \
            brp[j + 1] = brp[j] + brp[j + 1];                                      \
        for (i32 i = 0; i < nrows; i++) {                                          \
            for (i64 jj = rp[i]; jj < (i64)rp[i + 1]; jj++) {                      \
                i32 j = ci[jj];                                                    \
                bci[brp[j]] = i;                                                   \
                if (val_kind == 4)                                                 \
                    bvs[brp[j]] = (double)((const float *)vs)[jj];                 \
                else if (val_kind == 8)                                            \
                    bvs[brp[j]] = ((const double *)vs)[jj];                        \
                brp[j] += 1;                                                       \
            }                                                                      \
        }                                                                          \
        for (i32 i = ncols - 1; i > 0; i--)                                        \
            brp[i] = brp[i - 1];                                                   \
        if (ncols >= 0)                                                            \
            brp[0] = 0;                                                            \
    }

DEF_TRANSPOSE(transpose_r4, i32)
DEF_TRANSPOSE(transpose_r8, i64)

/* val_kind 0 => structure only (bvs may be NULL) */
int orc_transpose(i32 nrows, i32 ncols, i64 nnz, const void *rp, int rp_is64,
                  const i32 *ci, const void *vs, int val_kind,
                  void *brp, i32 *bci, double *bvs)
{
    if (val_kind != 0 && val_kind != 4 && val_kind != 8)
        return ORC_EARG;
    if (rp_is64)
        transpose_r8(nrows, ncols, nnz, (const i64 *)rp, ci, vs, val_kind, (i64 *)brp, bci, bvs);
    else
        transpose_r4(nrows, ncols, nnz, (const i32 *)rp, ci, vs, val_kind, (i32 *)brp, bci, bvs);
    return ORC_OK;
}

/* mult_abt, multiply.py:41-57: transpose B (values become float64), then mult_ab. */
int orc_mult_abt(i32 a_nrows, i32 a_ncols, i64 a_nnz, const void *a_rp, int a_rp64,
                 const i32 *a_ci, const void *a_vs, int a_vk,
                 i32 b_nrows, i32 b_ncols, i64 b_nnz, const void *b_rp, int b_rp64,
                 const i32 *b_ci, const void *b_vs, int b_vk,
                 i32 *c_rp, i32 **c_ci, double **c_vs, i64 *c_nnz)
{
    if (a_ncols != b_ncols) /* :53 */
        return ORC_EARG;
    if (b_vk != 4 && b_vk != 8)
        return ORC_EARG;
    size_t rpsz = b_rp64 ? sizeof(i64) : sizeof(i32);
    void *brp = malloc(rpsz * ((size_t)b_ncols + 1));
    i32 *bci = (i32 *)malloc(sizeof(i32) * (size_t)(b_nnz > 0 ? b_nnz : 1));
    double *bvs = (double *)malloc(sizeof(double) * (size_t)(b_nnz > 0 ? b_nnz : 1));
    if (!brp || !bci || !bvs) {
        free(brp);
        free(bci);
        free(bvs);
        return ORC_ENOMEM;
    }
    orc_transpose(b_nrows, b_ncols, b_nnz, b_rp, b_rp64, b_ci, b_vs, b_vk, brp, bci, bvs); /* :56 */
    int rc = orc_mult_ab(a_nrows, a_ncols, a_nnz, a_rp, a_rp64, a_ci, a_vs, a_vk,
                         b_ncols, b_nrows, b_nnz, brp, b_rp64, bci, bvs, 8,
                         c_rp, c_ci, c_vs, c_nnz); /* :57 */
    free(brp);
    free(bci);
    free(bvs);
    return rc;
}

/* --------------------------------------------------------------- sort_rows */
/* structure.py:156-169: per-row bubble sort on colinds, values swapped along
 * (_util.py:5-25).  In place. */
int orc_sort_rows(i32 nrows, const void *rp, int rp_is64, i32 *ci, void *vs, int val_kind)
{
    for (i32 i = 0; i < nrows; i++) {
        i64 sp = rp_is64 ? ((const i64 *)rp)[i] : ((const i32 *)rp)[i];
        i64 ep = rp_is64 ? ((const i64 *)rp)[i + 1] : ((const i32 *)rp)[i + 1];
        int swapped = 1;
        while (swapped) {
            swapped = 0;
            for (i64 j = sp; j < ep - 1; j++) {
                if (ci[j] > ci[j + 1]) {
                    i32 t = ci[j];
                    ci[j] = ci[j + 1];
                    ci[j + 1] = t;
                    if (val_kind == 4) {
                        float *v = (float *)vs;
                        float tv = v[j];
                        v[j] = v[j + 1];
                        v[j + 1] = tv;
                    } else if (val_kind == 8) {
                        double *v = (double *)vs;
                        double tv = v[j];
                        v[j] = v[j + 1];
                        v[j + 1] = tv;
                    }
                    swapped = 1;
                }
            }
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------ filter_zeros */
/* _struct.py:61-79: in-place compaction of stored zeros; rewrites rowptrs,
 * returns the new nnz (the caller truncates colinds/values). */
int orc_filter_zeros(i32 nrows, void *rp, int rp_is64, i32 *ci, void *vs, int val_kind, i64 *new_nnz)
{
    i64 nnz = 0;
    if (val_kind != 4 && val_kind != 8)
        return ORC_EARG;
    for (i32 i = 0; i < nrows; i++) {
        i64 sp = rp_is64 ? ((i64 *)rp)[i] : ((i32 *)rp)[i];
        i64 ep = rp_is64 ? ((i64 *)rp)[i + 1] : ((i32 *)rp)[i + 1];
        if (rp_is64) ((i64 *)rp)[i] = nnz;
        else ((i32 *)rp)[i] = (i32)nnz;
        for (i64 jp = sp; jp < ep; jp++) {
            int nz = val_kind == 4 ? (((float *)vs)[jp] != 0) : (((double *)vs)[jp] != 0);
            if (nz) {
                ci[nnz] = ci[jp];
                if (val_kind == 4) ((float *)vs)[nnz] = ((float *)vs)[jp];
                else ((double *)vs)[nnz] = ((double *)vs)[jp];
                nnz++;
            }
        }
    }
    if (rp_is64) ((i64 *)rp)[nrows] = nnz;
    else ((i32 *)rp)[nrows] = (i32)nnz;
    *new_nnz = nnz;
    return ORC_OK;
}

/* ----------------------------------------------------------- normalisation */
/* csr/transform.py:13-26 (center_rows): per non-empty row, m = mean of the stored values, subtracted in
 * place; returns the means in the values' dtype.  numba's np.mean accumulates sequentially in the ARRAY's
 * dtype and divides by the (int64) size, which promotes to float64; `values -= m` is evaluated in
 * float64 and stored back in the array's dtype. */
int orc_center_rows(i32 nrows, const void *rp, int rp_is64, void *vs, int val_kind, void *means)
{
    if (val_kind != 4 && val_kind != 8)
        return ORC_EARG;
    for (i32 i = 0; i < nrows; i++) {
        i64 sp = rp_is64 ? ((const i64 *)rp)[i] : ((const i32 *)rp)[i];
        i64 ep = rp_is64 ? ((const i64 *)rp)[i + 1] : ((const i32 *)rp)[i + 1];
        if (val_kind == 4) ((float *)means)[i] = 0.0f;
        else ((double *)means)[i] = 0.0;
        if (sp == ep)
            continue;
        if (val_kind == 4) {
            float *v = (float *)vs;
            float c = 0.0f;
            for (i64 k = sp; k < ep; k++)
                c += v[k];
            double m = (double)c / (double)(ep - sp);
            ((float *)means)[i] = (float)m;
            for (i64 k = sp; k < ep; k++)
                v[k] = (float)((double)v[k] - m);
        } else {
            double *v = (double *)vs;
            double c = 0.0;
            for (i64 k = sp; k < ep; k++)
                c += v[k];
            double m = c / (double)(ep - sp);
            ((double *)means)[i] = m;
            for (i64 k = sp; k < ep; k++)
                v[k] -= m;
        }
    }
    return ORC_OK;
}

/* csr/transform.py:29-66 (unit_rows): pre-normalise by a power of two taken from the largest magnitude
 * (so that tiny rows do not underflow when squared), Euclidean norm, divide.  The reference's norm is
 * BLAS nrm2 (np.linalg.norm under numba); here it is sqrt of a float64 sum of squares, so norms agree to
 * a few ulp, not bit for bit -- tests state the tolerance. */
int orc_unit_rows(i32 nrows, const void *rp, int rp_is64, void *vs, int val_kind, void *norms)
{
    if (val_kind != 4 && val_kind != 8)
        return ORC_EARG;
    const int maxexp = val_kind == 4 ? 128 : 1024, minexp = val_kind == 4 ? -126 : -1022; /* np.finfo */
    for (i32 i = 0; i < nrows; i++) {
        i64 sp = rp_is64 ? ((const i64 *)rp)[i] : ((const i32 *)rp)[i];
        i64 ep = rp_is64 ? ((const i64 *)rp)[i + 1] : ((const i32 *)rp)[i + 1];
        if (val_kind == 4) ((float *)norms)[i] = 0.0f;
        else ((double *)norms)[i] = 0.0;
        if (sp == ep)
            continue;
        double vmax = 0.0;
        for (i64 k = sp; k < ep; k++) {
            double a = fabs(val_kind == 4 ? (double)((float *)vs)[k] : ((double *)vs)[k]);
            if (a > vmax || a != a)
                vmax = a;
        }
        int ve = 0;
        (void)frexp(vmax, &ve);
        int pnexp = -ve < maxexp - 1 ? -ve : maxexp - 1;
        if (pnexp < minexp)
            pnexp = minexp;
        const double prenorm = ldexp(1.0, pnexp);
        double ss = 0.0;
        if (val_kind == 4) {
            float *v = (float *)vs;
            for (i64 k = sp; k < ep; k++) {
                v[k] = (float)((double)v[k] * prenorm);
                ss += (double)v[k] * (double)v[k];
            }
            const float inorm = (float)sqrt(ss);
            ((float *)norms)[i] = (float)((double)inorm / prenorm);
            for (i64 k = sp; k < ep; k++)
                v[k] = v[k] / inorm;
        } else {
            double *v = (double *)vs;
            for (i64 k = sp; k < ep; k++) {
                v[k] = v[k] * prenorm;
                ss += v[k] * v[k];
            }
            const double inorm = sqrt(ss);
            ((double *)norms)[i] = inorm / prenorm;
            for (i64 k = sp; k < ep; k++)
                v[k] = v[k] / inorm;
        }
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------- from_coo */
/* csr/structure.py:11-58 (_from_coo_structure / _from_coo_values): count per row, prefix sum,
 * stable scatter with a per-row cursor -- entries keep their COO order inside a row; values keep
 * their dtype.  rowptrs are int64 here; the CSR constructor narrows them (csr/csr.py:90-93). */
int orc_from_coo(i32 nrows, i64 nnz, const i32 *rows, const i32 *cols, const void *vals, int val_kind,
                 i64 *rowptrs, i32 *out_cols, void *out_vals)
{
    i64 *rpos = (i64 *)calloc((size_t)nrows + 1, sizeof(i64));
    if (!rpos)
        return ORC_ENOMEM;
    for (i32 i = 0; i <= nrows; i++)
        rowptrs[i] = 0;
    for (i64 k = 0; k < nnz; k++) {
        if (rows[k] < 0 || rows[k] >= nrows) {
            free(rpos);
            return ORC_EARG;
        }
        rowptrs[rows[k] + 1]++;
    }
    for (i32 i = 0; i < nrows; i++)
        rowptrs[i + 1] += rowptrs[i];
    memcpy(rpos, rowptrs, ((size_t)nrows + 1) * sizeof(i64));
    for (i64 k = 0; k < nnz; k++) {
        const i64 pos = rpos[rows[k]]++;
        out_cols[pos] = cols[k];
        if (val_kind == 4) ((float *)out_vals)[pos] = ((const float *)vals)[k];
        else if (val_kind == 8) ((double *)out_vals)[pos] = ((const double *)vals)[k];
    }
    free(rpos);
    return ORC_OK;
}
