/*
 * numcl_oracle.h -- CPU restatement of numcl's einsum / broadcast / reduce loop generators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under numcl_b200/ may include, link, import or execute
 * anything in oracle/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY PINNING: the reference is Common Lisp and no Lisp implementation exists in this
 * image, so the oracle cannot be validated against the reference executed here.  It is pinned
 * against every known-answer vector the reference's own test-suite holds for this path
 * (t/package.lisp, see tests/test_oracle_golden.py).  For shapes/values those vectors do not
 * cover (floats with rounding, >125 elements, f64, batched einsum) parity is pinned only by
 * this transcription being reviewable against the cited lines: "parity unpinned" there.
 *
 * All file:line citations are relative to the numcl source tree.
 */
#ifndef NUMCL_ORACLE_H
#define NUMCL_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCO_MAX_RANK 8
#define NCO_MAX_T    8     /* arrays per side */
#define NCO_MAX_IDX  24

/* Same numbering as include/numcl_cuda.h ncl_dtype (SBCL storage formats). */
enum {
  NCO_BIT = 0, NCO_UB2, NCO_UB4, NCO_UB7, NCO_UB8, NCO_UB15, NCO_UB16, NCO_UB31, NCO_UB32,
  NCO_UB62, NCO_UB63, NCO_UB64, NCO_SB8, NCO_SB16, NCO_SB32, NCO_FIXNUM, NCO_SB64,
  NCO_F32, NCO_F64, NCO_DTYPE_COUNT
};

typedef struct {
  void*   data;                 /* base vector, SBCL storage format                     */
  int32_t dtype;
  int32_t rank;
  int64_t dims[NCO_MAX_RANK];
  int64_t offset;               /* displaced-index-offset (einsum ignores it, A.7)      */
} nco_tensor;

const char* nco_last_error(void);

/* einsum-normalize-subscripts, src/3einsum.lisp:64-200, + sort-locality :523-565.
 * `spec` is the text of the subscripts list without the outer parens, e.g.
 * "ij jk -> ik" or "i -> (cl:+ @1 $1) -> -> ((i :start 0 :end 10 :step 2))".
 * Writes a canonical printout "iter=(0 1 2) i=((0 1) (1 2)) o=((0 2)) t=((+ @1 (* $1 $2)))
 * io=(nil nil) oo=(nil)" into out.  Returns 0, or -1 (type-error for a bad spec char etc). */
int nco_normalize(const char* spec, char* out, int out_cap);
int nco_spec_counts(const char* spec, int* n_in, int* n_out);

/* Output shapes of an einsum call (shape-resolver + %output-generator,
 * src/3einsum.lisp:287-367,439-468).  out[k].rank/dims are filled for outputs whose data
 * pointer is NULL; supplied outputs are checked instead. */
int nco_einsum_shapes(const char* spec, int n_in, const nco_tensor* in, int n_out, nco_tensor* out);

/* Runs the loop nest the reference's backend would generate
 * (src/4einsum-backends/common-lisp.lisp:44-240) over the given arrays.  Outputs must be
 * allocated by the caller (zero-filled to mimic `zeros`, or holding data to accumulate into). */
int nco_einsum(const char* spec, int n_in, const nco_tensor* in, int n_out, nco_tensor* out);

/* plan-broadcast, src/3einsum.lisp:378-425.  shapes: n arrays, ranks[], dims flattened
 * row by row with stride NCO_MAX_RANK.  plan_out receives 2*(n+1)*M int64 in row-major
 * (2, n+1, M) order; returns M or -1. */
int nco_plan_broadcast(int n, const int* ranks, const int64_t* dims, int64_t* plan_out);

/* broadcast / broadcast-core, src/4numeric.lisp:201-294.  op is a Lisp function name
 * ("+", "-", "*", "/", "max", "min", "=/bit", "logand", ...).  r must have the broadcast result
 * shape (broadcast-result-shape, :187-199) and the dtype infer-type decided. */
int nco_broadcast(const char* op, const nco_tensor* x, const nco_tensor* y, nco_tensor* r);
int nco_broadcast_shape(const nco_tensor* x, const nco_tensor* y, nco_tensor* r);

/* reduce-array's compiled nest, src/5reduce.lisp:67-87: r[kept] = fn(r[kept], a[all]) with the
 * loops nested in axis order.  r must be pre-filled with the initial element (:56-59). */
int nco_reduce(const char* op, const nco_tensor* a, const int* axes, int n_axes, nco_tensor* r);

/* map-array-into, src/4numeric.lisp:37-51 (exactly array-total-size elements). */
int nco_map(const char* form, int n, const nco_tensor* args, nco_tensor* r);

/* %coerce of one scalar into a 1-element array of dtype (src/1type.lisp:673-694). */
int nco_coerce_f64(double v, int dtype, void* out_slot8);
int nco_coerce_i64(int64_t v, int dtype, void* out_slot8);

/* element access helpers for the python harness (logical values <-> SBCL storage) */
size_t nco_storage_bytes(int dtype, int64_t n);   /* padded to 8 elements, 1instantiate.lisp:81-84 */
int nco_pack_i64(const int64_t* v, int64_t n, int dtype, void* storage);   /* stores via %coerce */
int nco_unpack_i64(const void* storage, int64_t n, int dtype, int64_t* v);

/* ---- typed fast loops: the "C transcription of numcl's loop nest, 1 thread" used as the
 * reported CPU baseline and as the checker at medium sizes.  Each is the nest the generic
 * interpreter above would run, specialised on element type exactly as SBCL's `specializing`
 * does; tests/test_oracle_fast.py checks them against the interpreter. ------------------ */
/* (numcl:+ a b): two passes, t = 0 + a; r = t + b  (src/4numeric.lisp:528, 349-356). */
void nco_fast_add2_f32(const float* a, const float* b, float* tmp, float* r,
                       int64_t rows, int64_t cols /* b is (cols) */);
/* (numcl:* x y) with x ub8 (A,B,C,D), y fixnum (D): t = 1 * x (ub8); r = t * y (fixnum, tagged) */
void nco_fast_mul2_u8_fix(const uint8_t* x, const int64_t* y_tagged, uint8_t* tmp,
                          int64_t* r_tagged, int64_t outer, int64_t inner);
/* (sum a :axes 1) and (sum a) for a (rows,cols) f32: sequential row-major (5reduce.lisp:78-87) */
void nco_fast_sum_axis1_f32(const float* a, float* r, int64_t rows, int64_t cols);
float nco_fast_sum_all_f32(const float* a, int64_t n);
/* (einsum '(ij jk -> ik)) i,j,k nest with A[i,j] hoisted and C[i,k] loaded+stored per k
 * (common-lisp.lisp:94-105,235-240).  c must be zeroed (or hold data to accumulate into). */
void nco_fast_matmul_f64(const double* a, const double* b, double* c, int64_t m, int64_t k, int64_t n);
void nco_fast_matmul_f32(const float* a, const float* b, float* c, int64_t m, int64_t k, int64_t n);
/* rows [r0,r1) of the same product only (sampled check at N=8192). */
void nco_fast_matmul_rows_f64(const double* a, const double* b, double* c_rows,
                              int64_t r0, int64_t r1, int64_t k, int64_t n);
void nco_fast_matmul_rows_f32(const float* a, const float* b, float* c_rows,
                              int64_t r0, int64_t r1, int64_t k, int64_t n);
/* (einsum '(bij bjk -> bik)) b,i,j,k nest */
void nco_fast_bmm_f32(const float* a, const float* b, float* c, int64_t nb, int64_t m, int64_t k, int64_t n);
/* (einsum '(abcd -> dbca)) on 64-bit words: out += in with out zero => copy; nest order is
 * sort-locality's for this spec (a,b,c,d by construction of the test) */
void nco_fast_transpose_abcd_dbca_w64(const int64_t* in_tagged, int64_t* out_tagged,
                                      int64_t A, int64_t B, int64_t C, int64_t D);

/* counter-based generator shared with the CUDA fill kernel:
 * splitmix64(seed ^ index). dist as in ncl_fill_random. */
int nco_fill_random(nco_tensor* t, uint64_t seed, int dist, double lo, double hi);
int nco_fill_random_at(nco_tensor* t, uint64_t seed, uint64_t first, int dist, double lo, double hi);

#ifdef __cplusplus
}
#endif
#endif
