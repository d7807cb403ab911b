/*
 * numcl_cuda.h -- C ABI of libnumcl_cuda.so, the B200 (sm_100a) backend for numcl's
 * einsum / broadcast / reduce hot path.
 *
 * Every entry point replaces (or is bound by) one reference interface; the reference
 * location is cited as file:line relative to the numcl source tree.  The reference-side
 * CFFI stubs a maintainer would add are shown in INTEGRATION.md and lisp/.
 *
 * Conventions
 *   - plain C, no C++/torch types; all sizes are explicit.
 *   - every int-returning function returns NCL_OK (0) or a negative ncl_status; the message
 *     is available per thread from ncl_last_error().  Nothing aborts, nothing throws.
 *   - there is NO CPU fallback: an op/dtype/transform the CUDA backend does not implement
 *     returns NCL_ERR_UNSUPPORTED.
 *   - calls are asynchronous on the library stream unless stated; ncl_sync() or any d2h
 *     copy orders the host after them.  Host-resident tensors (ncl_tensor.host != NULL)
 *     are staged inside the call and the call returns after the results are back on the host.
 */
#ifndef NUMCL_CUDA_H
#define NUMCL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCL_MAX_RANK     8   /* axes per array                                          */
#define NCL_MAX_INDEX    16  /* distinct einsum indices (iter-specs length)             */
#define NCL_MAX_OPERANDS 8   /* inputs (and, separately, outputs) per einsum            */
#define NCL_MAX_EXPR     48  /* nodes per TRANSFORM expression                          */
#define NCL_MAX_OPTS     8   /* iteration-option entries per array                      */

typedef enum {
  NCL_OK = 0,
  NCL_ERR_INVALID     = -1,  /* malformed descriptor / argument                         */
  NCL_ERR_SHAPE       = -2,  /* dims do not unify / not broadcast compatible
                                (reference: assert at src/3einsum.lisp:349-353,413-416,
                                 src/4numeric.lisp:217-219)                              */
  NCL_ERR_UNSUPPORTED = -3,  /* outside the CUDA whitelist (no CPU fallback)             */
  NCL_ERR_CUDA        = -4,  /* CUDA runtime / driver error, text in ncl_last_error()    */
  NCL_ERR_NOMEM       = -5,
  NCL_ERR_STATE       = -6,  /* ncl_init not called, wrong device, ...                   */
  NCL_ERR_DOMAIN      = -7   /* an element left the real domain of the function applied to it, where the
                                reference would have signalled a Lisp error or produced a complex number:
                                sqrt / log / log2 of a negative, asin / acos outside [-1,1], acosh below 1,
                                atanh outside [-1,1], isqrt of a negative, integer division / mod / rem by
                                zero, (expt 0 negative), an integer power that does not fit 64 bits.
                                Kernels record it in a sticky device flag; the next ncl_sync, ncl_d2h or
                                call with host-resident tensors reports it (and clears it).  The arrays
                                written by the offending call hold NaN / 0 in those elements.            */
} ncl_status;

/* Element types = SBCL's specialized array element types, in SBCL's storage format
 * (doc/DETAILS.org:74-97).  Sub-byte types are packed LSB-first.
 * FIXNUM and UB62 hold *tagged* words (value << 1); SB64/UB64 hold raw words.
 * %coerce (src/1type.lisp:673-694) wraps a stored integer to the type's bit width:
 * UB7 wraps mod 2^7 although it lives in an 8-bit slot, FIXNUM wraps to 63 bits signed. */
typedef enum {
  NCL_BIT = 0,   /* (unsigned-byte 1), 1-bit slots   */
  NCL_UB2,       /* 2-bit slots                      */
  NCL_UB4,       /* 4-bit slots                      */
  NCL_UB7,       /* 8-bit slots, 7 value bits        */
  NCL_UB8,
  NCL_UB15,      /* 16-bit slots                     */
  NCL_UB16,
  NCL_UB31,      /* 32-bit slots                     */
  NCL_UB32,
  NCL_UB62,      /* 64-bit slots, tagged             */
  NCL_UB63,      /* 64-bit slots, raw                */
  NCL_UB64,
  NCL_SB8,
  NCL_SB16,
  NCL_SB32,
  NCL_FIXNUM,    /* (signed-byte 63), tagged         */
  NCL_SB64,
  NCL_F32,       /* single-float                     */
  NCL_F64,       /* double-float                     */
  NCL_C64,       /* (complex single-float): 8-byte slots, real part first (src/1type.lisp:341-344: only float complexes exist) */
  NCL_C128,      /* (complex double-float): 16-byte slots */
  NCL_DTYPE_COUNT
} ncl_dtype;

/* Scalar functions allowed inside a TRANSFORM / as broadcast, map and reduce functions.
 * Semantics are Common Lisp's on SBCL x86-64: each float op is rounded separately
 * (no FMA contraction), integer ops are exact and wrapped only by the final %coerce. */
typedef enum {
  NCL_OP_ADD = 0, NCL_OP_SUB, NCL_OP_MUL, NCL_OP_DIV, NCL_OP_MAX, NCL_OP_MIN,
  /* comparisons -> 0/1 (src/4numeric.lisp:609-655) */
  NCL_OP_EQ, NCL_OP_NE, NCL_OP_LE, NCL_OP_GE, NCL_OP_LT, NCL_OP_GT,
  /* bitwise, integers only (src/4numeric.lisp:584-602) */
  NCL_OP_LOGAND, NCL_OP_LOGIOR, NCL_OP_LOGXOR,
  NCL_OP_LOGANDC1, NCL_OP_LOGANDC2, NCL_OP_LOGEQV, NCL_OP_LOGNAND, NCL_OP_LOGNOR, NCL_OP_LOGORC1, NCL_OP_LOGORC2,
  /* rounding family (src/4numeric.lisp:568-581): (floor number divisor) etc. -> the INTEGER quotient (first value);
   * mod / rem -> the remainder.  Integers: exact.  Floats: SBCL's definitions (truncate of the rounded quotient
   * number/divisor, then the adjustment by the sign of the remainder; numbers.lisp), each float op rounded on its own. */
  NCL_OP_FLOOR, NCL_OP_CEILING, NCL_OP_ROUND, NCL_OP_TRUNCATE, NCL_OP_MOD, NCL_OP_REM,
  /* (expt base power), src/4numeric.lisp:533: integer power -> SBCL's intexp (repeated squaring, exact for integers,
   * every multiplication rounded for floats; negative power = reciprocal); float power -> pow (tolerance only). */
  NCL_OP_EXPT,
  NCL_OP_BINARY_END,
  /* unary (src/4numeric.lisp:455-519) */
  NCL_OP_NEG = 32, NCL_OP_ABS, NCL_OP_SQUARE, NCL_OP_SQRT, NCL_OP_LOGNOT, NCL_OP_IDENTITY,
  NCL_OP_EXP, NCL_OP_LOG, NCL_OP_SIN, NCL_OP_COS, NCL_OP_TAN, NCL_OP_TANH, NCL_OP_SIGNUM,
  NCL_OP_ASIN, NCL_OP_ACOS, NCL_OP_ATAN, NCL_OP_SINH, NCL_OP_COSH, NCL_OP_ASINH, NCL_OP_ACOSH, NCL_OP_ATANH,
  NCL_OP_LOG2,                                   /* %log2 = (log x 2), src/4numeric.lisp:514-516 */
  NCL_OP_ISQRT, NCL_OP_LOGCOUNT, NCL_OP_INTEGER_LENGTH,   /* integers only, exact */
  /* complex numbers (src/4numeric.lisp:486-489; vdot conjugates its first argument, src/4linear-algebra2.lisp:54-59).
   * Complex operands are accepted by + - * / and by neg abs square identity conjugate realpart imagpart; every other
   * function of a complex value returns NCL_ERR_UNSUPPORTED. */
  NCL_OP_CONJUGATE, NCL_OP_REALPART, NCL_OP_IMAGPART,
  /* leaves of an expression */
  NCL_X_INPUT = 64,  /* $a : element of input  number a (1-based)  */
  NCL_X_OUTPUT,      /* @a : element of output number a (1-based)  */
  NCL_X_CONST_INT,   /* literal integer ival                        */
  NCL_X_CONST_F32,   /* literal single-float (float)fval            */
  NCL_X_CONST_F64    /* literal double-float fval                   */
} ncl_op;

/* One node of a TRANSFORM form, post-order; children are indices of earlier nodes.
 * N-ary Lisp forms (+ a b c) are flattened left-associatively by the caller. */
typedef struct {
  int32_t op;      /* ncl_op                                              */
  int32_t a, b;    /* child node indices; for NCL_X_INPUT/OUTPUT: a = n   */
  int32_t pad;
  int64_t ival;
  double  fval;
} ncl_expr_node;

typedef struct {
  int32_t n;                       /* node count; node[n-1] is the root */
  int32_t pad;
  ncl_expr_node node[NCL_MAX_EXPR];
} ncl_expr;

/* One entry of an iteration-option alist, (index :start s :end e :step k)
 * (src/3einsum.lisp:301-338, t/package.lisp:542-566).  `key` is what the normaliser left
 * in the car: the index NUMBER.  The reference looks it up by index number when sizing
 * loops (3einsum.lisp:307) and by AXIS POSITION when addressing (common-lisp.lisp:129);
 * the library reproduces both lookups. */
typedef struct {
  int32_t key;
  int32_t has_start, has_end;      /* 0 = default (0 / dimension)      */
  int32_t pad;
  int64_t start, end, step;        /* step defaults to 1               */
} ncl_iter_opt;

/* Serialised `einsum-specs` (src/3einsum.lisp:44-62), as produced by
 * einsum-normalize-subscripts (src/3einsum.lisp:64-200). */
typedef struct {
  int32_t n_in, n_out;
  int32_t n_iter;
  int32_t iter[NCL_MAX_INDEX];                         /* iter-specs; -1 = broadcast block */
  int32_t in_rank[NCL_MAX_OPERANDS];
  int32_t in_spec[NCL_MAX_OPERANDS][NCL_MAX_RANK];     /* i-specs; -1 = `-`                */
  int32_t out_rank[NCL_MAX_OPERANDS];
  int32_t out_spec[NCL_MAX_OPERANDS][NCL_MAX_RANK];    /* o-specs                          */
  int32_t in_nopt[NCL_MAX_OPERANDS];
  int32_t out_nopt[NCL_MAX_OPERANDS];
  ncl_iter_opt in_opt[NCL_MAX_OPERANDS][NCL_MAX_OPTS]; /* i-options                        */
  ncl_iter_opt out_opt[NCL_MAX_OPERANDS][NCL_MAX_OPTS];/* o-options                        */
  ncl_expr transform[NCL_MAX_OPERANDS];                /* transforms, one per output       */
} ncl_einsum_desc;

typedef struct ncl_buf ncl_buf;    /* device buffer owned by the library */

/* A numcl array as the backend sees it: the 1-D base vector plus the displaced N-D header
 * (src/1instantiate.lisp:81-96).  Exactly one of `buf` (device resident) and `host`
 * (host resident, staged inside the call) is non-NULL. */
typedef struct {
  ncl_buf* buf;
  void*    host;
  int32_t  dtype;                  /* ncl_dtype                                  */
  int32_t  rank;
  int64_t  dims[NCL_MAX_RANK];
  int64_t  elem_offset;            /* displaced-index-offset, in elements        */
} ncl_tensor;

/* ---- flags for ncl_einsum / ncl_reduce ------------------------------------------------ */
#define NCL_OUT_ACCUMULATE 0u   /* default: outputs hold caller data and are accumulated into,
                                   like a caller-supplied output (src/3einsum.lisp:468)       */
#define NCL_OUT_FRESH      1u   /* outputs are freshly created: the library treats them as the
                                   `zeros` the wrapper would have made (3einsum.lisp:468) /
                                   the `full` of the identity (5reduce.lisp:56-59) WITHOUT
                                   reading them, so no host/device fill pass is needed       */

/* ---- context ---------------------------------------------------------------------------- */
/* Bind this process to its CUDA device: one process per GPU, so ndev must be 1 (devs[0] = the
 * device ordinal); ndev > 1 returns NCL_ERR_UNSUPPORTED. */
int  ncl_init(int ndev, const int* devs);
void ncl_shutdown(void);
const char* ncl_last_error(void);
const char* ncl_version(void);
int  ncl_device_count(void);
/* Run subsequent work on a caller-owned cudaStream_t (e.g. torch's current stream). NULL
 * restores the library's own stream. */
int  ncl_set_stream(void* cuda_stream);
void* ncl_get_stream(void);
int  ncl_sync(void);
/* The library caches freed device blocks for re-use (every numcl operation returns a fresh array,
 * src/1instantiate.lisp:81-96); ncl_trim gives the cache back to the driver. */
int  ncl_trim(void);
/* Bytes the library itself has copied host->device / device->host since ncl_init (ncl_h2d, ncl_d2h,
 * staging of host-resident tensors, the residency table below): lets a caller verify that a warm
 * call moved nothing over PCIe. */
void ncl_transfer_stats(uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Calls on pinned host-resident arrays are cut into panels of their outermost kept loop: panel p is uploaded while
 * panel p-1 is computed and panel p-2 travels back (PCIe is full duplex, the copy engines run beside the SMs).  Number of
 * calls that ran that way and the panels they were cut into, since ncl_init.  NUMCL_NO_PIPELINE=1 disables it,
 * NUMCL_PANEL_MB sets the panel size (default 32). */
void ncl_pipeline_stats(int64_t* calls, int64_t* panels);
/* Number of kernels the library has launched since ncl_init (bench.py's gpu_launches). */
int64_t ncl_launch_count(void);
/* Name of the kernel family the last compute call dispatched to ("bcast_vec", "bcast_vec_flat", "bcast_widen",
 * "reduce_row_block", "reduce_rows_reg", "reduce_fewcols", "reduce_all", "reduce_col", "dot_all", "dot_rows",
 * "dot_matvec", "transpose_tile", "transpose_reg", "gemm_f64_dmma", "gemm_tf32x3_tcgen05", "generic_nest", "map_reduce", ...). */
const char* ncl_last_kernel(void);
/* Force the generic (thread-per-output, reference-ordered) nest kernel: test hook. */
int  ncl_force_generic(int on);

/* ---- CUDA graphs: a launch-bound sequence of calls recorded once and replayed with one launch --
 * (no reference counterpart: the reference compiles its loop nests with `compile` and caches them,
 * src/3einsum.lisp:30-31; the GPU analogue of that cache for a whole call sequence is a graph).
 * Between begin and end every compute call on DEVICE-resident tensors is recorded, not run; host
 * tensors, ncl_h2d/ncl_d2h/ncl_sync and anything that must allocate return NCL_ERR_STATE: run the
 * sequence once eagerly first.  The sharded reduce (mailbox exchange on the communication stream)
 * is captured too. */
typedef struct ncl_graph ncl_graph;
int     ncl_capture_begin(void);
int     ncl_capture_end(ncl_graph** out);
int     ncl_graph_launch(ncl_graph* g);
int64_t ncl_graph_kernels(const ncl_graph* g);   /* kernels one replay launches */
int     ncl_graph_free(ncl_graph* g);

/* ---- buffers: what src/1instantiate.lisp:81-96 (%make-array) gains ----------------------- */
ncl_buf* ncl_alloc(size_t bytes, int dev_slot);
int      ncl_free(ncl_buf* b);
void*    ncl_buf_ptr(const ncl_buf* b);       /* raw device pointer (interop)               */
size_t   ncl_buf_bytes(const ncl_buf* b);
ncl_buf* ncl_wrap(void* device_ptr, size_t bytes, int dev_slot);  /* non-owning             */
int      ncl_h2d(ncl_buf* dst, size_t dst_off, const void* host, size_t bytes);
int      ncl_d2h(void* host, const ncl_buf* src, size_t src_off, size_t bytes);
void*    ncl_host_alloc_pinned(size_t bytes);
int      ncl_host_free(void* p);
/* ---- residency: a device mirror that outlives the call, attached to a host base vector ------------------------
 * (north_star: "3array.lisp gains pinned-host and device buffers behind numcl's specialized arrays"; the hook is
 * %make-array, src/1instantiate.lisp:81-96: register the fresh base vector, unregister from its finalizer.)
 * A host-resident ncl_tensor whose `host` equals a registered base is served from the mirror: it is uploaded only
 * when the host copy is newer (after registration without NCL_RESIDENT_FRESH, or after ncl_host_dirty), so repeated
 * calls on the same arrays move no input bytes over PCIe.  Outputs are written back into the host vector when the call
 * returns (only the tensor's own window), unless the vector was registered NCL_RESIDENT_LAZY: then the result stays
 * on the device -- usable as an input of later calls at no cost -- until ncl_host_fetch.
 *   ncl_host_dirty : the host wrote into the vector ((setf aref) ...): the next use uploads it again.
 *   ncl_host_state : 0 in sync, 1 host newer, 2 device newer, -1 not registered. */
#define NCL_RESIDENT_FRESH 1u   /* the vector holds nothing yet (a result array): skip the first upload          */
#define NCL_RESIDENT_LAZY  2u   /* results stay on the device until ncl_host_fetch / ncl_host_unregister        */
int ncl_host_register(void* base, size_t bytes, unsigned flags);
int ncl_host_unregister(void* base);      /* fetches a pending result first */
int ncl_host_dirty(void* base);
int ncl_host_fetch(void* base);
int ncl_host_state(const void* base);
/* numcl's base vectors are padded to a multiple of 8 elements (1instantiate.lisp:81-84). */
size_t   ncl_storage_bytes(int dtype, int64_t n_elements);

/* full / zeros / ones (src/2zeros.lisp:41-58): fill with %coerce(value, dtype). */
int ncl_fill(ncl_tensor* t, int value_is_float, int64_t ival, double fval);
/* Synthetic inputs: counter-based splitmix64(seed ^ linear index).
 *  dist 0: floats uniform in [lo,hi) / integers uniform in [lo,hi] (inclusive)
 *  dist 1: integers in [lo,hi] stored in the tensor's dtype (floats get exact integers)
 *  dist 2: as 0 for floats, but 2^-10 of the elements replaced by specials
 *          {+0,-0,+inf,-inf,min denormal,max finite/2} (set R of SURVEY 8d)              */
int ncl_fill_random(ncl_tensor* t, uint64_t seed, int dist, double lo, double hi);
/* Same stream, but element k of t is stream element first_index + k: a rank generates its own row
 * shard of a larger synthetic array. */
int ncl_fill_random_at(ncl_tensor* t, uint64_t seed, uint64_t first_index, int dist, double lo, double hi);

/* ---- compute ---------------------------------------------------------------------------- */
/* The form a (defmethod einsum-body ((*compiler* (eql :cuda)) ev)) returns calls this:
 * replaces einsum-body (eql :common-lisp), src/4einsum-backends/common-lisp.lisp:44-240,
 * invoked from einsum-lambda, src/3einsum.lisp:491. */
int ncl_einsum(const ncl_einsum_desc* desc,
               const ncl_tensor* in, int n_in,
               ncl_tensor* out, int n_out, unsigned flags);

/* Replaces broadcast / broadcast-core, src/4numeric.lisp:201-399: r = (op x y) with numpy
 * right-aligned broadcasting, stored through %coerce into r's dtype. */
int ncl_broadcast(int op, const ncl_tensor* x, const ncl_tensor* y, ncl_tensor* r);

/* Fused n-ary fold used by numcl:+ * - / max min (src/4numeric.lisp:528-542):
 *   use_identity=1: ((ident op a0) op a1) ... with ident = 0 (ADD) / 1 (MUL), exactly the
 *                   reference's (reduce ... :initial-value ident), in ONE pass;
 *   use_identity=0: (a0 op a1) op a2 ...   (numcl:- / max min with >= 2 args).
 * n = 1 with use_identity gives the unary forms (- x) = 0 - x, (/ x) = 1 / x. */
int ncl_fold(int op, int use_identity, const ncl_tensor* args, int n, ncl_tensor* r);

/* Same-shape n-ary map: map-array-into, src/4numeric.lisp:37-113 (exactly
 * array-total-size elements; the reference's off-by-one at :102-103 is not reproduced). */
int ncl_map(const ncl_expr* fn, const ncl_tensor* args, int n, ncl_tensor* r);

/* Replaces reduce-array / reduce-lambda, src/5reduce.lisp:42-87:
 * r[kept] = op(r[kept], a[all]) over `axes` (n_axes = 0: all axes).
 * flags & NCL_OUT_FRESH: r starts from the op's numcl identity (sum 0, prod 1,
 * amax most-negative-value, amin most-positive-value; 5reduce.lisp:103-126). */
int ncl_reduce(int op, const ncl_tensor* a, const int* axes, int n_axes,
               ncl_tensor* r, unsigned flags);

/* numcl:mean / variance / standard-deviation (src/5reduce.lisp:128-167) as one fused launch: the sum kernels' final store
 * divides by the element count (then takes the square root); bit-identical to the reference's composition
 * (numcl:sqrt (numcl:/ (numcl:sum [(square a)] :axes axes) count)).  variance is the reference's mean of SQUARES.
 * r: single-float for single-float and integer arrays (integer / integer is a single-float, 2typeinfer.lisp:134-139),
 * double-float for double-float arrays. */
#define NCL_STAT_MEAN     1
#define NCL_STAT_VARIANCE 2
#define NCL_STAT_STDEV    3
int ncl_reduce_stat(int stat, const ncl_tensor* a, const int* axes, int n_axes, ncl_tensor* r);

/* astype / copy (src/3copy.lisp:27-43): r = %coerce(a, r.dtype), same shape. */
int ncl_convert(const ncl_tensor* a, ncl_tensor* r);

/* ---- multi-GPU: combine per-GPU reduction partials (no reference counterpart) ----------- */
/* NCCL is loaded lazily with dlopen; the unique id travels through the caller's own
 * rendezvous (torch.distributed / MPI / a file). */
int ncl_comm_unique_id(void* id128);                       /* rank 0: fill 128 bytes       */
int ncl_comm_init(const void* id128, int rank, int world); /* all ranks                    */
int ncl_comm_destroy(void);
/* In-place all-reduce of t with op in {ADD, MUL, MAX, MIN}. */
int ncl_allreduce(int op, ncl_tensor* t);
/* Full reduce of a row-sharded array: local two-phase reduce to a float64/int64 partial,
 * all-reduce of that one element, one final rounding into r (rank-0 tensor, every rank). */
int ncl_reduce_all_sharded(int op, const ncl_tensor* a_local, ncl_tensor* r);
/* Same, but the exchange and the final rounding are issued on the library's communication stream
 * and overlap with later work on the compute stream; ncl_comm_join() (or ncl_sync / any d2h)
 * orders the compute stream after them.  With NVLink peer access the local kernel's last block
 * stores the partial into every rank's mailbox (P2P stores) and a one-CTA kernel on the
 * communication stream collects them; without it the partial goes through ncclAllReduce. */
int ncl_reduce_all_sharded_async(int op, const ncl_tensor* a_local, ncl_tensor* r);
int ncl_comm_join(void);

#ifdef __cplusplus
}
#endif
#endif /* NUMCL_CUDA_H */
