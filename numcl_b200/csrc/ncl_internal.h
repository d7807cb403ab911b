// ncl_internal.h -- host-side internals of libnumcl_cuda.so (not part of the ABI).
//
// Every public entry point lowers its arguments to ONE intermediate form, the Nest: the loop
// nest the reference's generators would have emitted (loops outer->inner with extents, one
// element-stride per array per loop, which loops are contracted for each output, and a
// statically typed TRANSFORM expression).  The planner (ncl_plan.cu) then picks a kernel family:
//   K1/K2/K6 elementwise (k_elementwise.cu), K3 reductions (k_reduce.cu), K5 permutation
//   (k_transpose.cu), K4 contractions (k_gemm_*.cu), or the reference-ordered generic nest kernel
//   (k_generic.cu) for everything else on the whitelist.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include "../../include/numcl_cuda.h"

#define NCL_MAX_LOOPS 24          // einsum indices + expanded broadcast-block axes
#define NCL_MAX_ARR   9           // arrays in one nest: outputs + inputs

// ---- value classes: the three machine representations a Lisp real takes inside a kernel -------
enum NclCls { CLS_I = 0, CLS_F = 1, CLS_D = 2, CLS_R = 3 /* ratio: int/int before %coerce */, CLS_CF = 4 /* complex single */, CLS_CD = 5 /* complex double */ };
inline bool cls_complex(int c) { return c == CLS_CF || c == CLS_CD; }

struct DtInfo { int slot_bits, value_bits, is_signed, tagged, cls; const char* name; };
extern const DtInfo g_dt[NCL_DTYPE_COUNT];

// ---- statically typed expression ----------------------------------------------------------------
struct XNode { int op, a, b, cls; int64_t ival; double fval; };
struct Expr {
  int n = 0;
  XNode node[NCL_MAX_EXPR];
};

struct Operand {
  char*   base = nullptr;          // device pointer to the base vector
  int     dtype = 0;
  int64_t offset = 0;              // element offset of the nest origin
  int64_t stride[NCL_MAX_LOOPS];   // elements, per loop
  int64_t n_storage = 0;           // elements addressable in the base vector (bounds check)
  int     stage_item = -1;         // host-resident tensor: index of its staging item (ncl_api.cu: Stager), else -1
};

struct Nest {
  int nloops = 0;
  int64_t extent[NCL_MAX_LOOPS];
  int n_in = 0, n_out = 0;
  Operand in[NCL_MAX_OPERANDS];
  Operand out[NCL_MAX_OPERANDS];
  Expr transform[NCL_MAX_OPERANDS];     // one per output, typed
  bool fresh = false;                   // outputs are not read: they start from zero / identity
  // for reductions through ncl_reduce with fresh=true the accumulator starts from this identity
  bool has_identity = false; int64_t ident_i = 0; double ident_f = 0;
  // fused statistics (ncl_reduce_stat): the reduction kernels divide the finished sum by stat_count (and take the square root)
  // in their final store; only the reduction families understand this, every other family refuses such a nest
  int stat = 0; int64_t stat_count = 0;
};

// ---- context ------------------------------------------------------------------------------------
struct ncl_buf { void* ptr; size_t bytes; int dev_slot; bool owned; };

struct Context {
  bool inited = false;
  int ndev = 0; int devs[16];
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  const char* last_kernel = "";
  bool force_generic = false;
  bool capturing = false; int64_t capture_launch0 = 0;   // ncl_capture_begin .. ncl_capture_end
  uint64_t h2d_bytes = 0, d2h_bytes = 0;                 // bytes the library itself copied over PCIe (ncl_transfer_stats)
  // scratch: reduction partials, GEMM operand splits, host staging
  void* scratch = nullptr; size_t scratch_bytes = 0;
  void* scratch2 = nullptr; size_t scratch2_bytes = 0;   // host staging arena: lives for the whole call
  void* scratch3 = nullptr; size_t scratch3_bytes = 0;   // intermediate rows of multi-pass reductions
  unsigned int* ticket = nullptr;       // zero-initialised counters for last-block reductions
  void* consts = nullptr;               // [sb64 0][sb64 1]
  unsigned int* fault = nullptr;        // sticky DOMAIN flag written by kernels (ncl_fn.cuh); checked at sync points when fault_possible
  bool fault_possible = false;
  void* pinned = nullptr; size_t pinned_bytes = 0;
  // panel pipeline of host-resident tensors (ncl_api.cu: Stager::try_pipeline): copy streams and a grow-only event pool
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t* events = nullptr; int n_events = 0;
  int64_t pipelined_calls = 0, pipelined_panels = 0;
};
Context& ctx();
std::recursive_mutex& ncl_mutex();          // the library's one lock (ncl_api.cu): every entry point that touches the context takes it
int  set_error(int code, const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
#define NCL_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)
#define NCL_TRY(call) do { int r__ = (call); if (r__ != NCL_OK) return r__; } while (0)
int  ensure_scratch(size_t bytes);       // ctx().scratch  >= bytes
int  ensure_scratch2(size_t bytes);
int  ensure_scratch3(size_t bytes);
// ---- programmatic dependent launch (PDL): consecutive INDEPENDENT kernels overlap their tail and ramp ----------------
// The streaming reduction kernels execute griddepcontrol.launch_dependents when they start and griddepcontrol.wait before
// their first write.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, such a kernel may begin READING its
// inputs while the previous kernel of the stream is still draining (its blocks fill the SMs the predecessor's tail leaves
// idle), but writes nothing until every earlier kernel has completed.  That is only legal when its inputs are not written
// by a kernel that may still be in flight, so the host keeps the byte ranges written by every launch since the last fully
// serialised one; a launch whose reads touch none of them gets the attribute, anything else is launched normally.  Every
// launch site that does not take part (note_launch) marks the state unknown, which forces the next launch to serialise.
struct PdlRange { const char* lo; const char* hi; };
struct PdlTracker {
  bool unknown = true; int n = 0; PdlRange w[16];
  bool enabled = true;
};
PdlTracker& pdl();
bool pdl_eligible(const PdlRange* reads, int nreads);
void pdl_commit(bool launched_pdl, const PdlRange* writes, int nwrites);
int  pdl_launch(const void* kfn, dim3 grid, dim3 block, void** args, cudaStream_t stream, bool with_attribute);
inline PdlRange pdl_range(const Operand& o) { PdlRange r; r.lo = o.base; r.hi = o.base + (size_t)((o.n_storage * g_dt[o.dtype].slot_bits + 7) / 8); return r; }
inline void note_launch(const char* name, int n = 1) { ctx().launches += n; ctx().last_kernel = name; pdl().unknown = true; }
// a kernel on the communication stream: counted, but it does not sit between two kernels of the compute stream
inline void note_side_launch(const char* name) { ctx().launches += 1; ctx().last_kernel = name; }

// ---- expression helpers (ncl_plan.cu) -------------------------------------------------------------
int  type_expr(const ncl_expr* src, const int* in_dtype, int n_in, const int* out_dtype, int n_out,
               int self_out, Expr* dst);
bool expr_reads_output(const Expr& e);
// functions whose kernels may raise the DOMAIN flag (NCL_ERR_DOMAIN)
inline bool op_may_fault(int op) {
  switch (op) {
    case NCL_OP_SQRT: case NCL_OP_LOG: case NCL_OP_LOG2: case NCL_OP_ASIN: case NCL_OP_ACOS: case NCL_OP_ACOSH: case NCL_OP_ATANH: case NCL_OP_ISQRT:
    case NCL_OP_FLOOR: case NCL_OP_CEILING: case NCL_OP_ROUND: case NCL_OP_TRUNCATE: case NCL_OP_MOD: case NCL_OP_REM: case NCL_OP_EXPT: return true;
    default: return false;
  }
}
// residency table (ncl_api.cu): device mirror of a registered host base vector (uploading it first when the host copy is newer), or nullptr
void* resident_mirror(const void* host, size_t* bytes);
int check_fault();                       // ncl_api.cu: reads and clears the flag; NCL_OK or NCL_ERR_DOMAIN (synchronises the stream)

// ---- kernel families; each returns NCL_OK, an error, or NCL_ERR_UNSUPPORTED to fall through ------
int launch_generic(const Nest& n);                        // k_generic.cu
int try_elementwise(const Nest& n);                       // k_elementwise.cu
int try_reduce(const Nest& n);                            // k_reduce.cu
int try_dot_reduce(const Nest& n);                        // k_reduce.cu: contractions without an M x N structure
int try_transpose(const Nest& n);                         // k_transpose.cu
int try_gemm(const Nest& n);                              // k_gemm.cu (dispatch to f64/tf32/simt)
int run_nest(Nest& n);                                    // ncl_plan.cu: canonicalise + dispatch

int fill_const(void* base, int dtype, int64_t offset, int64_t n, bool is_float, int64_t iv, double fv);   // k_fill.cu
int fill_random(void* base, int dtype, int64_t offset, int64_t n, uint64_t seed, uint64_t first_index, int dist, double lo, double hi);

// full reduce of a contiguous array into a float64/int64 partial on the device (k_reduce.cu):
// used by ncl_reduce_all_sharded before the all-reduce.
int reduce_all_to_partial(int op, const Operand& a, int64_t n, void* partial8);
int finalize_partial(int op, const void* partial8, int cls, Operand& r, bool fresh);
int reduce_all_p2p(int op, const Operand& a, int64_t n, const Operand& r, unsigned long long* const* peers, int rank, int world,
                   unsigned long long seq);
// numcl's identity of a reduction into element type `dtype` (src/5reduce.lisp:103-126, src/1constants.lisp:30-89): sum 0,
// prod 1, amax most-negative-value, amin most-positive-value.  ONE definition for ncl_reduce, the sharded reduce and
// both of its finalize paths.
void reduce_identity(int op, int dtype, long long* ident_i, double* ident_f);
