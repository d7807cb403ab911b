// ncl_plan.cu -- static typing of TRANSFORM expressions and kernel-family dispatch.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ncl_internal.h"

const DtInfo g_dt[NCL_DTYPE_COUNT] = {
  {1, 1, 0, 0, CLS_I, "bit"},   {2, 2, 0, 0, CLS_I, "ub2"},   {4, 4, 0, 0, CLS_I, "ub4"},
  {8, 7, 0, 0, CLS_I, "ub7"},   {8, 8, 0, 0, CLS_I, "ub8"},   {16, 15, 0, 0, CLS_I, "ub15"},
  {16, 16, 0, 0, CLS_I, "ub16"}, {32, 31, 0, 0, CLS_I, "ub31"}, {32, 32, 0, 0, CLS_I, "ub32"},
  {64, 62, 0, 1, CLS_I, "ub62"}, {64, 63, 0, 0, CLS_I, "ub63"}, {64, 64, 0, 0, CLS_I, "ub64"},
  {8, 8, 1, 0, CLS_I, "sb8"},   {16, 16, 1, 0, CLS_I, "sb16"}, {32, 32, 1, 0, CLS_I, "sb32"},
  {64, 63, 1, 1, CLS_I, "fixnum"}, {64, 64, 1, 0, CLS_I, "sb64"},
  {32, 0, 1, 0, CLS_F, "f32"},  {64, 0, 1, 0, CLS_D, "f64"},
  {64, 0, 1, 0, CLS_CF, "c64"}, {128, 0, 1, 0, CLS_CD, "c128"},
};

// ---- PDL tracker (ncl_internal.h) ----------------------------------------------------------------------------
static PdlTracker g_pdl;
PdlTracker& pdl() {
  static bool init = false;
  if (!init) { init = true; const char* e = getenv("NUMCL_PDL"); if (e && atoi(e) == 0) g_pdl.enabled = false; }
  return g_pdl;
}
bool pdl_eligible(const PdlRange* reads, int nreads) {
  PdlTracker& t = pdl();
  if (!t.enabled || t.unknown || t.n >= 14) return false;
  for (int i = 0; i < nreads; i++)
    for (int j = 0; j < t.n; j++)
      if (reads[i].lo < t.w[j].hi && t.w[j].lo < reads[i].hi) return false;
  return true;
}
void pdl_commit(bool launched_pdl, const PdlRange* writes, int nwrites) {
  PdlTracker& t = pdl();
  if (!launched_pdl) t.n = 0;                                  // a serialised launch: everything before it has completed when it starts
  for (int i = 0; i < nwrites && t.n < 16; i++) t.w[t.n++] = writes[i];
  t.unknown = false;
}
int pdl_launch(const void* kfn, dim3 grid, dim3 block, void** args, cudaStream_t stream, bool with_attribute) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = with_attribute ? 1 : 0;
  return cuda_fail(cudaLaunchKernelExC(&cfg, kfn, args), "cudaLaunchKernelExC");
}

// ---- expression typing: Common Lisp contagion, decided once on the host ----------------------------
static int contagion(int a, int b) {
  if (cls_complex(a) || cls_complex(b)) {                       // float-contagion, src/1type.lisp:446-451: (complex <contagion of the parts>)
    if (a == CLS_R || b == CLS_R) return CLS_R;
    return (a == CLS_CD || b == CLS_CD || a == CLS_D || b == CLS_D) ? CLS_CD : CLS_CF;
  }
  if (a == CLS_D || b == CLS_D) return CLS_D;
  if (a == CLS_F || b == CLS_F) return CLS_F;
  if (a == CLS_R || b == CLS_R) return CLS_R;
  return CLS_I;
}

// ub64 words at or above 2^63 do not fit the signed 64-bit lanes the integer class computes in.  Functions that only need
// the value modulo 2^64 (+ - * and the bitwise family, whose results are wrapped by %coerce anyway) are exact regardless;
// anything that looks at the ORDER or MAGNITUDE of such a value (max min comparisons / abs division conversion to float
// rounding ...) would silently be wrong, so it is refused.
static bool mod64_safe(int op) {
  switch (op) {
    case NCL_OP_ADD: case NCL_OP_SUB: case NCL_OP_MUL: case NCL_OP_LOGAND: case NCL_OP_LOGIOR: case NCL_OP_LOGXOR:
    case NCL_OP_LOGANDC1: case NCL_OP_LOGANDC2: case NCL_OP_LOGEQV: case NCL_OP_LOGNAND: case NCL_OP_LOGNOR: case NCL_OP_LOGORC1: case NCL_OP_LOGORC2:
    case NCL_OP_LOGNOT: case NCL_OP_NEG: case NCL_OP_SQUARE: case NCL_OP_IDENTITY: return true;
    default: return false;
  }
}

int type_expr(const ncl_expr* src, const int* in_dtype, int n_in, const int* out_dtype, int n_out,
              int self_out, Expr* dst) {
  if (src->n <= 0 || src->n > NCL_MAX_EXPR) return set_error(NCL_ERR_INVALID, "transform has %d nodes", src->n);
  dst->n = src->n;
  bool u64[NCL_MAX_EXPR];
  for (int k = 0; k < src->n; k++) {
    const ncl_expr_node& s = src->node[k]; XNode& x = dst->node[k];
    x.op = s.op; x.a = s.a; x.b = s.b; x.ival = s.ival; x.fval = s.fval; x.cls = CLS_I;
    u64[k] = false;
    switch (s.op) {
      case NCL_X_INPUT:
        if (s.a < 1 || s.a > n_in) return set_error(NCL_ERR_INVALID, "transform refers to $%d but there are %d inputs", s.a, n_in);
        x.cls = g_dt[in_dtype[s.a - 1]].cls; u64[k] = in_dtype[s.a - 1] == NCL_UB64; break;
      case NCL_X_OUTPUT:
        if (s.a < 1 || s.a > n_out) return set_error(NCL_ERR_INVALID, "transform refers to @%d but there are %d outputs", s.a, n_out);
        x.cls = g_dt[out_dtype[s.a - 1]].cls; u64[k] = out_dtype[s.a - 1] == NCL_UB64; break;
      case NCL_X_CONST_INT: x.cls = CLS_I; break;
      case NCL_X_CONST_F32: x.cls = CLS_F; break;
      case NCL_X_CONST_F64: x.cls = CLS_D; break;
      default: {
        if (s.a < 0 || s.a >= k) return set_error(NCL_ERR_INVALID, "transform node %d: bad child", k);
        int ca = dst->node[s.a].cls;
        bool tainted = u64[s.a];
        if (s.op >= 0 && s.op < NCL_OP_BINARY_END) {
          if (s.b < 0 || s.b >= k) return set_error(NCL_ERR_INVALID, "transform node %d: bad child", k);
          int cb = dst->node[s.b].cls;
          tainted = tainted || u64[s.b];
          int cc = contagion(ca, cb);
          if (cc == CLS_R) return set_error(NCL_ERR_UNSUPPORTED, "arithmetic on an unreduced integer ratio is outside the CUDA whitelist");
          if (cls_complex(cc) && !(s.op == NCL_OP_ADD || s.op == NCL_OP_SUB || s.op == NCL_OP_MUL || s.op == NCL_OP_DIV))
            return set_error(NCL_ERR_UNSUPPORTED, "function code %d of a complex number is outside the CUDA whitelist (+ - * / only)", s.op);
          if (s.op >= NCL_OP_LOGAND && s.op <= NCL_OP_LOGORC2) { if (ca != CLS_I || cb != CLS_I) return set_error(NCL_ERR_INVALID, "bitwise function on non-integers"); x.cls = CLS_I; }
          else if (s.op >= NCL_OP_EQ && s.op <= NCL_OP_GT) x.cls = CLS_I;
          else if (s.op >= NCL_OP_FLOOR && s.op <= NCL_OP_TRUNCATE) x.cls = CLS_I;           // the integer quotient
          else if (s.op == NCL_OP_MOD || s.op == NCL_OP_REM) x.cls = cc;
          else if (s.op == NCL_OP_EXPT) x.cls = (cc == CLS_I) ? CLS_R : cc;                   // integer ^ negative integer is a ratio
          else if (s.op == NCL_OP_DIV && cc == CLS_I) x.cls = CLS_R;
          else x.cls = cc;
        } else {
          if (cls_complex(ca) && !(s.op == NCL_OP_NEG || s.op == NCL_OP_ABS || s.op == NCL_OP_SQUARE || s.op == NCL_OP_IDENTITY ||
                                   s.op == NCL_OP_CONJUGATE || s.op == NCL_OP_REALPART || s.op == NCL_OP_IMAGPART))
            return set_error(NCL_ERR_UNSUPPORTED, "function code %d of a complex number is outside the CUDA whitelist", s.op);
          switch (s.op) {
            case NCL_OP_CONJUGATE:                                   // the identity on reals
              if (ca == CLS_R) return set_error(NCL_ERR_UNSUPPORTED, "function of an integer ratio");
              x.cls = ca; break;
            case NCL_OP_REALPART: case NCL_OP_IMAGPART:              // complex-part-inferer, src/2typeinfer.lisp:283-296: the part's type
              if (ca == CLS_R) return set_error(NCL_ERR_UNSUPPORTED, "function of an integer ratio");
              x.cls = ca == CLS_CF ? CLS_F : ca == CLS_CD ? CLS_D : ca; break;
            case NCL_OP_NEG: case NCL_OP_ABS: case NCL_OP_SQUARE: case NCL_OP_SIGNUM: case NCL_OP_IDENTITY:
              if (ca == CLS_R && s.op != NCL_OP_IDENTITY) return set_error(NCL_ERR_UNSUPPORTED, "function of an integer ratio");
              x.cls = (s.op == NCL_OP_ABS && cls_complex(ca)) ? (ca == CLS_CF ? CLS_F : CLS_D) : ca; break;
            case NCL_OP_LOGNOT: case NCL_OP_ISQRT: case NCL_OP_LOGCOUNT: case NCL_OP_INTEGER_LENGTH:
              if (ca != CLS_I) return set_error(NCL_ERR_INVALID, "integer function of a non-integer"); x.cls = CLS_I; break;
            case NCL_OP_SQRT: case NCL_OP_EXP: case NCL_OP_LOG: case NCL_OP_SIN: case NCL_OP_COS: case NCL_OP_TAN: case NCL_OP_TANH:
            case NCL_OP_ASIN: case NCL_OP_ACOS: case NCL_OP_ATAN: case NCL_OP_SINH: case NCL_OP_COSH: case NCL_OP_ASINH: case NCL_OP_ACOSH:
            case NCL_OP_ATANH: case NCL_OP_LOG2:
              x.cls = (ca == CLS_D) ? CLS_D : CLS_F; break;      // float substitution: rational -> single-float
            default:
              return set_error(NCL_ERR_UNSUPPORTED, "function code %d is not on the CUDA whitelist", s.op);
          }
        }
        if (tainted) {
          if (!mod64_safe(s.op) || x.cls != CLS_I)
            return set_error(NCL_ERR_UNSUPPORTED, "an (unsigned-byte 64) operand reaches a function that depends on its order or magnitude (function code %d): "
                                                  "values >= 2^63 do not fit the signed 64-bit lanes; only + - * and the bitwise family are exact", s.op);
          u64[k] = true;
        }
      }
    }
  }
  if (self_out >= 0 && self_out < n_out && cls_complex(dst->node[src->n - 1].cls) && !cls_complex(g_dt[out_dtype[self_out]].cls))
    return set_error(NCL_ERR_INVALID, "a complex value cannot be stored into a %s array", g_dt[out_dtype[self_out]].name);
  if (u64[src->n - 1] && self_out >= 0 && self_out < n_out && g_dt[out_dtype[self_out]].cls != CLS_I)
    return set_error(NCL_ERR_UNSUPPORTED, "an (unsigned-byte 64) value is stored into a float array: values >= 2^63 do not fit the signed 64-bit lanes");
  return NCL_OK;
}

bool expr_reads_output(const Expr& e) {
  for (int k = 0; k < e.n; k++) if (e.node[k].op == NCL_X_OUTPUT) return true;
  return false;
}

// ---- dispatch -------------------------------------------------------------------------------------
static int check_bounds(const Operand& a, const Nest& n, const char* what, int idx) {
  long long lo = a.offset, hi = a.offset;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] <= 0) return NCL_OK;       // empty nest touches nothing
    long long span = (n.extent[l] - 1) * a.stride[l];
    if (span > 0) hi += span; else lo += span;
  }
  if (lo < 0 || hi >= a.n_storage)
    return set_error(NCL_ERR_SHAPE, "%s %d would be addressed at elements %lld..%lld of a %lld-element base vector",
                     what, idx, lo, hi, (long long)a.n_storage);
  return NCL_OK;
}

// ---- complex element types -------------------------------------------------------------------------------------------
// A (complex single-float) array is pairs (real, imaginary) of floats.  Wherever every array of a nest is complex of ONE
// precision and the transforms only add, subtract, negate or copy -- array + array, array - array, (numcl:sum a :axes ..),
// transposes and copies -- real and imaginary parts never meet: the nest is the same nest over float arrays with one more
// innermost loop of extent 2, and runs on the real kernel families at their speed.  Anything that multiplies, divides,
// conjugates or mixes in a real operand or a constant (the real takes part in ONE component only) stays a complex nest and
// runs on the generic kernel, which carries complex values.
static bool lower_complex(Nest& n) {
  int prec = -1;
  for (int k = 0; k < n.n_in + n.n_out; k++) {
    const Operand& a = k < n.n_in ? n.in[k] : n.out[k - n.n_in];
    const int c = g_dt[a.dtype].cls;
    if (!cls_complex(c) || (prec >= 0 && c != prec)) return false;
    prec = c;
  }
  if (prec < 0 || n.nloops >= NCL_MAX_LOOPS || n.stat) return false;
  for (int o = 0; o < n.n_out; o++)
    for (int k = 0; k < n.transform[o].n; k++) {
      const int op = n.transform[o].node[k].op;
      if (!(op == NCL_X_INPUT || op == NCL_X_OUTPUT || op == NCL_OP_ADD || op == NCL_OP_SUB || op == NCL_OP_NEG || op == NCL_OP_IDENTITY)) return false;
    }
  if (n.has_identity && (n.ident_f != 0.0 || n.ident_i != 0)) return false;
  const int real_dt = prec == CLS_CF ? NCL_F32 : NCL_F64, real_cls = prec == CLS_CF ? CLS_F : CLS_D;
  const int L = n.nloops++;
  n.extent[L] = 2;
  for (int k = 0; k < n.n_in + n.n_out; k++) {
    Operand& a = k < n.n_in ? n.in[k] : n.out[k - n.n_in];
    a.dtype = real_dt; a.offset *= 2; a.n_storage *= 2;
    for (int l = 0; l < L; l++) a.stride[l] *= 2;
    a.stride[L] = 1;
  }
  for (int o = 0; o < n.n_out; o++) for (int k = 0; k < n.transform[o].n; k++) n.transform[o].node[k].cls = real_cls;
  return true;
}

// ---- bit arrays ---------------------------------------------------------------------------------------------------------
// Every numcl comparison returns a BIT array and masks are combined with logand / logior / logxor / lognot ...: on arrays of
// one dense layout that is the same function on the 32-bit words the bits are packed in -- every bit position is independent,
// and %coerce into (unsigned-byte 1) keeps exactly the low bit the word operation produces ((lognot 0) = -1 -> 1).  The nest
// becomes one loop over (unsigned-byte 32) words on the elementwise engine (32 elements per lane operation instead of the
// generic kernel's one element per thread: 0.001 of the HBM peak); fewer than 32 trailing elements stay a nest of their own.
static int lower_bitwise(Nest& n, Nest* tail, bool* has_tail) {
  *has_tail = false;
  if (n.n_out != 1 || n.n_in < 1 || n.stat) return 0;
  if (n.out[0].dtype != NCL_BIT) return 0;
  for (int i = 0; i < n.n_in; i++) if (n.in[i].dtype != NCL_BIT) return 0;
  const Expr& e = n.transform[0];
  for (int k = 0; k < e.n; k++) {
    const int op = e.node[k].op;
    if (!(op == NCL_X_INPUT || op == NCL_OP_LOGNOT || (op >= NCL_OP_LOGAND && op <= NCL_OP_LOGORC2))) return 0;
  }
  // one dense row-major layout shared by every array
  int order[NCL_MAX_LOOPS], nl = 0;
  for (int l = 0; l < n.nloops; l++) { if (n.extent[l] == 0) return 0; if (n.extent[l] > 1) order[nl++] = l; }
  for (int i = 1; i < nl; i++) { int k = order[i], j = i - 1; while (j >= 0 && n.out[0].stride[order[j]] < n.out[0].stride[k]) { order[j + 1] = order[j]; j--; } order[j + 1] = k; }
  long long total = 1;
  for (int i = nl - 1; i >= 0; i--) {
    const int l = order[i];
    if (n.out[0].stride[l] != total) return 0;
    for (int k = 0; k < n.n_in; k++) if (n.in[k].stride[l] != total) return 0;
    total *= n.extent[l];
  }
  if (total < 32) return 0;
  if (n.out[0].offset % 32) return 0;
  for (int k = 0; k < n.n_in; k++) if (n.in[k].offset % 32) return 0;
  const long long body = total / 32 * 32, rest = total - body;
  if (rest) {
    *tail = n; *has_tail = true;
    tail->nloops = 1; tail->extent[0] = rest;
    tail->out[0].stride[0] = 1; tail->out[0].offset += body;
    for (int k = 0; k < n.n_in; k++) { tail->in[k].stride[0] = 1; tail->in[k].offset += body; }
  }
  n.nloops = 1; n.extent[0] = body / 32;
  auto words = [](Operand& a) { a.dtype = NCL_UB32; a.offset /= 32; a.n_storage /= 32; a.stride[0] = 1; };
  words(n.out[0]);
  for (int k = 0; k < n.n_in; k++) words(n.in[k]);
  return 1;
}

// ---- composite transforms -----------------------------------------------------------------------------------------------
// (map-array '(sqrt (+ (* $1 $1) (* $2 $2))) a b), (einsum '(i -> (* 2.0 (+ $1 $2)) -> i) ...): the elementwise engine has
// one-function shapes, so a transform of several functions ran on the generic kernel (an interpreter per element: 0.05 of the
// HBM peak).  It is evaluated node by node instead, every node one pass of the engine into a temporary of the node's own
// class (raw 64-bit words / single-float / double-float: what the interpreter holds in a register, so the bits are the same),
// the root straight into the result through %coerce.  Four passes at stream speed beat one pass at a twentieth of it.
static int run_composite_map(Nest& n) {
  if (n.n_out != 1 || n.stat || ctx().capturing) return NCL_ERR_UNSUPPORTED;
  const Expr& e = n.transform[0];
  int nops = 0;
  for (int k = 0; k < e.n; k++) {
    const XNode& x = e.node[k];
    if (x.op == NCL_X_OUTPUT || x.cls == CLS_R || x.cls > CLS_D) return NCL_ERR_UNSUPPORTED;
    if (x.op < NCL_X_INPUT) nops++;
  }
  if (nops < 2) return NCL_ERR_UNSUPPORTED;
  long long total = 1;
  for (int l = 0; l < n.nloops; l++) { if (n.extent[l] <= 0 || n.out[0].stride[l] == 0) { if (n.extent[l] != 1) return NCL_ERR_UNSUPPORTED; } total *= n.extent[l] > 0 ? n.extent[l] : 1; }
  if (total < 65536) return NCL_ERR_UNSUPPORTED;             // small arrays: one launch of the interpreter is cheaper than several passes
  for (int k = 0; k < n.n_in; k++) { const int c = g_dt[n.in[k].dtype].cls; if (c > CLS_D || n.in[k].dtype == NCL_UB64) return NCL_ERR_UNSUPPORTED; }
  // dense row-major layout of the temporaries over the nest's loops, in the order of the result's strides
  int ord[NCL_MAX_LOOPS], nl = 0; long long tstr[NCL_MAX_LOOPS];
  for (int l = 0; l < n.nloops; l++) { tstr[l] = 0; if (n.extent[l] > 1) ord[nl++] = l; }
  for (int i = 1; i < nl; i++) { int k = ord[i], j = i - 1; while (j >= 0 && llabs(n.out[0].stride[ord[j]]) < llabs(n.out[0].stride[k])) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = k; }
  { long long acc = 1; for (int i = nl - 1; i >= 0; i--) { tstr[ord[i]] = acc; acc *= n.extent[ord[i]]; } }
  auto tdtype = [](int cls) { return cls == CLS_D ? NCL_F64 : cls == CLS_F ? NCL_F32 : NCL_SB64; };
  ncl_buf* buf[NCL_MAX_EXPR]; Operand val[NCL_MAX_EXPR]; int kind[NCL_MAX_EXPR];     // kind: 0 not materialised, 1 array operand
  for (int k = 0; k < e.n; k++) { buf[k] = nullptr; kind[k] = 0; }
  auto release = [&](int rc) { for (int k = 0; k < e.n; k++) if (buf[k]) ncl_free(buf[k]); return rc; };
  auto scalar = [&](int k, bool is_float, long long iv, double fv, int cls) -> int {      // a constant as a rank-0 array
    buf[k] = ncl_alloc(64, 0); if (!buf[k]) return NCL_ERR_NOMEM;
    Operand& o = val[k]; memset(&o, 0, sizeof o); o.base = (char*)ncl_buf_ptr(buf[k]); o.dtype = tdtype(cls); o.n_storage = 8; o.stage_item = -1;
    kind[k] = 1;
    return fill_const(o.base, o.dtype, 0, 1, is_float, iv, fv);
  };
  int rc = NCL_OK;
  for (int k = 0; k < e.n && rc == NCL_OK; k++) {
    const XNode& x = e.node[k];
    if (x.op == NCL_X_INPUT) { val[k] = n.in[x.a - 1]; kind[k] = 1; continue; }
    if (x.op == NCL_X_CONST_INT) { rc = scalar(k, false, x.ival, 0.0, CLS_I); continue; }
    if (x.op == NCL_X_CONST_F32) { rc = scalar(k, true, 0, (double)(float)x.fval, CLS_F); continue; }
    if (x.op == NCL_X_CONST_F64) { rc = scalar(k, true, 0, x.fval, CLS_D); continue; }
    const bool binary = x.op < NCL_OP_BINARY_END, root = k == e.n - 1;
    Nest* t = new Nest;
    t->nloops = n.nloops; for (int l = 0; l < n.nloops; l++) t->extent[l] = n.extent[l];
    t->n_in = binary ? 2 : 1; t->n_out = 1; t->fresh = true;
    t->in[0] = val[x.a]; if (binary) t->in[1] = val[x.b];
    if (root) t->out[0] = n.out[0];
    else {
      buf[k] = ncl_alloc(ncl_storage_bytes(tdtype(x.cls), total), 0);
      if (!buf[k]) { delete t; rc = NCL_ERR_NOMEM; break; }
      Operand& o = val[k]; memset(&o, 0, sizeof o); o.base = (char*)ncl_buf_ptr(buf[k]); o.dtype = tdtype(x.cls); o.stage_item = -1;
      o.n_storage = (int64_t)((ncl_buf_bytes(buf[k]) * 8) / g_dt[o.dtype].slot_bits);
      for (int l = 0; l < n.nloops; l++) o.stride[l] = tstr[l];
      kind[k] = 1; t->out[0] = o;
    }
    Expr& te = t->transform[0]; te.n = 0;
    auto leaf = [&](int src, int slot) { XNode& nd = te.node[te.n++]; memset(&nd, 0, sizeof nd); nd.op = NCL_X_INPUT; nd.a = slot; nd.cls = e.node[src].cls; };
    leaf(x.a, 1); if (binary) leaf(x.b, 2);
    { XNode& nd = te.node[te.n++]; nd = x; nd.a = 0; nd.b = binary ? 1 : 0; }
    if (op_may_fault(x.op)) ctx().fault_possible = true;
    rc = try_elementwise(*t);
    if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*t);
    delete t;
    // operands that are no longer needed could be freed here; they are a few arrays and go back to the cache below
  }
  return release(rc);
}

// ---- reductions of a FUNCTION of the arrays ------------------------------------------------------------------------------
// `(ij -> (+ @1 (abs $1)) -> )` (the L1 norm), `(max @1 (abs $1))` (the infinity norm), `(+ @1 (abs (- $1 $2)))` (an L1
// distance), and plain sums whose result has another class than the array (a single-float array summed into a double-float
// scalar): reduction kernels fold ARRAYS of the result's class, so these nests ran on the generic kernel, where ONE thread
// walks all reduced loops of its result element -- 19 .. 36 SECONDS for a 4096 x 4096 array.  Three fast steps instead:
//   1. the function of the inputs, evaluated over the whole iteration space into a temporary (elementwise engine / composite
//      passes) -- only when some input spans every loop, so that the temporary is no larger than that input;
//   2. the temporary folded over the reduced loops into a small temporary of its own class (reduction kernels);
//   3. result <- (op result partial) over the kept loops: accumulation into a supplied result, the start value of a fresh
//      one and the final %coerce are exactly the nest's own.
static int run_map_reduce(Nest& n) {
  if (n.n_out != 1 || n.stat || ctx().capturing) return NCL_ERR_UNSUPPORTED;
  const Expr& e = n.transform[0];
  const XNode& root = e.node[e.n - 1];
  if (!(root.op == NCL_OP_ADD || root.op == NCL_OP_MUL || root.op == NCL_OP_MAX || root.op == NCL_OP_MIN)) return NCL_ERR_UNSUPPORTED;
  int sub = -1;
  if (e.node[root.a].op == NCL_X_OUTPUT && e.node[root.a].a == 1) sub = root.b;
  else if (e.node[root.b].op == NCL_X_OUTPUT && e.node[root.b].a == 1 && (root.op == NCL_OP_ADD || root.op == NCL_OP_MUL)) sub = root.a;
  if (sub < 0) return NCL_ERR_UNSUPPORTED;
  // nodes of the subtree, remapped; it may not read the result
  int map[NCL_MAX_EXPR]; bool used[NCL_MAX_EXPR];
  for (int k = 0; k < e.n; k++) { used[k] = false; map[k] = -1; }
  used[sub] = true;
  for (int k = sub; k >= 0; k--) if (used[k]) {
    const XNode& x = e.node[k];
    if (x.op == NCL_X_OUTPUT || x.cls == CLS_R || x.cls > CLS_D) return NCL_ERR_UNSUPPORTED;
    if (x.op < NCL_X_INPUT) { used[x.a] = true; if (x.op < NCL_OP_BINARY_END) used[x.b] = true; }
  }
  const int csub = e.node[sub].cls, cc = g_dt[n.out[0].dtype].cls;
  if (cc > CLS_D) return NCL_ERR_UNSUPPORTED;
  bool contracted = false; long long total = 1;
  for (int l = 0; l < n.nloops; l++) { if (n.extent[l] <= 0) return NCL_ERR_UNSUPPORTED; if (n.extent[l] > 1 && n.out[0].stride[l] == 0) contracted = true; total *= n.extent[l]; }
  if (!contracted || total < 65536 || total > (1LL << 32)) return NCL_ERR_UNSUPPORTED;
  int span = -1;                                             // an input that moves with every loop
  for (int k = 0; k < n.n_in && span < 0; k++) {
    bool all = true; for (int l = 0; l < n.nloops; l++) if (n.extent[l] > 1 && n.in[k].stride[l] <= 0) all = false;
    if (all && g_dt[n.in[k].dtype].cls <= CLS_D && n.in[k].dtype != NCL_UB64) span = k;
  }
  for (int k = 0; k < n.n_in; k++) if (g_dt[n.in[k].dtype].cls > CLS_D || n.in[k].dtype == NCL_UB64) return NCL_ERR_UNSUPPORTED;
  // no input spans every loop (`(i j -> )`, the sum of an outer product): the temporary is larger than any input, so it is
  // only built when it is small (<= 1 GiB) and the results are few; the loops then keep the nest's own order
  if (span < 0) {
    long long results = 1; for (int l = 0; l < n.nloops; l++) if (n.out[0].stride[l] != 0) results *= n.extent[l];
    if ((double)total * 8.0 > 1073741824.0 || results >= 4096) return NCL_ERR_UNSUPPORTED;     // many results: the generic kernel has a thread for each
  }
  auto tdtype = [](int cls) { return cls == CLS_D ? NCL_F64 : cls == CLS_F ? NCL_F32 : NCL_SB64; };
  // class of the temporaries: the higher of the function's and the result's.  (+ @1 $1) into a double-float result converts
  // every single-float element before it is added (contagion), i.e. the sum runs in double-float
  const int ct = csub > cc ? csub : cc;
  const int tdt = tdtype(ct);
  if ((double)total * (g_dt[tdt].slot_bits / 8) > 8.0 * 1073741824.0) return NCL_ERR_UNSUPPORTED;
  // loop order of the spanning input; dense strides of the big temporary (all loops) and the small one (kept loops)
  int ord[NCL_MAX_LOOPS], nl = 0; long long tstr[NCL_MAX_LOOPS], ustr[NCL_MAX_LOOPS], kept_total = 1;
  for (int l = 0; l < n.nloops; l++) { tstr[l] = 0; ustr[l] = 0; if (n.extent[l] > 1) ord[nl++] = l; }
  if (span >= 0)
    for (int i = 1; i < nl; i++) { int k = ord[i], j = i - 1; while (j >= 0 && n.in[span].stride[ord[j]] < n.in[span].stride[k]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = k; }
  { long long acc = 1, ua = 1; for (int i = nl - 1; i >= 0; i--) { const int l = ord[i]; tstr[l] = acc; acc *= n.extent[l]; if (n.out[0].stride[l] != 0) { ustr[l] = ua; ua *= n.extent[l]; } } kept_total = ua; }
  ncl_buf* tb = ncl_alloc(ncl_storage_bytes(tdt, total), 0);
  ncl_buf* ub = ncl_alloc(ncl_storage_bytes(tdt, kept_total), 0);
  auto release = [&](int rc) { if (tb) ncl_free(tb); if (ub) ncl_free(ub); return rc; };
  if (!tb || !ub) return release(NCL_ERR_NOMEM);
  auto operand = [&](ncl_buf* b, const long long* str) { Operand o; memset(&o, 0, sizeof o); o.base = (char*)ncl_buf_ptr(b); o.dtype = tdt; o.stage_item = -1;
    o.n_storage = (int64_t)((ncl_buf_bytes(b) * 8) / g_dt[tdt].slot_bits); for (int l = 0; l < n.nloops; l++) o.stride[l] = str[l]; return o; };
  const Operand T = operand(tb, tstr), U = operand(ub, ustr);
  // 1. the function of the inputs over the whole iteration space
  int rc;
  {
    Nest* m = new Nest(n);
    m->out[0] = T; m->fresh = true; m->has_identity = false;
    Expr& me = m->transform[0]; me.n = 0;
    for (int k = 0; k <= sub; k++) if (used[k]) { map[k] = me.n; XNode x = e.node[k]; if (x.op < NCL_X_INPUT) { x.a = map[x.a]; if (x.op < NCL_OP_BINARY_END) x.b = map[x.b]; } me.node[me.n++] = x; }
    for (int k = 0; k < me.n; k++) if (op_may_fault(me.node[k].op)) ctx().fault_possible = true;
    rc = try_elementwise(*m);
    if (rc == NCL_ERR_UNSUPPORTED) rc = run_composite_map(*m);
    if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*m);
    delete m;
  }
  if (rc != NCL_OK) return release(rc);
  // 2. the temporary folded over the reduced loops
  {
    Nest* r = new Nest(n);
    r->n_in = 1; r->in[0] = T; r->out[0] = U; r->fresh = true; r->has_identity = true;
    { long long ii; double ff; reduce_identity(root.op, tdt, &ii, &ff); r->ident_i = ii; r->ident_f = ff; }
    Expr& re = r->transform[0]; re.n = 3;
    for (int k = 0; k < 3; k++) { memset(&re.node[k], 0, sizeof re.node[k]); re.node[k].cls = ct; }
    re.node[0].op = NCL_X_OUTPUT; re.node[0].a = 1; re.node[1].op = NCL_X_INPUT; re.node[1].a = 1; re.node[2].op = root.op; re.node[2].a = 0; re.node[2].b = 1;
    rc = try_reduce(*r);
    if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*r);
    delete r;
  }
  if (rc != NCL_OK) return release(rc);
  // 3. result <- (op result partial) over the kept loops
  {
    Nest* c = new Nest(n);
    c->n_in = 1; c->in[0] = U;
    for (int l = 0; l < n.nloops; l++) if (n.out[0].stride[l] == 0) c->extent[l] = 1;      // the reduced loops are gone
    ncl_expr src; src.n = 3;
    for (int k = 0; k < 3; k++) memset(&src.node[k], 0, sizeof src.node[k]);
    src.node[0].op = NCL_X_OUTPUT; src.node[0].a = 1; src.node[1].op = NCL_X_INPUT; src.node[1].a = 1; src.node[2].op = root.op; src.node[2].a = 0; src.node[2].b = 1;
    int in_dt = tdt, out_dt = n.out[0].dtype;
    rc = type_expr(&src, &in_dt, 1, &out_dt, 1, 0, &c->transform[0]);
    if (rc == NCL_OK) { rc = try_elementwise(*c); if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*c); }
    delete c;
  }
  if (rc == NCL_OK) ctx().last_kernel = "map_reduce";       // (ncl_last_kernel names the family: the last launch may be the tiny combine)
  return release(rc);
}

// ---- contractions of two arrays of different element types --------------------------------------------------------------
// (vdot a b) with a single-float and a double-float vector, (inner ub8-array fixnum-array), matrix x vector of mixed types:
// the dot-like reduction kernels and the GEMMs want both operands in ONE element type, and the generic kernel walks the
// contracted loops with one thread per result -- a full dot product of mixed types was a single thread.  Common Lisp's
// contagion converts the lower operand to the other's class before the product, which is what converting the ARRAY does; so
// the lower operand is copied into the other's type first (one elementwise pass) when the result has that class.  Integers of
// different widths are both brought to fixnum (values that fit 62 bits only).
static int materialize_as(const Nest& n, const Operand& X, int odt, Operand* out, ncl_buf** buf) {
  int ord[NCL_MAX_LOOPS], nl = 0; long long total = 1;
  for (int l = 0; l < n.nloops; l++) if (n.extent[l] > 1 && X.stride[l] != 0) { if (X.stride[l] < 0) return NCL_ERR_UNSUPPORTED; ord[nl++] = l; total *= n.extent[l]; }
  for (int i = 1; i < nl; i++) { int k = ord[i], j = i - 1; while (j >= 0 && X.stride[ord[j]] < X.stride[k]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = k; }
  *buf = ncl_alloc(ncl_storage_bytes(odt, total), 0);
  if (!*buf) return NCL_ERR_NOMEM;
  Operand o; memset(&o, 0, sizeof o);
  o.base = (char*)ncl_buf_ptr(*buf); o.dtype = odt; o.stage_item = -1; o.n_storage = (int64_t)((ncl_buf_bytes(*buf) * 8) / g_dt[odt].slot_bits);
  { long long acc = 1; for (int i = nl - 1; i >= 0; i--) { o.stride[ord[i]] = acc; acc *= n.extent[ord[i]]; } }
  Nest* t = new Nest;
  t->nloops = nl; t->n_in = 1; t->n_out = 1; t->fresh = true;
  t->in[0] = X; t->out[0] = o;
  for (int i = 0; i < nl; i++) { t->extent[i] = n.extent[ord[i]]; t->in[0].stride[i] = X.stride[ord[i]]; t->out[0].stride[i] = o.stride[ord[i]]; }
  for (int i = nl; i < NCL_MAX_LOOPS; i++) { t->in[0].stride[i] = 0; t->out[0].stride[i] = 0; }
  if (nl == 0) { t->nloops = 1; t->extent[0] = 1; }
  Expr& x = t->transform[0]; x.n = 1; memset(&x.node[0], 0, sizeof x.node[0]); x.node[0].op = NCL_X_INPUT; x.node[0].a = 1; x.node[0].cls = g_dt[X.dtype].cls;
  int rc = try_elementwise(*t);
  if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*t);
  delete t;
  *out = o;
  return rc;
}
static int harmonize_contraction(Nest& n, ncl_buf** tmp) {
  tmp[0] = tmp[1] = nullptr;
  if (n.n_in != 2 || n.n_out != 1 || n.stat || ctx().capturing) return NCL_OK;
  const Expr& e = n.transform[0];
  if (e.n != 5) return NCL_OK;
  const XNode& root = e.node[4];
  if (root.op != NCL_OP_ADD || e.node[root.a].op != NCL_X_OUTPUT || e.node[root.b].op != NCL_OP_MUL) return NCL_OK;
  const XNode& mul = e.node[root.b];
  if (e.node[mul.a].op != NCL_X_INPUT || e.node[mul.b].op != NCL_X_INPUT || e.node[mul.a].a == e.node[mul.b].a) return NCL_OK;
  if (n.in[0].dtype == n.in[1].dtype) return NCL_OK;
  bool contracted = false;
  for (int l = 0; l < n.nloops; l++) if (n.extent[l] > 1 && n.out[0].stride[l] == 0) contracted = true;
  if (!contracted) return NCL_OK;
  const int c0 = g_dt[n.in[0].dtype].cls, c1 = g_dt[n.in[1].dtype].cls, cc = g_dt[n.out[0].dtype].cls;
  if (c0 > CLS_D || c1 > CLS_D || cc > CLS_D) return NCL_OK;
  int want[2] = {n.in[0].dtype, n.in[1].dtype};
  if (c0 != c1) {
    const int hi = c0 > c1 ? 0 : 1;
    if (cc != g_dt[n.in[hi].dtype].cls || !(n.in[hi].dtype == NCL_F32 || n.in[hi].dtype == NCL_F64) || n.in[1 - hi].dtype == NCL_UB64) return NCL_OK;
    want[1 - hi] = n.in[hi].dtype;
  } else if (c0 == CLS_I && cc == CLS_I) {
    for (int k = 0; k < 2; k++) {
      const DtInfo& d = g_dt[n.in[k].dtype];
      if (n.in[k].dtype == NCL_FIXNUM) continue;
      if (d.value_bits > (d.is_signed ? 63 : 62)) return NCL_OK;       // sb64 / ub63 / ub64 values need not fit a fixnum
      want[k] = NCL_FIXNUM;
    }
  } else return NCL_OK;
  for (int k = 0; k < 2; k++) if (want[k] != n.in[k].dtype) {
    Operand o; int rc = materialize_as(n, n.in[k], want[k], &o, &tmp[k]);
    if (rc != NCL_OK) return rc;
    n.in[k] = o;
  }
  // the typed transform: both products now have the operands' common class
  Expr& te = n.transform[0];
  const int pc = g_dt[n.in[0].dtype].cls;
  te.node[mul.a].cls = pc; te.node[mul.b].cls = pc; te.node[root.b].cls = pc;
  return NCL_OK;
}

int run_nest(Nest& n) {
  for (int i = 0; i < n.n_in; i++) NCL_TRY(check_bounds(n.in[i], n, "input", i));
  for (int o = 0; o < n.n_out; o++) NCL_TRY(check_bounds(n.out[o], n, "output", o));
  bool any_complex = false;
  for (int k = 0; k < n.n_in + n.n_out; k++) if (cls_complex(g_dt[(k < n.n_in ? n.in[k] : n.out[k - n.n_in]).dtype].cls)) any_complex = true;
  if (any_complex && !(!ctx().force_generic && lower_complex(n))) {
    if (n.stat) return NCL_ERR_UNSUPPORTED;
    return launch_generic(n);
  }
  bool empty = false;
  for (int l = 0; l < n.nloops; l++) if (n.extent[l] == 0) empty = true;
  if (!empty && !ctx().force_generic) {
    Nest* tail = new Nest; bool has_tail = false;
    if (lower_bitwise(n, tail, &has_tail) && has_tail) { int rc = launch_generic(*tail); if (rc != NCL_OK) { delete tail; return rc; } }
    delete tail;
  }
  if (n.stat) {                                        // fused statistics: reduction kernels only (ncl_reduce_stat composes otherwise)
    if (empty || ctx().force_generic) return NCL_ERR_UNSUPPORTED;
    int rc = try_dot_reduce(n);
    return rc != NCL_ERR_UNSUPPORTED ? rc : try_reduce(n);
  }
  if (!ctx().force_generic && !empty) {
    ncl_buf* tmp[2];
    int rc = harmonize_contraction(n, tmp);
    auto done = [&](int code) { for (int k = 0; k < 2; k++) if (tmp[k]) ncl_free(tmp[k]); return code; };
    if (rc != NCL_OK) return done(rc);
    if ((rc = try_dot_reduce(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    if ((rc = try_gemm(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    if ((rc = try_reduce(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    if ((rc = try_transpose(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    if ((rc = try_elementwise(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    if ((rc = run_composite_map(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    if ((rc = run_map_reduce(n)) != NCL_ERR_UNSUPPORTED) return done(rc);
    return done(launch_generic(n));
  }
  return launch_generic(n);
}
