// k_elementwise.cu -- host side of K1/K2/K6: expression pattern -> functor, nest -> EwParams.
// The kernel template and its documentation live in k_elementwise.cuh; the instantiations are
// split by value class over k_ew_*.cu.
#include "k_elementwise.cuh"

// ---- host: expression pattern -> functor, nest -> EwParams ---------------------------------------------------


struct Slot { int kind; int idx; };   // kind 0: input idx (0-based); 1: the output (read); 2: const 0; 3: const 1
struct Match { int shape; int n; Slot s[3]; int rt_op; };

static bool leaf_slot(const Expr& e, int k, int self_out, Slot* s) {
  const XNode& x = e.node[k];
  if (x.op == NCL_X_INPUT) { s->kind = 0; s->idx = x.a - 1; return true; }
  if (x.op == NCL_X_OUTPUT && x.a == self_out + 1) { s->kind = 1; s->idx = 0; return true; }
  if (x.op == NCL_X_CONST_INT && (x.ival == 0 || x.ival == 1)) { s->kind = 2 + (int)x.ival; s->idx = 0; return true; }
  return false;
}
static bool is_const(const Expr& e, int k, long long v) { return e.node[k].op == NCL_X_CONST_INT && e.node[k].ival == v; }
static bool is_self_out(const Expr& e, int k, int self_out) { return e.node[k].op == NCL_X_OUTPUT && e.node[k].a == self_out + 1; }

static int bin_shape(int op) {
  switch (op) { case NCL_OP_ADD: return S_ADD; case NCL_OP_SUB: return S_SUB; case NCL_OP_MUL: return S_MUL; case NCL_OP_DIV: return S_DIV;
    case NCL_OP_MAX: return S_MAX; case NCL_OP_MIN: return S_MIN; case NCL_OP_EQ: return S_EQ; case NCL_OP_NE: return S_NE;
    case NCL_OP_LE: return S_LE; case NCL_OP_GE: return S_GE; case NCL_OP_LT: return S_LT; case NCL_OP_GT: return S_GT;
    case NCL_OP_LOGAND: return S_AND; case NCL_OP_LOGIOR: return S_IOR; case NCL_OP_LOGXOR: return S_XOR;
    case NCL_OP_LOGANDC1: case NCL_OP_LOGANDC2: case NCL_OP_LOGEQV: case NCL_OP_LOGNAND: case NCL_OP_LOGNOR: case NCL_OP_LOGORC1: case NCL_OP_LOGORC2:
    case NCL_OP_FLOOR: case NCL_OP_CEILING: case NCL_OP_ROUND: case NCL_OP_TRUNCATE: case NCL_OP_MOD: case NCL_OP_REM: return S_RT2;
    default: return S_NONE; }       // expt: generic kernel (integer and float powers take different algorithms)
}
static int un_shape(int op) {
  switch (op) { case NCL_OP_NEG: return S_NEG; case NCL_OP_ABS: return S_ABS; case NCL_OP_SQUARE: return S_SQUARE; case NCL_OP_LOGNOT: return S_NOT;
    case NCL_OP_IDENTITY: return S_COPY; case NCL_OP_EXP: return S_EXP;
    case NCL_OP_SIN: return S_SIN; case NCL_OP_COS: return S_COS; case NCL_OP_TAN: return S_TAN; case NCL_OP_TANH: return S_TANH;
    // the runtime-selected families (k_ew_rt.cu); sqrt and log live there too because they can raise the DOMAIN flag
    case NCL_OP_SQRT: case NCL_OP_LOG: case NCL_OP_LOG2: case NCL_OP_ASIN: case NCL_OP_ACOS: case NCL_OP_ATAN: case NCL_OP_SINH: case NCL_OP_COSH:
    case NCL_OP_ASINH: case NCL_OP_ACOSH: case NCL_OP_ATANH: case NCL_OP_ISQRT: case NCL_OP_LOGCOUNT: case NCL_OP_INTEGER_LENGTH: return S_RT1;
    default: return S_NONE; }
}

static bool match_expr(const Expr& e, int self_out, bool fresh, Match* m) {
  int r = e.n - 1; const XNode& root = e.node[r];
  m->n = 0; m->shape = S_NONE; m->rt_op = root.op;
  if (root.op == NCL_X_INPUT) { m->shape = S_COPY; m->n = 1; return leaf_slot(e, r, self_out, &m->s[0]); }
  if (root.op >= 0 && root.op < NCL_OP_BINARY_END) {
    const XNode& A = e.node[root.a]; const XNode& B = e.node[root.b];
    // (+ (+ 0 $a) $b) / (* (* 1 $a) $b): identity-seeded folds of numcl:+ and numcl:*
    if ((root.op == NCL_OP_ADD || root.op == NCL_OP_MUL) && A.op == root.op && is_const(e, A.a, root.op == NCL_OP_ADD ? 0 : 1)
        && e.node[A.b].op == NCL_X_INPUT && B.op == NCL_X_INPUT) {
      m->shape = root.op == NCL_OP_ADD ? S_ADD0 : S_MUL; m->n = 2;
      return leaf_slot(e, A.b, self_out, &m->s[0]) && leaf_slot(e, root.b, self_out, &m->s[1]);
    }
    // default transform (+ @1 (* $a $b)) / (+ @1 $a)
    if (root.op == NCL_OP_ADD && is_self_out(e, root.a, self_out)) {
      if (B.op == NCL_OP_MUL && e.node[B.a].op == NCL_X_INPUT && e.node[B.b].op == NCL_X_INPUT) {
        if (fresh) { m->shape = S_MUL0; m->n = 2; return leaf_slot(e, B.a, self_out, &m->s[0]) && leaf_slot(e, B.b, self_out, &m->s[1]); }
        m->shape = S_ADDMUL; m->n = 3;
        return leaf_slot(e, root.a, self_out, &m->s[0]) && leaf_slot(e, B.a, self_out, &m->s[1]) && leaf_slot(e, B.b, self_out, &m->s[2]);
      }
      if (B.op == NCL_X_INPUT && fresh) { m->shape = S_ADD0U; m->n = 1; return leaf_slot(e, root.b, self_out, &m->s[0]); }
    }
    // (+ (* $a $b) $c) / (+ $c (* $a $b)): a map over three arrays (map-array with a fused multiply-add form): the product is
    // rounded, then added -- the accumulate functor with the addend in the accumulator's place (float + is commutative)
    if (root.op == NCL_OP_ADD) {
      for (int side = 0; side < 2; side++) {
        const int im = side ? root.b : root.a, ic = side ? root.a : root.b;
        const XNode& M = e.node[im];
        Slot c0;
        if (M.op == NCL_OP_MUL && e.node[M.a].op == NCL_X_INPUT && e.node[M.b].op == NCL_X_INPUT && e.node[ic].op == NCL_X_INPUT && leaf_slot(e, ic, self_out, &c0) && c0.kind == 0) {
          m->shape = S_ADDMUL; m->n = 3; m->s[0] = c0;
          return leaf_slot(e, M.a, self_out, &m->s[1]) && leaf_slot(e, M.b, self_out, &m->s[2]);
        }
      }
    }
    Slot a, b;
    if (leaf_slot(e, root.a, self_out, &a) && leaf_slot(e, root.b, self_out, &b)) {
      if ((a.kind == 1 || b.kind == 1) && fresh) return false;      // reads a fresh output: let the generic path define it
      m->shape = bin_shape(root.op); m->n = 2; m->s[0] = a; m->s[1] = b; return m->shape != S_NONE;
    }
    return false;
  }
  if (root.op >= NCL_OP_NEG && root.op <= NCL_OP_INTEGER_LENGTH) {
    Slot a; if (!leaf_slot(e, root.a, self_out, &a) || a.kind != 0) return false;
    m->shape = un_shape(root.op); m->n = 1; m->s[0] = a; return m->shape != S_NONE;
  }
  return false;
}

int ew_narrow_launch(int op, bool wide, const char* a, const char* b, char* out, long long n, bool b_scalar, long long b_off, int b_lk, bool swap,
                     bool signed8, unsigned mask16);                      // k_ew_narrow.cu
int ew_widen_launch(int a_lk, int op, const char* a, const char* b, char* out, unsigned nchunks, int bmode, unsigned period, int btag, int swap,
                    unsigned long long mask, int shift, const Operand* oa, const Operand* ob, const Operand* oout);   // k_ew_widen.cu

static int try_elementwise_32(const Nest& n);

// The engine decodes chunk and row numbers with 32-bit multiply-shift arithmetic, so one launch covers fewer than 2^31
// elements.  Larger nests (a 2 GiB (unsigned-byte 8) array is ordinary on a 180 GB part) are cut on the host along the loop with
// the largest output stride -- or, for one contiguous run, into runs of 2^30 elements -- and every piece takes the same fast
// path: same kernels, same results, a few launches instead of the thread-per-element generic kernel.
int try_elementwise(const Nest& n) {
  if (n.n_out != 1) return NCL_ERR_UNSUPPORTED;
  long long total = 1; int big = -1; long long bigstride = -1;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] <= 1) continue;
    if (n.out[0].stride[l] == 0) return try_elementwise_32(n);            // a contracted loop: not elementwise at all
    total *= n.extent[l];
    if (llabs(n.out[0].stride[l]) > bigstride) { bigstride = llabs(n.out[0].stride[l]); big = l; }
  }
  const long long LIMIT = 1LL << 30;
  if (total <= LIMIT || big < 0) return try_elementwise_32(n);
  // pieces of the outermost loop; a single contiguous run is cut at multiples of 2^20 elements (keeps every vector alignment)
  const long long per_step = total / n.extent[big];
  long long step = per_step >= LIMIT ? 1 : LIMIT / per_step;
  if (n.out[0].stride[big] == 1 || n.out[0].stride[big] == -1) step = step >> 20 << 20;
  if (step < 1) step = 1;
  for (long long s0 = 0; s0 < n.extent[big]; s0 += step) {
    Nest* x = new Nest(n);
    x->extent[big] = n.extent[big] - s0 < step ? n.extent[big] - s0 : step;
    x->out[0].offset += s0 * n.out[0].stride[big];
    for (int i = 0; i < n.n_in; i++) x->in[i].offset += s0 * n.in[i].stride[big];
    int rc = try_elementwise(*x);                                             // a piece may need cutting again along another loop
    if (rc == NCL_ERR_UNSUPPORTED && s0 > 0) rc = launch_generic(*x);         // never hand on a half-written result
    delete x;
    if (rc != NCL_OK) return rc;
  }
  return NCL_OK;
}

static int try_elementwise_32(const Nest& n) {
  if (n.n_out != 1) return NCL_ERR_UNSUPPORTED;
  const Operand& out = n.out[0];
  for (int l = 0; l < n.nloops; l++) if (n.extent[l] > 1 && out.stride[l] == 0) return NCL_ERR_UNSUPPORTED;   // contracted loop
  Match m;
  if (!match_expr(n.transform[0], 0, n.fresh, &m)) return NCL_ERR_UNSUPPORTED;
  Context& c = ctx();
  // operands
  Operand ops[3]; int ocls[3];
  for (int k = 0; k < m.n; k++) {
    const Slot& s = m.s[k];
    if (s.kind == 0) { if (s.idx < 0 || s.idx >= n.n_in) return NCL_ERR_UNSUPPORTED; ops[k] = n.in[s.idx]; }
    else if (s.kind == 1) ops[k] = out;
    else { memset(&ops[k], 0, sizeof(Operand)); ops[k].base = (char*)c.consts; ops[k].dtype = NCL_SB64; ops[k].offset = s.kind - 2; ops[k].n_storage = 8; }
    ocls[k] = g_dt[ops[k].dtype].cls;
  }
  int cc = CLS_I;
  for (int k = 0; k < m.n; k++) if (ocls[k] > cc) cc = ocls[k];
  if ((m.shape == S_ADDMUL || m.shape == S_MUL0) && m.n >= 2) {
    // a*b must be computed in the operands' own class before it meets the accumulator
    int pc = ocls[m.n - 1] > ocls[m.n - 2] ? ocls[m.n - 1] : ocls[m.n - 2];
    if (pc != cc && !(ocls[m.n - 1] == ocls[m.n - 2])) return NCL_ERR_UNSUPPORTED;
    if (ocls[m.n - 1] != cc || ocls[m.n - 2] != cc) return NCL_ERR_UNSUPPORTED;
  }
  int out_cls = g_dt[out.dtype].cls;
  int shape = m.shape;
  // astype / copy into a float type (src/3copy.lisp:27-43): the operand is converted on load exactly as `coerce` does
  // (integer -> float and double -> single round to nearest, single -> double is exact), so the copy runs in the output's class
  if (shape == S_COPY && out_cls != cc && out_cls != CLS_I) cc = out_cls;
  bool is_cmp = shape >= S_EQ && shape <= S_GT;
  int rt_family = -1;
  if (shape == S_RT1) {
    const int op = m.rt_op;
    if (op == NCL_OP_ISQRT || op == NCL_OP_LOGCOUNT || op == NCL_OP_INTEGER_LENGTH) { if (cc != CLS_I) return NCL_ERR_UNSUPPORTED; rt_family = RT_UN_I; }
    else { if (cc == CLS_I) cc = CLS_F; rt_family = cc == CLS_F ? RT_UN_F : RT_UN_D; }      // a rational argument is converted to single-float first
  } else if (shape == S_RT2) {
    const int op = m.rt_op;
    if (cc == CLS_I) rt_family = RT_BIN_I;
    else if (op == NCL_OP_MOD || op == NCL_OP_REM) rt_family = cc == CLS_F ? RT_BIN_F : RT_BIN_D;
    else return NCL_ERR_UNSUPPORTED;                       // float -> integer quotients: generic kernel
  }
  if (is_cmp && cc != CLS_I) {
    // Common Lisp compares an integer with a float EXACTLY (CLHS 12.1.4.1).  Converting the integer to the float's format first
    // is only the same thing while every value of the integer's type is representable there: 24 bits in single-float lanes,
    // 53 in double-float lanes.  Wider 32-bit integers against single-floats are compared in double-float lanes (exact for
    // both); 64-bit integers against any float go to the generic kernel, which compares exactly (cmp_i64_f64).
    int wide = 0;
    for (int k = 0; k < m.n; k++) if (ocls[k] == CLS_I && m.s[k].kind < 2) { const int vb = g_dt[ops[k].dtype].value_bits; if (vb > wide) wide = vb; }
    if (wide > 53) return NCL_ERR_UNSUPPORTED;
    if (wide > 24 && cc == CLS_F) cc = CLS_D;
  }
  if ((shape >= S_SQRT && shape <= S_TANH) && cc == CLS_I) cc = CLS_F;      // a rational argument is converted to single-float first (float substitution)
  if (shape == S_DIV && cc == CLS_I) { shape = S_IDIV; if (out_cls != CLS_F) return NCL_ERR_UNSUPPORTED; }
  else if (is_cmp) { if (out_cls != CLS_I) return NCL_ERR_UNSUPPORTED; }
  else if (out_cls != cc) return NCL_ERR_UNSUPPORTED;                    // float->int rounding stores etc.: generic path

  // ---- collapse the nest: drop unit loops, sort by output stride, merge ----
  struct D { long long ext; long long st[4]; };
  D d[NCL_MAX_LOOPS]; int nd = 0;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] == 1) continue;
    d[nd].ext = n.extent[l]; d[nd].st[0] = out.stride[l];
    for (int k = 0; k < 3; k++) d[nd].st[1 + k] = k < m.n ? ops[k].stride[l] : 0;
    nd++;
  }
  for (int i = 1; i < nd; i++) { D t = d[i]; int j = i - 1; while (j >= 0 && llabs(d[j].st[0]) < llabs(t.st[0])) { d[j + 1] = d[j]; j--; } d[j + 1] = t; }
  for (int i = nd - 1; i > 0; i--) {             // merge d[i-1] (outer) with d[i] (inner)
    bool ok = true;
    for (int k = 0; k < 4; k++) if (d[i - 1].st[k] != d[i].st[k] * d[i].ext) ok = false;
    if (ok) { d[i - 1].ext *= d[i].ext; for (int k = 0; k < 4; k++) d[i - 1].st[k] = d[i].st[k]; for (int j = i; j + 1 < nd; j++) d[j] = d[j + 1]; nd--; }
  }
  if (nd == 0) { d[0].ext = 1; for (int k = 0; k < 4; k++) d[0].st[k] = 0; d[0].st[0] = 1; nd = 1; }
  if (nd > EW_MAXD) return NCL_ERR_UNSUPPORTED;
  if (d[nd - 1].st[0] != 1) return NCL_ERR_UNSUPPORTED;
  for (int i = 0; i < nd; i++) {
    if (d[i].ext >= 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
    for (int k = 0; k < 4; k++) if (llabs(d[i].st[k]) >= 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
  }

  StoreSpec sk = dtype_store_spec(out.dtype);
  int out_bits = g_dt[out.dtype].slot_bits;
  // elements per chunk: 16 bytes of byte-addressable slots, or ONE 32-bit word of a packed result (32 bits, 16 ub2, 8 ub4 --
  // numcl infers (unsigned-byte 2 / 4) for small ranges: (mod a 7), mask + mask, (arange 5))
  int E = out_bits >= 8 ? 128 / out_bits : 32 / out_bits;
  // vector eligibility: inner extent divisible, every vectorised operand 16B-aligned at every row start
  const D& in0 = d[nd - 1];
  auto aligned = [&](int e) {
    if (in0.ext % e) return false;
    auto chk = [&](const Operand& o, int slot, int k) {
      int sb = g_dt[o.dtype].slot_bits;
      if (sb < 8 && slot != 0) return e * sb <= 32;     // packed inputs: E consecutive fields from one or two words (load_bits_raw), any offset
      if (sb < 8 && e * sb >= 32) {                     // packed result: every chunk is a whole, aligned 32-bit word
        if (e * sb != 32 || ((uintptr_t)o.base % 4) || (o.offset * sb) % 32) return false;
        for (int i = 0; i < nd - 1; i++) if ((d[i].st[k] * sb) % 32) return false;
        return true;
      }
      long long bytes = (long long)e * sb / 8; long long a = bytes >= 16 ? 16 : bytes;
      if (a <= 1) return true;
      if (((uintptr_t)o.base + (o.offset * sb) / 8) % a) return false;
      if ((o.offset * sb) % 8) return false;
      for (int i = 0; i < nd - 1; i++) if (((d[i].st[k] * sb) / 8) % a || (d[i].st[k] * sb) % 8) return false;
      (void)slot; return true;
    };
    if (!chk(out, 0, 0)) return false;
    for (int k = 0; k < m.n; k++) { if (in0.st[1 + k] == 1) { if (!chk(ops[k], 1 + k, 1 + k)) return false; } else if (in0.st[1 + k] != 0) return false; }
    return true;
  };
  if (E > 1 && !aligned(E) && nd == 1 && d[0].ext % E != 0 && d[0].ext > 4LL * E) {
    // One contiguous run whose length is not a multiple of the 16-byte chunk (an odd-sized array): vector body +
    // scalar tail, as two nests of one loop each.  (The all-scalar kernel ran at 0.53 of the HBM peak on 16383^2.)
    // Packed comparison results likewise: whole 32-bit result words, then the last partial word ((numcl:< a 0.5) of an array
    // whose size is not a multiple of 32 used to run on the generic kernel altogether).
    bool unit = d[0].st[0] == 1;
    for (int k = 0; k < m.n; k++) if (d[0].st[1 + k] != 0 && d[0].st[1 + k] != 1) unit = false;
    if (unit) {
      const long long tail = d[0].ext % E, body = d[0].ext - tail;
      auto piece = [&](long long len, long long shift) {
        Nest x = n;
        x.nloops = 1; x.extent[0] = len;
        for (int i = 0; i < x.n_in; i++) {
          long long st = 0;
          for (int k = 0; k < m.n; k++) if (m.s[k].kind == 0 && m.s[k].idx == i) st = d[0].st[1 + k];
          x.in[i].stride[0] = st; x.in[i].offset += shift * st;
        }
        x.out[0].stride[0] = 1; x.out[0].offset += shift;
        return x;
      };
      Nest nb = piece(body, 0);
      int rc = try_elementwise(nb);
      if (rc != NCL_OK) return rc;
      const char* body_kernel = c.last_kernel;
      Nest nt = piece(tail, body);
      rc = try_elementwise(nt);
      if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(nt);           // the body is done: never hand the whole nest on
      c.last_kernel = body_kernel;                                      // (ncl_last_kernel names the family that did the work)
      return rc;
    }
  }
  // Element-granular flat mode: the output is contiguous over a (rows x P) nest whose row length P is not a multiple of
  // the chunk (a trailing axis of 3, an odd-length row vector), every operand is either the same contiguous shape (streamed
  // as flat vectors) or anything else with one stride per loop (a row vector, one value per row, a scalar: gathered per
  // element, (row, col) decoded once per chunk).  Without it these nests ran element-wise, 3 lanes of a warp busy:
  // (16M,3) + (3) at 0.018 of the HBM peak.
  bool eflat = false; int gather[3] = {0, 0, 0}; long long g_outer[3] = {0, 0, 0}; long long eflat_tail = 0, eflat_P = 0, eflat_R = 0;
  // (packed comparison results too: `(N,3) < (3)` has rows that are not whole result words and used to run on the generic kernel,
  // 0.05 of the HBM peak; a lane then owns one 32-bit result word = 32 consecutive flat elements)
  if (E > 1 && (out_bits >= 8 || (is_cmp && out_bits == 1 && E == 32)) && nd == 2 && !aligned(E) && d[0].st[0] == d[1].ext && d[1].st[0] == 1) {
    const long long P = d[1].ext, R = d[0].ext, total = P * R;
    bool ok = total < 0x7fffffffLL && total >= 64LL * E && P < 0x40000000LL;
    ok = ok && (((uintptr_t)out.base + (out.offset * out_bits) / 8) % 16) == 0;
    for (int k = 0; k < m.n && ok; k++) {
      const long long s_o = d[0].st[1 + k], s_i = d[1].st[1 + k];
      const int sb = g_dt[ops[k].dtype].slot_bits;
      const long long cb = (long long)E * sb / 8;
      if (s_i == 1 && s_o == P && sb >= 8 && (((uintptr_t)ops[k].base + (ops[k].offset * sb) / 8) % (cb >= 16 ? 16 : cb)) == 0) gather[k] = 0;
      else { gather[k] = 1; g_outer[k] = s_o; }
    }
    if (ok) {
      eflat = true; eflat_P = P; eflat_R = R; eflat_tail = total % E;
      d[0].ext = total - eflat_tail; d[0].st[0] = 1;
      for (int k = 0; k < m.n; k++) d[0].st[1 + k] = gather[k] ? d[1].st[1 + k] : 1;
      nd = 1;
    }
  }
  if (!eflat && E > 1 && !aligned(E)) { if (out_bits < 8) return NCL_ERR_UNSUPPORTED; E = 1; }

  // 32-bit lanes when every slot involved is at most 32 bits wide (ub8 + ub8 -> ub15, sb16 * sb16 -> sb32, ub8 < ub8 -> bit):
  // half the registers and no 64-bit multiply / add pairs
  bool i32 = false;
  if (cc == CLS_I && E > 1 && shape != S_IDIV && rt_family < 0 && (is_cmp || (out_bits <= 32 && !g_dt[out.dtype].tagged))) {
    i32 = true;
    const bool ordered = is_cmp || shape == S_MAX || shape == S_MIN || shape == S_ABS;      // needs the true value in a signed 32-bit lane
    for (int k = 0; k < m.n; k++) {
      if (m.s[k].kind >= 2) continue;                                                      // the constants 0 / 1
      const int lk = dtype_load_kind(ops[k].dtype);
      const bool packed_scalar = lk >= LK_P1;                                                // (full nil 3) is a rank-0 ub2 array: read through load_scalar
      if (!((lk >= LK_U8 && lk <= LK_S32) || packed_scalar) || (ordered && lk == LK_U32)) i32 = false;
    }
  }
  EwKernel kfn = nullptr;
  // comparisons into a packed bit array: the input-centric kernel (k_ew_cmppack.cu) when the rows are whole result words and
  // every streamed operand is chunk aligned; E becomes the elements per 16-byte INPUT chunk
  bool cmppack = false; int cmp_U = 0;
  if (is_cmp && out_bits == 1 && m.n == 2 && !eflat && E == 32) {
    int maxb = 0;
    for (int k = 0; k < 2; k++) if (in0.st[1 + k] == 1) { const int b = g_dt[ops[k].dtype].slot_bits / 8; if (b > maxb) maxb = b; }
    if (maxb >= 1) {
      const int EI = 16 / maxb;
      kfn = ew_pick_cmppack(i32 ? 4 : cc, EI, &cmp_U);
      // 8- / 16-bit operands of one kind (the other may be a broadcast scalar that fits the field): SIMD-in-register compare
      if (i32 && (maxb == 1 || maxb == 2)) {
        bool same = true; int kind = -1;
        for (int k = 0; k < 2 && same; k++) {
          const int lk = dtype_load_kind(ops[k].dtype); const DtInfo& dk = g_dt[ops[k].dtype];
          if (in0.st[1 + k] == 1) { if (dk.slot_bits != maxb * 8 || (kind >= 0 && kind != lk)) same = false; kind = lk; }
          else if (in0.st[1 + k] != 0 || m.s[k].kind >= 2) same = false;
        }
        for (int k = 0; k < 2 && same; k++) if (in0.st[1 + k] == 0) {        // the scalar's every value must fit the array's fields with the same meaning
          const DtInfo& dk = g_dt[ops[k].dtype]; const bool arr_signed = (lk_flags(kind) & 1u) != 0;
          if (dk.is_signed ? !(arr_signed && dk.value_bits <= maxb * 8) : dk.value_bits > (arr_signed ? maxb * 8 - 1 : maxb * 8)) same = false;
        }
        if (same && kind >= 0) { EwKernel k2 = ew_pick_cmppack_swar(maxb * 8, &cmp_U); if (k2) kfn = k2; }
      }
      if (kfn && aligned(EI)) { cmppack = true; E = EI; } else kfn = nullptr;
    }
  }
  if (i32 && !cmppack) { kfn = is_cmp ? (E == 32 ? ew_pick_cmp(shape, 4) : nullptr) : ew_pick_i32(shape, E); if (!kfn) i32 = false; }
  if (!kfn) {
    if (rt_family >= 0) kfn = ew_pick_rt(rt_family, E);
    else if (shape == S_IDIV) kfn = ew_pick_idiv(E);
    else if (is_cmp) { if (E != 32) return NCL_ERR_UNSUPPORTED; kfn = ew_pick_cmp(shape, cc); }
    else if (cc == CLS_F) kfn = ew_pick_f32(shape, E);
    else if (cc == CLS_D) kfn = ew_pick_f64(shape, E);
    else kfn = ew_pick_int(shape, E);
  }
  if (!kfn) return NCL_ERR_UNSUPPORTED;

  EwParams p; memset(&p, 0, sizeof p);
  // Flat mode: two dims where every operand is either contiguous across both (same shape as the
  // output), a row vector repeated over the outer dim (stride 0 there: numpy broadcasting), or a
  // scalar.  The nest is then walked as ONE flat run of chunks; the row-vector operands are read at
  // (chunk mod period).  This keeps warp segments full when rows are short (C5a: 256 elements).
  if (!cmppack && nd == 2 && E > 1 && d[0].st[0] == d[1].ext && d[1].st[0] == 1 && d[1].ext * d[0].ext < 0x7fffffffLL) {
    bool ok = true; unsigned period[3] = {0, 0, 0};
    for (int k = 0; k < m.n && ok; k++) {
      long long s_outer = d[0].st[1 + k], s_inner = d[1].st[1 + k];
      if (s_inner == 1 && s_outer == d[1].ext) period[k] = 0;                 // flat
      else if (s_inner == 1 && s_outer == 0) period[k] = (unsigned)(d[1].ext / E);   // row vector
      else if (s_inner == 0 && s_outer == 0) period[k] = 0;                   // scalar
      else ok = false;
    }
    if (ok) {
      d[0].ext = d[0].ext * d[1].ext; d[0].st[0] = 1;
      for (int k = 0; k < m.n; k++) d[0].st[1 + k] = d[1].st[1 + k];
      nd = 1;
      for (int k = 0; k < m.n; k++) if (period[k] >= 1) {
        p.period[k] = period[k];
        if (period[k] > 1) {
          unsigned s2 = 0; while ((1ull << s2) < period[k]) s2++;
          p.pmul[k] = (unsigned)(((1ull << (31 + s2)) + period[k] - 1) / period[k]); p.pshr[k] = s2 - 1;
        }                                                                     // period 1: ew_div(q, 0, 0) = q, so q mod 1 = 0
      }
    }
  }
  // ---- 8-bit arrays, SIMD within a register (k_ew_narrow.cu): ub8 (op) ub8 / scalar -> 8- or 16-bit result, one flat run ----
  if (nd == 1 && !eflat && !cmppack && cc == CLS_I && out_cls == CLS_I && m.n == 2 && !is_cmp && rt_family < 0 && (out_bits == 8 || out_bits == 16) &&
      !g_dt[out.dtype].tagged && d[0].ext % 16 == 0 && d[0].ext >= 4096 && p.period[0] == 0 && p.period[1] == 0) {
    const int nop = (shape == S_ADD || shape == S_ADD0) ? 0 : shape == S_SUB ? 1 : (shape == S_MUL || shape == S_MUL0) ? 2 : shape == S_MAX ? 3 : shape == S_MIN ? 4
                    : shape == S_AND ? 5 : shape == S_IOR ? 6 : shape == S_XOR ? 7 : -1;
    const bool wide = out_bits == 16;
    // which operand streams (stride 1, an 8-bit array) and which is the other one (the same, or a broadcast scalar)
    int ia = -1, ib = -1;
    for (int k = 0; k < 2; k++) if (d[0].st[1 + k] == 1 && g_dt[ops[k].dtype].slot_bits == 8 && m.s[k].kind == 0) { if (ia < 0) ia = k; else ib = k; }
    if (ia >= 0 && ib < 0) { const int o = 1 - ia; if (d[0].st[1 + o] == 0 && m.s[o].kind == 0 && g_dt[ops[o].dtype].slot_bits <= 16) ib = o; else ia = -1; }
    bool ok = nop >= 0 && ia >= 0 && ib >= 0 && (wide ? nop <= 2 : nop >= 3);
    const bool b_scalar = ok && d[0].st[1 + ib] == 0;
    const DtInfo& da = g_dt[ops[ia >= 0 ? ia : 0].dtype]; const DtInfo& db = g_dt[ops[ib >= 0 ? ib : 0].dtype];
    bool signed8 = false;
    if (ok && !wide) {
      // max min and ior xor: both operands of one signedness, every value representable in the result's type
      if (b_scalar) ok = false;                                        // (a rank-0 operand of another width: the general engine)
      else if (da.is_signed != db.is_signed) ok = false;
      else { signed8 = da.is_signed != 0; if (g_dt[out.dtype].is_signed != da.is_signed || da.value_bits > g_dt[out.dtype].value_bits || db.value_bits > g_dt[out.dtype].value_bits) ok = false; }
    }
    if (ok && wide) {
      // unsigned bytes (a scalar: any unsigned rank-0 array whose values keep the field arithmetic inside its 16 bits:
      // + and - up to 15 bits, * up to 8 bits so that 255 * v < 2^16); - is two's complement in the 16-bit field
      if (da.is_signed || db.is_signed) ok = false;
      if (ok && b_scalar && db.value_bits > (nop == 2 ? 8 : 15)) ok = false;
      if (ok && nop == 2 && !b_scalar) ok = false;                      // array * array: the general engine
      if (ok && nop == 1 && !g_dt[out.dtype].is_signed) ok = false;     // a difference needs a signed result
    }
    if (ok) {
      const unsigned long long mk = sk.mask & 0xffffull;
      const char* ap = ops[ia].base + ops[ia].offset;
      char* op = out.base + out.offset * (out_bits / 8);
      const bool swap = ia == 1;                                        // the streamed array is the SECOND argument of the function
      if (b_scalar) {
        if ((((uintptr_t)ap | (uintptr_t)op) & 15) == 0) {
          int rc = ew_narrow_launch(nop, wide, ap, ops[ib].base, op, d[0].ext, true, ops[ib].offset, dtype_load_kind(ops[ib].dtype), swap, signed8, (unsigned)(mk | (mk << 16)));
          if (rc != NCL_ERR_UNSUPPORTED) return rc;
        }
      } else {
        const char* p0 = ops[0].base + ops[0].offset; const char* p1 = ops[1].base + ops[1].offset;
        if ((((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)op) & 15) == 0) {
          int rc = ew_narrow_launch(nop, wide, p0, p1, op, d[0].ext, false, 0, 0, false, signed8, (unsigned)(mk | (mk << 16)));
          if (rc != NCL_ERR_UNSUPPORTED) return rc;
        }
      }
    }
  }
  // ---- widening integer family (k_ew_widen.cu): narrow int array (op) 64-bit operand -> 64-bit result, flat ----
  if (nd == 1 && !eflat && E == 2 && cc == CLS_I && out_cls == CLS_I && m.n == 2 && !is_cmp && d[0].ext / 2 < 0x7fffffffLL) {
    int wop = shape == S_ADD || shape == S_ADD0 ? 0 : shape == S_SUB ? 1 : shape == S_MUL || shape == S_MUL0 ? 2 : shape == S_MAX ? 3 : shape == S_MIN ? 4 : -1;
    int lk0 = dtype_load_kind(ops[0].dtype), lk1 = dtype_load_kind(ops[1].dtype);
    auto narrow = [&](int k, int lk) { return lk >= LK_U8 && lk <= LK_S32 && d[0].st[1 + k] == 1 && p.period[k] == 0; };
    auto wide = [&](int lk) { return lk == LK_T64 || lk == LK_W64; };
    int ia = narrow(0, lk0) && wide(lk1) ? 0 : (narrow(1, lk1) && wide(lk0) ? 1 : -1);
    if (wop >= 0 && ia >= 0) {
      const int ib = 1 - ia;
      int bmode = -1;
      if (p.period[ib] != 0) bmode = (128u % p.period[ib]) == 0 ? 1 : ((p.period[ib] % 128u) == 0 ? 2 : -1);
      else if (d[0].st[1 + ib] == 1) bmode = 0;
      else if (d[0].st[1 + ib] == 0) bmode = 3;
      if (bmode >= 0) {
        const int la = ia == 0 ? lk0 : lk1, lb = ia == 0 ? lk1 : lk0;
        const char* ap = ops[ia].base + (ops[ia].offset * g_dt[ops[ia].dtype].slot_bits) / 8;
        const char* bp = ops[ib].base + ops[ib].offset * 8;
        char* op = out.base + out.offset * 8;
        int rc = ew_widen_launch(la, wop, ap, bp, op, (unsigned)(d[0].ext / 2), bmode, p.period[ib], lb == LK_T64 ? 1 : 0, ia == 1 ? 1 : 0, sk.mask, sk.shift,
                                 &ops[ia], &ops[ib], &out);
        if (rc != NCL_ERR_UNSUPPORTED) return rc;
      }
    }
  }
  p.nd = nd; p.n_in = m.n; p.sk = sk;
  p.rt_op = m.rt_op; p.fault = c.fault;
  if (rt_family >= 0 && op_may_fault(m.rt_op)) c.fault_possible = true;
  for (int i = 0; i < nd; i++) {                      // reverse: params dim 0 is the innermost
    const D& s = d[nd - 1 - i];
    p.ext[i] = (unsigned)(i == 0 ? s.ext / E : s.ext);
    if (p.ext[i] <= 1) { p.mul[i] = 0; p.shr[i] = 0; }
    else {                                            // q / ext = umulhi(q, ceil(2^(31+s)/ext)) >> (s-1), s = ceil(log2 ext), q < 2^31
      unsigned s2 = 0; while ((1ull << s2) < p.ext[i]) s2++;
      p.mul[i] = (unsigned)(((1ull << (31 + s2)) + p.ext[i] - 1) / p.ext[i]); p.shr[i] = s2 - 1;
    }
    for (int k = 0; k < 4; k++) p.st[k][i] = (int)s.st[k];
  }
  p.base[0] = out.base; p.off[0] = out.offset;
  for (int k = 0; k < m.n; k++) { p.base[1 + k] = ops[k].base; p.off[1 + k] = ops[k].offset; p.lk[k] = dtype_load_kind(ops[k].dtype); }
  if (eflat) {
    p.eper = (unsigned)eflat_P;
    unsigned s2 = 0; while ((1ull << s2) < p.eper) s2++;
    p.epmul = (unsigned)(((1ull << (31 + s2)) + p.eper - 1) / p.eper); p.epshr = s2 - 1;
    for (int k = 0; k < m.n; k++) { p.gather[k] = gather[k]; p.gst[k] = (int)g_outer[k]; }
  }
  int U = cmppack ? cmp_U : ew_u_host(cc != CLS_I ? 0 : i32 ? 2 : 1, E);
  unsigned long long rows = 1; for (int i = 1; i < nd; i++) rows *= p.ext[i];
  unsigned segs = (p.ext[0] + 32 * U - 1) / (32 * U);
  unsigned long long nitems = rows * segs;
  if (nitems >= 0x7fffffffull || (unsigned long long)p.ext[0] * E >= 0x7fffffffull) return NCL_ERR_UNSUPPORTED;   // beyond 2^31: generic path
  if (nitems == 0) return NCL_OK;
  p.segs_per_row = segs; p.nitems = (unsigned)nitems;
  if (segs <= 1) { p.seg_mul = 0; p.seg_shr = 0; }
  else { unsigned s2 = 0; while ((1ull << s2) < segs) s2++; p.seg_mul = (unsigned)(((1ull << (31 + s2)) + segs - 1) / segs); p.seg_shr = s2 - 1; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kfn, EW_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
  unsigned long long blocks = (nitems + EW_THREADS / 32 - 1) / (EW_THREADS / 32);
  // Four waves instead of one when no operand is hoisted out of the item loop (measured on C1: 23.7 -> 22.5 us,
  // better tail balance); a hoisted row vector is re-read per CTA, and there one resident wave is faster (C5a).
  int over = 4;
  for (int k = 0; k < m.n; k++) if (p.period[k] != 0 && ((32u * U) % p.period[k]) == 0) over = 1;
  unsigned long long cap = (unsigned long long)c.sm_count * per_sm * over;
  if (blocks > cap) blocks = cap;
  // warp-invariant row vectors (EwParams::whoist): the vector spans `span` segments; with the warp count a multiple of span
  // every item of a warp sees the same piece
  p.whoist = 0;
  if (!cmppack && !eflat) {
    const unsigned wpb = EW_THREADS / 32, SEGc = 32u * (unsigned)U;
    unsigned long long need = 1;                                                   // blocks must be a multiple of this
    unsigned mask = 0;
    for (int k = 0; k < m.n; k++) {
      unsigned span = 0;
      // (flat mode with a period of several segments -- C1 -- was tried too: the per-item fetch of the 2 KB piece is an L2 hit
      // there, and keeping it in registers cost 1 %: 22.45 -> 22.70 us)
      if (nd == 1) continue;
      else if (p.period[k] == 0 && !p.gather[k] && p.st[1 + k][0] == 1) {
        bool rowvec = true; for (int dd = 1; dd < nd; dd++) if (p.st[1 + k][dd] != 0) rowvec = false;
        if (rowvec && segs > 1) span = segs;
      }
      if (!span) continue;
      unsigned long long g = span, x = wpb; while (x) { unsigned long long t = g % x; g = x; x = t; }     // gcd(span, warps per block)
      const unsigned long long nb = span / g;                                       // blocks * wpb % span == 0  <=>  blocks % nb == 0
      unsigned long long l = need, y = nb; while (y) { unsigned long long t = l % y; l = y; y = t; }
      const unsigned long long lcm = need / l * nb;
      if (lcm <= blocks && lcm <= 4096) { need = lcm; mask |= 1u << k; }
    }
    if (mask) { blocks = blocks / need * need; p.whoist = mask; }
  }
  const char* kname = cmppack ? "cmp_pack" : eflat ? "bcast_vec_flat" : E == 1 ? "bcast_scalar" : "bcast_vec";
  {
    // programmatic dependent launch (ncl_internal.h): ew_kernel reads its first item while the previous kernel drains and
    // writes nothing before griddepcontrol.wait, provided nothing it reads is written by a kernel that may still be in flight.
    // The packed-comparison kernels do not take part.
    PdlRange reads[3]; int nr = 0;
    for (int k = 0; k < m.n; k++) if (m.s[k].kind < 2) reads[nr++] = pdl_range(ops[k]);
    const bool early = !cmppack && pdl_eligible(reads, nr);
    void* args[1] = {(void*)&p};
    NCL_TRY(pdl_launch((const void*)kfn, dim3((unsigned)blocks), dim3(EW_THREADS), args, c.stream, early));
    note_launch(kname);
    if (!cmppack) { PdlRange w = pdl_range(out); pdl_commit(early, &w, 1); }
  }
  if (eflat && eflat_tail) {
    // fewer than E trailing elements: the end of the last partial row, then whole rows, as two tiny nests
    const long long P = eflat_P, R = eflat_R, fr = eflat_tail / P, part = eflat_tail - fr * P, body = P * R - eflat_tail;
    auto piece = [&](int loops, long long e0, long long e1, long long row, long long col, long long out_shift) {
      Nest x = n;
      x.nloops = loops; x.extent[0] = e0; x.extent[1] = e1;
      for (int i = 0; i < x.n_in; i++) {
        long long so = 0, si = 0;
        for (int k = 0; k < m.n; k++) if (m.s[k].kind == 0 && m.s[k].idx == i) { so = gather[k] ? g_outer[k] : P; si = gather[k] ? d[0].st[1 + k] : 1; }
        x.in[i].offset += row * so + col * si;
        if (loops == 1) x.in[i].stride[0] = si; else { x.in[i].stride[0] = so; x.in[i].stride[1] = si; }
      }
      x.out[0].offset += out_shift;
      if (loops == 1) x.out[0].stride[0] = 1; else { x.out[0].stride[0] = P; x.out[0].stride[1] = 1; }
      return x;
    };
    auto run_piece = [&](Nest& t) { int rc = try_elementwise(t); return rc == NCL_ERR_UNSUPPORTED ? launch_generic(t) : rc; };
    if (part) { Nest t = piece(1, part, 1, R - fr - 1, P - part, body); NCL_TRY(run_piece(t)); }
    if (fr) { Nest t = piece(2, fr, P, R - fr, 0, body + part); NCL_TRY(run_piece(t)); }
  }
  c.last_kernel = kname;                                  // (the tail pieces above launched kernels of their own)
  return NCL_OK;
}
