// k_reduce.cu -- host side of K3: nest -> reduction form -> kernel (kernels and their documentation: k_reduce.cuh).
#include "k_reduce.cuh"

static RedKernel kernel_for_dot(int cls, int E, int which) {
  if (cls == CLS_F || cls == CLS_D) return red_pick_fd_dot(cls, E, which);
  return E >= 4 ? red_pick_int_narrow_dot(E, which) : red_pick_int_wide_dot(E, which);
}
static RedKernel kernel_for(int cls, int E, int op, int which) {
  if (cls == CLS_F || cls == CLS_D) return red_pick_fd(cls, E, op, which);
  return E >= 4 ? red_pick_int_narrow(E, op, which) : red_pick_int_wide(E, op, which);
}

static bool match_reduce(const Expr& e, int* op) {
  if (e.n != 3) return false;
  const XNode& r = e.node[2];
  int o;
  switch (r.op) { case NCL_OP_ADD: o = R_ADD; break; case NCL_OP_MUL: o = R_MUL; break; case NCL_OP_MAX: o = R_MAX; break; case NCL_OP_MIN: o = R_MIN; break; default: return false; }
  const XNode& a = e.node[r.a]; const XNode& b = e.node[r.b];
  bool ok = (a.op == NCL_X_OUTPUT && a.a == 1 && b.op == NCL_X_INPUT && b.a == 1) ||
            (o <= R_MUL && b.op == NCL_X_OUTPUT && b.a == 1 && a.op == NCL_X_INPUT && a.a == 1);
  if (!ok) return false;
  *op = o; return true;
}

struct RDim { long long ext, sin, sout, sin2; };       // sin2: stride of the second operand (DOT), 0 otherwise
static void sort_merge(RDim* d, int& n, bool kept) {
  for (int i = 1; i < n; i++) { RDim t = d[i]; int j = i - 1; while (j >= 0 && llabs(d[j].sin) < llabs(t.sin)) { d[j + 1] = d[j]; j--; } d[j + 1] = t; }
  for (int i = n - 1; i > 0; i--) {
    bool ok = d[i - 1].sin == d[i].sin * d[i].ext && (!kept || d[i - 1].sout == d[i].sout * d[i].ext) && d[i - 1].sin2 == d[i].sin2 * d[i].ext;
    if (ok) { d[i - 1].ext *= d[i].ext; d[i - 1].sin = d[i].sin; d[i - 1].sout = d[i].sout; d[i - 1].sin2 = d[i].sin2; for (int j = i; j + 1 < n; j++) d[j] = d[j + 1]; n--; }
  }
}

static int fill_common(RedParams& p, const Nest& n, int op, int ia) {
  const Operand& in = n.in[ia]; const Operand& out = n.out[0];
  memset(&p, 0, sizeof p);
  p.in = in.base; p.in_off = in.offset; p.lk = dtype_load_kind(in.dtype);
  p.out = out.base; p.out_off = out.offset; p.sk = dtype_store_spec(out.dtype); p.out_lk = dtype_load_kind(out.dtype);
  p.fresh = n.fresh ? 1 : 0;
  if (n.fresh) {
    if (n.has_identity) { p.ident_i = n.ident_i; p.ident_f = n.ident_f; }
    else { p.ident_i = 0; p.ident_f = 0.0; }     // einsum into a fresh `zeros` output
  }
  p.stat = n.stat; p.count = n.stat_count;
  (void)op; return NCL_OK;
}

// one resident wave: sm_count x (blocks of 256 threads that fit one SM for this kernel)
static long long persistent_blocks(const void* kfn) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  return (long long)ctx().sm_count * per_sm;
}

static bool aligned16(const Operand& o, long long elem_off) {
  int sb = g_dt[o.dtype].slot_bits;
  return (((uintptr_t)o.base + (elem_off * sb) / 8) % 16) == 0;
}

static int reduce_impl(const Nest& n, int op, int ia, int ib);

// Launch of a streaming reduction kernel that takes part in programmatic dependent launch (ncl_internal.h): `reads` are
// the arrays it starts streaming right away, `write` the array it stores into after griddepcontrol.wait.
static int launch_red(RedKernel kfn, dim3 grid, const RedParams& p, const Operand* in_a, const Operand* in_b, const Operand& out, const char* name) {
  PdlRange reads[2]; int nr = 0;
  if (in_a) reads[nr++] = pdl_range(*in_a);
  if (in_b) reads[nr++] = pdl_range(*in_b);
  const bool early = pdl_eligible(reads, nr);
  void* args[1] = {(void*)&p};
  NCL_TRY(pdl_launch((const void*)kfn, grid, dim3(256), args, ctx().stream, early));
  note_launch(name);
  PdlRange w = pdl_range(out);
  pdl_commit(early, &w, 1);
  return NCL_OK;
}

// Reduced loops that do not form ONE run of neighbouring loops of the input -- `(ijk -> j)`, `(ijkl -> jl)`: the kernels merge the
// reduced loops into one dimension and cannot, so such a nest fell to the generic kernel, where one thread walks all of
// the reduced loops of its result element (85 ms for 256^3 single-floats).  The innermost run is reduced into a temporary
// of the result's element type first, the rest from there: fast passes only.  Exact for integers (modular) and max / min;
// floats round once more between the passes, within the reduction tolerance (ncl_reduce does the same for :axes (0 2)).
static int reduce_two_pass(const Nest& n, int op) {
  const Operand& in = n.in[0]; const Operand& out = n.out[0];
  if (n.stat || g_dt[in.dtype].cls != g_dt[out.dtype].cls || g_dt[in.dtype].slot_bits < 8 || ctx().capturing) return NCL_ERR_UNSUPPORTED;
  int ord[NCL_MAX_LOOPS], nl = 0;
  for (int l = 0; l < n.nloops; l++) { if (n.extent[l] <= 0) return NCL_ERR_UNSUPPORTED; if (n.extent[l] > 1) { if (in.stride[l] <= 0) return NCL_ERR_UNSUPPORTED; ord[nl++] = l; } }
  for (int i = 1; i < nl; i++) { int k = ord[i], j = i - 1; while (j >= 0 && in.stride[ord[j]] < in.stride[k]) { ord[j + 1] = ord[j]; j--; } ord[j + 1] = k; }
  int runs = 0, last0 = -1;
  for (int i = 0; i < nl; i++) { const bool red = out.stride[ord[i]] == 0; if (red && (i == 0 || out.stride[ord[i - 1]] != 0)) { runs++; last0 = i; } }
  if (runs < 2) return NCL_ERR_UNSUPPORTED;
  int last1 = last0; while (last1 + 1 < nl && out.stride[ord[last1 + 1]] == 0) last1++;
  // temporary over every loop outside the innermost reduced run, row-major in the input's order
  long long tstride[NCL_MAX_LOOPS], total = 1;
  for (int i = nl - 1; i >= 0; i--) if (i < last0 || i > last1) { tstride[i] = total; total *= n.extent[ord[i]]; }
  ncl_buf* tb = ncl_alloc(ncl_storage_bytes(out.dtype, total), 0);
  if (!tb) return NCL_ERR_NOMEM;
  Operand tmp; memset(&tmp, 0, sizeof tmp);
  tmp.base = (char*)ncl_buf_ptr(tb); tmp.dtype = out.dtype; tmp.offset = 0; tmp.stage_item = -1;
  tmp.n_storage = (int64_t)((ncl_buf_bytes(tb) * 8) / g_dt[out.dtype].slot_bits);
  const int nop = op == R_ADD ? NCL_OP_ADD : op == R_MUL ? NCL_OP_MUL : op == R_MAX ? NCL_OP_MAX : NCL_OP_MIN;
  Nest* p1 = new Nest(n);                                   // pass 1: the innermost reduced run -> tmp
  p1->nloops = nl; p1->fresh = true; p1->has_identity = true;
  { long long ii; double ff; reduce_identity(nop, out.dtype, &ii, &ff); p1->ident_i = ii; p1->ident_f = ff; }
  p1->out[0] = tmp;
  for (int i = 0; i < nl; i++) { p1->extent[i] = n.extent[ord[i]]; p1->in[0].stride[i] = in.stride[ord[i]]; p1->out[0].stride[i] = (i < last0 || i > last1) ? tstride[i] : 0; }
  int rc = try_reduce(*p1);
  if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*p1);
  delete p1;
  if (rc == NCL_OK) {
    Nest* p2 = new Nest(n);                                 // pass 2: the rest, from tmp into the result
    int k = 0;
    p2->in[0] = tmp;
    for (int i = 0; i < nl; i++) if (i < last0 || i > last1) { p2->extent[k] = n.extent[ord[i]]; p2->in[0].stride[k] = tstride[i]; p2->out[0].stride[k] = out.stride[ord[i]]; k++; }
    p2->nloops = k;
    rc = try_reduce(*p2);
    if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*p2);
    delete p2;
  }
  ncl_free(tb);
  return rc;
}

int try_reduce(const Nest& n) {
  if (n.n_out != 1 || n.n_in != 1) return NCL_ERR_UNSUPPORTED;
  int op;
  if (!match_reduce(n.transform[0], &op)) return NCL_ERR_UNSUPPORTED;
  int rc = reduce_impl(n, op, 0, -1);
  if (rc == NCL_ERR_UNSUPPORTED) rc = reduce_two_pass(n, op);
  return rc;
}

// Contractions that are really reductions: the default transform (+ @1 (* $1 $2)) over two arrays of one element type
// where no loop is private to one input and kept (no M x N outer structure, k_gemm.cu): vdot / inner `(i i ->)`, row-wise
// dots `(ij ij -> i)`, matrix x vector `(ij j -> i)`, vector x matrix `(i ij -> j)`.  Every element of the larger input is
// used once, so they are HBM-bound streams and run on the reduction kernels with the product formed on load.
int try_dot_reduce(const Nest& n) {
  if (n.n_in == 1 && n.n_out == 1 && n.transform[0].n == 5) {
    // (+ @1 (* $1 $1)): the sum of squares / squared norm of ONE array, `(i -> (+ @1 (* $1 $1)) -> )`.  The same array is both
    // operands of the dot-like reduction (it used to fall to the generic kernel: one thread for the whole norm).
    const Expr& e1 = n.transform[0];
    const XNode& root = e1.node[4];
    if (root.op == NCL_OP_ADD && e1.node[root.a].op == NCL_X_OUTPUT && e1.node[root.b].op == NCL_OP_MUL) {
      const XNode& mul = e1.node[root.b];
      if (e1.node[mul.a].op == NCL_X_INPUT && e1.node[mul.b].op == NCL_X_INPUT && e1.node[mul.a].a == 1 && e1.node[mul.b].a == 1) {
        Nest* t = new Nest(n);
        t->n_in = 2; t->in[1] = n.in[0];
        t->transform[0].node[mul.b].a = 2;
        int rc = try_dot_reduce(*t);
        delete t; return rc;
      }
    }
  }
  if (n.n_in != 2 || n.n_out != 1) return NCL_ERR_UNSUPPORTED;
  const Expr& e = n.transform[0];
  if (e.n != 5) return NCL_ERR_UNSUPPORTED;
  const XNode& root = e.node[4];
  if (root.op != NCL_OP_ADD || e.node[root.a].op != NCL_X_OUTPUT || e.node[root.b].op != NCL_OP_MUL) return NCL_ERR_UNSUPPORTED;
  const XNode& mul = e.node[root.b];
  if (e.node[mul.a].op != NCL_X_INPUT || e.node[mul.b].op != NCL_X_INPUT || e.node[mul.a].a == e.node[mul.b].a) return NCL_ERR_UNSUPPORTED;
  const int i0 = e.node[mul.a].a - 1, i1 = e.node[mul.b].a - 1;
  const Operand& A = n.in[i0]; const Operand& B = n.in[i1];
  if (A.dtype != B.dtype) return NCL_ERR_UNSUPPORTED;
  bool a_all = true, b_all = true, m_loop = false, n_loop = false, k_loop = false;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] == 1) continue;
    const bool a = A.stride[l] != 0, b = B.stride[l] != 0, c = n.out[0].stride[l] != 0;
    if (!a) a_all = false;
    if (!b) b_all = false;
    if (a && !b && c) m_loop = true;
    if (!a && b && c) n_loop = true;
    if (!c) k_loop = true;
    if (!a && !b) return NCL_ERR_UNSUPPORTED;
  }
  if (!k_loop || (m_loop && n_loop) || (!a_all && !b_all)) return NCL_ERR_UNSUPPORTED;
  return a_all ? reduce_impl(n, R_ADD, i0, i1) : reduce_impl(n, R_ADD, i1, i0);
}

// ia: the array every loop moves; ib: second operand of a DOT (or -1)
static int reduce_impl(const Nest& n, int op, int ia, int ib) {
  const bool dot = ib >= 0;
  const Operand& in = n.in[ia]; const Operand& out = n.out[0];
  int icls = g_dt[in.dtype].cls, ocls = g_dt[out.dtype].cls;
  if (n.stat) { if (!((icls == CLS_I && ocls == CLS_F) || (icls == ocls && icls != CLS_I))) return NCL_ERR_UNSUPPORTED; }     // mean of integers is a single-float
  else if (icls != ocls) return NCL_ERR_UNSUPPORTED;                  // e.g. integer sum into a float array: generic path
  if (g_dt[in.dtype].slot_bits < 8) return NCL_ERR_UNSUPPORTED;       // packed inputs: generic path
  if (n.fresh && !n.has_identity && op != R_ADD) return NCL_ERR_UNSUPPORTED;   // (* @1 $1) into zeros etc.: generic defines it
  RDim kd[NCL_MAX_LOOPS], rd[NCL_MAX_LOOPS]; int nk = 0, nr = 0;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] == 1) continue;
    RDim d = {n.extent[l], in.stride[l], out.stride[l], dot ? n.in[ib].stride[l] : 0};
    if (out.stride[l] != 0) kd[nk++] = d; else rd[nr++] = d;
  }
  if (nr == 0) return NCL_ERR_UNSUPPORTED;
  for (int i = 0; i < nk; i++) if (kd[i].sin == 0) return NCL_ERR_UNSUPPORTED;
  for (int i = 0; i < nr; i++) if (rd[i].sin == 0) return NCL_ERR_UNSUPPORTED;   // broadcast input under a reduction: generic
  sort_merge(kd, nk, true); sort_merge(rd, nr, false);
  if (nr != 1) return NCL_ERR_UNSUPPORTED;
  Context& c = ctx();
  int sb = g_dt[in.dtype].slot_bits;
  int Evec = 128 / sb;
  RedParams p; fill_common(p, n, op, ia);
  long long R = rd[0].ext, sR = rd[0].sin;
  int cls = icls;
  if (dot) {
    const Operand& b = n.in[ib];
    p.in2 = b.base; p.in2_off = b.offset; p.lk2 = dtype_load_kind(b.dtype);
    p.r_stride2 = rd[0].sin2;
  }
  // second operand: vector loads need the same 16-byte phase as the first along every loop it moves with
  auto b_vec_ok = [&](long long s2) { return (((s2 * sb) / 8) % 16) == 0; };
  auto kfor = [&](int E, int which) { return dot ? kernel_for_dot(cls, E, which) : kernel_for(cls, E, op, which); };

  if (sR == 1 && nk == 0) {
    // ---- full reduce --------------------------------------------------------------------------
    p.cols = R;
    if (dot && rd[0].sin2 != 1) return NCL_ERR_UNSUPPORTED;
    p.vec = R >= Evec && aligned16(in, in.offset) && (!dot || aligned16(n.in[ib], n.in[ib].offset));   // a ragged tail (< Evec elements) is folded by block 0
    int E = p.vec ? Evec : 1;
    long long work = p.vec ? R / E : R;
    RedKernel kfn = kfor(E, KALL);
    long long blocks = (work + 256 * 4 - 1) / (256 * 4);
    long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
    NCL_TRY(ensure_scratch((size_t)blocks * 8));
    p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks;
    return launch_red(kfn, dim3((unsigned)blocks), p, &in, dot ? &n.in[ib] : nullptr, out, dot ? "dot_all" : "reduce_all");
  }
  // the contiguous kept dim must be the innermost of the kept group for the column form
  bool inner_reduced = (sR == 1);
  if (inner_reduced) {
    // ---- row reduce: kept dims (any strides) x contiguous reduced axis ------------------------------
    if (nk > 4) return NCL_ERR_UNSUPPORTED;
    p.nk = nk; p.rows = 1;
    for (int i = 0; i < nk; i++) { p.kext[i] = kd[i].ext; p.kin[i] = kd[i].sin; p.kout[i] = kd[i].sout; p.kin2[i] = kd[i].sin2; p.rows *= kd[i].ext; }
    p.cols = R;
    if (dot && rd[0].sin2 != 1 && rd[0].sin2 != 0) return NCL_ERR_UNSUPPORTED;
    // ---- few long rows ((sum a :axes 1) of a (3,N) array): one CTA per row would leave the GPU empty.  Every row is cut into
    // S pieces of L elements (L a multiple of the chunk): [rows x S x L] reduces over L into a scratch [rows x S], which then
    // reduces over S into the result; the R - S*L trailing elements of every row are accumulated last.
    if (!dot && !n.stat && p.rows < 2LL * c.sm_count && R >= 65536 && (n.fresh || op == R_ADD) && nk <= 3) {
      long long S = (8LL * c.sm_count + p.rows - 1) / p.rows; if (S > R / 16384) S = R / 16384;
      const long long L = R / S / (4 * Evec) * (4 * Evec), rem = R - S * L;
      if (S >= 2 && L > 0) {
        const int osb = g_dt[out.dtype].slot_bits;
        NCL_TRY(ensure_scratch3((size_t)(p.rows * S) * osb / 8 + 64));
        auto run = [&](Nest& x) { int r2 = try_reduce(x); return r2 == NCL_ERR_UNSUPPORTED ? launch_generic(x) : r2; };
        long long tst[4]; { long long acc = S; for (int i = nk - 1; i >= 0; i--) { tst[i] = acc; acc *= kd[i].ext; } }
        Nest a = n;                                         // kept..., S, L -> scratch[kept..., S]
        a.nloops = nk + 2;
        for (int i = 0; i < nk; i++) { a.extent[i] = kd[i].ext; a.in[0].stride[i] = kd[i].sin; a.out[0].stride[i] = tst[i]; }
        a.extent[nk] = S; a.in[0].stride[nk] = L; a.out[0].stride[nk] = 1;
        a.extent[nk + 1] = L; a.in[0].stride[nk + 1] = 1; a.out[0].stride[nk + 1] = 0;
        a.out[0].base = (char*)c.scratch3; a.out[0].offset = 0; a.out[0].n_storage = p.rows * S;
        a.fresh = true; if (!n.fresh) a.has_identity = false;
        NCL_TRY(run(a));
        Nest b = n;                                         // scratch[kept..., S] -> result
        b.nloops = nk + 1;
        for (int i = 0; i < nk; i++) { b.extent[i] = kd[i].ext; b.in[0].stride[i] = tst[i]; b.out[0].stride[i] = kd[i].sout; }
        b.extent[nk] = S; b.in[0].stride[nk] = 1; b.out[0].stride[nk] = 0;
        b.in[0].base = (char*)c.scratch3; b.in[0].dtype = out.dtype; b.in[0].offset = 0; b.in[0].n_storage = p.rows * S;
        NCL_TRY(run(b));
        if (rem > 0) {
          Nest r3 = n;
          r3.nloops = nk + 1;
          for (int i = 0; i < nk; i++) { r3.extent[i] = kd[i].ext; r3.in[0].stride[i] = kd[i].sin; r3.out[0].stride[i] = kd[i].sout; }
          r3.extent[nk] = rem; r3.in[0].stride[nk] = 1; r3.out[0].stride[nk] = 0;
          r3.in[0].offset += S * L; r3.fresh = false;
          NCL_TRY(run(r3));
        }
        return NCL_OK;
      }
    }
    bool vec = (R % Evec == 0) && aligned16(in, in.offset);
    for (int i = 0; i < nk && vec; i++) if (((kd[i].sin * sb) / 8) % 16) vec = false;
    if (dot && vec && rd[0].sin2 == 1) { vec = aligned16(n.in[ib], n.in[ib].offset); for (int i = 0; i < nk && vec; i++) if (!b_vec_ok(kd[i].sin2)) vec = false; }
    // rows that do not start / end on a chunk boundary: scalar head and tail around an aligned vector body (vec = 2),
    // unless they are so short that one thread folds a whole row element by element
    if (dot && !n.stat && vec && nk == 1 && kd[0].sin2 == 0 && rd[0].sin2 == 1 && R / Evec >= 512 && p.rows >= 4 && (cls == CLS_F || cls == CLS_D)) {
      p.vec = 1;
      const unsigned blocks4 = (unsigned)((p.rows + 3) / 4);
      red_dot_matvec_launch(cls, blocks4, c.stream, p);
      note_launch("dot_matvec");
      return cuda_fail(cudaGetLastError(), "dot_matvec_kernel");
    }
    // ---- rows of 2..8 elements of at least 4 bytes, back to back: the register kernel -------------------------------------------
    if (!dot && nk == 1 && kd[0].sin == R && R >= 2 && R <= 8 && sb >= 32 && p.rows >= 1024 && aligned16(in, in.offset)) {
      RedKernel kfn = red_rows_reg_pick(cls, sb, op, (int)R);
      const long long threads = (p.rows + Evec - 1) / Evec, blocks = (threads + 255) / 256;
      if (kfn && blocks <= 0x7fffffffLL) {
        p.vec = 1;
        kfn<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
        note_launch("reduce_rows_reg");
        return cuda_fail(cudaGetLastError(), "reduce_rows_reg_kernel");
      }
    }
    // ---- rows back to back, at most 2 KB long and not on chunk boundaries: the staged kernel -------------------------------
    if (!dot && !vec && nk == 1 && kd[0].sin == R && R * sb / 8 <= 2048 && p.rows >= 64 && aligned16(in, in.offset)) {
      const long long row_bytes = R * sb / 8;
      long long RT = 32768 / row_bytes / Evec * Evec;
      p.seg = RT; p.nseg = R <= 8 ? 1 : R <= 16 ? 2 : R <= 32 ? 4 : R <= 64 ? 8 : R <= 128 ? 16 : 32;
      RedKernel kfn = red_staged_pick(cls, op);
      long long tiles = (p.rows + RT - 1) / RT, cap = persistent_blocks((const void*)kfn);
      kfn<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, c.stream>>>(p);
      note_launch("reduce_rows_staged");
      return cuda_fail(cudaGetLastError(), "reduce_rows_staged_kernel");
    }
    int vmode = vec ? 1 : ((!dot && R > 16 && Evec > 1 && ((uintptr_t)in.base % 16) == 0) ? 2 : 0);
    p.vec = vmode; int E = vmode ? Evec : 1;
    long long chunks = vmode ? R / E : R;
    int which; long long blocks;
    if (vmode == 0 && R <= 16) { which = KROW1; blocks = (p.rows + 1023) / 1024; }
    else if (vmode == 1 && chunks <= 4) { which = KROW1; blocks = (p.rows + 1023) / 1024; }
    else if (chunks <= 32 && vmode) { which = KROW4; blocks = (p.rows + 63) / 64; }      // up to 8 chunks per thread, 4 + 3 in flight
    else if (chunks >= 512) { which = KROW256; blocks = p.rows; }
    else { which = KROW32; blocks = (p.rows + 7) / 8; }
    if (blocks > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
    const char* kname = dot ? "dot_rows" : which == KROW256 ? "reduce_row_block" : which == KROW32 ? "reduce_row_warp" : which == KROW4 ? "reduce_row_quad" : "reduce_row_thread";
    if (which != KROW1) return launch_red(kfor(E, which), dim3((unsigned)blocks), p, &in, dot ? &n.in[ib] : nullptr, out, kname);
    kfor(E, which)<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
    note_launch(kname);
    return cuda_fail(cudaGetLastError(), "reduce_row_kernel");
  }
  // ---- column reduce: [outer kept...][R][contiguous kept] ---------------------------------------------
  if (nk < 1 || kd[nk - 1].sin != 1 || kd[nk - 1].sout != 1) return NCL_ERR_UNSUPPORTED;
  if (nk - 1 > 4) return NCL_ERR_UNSUPPORTED;
  long long ncols = kd[nk - 1].ext;
  p.ncols = ncols; p.r_ext = R; p.r_stride = sR;
  p.nk = nk - 1; p.rows = 1;
  for (int i = 0; i < nk - 1; i++) { p.kext[i] = kd[i].ext; p.kin[i] = kd[i].sin; p.kout[i] = kd[i].sout; p.kin2[i] = kd[i].sin2; p.rows *= kd[i].ext; }
  if (dot) { p.c_stride2 = kd[nk - 1].sin2; if (p.c_stride2 != 0 && p.c_stride2 != 1) return NCL_ERR_UNSUPPORTED; }
  // ---- few columns ((sum a :axes 0) of an (N,3) array): the whole array as one stream, reduce_fewcols_kernel ----------
  if (!dot && !n.stat && nk == 1 && sR == ncols && ncols <= 128 && Evec > 1 && aligned16(in, in.offset) && R * ncols >= 32768 && R * ncols < (1LL << 40)) {
    const long long Rr = (R * ncols) % Evec ? R % Evec : 0, Rm = R - Rr;      // Rm * ncols is a multiple of the chunk
    RedKernel kfn = kernel_for(cls, Evec, op, KFEWCOLS);
    if (kfn && Rm > 0) {
      p.cols = Rm * ncols; p.vec = 1; p.kout[0] = kd[0].sout;
      long long work = p.cols / Evec, blocks = (work + 1023) / 1024;
      long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap;
      long long g = ncols, x = 256LL * Evec; while (x) { long long t = g % x; g = x; x = t; }     // gcd(ncols, 256 * Evec)
      const long long m = ncols / g;                                        // grid must be a multiple of m
      blocks = blocks / m * m; if (blocks < m) blocks = m;
      NCL_TRY(ensure_scratch((size_t)blocks * ncols * 8));
      p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks;
      kfn<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
      note_launch("reduce_fewcols");
      NCL_TRY(cuda_fail(cudaGetLastError(), "reduce_fewcols_kernel"));
      if (Rr > 0) {                                                         // fewer than Evec left-over rows, accumulated
        Nest r3 = n;
        r3.nloops = 2; r3.extent[0] = Rr; r3.extent[1] = ncols;
        r3.in[0].offset += Rm * ncols; r3.in[0].stride[0] = ncols; r3.in[0].stride[1] = 1;
        r3.out[0].stride[0] = 0; r3.out[0].stride[1] = kd[0].sout;
        r3.fresh = false;
        int rc = try_reduce(r3);
        if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(r3);
        NCL_TRY(rc);
      }
      return NCL_OK;
    }
  }
  bool vec = (ncols % Evec == 0) && aligned16(in, in.offset) && ((sR * sb) / 8) % 16 == 0;
  for (int i = 0; i < nk - 1 && vec; i++) if (((kd[i].sin * sb) / 8) % 16) vec = false;
  if (dot && vec && p.c_stride2 == 1) { vec = aligned16(n.in[ib], n.in[ib].offset) && b_vec_ok(rd[0].sin2); for (int i = 0; i < nk - 1 && vec; i++) if (!b_vec_ok(kd[i].sin2)) vec = false; }
  // ---- a plain matrix whose rows are not whole chunks: rows dealt into Evec alignment classes (k_reduce_rot.cu) ---------------
  if (!vec && !dot && nk == 1 && (sb == 32 || sb == 64) && Evec > 1 && ncols >= 256 && R >= 8 * Evec && sR > 0 && ((uintptr_t)in.base % (sb / 8)) == 0) {
    RedKernel kmain, kfin; red_col_rot_pick(cls, sb, op, &kmain, &kfin);
    const long long nch = (ncols + 2 * Evec - 2) / Evec, bx = (nch + 255) / 256;      // chunk positions: up to Evec - 1 elements of slack on both sides
    const long long want = persistent_blocks((const void*)kmain);
    long long ny = (want + bx - 1) / bx; if (ny < Evec) ny = Evec;                      // CTAs along y: one resident wave
    long long nseg = ny / Evec; const long long maxseg = R / (8 * Evec); if (nseg > maxseg) nseg = maxseg; if (nseg < 1) nseg = 1; if (nseg > 4096) nseg = 4096;
    long long seg = (R + nseg - 1) / nseg; seg = (seg + Evec - 1) / Evec * Evec;        // whole groups of Evec rows
    nseg = (R + seg - 1) / seg;
    if (bx <= 0x7fffffffLL && nseg * Evec <= 65535) {
      p.seg = seg; p.nseg = (int)nseg; p.kout[0] = kd[0].sout;
      NCL_TRY(ensure_scratch((size_t)nseg * Evec * ncols * 8));
      p.partial = c.scratch;
      kmain<<<dim3((unsigned)bx, (unsigned)(nseg * Evec)), 256, 0, c.stream>>>(p);
      note_launch("reduce_col_rot");
      NCL_TRY(cuda_fail(cudaGetLastError(), "reduce_col_rot_kernel"));
      kfin<<<(unsigned)((ncols + 255) / 256), 256, 0, c.stream>>>(p);
      note_launch("reduce_col_rot");
      return cuda_fail(cudaGetLastError(), "reduce_col_rot_finish_kernel");
    }
  }
  int E = vec ? Evec : 1; p.vec = vec;
  long long nq = p.rows * (ncols / E);
  // split R so that the grid is ONE resident wave of equal segments (2048 threads per SM assumed here before: two
  // uneven waves at 60 registers)
  long long want = persistent_blocks((const void*)kfor(E, KCOL)) * 256;
  int nseg = 1;
  if (nq < want && R >= 128) { nseg = (int)((want + nq - 1) / nq); long long maxseg = R / 32; if (nseg > maxseg) nseg = (int)maxseg; if (nseg > 1024) nseg = 1024; if (nseg < 1) nseg = 1; }
  p.nseg = nseg; p.seg = (R + nseg - 1) / nseg;
  p.nseg = (int)((R + p.seg - 1) / p.seg);
  long long bx = (nq + 255) / 256;
  if (bx > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
  if (p.nseg > 1) { NCL_TRY(ensure_scratch((size_t)p.nseg * nq * E * 8)); p.partial = c.scratch; }
  dim3 grid((unsigned)bx, (unsigned)p.nseg);
  kfor(E, KCOL)<<<grid, 256, 0, c.stream>>>(p);
  note_launch(dot ? "dot_cols" : "reduce_col");
  NCL_TRY(cuda_fail(cudaGetLastError(), "reduce_col_kernel"));
  if (p.nseg > 1) {
    kfor(E, KCOLFIN)<<<(unsigned)((nq * E + 255) / 256), 256, 0, c.stream>>>(p);
    note_launch(dot ? "dot_cols" : "reduce_col");
    NCL_TRY(cuda_fail(cudaGetLastError(), "reduce_col_finish_kernel"));
  }
  return NCL_OK;
}

// sharded full reduce with the exchange fused into the kernel (peer mailboxes over NVLink)
int reduce_all_p2p(int op, const Operand& a, int64_t count, const Operand& r, unsigned long long* const* peers, int rank, int world,
                   unsigned long long seq) {
  Context& c = ctx();
  int cls = g_dt[a.dtype].cls; int sb = g_dt[a.dtype].slot_bits;
  if (sb < 8 || world > 63) return NCL_ERR_UNSUPPORTED;
  int rop = op == NCL_OP_ADD ? R_ADD : op == NCL_OP_MUL ? R_MUL : op == NCL_OP_MAX ? R_MAX : op == NCL_OP_MIN ? R_MIN : -1;
  if (rop < 0) return NCL_ERR_UNSUPPORTED;
  RedParams p; memset(&p, 0, sizeof p);
  p.in = a.base; p.in_off = a.offset; p.lk = dtype_load_kind(a.dtype); p.cols = count;
  p.out = r.base; p.out_off = r.offset; p.sk = dtype_store_spec(r.dtype); p.out_lk = dtype_load_kind(r.dtype);
  p.fresh = 1; reduce_identity(op, r.dtype, &p.ident_i, &p.ident_f);
  int Evec = 128 / sb;
  p.vec = count >= Evec && aligned16(a, a.offset);
  int E = p.vec ? Evec : 1;
  long long work = p.vec ? count / E : count;
  RedKernel kfn = kernel_for(cls, E, rop, KALL);
  long long blocks = (work + 1023) / 1024; long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
  NCL_TRY(ensure_scratch((size_t)blocks * 8));
  p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks;
  p.peers = peers; p.rank = rank; p.world = world; p.seq = seq;
  Operand ain = a; if (ain.n_storage == 0) ain.n_storage = a.offset + count;
  Operand rout = r; if (rout.n_storage == 0) rout.n_storage = r.offset + 1;
  return launch_red(kfn, dim3((unsigned)blocks), p, &ain, nullptr, rout, "reduce_all_p2p");
}

// local phase of a sharded full reduce: contiguous a -> one float64 / int64 partial on the device
int reduce_all_to_partial(int op, const Operand& a, int64_t count, void* partial8) {
  Context& c = ctx();
  int cls = g_dt[a.dtype].cls; int sb = g_dt[a.dtype].slot_bits;
  if (sb < 8) return set_error(NCL_ERR_UNSUPPORTED, "sharded reduce of packed arrays");
  int rop = op == NCL_OP_ADD ? R_ADD : op == NCL_OP_MUL ? R_MUL : op == NCL_OP_MAX ? R_MAX : op == NCL_OP_MIN ? R_MIN : -1;
  if (rop < 0) return set_error(NCL_ERR_UNSUPPORTED, "reduce op");
  RedParams p; memset(&p, 0, sizeof p);
  p.in = a.base; p.in_off = a.offset; p.lk = dtype_load_kind(a.dtype); p.cols = count;
  int Evec = 128 / sb;
  p.vec = count >= Evec && aligned16(a, a.offset);
  int E = p.vec ? Evec : 1;
  long long work = p.vec ? count / E : count;
  RedKernel kfn = kernel_for(cls, E, rop, KALL);
  long long blocks = (work + 1023) / 1024; long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
  NCL_TRY(ensure_scratch((size_t)blocks * 8));
  p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks; p.raw_partial = partial8; p.fresh = 1;
  reduce_identity(op, a.dtype, &p.ident_i, &p.ident_f);       // an empty slab contributes the identity (max / min seed)
  kfn<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  note_launch("reduce_all");
  return cuda_fail(cudaGetLastError(), "reduce_all_kernel");
}
