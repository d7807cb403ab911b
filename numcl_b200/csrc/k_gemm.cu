// k_gemm.cu -- K4 dispatch: recognise a dense contraction in a nest and route it.
//
// A nest is a (batched) GEMM when it has two inputs, one output, the default transform
// (+ @1 (* $1 $2)) (src/3einsum.lisp:176-178) and every loop falls into exactly one class by which
// arrays it moves:  batch (A, B, C) | M (A, C) | N (B, C) | K (A, B).  Each class must merge into
// one strided dimension, with K contiguous in A, N contiguous in B and C (row-major operands, the
// layout of `matmul` and `(bij bjk -> bik)`).  Anything else stays on the generic nest kernel.
//   double-float            -> DMMA          (k_gemm_f64.cu)
//   single-float, large     -> 3xTF32 tcgen05 (k_gemm_tf32.cu: pack + GEMM, or the split fused into the GEMM
//                              -- on CTA pairs when M % 256 == 0 -- for small tile-aligned matrices)
//   single-float skinny/odd, all integer types -> FMA / IMAD tiles (k_gemm_simt.cu)
#include "k_gemm.cuh"

struct GDim { long long ext, sa, sb, sc; };

static bool merge_group(GDim* d, int& n) {
  // sort by |sc| (or |sa| for K, whose sc is 0) descending, then merge neighbours
  auto key = [](const GDim& x) { return x.sc != 0 ? llabs(x.sc) : llabs(x.sa); };
  for (int i = 1; i < n; i++) { GDim t = d[i]; int j = i - 1; while (j >= 0 && key(d[j]) < key(t)) { d[j + 1] = d[j]; j--; } d[j + 1] = t; }
  for (int i = n - 1; i > 0; i--) {
    if (d[i - 1].sa == d[i].sa * d[i].ext && d[i - 1].sb == d[i].sb * d[i].ext && d[i - 1].sc == d[i].sc * d[i].ext) {
      d[i - 1].ext *= d[i].ext; d[i - 1].sa = d[i].sa; d[i - 1].sb = d[i].sb; d[i - 1].sc = d[i].sc;
      for (int j = i; j + 1 < n; j++) d[j] = d[j + 1]; n--;
    }
  }
  return n <= 1;
}

int try_gemm(const Nest& n) {
  if (n.n_in != 2 || n.n_out != 1) return NCL_ERR_UNSUPPORTED;
  const Expr& e = n.transform[0];
  if (e.n != 5) return NCL_ERR_UNSUPPORTED;
  const XNode& root = e.node[4];
  if (root.op != NCL_OP_ADD || e.node[root.a].op != NCL_X_OUTPUT || e.node[root.b].op != NCL_OP_MUL) return NCL_ERR_UNSUPPORTED;
  const XNode& mul = e.node[root.b];
  if (e.node[mul.a].op != NCL_X_INPUT || e.node[mul.b].op != NCL_X_INPUT || e.node[mul.a].a == e.node[mul.b].a) return NCL_ERR_UNSUPPORTED;
  const Operand& C = n.out[0];
  // C[m][n] with n contiguous is what the kernels write.  When the result is m-contiguous instead -- `(ij jk -> ki)` -- the
  // operands swap roles: C^T = B^T A^T is the same product with the loops private to B as rows.
  bool swap_roles = false;
  {
    long long best_a = 0, best_b = 0;                                    // smallest result stride among the loops private to A / to B
    for (int l = 0; l < n.nloops; l++) {
      if (n.extent[l] == 1 || C.stride[l] == 0) continue;
      const bool a = n.in[e.node[mul.a].a - 1].stride[l] != 0, b = n.in[e.node[mul.b].a - 1].stride[l] != 0;
      const long long sc = llabs(C.stride[l]);
      if (a && !b && (best_a == 0 || sc < best_a)) best_a = sc;
      if (!a && b && (best_b == 0 || sc < best_b)) best_b = sc;
    }
    swap_roles = best_a == 1 && best_b != 1 && best_b != 0;
  }
  const Operand& A = n.in[(swap_roles ? e.node[mul.b].a : e.node[mul.a].a) - 1]; const Operand& B = n.in[(swap_roles ? e.node[mul.a].a : e.node[mul.b].a) - 1];
  int ca = g_dt[A.dtype].cls, cb = g_dt[B.dtype].cls, cc = g_dt[C.dtype].cls;
  // Mixed classes (single-float x double-float, integer x float): Common Lisp's contagion converts the lower operand to the
  // class of the other BEFORE the product, exactly what converting the ARRAY first does -- so one operand may be of a lower
  // class than the result if the other already has the result's class (both lower: the product would be formed in the lower
  // class first; that stays on the generic kernel).
  int dtA = A.dtype, dtB = B.dtype;
  if (ca != cc || cb != cc) {
    if (!(cc == CLS_F || cc == CLS_D) || ca > cc || cb > cc || (ca != cc && cb != cc) || ca > CLS_D || cb > CLS_D) return NCL_ERR_UNSUPPORTED;
    if (ca != cc) dtA = cc == CLS_F ? NCL_F32 : NCL_F64;
    if (cb != cc) dtB = cc == CLS_F ? NCL_F32 : NCL_F64;
    if (A.dtype == NCL_UB64 || B.dtype == NCL_UB64) return NCL_ERR_UNSUPPORTED;
  }
  if (g_dt[A.dtype].slot_bits < 8 || g_dt[B.dtype].slot_bits < 8 || g_dt[C.dtype].slot_bits < 8) return NCL_ERR_UNSUPPORTED;
  GDim gb[NCL_MAX_LOOPS], gm[NCL_MAX_LOOPS], gn[NCL_MAX_LOOPS], gk[NCL_MAX_LOOPS]; int nb = 0, nm = 0, nn = 0, nk = 0;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] == 1) continue;
    GDim d = {n.extent[l], A.stride[l], B.stride[l], C.stride[l]};
    bool a = d.sa != 0, b = d.sb != 0, c = d.sc != 0;
    if (a && b && c) gb[nb++] = d; else if (a && !b && c) gm[nm++] = d; else if (!a && b && c) gn[nn++] = d; else if (a && b && !c) gk[nk++] = d;
    else return NCL_ERR_UNSUPPORTED;
  }
  if (nk == 0) return NCL_ERR_UNSUPPORTED;                              // no contraction: elementwise / outer product
  if (!merge_group(gb, nb) || !merge_group(gm, nm) || !merge_group(gn, nn) || !merge_group(gk, nk)) return NCL_ERR_UNSUPPORTED;
  GemmParams p; memset(&p, 0, sizeof p);
  p.M = nm ? (int)gm[0].ext : 1; p.N = nn ? (int)gn[0].ext : 1; p.K = (int)gk[0].ext; p.batch = nb ? (int)gb[0].ext : 1;
  if ((nm && gm[0].ext > 0x7fffffffLL) || (nn && gn[0].ext > 0x7fffffffLL) || gk[0].ext > 0x7fffffffLL || (nb && gb[0].ext > 0x7fffffffLL)) return NCL_ERR_UNSUPPORTED;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batch <= 0) return NCL_ERR_UNSUPPORTED;   // empty loops: the generic kernel defines them
  if (nn && gn[0].sc != 1) return NCL_ERR_UNSUPPORTED;                  // C n-contiguous
  p.lda = nm ? gm[0].sa : 0; p.ldb = gk[0].sb; p.ldc = nm ? gm[0].sc : 0;
  p.sA = nb ? gb[0].sa : 0; p.sB = nb ? gb[0].sb : 0; p.sC = nb ? gb[0].sc : 0;
  p.A = A.base + (A.offset * g_dt[A.dtype].slot_bits) / 8; p.B = B.base + (B.offset * g_dt[B.dtype].slot_bits) / 8;
  p.C = C.base + (C.offset * g_dt[C.dtype].slot_bits) / 8;
  // Transposed operands: `(ij ik -> jk)` = A^T B (A is m-contiguous) and `(ij kj -> ik)` = A B^T (B is k-contiguous) are as common
  // as matmul itself (normal equations, covariances, the backward products of a layer) and used to fall to the thread-per-
  // output generic kernel -- 20.6 ms for 8192 x 256 x 256.  The operand is copied into the k- / n-contiguous layout first (one
  // HBM-bound pass through the transpose kernels), then the same GEMM kernels run.
  ncl_buf* tmp[2] = {nullptr, nullptr};
  auto release = [&](int rc) { for (int i = 0; i < 2; i++) if (tmp[i]) ncl_free(tmp[i]); return rc; };
  auto relayout = [&](const Operand& X, int odt, long long s_b, long long s_r, long long s_c, long long rows, long long cols, ncl_buf** out) -> int {
    // X[b][r][c] with element strides (s_b, s_r, s_c) -> contiguous [batch][rows][cols] of element type odt
    const long long total = (long long)p.batch * rows * cols;
    *out = ncl_alloc(ncl_storage_bytes(odt, total), 0);
    if (!*out) return NCL_ERR_NOMEM;
    Nest* t = new Nest;
    t->nloops = 3; t->extent[0] = p.batch; t->extent[1] = rows; t->extent[2] = cols;
    t->n_in = 1; t->n_out = 1; t->fresh = true;
    t->in[0] = X; t->in[0].stride[0] = s_b; t->in[0].stride[1] = s_r; t->in[0].stride[2] = s_c;
    Operand& o = t->out[0]; memset(&o, 0, sizeof o);
    o.base = (char*)ncl_buf_ptr(*out); o.dtype = odt; o.offset = 0; o.n_storage = (int64_t)((ncl_buf_bytes(*out) * 8) / g_dt[odt].slot_bits); o.stage_item = -1;
    o.stride[0] = rows * cols; o.stride[1] = cols; o.stride[2] = 1;
    Expr& x = t->transform[0]; x.n = 1; memset(&x.node[0], 0, sizeof x.node[0]); x.node[0].op = NCL_X_INPUT; x.node[0].a = 1; x.node[0].cls = g_dt[X.dtype].cls;
    int rc = try_transpose(*t);
    if (rc == NCL_ERR_UNSUPPORTED) rc = try_elementwise(*t);
    if (rc == NCL_ERR_UNSUPPORTED) rc = launch_generic(*t);
    delete t; return rc;
  };
  if (gk[0].sa != 1 || dtA != A.dtype) {                               // A must be k-contiguous and of the result's class
    if ((gk[0].sa != 1 && !(nm && gm[0].sa == 1)) || ctx().capturing) return NCL_ERR_UNSUPPORTED;
    int rc = relayout(A, dtA, p.sA, nm ? gm[0].sa : 0, gk[0].sa, p.M, p.K, &tmp[0]);
    if (rc != NCL_OK) return release(rc);
    p.A = (const char*)ncl_buf_ptr(tmp[0]); p.lda = p.K; p.sA = nb ? (long long)p.M * p.K : 0;
  }
  if ((nn && gn[0].sb != 1) || dtB != B.dtype) {                       // B must be n-contiguous and of the result's class
    if ((nn && gn[0].sb != 1 && gk[0].sb != 1) || ctx().capturing) return release(NCL_ERR_UNSUPPORTED);
    int rc = relayout(B, dtB, p.sB, gk[0].sb, nn ? gn[0].sb : 0, p.K, p.N, &tmp[1]);
    if (rc != NCL_OK) return release(rc);
    p.B = (const char*)ncl_buf_ptr(tmp[1]); p.ldb = p.N; p.sB = nb ? (long long)p.K * p.N : 0;
  }
  if (p.lda < 0 || p.ldb < 0 || p.ldc < 0 || p.sA < 0 || p.sB < 0 || p.sC < 0) return release(NCL_ERR_UNSUPPORTED);
  p.accumulate = n.fresh ? 0 : 1;
  p.lkA = dtype_load_kind(dtA); p.lkB = dtype_load_kind(dtB); p.lkC = dtype_load_kind(C.dtype); p.skC = dtype_store_spec(C.dtype);
  int rc = NCL_ERR_UNSUPPORTED;
  long long work = (long long)p.M * p.N;
  if (cc == CLS_D && dtA == NCL_F64 && p.M >= 32 && p.N >= 32) rc = gemm_f64_dmma(p);
  else if (cc == CLS_F && work >= 128 * 128 && p.K >= 32) rc = gemm_f32_tf32x3(p);
  if (rc != NCL_ERR_UNSUPPORTED) return release(rc);
  return release(gemm_simt(p, cc));
}
