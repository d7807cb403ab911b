// ncl_comm.cu -- multi-GPU: combining per-GPU reduction partials over NVLink.
//
// The reference has no collectives.  The path shards naturally along the outer axis (one process
// per GPU, each holding a row slab), so the only exchange step is the all-reduce of reduction
// partials: one float64/int64 element for a full reduce, the kept-axis vector for a reduction over
// the sharded axis.  NCCL is bound lazily with dlopen so that the library has no link-time
// dependency and shares the NCCL instance a host process (e.g. torch) may already have loaded.
#include <dlfcn.h>
#include <string.h>
#include "ncl_device.cuh"

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_errstr)(int);

static struct {
  void* h; fn_get_uid get_uid; fn_init_rank init_rank; fn_allreduce allreduce; fn_destroy destroy; fn_errstr errstr; fn_allgather allgather;
  // peer mailboxes for the fused one-element all-reduce (NVLink P2P, CUDA IPC)
  unsigned long long* mbox; unsigned long long** peers_host; unsigned long long** peers_dev; bool p2p;
  nccl_comm comm; int rank, world; void* partial;   // 8-byte device slot for the sharded full reduce
  cudaStream_t comm_stream; cudaEvent_t ev_partial, ev_done[2]; int outstanding; unsigned long long issued;   // finalizes in flight (<= 2) / issued so far
} g_nccl;

static int nccl_load() {
  if (g_nccl.h) return NCL_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.h; i++) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.h) return set_error(NCL_ERR_STATE, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.h, "ncclGetErrorString");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.h, "ncclAllGather");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy)
    return set_error(NCL_ERR_STATE, "libnccl.so.2 lacks the expected symbols");
  return NCL_OK;
}
static int nccl_fail(int rc, const char* what) {
  if (rc == 0) return NCL_OK;
  return set_error(NCL_ERR_CUDA, "%s: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "NCCL error");
}

// r <- fn(init, partial) with one rounding into r's element type
template <class W>
__global__ void finalize_kernel(const W* partial, char* out, long long off, StoreSpec sk, int out_lk, int fresh, int op, long long ident_i, double ident_f) {
  W s = *partial;
  W init;
  if (fresh) init = std::is_integral<W>::value ? (W)ident_i : (W)ident_f;
  else init = load_scalar<W>(out, off, out_lk);
  W r = op == NCL_OP_ADD ? Ops<W>::add(init, s) : op == NCL_OP_MUL ? Ops<W>::mul(init, s) : op == NCL_OP_MAX ? Ops<W>::mx(init, s) : Ops<W>::mn(init, s);
  store_scalar<W>(out, off, sk, r);
}

static int finalize_partial_on(cudaStream_t stream, int op, const void* partial8, int cls, Operand& r, bool fresh);
int finalize_partial(int op, const void* partial8, int cls, Operand& r, bool fresh) { return finalize_partial_on(ctx().stream, op, partial8, cls, r, fresh); }
static int finalize_partial_on(cudaStream_t stream, int op, const void* partial8, int cls, Operand& r, bool fresh) {
  StoreSpec sk = dtype_store_spec(r.dtype); int lk = dtype_load_kind(r.dtype);
  long long ident_i; double ident_f; reduce_identity(op, r.dtype, &ident_i, &ident_f);
  if (cls == CLS_I) finalize_kernel<long long><<<1, 1, 0, stream>>>((const long long*)partial8, r.base, r.offset, sk, lk, fresh, op, ident_i, ident_f);
  else finalize_kernel<double><<<1, 1, 0, stream>>>((const double*)partial8, r.base, r.offset, sk, lk, fresh, op, ident_i, ident_f);
  note_launch("reduce_finalize");
  return cuda_fail(cudaGetLastError(), "finalize_kernel");
}

// The collective half of the sharded full reduce when the mailboxes are mapped (k_reduce.cuh: reduce_all_kernel pushes):
// one CTA on the communication stream waits for the W messages of the current sequence number, folds them in rank order
// (every rank computes the same bits) and stores the result with one rounding.  It runs beside whatever the compute
// stream launched after the local reduction, so neither the NVLink latency nor the skew between ranks is on that stream.
template <class W>
__global__ void mailbox_finalize_kernel(unsigned long long* const* peers, int rank, int world, char* out, long long off, StoreSpec sk, int op,
                                        long long ident_i, double ident_f) {
  __shared__ unsigned long long xch[64];
  unsigned long long* mb = peers[rank];
  // The finalizes run in order on the communication stream and count for themselves (device-resident, so the launch is
  // identical on every graph replay): with two exchanges in flight the rank's push counter may already be one ahead.
  __shared__ unsigned long long xseq;
  if (threadIdx.x == 0) { xseq = ((volatile unsigned long long*)mb)[(size_t)8 * world + 1] + 1ull; ((volatile unsigned long long*)mb)[(size_t)8 * world + 1] = xseq; }
  __syncthreads();
  const unsigned long long seq = xseq;
  const int t = threadIdx.x;
  if (t < world) {
    unsigned long long bits = 0;
    if (!mbox_recv(mb, world, seq, t, &bits)) { printf("numcl-cuda: rank %d never received partial %llu of rank %d\n", rank, seq, t); __trap(); }
    xch[t] = bits;
  }
  __syncthreads();
  if (t != 0) return;
  auto val = [&](int r) -> W { if constexpr (std::is_integral<W>::value) return (W)(long long)xch[r]; else return (W)__longlong_as_double((long long)xch[r]); };
  W tot = val(0);
  for (int r = 1; r < world; r++) {
    const W v = val(r);
    tot = op == NCL_OP_ADD ? Ops<W>::add(tot, v) : op == NCL_OP_MUL ? Ops<W>::mul(tot, v) : op == NCL_OP_MAX ? Ops<W>::mx(tot, v) : Ops<W>::mn(tot, v);
  }
  const W init = std::is_integral<W>::value ? (W)ident_i : (W)ident_f;
  const W res = op == NCL_OP_ADD ? Ops<W>::add(init, tot) : op == NCL_OP_MUL ? Ops<W>::mul(init, tot) : op == NCL_OP_MAX ? Ops<W>::mx(init, tot) : Ops<W>::mn(init, tot);
  store_scalar<W>(out, off, sk, res);
}

// Map every rank's mailbox into this process (CUDA IPC over NVLink).  The 64-byte IPC handles travel
// through one ncclAllGather; any failure simply leaves p2p off.
static void setup_p2p() {
  g_nccl.p2p = false;
  if (!g_nccl.allgather || g_nccl.world < 2 || g_nccl.world > 63 || getenv("NUMCL_CUDA_NO_P2P")) return;
  const int W = g_nccl.world; cudaStream_t st = ctx().stream;
  size_t mbytes = mbox_words(W) * 8;                      // 4 sequence slots x W sources x 2 words + counters
  if (cudaMalloc((void**)&g_nccl.mbox, mbytes) != cudaSuccess) return;
  cudaMemset(g_nccl.mbox, 0, mbytes);
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, g_nccl.mbox) != cudaSuccess) { cudaGetLastError(); return; }
  char* dev_handles = nullptr;
  if (cudaMalloc((void**)&dev_handles, (size_t)W * sizeof mine) != cudaSuccess) return;
  cudaMemcpy(dev_handles + (size_t)g_nccl.rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice);
  cudaDeviceSynchronize();
  if (g_nccl.allgather(dev_handles + (size_t)g_nccl.rank * sizeof mine, dev_handles, sizeof mine, 0 /* ncclChar */, g_nccl.comm, st) != 0) { cudaFree(dev_handles); return; }
  if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(dev_handles); return; }
  cudaIpcMemHandle_t* all = new cudaIpcMemHandle_t[W];
  cudaMemcpy(all, dev_handles, (size_t)W * sizeof mine, cudaMemcpyDeviceToHost);
  cudaFree(dev_handles);
  g_nccl.peers_host = new unsigned long long*[W];
  bool ok = true;
  for (int r = 0; r < W && ok; r++) {
    if (r == g_nccl.rank) { g_nccl.peers_host[r] = g_nccl.mbox; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; }
    g_nccl.peers_host[r] = (unsigned long long*)p;
  }
  delete[] all;
  if (!ok) return;
  if (cudaMalloc((void**)&g_nccl.peers_dev, (size_t)W * sizeof(void*)) != cudaSuccess) return;
  cudaMemcpy(g_nccl.peers_dev, g_nccl.peers_host, (size_t)W * sizeof(void*), cudaMemcpyHostToDevice);
  // nobody may push before every rank has zeroed and mapped its mailbox: one more (tiny) collective as a barrier
  int* flag = nullptr;
  if (cudaMalloc((void**)&flag, 64) != cudaSuccess) return;
  cudaMemset(flag, 0, 64);
  bool synced = g_nccl.allreduce(flag, flag, 1, 2 /* ncclInt32 */, 0 /* sum */, g_nccl.comm, st) == 0 && cudaStreamSynchronize(st) == cudaSuccess;
  cudaFree(flag);
  if (!synced) return;
  g_nccl.p2p = true;
}

extern "C" {

int ncl_comm_unique_id(void* id128) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  NCL_TRY(nccl_load());
  nccl_uid u; NCL_TRY(nccl_fail(g_nccl.get_uid(&u), "ncclGetUniqueId"));
  memcpy(id128, &u, sizeof u); return NCL_OK;
}
int ncl_comm_init(const void* id128, int rank, int world) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  if (!ctx().inited) return set_error(NCL_ERR_STATE, "ncl_init has not been called");
  NCL_TRY(nccl_load());
  if (g_nccl.comm) return NCL_OK;
  nccl_uid u; memcpy(&u, id128, sizeof u);
  NCL_TRY(nccl_fail(g_nccl.init_rank(&g_nccl.comm, world, u, rank), "ncclCommInitRank"));
  g_nccl.rank = rank; g_nccl.world = world;
  NCL_CUDA(cudaMalloc(&g_nccl.partial, 64));
  // highest priority: the collective's single CTA must be scheduled ahead of the thousands of pending
  // blocks of whatever compute kernel it is meant to overlap
  int prio_lo = 0, prio_hi = 0; NCL_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  NCL_CUDA(cudaStreamCreateWithPriority(&g_nccl.comm_stream, cudaStreamNonBlocking, prio_hi));
  NCL_CUDA(cudaEventCreateWithFlags(&g_nccl.ev_partial, cudaEventDisableTiming));
  NCL_CUDA(cudaEventCreateWithFlags(&g_nccl.ev_done[0], cudaEventDisableTiming));
  NCL_CUDA(cudaEventCreateWithFlags(&g_nccl.ev_done[1], cudaEventDisableTiming));
  g_nccl.outstanding = 0; g_nccl.issued = 0;
  setup_p2p();                                          // best effort: without it the NCCL path is used
  return NCL_OK;
}
int ncl_comm_destroy(void) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  if (g_nccl.comm) {
    cudaStreamSynchronize(ctx().stream); cudaStreamSynchronize(g_nccl.comm_stream);
    g_nccl.destroy(g_nccl.comm); g_nccl.comm = nullptr;
    cudaStreamDestroy(g_nccl.comm_stream); cudaEventDestroy(g_nccl.ev_partial); cudaEventDestroy(g_nccl.ev_done[0]); cudaEventDestroy(g_nccl.ev_done[1]); g_nccl.outstanding = 0;
  }
  if (g_nccl.partial) { cudaFree(g_nccl.partial); g_nccl.partial = nullptr; }
  return NCL_OK;
}

static int nccl_dtype(int dt, int op, int* out) {
  switch (dt) {
    case NCL_F32: *out = 7; return NCL_OK;
    case NCL_F64: *out = 8; return NCL_OK;
    case NCL_SB64: case NCL_UB63: *out = 4; return NCL_OK;
    case NCL_UB64: *out = 5; return NCL_OK;
    case NCL_FIXNUM: case NCL_UB62:      // tagged words: (a<<1)+(b<<1) = (a+b)<<1, order preserved for max/min
      if (op == NCL_OP_MUL) return set_error(NCL_ERR_UNSUPPORTED, "all-reduce product of tagged fixnums");
      *out = 4; return NCL_OK;
    case NCL_SB32: *out = 2; return NCL_OK;
    case NCL_UB31: case NCL_UB32: *out = 3; return NCL_OK;
    case NCL_SB8: *out = 0; return NCL_OK;
    case NCL_UB7: case NCL_UB8: *out = 1; return NCL_OK;
    default: return set_error(NCL_ERR_UNSUPPORTED, "all-reduce of element type %s", g_dt[dt].name);
  }
}
static int nccl_op(int op, int* out) {
  switch (op) { case NCL_OP_ADD: *out = 0; return NCL_OK; case NCL_OP_MUL: *out = 1; return NCL_OK;
    case NCL_OP_MAX: *out = 2; return NCL_OK; case NCL_OP_MIN: *out = 3; return NCL_OK;
    default: return set_error(NCL_ERR_UNSUPPORTED, "all-reduce op %d", op); }
}

int ncl_allreduce(int op, ncl_tensor* t) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  if (!g_nccl.comm) return set_error(NCL_ERR_STATE, "ncl_comm_init has not been called");
  if (!t || !t->buf) return set_error(NCL_ERR_INVALID, "ncl_allreduce needs a device-resident tensor");
  int ndt, nop; NCL_TRY(nccl_dtype(t->dtype, op, &ndt)); NCL_TRY(nccl_op(op, &nop));
  int64_t n = 1; for (int i = 0; i < t->rank; i++) n *= t->dims[i];
  char* p = (char*)t->buf->ptr + (t->elem_offset * g_dt[t->dtype].slot_bits) / 8;
  return nccl_fail(g_nccl.allreduce(p, p, (size_t)n, ndt, nop, g_nccl.comm, ctx().stream), "ncclAllReduce");
}

// The collective and the final rounding run on the library's communication stream so that the
// ~30 us latency of a one-element all-reduce hides behind whatever the caller launches next on the
// compute stream (e.g. the axis-1 reduction of the same slab).  ncl_comm_join() orders the compute
// stream after it.
int ncl_reduce_all_sharded_async(int op, const ncl_tensor* a, ncl_tensor* r) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  if (!ctx().inited) return set_error(NCL_ERR_STATE, "ncl_init has not been called");
  // the slab may be a host array with a resident device mirror (ncl_host_register): the mirror is what gets reduced
  void* amirror = (a && !a->buf && a->host) ? resident_mirror(a->host, nullptr) : nullptr;
  if (!a || !r || (!a->buf && !amirror) || !r->buf) return set_error(NCL_ERR_INVALID, "ncl_reduce_all_sharded needs a device-resident (or registered host) slab and a device-resident result");
  if (r->rank != 0) return set_error(NCL_ERR_SHAPE, "result of a full reduce is a rank-0 array");
  if (g_dt[a->dtype].cls != g_dt[r->dtype].cls) return set_error(NCL_ERR_UNSUPPORTED, "element classes differ");
  int nop; NCL_TRY(nccl_op(op, &nop));
  Operand ao; memset(&ao, 0, sizeof ao); ao.base = a->buf ? (char*)a->buf->ptr : (char*)amirror; ao.dtype = a->dtype; ao.offset = a->elem_offset;
  int64_t n = 1; for (int i = 0; i < a->rank; i++) n *= a->dims[i];
  int cls = g_dt[a->dtype].cls;
  Operand ro; memset(&ro, 0, sizeof ro); ro.base = (char*)r->buf->ptr; ro.dtype = r->dtype; ro.offset = r->elem_offset;
  if (!(g_nccl.comm && g_nccl.world > 1)) {                              // single process: no exchange
    void* slot = nullptr; NCL_CUDA(cudaMalloc(&slot, 64));
    int rc = reduce_all_to_partial(op, ao, n, slot);
    if (rc == NCL_OK) rc = finalize_partial(op, slot, cls, ro, true);
    cudaStreamSynchronize(ctx().stream); cudaFree(slot);
    return rc;
  }
  if (g_nccl.p2p) {
    // The local reduction's last block pushes the partial into every rank's mailbox over NVLink (P2P stores, no NCCL
    // call); the one-CTA finalize collects the W messages on the communication stream.  The mailbox has FOUR sequence
    // slots and a rank's push n is ordered after its own finalize n - 2 (see ncl_device.cuh): two exchanges may be in
    // flight, so the compute stream never waits for the exchange it has just started -- only for the one before it.
    if (g_nccl.outstanding >= 2) { NCL_CUDA(cudaStreamWaitEvent(ctx().stream, g_nccl.ev_done[g_nccl.issued & 1], 0)); g_nccl.outstanding = 1; }
    int rc = reduce_all_p2p(op, ao, n, ro, g_nccl.peers_dev, g_nccl.rank, g_nccl.world, 0);
    if (rc != NCL_ERR_UNSUPPORTED) {
      NCL_TRY(rc);
      NCL_CUDA(cudaEventRecord(g_nccl.ev_partial, ctx().stream));
      NCL_CUDA(cudaStreamWaitEvent(g_nccl.comm_stream, g_nccl.ev_partial, 0));
      long long ident_i; double ident_f; reduce_identity(op, ro.dtype, &ident_i, &ident_f);
      const StoreSpec sk = dtype_store_spec(ro.dtype);
      if (cls == CLS_I) mailbox_finalize_kernel<long long><<<1, 64, 0, g_nccl.comm_stream>>>(g_nccl.peers_dev, g_nccl.rank, g_nccl.world, ro.base, ro.offset, sk, op, ident_i, ident_f);
      else mailbox_finalize_kernel<double><<<1, 64, 0, g_nccl.comm_stream>>>(g_nccl.peers_dev, g_nccl.rank, g_nccl.world, ro.base, ro.offset, sk, op, ident_i, ident_f);
      note_side_launch("reduce_all_p2p");
      NCL_TRY(cuda_fail(cudaGetLastError(), "mailbox_finalize_kernel"));
      NCL_CUDA(cudaEventRecord(g_nccl.ev_done[g_nccl.issued & 1], g_nccl.comm_stream));
      g_nccl.issued++; g_nccl.outstanding++;
      return NCL_OK;
    }
  }
  NCL_TRY(ncl_comm_join());                                                        // the NCCL path's partial slot is single-buffered
  NCL_TRY(reduce_all_to_partial(op, ao, n, g_nccl.partial));
  NCL_CUDA(cudaEventRecord(g_nccl.ev_partial, ctx().stream));
  NCL_CUDA(cudaStreamWaitEvent(g_nccl.comm_stream, g_nccl.ev_partial, 0));
  NCL_TRY(nccl_fail(g_nccl.allreduce(g_nccl.partial, g_nccl.partial, 1, cls == CLS_I ? 4 : 8, nop, g_nccl.comm, g_nccl.comm_stream), "ncclAllReduce"));
  NCL_TRY(finalize_partial_on(g_nccl.comm_stream, op, g_nccl.partial, cls, ro, true));
  NCL_CUDA(cudaEventRecord(g_nccl.ev_done[g_nccl.issued & 1], g_nccl.comm_stream));
  g_nccl.issued++; g_nccl.outstanding++;
  return NCL_OK;
}
int ncl_comm_join(void) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  // order the compute stream after every finalize still in flight (the last one, and the one before it)
  if (g_nccl.outstanding >= 1) NCL_CUDA(cudaStreamWaitEvent(ctx().stream, g_nccl.ev_done[(g_nccl.issued - 1) & 1], 0));
  if (g_nccl.outstanding >= 2) NCL_CUDA(cudaStreamWaitEvent(ctx().stream, g_nccl.ev_done[g_nccl.issued & 1], 0));
  g_nccl.outstanding = 0;
  return NCL_OK;
}
int ncl_reduce_all_sharded(int op, const ncl_tensor* a, ncl_tensor* r) {
  std::lock_guard<std::recursive_mutex> lock__(ncl_mutex());
  NCL_TRY(ncl_reduce_all_sharded_async(op, a, r));
  return ncl_comm_join();
}

}  // extern "C"
