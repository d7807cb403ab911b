// ncl_api.cu -- the C ABI (include/numcl_cuda.h): context, buffers, and the lowering of every
// entry point to the Nest the reference's loop generators would have emitted.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>
#include "ncl_internal.h"

static Context g_ctx;
static std::recursive_mutex g_mu;
static thread_local char g_err[1024] = "";
Context& ctx() { return g_ctx; }
std::recursive_mutex& ncl_mutex() { return g_mu; }

int set_error(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NCL_OK;
  return set_error(NCL_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
// ---- residency: a device mirror attached to a host base vector (what %make-array gains, src/1instantiate.lisp:81-96) ---------
// A numcl array is an ordinary Lisp array, so the host vector stays the array; ncl_host_register gives it a device mirror that
// SURVIVES the call.  A host-resident tensor whose `host` pointer is a registered base vector is not staged through the arena:
// it is uploaded only when the host copy is newer (first use, or after ncl_host_dirty), so the second and every later call on
// the same array moves nothing over PCIe.  Results are written back into the host vector at the end of the call (window
// only), or -- NCL_RESIDENT_LAZY -- left on the device until ncl_host_fetch.
enum { RES_IN_SYNC = 0, RES_HOST_NEWER = 1, RES_DEVICE_NEWER = 2 };
struct Resident { void* host; size_t bytes; ncl_buf* dev; int state; unsigned flags; };
static std::map<const void*, Resident> g_res;
static Resident* find_resident(const void* host) { auto it = g_res.find(host); return it == g_res.end() ? nullptr : &it->second; }

void* resident_mirror(const void* host, size_t* bytes) {
  Resident* r = find_resident(host);
  if (!r) return nullptr;
  if (r->state == RES_HOST_NEWER) {
    if (cudaMemcpyAsync(r->dev->ptr, r->host, r->bytes, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess) return nullptr;
    ctx().h2d_bytes += r->bytes; r->state = RES_IN_SYNC;
  }
  if (bytes) *bytes = r->bytes;
  return r->dev->ptr;
}
// ---- device buffer cache (see "buffers" below) ----
static std::multimap<size_t, void*> g_pool;
static size_t g_pool_bytes = 0;
static const size_t POOL_MAX_BYTES = (size_t)48 << 30;
static size_t pool_round(size_t bytes) {
  if (bytes <= ((size_t)1 << 20)) return bytes ? ((bytes + 511) & ~(size_t)511) : 512;
  return (bytes + (((size_t)2 << 20) - 1)) & ~(((size_t)2 << 20) - 1);      // 2 MiB granules: few distinct sizes, little slack
}
static bool pool_enabled() { static int on = -1; if (on < 0) { const char* e = getenv("NUMCL_NO_POOL"); on = (e && atoi(e)) ? 0 : 1; } return on == 1; }
static void pool_release_all() {
  if (g_pool.empty()) return;
  cudaStreamSynchronize(g_ctx.stream);
  for (auto& kv : g_pool) cudaFree(kv.second);
  g_pool.clear(); g_pool_bytes = 0;
}

// pinned host blocks (ncl_host_alloc_pinned below)
static std::multimap<size_t, void*> g_hpool;                 // free blocks by size
static std::map<const void*, size_t> g_hsize;                // every live or cached block -> its size
static size_t g_hpool_bytes = 0;
static size_t hpool_cap() { static size_t v = (size_t)-1; if (v == (size_t)-1) { const char* e = getenv("NUMCL_PINNED_POOL_MB"); v = (size_t)(e ? atol(e) : 4096) << 20; } return v; }
static void hpool_release_all() {
  for (auto& kv : g_hpool) { g_hsize.erase(kv.second); cudaFreeHost(kv.second); }
  g_hpool.clear(); g_hpool_bytes = 0;
}

static int grow(void** p, size_t* have, size_t want) {
  if (*have >= want) return NCL_OK;
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "a scratch arena must grow (%zu bytes) inside ncl_capture_begin/end: run the sequence once before capturing it", want);
  if (*p) { NCL_CUDA(cudaStreamSynchronize(g_ctx.stream)); NCL_CUDA(cudaFree(*p)); *p = nullptr; *have = 0; }
  size_t sz = want + want / 4 + (1u << 20);
  if (cudaMalloc(p, sz) != cudaSuccess) { cudaGetLastError(); pool_release_all(); NCL_CUDA(cudaMalloc(p, sz)); }
  *have = sz; return NCL_OK;
}
int ensure_scratch(size_t bytes) { return grow(&g_ctx.scratch, &g_ctx.scratch_bytes, bytes); }
int ensure_scratch2(size_t bytes) { return grow(&g_ctx.scratch2, &g_ctx.scratch2_bytes, bytes); }
int ensure_scratch3(size_t bytes) { return grow(&g_ctx.scratch3, &g_ctx.scratch3_bytes, bytes); }

void reduce_identity(int op, int dtype, long long* ident_i, double* ident_f) {
  const DtInfo& d = g_dt[dtype];
  *ident_i = op == NCL_OP_MUL ? 1 : 0; *ident_f = op == NCL_OP_MUL ? 1.0 : 0.0;
  if ((op != NCL_OP_MAX && op != NCL_OP_MIN) || cls_complex(d.cls)) return;      // complex numbers are not ordered: the typer refuses max / min
  const bool mx = op == NCL_OP_MAX;
  if (d.cls == CLS_F) *ident_f = mx ? -3.4028234663852886e38 : 3.4028234663852886e38;
  else if (d.cls == CLS_D) *ident_f = mx ? -1.7976931348623157e308 : 1.7976931348623157e308;
  else {
    long long lo, hi;
    if (d.is_signed) { lo = d.value_bits >= 64 ? INT64_MIN : -((long long)1 << (d.value_bits - 1)); hi = d.value_bits >= 64 ? INT64_MAX : ((long long)1 << (d.value_bits - 1)) - 1; }
    else { lo = 0; hi = d.value_bits >= 63 ? INT64_MAX : ((long long)1 << d.value_bits) - 1; }
    *ident_i = mx ? lo : hi;
  }
}

// The sticky DOMAIN flag (ncl_fn.cuh): read back only when a kernel that can raise it has been launched since the last check.
int check_fault() {
  if (!g_ctx.fault_possible || g_ctx.capturing) return NCL_OK;
  unsigned int h = 0;
  NCL_CUDA(cudaMemcpyAsync(&h, g_ctx.fault, sizeof h, cudaMemcpyDeviceToHost, g_ctx.stream));
  NCL_CUDA(cudaStreamSynchronize(g_ctx.stream));
  g_ctx.fault_possible = false;
  if (!h) return NCL_OK;
  NCL_CUDA(cudaMemsetAsync(g_ctx.fault, 0, sizeof h, g_ctx.stream));
  return set_error(NCL_ERR_DOMAIN, "an earlier call left the real domain of its function:%s%s%s (the reference signals a Lisp error or returns a complex number there)",
                   (h & 1) ? " argument outside the function's real domain (sqrt / log of a negative, asin / acos / atanh beyond 1, acosh below 1, isqrt of a negative, negative base to a fractional power);" : "",
                   (h & 2) ? " division by zero (floor / ceiling / round / truncate / mod / rem / expt);" : "",
                   (h & 4) ? " integer power beyond 64 bits;" : "");
}

#define LOCK std::lock_guard<std::recursive_mutex> lock__(g_mu)
#define NEED_INIT if (!g_ctx.inited) return set_error(NCL_ERR_STATE, "ncl_init has not been called")

extern "C" {

const char* ncl_last_error(void) { return g_err; }
const char* ncl_version(void) { return "numcl-cuda 0.1 (sm_100a)"; }

int ncl_init(int ndev, const int* devs) {
  LOCK;
  if (g_ctx.inited) return NCL_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(NCL_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  if (ndev <= 0) ndev = 1;
  // one process per GPU: the buffer cache, the scratch arenas and the stream all belong to ONE device
  if (ndev > 1) return set_error(NCL_ERR_UNSUPPORTED, "ndev = %d: bind one process to one GPU (one rank per device) and shard across processes", ndev);
  g_ctx.ndev = ndev;
  for (int i = 0; i < ndev; i++) g_ctx.devs[i] = devs ? devs[i] : i;
  NCL_CUDA(cudaSetDevice(g_ctx.devs[0]));
  cudaDeviceProp prop; NCL_CUDA(cudaGetDeviceProperties(&prop, g_ctx.devs[0]));
  if (prop.major != 10) return set_error(NCL_ERR_STATE, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  g_ctx.sm_count = prop.multiProcessorCount;
  NCL_CUDA(cudaStreamCreateWithFlags(&g_ctx.own_stream, cudaStreamNonBlocking));
  g_ctx.stream = g_ctx.own_stream;
  NCL_CUDA(cudaMalloc((void**)&g_ctx.ticket, 64 * sizeof(unsigned int)));
  NCL_CUDA(cudaMemset(g_ctx.ticket, 0, 64 * sizeof(unsigned int)));
  long long consts[8] = {0, 1, 0, 0, 0, 0, 0, 0};
  NCL_CUDA(cudaMalloc(&g_ctx.consts, sizeof consts));
  NCL_CUDA(cudaMemcpy(g_ctx.consts, consts, sizeof consts, cudaMemcpyHostToDevice));
  NCL_CUDA(cudaMalloc((void**)&g_ctx.fault, 64));
  NCL_CUDA(cudaMemset(g_ctx.fault, 0, 64));
  g_ctx.fault_possible = false;
  g_ctx.launches = 0; g_ctx.last_kernel = ""; g_ctx.inited = true;
  return NCL_OK;
}

void ncl_shutdown(void) {
  LOCK;
  if (!g_ctx.inited) return;
  cudaStreamSynchronize(g_ctx.stream);
  for (auto& kv : g_res) if (kv.second.dev) { cudaFree(kv.second.dev->ptr); delete kv.second.dev; }
  g_res.clear();
  pool_release_all(); hpool_release_all();
  if (g_ctx.scratch) cudaFree(g_ctx.scratch);
  if (g_ctx.scratch2) cudaFree(g_ctx.scratch2);
  if (g_ctx.scratch3) cudaFree(g_ctx.scratch3);
  if (g_ctx.ticket) cudaFree(g_ctx.ticket);
  if (g_ctx.consts) cudaFree(g_ctx.consts);
  if (g_ctx.fault) cudaFree(g_ctx.fault);
  if (g_ctx.pinned) cudaFreeHost(g_ctx.pinned);
  if (g_ctx.own_stream) cudaStreamDestroy(g_ctx.own_stream);
  if (g_ctx.copy_in) cudaStreamDestroy(g_ctx.copy_in);
  if (g_ctx.copy_out) cudaStreamDestroy(g_ctx.copy_out);
  for (int k = 0; k < g_ctx.n_events; k++) cudaEventDestroy(g_ctx.events[k]);
  free(g_ctx.events);
  g_ctx = Context();
}

int ncl_device_count(void) { int c = 0; if (cudaGetDeviceCount(&c) != cudaSuccess) return 0; return c; }
int ncl_set_stream(void* s) {
  LOCK; NEED_INIT;
  cudaStream_t ns = s ? (cudaStream_t)s : g_ctx.own_stream;
  if (ns != g_ctx.stream) {
    // blocks recycled by the buffer cache (and the scratch arenas) were last used on the old stream: order the new one after it
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
      cudaEventRecord(ev, g_ctx.stream); cudaStreamWaitEvent(ns, ev, 0); cudaEventDestroy(ev);
    }
    g_ctx.stream = ns;
  }
  pdl().unknown = true;
  return NCL_OK;
}
void* ncl_get_stream(void) { return (void*)g_ctx.stream; }
int ncl_sync(void) {
  LOCK; NEED_INIT;
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_sync inside ncl_capture_begin/end");
  ncl_comm_join(); NCL_CUDA(cudaStreamSynchronize(g_ctx.stream)); return check_fault();
}
// Give every cached device block back to the driver (the cache keeps freed base vectors for re-use and would otherwise
// only shrink when one of the library's own allocations fails): for hosts that share the GPU with another allocator.
int ncl_trim(void) { LOCK; NEED_INIT; if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_trim inside a capture"); pool_release_all(); hpool_release_all(); return NCL_OK; }
void ncl_pipeline_stats(int64_t* calls, int64_t* panels) { if (calls) *calls = g_ctx.pipelined_calls; if (panels) *panels = g_ctx.pipelined_panels; }
void ncl_transfer_stats(uint64_t* h2d_bytes, uint64_t* d2h_bytes) { if (h2d_bytes) *h2d_bytes = g_ctx.h2d_bytes; if (d2h_bytes) *d2h_bytes = g_ctx.d2h_bytes; }

// ---- CUDA graphs: a launch-bound sequence of calls recorded once, replayed with ONE launch ---------------------------------
// Between ncl_capture_begin and ncl_capture_end every compute call on device-resident tensors is recorded into a CUDA
// graph instead of being executed (the library's communication stream joins the capture through its events, so the sharded
// reduce with its mailbox exchange is captured too).  Host-resident tensors, ncl_d2h, ncl_sync and anything that has to
// allocate are refused while capturing: run the sequence once eagerly first.  ncl_graph_launch replays on the library
// stream; the arrays named at capture time are the ones read and written on every replay.
struct ncl_graph { cudaGraph_t g; cudaGraphExec_t e; int64_t launches; };
int ncl_capture_begin(void) {
  LOCK; NEED_INIT;
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_capture_begin: already capturing");
  NCL_TRY(ncl_comm_join());
  NCL_CUDA(cudaStreamBeginCapture(g_ctx.stream, cudaStreamCaptureModeRelaxed));
  g_ctx.capturing = true; g_ctx.capture_launch0 = g_ctx.launches;
  pdl().unknown = true;                                    // the first recorded kernel is fully serialised
  return NCL_OK;
}
int ncl_capture_end(ncl_graph** out) {
  LOCK; NEED_INIT;
  if (!g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_capture_end without ncl_capture_begin");
  int jrc = ncl_comm_join();                               // the communication stream must re-join before the capture ends
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(g_ctx.stream, &g);
  g_ctx.capturing = false; pdl().unknown = true;
  if (jrc != NCL_OK) { if (g) cudaGraphDestroy(g); return jrc; }
  if (e != cudaSuccess || !g) { cudaGetLastError(); return set_error(NCL_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e)); }
  cudaGraphExec_t x = nullptr;
  e = cudaGraphInstantiate(&x, g, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(g); cudaGetLastError(); return set_error(NCL_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
  if (!out) { cudaGraphExecDestroy(x); cudaGraphDestroy(g); return set_error(NCL_ERR_INVALID, "ncl_capture_end: out is NULL"); }
  *out = new ncl_graph{g, x, g_ctx.launches - g_ctx.capture_launch0};
  g_ctx.launches = g_ctx.capture_launch0;                  // nothing ran yet: replays are what gets counted
  return NCL_OK;
}
int ncl_graph_launch(ncl_graph* gr) {
  LOCK; NEED_INIT;
  if (!gr) return set_error(NCL_ERR_INVALID, "ncl_graph_launch: NULL graph");
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_graph_launch inside a capture");
  NCL_TRY(ncl_comm_join());
  NCL_CUDA(cudaGraphLaunch(gr->e, g_ctx.stream));
  pdl().unknown = true;
  g_ctx.launches += gr->launches;
  return NCL_OK;
}
int64_t ncl_graph_kernels(const ncl_graph* gr) { return gr ? gr->launches : 0; }
int ncl_graph_free(ncl_graph* gr) {
  LOCK;
  if (!gr) return NCL_OK;
  if (g_ctx.inited) cudaStreamSynchronize(g_ctx.stream);
  cudaGraphExecDestroy(gr->e); cudaGraphDestroy(gr->g); delete gr; return NCL_OK;
}
int64_t ncl_launch_count(void) { return g_ctx.launches; }
const char* ncl_last_kernel(void) { return g_ctx.last_kernel; }
int ncl_force_generic(int on) { g_ctx.force_generic = on != 0; return NCL_OK; }

// ---- buffers ----------------------------------------------------------------------------------
// Every numcl operation returns a fresh array (src/1instantiate.lisp:81-96), so allocation is on the path of every
// call.  cudaMalloc / cudaFree cost 0.1..1 ms and cudaFree synchronises the device; freed base vectors are therefore
// kept in a size-bucketed cache and handed out again.  Re-use is stream-ordered: all work of the library is issued
// on one stream, so a kernel writing a recycled block is ordered after every kernel that still reads it
// (ncl_set_stream chains the old and the new stream with an event).  NUMCL_NO_POOL=1 disables the cache.
ncl_buf* ncl_alloc(size_t bytes, int dev_slot) {
  LOCK;
  if (!g_ctx.inited) { set_error(NCL_ERR_STATE, "ncl_init has not been called"); return nullptr; }
  if (dev_slot < 0 || dev_slot >= g_ctx.ndev) { set_error(NCL_ERR_INVALID, "bad device slot %d", dev_slot); return nullptr; }
  void* p = nullptr;
  const size_t sz = pool_round(bytes);
  auto it = g_pool.find(sz);
  if (it != g_pool.end()) { p = it->second; g_pool.erase(it); g_pool_bytes -= sz; return new ncl_buf{p, sz, dev_slot, true}; }
  cudaError_t e = cudaMalloc(&p, sz);
  if (e != cudaSuccess && !g_pool.empty()) { cudaGetLastError(); pool_release_all(); e = cudaMalloc(&p, sz); }   // give the cache back first
  if (e != cudaSuccess) { cudaGetLastError(); set_error(NCL_ERR_NOMEM, "cudaMalloc(%zu): %s", sz, cudaGetErrorString(e)); return nullptr; }
  return new ncl_buf{p, sz, dev_slot, true};
}
int ncl_free(ncl_buf* b) {
  LOCK;
  if (!b) return NCL_OK;
  if (b->owned && b->ptr) {
    if (g_ctx.inited && pool_enabled() && g_pool_bytes + b->bytes <= POOL_MAX_BYTES && b->bytes == pool_round(b->bytes)) {
      g_pool.emplace(b->bytes, b->ptr); g_pool_bytes += b->bytes;
    } else { cudaStreamSynchronize(g_ctx.stream); cudaFree(b->ptr); }
  }
  delete b; return NCL_OK;
}
void* ncl_buf_ptr(const ncl_buf* b) { return b ? b->ptr : nullptr; }
size_t ncl_buf_bytes(const ncl_buf* b) { return b ? b->bytes : 0; }
ncl_buf* ncl_wrap(void* p, size_t bytes, int dev_slot) { return new ncl_buf{p, bytes, dev_slot, false}; }
int ncl_h2d(ncl_buf* dst, size_t off, const void* host, size_t bytes) {
  LOCK; NEED_INIT;
  if (!dst || off + bytes > dst->bytes) return set_error(NCL_ERR_INVALID, "ncl_h2d out of range");
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_h2d inside ncl_capture_begin/end");
  NCL_CUDA(cudaMemcpyAsync((char*)dst->ptr + off, host, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  g_ctx.h2d_bytes += bytes; pdl().unknown = true;
  return NCL_OK;
}
int ncl_d2h(void* host, const ncl_buf* src, size_t off, size_t bytes) {
  LOCK; NEED_INIT;
  if (!src || off + bytes > src->bytes) return set_error(NCL_ERR_INVALID, "ncl_d2h out of range");
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_d2h inside ncl_capture_begin/end");
  g_ctx.d2h_bytes += bytes;
  ncl_comm_join();
  NCL_CUDA(cudaMemcpyAsync(host, (const char*)src->ptr + off, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
  NCL_CUDA(cudaStreamSynchronize(g_ctx.stream));
  return check_fault();
}
// Pinned host vectors are what %make-array hands out on the :cuda backend, once per RESULT of every numcl call -- and
// cudaMallocHost / cudaFreeHost cost about 0.4 ms per MiB (a 64 MiB result: 25 ms, ten times its PCIe time).  Freed blocks
// are therefore kept in a size-bucketed cache like the device blocks (ncl_alloc); NUMCL_PINNED_POOL_MB caps what it may
// hold (default 4096), ncl_trim gives it back.
void* ncl_host_alloc_pinned(size_t bytes) {
  LOCK;
  const size_t sz = pool_round(bytes);
  auto it = g_hpool.find(sz);
  if (it != g_hpool.end()) { void* p = it->second; g_hpool.erase(it); g_hpool_bytes -= sz; return p; }
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, sz);
  if (e != cudaSuccess && !g_hpool.empty()) { cudaGetLastError(); hpool_release_all(); e = cudaMallocHost(&p, sz); }
  if (e != cudaSuccess) { cudaGetLastError(); set_error(NCL_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", sz, cudaGetErrorString(e)); return nullptr; }
  g_hsize[p] = sz;
  return p;
}
int ncl_host_free(void* p) {
  LOCK;
  if (!p) return NCL_OK;
  auto it = g_hsize.find(p);
  if (it == g_hsize.end()) { cudaFreeHost(p); return NCL_OK; }
  const size_t sz = it->second;
  // a copy into or out of the block may still be in flight on the library's streams only if the caller freed it before the
  // call returned, which the ABI forbids; the block can be handed out again at once
  if (g_ctx.inited && pool_enabled() && g_hpool_bytes + sz <= hpool_cap()) { g_hpool.emplace(sz, p); g_hpool_bytes += sz; return NCL_OK; }
  g_hsize.erase(it); cudaFreeHost(p); return NCL_OK;
}
int ncl_host_register(void* base, size_t bytes, unsigned flags) {
  LOCK; NEED_INIT;
  if (!base || !bytes) return set_error(NCL_ERR_INVALID, "ncl_host_register: NULL base or zero size");
  if (find_resident(base)) return set_error(NCL_ERR_STATE, "ncl_host_register: this base vector is registered already");
  ncl_buf* d = ncl_alloc(bytes, 0);
  if (!d) return NCL_ERR_NOMEM;
  // NCL_RESIDENT_FRESH: the vector was just allocated and holds nothing yet (a result array): no first upload
  g_res[base] = Resident{base, bytes, d, (flags & NCL_RESIDENT_FRESH) ? RES_IN_SYNC : RES_HOST_NEWER, flags};
  return NCL_OK;
}
int ncl_host_unregister(void* base) {
  LOCK;
  auto it = g_res.find(base);
  if (it == g_res.end()) return NCL_OK;
  int rc = NCL_OK;
  if (it->second.state == RES_DEVICE_NEWER && g_ctx.inited) rc = ncl_host_fetch(base);         // never lose a result
  ncl_free(it->second.dev); g_res.erase(base);
  return rc;
}
int ncl_host_dirty(void* base) {
  LOCK;
  Resident* r = find_resident(base);
  if (!r) return set_error(NCL_ERR_INVALID, "ncl_host_dirty: not a registered base vector");
  r->state = RES_HOST_NEWER; return NCL_OK;
}
int ncl_host_fetch(void* base) {
  LOCK; NEED_INIT;
  Resident* r = find_resident(base);
  if (!r) return set_error(NCL_ERR_INVALID, "ncl_host_fetch: not a registered base vector");
  if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "ncl_host_fetch inside ncl_capture_begin/end");
  if (r->state != RES_DEVICE_NEWER) return NCL_OK;
  ncl_comm_join();
  NCL_CUDA(cudaMemcpyAsync(r->host, r->dev->ptr, r->bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
  NCL_CUDA(cudaStreamSynchronize(g_ctx.stream));
  g_ctx.d2h_bytes += r->bytes; r->state = RES_IN_SYNC;
  return check_fault();
}
int ncl_host_state(const void* base) { Resident* r = find_resident(base); return r ? r->state : -1; }
size_t ncl_storage_bytes(int dtype, int64_t n) {
  if (dtype < 0 || dtype >= NCL_DTYPE_COUNT) return 0;
  int64_t padded = 8 * ((n + 7) / 8); if (padded == 0) padded = 8;
  return (size_t)((padded * g_dt[dtype].slot_bits + 7) / 8);
}

}  // extern "C"

// ---- tensors -> operands, with staging of host-resident tensors ---------------------------------
static int64_t total_elems(const ncl_tensor* t) { int64_t n = 1; for (int i = 0; i < t->rank; i++) n *= t->dims[i]; return n; }

static int check_tensor(const ncl_tensor* t, const char* what) {
  if (!t) return set_error(NCL_ERR_INVALID, "%s is NULL", what);
  if (t->dtype < 0 || t->dtype >= NCL_DTYPE_COUNT) return set_error(NCL_ERR_INVALID, "%s: bad dtype %d", what, t->dtype);
  if (t->rank < 0 || t->rank > NCL_MAX_RANK) return set_error(NCL_ERR_INVALID, "%s: bad rank %d", what, t->rank);
  for (int i = 0; i < t->rank; i++) if (t->dims[i] < 0) return set_error(NCL_ERR_INVALID, "%s: negative dimension", what);
  if ((t->buf == nullptr) == (t->host == nullptr)) return set_error(NCL_ERR_INVALID, "%s: exactly one of buf / host must be set", what);
  if (t->elem_offset < 0) return set_error(NCL_ERR_INVALID, "%s: negative offset", what);
  return NCL_OK;
}

// Host-resident tensors are staged through a device arena that lives for the call (or use their resident mirror).
struct Stager {
  struct Item { const ncl_tensor* t; char* dev; size_t bytes; bool is_out; Resident* res; };
  std::vector<Item> items; size_t used = 0;
  static size_t need(const ncl_tensor* t) { return (ncl_storage_bytes(t->dtype, t->elem_offset + total_elems(t)) + 255) & ~(size_t)255; }
  int plan(const ncl_tensor* ts, int n, bool is_out) {
    for (int i = 0; i < n; i++) if (ts[i].host) {
      Resident* r = find_resident(ts[i].host);
      if (r && need(&ts[i]) > ((r->bytes + 255) & ~(size_t)255)) return set_error(NCL_ERR_INVALID, "a tensor over a registered base vector addresses %zu bytes but %zu were registered", need(&ts[i]), r->bytes);
      items.push_back({&ts[i], r ? (char*)r->dev->ptr : nullptr, r ? 0 : need(&ts[i]), is_out, r}); used += items.back().bytes;
    }
    return NCL_OK;
  }
  // byte window of the base vector a tensor's elements occupy: [lo, hi)
  static void window(const ncl_tensor* t, size_t* lo, size_t* hi, bool* ragged) {
    const size_t bits = (size_t)g_dt[t->dtype].slot_bits, b0 = (size_t)t->elem_offset * bits, b1 = ((size_t)t->elem_offset + (size_t)total_elems(t)) * bits;
    *lo = b0 / 8; *hi = (b1 + 7) / 8; *ragged = (b0 % 8) != 0 || (b1 % 8) != 0;
  }
  bool prepared = false, uploaded = false, downloaded = false, outputs_too = false;
  // arena addresses are fixed before anything is copied, so that the nest can be built first and the copies can then be
  // issued either all at once (upload) or panel by panel under the kernels (try_pipeline)
  int prepare(bool outputs_too_) {
    outputs_too = outputs_too_;
    if (prepared || items.empty()) { prepared = true; return NCL_OK; }
    if (g_ctx.capturing) return set_error(NCL_ERR_STATE, "host-resident tensors cannot be staged inside ncl_capture_begin/end");
    if (used) NCL_TRY(ensure_scratch2(used));
    size_t off = 0;
    for (auto& it : items) if (!it.res) { it.dev = (char*)g_ctx.scratch2 + off; off += it.bytes; }
    prepared = true; return NCL_OK;
  }
  int upload() {
    if (uploaded) return NCL_OK;
    uploaded = true;
    if (items.empty()) return NCL_OK;
    for (auto& it : items) {
      if (it.res) {                                                  // resident mirror: upload only when the host copy is newer
        if (it.res->state == RES_HOST_NEWER) {
          NCL_CUDA(cudaMemcpyAsync(it.res->dev->ptr, it.res->host, it.res->bytes, cudaMemcpyHostToDevice, g_ctx.stream));
          g_ctx.h2d_bytes += it.res->bytes; it.res->state = RES_IN_SYNC;
        }
        continue;
      }
      size_t live = ncl_storage_bytes(it.t->dtype, it.t->elem_offset + total_elems(it.t));
      if (!it.is_out || outputs_too) {
        NCL_CUDA(cudaMemcpyAsync(it.dev, it.t->host, live, cudaMemcpyHostToDevice, g_ctx.stream));
        g_ctx.h2d_bytes += live;
      } else {
        // a write-only output is not uploaded; but a packed (bit / ub2 / ub4) window that does not start and end on byte
        // boundaries shares its first / last byte with neighbouring elements of the base vector, which the download
        // must hand back unchanged: stage those bytes (the kernels update sub-byte fields read-modify-write)
        size_t lo, hi; bool ragged; window(it.t, &lo, &hi, &ragged);
        if (ragged && hi > lo) NCL_CUDA(cudaMemcpyAsync(it.dev + lo, (const char*)it.t->host + lo, hi - lo, cudaMemcpyHostToDevice, g_ctx.stream));
      }
    }
    return NCL_OK;
  }
  int try_pipeline(Nest& nest);
  // the nest is built; copy what it reads, run it -- in panels of its outermost kept loop when that pays (try_pipeline)
  int run(Nest& nest) {
    if (!uploaded) {
      int rc = try_pipeline(nest);
      if (rc != NCL_ERR_UNSUPPORTED) return rc;
      NCL_TRY(upload());
    }
    return run_nest(nest);
  }
  char* dev_of(const ncl_tensor* t) { for (auto& it : items) if (it.t == t) return it.dev; return nullptr; }
  int index_of(const ncl_tensor* t) { for (size_t i = 0; i < items.size(); i++) if (items[i].t == t) return (int)i; return -1; }
  // Only the bytes of the tensor's own window travel back: the arena in front of a displaced view (elem_offset > 0) and the
  // padding behind it were never initialised when the output was not uploaded, and must not overwrite the host base vector.
  int download() {
    if (downloaded) return NCL_OK;
    bool any = false;
    for (auto& it : items) if (it.is_out) {
      if (it.res && (it.res->flags & NCL_RESIDENT_LAZY)) { it.res->state = RES_DEVICE_NEWER; continue; }     // stays on the device until ncl_host_fetch
      size_t lo, hi; bool ragged; window(it.t, &lo, &hi, &ragged);
      if (hi > lo) { NCL_CUDA(cudaMemcpyAsync((char*)it.t->host + lo, it.dev + lo, hi - lo, cudaMemcpyDeviceToHost, g_ctx.stream)); any = true; g_ctx.d2h_bytes += hi - lo; }
    }
    if (any) { NCL_CUDA(cudaStreamSynchronize(g_ctx.stream)); return check_fault(); }
    return NCL_OK;
  }
};

static void operand_from(const ncl_tensor* t, Stager& st, Operand* o) {
  memset(o, 0, sizeof *o);
  o->dtype = t->dtype; o->offset = t->elem_offset; o->stage_item = -1;
  if (t->buf) { o->base = (char*)t->buf->ptr; o->n_storage = (int64_t)((t->buf->bytes * 8) / g_dt[t->dtype].slot_bits); }
  else {
    o->base = st.dev_of(t); o->stage_item = st.index_of(t);
    Resident* r = find_resident(t->host);
    o->n_storage = (int64_t)(((r ? r->bytes : Stager::need(t)) * 8) / g_dt[t->dtype].slot_bits);
  }
}
static void rowmajor_strides(const ncl_tensor* t, int64_t* s) { int64_t acc = 1; for (int i = t->rank - 1; i >= 0; i--) { s[i] = acc; acc *= t->dims[i]; } }

// ---- panel pipeline of host-resident tensors ----------------------------------------------------------------------------
// A call on host arrays used to be three phases in a row on one stream: copy in, kernels, copy out.  PCIe is full duplex and
// the copy engines run beside the SMs, so the nest is cut into PANELS of its outermost kept loop -- rows of an elementwise
// result or of a row-wise reduction, row panels of A and C in a matrix product, batches of a batched product: panel p goes up
// on one copy stream while panel p-1 is computed and panel p-2 travels back on another.  Operands that do not move with the
// loop (the row vector of C1, B of a matrix product) go up once, first.  Conditions: pinned host memory (pageable copies block
// the host), byte-addressable element types, the loop is kept by every output (panels write disjoint elements) and is the
// slowest-varying axis of every staged array (a panel is one contiguous byte range).  Everything else takes the plain path.
static size_t panel_target_bytes() { static size_t v = 0; if (!v) { const char* e = getenv("NUMCL_PANEL_MB"); long mb = e ? atol(e) : 32; if (mb < 1) mb = 1; v = (size_t)mb << 20; } return v; }
static bool pipeline_enabled() { static int on = -1; if (on < 0) { const char* e = getenv("NUMCL_NO_PIPELINE"); on = (e && atoi(e)) ? 0 : 1; } return on == 1; }
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
static int pipeline_event(int i, cudaEvent_t* ev) {
  if (i >= g_ctx.n_events) {
    int want = (i + 64) & ~63;
    cudaEvent_t* ne = (cudaEvent_t*)realloc(g_ctx.events, (size_t)want * sizeof(cudaEvent_t));
    if (!ne) return set_error(NCL_ERR_NOMEM, "event pool");
    g_ctx.events = ne;
    for (int k = g_ctx.n_events; k < want; k++) NCL_CUDA(cudaEventCreateWithFlags(&g_ctx.events[k], cudaEventDisableTiming));
    g_ctx.n_events = want;
  }
  *ev = g_ctx.events[i]; return NCL_OK;
}

int Stager::try_pipeline(Nest& nest) {
  if (items.empty() || !prepared || !pipeline_enabled() || g_ctx.capturing || g_ctx.force_generic || nest.nloops < 1) return NCL_ERR_UNSUPPORTED;
  struct Mov { Operand* op; Item* it; bool is_out; int64_t span; size_t sb; bool up, down; };
  Mov mv[2 * NCL_MAX_OPERANDS]; int nm = 0;
  Operand* all[2 * NCL_MAX_OPERANDS]; int na = 0;
  std::vector<char> seen(items.size(), 0);
  for (int k = 0; k < nest.n_in + nest.n_out; k++) {
    const bool is_out = k >= nest.n_in;
    Operand* op = is_out ? &nest.out[k - nest.n_in] : &nest.in[k];
    all[na++] = op;
    if (op->stage_item < 0) continue;
    if (op->stage_item >= (int)items.size() || seen[op->stage_item]) return NCL_ERR_UNSUPPORTED;
    seen[op->stage_item] = 1;
    Item* it = &items[op->stage_item];
    if (it->is_out != is_out || g_dt[op->dtype].slot_bits < 8) return NCL_ERR_UNSUPPORTED;
    bool up = it->res ? it->res->state == RES_HOST_NEWER && (!is_out || outputs_too) : (!is_out || outputs_too);
    bool down = is_out && !(it->res && (it->res->flags & NCL_RESIDENT_LAZY));
    mv[nm++] = Mov{op, it, is_out, 0, (size_t)g_dt[op->dtype].slot_bits / 8, up, down};
  }
  for (char c : seen) if (!c) return NCL_ERR_UNSUPPORTED;
  // the loop to cut
  int L = -1;
  for (int l = 0; l < nest.nloops && L < 0; l++) {
    if (nest.extent[l] < 2) continue;
    bool ok = true;
    for (int o = 0; o < nest.n_out && ok; o++) if (nest.out[o].stride[l] == 0) ok = false;
    for (int m = 0; m < nm && ok; m++) {
      const Operand& X = *mv[m].op; const int64_t s = X.stride[l];
      if (s == 0) continue;                                        // an input that does not move with the loop: uploaded whole
      if (s < 0) { ok = false; break; }
      int64_t span = 0;
      for (int k = 0; k < nest.nloops; k++) if (k != l) { if (X.stride[k] < 0) ok = false; span += (nest.extent[k] - 1) * X.stride[k]; }
      if (span >= s) ok = false;
      mv[m].span = span;
    }
    if (ok) L = l;
  }
  if (L < 0) return NCL_ERR_UNSUPPORTED;
  for (int k = 0; k < nest.nloops; k++) if (nest.extent[k] <= 0) return NCL_ERR_UNSUPPORTED;
  const int64_t E = nest.extent[L];
  size_t row_bytes = 0, moved = 0, out_bytes = 0;
  for (int m = 0; m < nm; m++) {
    const int64_t s = mv[m].op->stride[L];
    const size_t whole = s ? (size_t)s * mv[m].sb * (size_t)E : (size_t)(mv[m].span + 1) * mv[m].sb;
    if (mv[m].up) { moved += whole; if (s) row_bytes += (size_t)s * mv[m].sb; }
    if (mv[m].down) { moved += whole; out_bytes += whole; row_bytes += (size_t)s * mv[m].sb; }
    if ((mv[m].up || mv[m].down) && !is_pinned(mv[m].it->t->host)) return NCL_ERR_UNSUPPORTED;
  }
  if (moved < ((size_t)8 << 20) || row_bytes == 0) return NCL_ERR_UNSUPPORTED;
  // an output written back while an input that overlaps it on the host is still being read: not worth reasoning about
  for (int a = 0; a < nm; a++) if (mv[a].down) for (int b = 0; b < nm; b++) if (!mv[b].is_out) {
    size_t alo, ahi, blo, bhi; bool rg; window(mv[a].it->t, &alo, &ahi, &rg); window(mv[b].it->t, &blo, &bhi, &rg);
    const char* A = (const char*)mv[a].it->t->host; const char* B = (const char*)mv[b].it->t->host;
    if (A + alo < B + bhi && B + blo < A + ahi) return NCL_ERR_UNSUPPORTED;
  }
  bool contraction = false;
  for (int l = 0; l < nest.nloops; l++) if (nest.extent[l] > 1) { bool red = true; for (int o = 0; o < nest.n_out; o++) if (nest.out[o].stride[l] != 0) red = false; if (red) contraction = true; }
  // panel size: NUMCL_PANEL_MB (32 MiB: a copy of that size runs at the link rate and its fixed cost is below 1 %), but at least
  // eight panels for arrays of a few tens of MiB; a contraction with a stationary operand re-reads it per panel, so it gets four
  size_t target = panel_target_bytes();
  if (moved / 8 < target) target = moved / 8 > ((size_t)4 << 20) ? moved / 8 : ((size_t)4 << 20);
  bool stationary = false;                                         // e.g. B of (ij jk -> ik): re-packed / re-read by every panel product
  for (int k = 0; k < nest.n_in; k++) if (nest.in[k].stride[L] == 0) stationary = true;
  int64_t rows = contraction && stationary ? (E + 3) / 4 : (int64_t)(target / row_bytes);
  if (rows < 1) rows = 1;
  if ((E + rows - 1) / rows > 256) rows = (E + 255) / 256;
  if (rows >= 256) rows &= ~(int64_t)255; else if (rows >= 8) rows &= ~(int64_t)7;
  const int np = (int)((E + rows - 1) / rows);
  if (np < 2) return NCL_ERR_UNSUPPORTED;

  if (!g_ctx.copy_in) { NCL_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_in, cudaStreamNonBlocking)); NCL_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_out, cudaStreamNonBlocking)); }
  const bool pipe_out = out_bytes >= ((size_t)1 << 20);           // small results travel back in one piece at the end (download)
  cudaEvent_t ev;
  NCL_TRY(pipeline_event(0, &ev));
  NCL_CUDA(cudaEventRecord(ev, g_ctx.stream));                    // the arena / the mirrors may still be read by earlier kernels
  NCL_CUDA(cudaStreamWaitEvent(g_ctx.copy_in, ev, 0));
  if (pipe_out) NCL_CUDA(cudaStreamWaitEvent(g_ctx.copy_out, ev, 0));
  uploaded = true;
  auto full_hi = [&](const Mov& m) { return m.it->res ? m.it->res->bytes : ncl_storage_bytes(m.it->t->dtype, m.it->t->elem_offset + total_elems(m.it->t)); };
  auto dev_base = [&](const Mov& m) { return m.it->res ? (char*)m.it->res->dev->ptr : m.it->dev; };
  for (int m = 0; m < nm; m++) if (mv[m].up && mv[m].op->stride[L] == 0) {          // stationary operands, whole
    const size_t n = full_hi(mv[m]);
    NCL_CUDA(cudaMemcpyAsync(dev_base(mv[m]), mv[m].it->t->host, n, cudaMemcpyHostToDevice, g_ctx.copy_in)); g_ctx.h2d_bytes += n;
  }
  int rc = NCL_OK;
  for (int p = 0; p < np && rc == NCL_OK; p++) {
    const int64_t r0 = (int64_t)p * rows, r1 = r0 + rows < E ? r0 + rows : E;
    for (int m = 0; m < nm; m++) if (mv[m].up && mv[m].op->stride[L] != 0) {
      const Operand& X = *mv[m].op;
      // the first panel also carries what lies in front of the window, the last what lies behind it (padding, the rest of a mirror)
      const size_t lo = p == 0 ? 0 : (size_t)(X.offset + r0 * X.stride[L]) * mv[m].sb;
      const size_t hi = p == np - 1 ? full_hi(mv[m]) : (size_t)(X.offset + r1 * X.stride[L]) * mv[m].sb;
      if (hi > lo) { NCL_CUDA(cudaMemcpyAsync(dev_base(mv[m]) + lo, (const char*)mv[m].it->t->host + lo, hi - lo, cudaMemcpyHostToDevice, g_ctx.copy_in)); g_ctx.h2d_bytes += hi - lo; }
    }
    NCL_TRY(pipeline_event(1 + 2 * p, &ev));
    NCL_CUDA(cudaEventRecord(ev, g_ctx.copy_in));
    NCL_CUDA(cudaStreamWaitEvent(g_ctx.stream, ev, 0));
    Nest* sub = new Nest(nest);
    sub->extent[L] = r1 - r0;
    for (int k = 0; k < sub->n_in; k++) sub->in[k].offset += r0 * sub->in[k].stride[L];
    for (int k = 0; k < sub->n_out; k++) sub->out[k].offset += r0 * sub->out[k].stride[L];
    rc = run_nest(*sub);
    delete sub;
    if (rc != NCL_OK || !pipe_out) continue;
    NCL_TRY(pipeline_event(2 + 2 * p, &ev));
    NCL_CUDA(cudaEventRecord(ev, g_ctx.stream));
    NCL_CUDA(cudaStreamWaitEvent(g_ctx.copy_out, ev, 0));
    for (int m = 0; m < nm; m++) if (mv[m].down) {
      const Operand& X = *mv[m].op;
      const size_t lo = (size_t)(X.offset + r0 * X.stride[L]) * mv[m].sb, hi = (size_t)(X.offset + (r1 - 1) * X.stride[L] + mv[m].span + 1) * mv[m].sb;
      NCL_CUDA(cudaMemcpyAsync((char*)mv[m].it->t->host + lo, dev_base(mv[m]) + lo, hi - lo, cudaMemcpyDeviceToHost, g_ctx.copy_out)); g_ctx.d2h_bytes += hi - lo;
    }
  }
  for (int m = 0; m < nm; m++) if (mv[m].it->res && mv[m].up) mv[m].it->res->state = RES_IN_SYNC;
  g_ctx.pipelined_calls++; g_ctx.pipelined_panels += np;
  if (rc != NCL_OK) { cudaStreamSynchronize(g_ctx.copy_in); cudaStreamSynchronize(g_ctx.copy_out); return rc; }
  if (pipe_out) {
    for (int m = 0; m < nm; m++) if (mv[m].is_out && !mv[m].down && mv[m].it->res) mv[m].it->res->state = RES_DEVICE_NEWER;
    downloaded = true;
    NCL_CUDA(cudaStreamSynchronize(g_ctx.copy_out));
    return check_fault();
  }
  return NCL_OK;                                                   // download() brings the (small) results back and synchronises
}

// ---- einsum: shape-resolver + plan-broadcast + nest construction -----------------------------------
static int64_t normalize_index(int64_t index, int64_t dim) {            // src/3einsum.lisp:274-278
  if (index < 0) { if (dim == 0) return 0; int64_t m = index % dim; if (m < 0) m += dim; return m; }
  return index < dim ? index : dim;
}
static int64_t ceil_div(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; if (r != 0 && ((r > 0) == (b > 0))) q++; return q; }
static const ncl_iter_opt* find_opt(const ncl_iter_opt* o, int n, int key) { for (int i = 0; i < n; i++) if (o[i].key == key) return &o[i]; return nullptr; }

#define QMAX 64
struct Resolver {
  int64_t q[QMAX]; bool set[QMAX];
  int M = 0; bool bc = false;
  int64_t plan_dim[NCL_MAX_OPERANDS + 1][NCL_MAX_RANK], plan_step[NCL_MAX_OPERANDS + 1][NCL_MAX_RANK];
  int unify(int idx, int64_t w) {
    if (idx < 0 || idx >= QMAX) return set_error(NCL_ERR_INVALID, "index number %d out of range", idx);
    if (set[idx] && q[idx] != w) return set_error(NCL_ERR_SHAPE, "index %d is used with widths %lld and %lld", idx, (long long)q[idx], (long long)w);
    q[idx] = w; set[idx] = true; return NCL_OK;
  }
};
static int range_width(const ncl_tensor* t, int axis, const ncl_iter_opt* o, int64_t* w) {   // :280-285
  int64_t D = t->dims[axis];
  int64_t start = (o && o->has_start) ? o->start : 0, end = (o && o->has_end) ? o->end : D, step = o ? o->step : 1;
  if (step == 0) return set_error(NCL_ERR_INVALID, ":step 0");
  *w = ceil_div(normalize_index(end, D) - normalize_index(start, D), step);
  if (*w < 0) *w = 0;
  return NCL_OK;
}
// one array through shape-resolver (src/3einsum.lisp:287-338); block = dims of the `-` block
static int resolve_array(Resolver& R, const ncl_tensor* t, const int32_t* spec, int rank, const ncl_iter_opt* opts, int nopt,
                         const char* what, int which, int* has_block, int* block_rank, int64_t* block) {
  int bpos = -1;
  for (int k = 0; k < rank; k++) {
    if (spec[k] == -1) { if (bpos >= 0) return set_error(NCL_ERR_INVALID, "%s %d: more than one `-` in a spec", what, which); bpos = k; }
    else if (spec[k] < 0) return set_error(NCL_ERR_INVALID, "%s %d: bad index %d", what, which, spec[k]);
  }
  *has_block = bpos >= 0;
  if (bpos < 0) {
    if (t->rank != rank) return set_error(NCL_ERR_SHAPE, "%s %d has rank %d but its spec has %d indices", what, which, t->rank, rank);
    for (int i = 0; i < rank; i++) { int64_t w; NCL_TRY(range_width(t, i, find_opt(opts, nopt, spec[i]), &w)); NCL_TRY(R.unify(spec[i], w)); }
    return NCL_OK;
  }
  // The reference sizes the strides of axes that precede `-` from the plan's first step only
  // (%dimension-step-size, common-lisp.lisp:60-66), which addresses garbage unless `-` leads.
  if (bpos != 0) return set_error(NCL_ERR_UNSUPPORTED, "%s %d: `-` must be the leading axis of a spec (reference quirk, see DESIGN.md)", what, which);
  int after = rank - 1;
  if (t->rank < after) return set_error(NCL_ERR_SHAPE, "%s %d has too few axes for its spec", what, which);
  for (int i = 0; i < after; i++) {
    int axis = t->rank - 1 - i; int idx = spec[rank - 1 - i];
    int64_t w; NCL_TRY(range_width(t, axis, find_opt(opts, nopt, idx), &w)); NCL_TRY(R.unify(idx, w));
  }
  *block_rank = t->rank - after;
  for (int k = 0; k < *block_rank; k++) block[k] = t->dims[k];
  return NCL_OK;
}

static int build_einsum_nest(const ncl_einsum_desc* d, const ncl_tensor* in, ncl_tensor* out, unsigned flags,
                             Stager& st, Nest* nest) {
  if (d->n_in < 0 || d->n_in > NCL_MAX_OPERANDS || d->n_out < 1 || d->n_out > NCL_MAX_OPERANDS) return set_error(NCL_ERR_INVALID, "bad operand counts");
  if (d->n_iter < 0 || d->n_iter > NCL_MAX_INDEX) return set_error(NCL_ERR_INVALID, "bad n_iter");
  Resolver R; memset(R.set, 0, sizeof R.set); memset(R.q, 0, sizeof R.q);
  int has_block[2 * NCL_MAX_OPERANDS] = {0}, brank[2 * NCL_MAX_OPERANDS] = {0};
  int64_t block[2 * NCL_MAX_OPERANDS][NCL_MAX_RANK];
  int nb = 0;
  for (int i = 0; i < d->n_in; i++) {
    if (d->in_rank[i] < 0 || d->in_rank[i] > NCL_MAX_RANK) return set_error(NCL_ERR_INVALID, "bad spec rank");
    NCL_TRY(resolve_array(R, &in[i], d->in_spec[i], d->in_rank[i], d->in_opt[i], d->in_nopt[i], "input", i, &has_block[i], &brank[i], block[i]));
    if (has_block[i]) nb++;
  }
  if (nb) {
    // plan-broadcast (src/3einsum.lisp:378-425); the generated code indexes the plan rows by array
    // position (common-lisp.lisp:219-223), which is only right with one output and `-` on every input
    if (nb != d->n_in || d->n_out != 1) return set_error(NCL_ERR_UNSUPPORTED, "`-` needs every input to carry it and exactly one output (reference quirk)");
    R.bc = true; R.M = 0;
    for (int i = 0; i < d->n_in; i++) if (brank[i] > R.M) R.M = brank[i];
    for (int r = 0; r <= d->n_in; r++) for (int j = 0; j < NCL_MAX_RANK; j++) { R.plan_dim[r][j] = 1; R.plan_step[r][j] = 1; }
    for (int i = 0; i < d->n_in; i++) for (int k = 0; k < brank[i]; k++) R.plan_dim[i + 1][R.M - brank[i] + k] = block[i][k];
    for (int i = 0; i < d->n_in; i++) for (int j = 0; j < R.M; j++) if (R.plan_dim[i + 1][j] > R.plan_dim[0][j]) R.plan_dim[0][j] = R.plan_dim[i + 1][j];
    for (int i = 0; i < d->n_in; i++) for (int j = 0; j < R.M; j++)
      if (!(R.plan_dim[i + 1][j] == 1 || R.plan_dim[i + 1][j] == R.plan_dim[0][j])) return set_error(NCL_ERR_SHAPE, "broadcast blocks are incompatible on axis %d", j);
    for (int r = 0; r <= d->n_in; r++) for (int j = R.M - 2; j >= 0; j--) R.plan_step[r][j] = R.plan_dim[r][j + 1] * R.plan_step[r][j + 1];
  }
  for (int o = 0; o < d->n_out; o++) {
    if (d->out_rank[o] < 0 || d->out_rank[o] > NCL_MAX_RANK) return set_error(NCL_ERR_INVALID, "bad spec rank");
    for (int k = 0; k < d->out_rank[o]; k++) { int idx = d->out_spec[o][k]; if (idx >= 0 && (idx >= QMAX || !R.set[idx])) return set_error(NCL_ERR_INVALID, "output index %d is not bound by any input", idx); }
    int hb, br; int64_t bl[NCL_MAX_RANK];
    NCL_TRY(resolve_array(R, &out[o], d->out_spec[o], d->out_rank[o], d->out_opt[o], d->out_nopt[o], "output", o, &hb, &br, bl));
    has_block[NCL_MAX_OPERANDS + o] = hb;
    if (hb) {                                                   // %output-result-shapes-match-p :431-437
      if (!R.bc || br != R.M) return set_error(NCL_ERR_SHAPE, "The output shapes do not match the broadcast plan!");
      for (int k = 0; k < br; k++) if (bl[k] != R.plan_dim[0][k]) return set_error(NCL_ERR_SHAPE, "The output shapes do not match the broadcast plan!");
    }
  }
  // loops
  nest->nloops = 0; nest->n_in = d->n_in; nest->n_out = d->n_out; nest->fresh = (flags & NCL_OUT_FRESH) != 0;
  for (int i = 0; i < d->n_in; i++) operand_from(&in[i], st, &nest->in[i]);
  for (int o = 0; o < d->n_out; o++) operand_from(&out[o], st, &nest->out[o]);
  auto later_product = [&](const int32_t* spec, int rank, int ax) { int64_t ss = 1; for (int r2 = ax + 1; r2 < rank; r2++) ss *= R.q[spec[r2]]; return ss; };
  for (int it = 0; it < d->n_iter; it++) {
    int idx = d->iter[it];
    if (idx >= 0) {
      if (idx >= QMAX || !R.set[idx]) return set_error(NCL_ERR_INVALID, "iter index %d is unbound", idx);
      if (nest->nloops >= NCL_MAX_LOOPS) return set_error(NCL_ERR_UNSUPPORTED, "too many loops");
      int L = nest->nloops++;
      nest->extent[L] = R.q[idx];
      auto one = [&](Operand& op, const int32_t* spec, int rank, const ncl_iter_opt* opts, int nopt) {
        int64_t stride = 0; bool first = true;
        for (int ax = 0; ax < rank; ax++) {
          if (spec[ax] != idx) continue;
          const ncl_iter_opt* o = find_opt(opts, nopt, ax);              // (assoc a-axis a-opts): by AXIS POSITION, common-lisp.lisp:129
          int64_t start = (o && o->has_start) ? o->start : 0, step = o ? o->step : 1;
          int64_t ss = later_product(spec, rank, ax);                    // Π of the later ?dims, :132-134
          stride += step * ss;                                            // repeated index: strides add up, :137-143
          if (first) { op.offset += start * ss; first = false; }
        }
        op.stride[L] = stride;
      };
      for (int o = 0; o < d->n_out; o++) one(nest->out[o], d->out_spec[o], d->out_rank[o], d->out_opt[o], d->out_nopt[o]);
      for (int i = 0; i < d->n_in; i++) one(nest->in[i], d->in_spec[i], d->in_rank[i], d->in_opt[i], d->in_nopt[i]);
    } else if (idx == -1) {
      if (!R.bc) return set_error(NCL_ERR_INVALID, "iter-specs has `-` but no input does");
      for (int j = 0; j < R.M; j++) {                                     // %einsum-broadcast: rec over the plan axes, :176-224
        if (nest->nloops >= NCL_MAX_LOOPS) return set_error(NCL_ERR_UNSUPPORTED, "too many loops");
        int L = nest->nloops++;
        nest->extent[L] = R.plan_dim[0][j];
        for (int o = 0; o < d->n_out; o++)
          nest->out[o].stride[L] = has_block[NCL_MAX_OPERANDS + o] ? R.plan_step[0][j] * later_product(d->out_spec[o], d->out_rank[o], 0) : 0;
        for (int i = 0; i < d->n_in; i++)
          nest->in[i].stride[L] = (has_block[i] && R.plan_dim[i + 1][j] != 1) ? R.plan_step[i + 1][j] * later_product(d->in_spec[i], d->in_rank[i], 0) : 0;
      }
    } else return set_error(NCL_ERR_INVALID, "bad iter entry %d", idx);
  }
  // transforms
  int in_dt[NCL_MAX_OPERANDS], out_dt[NCL_MAX_OPERANDS];
  for (int i = 0; i < d->n_in; i++) in_dt[i] = in[i].dtype;
  for (int o = 0; o < d->n_out; o++) out_dt[o] = out[o].dtype;
  for (int o = 0; o < d->n_out; o++) NCL_TRY(type_expr(&d->transform[o], in_dt, d->n_in, out_dt, d->n_out, o, &nest->transform[o]));
  return NCL_OK;
}

// ---- small expression builders ------------------------------------------------------------------------
static void x_leaf(ncl_expr* e, int op, int a) { ncl_expr_node& n = e->node[e->n++]; memset(&n, 0, sizeof n); n.op = op; n.a = a; }
static void x_const_i(ncl_expr* e, int64_t v) { ncl_expr_node& n = e->node[e->n++]; memset(&n, 0, sizeof n); n.op = NCL_X_CONST_INT; n.ival = v; }
static void x_op(ncl_expr* e, int op, int a, int b) { ncl_expr_node& n = e->node[e->n++]; memset(&n, 0, sizeof n); n.op = op; n.a = a; n.b = b; }

static int broadcast_nest(const ncl_tensor* args, int n_args, ncl_tensor* r, Stager& st, Nest* nest) {
  // broadcast-p / broadcast-result-shape, src/4numeric.lisp:174-199: right-aligned, dims equal or 1
  int rr = r->rank;
  for (int a = 0; a < n_args; a++) if (args[a].rank > rr) return set_error(NCL_ERR_SHAPE, "result rank %d is smaller than an argument's rank %d", rr, args[a].rank);
  nest->nloops = rr; nest->n_in = n_args; nest->n_out = 1; nest->fresh = true;
  operand_from(r, st, &nest->out[0]);
  int64_t rs[NCL_MAX_RANK]; rowmajor_strides(r, rs);
  for (int k = 0; k < rr; k++) { nest->extent[k] = r->dims[k]; nest->out[0].stride[k] = rs[k]; }
  for (int a = 0; a < n_args; a++) {
    operand_from(&args[a], st, &nest->in[a]);
    int64_t as[NCL_MAX_RANK]; rowmajor_strides(&args[a], as);
    for (int k = 0; k < rr; k++) {
      int ai = args[a].rank - rr + k;
      int64_t da = ai >= 0 ? args[a].dims[ai] : 1;
      if (!(da == r->dims[k] || da == 1))
        return set_error(NCL_ERR_SHAPE, "Given arrays have incompatible shape for broadcasting: axis %d has %lld vs result %lld", k, (long long)da, (long long)r->dims[k]);
      nest->in[a].stride[k] = (ai >= 0 && da != 1) ? as[ai] : 0;
    }
  }
  for (int k = 0; k < rr; k++) {                                 // result dim must be the max of the arguments'
    int64_t mx = 1; for (int a = 0; a < n_args; a++) { int ai = args[a].rank - rr + k; if (ai >= 0 && args[a].dims[ai] > mx) mx = args[a].dims[ai]; }
    bool any_zero = false; for (int a = 0; a < n_args; a++) { int ai = args[a].rank - rr + k; if (ai >= 0 && args[a].dims[ai] == 0) any_zero = true; }
    if (!any_zero && r->dims[k] != mx) return set_error(NCL_ERR_SHAPE, "result axis %d is %lld, expected %lld", k, (long long)r->dims[k], (long long)mx);
  }
  return NCL_OK;
}

template <class F> static int with_staging(const ncl_tensor* in, int n_in, ncl_tensor* out, int n_out, bool upload_outputs, F body) {
  Stager st;
  NCL_TRY(st.plan(in, n_in, false)); NCL_TRY(st.plan(out, n_out, true));
  NCL_TRY(st.prepare(upload_outputs));
  NCL_TRY(body(st));
  return st.download();
}

extern "C" {

int ncl_einsum(const ncl_einsum_desc* desc, const ncl_tensor* in, int n_in, ncl_tensor* out, int n_out, unsigned flags) {
  LOCK; NEED_INIT;
  if (!desc) return set_error(NCL_ERR_INVALID, "desc is NULL");
  if (n_in != desc->n_in || n_out != desc->n_out) return set_error(NCL_ERR_INVALID, "einsum: %d/%d arrays for %d/%d specs", n_in, n_out, desc->n_in, desc->n_out);
  for (int i = 0; i < n_in; i++) NCL_TRY(check_tensor(&in[i], "input"));
  for (int o = 0; o < n_out; o++) NCL_TRY(check_tensor(&out[o], "output"));
  return with_staging(in, n_in, out, n_out, !(flags & NCL_OUT_FRESH), [&](Stager& st) {
    Nest* nest = new Nest;
    int rc = build_einsum_nest(desc, in, out, flags, st, nest);
    // A fresh output stands for the `zeros` the wrapper would have made (src/3einsum.lisp:468), but the nest only writes
    // the elements it visits.  When it cannot visit all of them -- an index repeated in the o-spec (`i -> ii` writes the
    // diagonal only) or an o-option (:start / :end / :step) -- the rest must read as 0, not as whatever the recycled
    // block held: make the zeros for real and accumulate into them.
    if (rc == NCL_OK && nest->fresh) {
      bool sparse = false;
      for (int o = 0; o < desc->n_out && !sparse; o++) {
        if (desc->out_nopt[o] > 0) sparse = true;
        for (int k = 0; k < desc->out_rank[o] && !sparse; k++)
          for (int k2 = k + 1; k2 < desc->out_rank[o]; k2++) if (desc->out_spec[o][k] >= 0 && desc->out_spec[o][k] == desc->out_spec[o][k2]) sparse = true;
      }
      if (sparse) {
        rc = st.upload();                                          // boundary bytes of packed windows first, then the zeros
        for (int o = 0; o < desc->n_out && rc == NCL_OK; o++)
          rc = fill_const(nest->out[o].base, out[o].dtype, out[o].elem_offset, total_elems(&out[o]), false, 0, 0.0);
        nest->fresh = false;
      }
    }
    if (rc == NCL_OK) rc = st.run(*nest);
    delete nest; return rc;
  });
}

int ncl_fold(int op, int use_identity, const ncl_tensor* args, int n, ncl_tensor* r) {
  LOCK; NEED_INIT;
  if (n < 1 || n > NCL_MAX_OPERANDS) return set_error(NCL_ERR_INVALID, "ncl_fold: %d arguments", n);
  if (op < 0 || op >= NCL_OP_BINARY_END) return set_error(NCL_ERR_INVALID, "ncl_fold: op %d is not binary", op);
  if (n == 1 && !use_identity) return set_error(NCL_ERR_INVALID, "ncl_fold: one argument needs the identity");
  for (int i = 0; i < n; i++) NCL_TRY(check_tensor(&args[i], "argument"));
  NCL_TRY(check_tensor(r, "result"));
  {
    // Three or more arrays: (numcl:+ a b c) is a chain of binary BROADCAST passes in the reference (src/4numeric.lisp:528-542),
    // each with an intermediate array of the type inferred for it.  One fused pass over three or more arrays is not a shape the
    // elementwise engine has (it ran on the generic kernel: 0.07 of the HBM peak); the chain is: every step at stream speed,
    // the intermediates in the class of the partial result (fixnum / single-float / double-float -- what the reference's own
    // intermediates hold, exactly), the last step stored through %coerce into the result.
    bool dev = r->buf != nullptr && !g_ctx.capturing && !g_ctx.force_generic;
    for (int i = 0; i < n && dev; i++) if (!args[i].buf || cls_complex(g_dt[args[i].dtype].cls) || args[i].dtype == NCL_UB64) dev = false;
    if (n >= 3 && dev && !cls_complex(g_dt[r->dtype].cls)) {
      auto join = [&](int a, int b) { int c = a > b ? a : b; if (op == NCL_OP_DIV && c == CLS_I) c = CLS_F; return c; };    // int / int is a single-float
      auto tdtype = [](int cls) { return cls == CLS_D ? NCL_F64 : cls == CLS_F ? NCL_F32 : NCL_FIXNUM; };
      // shape of a partial result: the broadcast of the arguments folded so far (right-aligned, src/4numeric.lisp:174-199)
      auto bshape = [&](ncl_tensor* t, const ncl_tensor* x) {
        const int rk = t->rank > x->rank ? t->rank : x->rank;
        int64_t d[NCL_MAX_RANK];
        for (int k = 0; k < rk; k++) {
          const int it = t->rank - rk + k, ix = x->rank - rk + k;
          const int64_t a = it >= 0 ? t->dims[it] : 1, b = ix >= 0 ? x->dims[ix] : 1;
          d[k] = a == 1 ? b : a;
        }
        t->rank = rk; for (int k = 0; k < rk; k++) t->dims[k] = d[k];
      };
      ncl_tensor acc; memset(&acc, 0, sizeof acc);
      acc.rank = args[0].rank; for (int k = 0; k < acc.rank; k++) acc.dims[k] = args[0].dims[k];
      bshape(&acc, &args[1]);
      ncl_buf* held = nullptr;
      int cls = join(g_dt[args[0].dtype].cls, g_dt[args[1].dtype].cls), rc = NCL_OK;
      acc.dtype = tdtype(cls);
      held = ncl_alloc(ncl_storage_bytes(acc.dtype, total_elems(&acc)), 0); if (!held) return NCL_ERR_NOMEM;
      acc.buf = held;
      rc = ncl_fold(op, use_identity, args, 2, &acc);
      for (int i = 2; i < n && rc == NCL_OK; i++) {
        ncl_tensor pair[2] = {acc, args[i]};
        if (i == n - 1) { rc = ncl_fold(op, 0, pair, 2, r); break; }
        const int c2 = join(cls, g_dt[args[i].dtype].cls);
        ncl_tensor nxt = acc; bshape(&nxt, &args[i]);
        bool same_shape = nxt.rank == acc.rank; for (int k = 0; k < nxt.rank && same_shape; k++) if (nxt.dims[k] != acc.dims[k]) same_shape = false;
        if (c2 == cls && same_shape) rc = ncl_fold(op, 0, pair, 2, &acc);   // in place: an element is read and written by the same thread
        else {
          nxt.dtype = tdtype(c2);
          nxt.buf = ncl_alloc(ncl_storage_bytes(nxt.dtype, total_elems(&nxt)), 0); if (!nxt.buf) { rc = NCL_ERR_NOMEM; break; }
          rc = ncl_fold(op, 0, pair, 2, &nxt);
          ncl_free(held); held = nxt.buf; acc = nxt; cls = c2;
        }
      }
      ncl_free(held);
      return rc;
    }
  }
  return with_staging(args, n, r, 1, false, [&](Stager& st) {
    Nest* nest = new Nest;
    int rc = broadcast_nest(args, n, r, st, nest);
    if (rc == NCL_OK) {
      ncl_expr e; e.n = 0;
      int acc;
      if (use_identity) {                                        // (reduce ... :initial-value ident), src/4numeric.lisp:528-542
        int64_t ident = (op == NCL_OP_MUL || op == NCL_OP_DIV) ? 1 : 0;
        x_const_i(&e, ident); x_leaf(&e, NCL_X_INPUT, 1); x_op(&e, op, 0, 1); acc = 2;
      } else { x_leaf(&e, NCL_X_INPUT, 1); acc = 0; }
      for (int i = 1; i < n; i++) { x_leaf(&e, NCL_X_INPUT, i + 1); x_op(&e, op, acc, e.n - 1); acc = e.n - 1; }
      int in_dt[NCL_MAX_OPERANDS]; for (int i = 0; i < n; i++) in_dt[i] = args[i].dtype;
      int out_dt = r->dtype;
      rc = type_expr(&e, in_dt, n, &out_dt, 1, 0, &nest->transform[0]);
    }
    if (rc == NCL_OK) rc = st.run(*nest);
    delete nest; return rc;
  });
}

int ncl_broadcast(int op, const ncl_tensor* x, const ncl_tensor* y, ncl_tensor* r) {
  if (!x || !y) return set_error(NCL_ERR_INVALID, "NULL argument");
  ncl_tensor args[2] = {*x, *y};
  return ncl_fold(op, 0, args, 2, r);
}

int ncl_map(const ncl_expr* fn, const ncl_tensor* args, int n, ncl_tensor* r) {
  LOCK; NEED_INIT;
  if (!fn || n < 1 || n > NCL_MAX_OPERANDS) return set_error(NCL_ERR_INVALID, "ncl_map: bad arguments");
  for (int i = 0; i < n; i++) NCL_TRY(check_tensor(&args[i], "argument"));
  NCL_TRY(check_tensor(r, "result"));
  for (int i = 0; i < n; i++) {                                  // (assert (equal (shape sequence) result-shape)), :41-44
    if (args[i].rank != r->rank) return set_error(NCL_ERR_SHAPE, "map-array: shapes differ");
    for (int k = 0; k < r->rank; k++) if (args[i].dims[k] != r->dims[k]) return set_error(NCL_ERR_SHAPE, "map-array: shapes differ");
  }
  return with_staging(args, n, r, 1, false, [&](Stager& st) {
    Nest* nest = new Nest;
    int rc = broadcast_nest(args, n, r, st, nest);
    int in_dt[NCL_MAX_OPERANDS]; for (int i = 0; i < n; i++) in_dt[i] = args[i].dtype;
    int out_dt = r->dtype;
    if (rc == NCL_OK) rc = type_expr(fn, in_dt, n, &out_dt, 1, 0, &nest->transform[0]);
    if (rc == NCL_OK && expr_reads_output(nest->transform[0])) rc = set_error(NCL_ERR_INVALID, "map function may not read @");
    if (rc == NCL_OK) rc = st.run(*nest);
    delete nest; return rc;
  });
}

int ncl_convert(const ncl_tensor* a, ncl_tensor* r) {
  ncl_expr e; e.n = 0; x_leaf(&e, NCL_X_INPUT, 1);
  return ncl_map(&e, a, 1, r);
}

int ncl_reduce(int op, const ncl_tensor* a, const int* axes, int n_axes, ncl_tensor* r, unsigned flags) {
  LOCK; NEED_INIT;
  NCL_TRY(check_tensor(a, "array")); NCL_TRY(check_tensor(r, "result"));
  if (!(op == NCL_OP_ADD || op == NCL_OP_MUL || op == NCL_OP_MAX || op == NCL_OP_MIN))
    return set_error(NCL_ERR_UNSUPPORTED, "reduce-array function must be one of + * max min on the CUDA backend");
  bool summed[NCL_MAX_RANK] = {false};
  if (n_axes == 0) for (int k = 0; k < a->rank; k++) summed[k] = true;            // axes nil -> all, src/5reduce.lisp:48-55
  for (int k = 0; k < n_axes; k++) { if (axes[k] < 0 || axes[k] >= a->rank) return set_error(NCL_ERR_INVALID, "bad axis %d", axes[k]); summed[axes[k]] = true; }
  int rk = 0;
  for (int ax = 0; ax < a->rank; ax++) if (!summed[ax]) { if (rk >= r->rank || r->dims[rk] != a->dims[ax]) return set_error(NCL_ERR_SHAPE, "reduce: result shape mismatch"); rk++; }
  if (rk != r->rank) return set_error(NCL_ERR_SHAPE, "reduce: result rank %d, expected %d", r->rank, rk);
  bool fresh = (flags & NCL_OUT_FRESH) != 0;
  if (g_dt[a->dtype].slot_bits < 8 && total_elems(a) > 4096) {
    // Packed arrays (bit: the result of every comparison; ub2 / ub4): the streaming reduction kernels read byte-addressable
    // slots, and the generic kernel walks a reduced axis with ONE thread per result -- (numcl:sum (numcl:< a b)) over 64 M
    // elements took 73 seconds.  Counting the ones of a whole bit array is a sum of word popcounts; everything else goes through
    // an (unsigned-byte 8) copy, which the 8-bit reduction kernels take (dp4a sums).
    if (a->host) {                                               // stage the packed bytes once, then the device path
      const size_t nb = ncl_storage_bytes(a->dtype, a->elem_offset + total_elems(a));
      ncl_buf* tb = ncl_alloc(nb, 0); if (!tb) return NCL_ERR_NOMEM;
      int rc = ncl_h2d(tb, 0, a->host, nb);
      ncl_tensor a2 = *a; a2.host = nullptr; a2.buf = tb;
      if (rc == NCL_OK) rc = ncl_reduce(op, &a2, axes, n_axes, r, flags);
      ncl_free(tb); return rc;
    }
    const int64_t total = total_elems(a);
    bool all = true; for (int ax = 0; ax < a->rank; ax++) if (!summed[ax] && a->dims[ax] != 1) all = false;
    if (a->dtype == NCL_BIT && op == NCL_OP_ADD && all && total % 32 == 0 && a->elem_offset % 32 == 0) {
      ncl_tensor w; memset(&w, 0, sizeof w); w.buf = a->buf; w.dtype = NCL_UB32; w.rank = 1; w.dims[0] = total / 32; w.elem_offset = a->elem_offset / 32;
      ncl_tensor cnt; memset(&cnt, 0, sizeof cnt); cnt.dtype = NCL_FIXNUM; cnt.rank = 1; cnt.dims[0] = total / 32;
      cnt.buf = ncl_alloc(ncl_storage_bytes(NCL_FIXNUM, total / 32), 0); if (!cnt.buf) return NCL_ERR_NOMEM;
      ncl_expr e; e.n = 0; x_leaf(&e, NCL_X_INPUT, 1); x_op(&e, NCL_OP_LOGCOUNT, 0, 0);
      int rc = ncl_map(&e, &w, 1, &cnt);
      const int all_axes = 0;
      if (rc == NCL_OK) rc = ncl_reduce(NCL_OP_ADD, &cnt, &all_axes, 0, r, flags);
      ncl_free(cnt.buf); return rc;
    }
    ncl_tensor b8 = *a; b8.dtype = NCL_UB8; b8.elem_offset = 0;
    b8.buf = ncl_alloc(ncl_storage_bytes(NCL_UB8, total), 0); if (!b8.buf) return NCL_ERR_NOMEM;
    int rc = ncl_convert(a, &b8);
    if (rc == NCL_OK) rc = ncl_reduce(op, &b8, axes, n_axes, r, flags);
    ncl_free(b8.buf); return rc;
  }
  {
    // Summed axes that do not form ONE run of adjacent axes (e.g. axes (0 2) of a rank-3 array) cannot be merged into a
    // single reduced dimension, and the nest used to fall to the thread-per-output generic kernel (130 ms for
    // 256x512x512).  Peel the last run off into a temporary of the result's element type, then reduce the rest from
    // there: two fast kernels.  Exact for integers (modular) and max/min; floats round once more, within tolerance.
    int runs = 0, t0 = -1, t1 = -1; bool empty = false;
    for (int ax = 0; ax < a->rank; ax++) {
      if (a->dims[ax] == 0) empty = true;
      if (summed[ax] && (ax == 0 || !summed[ax - 1])) { runs++; t0 = ax; }
      if (summed[ax]) t1 = ax;
    }
    if (runs > 1 && !empty && !a->host && a->buf) {                  // (a host-resident result is staged by the second pass)
      ncl_tensor tmp; memset(&tmp, 0, sizeof tmp);
      tmp.dtype = r->dtype; tmp.rank = 0;
      int64_t n = 1;
      for (int ax = 0; ax < a->rank; ax++) if (ax < t0 || ax > t1) { tmp.dims[tmp.rank++] = a->dims[ax]; n *= a->dims[ax]; }
      tmp.buf = ncl_alloc(ncl_storage_bytes(r->dtype, n), 0);
      if (!tmp.buf) return NCL_ERR_NOMEM;
      int ax1[NCL_MAX_RANK], n1 = 0, ax2[NCL_MAX_RANK], n2 = 0;
      for (int ax = t0; ax <= t1; ax++) ax1[n1++] = ax;
      for (int ax = 0; ax < t0; ax++) if (summed[ax]) ax2[n2++] = ax;
      int rc = ncl_reduce(op, a, ax1, n1, &tmp, NCL_OUT_FRESH);
      if (rc == NCL_OK) rc = ncl_reduce(op, &tmp, ax2, n2, r, flags);
      ncl_free(tmp.buf);
      return rc;
    }
  }
  return with_staging(a, 1, r, 1, !fresh, [&](Stager& st) {
    Nest* nest = new Nest;
    nest->nloops = a->rank; nest->n_in = 1; nest->n_out = 1; nest->fresh = fresh;
    operand_from(a, st, &nest->in[0]); operand_from(r, st, &nest->out[0]);
    int64_t as[NCL_MAX_RANK], rs[NCL_MAX_RANK]; rowmajor_strides(a, as); rowmajor_strides(r, rs);
    int k = 0;
    for (int ax = 0; ax < a->rank; ax++) {                       // loops nest in axis order 0..rank-1, :82-87
      nest->extent[ax] = a->dims[ax]; nest->in[0].stride[ax] = as[ax];
      nest->out[0].stride[ax] = summed[ax] ? 0 : rs[k++];
    }
    if (fresh) {                                                  // identities of sum / prod / amax / amin, :103-126 + 1constants.lisp
      nest->has_identity = true;
      long long ii; double ff; reduce_identity(op, r->dtype, &ii, &ff);
      nest->ident_i = ii; nest->ident_f = ff;
    }
    ncl_expr e; e.n = 0; x_leaf(&e, NCL_X_OUTPUT, 1); x_leaf(&e, NCL_X_INPUT, 1); x_op(&e, op, 0, 1);   // (funcall fn r a)
    int in_dt = a->dtype, out_dt = r->dtype;
    int rc = type_expr(&e, &in_dt, 1, &out_dt, 1, 0, &nest->transform[0]);
    if (rc == NCL_OK) rc = st.run(*nest);
    delete nest; return rc;
  });
}

// numcl:mean / variance / standard-deviation, src/5reduce.lisp:128-167, as ONE launch: the reduction kernels' final store
// divides the finished sum by the element count (and takes the square root).  variance is the reference's: the mean of
// SQUARES, no mean subtraction, computed as the dot-like reduction of the array with itself (the product of
// each element with itself is rounded, then added -- what (numcl:sum (square array)) does in two passes).  Shapes or
// element types the fused kernels do not take are composed from the library's own passes, exactly as the reference does.
int ncl_reduce_stat(int stat, const ncl_tensor* a, const int* axes, int n_axes, ncl_tensor* r) {
  LOCK; NEED_INIT;
  NCL_TRY(check_tensor(a, "array")); NCL_TRY(check_tensor(r, "result"));
  if (stat < NCL_STAT_MEAN || stat > NCL_STAT_STDEV) return set_error(NCL_ERR_INVALID, "ncl_reduce_stat: bad statistic %d", stat);
  const int icls = g_dt[a->dtype].cls, ocls = g_dt[r->dtype].cls;
  if (!((icls == CLS_I && r->dtype == NCL_F32) || (icls == CLS_F && r->dtype == NCL_F32) || (icls == CLS_D && r->dtype == NCL_F64)))
    return set_error(NCL_ERR_INVALID, "ncl_reduce_stat: the result of a %s array is %s", g_dt[a->dtype].name, icls == CLS_D ? "double-float" : "single-float");
  (void)ocls;
  bool summed[NCL_MAX_RANK] = {false};
  if (n_axes == 0) for (int k = 0; k < a->rank; k++) summed[k] = true;
  for (int k = 0; k < n_axes; k++) { if (axes[k] < 0 || axes[k] >= a->rank) return set_error(NCL_ERR_INVALID, "bad axis %d", axes[k]); summed[axes[k]] = true; }
  int rk = 0; int64_t count = 1;
  for (int ax = 0; ax < a->rank; ax++) {
    if (summed[ax]) { count *= a->dims[ax]; continue; }
    if (rk >= r->rank || r->dims[rk] != a->dims[ax]) return set_error(NCL_ERR_SHAPE, "reduce: result shape mismatch");
    rk++;
  }
  if (rk != r->rank) return set_error(NCL_ERR_SHAPE, "reduce: result rank %d, expected %d", r->rank, rk);
  if (count == 0) return set_error(NCL_ERR_DOMAIN, "mean over zero elements: division by zero");
  int rc = NCL_ERR_UNSUPPORTED;
  if (a->buf && r->buf && g_dt[a->dtype].slot_bits >= 8 && a->dtype != NCL_UB64) {
    Stager st;
    Nest* nest = new Nest;
    const bool sq = stat != NCL_STAT_MEAN;
    nest->nloops = a->rank; nest->n_in = sq ? 2 : 1; nest->n_out = 1; nest->fresh = true;
    nest->stat = stat == NCL_STAT_STDEV ? 3 : 1; nest->stat_count = count;
    operand_from(a, st, &nest->in[0]); if (sq) operand_from(a, st, &nest->in[1]);
    operand_from(r, st, &nest->out[0]);
    int64_t as[NCL_MAX_RANK], rs[NCL_MAX_RANK]; rowmajor_strides(a, as); rowmajor_strides(r, rs);
    int k = 0;
    for (int ax = 0; ax < a->rank; ax++) {
      nest->extent[ax] = a->dims[ax]; nest->in[0].stride[ax] = as[ax]; if (sq) nest->in[1].stride[ax] = as[ax];
      nest->out[0].stride[ax] = summed[ax] ? 0 : rs[k++];
    }
    nest->has_identity = true; nest->ident_i = 0; nest->ident_f = 0.0;
    // the transform only has to have the SHAPE the reduction families match: (+ @1 $1) / (+ @1 (* $1 $2)); the accumulator's
    // class is the input's, the epilogue converts (k_reduce.cuh: store_result)
    Expr& e = nest->transform[0]; e.n = 0;
    auto node = [&](int op, int x, int y, int cls) { XNode& nd = e.node[e.n++]; memset(&nd, 0, sizeof nd); nd.op = op; nd.a = x; nd.b = y; nd.cls = cls; };
    node(NCL_X_OUTPUT, 1, 0, icls); node(NCL_X_INPUT, 1, 0, icls);
    if (sq) { node(NCL_X_INPUT, 2, 0, icls); node(NCL_OP_MUL, 1, 2, icls); node(NCL_OP_ADD, 0, 3, icls); }
    else node(NCL_OP_ADD, 0, 1, icls);
    bool runs_ok = true;                                                  // one run of adjacent summed axes (the kernels merge it)
    { int runs = 0; for (int ax = 0; ax < a->rank; ax++) if (summed[ax] && a->dims[ax] != 1 && (ax == 0 || !summed[ax - 1])) runs++; if (runs > 1) runs_ok = false; }
    if (runs_ok) rc = run_nest(*nest);
    delete nest;
  }
  if (rc != NCL_ERR_UNSUPPORTED) return rc;
  // ---- composition: (numcl:sqrt (numcl:/ (numcl:sum [(square a)]) count)), src/5reduce.lisp:128-167 ----
  if (!a->buf || !r->buf) return set_error(NCL_ERR_UNSUPPORTED, "ncl_reduce_stat works on device-resident tensors");
  ncl_tensor sqt = *a, sumt = *r, cnt; memset(&cnt, 0, sizeof cnt);
  ncl_buf *b1 = nullptr, *b2 = nullptr, *b3 = nullptr;
  auto cleanup = [&](int code) { if (b1) ncl_free(b1); if (b2) ncl_free(b2); if (b3) ncl_free(b3); return code; };
  const ncl_tensor* src = a;
  if (stat != NCL_STAT_MEAN) {                                            // (square a): fixnum for integers (any product wraps the same way modulo 2^63)
    sqt.dtype = icls == CLS_I ? NCL_FIXNUM : a->dtype; sqt.elem_offset = 0; sqt.host = nullptr;
    b1 = ncl_alloc(ncl_storage_bytes(sqt.dtype, total_elems(a)), 0); if (!b1) return NCL_ERR_NOMEM; sqt.buf = b1;
    ncl_expr e; e.n = 0; x_leaf(&e, NCL_X_INPUT, 1); x_op(&e, NCL_OP_SQUARE, 0, 0);
    rc = ncl_map(&e, a, 1, &sqt); if (rc != NCL_OK) return cleanup(rc);
    src = &sqt;
  }
  sumt.dtype = icls == CLS_I ? NCL_FIXNUM : a->dtype; sumt.elem_offset = 0;
  b2 = ncl_alloc(ncl_storage_bytes(sumt.dtype, total_elems(r)), 0); if (!b2) return cleanup(NCL_ERR_NOMEM); sumt.buf = b2;
  rc = ncl_reduce(NCL_OP_ADD, src, axes, n_axes, &sumt, NCL_OUT_FRESH); if (rc != NCL_OK) return cleanup(rc);
  cnt.dtype = NCL_FIXNUM; cnt.rank = 0;
  b3 = ncl_alloc(64, 0); if (!b3) return cleanup(NCL_ERR_NOMEM); cnt.buf = b3;
  rc = ncl_fill(&cnt, 0, count, 0.0); if (rc != NCL_OK) return cleanup(rc);
  rc = ncl_broadcast(NCL_OP_DIV, &sumt, &cnt, r); if (rc != NCL_OK) return cleanup(rc);
  if (stat == NCL_STAT_STDEV) { ncl_expr e; e.n = 0; x_leaf(&e, NCL_X_INPUT, 1); x_op(&e, NCL_OP_SQRT, 0, 0); ncl_tensor rin = *r; rc = ncl_map(&e, &rin, 1, r); }
  return cleanup(rc);
}

int ncl_fill(ncl_tensor* t, int value_is_float, int64_t ival, double fval) {
  LOCK; NEED_INIT; NCL_TRY(check_tensor(t, "tensor"));
  if (!t->buf) return set_error(NCL_ERR_UNSUPPORTED, "ncl_fill works on device-resident tensors");
  if (cls_complex(g_dt[t->dtype].cls)) {                          // #C(value 0): a one-loop nest whose transform is the constant
    Nest* nest = new Nest; Stager st;
    nest->nloops = 1; nest->extent[0] = total_elems(t); nest->n_in = 0; nest->n_out = 1; nest->fresh = true;
    operand_from(t, st, &nest->out[0]); nest->out[0].stride[0] = 1;
    Expr& e = nest->transform[0]; e.n = 1; memset(&e.node[0], 0, sizeof e.node[0]);
    e.node[0].op = value_is_float ? NCL_X_CONST_F64 : NCL_X_CONST_INT; e.node[0].cls = value_is_float ? CLS_D : CLS_I; e.node[0].ival = ival; e.node[0].fval = fval;
    int rc = nest->extent[0] > 0 ? run_nest(*nest) : NCL_OK;
    delete nest; return rc;
  }
  return fill_const(t->buf->ptr, t->dtype, t->elem_offset, total_elems(t), value_is_float != 0, ival, fval);
}
int ncl_fill_random_at(ncl_tensor* t, uint64_t seed, uint64_t first_index, int dist, double lo, double hi) {
  LOCK; NEED_INIT; NCL_TRY(check_tensor(t, "tensor"));
  if (!t->buf) return set_error(NCL_ERR_UNSUPPORTED, "ncl_fill_random works on device-resident tensors");
  if (cls_complex(g_dt[t->dtype].cls))                            // real and imaginary parts: the stream of a float array twice as long
    return fill_random(t->buf->ptr, t->dtype == NCL_C64 ? NCL_F32 : NCL_F64, 2 * t->elem_offset, 2 * total_elems(t), seed, 2 * first_index, dist, lo, hi);
  return fill_random(t->buf->ptr, t->dtype, t->elem_offset, total_elems(t), seed, first_index, dist, lo, hi);
}
int ncl_fill_random(ncl_tensor* t, uint64_t seed, int dist, double lo, double hi) { return ncl_fill_random_at(t, seed, 0, dist, lo, hi); }

}  // extern "C"
