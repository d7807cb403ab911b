// k_fill.cu -- `full` / `zeros` / `ones` (src/2zeros.lisp:41-58) and the synthetic-input generator.
//
// Bandwidth-bound, write-only.  One thread produces one 32-bit word worth of elements for the
// sub-word types (so that no two threads share a word) and one element otherwise.
#include "ncl_device.cuh"

struct FillParams {
  char* base; long long offset, n;
  StoreSpec st; int slot_bits; int cls;
  int is_float; long long iv; double fv;        // constant fill
  unsigned long long seed; int dist; double lo, hi; int random;
  unsigned long long first;                     // element k of the tensor is element first + k of the synthetic stream
};

template <class T> __device__ __forceinline__ void fill_store_direct(char* base, long long idx, const StoreSpec& s, T v) {
  // like store_scalar but without atomics: the caller owns whole words
  if (s.kind == SK_F32) { ((float*)base)[idx] = to_f32(v); return; }
  if (s.kind == SK_F64) { ((double*)base)[idx] = to_f64(v); return; }
  unsigned long long w = (((unsigned long long)round_to_i64(v)) & s.mask) << s.shift;
  switch (s.kind) {
    case SK_8:  ((unsigned char*)base)[idx] = (unsigned char)w; break;
    case SK_16: ((unsigned short*)base)[idx] = (unsigned short)w; break;
    case SK_32: ((unsigned int*)base)[idx] = (unsigned int)w; break;
    default:    ((unsigned long long*)base)[idx] = w; break;
  }
}

// element i of the synthetic stream; mirrors nco_fill_random in oracle/numcl_oracle.c exactly
__device__ __forceinline__ void synth(const FillParams& p, long long i, long long& iv, float& f, double& d) {
  unsigned long long h = splitmix64(p.seed ^ (p.first + (unsigned long long)i));
  if (p.cls != CLS_I && p.dist != 1) {
    if (p.cls == CLS_F) {
      float u = __fmul_rn((float)(h >> 40), 0x1p-24f);
      float w = __fsub_rn((float)p.hi, (float)p.lo);
      f = __fadd_rn((float)p.lo, __fmul_rn(w, u));
    } else {
      double u = __dmul_rn((double)(h >> 11), 0x1p-53);
      double w = __dsub_rn(p.hi, p.lo);
      d = __dadd_rn(p.lo, __dmul_rn(w, u));
    }
    if (p.dist == 2) {
      unsigned long long h2 = splitmix64(h);
      if ((h2 & 1023ull) == 0) {
        int sel = (int)((h2 >> 10) % 6ull);
        const unsigned int sf[6] = {0x00000000u, 0x80000000u, 0x7f800000u, 0xff800000u, 0x00000001u, 0x7effffffu};
        const unsigned long long sd[6] = {0x0ull, 0x8000000000000000ull, 0x7ff0000000000000ull, 0xfff0000000000000ull,
                                          0x1ull, 0x7fdfffffffffffffull};
        if (p.cls == CLS_F) f = __uint_as_float(sf[sel]); else d = __longlong_as_double((long long)sd[sel]);
      }
    }
  } else {
    long long l = (long long)p.lo, hh = (long long)p.hi;
    unsigned long long span = (unsigned long long)(hh - l) + 1ull;
    iv = l + (long long)(span ? h % span : h);
    f = __ll2float_rn(iv); d = __ll2double_rn(iv);
  }
}

__global__ void fill_kernel(FillParams p) {
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  if (p.slot_bits >= 8) {
    for (long long i = tid; i < p.n; i += stride) {
      long long iv = p.iv; float f = (float)p.fv; double d = p.fv;
      if (p.random) synth(p, i, iv, f, d);
      else if (!p.is_float) { f = __ll2float_rn(iv); d = __ll2double_rn(iv); }
      if (p.cls == CLS_F) fill_store_direct<float>(p.base, p.offset + i, p.st, f);
      else if (p.cls == CLS_D) fill_store_direct<double>(p.base, p.offset + i, p.st, d);
      else if (p.random || !p.is_float) fill_store_direct<long long>(p.base, p.offset + i, p.st, iv);
      else fill_store_direct<double>(p.base, p.offset + i, p.st, p.fv);       // (round x) into an integer type
    }
  } else {
    // packed: thread owns 32-bit words; offset must be word aligned (checked on the host)
    int per = 32 / p.slot_bits;
    long long nwords = (p.n + per - 1) / per;
    unsigned int* words = (unsigned int*)p.base + (p.offset * p.slot_bits) / 32;
    for (long long w = tid; w < nwords; w += stride) {
      unsigned int acc = 0;
      for (int k = 0; k < per; k++) {
        long long i = w * per + k;
        if (i >= p.n) break;
        long long iv = p.is_float ? __double2ll_rn(p.fv) : p.iv; float f; double d;
        if (p.random) synth(p, i, iv, f, d);
        acc |= ((unsigned int)((unsigned long long)iv & p.st.mask)) << (k * p.slot_bits);
      }
      if ((w + 1) * per <= p.n) words[w] = acc;
      else {   // partial last word: keep the padding bits untouched
        unsigned int m = (unsigned int)((1ull << ((p.n - w * per) * p.slot_bits)) - 1ull);
        words[w] = (words[w] & ~m) | (acc & m);
      }
    }
  }
}

// Constant fill of a byte-addressable type: the slot pattern is produced once per thread by the SAME store
// routine the scalar kernel uses (bit-identical by construction), replicated to 16 bytes, and streamed
// with 128-bit stores over the 16-byte aligned body; the unaligned head and tail are scalar.
// (write-only streams want many CTAs: 32 per SM measured best, tools/micro/write_bw.cu)
__global__ void __launch_bounds__(256) fill_const_vec_kernel(FillParams p) {
  unsigned long long tmp[2] = {0ull, 0ull};
  {
    long long iv = p.iv; float f = (float)p.fv; double d = p.fv;
    if (!p.is_float) { f = __ll2float_rn(iv); d = __ll2double_rn(iv); }
    if (p.cls == CLS_F) fill_store_direct<float>((char*)tmp, 0, p.st, f);
    else if (p.cls == CLS_D) fill_store_direct<double>((char*)tmp, 0, p.st, d);
    else if (!p.is_float) fill_store_direct<long long>((char*)tmp, 0, p.st, iv);
    else fill_store_direct<double>((char*)tmp, 0, p.st, p.fv);
  }
  const int sb = p.slot_bits >> 3;                                  // slot bytes: 1, 2, 4, 8
  unsigned long long pat = tmp[0];
  if (sb == 1) pat = (pat & 0xffull) * 0x0101010101010101ull;
  else if (sb == 2) pat = (pat & 0xffffull) * 0x0001000100010001ull;
  else if (sb == 4) pat = (pat & 0xffffffffull) * 0x0000000100000001ull;
  char* start = p.base + p.offset * sb;
  const long long nbytes = p.n * sb;
  long long head = (16 - (long long)((uintptr_t)start & 15)) & 15;
  if (head > nbytes) head = nbytes;
  const long long body = (nbytes - head) >> 4;                      // 16-byte chunks
  const long long tail0 = head + (body << 4);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const uint4 v = make_uint4((unsigned)pat, (unsigned)(pat >> 32), (unsigned)pat, (unsigned)(pat >> 32));
  uint4* b = (uint4*)(start + head);
  for (long long i = tid; i < body; i += stride) b[i] = v;
  // head and tail: at most 15 bytes each
  for (long long o = tid * sb; o < head; o += stride * sb) memcpy(start + o, &pat, sb);
  for (long long o = tail0 + tid * sb; o < nbytes; o += stride * sb) memcpy(start + o, &pat, sb);
}

static int launch_fill(FillParams& p) {
  Context& c = ctx();
  if (p.n <= 0) return NCL_OK;
  if (p.slot_bits < 8 && ((p.offset * p.slot_bits) % 32) != 0)
    return set_error(NCL_ERR_UNSUPPORTED, "fill of a packed array needs a 32-bit aligned element offset");
  if (!p.random && p.slot_bits >= 8 && p.n * (p.slot_bits >> 3) >= 4096) {
    long long chunks = (p.n * (p.slot_bits >> 3)) >> 4;
    long long blocks = (chunks + 255) / 256, cap = (long long)c.sm_count * 32;
    if (blocks > cap) blocks = cap;
    fill_const_vec_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
    note_launch("fill_vec");
    return cuda_fail(cudaGetLastError(), "fill_const_vec_kernel");
  }
  long long work = p.slot_bits >= 8 ? p.n : (p.n * p.slot_bits + 31) / 32;
  long long blocks = (work + 255) / 256;
  long long cap = (long long)c.sm_count * 16;
  if (blocks > cap) blocks = cap;
  fill_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  note_launch("fill");
  return cuda_fail(cudaGetLastError(), "fill_kernel");
}

int fill_const(void* base, int dtype, int64_t offset, int64_t n, bool is_float, int64_t iv, double fv) {
  if (cls_complex(g_dt[dtype].cls)) {                              // zeros of a complex type are zero words; other constants: ncl_fill builds a nest
    if ((is_float && fv != 0.0) || (!is_float && iv != 0)) return set_error(NCL_ERR_UNSUPPORTED, "fill_const: non-zero complex constant");
    return dtype == NCL_C64 ? fill_const(base, NCL_F64, offset, n, true, 0, 0.0) : fill_const(base, NCL_F64, 2 * offset, 2 * n, true, 0, 0.0);
  }
  FillParams p; memset(&p, 0, sizeof p);
  p.base = (char*)base; p.offset = offset; p.n = n; p.st = dtype_store_spec(dtype);
  p.slot_bits = g_dt[dtype].slot_bits; p.cls = g_dt[dtype].cls;
  p.is_float = is_float; p.iv = iv; p.fv = is_float ? fv : (double)iv; p.random = 0;
  return launch_fill(p);
}

int fill_random(void* base, int dtype, int64_t offset, int64_t n, uint64_t seed, uint64_t first_index, int dist, double lo, double hi) {
  FillParams p; memset(&p, 0, sizeof p);
  p.base = (char*)base; p.offset = offset; p.n = n; p.st = dtype_store_spec(dtype);
  p.slot_bits = g_dt[dtype].slot_bits; p.cls = g_dt[dtype].cls;
  p.seed = seed; p.first = first_index; p.dist = dist; p.lo = lo; p.hi = hi; p.random = 1;
  return launch_fill(p);
}
