// common.cuh -- arithmetic policies, exact division, look-back scan and staging helpers shared by
// the sm_100a kernels of the DCTZ hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dctz {

constexpr int BLK = 64;            // BLK_SZ, dctz.h:28
constexpr int TILE_BLOCKS = 128;   // blocks per CTA tile == threads per CTA (one block per thread)
constexpr int NWARPS = TILE_BLOCKS / 32;

// ------------------------------------------------------------------------------------------
// Arithmetic policies for the generated DCT (dct64_gen.cuh).  Only *_rn intrinsics: nvcc never
// contracts or re-associates them, so the flow graph is executed exactly as generated.
// ------------------------------------------------------------------------------------------
struct ArithD {
  typedef double V;
  static __device__ __forceinline__ constexpr double cst(double k) { return k; }
  static __device__ __forceinline__ V add(V a, V b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ V sub(V a, V b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ V mul(V a, double k) { return __dmul_rn(a, k); }
  static __device__ __forceinline__ V fma(V a, double k, V b) { return __fma_rn(a, k, b); }
  static __device__ __forceinline__ V fms(V a, double k, V b) { return __fma_rn(a, k, -b); }
  static __device__ __forceinline__ V neg(V a) { return -a; }
};

struct ArithF {
  typedef float V;
  static __device__ __forceinline__ constexpr float cst(double k) { return (float)k; }
  static __device__ __forceinline__ V add(V a, V b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ V sub(V a, V b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ V mul(V a, float k) { return __fmul_rn(a, k); }
  static __device__ __forceinline__ V fma(V a, float k, V b) { return __fmaf_rn(a, k, b); }
  static __device__ __forceinline__ V fms(V a, float k, V b) { return __fmaf_rn(a, k, -b); }
  static __device__ __forceinline__ V neg(V a) { return -a; }
};

template <typename T> struct ArithOf;
template <> struct ArithOf<double> { typedef ArithD type; };
template <> struct ArithOf<float> { typedef ArithF type; };

// ------------------------------------------------------------------------------------------
// Exact division by a loop-invariant divisor b with y = RN(1/b) precomputed.
//   q0 = RN(a*y); r = a - q0*b (exact, FMA); q1 = RN(q0 + r*y)
// Markstein's theorem: if y is the correctly rounded reciprocal and q0 is a faithful rounding of
// a/b, then q1 == RN(a/b) (b's significand not all ones; no under/overflow).  q0 is faithful when
// the relative error rho of y is < 2^-(p+1); otherwise one more correction step makes it so.
// The host / finalize kernel measures rho exactly and picks `iters` (see make_divisor()).
//   iters: 0 -> b == 1 (identity), 1, 2 -> FMA corrections, 3 -> IEEE division (degenerate b).
// Deviation window (documented in DESIGN.md): |a| below 2^-969 (double) / 2^-102 (float) lets r
// underflow, where the last bit of the quotient may differ from IEEE division.
// ------------------------------------------------------------------------------------------
template <typename T> struct Divisor { T b, y; int iters; };

__device__ __forceinline__ double div_exact(double a, const Divisor<double> &d) {
  double q = __dmul_rn(a, d.y);
  double r = __fma_rn(-q, d.b, a);
  q = __fma_rn(r, d.y, q);
  if (d.iters >= 2) {
    r = __fma_rn(-q, d.b, a);
    q = __fma_rn(r, d.y, q);
    if (d.iters == 3) q = __ddiv_rn(a, d.b);
  }
  return q;
}
__device__ __forceinline__ float div_exact(float a, const Divisor<float> &d) {
  float q = __fmul_rn(a, d.y);
  float r = __fmaf_rn(-q, d.b, a);
  q = __fmaf_rn(r, d.y, q);
  if (d.iters >= 2) {
    r = __fmaf_rn(-q, d.b, a);
    q = __fmaf_rn(r, d.y, q);
    if (d.iters == 3) q = __fdiv_rn(a, d.b);
  }
  return q;
}

__host__ __device__ inline Divisor<double> make_divisor(double b) {
  Divisor<double> d;
  d.b = b;
#ifdef __CUDA_ARCH__
  d.y = __drcp_rn(b);
  const double rho = fabs(__fma_rn(d.y, b, -1.0));
  const unsigned long long m = (unsigned long long)__double_as_longlong(b) & 0xFFFFFFFFFFFFFull;
#else
  d.y = 1.0 / b;
  const double rho = fabs(fma(d.y, b, -1.0));
  unsigned long long bits; memcpy(&bits, &b, 8);
  const unsigned long long m = bits & 0xFFFFFFFFFFFFFull;
#endif
  if (b == 1.0) d.iters = 0;
  else if (!(b == b) || b == 0.0 || m == 0xFFFFFFFFFFFFFull || !(rho < 1.0)) d.iters = 3;
  else d.iters = (rho < 5.5511151231257e-17 /* 2^-54 (1 - 2^-20) */) ? 1 : 2;
  return d;
}
__host__ __device__ inline Divisor<float> make_divisor(float b) {
  Divisor<float> d;
  d.b = b;
#ifdef __CUDA_ARCH__
  d.y = __frcp_rn(b);
  const float rho = fabsf(__fmaf_rn(d.y, b, -1.0f));
  const unsigned m = (unsigned)__float_as_int(b) & 0x7FFFFFu;
#else
  d.y = 1.0f / b;
  const float rho = fabsf(fmaf(d.y, b, -1.0f));
  unsigned bits; memcpy(&bits, &b, 4);
  const unsigned m = bits & 0x7FFFFFu;
#endif
  if (b == 1.0f) d.iters = 0;
  else if (!(b == b) || b == 0.0f || m == 0x7FFFFFu || !(rho < 1.0f)) d.iters = 3;
  else d.iters = (rho < 2.9802294e-08f /* 2^-25 (1 - 2^-20) */) ? 1 : 2;
  return d;
}

// ------------------------------------------------------------------------------------------
// conv_tbl (dctz-comp-lib.c:27-43) in closed form: ordinal t (0..254) -> centre-out id,
// conv(t) = (t <= 127) ? 254 - 2t : 2t - 255 = max(254 - 2t, 2t - 255).
// Its inverse is the identity conv_tbl_i (dctz-decomp-lib.c:23-39) followed by gen_bins'
// centre(id) = ((id & 1) ? id/2 + 1 : -(id/2)) * bin_width (binning.c:19-22).
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned conv_ordinal(int t) {
  const int a = 254 - 2 * t, b = 2 * t - 255;
  return (unsigned)(a > b ? a : b);
}
__host__ __device__ __forceinline__ int center_multiple(unsigned id) {
  return (id & 1u) ? (int)(id >> 1) + 1 : -(int)(id >> 1);
}

// ------------------------------------------------------------------------------------------
// cp.async (LDGSTS) helpers: 16-byte global -> shared copies with a per-thread destination, which
// lets us store a contiguous tile with an XOR swizzle so that the later one-row-per-thread
// 128-bit shared loads are bank-conflict free.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Row r (one 64-element block) of a tile holds CH = 64*sizeof(T)/16 chunks of 16 bytes; logical
// chunk c is stored at physical chunk c ^ (r & (CH-1)).
template <typename T> struct TileLayout {
  static constexpr int ROW_BYTES = BLK * (int)sizeof(T);
  static constexpr int CH = ROW_BYTES / 16;
  static constexpr int TILE_BYTES = TILE_BLOCKS * ROW_BYTES;
  static __device__ __forceinline__ unsigned offset(int row, int chunk) {
    return (unsigned)(row * ROW_BYTES + ((chunk ^ (row & (CH - 1))) << 4));
  }
};

// ------------------------------------------------------------------------------------------
// Decoupled look-back state, one 64-bit word per tile: [63:62] flag, [61:46] launch epoch,
// [45:0] value.  A word from another epoch reads as "not ready", so the array never needs to be
// cleared between launches (the host wraps the epoch and clears once every 65535 launches).
// ------------------------------------------------------------------------------------------
constexpr unsigned long long LB_AGG = 1ull, LB_INC = 2ull;
constexpr unsigned long long LB_VALUE_MASK = (1ull << 46) - 1;
__device__ __forceinline__ unsigned long long lb_pack(unsigned long long flag, unsigned epoch, unsigned long long v) {
  return (flag << 62) | ((unsigned long long)(epoch & 0xFFFFu) << 46) | (v & LB_VALUE_MASK);
}
__device__ __forceinline__ void lb_store(unsigned long long *p, unsigned long long w) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(w) : "l"(p) : "memory");
  return w;
}

// Control block shared by all CTAs of a persistent kernel.
struct TileControl {
  unsigned ticket;   // next tile index (dynamic scheduling; tickets are handed out in tile order)
  unsigned done;     // CTAs that have exited; the last one resets both fields for the next launch
};

// Executed by warp 0 of a CTA: exclusive prefix of this tile's `total` over all previous tiles.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *status, unsigned tile,
                                                                 unsigned epoch, unsigned long long total,
                                                                 int lane) {
  if (tile == 0) {
    if (lane == 0) lb_store(&status[0], lb_pack(LB_INC, epoch, total));
    return 0ull;
  }
  if (lane == 0) lb_store(&status[tile], lb_pack(LB_AGG, epoch, total));
  unsigned long long excl = 0;
  long long idx = (long long)tile - 1;
  while (true) {
    const long long my = idx - lane;
    unsigned long long w;
    bool ready;
    do {
      w = (my >= 0) ? lb_load(&status[my]) : lb_pack(LB_INC, epoch, 0);
      ready = ((unsigned)(w >> 46) & 0xFFFFu) == (epoch & 0xFFFFu) && (w >> 62) != 0;
    } while (!__all_sync(0xFFFFFFFFu, ready));
    const unsigned inc_mask = __ballot_sync(0xFFFFFFFFu, (w >> 62) == LB_INC);
    unsigned long long v = w & LB_VALUE_MASK;
    if (inc_mask) {
      const int first = __ffs(inc_mask) - 1;
      if (lane > first) v = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    excl += v;
    if (inc_mask) break;
    idx -= 32;
  }
  if (lane == 0) lb_store(&status[tile], lb_pack(LB_INC, epoch, excl + total));
  return excl;
}

// CTA-wide exclusive scan of one small count per thread (TILE_BLOCKS threads) + look-back.
// Returns this thread's exclusive offset inside the tile; *tile_total and *tile_base (global
// exclusive prefix of the tile) are broadcast to every thread.  Contains two __syncthreads().
struct ScanSmem {
  unsigned wsum[NWARPS];
  unsigned woff[NWARPS];
  unsigned total;
  unsigned long long base;
};
__device__ __forceinline__ unsigned tile_scan(unsigned cnt, ScanSmem &s, unsigned long long *status, unsigned tile,
                                              unsigned epoch, unsigned *tile_total, unsigned long long *tile_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s.wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned w = (lane < NWARPS) ? s.wsum[lane] : 0u;
    unsigned wi = w;
#pragma unroll
    for (int o = 1; o < NWARPS; o <<= 1) {
      const unsigned n = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= o) wi += n;
    }
    const unsigned total = __shfl_sync(0xFFFFFFFFu, wi, NWARPS - 1);
    if (lane < NWARPS) s.woff[lane] = wi - w;
    const unsigned long long base = lookback_exclusive(status, tile, epoch, total, lane);
    if (lane == 0) { s.total = total; s.base = base; }
  }
  __syncthreads();
  *tile_total = s.total;
  *tile_base = s.base;
  return s.woff[warp] + incl - cnt;
}

}  // namespace dctz
