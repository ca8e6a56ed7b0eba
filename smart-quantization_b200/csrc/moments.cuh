// Streaming moments (count, mean, M2) and their combination: shared by the grid statistics
// kernel, the single-block small-tensor kernels and the multi-tensor kernel.
#pragma once
#include "common.cuh"

namespace smaq {

struct Moments {
  double n, mean, m2;
};

// 1 / n for a count n (an integer below 2^53 held in a double), to within an ulp or two: float reciprocal plus
// two Newton steps.  A correctly rounded fp64 division is a ~40-instruction dependent chain, and a statistics
// pass over a SMALL tensor is nothing but ~30 merges in sequence (per-thread chunks, five shuffle rounds, the
// warps, the blocks): with the division the pass took 12 us however small the tensor.  The weight only scales
// the difference of two partial means, so an ulp in it moves the mean by ~1e-16 of that difference.
__device__ __forceinline__ double count_reciprocal(double n) {
  double r = (double)__frcp_rn((float)n);
  r = fma(r, fma(-n, r, 1.0), r);
  r = fma(r, fma(-n, r, 1.0), r);
  return r;
}

__device__ __forceinline__ Moments merge(const Moments& a, const Moments& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Moments r;
  r.n = a.n + b.n;
  double delta = b.mean - a.mean;
  double w = b.n * count_reciprocal(r.n);
  r.mean = a.mean + delta * w;
  r.m2 = a.m2 + b.m2 + delta * delta * (a.n * w);
  if (!(fabs(delta) <= 1.7976931348623157e308)) {
    // an infinite (or NaN) partial mean: torch's sum-then-divide keeps inf + inf = inf where the difference form
    // would make NaN of it, and its second moment about an infinite mean is NaN
    r.mean = a.mean * (a.n * count_reciprocal(r.n)) + b.mean * w;
    r.m2 = delta - delta;
  }
  return r;
}

__device__ __forceinline__ Moments shfl_xor(const Moments& m, int o) {
  Moments r;
  r.n = __shfl_xor_sync(0xffffffffu, m.n, o);
  r.mean = __shfl_xor_sync(0xffffffffu, m.mean, o);
  r.m2 = __shfl_xor_sync(0xffffffffu, m.m2, o);
  return r;
}

// What one thread carries through the pass.  kind 0: moments of x; kind 1: moments of x plus
// min/max (range std); kind 2: S2FP8's L = (x == 0 ? 0 : log2|x|): sum of L and max of L.
struct Acc {
  Moments m;
  float lo, hi;
};

// S2FP8's L = (x == 0 ? 0 : log2|x|) (s2fp8.py:34-35) for the MEAN: lg2.approx (MUFU.LG2, absolute error ~2^-22)
// — the mean of 2^30 of them stays within 1e-7 of the exact one, well inside the 1e-6 bar on statistics, and the
// pass becomes HBM-bound instead of libdevice-bound.  The MAXIMUM of L, which fixes alpha, is not taken from
// these: it is log2f (libdevice, what torch evaluates) of max|x|, see finalize<2>.

// ---- kind 2 (S2FP8): sum of L and min/max of |x| over a chunk held in registers ----------------------------
// Only the MEAN of L is needed (no second moment), so the pass carries a double sum and an element count and
// turns them into a Moments {n, mean, 0} at the end.  min|x| / max|x| propagate NaN like torch.max.
struct LogSum {
  double sum;
  long long count;
};
__device__ __forceinline__ float lg2_ftz(float a) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float stats_min3_nan(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float stats_max3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// A chunk whose smallest magnitude is a normal number (the common case) costs one MUFU.LG2 and one add per
// element plus one three-input min/max: flush-to-zero lg2 is exact enough only there.  A chunk that holds a zero
// (L = 0 by the reference's torch.where), a subnormal (needs the pre-scaled lg2) or a NaN takes the per-element
// form.  kCount is 1 + 2k so that the running lo/hi complete the last three-input operation.
template <int kCount>
__device__ __forceinline__ void log_chunk(LogSum& ls, float& lo, float& hi, const float (&raw)[kCount]) {
  float a[kCount];
#pragma unroll
  for (int i = 0; i < kCount; ++i) a[i] = fabsf(raw[i]);
  float cmin = a[0], cmax = a[0];
#pragma unroll
  for (int i = 1; i + 1 < kCount; i += 2) {
    cmin = stats_min3_nan(cmin, a[i], a[i + 1]);
    cmax = stats_max3_nan(cmax, a[i], a[i + 1]);
  }
  if ((kCount & 1) == 0) {  // even count: the last element pairs with the running value
    lo = stats_min3_nan(lo, cmin, a[kCount - 1]);
    hi = stats_max3_nan(hi, cmax, a[kCount - 1]);
  } else {
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(lo) : "f"(lo), "f"(cmin));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(hi) : "f"(hi), "f"(cmax));
  }
  float chunk_min = cmin;
  if ((kCount & 1) == 0) asm("min.NaN.f32 %0, %1, %2;" : "=f"(chunk_min) : "f"(cmin), "f"(a[kCount - 1]));
  float l[kCount];
  if (chunk_min >= 1.17549435e-38f) {  // false for NaN
#pragma unroll
    for (int i = 0; i < kCount; ++i) l[i] = lg2_ftz(a[i]);
  } else {
#pragma unroll
    for (int i = 0; i < kCount; ++i) l[i] = a[i] == 0.0f ? 0.0f : __log2f(a[i]);
  }
  // pairwise tree: short dependency chains, fixed order
#pragma unroll
  for (int w = 1; w < kCount; w <<= 1) {
#pragma unroll
    for (int i = 0; i + w < kCount; i += 2 * w) l[i] += l[i + w];
  }
  ls.sum += (double)l[0];
  ls.count += kCount;
}

// Merge a chunk of kCount transformed values held in registers.
template <int kKind, int kCount>
__device__ __forceinline__ void merge_chunk(Acc& acc, const float (&v)[kCount]) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kCount; ++i) s += v[i];
  float cm = s * (1.0f / kCount);
  float m2 = 0.0f;
#pragma unroll
  for (int i = 0; i < kCount; ++i) {
    float d = v[i] - cm;
    m2 = __fmaf_rn(d, d, m2);
  }
  Moments c;
  c.n = (double)kCount;
  // use the exact fp32 sum for the mean so that no chunk-level rounding of s/kCount leaks in
  c.mean = (double)s * (1.0 / kCount);
  double corr = c.mean - (double)cm;  // M2 about cm -> M2 about c.mean
  double m2d = (double)m2;
  // fp32 squares are only good between underflow and overflow: a chunk whose sum of squares is below
  // 16 * FLT_MIN * 2^24 (deviations under ~1e-15: a tensor of 1e-30s would get std 0), huge or NaN is summed
  // again in fp64 — one compare per chunk for ordinary data; an all-equal chunk (zeros) stops at the max
  if (!(m2 >= 3.2e-30f && m2 <= 1e30f)) {
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < kCount; ++i) amax = fmaxf(amax, fabsf(v[i] - cm));  // NaN-ignoring: NaN handled below
    if (amax > 0.0f || !(m2 == m2)) {
      m2d = 0.0;
#pragma unroll
      for (int i = 0; i < kCount; ++i) {
        const double d = (double)v[i] - (double)cm;
        m2d = fma(d, d, m2d);
      }
    }
  }
  c.m2 = m2d - (double)kCount * corr * corr;
  if (c.m2 < 0.0) c.m2 = 0.0;
  acc.m = merge(acc.m, c);
  if (kKind == 1) {
#pragma unroll
    for (int i = 0; i < kCount; ++i) {
      // NaN-propagating min/max like torch.max / torch.min
      acc.hi = (v[i] > acc.hi || v[i] != v[i]) ? v[i] : acc.hi;
      acc.lo = (v[i] < acc.lo || v[i] != v[i]) ? v[i] : acc.lo;
    }
  }
}

template <int kKind>
__device__ __forceinline__ void merge_one(Acc& acc, float v) {
  float a[1] = {v};
  merge_chunk<kKind, 1>(acc, a);
}

constexpr int kStatsThreads = 256;
constexpr int kPartialDoubles = 5;  // n, mean, m2, lo, hi

__device__ __forceinline__ float nanmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ float nanmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }

// ---- block-level combination without merges -----------------------------------------------------------------
// Chan's pairwise merge costs a reciprocal and ~20 dependent fp64 operations, and a tree over a block is 8 of
// them in sequence — twice per statistics kernel, which made a pass over a 256 KB tensor take 9 us.  The same
// combination as two plain sums:   N = sum n_t,  mean = sum(n_t mean_t) / N,
//                                  M2 = sum(m2_t + n_t (mean_t - mean)^2)
// (the second needs the first's result): fp64 additions through a fixed shuffle tree — deterministic — and one
// division.  Infinite / NaN partial means propagate as in merge(): mean = inf, M2 = NaN.
__device__ __forceinline__ double shfl_xor_f64(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// Sum of `a` and of `b` over the block, NaN-propagating max of hi and min of lo; the result is valid in EVERY
// thread.  smem: one Acc per warp (fields reused: m.n <- a, m.mean <- b).  Ends with a barrier: smem is free.
template <bool kMinMax>
__device__ __forceinline__ void block_sum2(double& a, double& b, float& hi, float& lo, Acc* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += shfl_xor_f64(a, o);
    b += shfl_xor_f64(b, o);
    if (kMinMax) {
      hi = nanmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      lo = nanmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    }
  }
  if (lane_id() == 0) {
    smem[warp_id()].m.n = a;
    smem[warp_id()].m.mean = b;
    smem[warp_id()].hi = hi;
    smem[warp_id()].lo = lo;
  }
  __syncthreads();
  const int warps = (int)(blockDim.x >> 5);
  a = smem[0].m.n;
  b = smem[0].m.mean;
  hi = smem[0].hi;
  lo = smem[0].lo;
  for (int w = 1; w < warps; ++w) {  // same order in every thread
    a += smem[w].m.n;
    b += smem[w].m.mean;
    if (kMinMax) {
      hi = nanmax(hi, smem[w].hi);
      lo = nanmin(lo, smem[w].lo);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double weighted_mean(double n, double s1) { return n > 0.0 ? s1 / n : 0.0; }
// a partial's contribution to M2 about the combined mean
__device__ __forceinline__ double m2_about(const Moments& m, double mean) {
  if (m.n == 0.0) return 0.0;
  const double d = m.mean - mean;
  return m.m2 + m.n * (d * d);
}

template <int kKind>
__device__ __forceinline__ Acc block_combine(Acc acc, Acc* smem /* [warps] */) {
  double n = acc.m.n, s1 = acc.m.n == 0.0 ? 0.0 : acc.m.n * acc.m.mean;
  float hi = acc.hi, lo = acc.lo;
  block_sum2<kKind != 0>(n, s1, hi, lo, smem);
  const double mean = weighted_mean(n, s1);
  double q = m2_about(acc.m, mean), unused = 0.0;
  float h2 = 0.f, l2 = 0.f;
  block_sum2<false>(q, unused, h2, l2, smem);
  Acc r;
  r.m = Moments{n, mean, q};
  r.hi = hi;
  r.lo = lo;
  return r;  // valid in every thread
}

// ---- one pass over a tensor, and the grid-level hand-over --------------------------------------------------
// Shared by the statistics kernel (smaq_stats.cu) and the single-launch statistics + round trip kernel
// (smaq_roundtrip.cu).  Workspace: an arrival ticket and one record of kPartialDoubles doubles per block.
struct StatsWs {
  unsigned int ticket;
  unsigned int pad[3];
  double partials[1];  // [grid][kPartialDoubles]
};

// Thread `tid` of `nthreads`: a strided walk over the tensor, 4 independent 128-bit loads in flight per thread
// (each warp-level load is 512 contiguous bytes), 16 values per chunked-Welford update.
template <int kKind, bool kAligned>
__device__ __forceinline__ Acc accumulate_tensor(const float* __restrict__ x, int64_t n, int64_t tid, int64_t nthreads) {
  Acc acc;
  acc.m = Moments{0.0, 0.0, 0.0};
  acc.hi = -INFINITY;
  acc.lo = INFINITY;
  LogSum ls = {0.0, 0};  // kind 2 only
  if (kAligned) {
    const float4* xv = reinterpret_cast<const float4*>(x);
    const int64_t nvec = n >> 2;
    int64_t v = tid;
    // Software pipeline: the four loads of the NEXT 16-element chunk are issued before the current chunk is
    // merged, and the up to three single vectors left at the end go out together, ahead of the last chunk's
    // merge — a thread's pass is one memory round trip plus arithmetic, not one round trip per chunk (a
    // 2^24-element tensor was ten dependent round trips per thread: 20 us for 64 MB).  Chunks are merged in the
    // same order as a plain loop would merge them: same bits.
    float4 a, b, c, d;
    bool have = v + 3 * nthreads < nvec;
    if (have) {
      a = ldg_stream(xv + v), b = ldg_stream(xv + v + nthreads), c = ldg_stream(xv + v + 2 * nthreads),
      d = ldg_stream(xv + v + 3 * nthreads);
      v += 4 * nthreads;
    }
    while (have) {
      const float r[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
      have = v + 3 * nthreads < nvec;
      if (have) {
        a = ldg_stream(xv + v), b = ldg_stream(xv + v + nthreads), c = ldg_stream(xv + v + 2 * nthreads),
        d = ldg_stream(xv + v + 3 * nthreads);
        v += 4 * nthreads;
      } else {
        // the last whole chunk: its merge overlaps the single vectors' flight
        const bool h0 = v < nvec, h1 = v + nthreads < nvec, h2 = v + 2 * nthreads < nvec;
        if (h0) a = ldg_stream(xv + v);
        if (h1) b = ldg_stream(xv + v + nthreads);
        if (h2) c = ldg_stream(xv + v + 2 * nthreads);
        if (kKind == 2) log_chunk<16>(ls, acc.lo, acc.hi, r);
        else merge_chunk<kKind, 16>(acc, r);
        v = nvec;  // consumed below
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const bool ht = t == 0 ? h0 : t == 1 ? h1 : h2;
          const float4 q = t == 0 ? a : t == 1 ? b : c;
          if (ht) {
            const float r4[4] = {q.x, q.y, q.z, q.w};
            if (kKind == 2) log_chunk<4>(ls, acc.lo, acc.hi, r4);
            else merge_chunk<kKind, 4>(acc, r4);
          }
        }
        break;
      }
      if (kKind == 2) log_chunk<16>(ls, acc.lo, acc.hi, r);
      else merge_chunk<kKind, 16>(acc, r);
    }
    // a thread with fewer than four vectors in all: they go out together
    if (v < nvec) {
      const bool h1 = v + nthreads < nvec, h2 = v + 2 * nthreads < nvec;
      a = ldg_stream(xv + v);
      if (h1) b = ldg_stream(xv + v + nthreads);
      if (h2) c = ldg_stream(xv + v + 2 * nthreads);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const bool ht = t == 0 ? true : t == 1 ? h1 : h2;
        const float4 q = t == 0 ? a : t == 1 ? b : c;
        if (ht) {
          const float r4[4] = {q.x, q.y, q.z, q.w};
          if (kKind == 2) log_chunk<4>(ls, acc.lo, acc.hi, r4);
          else merge_chunk<kKind, 4>(acc, r4);
        }
      }
    }
    const int64_t tail = nvec << 2;
    if (tid < n - tail) {
      float raw[1] = {x[tail + tid]};
      if (kKind == 2) log_chunk<1>(ls, acc.lo, acc.hi, raw);
      else merge_one<kKind>(acc, raw[0]);
    }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) {
      float raw[1] = {x[i]};
      if (kKind == 2) log_chunk<1>(ls, acc.lo, acc.hi, raw);
      else merge_one<kKind>(acc, raw[0]);
    }
  }
  if (kKind == 2 && ls.count > 0) acc.m = Moments{(double)ls.count, ls.sum / (double)ls.count, 0.0};
  return acc;
}

__device__ __forceinline__ void store_partial(StatsWs* ws, unsigned int block, const Acc& acc) {
  double* p = ws->partials + (size_t)block * kPartialDoubles;
  p[0] = acc.m.n; p[1] = acc.m.mean; p[2] = acc.m.m2; p[3] = (double)acc.lo; p[4] = (double)acc.hi;
}

// The per-block records of `count` blocks combined by one block, by the same two sums as block_combine; fixed
// order, the result is valid in every thread.  The second sum needs the first's result: a thread keeps its first
// three records in registers (every grid the statistics kernel launches: <= 3 per thread), so the records make ONE
// trip from L2, and re-reads only what lies beyond.
__device__ __forceinline__ Acc combine_partials(const StatsWs* ws, int count, Acc* smem) {
  const double* parts = ws->partials;
  constexpr int kHeld = 3;
  Moments held[kHeld];
  double tn = 0.0, s1 = 0.0;
  float hi = -INFINITY, lo = INFINITY;
  int slot = 0;
  for (int b = threadIdx.x; b < count; b += blockDim.x, ++slot) {
    const double* p = parts + (size_t)b * kPartialDoubles;
    const Moments m{__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)};
#pragma unroll
    for (int k = 0; k < kHeld; ++k)
      if (slot == k) held[k] = m;
    tn += m.n;
    s1 += m.n == 0.0 ? 0.0 : m.n * m.mean;
    lo = nanmin(lo, (float)__ldcg(p + 3));
    hi = nanmax(hi, (float)__ldcg(p + 4));
  }
  block_sum2<true>(tn, s1, hi, lo, smem);
  const double mean = weighted_mean(tn, s1);
  double q = 0.0, unused = 0.0;
  slot = 0;
  for (int b = threadIdx.x; b < count; b += blockDim.x, ++slot) {
    Moments m;
    if (slot < kHeld) {
#pragma unroll
      for (int k = 0; k < kHeld; ++k)
        if (slot == k) m = held[k];
    } else {
      const double* p = parts + (size_t)b * kPartialDoubles;
      m = Moments{__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)};
    }
    q += m2_about(m, mean);
  }
  float h2 = 0.f, l2 = 0.f;
  block_sum2<false>(q, unused, h2, l2, smem);
  Acc f;
  f.m = Moments{tn, mean, q};
  f.hi = hi;
  f.lo = lo;
  return f;
}

// Final scalar step shared by the grid kernel and the small-tensor kernels.
template <int kKind>
__device__ __forceinline__ void finalize(const Acc& a, int unbiased, float* out) {
  if (kKind == 2) {
    out[0] = (float)a.m.mean;  // mu  (s2fp8.py:36)
    // m = max(L) (s2fp8.py:37).  log2 is monotone, so max over the non-zero elements is log2f(max|x|); an exact
    // zero contributes L = 0 (the reference's torch.where), which wins when every non-zero |x| is below 1
    float m = (a.hi == 0.0f) ? 0.0f : log2f(a.hi);   // hi = max|x| (NaN propagates)
    if (a.lo == 0.0f && m < 0.0f) m = 0.0f;          // lo = min|x|: a zero is present
    out[1] = m;
    return;
  }
  out[0] = (float)a.m.mean;
  if (kKind == 1) {
    // (max - min) * (1 / sqrt(2 ln n)), every step in fp32 as in smart.py:102-106
    float range = a.hi - a.lo;
    float c = 1.0f / sqrtf(2.0f * logf((float)a.m.n));
    out[1] = range * c;
  } else {
    double denom = unbiased ? (a.m.n - 1.0) : a.m.n;
    out[1] = (float)sqrt(a.m.m2 / denom);  // 0/0 -> NaN like torch for n == 1
  }
}

}  // namespace smaq
