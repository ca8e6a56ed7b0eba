// Per-call parameters as the kernels see them (by value in the launch), shared by the round-trip
// and the packed encode/decode kernels.
#pragma once
#include "common.cuh"
#include "smaq_math.cuh"

namespace smaq {

struct KernelParams {
  float thr, range_main, range_out, clamp_lo, clamp_hi;
  int bits_main, bits_outlier;
  int all_positive, saturate;
  int zero_on_grid;
  uint64_t offset;
  const unsigned long long* offset_base;  // optional device counter added to `offset` (CUDA-graph replays)
  PhiloxKeys keys;
};

// The stream offset a kernel works with: the launch parameter plus the device counter.  The two hot kernels (packed
// encoder, fused round trip) take it as a separate value and keep reading everything else — the Philox keys above
// all — straight from the launch parameters (constant-bank operands); copying the struct cost them 1-2 %.
__device__ __forceinline__ uint64_t resolved_offset(const KernelParams& kp) {
  return kp.offset + (kp.offset_base ? __ldg(kp.offset_base) : 0ull);
}

// The parameters a kernel works with (everything that is not bandwidth-critical): a copy of the launch parameters whose stream offset includes the device
// counter (the copy's other fields stay what they are — loads from the constant bank)
__device__ __forceinline__ KernelParams resolved(const KernelParams& kp) {
  KernelParams k = kp;
  if (kp.offset_base) k.offset += __ldg(kp.offset_base);
  return k;
}

inline KernelParams to_kernel_params(const smaq_codec_params& p) {
  KernelParams k;
  k.thr = p.threshold;
  k.range_main = p.range_main;
  k.range_out = p.range_outlier;
  k.clamp_lo = p.clamp_lo;
  k.clamp_hi = p.clamp_hi;
  k.bits_main = p.bits_main;
  k.bits_outlier = p.bits_outlier;
  k.all_positive = p.all_positive;
  k.saturate = p.saturate;
  k.zero_on_grid = p.zero_on_grid;
  k.offset = p.offset;
  k.offset_base = (const unsigned long long*)p.offset_base;
  k.keys = make_philox_keys(p.seed);
  return k;
}

__device__ __forceinline__ Scalars scalars_from(float mean, float std_raw, const KernelParams& kp) {
  return make_scalars(mean, std_raw, kp.thr, kp.range_main, kp.range_out, kp.clamp_lo, kp.clamp_hi, kp.bits_main,
                      kp.bits_outlier);
}

// zero_on_grid: the mean m' closest to `mean` for which x == 0 decodes to exactly 0.  With c0 the integer code
// nearest to zero's scaled z-score, the decoder computes fl(fl(fl(c0 / range) - shift) * std) + m' for it
// (smart.py:171-172), so m' is minus that product: the final addition cancels exactly.  Zero keeps its code with
// probability ~1 - 1e-6 under stochastic rounding (its scaled value is an integer up to rounding noise).
__device__ __forceinline__ float snap_mean_to_zero(float mean, float std_raw, const KernelParams& kp) {
  const Scalars s = scalars_from(mean, std_raw, kp);
  const float z0 = true_div(sub_rn(0.0f, mean), s.div.b);
  if (!(fabsf(z0) <= 1e30f) || !(s.std_mul > 0.0f) || !(fabsf(s.std_mul) <= 1e30f)) return mean;
  const bool hi = z0 > s.thr, lo = z0 < s.neg_thr;
  const float shift = hi ? s.shift_hi : (lo ? s.shift_lo : s.shift_mid);
  const float range = (hi || lo) ? s.range_out.b : s.range_main.b;
  const float lim = (hi || lo) ? s.lim_out : s.lim_main;
  const float c0 = rintf(mul_rn(add_rn(z0, shift), range));
  if (!(fabsf(c0) <= lim)) return mean;   // zero is beyond what a code can hold: nothing to align
  return -mul_rn(sub_rn(true_div(c0, range), shift), s.std_mul);
}

inline int check_params(const smaq_codec_params* p) {
  if (!p) return fail(SMAQ_ERR_ARG, "params is NULL");
  if (!(p->threshold > 0.0f)) return fail(SMAQ_ERR_ARG, "threshold must be > 0");
  // the fake-quantisation round trip needs the widths only for the optional saturation limit 2^(bits-2)-1 (the
  // reference never checks them; its ranges are derived on the host): anything that limit can be formed for is
  // accepted here.  The packed encoder has its own, narrower, list (smaq_packed_layout_for).
  if (p->bits_main < 2 || p->bits_main > 32 || p->bits_outlier < 2 || p->bits_outlier > 32)
    return fail(SMAQ_ERR_ARG, "unsupported bit widths main=%d outlier=%d (2..32)", p->bits_main, p->bits_outlier);
  return SMAQ_OK;
}

}  // namespace smaq
