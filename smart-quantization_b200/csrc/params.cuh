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
  uint64_t offset;
  PhiloxKeys keys;
};

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
  k.offset = p.offset;
  k.keys = make_philox_keys(p.seed);
  return k;
}

__device__ __forceinline__ Scalars scalars_from(float mean, float std_raw, const KernelParams& kp) {
  return make_scalars(mean, std_raw, kp.thr, kp.range_main, kp.range_out, kp.clamp_lo, kp.clamp_hi, kp.bits_main,
                      kp.bits_outlier);
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
