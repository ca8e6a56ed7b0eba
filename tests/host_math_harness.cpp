// TEST-ONLY: compiles the product's per-element arithmetic (csrc/smaq_math.cuh) for the host so
// the operation sequence can be checked against the oracle on a machine without a GPU.  Nothing
// in the product links or loads this file.
//   g++ -O2 -ffp-contract=off -mfma -shared -fPIC tests/host_math_harness.cpp -o <out>.so
#include <cstdint>
#include <cstring>

#include "../smart-quantization_b200/csrc/smaq_math.cuh"

using namespace smaq;

extern "C" {

// Round trip of n values with explicit probs; variant: 0 = IEEE divide everywhere, 1 = as the kernel decides.
void harness_roundtrip(const float* x, const float* probs, float* y, float* codes, int64_t n, float mean,
                       float std_raw, float thr, float range_main, float range_out, float clamp_lo, float clamp_hi,
                       int bits_main, int bits_outlier, int stochastic, int all_positive, int saturate, int variant,
                       int* used_fast) {
  Scalars s = make_scalars(mean, std_raw, thr, range_main, range_out, clamp_lo, clamp_hi, bits_main, bits_outlier);
  bool fast = variant == 1 && s.fast;
  if (used_fast) *used_fast = fast;
  for (int64_t i = 0; i < n; ++i) {
    Classified k;
    float p = probs ? probs[i] : 0.0f;
    float code;
    if (fast) code = stochastic ? encode_value<true, true>(x[i], s, p, k) : encode_value<false, true>(x[i], s, p, k);
    else code = stochastic ? encode_value<true, false>(x[i], s, p, k) : encode_value<false, false>(x[i], s, p, k);
    if (saturate) code = saturate_code(code, s, k.hi || k.lo);
    if (codes) codes[i] = code;
    y[i] = fast ? decode_value<true>(code, k.shift, k.range, s, all_positive)
                : decode_value<false>(code, k.shift, k.range, s, all_positive);
  }
}

// Exhaustive-style check of the three-instruction division against the IEEE divide.
// Returns the number of mismatching bit patterns (NaN == NaN).
int64_t harness_div_check(const float* a, int64_t n, float b, int tiny_guard) {
  Divisor d = make_divisor(b);
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    float q = tiny_guard ? div_rn<true>(a[i], d) : div_rn<false>(a[i], d);
    float t = a[i] / b;
    uint32_t qb, tb;
    memcpy(&qb, &q, 4);
    memcpy(&tb, &t, 4);
    bool both_nan = (q != q) && (t != t);
    bool both_zero = (q == 0.0f) && (t == 0.0f);  // sign of zero is handled by the caller's contract
    if (qb != tb && !both_nan && !both_zero) ++bad;
  }
  return bad;
}

float harness_uniform24(uint32_t r) { return uniform24(r); }
}
