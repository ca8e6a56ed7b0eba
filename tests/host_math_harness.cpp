// TEST-ONLY: compiles the product's per-element arithmetic (csrc/smaq_math.cuh) for the host so
// the operation sequence can be checked against the oracle on a machine without a GPU.  Nothing
// in the product links or loads this file.
//   g++ -O2 -ffp-contract=off -mfma -shared -fPIC tests/host_math_harness.cpp -o <out>.so
#include <cstdint>
#include <cstring>

#include "../smart-quantization_b200/csrc/smaq_math.cuh"

using namespace smaq;

extern "C" {

// Round trip of n values with explicit probs, processed in pairs exactly as the kernels do;
// variant: 0 = IEEE divide everywhere, 1 = as the kernel decides (fast path + exact redo of flagged pairs).
void harness_roundtrip(const float* x, const float* probs, float* y, float* codes, int64_t n, float mean,
                       float std_raw, float thr, float range_main, float range_out, float clamp_lo, float clamp_hi,
                       int bits_main, int bits_outlier, int stochastic, int all_positive, int saturate, int variant,
                       int* used_fast) {
  Scalars s = make_scalars(mean, std_raw, thr, range_main, range_out, clamp_lo, clamp_hi, bits_main, bits_outlier);
  bool fast = variant == 1 && s.fast;
  int64_t redone = 0;
  for (int64_t i = 0; i < n; i += 2) {
    const bool has2 = i + 1 < n;
    f32x2 xv = pair(x[i], has2 ? x[i + 1] : x[i]);
    f32x2 pv = pair(probs ? probs[i] : 0.0f, probs ? (has2 ? probs[i + 1] : probs[i]) : 0.0f);
    PairClass k;
    bool suspect = !fast;
    f32x2 code = pair(0.f, 0.f), out = pair(0.f, 0.f);
    if (fast) {
      code = stochastic ? encode_pair<true, true>(xv, pv, s, k, suspect) : encode_pair<false, true>(xv, pv, s, k, suspect);
      if (saturate) code = pair(saturate_code(code.x, s, is_outlier0(k)), saturate_code(code.y, s, is_outlier1(k)));
      out = decode_pair<true, true>(code, k.shift, k.range_b, k.range_r, s, all_positive != 0, suspect);
    }
    if (suspect) {
      bool unused = false;
      redone += fast;
      code = stochastic ? encode_pair<true, false>(xv, pv, s, k, unused) : encode_pair<false, false>(xv, pv, s, k, unused);
      if (saturate) code = pair(saturate_code(code.x, s, is_outlier0(k)), saturate_code(code.y, s, is_outlier1(k)));
      out = decode_pair<false, false>(code, k.shift, k.range_b, k.range_r, s, all_positive != 0, unused);
    }
    y[i] = out.x;
    if (codes) codes[i] = code.x;
    if (has2) {
      y[i + 1] = out.y;
      if (codes) codes[i + 1] = code.y;
    }
  }
  if (used_fast) *used_fast = fast ? (int)(1 + (redone > n / 4)) : 0;  // 2: mostly redone (degenerate input)
}

// Exhaustive-style check of the three-instruction division against the IEEE divide.
// Returns the number of mismatching bit patterns (NaN == NaN).
int64_t harness_div_check(const float* a, int64_t n, float b, int tiny_guard) {
  Divisor d = make_divisor(b);
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    // the kernels' contract: div3 unless the guard fires, then the IEEE divide
    f32x2 q2 = div3(splat(a[i]), splat(d.b), splat(d.r));
    float q = q2.x;
    const bool redo = tiny_guard ? not_at_least(q, 9.094947017729282e-13f) : not_at_most(a[i], 1.2676506e30f);
    if (redo) q = true_div(a[i], d.b);
    float t = a[i] / b;
    uint32_t qb, tb;
    memcpy(&qb, &q, 4);
    memcpy(&tb, &t, 4);
    bool both_nan = (q != q) && (t != t);
    bool both_zero = (q == 0.0f) && (t == 0.0f);  // the sign of a zero quotient never reaches the output (Scalars::fast)
    if (qb != tb && !both_nan && !both_zero) ++bad;
  }
  return bad;
}

float harness_uniform24(uint32_t r) { return uniform24(r); }
}
