/*
 * smaq_b200 — C ABI of the B200-native SmaQ / FP8 / S2FP8 compress->decompress path.
 *
 * The reference (nimashoghi/smart-quantization) has no FFI: its hot path is ~30 eager
 * torch ops per call inside three Python plugin classes.  The entry points below are what
 * a binding for that path replaces; each cites the reference lines it stands in for.
 * The Python host side (smart-quantization_b200/smart_compress/_native.py) binds them with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named x/y/packed/ws/mean_std/... is a DEVICE pointer unless it says "host";
 *   - the caller owns every buffer (outputs and workspace included); nothing here allocates;
 *   - nothing here synchronises: work is enqueued on `stream` (a cudaStream_t) and returns;
 *   - no global mutable state: safe to call from several threads (forward on the Python main
 *     thread, backward on autograd's per-device worker thread);
 *   - return value: SMAQ_OK or a SMAQ_ERR_* code; smaq_b200_last_error() gives the text of the
 *     calling thread's last failure.  No exception crosses this boundary.
 *   - float data is IEEE binary32; element counts are int64_t (4 GiB+ tensors are in scope).
 */
#ifndef SMAQ_B200_H
#define SMAQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMAQ_B200_ABI_VERSION 4 /* 2: packed stream SQB2, count_saturated, smaq_compress, tensor_desc.stream; 3: smaq_float_quantize_multi;
                                    4: smaq_roundtrip_bn, smaq_roundtrip_multi(mean_std_out), range_std in the sampled statistics,
                                       packed stream SQB3 (one-pass encoder, smaq_encode_workspace_init), offset_base + smaq_counter_add, zero_on_grid, smaq_decode_sum, smaq_s2fp8_multi,
                                       split buffers + smaq_extras_compact */

typedef void* smaq_stream_t; /* cudaStream_t */

enum {
  SMAQ_OK = 0,
  SMAQ_ERR_ARG = 1,         /* null pointer, negative size, unsupported bit width ... */
  SMAQ_ERR_CUDA = 2,        /* a CUDA runtime call failed (launch error included) */
  SMAQ_ERR_WORKSPACE = 3,   /* caller's workspace / packed buffer is too small */
  SMAQ_ERR_UNSUPPORTED = 4  /* valid request the library does not implement */
};

int smaq_b200_abi_version(void);
const char* smaq_b200_last_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of this); <0 on error. */
int smaq_b200_sm_count(void);

/* ---- SmaQ ------------------------------------------------------------------------------- */

/* The per-call constants SmartFP derives from its flags (smart.py:72-84) plus the call's kwargs
 * (smart.py:111-118).  Plain data, read on the host at call time. */
typedef struct smaq_codec_params {
  float threshold;      /* fp32(--main_std_dev_threshold), > 0 */
  float range_main;     /* fp32(range_normal)   smart.py:78-80 */
  float range_outlier;  /* fp32(range_outlier)  smart.py:75-77 */
  float clamp_lo;       /* clamped_range        smart.py:82-84 */
  float clamp_hi;
  int32_t bits_main;    /* --num_bits_main    (2..32; the packed stream: 4..8) */
  int32_t bits_outlier; /* --num_bits_outlier (2..32; the packed stream: bits_main + 0..4) */
  int32_t stochastic;   /* 0: trunc (smart.py:169); 1: stochastic rounding (smart.py:93-98) */
  int32_t all_positive; /* clamp_min(0) at the end (smart.py:181-182) */
  int32_t saturate;     /* round trip only: clamp codes to what the packed format holds (not in
                           the reference; 0 reproduces smart.py exactly) */
  int32_t count_saturated; /* packed encoder only: fill header.n_saturated (costs ~5 % of the encode
                           kernel; the plugin sets it under --measure_compression_ratio).  0: the
                           header field is all ones */
  int32_t zero_on_grid; /* packed encoder only (not in the reference): move the mean the codes are relative to by at
                           most half a quantisation step so that x == 0.0 lies exactly on the grid and decodes to
                           exactly 0.0 (smaq_packed_header.mean holds the mean used).  For saved ReLU outputs, whose
                           backward mask is `output > 0`.  Applies when |0 - mean| <= outlier range; else no-op */
  uint64_t seed;        /* Philox4x32-7 key, used when stochastic && probs == NULL */
  uint64_t offset;      /* Philox stream offset (added to the counter's high words) */
  const uint64_t* offset_base; /* optional DEVICE counter: the stream is offset + *offset_base, read by the kernel.
                           What lets a training step be captured in a CUDA graph: the per-call offsets are baked into
                           the graph, the base advances on the device between replays (smaq_counter_add). NULL: 0 */
} smaq_codec_params;

/* *counter += delta on `stream` (one thread).  Captured at the end of a graphed training step with delta = the
 * number of random streams the step used, so every replay draws fresh numbers. */
int smaq_counter_add(uint64_t* counter, uint64_t delta, smaq_stream_t stream);

/* Full-tensor mean and standard deviation in one pass — replaces data.mean() and data.std()
 * (smart.py:130-132; two read passes there).  Welford chunks per thread, warp-shuffle and
 * block combine, last-block grid combine in fp64.  Writes mean_std[0]=mean, mean_std[1]=std
 * (std with divisor n-1 when unbiased!=0, n otherwise).  ws: smaq_stats_workspace_bytes(n). */
size_t smaq_stats_workspace_bytes(int64_t n);
int smaq_stats_full(const float* x, int64_t n, int unbiased, float* mean_std, void* ws, size_t ws_bytes,
                    smaq_stream_t stream);

/* --use_range_std_dev (smart.py:100-106): mean as above, std = (max-min)/sqrt(2 ln n). */
int smaq_stats_range(const float* x, int64_t n, float* mean_std, void* ws, size_t ws_bytes,
                     smaq_stream_t stream);

/* --use_sample_stats (smart.py:86-91): mean and BIASED std of x[idx[0..k)].  idx is a device
 * array of int64 indices (the caller's permutation prefix).  range_std != 0: --use_range_std_dev as well, i.e.
 * std = (max - min) / sqrt(2 ln k) over the k samples (smart.py:91 calls _get_std on the sample, :100-106). */
int smaq_stats_sampled(const float* x, int64_t n, const int64_t* idx, int32_t k, int32_t range_std, float* mean_std,
                       smaq_stream_t stream);
/* Same, but the k distinct indices are drawn on the device (Philox + Floyd's algorithm): a
 * uniform k-subset, the law of randperm(n)[:k], without materialising an n-element permutation. */
int smaq_stats_sampled_draw(const float* x, int64_t n, int32_t k, int32_t range_std, uint64_t seed, uint64_t offset,
                            float* mean_std, smaq_stream_t stream);

/* Fused fake-quantisation round trip — replaces smart.py:151-182 (~26 elementwise kernels, a
 * host sync and two H2D scalar copies) with one read and one write of the tensor.
 * mean_std: device float[2] from a stats call (raw std; the ==0 fix-up happens on device).
 * probs: optional device float[n] of U[0,1) numbers (parity mode); NULL -> in-kernel Philox.
 * y may alias x. */
int smaq_roundtrip(const float* x, float* y, int64_t n, const float* mean_std, const float* probs,
                   const smaq_codec_params* params, smaq_stream_t stream);

/* --use_batch_norm (smart.py:121,136-149,174-179; the hook passes the producing BatchNorm2d's weight and bias,
 * util/pytorch/autograd.py:64-72): x is an NCHW feature map, element i belongs to channel (i / inner) % channels
 * (inner = H * W).  Per element: x' = (x - beta_c) / gamma_c, the round trip of x' under mean_std (statistics of
 * x itself: the reference takes them BEFORE the un-affine), y = y' * gamma_c + beta_c, then clamp_min(0) if
 * params->all_positive.  --bn_scalar_params: pass the two means as 1-element arrays with channels = 1.
 * Replaces four permute+clone copies and four elementwise passes around the chain.  y may alias x. */
int smaq_roundtrip_bn(const float* x, float* y, int64_t n, const float* mean_std, const float* probs, const float* gamma,
                      const float* beta, int64_t channels, int64_t inner, const smaq_codec_params* params,
                      smaq_stream_t stream);

/* The whole default call in one entry point — full-tensor unbiased statistics, then the round trip
 * (smart.py:130-182 with the reference's default flags): what the training hooks issue hundreds of
 * times per step.  Two kernels on `stream`: the statistics kernel and, as a programmatic dependent launch that
 * overlaps its tail (up to 2^27 elements; SMAQ_DEPENDENT_LAUNCH=0 in the environment: an ordinary launch), the
 * round trip; bit-identical to smaq_stats_full followed by smaq_roundtrip.  The statistics kernel itself (and the
 * kernel of smaq_float_quantize) is launched as a programmatic dependent of whatever precedes it on `stream`: it
 * executes griddepcontrol.wait before touching memory, so any producer is safe, and the launch gap behind the
 * producer disappears (SMAQ_DEPENDENT_LAUNCH=1: only the launches inside a call are dependent).  The statistics live in the workspace
 * (last 256 bytes: mean, std).  y may alias x.
 * The workspace is initialised ONCE after allocation with smaq_compress_workspace_init (it zeroes the arrival
 * ticket of the statistics pass; every call leaves it zero again, so no memset node is issued per call).  Calls
 * sharing a workspace must be ordered on one stream.  An uninitialised workspace gives undefined statistics. */
size_t smaq_compress_workspace_bytes(int64_t n);
int smaq_compress_workspace_init(void* ws, size_t ws_bytes, smaq_stream_t stream);
int smaq_compress(const float* x, float* y, int64_t n, const float* probs, const smaq_codec_params* params,
                  void* ws, size_t ws_bytes, smaq_stream_t stream);

/* Number of outliers (|z| > threshold) under the given statistics — the only quantity the
 * reference's size accounting needs (smart.py:184-187: bits = 8*n_out + 6*(n - n_out)).  Adds the
 * count to *counter (device uint64, caller zeroes it).  Only used under --measure_compression_ratio. */
int smaq_count_outliers(const float* x, int64_t n, const float* mean_std, const smaq_codec_params* params,
                        unsigned long long* counter, smaq_stream_t stream);

/* Small-tensor path: statistics (full-tensor, unbiased) + round trip in ONE launch of one
 * thread block; n <= smaq_fused_small_max(). */
int64_t smaq_fused_small_max(void);
int smaq_roundtrip_small(const float* x, float* y, int64_t n, const float* probs,
                         const smaq_codec_params* params, float* mean_std_out /* may be NULL */,
                         smaq_stream_t stream);

/* Many tensors, one launch (the optimizer side: smart_compress/util/pytorch/optimizer.py:69-127
 * loops compress_fn over every parameter).  descs is a DEVICE array of `count` descriptors;
 * each tensor gets its own full-tensor statistics and its own round trip, in place or not. */
typedef struct smaq_tensor_desc {
  const float* x;
  float* y;
  int64_t n;
  int32_t all_positive;
  int32_t stream;       /* Philox stream of this tensor = params->offset + stream */
} smaq_tensor_desc;
size_t smaq_multi_workspace_bytes(int32_t count, int64_t total_elems);
/* mean_std_out: optional device float[count][2]; entry i receives the (mean, std) tensor i was quantised with
 * (untouched for tensors below min_size), so a caller — the parity tests — can check the codes against the
 * reference given those statistics. */
int smaq_roundtrip_multi(const smaq_tensor_desc* descs, int32_t count, int64_t max_n, int64_t total_elems,
                         const smaq_codec_params* params, int64_t min_size, void* ws, size_t ws_bytes,
                         float* mean_std_out, smaq_stream_t stream);

/* ---- packed SmaQ stream ------------------------------------------------------------------ */

/* Byte offsets of the sections of one packed tensor (stream "SQB3", DESIGN.md "Packed stream"). */
typedef struct smaq_packed_layout {
  int64_t n;
  int32_t bits_main, bits_outlier;
  int64_t n_warp_tiles;    /* ceil(n / 1024) */
  int64_t n_cta_tiles;     /* ceil(n_warp_tiles / 8) */
  int64_t header_off, header_bytes;   /* smaq_packed_header */
  int64_t planes_off, planes_bytes;   /* tag word + (bits_main-1) base words per lane, per warp tile */
  int64_t extras_off;                 /* (bits_outlier-bits_main) bits per outlier, dense inside a warp tile */
  int64_t extras_stride_bytes;        /* warp tile t's segment starts at extras_off + t * extras_stride_bytes */
  int64_t extras_capacity_bytes;
  int64_t total_capacity_bytes;       /* what the caller must allocate */
  int64_t workspace_bytes;            /* scratch for smaq_encode (64 bytes) */
} smaq_packed_layout;

/* First bytes of a packed buffer, written by smaq_encode on the device. */
typedef struct smaq_packed_header {
  uint32_t magic;            /* 'SQB3' */
  int32_t bits_main, bits_outlier;
  int32_t stochastic;
  int64_t n;
  float mean, std_raw;       /* statistics the codes are relative to */
  float threshold, range_main, range_outlier, clamp_lo, clamp_hi, pad0;
  uint64_t n_outlier;        /* smart.py:184-187: compressed bits = 8*n_outlier + 6*(n-n_outlier) */
  uint64_t n_saturated;      /* scaled values the field width cannot hold, or NaN (H1); ~0 when
                                not counted (smaq_codec_params.count_saturated == 0) */
  uint64_t extras_words;     /* 32-bit words actually used in the extras section (sum over the warp tiles) */
  uint64_t status;           /* 0 ok (reserved for failure codes) */
} smaq_packed_header;

int smaq_packed_layout_for(int64_t n, int32_t bits_main, int32_t bits_outlier, smaq_packed_layout* out);

/* Quantise and pack — the materialised form of the code the reference only ever holds as an
 * fp32 value (smart.py:164-169).  Codes saturate at the field width.  ONE kernel, one read of x.
 * ws: layout.workspace_bytes of scratch (an arrival ticket and three counters), zeroed ONCE after allocation
 * with smaq_encode_workspace_init; every call leaves it zero again, so no memset node is issued per call.
 * Calls sharing a workspace must be ordered on one stream. */
int smaq_encode_workspace_init(void* ws, size_t ws_bytes, smaq_stream_t stream);
int smaq_encode(const float* x, int64_t n, const float* mean_std, const float* probs,
                const smaq_codec_params* params, void* packed, size_t packed_bytes, void* ws,
                size_t ws_bytes, smaq_stream_t stream);

/* Unpack and de-normalise (smart.py:171-172,181-182).  Parameters come from the header. */
int smaq_decode(const void* packed, size_t packed_bytes, int64_t n, int32_t bits_main, int32_t bits_outlier,
                int32_t all_positive, float* y, smaq_stream_t stream);

/* The same with the stream in TWO buffers: `head` = header + planes (layout.extras_off bytes, exact for n) and
 * `extras` (layout.extras_capacity_bytes for smaq_encode_split).  A consumer that KEEPS the stream (saved
 * activations) compacts the extras afterwards and frees the capacity-sized buffer:
 *   smaq_extras_compact: seg_table[t] = first word of warp tile t's segment in the compacted extras, t = 0 ..
 *                        n_warp_tiles (the last entry is the total = header.extras_words), computed from the tag
 *                        words, and every tile's used words copied from the fixed-stride buffer to
 *                        extras_dst[seg_table[t] ...].  seg_table holds smaq_extras_table_entries(n) uint32 (table +
 *                        scratch).  A destination too small for the stream is reported in *overflow (device,
 *                        optional: the largest word index needed) and nothing is written past it.  Two launches;
 *   smaq_decode_split  : seg_table == NULL -> fixed-stride extras; else the compacted form.
 * Stored bits then: bits_main * n_main + bits_outlier * n_outlier (+ < 32 pad bits and 32 table bits per 1024
 * elements) — the reference's own accounting (smart.py:184-187) delivered in memory, not only logged. */
int smaq_encode_split(const float* x, int64_t n, const float* mean_std, const float* probs,
                      const smaq_codec_params* params, void* head, size_t head_bytes, void* extras, size_t extras_bytes,
                      void* ws, size_t ws_bytes, smaq_stream_t stream);
int smaq_decode_split(const void* head, size_t head_bytes, const void* extras, size_t extras_bytes,
                      const uint32_t* seg_table, int64_t n, int32_t bits_main, int32_t bits_outlier, int32_t all_positive,
                      float* y, smaq_stream_t stream);
int64_t smaq_extras_table_entries(int64_t n);
int smaq_extras_compact(const void* head, size_t head_bytes, const void* extras_src, int64_t n, int32_t bits_main,
                        int32_t bits_outlier, uint32_t* seg_table, void* extras_dst, size_t dst_bytes,
                        unsigned long long* overflow, smaq_stream_t stream);

/* The reduce step of a COMPRESSED all-reduce (not in the reference, which compresses after DDP's
 * fp32 all-reduce, optimizer.py:135-141 — so this is opt-in and changes numerics).  `packed` is a HOST array of
 * `count` (<= 8) device pointers to packed tensors of identical geometry (n, widths), typically one per rank and
 * living in that rank's memory, mapped into this process (symmetric memory / CUDA IPC over NVLink).  For the CTA
 * tiles [first_cta_tile, first_cta_tile + n_cta_tiles) — 8192 elements each — it writes
 *     y[i] = scale * (decode(packed[0])[i] + decode(packed[1])[i] + ...)
 * y is indexed like the whole tensor (pass the tensor's base pointer).  One kernel: its loads are the transfer. */
int smaq_decode_sum(const void* const* packed, int32_t count, size_t packed_bytes, int64_t n, int32_t bits_main,
                    int32_t bits_outlier, int64_t first_cta_tile, int64_t n_cta_tiles, float scale, float* y,
                    smaq_stream_t stream);

/* ---- low-precision float emulation --------------------------------------------------------- */

/* What qtorch 0.2.0's float_quantize does to each fp32 (the reference calls it at
 * smart_compress/util/pytorch/quantization.py:191-193), fused with the reference's own
 * "+max -> +inf" fix-up (quantization.py:195-199).
 * rounding: 0 nearest, 1 stochastic.  rand_bits: optional device int32[n] (what qtorch's CUDA
 * path draws with randint_like); NULL -> in-kernel Philox.  y may alias x. */
typedef struct smaq_floatq_params {
  int32_t exp_bits, man_bits;
  int32_t rounding;
  int32_t check_inf;        /* hparams.float_quantize_check_inf */
  int32_t max_exp_bias;     /* max stored exponent = 127 + 2^(exp_bits-1) + max_exp_bias; 0 for qtorch 0.2.0 */
  int32_t reserved;
  uint64_t seed, offset;
  const uint64_t* offset_base; /* as in smaq_codec_params */
} smaq_floatq_params;
int smaq_float_quantize(const float* x, float* y, int64_t n, const int32_t* rand_bits,
                        const smaq_floatq_params* params, smaq_stream_t stream);

/* float_quantize over MANY tensors in two launches — the loop OptimLP runs over every parameter, gradient and
 * state tensor (smart_compress/util/pytorch/optimizer.py:69-127) with --compress fp8 | fp16 | bf16.  descs is a
 * DEVICE array of `count` descriptors (x, y, n; all_positive is ignored); tensor i is rounded with the Philox
 * stream params->offset + descs[i].stream, i.e. exactly as smaq_float_quantize would round it with that offset.
 * In-kernel random numbers only (no rand_bits).  total_elems = sum of descs[i].n. */
size_t smaq_floatq_multi_workspace_bytes(int32_t count);
int smaq_float_quantize_multi(const smaq_tensor_desc* descs, int32_t count, int64_t total_elems,
                              const smaq_floatq_params* params, void* ws, size_t ws_bytes, smaq_stream_t stream);

/* S2FP8 (smart_compress/compress/s2fp8.py:31-48).  Pass 1: mu = mean(L), m = max(L) with
 * L = log2|x| and L := 0 where x == 0; writes mu_max[0..1].  Pass 2: alpha = 15/(m-mu),
 * beta = -alpha*mu, y = sign(x) * (Q(|x|^alpha * 2^beta) / 2^beta)^(1/alpha), Q = e5m2 above. */
int smaq_s2fp8_stats(const float* x, int64_t n, float* mu_max, void* ws, size_t ws_bytes,
                     smaq_stream_t stream);
int smaq_s2fp8_apply(const float* x, float* y, int64_t n, const float* mu_max, const int32_t* rand_bits,
                     const smaq_floatq_params* params, smaq_stream_t stream);

/* S2FP8 over MANY tensors in three launches (set-up, per-item log-domain moments, apply) — the loops OptimLP runs
 * with --compress s2fp8 (optimizer.py:69-127).  descs as in smaq_float_quantize_multi (all_positive ignored);
 * tensor i is rounded with the Philox stream params->offset + descs[i].stream.  mu_max_out: optional device
 * float[count][2] receiving each tensor's (mu, m).  In-kernel random numbers only. */
size_t smaq_s2fp8_multi_workspace_bytes(int32_t count, int64_t total_elems);
int smaq_s2fp8_multi(const smaq_tensor_desc* descs, int32_t count, int64_t total_elems, const smaq_floatq_params* params,
                     void* ws, size_t ws_bytes, float* mu_max_out, smaq_stream_t stream);

/* Test hook, not part of the reference's surface: out[i] = a[i] ** y[0] evaluated by the S2FP8 apply kernel's
 * packed fast path for every group of four elements it accepts (accepted[i / 4] = 1), by powf otherwise.  The
 * parity tests compare it bit for bit with torch.pow on the same GPU over all magnitudes: the quantised output
 * of smaq_s2fp8_apply alone would hide a one-ulp difference with probability 1 - 2^-21 per element. */
int smaq_selftest_pow(const float* a, const float* y, float* out, int32_t* accepted, int64_t n, smaq_stream_t stream);

/* Test hook for the S2FP8 apply kernel's screened path (an approximate |x|^alpha decides every element whose
 * rounding it cannot change; the rest are recomputed exactly).  Measures on the device, into six floats:
 *   out[0] max |lg2.approx(a) - log2 a| over EVERY normal a in [0.5, 2];  out[1] max relative error of lg2.approx
 *   over every other normal a;  out[2] max relative error of ex2.approx over EVERY t in [-126, 128);
 *   out[3] max |bits(v~) - bits(v)| / margin over `samples` random (alpha, beta, a) — must stay below 1;
 *   out[4] the largest margin met;  out[5] the number of samples that qualified.
 * The kernel's error constants are asserted against out[0..2] by the tests. */
int smaq_selftest_s2_screen(float* out6, int64_t samples, smaq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SMAQ_B200_H */
