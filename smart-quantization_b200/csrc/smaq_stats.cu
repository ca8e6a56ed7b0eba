// Kernel (1): tensor statistics for SmaQ.
//
// Replaces data.mean() + data.std() (reference smart_compress/compress/smart.py:130-132, two
// eager reductions = two reads of the tensor) with ONE pass: every thread runs a chunked
// Welford update (16 values at a time: fp32 chunk mean and chunk M2, merged into fp64 running
// moments with Chan's formula), lanes are combined with warp shuffles, warps through shared
// memory, and the last block to finish (atomic ticket) combines the per-block moments in a fixed
// order — so the result is independent of block scheduling and reproducible run to run.
//
// Also here: the sampled statistics of --use_sample_stats (smart.py:86-91), the range estimate of
// --use_range_std_dev (smart.py:100-106) and the log2-domain statistics of S2FP8
// (smart_compress/compress/s2fp8.py:34-37).
//
// HBM roofline: 4 bytes read per element, nothing written.
#include "common.cuh"
#include "moments.cuh"
#include "smaq_math.cuh"

namespace smaq {

// resident CTAs per SM the statistics kernel is compiled for (<= 64 registers) and its grid is capped at
#ifndef SMAQ_STATS_CTAS_PER_SM
#define SMAQ_STATS_CTAS_PER_SM 4
#endif

template <int kKind, bool kAligned>
__global__ void __launch_bounds__(kStatsThreads, SMAQ_STATS_CTAS_PER_SM) stats_kernel(const float* __restrict__ x, int64_t n, int unbiased,
                                                              float* __restrict__ out, StatsWs* ws) {
  __shared__ Acc smem[kStatsThreads / 32];
  __shared__ bool is_last;
  // launched as a programmatic dependent of the tensor's producer (set_first_kernel_dependent): nothing is read
  // before that kernel has completed; behind an ordinary launch this returns at once
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // a kernel launched behind this one as a programmatic dependent (smaq_compress's round trip) may become resident
  // as soon as this grid's CTAs leave; it waits for this grid's completion before it reads mean/std
  asm volatile("griddepcontrol.launch_dependents;");
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  Acc acc = accumulate_tensor<kKind, kAligned>(x, n, tid, nthreads);

  acc = block_combine<kKind>(acc, smem);
  if (gridDim.x == 1) {  // one block: nothing to hand over
    if (threadIdx.x == 0) finalize<kKind>(acc, unbiased, out);
    return;
  }
  if (threadIdx.x == 0) {
    store_partial(ws, blockIdx.x, acc);
    __threadfence();
    unsigned int t = atomicAdd(&ws->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();

  // last block: the per-block moments by the same two sums (moments.cuh)
  const Acc f = combine_partials(ws, (int)gridDim.x, smem);
  if (threadIdx.x == 0) {
    finalize<kKind>(f, unbiased, out);
    ws->ticket = 0;  // leave the workspace reusable
  }
}

// --use_sample_stats: k gathered values, one block.  kKind 1: --use_range_std_dev as well — the reference's
// _get_sample_mean_std calls _get_std(sample), i.e. (max - min) / sqrt(2 ln k) over the k samples (smart.py:86-108).
template <int kKind>
__global__ void __launch_bounds__(kStatsThreads) sampled_stats_kernel(const float* __restrict__ x, int64_t n,
                                                                      const int64_t* __restrict__ idx, int k,
                                                                      float* __restrict__ out) {
  __shared__ Acc smem[kStatsThreads / 32];
  Acc acc;
  acc.m = Moments{0.0, 0.0, 0.0};
  acc.hi = -INFINITY;
  acc.lo = INFINITY;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    int64_t j = idx[i];
    if (j >= 0 && j < n) merge_one<kKind>(acc, x[j]);
  }
  acc = block_combine<kKind>(acc, smem);
  if (threadIdx.x == 0) finalize<kKind>(acc, /*unbiased=*/0, out);  // smart.py:91: unbiased=False
}

// Same, drawing the k distinct indices on the device with Floyd's algorithm (a uniform k-subset,
// which is the law of randperm(n)[:k]; mean and std do not depend on the order).
template <int kKind>
__global__ void __launch_bounds__(32) sampled_draw_stats_kernel(const float* __restrict__ x, int64_t n, int k,
                                                                uint64_t seed, uint64_t offset,
                                                                float* __restrict__ out) {
  __shared__ int64_t chosen[1024];
  const int lane = lane_id();
  const PhiloxKeys keys = make_philox_keys(seed);
  for (int i = 0; i < k; ++i) {
    const int64_t j = n - k + i;  // Floyd: t ~ U{0..j}; take t unless already chosen, else j
    uint4 r = philox_group(keys, (uint64_t)i, offset);
    uint64_t r64 = ((uint64_t)r.x << 32) | r.y;
    int64_t t = (int64_t)__umul64hi(r64, (uint64_t)j + 1);
    bool hit = false;
    for (int q = lane; q < i; q += 32) hit |= (chosen[q] == t);
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0) chosen[i] = hit ? j : t;
    __syncwarp();
  }
  Acc acc;
  acc.m = Moments{0.0, 0.0, 0.0};
  acc.hi = -INFINITY;
  acc.lo = INFINITY;
  for (int i = lane; i < k; i += 32) merge_one<kKind>(acc, x[chosen[i]]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Moments other = shfl_xor(acc.m, o);
    acc.m = (lane & o) ? merge(other, acc.m) : merge(acc.m, other);
    if (kKind == 1) {  // NaN-propagating like torch's max / min
      acc.hi = max_nan(acc.hi, __shfl_xor_sync(0xffffffffu, acc.hi, o));
      acc.lo = min_nan(acc.lo, __shfl_xor_sync(0xffffffffu, acc.lo, o));
    }
  }
  if (lane == 0) finalize<kKind>(acc, 0, out);
}

static int stats_grid(int64_t n) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
#ifndef SMAQ_STATS_EPT
#define SMAQ_STATS_EPT 16
#endif
  // one 16-element chunk per thread until the grid fills ONE resident wave (4 CTAs per SM at ~56 registers), then
  // threads loop.  Measured over 2^16..2^30 elements (tools/midsize_bench.py): a second wave only starts late, and
  // 128 elements per thread made a 256 KB tensor wait out eight dependent memory round trips in two CTAs
  int64_t want = (n + (int64_t)kStatsThreads * SMAQ_STATS_EPT - 1) / ((int64_t)kStatsThreads * SMAQ_STATS_EPT);
  int64_t cap = (int64_t)sms * SMAQ_STATS_CTAS_PER_SM;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

int stats_grid_for(int64_t n) { return stats_grid(n); }

template <int kKind>
static int launch_stats(const float* x, int64_t n, int unbiased, float* out, void* ws, size_t ws_bytes,
                        cudaStream_t stream, bool ticket_is_zero = false) {
  if (!x || !out || !ws || n <= 0) return fail(SMAQ_ERR_ARG, "stats: null pointer or n <= 0");
  if (ws_bytes < smaq_stats_workspace_bytes(n)) return fail(SMAQ_ERR_WORKSPACE, "stats: workspace too small");
  int grid = stats_grid(n);
  // the arrival ticket must start at zero; the kernel's last block leaves it at zero again, so a workspace
  // that was zeroed once (smaq_compress_workspace_init) needs no memset node per call
  if (!ticket_is_zero && grid > 1) SMAQ_CUDA_OK(cudaMemsetAsync(ws, 0, 16, stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStatsThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  set_first_kernel_dependent(cfg, attr);
  StatsWs* sws = (StatsWs*)ws;
  if (aligned16(x)) SMAQ_CUDA_OK(cudaLaunchKernelEx(&cfg, stats_kernel<kKind, true>, x, n, unbiased, out, sws));
  else SMAQ_CUDA_OK(cudaLaunchKernelEx(&cfg, stats_kernel<kKind, false>, x, n, unbiased, out, sws));
  return SMAQ_OK;
}

int stats_full_zeroed_ws(const float* x, int64_t n, int unbiased, float* mean_std, void* ws, size_t ws_bytes,
                         cudaStream_t stream) {
  return launch_stats<0>(x, n, unbiased, mean_std, ws, ws_bytes, stream, /*ticket_is_zero=*/true);
}

}  // namespace smaq

extern "C" {

size_t smaq_stats_workspace_bytes(int64_t n) {
  (void)n;
  int sms = smaq::sm_count();
  if (sms <= 0) sms = 148;
  return 16 + (size_t)sms * 8 * smaq::kPartialDoubles * sizeof(double);
}

int smaq_stats_full(const float* x, int64_t n, int unbiased, float* mean_std, void* ws, size_t ws_bytes,
                    smaq_stream_t stream) {
  return smaq::launch_stats<0>(x, n, unbiased, mean_std, ws, ws_bytes, (cudaStream_t)stream);
}

int smaq_stats_range(const float* x, int64_t n, float* mean_std, void* ws, size_t ws_bytes, smaq_stream_t stream) {
  return smaq::launch_stats<1>(x, n, 1, mean_std, ws, ws_bytes, (cudaStream_t)stream);
}

int smaq_s2fp8_stats(const float* x, int64_t n, float* mu_max, void* ws, size_t ws_bytes, smaq_stream_t stream) {
  return smaq::launch_stats<2>(x, n, 0, mu_max, ws, ws_bytes, (cudaStream_t)stream);
}

int smaq_stats_sampled(const float* x, int64_t n, const int64_t* idx, int32_t k, int32_t range_std, float* mean_std,
                       smaq_stream_t stream) {
  if (!x || !idx || !mean_std || n <= 0 || k <= 0) return smaq::fail(SMAQ_ERR_ARG, "stats_sampled: bad argument");
  if (range_std) smaq::sampled_stats_kernel<1><<<1, smaq::kStatsThreads, 0, (cudaStream_t)stream>>>(x, n, idx, k, mean_std);
  else smaq::sampled_stats_kernel<0><<<1, smaq::kStatsThreads, 0, (cudaStream_t)stream>>>(x, n, idx, k, mean_std);
  SMAQ_LAUNCH_OK();
  return SMAQ_OK;
}

int smaq_stats_sampled_draw(const float* x, int64_t n, int32_t k, int32_t range_std, uint64_t seed, uint64_t offset,
                            float* mean_std, smaq_stream_t stream) {
  if (!x || !mean_std || n <= 0 || k <= 0) return smaq::fail(SMAQ_ERR_ARG, "stats_sampled_draw: bad argument");
  if (k > 1024) return smaq::fail(SMAQ_ERR_UNSUPPORTED, "stats_sampled_draw: k > 1024");
  if (k > n) k = (int32_t)n;
  if (range_std) smaq::sampled_draw_stats_kernel<1><<<1, 32, 0, (cudaStream_t)stream>>>(x, n, k, seed, offset, mean_std);
  else smaq::sampled_draw_stats_kernel<0><<<1, 32, 0, (cudaStream_t)stream>>>(x, n, k, seed, offset, mean_std);
  SMAQ_LAUNCH_OK();
  return SMAQ_OK;
}

}  // extern "C"
