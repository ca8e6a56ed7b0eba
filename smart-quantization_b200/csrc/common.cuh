// Shared device/host helpers for the smaq_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/smaq_b200.h"

namespace smaq {

// ---- error plumbing (thread-local; no global mutable state shared between threads) ----------
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int sm_count();  // cached per device, immutable once read

// Programmatic dependent launch (sm_90+): a kernel that consumes the result of the kernel launched just before it
// on the same stream is launched with this attribute; its CTAs become resident while the first grid drains — the
// first kernel executes griddepcontrol.launch_dependents — and block in griddepcontrol.wait until that grid has
// completed and its writes are visible.  Launch latency, CTA set-up and whatever the second kernel can do without
// the first one's result overlap the first one's tail.  Used for statistics -> round trip (smaq_compress),
// encode pass 1 -> pass 2 and S2FP8 statistics -> apply.  Development switch, read once:
// SMAQ_DEPENDENT_LAUNCH=0 makes them ordinary launches (tools/compress_ab.sh).
bool dependent_launch_enabled();
void set_dependent_launch(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr);
// The FIRST kernel of a codec call as a programmatic dependent of whatever kernel precedes it in the stream (level 2,
// the default; SMAQ_DEPENDENT_LAUNCH=1 keeps only the launches inside a call).  Such a kernel executes
// griddepcontrol.wait before it touches global memory, so it is correct behind any producer; what it gains is that
// its launch no longer waits for the producer's completion + memory flush to be processed by the front end.
void set_first_kernel_dependent(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr);

#define SMAQ_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return ::smaq::fail(SMAQ_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                          \
  } while (0)

#define SMAQ_LAUNCH_OK()                                                                \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess)                                                              \
      return ::smaq::fail(SMAQ_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                          \
  } while (0)

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- streaming 128-bit global accesses -------------------------------------------------------
// Every tensor on this path is touched once per kernel: keep it out of L1.  Plain (coherent)
// loads, not .nc: the round trip may run in place (y aliasing x).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
// 256-bit accesses (sm_100): eight consecutive fp32 per lane in one instruction; needs 32-byte
// alignment.  The load also marks the line evict-first in L2 (the qualifier exists for .v8 only).
struct f32x8 {
  float4 a, b;
};
__device__ __forceinline__ f32x8 ldg_stream8(const float* p) {
  f32x8 v;
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream8(float* p, const f32x8& v) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y),
               "f"(v.a.z), "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
               : "memory");
}
__host__ __device__ inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ---- Philox4x32-7 (Salmon et al., SC'11): counter-based, so the random number of element i
// depends only on (seed, offset, i) and never on the launch geometry.  The ten round keys are
// derived from the seed on the host and travel in the kernel parameters: in SASS they are
// constant-bank operands of the LOP3s, not per-round additions. -------------------------------
// Seven rounds: the fewest for which Salmon et al. report Philox4x32 passing BigCrush (their
// default of ten is a safety margin).  The numbers only decide stochastic-rounding ties; on these
// issue-bound kernels three rounds are 4-5 % of the run time (profiles/).
#ifndef SMAQ_PHILOX_ROUNDS
#define SMAQ_PHILOX_ROUNDS 7
#endif
constexpr int kPhiloxRounds = SMAQ_PHILOX_ROUNDS;
struct PhiloxKeys {
  uint32_t k0[kPhiloxRounds], k1[kPhiloxRounds];
};
__host__ __device__ __forceinline__ PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < kPhiloxRounds; ++r) {
    K.k0[r] = a;
    K.k1[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return K;
}
__host__ __device__ __forceinline__ void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}
// counter = (c0,c1,c2,c3); returns 4 x 32 random bits
__host__ __device__ __forceinline__ uint4 philox4x32(const PhiloxKeys& K, uint32_t c0, uint32_t c1, uint32_t c2,
                                                     uint32_t c3) {
#pragma unroll
  for (int r = 0; r < kPhiloxRounds; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo32(0xD2511F53u, c0, h0, l0);
    mulhilo32(0xCD9E8D57u, c2, h1, l1);
    const uint32_t n0 = h1 ^ c1 ^ K.k0[r], n2 = h0 ^ c3 ^ K.k1[r];
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
  }
  uint4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}
// the 4 random words for elements [4*g, 4*g+4) of a tensor, stream `offset`
__host__ __device__ __forceinline__ uint4 philox_group(const PhiloxKeys& K, uint64_t g, uint64_t offset) {
  return philox4x32(K, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)offset, (uint32_t)(offset >> 32));
}
__host__ __device__ __forceinline__ uint32_t philox_word(const uint4& q, int j) {
  return j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w;
}

// ---- SmaQ's in-kernel uniforms: 16 random bits per element, ONE Philox4x32-7 call per 16 elements -------------
// Elements are taken in groups of eight (group g = elements 8g .. 8g+7: one 256-bit access, one "chunk" of the
// packed encoder).  Groups g and g ^ 32 share a call — in the packed encoder's warp tile these are chunks k and k+1
// of the SAME lane, so a lane draws twice per 32 elements instead of four times (the Philox calls were 19 % of the
// encoder's time).  Word q of the call serves elements 2q and 2q+1 of both groups; a 32-bit word
// b3 b2 b1 b0 holds four 16-bit windows at byte stride, (b1 b0), (b2 b1), (b3 b2), (b0 b3):
//     group with bit 5 clear: element 2q -> (b1 b0), element 2q+1 -> (b3 b2)      [the two disjoint halves]
//     group with bit 5 set  : element 2q -> (b2 b1), element 2q+1 -> (b0 b3)      [the two windows in between]
// Inside a group the eight values are disjoint bit fields of the call (independent); each value is exactly
// uniform on 0..65535; a value of the second group shares one byte with each of two values of the first group,
// 256 elements away (its low byte is the high byte of one of them: a dependence of order 2^-8 between two
// rounding decisions).  The number depends on (seed, stream offset, element index) only, never on the launch
// geometry, and every kernel (round trip, small / multi-tensor, packed encoder) draws the same one for the same
// element.  oracle/rng.py restates the scheme in numpy; the parity tests feed its numbers to the CPU oracle.
__host__ __device__ __forceinline__ uint64_t rnd16_call_index(uint64_t g) { return ((g >> 6) << 5) | (g & 31u); }
__host__ __device__ __forceinline__ uint32_t rnd16_sub(uint64_t g) { return (uint32_t)(g >> 5) & 1u; }
__host__ __device__ __forceinline__ uint4 rnd16_call(const PhiloxKeys& K, uint64_t g, uint64_t offset) {
  return philox_group(K, rnd16_call_index(g), offset);
}
// the 16-bit value of element j (0..7) of a group whose call returned r
__host__ __device__ __forceinline__ uint32_t rnd16_k(const uint4& r, uint32_t sub, int j) {
  const uint32_t w = philox_word(r, j >> 1);
  const uint32_t t = 2u * (uint32_t)(j & 1) + sub;              // window: bytes t and (t + 1) & 3
  return ((w >> (8u * t)) & 0xFFu) | (((w >> (8u * ((t + 1u) & 3u))) & 0xFFu) << 8);
}
// PRMT selectors that build the float 2^e + k * 2^(e-23) from a word and a constant 0xEE000000 (exponent byte):
// result bytes = [low byte of k, high byte of k, 0x00, exponent byte]
__host__ __device__ __forceinline__ uint32_t rnd16_sel_even(uint32_t sub) { return sub ? 0x7621u : 0x7610u; }
__host__ __device__ __forceinline__ uint32_t rnd16_sel_odd(uint32_t sub) { return sub ? 0x7603u : 0x7632u; }

// ---- warp helpers --------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= o) v += t;
  }
  return v;
}

// smaq_stats_full for a workspace whose ticket word is known to be zero (smaq_stats.cu; used by smaq_compress)
int stats_full_zeroed_ws(const float* x, int64_t n, int unbiased, float* mean_std, void* ws, size_t ws_bytes,
                         cudaStream_t stream);

}  // namespace smaq
