// Kernels (2) and (3): the SmaQ quantizer that MATERIALISES its codes as a dense bit stream,
// and the matching dequantizer.
//
// The reference only ever holds the code as an fp32 value (smart_compress/compress/smart.py:164-169)
// and accounts for 6 bits per main element and 8 per outlier (smart.py:184-187).  The stream
// written here has exactly that size plus a 4-byte table entry per 8192 elements and at most 31
// padding bits per 1024 elements; its layout ("SQB1") is specified in DESIGN.md and restated
// executable in oracle/pack.py.
//
// Mapping to the hardware
//   * one warp owns 1024 consecutive elements; the tile is brought into shared memory by ONE TMA
//     bulk copy (cp.async.bulk + mbarrier), double-buffered, so the next tile is in flight while
//     this one is quantised — these kernels are bound by instruction issue, and without the
//     prefetch too few bytes are in flight to keep HBM busy (profiles/);
//   * lane l takes elements 256k + 8l + j (two 128-bit shared loads per chunk; eight elements share
//     one Philox call) and packs its own
//     32 codes in registers: one tag word and PM words of base fields — the fixed-position part
//     of the stream (PM+1 bits per element), stored with coalesced 32-bit row writes.  No
//     __ballot_sync transposition: 6 votes per element cost more issue slots than the quantiser;
//   * the variable part — XB extra bits per OUTLIER — is compacted per lane in registers, placed
//     inside the warp tile by a shuffle prefix scan of the per-lane outlier counts (popc of the
//     tag word), and placed in the tensor by a single-pass decoupled look-back over groups of CTA
//     tiles (groups are numbered by an atomic ticket, so a group only ever waits for groups that
//     are already running).  Placement is deterministic: the stream is byte-identical run to run;
//   * the decoder needs no scan across tiles: the per-tile word offset is in the table.
//
// HBM roofline (6/8 bits, fraction f of outliers): encode reads 4 B and writes (6 + 2f)/8 B per
// element; decode the reverse.  f = 0.165 on the benchmark input -> 4.79 B per element each way.
#include "params.cuh"

namespace smaq {

constexpr int kWarpTile = 1024;
constexpr int kWarpsPerCta = 8;
constexpr int kCtaTile = kWarpTile * kWarpsPerCta;
constexpr int kPackThreads = 32 * kWarpsPerCta;
constexpr uint32_t kMagic = 0x31425153u;  // 'SQB1'

// groups of lane-local elements that share one 32-bit extras accumulator: XB * (32 / G) <= 32
__host__ __device__ constexpr int ext_groups(int xb) { return xb <= 1 ? 1 : xb <= 2 ? 2 : xb <= 4 ? 4 : 8; }
// 32-bit words of extras a warp tile can need
__host__ __device__ constexpr int seg_words(int xb) { return xb == 0 ? 1 : kWarpTile * xb / 32; }

template <int PM>
__device__ __forceinline__ void put_field(uint32_t (&bw)[PM], int i, uint32_t field) {
  const int pos = PM * i, w = pos >> 5, sh = pos & 31;  // compile-time after unrolling
  bw[w] += field << sh;                                  // fields never overlap: add == or (one LEA)
  if (sh + PM > 32) bw[w + 1] += field >> (32 - sh);
}
template <int PM>
__device__ __forceinline__ uint32_t get_field(const uint32_t (&bw)[PM], int i) {
  const int pos = PM * i, w = pos >> 5, sh = pos & 31;
  uint32_t v = bw[w] >> sh;
  if (sh + PM > 32) v |= bw[w + 1] << (32 - sh);
  return v & ((1u << PM) - 1u);
}

// What a lane accumulates for its 32 elements of one warp tile.
template <int PM>
struct LaneWords {
  uint32_t tagw;      // bit i: element i is an outlier
  uint32_t bw[PM];    // 32 base fields of PM bits
  uint32_t ea[8];     // extras accumulators (ext_groups(XB) of them are used)
  uint32_t ecnt[8];   // bits held by each
  uint32_t n_sat;     // codes clipped at the field width / not a number
};

// Integer side of one element: rounded code (fp32) -> stored payload; updates the lane's words.
//   payload = (min(|code|, limit) << 1) | s,  s = sign bit of the code (main) / of z (outlier: lower side)
// Branch-free: `m` is the all-ones/zero outlier mask, `zb` the bits of z.  F2I saturates large
// codes and maps NaN to 0, which is the format's rule; `vmask` is zero for padding elements.
template <int PM, int XB, bool kCheckNan>
__device__ __forceinline__ void pack_element(int i, float code, uint32_t m, uint32_t zb, uint32_t vmask,
                                             const uint32_t lim_main, const uint32_t lim_out, LaneWords<PM>& L) {
  constexpr int EPG = 32 / ext_groups(XB);
  const uint32_t a = (uint32_t)abs(__float2int_rz(code));
  const uint32_t lim = (lim_out & m) | (lim_main & ~m);
  const uint32_t mag = min(a, lim);
  uint32_t cb = bits_of(code);
  bool bad = a > lim;
  if (kCheckNan) {  // exact path only: the fast path never sees a NaN code (it is flagged upstream)
    const bool isnan_ = code != code;
    bad = bad || isnan_;
    cb = isnan_ ? 0u : cb;
  }
  const uint32_t sbit = ((zb & m) | (cb & ~m)) >> 31;
  const uint32_t payload = ((mag << 1) | sbit) & vmask;
  L.n_sat += (bad && vmask) ? 1u : 0u;
  L.tagw |= m & vmask & (1u << i);
  put_field<PM>(L.bw, i, payload & ((1u << PM) - 1u));
  if (XB > 0) {
    const int g = i / EPG;
    L.ea[g] |= (payload >> PM) << L.ecnt[g];  // payload >> PM is 0 for a main element
    L.ecnt[g] += m & vmask & (uint32_t)XB;
  }
}

// Exact (IEEE-divide) re-computation of one 4-element chunk: degenerate statistics or a flagged
// chunk.  Out of line; the result comes back in registers.
struct ExactChunk {
  float4 code;
  uint32_t cls;  // bit j: outlier; bit 4+j: z negative
};
template <bool kStochastic>
__device__ __noinline__ ExactChunk encode_chunk_exact(float4 v, float4 pr, const Scalars& s) {
  PairClass k0, k1;
  bool unused = false;
  const f32x2 c01 = encode_pair<kStochastic, false>(pair(v.x, v.y), pair(pr.x, pr.y), s, k0, unused);
  const f32x2 c23 = encode_pair<kStochastic, false>(pair(v.z, v.w), pair(pr.z, pr.w), s, k1, unused);
  ExactChunk r;
  r.code = make_float4(c01.x, c01.y, c23.x, c23.y);
  r.cls = (k0.m0 & 1u) | (k0.m1 & 2u) | (k1.m0 & 4u) | (k1.m1 & 8u) | ((k0.zb0 >> 31) << 4) | ((k0.zb1 >> 31) << 5) |
          ((k1.zb0 >> 31) << 6) | ((k1.zb1 >> 31) << 7);
  return r;
}

// One warp tile: quantise 1024 values and pack them into the lane's words.  `staged` points at the
// tile in shared memory (TMA), or is null: direct global loads (unaligned tensors, the ragged
// last tile).
template <int PM, int XB, bool kStochastic, bool kHasProbs, bool kFast>
__device__ __forceinline__ void encode_tile(const float* __restrict__ x, int64_t n, const float* __restrict__ probs,
                                            const KernelParams& kp, const Scalars& s, const float4* staged,
                                            int64_t base, LaneWords<PM>& L) {
  const int lane = lane_id();
  const uint32_t lim_main = (uint32_t)s.lim_main, lim_out = (uint32_t)s.lim_out;
  L.tagw = 0;
  L.n_sat = 0;
#pragma unroll
  for (int w = 0; w < PM; ++w) L.bw[w] = 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) { L.ea[g] = 0; L.ecnt[g] = 0; }

  const bool full = staged != nullptr;
  const bool probs_vec = kStochastic && kHasProbs && full && aligned16(probs);
  // lane l owns elements 256k + 8l + j of the tile (k = 0..3, j = 0..7): local index i = 8k + j
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t e = base + 256 * k + 8 * lane;
    float4 v[2], p4[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t eh = e + 4 * h;
      p4[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full) v[h] = staged[64 * k + 2 * lane + h];
      else v[h] = make_float4(eh < n ? x[eh] : 0.f, eh + 1 < n ? x[eh + 1] : 0.f, eh + 2 < n ? x[eh + 2] : 0.f,
                              eh + 3 < n ? x[eh + 3] : 0.f);
      if (kStochastic && kHasProbs) {
        if (probs_vec) p4[h] = ldg_stream(reinterpret_cast<const float4*>(probs + eh));
        else p4[h] = make_float4(eh < n ? probs[eh] : 0.f, eh + 1 < n ? probs[eh + 1] : 0.f,
                                 eh + 2 < n ? probs[eh + 2] : 0.f, eh + 3 < n ? probs[eh + 3] : 0.f);
      }
    }
    if (kStochastic && !kHasProbs) {
      const uint4 r = philox_group(kp.keys, (uint64_t)(e >> 3), kp.offset);
      const f32x2 q0 = uniform16_pair(r.x), q1 = uniform16_pair(r.y), q2 = uniform16_pair(r.z), q3 = uniform16_pair(r.w);
      p4[0] = make_float4(q0.x, q0.y, q1.x, q1.y);
      p4[1] = make_float4(q2.x, q2.y, q3.x, q3.y);
    }
    bool suspect = false;
    if (kFast && full) {
      PairClass k0, k1, k2, k3;
      const f32x2 c0 = encode_pair<kStochastic, true>(pair(v[0].x, v[0].y), pair(p4[0].x, p4[0].y), s, k0, suspect);
      const f32x2 c1 = encode_pair<kStochastic, true>(pair(v[0].z, v[0].w), pair(p4[0].z, p4[0].w), s, k1, suspect);
      const f32x2 c2 = encode_pair<kStochastic, true>(pair(v[1].x, v[1].y), pair(p4[1].x, p4[1].y), s, k2, suspect);
      const f32x2 c3 = encode_pair<kStochastic, true>(pair(v[1].z, v[1].w), pair(p4[1].z, p4[1].w), s, k3, suspect);
      if (!suspect) {  // the hot path: branch-free packing of eight elements
        pack_element<PM, XB, false>(8 * k + 0, c0.x, k0.m0, k0.zb0, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 1, c0.y, k0.m1, k0.zb1, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 2, c1.x, k1.m0, k1.zb0, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 3, c1.y, k1.m1, k1.zb1, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 4, c2.x, k2.m0, k2.zb0, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 5, c2.y, k2.m1, k2.zb1, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 6, c3.x, k3.m0, k3.zb0, 0xFFFFFFFFu, lim_main, lim_out, L);
        pack_element<PM, XB, false>(8 * k + 7, c3.y, k3.m1, k3.zb1, 0xFFFFFFFFu, lim_main, lim_out, L);
      }
    } else {
      suspect = true;
    }
    if (suspect) {  // rare: flagged chunk, degenerate statistics, or the ragged last tile
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const ExactChunk ex = encode_chunk_exact<kStochastic>(v[h], p4[h], s);
        const float cj[4] = {ex.code.x, ex.code.y, ex.code.z, ex.code.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t vmask = (full || (e + 4 * h + j) < n) ? 0xFFFFFFFFu : 0u;
          const uint32_t m = ((ex.cls >> j) & 1u) ? 0xFFFFFFFFu : 0u;
          const uint32_t zb = ((ex.cls >> (4 + j)) & 1u) << 31;
          pack_element<PM, XB, true>(8 * k + 4 * h + j, cj[j], m, zb, vmask, lim_main, lim_out, L);
        }
      }
    }
  }
}

// ---- TMA bulk copy of one warp tile (4 KB, contiguous) into shared memory ----------------------
// One lane issues cp.async.bulk; every lane of the warp waits on the warp's mbarrier.  The copy
// is linear, so lane l finds its chunk k at float4 index 32k + l — the same coalesced mapping a
// direct 128-bit load would use.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_tile(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

constexpr int kMaxRounds = 4;  // CTA tiles per CTA ("group")
constexpr int kStages = 2;
constexpr int kStageBytes = kWarpTile * 4;                                  // 4 KB per warp tile
constexpr int kEncodeDynSmem = kWarpsPerCta * kStages * kStageBytes + 1024;  // + alignment slack

// Scratch of one encode call (device).  Placement of the variable-length part is
// reduce-then-scan: pass 1 (the expensive one, one read of x) quantises, writes the fixed part of
// the stream and parks every warp tile's extras segment at a fixed stride; pass 2 scans the
// per-group word counts (one small block); pass 3 moves the segments to their dense positions.
// No kernel ever waits on another block, so nothing here can hang, and the placement is a pure
// function of the input: the stream is byte-identical run to run.  (A single-pass decoupled
// look-back was measured first: with ~450 groups resident, each group's look-back walk under full
// HBM load cost ~10x its own compute time — profiles/r1_encode_lookback.txt.)
struct EncodeScratch {
  uint32_t* group_words;  // [n_groups] extras words of each group
  uint32_t* group_nout;   // [n_groups] outliers
  uint32_t* group_nsat;   // [n_groups] clipped / NaN codes
  uint32_t* group_off;    // [n_groups] exclusive prefix of group_words (pass 2)
  uint32_t* seg_words;    // [n_warp_tiles] words of each warp tile's segment
  uint32_t* staging;      // [n_warp_tiles][seg_words(XB)] parked segments
};

// Pass 1.  One CTA = one group of `rounds` consecutive CTA tiles (8 warp tiles each); each warp
// runs through its `rounds` warp tiles on its own: while it quantises tile r out of shared
// memory, TMA is already filling the other stage with tile r+1.  No block barrier in the loop.
template <int PM, int XB, bool kStochastic, bool kHasProbs>
__global__ void __launch_bounds__(kPackThreads, 3)
    encode_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ mean_std,
                  const float* __restrict__ probs, const __grid_constant__ KernelParams kp,
                  uint32_t* __restrict__ planes, EncodeScratch sc, long long n_cta_tiles, int rounds, int aligned) {
  constexpr int G = ext_groups(XB);
  constexpr int kSeg = seg_words(XB);
  extern __shared__ unsigned char dyn_smem[];
  __shared__ uint32_t s_seg[kWarpsPerCta][kSeg + 1];  // +1: spill word of the last atomicOr
  __shared__ uint32_t s_tot[3][kWarpsPerCta];
  __shared__ __align__(8) uint64_t s_bar[kWarpsPerCta][kStages];

  const int lane = lane_id(), warp = warp_id();
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) mbar_init(&s_bar[warp][st], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const long long group = blockIdx.x;
  const long long first_tile = group * rounds;
  const int nrounds = (int)min((long long)rounds, n_cta_tiles - first_tile);

  // this warp's two 4 KB stages (1 KB aligned)
  unsigned char* stage_base =
      (unsigned char*)(((uintptr_t)dyn_smem + 1023) & ~(uintptr_t)1023) + (size_t)warp * kStages * kStageBytes;

  const Scalars s = scalars_from(mean_std[0], mean_std[1], kp);
  auto tile_base = [&](int r) { return ((int64_t)(first_tile + r) * kWarpsPerCta + warp) * kWarpTile; };
  // a warp tile is staged through TMA when it lies wholly inside the tensor and x is 16-byte aligned
  auto tile_is_full = [&](int r) { return aligned && (tile_base(r) + kWarpTile <= n); };
  auto issue = [&](int r) {
    if (lane == 0 && tile_is_full(r))
      tma_load_tile(stage_base + (r % kStages) * kStageBytes, x + tile_base(r), kStageBytes, &s_bar[warp][r % kStages]);
  };

  issue(0);
  uint32_t n_out_total = 0, n_sat_total = 0, words_total = 0;
  uint32_t* seg = s_seg[warp];
  for (int r = 0; r < nrounds; ++r) {
    if (r + 1 < nrounds) issue(r + 1);  // stage (r+1)%2 was fully consumed in round r-1 (__syncwarp below)
    const int64_t base = tile_base(r);
    if (base >= n) break;  // warp tiles past the end of the tensor: nothing stored (uniform per warp)
    const int64_t wt = base / kWarpTile;
    const float4* staged = nullptr;
    if (tile_is_full(r)) {
      mbar_wait(&s_bar[warp][r % kStages], (uint32_t)((r / kStages) & 1));
      staged = reinterpret_cast<const float4*>(stage_base + (r % kStages) * kStageBytes);
    }
    if (XB > 0) {
      for (int j = lane; j < kSeg + 1; j += 32) seg[j] = 0;
    }
    LaneWords<PM> L;
    if (s.fast) encode_tile<PM, XB, kStochastic, kHasProbs, true>(x, n, probs, kp, s, staged, base, L);
    else encode_tile<PM, XB, kStochastic, kHasProbs, false>(x, n, probs, kp, s, staged, base, L);
    __syncwarp();  // stage fully read (lane 0 may refill it); segment zeroing visible to the whole warp

    // fixed-position part of the stream: (1 + PM) rows of 32 words per warp tile, coalesced
    uint32_t* rec = planes + wt * (int64_t)((1 + PM) * 32) + lane;
    rec[0] = L.tagw;
#pragma unroll
    for (int w = 0; w < PM; ++w) rec[32 * (w + 1)] = L.bw[w];

    // variable part: the lane's extras go into this warp tile's word-aligned segment
    const uint32_t n_out = __popc(L.tagw);
    n_out_total += n_out;
    n_sat_total += L.n_sat;
    if (XB > 0) {
      const uint32_t inc = warp_inclusive_scan(n_out * XB);
      uint32_t pos = inc - n_out * XB;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (L.ecnt[g]) {
          const uint32_t w = pos >> 5, sh = pos & 31;
          atomicOr(&seg[w], L.ea[g] << sh);
          if (sh + L.ecnt[g] > 32) atomicOr(&seg[w + 1], L.ea[g] >> (32 - sh));
          pos += L.ecnt[g];
        }
      }
      const uint32_t nwords = (__shfl_sync(0xffffffffu, inc, 31) + 31) >> 5;
      __syncwarp();
      uint32_t* park = sc.staging + wt * (int64_t)kSeg;
      for (uint32_t j = lane; j < nwords; j += 32) park[j] = seg[j];
      if (lane == 0) sc.seg_words[wt] = nwords;
      words_total += nwords;
      __syncwarp();  // the copy is done before the next round zeroes the segment
    }
  }

  // per-group totals (plain stores: no atomics, no pre-zeroed memory)
  const uint32_t w_out = warp_sum(n_out_total), w_sat = warp_sum(n_sat_total);
  if (lane == 0) {
    s_tot[0][warp] = words_total;
    s_tot[1][warp] = w_out;
    s_tot[2][warp] = w_sat;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) t += s_tot[threadIdx.x][w];
    uint32_t* dst = threadIdx.x == 0 ? sc.group_words : threadIdx.x == 1 ? sc.group_nout : sc.group_nsat;
    dst[group] = t;
  }
}

// Pass 2: one block scans the per-group word counts and finishes the header.
__global__ void __launch_bounds__(1024) encode_scan_kernel(EncodeScratch sc, long long n_groups, long long n_cta_tiles,
                                                           smaq_packed_header* __restrict__ hdr,
                                                           uint32_t* __restrict__ table,
                                                           const float* __restrict__ mean_std,
                                                           const __grid_constant__ KernelParams kp, int64_t n,
                                                           int stochastic) {
  __shared__ unsigned long long s_w[32], s_o[32], s_s[32];
  __shared__ unsigned long long s_carry[3];
  const int lane = lane_id(), warp = warp_id();
  if (threadIdx.x < 3) s_carry[threadIdx.x] = 0;
  __syncthreads();
  for (long long base = 0; base < n_groups; base += 1024) {
    const long long g = base + threadIdx.x;
    const unsigned long long w = g < n_groups ? sc.group_words[g] : 0ull;
    const unsigned long long o = g < n_groups ? sc.group_nout[g] : 0ull;
    const unsigned long long sa = g < n_groups ? sc.group_nsat[g] : 0ull;
    // inclusive scan of w inside the warp; plain sums of o and sa
    unsigned long long inc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    const unsigned long long so = warp_sum(o), ss = warp_sum(sa);
    if (lane == 31) s_w[warp] = inc;
    if (lane == 0) { s_o[warp] = so; s_s[warp] = ss; }
    __syncthreads();
    unsigned long long before = s_carry[0];
    for (int i = 0; i < warp; ++i) before += s_w[i];
    if (g < n_groups) sc.group_off[g] = (uint32_t)(before + inc - w);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tw = 0, to = 0, ts = 0;
      for (int i = 0; i < 32; ++i) { tw += s_w[i]; to += s_o[i]; ts += s_s[i]; }
      s_carry[0] += tw; s_carry[1] += to; s_carry[2] += ts;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    hdr->magic = kMagic;
    hdr->bits_main = kp.bits_main;
    hdr->bits_outlier = kp.bits_outlier;
    hdr->stochastic = stochastic;
    hdr->n = n;
    hdr->mean = mean_std[0];
    hdr->std_raw = mean_std[1];
    hdr->threshold = kp.thr;
    hdr->range_main = kp.range_main;
    hdr->range_outlier = kp.range_out;
    hdr->clamp_lo = kp.clamp_lo;
    hdr->clamp_hi = kp.clamp_hi;
    hdr->pad0 = 0.0f;
    hdr->n_outlier = s_carry[1];
    hdr->n_saturated = s_carry[2];
    hdr->extras_words = s_carry[0];
    hdr->status = 0;
    table[n_cta_tiles] = (uint32_t)s_carry[0];
  }
}

// Pass 3: move every parked segment to its dense position and write the per-CTA-tile table.
template <int XB>
__global__ void __launch_bounds__(kPackThreads) encode_place_kernel(EncodeScratch sc, uint32_t* __restrict__ table,
                                                                    uint32_t* __restrict__ extras,
                                                                    long long n_cta_tiles, long long n_warp_tiles,
                                                                    int rounds) {
  constexpr int kSeg = seg_words(XB);
  __shared__ uint32_t s_off[kMaxRounds * kWarpsPerCta + 1];
  const int lane = lane_id(), warp = warp_id();
  const long long group = blockIdx.x;
  const long long first_wt = group * rounds * kWarpsPerCta;
  const int nseg = rounds * kWarpsPerCta;  // <= 32
  if (warp == 0) {
    const long long wt = first_wt + lane;
    const uint32_t wds = (lane < nseg && wt < n_warp_tiles) ? sc.seg_words[wt] : 0u;
    const uint32_t inc = warp_inclusive_scan(wds);
    s_off[lane] = sc.group_off[group] + inc - wds;
    if (lane == 31) s_off[32] = sc.group_off[group] + inc;
  }
  __syncthreads();
  for (int i = warp; i < nseg; i += kWarpsPerCta) {
    const long long wt = first_wt + i;
    if (wt >= n_warp_tiles) break;
    const uint32_t off = s_off[i], cnt = s_off[i + 1] - off;
    const uint32_t* src = sc.staging + wt * (int64_t)kSeg;
    for (uint32_t j = lane; j < cnt; j += 32) extras[(size_t)off + j] = src[j];
  }
  if ((int)threadIdx.x < rounds) {
    const long long tile = group * rounds + threadIdx.x;
    if (tile < n_cta_tiles) table[tile] = s_off[threadIdx.x * kWarpsPerCta];
  }
}

// ---- decoder ----------------------------------------------------------------------------------------
template <int PM, int XB, bool kFast, bool kAllPos>
__device__ __forceinline__ void decode_tile(const uint32_t* __restrict__ seg, uint32_t pos, uint32_t tagw,
                                            const uint32_t (&bw)[PM], const Scalars& s, float* __restrict__ y,
                                            int64_t n, int64_t base, bool aligned) {
  constexpr int G = ext_groups(XB);
  constexpr int EPG = 32 / G;
  const int lane = lane_id();
  uint32_t win[G];
  if (XB > 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const uint32_t gmask = (EPG == 32) ? 0xffffffffu : (((1u << EPG) - 1u) << (g * EPG));
      const uint32_t cnt = __popc(tagw & gmask) * XB;
      const uint32_t w = pos >> 5, sh = pos & 31;
      win[g] = __funnelshift_r(seg[w], seg[w + 1], sh);
      pos += cnt;
    }
  }
  const uint32_t tb = bits_of(s.thr);
  const bool full = aligned && (base + kWarpTile <= n);
  // lane l owns elements 256k + 8l + j of the tile (k = 0..3, j = 0..7): local index i = 8k + j
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float out[8];
#pragma unroll
    for (int h = 0; h < 4; ++h) {  // four packed pairs
      float code[2], shift[2], rb[2], rr[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = 8 * k + 2 * h + q;
        const uint32_t m = 0u - ((tagw >> i) & 1u);  // outlier mask
        uint32_t payload = get_field<PM>(bw, i);
        if (XB > 0) {
          const int g = i / EPG;
          payload |= ((win[g] & ((1u << XB) - 1u)) << PM) & m;
          win[g] >>= (m & (uint32_t)XB);
        }
        const uint32_t sb = payload << 31;  // s bit moved to the sign position
        const float mag = (float)(payload >> 1);
        code[q] = from_bits(bits_of(mag) | sb);               // -0.0 when s and mag == 0
        shift[q] = from_bits(((sb ^ 0x80000000u) | tb) & m);  // -t above, +t below, +0 inside
        rb[q] = select_f(m, s.range_out.b, s.range_main.b);
        rr[q] = select_f(m, s.range_out.r, s.range_main.r);
      }
      bool unused = false;
      const f32x2 yv = decode_pair<kFast, false>(pair(code[0], code[1]), pair(shift[0], shift[1]), pair(rb[0], rb[1]),
                                                 pair(rr[0], rr[1]), s, kAllPos, unused);
      out[2 * h] = yv.x;
      out[2 * h + 1] = yv.y;
    }
    const int64_t e = base + 256 * k + 8 * lane;
    if (full) {
      f32x8 o;
      o.a = make_float4(out[0], out[1], out[2], out[3]);
      o.b = make_float4(out[4], out[5], out[6], out[7]);
      stg_stream8(y + e, o);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (e + j < n) y[e + j] = out[j];
    }
  }
}

template <int PM, int XB>
__global__ void __launch_bounds__(kPackThreads, 4)
    decode_kernel(const smaq_packed_header* __restrict__ hdr, const uint32_t* __restrict__ table,
                  const uint32_t* __restrict__ planes, const uint32_t* __restrict__ extras, float* __restrict__ y,
                  int64_t n, int all_positive, int aligned) {
  constexpr int kSeg = seg_words(XB);
  __shared__ uint32_t s_seg[kWarpsPerCta][kSeg + 2];
  __shared__ uint32_t s_warp[kWarpsPerCta];
  const long long tile = blockIdx.x;
  const int lane = lane_id(), warp = warp_id();
  const int64_t wt = (int64_t)tile * kWarpsPerCta + warp;
  const int64_t base = wt * kWarpTile;

  const Scalars s = make_scalars(hdr->mean, hdr->std_raw, hdr->threshold, hdr->range_main, hdr->range_outlier,
                                 hdr->clamp_lo, hdr->clamp_hi, hdr->bits_main, hdr->bits_outlier);
  uint32_t tagw = 0, bw[PM];
#pragma unroll
  for (int w = 0; w < PM; ++w) bw[w] = 0;
  if (base < n) {
    const uint32_t* rec = planes + wt * (int64_t)((1 + PM) * 32) + lane;
    tagw = rec[0];
#pragma unroll
    for (int w = 0; w < PM; ++w) bw[w] = rec[32 * (w + 1)];
  }
  // lane offsets inside the warp tile's segment; the segment's place inside the CTA tile
  const uint32_t nb = __popc(tagw) * XB;
  const uint32_t inc = warp_inclusive_scan(nb);
  const uint32_t my_words = (__shfl_sync(0xffffffffu, inc, 31) + 31) >> 5;
  if (lane == 0) s_warp[warp] = my_words;
  __syncthreads();
  uint32_t seg_off = table[tile];
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) seg_off += (w < warp) ? s_warp[w] : 0u;
  for (uint32_t i = lane; i < my_words; i += 32) s_seg[warp][i] = extras[(size_t)seg_off + i];
  if (lane < 2) s_seg[warp][my_words + lane] = 0;
  __syncwarp();
  if (base >= n) return;
  const uint32_t pos = inc - nb;
  if (s.fast) {
    if (all_positive) decode_tile<PM, XB, true, true>(s_seg[warp], pos, tagw, bw, s, y, n, base, aligned != 0);
    else decode_tile<PM, XB, true, false>(s_seg[warp], pos, tagw, bw, s, y, n, base, aligned != 0);
  } else {
    if (all_positive) decode_tile<PM, XB, false, true>(s_seg[warp], pos, tagw, bw, s, y, n, base, aligned != 0);
    else decode_tile<PM, XB, false, false>(s_seg[warp], pos, tagw, bw, s, y, n, base, aligned != 0);
  }
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static bool width_supported(int bits_main, int bits_outlier) {
  const int pm = bits_main - 1, xb = bits_outlier - bits_main;
#ifdef SMAQ_PACK_MINIMAL
  return pm == 5 && xb == 2;
#endif
  return pm >= 3 && pm <= 7 && xb >= 0 && xb <= 4;
}

static int rounds_for(int64_t n_cta_tiles) {
  // CTA tiles per CTA: enough groups for >= 4 waves over 3 resident CTAs per SM, at most kMaxRounds
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int64_t r = n_cta_tiles / ((int64_t)sms * 3 * 4);
  return (int)(r < 1 ? 1 : (r > kMaxRounds ? kMaxRounds : r));
}

}  // namespace smaq

extern "C" {

int smaq_packed_layout_for(int64_t n, int32_t bits_main, int32_t bits_outlier, smaq_packed_layout* out) {
  using namespace smaq;
  if (!out || n < 0) return fail(SMAQ_ERR_ARG, "packed_layout: bad argument");
  if (!width_supported(bits_main, bits_outlier))
    return fail(SMAQ_ERR_UNSUPPORTED,
                "packed stream supports num_bits_main 4..8 and num_bits_outlier - num_bits_main 0..4 (got %d/%d); "
                "the fake-quantisation round trip has no such limit",
                bits_main, bits_outlier);
  const int pm = bits_main - 1, xb = bits_outlier - bits_main;
  smaq_packed_layout l;
  l.n = n;
  l.bits_main = bits_main;
  l.bits_outlier = bits_outlier;
  l.n_warp_tiles = (n + kWarpTile - 1) / kWarpTile;
  l.n_cta_tiles = (l.n_warp_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  l.header_off = 0;
  l.header_bytes = 128;
  l.table_off = l.header_off + l.header_bytes;
  l.table_bytes = align_up((l.n_cta_tiles + 1) * 4, 128);
  l.planes_off = l.table_off + l.table_bytes;
  l.planes_bytes = l.n_warp_tiles * (int64_t)(1 + pm) * 128;
  l.extras_off = l.planes_off + l.planes_bytes;
  l.extras_capacity_bytes = align_up(l.n_warp_tiles * (int64_t)(kWarpTile * xb / 32) * 4 + 4, 128);
  l.total_capacity_bytes = l.extras_off + l.extras_capacity_bytes;
  // scratch: 4 uint32 per group (at most one group per CTA tile) + 1 per warp tile + the parked segments
  l.workspace_bytes = align_up(l.n_cta_tiles * 16 + l.n_warp_tiles * 4, 256) + l.n_warp_tiles * (int64_t)seg_words(xb) * 4 + 256;
  *out = l;
  return SMAQ_OK;
}

int smaq_encode(const float* x, int64_t n, const float* mean_std, const float* probs, const smaq_codec_params* params,
                void* packed, size_t packed_bytes, void* ws, size_t ws_bytes, smaq_stream_t stream_) {
  using namespace smaq;
  static_assert(sizeof(smaq_packed_header) <= 128, "header must fit its slot");
  if (int rc = check_params(params)) return rc;
  if (!x || !mean_std || !packed || !ws || n <= 0) return fail(SMAQ_ERR_ARG, "encode: null pointer or n <= 0");
  smaq_packed_layout l;
  if (int rc = smaq_packed_layout_for(n, params->bits_main, params->bits_outlier, &l)) return rc;
  if (packed_bytes < (size_t)l.total_capacity_bytes) return fail(SMAQ_ERR_WORKSPACE, "encode: packed buffer too small");
  if (ws_bytes < (size_t)l.workspace_bytes) return fail(SMAQ_ERR_WORKSPACE, "encode: workspace too small");
  if (!aligned16(packed)) return fail(SMAQ_ERR_ARG, "encode: packed buffer must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rounds = rounds_for(l.n_cta_tiles);
  const long long n_groups = (l.n_cta_tiles + rounds - 1) / rounds;
  char* pb = (char*)packed;
  auto* hdr = (smaq_packed_header*)(pb + l.header_off);
  auto* table = (uint32_t*)(pb + l.table_off);
  auto* planes = (uint32_t*)(pb + l.planes_off);
  auto* extras = (uint32_t*)(pb + l.extras_off);
  const KernelParams kp = to_kernel_params(*params);
  const int aligned = aligned16(x);
  const int pm = params->bits_main - 1, xb = params->bits_outlier - params->bits_main;
  const bool st = params->stochastic != 0;
  const bool hp = st && probs != nullptr;
  const unsigned grid = (unsigned)n_groups;
  EncodeScratch sc;
  {
    uint32_t* w = (uint32_t*)ws;
    sc.group_words = w;
    sc.group_nout = w + l.n_cta_tiles;
    sc.group_nsat = w + 2 * l.n_cta_tiles;
    sc.group_off = w + 3 * l.n_cta_tiles;
    sc.seg_words = w + 4 * l.n_cta_tiles;
    sc.staging = (uint32_t*)((char*)ws + align_up(l.n_cta_tiles * 16 + l.n_warp_tiles * 4, 256));
  }

#define SMAQ_ENC_LAUNCH(PM_, XB_, ST_, HP_)                                                                        \
  {                                                                                                                \
    auto kern = encode_kernel<PM_, XB_, ST_, HP_>;                                                                 \
    SMAQ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncodeDynSmem));         \
    kern<<<grid, kPackThreads, kEncodeDynSmem, stream>>>(x, n, mean_std, probs, kp, planes, sc, l.n_cta_tiles,     \
                                                         rounds, aligned);                                         \
    SMAQ_LAUNCH_OK();                                                                                              \
    encode_scan_kernel<<<1, 1024, 0, stream>>>(sc, n_groups, l.n_cta_tiles, hdr, table, mean_std, kp, n, ST_);     \
    encode_place_kernel<XB_><<<grid, kPackThreads, 0, stream>>>(sc, table, extras, l.n_cta_tiles, l.n_warp_tiles,  \
                                                                rounds);                                           \
  }
#define SMAQ_ENC(PM_, XB_)                                                                                         \
  if (pm == PM_ && xb == XB_) {                                                                                    \
    if (!st) SMAQ_ENC_LAUNCH(PM_, XB_, false, false)                                                               \
    else if (hp) SMAQ_ENC_LAUNCH(PM_, XB_, true, true)                                                             \
    else SMAQ_ENC_LAUNCH(PM_, XB_, true, false)                                                                    \
  }
#define SMAQ_ENC_ROW(PM_) SMAQ_ENC(PM_, 0) SMAQ_ENC(PM_, 1) SMAQ_ENC(PM_, 2) SMAQ_ENC(PM_, 3) SMAQ_ENC(PM_, 4)
#ifdef SMAQ_PACK_MINIMAL
  SMAQ_ENC(5, 2)
#else
  SMAQ_ENC_ROW(3) SMAQ_ENC_ROW(4) SMAQ_ENC_ROW(5) SMAQ_ENC_ROW(6) SMAQ_ENC_ROW(7)
#endif
#undef SMAQ_ENC_ROW
#undef SMAQ_ENC
#undef SMAQ_ENC_LAUNCH
  SMAQ_LAUNCH_OK();
  return SMAQ_OK;
}

int smaq_decode(const void* packed, size_t packed_bytes, int64_t n, int32_t bits_main, int32_t bits_outlier,
                int32_t all_positive, float* y, smaq_stream_t stream_) {
  using namespace smaq;
  if (!packed || !y || n <= 0) return fail(SMAQ_ERR_ARG, "decode: null pointer or n <= 0");
  smaq_packed_layout l;
  if (int rc = smaq_packed_layout_for(n, bits_main, bits_outlier, &l)) return rc;
  if (packed_bytes < (size_t)l.extras_off) return fail(SMAQ_ERR_WORKSPACE, "decode: packed buffer too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  const char* pb = (const char*)packed;
  auto* hdr = (const smaq_packed_header*)(pb + l.header_off);
  auto* table = (const uint32_t*)(pb + l.table_off);
  auto* planes = (const uint32_t*)(pb + l.planes_off);
  auto* extras = (const uint32_t*)(pb + l.extras_off);
  const int aligned = aligned32(y);
  const int pm = bits_main - 1, xb = bits_outlier - bits_main;
  const unsigned grid = (unsigned)l.n_cta_tiles;
#define SMAQ_DEC(PM_, XB_)                                                                                         \
  if (pm == PM_ && xb == XB_)                                                                                      \
    decode_kernel<PM_, XB_><<<grid, kPackThreads, 0, stream>>>(hdr, table, planes, extras, y, n, all_positive, aligned);
#define SMAQ_DEC_ROW(PM_) SMAQ_DEC(PM_, 0) SMAQ_DEC(PM_, 1) SMAQ_DEC(PM_, 2) SMAQ_DEC(PM_, 3) SMAQ_DEC(PM_, 4)
#ifdef SMAQ_PACK_MINIMAL
  SMAQ_DEC(5, 2)
#else
  SMAQ_DEC_ROW(3) SMAQ_DEC_ROW(4) SMAQ_DEC_ROW(5) SMAQ_DEC_ROW(6) SMAQ_DEC_ROW(7)
#endif
#undef SMAQ_DEC_ROW
#undef SMAQ_DEC
  SMAQ_LAUNCH_OK();
  return SMAQ_OK;
}

}  // extern "C"
