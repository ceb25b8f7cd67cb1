// radix_sort.cuh — hand-written onesweep LSD radix sort for (u64 key, u32 value) pairs.
//
// Replaces nothing the reference wrote itself: the sort is the first half of the LBVH builder that
// stands in for optixAccelBuild (owl/UserGeomGroup.cpp:199-215).  One upfront histogram pass for
// all 8 digits, then 8 scatter passes; each pass reads and writes every pair exactly once and
// resolves the cross-tile digit offsets with decoupled look-back (single sweep), so the whole
// sort moves 8 x 24 B per pair plus one 8 B histogram read.  Stable.
#pragma once
#include "common.cuh"

namespace tknn {
namespace rsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int PASSES = 64 / RADIX_BITS;
#ifndef TKNN_SORT_THREADS
#define TKNN_SORT_THREADS 256
#endif
#ifndef TKNN_SORT_ITEMS
#define TKNN_SORT_ITEMS 16
#endif
constexpr int THREADS = TKNN_SORT_THREADS;
static_assert(THREADS >= RADIX && THREADS % 32 == 0, "onesweep_kernel: threads 0 .. RADIX-1 own the digits (status, histogram and scatter rows)");
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = TKNN_SORT_ITEMS;      // pairs per thread
constexpr int TILE = THREADS * ITEMS;       // 4096 pairs per tile
constexpr uint32_t FLAG_AGG = 1u << 30;     // tile aggregate published
constexpr uint32_t FLAG_PREFIX = 2u << 30;  // inclusive prefix published
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;
constexpr int LOOKBACK = 8;                   // predecessors fetched per look-back step
constexpr uint64_t MAX_N = (1ull << 30) - 1;  // counts share a word with the two flag bits

constexpr size_t DYN_SMEM = (size_t)TILE * (sizeof(uint64_t) + sizeof(uint32_t));

__host__ __device__ inline uint32_t num_tiles(uint64_t n) { return (uint32_t)((n + TILE - 1) / TILE); }

// hist[pass][digit] += count, all passes in one read of the keys.
static __global__ void __launch_bounds__(THREADS) histogram_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                            uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_hist[PASSES * RADIX];
  for (int i = threadIdx.x; i < PASSES * RADIX; i += THREADS) s_hist[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * THREADS;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    const uint64_t key = keys[i];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) atomicAdd(&s_hist[p * RADIX + (uint32_t)((key >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PASSES * RADIX; i += THREADS) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// in-place exclusive scan of each pass's 256 bins: one block per pass.
// Block 0 also writes the LOOKBACK "prefix 0" rows that sit in front of the tile status array.
static __global__ void __launch_bounds__(RADIX) scan_histogram_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ status_pad) {
  __shared__ uint32_t s_warp[RADIX / 32];
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < LOOKBACK * RADIX; i += RADIX) status_pad[i] = FLAG_PREFIX;
  uint32_t* h = hist + blockIdx.x * RADIX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t v = h[threadIdx.x];
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t off = 0;
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  h[threadIdx.x] = off + inc - v;
}

// Lanes holding the same 8-bit digit.  Eight ballots (one per digit bit) instead of `match.any`: the
// instruction loops over the distinct values in the warp (~30 of them for 8-bit digits) and its latency
// sat on the ranking loop's critical path — 38 % of the pass's stall samples (round-1 capture; the kernel as it is now:
// profiles/r1_sort_v8_onesweep_ncu_summary.json, profiles/r2_build_onesweep_kernel_ncu_summary.json + _regions.txt).
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
  uint32_t peers = FULL_MASK;
#pragma unroll
  for (int b = 0; b < RADIX_BITS; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t m = __ballot_sync(FULL_MASK, bit);
    peers &= bit ? m : ~m;
  }
  return peers;
}

// One onesweep pass over digit `shift / 8`.
#ifndef TKNN_SORT_MINBLOCKS
#define TKNN_SORT_MINBLOCKS 3
#endif
static __global__ void __launch_bounds__(THREADS, TKNN_SORT_MINBLOCKS)
    onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint64_t n, int shift,
                    const uint32_t* __restrict__ bin_offset, uint32_t* status, uint32_t* tile_counter) {
  __shared__ uint32_t s_warp_hist[WARPS][RADIX];
  __shared__ uint32_t s_digit_start[RADIX];
  __shared__ uint32_t s_scatter[RADIX];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_dyn);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);

  // dynamic tile ids: a tile only ever waits on tiles that started before it (deadlock freedom)
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = threadIdx.x; i < WARPS * RADIX; i += THREADS) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;

  uint64_t key[ITEMS];
  uint32_t val[ITEMS];
  uint32_t rank[ITEMS];
  const uint64_t wbase = base + (uint64_t)warp * (32 * ITEMS);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint64_t idx = wbase + i * 32 + lane;
    const bool ok = idx < n;
    key[i] = ok ? keys_in[idx] : ~0ull;  // padding ranks after every real key of digit 255
    val[i] = ok ? vals_in[idx] : 0u;
  }
  // warp-local stable ranking: lanes holding the same digit find each other (match_digit); the
  // lowest of them bumps the warp's digit counter for the whole peer group.
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    const uint32_t peers = match_digit(d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) {
      old = s_warp_hist[warp][d];
      s_warp_hist[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(FULL_MASK, old, leader);
    rank[i] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // thread d < RADIX owns digit d (the first RADIX / 32 warps, all of them when THREADS == RADIX): exclusive scan of the
  // warp counters, publish the tile aggregate
  const int d = threadIdx.x;
  const bool owner = d < RADIX;
  uint32_t total = 0, inc = 0;
  uint32_t* st = status + (size_t)tile * RADIX;
  if (owner) {
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const uint32_t c = s_warp_hist[w][d];
      s_warp_hist[w][d] = total;
      total += c;
    }
    atomicExch(&st[d], total | (tile == 0 ? FLAG_PREFIX : FLAG_AGG));

    // exclusive scan of `total` over the 256 digits -> where each digit starts inside the tile
    inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
  }
  __syncthreads();
  if (owner) {
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_wsum[w];
    const uint32_t dstart = woff + inc - total;
    s_digit_start[d] = dstart;

    // decoupled look-back over the preceding tiles for digit d
    // The walk is latency bound (one dependent global load per predecessor: 30 % of the pass's stall samples),
    // so LOOKBACK predecessors are fetched at once and consumed in order; a predecessor that has published
    // nothing yet restarts the window at that tile.
    uint32_t excl = 0;
    if (tile > 0) {
      int64_t t = (int64_t)tile - 1;
      bool done = false;
      while (!done) {
        uint32_t v[LOOKBACK];
        // rows -1 .. -LOOKBACK in front of tile 0 hold "prefix 0" (scan_histogram_kernel), so the window needs
        // no range test: the walk itself always ends at tile 0, which publishes FLAG_PREFIX directly
        const volatile uint32_t* p = status + t * RADIX + d;
#pragma unroll
        for (int w = 0; w < LOOKBACK; ++w) v[w] = *(p - w * RADIX);
        int used = 0;
#pragma unroll
        for (int w = 0; w < LOOKBACK; ++w) {
          if (!done && used == w) {
            if ((v[w] & FLAG_MASK) != 0) {
              excl += v[w] & VALUE_MASK;
              ++used;
              if (v[w] & FLAG_PREFIX) done = true;
            }
          }
        }
        t -= used;  // used < LOOKBACK: tile t - used was not ready — poll again from there
      }
      atomicExch(&st[d], ((excl + total) & VALUE_MASK) | FLAG_PREFIX);
    }
    s_scatter[d] = bin_offset[d] + excl - dstart;  // global slot of tile-local position s is s_scatter[d] + s
  }
  __syncthreads();

  // stage the tile in sorted order, then stream it out: runs of equal digits go to consecutive addresses
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t dg = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    const uint32_t pos = s_digit_start[dg] + s_warp_hist[warp][dg] + rank[i];
    s_keys[pos] = key[i];
    s_vals[pos] = val[i];
  }
  __syncthreads();
  const uint32_t tile_n = (uint32_t)((n - base) < (uint64_t)TILE ? (n - base) : (uint64_t)TILE);
  for (uint32_t s = threadIdx.x; s < tile_n; s += THREADS) {
    const uint64_t k = s_keys[s];
    const uint32_t dg = (uint32_t)(k >> shift) & (RADIX - 1);
    const uint32_t dst = s_scatter[dg] + s;
    keys_out[dst] = k;
    vals_out[dst] = s_vals[s];
  }
}

// temp layout (uint32 words): hist[PASSES*RADIX] | counters[PASSES] (+pad to 16) | LOOKBACK rows of "prefix 0" |
// status[tiles*RADIX]
inline size_t temp_words(uint64_t n) { return (size_t)PASSES * RADIX + 16 + (size_t)(LOOKBACK + num_tiles(n)) * RADIX; }

// Sorts (keys_a, vals_a) on the low 8 * passes key bits (keys must be zero above them) using
// (keys_b, vals_b) as the alternate buffer.  The result ends in the `a` buffers when `passes` is even
// and in the `b` buffers when it is odd (*in_b tells which).  Returns the number of kernels launched.
inline int sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n,
                      uint32_t* temp, int sm_count, cudaStream_t stream, int passes = PASSES, bool* in_b = nullptr) {
  if (in_b) *in_b = false;
  if (n == 0) return 0;
  if (passes < 1) passes = 1;
  if (passes > PASSES) passes = PASSES;
  uint32_t* hist = temp;
  uint32_t* counters = temp + PASSES * RADIX;
  uint32_t* status_pad = counters + 16;
  uint32_t* status = status_pad + LOOKBACK * RADIX;
  const uint32_t tiles = num_tiles(n);
  cudaMemsetAsync(temp, 0, sizeof(uint32_t) * (PASSES * RADIX + 16), stream);
  cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DYN_SMEM);
  uint64_t hb = (n + THREADS * 8 - 1) / (THREADS * 8);
  const uint64_t hmax = (uint64_t)sm_count * 8;
  if (hb > hmax) hb = hmax;
  histogram_kernel<<<(unsigned)hb, THREADS, 0, stream>>>(keys_a, n, hist);
  scan_histogram_kernel<<<PASSES, RADIX, 0, stream>>>(hist, status_pad);
  int launches = 2;
  uint64_t *kin = keys_a, *kout = keys_b;
  uint32_t *vin = vals_a, *vout = vals_b;
  for (int p = 0; p < passes; ++p) {
    cudaMemsetAsync(status, 0, sizeof(uint32_t) * (size_t)tiles * RADIX, stream);
    onesweep_kernel<<<tiles, THREADS, DYN_SMEM, stream>>>(kin, vin, kout, vout, n, p * RADIX_BITS, hist + p * RADIX,
                                                          status, counters + p);
    ++launches;
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  if (in_b) *in_b = (passes & 1) != 0;
  return launches;
}

}  // namespace rsort
}  // namespace tknn
