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
#ifndef TKNN_SORT_LOOKBACK
#define TKNN_SORT_LOOKBACK 4
#endif
constexpr int LOOKBACK = TKNN_SORT_LOOKBACK;  // predecessors fetched per look-back step
constexpr uint64_t MAX_N = (1ull << 30) - 1;  // counts share a word with the two flag bits

constexpr size_t DYN_SMEM = (size_t)TILE * (sizeof(uint64_t) + sizeof(uint32_t));
constexpr size_t DYN_SMEM_KEYS = (size_t)TILE * sizeof(uint64_t);

__host__ __device__ inline uint32_t num_tiles(uint64_t n) { return (uint32_t)((n + TILE - 1) / TILE); }

// hist[pass][digit] += count, all passes in one read of the keys.
// Digit p is bits [shift0 + 8p, shift0 + 8p + 8) of the key (shift0 > 0: packed (code, index) keys, sort_keys below).
static __global__ void __launch_bounds__(THREADS) histogram_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                            uint32_t* __restrict__ hist, int shift0, int passes) {
  __shared__ uint32_t s_hist[PASSES * RADIX];
  for (int i = threadIdx.x; i < PASSES * RADIX; i += THREADS) s_hist[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * THREADS;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    const uint64_t key = keys[i] >> shift0;
#pragma unroll
    for (int p = 0; p < PASSES; ++p)
      if (p < passes) atomicAdd(&s_hist[p * RADIX + (uint32_t)((key >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
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
  // diff = lanes whose digit differs from mine in some bit.  Per bit: one predicate (and + setp fuse into LOP3.P), one vote,
  // one select and ONE three-input LOP3 (diff | (ballot ^ mine)).  Written in PTX because the C form — in either the
  // `peers &= bit ? m : ~m` or the `diff |= m ^ -bit` spelling — compiled to 6 instructions per bit (a shift, a mask, a
  // compare and a negate around the vote).
  uint32_t diff = 0;
#pragma unroll
  for (int b = 0; b < RADIX_BITS; ++b) {
    uint32_t m, mine;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "selp.b32 %1, 0xffffffff, 0, p;\n\t"
        "}"
        : "=r"(m), "=r"(mine)
        : "r"(d), "r"(1u << b));
    diff |= m ^ mine;
  }
  return ~diff;
}

// One onesweep pass over the 8-bit digit at bit `shift`.
// PAIRS = true: (u64 key, u32 value) pairs, 24 B moved per pair.  PAIRS = false: keys only (sort_keys: the builder packs
// (curve code, point index) into one u64, the index in the low bits that no pass touches), 16 B moved per key, no value
// registers, a 32 KB instead of a 48 KB tile, 4 instead of 3 resident blocks.
//
// The inter-tile dependency of a single-sweep sort is the look-back: a tile cannot place a key before every preceding tile
// has published its digit counts.  Round 1's kernel published a tile's aggregate only after it had RANKED its keys (half of
// its life) and its inclusive prefix after its own look-back, so tiles waited on work that has nothing to do with the
// counts.  Here the per-warp digit counts are taken first, with shared-memory atomics (order-free: only the totals matter),
// the aggregate is published right behind the load, and the look-back runs BEFORE the ranking: the chain between consecutive
// tiles shrinks to load + count + one window, and everything expensive (ranking, staging, streaming out) is independent
// per tile.  Because the counters then already hold each (warp, digit) run's first slot in the sorted tile, the ranking
// loop writes every key straight to its slot: no rank registers, no second pass over the keys.
// Measured, sort of 10 M / 100 M packed keys, 5 passes (profiles/r2_ab_sortkeys*.jsonl): round 1's order of phases 0.468 / 3.36 ms,
// this kernel 0.441 / 3.66 ms with windows of 8 predecessors, 0.427 / 3.48 ms with windows of 4 (kept), 0.431 / 3.42 with 2,
// 0.466 / 3.73 with 1, 0.476 / 4.02 with 16, 0.554 / 4.84 with 32: the window loads are L2 traffic of the size of the keys
// themselves (a window of 8 rows is 8 KB per tile against a 32 KB tile), so wider windows lose more than they hide.
// 5 resident blocks (51 registers), tiles of 3072 or 2048 keys: slower (0.45, 0.48, 0.57 ms).
#ifndef TKNN_SORT_MINBLOCKS
#define TKNN_SORT_MINBLOCKS 3
#endif
#ifndef TKNN_SORT_MINBLOCKS_KEYS
#define TKNN_SORT_MINBLOCKS_KEYS 4
#endif
template <bool PAIRS>
static __global__ void __launch_bounds__(THREADS, PAIRS ? TKNN_SORT_MINBLOCKS : TKNN_SORT_MINBLOCKS_KEYS)
    onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                          uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint64_t n, int shift,
                          const uint32_t* __restrict__ bin_offset, uint32_t* status, uint32_t* tile_counter) {
  static_assert(THREADS == RADIX, "onesweep_kernel: one digit owner per thread");
  __shared__ uint32_t s_warp_hist[WARPS][RADIX];
  __shared__ uint32_t s_scatter[RADIX];
  __shared__ uint32_t s_wsum[WARPS];
  __shared__ uint32_t s_tile;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_dyn);
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);  // PAIRS only

  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = threadIdx.x; i < WARPS * RADIX; i += THREADS) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;

  uint64_t key[ITEMS];
  uint32_t val[PAIRS ? ITEMS : 1];
  const uint64_t wbase = base + (uint64_t)warp * (32 * ITEMS);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint64_t idx = wbase + i * 32 + lane;
    const bool ok = idx < n;
    key[i] = ok ? keys_in[idx] : ~0ull;  // padding ranks after every real key of digit 255
    if (PAIRS) val[i] = ok ? vals_in[idx] : 0u;
  }
  // the digits, extracted once and kept four to a register; early counts
  static_assert(ITEMS % 4 == 0, "digits are packed four to a register");
  uint32_t dig[ITEMS / 4];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t dg = (uint32_t)(key[i] >> shift) & (RADIX - 1);
    dig[i >> 2] = (i & 3) ? (dig[i >> 2] | (dg << (8 * (i & 3)))) : dg;
    atomicAdd(&s_warp_hist[warp][dg], 1u);
  }
  __syncthreads();

  // thread d owns digit d: tile total -> aggregate, published at once
  const int d = threadIdx.x;
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) total += s_warp_hist[w][d];
  uint32_t* st = status + (size_t)tile * RADIX;
  atomicExch(&st[d], total | (tile == 0 ? FLAG_PREFIX : FLAG_AGG));
  // first look-back window in flight while the tile-local scan runs
  uint32_t v[LOOKBACK];
  int64_t lt = (int64_t)tile - 1;
  {
    const volatile uint32_t* p = status + lt * RADIX + d;  // tile 0 reads the "prefix 0" rows in front of the array
#pragma unroll
    for (int w = 0; w < LOOKBACK; ++w) v[w] = *(p - w * RADIX);
  }
  uint32_t inc = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  uint32_t dstart;
  {
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_wsum[w];
    dstart = woff + inc - total;  // where digit d starts inside the sorted tile
    uint32_t run = dstart;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {  // counter (w, d) := first slot of warp w's run of digit d
      const uint32_t c = s_warp_hist[w][d];
      s_warp_hist[w][d] = run;
      run += c;
    }
  }
  // decoupled look-back (LOOKBACK predecessors per step, consumed in order; a predecessor that has published nothing yet
  // restarts the window at that tile)
  {
    uint32_t excl = 0;
    if (tile > 0) {
      bool done = false, loaded = true;
      while (!done) {
        if (!loaded) {
          const volatile uint32_t* p = status + lt * RADIX + d;
#pragma unroll
          for (int w = 0; w < LOOKBACK; ++w) v[w] = *(p - w * RADIX);
        }
        loaded = false;
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
        lt -= used;
      }
      atomicExch(&st[d], ((excl + total) & VALUE_MASK) | FLAG_PREFIX);
    }
    s_scatter[d] = bin_offset[d] + excl - dstart;  // global slot of tile-local position s is s_scatter[d] + s
  }
  __syncthreads();

  // warp-local stable ranking straight into the sorted tile: lanes holding the same digit find each other (match_digit);
  // every lane reads its digit's (warp, digit) counter, then the lowest lane of each peer group advances it for the group
  // (a read by all lanes and a predicated store instead of a leader's branch, read-modify-write and a shuffle)
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t dg = (dig[i >> 2] >> (8 * (i & 3))) & (RADIX - 1);
    const uint32_t peers = match_digit(dg);
    const uint32_t lower = peers & lt_mask;
    const uint32_t old = s_warp_hist[warp][dg];
    __syncwarp();
    if (lower == 0u) s_warp_hist[warp][dg] = old + __popc(peers);
    const uint32_t pos = old + __popc(lower);
    s_keys[pos] = key[i];
    if (PAIRS) s_vals[pos] = val[i];
    __syncwarp();
  }
  __syncthreads();

  // stream the tile out: runs of equal digits go to consecutive addresses
  const uint32_t tile_n = (uint32_t)((n - base) < (uint64_t)TILE ? (n - base) : (uint64_t)TILE);
  for (uint32_t s = threadIdx.x; s < tile_n; s += THREADS) {
    const uint64_t k = s_keys[s];
    const uint32_t dg = (uint32_t)(k >> shift) & (RADIX - 1);
    const uint32_t dst = s_scatter[dg] + s;
    keys_out[dst] = k;
    if (PAIRS) vals_out[dst] = s_vals[s];
  }
}

// temp layout (uint32 words): hist[PASSES*RADIX] | counters[PASSES] (+pad to 16) | LOOKBACK rows of "prefix 0" |
// status[tiles*RADIX]
inline size_t temp_words(uint64_t n) { return (size_t)PASSES * RADIX + 16 + (size_t)(LOOKBACK + num_tiles(n)) * RADIX; }

// Sorts (keys_a, vals_a) on the key bits [shift0, shift0 + 8 * passes) (stable; keys must agree above them) using
// (keys_b, vals_b) as the alternate buffer.  vals_a == nullptr: keys only (PAIRS = false).  The result ends in the `a`
// buffers when `passes` is even and in the `b` buffers when it is odd (*in_b tells which).  Returns the number of
// kernels launched.
inline int sort_impl(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n, uint32_t* temp,
                     int sm_count, cudaStream_t stream, int shift0, int passes, bool* in_b) {
  if (in_b) *in_b = false;
  if (n == 0) return 0;
  if (passes < 1) passes = 1;
  if (passes > PASSES) passes = PASSES;
  if (shift0 < 0) shift0 = 0;
  if (shift0 + RADIX_BITS * (passes - 1) > 63) passes = (63 - shift0) / RADIX_BITS + 1;
  const bool pairs = vals_a != nullptr;
  uint32_t* hist = temp;
  uint32_t* counters = temp + PASSES * RADIX;
  uint32_t* status_pad = counters + 16;
  uint32_t* status = status_pad + LOOKBACK * RADIX;
  const uint32_t tiles = num_tiles(n);
  cudaMemsetAsync(temp, 0, sizeof(uint32_t) * (PASSES * RADIX + 16), stream);
  const size_t dyn = pairs ? DYN_SMEM : DYN_SMEM_KEYS;
  if (pairs) cudaFuncSetAttribute(onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  else cudaFuncSetAttribute(onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  uint64_t hb = (n + THREADS * 8 - 1) / (THREADS * 8);
  const uint64_t hmax = (uint64_t)sm_count * 8;
  if (hb > hmax) hb = hmax;
  histogram_kernel<<<(unsigned)hb, THREADS, 0, stream>>>(keys_a, n, hist, shift0, passes);
  scan_histogram_kernel<<<PASSES, RADIX, 0, stream>>>(hist, status_pad);
  int launches = 2;
  uint64_t *kin = keys_a, *kout = keys_b;
  uint32_t *vin = vals_a, *vout = vals_b;
  for (int p = 0; p < passes; ++p) {
    cudaMemsetAsync(status, 0, sizeof(uint32_t) * (size_t)tiles * RADIX, stream);
    if (pairs)
      onesweep_kernel<true><<<tiles, THREADS, dyn, stream>>>(kin, vin, kout, vout, n, shift0 + p * RADIX_BITS, hist + p * RADIX,
                                                            status, counters + p);
    else
      onesweep_kernel<false><<<tiles, THREADS, dyn, stream>>>(kin, nullptr, kout, nullptr, n, shift0 + p * RADIX_BITS,
                                                             hist + p * RADIX, status, counters + p);
    ++launches;
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  if (in_b) *in_b = (passes & 1) != 0;
  return launches;
}

inline int sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint64_t n,
                      uint32_t* temp, int sm_count, cudaStream_t stream, int passes = PASSES, bool* in_b = nullptr) {
  return sort_impl(keys_a, vals_a, keys_b, vals_b, n, temp, sm_count, stream, 0, passes, in_b);
}

// Keys only, digits from bit shift0 up: the builder's packed (curve code << index bits | point index) keys — the low
// `shift0` bits (the index) ride along untouched, and because the input is in index order the result equals the stable
// pair sort of (code, index).
inline int sort_keys(uint64_t* keys_a, uint64_t* keys_b, uint64_t n, uint32_t* temp, int sm_count, cudaStream_t stream,
                     int shift0, int passes, bool* in_b = nullptr) {
  return sort_impl(keys_a, nullptr, keys_b, nullptr, n, temp, sm_count, stream, shift0, passes, in_b);
}

}  // namespace rsort
}  // namespace tknn
