// trueknn.cu — the C ABI of libtrueknn (include/trueknn.h) and the host-side round driver.
//
// Host logic replaced (reference file:line):
//   tknn_build   <- hostCode.cpp:165-175,199-212  + owl/UserGeomGroup.cpp:39-241 + owl/UserGeom.cu:172-232
//   tknn_search  <- hostCode.cpp:285-340 (round loop, termination scan, radius doubling, refits)
//   error model  <- owl/helper/cuda.h:22-59 / owl/helper/optix.h:34-55 (throw / exit) -> return codes
#include "../../include/trueknn.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "brute.cuh"
#include "common.cuh"
#include "ctx.cuh"
#include "lbvh.cuh"
#include "radix_sort.cuh"
#include "traverse.cuh"

using namespace tknn;

// ---------------------------------------------------------------------------------------------
// context helpers (declared in ctx.cuh)
// ---------------------------------------------------------------------------------------------
namespace tknn {
namespace host {

int fail(tknn_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    c->err = buf;
  }
  return code;
}

int ensure(tknn_ctx* c, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return TKNN_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
  if (bytes == 0) bytes = 16;
  TK_CUDA(c, cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return TKNN_OK;
}

void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace host
}  // namespace tknn

using namespace tknn::host;

namespace {

// exclusive scan of popc(words[0..nw)) into offsets, total into scalars[SC_TOTAL]
int popc_scan(tknn_ctx* c, const uint32_t* words, uint64_t nw, uint32_t* offsets, int* launches) {
  const uint32_t nb = (uint32_t)((nw + lbvh::SCAN_CHUNK - 1) / lbvh::SCAN_CHUNK);
  TK_TRY(ensure(c, c->block_sums, sizeof(uint32_t) * (size_t)(nb + 1)));
  uint32_t* sums = c->block_sums.as<uint32_t>();
  uint32_t* sc = c->scalars.as<uint32_t>();
  lbvh::popc_reduce_kernel<<<nb, lbvh::THREADS, 0, c->stream>>>(words, nw, sums);
  lbvh::scan_sums_kernel<<<1, lbvh::THREADS, 0, c->stream>>>(sums, nb, sc + SC_TOTAL);
  lbvh::popc_apply_kernel<<<nb, lbvh::THREADS, 0, c->stream>>>(words, nw, sums, offsets);
  if (launches) *launches += 3;
  return TKNN_OK;
}

// ---------------------------------------------------------------------------------------------
// the round driver (shared by search / shard / query / estimator)
// ---------------------------------------------------------------------------------------------
int fetch_words(tknn_ctx* c, const uint32_t* a, int na, const uint32_t* b, int nb, uint32_t* host_out);  // below

struct Job {
  const float4* queries = nullptr;
  const int32_t* self_ids = nullptr;
  const float* query_r2 = nullptr;
  const uint32_t* first_queue = nullptr;  // optional explicit queue for round 1
  uint64_t n_queries = 0;                 // queries in round 1
  uint64_t q_begin = 0;
  int self_is_row = 0;
  int row_mode = 0;
  int k = 0;
  float start_radius = 0.f;  // > 0 or +inf; ignored while r_dev is set
  const float* r_dev = nullptr;  // {r, r * r} on the device (the estimator's output): round 1 reads r2 from there and the
                                 // host learns r with the first read-back it needs anyway
  float* r_host = nullptr;       // receives the radius round 1 ran with
  int squared = 0;
  int32_t* idx_out = nullptr;
  float* dist_out = nullptr;
  int32_t* qid_out = nullptr;
  bool record_stats = true;
  bool force_sparse = false;  // thread-per-query kernel from round 1 (the start-radius sample)
};

template <int MODE>
int launch_traverse(tknn_ctx* c, const trav::Params& P) {
  const int k = MODE == trav::MODE_KNN ? P.k : 0;
  const size_t per_warp = trav::smem_per_warp(k);
  // pick the kernel variant first: its register count enters the block-shape choice
  void (*kern)(const trav::Params) = nullptr;
  const bool ties = c->tie_pruning == 1 || (c->tie_pruning == 0 && c->has_dup_leaves);
  const int variant = MODE != trav::MODE_KNN ? 0 : (ties ? 2 : (c->approx_filter ? 1 : 0));
  const bool hp = MODE == trav::MODE_KNN && k > trav::LIST_MAX_K;
#define TK_PICK(CNT, VAR) (hp ? trav::traverse_kernel<MODE, CNT, VAR, true> : trav::traverse_kernel<MODE, CNT, VAR, false>)
  if (c->counters) kern = variant == 2 ? TK_PICK(true, 2) : variant == 1 ? TK_PICK(true, 1) : TK_PICK(true, 0);
  else kern = variant == 2 ? TK_PICK(false, 2) : variant == 1 ? TK_PICK(false, 1) : TK_PICK(false, 0);
#undef TK_PICK
  cudaFuncAttributes fa;
  TK_CUDA(c, cudaFuncGetAttributes(&fa, kern));
  const int regs_alloc = ((fa.numRegs + 7) / 8) * 8;  // registers are allocated in units of 8 per thread
  // block shape: the warps-per-block that keeps the most warps resident per SM
  // (227 KB shared memory, 64 K registers, 64 warps, 32 blocks)
  const size_t sm_smem = 227 * 1024;
  int warps = 0, bps = 1, best = 0;
  for (int w = 8; w >= 1; --w) {
    const size_t need_b = per_warp * w + 1024;  // + per-block reservation
    if (need_b > sm_smem) continue;
    int b = (int)std::min<size_t>(sm_smem / need_b, (size_t)(64 / w));
    // registers: each of the 4 SM sub-partitions owns 16 K of them and the warps of successive blocks go to
    // the sub-partitions round-robin, so a block fits only while no sub-partition overflows (ncu: at 48
    // registers 224-thread blocks were limited to 5 per SM and 192-thread blocks to 6, not 6 and 7)
    {
      const int per_smsp = 16384 / (regs_alloc * 32);
      int load[4] = {0, 0, 0, 0}, next = 0, fit = 0;
      for (; fit < b; ++fit) {
        int trial[4] = {load[0], load[1], load[2], load[3]};
        for (int i = 0; i < w; ++i) ++trial[(next + i) & 3];
        if (std::max(std::max(trial[0], trial[1]), std::max(trial[2], trial[3])) > per_smsp) break;
        for (int i = 0; i < 4; ++i) load[i] = trial[i];
        next = (next + w) & 3;
      }
      b = fit;
    }
    b = std::min(b, 32);
    if (b * w > best) { best = b * w; warps = w; bps = b; }
  }
  if (warps < 1) return fail(c, TKNN_EINVAL, "k = %d needs %zu B of shared memory per warp", k, per_warp);
  const size_t smem = per_warp * warps;
  if (c->blocks_per_sm > 0) bps = c->blocks_per_sm;
  uint64_t grid = (uint64_t)c->sm_count * bps;
  const uint64_t need = ((uint64_t)P.n_groups + warps - 1) / warps;
  if (grid > need) grid = std::max<uint64_t>(1, need);
  TK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)grid, warps * 32, smem, c->stream>>>(P);
  TK_CUDA(c, cudaGetLastError());
  return TKNN_OK;
}

// thread-per-query variant for rounds whose active set is a small, spatially incoherent remainder
int launch_traverse_sparse(tknn_ctx* c, const trav::Params& P) {
  const size_t smem = trav::sparse_smem(P.k);
  if (smem > 200 * 1024) return TKNN_EINVAL;  // caller falls back to the cooperative kernel
  const unsigned grid = (unsigned)((P.n_active + trav::SPARSE_THREADS - 1) / trav::SPARSE_THREADS);
  if (c->counters) {
    auto kern = trav::traverse_sparse_kernel<true>;
    TK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, trav::SPARSE_THREADS, smem, c->stream>>>(P);
  } else {
    auto kern = trav::traverse_sparse_kernel<false>;
    TK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, trav::SPARSE_THREADS, smem, c->stream>>>(P);
  }
  TK_CUDA(c, cudaGetLastError());
  return TKNN_OK;
}

// warp-per-query (T = 32: tiny rounds, the start-radius sample, small query sets) or team-per-query (T = 4 / 8 / 16: sparse
// rounds) variant
template <int T>
int launch_traverse_team(tknn_ctx* c, const trav::Params& P) {
  const size_t smem = trav::warpq_smem(P.k, T);
  const uint64_t per_block = (uint64_t)trav::WQ_WARPS * (32 / T);
  const unsigned grid = (unsigned)((P.n_active + per_block - 1) / per_block);
  if (P.unresolved)  // the kernel ORs bits into the ballot words
    TK_CUDA(c, cudaMemsetAsync(P.unresolved, 0, sizeof(uint32_t) * (size_t)P.n_groups, c->stream));
  if (c->counters) {
    auto kern = trav::traverse_warp_kernel<true, T>;
    TK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, trav::WQ_WARPS * 32, smem, c->stream>>>(P);
  } else {
    auto kern = trav::traverse_warp_kernel<false, T>;
    TK_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, trav::WQ_WARPS * 32, smem, c->stream>>>(P);
  }
  TK_CUDA(c, cudaGetLastError());
  return TKNN_OK;
}

int launch_traverse_warp(tknn_ctx* c, const trav::Params& P) { return launch_traverse_team<32>(c, P); }

// sparse rounds: teams of c->sparse_team lanes per query, or (0) the thread-per-query kernel
int launch_traverse_sparse_round(tknn_ctx* c, const trav::Params& P) {
  const size_t team_smem = trav::warpq_smem(P.k, std::max(4, c->sparse_team));
  if (c->sparse_team == 0 || team_smem > 200 * 1024) return launch_traverse_sparse(c, P);
  switch (c->sparse_team) {
    case 4: return launch_traverse_team<4>(c, P);
    case 16: return launch_traverse_team<16>(c, P);
    default: return launch_traverse_team<8>(c, P);
  }
}

// One round = one traversal launch over the active queries; unresolved ones are compacted (order preserved) into the
// next round's queue and the radius doubles (hostCode.cpp:285-340 without the refits).
//
// Small searches are latency bound (cfg1: 100 K queries, 0.15 ms of kernels), so the common case makes NO host decision
// between its launches: round 1 reads its radius from the device (the estimator's output), the compaction runs
// unconditionally, and round 2 is launched SPECULATIVELY as the final, unbounded round of one warp per query over at most
// `cap` queue entries with the true count read on the device.  The host then reads {count, radius} once.  Only when more
// than `cap` queries were left (a far too small user radius) does the doubling loop continue for the rest.
int run_rounds(tknn_ctx* c, const Job& job, int* launches_io) {
  uint32_t* sc = c->scalars.as<uint32_t>();
  const uint64_t n0 = job.n_queries;
  const uint32_t g0 = (uint32_t)((n0 + 31) / 32);
  TK_TRY(ensure(c, c->unresolved, sizeof(uint32_t) * (size_t)(g0 + 1)));
  TK_TRY(ensure(c, c->offsets, sizeof(uint32_t) * (size_t)(g0 + 1)));

  const float diag = std::sqrt((c->scene_box[3] - c->scene_box[0]) * (c->scene_box[3] - c->scene_box[0]) +
                               (c->scene_box[4] - c->scene_box[1]) * (c->scene_box[4] - c->scene_box[1]) +
                               (c->scene_box[5] - c->scene_box[2]) * (c->scene_box[5] - c->scene_box[2]));
  float radius = job.r_dev ? 0.0f : job.start_radius;  // with r_dev: known after the first read-back
  bool radius_known = job.r_dev == nullptr;
  uint64_t active = n0;
  const uint32_t* queue = job.first_queue;
  DevBuf* qbuf = nullptr;  // which of queue_a / queue_b holds `queue` (nullptr: the caller's first queue, or none)
  int round = 0;
  int launches = 0;
  const uint64_t spec_cap = std::min<uint64_t>(n0, (uint64_t)std::max(0, c->warp_round_max));
  const bool speculate = c->speculative_max > 0 && n0 <= (uint64_t)c->speculative_max && spec_cap > 0;

  auto fill = [&](trav::Params& P, bool last) {
    std::memset(&P, 0, sizeof(P));
    P.nodes = c->nodes.as<Node>();
    P.node_min_idx = c->node_min_idx.as<int2>();
    P.pts = c->pts.as<float4>();
    P.queries = job.queries;
    P.queue = queue;
    P.self_ids = job.self_ids;
    P.query_r2 = job.query_r2;
    P.n_active = active;
    P.q_begin = job.q_begin;
    P.n_groups = (uint32_t)((active + 31) / 32);
    P.r2 = last ? INFINITY : radius * radius;
    P.k = job.k;
    P.self_is_row = job.self_is_row;
    P.row_mode = job.row_mode;
    P.final_round = last ? 1 : 0;
    P.squared = job.squared;
    P.error = sc + SC_ERROR;
    P.idx_out = job.idx_out;
    P.dist_out = job.dist_out;
    P.qid_out = job.qid_out;
    P.unresolved = last ? nullptr : c->unresolved.as<uint32_t>();
    P.group_counter = sc + SC_GROUP_COUNTER;
    P.counters = c->counters ? reinterpret_cast<unsigned long long*>(sc + SC_COUNTERS) : nullptr;
  };
  auto round_events = [&](int r) -> int {
    while ((int)c->round_ev.size() < 4 * (r + 1)) {
      cudaEvent_t a;
      TK_CUDA(c, cudaEventCreate(&a));
      c->round_ev.push_back(a);
    }
    return TKNN_OK;
  };
  // {unresolved count, radius} in one read-back: the one host decision of a round (hostCode.cpp:310-330)
  auto read_back = [&](uint32_t* count) -> int {
    struct { uint32_t total; float r; } h = {0, 0.f};
    static_assert(sizeof(h) == 2 * sizeof(uint32_t), "mailbox layout");
    TK_TRY(fetch_words(c, sc + SC_TOTAL, 1, reinterpret_cast<const uint32_t*>(job.r_dev), radius_known ? 0 : 1,
                       reinterpret_cast<uint32_t*>(&h)));
    *count = h.total;
    if (!radius_known) { radius = h.r; radius_known = true; if (job.r_host) *job.r_host = h.r; }
    return TKNN_OK;
  };

  while (active > 0) {
    // a radius the host does not know yet (r_dev) is never treated as the last round: the estimator's fallback for a
    // degenerate sample is +inf, and a round at r2 = +inf resolves every query that has k neighbours at all
    const bool last = radius_known && (std::isinf(radius) || radius > 2.0f * diag || round >= TKNN_MAX_ROUNDS - 1);
    trav::Params P;
    fill(P, last);
    if (!radius_known) P.r2_dev = job.r_dev + 1;
    const bool timed = job.record_stats && round < TKNN_MAX_ROUNDS;
    if (timed) {
      TK_TRY(round_events(round));
      TK_CUDA(c, cudaEventRecord(c->round_ev[4 * round], c->stream));
      c->stats.round_queries[round] = active;
    }
    TK_CUDA(c, cudaMemsetAsync(sc + SC_GROUP_COUNTER, 0, sizeof(uint32_t), c->stream));
    if (timed) TK_CUDA(c, cudaEventRecord(c->round_ev[4 * round + 2], c->stream));
    // sparse remainder (< 1/sparse_divisor of the round-1 queries, taken from a queue): one thread per query
    const bool sparse = trav::sparse_smem(job.k) <= 200 * 1024 &&
                        (job.force_sparse || (c->sparse_divisor > 0 && queue != nullptr && round > 0 &&
                                              active * (uint64_t)c->sparse_divisor <= n0));
    // one warp per query: small rounds; four times further for k > 24, where a private heap per THREAD is slow
    // (cfg3, 51 725 leftover queries: 3.6 ms on the thread-per-query kernel, profiles/r2_ab_opts.jsonl)
    const bool tiny = c->warp_round_max > 0 &&
                      active <= (uint64_t)c->warp_round_max * (job.k > trav::LIST_MAX_K ? 4u : 1u);
    if (tiny) TK_TRY(launch_traverse_warp(c, P));
    else if (sparse) TK_TRY(launch_traverse_sparse_round(c, P));
    else TK_TRY(launch_traverse<trav::MODE_KNN>(c, P));
    if (timed) TK_CUDA(c, cudaEventRecord(c->round_ev[4 * round + 3], c->stream));
    ++launches;

    uint32_t next_active = 0;
    if (!last) {
      // order-preserving compaction of the unresolved queries (ballot words + prefix sums)
      TK_TRY(popc_scan(c, c->unresolved.as<uint32_t>(), P.n_groups, c->offsets.as<uint32_t>(), &launches));
      DevBuf& qout = (qbuf == &c->queue_a) ? c->queue_b : c->queue_a;
      if (speculate && round == 0) {
        // ---- no host decision: compact everything, run the final round over <= spec_cap entries, then look ----
        TK_TRY(ensure(c, qout, sizeof(uint32_t) * (size_t)n0));
        trav::compact_queue_kernel<<<blocks_for((uint64_t)P.n_groups * 32, 256), 256, 0, c->stream>>>(
            c->unresolved.as<uint32_t>(), c->offsets.as<uint32_t>(), P.n_groups, queue, job.q_begin, qout.as<uint32_t>());
        ++launches;
        if (timed) TK_CUDA(c, cudaEventRecord(c->round_ev[4 * round + 1], c->stream));
        queue = qout.as<uint32_t>();
        qbuf = &qout;
        active = spec_cap;
        trav::Params F;
        fill(F, true);
        F.n_active_dev = sc + SC_TOTAL;
        const bool timed2 = job.record_stats && round + 1 < TKNN_MAX_ROUNDS;
        if (timed2) {
          TK_TRY(round_events(round + 1));
          TK_CUDA(c, cudaEventRecord(c->round_ev[4 * (round + 1)], c->stream));
          TK_CUDA(c, cudaEventRecord(c->round_ev[4 * (round + 1) + 2], c->stream));
        }
        TK_TRY(launch_traverse_warp(c, F));
        ++launches;
        if (timed2) {
          TK_CUDA(c, cudaEventRecord(c->round_ev[4 * (round + 1) + 3], c->stream));
          TK_CUDA(c, cudaEventRecord(c->round_ev[4 * (round + 1) + 1], c->stream));
        }
        TK_TRY(read_back(&next_active));
        ++round;  // round 1 is done
        if (next_active > 0) {
          if (timed2) c->stats.round_queries[round] = std::min<uint64_t>(next_active, spec_cap);
          ++round;  // ... and so is the speculative final round
        }
        radius *= 2.0f;
        if (next_active <= spec_cap) { active = 0; break; }
        // more queries were left than the speculative round covered: the doubling loop takes the rest of the queue
        queue = queue + spec_cap;
        active = next_active - spec_cap;
        continue;
      }
      TK_TRY(read_back(&next_active));
      if (next_active > 0) {
        TK_TRY(ensure(c, qout, sizeof(uint32_t) * (size_t)next_active));
        trav::compact_queue_kernel<<<blocks_for((uint64_t)P.n_groups * 32, 256), 256, 0, c->stream>>>(
            c->unresolved.as<uint32_t>(), c->offsets.as<uint32_t>(), P.n_groups, queue, job.q_begin, qout.as<uint32_t>());
        ++launches;
        queue = qout.as<uint32_t>();
        qbuf = &qout;
      }
    }
    if (timed) TK_CUDA(c, cudaEventRecord(c->round_ev[4 * round + 1], c->stream));
    ++round;
    active = next_active;
    if (!last) radius *= 2.0f;
  }
  if (!radius_known && job.r_host) {  // a single final round ran on a device radius: fetch it for the statistics
    uint32_t dummy = 0;
    TK_TRY(read_back(&dummy));
  }
  if (job.record_stats) {
    c->stats.rounds = round;
    c->stats.final_radius = radius;
  }
  if (launches_io) *launches_io += launches;
  return TKNN_OK;
}

// Sampled k-th-neighbour distance -> start radius (role of Util/random_sample.py:5-32), entirely on the device: the
// sample queue is generated by a kernel, the sampled queries run one unbounded final round (one warp per query), and a
// one-block radix select leaves {r, r * r} at sc[SC_QUANTILE].  Nothing is copied and nothing waits: round 1 reads r2
// from there (Job::r_dev).
int estimate_radius_device(tknn_ctx* c, const float4* queries, uint64_t q_begin, uint64_t nq, const int32_t* self_ids,
                           int self_is_row, int k, int* launches) {
  // sample_groups runs of 32 consecutive sorted positions: neighbouring queries walk nearly the same nodes and leaves
  const uint64_t groups = (nq + 31) / 32;
  const uint64_t sg = std::min<uint64_t>((uint64_t)std::max(1, c->sample_groups), groups);
  uint64_t m = sg * 32;  // only the last group of the range can be short, and only the last run can be that group
  if (groups * (sg - 1) / sg == groups - 1 && (nq & 31)) m -= 32 - (nq & 31);
  TK_TRY(ensure(c, c->sample, m * sizeof(uint32_t) + m * (size_t)k * (sizeof(int32_t) + sizeof(float))));
  uint32_t* dq = c->sample.as<uint32_t>();
  int32_t* di = reinterpret_cast<int32_t*>(dq + m);
  float* dd = reinterpret_cast<float*>(di + m * (size_t)k);
  trav::sample_queue_kernel<<<blocks_for(sg * 32, 256), 256, 0, c->stream>>>(groups, (uint32_t)sg, q_begin, nq, dq);
  Job job;
  job.queries = queries;
  job.self_ids = self_ids;
  job.first_queue = dq;
  job.n_queries = m;
  job.q_begin = q_begin;
  job.self_is_row = self_is_row;
  job.row_mode = 2;
  job.k = k;
  job.start_radius = INFINITY;
  job.squared = 0;
  job.idx_out = di;
  job.dist_out = dd;
  job.record_stats = false;
  job.force_sparse = true;
  const int saved = c->counters;
  c->counters = 0;
  int rc = run_rounds(c, job, launches);
  c->counters = saved;
  TK_TRY(rc);
  const uint32_t pos = (uint32_t)((double)(m - 1) * std::min(1000, std::max(0, c->radius_quantile)) / 1000.0);
  trav::radius_quantile_kernel<<<1, 1024, 0, c->stream>>>(dd, (uint32_t)m, k, pos,
                                                          reinterpret_cast<float*>(c->scalars.as<uint32_t>() + SC_QUANTILE));
  TK_CUDA(c, cudaGetLastError());
  if (launches) *launches += 2;
  return TKNN_OK;
}

int auto_morton_bits(uint64_t n) {
  int lg = 0;
  while (((uint64_t)1 << lg) < n) ++lg;
  return std::min(21, std::max(10, (lg + 2) / 3 + 8));
}

// Key layout of a build (TKNN_OPT_SORT_MODE, TKNN_OPT_MORTON_BITS).  Packed: the curve code sits above the point's index
// in one u64 and the sort moves keys only.  The code gets as many bits per axis as fit beside the index, at most 13
// (39 bits: five 8-bit passes) — 13 bits at 10 M points are cells 32x finer per axis than the mean point spacing, 12 bits
// at 100 M points 9x — and one bit per axis less where that saves a whole pass and still leaves log2(n)/3 + 4.  Points
// that share a cell are ordered by index: tree quality at the scale of one cell, never exactness.
// *idx_bits == 0: pair sort.
void key_layout(uint64_t n, int morton_bits, int sort_mode, int* mbits, int* idx_bits) {
  int lg = 0;
  while (((uint64_t)1 << lg) < n) ++lg;
  const int ib = std::max(1, lg), spacing = (lg + 2) / 3, fit = (64 - ib) / 3;
  *idx_bits = 0;
  if (morton_bits > 0) {
    *mbits = morton_bits;
    if (sort_mode == 0 && 3 * morton_bits + ib <= 64) *idx_bits = ib;
    return;
  }
  if (sort_mode == 0 && fit >= spacing + 3) {
    int b = std::min(13, fit);
    const int spare = 3 * b - 8 * ((3 * b - 1) / 8);  // code bits in the last, partial pass
    if (spare <= 3 && b - 1 >= spacing + 4) --b;
    *mbits = b;
    *idx_bits = ib;
    return;
  }
  *mbits = auto_morton_bits(n);
}

void choose_key_layout(const tknn_ctx* c, uint64_t n, int* mbits, int* idx_bits) {
  key_layout(n, c->morton_bits, c->sort_mode, mbits, idx_bits);
}

// Hilbert levels of the key kernel (lbvh.cuh: morton_kernel): two levels below the one where a cell holds one point
int hilbert_levels(uint64_t n, int bits, int forced = 0) {
  if (forced > 0) return std::max(2, std::min(bits, forced));
  int lg = 0;
  while (((uint64_t)1 << lg) < n) ++lg;
  return std::max(2, std::min(bits, (lg + 2) / 3 + 2));
}

int check_ctx(tknn_ctx* c) { return c ? TKNN_OK : TKNN_EINVAL; }

// out[0 .. na) = a[0 .. na), out[na .. na + nb) = b[0 .. nb): the mailbox writer (ctx.cuh)
static __global__ void mailbox_kernel(const uint32_t* __restrict__ a, int na, const uint32_t* __restrict__ b, int nb,
                                      volatile uint32_t* out) {
  for (int i = 0; i < na; ++i) out[i] = a[i];
  for (int i = 0; i < nb; ++i) out[na + i] = b[i];
  __threadfence_system();
}

// Brings na + nb (<= 16) device words to the host and waits for the stream.  With host outputs the bulk result copies of
// the previous slice occupy the device->host copy engine for milliseconds; a 4-byte cudaMemcpyAsync on the compute stream
// would queue behind them and stall the round loop (measured: cfg2's four file-order slices searched in 15.2 ms with
// idx + dist copies in flight against 12.6 ms with half the bytes).  A kernel's stores to mapped host memory do not.
int fetch_words(tknn_ctx* c, const uint32_t* a, int na, const uint32_t* b, int nb, uint32_t* host_out) {
  if (na + nb > 16 || na < 0 || nb < 0) return fail(c, TKNN_EINVAL, "internal: mailbox overflow");
  if (c->mailbox_h) {
    mailbox_kernel<<<1, 1, 0, c->stream>>>(a, na, b, nb, c->mailbox_d);
    TK_CUDA(c, cudaGetLastError());
    ++c->mailbox_launches;
    TK_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < na + nb; ++i) host_out[i] = c->mailbox_h[i];
    return TKNN_OK;
  }
  if (na) TK_CUDA(c, cudaMemcpyAsync(host_out, a, na * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (nb) TK_CUDA(c, cudaMemcpyAsync(host_out + na, b, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  TK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TKNN_OK;
}

int read_error_flag(tknn_ctx* c) {
  uint32_t e[2] = {0, 0};  // SC_ERROR, SC_QBAD
  TK_TRY(fetch_words(c, c->scalars.as<uint32_t>() + SC_ERROR, 2, nullptr, 0, e));
  if (e[0]) return fail(c, TKNN_ECUDA, "traversal stack overflow (BVH deeper than %d)", STACK_DEPTH);
  if (e[1]) return fail(c, TKNN_EINVAL, "non-finite coordinate in the query points");
  return TKNN_OK;
}

int collect_counters(tknn_ctx* c) {
  if (!c->counters) return TKNN_OK;
  unsigned long long h[8];
  TK_CUDA(c, cudaMemcpyAsync(h, c->scalars.as<uint32_t>() + SC_COUNTERS, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  TK_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stats.nodes_visited = h[0];
  c->stats.points_tested = h[1];
  c->stats.heap_inserts = h[2];
  c->stats.warp_node_visits = h[3];
  c->stats.warp_leaf_visits = h[4];
  c->stats.warp_point_loads = h[5];
  c->stats.filter_violations = h[6];
  return TKNN_OK;
}

int finish_round_stats(tknn_ctx* c) {
  for (int r = 0; r < c->stats.rounds && r < TKNN_MAX_ROUNDS; ++r) {
    float ms = 0.f;
    TK_CUDA(c, cudaEventElapsedTime(&ms, c->round_ev[4 * r], c->round_ev[4 * r + 1]));
    c->stats.round_ms[r] = ms;
    TK_CUDA(c, cudaEventElapsedTime(&ms, c->round_ev[4 * r + 2], c->round_ev[4 * r + 3]));
    c->stats.kernel_ms[r] = ms;
  }
  return TKNN_OK;
}

void reset_search_stats(tknn_ctx* c) {
  tknn_stats& s = c->stats;
  s.n_queries = 0; s.k = 0; s.rounds = 0; s.start_radius = s.final_radius = 0.f;
  s.estimate_ms = s.search_ms = s.d2h_ms = 0.f;
  std::memset(s.round_ms, 0, sizeof(s.round_ms));
  std::memset(s.kernel_ms, 0, sizeof(s.kernel_ms));
  std::memset(s.round_queries, 0, sizeof(s.round_queries));
  s.kernel_launches = 0;
  s.nodes_visited = s.points_tested = s.heap_inserts = 0;
  s.warp_node_visits = s.warp_leaf_visits = s.warp_point_loads = s.filter_violations = 0;
  s.d2h_bytes = 0;
  c->mailbox_launches = 0;
}

}  // namespace

// all-points search over query positions [q_begin, q_begin + nq) of the sorted order
int tknn::host::search_range(tknn_ctx* c, int k, float start_radius, uint64_t q_begin, uint64_t nq, int row_mode, int32_t* qid_out,
                 int32_t* idx_out, float* dist_out, uint64_t rows) {
  if (c->n == 0) return fail(c, TKNN_ESTATE, "tknn_search before tknn_build");
  if (k < 1 || k > TKNN_MAX_K) return fail(c, TKNN_EINVAL, "k = %d outside [1, %d]", k, TKNN_MAX_K);
  if ((uint64_t)k > c->n - 1) return fail(c, TKNN_EINVAL, "k = %d > n - 1 = %llu: fewer than k neighbours exist", k,
                                          (unsigned long long)(c->n - 1));
  if (std::isnan(start_radius)) return fail(c, TKNN_EINVAL, "start_radius is NaN");
  if (!idx_out) return fail(c, TKNN_EINVAL, "null output array");
  if (c->ids_in_w && row_mode == 0)
    return fail(c, TKNN_ESTATE, "this BVH carries caller-chosen point ids (point-partitioned build): rows in build order do not exist");
  reset_search_stats(c);
  c->stats.n_queries = nq;
  c->stats.k = k;
  uint32_t* sc = c->scalars.as<uint32_t>();

  // dist_out == nullptr: indices only — the distances stay in device scratch and are never copied (half the
  // device->host bytes of a host-output call; a distance is recomputable from its index)
  const bool idx_dev = is_device_ptr(idx_out), dist_dev = !dist_out || is_device_ptr(dist_out);
  const bool qid_dev = qid_out ? is_device_ptr(qid_out) : true;
  int32_t* d_idx = idx_out;
  float* d_dist = dist_out;
  int32_t* d_qid = qid_out;
  const size_t out_elems = (size_t)rows * (size_t)k;
  if (!idx_dev || (qid_out && !qid_dev)) {
    TK_TRY(ensure(c, c->stage_idx, out_elems * sizeof(int32_t) + (qid_out ? rows * sizeof(int32_t) : 0)));
    if (!idx_dev) d_idx = c->stage_idx.as<int32_t>();
    if (qid_out && !qid_dev) d_qid = c->stage_idx.as<int32_t>() + out_elems;
  }
  if (!dist_dev || !dist_out) {
    TK_TRY(ensure(c, c->stage_dist, out_elems * sizeof(float)));
    d_dist = c->stage_dist.as<float>();
  }

  TK_CUDA(c, cudaMemsetAsync(sc + SC_ERROR, 0, 2 * sizeof(uint32_t), c->stream));
  TK_CUDA(c, cudaMemsetAsync(sc + SC_COUNTERS, 0, 8 * sizeof(unsigned long long), c->stream));
  TK_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  int launches = 0;
  float r0 = start_radius;
  const bool estimated = nq > 0 && !(r0 > 0.0f);
  if (estimated) TK_TRY(estimate_radius_device(c, c->pts.as<float4>(), q_begin, nq, nullptr, 1, k, &launches));
  TK_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  c->stats.start_radius = r0;

  Job job;
  if (estimated) {
    job.r_dev = reinterpret_cast<const float*>(sc + SC_QUANTILE);
    job.r_host = &c->stats.start_radius;
  }
  job.queries = c->pts.as<float4>();
  job.n_queries = nq;
  job.q_begin = q_begin;
  job.self_is_row = 1;
  job.row_mode = row_mode;
  job.k = k;
  job.start_radius = r0;
  job.squared = c->squared;
  job.idx_out = d_idx;
  job.dist_out = d_dist;
  job.qid_out = d_qid;

  // Compact rows (row_mode 1) are in Morton order, so a contiguous slice of the queries yields a
  // contiguous, FINAL slice of the output: with host outputs the shard is searched in chunks and each
  // chunk's device->host copy (copy stream) overlaps the search of the next one (compute stream).
  const uint64_t groups = (nq + 31) / 32;
  const bool host_out = !idx_dev || !dist_dev || (qid_out && !qid_dev);
  const int chunks = (row_mode == 1 && host_out && c->output_chunks > 1 && nq >= (1u << 18))
                         ? (int)std::min<uint64_t>((uint64_t)c->output_chunks, groups) : 1;
  // file-order slices: the device->host link moves the rows of a slice about as fast as a slice of that sparsity is
  // searched, so the step ends one slice-search after the link could have started: 5 slices with distances (cfg2: 21.2 /
  // 20.6 / 20.3 / 20.9 / 21.8 ms end to end with 3 / 4 / 5 / 6 / 8), 2 when only indices leave the device (17.9 / 16.5 / 16.9
  // ms with 1 / 2 / 3)
  // The heap kernel (k > 24) pays far more for a sparse slice (cfg3, k = 64: halves cost 1.7x per query, quarters 2.5x), so
  // there the search is the slower side from three slices on: 140.2 / 136.0 / 136.5 / 144.6 ms with 1 / 2 / 3 / 4 slices, and
  // 95.3 / 106.3 / 121.5 ms indices-only with 1 / 2 / 3 (profiles/r2_e2e_cfg3.txt).
  const bool heap_k = k > trav::LIST_MAX_K;
  const int fo_chunks = c->file_order_chunks > 0 ? c->file_order_chunks : (heap_k ? (dist_out ? 2 : 1) : (dist_out ? 5 : 2));
  if (chunks > 1) {
    while ((int)c->chunk_ev.size() < chunks) {
      cudaEvent_t e;
      TK_CUDA(c, cudaEventCreate(&e));
      c->chunk_ev.push_back(e);
    }
    for (int ci = 0; ci < chunks; ++ci) {
      const uint64_t g0 = groups * (uint64_t)ci / chunks, g1 = groups * (uint64_t)(ci + 1) / chunks;
      const uint64_t r_lo = g0 * 32, r_hi = std::min<uint64_t>(nq, g1 * 32);  // rows of this chunk
      if (r_hi <= r_lo) continue;
      Job cj = job;
      cj.q_begin = q_begin + r_lo;
      cj.n_queries = r_hi - r_lo;
      cj.idx_out = d_idx + r_lo * (uint64_t)k;
      cj.dist_out = d_dist + r_lo * (uint64_t)k;
      cj.qid_out = d_qid ? d_qid + r_lo : nullptr;
      cj.record_stats = (ci == 0);  // per-round figures describe the first chunk
      TK_TRY(run_rounds(c, cj, &launches));
      TK_CUDA(c, cudaEventRecord(c->chunk_ev[ci], c->stream));
      TK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[ci], 0));
      const size_t e0 = (size_t)r_lo * k, ne = (size_t)(r_hi - r_lo) * k;
      if (!idx_dev) {
        TK_CUDA(c, cudaMemcpyAsync(idx_out + e0, d_idx + e0, ne * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
        c->stats.d2h_bytes += ne * sizeof(int32_t);
      }
      if (!dist_dev) {
        TK_CUDA(c, cudaMemcpyAsync(dist_out + e0, d_dist + e0, ne * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
        c->stats.d2h_bytes += ne * sizeof(float);
      }
      if (qid_out && !qid_dev) {
        TK_CUDA(c, cudaMemcpyAsync(qid_out + r_lo, d_qid + r_lo, (r_hi - r_lo) * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                   c->copy_stream));
        c->stats.d2h_bytes += (r_hi - r_lo) * sizeof(int32_t);
      }
    }
    TK_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    // the caller's stream must not run ahead of the copies
    TK_CUDA(c, cudaEventRecord(c->chunk_ev[0], c->copy_stream));
    TK_CUDA(c, cudaStreamWaitEvent(c->stream, c->chunk_ev[0], 0));
  } else if (row_mode == 0 && host_out && fo_chunks > 1 && nq == c->n && nq >= (1u << 18)) {
    // File-order rows: slice the queries by ORIGINAL index so that each slice's rows are one contiguous,
    // final block of the output.  A slice is every C-th point of the cloud in Morton order, so its groups
    // are C times less dense in space (costlier per query) — the price of overlapping the result copy.
    const int fc = fo_chunks;
    while ((int)c->chunk_ev.size() < fc) {
      cudaEvent_t e;
      TK_CUDA(c, cudaEventCreate(&e));
      c->chunk_ev.push_back(e);
    }
    TK_TRY(ensure(c, c->chunk_queue, (size_t)nq * sizeof(uint32_t)));
    TK_TRY(ensure(c, c->unresolved, sizeof(uint32_t) * (size_t)(groups + 1)));
    TK_TRY(ensure(c, c->offsets, sizeof(uint32_t) * (size_t)(groups + 1)));
    for (int ci = 0; ci < fc; ++ci) {
      const uint64_t lo = nq * (uint64_t)ci / fc, hi = nq * (uint64_t)(ci + 1) / fc;
      if (hi <= lo) continue;
      uint32_t* cq = c->chunk_queue.as<uint32_t>() + lo;
      trav::index_range_flag_kernel<<<blocks_for(nq, 256), 256, 0, c->stream>>>(c->pts.as<float4>(), nq, (uint32_t)lo,
                                                                               (uint32_t)hi, c->unresolved.as<uint32_t>());
      TK_TRY(popc_scan(c, c->unresolved.as<uint32_t>(), groups, c->offsets.as<uint32_t>(), &launches));
      trav::compact_queue_kernel<<<blocks_for(groups * 32, 256), 256, 0, c->stream>>>(
          c->unresolved.as<uint32_t>(), c->offsets.as<uint32_t>(), (uint32_t)groups, nullptr, 0, cq);
      launches += 2;
      Job cj = job;
      cj.first_queue = cq;
      cj.n_queries = hi - lo;
      cj.record_stats = (ci == 0);
      TK_TRY(run_rounds(c, cj, &launches));
      TK_CUDA(c, cudaEventRecord(c->chunk_ev[ci], c->stream));
      TK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[ci], 0));
      const size_t e0 = (size_t)lo * k, ne = (size_t)(hi - lo) * k;
      if (!idx_dev) {
        TK_CUDA(c, cudaMemcpyAsync(idx_out + e0, d_idx + e0, ne * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
        c->stats.d2h_bytes += ne * sizeof(int32_t);
      }
      if (!dist_dev) {
        TK_CUDA(c, cudaMemcpyAsync(dist_out + e0, d_dist + e0, ne * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
        c->stats.d2h_bytes += ne * sizeof(float);
      }
    }
    TK_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    TK_CUDA(c, cudaEventRecord(c->chunk_ev[0], c->copy_stream));
    TK_CUDA(c, cudaStreamWaitEvent(c->stream, c->chunk_ev[0], 0));
  } else {
    TK_TRY(run_rounds(c, job, &launches));
    TK_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    if (!idx_dev) {
      TK_CUDA(c, cudaMemcpyAsync(idx_out, d_idx, out_elems * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      c->stats.d2h_bytes += out_elems * sizeof(int32_t);
    }
    if (!dist_dev) {
      TK_CUDA(c, cudaMemcpyAsync(dist_out, d_dist, out_elems * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      c->stats.d2h_bytes += out_elems * sizeof(float);
    }
    if (qid_out && !qid_dev) {
      TK_CUDA(c, cudaMemcpyAsync(qid_out, d_qid, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      c->stats.d2h_bytes += rows * sizeof(int32_t);
    }
  }
  TK_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  TK_CUDA(c, cudaStreamSynchronize(c->stream));
  TK_CUDA(c, cudaEventElapsedTime(&c->stats.estimate_ms, c->ev[0], c->ev[1]));
  TK_CUDA(c, cudaEventElapsedTime(&c->stats.search_ms, c->ev[0], c->ev[2]));
  TK_CUDA(c, cudaEventElapsedTime(&c->stats.d2h_ms, c->ev[2], c->ev[3]));
  if (start_radius > 0.0f) c->stats.estimate_ms = 0.f;
  c->stats.kernel_launches = (uint32_t)(launches + c->mailbox_launches + (c->mailbox_h ? 1 : 0));  // + the error-flag fetch below
  TK_TRY(finish_round_stats(c));
  TK_TRY(collect_counters(c));
  TK_TRY(read_error_flag(c));
  return TKNN_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int tknn_version(void) { return TKNN_VERSION; }

const char* tknn_last_error(const tknn_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int tknn_create(int device, tknn_ctx** out) {
  if (!out) return TKNN_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) { cudaGetLastError(); return TKNN_ECUDA; }  // no CPU fallback
  if (device < 0 || device >= count) return TKNN_EINVAL;
  tknn_ctx* c = new (std::nothrow) tknn_ctx();
  if (!c) return TKNN_ENOMEM;
  c->device = device;
  ScopedDevice sd(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return TKNN_ECUDA; }
  if (prop.major < 10) { delete c; return TKNN_ECUDA; }  // sm_100a code only
  c->sm_count = prop.multiProcessorCount;
  c->l2_bytes = (size_t)prop.l2CacheSize;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return TKNN_ECUDA; }
  c->stream = c->own_stream;
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return TKNN_ECUDA; }
  for (auto& ev : c->ev)
    if (cudaEventCreate(&ev) != cudaSuccess) { delete c; return TKNN_ECUDA; }
  if (cudaMalloc(&c->scalars.p, SC_WORDS * sizeof(uint32_t)) != cudaSuccess) { delete c; return TKNN_ENOMEM; }
  c->scalars.bytes = SC_WORDS * sizeof(uint32_t);
  {  // the mailbox is an optimisation: without mapped memory fetch_words falls back to small copies
    void* h = nullptr;
    void* d = nullptr;
    if (cudaHostAlloc(&h, 16 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
        cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
      c->mailbox_h = static_cast<uint32_t*>(h);
      c->mailbox_d = static_cast<uint32_t*>(d);
    } else {
      if (h) cudaFreeHost(h);
      cudaGetLastError();
    }
  }
  *out = c;
  return TKNN_OK;
}

int tknn_destroy(tknn_ctx* c) {
  if (!c) return TKNN_EINVAL;
  ScopedDevice sd(c->device);
  cudaStreamSynchronize(c->stream);
  tknn_internal_free_dist(c);
  for (DevBuf* b : {&c->pts, &c->nodes, &c->leaf_start, &c->node_min_idx, &c->queue_a, &c->queue_b, &c->unresolved, &c->offsets,
                    &c->block_sums, &c->scalars, &c->stage_idx, &c->stage_dist, &c->sample, &c->chunk_queue, &c->b_in, &c->b_keys_a,
                    &c->b_keys_b, &c->b_vals_a, &c->b_vals_b, &c->b_sort_tmp, &c->b_delta, &c->b_ballots, &c->b_leaf_key,
                    &c->b_child_info, &c->b_parent_leaf, &c->b_parent_node, &c->b_arrive, &c->q_stage, &c->q_keys_a, &c->q_keys_b,
                    &c->q_vals_a, &c->q_vals_b, &c->q_sort_tmp, &c->q_pts, &c->q_sid_stage, &c->q_sid_sorted, &c->q_rad_stage,
                    &c->q_r2_sorted})
    release(*b);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->round_ev) cudaEventDestroy(ev);
  for (auto& ev : c->chunk_ev) cudaEventDestroy(ev);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->mailbox_h) cudaFreeHost(c->mailbox_h);
  delete c;
  return TKNN_OK;
}

int tknn_set_stream(tknn_ctx* c, void* cuda_stream) {
  TK_TRY(check_ctx(c));
  c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return TKNN_OK;
}

int tknn_set_option(tknn_ctx* c, int key, int64_t value) {
  TK_TRY(check_ctx(c));
  switch (key) {
    case TKNN_OPT_LEAF_SIZE:
      if (value < 2 || value > MAX_LEAF) return fail(c, TKNN_EINVAL, "leaf size %lld outside [2, %d]", (long long)value, MAX_LEAF);
      c->leaf_size = (int)value;
      return TKNN_OK;
    case TKNN_OPT_COUNTERS: c->counters = value ? 1 : 0; return TKNN_OK;
    case TKNN_OPT_LEAF_POLICY:
      if (value != 0 && value != 1) return fail(c, TKNN_EINVAL, "leaf policy must be 0 or 1");
      c->leaf_policy = (int)value;
      return TKNN_OK;
    case TKNN_OPT_SAMPLE_GROUPS:
      if (value < 1 || value > 65536) return fail(c, TKNN_EINVAL, "sample groups outside [1, 65536]");
      c->sample_groups = (int)value;
      return TKNN_OK;
    case TKNN_OPT_BLOCKS_PER_SM:
      if (value < 0 || value > 32) return fail(c, TKNN_EINVAL, "blocks per SM outside [0, 32]");
      c->blocks_per_sm = (int)value;
      return TKNN_OK;
    case TKNN_OPT_SQUARED_DIST: c->squared = value ? 1 : 0; return TKNN_OK;
    case TKNN_OPT_KEEP_SCRATCH: c->keep_scratch = value ? 1 : 0; return TKNN_OK;
    case TKNN_OPT_APPROX_FILTER: c->approx_filter = value ? 1 : 0; return TKNN_OK;
    case TKNN_OPT_TIE_PRUNING:
      if (value < 0 || value > 2) return fail(c, TKNN_EINVAL, "tie pruning must be 0 (auto), 1 (on) or 2 (off)");
      c->tie_pruning = (int)value;
      return TKNN_OK;
    case TKNN_OPT_MORTON_BITS:
      if (value != 0 && (value < 4 || value > 21)) return fail(c, TKNN_EINVAL, "morton bits per axis must be 0 (auto) or in [4, 21]");
      c->morton_bits = (int)value;
      return TKNN_OK;
    case TKNN_OPT_SPECULATIVE_MAX:
      if (value < 0 || value > (int64_t)1 << 30) return fail(c, TKNN_EINVAL, "speculative maximum outside [0, 2^30]");
      c->speculative_max = (int)value;
      return TKNN_OK;
    case TKNN_OPT_SORT_MODE:
      if (value != 0 && value != 1) return fail(c, TKNN_EINVAL, "sort mode must be 0 (packed keys when they fit) or 1 (pairs)");
      c->sort_mode = (int)value;
      return TKNN_OK;
    case TKNN_OPT_SPARSE_TEAM:
      if (value != 0 && value != 4 && value != 8 && value != 16) return fail(c, TKNN_EINVAL, "sparse team must be 0, 4, 8 or 16 lanes");
      c->sparse_team = (int)value;
      return TKNN_OK;
    case TKNN_OPT_CURVE:
      if (value != 0 && value != 1 && !(value >= 101 && value <= 121))
        return fail(c, TKNN_EINVAL, "curve must be 0 (Hilbert), 1 (Morton) or 100 + L (Hilbert on exactly L levels)");
      c->curve = value == 1 ? 0 : 1;  // internal: 1 = Hilbert
      c->curve_levels = value > 100 ? (int)value - 100 : 0;
      return TKNN_OK;
    case TKNN_OPT_FILE_ORDER_CHUNKS:
      if (value < 0 || value > 64) return fail(c, TKNN_EINVAL, "file-order chunks outside [0, 64]");
      c->file_order_chunks = (int)value;
      return TKNN_OK;
    case TKNN_OPT_OUTPUT_CHUNKS:
      if (value < 1 || value > 64) return fail(c, TKNN_EINVAL, "output chunks outside [1, 64]");
      c->output_chunks = (int)value;
      return TKNN_OK;
    case TKNN_OPT_SPARSE_DIVISOR:
      if (value < 0 || value > 1000000) return fail(c, TKNN_EINVAL, "sparse divisor outside [0, 1e6]");
      c->sparse_divisor = (int)value;
      return TKNN_OK;
    case TKNN_OPT_WARP_ROUND_MAX:
      if (value < 0 || value > (int64_t)1 << 30) return fail(c, TKNN_EINVAL, "warp round maximum outside [0, 2^30]");
      c->warp_round_max = (int)value;
      return TKNN_OK;
    case TKNN_OPT_RADIUS_QUANTILE:
      if (value < 0 || value > 1000) return fail(c, TKNN_EINVAL, "radius quantile outside [0, 1000] per mille");
      c->radius_quantile = (int)value;
      return TKNN_OK;
    default: return fail(c, TKNN_EINVAL, "unknown option %d", key);
  }
}

int tknn_build(tknn_ctx* c, const float* xyz, uint64_t n, int dim, int stride_floats) {
  tknn_internal_dist_invalidate(c);
  return build_core(c, xyz, n, dim, stride_floats, false);
}

}  // extern "C"

int tknn::host::build_core(tknn_ctx* c, const float* xyz, uint64_t n, int dim, int stride_floats, bool ids_in_w) {
  TK_TRY(check_ctx(c));
  if (!xyz) return fail(c, TKNN_EINVAL, "null point array");
  if (ids_in_w && (dim != 3 || stride_floats < 4)) return fail(c, TKNN_EINVAL, "ids in w need dim 3 and stride >= 4");
  if (dim != 2 && dim != 3) return fail(c, TKNN_EINVAL, "dim = %d (must be 2 or 3, hostCode.cpp:114-124)", dim);
  if (stride_floats < dim) return fail(c, TKNN_EINVAL, "stride %d < dim %d", stride_floats, dim);
  if (n < 2) return fail(c, TKNN_EINVAL, "need at least 2 points (got %llu)", (unsigned long long)n);
  if (n > rsort::MAX_N) return fail(c, TKNN_EINVAL, "n = %llu exceeds %llu points per device", (unsigned long long)n,
                                    (unsigned long long)rsort::MAX_N);
  ScopedDevice sd(c->device);
  cudaStream_t st = c->stream;
  c->n = 0;
  c->n_leaves = 0;
  tknn_stats& S = c->stats;
  std::memset(&S, 0, sizeof(S));
  c->mailbox_launches = 0;
  int launches = 0;

  // ---- stage the input ----
  DevBuf &in_stage = c->b_in, &keys_a = c->b_keys_a, &keys_b = c->b_keys_b, &vals_a = c->b_vals_a, &vals_b = c->b_vals_b,
         &sort_tmp = c->b_sort_tmp, &delta = c->b_delta, &ballots = c->b_ballots, &leaf_key = c->b_leaf_key,
         &child_info = c->b_child_info, &parent_leaf = c->b_parent_leaf, &parent_node = c->b_parent_node, &arrive = c->b_arrive;
  auto cleanup = [&]() {
    if (c->keep_scratch) return;
    for (DevBuf* b : {&in_stage, &keys_a, &keys_b, &vals_a, &vals_b, &sort_tmp, &delta, &ballots, &leaf_key, &child_info,
                      &parent_leaf, &parent_node, &arrive})
      release(*b);
  };
  // size-only allocations happen before the timed region (cudaMalloc is a synchronous host call)
  {
    const uint64_t nw0 = (n + 31) / 32;
    int rc0 = TKNN_OK;
    if (!is_device_ptr(xyz)) rc0 = ensure(c, in_stage, (size_t)n * stride_floats * sizeof(float));
    if (rc0 == TKNN_OK) rc0 = ensure(c, keys_a, n * sizeof(uint64_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, keys_b, n * sizeof(uint64_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, vals_a, n * sizeof(uint32_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, vals_b, n * sizeof(uint32_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, sort_tmp, rsort::temp_words(n) * sizeof(uint32_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, ballots, nw0 * sizeof(uint32_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, c->offsets, (nw0 + 1) * sizeof(uint32_t));
    if (rc0 == TKNN_OK) rc0 = ensure(c, c->block_sums, sizeof(uint32_t) * (size_t)(nw0 / lbvh::SCAN_CHUNK + 2));
    if (rc0 == TKNN_OK) rc0 = ensure(c, c->pts, n * sizeof(float4));
    if (rc0 != TKNN_OK) { cleanup(); return rc0; }
  }
#define TK_B(expr) do { int rc_ = (expr); if (rc_ != TKNN_OK) { cleanup(); return rc_; } } while (0)
#define TK_BC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); cleanup(); \
    return fail(c, e_ == cudaErrorMemoryAllocation ? TKNN_ENOMEM : TKNN_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } } while (0)

  TK_BC(cudaEventRecord(c->ev[0], st));
  const float* d_xyz = xyz;
  if (!is_device_ptr(xyz)) {
    const size_t bytes = (size_t)n * stride_floats * sizeof(float);
    TK_B(ensure(c, in_stage, bytes));
    TK_BC(cudaMemcpyAsync(in_stage.p, xyz, bytes, cudaMemcpyHostToDevice, st));
    d_xyz = in_stage.as<float>();
    S.h2d_bytes = bytes;
  }
  TK_BC(cudaEventRecord(c->ev[1], st));

  // ---- scene bounds ----
  uint32_t* sc = c->scalars.as<uint32_t>();
  const uint32_t binit[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
  TK_BC(cudaMemcpyAsync(sc + SC_BOUNDS, binit, sizeof(binit), cudaMemcpyHostToDevice, st));
  {
    const unsigned nb = (unsigned)std::min<uint64_t>((uint64_t)c->sm_count * 8, blocks_for(n, lbvh::THREADS));
    lbvh::bounds_kernel<<<nb, lbvh::THREADS, 0, st>>>(d_xyz, n, dim, stride_floats, sc + SC_BOUNDS);
    ++launches;
  }
  TK_BC(cudaEventRecord(c->ev[2], st));

  // ---- Morton codes ----
  TK_B(ensure(c, keys_a, n * sizeof(uint64_t)));
  TK_B(ensure(c, keys_b, n * sizeof(uint64_t)));
  TK_B(ensure(c, vals_a, n * sizeof(uint32_t)));
  TK_B(ensure(c, vals_b, n * sizeof(uint32_t)));
  int mbits = 0, idx_bits = 0;
  choose_key_layout(c, n, &mbits, &idx_bits);
  lbvh::morton_kernel<<<blocks_for(n, lbvh::THREADS), lbvh::THREADS, 0, st>>>(d_xyz, n, dim, stride_floats, sc + SC_BOUNDS, mbits,
                                                                             c->curve ? hilbert_levels(n, mbits, c->curve_levels) : 0, idx_bits,
                                                                             keys_a.as<uint64_t>(), vals_a.as<uint32_t>());
  ++launches;
  TK_BC(cudaEventRecord(c->ev[3], st));

  // ---- onesweep radix sort ----
  TK_B(ensure(c, sort_tmp, rsort::temp_words(n) * sizeof(uint32_t)));
  bool sorted_in_b = false;
  if (idx_bits > 0)
    launches += rsort::sort_keys(keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), n, sort_tmp.as<uint32_t>(), c->sm_count, st, idx_bits,
                                 (3 * mbits + 7) / 8, &sorted_in_b);
  else
    launches += rsort::sort_pairs(keys_a.as<uint64_t>(), vals_a.as<uint32_t>(), keys_b.as<uint64_t>(), vals_b.as<uint32_t>(), n,
                                  sort_tmp.as<uint32_t>(), c->sm_count, st, (3 * mbits + 7) / 8, &sorted_in_b);
  const uint64_t* skeys = sorted_in_b ? keys_b.as<uint64_t>() : keys_a.as<uint64_t>();
  const uint32_t* svals = idx_bits > 0 ? nullptr : (sorted_in_b ? vals_b.as<uint32_t>() : vals_a.as<uint32_t>());
  TK_BC(cudaGetLastError());
  TK_BC(cudaEventRecord(c->ev[4], st));

  // ---- leaf cut + point gather ----
  const uint64_t nw = (n + 31) / 32;
  TK_B(ensure(c, ballots, nw * sizeof(uint32_t)));
  TK_B(ensure(c, c->offsets, (nw + 1) * sizeof(uint32_t)));
  const uint64_t force_split = (n <= (uint64_t)c->leaf_size) ? n / 2 : 0;
  lbvh::leaf_flag_kernel<<<blocks_for(nw * 32, lbvh::THREADS), lbvh::THREADS, 0, st>>>(
      skeys, n, c->leaf_size, c->leaf_policy, force_split, ballots.as<uint32_t>());
  launches += 1;
  TK_B(popc_scan(c, ballots.as<uint32_t>(), nw, c->offsets.as<uint32_t>(), &launches));
  uint32_t mb[2] = {0, 0};
  TK_B(fetch_words(c, sc + SC_TOTAL, 1, sc + SC_BOUNDS + 6, 1, mb));  // the leaf count sizes every later launch
  const uint32_t m = mb[0];
  const uint32_t bad[1] = {mb[1]};
  if (bad[0]) { cleanup(); return fail(c, TKNN_EINVAL, "non-finite coordinate in the input points"); }
  if (m < 2) { cleanup(); return fail(c, TKNN_ECUDA, "internal: leaf cut produced %u leaves", m); }
  TK_B(ensure(c, c->leaf_start, (size_t)(m + 1) * sizeof(uint32_t)));
  TK_B(ensure(c, leaf_key, (size_t)m * sizeof(uint64_t)));
  TK_B(ensure(c, c->pts, n * sizeof(float4)));
  lbvh::leaf_emit_kernel<<<blocks_for(n, lbvh::THREADS), lbvh::THREADS, 0, st>>>(
      ballots.as<uint32_t>(), c->offsets.as<uint32_t>(), skeys, n, m, c->leaf_start.as<uint32_t>(),
      leaf_key.as<uint64_t>());
  lbvh::gather_points_kernel<<<blocks_for(n, lbvh::THREADS), lbvh::THREADS, 0, st>>>(d_xyz, dim, stride_floats, svals, skeys, idx_bits,
                                                                                   n, ids_in_w ? 1 : 0, c->pts.as<float4>());
  launches += 2;
  TK_BC(cudaEventRecord(c->ev[5], st));

  // ---- Karras hierarchy ----
  TK_B(ensure(c, c->nodes, (size_t)(m - 1) * sizeof(Node)));
  TK_B(ensure(c, c->node_min_idx, (size_t)(m - 1) * 2 * sizeof(int)));
  TK_B(ensure(c, child_info, (size_t)(m - 1) * sizeof(int4)));
  TK_B(ensure(c, parent_leaf, (size_t)m * sizeof(int32_t)));
  TK_B(ensure(c, parent_node, (size_t)m * sizeof(int32_t)));
  TK_B(ensure(c, arrive, (size_t)m * sizeof(uint32_t)));
  TK_BC(cudaMemsetAsync(arrive.p, 0, (size_t)m * sizeof(uint32_t), st));
  TK_BC(cudaMemsetAsync(sc + SC_DUPLEAF, 0, sizeof(uint32_t), st));
  lbvh::karras_kernel<<<blocks_for(m - 1, lbvh::THREADS), lbvh::THREADS, 0, st>>>(
      leaf_key.as<uint64_t>(), c->leaf_start.as<uint32_t>(), m, child_info.as<int4>(), parent_leaf.as<int32_t>(),
      parent_node.as<int32_t>());
  ++launches;
  TK_BC(cudaEventRecord(c->ev[6], st));

  // ---- bottom-up refit ----
  lbvh::refit_kernel<<<blocks_for(m, lbvh::THREADS), lbvh::THREADS, 0, st>>>(
      c->pts.as<float4>(), c->leaf_start.as<uint32_t>(), m, child_info.as<int4>(), parent_leaf.as<int32_t>(),
      parent_node.as<int32_t>(), arrive.as<uint32_t>(), c->nodes.as<Node>(), c->node_min_idx.as<int>(), sc + SC_DUPLEAF,
      (uint32_t)std::max(2, std::min(8, c->leaf_size)), reinterpret_cast<float*>(sc + SC_SCENE));
  ++launches;
  TK_BC(cudaGetLastError());
  TK_BC(cudaEventRecord(c->ev[7], st));
  uint32_t tail[7] = {0, 0, 0, 0, 0, 0, 0};
  TK_B(fetch_words(c, sc + SC_SCENE, 6, sc + SC_DUPLEAF, 1, tail));
  std::memcpy(c->scene_box, tail, 6 * sizeof(float));
  const uint32_t dupleaf = tail[6];
  cleanup();
#undef TK_B
#undef TK_BC

  TK_CUDA(c, cudaEventElapsedTime(&S.h2d_ms, c->ev[0], c->ev[1]));
  TK_CUDA(c, cudaEventElapsedTime(&S.bounds_ms, c->ev[1], c->ev[2]));
  TK_CUDA(c, cudaEventElapsedTime(&S.morton_ms, c->ev[2], c->ev[3]));
  TK_CUDA(c, cudaEventElapsedTime(&S.sort_ms, c->ev[3], c->ev[4]));
  TK_CUDA(c, cudaEventElapsedTime(&S.leaves_ms, c->ev[4], c->ev[5]));
  TK_CUDA(c, cudaEventElapsedTime(&S.hierarchy_ms, c->ev[5], c->ev[6]));
  TK_CUDA(c, cudaEventElapsedTime(&S.refit_ms, c->ev[6], c->ev[7]));
  TK_CUDA(c, cudaEventElapsedTime(&S.build_ms, c->ev[1], c->ev[7]));
  S.n_points = n;
  S.n_leaves = m;
  S.n_nodes = m - 1;
  c->built_morton_bits = mbits;
  c->built_idx_bits = idx_bits;
  c->built_curve = c->curve ? hilbert_levels(n, mbits, c->curve_levels) : 0;
  c->has_dup_leaves = dupleaf != 0;
  S.build_launches = (uint32_t)(launches + c->mailbox_launches);
  c->n = n;
  c->n_leaves = m;
  c->ids_in_w = ids_in_w;
  return TKNN_OK;
}

extern "C" {

int tknn_search(tknn_ctx* c, int k, float start_radius, int32_t* idx_out, float* dist_out) {
  TK_TRY(check_ctx(c));
  ScopedDevice sd(c->device);
  return search_range(c, k, start_radius, 0, c->n, 0, nullptr, idx_out, dist_out, c->n);
}

int tknn_key_layout(uint64_t n, int morton_bits, int sort_mode, int* code_bits_per_axis, int* index_bits, int* sort_passes) {
  if (n < 2 || n > rsort::MAX_N || morton_bits < 0 || morton_bits > 21 || (morton_bits > 0 && morton_bits < 4) ||
      (sort_mode != 0 && sort_mode != 1) || !code_bits_per_axis || !index_bits || !sort_passes)
    return TKNN_EINVAL;
  key_layout(n, morton_bits, sort_mode, code_bits_per_axis, index_bits);
  *sort_passes = (3 * *code_bits_per_axis + 7) / 8;
  return TKNN_OK;
}

uint64_t tknn_shard_capacity(uint64_t n, int n_shards) {
  if (n_shards < 1) return 0;
  const uint64_t groups = (n + 31) / 32;
  return ((groups + n_shards - 1) / n_shards + 1) * 32;
}

int tknn_search_shard(tknn_ctx* c, int k, float start_radius, int shard, int n_shards, int32_t* qid_out, int32_t* idx_out,
                      float* dist_out, uint64_t* n_out) {
  TK_TRY(check_ctx(c));
  if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(c, TKNN_EINVAL, "shard %d of %d", shard, n_shards);
  if (!n_out) return fail(c, TKNN_EINVAL, "null n_out");
  ScopedDevice sd(c->device);
  const uint64_t groups = (c->n + 31) / 32;
  const uint64_t g0 = groups * (uint64_t)shard / (uint64_t)n_shards, g1 = groups * (uint64_t)(shard + 1) / (uint64_t)n_shards;
  const uint64_t q0 = g0 * 32, q1 = std::min<uint64_t>(c->n, g1 * 32);
  const uint64_t nq = q1 > q0 ? q1 - q0 : 0;
  *n_out = nq;
  if (nq == 0) {
    if (c->n == 0) return fail(c, TKNN_ESTATE, "tknn_search_shard before tknn_build");
    reset_search_stats(c);
    return TKNN_OK;
  }
  return search_range(c, k, start_radius, q0, nq, 1, qid_out, idx_out, dist_out, nq);
}

int tknn_estimate_start_radius(tknn_ctx* c, int k, float* radius_out) {
  TK_TRY(check_ctx(c));
  if (!radius_out) return fail(c, TKNN_EINVAL, "null radius_out");
  if (c->n == 0) return fail(c, TKNN_ESTATE, "tknn_estimate_start_radius before tknn_build");
  if (k < 1 || k > TKNN_MAX_K || (uint64_t)k > c->n - 1) return fail(c, TKNN_EINVAL, "k = %d invalid for n = %llu", k,
                                                                     (unsigned long long)c->n);
  ScopedDevice sd(c->device);
  int launches = 0;
  TK_TRY(estimate_radius_device(c, c->pts.as<float4>(), 0, c->n, nullptr, 1, k, &launches));
  uint32_t rb = 0;
  TK_TRY(fetch_words(c, c->scalars.as<uint32_t>() + SC_QUANTILE, 1, nullptr, 0, &rb));
  std::memcpy(radius_out, &rb, sizeof(float));
  return TKNN_OK;
}

int tknn_query(tknn_ctx* c, const float* queries, uint64_t nq, int dim, int stride_floats, const int32_t* self_ids,
               const float* init_radius2, int k, float start_radius, int32_t* idx_out, float* dist_out) {
  TK_TRY(check_ctx(c));
  if (c->n == 0) return fail(c, TKNN_ESTATE, "tknn_query before tknn_build");
  if (!queries && nq) return fail(c, TKNN_EINVAL, "null query array");
  if (dim != 2 && dim != 3) return fail(c, TKNN_EINVAL, "dim = %d (must be 2 or 3)", dim);
  if (stride_floats < dim) return fail(c, TKNN_EINVAL, "stride %d < dim %d", stride_floats, dim);
  if (k < 1 || k > TKNN_MAX_K) return fail(c, TKNN_EINVAL, "k = %d outside [1, %d]", k, TKNN_MAX_K);
  if ((uint64_t)k > c->n) return fail(c, TKNN_EINVAL, "k = %d > n = %llu", k, (unsigned long long)c->n);
  if (nq > rsort::MAX_N) return fail(c, TKNN_EINVAL, "too many queries");
  if (nq && (!idx_out || !dist_out)) return fail(c, TKNN_EINVAL, "null output array");
  if (std::isnan(start_radius)) return fail(c, TKNN_EINVAL, "start_radius is NaN");
  reset_search_stats(c);
  c->stats.n_queries = nq;
  c->stats.k = k;
  if (nq == 0) return TKNN_OK;
  ScopedDevice sd(c->device);
  cudaStream_t st = c->stream;
  uint32_t* sc = c->scalars.as<uint32_t>();

  // scratch lives in the context (grow-only): a call allocates nothing once the buffers have reached their size
  DevBuf &q_stage = c->q_stage, &keys_a = c->q_keys_a, &keys_b = c->q_keys_b, &vals_a = c->q_vals_a, &vals_b = c->q_vals_b,
         &sort_tmp = c->q_sort_tmp, &qpts = c->q_pts, &sid_stage = c->q_sid_stage, &sid_sorted = c->q_sid_sorted,
         &rad_stage = c->q_rad_stage, &r2_sorted = c->q_r2_sorted;
  auto cleanup = [&]() {
    if (c->keep_scratch) return;
    for (DevBuf* b : {&q_stage, &keys_a, &keys_b, &vals_a, &vals_b, &sort_tmp, &qpts, &sid_stage, &sid_sorted, &rad_stage,
                      &r2_sorted})
      release(*b);
  };
#define TK_B(expr) do { int rc_ = (expr); if (rc_ != TKNN_OK) { cleanup(); return rc_; } } while (0)
#define TK_BC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); cleanup(); \
    return fail(c, e_ == cudaErrorMemoryAllocation ? TKNN_ENOMEM : TKNN_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } } while (0)

  TK_BC(cudaEventRecord(c->ev[0], st));
  const float* d_q = queries;
  if (!is_device_ptr(queries)) {
    const size_t bytes = (size_t)nq * stride_floats * sizeof(float);
    TK_B(ensure(c, q_stage, bytes));
    TK_BC(cudaMemcpyAsync(q_stage.p, queries, bytes, cudaMemcpyHostToDevice, st));
    d_q = q_stage.as<float>();
  }
  const int32_t* d_sid = self_ids;
  if (self_ids && !is_device_ptr(self_ids)) {
    TK_B(ensure(c, sid_stage, nq * sizeof(int32_t)));
    TK_BC(cudaMemcpyAsync(sid_stage.p, self_ids, nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    d_sid = sid_stage.as<int32_t>();
  }
  const float* d_rad = init_radius2;
  if (init_radius2 && !is_device_ptr(init_radius2)) {
    TK_B(ensure(c, rad_stage, nq * sizeof(float)));
    TK_BC(cudaMemcpyAsync(rad_stage.p, init_radius2, nq * sizeof(float), cudaMemcpyHostToDevice, st));
    d_rad = rad_stage.as<float>();
  }
  // Morton-sort the queries on the DATA's grid so that groups of 32 are spatially coherent
  TK_B(ensure(c, keys_a, nq * sizeof(uint64_t)));
  TK_B(ensure(c, keys_b, nq * sizeof(uint64_t)));
  TK_B(ensure(c, vals_a, nq * sizeof(uint32_t)));
  TK_B(ensure(c, vals_b, nq * sizeof(uint32_t)));
  TK_B(ensure(c, sort_tmp, rsort::temp_words(nq) * sizeof(uint32_t)));
  TK_B(ensure(c, qpts, nq * sizeof(float4)));
  int launches = 0;
  TK_BC(cudaMemsetAsync(sc + SC_ERROR, 0, 2 * sizeof(uint32_t), st));
  const int qbits = c->built_morton_bits;
  lbvh::morton_kernel<<<blocks_for(nq, lbvh::THREADS), lbvh::THREADS, 0, st>>>(d_q, nq, dim, stride_floats, sc + SC_BOUNDS, qbits,
                                                                              c->built_curve, 0, keys_a.as<uint64_t>(), vals_a.as<uint32_t>());
  ++launches;
  bool q_in_b = false;
  launches += rsort::sort_pairs(keys_a.as<uint64_t>(), vals_a.as<uint32_t>(), keys_b.as<uint64_t>(), vals_b.as<uint32_t>(), nq,
                                sort_tmp.as<uint32_t>(), c->sm_count, st, (3 * qbits + 7) / 8, &q_in_b);
  const uint32_t* qorder = q_in_b ? vals_b.as<uint32_t>() : vals_a.as<uint32_t>();
  lbvh::gather_points_kernel<<<blocks_for(nq, lbvh::THREADS), lbvh::THREADS, 0, st>>>(d_q, dim, stride_floats,
                                                                                    qorder, nullptr, 0, nq, 0, qpts.as<float4>(), sc + SC_QBAD);
  TK_BC(cudaGetLastError());
  ++launches;
  const int32_t* d_sid_sorted = nullptr;
  const float* d_r2_sorted = nullptr;
  if (d_sid) {
    TK_B(ensure(c, sid_sorted, nq * sizeof(int32_t)));
    brute::gather_i32_kernel<<<blocks_for(nq, 256), 256, 0, st>>>(d_sid, qorder, nq, sid_sorted.as<int32_t>());
    d_sid_sorted = sid_sorted.as<int32_t>();
    ++launches;
  }
  if (d_rad) {
    TK_B(ensure(c, r2_sorted, nq * sizeof(float)));
    brute::gather_r2_kernel<<<blocks_for(nq, 256), 256, 0, st>>>(d_rad, qorder, nq, r2_sorted.as<float>());
    d_r2_sorted = r2_sorted.as<float>();
    ++launches;
  }

  const bool idx_dev = is_device_ptr(idx_out), dist_dev = is_device_ptr(dist_out);
  int32_t* d_idx = idx_out;
  float* d_dist = dist_out;
  const size_t out_elems = (size_t)nq * k;
  if (!idx_dev) { TK_B(ensure(c, c->stage_idx, out_elems * sizeof(int32_t))); d_idx = c->stage_idx.as<int32_t>(); }
  if (!dist_dev) { TK_B(ensure(c, c->stage_dist, out_elems * sizeof(float))); d_dist = c->stage_dist.as<float>(); }

  TK_BC(cudaMemsetAsync(sc + SC_COUNTERS, 0, 8 * sizeof(unsigned long long), st));
  float r0 = start_radius;
  // a query set can hold fewer than k reachable neighbours only through self exclusion / radius caps
  const bool may_underfill = (uint64_t)k > c->n - (self_ids ? 1 : 0);
  bool estimated = false;
  if (!(r0 > 0.0f)) {
    if (init_radius2 || may_underfill) r0 = INFINITY;
    else { TK_B(estimate_radius_device(c, qpts.as<float4>(), 0, nq, d_sid_sorted, 0, k, &launches)); estimated = true; }
  }
  TK_BC(cudaEventRecord(c->ev[1], st));
  c->stats.start_radius = r0;
  Job job;
  if (estimated) {
    job.r_dev = reinterpret_cast<const float*>(sc + SC_QUANTILE);
    job.r_host = &c->stats.start_radius;
  }
  job.queries = qpts.as<float4>();
  job.self_ids = d_sid_sorted;
  job.query_r2 = d_r2_sorted;
  job.n_queries = nq;
  job.q_begin = 0;
  job.self_is_row = 0;
  job.row_mode = 0;
  job.k = k;
  job.start_radius = r0;
  job.squared = c->squared;
  job.idx_out = d_idx;
  job.dist_out = d_dist;
  TK_B(run_rounds(c, job, &launches));
  TK_BC(cudaEventRecord(c->ev[2], st));
  if (!idx_dev) { TK_BC(cudaMemcpyAsync(idx_out, d_idx, out_elems * sizeof(int32_t), cudaMemcpyDeviceToHost, st)); c->stats.d2h_bytes += out_elems * 4; }
  if (!dist_dev) { TK_BC(cudaMemcpyAsync(dist_out, d_dist, out_elems * sizeof(float), cudaMemcpyDeviceToHost, st)); c->stats.d2h_bytes += out_elems * 4; }
  TK_BC(cudaEventRecord(c->ev[3], st));
  TK_BC(cudaStreamSynchronize(st));
  cleanup();
#undef TK_B
#undef TK_BC
  TK_CUDA(c, cudaEventElapsedTime(&c->stats.search_ms, c->ev[0], c->ev[2]));
  TK_CUDA(c, cudaEventElapsedTime(&c->stats.d2h_ms, c->ev[2], c->ev[3]));
  c->stats.kernel_launches = (uint32_t)(launches + c->mailbox_launches + (c->mailbox_h ? 1 : 0));  // + the error-flag fetch below
  TK_TRY(finish_round_stats(c));
  TK_TRY(collect_counters(c));
  TK_TRY(read_error_flag(c));
  return TKNN_OK;
}

int tknn_range_count(tknn_ctx* c, float radius, uint32_t* count_out) {
  TK_TRY(check_ctx(c));
  if (c->n == 0) return fail(c, TKNN_ESTATE, "tknn_range_count before tknn_build");
  if (!count_out) return fail(c, TKNN_EINVAL, "null output array");
  if (!(radius >= 0.0f)) return fail(c, TKNN_EINVAL, "radius must be >= 0");
  ScopedDevice sd(c->device);
  reset_search_stats(c);
  c->stats.n_queries = c->n;
  uint32_t* sc = c->scalars.as<uint32_t>();
  const bool dev = is_device_ptr(count_out);
  uint32_t* d_out = count_out;
  if (!dev) { TK_TRY(ensure(c, c->stage_idx, c->n * sizeof(uint32_t))); d_out = c->stage_idx.as<uint32_t>(); }
  trav::Params P;
  std::memset(&P, 0, sizeof(P));
  P.nodes = c->nodes.as<Node>();
  P.node_min_idx = c->node_min_idx.as<int2>();
  P.pts = c->pts.as<float4>();
  P.queries = c->pts.as<float4>();
  P.n_active = c->n;
  P.n_groups = (uint32_t)((c->n + 31) / 32);
  P.r2 = radius * radius;
  P.k = 0;
  P.self_is_row = 1;
  P.final_round = 1;
  P.error = sc + SC_ERROR;
  P.count_out = d_out;
  P.group_counter = sc + SC_GROUP_COUNTER;
  P.counters = c->counters ? reinterpret_cast<unsigned long long*>(sc + SC_COUNTERS) : nullptr;
  TK_CUDA(c, cudaMemsetAsync(sc + SC_ERROR, 0, 2 * sizeof(uint32_t), c->stream));
  TK_CUDA(c, cudaMemsetAsync(sc + SC_COUNTERS, 0, 8 * sizeof(unsigned long long), c->stream));
  TK_CUDA(c, cudaMemsetAsync(sc + SC_GROUP_COUNTER, 0, sizeof(uint32_t), c->stream));
  TK_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
  TK_TRY(launch_traverse<trav::MODE_RANGE_COUNT>(c, P));
  TK_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  if (!dev) {
    TK_CUDA(c, cudaMemcpyAsync(count_out, d_out, c->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes = c->n * sizeof(uint32_t);
  }
  TK_CUDA(c, cudaStreamSynchronize(c->stream));
  TK_CUDA(c, cudaEventElapsedTime(&c->stats.search_ms, c->ev[0], c->ev[2]));
  c->stats.rounds = 1;
  c->stats.kernel_launches = 1;
  c->stats.start_radius = c->stats.final_radius = radius;
  TK_TRY(collect_counters(c));
  TK_TRY(read_error_flag(c));
  return TKNN_OK;
}

int tknn_brute_force(tknn_ctx* c, const int32_t* query_ids, uint64_t nq, int k, int32_t* idx_out, float* dist_out) {
  TK_TRY(check_ctx(c));
  if (c->n == 0) return fail(c, TKNN_ESTATE, "tknn_brute_force before tknn_build");
  if (k < 1 || k > TKNN_MAX_K || (uint64_t)k > c->n - 1) return fail(c, TKNN_EINVAL, "k = %d invalid for n = %llu", k,
                                                                     (unsigned long long)c->n);
  if (nq == 0) return TKNN_OK;
  if (!query_ids || !idx_out || !dist_out) return fail(c, TKNN_EINVAL, "null array");
  if (nq > (1u << 20)) return fail(c, TKNN_EINVAL, "at most 2^20 brute-force queries per call");
  if (c->ids_in_w) return fail(c, TKNN_ESTATE, "this BVH carries caller-chosen point ids: use tknn_partition_verify");
  ScopedDevice sd(c->device);
  cudaStream_t st = c->stream;
  std::vector<int32_t> ids(nq);
  if (is_device_ptr(query_ids)) {
    TK_CUDA(c, cudaMemcpyAsync(ids.data(), query_ids, nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TK_CUDA(c, cudaStreamSynchronize(st));
  } else {
    std::memcpy(ids.data(), query_ids, nq * sizeof(int32_t));
  }
  std::vector<uint32_t> perm(nq);
  for (uint32_t i = 0; i < nq; ++i) perm[i] = i;
  std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
  std::vector<int32_t> sorted(nq);
  for (uint32_t i = 0; i < nq; ++i) sorted[i] = ids[perm[i]];
  for (uint32_t i = 0; i < nq; ++i) {
    if (sorted[i] < 0 || (uint64_t)sorted[i] >= c->n) return fail(c, TKNN_EINVAL, "query id %d out of range", sorted[i]);
    if (i && sorted[i] == sorted[i - 1]) return fail(c, TKNN_EINVAL, "duplicate query id %d", sorted[i]);
  }
  DevBuf d_ids, d_perm, qpts, o_idx, o_dist;
  auto cleanup = [&]() { for (DevBuf* b : {&d_ids, &d_perm, &qpts, &o_idx, &o_dist}) release(*b); };
#define TK_B(expr) do { int rc_ = (expr); if (rc_ != TKNN_OK) { cleanup(); return rc_; } } while (0)
#define TK_BC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); cleanup(); \
    return fail(c, e_ == cudaErrorMemoryAllocation ? TKNN_ENOMEM : TKNN_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } } while (0)
  TK_B(ensure(c, d_ids, nq * sizeof(int32_t)));
  TK_B(ensure(c, d_perm, nq * sizeof(uint32_t)));
  TK_B(ensure(c, qpts, nq * sizeof(float4)));
  TK_BC(cudaMemcpyAsync(d_ids.p, sorted.data(), nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  TK_BC(cudaMemcpyAsync(d_perm.p, perm.data(), nq * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  const bool idx_dev = is_device_ptr(idx_out), dist_dev = is_device_ptr(dist_out);
  int32_t* d_idx = idx_out;
  float* d_dist = dist_out;
  if (!idx_dev) { TK_B(ensure(c, o_idx, nq * k * sizeof(int32_t))); d_idx = o_idx.as<int32_t>(); }
  if (!dist_dev) { TK_B(ensure(c, o_dist, nq * k * sizeof(float))); d_dist = o_dist.as<float>(); }
  brute::lookup_queries_kernel<<<blocks_for(c->n, 256), 256, 0, st>>>(c->pts.as<float4>(), c->n, d_ids.as<int32_t>(),
                                                                       (uint32_t)nq, qpts.as<float4>());
  TK_B(brute_core(c, qpts.as<float4>(), nq, k, d_perm.as<uint32_t>(), d_idx, d_dist, 0));
  if (!idx_dev) TK_BC(cudaMemcpyAsync(idx_out, d_idx, nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (!dist_dev) TK_BC(cudaMemcpyAsync(dist_out, d_dist, nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  TK_BC(cudaStreamSynchronize(st));
  cleanup();
#undef TK_B
#undef TK_BC
  return TKNN_OK;
}

}  // extern "C"

// Exact tiled brute force of external queries (x, y, z, id to exclude) against the sorted points: one warp per
// (32 queries, split of the point range), then a per-query merge of the splits.  Fills the grid twice over.
int tknn::host::brute_core(tknn_ctx* c, const float4* d_qpts, uint64_t nq, int k, const uint32_t* d_row_of, int32_t* d_idx,
                           float* d_dist, int squared) {
  if (nq == 0) return TKNN_OK;
  cudaStream_t st = c->stream;
  const uint64_t groups = (nq + 31) / 32;
  int splits = (int)std::max<uint64_t>(1, std::min<uint64_t>(256, ((uint64_t)c->sm_count * 64 + groups - 1) / groups));
  splits = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)splits, (c->n + 1023) / 1024));
  DevBuf partial;
  int rc = ensure(c, partial, (size_t)splits * nq * k * sizeof(uint64_t));
  if (rc != TKNN_OK) return rc;
  const size_t smem = 32 * sizeof(float4) + (size_t)k * 32 * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(brute::brute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(brute::merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    dim3 grid((unsigned)groups, (unsigned)splits);
    brute::brute_kernel<<<grid, 32, smem, st>>>(c->pts.as<float4>(), c->n, d_qpts, (uint32_t)nq, k, splits, partial.as<uint64_t>());
    brute::merge_keys_kernel<<<(unsigned)groups, 32, smem, st>>>(partial.as<uint64_t>(), splits, (uint32_t)nq, k, d_row_of, squared,
                                                                 d_idx, d_dist);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `partial` is released below
  release(partial);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(c, TKNN_ECUDA, "brute force: %s", cudaGetErrorString(e)); }
  return TKNN_OK;
}

extern "C" {

int tknn_merge_topk(tknn_ctx* c, const int32_t* idx_parts, const float* d2_parts, int parts, uint64_t nq, int k,
                    int32_t* idx_out, float* dist_out) {
  TK_TRY(check_ctx(c));
  if (parts < 1 || k < 1 || k > TKNN_MAX_K) return fail(c, TKNN_EINVAL, "bad parts / k");
  if (nq == 0) return TKNN_OK;
  if (!idx_parts || !d2_parts || !idx_out || !dist_out) return fail(c, TKNN_EINVAL, "null array");
  if (!is_device_ptr(idx_parts) || !is_device_ptr(d2_parts) || !is_device_ptr(idx_out) || !is_device_ptr(dist_out))
    return fail(c, TKNN_EINVAL, "tknn_merge_topk takes device arrays");
  ScopedDevice sd(c->device);
  const size_t smem = (size_t)k * 32 * sizeof(uint64_t);
  TK_CUDA(c, cudaFuncSetAttribute(brute::merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  brute::merge_lists_kernel<<<(unsigned)((nq + 31) / 32), 32, smem, c->stream>>>(idx_parts, d2_parts, parts, (uint32_t)nq, k,
                                                                               c->squared, idx_out, dist_out);
  TK_CUDA(c, cudaGetLastError());
  TK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TKNN_OK;
}

int tknn_get_stats(const tknn_ctx* c, tknn_stats* out) {
  if (!c || !out) return TKNN_EINVAL;
  *out = c->stats;
  return TKNN_OK;
}

int tknn_sort_pairs(tknn_ctx* c, uint64_t* keys, uint32_t* values, uint64_t n) {
  TK_TRY(check_ctx(c));
  if (n == 0) return TKNN_OK;
  if (!keys) return fail(c, TKNN_EINVAL, "null array");
  if (n > rsort::MAX_N) return fail(c, TKNN_EINVAL, "n too large");
  ScopedDevice sd(c->device);
  cudaStream_t st = c->stream;
  const bool kd = is_device_ptr(keys), vd = !values || is_device_ptr(values);  // values == nullptr: keys only
  DevBuf ka, kb, va, vb, tmp;
  auto cleanup = [&]() { for (DevBuf* b : {&ka, &kb, &va, &vb, &tmp}) release(*b); };
#define TK_B(expr) do { int rc_ = (expr); if (rc_ != TKNN_OK) { cleanup(); return rc_; } } while (0)
#define TK_BC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); cleanup(); \
    return fail(c, e_ == cudaErrorMemoryAllocation ? TKNN_ENOMEM : TKNN_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } } while (0)
  uint64_t* dk = keys;
  uint32_t* dv = values;
  if (!kd) { TK_B(ensure(c, ka, n * sizeof(uint64_t))); dk = ka.as<uint64_t>(); TK_BC(cudaMemcpyAsync(dk, keys, n * 8, cudaMemcpyHostToDevice, st)); }
  if (!vd) { TK_B(ensure(c, va, n * sizeof(uint32_t))); dv = va.as<uint32_t>(); TK_BC(cudaMemcpyAsync(dv, values, n * 4, cudaMemcpyHostToDevice, st)); }
  TK_B(ensure(c, kb, n * sizeof(uint64_t)));
  if (values) TK_B(ensure(c, vb, n * sizeof(uint32_t)));
  TK_B(ensure(c, tmp, rsort::temp_words(n) * sizeof(uint32_t)));
  if (values) rsort::sort_pairs(dk, dv, kb.as<uint64_t>(), vb.as<uint32_t>(), n, tmp.as<uint32_t>(), c->sm_count, st);
  else rsort::sort_keys(dk, kb.as<uint64_t>(), n, tmp.as<uint32_t>(), c->sm_count, st, 0, rsort::PASSES);
  TK_BC(cudaGetLastError());
  if (!kd) TK_BC(cudaMemcpyAsync(keys, dk, n * 8, cudaMemcpyDeviceToHost, st));
  if (!vd) TK_BC(cudaMemcpyAsync(values, dv, n * 4, cudaMemcpyDeviceToHost, st));
  TK_BC(cudaStreamSynchronize(st));
  cleanup();
#undef TK_B
#undef TK_BC
  return TKNN_OK;
}

int tknn_get_bvh(const tknn_ctx* c, void* nodes_out, void* points_out, uint32_t* leaf_start_out) {
  if (!c) return TKNN_EINVAL;
  if (c->n == 0) return TKNN_ESTATE;
  ScopedDevice sd(c->device);
  cudaStreamSynchronize(c->stream);
  if (nodes_out && cudaMemcpy(nodes_out, c->nodes.p, (size_t)(c->n_leaves - 1) * sizeof(Node), cudaMemcpyDefault) != cudaSuccess)
    return TKNN_ECUDA;
  if (points_out && cudaMemcpy(points_out, c->pts.p, c->n * sizeof(float4), cudaMemcpyDefault) != cudaSuccess) return TKNN_ECUDA;
  if (leaf_start_out &&
      cudaMemcpy(leaf_start_out, c->leaf_start.p, (size_t)(c->n_leaves + 1) * sizeof(uint32_t), cudaMemcpyDefault) != cudaSuccess)
    return TKNN_ECUDA;
  return TKNN_OK;
}

int tknn_morton_codes(tknn_ctx* c, const float* xyz, uint64_t n, int dim, int stride_floats, const float* box6,
                      uint64_t* codes_out) {
  TK_TRY(check_ctx(c));
  if (n == 0) return TKNN_OK;
  if (!xyz || !box6 || !codes_out) return fail(c, TKNN_EINVAL, "null array");
  if (dim != 2 && dim != 3) return fail(c, TKNN_EINVAL, "dim = %d (must be 2 or 3)", dim);
  if (stride_floats < dim) return fail(c, TKNN_EINVAL, "stride %d < dim %d", stride_floats, dim);
  ScopedDevice sd(c->device);
  cudaStream_t st = c->stream;
  float hbox[6];
  if (is_device_ptr(box6)) {
    TK_CUDA(c, cudaMemcpyAsync(hbox, box6, sizeof(hbox), cudaMemcpyDeviceToHost, st));
    TK_CUDA(c, cudaStreamSynchronize(st));
  } else {
    std::memcpy(hbox, box6, sizeof(hbox));
  }
  // same order-preserving encoding the builder's bounds kernel produces
  auto enc = [](float f) { uint32_t b; std::memcpy(&b, &f, 4); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); };
  uint32_t ob[8] = {enc(hbox[0]), enc(hbox[1]), enc(hbox[2]), enc(hbox[3]), enc(hbox[4]), enc(hbox[5]), 0u, 0u};
  DevBuf in, keys, vals, dob;
  auto cleanup = [&]() { for (DevBuf* b : {&in, &keys, &vals, &dob}) release(*b); };
#define TK_B(expr) do { int rc_ = (expr); if (rc_ != TKNN_OK) { cleanup(); return rc_; } } while (0)
#define TK_BC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); cleanup(); \
    return fail(c, e_ == cudaErrorMemoryAllocation ? TKNN_ENOMEM : TKNN_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } } while (0)
  const float* d_xyz = xyz;
  if (!is_device_ptr(xyz)) {
    TK_B(ensure(c, in, (size_t)n * stride_floats * sizeof(float)));
    TK_BC(cudaMemcpyAsync(in.p, xyz, (size_t)n * stride_floats * sizeof(float), cudaMemcpyHostToDevice, st));
    d_xyz = in.as<float>();
  }
  const bool out_dev = is_device_ptr(codes_out);
  uint64_t* d_keys = codes_out;
  if (!out_dev) { TK_B(ensure(c, keys, n * sizeof(uint64_t))); d_keys = keys.as<uint64_t>(); }
  TK_B(ensure(c, vals, n * sizeof(uint32_t)));
  TK_B(ensure(c, dob, sizeof(ob)));
  TK_BC(cudaMemcpyAsync(dob.p, ob, sizeof(ob), cudaMemcpyHostToDevice, st));
  lbvh::morton_kernel<<<blocks_for(n, lbvh::THREADS), lbvh::THREADS, 0, st>>>(d_xyz, n, dim, stride_floats, dob.as<uint32_t>(), 21,
                                                                             0, 0, d_keys, vals.as<uint32_t>());
  TK_BC(cudaGetLastError());
  if (!out_dev) TK_BC(cudaMemcpyAsync(codes_out, d_keys, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  TK_BC(cudaStreamSynchronize(st));
  cleanup();
#undef TK_B
#undef TK_BC
  return TKNN_OK;
}

int tknn_reach_mask(tknn_ctx* c, const float* xyz, uint64_t n, int stride_floats, const float* reach2, const float* box6,
                    const float* summaries, int n_ranks, int cell_bits, int self_rank, uint32_t* mask_out) {
  TK_TRY(check_ctx(c));
  if (n == 0) return TKNN_OK;
  if (!xyz || !reach2 || !box6 || !summaries || !mask_out) return fail(c, TKNN_EINVAL, "null array");
  if (n_ranks < 1 || n_ranks > 32 || cell_bits < 1 || cell_bits > 5 || stride_floats < 3)
    return fail(c, TKNN_EINVAL, "n_ranks in [1,32], cell_bits in [1,5], stride >= 3");
  if (!is_device_ptr(xyz) || !is_device_ptr(reach2) || !is_device_ptr(box6) || !is_device_ptr(summaries) ||
      !is_device_ptr(mask_out))
    return fail(c, TKNN_EINVAL, "tknn_reach_mask takes device arrays");
  ScopedDevice sd(c->device);
  brute::reach_mask_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(xyz, n, stride_floats, reach2, box6, summaries, n_ranks,
                                                                     cell_bits, self_rank, mask_out);
  TK_CUDA(c, cudaGetLastError());
  TK_CUDA(c, cudaStreamSynchronize(c->stream));
  return TKNN_OK;
}

int tknn_generate_uniform(tknn_ctx* c, uint64_t seed, uint64_t first, uint64_t n, float* xyz_out) {
  TK_TRY(check_ctx(c));
  if (n == 0) return TKNN_OK;
  if (!xyz_out) return fail(c, TKNN_EINVAL, "null output array");
  ScopedDevice sd(c->device);
  const bool dev = is_device_ptr(xyz_out);
  DevBuf tmp;
  float* d = xyz_out;
  if (!dev) {
    int rc = ensure(c, tmp, 3 * n * sizeof(float));
    if (rc != TKNN_OK) return rc;
    d = tmp.as<float>();
  }
  brute::generate_uniform_kernel<<<blocks_for(3 * n, 256), 256, 0, c->stream>>>(seed, first, n, d);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && !dev) e = cudaMemcpyAsync(xyz_out, d, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  release(tmp);
  if (e != cudaSuccess) return fail(c, TKNN_ECUDA, "generate_uniform: %s", cudaGetErrorString(e));
  return TKNN_OK;
}

int tknn_measure_bandwidth(tknn_ctx* c, double* l2_gbs, double* hbm_gbs, int* sm_count, uint64_t* l2_bytes) {
  TK_TRY(check_ctx(c));
  ScopedDevice sd(c->device);
  if (sm_count) *sm_count = c->sm_count;
  if (l2_bytes) *l2_bytes = c->l2_bytes;
  cudaStream_t st = c->stream;
  DevBuf buf;
  const size_t big = (size_t)2 << 30;  // HBM-sized: 2 GiB >> L2
  int rc = ensure(c, buf, big);
  if (rc != TKNN_OK) return rc;
  cudaError_t e = cudaMemsetAsync(buf.p, 1, big, st);
  uint32_t* sink = c->scalars.as<uint32_t>() + SC_WORDS - 1;
  const unsigned grid = (unsigned)c->sm_count * 16;
  float ms = 0.f;
  auto run = [&](size_t bytes, int passes, double* out) {
    const uint64_t words = bytes / 16;
    brute::read_probe_kernel<<<grid, 256, 0, st>>>(buf.as<uint4>(), words, 1, sink);  // warm
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(c->ev[0], st);
      brute::read_probe_kernel<<<grid, 256, 0, st>>>(buf.as<uint4>(), words, passes, sink);
      cudaEventRecord(c->ev[1], st);
      cudaEventSynchronize(c->ev[1]);
      cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
      const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
      if (gbs > best) best = gbs;
    }
    if (out) *out = best;
  };
  if (e == cudaSuccess) {
    run((size_t)32 << 20, 64, l2_gbs);  // 32 MiB stays L2-resident
    run(big, 1, hbm_gbs);
    e = cudaGetLastError();
  }
  release(buf);
  if (e != cudaSuccess) return fail(c, TKNN_ECUDA, "measure_bandwidth: %s", cudaGetErrorString(e));
  return TKNN_OK;
}

// Aggregate shared-memory bandwidth in GB/s of bytes delivered to lanes: conflict-free LDS.128 (the 128 B/clk/SM
// crossbar) and warp-broadcast LDS.128 (the traversal kernel's leaf filter).  Best of 5 runs of ~1 ms each.
int tknn_measure_smem_bandwidth(tknn_ctx* c, double* conflict_free_gbs, double* broadcast_gbs) {
  TK_TRY(check_ctx(c));
  ScopedDevice sd(c->device);
  cudaStream_t st = c->stream;
  uint32_t* sink = c->scalars.as<uint32_t>() + SC_WORDS - 1;
  const unsigned grid = (unsigned)c->sm_count * 8;
  const int iters = 2048;
  for (int mode = 0; mode < 2; ++mode) {
    double best = 0;
    brute::smem_probe_kernel<<<grid, 256, 0, st>>>(64, mode, 264, sink);  // warm
    for (int rep = 0; rep < 5; ++rep) {
      float ms = 0.f;
      TK_CUDA(c, cudaEventRecord(c->ev[0], st));
      brute::smem_probe_kernel<<<grid, 256, 0, st>>>(iters, mode, 264, sink);
      TK_CUDA(c, cudaEventRecord(c->ev[1], st));
      TK_CUDA(c, cudaEventSynchronize(c->ev[1]));
      TK_CUDA(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
      const double gbs = (double)grid * 256 * iters * 8 * 16 / (ms * 1e-3) / 1e9;
      if (gbs > best) best = gbs;
    }
    if (mode == 0 && conflict_free_gbs) *conflict_free_gbs = best;
    if (mode == 1 && broadcast_gbs) *broadcast_gbs = best;
  }
  TK_CUDA(c, cudaGetLastError());
  return TKNN_OK;
}

}  // extern "C"
