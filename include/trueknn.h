/*
 * trueknn.h — C ABI of libtrueknn: the B200-native (sm_100a) replacement for the hot path of
 * vani-nag/OWLRayTracing's samples/s01-trueknn (TrueKNN): accel build -> traverse / intersect /
 * k-list -> radius-doubling rounds.  Citations are relative to the reference repository root.
 *
 * The reference has no kNN library interface: its sample drives ~25 OWL C-API calls
 * (owl/include/owl/owl_host.h:336-1240, call sites samples/s01-trueknn/hostCode.cpp:141-362).
 * Each entry point below names the slice of that sequence it replaces.
 *
 * Conventions (all functions): plain pointers and sizes, no C++ types; returns TKNN_OK (0) or a
 * TKNN_E* code and never throws or exits (the reference throws std::runtime_error through
 * extern "C" — owl/helper/cuda.h:22-31 — and exit(2)s on OptiX errors — owl/helper/optix.h:34-42);
 * the message of the last failure is available from tknn_last_error().  A context is bound to one
 * CUDA device and is not thread-safe; distinct contexts are independent.  Input/output arrays may
 * be HOST or DEVICE pointers (detected with cudaPointerGetAttributes); host buffers are staged
 * through device memory inside the call.  Calls are synchronous on return, like owlLaunch2D
 * (owl/impl.cpp:168-176).  There is no CPU fallback: without a CUDA device tknn_create fails.
 *
 * Result semantics (north star): for query q the k data points p != q BY INDEX minimising
 * (d2(q,p), index(p)) lexicographically, d2 = fmaf(dz,dz,fmaf(dy,dy,dx*dx)) in fp32;
 * idx_out[q*k+i] ascending, dist_out[q*k+i] = sqrtf(d2); unfilled slots are -1 / FLT_MAX
 * (hostCode.cpp:129).  Indices refer to the order of the points passed to tknn_build
 * (= file order in the sample, hostCode.cpp:115-124).
 */
#ifndef TRUEKNN_H
#define TRUEKNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TKNN_API
#else
#define TKNN_API __attribute__((visibility("default")))
#endif

#define TKNN_VERSION 100

enum {
  TKNN_OK = 0,
  TKNN_EINVAL = 1, /* bad argument (k >= n, dim not 2|3, non-finite coordinate, ...) */
  TKNN_ENOMEM = 2, /* host or device allocation failed */
  TKNN_ECUDA = 3,  /* CUDA runtime error (text in tknn_last_error) */
  TKNN_ENCCL = 4,  /* NCCL error, or libnccl.so.2 could not be loaded (multi-GPU calls only) */
  TKNN_ESTATE = 5  /* call order (search before build, ...) */
};

#define TKNN_MAX_K 512
#define TKNN_MAX_ROUNDS 48

/* option keys for tknn_set_option */
enum {
  TKNN_OPT_LEAF_SIZE = 1,   /* max points per BVH leaf, 4..32 (default 32)                         */
  TKNN_OPT_COUNTERS = 2,    /* 1: run the counting variant of the traversal kernel (V/T of §8d)   */
  TKNN_OPT_LEAF_POLICY = 3, /* 0: Morton-aligned maximal subtrees (default); 1: fixed chunks       */
  TKNN_OPT_SAMPLE_GROUPS = 4, /* groups of 32 queries sampled by the start-radius estimator (128) */
  TKNN_OPT_BLOCKS_PER_SM = 5, /* persistent-grid size of the traversal kernel, 0 = auto            */
  TKNN_OPT_SQUARED_DIST = 6,  /* 1: dist_out receives d2 instead of sqrtf(d2)                      */
  TKNN_OPT_RADIUS_QUANTILE = 7 /* start-radius estimator: per-mille quantile of the sampled k-th
                                  neighbour distance (default 990)                                */
  ,TKNN_OPT_KEEP_SCRATCH = 8   /* 1 (default): keep the build scratch buffers for the next tknn_build    */
  ,TKNN_OPT_APPROX_FILTER = 10 /* 1: conservative 3-FMA pre-filter before the exact distance test
                                  (audited, -1 % time); 0 (default): exact test on every pair */
  ,TKNN_OPT_OUTPUT_CHUNKS = 11 /* tknn_search_shard with HOST outputs: number of Morton slices whose
                                  device->host copies overlap the search of the next slice (default 4; 1 = off) */
  ,TKNN_OPT_FILE_ORDER_CHUNKS = 12 /* tknn_search with HOST outputs: slices by original index (rows of a slice are
                                  contiguous and final), copies overlapped like above (default 0 = auto: 5, or 2 when
                                  dist_out is NULL; for k > 24: 2 and 1; 1 = off)   */
  ,TKNN_OPT_MORTON_BITS = 13   /* curve-code bits per axis for the next build: 0 = auto (see TKNN_OPT_SORT_MODE), else 4..21 */
  ,TKNN_OPT_TIE_PRUNING = 14   /* index-aware pruning of exact distance ties: 0 = auto (kernel variant used when the
                                  build found a leaf of >= 8 coincident points, i.e. a duplicate cluster), 1 = always,
                                  2 = never           */
  ,TKNN_OPT_WARP_ROUND_MAX = 15 /* rounds with at most this many active queries (the start-radius sample, late rounds of a
                                  few stragglers, small query sets) run one WARP per query (default 49152; 0 = never) */
  ,TKNN_OPT_CURVE = 16          /* space-filling curve the next build sorts the points by: 0 = Hilbert (default; consecutive
                                  points are always neighbours in space, so a 32-query group is compact), 1 = Morton */
  ,TKNN_OPT_SPECULATIVE_MAX = 17 /* searches of at most this many queries run without a host decision between their launches:
                                  round 2 is launched as the final round before the host knows how many queries round 1 left
                                  (default 2^20; 0 = always wait for the count) */
  ,TKNN_OPT_SPARSE_TEAM = 18    /* sparse rounds (more than WARP_ROUND_MAX leftover queries, fewer than n / SPARSE_DIVISOR): 0 (default)
                                  = one thread per query; 4, 8, 16 = a team of that many lanes shares one traversal (measured
                                  slower: cfg2's 97 478 leftovers 0.45 ms with threads, 0.78 / 0.65 / 0.63 ms with teams) */
  ,TKNN_OPT_SORT_MODE = 19      /* build sort: 0 (default) = packed keys when they fit — (curve code << index bits | point index) in one
                                  u64, as many code bits per axis as fit beside the index (at most 13: five 8-bit passes), sorted
                                  keys-only at 16 B per point and pass; falls back to 1 when fewer than log2(n)/3 + 3 bits per axis
                                  would fit (n > 2^28) or TKNN_OPT_MORTON_BITS asks for more than fit; 1 = (u64 code, u32 index)
                                  pairs, log2(n)/3 + 8 bits per axis, 24 B per point and pass (round 1's sort) */
  ,TKNN_OPT_SPARSE_DIVISOR = 9 /* rounds >= 2 with fewer than n/divisor active queries run the
                                  thread-per-query kernel (default 8; 0 = never)                  */
};

typedef struct tknn_ctx tknn_ctx;

typedef struct tknn_stats {
  /* build (replaces "Build time", hostCode.cpp:201-212) */
  uint64_t n_points;
  uint32_t n_leaves;
  uint32_t n_nodes;
  float build_ms;     /* points on device -> BVH ready (sum of the phases below) */
  float bounds_ms;    /* scene AABB reduce                                        */
  float morton_ms;    /* Morton codes                                             */
  float sort_ms;      /* onesweep radix sort (8 passes)                           */
  float leaves_ms;    /* leaf cut + point gather                                  */
  float hierarchy_ms; /* Karras hierarchy                                         */
  float refit_ms;     /* bottom-up AABB refit                                     */
  float h2d_ms;       /* host->device staging of the points (0 for device input)  */
  /* last search (replaces "True KNN time", hostCode.cpp:279-347) */
  uint64_t n_queries;
  int32_t k;
  int32_t rounds;
  float start_radius; /* the radius round 1 ran with (after auto-estimation)      */
  float final_radius;
  float estimate_ms;  /* start-radius estimator (0 when the caller gave a radius) */
  float search_ms;    /* first launch -> last result written on device, all rounds, incl. estimate */
  float d2h_ms;       /* device->host copy of results (0 for device output)       */
  float round_ms[TKNN_MAX_ROUNDS];  /* round incl. compaction + the host termination check */
  float kernel_ms[TKNN_MAX_ROUNDS]; /* the traversal kernel alone, CUDA events on the launching stream */
  uint64_t round_queries[TKNN_MAX_ROUNDS]; /* queries active in each round */
  uint32_t kernel_launches; /* kernels launched by the last search */
  uint32_t build_launches;  /* kernels launched by the last build  */
  /* counters (TKNN_OPT_COUNTERS=1), summed over all rounds; per-query, i.e. a 32-query group
   * visiting a node counts once per active query (V(q), T(q)) */
  uint64_t nodes_visited;   /* sum_q V(q): 64-byte node records a query's warp read           */
  uint64_t points_tested;   /* sum_q T(q): float4 points distance-tested per query            */
  uint64_t heap_inserts;    /* accepted candidates                                            */
  uint64_t warp_node_visits;/* node records actually loaded (once per warp)                   */
  uint64_t warp_leaf_visits;/* leaves actually loaded (once per warp)                         */
  uint64_t warp_point_loads;/* float4 points actually loaded (once per warp)                  */
  uint64_t filter_violations;/* counting build: pairs the pre-filter rejected but the exact test accepts; must be 0 */
  uint64_t h2d_bytes, d2h_bytes; /* staging traffic of the last build / search */
} tknn_stats;

/* Replaces owlContextCreate + module/program/pipeline/SBT setup (hostCode.cpp:141-159,267-269).
 * device = CUDA ordinal.  Fails with TKNN_ECUDA when no usable device exists. */
TKNN_API int tknn_create(int device, tknn_ctx** out);

/* Replaces owlContextDestroy (hostCode.cpp:362): frees every device allocation of the context. */
TKNN_API int tknn_destroy(tknn_ctx* ctx);

/* Run all work of this context on `cuda_stream` (a cudaStream_t; NULL = the context's own
 * non-blocking stream; pass cudaStreamLegacy / cudaStreamPerThread for the default streams).  Inputs
 * produced on another stream must be complete before a call: the library only orders its work on the
 * stream it was given.  The reference uses one stream per LaunchParams (owl/LaunchParams.cpp:46). */
TKNN_API int tknn_set_stream(tknn_ctx* ctx, void* cuda_stream);

TKNN_API int tknn_set_option(tknn_ctx* ctx, int key, int64_t value);

/* Replaces owlDeviceBufferCreate(points) + owlUserGeomGroupCreate + owlGroupBuildAccel(GAS) +
 * owlInstanceGroupCreate + owlGroupBuildAccel(IAS) (hostCode.cpp:165-175,199-212), i.e.
 * UserGeomGroup::buildAccel (owl/UserGeomGroup.cpp:39-241) and the bounds program
 * (deviceCode.cu:38-56): builds an LBVH over the n points.
 * xyz: n rows of `stride_floats` floats (>= dim); dim 2 => z = 0 (hostCode.cpp:114-118), dim 3. */
TKNN_API int tknn_build(tknn_ctx* ctx, const float* xyz, uint64_t n, int dim, int stride_floats);

/* Replaces the whole round loop (hostCode.cpp:285-340: owlLaunch2D + host termination scan +
 * radius *= 2 + 2x owlGroupRefitAccel) and the raygen / intersection programs
 * (deviceCode.cu:62-152).  All n points are the queries (samples/s01-trueknn/README.md:2).
 * start_radius > 0: round 1 searches the closed ball of that radius, unresolved queries double it
 * (hostCode.cpp:323); start_radius <= 0: estimated from a sample (Util/random_sample.py's role);
 * start_radius = +inf: one unbounded round.  idx_out/dist_out: n*k each, rows in build order.
 * dist_out may be NULL (tknn_search and tknn_search_shard): indices only — with host outputs that halves the
 * device->host copy, which bounds the end-to-end time of a large search.
 * k must satisfy 1 <= k <= min(n-1, TKNN_MAX_K) (the reference never terminates for k > n-1). */
TKNN_API int tknn_search(tknn_ctx* ctx, int k, float start_radius, int32_t* idx_out, float* dist_out);

/* Query-sharded variant: this context answers shard `shard` of `n_shards`
 * contiguous Morton slices of the queries over its (replicated) BVH.  Rows are written compactly:
 * row r holds the neighbours of query qid_out[r]; *n_out rows are produced.  Capacity of each
 * output array must be at least tknn_shard_capacity(n, n_shards) rows. */
TKNN_API int tknn_search_shard(tknn_ctx* ctx, int k, float start_radius, int shard, int n_shards, int32_t* qid_out,
                               int32_t* idx_out, float* dist_out, uint64_t* n_out);
TKNN_API uint64_t tknn_shard_capacity(uint64_t n, int n_shards);

/* Separate query set on the built BVH.  queries: nq rows; self_ids (may be
 * NULL) names the data index to exclude per query (-1: none); init_radius2 (may be NULL): per-query
 * cap on the SQUARED distance, closed (d2 <= cap; negative = no cap) — squared so that the
 * point-partitioned driver can pass a k-th-neighbour d2 bit-exactly for its boundary queries.
 * Output rows follow the query order; rows with fewer than k neighbours end in -1 / FLT_MAX. */
TKNN_API int tknn_query(tknn_ctx* ctx, const float* queries, uint64_t nq, int dim, int stride_floats,
                        const int32_t* self_ids, const float* init_radius2, int k, float start_radius,
                        int32_t* idx_out, float* dist_out);

/* Fixed-radius neighbour count per point (closed ball, self excluded) on the same traversal
 * (the DBSCAN core-point test).  count_out: n entries, build order. */
TKNN_API int tknn_range_count(tknn_ctx* ctx, float radius, uint32_t* count_out);

/* Start-radius estimator alone (replaces samples/s01-trueknn/Util/random_sample.py:5-32). */
TKNN_API int tknn_estimate_start_radius(tknn_ctx* ctx, int k, float* radius_out);

/* Exact brute-force kNN of selected data points on the GPU (tiled, no BVH): the second oracle for
 * sampled queries at sizes no CPU oracle reaches.  query_ids: nq data
 * indices (host or device).  Output rows follow query_ids. */
TKNN_API int tknn_brute_force(tknn_ctx* ctx, const int32_t* query_ids, uint64_t nq, int k, int32_t* idx_out,
                              float* dist_out);

/* k-way merge of partial neighbour lists on (d2, index): for each of nq rows, merge `parts` lists
 * of k (idx, d2) pairs (layout [parts][nq][k], a list ends at its first -1 index, duplicates by
 * index removed) into one list of k.  d2_parts holds SQUARED distances (produced with
 * TKNN_OPT_SQUARED_DIST = 1: sqrtf is not injective, so only d2 gives the exact order); dist_out
 * receives sqrtf(d2), or d2 again while TKNN_OPT_SQUARED_DIST is set (chained merges).  Used by the
 * point-partitioned driver. */
TKNN_API int tknn_merge_topk(tknn_ctx* ctx, const int32_t* idx_parts, const float* d2_parts, int parts, uint64_t nq,
                             int k, int32_t* idx_out, float* dist_out);

TKNN_API int tknn_get_stats(const tknn_ctx* ctx, tknn_stats* out);
TKNN_API const char* tknn_last_error(const tknn_ctx* ctx);
TKNN_API int tknn_version(void);

/* ================================ multi-GPU ================================
 * The reference's multi-GPU model is "replicate every object on every device and split the launch"
 * (owl/RayGen.cpp:150-200, owl/Context.cpp:75-120), created from a device list (owlContextCreate(ids, n),
 * owl/include/owl/owl_host.h:360) and unused by its sample (hostCode.cpp:141 passes one device).  Both
 * variants below live INSIDE the library, over NCCL (NVLink 5 / NVSwitch):
 *
 *   TKNN_SHARD_QUERIES    the BVH is replicated on every GPU, GPU g answers the g-th contiguous Morton slice of
 *                         the queries; no collective on the search path (one all-gather of the points at build).
 *   TKNN_PARTITION_POINTS every GPU owns one Morton range of the points (clouds that should not be replicated):
 *                         redistribution by Morton range, local LBVH, local search, boundary-query exchange,
 *                         remote bounded search, partial top-k merge on (d2, GLOBAL index).
 *
 * Two ways in: (a) ONE PROCESS drives all devices — tknn_create_multi (ncclCommInitAll inside) + tknn_multi_build
 * + tknn_multi_search, host arrays in, host arrays in FILE order out: the drop-in for the sample; (b) ONE RANK
 * PER PROCESS (torchrun, MPI): every rank creates its own tknn_ctx, rank 0 makes a unique id, everyone calls
 * tknn_comm_init, then the per-rank calls below; results stay sharded on the ranks.
 * Collective calls must be made by every rank in the same order.  NCCL failures return TKNN_ENCCL. */

enum { TKNN_SHARD_QUERIES = 1, TKNN_PARTITION_POINTS = 2 };

#define TKNN_UNIQUE_ID_BYTES 128

typedef struct tknn_dist_stats {
  int32_t n_ranks, rank;
  uint64_t n_global, n_owned;
  /* build, ms between CUDA events on this rank's stream (time spent waiting for peers inside a collective included) */
  float h2d_ms;          /* staging of this rank's slice (0 for device input)                              */
  float allgather_ms;    /* replicated build: the one all-gather of the points                             */
  float box_ms;          /* partitioned build: global box (local reduce + all-reduce)                      */
  float codes_ms;        /* Morton codes on the global grid + cell histogram + its all-reduce              */
  float splitters_ms;    /* histogram read-back and the host's splitter choice                             */
  float bucket_ms;       /* destination ranks, one onesweep pass on them, row packing                      */
  float exchange_ms;     /* all-to-all of (x, y, z, global id) rows                                        */
  float lbvh_ms;         /* the local (or replicated) LBVH build                                           */
  float summaries_ms;    /* per-cell boxes + all-gather                                                    */
  float build_total_ms;
  /* partitioned search */
  float local_search_ms, reach_ms, exchange_out_ms, remote_search_ms, exchange_back_ms, merge_ms, finish_ms, d2h_ms;
  float search_total_ms; /* first launch -> last result written on the device                              */
  uint64_t boundary_sent, boundary_received;
  uint64_t bytes_sent_build, bytes_sent_search; /* bytes this rank handed to NCCL (self copies excluded)   */
} tknn_dist_stats;

/* (b) one rank per process.  tknn_comm_unique_id: rank 0 fills 128 bytes (ncclGetUniqueId) and ships them to the
 * other ranks by any means; tknn_comm_init: ncclCommInitRank on the context's device. */
TKNN_API int tknn_comm_unique_id(void* id128_out);
TKNN_API int tknn_comm_init(tknn_ctx* ctx, int n_ranks, int rank, const void* id128);
TKNN_API int tknn_get_dist_stats(const tknn_ctx* ctx, tknn_dist_stats* out);

/* TKNN_SHARD_QUERIES build: this rank holds rows [first, first + n_local) of the n_total-point cloud (host or
 * device); slices must tile the cloud in rank order.  Uploads the slice, assembles the cloud on every GPU with one
 * ncclAllGather, builds the replicated LBVH.  Search with tknn_search_shard(ctx, k, r, rank, n_ranks, ...). */
TKNN_API int tknn_build_replicated(tknn_ctx* ctx, const float* xyz_local, uint64_t n_local, uint64_t first,
                                   uint64_t n_total, int dim, int stride_floats);

/* TKNN_PARTITION_POINTS build: this rank STARTS with rows of global indices [first_index, first_index + n_local)
 * (host or device) and ENDS owning one Morton range of the global cloud (tknn_partition_owned rows).  Global
 * indices must stay below 2^31. */
TKNN_API int tknn_partition_build(tknn_ctx* ctx, const float* xyz_local, uint64_t n_local, uint64_t first_index, int dim,
                                  int stride_floats);
TKNN_API uint64_t tknn_partition_owned(const tknn_ctx* ctx);

/* Exact kNN of the owned points against the GLOBAL cloud.  Row r: gid_out[r] = global index of the query,
 * idx_out[r*k..] = global neighbour indices, dist_out[r*k..] = distances (d2 while TKNN_OPT_SQUARED_DIST is set).
 * Arrays (host or device) hold `capacity` >= tknn_partition_owned rows; *n_out rows are written. */
TKNN_API int tknn_partition_search(tknn_ctx* ctx, int k, float start_radius, int32_t* gid_out, int32_t* idx_out,
                                   float* dist_out, uint64_t capacity, uint64_t* n_out);

/* Proof at sizes no replica fits (2 B points): every rank samples `samples` of its result rows, all ranks brute-force
 * ALL sampled queries against their OWN points (exact tiled kernel, no BVH), the partial lists are all-gathered and
 * merged on (d2, global index), and each rank compares its rows bit for bit.  gid/idx/dist: the arrays
 * tknn_partition_search filled (n_rows rows).  *n_checked / *n_bad: GLOBAL sums (identical on every rank). */
TKNN_API int tknn_partition_verify(tknn_ctx* ctx, int k, int samples, const int32_t* gid, const int32_t* idx, const float* dist,
                                   uint64_t n_rows, uint64_t* n_checked, uint64_t* n_bad);

/* (a) one process, all devices.  device_ids: CUDA ordinals (owlContextCreate's device list); distinct devices talk
 * over NCCL (ncclCommInitAll); a list that NAMES ONE DEVICE SEVERAL TIMES runs that many ranks on it over an
 * in-process transport (device-to-device copies) — NCCL refuses two ranks on one GPU — which is how the single-GPU
 * test suite exercises the multi-rank paths.  mode: TKNN_SHARD_QUERIES | TKNN_PARTITION_POINTS. */
typedef struct tknn_multi tknn_multi;
TKNN_API int tknn_create_multi(const int* device_ids, int n_devices, int mode, tknn_multi** out);
TKNN_API int tknn_multi_destroy(tknn_multi* m);
TKNN_API int tknn_multi_set_option(tknn_multi* m, int key, int64_t value); /* applied to every rank's context */
/* xyz: n host rows; replaces the same calls as tknn_build on every device */
TKNN_API int tknn_multi_build(tknn_multi* m, const float* xyz, uint64_t n, int dim, int stride_floats);
/* idx_out / dist_out: HOST arrays, n*k each, rows in build (file) order — the layout of tknn_search.  Every rank
 * writes its result rows straight into the owner GPU's slice of the file-order array through peer pointers (P2P
 * stores over NVLink), then each GPU copies its contiguous slice to the host. */
TKNN_API int tknn_multi_search(tknn_multi* m, int k, float start_radius, int32_t* idx_out, float* dist_out);
TKNN_API int tknn_multi_ranks(const tknn_multi* m);
TKNN_API tknn_ctx* tknn_multi_ctx(tknn_multi* m, int rank); /* per-rank context: stats, options (owned by m) */
TKNN_API const char* tknn_multi_last_error(const tknn_multi* m);
/* wall-clock phases of the last tknn_multi_build / tknn_multi_search (ms): [0] build, [1] search (all ranks done),
 * [2] row exchange to the owners, [3] device->host copies */
TKNN_API int tknn_multi_get_times(const tknn_multi* m, float* ms4);

/* ---- introspection used by the tests and the bench (not part of the drop-in surface) ---- */

/* The sort-key layout tknn_build would choose for n points under TKNN_OPT_MORTON_BITS = morton_bits (0 = auto) and
 * TKNN_OPT_SORT_MODE = sort_mode: curve-code bits per axis, index bits packed below the code (0 = pair sort) and the
 * number of 8-bit onesweep passes.  Pure host arithmetic: needs no context and no device. */
TKNN_API int tknn_key_layout(uint64_t n, int morton_bits, int sort_mode, int* code_bits_per_axis, int* index_bits,
                             int* sort_passes);

/* Stand-alone onesweep radix sort of (u64 key, u32 value) pairs, stable, in place; device or host
 * pointers.  values == NULL: keys only (the kernel variant behind the builder's packed keys). */
TKNN_API int tknn_sort_pairs(tknn_ctx* ctx, uint64_t* keys, uint32_t* values, uint64_t n);

/* Copies of the built BVH: nodes (n_nodes x 16 words), sorted points
 * (n x float4: x, y, z, original index bits), leaf starts (n_leaves + 1).  NULL pointers skipped. */
TKNN_API int tknn_get_bvh(const tknn_ctx* ctx, void* nodes_out, void* points_out, uint32_t* leaf_start_out);

/* 63-bit Morton codes of n points on the cubic grid spanning box = {lo.xyz, hi.xyz} (the same code
 * the builder sorts by); the point-partitioned driver uses it with the GLOBAL box to assign Morton
 * ranges to ranks.  Device or host arrays. */
TKNN_API int tknn_morton_codes(tknn_ctx* ctx, const float* xyz, uint64_t n, int dim, int stride_floats,
                               const float* box6, uint64_t* codes_out);

/* Point-partitioned driver: for each of n points with squared reach reach2[i], the set of
 * REMOTE ranks whose published per-cell boxes the closed ball (p, sqrt(reach2)) touches, as a bit mask
 * (bit s = rank s, n_ranks <= 32).  summaries: [n_ranks][8^cell_bits][6] floats (lo.xyz, hi.xyz; an
 * untouched cell holds an inverted box); box6: the global box the Morton grid spans.  Device arrays. */
TKNN_API int tknn_reach_mask(tknn_ctx* ctx, const float* xyz, uint64_t n, int stride_floats, const float* reach2,
                             const float* box6, const float* summaries, int n_ranks, int cell_bits, int self_rank,
                             uint32_t* mask_out);

/* Synthetic clouds generated on the device by the stateless hash of SURVEY.md §8d:
 * u(i,a) = (mix64(seed ^ ((3i+a) * 0x9E3779B97F4A7C15)) >> 40) * 2^-24.  Writes n rows of xyz
 * for indices [first, first+n) to a DEVICE or HOST array. */
TKNN_API int tknn_generate_uniform(tknn_ctx* ctx, uint64_t seed, uint64_t first, uint64_t n, float* xyz_out);

/* Point-file ingest and neighbour writer (host only, no CUDA).
 * tknn_read_points: the sample's file grammar (hostCode.cpp:83-124), mmap + parallel parse; names
 * ending in ".f32" are raw little-endian float32 rows of `dim`.  xyz_out: n_cap rows of 3 floats; when the rows do
 * not fit (the grammar consumes the last line whole, so more than n rows can come back) the call returns
 * TKNN_EINVAL with *n_out = the capacity needed.
 * tknn_write_neighbours: `query,neighbourIndex,distance` lines (the dump commented out at
 * hostCode.cpp:312-319), or raw arrays (path + ".idx.i32" / ".dist.f32") when binary != 0. */
TKNN_API int tknn_read_points(const char* path, uint64_t n, int dim, float* xyz_out, uint64_t n_cap, uint64_t* n_out);
TKNN_API int tknn_write_neighbours(const char* path, const int32_t* idx, const float* dist, uint64_t n, int k, int binary);

/* Device properties the roofline needs: SM count, L2 bytes, and a measured L2 / HBM read
 * bandwidth (GB/s) from a short read loop over an L2-resident / HBM-sized buffer. */
TKNN_API int tknn_measure_bandwidth(tknn_ctx* ctx, double* l2_gbs, double* hbm_gbs, int* sm_count, uint64_t* l2_bytes);

/* Measured aggregate shared-memory bandwidth (GB/s of bytes delivered to lanes): conflict-free LDS.128 — the
 * 128 B/clk/SM crossbar — and warp-broadcast LDS.128, the access pattern of the traversal kernel's leaf filter.
 * The traversal kernel is bound by this data path (and by issue slots), not by HBM: bench.py's roofline uses it. */
TKNN_API int tknn_measure_smem_bandwidth(tknn_ctx* ctx, double* conflict_free_gbs, double* broadcast_gbs);

#ifdef __cplusplus
}
#endif
#endif /* TRUEKNN_H */
