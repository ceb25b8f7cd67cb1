/*
 * oracle/knn_oracle.c — CPU oracle for the TrueKNN hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; nothing under owlraytracing_b200/ links, imports or calls it.
 *
 * PARITY UNPINNED by the reference: vani-nag/OWLRayTracing ships no golden vectors, no
 * known-answer tests and no dataset for samples/s01-trueknn, and the
 * sample cannot be compiled here (it needs the OptiX SDK + libnvoptix; BVH build and
 * traversal live in that closed third-party library: OptiX 7.0-7.2 per README.md:22,
 * call sites owl/UserGeomGroup.cpp:175,199, owl/RayGen.cpp:191, owl/include/owl/owl_device.h:161).
 * What this file pins instead:
 *   - tko_knn_brute   : the exact all-points kNN the north star mandates ((d2, index)
 *                       lexicographic order, self excluded by index) — ground truth.
 *   - tko_knn_kdtree  : exact kd-tree, must equal tko_knn_brute bit-for-bit (tests do that)
 *                       before it is trusted at 10 M points; also the timed CPU baseline.
 *   - tko_reference_trueknn : a restatement of the reference's *own* rules (rounds,
 *                       AABB/L-inf candidate set, strict-< sorted insert, count-only
 *                       termination) from samples/s01-trueknn/deviceCode.cu:38-152 and
 *                       hostCode.cpp:126-130,285-340.  It is inexact by construction
 *                       (SURVEY.md F2); it becomes exact when the start radius covers the
 *                       whole cloud, which is the one point where the reference's semantics
 *                       and the exact oracle provably coincide — tests pin that equality.
 *   - tko_parse_points: the reference's point-file grammar (hostCode.cpp:83-124).
 *
 * Distance arithmetic (shared by every function here and by the CUDA kernels):
 *     dx = qx - px; dy = qy - py; dz = qz - pz          (fp32, round-to-nearest)
 *     d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx))
 * This is the reference's `(x*x) + (y*y) + (z*z)` (deviceCode.cu:110-113) under nvcc's
 * default FMA contraction, written out explicitly so host and device agree bit-for-bit.
 * Reported distance = sqrtf(d2) (IEEE, not the reference's --use_fast_math approximation).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <ctype.h>
#include <errno.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TKO_API __attribute__((visibility("default")))
/* runtime-dispatched clones: the "fma" clone inlines fmaf to vfmadd, the default clone calls libm's
 * correctly rounded fmaf — both give identical bits. */
#if defined(__x86_64__)
#define TKO_CLONES __attribute__((target_clones("fma", "default")))
#else
#define TKO_CLONES
#endif
#define TKO_INLINE static inline __attribute__((always_inline))

TKO_INLINE float tko_d2(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = qx - px, dy = qy - py, dz = qz - pz;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* (d2, idx) packed so that one unsigned compare is the lexicographic order; d2 >= 0 so its
 * bit pattern orders like the float. */
TKO_INLINE uint64_t tko_key(float d2, int32_t idx) {
  uint32_t b;
  memcpy(&b, &d2, 4);
  return ((uint64_t)b << 32) | (uint32_t)idx;
}

/* bounded sorted list: keep the k smallest keys, ascending.  cnt = how many are valid. */
TKO_INLINE void tko_list_insert(uint64_t* list, int* cnt, int k, uint64_t key) {
  int c = *cnt;
  if (c == k) {
    if (key >= list[k - 1]) return;
    c = k - 1;
  }
  int i = c;
  while (i > 0 && list[i - 1] > key) {
    list[i] = list[i - 1];
    --i;
  }
  list[i] = key;
  *cnt = c + 1;
}

static int tko_squared_output = 0; /* 1: report d2 instead of sqrtf(d2) (stand-in engine of the multi-GPU host tests) */

TKO_API void tko_set_squared_output(int on) { tko_squared_output = on ? 1 : 0; }

TKO_INLINE void tko_list_emit(const uint64_t* list, int cnt, int k, int32_t* idx_out, float* dist_out) {
  for (int i = 0; i < k; ++i) {
    if (i < cnt) {
      uint32_t b = (uint32_t)(list[i] >> 32);
      float d2;
      memcpy(&d2, &b, 4);
      idx_out[i] = (int32_t)(uint32_t)(list[i] & 0xffffffffu);
      dist_out[i] = tko_squared_output ? d2 : sqrtf(d2);
    } else { /* unfilled-slot sentinels, hostCode.cpp:129 */
      idx_out[i] = -1;
      dist_out[i] = FLT_MAX;
    }
  }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm of bench.py must still use every core */
TKO_API void tko_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

TKO_API int tko_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

TKO_API float tko_dist2(const float* q, const float* p) { return tko_d2(q[0], q[1], q[2], p[0], p[1], p[2]); }

/* ------------------------------------------------------------------------------------------
 * Ground truth: brute force.  Queries are rows of `q` (nq x 3); data are rows of `xyz` (n x 3).
 * self_ids[i] (or i itself when self_ids == NULL and q == xyz) is excluded BY INDEX, so
 * coincident duplicates are legal neighbours at distance 0 (deviceCode.cu:103).
 * self_ids[i] < 0 excludes nothing.  radius2 >= 0 restricts to d2 <= radius2 (closed ball);
 * pass INFINITY for plain kNN.
 * ---------------------------------------------------------------------------------------- */
TKO_CLONES
TKO_API int tko_knn_brute_queries(const float* xyz, int64_t n, const float* q, int64_t nq, const int32_t* self_ids,
                                  int k, float radius2, const float* caps2, int32_t* idx_out, float* dist_out) {
  if (!xyz || !q || n < 0 || nq < 0 || k <= 0) return 1;
#pragma omp parallel
  {
    uint64_t* list = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)k);
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < nq; ++i) {
      const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
      const int64_t self = self_ids ? self_ids[i] : (q == xyz ? i : -1);
      int cnt = 0;
      float worst = radius2;
      if (caps2 && caps2[i] >= 0.0f && caps2[i] < worst) worst = caps2[i]; /* per-query closed cap on d2 */
      for (int64_t j = 0; j < n; ++j) {
        float d2 = tko_d2(qx, qy, qz, xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]);
        if (d2 > worst || j == self) continue;
        tko_list_insert(list, &cnt, k, tko_key(d2, (int32_t)j));
        if (cnt == k) {
          uint32_t b = (uint32_t)(list[k - 1] >> 32);
          memcpy(&worst, &b, 4);
        }
      }
      tko_list_emit(list, cnt, k, idx_out + i * k, dist_out + i * k);
    }
    free(list);
  }
  return 0;
}

TKO_API int tko_knn_brute(const float* xyz, int64_t n, int k, int32_t* idx_out, float* dist_out) {
  return tko_knn_brute_queries(xyz, n, xyz, n, NULL, k, INFINITY, NULL, idx_out, dist_out);
}

/* ------------------------------------------------------------------------------------------
 * Exact kd-tree (implicit, balanced, points permuted into tree order, tight box per node).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float lo[3], hi[3];
} tko_box;

typedef struct tko_kdtree {
  int64_t n;
  int leaf;        /* max points per leaf */
  float* pts;      /* n x 3, tree order */
  int32_t* ids;    /* original index of each tree-order point */
  int64_t n_nodes; /* heap-numbered: node 0 = root, children 2i+1, 2i+2 */
  tko_box* boxes;
  int64_t* lo;     /* point range per node */
  int64_t* hi;
} tko_kdtree;

static void tko_box_of(const float* xyz, const int32_t* ids, int64_t lo, int64_t hi, tko_box* b) {
  for (int a = 0; a < 3; ++a) { b->lo[a] = INFINITY; b->hi[a] = -INFINITY; }
  for (int64_t i = lo; i < hi; ++i)
    for (int a = 0; a < 3; ++a) {
      float v = xyz[3 * (int64_t)ids[i] + a];
      if (v < b->lo[a]) b->lo[a] = v;
      if (v > b->hi[a]) b->hi[a] = v;
    }
}

/* quickselect on ids[lo,hi) by coordinate `a`, ties broken by id so the permutation is deterministic */
static inline int tko_less(const float* xyz, int a, int32_t i, int32_t j) {
  float vi = xyz[3 * (int64_t)i + a], vj = xyz[3 * (int64_t)j + a];
  return vi < vj || (vi == vj && i < j);
}
static void tko_select(const float* xyz, int32_t* ids, int64_t lo, int64_t hi, int64_t nth, int a) {
  while (hi - lo > 1) {
    int64_t mid = lo + (hi - lo) / 2;
    /* median of three */
    int32_t x = ids[lo], y = ids[mid], z = ids[hi - 1], p;
    if (tko_less(xyz, a, x, y)) p = tko_less(xyz, a, y, z) ? y : (tko_less(xyz, a, x, z) ? z : x);
    else p = tko_less(xyz, a, x, z) ? x : (tko_less(xyz, a, y, z) ? z : y);
    int64_t i = lo, j = hi - 1;
    while (i <= j) {
      while (tko_less(xyz, a, ids[i], p)) ++i;
      while (tko_less(xyz, a, p, ids[j])) --j;
      if (i <= j) { int32_t t = ids[i]; ids[i] = ids[j]; ids[j] = t; ++i; --j; }
    }
    if (nth <= j) hi = j + 1;
    else if (nth >= i) lo = i;
    else return;
  }
}

static void tko_build_rec(tko_kdtree* t, const float* xyz, int64_t node, int64_t lo, int64_t hi, int depth) {
  t->lo[node] = lo;
  t->hi[node] = hi;
  tko_box_of(xyz, t->ids, lo, hi, &t->boxes[node]);
  if (hi - lo <= t->leaf) return;
  const tko_box* b = &t->boxes[node];
  int a = 0;
  float e = b->hi[0] - b->lo[0];
  for (int c = 1; c < 3; ++c)
    if (b->hi[c] - b->lo[c] > e) { e = b->hi[c] - b->lo[c]; a = c; }
  int64_t mid = lo + (hi - lo) / 2;
  tko_select(xyz, t->ids, lo, hi, mid, a);
  if (depth < 6) {
#pragma omp task default(shared)
    tko_build_rec(t, xyz, 2 * node + 1, lo, mid, depth + 1);
#pragma omp task default(shared)
    tko_build_rec(t, xyz, 2 * node + 2, mid, hi, depth + 1);
#pragma omp taskwait
  } else {
    tko_build_rec(t, xyz, 2 * node + 1, lo, mid, depth + 1);
    tko_build_rec(t, xyz, 2 * node + 2, mid, hi, depth + 1);
  }
}

TKO_API tko_kdtree* tko_kdtree_build(const float* xyz, int64_t n, int leaf) {
  if (!xyz || n <= 0) return NULL;
  if (leaf < 1) leaf = 8;
  tko_kdtree* t = (tko_kdtree*)calloc(1, sizeof(*t));
  t->n = n;
  t->leaf = leaf;
  /* balanced halving: depth D with ceil(n / 2^D) <= leaf; heap numbering needs 2^(D+1) - 1 slots */
  int D = 0;
  while (((n + ((int64_t)1 << D) - 1) >> D) > leaf) ++D;
  t->n_nodes = ((int64_t)1 << (D + 1)) - 1;
  t->boxes = (tko_box*)malloc(sizeof(tko_box) * (size_t)t->n_nodes);
  t->lo = (int64_t*)malloc(sizeof(int64_t) * (size_t)t->n_nodes);
  t->hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)t->n_nodes);
  t->ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  t->pts = (float*)malloc(sizeof(float) * 3 * (size_t)n);
  for (int64_t i = 0; i < t->n_nodes; ++i) t->lo[i] = t->hi[i] = -1;
  for (int64_t i = 0; i < n; ++i) t->ids[i] = (int32_t)i;
#pragma omp parallel
#pragma omp single
  tko_build_rec(t, xyz, 0, 0, n, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int64_t s = t->ids[i];
    t->pts[3 * i] = xyz[3 * s];
    t->pts[3 * i + 1] = xyz[3 * s + 1];
    t->pts[3 * i + 2] = xyz[3 * s + 2];
  }
  return t;
}

TKO_API void tko_kdtree_free(tko_kdtree* t) {
  if (!t) return;
  free(t->boxes); free(t->lo); free(t->hi); free(t->ids); free(t->pts); free(t);
}

/* point-to-box squared distance with the SAME op order as tko_d2, so RN monotonicity gives
 * boxdist2 <= d2(point) for every point inside the box (ties survive: prune only on strict >). */
TKO_INLINE float tko_boxd2(const tko_box* b, float qx, float qy, float qz) {
  float dx = fmaxf(fmaxf(b->lo[0] - qx, qx - b->hi[0]), 0.0f);
  float dy = fmaxf(fmaxf(b->lo[1] - qy, qy - b->hi[1]), 0.0f);
  float dz = fmaxf(fmaxf(b->lo[2] - qz, qz - b->hi[2]), 0.0f);
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

TKO_INLINE void tko_kd_query(const tko_kdtree* t, float qx, float qy, float qz, int64_t self, int k, float radius2,
                             uint64_t* list, int* cnt_io) {
  int64_t stack[128];
  int sp = 0;
  int cnt = 0;
  float worst = radius2;
  stack[sp++] = 0;
  while (sp) {
    int64_t node = stack[--sp];
    if (tko_boxd2(&t->boxes[node], qx, qy, qz) > worst) continue;
    int64_t c0 = 2 * node + 1, c1 = c0 + 1;
    if (c1 >= t->n_nodes || t->lo[c0] < 0) { /* leaf */
      for (int64_t i = t->lo[node]; i < t->hi[node]; ++i) {
        float d2 = tko_d2(qx, qy, qz, t->pts[3 * i], t->pts[3 * i + 1], t->pts[3 * i + 2]);
        if (d2 > worst || t->ids[i] == self) continue;
        tko_list_insert(list, &cnt, k, tko_key(d2, t->ids[i]));
        if (cnt == k) {
          uint32_t b = (uint32_t)(list[k - 1] >> 32);
          memcpy(&worst, &b, 4);
        }
      }
      continue;
    }
    float d0 = tko_boxd2(&t->boxes[c0], qx, qy, qz), d1 = tko_boxd2(&t->boxes[c1], qx, qy, qz);
    if (d0 <= d1) { stack[sp++] = c1; stack[sp++] = c0; }
    else { stack[sp++] = c0; stack[sp++] = c1; }
  }
  *cnt_io = cnt;
}

/* queries in tree order when q == NULL (all points are queries, self excluded by index) */
TKO_CLONES
TKO_API int tko_kdtree_knn(const tko_kdtree* t, const float* q, int64_t nq, const int32_t* self_ids, int k,
                           float radius2, int32_t* idx_out, float* dist_out) {
  if (!t || k <= 0) return 1;
  const int all = (q == NULL);
  if (all) nq = t->n;
#pragma omp parallel
  {
    uint64_t* list = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; ++i) {
      float qx, qy, qz;
      int64_t self, row;
      if (all) {
        qx = t->pts[3 * i]; qy = t->pts[3 * i + 1]; qz = t->pts[3 * i + 2];
        self = t->ids[i]; row = self;
      } else {
        qx = q[3 * i]; qy = q[3 * i + 1]; qz = q[3 * i + 2];
        self = self_ids ? self_ids[i] : -1; row = i;
      }
      int cnt = 0;
      tko_kd_query(t, qx, qy, qz, self, k, radius2, list, &cnt);
      tko_list_emit(list, cnt, k, idx_out + row * k, dist_out + row * k);
    }
    free(list);
  }
  return 0;
}

/* Queries = the points at TREE positions pos[0..nq) (self excluded by index), rows in the order of pos.  Runs of
 * consecutive positions are neighbours in space, which is the order the GPU arm processes its queries in (Morton
 * order): the fair form of a sampled CPU baseline (bench.py) — index-sorted samples are cache-hostile. */
TKO_CLONES
TKO_API int tko_kdtree_knn_positions(const tko_kdtree* t, const int64_t* pos, int64_t nq, int k, float radius2,
                                     int32_t* idx_out, float* dist_out) {
  if (!t || !pos || k <= 0) return 1;
#pragma omp parallel
  {
    uint64_t* list = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < nq; ++i) {
      const int64_t p = pos[i];
      int cnt = 0;
      if (p >= 0 && p < t->n)
        tko_kd_query(t, t->pts[3 * p], t->pts[3 * p + 1], t->pts[3 * p + 2], t->ids[p], k, radius2, list, &cnt);
      tko_list_emit(list, cnt, k, idx_out + i * k, dist_out + i * k);
    }
    free(list);
  }
  return 0;
}

/* original index of the point at each tree position (to check sampled answers) */
TKO_API int tko_kdtree_ids(const tko_kdtree* t, const int64_t* pos, int64_t nq, int32_t* ids_out) {
  if (!t || !pos) return 1;
  for (int64_t i = 0; i < nq; ++i) ids_out[i] = (pos[i] >= 0 && pos[i] < t->n) ? t->ids[pos[i]] : -1;
  return 0;
}

/* one-call all-points kNN via the kd-tree; build_s / query_s are wall seconds (may be NULL) */
TKO_API int tko_knn_kdtree(const float* xyz, int64_t n, int k, int32_t* idx_out, float* dist_out, double* build_s,
                           double* query_s) {
  double t0 = 0, t1 = 0, t2 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  tko_kdtree* t = tko_kdtree_build(xyz, n, 8);
  if (!t) return 1;
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  int rc = tko_kdtree_knn(t, NULL, 0, NULL, k, INFINITY, idx_out, dist_out);
#ifdef _OPENMP
  t2 = omp_get_wtime();
#endif
  if (build_s) *build_s = t1 - t0;
  if (query_s) *query_s = t2 - t1;
  tko_kdtree_free(t);
  return rc;
}

/* fixed-radius neighbour count (closed ball, self excluded by index) — oracle for the range /
 * DBSCAN core-count row. */
TKO_CLONES
TKO_API int tko_range_count(const float* xyz, int64_t n, float radius, uint32_t* count_out) {
  if (!xyz || n < 0) return 1;
  const float r2 = radius * radius;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i) {
    uint32_t c = 0;
    for (int64_t j = 0; j < n; ++j)
      if (j != i && tko_d2(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]) <= r2) ++c;
    count_out[i] = c;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Restatement of the reference's own algorithm (inexact; for the equivalence pin and reporting).
 *   state        hostCode.cpp:126-130     k slots {ind=-1, dist=FLT_MAX}, numNeighbors = k
 *   candidates   deviceCode.cu:38-48      prim box = [c - rad, c + rad]; the degenerate ray at the
 *                deviceCode.cu:149-151    query reports every prim whose box contains the query
 *   per cand.    deviceCode.cu:77-85      skip if already in the list (by index)
 *                deviceCode.cu:103        skip self (by index)
 *                deviceCode.cu:110-116    dist = sqrt(...); accept iff dist < list[k-1].dist
 *                deviceCode.cu:118-134    numNeighbors = max(numNeighbors-1, 0); strict-< sorted insert
 *   rounds       hostCode.cpp:285-340     any query with numNeighbors > 0 => radius *= 2, again
 * Candidate arrival order inside a round is hardware traversal order (unspecified); this
 * restatement enumerates in index order.  O(n^2) per round: small n only.
 * ---------------------------------------------------------------------------------------- */
TKO_CLONES
TKO_API int tko_reference_trueknn(const float* xyz, int64_t n, int k, float radius, int max_rounds, int32_t* idx_out,
                                  float* dist_out, int* rounds_out, float* final_radius_out) {
  if (!xyz || n <= 0 || k <= 0 || !(radius > 0.0f)) return 1;
  int32_t* need = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    need[i] = k;
    for (int j = 0; j < k; ++j) { idx_out[i * k + j] = -1; dist_out[i * k + j] = FLT_MAX; }
  }
  int rounds = 0, found = 0;
  while (!found && rounds < max_rounds) {
    ++rounds;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t x = 0; x < n; ++x) {
      if (need[x] <= 0) continue; /* raygen early-out, deviceCode.cu:149 */
      const float qx = xyz[3 * x], qy = xyz[3 * x + 1], qz = xyz[3 * x + 2];
      int32_t* li = idx_out + x * k;
      float* ld = dist_out + x * k;
      for (int64_t p = 0; p < n; ++p) {
        const float cx = xyz[3 * p], cy = xyz[3 * p + 1], cz = xyz[3 * p + 2];
        if (qx < cx - radius || qx > cx + radius || qy < cy - radius || qy > cy + radius || qz < cz - radius ||
            qz > cz + radius)
          continue;
        int seen = 0;
        for (int i = 0; i < k; ++i)
          if (li[i] == (int32_t)p) { seen = 1; break; }
        if (seen || p == x) continue;
        float dx = cx - qx, dy = cy - qy, dz = cz - qz;
        float dist = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        if (dist < ld[k - 1]) {
          if (need[x] > 0) need[x] -= 1;
          int q = 0;
          for (; q < k; ++q)
            if (dist < ld[q]) break;
          for (int w = k - 1; w > q; --w) { ld[w] = ld[w - 1]; li[w] = li[w - 1]; }
          ld[q] = dist;
          li[q] = (int32_t)p;
        }
      }
    }
    found = 1;
    for (int64_t j = 0; j < n; ++j)
      if (need[j] > 0) { found = 0; radius *= 2; break; }
  }
  free(need);
  if (rounds_out) *rounds_out = rounds;
  if (final_radius_out) *final_radius_out = radius;
  return found ? 0 : 2;
}

/* ------------------------------------------------------------------------------------------
 * Point-file grammar, hostCode.cpp:83-124: read lines while floats are still owed
 * (count = n*dim, checked BEFORE each line, so a line is always consumed whole); on each
 * line extract floats with operator>>, skipping one ',' after a float; then chunk the
 * flat float vector by dim (dim 2 => z = 0; dim 3 => xyz; anything else => no points).
 * The reference throws std::out_of_range when the float count is not a multiple of dim;
 * this restatement returns -2 there.  Returns the number of points, or < 0 on error.
 * ---------------------------------------------------------------------------------------- */
TKO_API int64_t tko_parse_points(const char* text, int64_t len, int64_t n, int dim, float* xyz_out, int64_t cap) {
  if (!text || len < 0 || n < 0) return -1;
  int64_t owed = n * dim;
  int64_t nf = 0, fcap = 1024;
  float* v = (float*)malloc(sizeof(float) * (size_t)fcap);
  int64_t pos = 0;
  while (pos < len && owed > 0) {
    int64_t eol = pos;
    while (eol < len && text[eol] != '\n') ++eol;
    /* one line: [pos, eol) */
    int64_t p = pos;
    for (;;) {
      while (p < eol && isspace((unsigned char)text[p])) ++p;
      if (p >= eol) break;
      char buf[128];
      int64_t m = eol - p < 127 ? eol - p : 127;
      memcpy(buf, text + p, (size_t)m);
      buf[m] = 0;
      char* end = NULL;
      errno = 0;
      float f = strtof(buf, &end);
      if (end == buf) break; /* operator>> fails: rest of the line is dropped */
      if (nf == fcap) { fcap *= 2; v = (float*)realloc(v, sizeof(float) * (size_t)fcap); }
      v[nf++] = f;
      --owed;
      p += end - buf;
      if (p < eol && text[p] == ',') ++p;
    }
    pos = eol + 1;
  }
  int64_t np = 0;
  if (dim == 2 || dim == 3) {
    if (nf % dim) { free(v); return -2; }
    np = nf / dim;
    if (np > cap) { free(v); return -3; }
    for (int64_t i = 0; i < np; ++i) {
      xyz_out[3 * i] = v[dim * i];
      xyz_out[3 * i + 1] = v[dim * i + 1];
      xyz_out[3 * i + 2] = dim == 3 ? v[dim * i + 2] : 0.0f;
    }
  }
  free(v);
  return np;
}
