// common.cuh — shared device helpers for libtrueknn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtrueknn is written for sm_100a (B200) only"
#endif

namespace tknn {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int WARP = 32;
constexpr int MAX_LEAF = 32;       // a leaf is at most one warp-wide float4 load (512 B)
constexpr int STACK_DEPTH = 96;    // 63 Morton bits + 32 tie-break bits bounds the LBVH depth

// One BVH node = 64 B = 4 x float4, 64-byte aligned.  The two children's boxes are stored INTERLEAVED per
// axis so that every 8-byte pair is one packed fp32x2 operand (child 0 in the low half, child 1 in the
// high half) and the point-to-box distances of both children cost one FADD2/FMUL2/FFMA2 chain:
//   a = (lo0.x, lo1.x, lo0.y, lo1.y)   b = (lo0.z, lo1.z, hi0.x, hi1.x)
//   c = (hi0.y, hi1.y, hi0.z, hi1.z)   d = (ref0, ref1, count0, count1)        (d as int bits)
// count == 0 : internal child, ref = node index;  count > 0 : leaf child, ref = first sorted point.
struct __align__(64) Node {
  float4 a, b, c, d;
};
static_assert(sizeof(Node) == 64, "node record must be 64 bytes");

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// The one distance formula (see oracle/knn_oracle.c): explicit FMAs, never re-associated.
__device__ __forceinline__ float dist2(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Point-to-box squared distance with the same op order, so boxdist2 <= dist2(point) for every
// point inside the box under round-to-nearest monotonicity.
__device__ __forceinline__ float box_dist2(float qx, float qy, float qz, const float4& lo, const float4& hi) {
  const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.0f);
  const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.0f);
  const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.0f);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Both children of a node at once: the same chain as box_dist2 per component, on packed fp32x2 operands
// (x = child 0, y = child 1).  Bit-identical to two box_dist2 calls.
__device__ __forceinline__ float2 box_dist2_x2(float qx, float qy, float qz, const float4& a, const float4& b,
                                               const float4& c) {
  const float2 q_x = make_float2(qx, qx), q_y = make_float2(qy, qy), q_z = make_float2(qz, qz);
  const float2 n_x = make_float2(-qx, -qx), n_y = make_float2(-qy, -qy), n_z = make_float2(-qz, -qz);
  const float2 lx = __fadd2_rn(make_float2(a.x, a.y), n_x), hx = __fadd2_rn(q_x, make_float2(-b.z, -b.w));
  const float2 ly = __fadd2_rn(make_float2(a.z, a.w), n_y), hy = __fadd2_rn(q_y, make_float2(-c.x, -c.y));
  const float2 lz = __fadd2_rn(make_float2(b.x, b.y), n_z), hz = __fadd2_rn(q_z, make_float2(-c.z, -c.w));
  const float2 dx = make_float2(fmaxf(fmaxf(lx.x, hx.x), 0.0f), fmaxf(fmaxf(lx.y, hx.y), 0.0f));
  const float2 dy = make_float2(fmaxf(fmaxf(ly.x, hy.x), 0.0f), fmaxf(fmaxf(ly.y, hy.y), 0.0f));
  const float2 dz = make_float2(fmaxf(fmaxf(lz.x, hz.x), 0.0f), fmaxf(fmaxf(lz.y, hz.y), 0.0f));
  return __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
}

// (d2, idx) -> one u64 whose unsigned order is "smaller d2, then lower index" (d2 >= 0).
__device__ __forceinline__ uint64_t make_key(float d2, int idx) {
  return ((uint64_t)__float_as_uint(d2) << 32) | (uint32_t)idx;
}
__device__ __forceinline__ float key_d2(uint64_t key) { return __uint_as_float((uint32_t)(key >> 32)); }
__device__ __forceinline__ int key_idx(uint64_t key) { return (int)(uint32_t)(key & 0xffffffffu); }

// float <-> order-preserving uint (for atomicMin/Max on floats of either sign)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace tknn
