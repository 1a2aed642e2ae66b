// Ray casting against the map on the device: a linear BVH over the faces (Morton order, Karras' radix tree, bottom-up
// refit), one thread per ray.  Replaces the lvr2 raycaster behind MeshMap::raycaster() (mesh_map.h:318, built at
// mesh_map.cpp:317-321) for its two users on the hot path's input side:
//   * ObstacleLayer::processPointCloud (obstacle_layer.cpp:215-296): sensor points -> rays along the down axis -> lethal set
//   * lvr2::calcNormalClearance (clearance_layer.cpp:161): one ray per vertex along its normal -> free space above it
// The ray/triangle arithmetic is the one the oracle states (oracle.cpp "Ray casting against the map"): float, two-sided
// Moeller-Trumbore, evaluated only inside the triangle's box grown by eps, hit accepted for tn <= t <= tf of that box,
// nearest t wins, ties to the smallest face id.  Every node box is the exact union of the grown triangle boxes below it and
// box_span() is monotone in the box (float subtraction and multiplication by one factor are monotone), so a triangle whose
// own box passes is reachable through all its ancestors and pruning by "entry > best t" can never drop a hit the
// brute-force loop of the oracle would keep: the tree changes the cost, not the result.
#pragma once
#include <cstdint>

namespace mnb {

constexpr uint32_t RAY_NONE = 0xffffffffu;
constexpr int RAY_STACK = 96;      // Morton radix tree: depth <= 63 key bits + log2(F) index bits

struct RayBvh {
  uint32_t n = 0;                  // faces
  float eps = 0.0f;                // box growth: 1e-5f * max |coordinate|
  float4* lo = nullptr;            // 2n-1 nodes: xyz = box minimum, w = left child (internal) or face id (leaf) as bits
  float4* hi = nullptr;            //             xyz = box maximum, w = right child (internal) or RAY_NONE (leaf)
  uint32_t* parent = nullptr;      // 2n-1
  unsigned int* visits = nullptr;  // n-1 refit counters
  unsigned long long* keys = nullptr; uint32_t* order = nullptr;      // Morton keys / face ids, sorted
  unsigned long long* keys_tmp = nullptr; uint32_t* order_tmp = nullptr;
  unsigned int* scene = nullptr;   // [0..2] ordered-uint minimum per axis, [3..5] maximum, [6] bits of max |coordinate|
};

struct Ray { float o[3], d[3]; };

__device__ __forceinline__ unsigned int ordered_bits(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// per-axis slabs of a box: entry tn (clamped at 0) and exit tf of the ray; false when the ray misses the box
__device__ __forceinline__ bool box_span(const Ray& r, const float4 lo, const float4 hi, float& tn, float& tf) {
  tn = 0.0f; tf = __uint_as_float(0x7f800000u);
  const float blo[3] = {lo.x, lo.y, lo.z}, bhi[3] = {hi.x, hi.y, hi.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!(fabsf(r.d[k]) >= 1e-20f)) {
      if (r.o[k] < blo[k] || r.o[k] > bhi[k]) return false;
    } else {
      const float inv = 1.0f / r.d[k];
      const float t1 = (blo[k] - r.o[k]) * inv, t2 = (bhi[k] - r.o[k]) * inv;
      tn = fmaxf(tn, fminf(t1, t2));
      tf = fminf(tf, fmaxf(t1, t2));
    }
  }
  return tn <= tf;
}

__device__ __forceinline__ bool ray_triangle(const Ray& r, const float* __restrict__ a, const float* __restrict__ b,
                                             const float* __restrict__ c, float& t) {
  const float e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
  const float e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
  const float px = r.d[1] * e2z - r.d[2] * e2y, py = r.d[2] * e2x - r.d[0] * e2z, pz = r.d[0] * e2y - r.d[1] * e2x;
  const float det = (e1x * px + e1y * py) + e1z * pz;
  if (det == 0.0f) return false;
  const float inv = 1.0f / det;
  const float sx = r.o[0] - a[0], sy = r.o[1] - a[1], sz = r.o[2] - a[2];
  const float u = ((sx * px + sy * py) + sz * pz) * inv;
  if (!(u >= 0.0f && u <= 1.0f)) return false;
  const float qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
  const float v = ((r.d[0] * qx + r.d[1] * qy) + r.d[2] * qz) * inv;
  if (!(v >= 0.0f && u + v <= 1.0f)) return false;
  t = ((e2x * qx + e2y * qy) + e2z * qz) * inv;
  return t >= 0.0f;
}

// ---- build ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bvh_scene(const float* __restrict__ pos, uint32_t V, unsigned int* __restrict__ scene) {
  unsigned int lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u}, mx = 0u;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; ++k) {
      const float x = pos[3 * v + k];
      const unsigned int o = ordered_bits(x);
      lo[k] = min(lo[k], o); hi[k] = max(hi[k], o);
      mx = max(mx, __float_as_uint(fabsf(x)));
    }
  for (int k = 0; k < 3; ++k) { atomicMin(&scene[k], lo[k]); atomicMax(&scene[3 + k], hi[k]); }
  atomicMax(&scene[6], mx);
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {      // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(256) k_bvh_morton(const float* __restrict__ pos, const uint32_t* __restrict__ faces, uint32_t F,
                                                    const unsigned int* __restrict__ scene, unsigned long long* __restrict__ keys,
                                                    uint32_t* __restrict__ order) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const uint32_t a = faces[3 * (size_t)f], b = faces[3 * (size_t)f + 1], c = faces[3 * (size_t)f + 2];
  unsigned long long key = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = from_ordered_bits(scene[k]), hi = from_ordered_bits(scene[3 + k]);
    const float ctr = ((pos[3 * (size_t)a + k] + pos[3 * (size_t)b + k]) + pos[3 * (size_t)c + k]) * (1.0f / 3.0f);
    const float ext = hi - lo;
    float q = ext > 0.0f ? (ctr - lo) / ext : 0.0f;
    q = fminf(fmaxf(q * 2097152.0f, 0.0f), 2097151.0f);
    key |= spread21((unsigned long long)q) << (2 - k);
  }
  keys[f] = key; order[f] = f;
}

// leaves: the grown box of the face at sorted position k -> node n-1+k
__global__ void __launch_bounds__(256) k_bvh_leaves(const float* __restrict__ pos, const uint32_t* __restrict__ faces, RayBvh t) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.n) return;
  const uint32_t f = t.order[k];
  const float* A = pos + 3 * (size_t)faces[3 * (size_t)f];
  const float* B = pos + 3 * (size_t)faces[3 * (size_t)f + 1];
  const float* C = pos + 3 * (size_t)faces[3 * (size_t)f + 2];
  float lo[3], hi[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    lo[q] = fminf(fminf(A[q], B[q]), C[q]) - t.eps;
    hi[q] = fmaxf(fmaxf(A[q], B[q]), C[q]) + t.eps;
  }
  t.lo[t.n - 1 + k] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(f));
  t.hi[t.n - 1 + k] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(RAY_NONE));
}

// common-prefix length of the sorted keys at i and j (equal keys: the positions themselves extend the key); -1 out of range
__device__ __forceinline__ int bvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a != b) return __clzll((long long)(a ^ b));
  return 64 + __clz(i ^ j);
}

// Karras 2012: internal node i covers a key range found by doubling + binary search; its split is where the prefix grows
__global__ void __launch_bounds__(256) k_bvh_tree(RayBvh t) {
  const int n = (int)t.n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int dir = bvh_delta(t.keys, n, i, i + 1) > bvh_delta(t.keys, n, i, i - 1) ? 1 : -1;
  const int dmin = bvh_delta(t.keys, n, i, i - dir);
  int lmax = 2;
  while (bvh_delta(t.keys, n, i, i + lmax * dir) > dmin) lmax *= 2;
  int l = 0;
  for (int s = lmax / 2; s >= 1; s /= 2)
    if (bvh_delta(t.keys, n, i, i + (l + s) * dir) > dmin) l += s;
  const int j = i + l * dir;
  const int dnode = bvh_delta(t.keys, n, i, j);
  int s = 0;
  for (int step = (l + 1) / 2;; step = (step + 1) / 2) {
    if (bvh_delta(t.keys, n, i, i + (s + step) * dir) > dnode) s += step;
    if (step == 1) break;
  }
  const int gamma = i + s * dir + min(dir, 0);
  const int lo_ = min(i, j), hi_ = max(i, j);
  const uint32_t left = (lo_ == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
  const uint32_t right = (hi_ == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
  t.lo[i].w = __uint_as_float(left); t.hi[i].w = __uint_as_float(right);
  t.parent[left] = (uint32_t)i; t.parent[right] = (uint32_t)i;
  if (i == 0) t.parent[0] = RAY_NONE;
}

// bottom-up boxes: the second child to arrive at a node merges both boxes and moves on
__global__ void __launch_bounds__(256) k_bvh_refit(RayBvh t) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.n || t.n < 2) return;
  uint32_t node = t.parent[t.n - 1 + k];
  __threadfence();
  while (node != RAY_NONE) {
    if (atomicAdd(&t.visits[node], 1u) == 0u) return;
    __threadfence();
    const uint32_t l = __float_as_uint(__ldcg(&t.lo[node].w)), r = __float_as_uint(__ldcg(&t.hi[node].w));
    const float4 ll = __ldcg(&t.lo[l]), lh = __ldcg(&t.hi[l]), rl = __ldcg(&t.lo[r]), rh = __ldcg(&t.hi[r]);
    t.lo[node] = make_float4(fminf(ll.x, rl.x), fminf(ll.y, rl.y), fminf(ll.z, rl.z), __uint_as_float(l));
    t.hi[node] = make_float4(fmaxf(lh.x, rh.x), fmaxf(lh.y, rh.y), fmaxf(lh.z, rh.z), __uint_as_float(r));
    __threadfence();
    node = t.parent[node];
  }
}

// ---- traversal --------------------------------------------------------------------------------------------------------
struct RayHit { float t; uint32_t face; };

// nearest hit of one ray; faces with the corner skip_vertex are ignored.  *overflow is raised if the stack is too small
// (the caller reports an error: a result is never silently wrong).
__device__ inline RayHit bvh_cast(const RayBvh& t, const float* __restrict__ pos, const uint32_t* __restrict__ faces,
                                  const Ray& r, uint32_t skip_vertex, unsigned int* overflow) {
  RayHit best{__uint_as_float(0x7f800000u), RAY_NONE};
  if (t.n == 0) return best;
  uint32_t stack[RAY_STACK]; float entry[RAY_STACK];
  int sp = 0;
  const uint32_t first_leaf = t.n - 1;
  // a child is examined when its parent is visited: leaves are tested at once, internal nodes are pushed with their entry
  auto examine = [&](uint32_t c, bool& push, float& tn_out) {
    push = false;
    const float4 lo = __ldg(&t.lo[c]), hi = __ldg(&t.hi[c]);
    float tn, tf;
    if (!box_span(r, lo, hi, tn, tf) || tn > best.t) return;
    if (c >= first_leaf) {
      const uint32_t f = __float_as_uint(lo.w);
      const uint32_t a = __ldg(&faces[3 * (size_t)f]), b = __ldg(&faces[3 * (size_t)f + 1]), cc = __ldg(&faces[3 * (size_t)f + 2]);
      if (a == skip_vertex || b == skip_vertex || cc == skip_vertex) return;
      float th;
      if (!ray_triangle(r, pos + 3 * (size_t)a, pos + 3 * (size_t)b, pos + 3 * (size_t)cc, th)) return;
      if (!(th >= tn && th <= tf)) return;
      if (th < best.t || (th == best.t && f < best.face)) { best.t = th; best.face = f; }
    } else { push = true; tn_out = tn; }
  };
  bool push; float tn0 = 0.0f;
  examine(0u, push, tn0);
  if (push) { stack[0] = 0u; entry[0] = tn0; sp = 1; }
  while (sp > 0) {
    --sp;
    const uint32_t node = stack[sp];
    if (entry[sp] > best.t) continue;
    const uint32_t l = __float_as_uint(__ldg(&t.lo[node].w)), rr = __float_as_uint(__ldg(&t.hi[node].w));
    bool pl, pr; float tl = 0.0f, tr = 0.0f;
    examine(l, pl, tl);
    examine(rr, pr, tr);
    if (pl && pr) {
      if (sp + 2 > RAY_STACK) { atomicExch(overflow, 1u); return best; }
      const bool left_first = tl <= tr;                 // the nearer child is popped first
      stack[sp] = left_first ? rr : l; entry[sp] = left_first ? tr : tl; ++sp;
      stack[sp] = left_first ? l : rr; entry[sp] = left_first ? tl : tr; ++sp;
    } else if (pl || pr) {
      if (sp + 1 > RAY_STACK) { atomicExch(overflow, 1u); return best; }
      stack[sp] = pl ? l : rr; entry[sp] = pl ? tl : tr; ++sp;
    }
  }
  return best;
}

// lvr2::RaycasterBase::castRays (obstacle_layer.cpp:239): hit flag, distance, face id and hit point per ray
__global__ void __launch_bounds__(128) k_cast_rays(RayBvh t, const float* __restrict__ pos, const uint32_t* __restrict__ faces,
                                                   uint32_t n, const float* __restrict__ origins, const float* __restrict__ dirs,
                                                   uint32_t dir_stride, uint8_t* __restrict__ hit, float* __restrict__ dist,
                                                   uint32_t* __restrict__ face, float* __restrict__ point, unsigned int* overflow) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.o[k] = origins[3 * (size_t)i + k]; r.d[k] = dirs[(size_t)dir_stride * i + k]; }
  const RayHit h = bvh_cast(t, pos, faces, r, RAY_NONE, overflow);
  const bool any = h.face != RAY_NONE;
  if (hit) hit[i] = any ? 1 : 0;
  if (dist) dist[i] = h.t;
  if (face) face[i] = h.face;
  if (point)
    for (int k = 0; k < 3; ++k) point[3 * (size_t)i + k] = any ? r.o[k] + r.d[k] * h.t : __uint_as_float(0x7fc00000u);
}

struct ObstacleArgs {
  float tf[12];                 // row-major [R|t], message frame -> map frame (obstacle_layer.cpp:176-180)
  float axis[3];                // down axis in the map frame (:183-205)
  double max_obstacle_dist, robot_height;
};

// ObstacleLayer::processPointCloud, the per-point part (obstacle_layer.cpp:215-256): range filter, transform, ray along the
// down axis, hit within robot_height -> the three vertices of the face are marked
__global__ void __launch_bounds__(128) k_obstacle_rays(RayBvh t, const float* __restrict__ pos, const uint32_t* __restrict__ faces,
                                                       uint32_t n, const float* __restrict__ points, ObstacleArgs a,
                                                       uint8_t* __restrict__ mark, unsigned int* overflow) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = points[3 * (size_t)i], y = points[3 * (size_t)i + 1], z = points[3 * (size_t)i + 2];
  const float norm = sqrtf((x * x + y * y) + z * z);
  if (!((double)norm <= a.max_obstacle_dist)) return;                                     // :221 (NaN points are dropped)
  Ray r;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.o[k] = ((a.tf[4 * k] * x + a.tf[4 * k + 1] * y) + a.tf[4 * k + 2] * z) + a.tf[4 * k + 3];
    r.d[k] = a.axis[k];
  }
  const RayHit h = bvh_cast(t, pos, faces, r, RAY_NONE, overflow);
  if (h.face != RAY_NONE && (double)h.t <= a.robot_height)                                // :245
    for (int k = 0; k < 3; ++k) mark[faces[3 * (size_t)h.face + k]] = 1;                  // :248-252
}

// new lethal set against the previous one (obstacle_layer.cpp:258-273): membership flags for the ordered compaction
// (NaN = not a member, the encoding k_update_set_* already understand) and the roll-over of the mask
__global__ void __launch_bounds__(256) k_obstacle_diff(uint32_t V, const uint8_t* __restrict__ now, uint8_t* __restrict__ lethal_mask,
                                                       float* __restrict__ member_now, float* __restrict__ member_changed,
                                                       float* __restrict__ costs) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float nan = __uint_as_float(0x7fc00000u);
  const bool a = now[v] != 0, b = lethal_mask[v] != 0;
  member_now[v] = a ? 0.0f : nan;
  member_changed[v] = (a != b) ? 0.0f : nan;
  lethal_mask[v] = a ? 1 : 0;
  if (costs) costs[v] = a ? __uint_as_float(0x7f800000u) : nan;                            // costs.insert(vertex, inf) (:250)
}

// lvr2::calcNormalClearance (clearance_layer.cpp:161): ray from each vertex along its normal, faces at the vertex ignored
__global__ void __launch_bounds__(128) k_normal_clearance(RayBvh t, const float* __restrict__ pos, const uint32_t* __restrict__ faces,
                                                          uint32_t V, const float* __restrict__ vn, float* __restrict__ out,
                                                          unsigned int* overflow) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  Ray r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.o[k] = pos[3 * (size_t)v + k]; r.d[k] = vn[3 * (size_t)v + k]; }
  const float inf = __uint_as_float(0x7f800000u);
  if (!(((r.d[0] * r.d[0] + r.d[1] * r.d[1]) + r.d[2] * r.d[2]) > 0.0f)) { out[v] = inf; return; }
  const RayHit h = bvh_cast(t, pos, faces, r, v, overflow);
  out[v] = h.face != RAY_NONE ? h.t : inf;
}

}  // namespace mnb
