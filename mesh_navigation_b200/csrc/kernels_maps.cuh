// Small per-edge / per-corner map kernels (edge distances, edge weights, weight gathers, ELL adjacency).
// (part of libmeshnav_b200.so: included by meshnav.cu, which holds the C ABI and all host code)
#pragma once
#include "launch.cuh"
#include "problems.cuh"

using namespace mnb;

// ============================================================================
// small map kernels
// ============================================================================
// lvr2::calcVertexDistances equivalent (mesh_map.cpp:404-425): Euclidean edge length, float.
__global__ void k_edge_dist(const float* __restrict__ pos, const uint32_t* __restrict__ edges, uint32_t E,
                            float* __restrict__ out) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint32_t a = edges[2 * (size_t)e], b = edges[2 * (size_t)e + 1];
  const float dx = pos[3 * (size_t)a] - pos[3 * (size_t)b];
  const float dy = pos[3 * (size_t)a + 1] - pos[3 * (size_t)b + 1];
  const float dz = pos[3 * (size_t)a + 2] - pos[3 * (size_t)b + 2];
  out[e] = sqrtf(dx * dx + dy * dy + dz * dz);
}

// MeshMap::computeEdgeWeights (mesh_map.cpp:517-561)
__global__ void k_edge_weights(const float* __restrict__ cost, const uint32_t* __restrict__ edges,
                               const float* __restrict__ dist, double factor, uint32_t E, float* __restrict__ out) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float c1 = cost[edges[2 * (size_t)e]], c2 = cost[edges[2 * (size_t)e + 1]];
  if (isinf(c1) || isinf(c2)) {
    out[e] = __uint_as_float(INF_BITS);
  } else {
    const float vertex_dist = dist[e];
    const float edge_cost = (float)((double)(vertex_dist * (c1 + c2)) / 2.0);   // :550 (float product, /2.0 in double)
    out[e] = (float)((double)vertex_dist + factor * (double)edge_cost);         // :552
  }
}

// sum and count of the finite edge weights (the scale the band widths follow): out = {sum, count}
__global__ void k_weight_scale(const float* __restrict__ w, uint32_t E, double* __restrict__ out) {
  double s = 0.0, n = 0.0;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const float x = w[e];
    if (x > 0.0f && __float_as_uint(x) < INF_BITS) { s += (double)x; n += 1.0; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], n); }
}

// per-corner weight records {w(v1,v2), w(v1,c), w(v2,c), 0}
__global__ void k_gather_corner_w(const uint4* __restrict__ cor_eid, const float* __restrict__ w, size_t NC,
                                  float4* __restrict__ out) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= NC) return;
  const uint4 e = cor_eid[k];
  out[k] = make_float4(w[e.x], w[e.y], w[e.z], 0.0f);
}

// static half of the CVP unfolding per ELL slot {p, hc, t0a, -} in double (CvpEllProblem::face_geo): depends on the
// installed edge weights only, so it is computed once per mnb_set_costs instead of once per recompute
__global__ void k_corner_geo(const float4* __restrict__ w, size_t N, double4* __restrict__ out) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const float4 ww = w[k];
  const CvpEllProblem::FaceGeo g = CvpEllProblem::face_geo((double)ww.z, (double)ww.y, (double)ww.x);
  out[k] = make_double4(g.p, g.hc, g.t0a, 0.0);
}

// per-directed-edge records {neighbour, weight bits}
__global__ void k_gather_adj_w(const uint32_t* __restrict__ nbr, const uint32_t* __restrict__ eid,
                               const float* __restrict__ w, size_t NA, uint2* __restrict__ out) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= NA) return;
  out[k] = make_uint2(nbr[k], __float_as_uint(w[eid[k]]));
}

// ELL view of the same records for the 8-lanes-per-candidate Dijkstra: row v = 8 x {neighbour | -1, weight bits, -, degree}
__global__ void k_build_ell_adj(const uint32_t* __restrict__ adj_ptr, const uint2* __restrict__ adj_nw, uint32_t V,
                                uint4* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)V * ELL_W) return;
  const uint32_t v = (uint32_t)(t / ELL_W), j = (uint32_t)(t % ELL_W);
  const uint32_t kb = adj_ptr[v], deg = adj_ptr[v + 1] - kb;
  uint4 r = make_uint4(0xffffffffu, INF_BITS, 0u, deg);
  if (j < deg) { const uint2 nw = adj_nw[kb + j]; r.x = nw.x; r.y = nw.y; }
  out[t] = r;
}
