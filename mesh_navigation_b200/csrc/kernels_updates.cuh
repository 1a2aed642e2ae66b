// Incremental updates: layerChanged / updateEdgeWeights on the changed vertices, MaxCombination update, ordered update set.
// (part of libmeshnav_b200.so: included by meshnav.cu, which holds the C ABI and all host code)
#pragma once
#include "launch.cuh"
#include "problems.cuh"

using namespace mnb;

// ============================================================================
// Incremental updates: MeshMap::layerChanged + updateEdgeWeights (mesh_map.cpp:455-492, 563-618),
// MaxCombinationLayer::onInputChanged (combination_layer.cpp:87-147), the update set of InflationLayer::onInputChanged
// (inflation_layer.cpp:154-164).  Work is proportional to the changed set: 8 lanes per changed vertex walk its incident
// edges / faces and patch only the table entries that hold one of those edges' weights.
// ============================================================================
__global__ void k_update_costs(const uint32_t* __restrict__ changed, uint32_t n, const float* __restrict__ costs, int by_vertex,
                               float default_value, uint32_t V, float* __restrict__ cost, uint32_t* __restrict__ stamp, uint32_t call) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = changed[i];
  if (v >= V) return;
  stamp[v] = call;                                            // membership in this call's change set (k_refresh_weight_tables)
  float c = by_vertex ? costs[v] : costs[i];
  if (by_vertex && c != c) c = default_value;                 // cost_map.get(vH).value_or(default_value), mesh_map.cpp:486
  cost[v] = c;
}

__global__ void k_update_edge_weights(const uint32_t* __restrict__ changed, uint32_t n, uint32_t V, const uint32_t* __restrict__ adj_ptr,
                                      const uint32_t* __restrict__ adj_eid, const uint32_t* __restrict__ edges,
                                      const float* __restrict__ cost, const float* __restrict__ dist, double factor,
                                      float* __restrict__ w) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * ELL_W) return;
  const uint32_t v = changed[t / ELL_W], j = (uint32_t)(t % ELL_W);
  if (v >= V) return;
  for (uint32_t k = adj_ptr[v] + j; k < adj_ptr[v + 1]; k += ELL_W) {            // getEdgesOfVertex, mesh_map.cpp:580
    const uint32_t e = adj_eid[k];
    const float c1 = cost[edges[2 * (size_t)e]], c2 = cost[edges[2 * (size_t)e + 1]];
    if (isinf(c1) || isinf(c2)) {                                                // :598
      w[e] = __uint_as_float(INF_BITS);
    } else {
      const float vertex_dist = dist[e];
      const float edge_cost = (float)((double)(vertex_dist * (c1 + c2)) / 2.0);  // :609
      w[e] = (float)((double)vertex_dist + factor * (double)edge_cost);          // :611
    }
  }
}

struct RefreshArgs {
  const uint32_t* changed; uint32_t n, V;
  const uint32_t* faces;
  const uint32_t* cor_ptr; const int4* cor_idx; const uint4* cor_eid; const uint32_t* face_cor;
  const uint32_t* adj_ptr; const uint32_t* adj_nbr; const uint32_t* adj_eid;
  const float* w;
  float4* cor_w; float4* ell_w; double4* ell_geo; uint2* adj_nw; uint4* ell_adj;
  const uint32_t* stamp; uint32_t call;      // stamp[x] == call: x is in the change set of this call
};
// every table entry that stores the weight of an edge incident to a changed vertex: the corner records (CSR + ELL +
// precomputed unfolding geometry) of all three vertices of each incident face, and both directions of the adjacency.
// A face (an edge) whose vertices are all in the change set is patched once, by its smallest changed vertex: obstacle
// discs change compact regions, where every face used to be rewritten three times.
__global__ void k_refresh_weight_tables(const RefreshArgs a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)a.n * ELL_W) return;
  const uint32_t v = a.changed[t / ELL_W], j = (uint32_t)(t % ELL_W);
  if (v >= a.V) return;
  for (uint32_t k = a.cor_ptr[v] + j; k < a.cor_ptr[v + 1]; k += ELL_W) {
    const int4 ci = a.cor_idx[k];
    const uint32_t f = (uint32_t)ci.z;
    if (((uint32_t)ci.x < v && a.stamp[(uint32_t)ci.x] == a.call) || ((uint32_t)ci.y < v && a.stamp[(uint32_t)ci.y] == a.call)) continue;
    for (int c = 0; c < 3; ++c) {                      // the three corner records of the face, through the face -> corner table
      const uint32_t x = a.faces[3 * (size_t)f + c];
      const uint32_t kk = a.face_cor[3 * (size_t)f + c], kb = a.cor_ptr[x];
      const uint4 e = a.cor_eid[kk];
      const float4 ww = make_float4(a.w[e.x], a.w[e.y], a.w[e.z], 0.0f);
      a.cor_w[kk] = ww;
      if (kk - kb < ELL_W) {
        const size_t s = (size_t)x * ELL_W + (kk - kb);
        a.ell_w[s] = ww;
        const CvpEllProblem::FaceGeo g = CvpEllProblem::face_geo((double)ww.z, (double)ww.y, (double)ww.x);
        a.ell_geo[s] = make_double4(g.p, g.hc, g.t0a, 0.0);
      }
    }
  }
  if (!a.adj_nw) return;                       // the adjacency tables are stale as a whole: the next Dijkstra call rebuilds them
  const uint32_t ab = a.adj_ptr[v];
  for (uint32_t k = ab + j; k < a.adj_ptr[v + 1]; k += ELL_W) {
    const uint32_t u = a.adj_nbr[k];
    if (u < v && a.stamp[u] == a.call) continue;
    const uint32_t wb = __float_as_uint(a.w[a.adj_eid[k]]);
    a.adj_nw[k] = make_uint2(u, wb);
    if (k - ab < ELL_W) reinterpret_cast<uint32_t*>(&a.ell_adj[(size_t)v * ELL_W + (k - ab)])[1] = wb;
    const uint32_t ub = a.adj_ptr[u], ue = a.adj_ptr[u + 1];
    for (uint32_t kk = ub; kk < ue; ++kk) {
      if (a.adj_nbr[kk] != v) continue;
      a.adj_nw[kk] = make_uint2(v, wb);
      if (kk - ub < ELL_W) reinterpret_cast<uint32_t*>(&a.ell_adj[(size_t)u * ELL_W + (kk - ub)])[1] = wb;
      break;
    }
  }
}

constexpr int COMB_MAX_LAYERS = 8;
struct CombineArgs {
  const float* costs[COMB_MAX_LAYERS]; const uint8_t* lethal[COMB_MAX_LAYERS]; float def[COMB_MAX_LAYERS];
  float weight[COMB_MAX_LAYERS]; int average;      // AvgCombinationLayer: weighted sum in layer order (combination_layer.cpp:264-271)
  uint32_t n_layers;
  const uint32_t* changed; uint32_t n, V;
  float* io_costs; uint8_t* io_lethal;
};
__global__ void k_max_combination_update(const CombineArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const uint32_t v = a.changed[i];
  if (v >= a.V) return;
  float cost = 0.0f; bool lethal = false;
  for (uint32_t l = 0; l < a.n_layers; ++l) {
    float tmp = a.costs[l][v];
    if (tmp != tmp) tmp = a.def[l];                        // cm.get(v).value_or(def), combination_layer.cpp:116
    if (a.average) cost = cost + a.weight[l] * tmp;        // cost += combinationWeight * value (:269), float, no contraction
    else cost = fmaxf(tmp, cost);                          // std::max(tmp, cost) (:117)
    lethal = lethal || (a.lethal[l] && a.lethal[l][v]);
  }
  a.io_costs[v] = cost;
  if (a.io_lethal) a.io_lethal[v] = lethal ? 1 : 0;
}

// update set of InflationLayer::onInputChanged: keys(new riskiness) U keys(old riskiness), ascending.
// Ordered compaction in three small kernels: per-tile counts, one-CTA exclusive scan of the tile counts, ordered write.
constexpr int US_TILE = 2048;    // vertices per CTA (256 threads x 8)
__device__ __forceinline__ bool in_update_set(const float* __restrict__ nw, const float* __restrict__ old, uint32_t v) {
  const float a = nw[v];
  if (a == a) return true;
  if (old) { const float b = old[v]; return b == b; }
  return false;
}
__global__ void __launch_bounds__(256) k_update_set_count(const float* __restrict__ nw, const float* __restrict__ old, uint32_t V,
                                                          unsigned int* __restrict__ tile_count) {
  __shared__ unsigned int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  unsigned int mine = 0;
  const uint32_t base = blockIdx.x * (uint32_t)US_TILE;
  for (int r = 0; r < US_TILE / 256; ++r) {
    const uint32_t v = base + r * 256 + threadIdx.x;
    if (v < V && in_update_set(nw, old, v)) mine++;
  }
  const unsigned int wsum = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&cnt, wsum);
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = cnt;
}
__global__ void __launch_bounds__(1024) k_update_set_scan(unsigned int* __restrict__ tile_count, uint32_t n_tiles, unsigned int* __restrict__ total) {
  __shared__ unsigned int warp_sum[32];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_tiles; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const unsigned int x = i < n_tiles ? tile_count[i] : 0u;
    unsigned int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o); if ((int)(threadIdx.x & 31) >= o) incl += y; }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      const unsigned int ws = threadIdx.x < (blockDim.x >> 5) ? warp_sum[threadIdx.x] : 0u;
      unsigned int wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, wi, o); if ((int)threadIdx.x >= o) wi += y; }
      warp_sum[threadIdx.x] = wi - ws;                      // exclusive prefix of the warp sums
    }
    __syncthreads();
    const unsigned int excl = carry + warp_sum[threadIdx.x >> 5] + incl - x;
    if (i < n_tiles) tile_count[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_update_set_write(const float* __restrict__ nw, const float* __restrict__ old, uint32_t V,
                                                          const unsigned int* __restrict__ tile_offset, uint32_t* __restrict__ out) {
  __shared__ unsigned int warp_base[8];
  __shared__ unsigned int run;
  if (threadIdx.x == 0) run = tile_offset[blockIdx.x];
  __syncthreads();
  const uint32_t base = blockIdx.x * (uint32_t)US_TILE;
  for (int r = 0; r < US_TILE / 256; ++r) {
    const uint32_t v = base + r * 256 + threadIdx.x;
    const bool in = v < V && in_update_set(nw, old, v);
    const unsigned int bal = __ballot_sync(0xffffffffu, in);
    const unsigned int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_base[wid] = __popc(bal);
    __syncthreads();
    unsigned int before = 0;
    for (unsigned int q = 0; q < wid; ++q) before += warp_base[q];
    unsigned int row_total = 0;
    for (unsigned int q = 0; q < 8; ++q) row_total += warp_base[q];
    if (in) out[run + before + __popc(bal & ((1u << lane) - 1u))] = v;
    __syncthreads();
    if (threadIdx.x == 0) run += row_total;
    __syncthreads();
  }
}
