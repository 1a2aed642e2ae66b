// libmeshnav_b200.so -- C ABI (include/meshnav_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17
//             -Xcompiler -fPIC -shared -o libmeshnav_b200.so meshnav.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/meshnav_b200.h"
#include "band_engine.cuh"
#include "problems.cuh"
#include "topology.hpp"

using namespace mnb;

// Kernel launches go through one macro: the CPU interpreter behind the `-m "not gpu"` logic tests (tests/emu/) compiles
// this very file with g++, which has no <<<>>>.  MNB_EMU_ACTIVE is only ever defined by tests/emu/cuda_runtime.h; the
// shipped library is built by nvcc without it and contains no host execution path for any kernel.
#ifdef MNB_EMU_ACTIVE
#define MNB_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(kern, (unsigned)(grid), (unsigned)(block), (size_t)(smem), 1u, false, __VA_ARGS__)
#define MNB_DYNAMIC_SMEM(name) unsigned char* name = emu::g_cta.dyn_smem
#else
#define MNB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MNB_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

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

// ============================================================================
// wavefront kernels
// ============================================================================
// round watchdog (group-uniform): far above the dependency depth of any sane mesh (a grid needs ~ D/h rounds),
// small enough that a livelock is reported in seconds instead of hanging the device
static inline uint32_t watchdog_rounds(uint32_t V) { return 200000u + 256u * (uint32_t)sqrt((double)V); }

struct WaveWorkspace {     // per group (index g): state + g*V etc.
  uint4* state;
  uint32_t* minor;
  uint32_t* root;    // cascade roots of flagged labels (problems.cuh)
  uint32_t* last_eval; uint32_t* dirty; uint32_t* excl;   // clean-candidate skip stamps (problems.cuh)
  uint32_t* chg;
  uint32_t* ver;     // single-plan only (V entries): input versions for the in-round sweeps
  uint32_t* mark;
  uint32_t* list0;
  uint32_t* list1;
  GroupCtl* ctl;
};

struct CvpKernelArgs {
  uint32_t V;
  const float* pos;
  const uint32_t* faces;
  const uint32_t* cor_ptr; const int4* cor_idx; const float4* cor_w;
  const int4* ell_idx; const float4* ell_w; const double4* ell_geo;
  const float* cost; const uint8_t* invalid;
  WaveWorkspace ws;
  uint32_t n_queries;
  const uint32_t* seed_faces;   // [n_queries] device
  const float* seed_pos;        // [3 n_queries] device
  long long robot_face;         // single query only, -1 = none
  double cost_limit, goal_dist_offset;
  float delta;
  float* out_dist;              // [n_queries][V]
  uint32_t* out_pred;           // single query or null
  float* out_dir;
  int32_t* out_cut;
  unsigned int* next_query;
  const int* cancel_flag;
  uint32_t max_rounds;
  int sweeps;                   // in-round sweeps of a single plan (0 = off)
  int skip_clean;               // clean-candidate skip (band_engine.cuh), 0 = off
};

template <int CS>
__device__ __forceinline__ void group_coords(uint32_t& g, uint32_t& gthreads, uint32_t& gtid) {
  if constexpr (CS == 0) {
    g = 0; gthreads = gridDim.x * blockDim.x; gtid = blockIdx.x * blockDim.x + threadIdx.x;
  } else {
    g = blockIdx.x / CS; gthreads = CS * blockDim.x; gtid = (blockIdx.x % CS) * blockDim.x + threadIdx.x;
  }
}

__device__ __forceinline__ void ctl_reset(GroupCtl* ctl, unsigned int n0, float seed_min) {
  ctl->count[0] = n0; ctl->count[1] = 0; ctl->count[2] = 0;
  ctl->m_tau[0] = INF_BITS; ctl->m_tau[1] = INF_BITS; ctl->m_tau[2] = 0u;
  ctl->lo[0] = INF_BITS; ctl->lo[1] = INF_BITS; ctl->lo[2] = __float_as_uint(seed_min);
  ctl->goal_ring[0] = INF_BITS; ctl->goal_ring[1] = INF_BITS; ctl->stop_ring[0] = 0; ctl->stop_ring[1] = 0;
  ctl->goal_bits = INF_BITS; ctl->robot_left = 0;
}

#ifndef MNB_CVP_MINBLOCKS
#define MNB_CVP_MINBLOCKS 1
#endif
#ifndef MNB_CVP_THREADS
#define MNB_CVP_THREADS 512
#endif
template <int CS, bool SKIP>
__global__ void __launch_bounds__(MNB_CVP_THREADS, MNB_CVP_MINBLOCKS) k_cvp(const CvpKernelArgs a) {
  __shared__ Stage st;
  uint32_t g, gthreads, gtid;
  group_coords<CS>(g, gthreads, gtid);
  const uint32_t V = a.V;
  uint4* state = a.ws.state + (size_t)g * V;
  uint32_t* mark = a.ws.mark + (size_t)g * V;
  uint32_t* list0 = a.ws.list0 + (size_t)g * V;
  uint32_t* list1 = a.ws.list1 + (size_t)g * V;
  GroupCtl* ctl = a.ws.ctl + g;
  if (threadIdx.x == 0) { st.n = 0; st.m_tau = INF_BITS; st.lo = INF_BITS; }
  __syncthreads();

  for (;;) {
    if (gtid == 0) ctl->query = atomicAdd(a.next_query, 1u);
    group_sync<CS>();
    const uint32_t q = __ldcg(&ctl->query);
    if (q >= a.n_queries) break;
    const bool single = (a.n_queries == 1);
    uint32_t* chg = a.ws.chg + (size_t)g * V;
    const int sweeps = 0;                            // in-round sweeps are compiled into the whole-grid kernel only
    uint32_t* last_eval = a.ws.last_eval + (size_t)g * V; uint32_t* dirty = a.ws.dirty + (size_t)g * V;
    for (uint32_t v = gtid; v < V; v += gthreads) { state[v] = state_inf(); mark[v] = MARK_NONE; chg[v] = 0u; last_eval[v] = 0u; dirty[v] = 0u; if (sweeps) a.ws.ver[v] = 0u; }
    group_sync<CS>();

    const uint32_t sf = a.seed_faces[q];
    const uint32_t s0 = a.faces[3 * (size_t)sf], s1 = a.faces[3 * (size_t)sf + 1], s2 = a.faces[3 * (size_t)sf + 2];
    CvpEllProblemT<SKIP> prob;
    prob.cor_ptr = a.cor_ptr; prob.cor_idx = a.cor_idx; prob.cor_w = a.cor_w; prob.cost = a.cost; prob.invalid = a.invalid;
    prob.ell_idx = a.ell_idx; prob.ell_w = a.ell_w; prob.ell_geo = a.ell_geo;
    prob.state = state; prob.minor_arr = a.ws.minor + (size_t)g * V; prob.root_arr = a.ws.root + (size_t)g * V; prob.chg = chg; prob.ver = a.ws.ver; prob.deferred_m = __uint_as_float(INF_BITS); prob.pred = nullptr; prob.dir = nullptr; prob.cut = nullptr; prob.cost_limit = a.cost_limit;
    prob.s0 = s0; prob.s1 = s1; prob.s2 = s2; prob.seed_noexpand = 0;
    prob.last_eval = last_eval; prob.dirty_round = dirty; prob.excl_min = a.ws.excl + (size_t)g * V; prob.skip_clean = SKIP ? 1 : 0;
    float sd[3];
    {
      const uint32_t sv[3] = {s0, s1, s2};
      for (int k = 0; k < 3; ++k) {   // cvp:719-728
        const float dx = a.seed_pos[3 * (size_t)q] - a.pos[3 * (size_t)sv[k]];
        const float dy = a.seed_pos[3 * (size_t)q + 1] - a.pos[3 * (size_t)sv[k] + 1];
        const float dz = a.seed_pos[3 * (size_t)q + 2] - a.pos[3 * (size_t)sv[k] + 2];
        sd[k] = sqrtf(dx * dx + dy * dy + dz * dz);
        const bool noexp = ((double)a.cost[sv[k]] >= a.cost_limit) || (a.invalid && a.invalid[sv[k]]);  // cvp:757,760
        if (noexp) prob.seed_noexpand |= (1u << k);
      }
    }
    const float seed_min = fminf(sd[0], fminf(sd[1], sd[2]));
    const float seed_max = fmaxf(sd[0], fmaxf(sd[1], sd[2]));
    uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu, r2 = 0xffffffffu;
    const int has_robot = single && a.robot_face >= 0;
    if (has_robot) {
      r0 = a.faces[3 * (size_t)a.robot_face]; r1 = a.faces[3 * (size_t)a.robot_face + 1]; r2 = a.faces[3 * (size_t)a.robot_face + 2];
    }
    if (gtid == 0) {
      const uint32_t sv[3] = {s0, s1, s2};
      for (int k = 0; k < 3; ++k) {
        state[sv[k]] = make_uint4(__float_as_uint(sd[k]), __float_as_uint(sd[k]), 0u, 0u);
        mark[sv[k]] = MARK_FIXED;
      }
      unsigned int n0 = 0;
      for (int k = 0; k < 3; ++k)
        prob.activate(sv[k], [&](uint32_t x) {
          if (mark[x] == MARK_NONE && prob.eligible(x)) { mark[x] = MARK_CAND; list0[n0++] = x; }
        });
      ctl_reset(ctl, n0, seed_min);
      if (has_robot) {
        int left = 0; const uint32_t rv[3] = {r0, r1, r2};
        for (int k = 0; k < 3; ++k) if (mark[rv[k]] != MARK_FIXED) left++;
        ctl->robot_left = left;
        if (left == 0) {  // robot face == seed face: cutoff armed when the last seed pops (cvp:763-771)
          ctl->goal_ring[0] = __float_as_uint((float)((double)seed_max + a.goal_dist_offset));
        }
      }
    }
    group_sync<CS>();
    float delta = a.delta;
    if (has_robot && a.goal_dist_offset < (double)delta) delta = (float)fmax(a.goal_dist_offset, 1e-4);
    run_band_rounds_sub8<CS, false>(prob, ctl, list0, list1, mark, st, delta, gthreads, gtid, has_robot, r0, r1, r2,
                        a.goal_dist_offset, a.cancel_flag, nextafterf(seed_max, __uint_as_float(INF_BITS)), a.max_rounds, sweeps, nullptr, V);
    group_sync<CS>();
    if (a.out_dist) {
      float* od = a.out_dist + (size_t)q * V;
      for (uint32_t v = gtid; v < V; v += gthreads) od[v] = __uint_as_float(state[v].x);
    }
    group_sync<CS>();
  }
}

// Single plan on the whole GPU: cooperative launch, one CTA per SM (x occupancy), 8 lanes per
// candidate, grid-wide barrier per round.  Used when latency of ONE wavefront matters.
#ifndef MNB_GRID_MINBLOCKS
#define MNB_GRID_MINBLOCKS 1
#endif
template <bool SKIP>
__global__ void __launch_bounds__(512, MNB_GRID_MINBLOCKS) k_cvp_grid(const CvpKernelArgs a) {
  __shared__ Stage st;
  __shared__ SweepStage sws;
  uint32_t g, gthreads, gtid;
  group_coords<0>(g, gthreads, gtid);
  const uint32_t V = a.V;
  uint4* state = a.ws.state; uint32_t* mark = a.ws.mark; uint32_t* list0 = a.ws.list0; uint32_t* list1 = a.ws.list1;
  GroupCtl* ctl = a.ws.ctl;
  if (threadIdx.x == 0) { st.n = 0; st.m_tau = INF_BITS; st.lo = INF_BITS; sws.dn[0] = 0; sws.dn[1] = 0; }
  __syncthreads();
  for (uint32_t v = gtid; v < V; v += gthreads) { state[v] = state_inf(); mark[v] = MARK_NONE; a.ws.chg[v] = 0u; a.ws.ver[v] = 0u; a.ws.last_eval[v] = 0u; a.ws.dirty[v] = 0u; }
  group_sync<0>(ctl->barrier);
  const uint32_t sf = a.seed_faces[0];
  const uint32_t s0 = a.faces[3 * (size_t)sf], s1 = a.faces[3 * (size_t)sf + 1], s2 = a.faces[3 * (size_t)sf + 2];
  CvpEllProblemT<SKIP> prob;
  prob.cor_ptr = a.cor_ptr; prob.cor_idx = a.cor_idx; prob.cor_w = a.cor_w; prob.cost = a.cost; prob.invalid = a.invalid;
  prob.ell_idx = a.ell_idx; prob.ell_w = a.ell_w; prob.ell_geo = a.ell_geo;
  prob.state = state; prob.minor_arr = a.ws.minor; prob.root_arr = a.ws.root; prob.chg = a.ws.chg; prob.ver = a.ws.ver; prob.deferred_m = __uint_as_float(INF_BITS); prob.pred = nullptr; prob.dir = nullptr; prob.cut = nullptr; prob.cost_limit = a.cost_limit;
  prob.s0 = s0; prob.s1 = s1; prob.s2 = s2; prob.seed_noexpand = 0;
  prob.last_eval = a.ws.last_eval; prob.dirty_round = a.ws.dirty; prob.excl_min = a.ws.excl; prob.skip_clean = SKIP ? 1 : 0;
  float sd[3];
  {
    const uint32_t sv[3] = {s0, s1, s2};
    for (int k = 0; k < 3; ++k) {   // cvp:719-728
      const float dx = a.seed_pos[0] - a.pos[3 * (size_t)sv[k]];
      const float dy = a.seed_pos[1] - a.pos[3 * (size_t)sv[k] + 1];
      const float dz = a.seed_pos[2] - a.pos[3 * (size_t)sv[k] + 2];
      sd[k] = sqrtf(dx * dx + dy * dy + dz * dz);
      const bool noexp = ((double)a.cost[sv[k]] >= a.cost_limit) || (a.invalid && a.invalid[sv[k]]);
      if (noexp) prob.seed_noexpand |= (1u << k);
    }
  }
  const float seed_min = fminf(sd[0], fminf(sd[1], sd[2]));
  const float seed_max = fmaxf(sd[0], fmaxf(sd[1], sd[2]));
  uint32_t r0 = 0xffffffffu, r1 = 0xffffffffu, r2 = 0xffffffffu;
  const int has_robot = a.robot_face >= 0;
  if (has_robot) {
    r0 = a.faces[3 * (size_t)a.robot_face]; r1 = a.faces[3 * (size_t)a.robot_face + 1]; r2 = a.faces[3 * (size_t)a.robot_face + 2];
  }
  if (gtid == 0) {
    const uint32_t sv[3] = {s0, s1, s2};
    for (int k = 0; k < 3; ++k) { state[sv[k]] = make_uint4(__float_as_uint(sd[k]), __float_as_uint(sd[k]), 0u, 0u); mark[sv[k]] = MARK_FIXED; }
    unsigned int n0 = 0;
    for (int k = 0; k < 3; ++k)
      prob.activate(sv[k], [&](uint32_t x) {
        if (mark[x] == MARK_NONE && prob.eligible(x)) { mark[x] = MARK_CAND; list0[n0++] = x; }
      });
    ctl_reset(ctl, n0, seed_min);
    if (has_robot) {
      int left = 0; const uint32_t rv[3] = {r0, r1, r2};
      for (int k = 0; k < 3; ++k) if (mark[rv[k]] != MARK_FIXED) left++;
      ctl->robot_left = left;
      if (left == 0) ctl->goal_ring[0] = __float_as_uint((float)((double)seed_max + a.goal_dist_offset));
    }
  }
  group_sync<0>(ctl->barrier);
  const float delta = a.delta;      // not clamped to goal_dist_offset: the engine caps settling instead (settle_cap)
  // in-round sweeps pay off once the band is several dependency hops deep (one hop ~ 0.15 m of potential on these meshes)
  int sweeps = a.sweeps;
  if (sweeps < 0) sweeps = delta < 0.45f ? 0 : min(15, (int)(delta / 0.16f));
  run_band_rounds_sub8<0, true>(prob, ctl, list0, list1, mark, st, delta, gthreads, gtid, has_robot, r0, r1, r2,
                          a.goal_dist_offset, a.cancel_flag, nextafterf(seed_max, __uint_as_float(INF_BITS)), a.max_rounds, sweeps, &sws, V);
  group_sync<0>(ctl->barrier);
  if (a.out_dist)
    for (uint32_t v = gtid; v < V; v += gthreads) a.out_dist[v] = __uint_as_float(state[v].x);
}

// predecessors_ / direction_ / cutting_faces_ (cvp:423-431,493-517) from the FINAL labels: every vertex
// replays its faces once more in event order and evaluates the winning face with the literal acos form.
// Done after the wavefront so that the stored angles use the final source potentials.
__global__ void __launch_bounds__(256) k_cvp_epilogue(const CvpKernelArgs a, const GroupCtl* ctl) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.V) return;
  const uint32_t sf = a.seed_faces[0];
  CvpProblem prob;
  prob.cor_ptr = a.cor_ptr; prob.cor_idx = a.cor_idx; prob.cor_w = a.cor_w; prob.cost = a.cost; prob.invalid = a.invalid;
  prob.state = a.ws.state; prob.minor_arr = a.ws.minor; prob.root_arr = a.ws.root; prob.chg = a.ws.chg; prob.ver = nullptr; prob.deferred_m = __uint_as_float(INF_BITS); prob.pred = a.out_pred; prob.dir = a.out_dir; prob.cut = a.out_cut; prob.cost_limit = a.cost_limit;
  prob.s0 = a.faces[3 * (size_t)sf]; prob.s1 = a.faces[3 * (size_t)sf + 1]; prob.s2 = a.faces[3 * (size_t)sf + 2];
  prob.seed_noexpand = 0;
  {
    const uint32_t sv[3] = {prob.s0, prob.s1, prob.s2};
    for (int k = 0; k < 3; ++k)
      if (((double)a.cost[sv[k]] >= a.cost_limit) || (a.invalid && a.invalid[sv[k]])) prob.seed_noexpand |= (1u << k);
  }
  const float d = __uint_as_float(a.ws.state[c].x);
  if (prob.seed_index(c) >= 0) {                         // cvp:719-728
    a.out_pred[c] = c; a.out_dir[c] = 0.0f; a.out_cut[c] = (int32_t)sf;
    return;
  }
  int win = -1; float nd, wu1 = 0, wu2 = 0; EvTime nt;
  if (__float_as_uint(d) != INF_BITS && prob.eligible(c))
    prob.replay(c, __uint_as_float(INF_BITS), __uint_as_float(ctl->goal_bits), 0xfffffff0u /* final labels: nothing is deferred */, nd, nt, win, wu1, wu2);
  prob.write_aux(c, win, wu1, wu2);
}

struct DijkstraKernelArgs {
  uint32_t V;
  const uint32_t* adj_ptr; const uint2* adj_nw;
  const float* cost; const uint8_t* invalid;
  WaveWorkspace ws;
  uint32_t seed_vertex; long long robot_vertex;
  double cost_limit, goal_dist_offset;
  float delta;
  float* out_dist; uint32_t* out_pred;
  const int* cancel_flag;
  uint32_t max_rounds;
  const uint4* ell_adj; int sweeps;   // whole-grid kernel only
};

template <int CS>
__global__ void __launch_bounds__(512, 1) k_dijkstra(const DijkstraKernelArgs a) {
  __shared__ Stage st;
  uint32_t g, gthreads, gtid;
  group_coords<CS>(g, gthreads, gtid);
  const uint32_t V = a.V;
  uint4* state = a.ws.state;
  uint32_t* mark = a.ws.mark; uint32_t* list0 = a.ws.list0; uint32_t* list1 = a.ws.list1;
  GroupCtl* ctl = a.ws.ctl;
  if (threadIdx.x == 0) { st.n = 0; st.m_tau = INF_BITS; st.lo = INF_BITS; }
  __syncthreads();
  for (uint32_t v = gtid; v < V; v += gthreads) { state[v] = state_inf(); mark[v] = MARK_NONE; a.out_pred[v] = v; }
  group_sync<CS>(ctl->barrier);
  DijkstraProblem prob;
  prob.adj_ptr = a.adj_ptr; prob.adj_nw = a.adj_nw; prob.cost = a.cost; prob.invalid = a.invalid;
  prob.state = state; prob.pred = a.out_pred; prob.cost_limit = a.cost_limit; prob.deferred_m = __uint_as_float(INF_BITS);
  const int has_robot = a.robot_vertex >= 0;
  const uint32_t rv = has_robot ? (uint32_t)a.robot_vertex : 0xffffffffu;
  if (gtid == 0) {
    state[a.seed_vertex] = make_uint4(0u, 0u, 0u, 0u);     // dijkstra:276 (d = 0, tau = 0)
    mark[a.seed_vertex] = MARK_FIXED;
    unsigned int n0 = 0;
    prob.activate(a.seed_vertex, [&](uint32_t x) {
      if (mark[x] == MARK_NONE && prob.eligible(x)) { mark[x] = MARK_CAND; list0[n0++] = x; }
    });
    ctl_reset(ctl, n0, 0.0f);
    if (has_robot) ctl->robot_left = 1;
  }
  group_sync<CS>();
  float delta = a.delta;
  if (has_robot && a.goal_dist_offset < (double)delta) delta = (float)fmax(a.goal_dist_offset, 1e-4);
  run_band_rounds<CS>(prob, ctl, list0, list1, mark, st, delta, gthreads, gtid, has_robot, rv, rv, rv,
                      a.goal_dist_offset, a.cancel_flag, 1e-30f, a.max_rounds);
  group_sync<CS>();
  for (uint32_t v = gtid; v < V; v += gthreads) a.out_dist[v] = __uint_as_float(state[v].x);
}

// Single Dijkstra plan on the whole GPU: 8 lanes per candidate (one edge each), wide band + in-round sweeps,
// same engine instance as k_cvp_grid.
__global__ void __launch_bounds__(512, 1) k_dijkstra_grid(const DijkstraKernelArgs a) {
  __shared__ Stage st;
  __shared__ SweepStage sws;
  uint32_t g, gthreads, gtid;
  group_coords<0>(g, gthreads, gtid);
  const uint32_t V = a.V;
  uint4* state = a.ws.state;
  uint32_t* mark = a.ws.mark; uint32_t* list0 = a.ws.list0; uint32_t* list1 = a.ws.list1;
  GroupCtl* ctl = a.ws.ctl;
  if (threadIdx.x == 0) { st.n = 0; st.m_tau = INF_BITS; st.lo = INF_BITS; sws.dn[0] = 0; sws.dn[1] = 0; }
  __syncthreads();
  for (uint32_t v = gtid; v < V; v += gthreads) { state[v] = state_inf(); mark[v] = MARK_NONE; a.out_pred[v] = v; a.ws.ver[v] = 0u; }
  group_sync<0>(ctl->barrier);
  DijkstraEllProblem prob;
  prob.adj_ptr = a.adj_ptr; prob.adj_nw = a.adj_nw; prob.cost = a.cost; prob.invalid = a.invalid;
  prob.state = state; prob.pred = a.out_pred; prob.cost_limit = a.cost_limit; prob.deferred_m = __uint_as_float(INF_BITS);
  prob.strict = 0; prob.ell_adj = a.ell_adj; prob.ver = a.ws.ver;
  const int has_robot = a.robot_vertex >= 0;
  const uint32_t rv = has_robot ? (uint32_t)a.robot_vertex : 0xffffffffu;
  if (gtid == 0) {
    state[a.seed_vertex] = make_uint4(0u, 0u, 0u, 0u);     // dijkstra:276 (d = 0, tau = 0)
    mark[a.seed_vertex] = MARK_FIXED;
    unsigned int n0 = 0;
    prob.activate(a.seed_vertex, [&](uint32_t x) {
      if (mark[x] == MARK_NONE && prob.eligible(x)) { mark[x] = MARK_CAND; list0[n0++] = x; }
    });
    ctl_reset(ctl, n0, 0.0f);
    if (has_robot) ctl->robot_left = 1;
  }
  group_sync<0>(ctl->barrier);
  const float delta = a.delta;      // not clamped to goal_dist_offset: the engine caps settling instead (settle_cap)
  int sweeps = a.sweeps;
  if (sweeps < 0) sweeps = delta < 0.45f ? 0 : min(15, (int)(delta / 0.16f));
  run_band_rounds_sub8<0, true>(prob, ctl, list0, list1, mark, st, delta, gthreads, gtid, has_robot, rv, rv, rv,
                                a.goal_dist_offset, a.cancel_flag, 1e-30f, a.max_rounds, sweeps, &sws, V);
  group_sync<0>(ctl->barrier);
  for (uint32_t v = gtid; v < V; v += gthreads) a.out_dist[v] = __uint_as_float(state[v].x);
}

// ============================================================================
// Fused geometric cost layers (mesh_layers: HeightDiff, Roughness, Steepness, Ridge, Clearance cost
// mapping, Border) + MaxCombinationLayer + lethal masks: ONE pass over the radius neighbourhood per
// vertex instead of the reference's three independent visitLocalVertexNeighborhood runs with
// std::set bookkeeping (ridge_layer.cpp:166-175, height_diff_layer.cpp:108, roughness_layer.cpp:143).
// Definitions of the lvr2 pieces: see oracle/oracle.cpp (orc_layers).
// ============================================================================
__global__ void k_face_normals(const float* __restrict__ pos, const uint32_t* __restrict__ faces, uint32_t F,
                               float* __restrict__ fn) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float* p0 = pos + 3 * (size_t)faces[3 * (size_t)f];
  const float* p1 = pos + 3 * (size_t)faces[3 * (size_t)f + 1];
  const float* p2 = pos + 3 * (size_t)faces[3 * (size_t)f + 2];
  const float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
  const float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
  float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
  const float l = sqrtf(nx * nx + ny * ny + nz * nz);
  if (l > 0) { nx /= l; ny /= l; nz /= l; }
  fn[3 * (size_t)f] = nx; fn[3 * (size_t)f + 1] = ny; fn[3 * (size_t)f + 2] = nz;
}

__global__ void k_vertex_normals(const uint32_t* __restrict__ cor_ptr, const int4* __restrict__ cor_idx,
                                 const float* __restrict__ fn, uint32_t V, float* __restrict__ vn) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float nx = 0, ny = 0, nz = 0;
  for (uint32_t k = cor_ptr[v]; k < cor_ptr[v + 1]; ++k) {
    const int f = cor_idx[k].z;
    nx = nx + fn[3 * (size_t)f]; ny = ny + fn[3 * (size_t)f + 1]; nz = nz + fn[3 * (size_t)f + 2];
  }
  const float l = sqrtf(nx * nx + ny * ny + nz * nz);
  if (l > 0) { nx /= l; ny /= l; nz /= l; }
  vn[3 * (size_t)v] = nx; vn[3 * (size_t)v + 1] = ny; vn[3 * (size_t)v + 2] = nz;
}

struct LayerKernelArgs {
  uint32_t V;
  const float* pos; const float* vn;
  const uint32_t* adj_ptr; const uint32_t* adj_nbr;
  const uint8_t* border;
  const float* clearance;      // may be null
  mnb_layer_params P;
  float* costs;                // 6 x V
  float* combined; uint8_t* lethal_mask;
  unsigned int* overflow;      // neighbourhood larger than the per-thread scratch
};

constexpr int NB_SEEN = 320, NB_STACK = 160;

// (fallback for neighbourhoods that overflow the hashed set below: linear `seen` list, any size up to NB_SEEN)
// traversal shared by the three radius layers; WHICH selects the accumulators that are active (bit0 height
// diff, bit1 roughness, bit2 ridge) so that layers with equal radii share one walk
template <int WHICH>
__device__ __noinline__ void walk_linear(const LayerKernelArgs& a, uint32_t v, float radius, float& zmin, float& zmax,
                                     float& rsum, int& rcnt, float& value, int& num) {
  uint32_t seen[NB_SEEN]; uint32_t stack[NB_STACK];
  int ns = 0, sp = 0;
  seen[ns++] = v; stack[sp++] = v;
  const float px = a.pos[3 * (size_t)v], py = a.pos[3 * (size_t)v + 1], pz = a.pos[3 * (size_t)v + 2];
  const float nvx = a.vn[3 * (size_t)v], nvy = a.vn[3 * (size_t)v + 1], nvz = a.vn[3 * (size_t)v + 2];
  const float rx = px + nvx, ry = py + nvy, rz = pz + nvz;
  while (sp > 0) {
    const uint32_t u = stack[--sp];
    for (uint32_t k = a.adj_ptr[u]; k < a.adj_ptr[u + 1]; ++k) {
      const uint32_t n = a.adj_nbr[k];
      bool was = false;
      for (int s = 0; s < ns; ++s) if (seen[s] == n) { was = true; break; }
      if (was) continue;
      if (ns >= NB_SEEN) { atomicAdd(a.overflow, 1u); return; }
      seen[ns++] = n;
      const float qx = a.pos[3 * (size_t)n], qy = a.pos[3 * (size_t)n + 1], qz = a.pos[3 * (size_t)n + 2];
      const float dx = qx - px, dy = qy - py, dz = qz - pz;
      if (sqrtf(dx * dx + dy * dy + dz * dz) < radius) {
        if (WHICH & 1) { zmin = fminf(zmin, qz); zmax = fmaxf(zmax, qz); }
        if (WHICH & 6) {
          const float nnx = a.vn[3 * (size_t)n], nny = a.vn[3 * (size_t)n + 1], nnz = a.vn[3 * (size_t)n + 2];
          if (WHICH & 2) {
            float dot = nvx * nnx + nvy * nny + nvz * nnz;
            dot = fminf(1.0f, fmaxf(-1.0f, dot));
            rsum = rsum + acosf(dot); rcnt++;
          }
          if (WHICH & 4) {
            const float cx = (qx + nnx) - rx, cy = (qy + nny) - ry, cz = (qz + nnz) - rz;
            value += sqrtf(cx * cx + cy * cy + cz * cz); num++;
          }
        }
        if (sp >= NB_STACK) { atomicAdd(a.overflow, 1u); return; }
        stack[sp++] = n;
      }
    }
  }
}


// Same traversal, same visiting order (so the float sums are bit-identical to the oracle's), but the `seen` set is a
// 128-entry open-addressing hash in local memory instead of a linear list: ~2 probes per membership test instead of
// ~24 compares.  Neighbourhoods with more than NB_HSEEN seen vertices fall back to walk_linear.
constexpr int NB_HASH = 128, NB_HSEEN = 96;
template <int WHICH>
__device__ __forceinline__ void walk(const LayerKernelArgs& a, uint32_t v, float radius, float& zmin, float& zmax,
                                     float& rsum, int& rcnt, float& value, int& num) {
  uint32_t ht[NB_HASH]; uint32_t stack[NB_HSEEN];
#pragma unroll 8
  for (int i = 0; i < NB_HASH; ++i) ht[i] = 0xffffffffu;
  const float zmin0 = zmin, zmax0 = zmax, rsum0 = rsum, value0 = value; const int rcnt0 = rcnt, num0 = num;
  int ns = 0, sp = 0;
  ht[(v * 2654435761u) >> 25] = v; ns = 1; stack[sp++] = v;
  const float px = a.pos[3 * (size_t)v], py = a.pos[3 * (size_t)v + 1], pz = a.pos[3 * (size_t)v + 2];
  const float nvx = a.vn[3 * (size_t)v], nvy = a.vn[3 * (size_t)v + 1], nvz = a.vn[3 * (size_t)v + 2];
  const float rx = px + nvx, ry = py + nvy, rz = pz + nvz;
  bool overflow = false;
  while (sp > 0 && !overflow) {
    const uint32_t u = stack[--sp];
    for (uint32_t k = a.adj_ptr[u]; k < a.adj_ptr[u + 1]; ++k) {
      const uint32_t n = a.adj_nbr[k];
      uint32_t h = (n * 2654435761u) >> 25;
      bool was = false;
      for (;;) {
        const uint32_t e = ht[h];
        if (e == n) { was = true; break; }
        if (e == 0xffffffffu) break;
        h = (h + 1u) & (uint32_t)(NB_HASH - 1);
      }
      if (was) continue;
      if (ns >= NB_HSEEN) { overflow = true; break; }
      ht[h] = n; ++ns;
      const float qx = a.pos[3 * (size_t)n], qy = a.pos[3 * (size_t)n + 1], qz = a.pos[3 * (size_t)n + 2];
      const float dx = qx - px, dy = qy - py, dz = qz - pz;
      if (sqrtf(dx * dx + dy * dy + dz * dz) < radius) {
        if (WHICH & 1) { zmin = fminf(zmin, qz); zmax = fmaxf(zmax, qz); }
        if (WHICH & 6) {
          const float nnx = a.vn[3 * (size_t)n], nny = a.vn[3 * (size_t)n + 1], nnz = a.vn[3 * (size_t)n + 2];
          if (WHICH & 2) {
            float dot = nvx * nnx + nvy * nny + nvz * nnz;
            dot = fminf(1.0f, fmaxf(-1.0f, dot));
            rsum = rsum + acosf(dot); rcnt++;
          }
          if (WHICH & 4) {
            const float cx = (qx + nnx) - rx, cy = (qy + nny) - ry, cz = (qz + nnz) - rz;
            value += sqrtf(cx * cx + cy * cy + cz * cz); num++;
          }
        }
        stack[sp++] = n;            // sp <= ns <= NB_HSEEN
      }
    }
  }
  if (overflow) {
    zmin = zmin0; zmax = zmax0; rsum = rsum0; value = value0; rcnt = rcnt0; num = num0;
    walk_linear<WHICH>(a, v, radius, zmin, zmax, rsum, rcnt, value, num);
  }
}

// Variant with the `seen` hash set and the traversal stack in SHARED memory (slot-major, one bank per thread: conflict
// free) instead of thread-local memory: 1536 resident threads x ~0.9 KB of randomly probed local memory does not fit the
// L1, so every probe of the local-memory version is an L2 trip.  Same traversal, same visiting order, same sums.
// Opt-in (MNB_LAYERS_SMEM=1) until it has been timed on a B200.
constexpr int LS_THREADS = 128, LS_STACK = 48;
template <int WHICH>
__device__ __forceinline__ void walk_smem(const LayerKernelArgs& a, uint32_t v, float radius, float& zmin, float& zmax,
                                          float& rsum, int& rcnt, float& value, int& num, uint32_t* __restrict__ ht,
                                          uint32_t* __restrict__ stack) {
  // ht[slot * LS_THREADS], stack[i * LS_THREADS]: both already offset by threadIdx.x
#pragma unroll 8
  for (int i = 0; i < NB_HASH; ++i) ht[i * LS_THREADS] = 0xffffffffu;
  const float zmin0 = zmin, zmax0 = zmax, rsum0 = rsum, value0 = value; const int rcnt0 = rcnt, num0 = num;
  int ns = 0, sp = 0;
  ht[((v * 2654435761u) >> 25) * LS_THREADS] = v; ns = 1; stack[(sp++) * LS_THREADS] = v;
  const float px = a.pos[3 * (size_t)v], py = a.pos[3 * (size_t)v + 1], pz = a.pos[3 * (size_t)v + 2];
  const float nvx = a.vn[3 * (size_t)v], nvy = a.vn[3 * (size_t)v + 1], nvz = a.vn[3 * (size_t)v + 2];
  const float rx = px + nvx, ry = py + nvy, rz = pz + nvz;
  bool overflow = false;
  while (sp > 0 && !overflow) {
    const uint32_t u = stack[(--sp) * LS_THREADS];
    for (uint32_t k = a.adj_ptr[u]; k < a.adj_ptr[u + 1]; ++k) {
      const uint32_t n = a.adj_nbr[k];
      uint32_t h = (n * 2654435761u) >> 25;
      bool was = false;
      for (;;) {
        const uint32_t e = ht[h * LS_THREADS];
        if (e == n) { was = true; break; }
        if (e == 0xffffffffu) break;
        h = (h + 1u) & (uint32_t)(NB_HASH - 1);
      }
      if (was) continue;
      if (ns >= NB_HSEEN) { overflow = true; break; }
      ht[h * LS_THREADS] = n; ++ns;
      const float qx = a.pos[3 * (size_t)n], qy = a.pos[3 * (size_t)n + 1], qz = a.pos[3 * (size_t)n + 2];
      const float dx = qx - px, dy = qy - py, dz = qz - pz;
      if (sqrtf(dx * dx + dy * dy + dz * dz) < radius) {
        if (WHICH & 1) { zmin = fminf(zmin, qz); zmax = fmaxf(zmax, qz); }
        if (WHICH & 6) {
          const float nnx = a.vn[3 * (size_t)n], nny = a.vn[3 * (size_t)n + 1], nnz = a.vn[3 * (size_t)n + 2];
          if (WHICH & 2) {
            float dot = nvx * nnx + nvy * nny + nvz * nnz;
            dot = fminf(1.0f, fmaxf(-1.0f, dot));
            rsum = rsum + acosf(dot); rcnt++;
          }
          if (WHICH & 4) {
            const float cx = (qx + nnx) - rx, cy = (qy + nny) - ry, cz = (qz + nnz) - rz;
            value += sqrtf(cx * cx + cy * cy + cz * cz); num++;
          }
        }
        if (sp >= LS_STACK) { overflow = true; break; }
        stack[(sp++) * LS_THREADS] = n;
      }
    }
  }
  if (overflow) {
    zmin = zmin0; zmax = zmax0; rsum = rsum0; value = value0; rcnt = rcnt0; num = num0;
    walk_linear<WHICH>(a, v, radius, zmin, zmax, rsum, rcnt, value, num);
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(128) k_layers(const LayerKernelArgs a) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= a.V) return;
  const mnb_layer_params& P = a.P;
  const float pz = a.pos[3 * (size_t)v + 2];
  float zmin = pz, zmax = pz, rsum = 0.0f, value = 0.0f; int rcnt = 0, num = 0;
  const float r_hd = (float)P.height_diff_radius, r_ro = (float)P.roughness_radius, r_ri = (float)P.ridge_radius;
  if constexpr (SMEM) {
    MNB_DYNAMIC_SMEM(ls_raw);
    uint32_t* ht = reinterpret_cast<uint32_t*>(ls_raw) + threadIdx.x;
    uint32_t* stack = ht + NB_HASH * LS_THREADS;
    if (r_hd == r_ro && r_ro == r_ri) {
      walk_smem<7>(a, v, r_hd, zmin, zmax, rsum, rcnt, value, num, ht, stack);
    } else {
      walk_smem<1>(a, v, r_hd, zmin, zmax, rsum, rcnt, value, num, ht, stack);
      walk_smem<2>(a, v, r_ro, zmin, zmax, rsum, rcnt, value, num, ht, stack);
      walk_smem<4>(a, v, r_ri, zmin, zmax, rsum, rcnt, value, num, ht, stack);
    }
  } else if (r_hd == r_ro && r_ro == r_ri) {
    walk<7>(a, v, r_hd, zmin, zmax, rsum, rcnt, value, num);
  } else {
    walk<1>(a, v, r_hd, zmin, zmax, rsum, rcnt, value, num);
    walk<2>(a, v, r_ro, zmin, zmax, rsum, rcnt, value, num);
    walk<4>(a, v, r_ri, zmin, zmax, rsum, rcnt, value, num);
  }
  const float hd = zmax - zmin;
  const float ro = rcnt ? rsum / (float)rcnt : 0.0f;
  const float st = acosf(a.vn[3 * (size_t)v + 2]);                               // steepness_layer.cpp:165
  const float ri = num == 0 ? (float)(P.ridge_threshold + 0.1) : value / num;     // ridge_layer.cpp:177-184
  const float cl = a.clearance ? a.clearance[v] : __uint_as_float(INF_BITS);
  float cc; bool cl_lethal = false;                                              // clearance_layer.cpp:77-96
  const double inflated_height = P.clearance_robot_height + P.clearance_height_inflation;
  if (cl < P.clearance_robot_height) { cc = 1.0f; cl_lethal = true; }
  else if (cl < inflated_height) {
    const double diff = (cl - P.clearance_robot_height) / P.clearance_height_inflation;
    cc = (float)((cos(diff * 3.14159265358979323846) + 1.0) / 2.0);
  } else cc = 0.0f;
  const float bo = a.border[v] ? (float)P.border_cost : 0.0f;
  const size_t V = a.V;
  if (a.costs) {
    a.costs[v] = hd; a.costs[V + v] = ro; a.costs[2 * V + v] = st; a.costs[3 * V + v] = ri; a.costs[4 * V + v] = cc; a.costs[5 * V + v] = bo;
  }
  uint8_t mask = 0;
  if (hd > P.height_diff_threshold) mask |= 1;
  if (ro > P.roughness_threshold) mask |= 2;
  if (st > P.steepness_threshold) mask |= 4;
  if (ri > P.ridge_threshold) mask |= 8;
  if (cl_lethal) mask |= 16;
  if (bo > P.border_threshold) mask |= 32;
  if (a.lethal_mask) a.lethal_mask[v] = mask;
  if (a.combined) a.combined[v] = fmaxf(fmaxf(fmaxf(0.0f, hd), fmaxf(ro, st)), fmaxf(fmaxf(ri, cc), bo));   // combination_layer.cpp:60-71
}

// ============================================================================
// Vector-field epilogues: DijkstraMeshPlanner::computeVectorMap (dijkstra_mesh_planner.cpp:189-209) and
// CVPMeshPlanner::computeVectorMap (cvp_mesh_planner.cpp:204-239).  NaN = "no entry in the sparse map".
// ============================================================================
struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 f3cross(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float f3dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 f3load(const float* __restrict__ p, uint32_t v) { return {p[3 * (size_t)v], p[3 * (size_t)v + 1], p[3 * (size_t)v + 2]}; }
__device__ __forceinline__ F3 f3normalized(F3 v) {
  const float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (l > 0) { v.x /= l; v.y /= l; v.z /= l; }
  return v;
}

// one entry of the planners' vector map; false = "no entry" (pred == self or no cutting face)
__device__ __forceinline__ bool vertex_vector(const float* __restrict__ pos, const float* __restrict__ vn,
                                              const uint32_t* __restrict__ pred, const float* __restrict__ direction,
                                              const int32_t* __restrict__ cut, uint32_t v3, F3& out) {
  const uint32_t v1 = pred[v3];
  if (v1 == v3 || (cut && cut[v3] < 0)) return false;
  F3 v = f3sub(f3load(pos, v1), f3load(pos, v3));
  if (direction) {   // rotate about the vertex normal by the stored angle (Rodrigues; lvr2 BaseVector::rotated)
    const F3 n = f3load(vn, v3);
    const double alpha = (double)direction[v3];
    const float sina = (float)sin(alpha), cosa = (float)cos(alpha);
    const float ndotv = f3dot(n, v);
    const F3 c = f3cross(n, v);
    v = {v.x * cosa + c.x * sina + n.x * ndotv * (1.0f - cosa), v.y * cosa + c.y * sina + n.y * ndotv * (1.0f - cosa),
         v.z * cosa + c.z * sina + n.z * ndotv * (1.0f - cosa)};
  }
  out = f3normalized(v);
  return true;
}

__global__ void k_vector_map(const float* __restrict__ pos, const float* __restrict__ vn, const uint32_t* __restrict__ pred,
                             const float* __restrict__ direction, const int32_t* __restrict__ cut, uint32_t V,
                             float* __restrict__ out) {
  const uint32_t v3 = blockIdx.x * blockDim.x + threadIdx.x;
  if (v3 >= V) return;
  const float nan = __int_as_float(0x7fc00000);
  F3 v{nan, nan, nan};
  vertex_vector(pos, vn, pred, direction, cut, v3, v);
  float* o = out + 3 * (size_t)v3;
  o[0] = v.x; o[1] = v.y; o[2] = v.z;
}


// ============================================================================
// InflationLayer repulsive vector field, vector_map_ (inflation_layer.cpp:277-308), from the final labels of k_inflate.
// The reference accumulates it inside the sequential loop; the result factors into two phases (oracle: orc_inflation):
//  (1) while the lethal vertices pop (all at key 0, in id order) every face with exactly two lethal vertices adds its
//      direction to the vectors of its three vertices, once per (popping vertex, incident edge of the face, side of the
//      edge) -- `vec = (vec + dir).normalized()` in exactly that order (:277-295).  Per vertex this is an ordered fold over
//      at most 4 events per incident face: gathered, sorted by (popping vertex, edge position, side) and folded here.
//  (2) afterwards a vertex' vector is overwritten by every accepted update with a non-lethal source,
//      (vec[v1]*(u3-u1) + vec[v2]*(u3-u2)).normalized() (:301-308): the LAST accepted face of the event-ordered replay
//      decides; its sources popped earlier, so their vectors are final -- evaluated by fixed-point iteration over the
//      (acyclic) source relation.
// ============================================================================
struct InflVecArgs {
  uint32_t V;
  const float* pos; const uint32_t* faces;
  const uint32_t* cor_ptr; const int4* cor_idx; const float4* cor_wd; const uint4* cor_eid;
  const uint32_t* adj_ptr; const uint32_t* adj_nbr;
  const uint8_t* invalid;
  WaveWorkspace ws;
  float max_distance;
  float* vec;                  // 3V, zero = no entry
  int4* src;                   // {v1, v2, bits(u3-u1), bits(u3-u2)}; v1 = -1: no overwrite
  unsigned int* flag;          // [0] a vector changed in this sweep, [1] scratch overflow
};
constexpr int IV_MAXF = 24, IV_MAXE = 4 * IV_MAXF;

__device__ __forceinline__ float iv_len(const float* __restrict__ pos, uint32_t p, uint32_t q) {   // == k_edge_dist
  const float dx = pos[3 * (size_t)p] - pos[3 * (size_t)q], dy = pos[3 * (size_t)p + 1] - pos[3 * (size_t)q + 1],
              dz = pos[3 * (size_t)p + 2] - pos[3 * (size_t)q + 2];
  return sqrtf(dx * dx + dy * dy + dz * dz);
}

__global__ void __launch_bounds__(128) k_infl_vec_lethal(const InflVecArgs a) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= a.V) return;
  float* out = a.vec + 3 * (size_t)x;
  out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f;
  unsigned long long key[IV_MAXE]; uint8_t ev_face[IV_MAXE];
  F3 fdir[IV_MAXF];
  int ne = 0, nf = 0;
  for (uint32_t k = a.cor_ptr[x]; k < a.cor_ptr[x + 1]; ++k) {
    const uint32_t f = (uint32_t)a.cor_idx[k].z;
    const uint32_t fa = a.faces[3 * (size_t)f], fb = a.faces[3 * (size_t)f + 1], fc = a.faces[3 * (size_t)f + 2];
    const bool la = __uint_as_float(a.ws.state[fa].x) == 0.0f, lb = __uint_as_float(a.ws.state[fb].x) == 0.0f,
               lc = __uint_as_float(a.ws.state[fc].x) == 0.0f;
    uint32_t w1, w2, w3;                                       // argument order of waveFrontUpdate (:445-470)
    if (la && lb && !lc) { w1 = fa; w2 = fb; w3 = fc; }
    else if (la && !lb && lc) { w1 = fc; w2 = fa; w3 = fb; }
    else if (!la && lb && lc) { w1 = fb; w2 = fc; w3 = fa; }
    else continue;
    const float cand = inflation_candidate(0.0f, 0.0f, iv_len(a.pos, w2, w3), iv_len(a.pos, w1, w3), iv_len(a.pos, w1, w2));
    if (__float_as_uint(cand) == INF_BITS) continue;           // :271 non-finite update: the call returns before the vectors
    if (nf >= IV_MAXF) { atomicAdd(&a.flag[1], 1u); return; }
    const F3 p1 = f3load(a.pos, w1), p2 = f3load(a.pos, w2), p3 = f3load(a.pos, w3);
    fdir[nf] = f3normalized(F3{(p3.x - p2.x) + (p3.x - p1.x), (p3.y - p2.y) + (p3.y - p1.y), (p3.z - p2.z) + (p3.z - p1.z)});
    const uint32_t lv[2] = {w1, w2};
    for (int s = 0; s < 2; ++s) {
      const uint32_t p = lv[s];                                // the popping lethal vertex
      if (a.invalid && a.invalid[p]) continue;                 // pops but does not expand (:417)
      const uint32_t others[2] = {p == w1 ? w2 : w1, w3};
      for (int t = 0; t < 2; ++t) {
        const uint32_t q = others[t];
        uint32_t epos = 0;                                     // position of edge (p,q) among p's edges (ascending edge id)
        for (uint32_t kk = a.adj_ptr[p]; kk < a.adj_ptr[p + 1]; ++kk) if (a.adj_nbr[kk] == q) { epos = kk - a.adj_ptr[p]; break; }
        uint32_t side = 0;                                     // an edge lists its faces in ascending id
        for (uint32_t kk = a.cor_ptr[p]; kk < a.cor_ptr[p + 1]; ++kk) {
          const int4 ix = a.cor_idx[kk];
          if ((uint32_t)ix.z != f && ((uint32_t)ix.x == q || (uint32_t)ix.y == q)) { side = (uint32_t)ix.z < f ? 1u : 0u; break; }
        }
        key[ne] = ((unsigned long long)p << 32) | ((unsigned long long)epos << 1) | side;
        ev_face[ne] = (uint8_t)nf; ++ne;
      }
    }
    ++nf;
  }
  if (ne == 0) return;
  F3 v{0.0f, 0.0f, 0.0f};
  for (int i = 0; i < ne; ++i) {                               // selection sort: <= 96 events, usually 4-16
    int b = i;
    for (int j = i + 1; j < ne; ++j) if (key[j] < key[b]) b = j;
    const unsigned long long kb = key[b]; const uint8_t fbi = ev_face[b];
    key[b] = key[i]; ev_face[b] = ev_face[i]; key[i] = kb; ev_face[i] = fbi;
    const F3 d = fdir[fbi];
    v = f3normalized(F3{v.x + d.x, v.y + d.y, v.z + d.z});
  }
  out[0] = v.x; out[1] = v.y; out[2] = v.z;
}

__global__ void __launch_bounds__(128) k_infl_vec_sources(const InflVecArgs a) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.V) return;
  int4 r = make_int4(-1, -1, 0, 0);
  const float d = __uint_as_float(a.ws.state[c].x);
  if (d != 0.0f && __float_as_uint(d) != INF_BITS) {
    InflationProblem prob;
    prob.cor_ptr = a.cor_ptr; prob.cor_idx = a.cor_idx; prob.cor_wd = a.cor_wd; prob.cor_eid = a.cor_eid; prob.invalid = a.invalid;
    prob.state = a.ws.state; prob.minor_arr = a.ws.minor; prob.root_arr = a.ws.root; prob.chg = a.ws.chg;
    prob.deferred_m = __uint_as_float(INF_BITS); prob.strict = 0; prob.max_distance = a.max_distance;
    float nd, wu1, wu2; EvTime tc; int win;
    prob.replay(c, __uint_as_float(INF_BITS), 0xfffffff0u /* final labels: nothing is deferred */, nd, tc, win, wu1, wu2);
    if (win >= 0 && (wu1 != 0.0f || wu2 != 0.0f)) {           // :301 (an update from two lethal sources keeps the phase-1 vector)
      const int4 ix = a.cor_idx[win];
      r = make_int4(ix.x, ix.y, __float_as_int(nd - wu1), __float_as_int(nd - wu2));
    }
  }
  a.src[c] = r;
}

__global__ void __launch_bounds__(256) k_infl_vec_sweep(const InflVecArgs a) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.V) return;
  const int4 r = a.src[c];
  if (r.x < 0) return;
  const float d31 = __int_as_float(r.z), d32 = __int_as_float(r.w);
  const float* va = a.vec + 3 * (size_t)r.x; const float* vb = a.vec + 3 * (size_t)r.y;
  const F3 v = f3normalized(F3{va[0] * d31 + vb[0] * d32, va[1] * d31 + vb[1] * d32, va[2] * d31 + vb[2] * d32});   // :306
  float* out = a.vec + 3 * (size_t)c;
  if (__float_as_uint(out[0]) != __float_as_uint(v.x) || __float_as_uint(out[1]) != __float_as_uint(v.y) ||
      __float_as_uint(out[2]) != __float_as_uint(v.z)) {
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
    a.flag[0] = 1u;
  }
}

// InflationLayer::vectorAt(vertices, barycentric_coords) (inflation_layer.cpp:493-521); see oracle inflationVectorAt
struct RepulsiveField {
  const float* dist; const float* vec;   // null = no repulsive layer
  float inscribed_radius_f; double inscribed_radius, inflation_radius; float lethal_value, inscribed_value;
};
__device__ __forceinline__ F3 inflation_vector_at(const RepulsiveField& L, const uint32_t* __restrict__ t, const float bary[3]) {
  const float d0 = L.dist[t[0]], d1 = L.dist[t[1]], d2 = L.dist[t[2]];
  if (!isfinite(d0) || !isfinite(d1) || !isfinite(d2)) return F3{0.0f, 0.0f, 0.0f};
  const float distance = d0 * bary[0] + d1 * bary[1] + d2 * bary[2];
  if ((double)distance > L.inflation_radius) return F3{0.0f, 0.0f, 0.0f};
  const F3 va = f3load(L.vec, t[0]), vb = f3load(L.vec, t[1]), vc = f3load(L.vec, t[2]);
  const F3 v{va.x * bary[0] + vb.x * bary[1] + vc.x * bary[2], va.y * bary[0] + vb.y * bary[1] + vc.y * bary[2],
             va.z * bary[0] + vb.z * bary[1] + vc.z * bary[2]};
  if ((double)distance > L.inscribed_radius) {
    const float alpha = (float)(((double)sqrtf(distance) - L.inscribed_radius) / (L.inflation_radius - L.inscribed_radius) * 3.14159265358979323846);
    const float s1 = L.inscribed_value, s2 = cosf(alpha) + 1, s3 = 2.0f;
    return F3{v.x * s1 * s2 / s3, v.y * s1 * s2 / s3, v.z * s1 * s2 / s3};
  }
  const float s = distance > 0 ? L.inscribed_value : L.lethal_value;
  return F3{v.x * s, v.y * s, v.z * s};
}
__global__ void k_inflation_vector_at(const RepulsiveField L, const uint32_t* __restrict__ faces, uint32_t n,
                                      const uint32_t* __restrict__ faces_q, const float* __restrict__ bary, float* __restrict__ out) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const float b[3] = {bary[3 * (size_t)q], bary[3 * (size_t)q + 1], bary[3 * (size_t)q + 2]};
  const F3 v = inflation_vector_at(L, faces + 3 * (size_t)faces_q[q], b);
  out[3 * (size_t)q] = v.x; out[3 * (size_t)q + 1] = v.y; out[3 * (size_t)q + 2] = v.z;
}

// ============================================================================
// Vector-field back-tracking (cvp_mesh_planner.cpp:920-951): MeshMap::meshAhead (mesh_map.cpp:1070-1108),
// searchNeighbourFaces (:999-1068), directionAtPosition (:625-650), projectedBarycentricCoords (util.cpp:313-347).
// A strictly sequential walk of a few hundred steps: one thread follows the field on the device-resident result of
// the last plan, so a makePlan moves a few KB of poses over PCIe instead of four V-sized arrays.
// ============================================================================
struct BacktrackArgs {
  const float* pos; const float* vn; const uint32_t* faces; const uint32_t* cor_ptr; const int4* cor_idx;
  const uint32_t* pred; const float* direction; const int32_t* cut;
  float start[3]; uint32_t start_face; float goal[3]; uint32_t goal_face;
  double step_width; uint32_t max_points;
  float* path_pos; uint32_t* path_face; int32_t* result /* [0] outcome, [1] n_points */; const int* cancel_flag;
  RepulsiveField layer;        // InflationLayer::vectorAt added in meshAhead (mesh_map.cpp:1097-1102); dist == null: none
};

__device__ __forceinline__ bool projected_barycentric(F3 p, F3 a, F3 b, F3 c, float bary[3], float& dist) {
  const F3 u = f3sub(b, a), v = f3sub(c, a), w = f3sub(p, a), n = f3cross(u, v);
  const float oneOver4ASquared = (float)(1.0 / (double)f3dot(n, n));
  const float gamma = f3dot(f3cross(u, w), n) * oneOver4ASquared;
  const float beta = f3dot(f3cross(w, v), n) * oneOver4ASquared;
  const float alpha = 1 - gamma - beta;
  bary[0] = alpha; bary[1] = beta; bary[2] = gamma;
  dist = f3dot(n, w) / sqrtf(f3dot(n, n));
  const float EPSILON = 0.01f;
  return (0 - EPSILON <= alpha) && (alpha <= 1 + EPSILON) && (0 - EPSILON <= beta) && (beta <= 1 + EPSILON) &&
         (0 - EPSILON <= gamma) && (gamma <= 1 + EPSILON);
}

// One warp walks the path.  The scalar parts (position, current face, direction blend) are computed redundantly by
// all lanes; searchNeighbourFaces -- a breadth-first list of up to a few hundred faces per step -- is spread over
// the lanes 32 faces at a time: containment tests in parallel, expansions gathered per lane, deduplicated through a
// shared-memory hash set and appended in exactly the order the sequential loop of mesh_map.cpp:1031-1063 would
// produce (lane-major sequence numbers + atomicMin decide which duplicate came first), so the face that is returned
// is the same one.
constexpr int BT_LIST_CAP = 4096;      // faces in the search list
constexpr int BT_HASH_CAP = 8192;      // open-addressing set over face ids (power of two)
constexpr int BT_MAXC = 64;            // expansion candidates of one listed face (<= sum of its vertices' face counts)
struct BtShared {
  uint32_t list[BT_LIST_CAP];
  uint32_t hkey[BT_HASH_CAP];
  uint32_t hseq[BT_HASH_CAP];
  uint32_t cand[32][BT_MAXC];
  uint16_t cslot[32][BT_MAXC];
};

__global__ void __launch_bounds__(32) k_backtrack(BacktrackArgs a) {
  MNB_DYNAMIC_SMEM(bt_raw);
  BtShared& S = *reinterpret_cast<BtShared*>(bt_raw);
  constexpr unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x;
  uint32_t n = 0;
  auto push = [&](F3 p, uint32_t f) {
    if (lane == 0 && n < a.max_points) { a.path_pos[3 * n] = p.x; a.path_pos[3 * n + 1] = p.y; a.path_pos[3 * n + 2] = p.z; a.path_face[n] = f; }
    ++n;
  };
  uint32_t face = a.goal_face;
  F3 pos{a.goal[0], a.goal[1], a.goal[2]};
  const F3 st{a.start[0], a.start[1], a.start[2]};
  const float step = (float)a.step_width;
  push(pos, face);
  int32_t outcome = MNB_SUCCESS;
  for (;;) {
    const F3 d = f3sub(pos, st);
    if (!((double)f3dot(d, d) > a.step_width)) break;                            // cvp:925 (distance2 vs step_width, as written)
    if (a.cancel_flag && *(volatile const int*)a.cancel_flag) { outcome = MNB_CANCELED; break; }
    if (n + 1 >= a.max_points) { outcome = MNB_E_STATE; break; }
    // ---- meshAhead ----
    float bary[3], dist;
    const uint32_t* t = a.faces + 3 * (size_t)face;
    bool ok = projected_barycentric(pos, f3load(a.pos, t[0]), f3load(a.pos, t[1]), f3load(a.pos, t[2]), bary, dist);
    if (!ok) {                                                                    // searchNeighbourFaces(pos, face, step, 0.4)
      F3 center{0, 0, 0};
      for (int k = 0; k < 3; ++k) { const F3 q = f3load(a.pos, t[k]); center = {center.x + q.x, center.y + q.y, center.z + q.z}; }
      center = {center.x / 3, center.y / 3, center.z / 3};
      float vcm = 0;
      for (int k = 0; k < 3; ++k) { const F3 e = f3sub(f3load(a.pos, t[k]), center); vcm = fmaxf(vcm, sqrtf(f3dot(e, e))); }
      const float ext = step + vcm, rsq = ext * ext;
      for (uint32_t i = lane; i < (uint32_t)BT_HASH_CAP; i += 32) { S.hkey[i] = 0xffffffffu; S.hseq[i] = 0xffffffffu; }
      __syncwarp();
      auto slot_of = [](uint32_t key) { return (key * 2654435761u) >> (32 - 13); };
      static_assert(BT_HASH_CAP == (1 << 13), "hash shift");
      if (lane == 0) {
        S.list[0] = face;
        const uint32_t h = slot_of(face); S.hkey[h] = face; S.hseq[h] = 0u;
      }
      __syncwarp();
      uint32_t cnt = 1, it = 0;
      bool overflow = false;
      while (it < cnt && !ok) {
        const uint32_t chunk = min(32u, cnt - it);
        const bool act = lane < chunk;
        const uint32_t f = act ? S.list[it + lane] : 0u;
        const uint32_t* q = a.faces + 3 * (size_t)f;
        const uint32_t q0 = q[0], q1 = q[1], q2 = q[2];
        const F3 P0 = f3load(a.pos, q0), P1 = f3load(a.pos, q1), P2 = f3load(a.pos, q2);
        float lb[3], ld;
        const bool pass = act && projected_barycentric(pos, P0, P1, P2, lb, ld) && fabsf(ld) < 0.4f;
        const unsigned pm = __ballot_sync(FULL, pass);
        if (pm) {
          const int L = __ffs(pm) - 1;
          face = __shfl_sync(FULL, f, L);
          bary[0] = __shfl_sync(FULL, lb[0], L); bary[1] = __shfl_sync(FULL, lb[1], L); bary[2] = __shfl_sync(FULL, lb[2], L);
          ok = true;
          break;
        }
        // expansion candidates of my face, in the reference's order: vertex 0, 1, 2; faces of the vertex in CSR order
        uint32_t nc = 0;
        if (act) {
          const uint32_t qv[3] = {q0, q1, q2};
          const F3 PV[3] = {P0, P1, P2};
          for (int k = 0; k < 3; ++k) {
            const F3 e = f3sub(center, PV[k]);
            if (!(f3dot(e, e) < rsq)) continue;
            for (uint32_t jx = a.cor_ptr[qv[k]]; jx < a.cor_ptr[qv[k] + 1]; ++jx) {
              if (nc < (uint32_t)BT_MAXC) S.cand[lane][nc] = (uint32_t)a.cor_idx[jx].z;
              ++nc;
            }
          }
        }
        if (__any_sync(FULL, nc > (uint32_t)BT_MAXC)) { overflow = true; break; }
        __syncwarp();
        // insert all candidates; among duplicates the smallest sequence number (= first in sequential order) wins
        for (uint32_t i = 0; i < nc; ++i) {
          const uint32_t key = S.cand[lane][i], seq = lane * (uint32_t)BT_MAXC + i + 1u;
          uint32_t h = slot_of(key);
          for (;;) {
            const uint32_t old = atomicCAS(&S.hkey[h], 0xffffffffu, key);
            if (old == 0xffffffffu || old == key) break;
            h = (h + 1u) & (uint32_t)(BT_HASH_CAP - 1);
          }
          atomicMin(&S.hseq[h], seq);
          S.cslot[lane][i] = (uint16_t)h;
        }
        __syncwarp();
        uint32_t newc = 0;
        for (uint32_t i = 0; i < nc; ++i) newc += (S.hseq[S.cslot[lane][i]] == lane * (uint32_t)BT_MAXC + i + 1u) ? 1u : 0u;
        uint32_t incl = newc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += v; }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (cnt + total > (uint32_t)BT_LIST_CAP) { overflow = true; break; }
        uint32_t w = cnt + incl - newc;
        __syncwarp();
        for (uint32_t i = 0; i < nc; ++i) {
          const uint32_t h = S.cslot[lane][i];
          if (S.hseq[h] == lane * (uint32_t)BT_MAXC + i + 1u) S.list[w++] = S.cand[lane][i];
        }
        __syncwarp();
        for (uint32_t i = 0; i < nc; ++i) {            // commit: members of the list can never be "new" again
          const uint32_t h = S.cslot[lane][i];
          if (S.hseq[h] == lane * (uint32_t)BT_MAXC + i + 1u) S.hseq[h] = 0u;
        }
        __syncwarp();
        cnt += total; it += chunk;
      }
      if (!ok) { outcome = overflow ? MNB_E_STATE : MNB_NO_PATH_FOUND; break; }
      t = a.faces + 3 * (size_t)face;
      const F3 A = f3load(a.pos, t[0]), B = f3load(a.pos, t[1]), C = f3load(a.pos, t[2]);   // project onto the surface
      pos = {A.x * bary[0] + B.x * bary[1] + C.x * bary[2], A.y * bary[0] + B.y * bary[1] + C.y * bary[2],
             A.z * bary[0] + B.z * bary[1] + C.z * bary[2]};
    }
    // ---- directionAtPosition ----
    bool any = false;
    F3 vec{0, 0, 0};
    for (int k = 0; k < 3; ++k) {
      F3 e;
      if (!vertex_vector(a.pos, a.vn, a.pred, a.direction, a.cut, t[k], e)) continue;
      any = true;
      vec = {vec.x + e.x * bary[k], vec.y + e.y * bary[k], vec.z + e.z * bary[k]};
    }
    if (!any || !(isfinite(vec.x) && isfinite(vec.y) && isfinite(vec.z))) { outcome = MNB_NO_PATH_FOUND; break; }
    F3 dir = f3normalized(vec);                          // opt_dir.get().normalized()
    if (a.layer.dist) { const F3 lv = inflation_vector_at(a.layer, t, bary); dir = F3{dir.x + lv.x, dir.y + lv.y, dir.z + lv.z}; }
    dir = f3normalized(dir);                             // dir += layer->vectorAt(...); dir.normalize()
    pos = {pos.x + dir.x * step, pos.y + dir.y * step, pos.z + dir.z * step};
    push(pos, face);
  }
  if (outcome == MNB_SUCCESS) push(st, a.start_face);                            // cvp:951
  if (lane == 0) { a.result[0] = outcome; a.result[1] = (int32_t)n; }
}

// ============================================================================
// Localisation: MeshMap::getNearestVertexHandle (mesh_map.cpp:1161-1174; the reference walks a nanoflann KD-tree,
// here the 12 B/vertex position array is streamed once for ALL queries of the call -- HBM-bound, ~8 us per 5M-vertex
// pass) and MeshMap::searchContainingFace (mesh_map.cpp:1120-1159).
// Key = (squared distance bits << 32) | vertex id: non-negative floats order like their bit patterns, so one 64-bit
// atomicMin yields the nearest vertex with ties to the lowest id.
// ============================================================================
constexpr int LOC_Q = 32;       // queries per pass (registers)

__global__ void __launch_bounds__(256) k_nearest_vertex(const float* __restrict__ pos, uint32_t V, const float* __restrict__ points,
                                                        uint32_t q0, uint32_t nq, unsigned long long* __restrict__ keys) {
  __shared__ float sq[3 * LOC_Q];
  __shared__ unsigned long long sbest[LOC_Q];
  if (threadIdx.x < 3 * nq) sq[threadIdx.x] = points[3 * (size_t)q0 + threadIdx.x];
  if (threadIdx.x < LOC_Q) sbest[threadIdx.x] = ~0ull;
  __syncthreads();
  unsigned long long best[LOC_Q];
#pragma unroll
  for (int q = 0; q < LOC_Q; ++q) best[q] = ~0ull;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const float x = pos[3 * (size_t)v], y = pos[3 * (size_t)v + 1], z = pos[3 * (size_t)v + 2];
#pragma unroll
    for (int q = 0; q < LOC_Q; ++q) {
      if (q < (int)nq) {
        const float dx = sq[3 * q] - x, dy = sq[3 * q + 1] - y, dz = sq[3 * q + 2] - z;
        const float d = dx * dx + dy * dy + dz * dz;
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | v;
        if (!(d != d) && key < best[q]) best[q] = key;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < LOC_Q; ++q) {
    if (q < (int)nq) {
      unsigned long long b = best[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o); b = t < b ? t : b; }
      if ((threadIdx.x & 31) == 0 && b != ~0ull) atomicMin(&sbest[q], b);
    }
  }
  __syncthreads();
  if (threadIdx.x < nq && sbest[threadIdx.x] != ~0ull) atomicMin(&keys[q0 + threadIdx.x], sbest[threadIdx.x]);
}

__global__ void k_containing_face(const float* __restrict__ pos, const uint32_t* __restrict__ faces, const uint32_t* __restrict__ cor_ptr,
                                  const int4* __restrict__ cor_idx, const float* __restrict__ points, uint32_t n,
                                  const unsigned long long* __restrict__ keys, uint32_t* __restrict__ out_vertex,
                                  int32_t* __restrict__ out_face, float* __restrict__ out_bary) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t best = (uint32_t)(keys[q] & 0xffffffffu);
  if (out_vertex) out_vertex[q] = best;
  const F3 p = f3load(points, q);
  float lowest = 3.402823466e+38f;
  int32_t bf = -1; float bb[3] = {0, 0, 0};
  for (uint32_t k = cor_ptr[best]; k < cor_ptr[best + 1]; ++k) {
    const uint32_t f = (uint32_t)cor_idx[k].z;
    const uint32_t* t = faces + 3 * (size_t)f;
    float cb[3], dist = 0;
    if (projected_barycentric(p, f3load(pos, t[0]), f3load(pos, t[1]), f3load(pos, t[2]), cb, dist) && dist < lowest) {
      lowest = dist; bf = (int32_t)f; bb[0] = cb[0]; bb[1] = cb[1]; bb[2] = cb[2];
    }
  }
  if (out_face) out_face[q] = bf;
  if (out_bary) { out_bary[3 * (size_t)q] = bb[0]; out_bary[3 * (size_t)q + 1] = bb[1]; out_bary[3 * (size_t)q + 2] = bb[2]; }
}

// ============================================================================
// InflationLayer::waveCostInflation (inflation_layer.cpp:341-491): whole-grid cooperative kernel
// (multi-source: few, very wide rounds) + fading epilogue (:482-490, :315-339)
// ============================================================================
struct InflateKernelArgs {
  uint32_t V;
  const uint32_t* cor_ptr; const int4* cor_idx; const float4* cor_wd; const uint4* cor_eid;
  const uint8_t* invalid;
  WaveWorkspace ws;
  const uint32_t* lethals; uint32_t n_lethals;
  float max_distance;
  InflationParams params;
  float* out_dist; float* out_cost;
  uint32_t max_rounds;
};

__global__ void __launch_bounds__(512, 1) k_inflate(const InflateKernelArgs a) {
  __shared__ Stage st;
  uint32_t g, gthreads, gtid;
  group_coords<0>(g, gthreads, gtid);
  const uint32_t V = a.V;
  uint4* state = a.ws.state; uint32_t* mark = a.ws.mark; uint32_t* list0 = a.ws.list0; uint32_t* list1 = a.ws.list1;
  GroupCtl* ctl = a.ws.ctl;
  if (threadIdx.x == 0) { st.n = 0; st.m_tau = INF_BITS; st.lo = INF_BITS; }
  __syncthreads();
  for (uint32_t v = gtid; v < V; v += gthreads) { state[v] = state_inf(); mark[v] = MARK_NONE; a.ws.chg[v] = 0u; }
  group_sync<0>(ctl->barrier);
  for (uint32_t i = gtid; i < a.n_lethals; i += gthreads) {      // :397-402
    const uint32_t v = a.lethals[i];
    if (v < V) { state[v] = make_uint4(0u, 0u, 0u, 0u); mark[v] = MARK_FIXED; }
  }
  if (gtid == 0) ctl_reset(ctl, 0, 0.0f);
  group_sync<0>(ctl->barrier);
  InflationProblem prob;
  prob.cor_ptr = a.cor_ptr; prob.cor_idx = a.cor_idx; prob.cor_wd = a.cor_wd; prob.cor_eid = a.cor_eid; prob.invalid = a.invalid;
  prob.state = state; prob.minor_arr = a.ws.minor; prob.root_arr = a.ws.root; prob.chg = a.ws.chg; prob.deferred_m = __uint_as_float(INF_BITS); prob.max_distance = a.max_distance;
  for (uint32_t i = gtid; i < a.n_lethals; i += gthreads) {
    const uint32_t v = a.lethals[i];
    if (v >= V) continue;
    prob.activate(v, [&](uint32_t x) {
      if (__ldcg(&mark[x]) == MARK_NONE && atomicCAS(&mark[x], MARK_NONE, MARK_CAND) == MARK_NONE)
        stage_push(st, x, list0, &ctl->count[0]);
    });
  }
  stage_flush(st, list0, &ctl->count[0], &ctl->m_tau[0], &ctl->lo[0]);
  group_sync<0>(ctl->barrier);
  run_band_rounds<0>(prob, ctl, list0, list1, mark, st, __uint_as_float(INF_BITS), gthreads, gtid, 0, 0u, 0u, 0u, 0.0,
                     nullptr, 1e-30f, a.max_rounds);
  group_sync<0>(ctl->barrier);
  for (uint32_t v = gtid; v < V; v += gthreads) {
    const float d = __uint_as_float(state[v].x);
    if (a.out_dist) a.out_dist[v] = d;
    if (a.out_cost) a.out_cost[v] = (__float_as_uint(d) == INF_BITS) ? __int_as_float(0x7fc00000) : fading(a.params, d);
  }
}


// ============================================================================
// Incremental updates: MeshMap::layerChanged + updateEdgeWeights (mesh_map.cpp:455-492, 563-618),
// MaxCombinationLayer::onInputChanged (combination_layer.cpp:87-147), the update set of InflationLayer::onInputChanged
// (inflation_layer.cpp:154-164).  Work is proportional to the changed set: 8 lanes per changed vertex walk its incident
// edges / faces and patch only the table entries that hold one of those edges' weights.
// ============================================================================
__global__ void k_update_costs(const uint32_t* __restrict__ changed, uint32_t n, const float* __restrict__ costs, int by_vertex,
                               float default_value, uint32_t V, float* __restrict__ cost) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = changed[i];
  if (v >= V) return;
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
  const uint32_t* cor_ptr; const int4* cor_idx; const uint4* cor_eid;
  const uint32_t* adj_ptr; const uint32_t* adj_nbr; const uint32_t* adj_eid;
  const float* w;
  float4* cor_w; float4* ell_w; double4* ell_geo; uint2* adj_nw; uint4* ell_adj;
};
// every table entry that stores the weight of an edge incident to a changed vertex: the corner records (CSR + ELL +
// precomputed unfolding geometry) of all three vertices of each incident face, and both directions of the adjacency
__global__ void k_refresh_weight_tables(const RefreshArgs a) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)a.n * ELL_W) return;
  const uint32_t v = a.changed[t / ELL_W], j = (uint32_t)(t % ELL_W);
  if (v >= a.V) return;
  for (uint32_t k = a.cor_ptr[v] + j; k < a.cor_ptr[v + 1]; k += ELL_W) {
    const uint32_t f = (uint32_t)a.cor_idx[k].z;
    for (int c = 0; c < 3; ++c) {
      const uint32_t x = a.faces[3 * (size_t)f + c];
      const uint32_t kb = a.cor_ptr[x], ke = a.cor_ptr[x + 1];
      for (uint32_t kk = kb; kk < ke; ++kk) {
        if ((uint32_t)a.cor_idx[kk].z != f) continue;
        const uint4 e = a.cor_eid[kk];
        const float4 ww = make_float4(a.w[e.x], a.w[e.y], a.w[e.z], 0.0f);
        a.cor_w[kk] = ww;
        if (kk - kb < ELL_W) {
          const size_t s = (size_t)x * ELL_W + (kk - kb);
          a.ell_w[s] = ww;
          const CvpEllProblem::FaceGeo g = CvpEllProblem::face_geo((double)ww.z, (double)ww.y, (double)ww.x);
          a.ell_geo[s] = make_double4(g.p, g.hc, g.t0a, 0.0);
        }
        break;
      }
    }
  }
  const uint32_t ab = a.adj_ptr[v];
  for (uint32_t k = ab + j; k < a.adj_ptr[v + 1]; k += ELL_W) {
    const uint32_t u = a.adj_nbr[k], wb = __float_as_uint(a.w[a.adj_eid[k]]);
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
    cost = fmaxf(tmp, cost);                               // std::max(tmp, cost): NaN never enters (tmp is not NaN... unless def is)
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

// ============================================================================
// host side
// ============================================================================
struct mnb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int ptr_mode = MNB_PTR_HOST;
  int sm_count = 0;
  // mesh
  uint32_t V = 0, F = 0, E = 0;
  size_t NC = 0, NA = 0;
  HostTopology topo;
  float* d_pos = nullptr; uint32_t* d_faces = nullptr; uint32_t* d_edges = nullptr;
  uint32_t* d_cor_ptr = nullptr; int4* d_cor_idx = nullptr; uint4* d_cor_eid = nullptr;
  float4* d_cor_w = nullptr; float4* d_cor_wd = nullptr;
  int4* d_ell_idx = nullptr; uint4* d_ell_eid = nullptr; float4* d_ell_w = nullptr; float4* d_ell_wd = nullptr; double4* d_ell_geo = nullptr;
  uint32_t* d_adj_ptr = nullptr; uint32_t* d_adj_nbr = nullptr; uint32_t* d_adj_eid = nullptr; uint2* d_adj_nw = nullptr; uint4* d_ell_adj = nullptr;
  float* d_edge_dist = nullptr; float* d_edge_w = nullptr; float* d_cost = nullptr; uint8_t* d_invalid = nullptr;
  bool has_invalid = false, costs_set = false;
  // workspace
  uint32_t ws_groups = 0;
  WaveWorkspace ws{};
  unsigned int* d_next_query = nullptr;
  int* h_cancel = nullptr; int* d_cancel = nullptr;
  // scratch outputs for host-pointer mode
  float* d_out_dist = nullptr; size_t out_dist_cap = 0;
  uint32_t* d_out_pred = nullptr; float* d_out_dir = nullptr; int32_t* d_out_cut = nullptr;
  uint32_t* d_seed_faces = nullptr; float* d_seed_pos = nullptr; uint32_t seed_cap = 0;
  float* d_face_normals = nullptr; float* d_vertex_normals = nullptr; uint8_t* d_border = nullptr;
  float* d_layer_costs = nullptr; float* d_layer_combined = nullptr; uint8_t* d_layer_mask = nullptr; float* d_clearance = nullptr;
  unsigned int* d_overflow = nullptr;
  // device-resident result of the last single CVP plan (for mnb_cvp_backtrack)
  const uint32_t* last_pred = nullptr; const float* last_dir = nullptr; const int32_t* last_cut = nullptr;
  uint32_t last_seed_face = 0; float last_seed_pos[3] = {0, 0, 0}; bool last_valid = false;
  float* d_path_pos = nullptr; uint32_t* d_path_face = nullptr; int32_t* d_bt_result = nullptr; uint32_t path_cap = 0;
  uint32_t* d_lethals = nullptr; uint32_t lethal_cap = 0; uint8_t* d_infl_invalid = nullptr; float* d_out_cost = nullptr;
  // repulsive vector field of the last inflation (InflationLayer::vector_map_ / distances_)
  bool infl_labels_valid = false, infl_had_invalid = false, infl_field_valid = false, repulsive_on = false;
  mnb_inflation_params infl_params{}; uint64_t infl_rounds = 0;
  float* d_infl_vec = nullptr; float* d_infl_dist = nullptr; int4* d_infl_src = nullptr; unsigned int* d_infl_flag = nullptr;
  // incremental updates
  float* d_prev_risk = nullptr; bool prev_risk_valid = false;     // riskiness map of the previous inflation (NaN = no entry)
  uint32_t* d_upd_ids = nullptr; float* d_upd_costs = nullptr; size_t upd_cap = 0; size_t upd_cost_cap = 0;
  uint32_t* d_changed = nullptr; unsigned int* d_tile_count = nullptr; unsigned int* d_total = nullptr;
  // tuning
  float delta = 0.3f; int cluster = -1 /* -1: whole-grid cooperative kernel for single plans */; int batch_cluster = 1; int threads = 512;
  int grid_blocks_per_sm = 0;
  int layers_smem = 0;         // k_layers with the seen-set / stack in shared memory (MNB_LAYERS_SMEM=1, mnb_debug_set_layers_smem)
  int skip_clean = 0;          // clean-candidate skip of the CVP kernels (band_engine.cuh): bit-identical on the kernel interpreter,
                               // not yet timed on a B200 -> opt-in (MNB_SKIP_CLEAN=1 / mnb_debug_set_skip_clean)
  int sweeps = -1;             // in-round sweeps of the whole-grid single-plan kernel; -1 = derived from the band width
  float grid_delta = 1.8f;     // band width of the whole-grid single-plan kernel (wide band + in-round sweeps)
  float dijkstra_grid_delta = 3.0f;
  mnb_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
      return MNB_E_CUDA;                                                           \
    }                                                                              \
  } while (0)

template <class T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T) > 0 ? n * sizeof(T) : 1); }
template <class T>
static void dfree(T*& p) { if (p) cudaFree(p); p = nullptr; }

static void free_mesh(mnb_ctx* c) {
  dfree(c->d_pos); dfree(c->d_faces); dfree(c->d_edges); dfree(c->d_cor_ptr); dfree(c->d_cor_idx); dfree(c->d_cor_eid);
  dfree(c->d_cor_w); dfree(c->d_cor_wd); dfree(c->d_ell_idx); dfree(c->d_ell_eid); dfree(c->d_ell_w); dfree(c->d_ell_wd); dfree(c->d_ell_geo); dfree(c->d_adj_ptr); dfree(c->d_adj_nbr); dfree(c->d_adj_eid); dfree(c->d_adj_nw); dfree(c->d_ell_adj);
  dfree(c->d_edge_dist); dfree(c->d_edge_w); dfree(c->d_cost); dfree(c->d_invalid);
  dfree(c->ws.state); dfree(c->ws.minor); dfree(c->ws.root); dfree(c->ws.last_eval); dfree(c->ws.dirty); dfree(c->ws.excl); dfree(c->ws.chg); dfree(c->ws.ver); dfree(c->ws.mark); dfree(c->ws.list0); dfree(c->ws.list1); dfree(c->ws.ctl);
  c->ws_groups = 0;
  dfree(c->d_out_dist); c->out_dist_cap = 0; dfree(c->d_out_pred); dfree(c->d_out_dir); dfree(c->d_out_cut);
  dfree(c->d_infl_invalid); dfree(c->d_out_cost);
  dfree(c->d_infl_vec); dfree(c->d_infl_dist); dfree(c->d_infl_src); dfree(c->d_infl_flag);
  c->infl_labels_valid = false; c->infl_field_valid = false; c->repulsive_on = false;
  dfree(c->d_prev_risk); c->prev_risk_valid = false; dfree(c->d_upd_ids); dfree(c->d_upd_costs); c->upd_cap = 0; c->upd_cost_cap = 0;
  dfree(c->d_changed); dfree(c->d_tile_count); dfree(c->d_total);
  dfree(c->d_path_pos); dfree(c->d_path_face); dfree(c->d_bt_result); c->path_cap = 0; c->last_valid = false;
  dfree(c->d_face_normals); dfree(c->d_vertex_normals); dfree(c->d_border); dfree(c->d_layer_costs); dfree(c->d_layer_combined);
  dfree(c->d_layer_mask); dfree(c->d_clearance); dfree(c->d_overflow);
  c->costs_set = false;
}

static int32_t ensure_workspace(mnb_ctx* ctx, uint32_t groups) {
  if (groups <= ctx->ws_groups) return MNB_OK;
  dfree(ctx->ws.state); dfree(ctx->ws.minor); dfree(ctx->ws.root); dfree(ctx->ws.last_eval); dfree(ctx->ws.dirty); dfree(ctx->ws.excl); dfree(ctx->ws.chg); dfree(ctx->ws.ver); dfree(ctx->ws.mark); dfree(ctx->ws.list0); dfree(ctx->ws.list1); dfree(ctx->ws.ctl);
  ctx->ws_groups = 0;
  const size_t n = (size_t)groups * ctx->V;
  CK(dalloc(&ctx->ws.state, n)); CK(dalloc(&ctx->ws.minor, n)); CK(dalloc(&ctx->ws.root, n)); CK(dalloc(&ctx->ws.last_eval, n)); CK(dalloc(&ctx->ws.dirty, n)); CK(dalloc(&ctx->ws.excl, n)); CK(dalloc(&ctx->ws.chg, n)); CK(dalloc(&ctx->ws.ver, (size_t)ctx->V)); CK(dalloc(&ctx->ws.mark, n)); CK(dalloc(&ctx->ws.list0, n)); CK(dalloc(&ctx->ws.list1, n));
  CK(dalloc(&ctx->ws.ctl, groups));
  ctx->ws_groups = groups;
  return MNB_OK;
}

extern "C" {

int32_t mnb_create(int32_t device, mnb_ctx** out_ctx) {
  if (!out_ctx) return MNB_E_ARG;
  *out_ctx = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return MNB_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MNB_E_CUDA;
  if (prop.major < 10) return MNB_E_CUDA;   // sm_100a only; no fallback path exists
  if (cudaSetDevice(device) != cudaSuccess) return MNB_E_CUDA;
  mnb_ctx* c = new mnb_ctx();
  c->device = device; c->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("MNB_LAYERS_SMEM")) c->layers_smem = atoi(e) != 0;                                   // experiment knob
  if (const char* e = getenv("MNB_SKIP_CLEAN")) c->skip_clean = atoi(e) != 0;                                     // experiment knob
  if (const char* e = getenv("MNB_SWEEPS")) { const int k = atoi(e); if (k >= -1 && k <= 64) c->sweeps = k; }   // experiment knob
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MNB_E_CUDA; }
  cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
  cudaMalloc((void**)&c->d_next_query, sizeof(unsigned int));
  if (cudaHostAlloc((void**)&c->h_cancel, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    *c->h_cancel = 0;
    cudaHostGetDevicePointer((void**)&c->d_cancel, c->h_cancel, 0);
  }
  *out_ctx = c;
  return MNB_OK;
}

void mnb_destroy(mnb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_mesh(ctx);
  dfree(ctx->d_next_query); dfree(ctx->d_seed_faces); dfree(ctx->d_seed_pos); dfree(ctx->d_lethals);
  if (ctx->h_cancel) cudaFreeHost(ctx->h_cancel);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* mnb_last_error(mnb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int32_t mnb_set_pointer_mode(mnb_ctx* ctx, int32_t mode) {
  if (!ctx || (mode != MNB_PTR_HOST && mode != MNB_PTR_DEVICE)) return MNB_E_ARG;
  ctx->ptr_mode = mode; return MNB_OK;
}
void* mnb_stream(mnb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint32_t mnb_num_vertices(mnb_ctx* ctx) { return ctx ? ctx->V : 0; }
uint32_t mnb_num_faces(mnb_ctx* ctx) { return ctx ? ctx->F : 0; }
uint32_t mnb_num_edges(mnb_ctx* ctx) { return ctx ? ctx->E : 0; }

int32_t mnb_set_tuning(mnb_ctx* ctx, float band_delta, int32_t cluster_size, int32_t threads_per_cta) {
  if (!ctx) return MNB_E_ARG;
  if (band_delta > 0) { ctx->delta = band_delta; ctx->grid_delta = band_delta; ctx->dijkstra_grid_delta = band_delta; }
  if (cluster_size == 1 || cluster_size == 2 || cluster_size == 4 || cluster_size == 8 || cluster_size == 16) {
    ctx->cluster = cluster_size;
    ctx->batch_cluster = cluster_size > 8 ? 8 : cluster_size;
  } else if (cluster_size == -1) {
    ctx->cluster = -1;        // single plans on the whole grid (cooperative launch); batches keep their cluster size
  } else if (cluster_size != 0) return MNB_E_ARG;
  if (threads_per_cta == 128 || threads_per_cta == 256 || threads_per_cta == 512) ctx->threads = threads_per_cta;
  else if (threads_per_cta != 0) return MNB_E_ARG;
  return MNB_OK;
}

int32_t mnb_get_stats(mnb_ctx* ctx, mnb_stats* out) {
  if (!ctx || !out) return MNB_E_ARG;
  *out = ctx->stats; return MNB_OK;
}

int32_t mnb_set_mesh(mnb_ctx* ctx, uint32_t V, uint32_t F, const float* pos, const uint32_t* faces,
                     const uint32_t* edges, uint32_t E) {
  if (!ctx || !pos || !faces || V == 0 || F == 0) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  free_mesh(ctx);
  try {
    ctx->topo.build(V, F, faces, edges, E);
  } catch (const std::exception& ex) {
    ctx->err = ex.what();
    return MNB_E_ARG;
  }
  HostTopology& T = ctx->topo;
  ctx->V = V; ctx->F = F; ctx->E = T.E; ctx->NC = T.cor_v1.size(); ctx->NA = T.vadj_nbr.size();
  const size_t NC = ctx->NC, NA = ctx->NA;
  CK(dalloc(&ctx->d_pos, 3 * (size_t)V)); CK(dalloc(&ctx->d_faces, 3 * (size_t)F)); CK(dalloc(&ctx->d_edges, 2 * (size_t)T.E));
  CK(dalloc(&ctx->d_cor_ptr, (size_t)V + 1)); CK(dalloc(&ctx->d_cor_idx, NC)); CK(dalloc(&ctx->d_cor_eid, NC));
  CK(dalloc(&ctx->d_cor_w, NC)); CK(dalloc(&ctx->d_cor_wd, NC));
  CK(dalloc(&ctx->d_adj_ptr, (size_t)V + 1)); CK(dalloc(&ctx->d_adj_nbr, NA)); CK(dalloc(&ctx->d_adj_eid, NA)); CK(dalloc(&ctx->d_adj_nw, NA)); CK(dalloc(&ctx->d_ell_adj, (size_t)V * ELL_W));
  CK(dalloc(&ctx->d_edge_dist, (size_t)T.E)); CK(dalloc(&ctx->d_edge_w, (size_t)T.E)); CK(dalloc(&ctx->d_cost, (size_t)V));
  CK(dalloc(&ctx->d_invalid, (size_t)V));
  CK(cudaMemcpyAsync(ctx->d_pos, pos, sizeof(float) * 3 * (size_t)V, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_faces, faces, sizeof(uint32_t) * 3 * (size_t)F, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_edges, T.edges.data(), sizeof(uint32_t) * 2 * (size_t)T.E, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_cor_ptr, T.vcor_ptr.data(), sizeof(uint32_t) * ((size_t)V + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_adj_ptr, T.vadj_ptr.data(), sizeof(uint32_t) * ((size_t)V + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_adj_nbr, T.vadj_nbr.data(), sizeof(uint32_t) * NA, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_adj_eid, T.vadj_eid.data(), sizeof(uint32_t) * NA, cudaMemcpyHostToDevice, ctx->stream));
  {
    std::vector<int4> idx(NC); std::vector<uint4> eid(NC);
    for (size_t k = 0; k < NC; ++k) {
      idx[k] = make_int4((int)T.cor_v1[k], (int)T.cor_v2[k], (int)T.cor_face[k], (int)T.cor_side[k]);   // .w: edge-side bits (topology.hpp)
      eid[k] = make_uint4(T.cor_ec[k], T.cor_eb[k], T.cor_ea[k], 0);
    }
    CK(cudaMemcpyAsync(ctx->d_cor_idx, idx.data(), sizeof(int4) * NC, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_cor_eid, eid.data(), sizeof(uint4) * NC, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // ELL rows: 8 slots per vertex (one 128-byte line), slot 0 carries the degree in .w
    const size_t NE = (size_t)V * ELL_W;
    std::vector<int4> eidx(NE, make_int4(ELL_EMPTY, ELL_EMPTY, -1, 0)); std::vector<uint4> eeid(NE, make_uint4(0, 0, 0, 0));
    for (uint32_t v = 0; v < V; ++v) {
      const uint32_t kb = T.vcor_ptr[v], ke = T.vcor_ptr[v + 1];
      for (uint32_t k = kb; k < ke && k - kb < ELL_W; ++k) { eidx[(size_t)v * ELL_W + (k - kb)] = idx[k]; eeid[(size_t)v * ELL_W + (k - kb)] = eid[k]; }
      eidx[(size_t)v * ELL_W].w = (int)(ke - kb);
    }
    CK(dalloc(&ctx->d_ell_idx, NE)); CK(dalloc(&ctx->d_ell_eid, NE)); CK(dalloc(&ctx->d_ell_w, NE)); CK(dalloc(&ctx->d_ell_wd, NE));
    CK(dalloc(&ctx->d_ell_geo, NE));
    CK(cudaMemcpyAsync(ctx->d_ell_idx, eidx.data(), sizeof(int4) * NE, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ell_eid, eeid.data(), sizeof(uint4) * NE, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  // the host copies of the big per-corner arrays are no longer needed
  std::vector<uint32_t>().swap(T.cor_v1); std::vector<uint32_t>().swap(T.cor_v2); std::vector<uint32_t>().swap(T.cor_face);
  std::vector<uint32_t>().swap(T.cor_ec); std::vector<uint32_t>().swap(T.cor_eb); std::vector<uint32_t>().swap(T.cor_ea);
  std::vector<uint8_t>().swap(T.cor_side);
  std::vector<uint32_t>().swap(T.vadj_nbr); std::vector<uint32_t>().swap(T.vadj_eid); std::vector<uint32_t>().swap(T.face_edges);
  CK(dalloc(&ctx->d_face_normals, 3 * (size_t)F)); CK(dalloc(&ctx->d_vertex_normals, 3 * (size_t)V)); CK(dalloc(&ctx->d_border, (size_t)V));
  CK(cudaMemcpyAsync(ctx->d_border, T.border.data(), (size_t)V, cudaMemcpyHostToDevice, ctx->stream));
  MNB_LAUNCH(k_face_normals, (F + 255) / 256, 256, 0, ctx->stream, ctx->d_pos, ctx->d_faces, F, ctx->d_face_normals);
  MNB_LAUNCH(k_vertex_normals, (V + 255) / 256, 256, 0, ctx->stream, ctx->d_cor_ptr, ctx->d_cor_idx, ctx->d_face_normals, V, ctx->d_vertex_normals);
  MNB_LAUNCH(k_edge_dist, (T.E + 255) / 256, 256, 0, ctx->stream, ctx->d_pos, ctx->d_edges, T.E, ctx->d_edge_dist);
  MNB_LAUNCH(k_gather_corner_w, (unsigned)((NC + 255) / 256), 256, 0, ctx->stream, ctx->d_cor_eid, ctx->d_edge_dist, NC, ctx->d_cor_wd);
  MNB_LAUNCH(k_gather_corner_w, (unsigned)(((size_t)V * ELL_W + 255) / 256), 256, 0, ctx->stream, ctx->d_ell_eid, ctx->d_edge_dist, (size_t)V * ELL_W, ctx->d_ell_wd);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

int32_t mnb_get_edges(mnb_ctx* ctx, uint32_t* out_edges) {
  if (!ctx || !out_edges || !ctx->V) return MNB_E_ARG;
  std::memcpy(out_edges, ctx->topo.edges.data(), sizeof(uint32_t) * 2 * (size_t)ctx->E);
  return MNB_OK;
}

static cudaMemcpyKind in_kind(mnb_ctx* c) { return c->ptr_mode == MNB_PTR_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice; }
static cudaMemcpyKind out_kind(mnb_ctx* c) { return c->ptr_mode == MNB_PTR_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice; }

int32_t mnb_get_edge_distances(mnb_ctx* ctx, float* out) {
  if (!ctx || !out || !ctx->V) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->d_edge_dist, sizeof(float) * (size_t)ctx->E, out_kind(ctx), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

static int32_t install_weights(mnb_ctx* ctx) {
  MNB_LAUNCH(k_gather_corner_w, (unsigned)((ctx->NC + 255) / 256), 256, 0, ctx->stream, ctx->d_cor_eid, ctx->d_edge_w, ctx->NC, ctx->d_cor_w);
  MNB_LAUNCH(k_gather_corner_w, (unsigned)(((size_t)ctx->V * ELL_W + 255) / 256), 256, 0, ctx->stream, ctx->d_ell_eid, ctx->d_edge_w, (size_t)ctx->V * ELL_W, ctx->d_ell_w);
  MNB_LAUNCH(k_corner_geo, (unsigned)(((size_t)ctx->V * ELL_W + 255) / 256), 256, 0, ctx->stream, ctx->d_ell_w, (size_t)ctx->V * ELL_W, ctx->d_ell_geo);
  MNB_LAUNCH(k_gather_adj_w, (unsigned)((ctx->NA + 255) / 256), 256, 0, ctx->stream, ctx->d_adj_nbr, ctx->d_adj_eid, ctx->d_edge_w, ctx->NA, ctx->d_adj_nw);
  MNB_LAUNCH(k_build_ell_adj, (unsigned)(((size_t)ctx->V * ELL_W + 255) / 256), 256, 0, ctx->stream, ctx->d_adj_ptr, ctx->d_adj_nw, ctx->V, ctx->d_ell_adj);
  CK(cudaGetLastError());
  ctx->costs_set = true;
  return MNB_OK;
}

int32_t mnb_compute_edge_weights(mnb_ctx* ctx, const float* vertex_costs, double edge_cost_factor, float* out_w) {
  if (!ctx || !vertex_costs || !ctx->V) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->d_cost, vertex_costs, sizeof(float) * (size_t)ctx->V, in_kind(ctx), ctx->stream));
  MNB_LAUNCH(k_edge_weights, (ctx->E + 255) / 256, 256, 0, ctx->stream, ctx->d_cost, ctx->d_edges, ctx->d_edge_dist, edge_cost_factor, ctx->E, ctx->d_edge_w);
  CK(cudaGetLastError());
  int32_t rc = install_weights(ctx);
  if (rc != MNB_OK) return rc;
  if (out_w) CK(cudaMemcpyAsync(out_w, ctx->d_edge_w, sizeof(float) * (size_t)ctx->E, out_kind(ctx), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

int32_t mnb_set_costs(mnb_ctx* ctx, const float* vertex_costs, const float* edge_weights, const uint8_t* invalid) {
  if (!ctx || !vertex_costs || !edge_weights || !ctx->V) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->d_cost, vertex_costs, sizeof(float) * (size_t)ctx->V, in_kind(ctx), ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_edge_w, edge_weights, sizeof(float) * (size_t)ctx->E, in_kind(ctx), ctx->stream));
  if (invalid) CK(cudaMemcpyAsync(ctx->d_invalid, invalid, (size_t)ctx->V, in_kind(ctx), ctx->stream));
  ctx->has_invalid = invalid != nullptr;
  int32_t rc = install_weights(ctx);
  if (rc != MNB_OK) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

int32_t mnb_cancel(mnb_ctx* ctx) {
  if (!ctx || !ctx->h_cancel) return MNB_E_ARG;
  *(volatile int*)ctx->h_cancel = 1;
  return MNB_OK;
}

}  // extern "C"


static RepulsiveField repulsive_field_of(mnb_ctx* ctx) {
  RepulsiveField L{};
  L.dist = ctx->d_infl_dist; L.vec = ctx->d_infl_vec;
  L.inscribed_radius = ctx->infl_params.inscribed_radius; L.inflation_radius = ctx->infl_params.inflation_radius;
  L.inscribed_radius_f = (float)ctx->infl_params.inscribed_radius;
  L.lethal_value = (float)ctx->infl_params.lethal_value; L.inscribed_value = (float)ctx->infl_params.inscribed_value;
  return L;
}

extern "C" {

int32_t mnb_inflation_vector_map(mnb_ctx* ctx, float* out_vectors) {
  if (!ctx || !ctx->V) return MNB_E_ARG;
  if (!ctx->infl_labels_valid) {
    ctx->err = "mnb_inflation_vector_map needs the labels of the last mnb_inflate / mnb_inflation_update: call it before the next planner call on this context";
    return MNB_E_STATE;
  }
  CK(cudaSetDevice(ctx->device));
  const size_t V = ctx->V;
  if (!ctx->d_infl_vec) { CK(dalloc(&ctx->d_infl_vec, 3 * V)); CK(dalloc(&ctx->d_infl_src, V)); CK(dalloc(&ctx->d_infl_flag, (size_t)2)); }
  InflVecArgs a{};
  a.V = ctx->V; a.pos = ctx->d_pos; a.faces = ctx->d_faces; a.cor_ptr = ctx->d_cor_ptr; a.cor_idx = ctx->d_cor_idx; a.cor_wd = ctx->d_cor_wd;
  a.cor_eid = ctx->d_cor_eid; a.adj_ptr = ctx->d_adj_ptr; a.adj_nbr = ctx->d_adj_nbr; a.invalid = ctx->infl_had_invalid ? ctx->d_infl_invalid : nullptr; a.ws = ctx->ws;
  a.max_distance = (float)ctx->infl_params.inflation_radius; a.vec = ctx->d_infl_vec; a.src = ctx->d_infl_src; a.flag = ctx->d_infl_flag;
  CK(cudaMemsetAsync(ctx->d_infl_flag, 0, 2 * sizeof(unsigned int), ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  MNB_LAUNCH(k_infl_vec_lethal, (ctx->V + 127) / 128, 128, 0, ctx->stream, a);
  MNB_LAUNCH(k_infl_vec_sources, (ctx->V + 127) / 128, 128, 0, ctx->stream, a);
  CK(cudaGetLastError());
  unsigned launches = 2;
  // fixed point over the acyclic source relation: its depth is bounded by the number of rounds the wave took
  const unsigned max_sweeps = (unsigned)ctx->infl_rounds + 8u;
  unsigned int flag[2] = {1u, 0u};
  for (unsigned it = 0; it < max_sweeps && flag[0]; ++it, ++launches) {
    CK(cudaMemsetAsync(ctx->d_infl_flag, 0, sizeof(unsigned int), ctx->stream));
    MNB_LAUNCH(k_infl_vec_sweep, (ctx->V + 255) / 256, 256, 0, ctx->stream, a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(flag, ctx->d_infl_flag, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (flag[1]) { ctx->err = "inflation vector field: a vertex has more than 24 faces with two lethal vertices"; return MNB_E_NOMEM; }
  if (flag[0]) { ctx->err = "inflation vector field did not reach its fixed point"; return MNB_E_STATE; }
  if (out_vectors) CK(cudaMemcpyAsync(out_vectors, ctx->d_infl_vec, sizeof(float) * 3 * V, out_kind(ctx), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const uint64_t wave_rounds = ctx->infl_rounds;
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = launches; ctx->stats.settled = ctx->V; ctx->stats.rounds = wave_rounds;
  ctx->infl_field_valid = true;
  return MNB_OK;
}

int32_t mnb_set_repulsive_field(mnb_ctx* ctx, int32_t enable) {
  if (!ctx) return MNB_E_ARG;
  if (enable && !ctx->infl_field_valid) { ctx->err = "mnb_set_repulsive_field needs mnb_inflation_vector_map first"; return MNB_E_STATE; }
  ctx->repulsive_on = enable != 0;
  return MNB_OK;
}

int32_t mnb_inflation_vector_at(mnb_ctx* ctx, uint32_t n, const uint32_t* faces_q, const float* bary, float* out) {
  if (!ctx || !ctx->V || !faces_q || !bary || !out || n == 0) return MNB_E_ARG;
  if (!ctx->infl_field_valid) { ctx->err = "mnb_inflation_vector_at needs mnb_inflation_vector_map first"; return MNB_E_STATE; }
  for (uint32_t i = 0; ctx->ptr_mode == MNB_PTR_HOST && i < n; ++i) if (faces_q[i] >= ctx->F) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  uint32_t* d_f = nullptr; float* d_b = nullptr; float* d_o = nullptr;
  if (!dev) {
    CK(dalloc(&d_f, (size_t)n)); CK(dalloc(&d_b, 3 * (size_t)n)); CK(dalloc(&d_o, 3 * (size_t)n));
    CK(cudaMemcpyAsync(d_f, faces_q, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_b, bary, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  }
  MNB_LAUNCH(k_inflation_vector_at, (n + 127) / 128, 128, 0, ctx->stream, repulsive_field_of(ctx), (const uint32_t*)ctx->d_faces, n,
             dev ? faces_q : (const uint32_t*)d_f, dev ? bary : (const float*)d_b, dev ? out : d_o);
  CK(cudaGetLastError());
  if (!dev) CK(cudaMemcpyAsync(out, d_o, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  dfree(d_f); dfree(d_b); dfree(d_o);
  return MNB_OK;
}

}  // extern "C"

template <class KArgs>
static cudaError_t launch_cluster(void (*kern)(const KArgs), const KArgs& args, int cs, unsigned blocks, int threads,
                                  cudaStream_t stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = cs > 1 ? 1 : 0;
  if (cs > 8) {
    cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  return cudaLaunchKernelEx(&cfg, kern, args);
}

// cooperative (grid-synchronising) launch of a kernel that takes one argument struct
template <class KArgs>
static cudaError_t launch_cooperative(void (*kern)(const KArgs), const KArgs& args, unsigned blocks, int threads, cudaStream_t stream) {
#ifdef MNB_EMU_ACTIVE
  (void)stream;
  return emu::launch(kern, blocks, (unsigned)threads, (size_t)0, 1u, true, args);
#else
  void* kargs[] = {(void*)&args};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(blocks), dim3(threads), kargs, 0, stream);
#endif
}

static int32_t launch_cvp(mnb_ctx* ctx, const CvpKernelArgs& a, int cs, unsigned groups) {
  cudaError_t e;
  const unsigned blocks = groups * cs;
  const int threads = MNB_CVP_THREADS;
  switch (cs) {
    case 1: e = a.skip_clean ? launch_cluster(k_cvp<1, true>, a, 1, blocks, threads, ctx->stream) : launch_cluster(k_cvp<1, false>, a, 1, blocks, threads, ctx->stream); break;
    case 2: e = launch_cluster(k_cvp<2, false>, a, 2, blocks, threads, ctx->stream); break;     // (the skip variant is built for the two
    case 4: e = launch_cluster(k_cvp<4, false>, a, 4, blocks, threads, ctx->stream); break;     //  default configurations only: per-CTA batches
    case 8: e = launch_cluster(k_cvp<8, false>, a, 8, blocks, threads, ctx->stream); break;     //  and the whole-grid single plan)
    default: e = launch_cluster(k_cvp<16, false>, a, 16, blocks, threads, ctx->stream); break;
  }
  if (e != cudaSuccess) { ctx->err = std::string("cvp launch: ") + cudaGetErrorString(e); return MNB_E_CUDA; }
  return MNB_OK;
}

static int32_t finish_stats(mnb_ctx* ctx, unsigned groups, unsigned launches) {
  std::vector<GroupCtl> h(groups);
  CK(cudaMemcpyAsync(h.data(), ctx->ws.ctl, sizeof(GroupCtl) * groups, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stats.rounds = 0; ctx->stats.recomputes = 0; ctx->stats.settled = 0;
  ctx->stats.skipped = 0;
  for (auto& c : h) { ctx->stats.rounds += c.rounds; ctx->stats.recomputes += c.recomputes; ctx->stats.settled += c.settled; ctx->stats.skipped += c.skipped; }
  ctx->stats.kernel_launches = launches;
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->stats.kernel_ms = ms;
  if (getenv("MNB_PHASE_TIMING")) for (auto& c : h) fprintf(stderr, "[mnb] rounds %llu: CTA0 cycles work %llu flush %llu sync %llu (per round %.0f / %.0f / %.0f) main-pass cycles %llu (unused %llu) chunk-candidates %llu | sweeps: dirty %llu polled %llu poll-cycles %llu eval-cycles %llu (%llu)\n", c.rounds, c.t_work, c.t_flush, c.t_sync, (double)c.t_work / (double)(c.rounds ? c.rounds : 1), (double)c.t_flush / (double)(c.rounds ? c.rounds : 1), (double)c.t_sync / (double)(c.rounds ? c.rounds : 1), c.t_ph[0], c.t_ph[1], c.t_ph[2], c.t_ph[3], c.t_ph[4], c.t_ph[5], c.t_ph[6], c.t_ph[7]);
  for (auto& c : h)
    if (c.watchdog) { ctx->err = "wavefront did not converge within the round watchdog"; return MNB_E_STATE; }
  return MNB_OK;
}

static int32_t ensure_out(mnb_ctx* ctx, size_t n_dist, bool aux) {
  if (n_dist > ctx->out_dist_cap) { dfree(ctx->d_out_dist); CK(dalloc(&ctx->d_out_dist, n_dist)); ctx->out_dist_cap = n_dist; }
  if (aux && !ctx->d_out_pred) {
    CK(dalloc(&ctx->d_out_pred, (size_t)ctx->V)); CK(dalloc(&ctx->d_out_dir, (size_t)ctx->V)); CK(dalloc(&ctx->d_out_cut, (size_t)ctx->V));
  }
  return MNB_OK;
}

static int32_t ensure_seeds(mnb_ctx* ctx, uint32_t n) {
  if (n > ctx->seed_cap) {
    dfree(ctx->d_seed_faces); dfree(ctx->d_seed_pos);
    CK(dalloc(&ctx->d_seed_faces, (size_t)n)); CK(dalloc(&ctx->d_seed_pos, 3 * (size_t)n));
    ctx->seed_cap = n;
  }
  return MNB_OK;
}

static void fill_cvp_args(mnb_ctx* ctx, CvpKernelArgs& a) {
  a.V = ctx->V; a.pos = ctx->d_pos; a.faces = ctx->d_faces; a.cor_ptr = ctx->d_cor_ptr; a.cor_idx = ctx->d_cor_idx;
  a.cor_w = ctx->d_cor_w; a.ell_idx = ctx->d_ell_idx; a.ell_w = ctx->d_ell_w; a.ell_geo = ctx->d_ell_geo; a.cost = ctx->d_cost; a.invalid = ctx->has_invalid ? ctx->d_invalid : nullptr; a.ws = ctx->ws;
  a.seed_faces = ctx->d_seed_faces; a.seed_pos = ctx->d_seed_pos; a.delta = ctx->delta; a.next_query = ctx->d_next_query;
  a.cancel_flag = ctx->d_cancel; a.max_rounds = watchdog_rounds(ctx->V); a.sweeps = 0; a.skip_clean = ctx->skip_clean;
}

extern "C" {

int32_t mnb_cvp(mnb_ctx* ctx, uint32_t seed_face, const float seed_pos[3], int64_t robot_face, double cost_limit,
                double goal_dist_offset, float* out_dist, uint32_t* out_pred, float* out_direction, int32_t* out_cut) {
  if (!ctx || !seed_pos || !ctx->V) return MNB_E_ARG;
  if (!ctx->costs_set) { ctx->err = "mnb_set_costs / mnb_compute_edge_weights not called"; return MNB_E_STATE; }
  if (seed_face >= ctx->F) return MNB_INVALID_START;
  if (robot_face >= (int64_t)ctx->F) return MNB_INVALID_GOAL;
  CK(cudaSetDevice(ctx->device));
  int32_t rc;
  if ((rc = ensure_workspace(ctx, 1)) != MNB_OK) return rc;
  if ((rc = ensure_seeds(ctx, 1)) != MNB_OK) return rc;
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  if ((rc = ensure_out(ctx, dev ? 0 : (size_t)ctx->V, true)) != MNB_OK) return rc;
  if (ctx->h_cancel) *ctx->h_cancel = 0;     // cvp:679 "reset cancel planning"
  ctx->infl_labels_valid = false;            // the wavefront workspace is shared with the inflation wave
  CK(cudaMemcpyAsync(ctx->d_seed_faces, &seed_face, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_seed_pos, seed_pos, 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_next_query, 0, sizeof(unsigned int), ctx->stream));
  CK(cudaMemsetAsync(ctx->ws.ctl, 0, sizeof(GroupCtl), ctx->stream));
  CvpKernelArgs a{};
  fill_cvp_args(ctx, a);
  a.n_queries = 1; a.robot_face = robot_face; a.cost_limit = cost_limit; a.goal_dist_offset = goal_dist_offset;
  a.sweeps = ctx->sweeps;
  a.out_dist = dev ? out_dist : ctx->d_out_dist;
  // aux outputs are always produced for a single plan (the outcome code needs predecessors_)
  a.out_pred = (dev && out_pred) ? out_pred : ctx->d_out_pred;
  a.out_dir = (dev && out_direction) ? out_direction : ctx->d_out_dir;
  a.out_cut = (dev && out_cut) ? out_cut : ctx->d_out_cut;
  if (dev && !out_dist) { if ((rc = ensure_out(ctx, (size_t)ctx->V, true)) != MNB_OK) return rc; a.out_dist = ctx->d_out_dist; }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  if (ctx->cluster == -1) {
    a.delta = ctx->grid_delta;
    if (ctx->grid_blocks_per_sm == 0) {
      int nb = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_cvp_grid<false>, ctx->threads, 0));
      ctx->grid_blocks_per_sm = nb > MNB_GRID_MINBLOCKS ? MNB_GRID_MINBLOCKS : nb;
      if (nb <= 0) { ctx->err = "k_cvp_grid cannot be resident"; return MNB_E_CUDA; }
    }
    if (a.skip_clean) CK(launch_cooperative(k_cvp_grid<true>, a, (unsigned)(ctx->sm_count * ctx->grid_blocks_per_sm), ctx->threads, ctx->stream));
    else CK(launch_cooperative(k_cvp_grid<false>, a, (unsigned)(ctx->sm_count * ctx->grid_blocks_per_sm), ctx->threads, ctx->stream));
  } else {
    if ((rc = launch_cvp(ctx, a, ctx->cluster, 1)) != MNB_OK) return rc;
  }
  MNB_LAUNCH(k_cvp_epilogue, (ctx->V + 255) / 256, 256, 0, ctx->stream, a, ctx->ws.ctl);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) {
    if (out_dist) CK(cudaMemcpyAsync(out_dist, a.out_dist, sizeof(float) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_pred) CK(cudaMemcpyAsync(out_pred, a.out_pred, sizeof(uint32_t) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_direction) CK(cudaMemcpyAsync(out_direction, a.out_dir, sizeof(float) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_cut) CK(cudaMemcpyAsync(out_cut, a.out_cut, sizeof(int32_t) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
  }
  // outcome (cvp:888-918)
  uint32_t rf[3] = {0, 0, 0}, rp[3] = {0, 0, 0};
  if (robot_face >= 0) {
    CK(cudaMemcpyAsync(rf, ctx->d_faces + 3 * (size_t)robot_face, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; ++k)
      CK(cudaMemcpyAsync(&rp[k], a.out_pred + rf[k], sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  ctx->last_valid = false;
  if ((rc = finish_stats(ctx, 1, 2)) != MNB_OK) return rc;
  if (ctx->h_cancel && *ctx->h_cancel) return MNB_CANCELED;
  ctx->last_pred = a.out_pred; ctx->last_dir = a.out_dir; ctx->last_cut = a.out_cut; ctx->last_seed_face = seed_face;
  for (int k = 0; k < 3; ++k) ctx->last_seed_pos[k] = seed_pos[k];
  ctx->last_valid = true;
  if (robot_face >= 0) {
    bool any = false;
    for (int k = 0; k < 3; ++k) if (rp[k] != rf[k]) any = true;
    if (!any && (uint32_t)robot_face != seed_face) return MNB_NO_PATH_FOUND;
  }
  return MNB_SUCCESS;
}

int32_t mnb_cvp_batch(mnb_ctx* ctx, uint32_t n, const uint32_t* seed_faces, const float* seed_pos, double cost_limit,
                      float* out_dist) {
  if (!ctx || !seed_faces || !seed_pos || !out_dist || !ctx->V || n == 0) return MNB_E_ARG;
  if (!ctx->costs_set) { ctx->err = "costs not set"; return MNB_E_STATE; }
  for (uint32_t i = 0; i < n; ++i) if (seed_faces[i] >= ctx->F) return MNB_INVALID_START;
  CK(cudaSetDevice(ctx->device));
  const int cs = ctx->batch_cluster;
  int per_sm = 1;
  if (cs == 1) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cvp<1, false>, MNB_CVP_THREADS, 0)); if (per_sm < 1) per_sm = 1; if (per_sm > 2) per_sm = 2; }
  unsigned groups = (unsigned)(ctx->sm_count * per_sm / cs);
  if (groups > n) groups = n;
  if (groups == 0) groups = 1;
  int32_t rc;
  if ((rc = ensure_workspace(ctx, groups)) != MNB_OK) return rc;
  if ((rc = ensure_seeds(ctx, n)) != MNB_OK) return rc;
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  if ((rc = ensure_out(ctx, dev ? 0 : (size_t)n * ctx->V, false)) != MNB_OK) return rc;
  if (ctx->h_cancel) *ctx->h_cancel = 0;
  ctx->infl_labels_valid = false;
  CK(cudaMemcpyAsync(ctx->d_seed_faces, seed_faces, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_seed_pos, seed_pos, 3 * sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_next_query, 0, sizeof(unsigned int), ctx->stream));
  CK(cudaMemsetAsync(ctx->ws.ctl, 0, sizeof(GroupCtl) * groups, ctx->stream));
  CvpKernelArgs a{};
  fill_cvp_args(ctx, a);
  a.n_queries = n; a.robot_face = -1; a.cost_limit = cost_limit; a.goal_dist_offset = 0.0;
  a.out_dist = dev ? out_dist : ctx->d_out_dist;
  a.out_pred = nullptr; a.out_dir = nullptr; a.out_cut = nullptr;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  if ((rc = launch_cvp(ctx, a, cs, groups)) != MNB_OK) return rc;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) CK(cudaMemcpyAsync(out_dist, a.out_dist, sizeof(float) * (size_t)n * ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = finish_stats(ctx, groups, 1)) != MNB_OK) return rc;
  if (ctx->h_cancel && *ctx->h_cancel) return MNB_CANCELED;
  return MNB_SUCCESS;
}

int32_t mnb_dijkstra(mnb_ctx* ctx, uint32_t seed_vertex, int64_t robot_vertex, double cost_limit, double goal_dist_offset,
                     float* out_dist, uint32_t* out_pred) {
  if (!ctx || !ctx->V) return MNB_E_ARG;
  if (!ctx->costs_set) { ctx->err = "costs not set"; return MNB_E_STATE; }
  if (seed_vertex >= ctx->V) return MNB_INVALID_START;
  if (robot_vertex >= (int64_t)ctx->V) return MNB_INVALID_GOAL;
  CK(cudaSetDevice(ctx->device));
  int32_t rc;
  if ((rc = ensure_workspace(ctx, 1)) != MNB_OK) return rc;
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  if ((rc = ensure_out(ctx, (size_t)ctx->V, true)) != MNB_OK) return rc;
  if (ctx->h_cancel) *ctx->h_cancel = 0;     // dijkstra:238
  ctx->infl_labels_valid = false;
  CK(cudaMemsetAsync(ctx->ws.ctl, 0, sizeof(GroupCtl), ctx->stream));
  DijkstraKernelArgs a{};
  a.V = ctx->V; a.adj_ptr = ctx->d_adj_ptr; a.adj_nw = ctx->d_adj_nw; a.cost = ctx->d_cost;
  a.invalid = ctx->has_invalid ? ctx->d_invalid : nullptr; a.ws = ctx->ws; a.seed_vertex = seed_vertex;
  a.robot_vertex = robot_vertex; a.cost_limit = cost_limit; a.goal_dist_offset = goal_dist_offset; a.delta = ctx->delta;
  a.out_dist = (dev && out_dist) ? out_dist : ctx->d_out_dist;
  a.out_pred = (dev && out_pred) ? out_pred : ctx->d_out_pred;
  a.cancel_flag = ctx->d_cancel; a.max_rounds = watchdog_rounds(ctx->V);
  if (robot_vertex >= 0 && (uint32_t)robot_vertex == seed_vertex) return MNB_SUCCESS;   // dijkstra:252-255
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  cudaError_t e;
  const int cs = ctx->cluster;
  if (cs == -1) {   // single plan on the whole GPU (cooperative launch, one CTA per SM)
    a.delta = ctx->dijkstra_grid_delta; a.ell_adj = ctx->d_ell_adj; a.sweeps = ctx->sweeps;
    e = launch_cooperative(k_dijkstra_grid, a, (unsigned)ctx->sm_count, 512, ctx->stream);
  } else
  switch (cs) {
    case 1: e = launch_cluster(k_dijkstra<1>, a, 1, 1, ctx->threads, ctx->stream); break;
    case 2: e = launch_cluster(k_dijkstra<2>, a, 2, 2, ctx->threads, ctx->stream); break;
    case 4: e = launch_cluster(k_dijkstra<4>, a, 4, 4, ctx->threads, ctx->stream); break;
    case 8: e = launch_cluster(k_dijkstra<8>, a, 8, 8, ctx->threads, ctx->stream); break;
    default: e = launch_cluster(k_dijkstra<16>, a, 16, 16, ctx->threads, ctx->stream); break;
  }
  if (e != cudaSuccess) { ctx->err = std::string("dijkstra launch: ") + cudaGetErrorString(e); return MNB_E_CUDA; }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) {
    if (out_dist) CK(cudaMemcpyAsync(out_dist, a.out_dist, sizeof(float) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_pred) CK(cudaMemcpyAsync(out_pred, a.out_pred, sizeof(uint32_t) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
  }
  uint32_t rp = 0;
  if (robot_vertex >= 0) CK(cudaMemcpyAsync(&rp, a.out_pred + robot_vertex, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = finish_stats(ctx, 1, 1)) != MNB_OK) return rc;
  if (ctx->h_cancel && *ctx->h_cancel) return MNB_CANCELED;
  if (robot_vertex >= 0 && rp == (uint32_t)robot_vertex) return MNB_NO_PATH_FOUND;          // dijkstra:358-362
  return MNB_SUCCESS;
}

int32_t mnb_get_vertex_normals(mnb_ctx* ctx, float* out) {
  if (!ctx || !out || !ctx->V) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->d_vertex_normals, sizeof(float) * 3 * (size_t)ctx->V, out_kind(ctx), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

int32_t mnb_compute_layers(mnb_ctx* ctx, const mnb_layer_params* params, const float* clearance, float* out_costs,
                           float* out_combined, uint8_t* out_lethal_mask) {
  if (!ctx || !params || !ctx->V) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t V = ctx->V;
  if (!ctx->d_layer_costs) {
    CK(dalloc(&ctx->d_layer_costs, 6 * V)); CK(dalloc(&ctx->d_layer_combined, V)); CK(dalloc(&ctx->d_layer_mask, V));
    CK(dalloc(&ctx->d_overflow, (size_t)1));
  }
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  if (clearance) {
    if (!ctx->d_clearance) CK(dalloc(&ctx->d_clearance, V));
    CK(cudaMemcpyAsync(ctx->d_clearance, clearance, sizeof(float) * V, in_kind(ctx), ctx->stream));
  }
  CK(cudaMemsetAsync(ctx->d_overflow, 0, sizeof(unsigned int), ctx->stream));
  LayerKernelArgs a{};
  a.V = ctx->V; a.pos = ctx->d_pos; a.vn = ctx->d_vertex_normals; a.adj_ptr = ctx->d_adj_ptr; a.adj_nbr = ctx->d_adj_nbr;
  a.border = ctx->d_border; a.clearance = clearance ? ctx->d_clearance : nullptr; a.P = *params;
  a.costs = (dev && out_costs) ? out_costs : ctx->d_layer_costs;
  a.combined = (dev && out_combined) ? out_combined : ctx->d_layer_combined;
  a.lethal_mask = (dev && out_lethal_mask) ? out_lethal_mask : ctx->d_layer_mask;
  a.overflow = ctx->d_overflow;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  if (ctx->layers_smem) {
    const size_t smem = sizeof(uint32_t) * (size_t)(NB_HASH + LS_STACK) * LS_THREADS;       // 88 KB: two CTAs per SM
    CK(cudaFuncSetAttribute(k_layers<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNB_LAUNCH(k_layers<true>, (ctx->V + LS_THREADS - 1) / LS_THREADS, LS_THREADS, smem, ctx->stream, a);
  } else {
    MNB_LAUNCH(k_layers<false>, (ctx->V + 127) / 128, 128, 0, ctx->stream, a);
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) {
    if (out_costs) CK(cudaMemcpyAsync(out_costs, a.costs, sizeof(float) * 6 * V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_combined) CK(cudaMemcpyAsync(out_combined, a.combined, sizeof(float) * V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_lethal_mask) CK(cudaMemcpyAsync(out_lethal_mask, a.lethal_mask, V, cudaMemcpyDeviceToHost, ctx->stream));
  }
  unsigned int ovf = 0;
  CK(cudaMemcpyAsync(&ovf, ctx->d_overflow, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = 1; ctx->stats.settled = ctx->V;
  if (ovf) { ctx->err = "layer neighbourhood exceeds the per-vertex scratch (radius too large for the mesh resolution)"; return MNB_E_NOMEM; }
  return MNB_OK;
}

int32_t mnb_vector_map(mnb_ctx* ctx, const uint32_t* pred, const float* direction, const int32_t* cutting_face, float* out_vec) {
  if (!ctx || !ctx->V || !pred || !out_vec) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const size_t V = ctx->V;
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  const uint32_t* d_pred = pred; const float* d_dir = direction; const int32_t* d_cut = cutting_face; float* d_out = out_vec;
  float* tmp_out = nullptr; uint32_t* tmp_pred = nullptr; float* tmp_dir = nullptr; int32_t* tmp_cut = nullptr;
  if (!dev) {
    CK(dalloc(&tmp_out, 3 * V)); CK(dalloc(&tmp_pred, V));
    CK(cudaMemcpyAsync(tmp_pred, pred, sizeof(uint32_t) * V, cudaMemcpyHostToDevice, ctx->stream));
    d_pred = tmp_pred; d_out = tmp_out;
    if (direction) { CK(dalloc(&tmp_dir, V)); CK(cudaMemcpyAsync(tmp_dir, direction, sizeof(float) * V, cudaMemcpyHostToDevice, ctx->stream)); d_dir = tmp_dir; }
    if (cutting_face) { CK(dalloc(&tmp_cut, V)); CK(cudaMemcpyAsync(tmp_cut, cutting_face, sizeof(int32_t) * V, cudaMemcpyHostToDevice, ctx->stream)); d_cut = tmp_cut; }
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  MNB_LAUNCH(k_vector_map, (ctx->V + 255) / 256, 256, 0, ctx->stream, ctx->d_pos, ctx->d_vertex_normals, d_pred, d_dir, d_cut, ctx->V, d_out);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) CK(cudaMemcpyAsync(out_vec, d_out, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  dfree(tmp_out); dfree(tmp_pred); dfree(tmp_dir); dfree(tmp_cut);
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = 1; ctx->stats.settled = ctx->V;
  return MNB_OK;
}

int32_t mnb_cvp_backtrack(mnb_ctx* ctx, const float robot_pos[3], uint32_t robot_face, double step_width, uint32_t max_points,
                          float* path_pos, uint32_t* path_face, uint32_t* n_points) {
  if (!ctx || !ctx->V || !robot_pos || !path_pos || !n_points || max_points < 2) return MNB_E_ARG;
  if (!ctx->last_valid) { ctx->err = "mnb_cvp_backtrack needs a preceding successful mnb_cvp on this context"; return MNB_E_STATE; }
  if (robot_face >= ctx->F) return MNB_INVALID_GOAL;
  CK(cudaSetDevice(ctx->device));
  if (max_points > ctx->path_cap) {
    dfree(ctx->d_path_pos); dfree(ctx->d_path_face); ctx->path_cap = 0;
    CK(dalloc(&ctx->d_path_pos, 3 * (size_t)max_points)); CK(dalloc(&ctx->d_path_face, (size_t)max_points));
    ctx->path_cap = max_points;
  }
  if (!ctx->d_bt_result) CK(dalloc(&ctx->d_bt_result, (size_t)2));
  BacktrackArgs a{};
  a.pos = ctx->d_pos; a.vn = ctx->d_vertex_normals; a.faces = ctx->d_faces; a.cor_ptr = ctx->d_cor_ptr; a.cor_idx = ctx->d_cor_idx;
  a.pred = ctx->last_pred; a.direction = ctx->last_dir; a.cut = ctx->last_cut;
  for (int k = 0; k < 3; ++k) { a.start[k] = ctx->last_seed_pos[k]; a.goal[k] = robot_pos[k]; }
  a.start_face = ctx->last_seed_face; a.goal_face = robot_face; a.step_width = step_width; a.max_points = max_points;
  a.path_pos = ctx->d_path_pos; a.path_face = ctx->d_path_face; a.result = ctx->d_bt_result; a.cancel_flag = ctx->d_cancel;
  a.layer = RepulsiveField{};
  if (ctx->repulsive_on && ctx->infl_field_valid) a.layer = repulsive_field_of(ctx);
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(cudaFuncSetAttribute(k_backtrack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BtShared)));
  MNB_LAUNCH(k_backtrack, 1, 32, sizeof(BtShared), ctx->stream, a);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  int32_t res[2] = {0, 0};
  CK(cudaMemcpyAsync(res, ctx->d_bt_result, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const uint32_t n = (uint32_t)res[1] < max_points ? (uint32_t)res[1] : max_points;
  *n_points = n;
  const cudaMemcpyKind kind = ctx->ptr_mode == MNB_PTR_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CK(cudaMemcpyAsync(path_pos, ctx->d_path_pos, sizeof(float) * 3 * (size_t)n, kind, ctx->stream));
  if (path_face) CK(cudaMemcpyAsync(path_face, ctx->d_path_face, sizeof(uint32_t) * (size_t)n, kind, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = 1; ctx->stats.settled = n;
  if (res[0] == MNB_E_STATE) { ctx->err = "back-tracking exceeded max_points (cyclic vector field?) or the face search list"; return MNB_E_STATE; }
  return res[0];
}

int32_t mnb_locate(mnb_ctx* ctx, uint32_t n, const float* points, uint32_t* out_vertex, int32_t* out_face, float* out_bary) {
  if (!ctx || !ctx->V || !points || n == 0) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  float* d_pts = nullptr; unsigned long long* d_keys = nullptr; uint32_t* d_v = nullptr; int32_t* d_f = nullptr; float* d_b = nullptr;
  CK(dalloc(&d_keys, (size_t)n));
  CK(cudaMemsetAsync(d_keys, 0xff, sizeof(unsigned long long) * (size_t)n, ctx->stream));
  const float* pts = points;
  if (!dev) {
    CK(dalloc(&d_pts, 3 * (size_t)n)); CK(dalloc(&d_v, (size_t)n)); CK(dalloc(&d_f, (size_t)n)); CK(dalloc(&d_b, 3 * (size_t)n));
    CK(cudaMemcpyAsync(d_pts, points, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    pts = d_pts;
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  const uint32_t want = (ctx->V + 255) / 256, cap = (uint32_t)ctx->sm_count * 8;
  const uint32_t blocks = want < cap ? want : cap;
  uint32_t launches = 0;
  for (uint32_t q0 = 0; q0 < n; q0 += LOC_Q, ++launches)
    MNB_LAUNCH(k_nearest_vertex, blocks, 256, 0, ctx->stream, ctx->d_pos, ctx->V, pts, q0, n - q0 < (uint32_t)LOC_Q ? n - q0 : (uint32_t)LOC_Q, d_keys);
  MNB_LAUNCH(k_containing_face, (n + 127) / 128, 128, 0, ctx->stream, ctx->d_pos, ctx->d_faces, ctx->d_cor_ptr, ctx->d_cor_idx, pts, n, d_keys,
                                                           dev ? out_vertex : d_v, dev ? out_face : d_f, dev ? out_bary : d_b);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) {
    if (out_vertex) CK(cudaMemcpyAsync(out_vertex, d_v, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_face) CK(cudaMemcpyAsync(out_face, d_f, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_bary) CK(cudaMemcpyAsync(out_bary, d_b, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  dfree(d_pts); dfree(d_keys); dfree(d_v); dfree(d_f); dfree(d_b);
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = launches + 1; ctx->stats.settled = n;
  return MNB_OK;
}

// experiment knob (not part of the public header): in-round sweeps of the whole-grid single-plan kernel
int32_t mnb_debug_set_sweeps(mnb_ctx* ctx, int32_t k) { if (!ctx || k < -1 || k > 64) return MNB_E_ARG; ctx->sweeps = k; return MNB_OK; }

int32_t mnb_debug_set_layers_smem(mnb_ctx* ctx, int32_t on) { if (!ctx) return MNB_E_ARG; ctx->layers_smem = on != 0; return MNB_OK; }
int32_t mnb_debug_set_skip_clean(mnb_ctx* ctx, int32_t on) { if (!ctx) return MNB_E_ARG; ctx->skip_clean = on != 0; return MNB_OK; }

// debugging aid (not part of the public header): raw labels {d, a1, a2, a3|flag} of wavefront group 0
int32_t mnb_debug_get_labels(mnb_ctx* ctx, uint32_t* out4v) {
  if (!ctx || !ctx->ws.state) return MNB_E_ARG;
  CK(cudaMemcpy(out4v, ctx->ws.state, sizeof(uint4) * (size_t)ctx->V, cudaMemcpyDeviceToHost));
  return MNB_OK;
}

}  // extern "C"

// InflationLayer::waveCostInflation; with want_update additionally the update set of InflationLayer::onInputChanged
static int32_t inflate_impl(mnb_ctx* ctx, const uint32_t* lethals, uint32_t n, const uint8_t* invalid,
                            const mnb_inflation_params* params, float* out_dist, float* out_cost, bool want_update,
                            uint32_t* out_changed, uint32_t* n_changed) {
  if (!ctx || !ctx->V || !params || (n && !lethals)) return MNB_E_ARG;
  CK(cudaSetDevice(ctx->device));
  int32_t rc;
  if ((rc = ensure_workspace(ctx, 1)) != MNB_OK) return rc;
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  if ((rc = ensure_out(ctx, (size_t)ctx->V, false)) != MNB_OK) return rc;
  if (!ctx->d_out_cost) CK(dalloc(&ctx->d_out_cost, (size_t)ctx->V));
  if (n > ctx->lethal_cap) { dfree(ctx->d_lethals); CK(dalloc(&ctx->d_lethals, (size_t)n)); ctx->lethal_cap = n; }
  if (n) CK(cudaMemcpyAsync(ctx->d_lethals, lethals, sizeof(uint32_t) * n, in_kind(ctx), ctx->stream));
  if (invalid) {
    if (!ctx->d_infl_invalid) CK(dalloc(&ctx->d_infl_invalid, (size_t)ctx->V));
    CK(cudaMemcpyAsync(ctx->d_infl_invalid, invalid, (size_t)ctx->V, in_kind(ctx), ctx->stream));
  }
  CK(cudaMemsetAsync(ctx->ws.ctl, 0, sizeof(GroupCtl), ctx->stream));
  InflateKernelArgs a{};
  a.V = ctx->V; a.cor_ptr = ctx->d_cor_ptr; a.cor_idx = ctx->d_cor_idx; a.cor_wd = ctx->d_cor_wd; a.cor_eid = ctx->d_cor_eid;
  a.invalid = invalid ? ctx->d_infl_invalid : nullptr; a.ws = ctx->ws; a.lethals = ctx->d_lethals; a.n_lethals = n;
  a.max_distance = (float)params->inflation_radius;      // double -> `const float&` parameter (inflation_layer.cpp:240,450)
  a.params.inscribed_radius = params->inscribed_radius; a.params.inflation_radius = params->inflation_radius;
  a.params.lethal_value = params->lethal_value; a.params.inscribed_value = params->inscribed_value;
  a.params.cost_scaling_factor = params->cost_scaling_factor;
  a.out_dist = (dev && out_dist) ? out_dist : ctx->d_out_dist;
  a.out_cost = (dev && out_cost) ? out_cost : ctx->d_out_cost;
  a.max_rounds = watchdog_rounds(ctx->V);
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(launch_cooperative(k_inflate, a, (unsigned)ctx->sm_count, ctx->threads, ctx->stream));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!dev) {
    if (out_dist) CK(cudaMemcpyAsync(out_dist, a.out_dist, sizeof(float) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_cost) CK(cudaMemcpyAsync(out_cost, a.out_cost, sizeof(float) * (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
  }
  int32_t rc2 = finish_stats(ctx, 1, 1);
  if (rc2 != MNB_OK) return rc2;
  // the riskiness map of this run is the "previous" one of the next mnb_inflation_update (riskiness_ = std::move(new_costs))
  const size_t V = ctx->V;
  if (want_update) {
    const uint32_t n_tiles = (uint32_t)((V + US_TILE - 1) / US_TILE);
    if (!ctx->d_changed) { CK(dalloc(&ctx->d_changed, V)); CK(dalloc(&ctx->d_tile_count, (size_t)n_tiles)); CK(dalloc(&ctx->d_total, (size_t)1)); }
    const float* old = ctx->prev_risk_valid ? ctx->d_prev_risk : nullptr;
    uint32_t* d_out = (dev && out_changed) ? out_changed : ctx->d_changed;
    MNB_LAUNCH(k_update_set_count, n_tiles, 256, 0, ctx->stream, (const float*)a.out_cost, old, ctx->V, ctx->d_tile_count);
    MNB_LAUNCH(k_update_set_scan, 1, 1024, 0, ctx->stream, ctx->d_tile_count, n_tiles, ctx->d_total);
    MNB_LAUNCH(k_update_set_write, n_tiles, 256, 0, ctx->stream, (const float*)a.out_cost, old, ctx->V, (const unsigned int*)ctx->d_tile_count, d_out);
    CK(cudaGetLastError());
    unsigned int total = 0;
    CK(cudaMemcpyAsync(&total, ctx->d_total, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_changed) *n_changed = total;
    if (!dev && out_changed && total) CK(cudaMemcpyAsync(out_changed, d_out, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.kernel_launches += 3;
  }
  if (!ctx->d_prev_risk) CK(dalloc(&ctx->d_prev_risk, V));
  CK(cudaMemcpyAsync(ctx->d_prev_risk, a.out_cost, sizeof(float) * V, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->prev_risk_valid = true;
  ctx->infl_labels_valid = true; ctx->infl_had_invalid = invalid != nullptr; ctx->infl_params = *params; ctx->infl_field_valid = false; ctx->infl_rounds = ctx->stats.rounds;
  if (!ctx->d_infl_dist) CK(dalloc(&ctx->d_infl_dist, V));
  CK(cudaMemcpyAsync(ctx->d_infl_dist, a.out_dist, sizeof(float) * V, cudaMemcpyDeviceToDevice, ctx->stream));   // distances_
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

extern "C" {

int32_t mnb_inflate(mnb_ctx* ctx, const uint32_t* lethals, uint32_t n, const uint8_t* invalid,
                    const mnb_inflation_params* params, float* out_dist, float* out_cost) {
  return inflate_impl(ctx, lethals, n, invalid, params, out_dist, out_cost, false, nullptr, nullptr);
}

int32_t mnb_inflation_update(mnb_ctx* ctx, const uint32_t* lethals, uint32_t n, const uint8_t* invalid,
                             const mnb_inflation_params* params, float* out_dist, float* out_cost, uint32_t* out_changed,
                             uint32_t* n_changed) {
  if (!n_changed) return MNB_E_ARG;
  return inflate_impl(ctx, lethals, n, invalid, params, out_dist, out_cost, true, out_changed, n_changed);
}

int32_t mnb_get_costs(mnb_ctx* ctx, float* out_vertex_costs, float* out_edge_weights) {
  if (!ctx || !ctx->V) return MNB_E_ARG;
  if (!ctx->costs_set) { ctx->err = "costs not set"; return MNB_E_STATE; }
  CK(cudaSetDevice(ctx->device));
  if (out_vertex_costs) CK(cudaMemcpyAsync(out_vertex_costs, ctx->d_cost, sizeof(float) * (size_t)ctx->V, out_kind(ctx), ctx->stream));
  if (out_edge_weights) CK(cudaMemcpyAsync(out_edge_weights, ctx->d_edge_w, sizeof(float) * (size_t)ctx->E, out_kind(ctx), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MNB_OK;
}

int32_t mnb_update_vertex_costs(mnb_ctx* ctx, uint32_t n_changed, const uint32_t* changed, const float* costs,
                                int32_t costs_indexed_by_vertex, float default_value, double edge_cost_factor) {
  if (!ctx || !ctx->V || (n_changed && (!changed || !costs))) return MNB_E_ARG;
  if (!ctx->costs_set) { ctx->err = "mnb_update_vertex_costs needs mnb_set_costs / mnb_compute_edge_weights first"; return MNB_E_STATE; }
  if (n_changed == 0) return MNB_OK;
  CK(cudaSetDevice(ctx->device));
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  const uint32_t* d_ids = changed; const float* d_costs = costs;
  if (!dev) {
    const size_t nc = costs_indexed_by_vertex ? (size_t)ctx->V : (size_t)n_changed;
    if (n_changed > ctx->upd_cap) { dfree(ctx->d_upd_ids); ctx->upd_cap = 0; CK(dalloc(&ctx->d_upd_ids, (size_t)n_changed)); ctx->upd_cap = n_changed; }
    if (nc > ctx->upd_cost_cap) { dfree(ctx->d_upd_costs); ctx->upd_cost_cap = 0; CK(dalloc(&ctx->d_upd_costs, nc)); ctx->upd_cost_cap = nc; }
    CK(cudaMemcpyAsync(ctx->d_upd_ids, changed, sizeof(uint32_t) * (size_t)n_changed, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_upd_costs, costs, sizeof(float) * nc, cudaMemcpyHostToDevice, ctx->stream));
    d_ids = ctx->d_upd_ids; d_costs = ctx->d_upd_costs;
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  unsigned launches = 1;
  MNB_LAUNCH(k_update_costs, (n_changed + 255) / 256, 256, 0, ctx->stream, d_ids, n_changed, d_costs, (int)(costs_indexed_by_vertex != 0),
             default_value, ctx->V, ctx->d_cost);
  if (edge_cost_factor != 0) {                       // mesh_map.cpp:568-572: no edge update at all for a zero factor
    const unsigned blocks = (unsigned)(((size_t)n_changed * ELL_W + 255) / 256);
    MNB_LAUNCH(k_update_edge_weights, blocks, 256, 0, ctx->stream, d_ids, n_changed, ctx->V, (const uint32_t*)ctx->d_adj_ptr,
               (const uint32_t*)ctx->d_adj_eid, (const uint32_t*)ctx->d_edges, (const float*)ctx->d_cost, (const float*)ctx->d_edge_dist,
               edge_cost_factor, ctx->d_edge_w);
    RefreshArgs r{};
    r.changed = d_ids; r.n = n_changed; r.V = ctx->V; r.faces = ctx->d_faces; r.cor_ptr = ctx->d_cor_ptr; r.cor_idx = ctx->d_cor_idx;
    r.cor_eid = ctx->d_cor_eid; r.adj_ptr = ctx->d_adj_ptr; r.adj_nbr = ctx->d_adj_nbr; r.adj_eid = ctx->d_adj_eid; r.w = ctx->d_edge_w;
    r.cor_w = ctx->d_cor_w; r.ell_w = ctx->d_ell_w; r.ell_geo = ctx->d_ell_geo; r.adj_nw = ctx->d_adj_nw; r.ell_adj = ctx->d_ell_adj;
    MNB_LAUNCH(k_refresh_weight_tables, blocks, 256, 0, ctx->stream, r);
    launches = 3;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = launches; ctx->stats.settled = n_changed;
  ctx->last_valid = false;                           // the device-resident plan no longer belongs to the installed costs
  return MNB_OK;
}

int32_t mnb_max_combination_update(mnb_ctx* ctx, uint32_t n_layers, const float* const* layer_costs, const float* defaults,
                                   const uint8_t* const* layer_lethal, uint32_t n_changed, const uint32_t* changed,
                                   float* io_costs, uint8_t* io_lethal) {
  if (!ctx || !ctx->V || n_layers == 0 || n_layers > (uint32_t)COMB_MAX_LAYERS || !layer_costs || !defaults || !io_costs ||
      (n_changed && !changed)) return MNB_E_ARG;
  for (uint32_t l = 0; l < n_layers; ++l) if (!layer_costs[l]) return MNB_E_ARG;
  if (n_changed == 0) return MNB_OK;
  CK(cudaSetDevice(ctx->device));
  const bool dev = ctx->ptr_mode == MNB_PTR_DEVICE;
  const size_t V = ctx->V;
  CombineArgs a{};
  a.n_layers = n_layers; a.n = n_changed; a.V = ctx->V;
  std::vector<void*> tmp;                             // host-pointer mode: device copies of the maps
  auto cleanup = [&]() { for (void* q : tmp) cudaFree(q); };
  auto up = [&](const void* h, size_t bytes, void** d) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, bytes ? bytes : 1); if (e != cudaSuccess) return e;
    tmp.push_back(*d);
    return cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream);
  };
  cudaError_t e = cudaSuccess;
  for (uint32_t l = 0; l < n_layers && e == cudaSuccess; ++l) {
    a.def[l] = defaults[l];
    if (dev) { a.costs[l] = layer_costs[l]; a.lethal[l] = layer_lethal ? layer_lethal[l] : nullptr; continue; }
    void* d = nullptr;
    e = up(layer_costs[l], sizeof(float) * V, &d); a.costs[l] = (const float*)d;
    if (e == cudaSuccess && layer_lethal && layer_lethal[l]) { e = up(layer_lethal[l], V, &d); a.lethal[l] = (const uint8_t*)d; }
  }
  void* d_ids = nullptr; void* d_io = nullptr; void* d_il = nullptr;
  if (!dev && e == cudaSuccess) {
    e = up(changed, sizeof(uint32_t) * (size_t)n_changed, &d_ids);
    if (e == cudaSuccess) e = up(io_costs, sizeof(float) * V, &d_io);
    if (e == cudaSuccess && io_lethal) e = up(io_lethal, V, &d_il);
  }
  if (e != cudaSuccess) { cleanup(); ctx->err = std::string("mnb_max_combination_update: ") + cudaGetErrorString(e); return MNB_E_CUDA; }
  a.changed = dev ? changed : (const uint32_t*)d_ids;
  a.io_costs = dev ? io_costs : (float*)d_io;
  a.io_lethal = dev ? io_lethal : (uint8_t*)d_il;
  cudaEventRecord(ctx->ev0, ctx->stream);
  MNB_LAUNCH(k_max_combination_update, (n_changed + 255) / 256, 256, 0, ctx->stream, a);
  e = cudaGetLastError();
  cudaEventRecord(ctx->ev1, ctx->stream);
  if (e == cudaSuccess && !dev) {
    e = cudaMemcpyAsync(io_costs, a.io_costs, sizeof(float) * V, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && io_lethal) e = cudaMemcpyAsync(io_lethal, a.io_lethal, V, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) { ctx->err = std::string("mnb_max_combination_update: ") + cudaGetErrorString(e); return MNB_E_CUDA; }
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats = mnb_stats{}; ctx->stats.kernel_ms = ms; ctx->stats.kernel_launches = 1; ctx->stats.settled = n_changed;
  return MNB_OK;
}

}  // extern "C"
