// Per-algorithm "pull" recompute rules plugged into the band engine.
#pragma once
#include "band_engine.cuh"

namespace mnb {

// ---------------------------------------------------------------------------
// CVP: vertex c is recomputed from its corner records (one per incident face):
//   cor_idx[k] = {v1, v2, face, -}   v1 = next(c), v2 = next(v1) in the face's cyclic order
//   cor_w[k]   = {w(v1,v2), w(v1,c), w(v2,c), -}   edge *weights* (cvp:380-390)
// Reference: the face (v1,v2,c) updates c when the later of v1,v2 is popped and c
// is the only non-fixed vertex (cvp_mesh_planner.cpp:790-866).
// ---------------------------------------------------------------------------
struct CvpProblem {
  const uint32_t* __restrict__ cor_ptr;
  const int4* __restrict__ cor_idx;
  const float4* __restrict__ cor_w;
  const float* __restrict__ cost;
  const uint8_t* __restrict__ invalid;  // may be null
  unsigned long long* state;
  uint32_t* pred;    // may be null (batch: potentials only)
  float* dir;
  int32_t* cut;
  double cost_limit;
  uint32_t s0, s1, s2;       // seed vertices (pre-fixed, cvp:719-728)
  uint32_t seed_noexpand;    // bit k: seed k pops but does not expand (cvp:757,760)

  static constexpr int MAXF = 12;

  __device__ __forceinline__ unsigned long long load_state(uint32_t v) const { return __ldcg(&state[v]); }
  __device__ __forceinline__ bool eligible(uint32_t x) const {
    if (invalid && invalid[x]) return false;        // cvp:785 (no face with an invalid vertex)
    return !((double)cost[x] >= cost_limit);        // cvp:802,825,848
  }
  __device__ __forceinline__ int seed_index(uint32_t v) const { return v == s0 ? 0 : (v == s1 ? 1 : (v == s2 ? 2 : -1)); }

  template <class F>
  __device__ __forceinline__ void activate(uint32_t c, F push) const {
    const uint32_t kb = cor_ptr[c], ke = cor_ptr[c + 1];
    for (uint32_t k = kb; k < ke; ++k) {
      const int4 ix = __ldg(&cor_idx[k]);
      push((uint32_t)ix.x);
      push((uint32_t)ix.y);
    }
  }

  // event time of corner k for candidate c; returns false if the face cannot fire
  __device__ __forceinline__ bool corner_time(uint32_t k, float band_end, float goal, float& T, uint32_t& Tid,
                                              float& u1, float& u2) const {
    const int4 ix = __ldg(&cor_idx[k]);
    const uint32_t v1 = (uint32_t)ix.x, v2 = (uint32_t)ix.y;
    const unsigned long long a = load_state(v1), b = load_state(v2);
    u1 = state_d(a); u2 = state_d(b);
    if (!(u1 < band_end) || !(u2 < band_end)) return false;
    if (invalid && (invalid[v1] || invalid[v2])) return false;
    const float t1 = state_tau(a), t2 = state_tau(b);
    const int i1 = seed_index(v1), i2 = seed_index(v2);
    const bool v1_later = t1 > t2 || (t1 == t2 && v1 > v2);
    if (i1 >= 0 && i2 >= 0) {
      // both sources pre-fixed: the face fires at the FIRST of them that pops and expands
      const bool e1 = !((seed_noexpand >> i1) & 1u), e2 = !((seed_noexpand >> i2) & 1u);
      if (!e1 && !e2) return false;
      const bool use1 = e1 && (!e2 || !v1_later);
      T = use1 ? t1 : t2; Tid = use1 ? v1 : v2;
      return true;
    }
    const uint32_t later = v1_later ? v1 : v2;
    const int il = v1_later ? i1 : i2;
    if (il >= 0 && ((seed_noexpand >> il) & 1u)) return false;
    const float dl = v1_later ? u1 : u2;
    if (dl > goal) return false;                       // cvp:754
    T = v1_later ? t1 : t2; Tid = later;
    return true;
  }

  __device__ __forceinline__ void write_result(uint32_t c, float nd, float ntau, int win, const CvpResult& best) {
    __stcg(&state[c], pack_state(nd, ntau));
    if (pred) {
      if (win >= 0) {
        const int4 ix = __ldg(&cor_idx[win]);
        pred[c] = best.pred_sel == 1 ? (uint32_t)ix.x : (uint32_t)ix.y;
        dir[c] = best.direction;
        cut[c] = ix.z;
      } else {
        pred[c] = c; dir[c] = 0.0f; cut[c] = -1;
      }
    }
  }

  // generic path for vertices with more than MAXF incident faces: repeated selection of the next
  // corner in (T, Tid, corner index) order by rescanning the corner list (O(deg^2), rare)
  __device__ __noinline__ void recompute_big(uint32_t c, float band_end, float goal, float& nd, float& ntau, int& win,
                                             CvpResult& best) const {
    const uint32_t kb = cor_ptr[c], ke = cor_ptr[c + 1];
    float cur = __uint_as_float(INF_BITS), tcur = cur;
    float lastT = 0.0f; uint32_t lastId = 0, lastK = 0; bool have_last = false;
    win = -1;
    for (;;) {
      float bT = 0, bu1 = 0, bu2 = 0; uint32_t bId = 0, bk = 0; bool found = false;
      for (uint32_t k = kb; k < ke; ++k) {
        float T, u1, u2; uint32_t Tid;
        if (!corner_time(k, band_end, goal, T, Tid, u1, u2)) continue;
        if (have_last) {
          const bool after = T > lastT || (T == lastT && (Tid > lastId || (Tid == lastId && k > lastK)));
          if (!after) continue;
        }
        if (!found || T < bT || (T == bT && (Tid < bId || (Tid == bId && k < bk)))) {
          bT = T; bId = Tid; bk = k; bu1 = u1; bu2 = u2; found = true;
        }
      }
      if (!found) break;
      if (!(bT < tcur || (bT == tcur && bId < c))) break;
      const float4 w = __ldg(&cor_w[bk]);
      CvpResult r;
      if (cvp_update(bu1, bu2, cur, w.z, w.y, w.x, r)) { cur = r.value; tcur = fmaxf(cur, bT); best = r; win = (int)bk; }
      lastT = bT; lastId = bId; lastK = bk; have_last = true;
    }
    nd = cur; ntau = tcur;
  }

  __device__ __forceinline__ bool recompute(uint32_t c, float band_end, float goal, float d_old, float tau_old,
                                            float& nd, float& ntau) {
    const uint32_t kb = cor_ptr[c], ke = cor_ptr[c + 1];
    int win = -1;
    CvpResult best; best.value = 0; best.direction = 0; best.pred_sel = 1;
    if (ke - kb > (uint32_t)MAXF) {
      recompute_big(c, band_end, goal, nd, ntau, win, best);
    } else {
      float Tt[MAXF], U1[MAXF], U2[MAXF]; uint32_t Ti[MAXF], K[MAXF];
      int n = 0;
      for (uint32_t k = kb; k < ke; ++k) {
        float T, u1, u2; uint32_t Tid;
        if (!corner_time(k, band_end, goal, T, Tid, u1, u2)) continue;
        Tt[n] = T; Ti[n] = Tid; U1[n] = u1; U2[n] = u2; K[n] = k; ++n;
      }
      float cur = __uint_as_float(INF_BITS), tcur = cur;
      for (int i = 0; i < n; ++i) {
        int b = i;
        for (int j = i + 1; j < n; ++j)
          if (Tt[j] < Tt[b] || (Tt[j] == Tt[b] && (Ti[j] < Ti[b] || (Ti[j] == Ti[b] && K[j] < K[b])))) b = j;
        const float T = Tt[b]; const uint32_t Tid = Ti[b]; const float u1 = U1[b], u2 = U2[b]; const uint32_t k = K[b];
        Tt[b] = Tt[i]; Ti[b] = Ti[i]; U1[b] = U1[i]; U2[b] = U2[i]; K[b] = K[i];
        if (!(T < tcur || (T == tcur && Tid < c))) break;   // c has been popped before this face fires
        const float4 w = __ldg(&cor_w[k]);
        CvpResult r;
        if (cvp_update(u1, u2, cur, w.z, w.y, w.x, r)) { cur = r.value; tcur = fmaxf(cur, T); best = r; win = (int)k; }
      }
      nd = cur; ntau = tcur;
    }
    if (__float_as_uint(nd) == __float_as_uint(d_old) && __float_as_uint(ntau) == __float_as_uint(tau_old)) return false;
    write_result(c, nd, ntau, win, best);
    return true;
  }
};

// ---------------------------------------------------------------------------
// Dijkstra: d[c] = min over expandable neighbours u of fl(d[u] + w(u,c));
// among equal sums the neighbour that pops first wins (strict '<' at
// dijkstra_mesh_planner.cpp:332): order (d[u], u).
//   adj_nw[k] = {neighbour id, float bits of the edge weight}
// ---------------------------------------------------------------------------
struct DijkstraProblem {
  const uint32_t* __restrict__ adj_ptr;
  const uint2* __restrict__ adj_nw;
  const float* __restrict__ cost;
  const uint8_t* __restrict__ invalid;  // may be null
  unsigned long long* state;
  uint32_t* pred;
  double cost_limit;

  __device__ __forceinline__ unsigned long long load_state(uint32_t v) const { return __ldcg(&state[v]); }
  __device__ __forceinline__ bool eligible(uint32_t x) const { return !(invalid && invalid[x]); }  // :328

  template <class F>
  __device__ __forceinline__ void activate(uint32_t c, F push) const {
    const uint32_t kb = adj_ptr[c], ke = adj_ptr[c + 1];
    for (uint32_t k = kb; k < ke; ++k) push(__ldg(&adj_nw[k]).x);
  }

  __device__ __forceinline__ bool recompute(uint32_t c, float band_end, float goal, float d_old, float tau_old,
                                            float& nd, float& ntau) {
    const uint32_t kb = adj_ptr[c], ke = adj_ptr[c + 1];
    float best = __uint_as_float(INF_BITS), best_du = best; uint32_t best_u = c;
    for (uint32_t k = kb; k < ke; ++k) {
      const uint2 nw = __ldg(&adj_nw[k]);
      const uint32_t u = nw.x;
      const float du = state_d(load_state(u));
      if (!(du < band_end)) continue;
      if (du > goal) continue;                               // :299
      if ((double)__ldg(&cost[u]) > cost_limit) continue;    // :302
      const float tmp = __fadd_rn(du, __uint_as_float(nw.y)); // :331
      if (tmp < best || (tmp == best && __float_as_uint(tmp) != INF_BITS &&
                         (du < best_du || (du == best_du && u < best_u)))) {
        best = tmp; best_du = du; best_u = u;
      }
    }
    nd = best; ntau = best;
    if (__float_as_uint(nd) == __float_as_uint(d_old)) {
      // same potential; the predecessor can still change among exact ties
      if (__float_as_uint(nd) == INF_BITS || pred[c] == best_u) return false;
      pred[c] = best_u;
      return false;
    }
    __stcg(&state[c], pack_state(nd, ntau));
    pred[c] = best_u;
    return true;
  }
};

}  // namespace mnb
