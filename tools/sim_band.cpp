// DEVELOPMENT TOOL (not product, not oracle): sequential CPU emulation of the
// round structure of the GPU "sliding band" wavefront (Jacobi rounds), used in the
// build container (no GPU) to check the parallel semantics against the oracle
// before the CUDA kernel runs on a B200.  Mirrors CvpProblem::replay / face_time and
// run_band_rounds (band_engine.cuh, problems.cuh); shares wavefront_math.cuh /
// topology.hpp with the kernels so the per-face arithmetic is the same source.
//
// g++ -O2 -ffp-contract=off -shared -fPIC -o tools/libsimband.so tools/sim_band.cpp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../mesh_navigation_b200/csrc/topology.hpp"
#include "../mesh_navigation_b200/csrc/wavefront_math.cuh"

using namespace mnb;

namespace {
constexpr float FINF = std::numeric_limits<float>::infinity();

#ifndef SIM_ANGLES
#define SIM_ANGLES false
#endif
constexpr int MAXL = 12;
static int g_levels = 3;   // the CUDA labels track 3 water levels
struct Tm3 { float a[MAXL]; uint32_t minor; };
struct Lab { float d; Tm3 t; };
static inline bool tless(const Tm3& x, const Tm3& y) {
  for (int i = 0; i < g_levels; ++i) { if (x.a[i] < y.a[i]) return true; if (x.a[i] > y.a[i]) return false; }
  return x.minor < y.minor;
}
static inline bool teq(const Tm3& x, const Tm3& y) {
  for (int i = 0; i < g_levels; ++i) if (x.a[i] != y.a[i]) return false;
  return x.minor == y.minor;
}
static inline Tm3 tnormal(float key, uint32_t c) { Tm3 t{}; t.a[0] = key; t.minor = 2u * c; return t; }

struct Sim {
  HostTopology T;
  const float* w; const float* cost; const uint8_t* invalid;
  double cost_limit;
  std::vector<Lab> L;
  uint32_t s[3]; uint32_t seed_noexpand = 0;
  float band_end, goal;

  int seed_index(uint32_t v) const { return v == s[0] ? 0 : (v == s[1] ? 1 : (v == s[2] ? 2 : -1)); }

  std::vector<uint32_t> chg; uint32_t round = 0; mutable float blocked_m = FINF; mutable bool deferred_flag = false;
  bool face_time(uint32_t v1, uint32_t v2, Tm3& T, uint32_t& tv) const {
    const Lab &a = L[v1], &b = L[v2];
    if (!(a.d < band_end) || !(b.d < band_end)) return false;
    if (invalid && (invalid[v1] || invalid[v2])) return false;
    const int i1 = seed_index(v1), i2 = seed_index(v2);
    const Tm3 &ta = a.t, &tb = b.t;
    const bool v1_later = tless(tb, ta);
    if (i1 >= 0 && i2 >= 0) {
      const bool e1 = !((seed_noexpand >> i1) & 1u), e2 = !((seed_noexpand >> i2) & 1u);
      if (!e1 && !e2) return false;
      const bool use1 = e1 && (!e2 || !v1_later);
      T = use1 ? ta : tb; tv = use1 ? v1 : v2;
      return true;
    }
    const int il = v1_later ? i1 : i2;
    if (il >= 0 && ((seed_noexpand >> il) & 1u)) return false;
    if ((v1_later ? a.d : b.d) > goal) return false;
    T = v1_later ? ta : tb; tv = v1_later ? v1 : v2;
    return true;
  }

  mutable std::vector<uint64_t> sig; mutable size_t same_inputs = 0, total_replays = 0;
  Lab replay(uint32_t c) const {
    struct Cand { Tm3 T; uint32_t k; float u1, u2; uint32_t tv; };
    Cand cs[64]; int n = 0;
    for (uint32_t k = T.vcor_ptr[c]; k < T.vcor_ptr[c + 1] && n < 64; ++k) {
      Tm3 Tt; uint32_t tv;
      if (!face_time(T.cor_v1[k], T.cor_v2[k], Tt, tv)) continue;
      cs[n++] = {Tt, k, L[T.cor_v1[k]].d, L[T.cor_v2[k]].d, tv};
    }
    {
      uint64_t h = 1469598103934665603ull;
      auto mixf = [&](float f) { uint32_t u; memcpy(&u, &f, 4); h = (h ^ u) * 1099511628211ull; };
      for (int i = 0; i < n; ++i) { mixf(cs[i].u1); mixf(cs[i].u2); mixf(cs[i].T.a[0]); mixf(cs[i].T.a[1]); h = (h ^ cs[i].k) * 1099511628211ull; h = (h ^ cs[i].T.minor) * 1099511628211ull; }
      if (sig.size() != L.size()) sig.assign(L.size(), 0);
      total_replays++;
      if (sig[c] == h) same_inputs++;
      sig[c] = h;
    }
    float cur = FINF; Tm3 tc = tnormal(FINF, c);
    for (int i = 0; i < n; ++i) {
      int b = i;
      for (int j = i + 1; j < n; ++j)
        if (tless(cs[j].T, cs[b].T) || (!tless(cs[b].T, cs[j].T) && cs[j].k < cs[b].k)) b = j;
      std::swap(cs[i], cs[b]);
      if (!tless(cs[i].T, tc)) break;
      CvpResult r;
      const uint32_t k = cs[i].k;
      if (cvp_update_t<SIM_ANGLES>(cs[i].u1, cs[i].u2, cur, w[T.cor_ea[k]], w[T.cor_eb[k]], w[T.cor_ec[k]], r)) {
        // a back-step label is only taken from a trigger that has been stable for a whole round
        if (!(r.value > cs[i].T.a[0]) && !(chg[cs[i].tv] < round)) { blocked_m = std::fmin(blocked_m, cs[i].T.a[0]); deferred_flag = true; continue; }
        cur = r.value;
        const Tm3& F = cs[i].T;
        // monotonic stack of water levels: keep the levels above the new key, then the key itself
        Tm3 nt{}; int lvl = 0;
        while (lvl < g_levels && !(r.value > F.a[lvl])) { nt.a[lvl] = F.a[lvl]; ++lvl; }
        if (lvl < g_levels) { nt.a[lvl] = r.value; nt.minor = 2u * c; }
        else nt.minor = F.minor + 1u;          // deeper than we track: right after the trigger
        tc = nt;
      }
    }
    return {cur, tc};
  }
};
}  // namespace

static std::vector<Lab> g_last;
extern "C" void sim_get_label(uint32_t v, double* out) { out[0]=g_last[v].d; out[1]=g_last[v].t.a[0]; out[2]=g_last[v].t.a[1]; out[3]=g_last[v].t.minor; }

extern "C" int sim_cvp_band(uint32_t V, uint32_t F, const uint32_t* faces, const uint32_t* edges, uint32_t E,
                            const float* pos, const float* edge_weights, const float* vertex_costs,
                            const uint8_t* invalid, uint32_t seed_face, const float* seed_pos,
                            double cost_limit, double delta, int flags, float* out_dist, double* stats) {
  Sim S;
  g_levels = flags > 0 ? flags : 3;
  S.T.build(V, F, faces, edges, E);
  S.w = edge_weights; S.cost = vertex_costs; S.invalid = invalid; S.cost_limit = cost_limit;
  S.goal = FINF;
  S.L.resize(V); S.chg.assign(V, 0);
  for (uint32_t v = 0; v < V; ++v) { S.L[v].d = FINF; S.L[v].t = tnormal(FINF, v); }
  std::vector<uint8_t> mark(V, 0);   // 0 none, 1 cand, 2 fixed, 3 cand+activated
  std::vector<uint32_t> cand, next;
  auto eligible = [&](uint32_t x) { return !(invalid && invalid[x]) && !((double)vertex_costs[x] >= cost_limit); };
  auto activate = [&](uint32_t v, std::vector<uint32_t>& out) {
    for (uint32_t k = S.T.vcor_ptr[v]; k < S.T.vcor_ptr[v + 1]; ++k)
      for (uint32_t x : {S.T.cor_v1[k], S.T.cor_v2[k]})
        if (mark[x] == 0 && eligible(x)) { mark[x] = 1; out.push_back(x); }
  };
  float seed_min = FINF, seed_max = 0;
  for (int k = 0; k < 3; ++k) {
    const uint32_t v = faces[3 * (size_t)seed_face + k];
    S.s[k] = v;
    const float dx = seed_pos[0] - pos[3 * (size_t)v], dy = seed_pos[1] - pos[3 * (size_t)v + 1],
                dz = seed_pos[2] - pos[3 * (size_t)v + 2];
    const float d = std::sqrt(dx * dx + dy * dy + dz * dz);
    S.L[v].d = d; S.L[v].t = tnormal(d, v);
    mark[v] = 2;
    seed_min = std::fmin(seed_min, d); seed_max = std::fmax(seed_max, d);
    if (((double)vertex_costs[v] >= cost_limit) || (invalid && invalid[v])) S.seed_noexpand |= 1u << k;
  }
  for (int k = 0; k < 3; ++k) activate(S.s[k], cand);

  size_t rounds = 0, recomputes = 0, dirty_recomputes = 0, clean_but_changed = 0;
  const int n_sweeps = getenv("SIM_SWEEPS") ? atoi(getenv("SIM_SWEEPS")) : 0;
  std::vector<uint32_t> chg_stamp(V, 0), eval_stamp(V, 0); uint32_t stamp = 1;
  std::vector<uint32_t> dst(V, 0xffffffffu), evr(V, 0xffffffffu); std::vector<float> wake(V, 0.0f); size_t skipped = 0, skip_wrong = 0;
  float m_prev = 0.0f, lo_prev = seed_min, band_end_prev = std::nextafter(seed_max, FINF);
  const size_t max_rounds = 2 * (size_t)V + 64;
  std::vector<Lab> nl;
  if (stats) stats[2] = 0;
  while (!cand.empty()) {
    if (rounds > 0 && m_prev == FINF && lo_prev == FINF) break;
    if (rounds > max_rounds) { if (stats) stats[2] = 1; break; }
    float band_end = (float)(lo_prev + (float)delta);
    if (!(band_end > band_end_prev)) band_end = band_end_prev;
    S.band_end = band_end; S.round = (uint32_t)rounds;
    next.clear(); ++stamp;
    float m = FINF, lo = FINF; S.blocked_m = FINF;
    // Jacobi: decide "fixed" on the labels of the previous round, recompute the rest from old labels
    nl.resize(cand.size());
    std::vector<uint8_t> fixed_now(cand.size(), 0);
    for (size_t i = 0; i < cand.size(); ++i) {
      const uint32_t c = cand[i];
      const Lab old = S.L[c];
      if (old.t.a[0] < m_prev && old.t.a[0] < band_end_prev) { fixed_now[i] = 1; continue; }
      S.deferred_flag = false;
      nl[i] = S.replay(c); recomputes++;
      {
        // push-stamp skip rule (validation only: we still recompute and check the prediction)
        const bool need = evr[c] == 0xffffffffu || dst[c] != 0xffffffffu && dst[c] >= evr[c] || band_end > wake[c];
        if (!need) { skipped++; if (nl[i].d != old.d || !teq(nl[i].t, old.t)) skip_wrong++; }
        evr[c] = S.deferred_flag ? 0xffffffffu : (uint32_t)rounds;
        float wk = FINF;
        for (uint32_t k = S.T.vcor_ptr[c]; k < S.T.vcor_ptr[c + 1]; ++k) {
          const float da = S.L[S.T.cor_v1[k]].d, db = S.L[S.T.cor_v2[k]].d;
          if ((!(da < band_end) || !(db < band_end)) && da < FINF && db < FINF) wk = std::fmin(wk, std::fmax(da, db));
        }
        wake[c] = wk;
      }
    }
    for (size_t i = 0; i < cand.size(); ++i) {
      const uint32_t c = cand[i];
      if (fixed_now[i]) { mark[c] = 2; continue; }
      const Lab old = S.L[c];
      if (nl[i].d != old.d || !teq(nl[i].t, old.t)) {
        m = std::fmin(m, std::fmin(old.t.a[0], nl[i].t.a[0]));
        if (old.d < FINF) S.chg[c] = (uint32_t)rounds + 1;   // first-time labelling is not a re-label
        S.L[c] = nl[i]; chg_stamp[c] = stamp;
        for (uint32_t k = S.T.vcor_ptr[c]; k < S.T.vcor_ptr[c + 1]; ++k) { dst[S.T.cor_v1[k]] = (uint32_t)rounds; dst[S.T.cor_v2[k]] = (uint32_t)rounds; }
      }
      eval_stamp[c] = stamp;
      lo = std::fmin(lo, nl[i].d);
      next.push_back(c);
      if (nl[i].d < FINF && mark[c] == 1) { mark[c] = 3; activate(c, next); }
    }
    // ---- inner sweeps (experiment): extra Jacobi passes over the surviving list inside the same round ----
    for (int sw = 0; sw < n_sweeps; ++sw) {
      ++stamp;
      const size_t ns = next.size();
      std::vector<Lab> nl2(ns); std::vector<uint8_t> ev(ns, 0);
      for (size_t i = 0; i < ns; ++i) {
        const uint32_t c = next[i];
        // dirty rule: some corner source changed at or after my last evaluation
        bool dirty = false;
        for (uint32_t k = S.T.vcor_ptr[c]; k < S.T.vcor_ptr[c + 1] && !dirty; ++k)
          dirty = chg_stamp[S.T.cor_v1[k]] >= eval_stamp[c] || chg_stamp[S.T.cor_v2[k]] >= eval_stamp[c];
        nl2[i] = S.replay(c); ev[i] = 1;
        if (dirty) { dirty_recomputes++; }
        else if (nl2[i].d != S.L[c].d || !teq(nl2[i].t, S.L[c].t)) clean_but_changed++;
      }
      bool any = false;
      for (size_t i = 0; i < ns; ++i) {
        const uint32_t c = next[i];
        const Lab old = S.L[c];
        eval_stamp[c] = stamp;
        if (nl2[i].d != old.d || !teq(nl2[i].t, old.t)) {
          any = true;
          m = std::fmin(m, std::fmin(old.t.a[0], nl2[i].t.a[0]));
          if (old.d < FINF) S.chg[c] = (uint32_t)rounds + 1;
          S.L[c] = nl2[i]; chg_stamp[c] = stamp;
        }
        if (nl2[i].d < FINF && mark[c] == 1) { mark[c] = 3; activate(c, next); }
      }
      if (!any) break;
    }
    if (n_sweeps > 0) { lo = FINF; for (uint32_t c : next) lo = std::fmin(lo, S.L[c].d); }
    cand.swap(next);
    m = std::fmin(m, S.blocked_m);   // a deferred back-step is a pending change at its trigger's pop time
    m_prev = m; lo_prev = lo; band_end_prev = band_end;
    rounds++;
  }
  fprintf(stderr, "[sim] push-stamp skip: skipped/V=%.2f of recomputes/V=%.2f wrong=%zu\n", (double)skipped / V, (double)recomputes / V, skip_wrong);
  if (getenv("SIM_SWEEPS")) fprintf(stderr, "[sim] sweeps=%d dirty_recomputes/V=%.2f clean_but_changed=%zu\n", n_sweeps, (double)dirty_recomputes / V, clean_but_changed);
  for (uint32_t v = 0; v < V; ++v) out_dist[v] = S.L[v].d;
  g_last = S.L;
  if (stats) { stats[0] = (double)rounds; stats[1] = (double)recomputes; stats[3] = (double)S.same_inputs / (double)(S.total_replays ? S.total_replays : 1); }
  return 0;
}

// ---- inflation (mirrors InflationProblem::recompute + k_inflate round structure) -----------------
namespace {
static inline Tm3 accept_time3(uint32_t c, float X, const Tm3& F) {
  Tm3 nt{}; int lvl = 0;
  while (lvl < g_levels && !(X > F.a[lvl])) { nt.a[lvl] = F.a[lvl]; ++lvl; }
  if (lvl < g_levels) { nt.a[lvl] = X; nt.minor = 2u * c; }
  else nt.minor = F.minor + 1u;
  return nt;
}
}
extern "C" int sim_inflation(uint32_t V, uint32_t F, const uint32_t* faces, const uint32_t* edges, uint32_t E,
                             const float* edge_dist, const uint8_t* invalid, const uint32_t* lethals, uint32_t nl,
                             float max_distance, float* out_dist, double* stats, int trace_v) {
  g_levels = 3;
  HostTopology T; T.build(V, F, faces, edges, E);
  std::vector<Lab> L(V);
  for (uint32_t v = 0; v < V; ++v) { L[v].d = FINF; L[v].t = tnormal(FINF, v); }
  std::vector<uint8_t> mark(V, 0);
  std::vector<uint32_t> chg(V, 0);   // round of the last label change (0 = never / initial)
  size_t rounds = 0;
  std::vector<uint32_t> cand, next;
  for (uint32_t i = 0; i < nl; ++i) { const uint32_t v = lethals[i]; L[v].d = 0; L[v].t = tnormal(0.0f, v); mark[v] = 2; }
  auto activate = [&](uint32_t v, std::vector<uint32_t>& out) {
    for (uint32_t k = T.vcor_ptr[v]; k < T.vcor_ptr[v + 1]; ++k)
      for (uint32_t x : {T.cor_v1[k], T.cor_v2[k]}) if (mark[x] == 0) { mark[x] = 1; out.push_back(x); }
  };
  for (uint32_t i = 0; i < nl; ++i) activate(lethals[i], cand);
  float blocked_m = FINF;
  auto replay = [&](uint32_t c, bool trace) -> Lab {
    struct Cand { Tm3 T; uint32_t k; float u1, u2; uint32_t tv; };
    Cand cs[64]; int n = 0;
    for (uint32_t k = T.vcor_ptr[c]; k < T.vcor_ptr[c + 1] && n < 64; ++k) {
      const uint32_t v1 = T.cor_v1[k], v2 = T.cor_v2[k];
      const Lab &a = L[v1], &b = L[v2];
      if (a.t.a[0] == FINF || b.t.a[0] == FINF) continue;
      const bool l1 = a.d == 0.0f, l2 = b.d == 0.0f;
      const bool i1 = invalid && invalid[v1], i2 = invalid && invalid[v2];
      if ((i1 && !l1) || (i2 && !l2)) continue;
      const bool v1_later = tless(b.t, a.t);
      Tm3 Tt; uint32_t tv;
      if (l1 && l2) {
        const bool e1 = !i1, e2 = !i2;
        if (!e1 && !e2) continue;
        const bool use1 = e1 && (!e2 || !v1_later);
        Tt = use1 ? a.t : b.t; tv = use1 ? v1 : v2;
      } else {
        if (v1_later ? i1 : i2) continue;
        Tt = v1_later ? a.t : b.t; tv = v1_later ? v1 : v2;
      }
      cs[n++] = {Tt, k, a.d, b.d, tv};
    }
    float cur = FINF; Tm3 tc = tnormal(FINF, c);
    const bool never_fixed = invalid && invalid[c];   // popped but never fixed (:417): keeps receiving updates
    for (int i = 0; i < n; ++i) {
      int b = i;
      for (int j = i + 1; j < n; ++j) if (tless(cs[j].T, cs[b].T) || (!tless(cs[b].T, cs[j].T) && cs[j].k < cs[b].k)) b = j;
      std::swap(cs[i], cs[b]);
      if (!never_fixed && !tless(cs[i].T, tc)) break;
      const uint32_t k = cs[i].k;
      const float cnd = inflation_candidate(cs[i].u1, cs[i].u2, edge_dist[T.cor_ea[k]], edge_dist[T.cor_eb[k]], edge_dist[T.cor_ec[k]]);
      if (trace) printf("   v%u face%u src(%u:%g,%u:%g) T=(%g,%g,%g,%u) cand=%.9g cur=%.9g\n", c, T.cor_face[k], T.cor_v1[k], cs[i].u1, T.cor_v2[k], cs[i].u2, cs[i].T.a[0], cs[i].T.a[1], cs[i].T.a[2], cs[i].T.minor, cnd, cur);
      // a non-causal (back-step) label may only be taken from a trigger whose label has been stable for a
      // whole round: breaks self-sustaining cyclic dependencies between a trigger and its own child
      const bool backstep = !(cnd > cs[i].T.a[0]);
      if (backstep && !(chg[cs[i].tv] < (uint32_t)rounds)) { blocked_m = std::fmin(blocked_m, cs[i].T.a[0]); continue; }   // chg = 1 + round of last change
      if (cnd < cur) {
        cur = cnd;
        if (cs[i].u1 <= max_distance && cs[i].u2 <= max_distance) tc = accept_time3(c, cnd, cs[i].T);
      }
    }
    return {cur, tc};
  };
  float m_prev = 0.0f, lo_prev = 0.0f;
  const size_t max_rounds = 400;
  std::vector<Lab> nl2;
  if (stats) stats[2] = 0;
  while (!cand.empty()) {
    if (rounds > 0 && m_prev == FINF && lo_prev == FINF) break;
    if (rounds > max_rounds) { if (stats) stats[2] = 1; break; }
    next.clear();
    float m = FINF, lo = FINF; blocked_m = FINF;
    nl2.resize(cand.size());
    std::vector<uint8_t> fixed_now(cand.size(), 0);
    for (size_t i = 0; i < cand.size(); ++i) {
      const uint32_t c = cand[i];
      if (L[c].t.a[0] < m_prev) { fixed_now[i] = 1; continue; }
      nl2[i] = replay(c, (int)c == trace_v && rounds > max_rounds - 4);
    }
    for (size_t i = 0; i < cand.size(); ++i) {
      const uint32_t c = cand[i];
      if (fixed_now[i]) { mark[c] = 2; continue; }
      const Lab old = L[c];
      if (nl2[i].d != old.d || !teq(nl2[i].t, old.t)) { m = std::fmin(m, std::fmin(old.t.a[0], nl2[i].t.a[0])); if (old.d < FINF) chg[c] = (uint32_t)rounds + 1; L[c] = nl2[i]; }
      lo = std::fmin(lo, nl2[i].t.a[0]);
      next.push_back(c);
      if (nl2[i].d < FINF && mark[c] == 1) { mark[c] = 3; activate(c, next); }
    }
    cand.swap(next);
    m = std::fmin(m, blocked_m);
    m_prev = m; lo_prev = lo; rounds++;
  }
  for (uint32_t v = 0; v < V; ++v) out_dist[v] = L[v].d;
  if (stats) { stats[0] = (double)rounds; }
  return 0;
}
