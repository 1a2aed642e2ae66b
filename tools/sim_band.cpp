// DEVELOPMENT TOOL (not product, not oracle): sequential CPU emulation of the
// round structure of the GPU "sliding band" wavefront, used in the build
// container (no GPU) to check the parallel semantics against the oracle before
// the CUDA kernel runs on a B200.  Shares wavefront_math.cuh / topology.hpp
// with the kernels so the per-face arithmetic is the same source.
//
// g++ -O2 -ffp-contract=off -shared -fPIC -o tools/libsimband.so tools/sim_band.cpp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <vector>

#include "../mesh_navigation_b200/csrc/topology.hpp"
#include "../mesh_navigation_b200/csrc/wavefront_math.cuh"

using namespace mnb;

namespace {
constexpr float FINF = std::numeric_limits<float>::infinity();

struct Sim {
  HostTopology T;
  const float* w; const float* cost; const uint8_t* invalid;
  double cost_limit; int gate; int flags_;
  std::vector<float> d, tau;
  std::vector<uint8_t> fixed;
  float band_end;

  // recompute vertex c from scratch, emulating the reference's per-vertex event order
  float recompute(uint32_t c, float* tau_out) {
    struct Cand { float T; uint32_t Tid; float u1, u2, a, b, cw; };
    Cand cs[64]; int n = 0;
    for (uint32_t k = T.vcor_ptr[c]; k < T.vcor_ptr[c + 1] && n < 64; ++k) {
      const uint32_t v1 = T.cor_v1[k], v2 = T.cor_v2[k];
      if (invalid && (invalid[v1] || invalid[v2])) continue;
      const float d1 = d[v1], d2 = d[v2];
      const bool av1 = fixed[v1] || d1 < band_end, av2 = fixed[v2] || d2 < band_end;
      if (!av1 || !av2) continue;
      // event time of a vertex = (tau, id) lexicographic; the later of (v1,v2) is the one whose
      // pop visits the face: it must expand (cvp:757)
      const bool v1_later = tau[v1] > tau[v2] || (tau[v1] == tau[v2] && v1 > v2);
      const uint32_t later = v1_later ? v1 : v2;
      if (!(cost[later] < cost_limit)) continue;
      cs[n++] = {tau[later], later, d1, d2, w[T.cor_ea[k]], w[T.cor_eb[k]], w[T.cor_ec[k]]};
    }
    float cur = FINF, tcur = FINF;
    for (int i = 0; i < n; ++i) {
      int best = i;
      for (int j = i + 1; j < n; ++j)
        if (cs[j].T < cs[best].T || (cs[j].T == cs[best].T && cs[j].Tid < cs[best].Tid)) best = j;
      std::swap(cs[i], cs[best]);
      const bool before = cs[i].T < tcur || (cs[i].T == tcur && cs[i].Tid < c);
      if (gate && !before) break;   // c would already have been popped
      CvpResult r;
      if (cvp_update(cs[i].u1, cs[i].u2, cur, cs[i].a, cs[i].b, cs[i].cw, r)) {
        cur = r.value;
        tcur = std::fmax(cur, cs[i].T);
        if ((flags_ & 4) && tcur == cs[i].T && !(cs[i].Tid < c)) tcur = std::nextafter(tcur, FINF);
      }
    }
    *tau_out = tcur;
    return cur;
  }
};
}  // namespace

extern "C" int sim_cvp_band(uint32_t V, uint32_t F, const uint32_t* faces, const uint32_t* edges, uint32_t E,
                            const float* pos, const float* edge_weights, const float* vertex_costs,
                            const uint8_t* invalid, uint32_t seed_face, const float* seed_pos,
                            double cost_limit, double delta, int flags, float* out_dist, double* stats) {
  Sim S;
  S.T.build(V, F, faces, edges, E);
  S.w = edge_weights; S.cost = vertex_costs; S.invalid = invalid; S.cost_limit = cost_limit;
  S.gate = flags & 1; S.flags_ = flags;
  const bool use_tau = flags & 2;
  S.d.assign(V, FINF); S.tau.assign(V, FINF); S.fixed.assign(V, 0);
  std::vector<uint8_t> in_cand(V, 0);
  std::vector<uint32_t> cand;
  auto add_neighbours = [&](uint32_t v) {
    for (uint32_t k = S.T.vcor_ptr[v]; k < S.T.vcor_ptr[v + 1]; ++k)
      for (uint32_t x : {S.T.cor_v1[k], S.T.cor_v2[k]})
        if (!S.fixed[x] && !in_cand[x] && !(invalid && invalid[x]) && vertex_costs[x] < cost_limit) {
          in_cand[x] = 1; cand.push_back(x);
        }
  };
  float seed_min = FINF;
  for (int k = 0; k < 3; ++k) {
    const uint32_t v = faces[3 * (size_t)seed_face + k];
    const float dx = seed_pos[0] - pos[3 * (size_t)v], dy = seed_pos[1] - pos[3 * (size_t)v + 1],
                dz = seed_pos[2] - pos[3 * (size_t)v + 2];
    S.d[v] = std::sqrt(dx * dx + dy * dy + dz * dz);
    S.fixed[v] = 1;
    seed_min = std::fmin(seed_min, S.d[v]);
  }
  for (int k = 0; k < 3; ++k) { const uint32_t v = faces[3 * (size_t)seed_face + k]; S.tau[v] = seed_min; }
  for (int k = 0; k < 3; ++k) add_neighbours(faces[3 * (size_t)seed_face + k]);

  size_t rounds = 0, recomputes = 0;
  std::vector<float> nd, nt;
  std::vector<uint8_t> was_avail;
  const size_t max_rounds = 200000;
  while (!cand.empty()) {
    rounds++;
    if (rounds > max_rounds) { if (stats) { stats[0] = -1; } break; }
    float lo = FINF;
    for (uint32_t c : cand) lo = std::fmin(lo, S.d[c]);
    S.band_end = (lo == FINF) ? FINF : (float)(lo + delta);
    nd.resize(cand.size()); nt.resize(cand.size());
    for (size_t i = 0; i < cand.size(); ++i) { nd[i] = S.recompute(cand[i], &nt[i]); recomputes++; }
    float m = FINF; bool any = false;
    const size_t ncand = cand.size();
    for (size_t i = 0; i < ncand; ++i) {
      const uint32_t c = cand[i];
      const float old = S.d[c];
      if (nd[i] != old) { any = true; m = std::fmin(m, std::fmin(old, nd[i])); }
    }
    for (size_t i = 0; i < ncand; ++i) {
      const uint32_t c = cand[i];
      const bool newly_avail = !(S.d[c] < S.band_end) && (nd[i] < S.band_end);
      S.d[c] = nd[i]; S.tau[c] = use_tau ? nt[i] : nd[i];
      if (newly_avail || (nd[i] < S.band_end)) add_neighbours(c);
    }
    // everything strictly below the smallest changed value is a converged prefix
    const float fix_below = any ? m : S.band_end;
    size_t wr = 0;
    for (size_t i = 0; i < cand.size(); ++i) {
      const uint32_t c = cand[i];
      if (S.d[c] < fix_below && S.d[c] < S.band_end) { S.fixed[c] = 1; in_cand[c] = 0; }
      else cand[wr++] = c;
    }
    cand.resize(wr);
    if (!any && fix_below == FINF) {
      // nothing labelled and nothing changed: remaining candidates are unreachable
      bool labelled = false;
      for (uint32_t c : cand) if (S.d[c] < FINF) { labelled = true; break; }
      if (!labelled) break;
    }
  }
  for (uint32_t v = 0; v < V; ++v) out_dist[v] = S.d[v];
  if (stats) { stats[0] = (double)rounds; stats[1] = (double)recomputes; }
  return 0;
}
