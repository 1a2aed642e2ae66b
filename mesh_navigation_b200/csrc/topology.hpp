// Host-side flattening of a triangle mesh into the CSR layouts the kernels read.
//
// Replaces, for the hot path only, what the reference gets from the lvr2
// half-edge mesh (un-vendored): getEdgesOfVertex / getVerticesOfEdge
// (dijkstra_mesh_planner.cpp:308,324), getFacesOfVertex / getVerticesOfFace /
// getEdgeBetween / getFaceBetween (cvp_mesh_planner.cpp:776,780,380-388,423).
// The mesh is immutable after MeshMap::readMap (mesh_map.cpp:149-452), so this
// runs once per map, not per plan.
//
// Layouts
//   edges[2E]            (lo,hi) sorted by (lo,hi) unless the caller supplies its own edge order
//   vertex -> neighbours  CSR  vadj_ptr[V+1], vadj_nbr[], vadj_eid[]   (ascending edge id)
//   vertex -> corners     CSR  vcor_ptr[V+1], one record per incident face of vertex v3, with
//                         v1 = next(v3), v2 = next(v1) in the face's cyclic order -- the argument
//                         order of waveFrontUpdate(v1,v2,v3) at cvp:811,834,857 / inflation:450,458,466
//                         cor_v1, cor_v2, cor_face, and the three edge ids
//                         cor_ec = edge(v1,v2), cor_eb = edge(v1,v3), cor_ea = edge(v2,v3)
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

namespace mnb {

struct HostTopology {
  uint32_t V = 0, F = 0, E = 0;
  std::vector<uint32_t> edges;
  std::vector<uint32_t> face_edges;  // 3F: edge(v0,v1), edge(v1,v2), edge(v2,v0)
  std::vector<uint32_t> vadj_ptr, vadj_nbr, vadj_eid;
  std::vector<uint32_t> vcor_ptr, cor_v1, cor_v2, cor_face, cor_ec, cor_eb, cor_ea;
  std::vector<uint32_t> face_cor;    // 3F: index of the corner record of face f at its k-th vertex (the way back from a face to its three records)
  std::vector<uint8_t> cor_side;     // per corner record, bit i = this face is NOT the lowest-id face of edge i (0: ec, 1: eb, 2: ea),
                                     // i.e. the second face pmp lists for the halfedge pair (inflation_layer.cpp:427)
  std::vector<uint8_t> border;       // 1 = vertex lies on an edge with a single face

  static inline uint64_t ekey(uint32_t a, uint32_t b) {
    return a < b ? (((uint64_t)a << 32) | b) : (((uint64_t)b << 32) | a);
  }

  // host threads of the builder (the map is built once per MeshMap::readMap; 50M vertices = 300M half-edges)
  static unsigned workers(size_t n) {
    if (n < (1u << 20)) return 1;
    const unsigned hc = std::thread::hardware_concurrency();
    return hc < 2 ? 1 : (hc > 16 ? 16 : hc);
  }
  template <class Fn>
  static void parallel_for(unsigned T, Fn fn) {            // fn(t) for t in [0, T)
    if (T <= 1) { fn(0u); return; }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(fn, t);
    fn(0u);
    for (auto& x : th) x.join();
  }

  // LSD radix sort of (key, value) pairs by key, 16-bit digits; digits that are constant over the input are skipped.
  // Parallel over contiguous chunks (per-thread histograms, one prefix over (digit, thread)): stable, same result as the
  // sequential sort whatever the thread count.
  static void radix_sort_pairs(std::vector<uint64_t>& keys, std::vector<uint32_t>& vals) {
    const size_t n = keys.size();
    if (n < 2) return;
    std::vector<uint64_t> k2(n); std::vector<uint32_t> v2(n);
    const unsigned T = workers(n);
    std::vector<std::vector<size_t>> cnt(T, std::vector<size_t>(65536));
    auto lo = [&](unsigned t) { return n * t / T; };
    for (int pass = 0; pass < 4; ++pass) {
      const int sh = 16 * pass;
      parallel_for(T, [&](unsigned t) {
        auto& c = cnt[t]; std::fill(c.begin(), c.end(), (size_t)0);
        for (size_t i = lo(t), e = lo(t + 1); i < e; ++i) c[(keys[i] >> sh) & 0xffffu]++;
      });
      size_t first = 0;
      for (unsigned t = 0; t < T; ++t) first += cnt[t][(keys[0] >> sh) & 0xffffu];
      if (first == n) continue;                                   // all keys share this digit
      size_t sum = 0;
      for (size_t d = 0; d < 65536; ++d)
        for (unsigned t = 0; t < T; ++t) { const size_t c = cnt[t][d]; cnt[t][d] = sum; sum += c; }
      parallel_for(T, [&](unsigned t) {
        auto& c = cnt[t];
        for (size_t i = lo(t), e = lo(t + 1); i < e; ++i) { const size_t d = c[(keys[i] >> sh) & 0xffffu]++; k2[d] = keys[i]; v2[d] = vals[i]; }
      });
      keys.swap(k2); vals.swap(v2);
    }
  }

  void build(uint32_t V_, uint32_t F_, const uint32_t* faces, const uint32_t* edges_in, uint32_t E_in) {
    V = V_; F = F_;
    for (size_t i = 0; i < 3 * (size_t)F; ++i)
      if (faces[i] >= V) throw std::runtime_error("face index out of range");
    // Edge ids of the 3F half-edges without a per-half-edge binary search: radix-sort (key, slot) pairs, then one linear
    // pass assigns ids to runs of equal keys (own numbering: ascending (lo,hi); caller numbering: merge with the sorted
    // caller keys).  5 M vertices: ~3x faster than sort + unique + lower_bound per half-edge.
    const size_t H = 3 * (size_t)F;
    std::vector<uint64_t> hk(H); std::vector<uint32_t> hs(H);
    {
      const unsigned T = workers(H);
      parallel_for(T, [&](unsigned t) {
        for (uint32_t f = (uint32_t)((size_t)F * t / T), fe = (uint32_t)((size_t)F * (t + 1) / T); f < fe; ++f) {
          const uint32_t* v = faces + 3 * (size_t)f;
          for (int k = 0; k < 3; ++k) { hk[3 * (size_t)f + k] = ekey(v[k], v[(k + 1) % 3]); hs[3 * (size_t)f + k] = 3 * f + (uint32_t)k; }
        }
      });
    }
    radix_sort_pairs(hk, hs);
    face_edges.resize(H);
    std::vector<uint8_t> edge_face_cnt;
    std::vector<uint32_t> edge_first_face;
    if (edges_in) {
      E = E_in;
      for (size_t i = 0; i < 2 * (size_t)E; ++i)
        if (edges_in[i] >= V) throw std::runtime_error("edge endpoint out of range");
      edges.assign(edges_in, edges_in + 2 * (size_t)E);
      std::vector<uint64_t> ck(E); std::vector<uint32_t> cid(E);
      for (uint32_t e = 0; e < E; ++e) { ck[e] = ekey(edges[2 * (size_t)e], edges[2 * (size_t)e + 1]); cid[e] = e; }
      radix_sort_pairs(ck, cid);
      for (uint32_t e = 1; e < E; ++e)
        if (ck[e] == ck[e - 1]) throw std::runtime_error("duplicate edge in the caller's edge list");
      edge_face_cnt.assign(E, 0); edge_first_face.assign(E, 0xffffffffu);
      size_t c = 0;
      for (size_t i = 0; i < H; ++i) {
        while (c < E && ck[c] < hk[i]) ++c;
        if (c == E || ck[c] != hk[i]) throw std::runtime_error("face edge missing from edge list");
        const uint32_t e = cid[c], f = hs[i] / 3;
        face_edges[hs[i]] = e;
        if (edge_face_cnt[e] < 255) edge_face_cnt[e]++;
        if (f < edge_first_face[e]) edge_first_face[e] = f;
      }
    } else {
      size_t ne = 0;
      for (size_t i = 0; i < H; ++i) if (i == 0 || hk[i] != hk[i - 1]) ++ne;
      E = (uint32_t)ne;
      edges.resize(2 * (size_t)E);
      edge_face_cnt.assign(E, 0); edge_first_face.assign(E, 0xffffffffu);
      uint32_t e = 0xffffffffu;
      for (size_t i = 0; i < H; ++i) {
        if (i == 0 || hk[i] != hk[i - 1]) { ++e; edges[2 * (size_t)e] = (uint32_t)(hk[i] >> 32); edges[2 * (size_t)e + 1] = (uint32_t)hk[i]; }
        const uint32_t f = hs[i] / 3;
        face_edges[hs[i]] = e;
        if (edge_face_cnt[e] < 255) edge_face_cnt[e]++;
        if (f < edge_first_face[e]) edge_first_face[e] = f;
      }
    }
    std::vector<uint64_t>().swap(hk); std::vector<uint32_t>().swap(hs);
    border.assign(V, 0);
    for (uint32_t e = 0; e < E; ++e)
      if (edge_face_cnt[e] == 1) { border[edges[2 * (size_t)e]] = 1; border[edges[2 * (size_t)e + 1]] = 1; }

    vadj_ptr.assign((size_t)V + 1, 0);
    for (uint32_t e = 0; e < E; ++e) { vadj_ptr[edges[2 * (size_t)e] + 1]++; vadj_ptr[edges[2 * (size_t)e + 1] + 1]++; }
    for (uint32_t v = 0; v < V; ++v) vadj_ptr[v + 1] += vadj_ptr[v];
    vadj_nbr.resize(vadj_ptr[V]); vadj_eid.resize(vadj_ptr[V]);
    {
      std::vector<uint32_t> cur(vadj_ptr.begin(), vadj_ptr.end() - 1);
      const unsigned T = workers(2 * (size_t)E);
      parallel_for(T, [&](unsigned t) {                    // vertex ranges again: ascending edge id per vertex is kept
        const uint32_t vlo = (uint32_t)((size_t)V * t / T), vhi = (uint32_t)((size_t)V * (t + 1) / T);
        for (uint32_t e = 0; e < E; ++e) {
          const uint32_t a = edges[2 * (size_t)e], b = edges[2 * (size_t)e + 1];
          if (a >= vlo && a < vhi) { vadj_eid[cur[a]] = e; vadj_nbr[cur[a]++] = b; }
          if (b >= vlo && b < vhi) { vadj_eid[cur[b]] = e; vadj_nbr[cur[b]++] = a; }
        }
      });
    }
    vcor_ptr.assign((size_t)V + 1, 0);
    for (size_t i = 0; i < 3 * (size_t)F; ++i) vcor_ptr[faces[i] + 1]++;
    for (uint32_t v = 0; v < V; ++v) vcor_ptr[v + 1] += vcor_ptr[v];
    const size_t NC = vcor_ptr[V];
    cor_v1.resize(NC); cor_v2.resize(NC); cor_face.resize(NC);
    cor_ec.resize(NC); cor_eb.resize(NC); cor_ea.resize(NC); cor_side.resize(NC);
    face_cor.resize(3 * (size_t)F);
    {
      // every thread owns a range of vertices and scans all faces: the records of a vertex stay in ascending face order
      std::vector<uint32_t> cur(vcor_ptr.begin(), vcor_ptr.end() - 1);
      const unsigned T = workers(3 * (size_t)F);
      parallel_for(T, [&](unsigned t) {
      const uint32_t vlo = (uint32_t)((size_t)V * t / T), vhi = (uint32_t)((size_t)V * (t + 1) / T);
      for (uint32_t f = 0; f < F; ++f) {
        const uint32_t* v = faces + 3 * (size_t)f;
        const uint32_t* fe = &face_edges[3 * (size_t)f];
        for (int k = 0; k < 3; ++k) {  // v3 = v[k], v1 = v[k+1], v2 = v[k+2]
          if (v[k] < vlo || v[k] >= vhi) continue;
          const uint32_t slot = cur[v[k]]++;
          cor_v1[slot] = v[(k + 1) % 3];
          cor_v2[slot] = v[(k + 2) % 3];
          cor_face[slot] = f;
          face_cor[3 * (size_t)f + k] = slot;
          cor_ec[slot] = fe[(k + 1) % 3];  // edge(v1,v2)
          cor_eb[slot] = fe[k];            // edge(v3,v1)
          cor_ea[slot] = fe[(k + 2) % 3];  // edge(v2,v3)
          cor_side[slot] = (uint8_t)((edge_first_face[cor_ec[slot]] != f ? 1 : 0) | (edge_first_face[cor_eb[slot]] != f ? 2 : 0) |
                                     (edge_first_face[cor_ea[slot]] != f ? 4 : 0));
        }
      }
      });
    }
  }
};

}  // namespace mnb
