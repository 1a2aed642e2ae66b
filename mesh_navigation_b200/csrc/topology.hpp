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
#include <utility>
#include <vector>

namespace mnb {

struct HostTopology {
  uint32_t V = 0, F = 0, E = 0;
  std::vector<uint32_t> edges;
  std::vector<uint32_t> face_edges;  // 3F: edge(v0,v1), edge(v1,v2), edge(v2,v0)
  std::vector<uint32_t> vadj_ptr, vadj_nbr, vadj_eid;
  std::vector<uint32_t> vcor_ptr, cor_v1, cor_v2, cor_face, cor_ec, cor_eb, cor_ea;
  std::vector<uint8_t> cor_side;     // per corner record, bit i = this face is NOT the lowest-id face of edge i (0: ec, 1: eb, 2: ea),
                                     // i.e. the second face pmp lists for the halfedge pair (inflation_layer.cpp:427)
  std::vector<uint8_t> border;       // 1 = vertex lies on an edge with a single face

  static inline uint64_t ekey(uint32_t a, uint32_t b) {
    return a < b ? (((uint64_t)a << 32) | b) : (((uint64_t)b << 32) | a);
  }

  void build(uint32_t V_, uint32_t F_, const uint32_t* faces, const uint32_t* edges_in, uint32_t E_in) {
    V = V_; F = F_;
    for (size_t i = 0; i < 3 * (size_t)F; ++i)
      if (faces[i] >= V) throw std::runtime_error("face index out of range");
    std::vector<std::pair<uint64_t, uint32_t>> lut;
    if (edges_in) {
      E = E_in;
      edges.assign(edges_in, edges_in + 2 * (size_t)E);
      lut.resize(E);
      for (uint32_t e = 0; e < E; ++e) lut[e] = {ekey(edges[2 * (size_t)e], edges[2 * (size_t)e + 1]), e};
      std::sort(lut.begin(), lut.end());
    } else {
      std::vector<uint64_t> keys(3 * (size_t)F);
      for (uint32_t f = 0; f < F; ++f) {
        const uint32_t* v = faces + 3 * (size_t)f;
        keys[3 * (size_t)f] = ekey(v[0], v[1]);
        keys[3 * (size_t)f + 1] = ekey(v[1], v[2]);
        keys[3 * (size_t)f + 2] = ekey(v[2], v[0]);
      }
      std::sort(keys.begin(), keys.end());
      keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
      E = (uint32_t)keys.size();
      edges.resize(2 * (size_t)E);
      lut.resize(E);
      for (uint32_t e = 0; e < E; ++e) {
        edges[2 * (size_t)e] = (uint32_t)(keys[e] >> 32);
        edges[2 * (size_t)e + 1] = (uint32_t)keys[e];
        lut[e] = {keys[e], e};
      }
    }
    auto find_edge = [&](uint32_t a, uint32_t b) -> uint32_t {
      const uint64_t k = ekey(a, b);
      auto it = std::lower_bound(lut.begin(), lut.end(), std::make_pair(k, (uint32_t)0));
      if (it == lut.end() || it->first != k) throw std::runtime_error("face edge missing from edge list");
      return it->second;
    };
    face_edges.resize(3 * (size_t)F);
    std::vector<uint8_t> edge_face_cnt(E, 0);
    std::vector<uint32_t> edge_first_face(E, 0xffffffffu);
    for (uint32_t f = 0; f < F; ++f) {
      const uint32_t* v = faces + 3 * (size_t)f;
      for (int k = 0; k < 3; ++k) {
        const uint32_t e = find_edge(v[k], v[(k + 1) % 3]);
        face_edges[3 * (size_t)f + k] = e;
        if (edge_face_cnt[e] < 255) edge_face_cnt[e]++;
        if (edge_first_face[e] == 0xffffffffu) edge_first_face[e] = f;
      }
    }
    border.assign(V, 0);
    for (uint32_t e = 0; e < E; ++e)
      if (edge_face_cnt[e] == 1) { border[edges[2 * (size_t)e]] = 1; border[edges[2 * (size_t)e + 1]] = 1; }

    vadj_ptr.assign((size_t)V + 1, 0);
    for (uint32_t e = 0; e < E; ++e) { vadj_ptr[edges[2 * (size_t)e] + 1]++; vadj_ptr[edges[2 * (size_t)e + 1] + 1]++; }
    for (uint32_t v = 0; v < V; ++v) vadj_ptr[v + 1] += vadj_ptr[v];
    vadj_nbr.resize(vadj_ptr[V]); vadj_eid.resize(vadj_ptr[V]);
    {
      std::vector<uint32_t> cur(vadj_ptr.begin(), vadj_ptr.end() - 1);
      for (uint32_t e = 0; e < E; ++e) {
        const uint32_t a = edges[2 * (size_t)e], b = edges[2 * (size_t)e + 1];
        vadj_eid[cur[a]] = e; vadj_nbr[cur[a]++] = b;
        vadj_eid[cur[b]] = e; vadj_nbr[cur[b]++] = a;
      }
    }
    vcor_ptr.assign((size_t)V + 1, 0);
    for (size_t i = 0; i < 3 * (size_t)F; ++i) vcor_ptr[faces[i] + 1]++;
    for (uint32_t v = 0; v < V; ++v) vcor_ptr[v + 1] += vcor_ptr[v];
    const size_t NC = vcor_ptr[V];
    cor_v1.resize(NC); cor_v2.resize(NC); cor_face.resize(NC);
    cor_ec.resize(NC); cor_eb.resize(NC); cor_ea.resize(NC); cor_side.resize(NC);
    {
      std::vector<uint32_t> cur(vcor_ptr.begin(), vcor_ptr.end() - 1);
      for (uint32_t f = 0; f < F; ++f) {
        const uint32_t* v = faces + 3 * (size_t)f;
        const uint32_t* fe = &face_edges[3 * (size_t)f];
        for (int k = 0; k < 3; ++k) {  // v3 = v[k], v1 = v[k+1], v2 = v[k+2]
          const uint32_t slot = cur[v[k]]++;
          cor_v1[slot] = v[(k + 1) % 3];
          cor_v2[slot] = v[(k + 2) % 3];
          cor_face[slot] = f;
          cor_ec[slot] = fe[(k + 1) % 3];  // edge(v1,v2)
          cor_eb[slot] = fe[k];            // edge(v3,v1)
          cor_ea[slot] = fe[(k + 2) % 3];  // edge(v2,v3)
          cor_side[slot] = (uint8_t)((edge_first_face[cor_ec[slot]] != f ? 1 : 0) | (edge_first_face[cor_eb[slot]] != f ? 2 : 0) |
                                     (edge_first_face[cor_ea[slot]] != f ? 4 : 0));
        }
      }
    }
  }
};

}  // namespace mnb
