// Marshalling between mesh_map::MeshMap (lvr2 attribute maps) and the flat arrays of include/meshnav_b200.h.
// One B200Map per MeshMap instance, shared by every plugin of this package that is loaded for that map.
//
// Reference interfaces this header reads (paths relative to naturerobots/mesh_navigation):
//   mesh_map/include/mesh_map/mesh_map.h:276 mesh(), :292 vertexCosts(), :342 edgeWeights(), :350 edgeDistances(),
//   :447 invalid, :516 edge_cost_factor; lvr2::PMPMesh vertex / face / edge iteration as in mesh_map.cpp:404-425.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include <mesh_map/mesh_map.h>
#include <meshnav_b200.h>

namespace mesh_navigation_b200_plugins
{
class B200Map
{
public:
  mnb_ctx* ctx = nullptr;
  uint32_t V = 0, F = 0, E = 0;
  std::vector<uint32_t> edge_row;   // lvr2 EdgeHandle::idx() -> row of the flat edge arrays (lvr2 edge ids may have gaps)

  ~B200Map() { if (ctx) mnb_destroy(ctx); }

  // the one context of `map` (created and filled on first use)
  static std::shared_ptr<B200Map> of(const std::shared_ptr<mesh_map::MeshMap>& map, int device = 0)
  {
    static std::mutex mtx;
    static std::map<const mesh_map::MeshMap*, std::weak_ptr<B200Map>> registry;
    std::lock_guard<std::mutex> lock(mtx);
    auto& slot = registry[map.get()];
    if (auto alive = slot.lock()) return alive;
    auto fresh = std::make_shared<B200Map>();
    fresh->upload(*map, device);
    slot = fresh;
    return fresh;
  }

  // once per MeshMap::readMap: positions, faces in lvr2's cyclic order, edges in lvr2's EdgeHandle order
  void upload(mesh_map::MeshMap& map, int device)
  {
    const auto mesh = map.mesh();
    V = static_cast<uint32_t>(mesh->nextVertexIndex());
    std::vector<float> pos(3 * static_cast<size_t>(V), 0.0f);
    for (auto vH : mesh->vertices())
    {
      const auto p = mesh->getVertexPosition(vH);
      pos[3 * vH.idx()] = p.x; pos[3 * vH.idx() + 1] = p.y; pos[3 * vH.idx() + 2] = p.z;
    }
    std::vector<uint32_t> faces, edges;
    for (auto fH : mesh->faces())
      for (auto vH : mesh->getVerticesOfFace(fH)) faces.push_back(static_cast<uint32_t>(vH.idx()));
    for (auto eH : mesh->edges())
    {
      const auto vs = mesh->getVerticesOfEdge(eH);
      if (edge_row.size() <= eH.idx()) edge_row.resize(eH.idx() + 1, 0xffffffffu);
      edge_row[eH.idx()] = static_cast<uint32_t>(edges.size() / 2);
      edges.push_back(static_cast<uint32_t>(vs[0].idx()));
      edges.push_back(static_cast<uint32_t>(vs[1].idx()));
    }
    F = static_cast<uint32_t>(faces.size() / 3);
    E = static_cast<uint32_t>(edges.size() / 2);
    check(mnb_create(device, &ctx), "mnb_create");
    check(mnb_set_mesh(ctx, V, F, pos.data(), faces.data(), edges.data(), E), "mnb_set_mesh");
  }

  // per plan: MeshMap::vertexCosts() / edgeWeights() / invalid (cvp_mesh_planner.cpp:663-664, :245)
  void pushCosts(mesh_map::MeshMap& map)
  {
    std::vector<float> vc(V, 0.0f), ew(E, 0.0f);
    std::vector<uint8_t> inv(V, 0);
    const auto mesh = map.mesh();
    const auto& vertex_costs = map.vertexCosts();
    const auto& edge_weights = map.edgeWeights();
    for (auto vH : mesh->vertices()) { vc[vH.idx()] = vertex_costs[vH]; inv[vH.idx()] = map.invalid[vH] ? 1 : 0; }
    for (auto eH : mesh->edges()) ew[edge_row[eH.idx()]] = edge_weights[eH];
    check(mnb_set_costs(ctx, vc.data(), ew.data(), inv.data()), "mnb_set_costs");
    costs_pushed_ = true;
  }

  // MeshMap::layerChanged + updateEdgeWeights (mesh_map.cpp:455-492, 563-618) for the changed vertices only
  void layerChanged(mesh_map::MeshMap& map, const std::set<lvr2::VertexHandle>& changes)
  {
    if (!costs_pushed_) { pushCosts(map); return; }
    std::vector<uint32_t> ids; std::vector<float> costs;
    const auto& vertex_costs = map.vertexCosts();
    for (auto vH : changes) { ids.push_back(static_cast<uint32_t>(vH.idx())); costs.push_back(vertex_costs[vH]); }
    check(mnb_update_vertex_costs(ctx, static_cast<uint32_t>(ids.size()), ids.data(), costs.data(), 0, 0.0f, map.edge_cost_factor),
          "mnb_update_vertex_costs");
  }

  void check(int32_t rc, const char* what) const
  {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + (ctx ? mnb_last_error(ctx) : "no context") + " (" + std::to_string(rc) + ")");
  }

private:
  bool costs_pushed_ = false;
};
}  // namespace mesh_navigation_b200_plugins
