// pluginlib layer: the interface of mesh_layers::InflationLayer (mesh_layers/include/mesh_layers/inflation_layer.h:47-250)
// with waveCostInflation (inflation_layer.cpp:341-491), the update set of onInputChanged (:97-179) and the repulsive
// field (vector_map_, vectorAt :493-552) served by libmeshnav_b200.so.
#include <cmath>
#include <limits>
#include <set>
#include <string>
#include <vector>

#include <mesh_map/abstract_layer.h>
#include <mesh_map/mesh_map.h>
#include <mesh_map/util.h>
#include <pluginlib/class_list_macros.hpp>
#include <rclcpp/rclcpp.hpp>

#include <mesh_navigation_b200_plugins/b200_map.h>

namespace mesh_navigation_b200_plugins
{
class B200InflationLayer : public mesh_map::AbstractLayer
{
public:
  bool readLayer() override { return false; }     // always recomputed: a wave over 5M vertices takes milliseconds
  bool writeLayer() override { return true; }
  float defaultValue() override { return 0; }                                                     // inflation_layer.h:71-74
  float threshold() override { return std::numeric_limits<float>::quiet_NaN(); }                  // inflation_layer.cpp:91-94
  const lvr2::VertexMap<float>& costs() override { return riskiness_; }
  const std::set<lvr2::VertexHandle>& lethals() override { return lethal_vertices_; }
  const boost::optional<lvr2::VertexMap<mesh_map::Vector>&> vectorMap() override { return vector_map_; }

  bool computeLayer() override                                                                     // :563-596
  {
    std::set<lvr2::VertexHandle> lethals;
    if (!inputLethals(lethals)) return false;
    std::set<lvr2::VertexHandle> update;
    if (!inflate(lethals, update)) return false;
    lethal_vertices_ = std::move(lethals);
    return true;
  }

  void onInputChanged(const rclcpp::Time& timestamp, const std::set<lvr2::VertexHandle>& changed) override   // :97-179
  {
    (void)changed;                                   // the reference re-runs the whole wave as well (:140-149)
    std::set<lvr2::VertexHandle> lethals, update;
    if (!inputLethals(lethals)) return;
    {
      const auto wlock = this->writeLock();
      if (!inflate(lethals, update)) return;
      lethal_vertices_ = std::move(lethals);
    }
    this->notifyChange(timestamp, update);           // :176 -- keys(new riskiness) U keys(old riskiness), from the device
  }

  mesh_map::Vector vectorAt(const std::array<lvr2::VertexHandle, 3>& vertices, const std::array<float, 3>& barycentric_coords) override   // :493-521
  {
    if (!config_.repulsive_field) return mesh_map::Vector();
    // one face query: find the face of the three vertices through the map (the kernel takes face ids)
    const auto map = map_ptr_.lock();
    const auto face_opt = map->mesh()->getFaceBetween(vertices[0], vertices[1], vertices[2]);
    if (!face_opt) return mesh_map::Vector();
    const uint32_t face = static_cast<uint32_t>(face_opt.unwrap().idx());
    float out[3] = {0, 0, 0};
    if (mnb_inflation_vector_at(b200_->ctx, 1, &face, barycentric_coords.data(), out) != MNB_OK) return mesh_map::Vector();
    return mesh_map::Vector(out[0], out[1], out[2]);
  }

  mesh_map::Vector vectorAt(const lvr2::VertexHandle& vH) override                                  // :523-561
  {
    if (!config_.repulsive_field) return mesh_map::Vector();
    const auto dist_opt = distances_.get(vH);
    const auto vector_opt = vector_map_.get(vH);
    if (!dist_opt || !vector_opt) return mesh_map::Vector();
    const float distance = dist_opt.get();
    const mesh_map::Vector vec = vector_opt.get();
    if (distance > config_.inflation_radius)
      if (distance > config_.inscribed_radius)       // (the reference's dangling if, kept: :541-549)
      {
        const float alpha = (std::sqrt(distance) - config_.inscribed_radius) / (config_.inflation_radius - config_.inscribed_radius) * M_PI;
        return vec * config_.inscribed_value * (std::cos(alpha) + 1) / 2.0;
      }
    if (distance > 0) return vec * config_.inscribed_value;
    return vec * config_.lethal_value;
  }

protected:
  bool initialize() override                                                                        // :598-650
  {
    const std::string ns = mesh_map::MeshMap::MESH_MAP_NAMESPACE + "." + layer_name_;
    config_.inscribed_radius = node_->declare_parameter(ns + ".inscribed_radius", config_.inscribed_radius);
    config_.inflation_radius = node_->declare_parameter(ns + ".inflation_radius", config_.inflation_radius);
    config_.lethal_value = node_->declare_parameter(ns + ".lethal_value", config_.lethal_value);
    config_.inscribed_value = node_->declare_parameter(ns + ".inscribed_value", config_.inscribed_value);
    config_.cost_scaling_factor = node_->declare_parameter(ns + ".cost_scaling_factor", config_.cost_scaling_factor);
    config_.repulsive_field = node_->declare_parameter(ns + ".repulsive_field", config_.repulsive_field);
    const auto map = map_ptr_.lock();
    if (!map) return false;
    try { b200_ = B200Map::of(map, static_cast<int>(node_->declare_parameter(ns + ".cuda_device", 0))); }
    catch (const std::exception& ex) { RCLCPP_ERROR_STREAM(get_logger(), layer_name_ << ": " << ex.what()); return false; }
    return true;
  }

private:
  bool inputLethals(std::set<lvr2::VertexHandle>& lethals)
  {
    const std::vector<std::string> inputs =
        node_->get_parameter(mesh_map::MeshMap::MESH_MAP_NAMESPACE + "." + layer_name_ + ".inputs").as_string_array();
    if (inputs.size() != 1) { RCLCPP_ERROR(get_logger(), "[B200InflationLayer] Exactly one input layer is required!"); return false; }
    const auto map = map_ptr_.lock();
    if (!map) return false;
    const auto input = map->layer(inputs[0]);
    if (nullptr == input) { RCLCPP_ERROR(get_logger(), "[B200InflationLayer] Could not get layer '%s' from map!", inputs[0].c_str()); return false; }
    const auto input_lock = input->readLock();
    lethals = input->lethals();
    return true;
  }

  // the wave + the update set + (optionally) the repulsive field; fills riskiness_, distances_, vector_map_
  bool inflate(const std::set<lvr2::VertexHandle>& lethals, std::set<lvr2::VertexHandle>& update)
  {
    const auto map = map_ptr_.lock();
    const uint32_t V = b200_->V;
    std::vector<uint32_t> le; le.reserve(lethals.size());
    for (auto vH : lethals) le.push_back(static_cast<uint32_t>(vH.idx()));
    std::vector<uint8_t> inv(V, 0);
    for (auto vH : map->mesh()->vertices()) inv[vH.idx()] = map->invalid[vH] ? 1 : 0;
    const mnb_inflation_params p{config_.inscribed_radius, config_.inflation_radius, config_.lethal_value, config_.inscribed_value,
                                 config_.cost_scaling_factor};                                       // inflation_layer.h:240-248
    std::vector<float> dist(V), cost(V); std::vector<uint32_t> upd(V); uint32_t n_upd = 0;
    const int32_t rc = mnb_inflation_update(b200_->ctx, le.data(), static_cast<uint32_t>(le.size()), inv.data(), &p, dist.data(),
                                            cost.data(), upd.data(), &n_upd);
    if (rc != MNB_OK) { RCLCPP_ERROR_STREAM(get_logger(), layer_name_ << ": " << mnb_last_error(b200_->ctx)); return false; }
    riskiness_.clear(); distances_.clear(); vector_map_.clear();
    for (uint32_t v = 0; v < V; ++v)
    {
      if (!std::isnan(cost[v])) riskiness_.insert(lvr2::VertexHandle(v), cost[v]);                  // :484-490
      if (std::isfinite(dist[v])) distances_.insert(lvr2::VertexHandle(v), dist[v]);
    }
    for (uint32_t i = 0; i < n_upd; ++i) update.insert(lvr2::VertexHandle(upd[i]));
    if (config_.repulsive_field)
    {
      std::vector<float> vec(3 * static_cast<size_t>(V));
      if (mnb_inflation_vector_map(b200_->ctx, vec.data()) == MNB_OK)
        for (uint32_t v = 0; v < V; ++v)
        {
          const float* q = &vec[3 * static_cast<size_t>(v)];
          if (q[0] != 0 || q[1] != 0 || q[2] != 0) vector_map_.insert(lvr2::VertexHandle(v), mesh_map::Vector(q[0], q[1], q[2]));
        }
    }
    return true;
  }

  std::shared_ptr<B200Map> b200_;
  lvr2::DenseVertexMap<float> riskiness_;
  lvr2::DenseVertexMap<float> distances_;
  lvr2::DenseVertexMap<mesh_map::Vector> vector_map_;
  std::set<lvr2::VertexHandle> lethal_vertices_;
  struct {   // inflation_layer.h:240-248
    double inscribed_radius = 0.25;
    double inflation_radius = 0.4;
    double lethal_value = 1.0;
    double inscribed_value = 0.99;
    double cost_scaling_factor = 1.0;
    int min_contour_size = 3;
    bool repulsive_field = true;
  } config_;
};
}  // namespace mesh_navigation_b200_plugins

PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200InflationLayer, mesh_map::AbstractLayer)
