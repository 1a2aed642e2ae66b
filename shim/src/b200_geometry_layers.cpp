// pluginlib layers: HeightDiff / Roughness / Steepness / Ridge / Clearance / Border of mesh_layers
// (mesh_layers/src/{height_diff,roughness,steepness,ridge,clearance,border}_layer.cpp), all six computed by ONE fused kernel
// (mnb_compute_layers) whose result is shared by the six plugin objects of a map, so that an existing
// `mesh_map.layers` configuration keeps working layer by layer.
#include <cmath>
#include <limits>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include <mesh_map/abstract_layer.h>
#include <mesh_map/mesh_map.h>
#include <pluginlib/class_list_macros.hpp>
#include <rclcpp/rclcpp.hpp>

#include <mesh_navigation_b200_plugins/b200_map.h>

namespace mesh_navigation_b200_plugins
{
// the cached result of one mnb_compute_layers call per map
struct FusedLayers
{
  std::mutex mtx;
  bool valid = false;
  mnb_layer_params params{0.185, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3, 0.5, 0.3, 0.5, 1.0};   // defaults of the six layer headers
  std::vector<float> costs;        // 6 x V: height_diff, roughness, steepness, ridge, clearance, border
  std::vector<uint8_t> lethal_mask;

  static std::shared_ptr<FusedLayers> of(const mesh_map::MeshMap* map)
  {
    static std::mutex m; static std::map<const mesh_map::MeshMap*, std::shared_ptr<FusedLayers>> reg;
    std::lock_guard<std::mutex> lock(m);
    auto& s = reg[map];
    if (!s) s = std::make_shared<FusedLayers>();
    return s;
  }
  bool ensure(B200Map& b200, const std::vector<float>* clearance)
  {
    std::lock_guard<std::mutex> lock(mtx);
    if (valid) return true;
    costs.resize(6 * static_cast<size_t>(b200.V)); lethal_mask.resize(b200.V);
    valid = mnb_compute_layers(b200.ctx, &params, clearance ? clearance->data() : nullptr, costs.data(), nullptr, lethal_mask.data()) == MNB_OK;
    return valid;
  }
};

// one of the six maps; WHICH = row of FusedLayers::costs / bit of the lethal mask
template <int WHICH>
class B200GeometryLayer : public mesh_map::AbstractLayer
{
public:
  bool readLayer() override { return false; }
  bool writeLayer() override { return true; }
  float defaultValue() override { return WHICH == 4 ? std::numeric_limits<float>::infinity() : 0.0f; }   // clearance_layer.h:70-73: +inf; the others 0
  float threshold() override { return static_cast<float>(threshold_); }
  const lvr2::VertexMap<float>& costs() override { return costs_; }
  const std::set<lvr2::VertexHandle>& lethals() override { return lethal_vertices_; }

  bool computeLayer() override
  {
    const auto map = map_ptr_.lock();
    if (!map) return false;
    auto fused = FusedLayers::of(map.get());
    std::vector<float> clearance;
    if (!fused->valid && clearance_from_map_)
    {
      // ClearanceLayer::computeLayer (clearance_layer.cpp:122-164): lvr2::calcNormalClearance -> one ray per vertex along its
      // normal on the device BVH (NULL = the vertex normals mnb_set_mesh computed); the cost mapping (:67-99) is fused below
      clearance.assign(b200_->V, std::numeric_limits<float>::infinity());
      if (mnb_normal_clearance(b200_->ctx, nullptr, clearance.data()) != MNB_OK)
      {
        RCLCPP_ERROR_STREAM(get_logger(), layer_name_ << ": " << mnb_last_error(b200_->ctx));
        return false;
      }
    }
    if (!fused->ensure(*b200_, clearance.empty() ? nullptr : &clearance)) return false;
    costs_.clear(); lethal_vertices_.clear();
    const float* row = &fused->costs[static_cast<size_t>(WHICH) * b200_->V];
    for (auto vH : map->mesh()->vertices())
    {
      costs_.insert(vH, row[vH.idx()]);
      if (fused->lethal_mask[vH.idx()] & (1u << WHICH)) lethal_vertices_.insert(vH);
    }
    return true;
  }

protected:
  bool initialize() override
  {
    const auto map = map_ptr_.lock();
    if (!map) return false;
    const std::string ns = mesh_map::MeshMap::MESH_MAP_NAMESPACE + "." + layer_name_;
    auto fused = FusedLayers::of(map.get());
    auto& P = fused->params;
    // parameter names of the reference layers (mesh_layers/src/*_layer.cpp, initialize()): threshold / radius / factor
    if (WHICH == 0) { P.height_diff_threshold = threshold_ = node_->declare_parameter(ns + ".threshold", P.height_diff_threshold); P.height_diff_radius = node_->declare_parameter(ns + ".radius", P.height_diff_radius); }
    if (WHICH == 1) { P.roughness_threshold = threshold_ = node_->declare_parameter(ns + ".threshold", P.roughness_threshold); P.roughness_radius = node_->declare_parameter(ns + ".radius", P.roughness_radius); }
    if (WHICH == 2) { P.steepness_threshold = threshold_ = node_->declare_parameter(ns + ".threshold", P.steepness_threshold); }
    if (WHICH == 3) { P.ridge_threshold = threshold_ = node_->declare_parameter(ns + ".threshold", P.ridge_threshold); P.ridge_radius = node_->declare_parameter(ns + ".radius", P.ridge_radius); }
    if (WHICH == 4) { P.clearance_robot_height = node_->declare_parameter(ns + ".robot_height", P.clearance_robot_height); P.clearance_height_inflation = node_->declare_parameter(ns + ".height_inflation", P.clearance_height_inflation); threshold_ = 1.0; clearance_from_map_ = true; }
    if (WHICH == 5) { P.border_threshold = threshold_ = node_->declare_parameter(ns + ".threshold", P.border_threshold); P.border_cost = node_->declare_parameter(ns + ".border_cost", P.border_cost); }
    fused->valid = false;
    try { b200_ = B200Map::of(map, static_cast<int>(node_->declare_parameter(ns + ".cuda_device", 0))); }
    catch (const std::exception& ex) { RCLCPP_ERROR_STREAM(get_logger(), layer_name_ << ": " << ex.what()); return false; }
    return true;
  }

private:
  std::shared_ptr<B200Map> b200_;
  lvr2::DenseVertexMap<float> costs_;
  std::set<lvr2::VertexHandle> lethal_vertices_;
  double threshold_ = 0.0;
  bool clearance_from_map_ = false;
};

typedef B200GeometryLayer<0> B200HeightDiffLayer;
typedef B200GeometryLayer<1> B200RoughnessLayer;
typedef B200GeometryLayer<2> B200SteepnessLayer;
typedef B200GeometryLayer<3> B200RidgeLayer;
typedef B200GeometryLayer<4> B200ClearanceLayer;
typedef B200GeometryLayer<5> B200BorderLayer;
}  // namespace mesh_navigation_b200_plugins

PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200HeightDiffLayer, mesh_map::AbstractLayer)
PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200RoughnessLayer, mesh_map::AbstractLayer)
PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200SteepnessLayer, mesh_map::AbstractLayer)
PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200RidgeLayer, mesh_map::AbstractLayer)
PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200ClearanceLayer, mesh_map::AbstractLayer)
PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200BorderLayer, mesh_map::AbstractLayer)
