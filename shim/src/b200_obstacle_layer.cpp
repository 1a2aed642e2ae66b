// pluginlib layer: the interface of mesh_layers::ObstacleLayer (mesh_layers/include/mesh_layers/obstacle_layer.h:47-160).
// The ROS side of processPointCloud stays as in the reference (subscription, tf lookups, obstacle_layer.cpp:133-205); the
// computation from the cloud to the lethal set and the changed set (:207-273: range filter, transform, one ray per point
// through MeshMap::raycaster(), height filter, std::set differences) is one call into libmeshnav_b200.so.
#include <cmath>
#include <limits>
#include <set>
#include <string>
#include <vector>

#include <mesh_map/abstract_layer.h>
#include <mesh_map/mesh_map.h>
#include <pluginlib/class_list_macros.hpp>
#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/point_cloud2.hpp>
#include <tf2_ros/buffer.h>

#include <mesh_navigation_b200_plugins/b200_map.h>

namespace mesh_navigation_b200_plugins
{
class B200ObstacleLayer : public mesh_map::AbstractLayer
{
public:
  bool readLayer() override { return false; }                                                      // obstacle_layer.h:62: nothing to read
  bool writeLayer() override { return true; }                                                      // :69
  float defaultValue() override { return 0.0; }                                                    // :80
  float threshold() override { return std::numeric_limits<float>::infinity(); }                    // :89
  bool computeLayer() override { return true; }                                                    // :96: computed when sensor data arrives
  const lvr2::VertexMap<float>& costs() override { return costs_; }                                // :103
  const std::set<lvr2::VertexHandle>& lethals() override { return lethals_; }                      // :110

protected:
  bool initialize() override                                                                       // obstacle_layer.cpp:29-131
  {
    const std::string ns = mesh_map::MeshMap::MESH_MAP_NAMESPACE + "." + layer_name_;
    config_.robot_height = node_->declare_parameter(ns + ".robot_height", config_.robot_height);
    config_.max_obstacle_dist = node_->declare_parameter(ns + ".max_obstacle_dist", config_.max_obstacle_dist);
    config_.topic = node_->declare_parameter(ns + ".topic", std::string());
    const std::string qos = node_->declare_parameter(ns + ".qos", std::string("Reliable"));
    config_.tf_tolerance = node_->declare_parameter(ns + ".tf_tolerance", config_.tf_tolerance);
    const std::vector<double> dir = node_->declare_parameter(ns + ".down_axis", std::vector<double>{0.0, 0.0, -1.0});
    if (dir.size() != 3) { RCLCPP_ERROR(get_logger(), "Invalid parameter value for 'down_axis'! Must be exactly 3 values!"); return false; }
    const float len = std::sqrt(static_cast<float>(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]));
    for (int k = 0; k < 3; ++k) config_.down_axis[k] = static_cast<float>(dir[k]) / len;            // :110 normalized()
    config_.axis_frame_id = node_->declare_parameter(ns + ".axis_frame", node_->get_parameter("robot_frame").as_string());
    const auto map = map_ptr_.lock();
    if (!map) return false;
    try { b200_ = B200Map::of(map, static_cast<int>(node_->declare_parameter(ns + ".cuda_device", 0))); }
    catch (const std::exception& ex) { RCLCPP_ERROR_STREAM(get_logger(), layer_name_ << ": " << ex.what()); return false; }
    if (mnb_obstacle_reset(b200_->ctx) != MNB_OK) return false;
    rclcpp::QoS profile(10);
    if (qos == "BestEffort") profile.best_effort(); else profile.reliable();
    sub_ = node_->create_subscription<sensor_msgs::msg::PointCloud2>(
        config_.topic, profile, [this](const sensor_msgs::msg::PointCloud2::ConstSharedPtr& msg) { processPointCloud(msg); });
    return true;
  }

private:
  // rotation matrix of a unit quaternion (what Eigen::Quaternionf::toRotationMatrix yields for :176-180), row-major into m[3][4]
  static void quaternion_rows(double qx, double qy, double qz, double qw, float m[12])
  {
    const float n = std::sqrt(static_cast<float>(qx * qx + qy * qy + qz * qz + qw * qw));
    const float x = static_cast<float>(qx) / n, y = static_cast<float>(qy) / n, z = static_cast<float>(qz) / n, w = static_cast<float>(qw) / n;
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w);     m[2] = 2 * (x * z + y * w);
    m[4] = 2 * (x * y + z * w);     m[5] = 1 - 2 * (x * x + z * z); m[6] = 2 * (y * z - x * w);
    m[8] = 2 * (x * z - y * w);     m[9] = 2 * (y * z + x * w);     m[10] = 1 - 2 * (x * x + y * y);
  }

  void processPointCloud(const sensor_msgs::msg::PointCloud2::ConstSharedPtr& msg)                  // obstacle_layer.cpp:133-296
  {
    const auto map = map_ptr_.lock();
    if (nullptr == map) { RCLCPP_ERROR(get_logger(), "Could not update cost map: Failed to lock map_ptr_"); return; }
    mnb_obstacle_params p{};
    p.max_obstacle_dist = config_.max_obstacle_dist; p.robot_height = config_.robot_height;
    try
    {
      const auto tf = map->tf2Buffer().lookupTransform(map->mapFrame(), msg->header.frame_id, msg->header.stamp,
                                                       rclcpp::Duration::from_seconds(config_.tf_tolerance));   // :159-165
      quaternion_rows(tf.transform.rotation.x, tf.transform.rotation.y, tf.transform.rotation.z, tf.transform.rotation.w, p.tf);
      p.tf[3] = static_cast<float>(tf.transform.translation.x); p.tf[7] = static_cast<float>(tf.transform.translation.y);
      p.tf[11] = static_cast<float>(tf.transform.translation.z);
      const auto axis_tf = map->tf2Buffer().lookupTransform(map->mapFrame(), config_.axis_frame_id, msg->header.stamp,
                                                            rclcpp::Duration::from_seconds(config_.tf_tolerance));   // :186-196
      float r[12];
      quaternion_rows(axis_tf.transform.rotation.x, axis_tf.transform.rotation.y, axis_tf.transform.rotation.z, axis_tf.transform.rotation.w, r);
      for (int k = 0; k < 3; ++k)
        p.down_axis[k] = r[4 * k] * config_.down_axis[0] + r[4 * k + 1] * config_.down_axis[1] + r[4 * k + 2] * config_.down_axis[2];   // :195
    }
    catch (const tf2::TransformException& ex)
    {
      RCLCPP_ERROR_STREAM(get_logger(), "Failed to lookup transform into " << map->mapFrame() << ": " << ex.what());
      return;
    }
    std::vector<float> xyz;                                                                         // :207-226 (the range filter runs on the device)
    xyz.reserve(3 * static_cast<size_t>(msg->width) * msg->height);
    sensor_msgs::PointCloud2ConstIterator<float> x_it(*msg, "x"), y_it(*msg, "y"), z_it(*msg, "z");
    for (; x_it != x_it.end() && y_it != y_it.end() && z_it != z_it.end(); ++x_it, ++y_it, ++z_it)
    {
      xyz.push_back(*x_it); xyz.push_back(*y_it); xyz.push_back(*z_it);
    }
    const uint32_t V = b200_->V;
    std::vector<uint32_t> lethal_ids(V), changed_ids(V);
    uint32_t n_lethal = 0, n_changed = 0;
    const int32_t rc = mnb_obstacle_update(b200_->ctx, static_cast<uint32_t>(xyz.size() / 3), xyz.data(), &p, lethal_ids.data(), &n_lethal,
                                           changed_ids.data(), &n_changed, nullptr);              // :229-273
    if (rc != MNB_OK) { RCLCPP_ERROR_STREAM(get_logger(), layer_name_ << ": " << mnb_last_error(b200_->ctx)); return; }
    lvr2::SparseVertexMap<float> new_costs;
    std::set<lvr2::VertexHandle> new_lethals, changed;
    for (uint32_t i = 0; i < n_lethal; ++i)
    {
      new_costs.insert(lvr2::VertexHandle(lethal_ids[i]), std::numeric_limits<float>::infinity());  // :250
      new_lethals.insert(new_lethals.end(), lvr2::VertexHandle(lethal_ids[i]));                             // ascending ids: O(1) hinted inserts
    }
    for (uint32_t i = 0; i < n_changed; ++i) changed.insert(changed.end(), lvr2::VertexHandle(changed_ids[i]));
    {
      const auto wlock = this->writeLock();                                                         // :275-283
      costs_ = std::move(new_costs);
      lethals_ = std::move(new_lethals);
    }
    this->notifyChange(msg->header.stamp, changed);                                                 // :289
  }

  std::shared_ptr<B200Map> b200_;
  rclcpp::SubscriptionBase::SharedPtr sub_;
  lvr2::SparseVertexMap<float> costs_;
  std::set<lvr2::VertexHandle> lethals_;
  struct {   // obstacle_layer.h:142-151
    double robot_height = std::numeric_limits<float>::infinity();
    double max_obstacle_dist = std::numeric_limits<float>::infinity();
    std::string topic;
    double tf_tolerance = 0.1;
    float down_axis[3] = {0.0f, 0.0f, -1.0f};
    std::string axis_frame_id;
  } config_;
};
}  // namespace mesh_navigation_b200_plugins

PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200ObstacleLayer, mesh_map::AbstractLayer)
