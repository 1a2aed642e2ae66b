// pluginlib planner: the interface of cvp_mesh_planner::CVPMeshPlanner (cvp_mesh_planner/src/cvp_mesh_planner.cpp:50-239),
// with waveFrontPropagation (:651-970) served by libmeshnav_b200.so.  Drop-in: same plugin base class, same parameters
// (cost_limit, step_width, goal_dist_offset, publish_vector_field, publish_face_vectors), same GetPath result codes.
#include <list>
#include <string>
#include <vector>

#include <mbf_mesh_core/mesh_planner.h>
#include <mbf_msgs/action/get_path.hpp>
#include <mesh_map/util.h>
#include <pluginlib/class_list_macros.hpp>
#include <rclcpp/rclcpp.hpp>

#include <mesh_navigation_b200_plugins/b200_map.h>

namespace mesh_navigation_b200_plugins
{
class B200CVPMeshPlanner : public mbf_mesh_core::MeshPlanner
{
public:
  typedef std::shared_ptr<B200CVPMeshPlanner> Ptr;

  // mbf_mesh_core/mesh_planner.h:87
  bool initialize(const std::string& plugin_name, const std::shared_ptr<mesh_map::MeshMap>& mesh_map_ptr,
                  const rclcpp::Node::SharedPtr& node) override
  {
    mesh_map_ = mesh_map_ptr; name_ = plugin_name; node_ = node; map_frame_ = mesh_map_->mapFrame();
    config_.publish_vector_field = node_->declare_parameter(name_ + ".publish_vector_field", config_.publish_vector_field);
    config_.publish_face_vectors = node_->declare_parameter(name_ + ".publish_face_vectors", config_.publish_face_vectors);
    config_.goal_dist_offset = node_->declare_parameter(name_ + ".goal_dist_offset", config_.goal_dist_offset);
    config_.cost_limit = node_->declare_parameter(name_ + ".cost_limit", config_.cost_limit);
    config_.step_width = node_->declare_parameter(name_ + ".step_width", config_.step_width);
    const int device = static_cast<int>(node_->declare_parameter(name_ + ".cuda_device", 0));
    try { b200_ = B200Map::of(mesh_map_, device); }
    catch (const std::exception& ex) { RCLCPP_ERROR_STREAM(node_->get_logger(), name_ << ": " << ex.what()); return false; }
    return true;
  }

  // mbf_mesh_core/mesh_planner.h:71-73; body follows cvp_mesh_planner.cpp:62-140
  uint32_t makePlan(const geometry_msgs::msg::PoseStamped& start, const geometry_msgs::msg::PoseStamped& goal, double tolerance,
                    std::vector<geometry_msgs::msg::PoseStamped>& plan, double& cost, std::string& message) override
  {
    (void)tolerance;
    geometry_msgs::msg::PoseStamped start_in_map, goal_in_map;
    try { start_in_map = mesh_map_->transformToMapFrame(start); goal_in_map = mesh_map_->transformToMapFrame(goal); }
    catch (const std::exception& ex)
    {
      RCLCPP_ERROR_STREAM(node_->get_logger(), name_ << ": could not transform start or goal to '" << map_frame_ << "': " << ex.what());
      return mbf_msgs::action::GetPath::Result::TF_ERROR;
    }
    mesh_map::Vector start_vec = mesh_map::toVector(start_in_map.pose.position);
    mesh_map::Vector goal_vec = mesh_map::toVector(goal_in_map.pose.position);

    // the wave is seeded at the GOAL and stops once the robot's face is fixed (waveFrontPropagation(goal_vec, start_vec, ..), :90)
    const lvr2::OptionalFaceHandle seed_opt = mesh_map_->getContainingFace(goal_vec, 0.4);    // :673
    const lvr2::OptionalFaceHandle robot_opt = mesh_map_->getContainingFace(start_vec, 0.4);  // :674
    if (!seed_opt) { message = "Could not find a face close enough to the given start pose"; return mbf_msgs::action::GetPath::Result::INVALID_START; }
    if (!robot_opt) { message = "Could not find a face close enough to the given goal pose"; return mbf_msgs::action::GetPath::Result::INVALID_GOAL; }
    const uint32_t seed_face = static_cast<uint32_t>(seed_opt.unwrap().idx());
    const uint32_t robot_face = static_cast<uint32_t>(robot_opt.unwrap().idx());

    try { b200_->pushCosts(*mesh_map_); }
    catch (const std::exception& ex) { message = ex.what(); return mbf_msgs::action::GetPath::Result::INTERNAL_ERROR; }

    // potentials / predecessors / directions / cutting faces stay on the device; only the path comes back
    const float seed_pos[3] = {goal_vec.x, goal_vec.y, goal_vec.z};
    const int32_t rc = mnb_cvp(b200_->ctx, seed_face, seed_pos, static_cast<int64_t>(robot_face), config_.cost_limit,
                               config_.goal_dist_offset, nullptr, nullptr, nullptr, nullptr);
    if (rc < 0) { message = mnb_last_error(b200_->ctx); return mbf_msgs::action::GetPath::Result::INTERNAL_ERROR; }
    if (rc != MNB_SUCCESS) { message = rc == MNB_CANCELED ? "Wave front propagation has been canceled!" : "Predecessor of the goal is not set! No path found!"; return static_cast<uint32_t>(rc); }

    // vector-field back-tracking from the robot to the goal (:920-951), on the device
    constexpr uint32_t kMaxPoints = 1u << 18;
    std::vector<float> pts(3 * static_cast<size_t>(kMaxPoints)); std::vector<uint32_t> pfaces(kMaxPoints); uint32_t n_points = 0;
    const float robot_pos[3] = {start_vec.x, start_vec.y, start_vec.z};
    const int32_t brc = mnb_cvp_backtrack(b200_->ctx, robot_pos, robot_face, config_.step_width, kMaxPoints, pts.data(), pfaces.data(), &n_points);
    if (brc < 0) { message = mnb_last_error(b200_->ctx); return mbf_msgs::action::GetPath::Result::INTERNAL_ERROR; }
    if (brc != MNB_SUCCESS) { message = "Could not find a valid path, while back-tracking from the goal"; return static_cast<uint32_t>(brc); }

    // poses as the reference builds them (:104-124): orientation from consecutive points and the face normal
    std_msgs::msg::Header header; header.stamp = node_->now(); header.frame_id = mesh_map_->mapFrame();
    cost = 0; plan.clear();
    const auto& face_normals = mesh_map_->faceNormals();
    for (uint32_t i = 0; i + 1 < n_points; ++i)
    {
      const mesh_map::Vector cur(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), next(pts[3 * i + 3], pts[3 * i + 4], pts[3 * i + 5]);
      float dir_length = 0;
      geometry_msgs::msg::PoseStamped pose; pose.header = header;
      pose.pose = mesh_map::calculatePoseFromPosition(cur, next, face_normals[lvr2::FaceHandle(pfaces[i])], dir_length);
      cost += dir_length;
      plan.push_back(pose);
    }
    geometry_msgs::msg::PoseStamped last; last.header = header; last.pose = goal_in_map.pose;     // :117-120
    plan.push_back(last);
    RCLCPP_INFO_STREAM(node_->get_logger(), "Path length: " << cost << "m");
    if (config_.publish_vector_field) publishVectorField();
    return mbf_msgs::action::GetPath::Result::SUCCESS;
  }

  // mbf_mesh_core/mesh_planner.h:79; cvp_mesh_planner.cpp:142-146
  bool cancel() override { if (b200_ && b200_->ctx) mnb_cancel(b200_->ctx); return true; }

private:
  // computeVectorMap (:204-239) of the last plan, fetched only when it is to be published
  void publishVectorField()
  {
    std::vector<float> vec(3 * static_cast<size_t>(b200_->V));
    if (mnb_vector_map(b200_->ctx, nullptr, nullptr, nullptr, vec.data()) != MNB_OK) return;
    lvr2::DenseVertexMap<mesh_map::Vector> vector_map;
    for (auto vH : mesh_map_->mesh()->vertices())
    {
      const float* v = &vec[3 * vH.idx()];
      if (v[0] == v[0]) vector_map.insert(vH, mesh_map::Vector(v[0], v[1], v[2]));            // NaN = no entry
    }
    mesh_map_->publishVectorField("vector_field", vector_map, config_.publish_face_vectors);
  }

  std::shared_ptr<mesh_map::MeshMap> mesh_map_;
  std::shared_ptr<B200Map> b200_;
  std::string name_, map_frame_;
  rclcpp::Node::SharedPtr node_;
  struct {   // cvp_mesh_planner.h:201-212
    bool publish_vector_field = false;
    bool publish_face_vectors = false;
    double goal_dist_offset = 0.3;
    double cost_limit = 1.0;
    double step_width = 0.4;
  } config_;
};
}  // namespace mesh_navigation_b200_plugins

PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200CVPMeshPlanner, mbf_mesh_core::MeshPlanner)
