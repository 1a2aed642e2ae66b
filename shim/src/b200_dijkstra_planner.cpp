// pluginlib planner: the interface of dijkstra_mesh_planner::DijkstraMeshPlanner
// (dijkstra_mesh_planner/src/dijkstra_mesh_planner.cpp:57-215) with dijkstra() (:217-398) served by libmeshnav_b200.so.
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
class B200DijkstraMeshPlanner : public mbf_mesh_core::MeshPlanner
{
public:
  bool initialize(const std::string& plugin_name, const std::shared_ptr<mesh_map::MeshMap>& mesh_map_ptr,
                  const rclcpp::Node::SharedPtr& node) override
  {
    mesh_map_ = mesh_map_ptr; name_ = plugin_name; node_ = node; map_frame_ = mesh_map_->mapFrame();
    config_.publish_vector_field = node_->declare_parameter(name_ + ".publish_vector_field", config_.publish_vector_field);
    config_.publish_face_vectors = node_->declare_parameter(name_ + ".publish_face_vectors", config_.publish_face_vectors);
    config_.goal_dist_offset = node_->declare_parameter(name_ + ".goal_dist_offset", config_.goal_dist_offset);
    config_.cost_limit = node_->declare_parameter(name_ + ".cost_limit", config_.cost_limit);
    const int device = static_cast<int>(node_->declare_parameter(name_ + ".cuda_device", 0));
    try { b200_ = B200Map::of(mesh_map_, device); }
    catch (const std::exception& ex) { RCLCPP_ERROR_STREAM(node_->get_logger(), name_ << ": " << ex.what()); return false; }
    return true;
  }

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

    // dijkstra(goal_vec, start_vec, path) (:84): seeded at the goal, stops once the robot's vertex is fixed
    const auto seed_opt = mesh_map_->getNearestVertexHandle(goal_vec);     // :235
    const auto robot_opt = mesh_map_->getNearestVertexHandle(start_vec);   // :236
    if (!seed_opt) return mbf_msgs::action::GetPath::Result::INVALID_START;
    if (!robot_opt) return mbf_msgs::action::GetPath::Result::INVALID_GOAL;
    const uint32_t seed_vertex = static_cast<uint32_t>(seed_opt.unwrap().idx());
    const uint32_t robot_vertex = static_cast<uint32_t>(robot_opt.unwrap().idx());

    std_msgs::msg::Header header; header.stamp = node_->now(); header.frame_id = mesh_map_->mapFrame();
    cost = 0; plan.clear();
    if (seed_vertex == robot_vertex) return mbf_msgs::action::GetPath::Result::SUCCESS;   // :252-255 (empty path, empty plan)

    try { b200_->pushCosts(*mesh_map_); }
    catch (const std::exception& ex) { message = ex.what(); return mbf_msgs::action::GetPath::Result::INTERNAL_ERROR; }
    std::vector<float> dist(b200_->V); std::vector<uint32_t> pred(b200_->V);
    const int32_t rc = mnb_dijkstra(b200_->ctx, seed_vertex, static_cast<int64_t>(robot_vertex), config_.cost_limit, config_.goal_dist_offset,
                                    dist.data(), pred.data());
    if (rc < 0) { message = mnb_last_error(b200_->ctx); return mbf_msgs::action::GetPath::Result::INTERNAL_ERROR; }
    if (rc != MNB_SUCCESS) return static_cast<uint32_t>(rc);               // CANCELED / NO_PATH_FOUND, same codes (:353-365)

    // predecessor walk from the robot's vertex down to the seed (:367-373: the robot's own vertex is not part of the path,
    // the seed vertex is); after path.reverse() (:86) this is plan order
    std::list<lvr2::VertexHandle> path;
    for (uint32_t v = robot_vertex, guard = 0; v != seed_vertex && guard <= b200_->V; ++guard)
    {
      v = pred[v];
      path.push_back(lvr2::VertexHandle(v));
    }
    const auto mesh = mesh_map_->mesh();
    const auto& vertex_normals = mesh_map_->vertexNormals();
    mesh_map::Vector vec = start_vec;
    mesh_map::Normal normal = vertex_normals[path.front()];
    float dir_length = 0;
    geometry_msgs::msg::PoseStamped pose; pose.header = header;
    while (!path.empty())                                                   // :96-113
    {
      const lvr2::VertexHandle vH = path.front();
      const mesh_map::Vector next = mesh->getVertexPosition(vH);
      pose.pose = mesh_map::calculatePoseFromPosition(vec, next, normal, dir_length);
      cost += dir_length; vec = next; normal = vertex_normals[vH];
      plan.push_back(pose);
      path.pop_front();
    }
    pose.pose = mesh_map::calculatePoseFromPosition(vec, goal_vec, normal, dir_length);
    cost += dir_length;
    plan.push_back(pose);
    RCLCPP_INFO_STREAM(node_->get_logger(), "Path length: " << cost << "m");

    if (config_.publish_vector_field)                                       // computeVectorMap (:189-209)
    {
      std::vector<float> vec3(3 * static_cast<size_t>(b200_->V));
      if (mnb_vector_map(b200_->ctx, pred.data(), nullptr, nullptr, vec3.data()) == MNB_OK)
      {
        lvr2::DenseVertexMap<mesh_map::Vector> vector_map;
        for (auto vH : mesh->vertices())
        {
          const float* q = &vec3[3 * vH.idx()];
          if (q[0] == q[0]) vector_map.insert(vH, mesh_map::Vector(q[0], q[1], q[2]));
        }
        mesh_map_->publishVectorField("vector_field", vector_map, config_.publish_face_vectors);
      }
    }
    return mbf_msgs::action::GetPath::Result::SUCCESS;
  }

  bool cancel() override { if (b200_ && b200_->ctx) mnb_cancel(b200_->ctx); return true; }

private:
  std::shared_ptr<mesh_map::MeshMap> mesh_map_;
  std::shared_ptr<B200Map> b200_;
  std::string name_, map_frame_;
  rclcpp::Node::SharedPtr node_;
  struct {   // dijkstra_mesh_planner.h:178-187
    bool publish_vector_field = false;
    bool publish_face_vectors = false;
    double goal_dist_offset = 0.3;
    double cost_limit = 1.0;
  } config_;
};
}  // namespace mesh_navigation_b200_plugins

PLUGINLIB_EXPORT_CLASS(mesh_navigation_b200_plugins::B200DijkstraMeshPlanner, mbf_mesh_core::MeshPlanner)
