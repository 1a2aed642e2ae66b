// stub of mbf_mesh_core::MeshPlanner (mbf_mesh_core/include/mbf_mesh_core/mesh_planner.h:50-92), virtual interface verbatim
#pragma once
#include <memory>
#include <string>
#include <vector>
#include <mesh_map/mesh_map.h>
namespace mbf_abstract_core { class AbstractPlanner { public: virtual ~AbstractPlanner() {} }; }
namespace mbf_mesh_core {
class MeshPlanner : public mbf_abstract_core::AbstractPlanner {
public:
  typedef std::shared_ptr<mbf_mesh_core::MeshPlanner> Ptr;
  virtual ~MeshPlanner() {}
  virtual uint32_t makePlan(const geometry_msgs::msg::PoseStamped& start, const geometry_msgs::msg::PoseStamped& goal, double tolerance,
                            std::vector<geometry_msgs::msg::PoseStamped>& plan, double& cost, std::string& message) = 0;   // :71-73
  virtual bool cancel() = 0;                                                                                                // :80
  virtual bool initialize(const std::string& name, const std::shared_ptr<mesh_map::MeshMap>& mesh_map_ptr, const rclcpp::Node::SharedPtr& node) = 0;   // :88
protected:
  MeshPlanner() {}
};
}  // namespace mbf_mesh_core
