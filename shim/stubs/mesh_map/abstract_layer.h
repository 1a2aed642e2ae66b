// stub of mesh_map::AbstractLayer (mesh_map/include/mesh_map/abstract_layer.h:55-280): the virtual interface verbatim,
// the helpers the layers call
#pragma once
#include <array>
#include <functional>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <string>
#include <boost/optional.hpp>
#include <mesh_map/mesh_map.h>
namespace mesh_map {
typedef std::function<void(const std::string&, const rclcpp::Time&, const std::set<lvr2::VertexHandle>&)> notify_func;
class AbstractLayer {
public:
  typedef std::shared_ptr<mesh_map::AbstractLayer> Ptr;
  virtual ~AbstractLayer() {}
  virtual bool readLayer() = 0;                                                                   // :64
  virtual bool writeLayer() = 0;                                                                  // :70
  virtual float defaultValue() = 0;                                                               // :76
  virtual float threshold() = 0;                                                                  // :82
  virtual bool computeLayer() = 0;                                                                // :88
  virtual const lvr2::VertexMap<float>& costs() = 0;                                              // :94
  virtual const std::set<lvr2::VertexHandle>& lethals() = 0;                                      // :100
  virtual void onInputChanged(const rclcpp::Time& timestamp, const std::set<lvr2::VertexHandle>& changed) { (void)timestamp; (void)changed; }   // :111
  virtual Vector vectorAt(const std::array<lvr2::VertexHandle, 3>& vertices, const std::array<float, 3>& barycentric_coords) { (void)vertices; (void)barycentric_coords; return Vector(); }   // :127
  virtual const boost::optional<lvr2::VertexMap<Vector>&> vectorMap() { return boost::none; }    // :139
  virtual Vector vectorAt(const lvr2::VertexHandle& vertex) { (void)vertex; return Vector(); }   // :149
  bool initialize(const std::string& name, const notify_func notify_update, std::shared_ptr<mesh_map::MeshMap> map, const rclcpp::Node::SharedPtr node);   // :160
  std::shared_lock<std::shared_mutex> readLock() { return std::shared_lock<std::shared_mutex>(mutex_); }   // :171
protected:
  void notifyChange(const rclcpp::Time& timestamp, const std::set<lvr2::VertexHandle>& changed) { notify_(layer_name_, timestamp, changed); }   // :216
  std::unique_lock<std::shared_mutex> writeLock() { return std::unique_lock<std::shared_mutex>(mutex_); }   // :226
  virtual bool initialize() = 0;                                                                  // :234
  const rclcpp::Logger& get_logger() const { return logger_; }                                    // :239
  std::string layer_name_;
  std::weak_ptr<mesh_map::MeshMap> map_ptr_;
  rclcpp::Node::SharedPtr node_;
private:
  notify_func notify_;
  std::shared_mutex mutex_;
  rclcpp::Logger logger_;
};
}  // namespace mesh_map
