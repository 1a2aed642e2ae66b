// stub of mesh_map::MeshMap -- the members the plugins use, with the reference declarations' lines
// (mesh_map/include/mesh_map/mesh_map.h)
#pragma once
#include <memory>
#include <set>
#include <string>
#include <lvr2/lvr2_stub.hpp>
#include <rclcpp/rclcpp.hpp>
namespace std_msgs { namespace msg { struct Header { rclcpp::Time stamp; std::string frame_id; }; } }
#include <tf2_ros/buffer.h>
namespace geometry_msgs { namespace msg {
struct Point { double x = 0, y = 0, z = 0; }; struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; }; struct PoseStamped { std_msgs::msg::Header header; Pose pose; };
} }
namespace mesh_map {
typedef lvr2::BaseVector<float> Vector;      // mesh_map/util.h
typedef lvr2::Normal<float> Normal;
class AbstractLayer;
class MeshMap {
public:
  static const std::string MESH_MAP_NAMESPACE;                                                    // mesh_map.h:73
  lvr2::OptionalVertexHandle getNearestVertexHandle(const mesh_map::Vector& pos);                 // :97
  lvr2::OptionalFaceHandle getContainingFace(Vector& position, const float& max_dist);           // :114
  geometry_msgs::msg::PoseStamped transformToMapFrame(const geometry_msgs::msg::PoseStamped& pose);   // :262
  std::shared_ptr<lvr2::PMPMesh<Vector>> mesh();                                                  // :276
  const lvr2::DenseVertexMap<float>& vertexCosts();                                               // :292
  const std::string& mapFrame() const;                                                            // :300
  const tf2_ros::Buffer& tf2Buffer() const;                                                       // :308
  const lvr2::DenseFaceMap<Normal>& faceNormals();                                                // :326
  const lvr2::DenseVertexMap<Normal>& vertexNormals();                                            // :334
  const lvr2::DenseEdgeMap<float>& edgeWeights();                                                 // :342
  const lvr2::DenseEdgeMap<float>& edgeDistances();                                               // :350
  void publishVectorField(const std::string& name, const lvr2::DenseVertexMap<Vector>& vector_map, const bool publish_face_vectors = false);   // :388
  std::shared_ptr<AbstractLayer> layer(const std::string& layer_name);                            // :430
  lvr2::DenseVertexMap<bool> invalid;                                                             // :447
  double edge_cost_factor;                                                                        // :516
};
}  // namespace mesh_map
