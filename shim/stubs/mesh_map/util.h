// stub of mesh_map/util.h:94,122
#pragma once
#include <mesh_map/mesh_map.h>
namespace mesh_map {
Vector toVector(const geometry_msgs::msg::Point& point);
geometry_msgs::msg::Pose calculatePoseFromPosition(const Vector& current, const Vector& next, const Normal& normal, float& cost);
}
