// stub of tf2_ros::Buffer::lookupTransform / tf2::TransformException / geometry_msgs::msg::TransformStamped
#pragma once
#include <stdexcept>
#include <string>
#include <rclcpp/rclcpp.hpp>
namespace geometry_msgs { namespace msg {
struct TransformStamped {
  struct { struct { double x = 0, y = 0, z = 0; } translation; struct { double x = 0, y = 0, z = 0, w = 1; } rotation; } transform;
};
} }
namespace tf2 { struct TransformException : std::runtime_error { using std::runtime_error::runtime_error; }; }
namespace tf2_ros {
struct Buffer {
  geometry_msgs::msg::TransformStamped lookupTransform(const std::string& target_frame, const std::string& source_frame,
                                                       const rclcpp::Time& time, const rclcpp::Duration& timeout) const;
};
}  // namespace tf2_ros
