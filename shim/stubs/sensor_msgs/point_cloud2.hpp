// stub of sensor_msgs::msg::PointCloud2 and sensor_msgs::PointCloud2ConstIterator (sensor_msgs/point_cloud2_iterator.hpp)
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <rclcpp/rclcpp.hpp>
namespace sensor_msgs {
namespace msg {
struct PointCloud2 {
  typedef std::shared_ptr<const PointCloud2> ConstSharedPtr;
  struct { std::string frame_id; rclcpp::Time stamp; } header;
  uint32_t width = 0, height = 0;
};
}  // namespace msg
template <class T> struct PointCloud2ConstIterator {
  PointCloud2ConstIterator(const msg::PointCloud2& cloud, const std::string& field);
  const T& operator*() const; PointCloud2ConstIterator& operator++(); PointCloud2ConstIterator end() const;
  bool operator!=(const PointCloud2ConstIterator& other) const;
};
}  // namespace sensor_msgs
