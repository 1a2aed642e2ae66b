// stub of the rclcpp declarations the plugins touch (Node parameters, logging, time)
#pragma once
#include <cstddef>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
namespace rclcpp {
struct Time {};
struct Duration { static Duration from_seconds(double s); double seconds() const; };
struct QoS { explicit QoS(size_t depth); QoS& reliable(); QoS& best_effort(); };
struct SubscriptionBase { typedef std::shared_ptr<SubscriptionBase> SharedPtr; };
struct Logger {};
inline Logger get_logger(const std::string&) { return Logger(); }
struct Parameter { std::vector<std::string> as_string_array() const; std::string as_string() const; double as_double() const; std::string get_name() const; };
struct Node {
  typedef std::shared_ptr<Node> SharedPtr;
  template <class T> T declare_parameter(const std::string& name, const T& default_value);
  Parameter get_parameter(const std::string& name) const;
  Logger get_logger() const; Time now() const;
  template <class MsgT, class CallbackT> std::shared_ptr<SubscriptionBase> create_subscription(const std::string& topic, const QoS& qos, CallbackT&& callback)
  { (void)topic; (void)qos; (void)callback; return std::shared_ptr<SubscriptionBase>(); }
};
}  // namespace rclcpp
#define RCLCPP_ERROR_STREAM(logger, args) do { std::ostringstream _s; _s << args; (void)logger; } while (0)
#define RCLCPP_INFO_STREAM(logger, args) do { std::ostringstream _s; _s << args; (void)logger; } while (0)
#define RCLCPP_ERROR(logger, ...) do { (void)logger; } while (0)
