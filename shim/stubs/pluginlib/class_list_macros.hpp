// stub: the registration macro only has to name two existing types with the right base relationship
#pragma once
#include <type_traits>
#define PLUGINLIB_EXPORT_CLASS(Derived, Base) static_assert(std::is_base_of<Base, Derived>::value && !std::is_abstract<Derived>::value, "plugin class must implement its base interface");
