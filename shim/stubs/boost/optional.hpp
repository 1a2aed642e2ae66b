// stub of boost::optional (only what mesh_map/abstract_layer.h:120 needs)
#pragma once
namespace boost {
struct none_t {}; static const none_t none{};
template <class T> class optional;
template <class T> class optional<T&> { T* p_ = nullptr; public: optional() {} optional(none_t) {} optional(T& r) : p_(&r) {} explicit operator bool() const { return p_; } T& get() const { return *p_; } };
template <class T> class optional { bool has_ = false; T v_{}; public: optional() {} optional(none_t) {} optional(const T& v) : has_(true), v_(v) {} explicit operator bool() const { return has_; } const T& get() const { return v_; } };
}
