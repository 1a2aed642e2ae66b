// stub of the lvr2 types the reference's plugin interfaces expose (lvr2 is an un-vendored dependency, pinned only as
// `version: main` in source_dependencies.yaml:4-7): handles, attribute maps, PMPMesh iteration.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <vector>
#include <boost/optional.hpp>
namespace lvr2 {
template <class T> struct BaseVector { T x{}, y{}, z{}; BaseVector() {} BaseVector(T a, T b, T c) : x(a), y(b), z(c) {}
  BaseVector operator*(double s) const { return BaseVector((T)(x * s), (T)(y * s), (T)(z * s)); } BaseVector operator/(double s) const { return (*this) * (1.0 / s); } };
template <class T> using Normal = BaseVector<T>;
struct Handle { uint32_t i = 0; Handle() {} explicit Handle(uint32_t v) : i(v) {} uint32_t idx() const { return i; } bool operator<(const Handle& o) const { return i < o.i; } bool operator==(const Handle& o) const { return i == o.i; } };
struct VertexHandle : Handle { using Handle::Handle; };
struct FaceHandle : Handle { using Handle::Handle; };
struct EdgeHandle : Handle { using Handle::Handle; };
template <class H> struct OptionalHandle { bool has = false; H h; explicit operator bool() const { return has; } H unwrap() const { return h; } };
using OptionalVertexHandle = OptionalHandle<VertexHandle>; using OptionalFaceHandle = OptionalHandle<FaceHandle>;
template <class H, class V> struct AttributeMap {
  std::map<uint32_t, V> m;
  struct It { typename std::map<uint32_t, V>::const_iterator it; H operator*() const { return H(it->first); } It& operator++() { ++it; return *this; } bool operator!=(const It& o) const { return it != o.it; } };
  It begin() const { return It{m.begin()}; } It end() const { return It{m.end()}; }
  void insert(H h, const V& v) { m[h.idx()] = v; } void clear() { m.clear(); } size_t numValues() const { return m.size(); }
  const V& operator[](H h) const { return m.at(h.idx()); } V& operator[](H h) { return m[h.idx()]; }
  boost::optional<V> get(H h) const { auto f = m.find(h.idx()); return f == m.end() ? boost::optional<V>() : boost::optional<V>(f->second); }
  bool containsKey(H h) const { return m.count(h.idx()) != 0; }
};
template <class V> using VertexMap = AttributeMap<VertexHandle, V>;
template <class V> using DenseVertexMap = AttributeMap<VertexHandle, V>;
template <class V> using SparseVertexMap = AttributeMap<VertexHandle, V>;
template <class V> using DenseEdgeMap = AttributeMap<EdgeHandle, V>;
template <class V> using DenseFaceMap = AttributeMap<FaceHandle, V>;
template <class Vec> struct PMPMesh {
  template <class H> struct Range { std::vector<H> v; typename std::vector<H>::const_iterator begin() const { return v.begin(); } typename std::vector<H>::const_iterator end() const { return v.end(); } };
  Range<VertexHandle> vertices() const; Range<FaceHandle> faces() const; Range<EdgeHandle> edges() const;
  size_t nextVertexIndex() const; bool containsVertex(VertexHandle) const;
  Vec getVertexPosition(VertexHandle) const;
  std::array<VertexHandle, 3> getVerticesOfFace(FaceHandle) const;
  std::array<VertexHandle, 2> getVerticesOfEdge(EdgeHandle) const;
  OptionalFaceHandle getFaceBetween(VertexHandle, VertexHandle, VertexHandle) const;
};
template <class Vec> DenseVertexMap<float> calcNormalClearance(const PMPMesh<Vec>& mesh, const DenseVertexMap<Normal<float>>& normals);   // clearance_layer.cpp:161
}  // namespace lvr2
