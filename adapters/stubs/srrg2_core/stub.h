// MINIMAL STAND-INS for the srrg2_core symbols the adapters touch.  NOT upstream code: only
// the shapes the reference's own sources prove (file:line in the comments, R/ = srrg2_slam_interfaces/src/
// srrg2_slam_interfaces/), enough for `g++ -fsyntax-only` of adapters/*.h where srrg2_core is not installed.
// A maintainer compiles the adapters against the real headers instead (drop -I adapters/stubs).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace srrg2_core {

// PARAM(PropertyType, name, doc, default, flag*) -> member param_<name> with .value() / .setValue()
// (R/registration/aligners/aligner.h:30-35, multi_aligner.h:34-57)
template <typename T>
struct Property_ {
  T _v{};
  Property_() = default;
  explicit Property_(const T& v) : _v(v) {}
  const T& value() const { return _v; }
  T& value() { return _v; }
  void setValue(const T& v) { _v = v; }
};
using PropertyFloat = Property_<float>;
using PropertyInt = Property_<int>;
using PropertyBool = Property_<bool>;
using PropertyString = Property_<std::string>;
#define PARAM(TYPE, NAME, DOC, DEFAULT, FLAG) TYPE param_##NAME{DEFAULT}
#define BOSS_REGISTER_CLASS(CLASS) /* class registry of the configuration system: no-op in the stub build */

// PropertyConfigurable_<T>: .value() (shared_ptr), operator-> ; PropertyConfigurableVector_<T>: .size(), .value(i)
// (R/registration/aligners/multi_aligner_impl.cpp:10-12,29,59)
template <typename T>
struct PropertyConfigurable_ {
  std::shared_ptr<T> _p;
  const std::shared_ptr<T>& value() const { return _p; }
  void setValue(std::shared_ptr<T> p) { _p = std::move(p); }
  T* operator->() const { return _p.get(); }
};
template <typename T>
struct PropertyConfigurableVector_ {
  std::vector<std::shared_ptr<T>> _v;
  size_t size() const { return _v.size(); }
  const std::shared_ptr<T>& value(size_t i) const { return _v.at(i); }
  void pushBack(std::shared_ptr<T> p) { _v.push_back(std::move(p)); }
};
struct Configurable { virtual ~Configurable() = default; };

// PropertyContainerBase: the dynamic properties of a local map / a measurement; an aligner slice looks its cloud up by
// name (R/registration/aligners/aligner_slice_processor_base.h:160-168 bindSlice, R/mapping/local_map.h:15-31)
struct PropertyContainerBase {
  std::map<std::string, void*> _properties;
  template <typename T>
  T* property(const std::string& name) const {
    auto it = _properties.find(name);
    return it == _properties.end() ? nullptr : static_cast<T*>(it->second);
  }
};
using PropertyContainerDynamic = PropertyContainerBase;

// Correspondence(fixed_idx, moving_idx, response) (R/registration/loop_detector/multi_loop_detector_hbst_impl.cpp:183-191)
struct Correspondence {
  int fixed_idx = -1, moving_idx = -1;
  float response = 0.f;
  Correspondence() = default;
  Correspondence(int f, int m, float r) : fixed_idx(f), moving_idx(m), response(r) {}
};
using CorrespondenceVector = std::vector<Correspondence>;

// Isometry{2,3}f: an Eigen::Transform upstream (column-major .matrix()); here a plain column-major array with the
// members the adapters use: Identity(), matrix().data(), ::Dim (R/registration/aligners/aligner_slice_processor.h:161-167)
template <int D>
struct IsometryStub {
  static constexpr int Dim = D;
  struct Matrix {
    std::array<float, (D + 1) * (D + 1)> a{};
    const float* data() const { return a.data(); }
    float* data() { return a.data(); }
    float operator()(int r, int c) const { return a[(size_t) (c * (D + 1) + r)]; }
    float& operator()(int r, int c) { return a[(size_t) (c * (D + 1) + r)]; }
  } _m;
  static IsometryStub Identity() {
    IsometryStub T;
    for (int i = 0; i <= D; ++i) T._m(i, i) = 1.f;
    return T;
  }
  const Matrix& matrix() const { return _m; }
  Matrix& matrix() { return _m; }
  // rigid inverse and composition (Eigen::Transform::inverse / operator* upstream;
  // R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:115)
  IsometryStub inverse() const {
    IsometryStub T = Identity();
    for (int r = 0; r < D; ++r) {
      float t = 0.f;
      for (int c = 0; c < D; ++c) { T._m(r, c) = _m(c, r); t -= _m(c, r) * _m(c, D); }
      T._m(r, D) = t;
    }
    return T;
  }
  IsometryStub operator*(const IsometryStub& o) const {
    IsometryStub T;
    for (int r = 0; r <= D; ++r)
      for (int c = 0; c <= D; ++c) {
        float v = 0.f;
        for (int k = 0; k <= D; ++k) v += _m(r, k) * o._m(k, c);
        T._m(r, c) = v;
      }
    return T;
  }
};
using Isometry2f = IsometryStub<2>;
using Isometry3f = IsometryStub<3>;

// PointNormal{2,3}f: .coordinates(), .normal(), .status == Valid (R/mapping/merger_correspondence_homo_impl.cpp:35-39,62-74)
enum POINT_STATUS { Valid = 0, Invalid = 1 };
template <int D>
struct PointNormalStub {
  static constexpr int Dim = D;
  std::array<float, D> _c{}, _n{};
  POINT_STATUS status = Valid;
  const std::array<float, D>& coordinates() const { return _c; }
  const std::array<float, D>& normal() const { return _n; }
};
using PointNormal2fVectorCloud = std::vector<PointNormalStub<2>>;
using PointNormal3fVectorCloud = std::vector<PointNormalStub<3>>;

}  // namespace srrg2_core
