// srrg2b_flatten.h -- what every adapter needs between the reference's types and the C ABI of include/srrg2b.h:
// Eigen isometries are column-major, the ABI takes row-major (d+1)x(d+1); point clouds are AoS vectors of
// PointNormal{2,3}f, the ABI takes flat fp32 arrays + a validity mask.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <srrg2b.h>

namespace srrg2b_adapters {

template <typename Isometry>
inline void to_row_major(const Isometry& T, float* out /* (Dim+1)^2, zero-padded to 16 by the caller if needed */) {
  constexpr int N = Isometry::Dim + 1;
  const auto& M = T.matrix();
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) out[r * N + c] = M(r, c);
}
template <typename Isometry>
inline Isometry from_row_major(const float* in) {
  constexpr int N = Isometry::Dim + 1;
  Isometry T = Isometry::Identity();
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) T.matrix()(r, c) = in[r * N + c];
  return T;
}
template <typename Isometry>
inline void embed16(const Isometry& T, float* out16) {  // srrg2b_slice carries 16 floats whatever the dimension
  std::memset(out16, 0, 16 * sizeof(float));
  to_row_major(T, out16);
}

// PointNormal{2,3}fVectorCloud -> flat arrays (kept alive by the caller until srrg2b_set_cloud returns)
struct FlatCloud {
  std::vector<float> coords, normals;
  std::vector<uint8_t> valid;
  srrg2b_cloud describe() const {
    srrg2b_cloud d;
    std::memset(&d, 0, sizeof(d));
    d.coords = coords.data();
    d.normals = normals.data();
    d.valid = valid.data();
    d.n = (int64_t) valid.size();
    return d;
  }
};
template <typename Cloud>
inline FlatCloud flatten(const Cloud& cloud) {
  using Point = typename Cloud::value_type;
  constexpr int D = Point::Dim;
  FlatCloud f;
  f.coords.resize(cloud.size() * D);
  f.normals.resize(cloud.size() * D);
  f.valid.resize(cloud.size());
  for (size_t i = 0; i < cloud.size(); ++i) {
    for (int k = 0; k < D; ++k) {
      f.coords[i * D + k] = cloud[i].coordinates()[k];
      f.normals[i * D + k] = cloud[i].normal()[k];
    }
    f.valid[i] = cloud[i].status == srrg2_core::Valid ? 1 : 0;
  }
  return f;
}

inline void check(srrg2b_ctx* ctx, int rc, const char* who) {
  // configuration errors throw, like the reference (R/registration/aligners/aligner_slice_processor_impl.cpp:13-16);
  // numeric outcomes travel as status values
  if (rc != SRRG2B_OK) throw std::runtime_error(std::string(who) + "|" + (ctx ? srrg2b_last_error(ctx) : "no context"));
}

}  // namespace srrg2b_adapters
