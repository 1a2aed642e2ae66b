// Per-face wavefront update rules, usable from device code and (for the
// host-side semantics simulator under tools/) from plain C++.
//
//   cvp_update      : CVPMeshPlanner::waveFrontUpdate, default variant
//                     (reference cvp_mesh_planner/src/cvp_mesh_planner.cpp:369-556)
//   sethian_update  : InflationLayer::computeUpdateSethianMethod
//                     (reference mesh_layers/src/inflation_layer.cpp:181-234)
//   fading          : InflationLayer::fading (inflation_layer.cpp:315-339)
//
// Both keep the reference's evaluation order and precision (double inside the
// CVP update with float stores; float throughout for the Sethian update) so
// that branch decisions match.  Compile device code with -fmad=false.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MNB_HD __host__ __device__ __forceinline__
#else
#define MNB_HD inline
#endif

namespace mnb {

struct CvpResult {
  float value;      // new potential of v3 (float store of the double result)
  float direction;  // theta stored in direction_[v3]
  int pred_sel;     // 1 -> predecessor v1, 2 -> predecessor v2
};

// u1,u2 = potentials of v1,v2 ; u3 = current potential of v3 (may be +inf)
// c = |v1v2|, b = |v1v3|, a = |v2v3| as edge *weights* (cvp:380-390).
// Returns true iff the reference would have written distances[v3].
//
// ANGLES = true : literal restatement, three acos calls (cvp:456-458), fills r.direction.
// ANGLES = false: the wavefront hot loop.  theta_i = acos(t_i) is strictly decreasing, so the
//   reference's comparisons  theta1 < theta0, theta2 < theta0, theta1 < theta2  are evaluated on the
//   cosines (acos_less below) -- except where the two cosines are so close that their ROUNDED angles
//   may coincide: |d acos| >= |dx|, and angles in (0, pi] are spaced up to 4.4e-16 apart, so cosines
//   more than that apart have distinct, ordered angles; closer ones (a right angle at v3 puts t0a at
//   ~0, where cosines are spaced 1e-300 apart but all map to pi/2) take the literal acos comparison.
//   r.direction is NOT filled; the caller re-evaluates the winning face once with ANGLES = true when
//   it stores the result.
MNB_HD bool acos_less(double x, double y) {          // acos(x) < acos(y) as the reference's doubles decide it; |x|, |y| <= 1 or NaN
  if (fabs(x - y) > 1e-15) return x > y;
  return acos(x) < acos(y);
}
template <bool ANGLES>
MNB_HD bool cvp_update_t(double u1, double u2, double u3, double a, double b, double c, CvpResult& r) {
  const double c_sq = c * c, b_sq = b * b, a_sq = a * a;
  const double u1_sq = u1 * u1, u2_sq = u2 * u2;
  const double sx = (c_sq + u1_sq - u2_sq) / (2 * c);
  const double sy = -sqrt(fmax(u1_sq - sx * sx, 0.0));
  const double p = (b_sq + c_sq - a_sq) / (2 * c);
  const double hc = sqrt(fmax(b_sq - p * p, 0.0));
  const double dy = hc - sy;
  const double dx = p - sx;
  const double u3tmp_sq = dx * dx + dy * dy;
  double u3tmp = sqrt(u3tmp_sq);
  if (!(u3tmp < u3)) return false;
  const double t0a = (a_sq + b_sq - c_sq) / (2 * a * b);
  const double t1a = (u3tmp_sq + b_sq - u1_sq) / (2 * u3tmp * b);
  const double t2a = (a_sq + u3tmp_sq - u2_sq) / (2 * a * u3tmp);
  int edge_fallback = 0;  // 1: u1 + b via v1, 2: u2 + a via v2
  if (fabs(t1a) > 1) {
    edge_fallback = 1;
  } else if (fabs(t2a) > 1) {
    edge_fallback = 2;
  } else if (ANGLES) {
    const double theta0 = acos(t0a);
    const double theta1 = acos(t1a);
    const double theta2 = acos(t2a);
    if (theta1 < theta0 && theta2 < theta0) {
      r.value = (float)u3tmp;
      if (theta1 < theta2) { r.pred_sel = 1; r.direction = (float)theta1; }
      else { r.pred_sel = 2; r.direction = (float)(-theta2); }
      return true;
    }
    edge_fallback = (theta1 < theta2) ? 1 : 2;
  } else {
    // NaN cosines (degenerate triangles) make every acos comparison false in the reference:
    // '>' on NaN is false as well, so the same branch (u2 + a) is taken.  The reference guards
    // |t1a|, |t2a| <= 1 but NOT t0a: with cost-weighted (non-geometric) edge weights |t0a| can exceed
    // 1, acos(t0a) is NaN and both comparisons against theta0 are false.
    if (fabs(t0a) <= 1 && acos_less(t1a, t0a) && acos_less(t2a, t0a)) {
      r.value = (float)u3tmp;
      r.pred_sel = acos_less(t1a, t2a) ? 1 : 2;
      r.direction = 0.0f;
      return true;
    }
    edge_fallback = acos_less(t1a, t2a) ? 1 : 2;
  }
  u3tmp = (edge_fallback == 1) ? (u1 + b) : (u2 + a);
  if (!(u3tmp < u3)) return false;
  r.value = (float)u3tmp;
  r.pred_sel = edge_fallback;
  r.direction = 0.0f;
  return true;
}

MNB_HD bool cvp_update(double u1, double u2, double u3, double a, double b, double c, CvpResult& r) {
  return cvp_update_t<true>(u1, u2, u3, a, b, c, r);
}

// d1,d2 = distances of v1,v2 ; a = |v2v3| ; b = |v1v3| ; dot = cos of the angle at v3.
MNB_HD float sethian_update(float d1, float d2, float a, float b, float dot, float F) {
  const float INF = __builtin_huge_valf();
  float t = INF;
  const float r_cos_angle = dot;
  const float r_sin_angle = sqrtf(1 - dot * dot);
  const float u = d2 - d1;
  const float f2 = a * a + b * b - 2 * a * b * r_cos_angle;
  const float f1 = b * u * (a * r_cos_angle - b);
  const float f0 = b * b * (u * u - F * F * a * a * r_sin_angle);  // `sin`, not sin^2: literal (inflation_layer.cpp:195)
  const float delta = f1 * f1 - f0 * f2;
  if (delta >= 0) {
    if (fabsf(f2) > 1e-9f) {
      t = (-f1 - sqrtf(delta)) / f2;
      if (t < u || b * (t - u) / t < a * r_cos_angle || a / r_cos_angle < b * (t - u) / 2) {
        t = (-f1 + sqrtf(delta)) / f2;
      } else {
        if (f1 != 0) t = -f0 / f1; else t = -INF;
      }
    }
  } else {
    t = -INF;
  }
  if (u < t && a * r_cos_angle < b * (t - u) / t && b * (t - u) / t < a / r_cos_angle) return t + d1;
  return fminf(b * F + d1, a * F + d2);
}

// InflationLayer::waveFrontUpdate's distance part (inflation_layer.cpp:248-275, 297-311).
// Returns the candidate (finite) or +inf when the reference returns early.
MNB_HD float inflation_candidate(float u1, float u2, float a, float b, float c) {
  const float dot = (a * a + b * b - c * c) / (2 * a * b);
  const float u3tmp = sethian_update(u1, u2, a, b, dot, 1.0f);
  return (fabsf(u3tmp) <= 3.402823466e+38f) ? u3tmp : __builtin_huge_valf();  // isfinite
}

struct InflationParams {  // reference config members are doubles (inflation_layer.h:240-248)
  double inscribed_radius, inflation_radius, lethal_value, inscribed_value, cost_scaling_factor;
};

MNB_HD float fading(const InflationParams& cfg, float distance) {
  if (distance > cfg.inflation_radius) return 0.0f;
  if (distance > cfg.inscribed_radius) {
    const float factor = (float)exp(-1.0 * cfg.cost_scaling_factor * (distance - cfg.inscribed_radius));
    return (float)(cfg.inscribed_value * factor);
  }
  if (distance > 0) return (float)cfg.inscribed_value;
  return (float)cfg.lethal_value;
}

// order-preserving float <-> uint for non-negative floats and +inf
MNB_HD uint32_t f2u(float f) { union { float f; uint32_t u; } x; x.f = f; return x.u; }
MNB_HD float u2f(uint32_t u) { union { float f; uint32_t u; } x; x.u = u; return x.f; }

}  // namespace mnb
