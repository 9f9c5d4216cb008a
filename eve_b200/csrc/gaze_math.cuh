// Gaze geometry of the EVE hot path as scalar-templated host/device functions: the value and the
// hand-derived vector-Jacobian products of
//   pitchyaw_to_vector (common.py:32-40), vector_to_pitchyaw (:43-54), pitchyaw_to_rotation
//   (:57-76), calculate_combined_gaze_direction (:129-146), apply_offset_augmentation (:182-218)
// and of the angular loss (losses/angular.py:29-38).  The kernels in geometry.cu / losses.cu
// instantiate them with float; tests/test_host_math.py compiles this header for the host with
// double and checks every function and every VJP against torch autograd (no GPU needed).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EVE_HD __host__ __device__ __forceinline__
#else
#define EVE_HD inline
#endif

namespace eve {
namespace gm {

template <typename T>
EVE_HD void pitchyaw_to_vector(const T* a, T* v) {
  const T sp = sin(a[0]), cp = cos(a[0]), sy = sin(a[1]), cy = cos(a[1]);
  v[0] = cp * sy;
  v[1] = sp;
  v[2] = cp * cy;
}
// da (+)= J^T gv
template <typename T>
EVE_HD void pitchyaw_to_vector_vjp(const T* a, const T* gv, T* da) {
  const T sp = sin(a[0]), cp = cos(a[0]), sy = sin(a[1]), cy = cos(a[1]);
  da[0] += gv[0] * (-sp * sy) + gv[1] * cp + gv[2] * (-sp * cy);
  da[1] += gv[0] * (cp * cy) + gv[2] * (-cp * sy);
}

// n = v / (|v| + 1e-7); [asin(n1), atan2(n0, n2)]
template <typename T>
EVE_HD void vector_to_pitchyaw(const T* v, T* a) {
  const T nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const T s = nrm + (T)1e-7;
  a[0] = asin(v[1] / s);
  a[1] = atan2(v[0] / s, v[2] / s);
}
template <typename T>
EVE_HD void vector_to_pitchyaw_vjp(const T* v, const T* ga, T* dv) {
  const T nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const T s = nrm + (T)1e-7;
  const T n0 = v[0] / s, n1 = v[1] / s, n2 = v[2] / s;
  const T r2 = n0 * n0 + n2 * n2;
  T dn[3];
  dn[0] = ga[1] * n2 / r2;
  dn[1] = ga[0] / sqrt((T)1 - n1 * n1);
  dn[2] = -ga[1] * n0 / r2;
  // n_i = v_i / s, ds/dv_j = v_j / |v|
  const T dot = (dn[0] * v[0] + dn[1] * v[1] + dn[2] * v[2]) / (s * s * nrm);
  dv[0] += dn[0] / s - dot * v[0];
  dv[1] += dn[1] / s - dot * v[1];
  dv[2] += dn[2] / s - dot * v[2];
}

// w = (Ry(yaw) Rx(pitch)) k  with the reference's matrices (common.py:57-76)
template <typename T>
EVE_HD void rotate_by_pitchyaw(const T* a, const T* k, T* w) {
  const T s0 = sin(a[0]), c0 = cos(a[0]), s1 = sin(a[1]), c1 = cos(a[1]);
  w[0] = c1 * k[0] - s1 * s0 * k[1] + s1 * c0 * k[2];
  w[1] = c0 * k[1] + s0 * k[2];
  w[2] = -s1 * k[0] - c1 * s0 * k[1] + c1 * c0 * k[2];
}
template <typename T>
EVE_HD void rotate_by_pitchyaw_vjp(const T* a, const T* k, const T* gw, T* da) {
  const T s0 = sin(a[0]), c0 = cos(a[0]), s1 = sin(a[1]), c1 = cos(a[1]);
  const T w0 = c1 * k[0] - s1 * s0 * k[1] + s1 * c0 * k[2];
  const T w2 = -s1 * k[0] - c1 * s0 * k[1] + c1 * c0 * k[2];
  da[0] += gw[0] * (-s1 * c0 * k[1] - s1 * s0 * k[2]) + gw[1] * (-s0 * k[1] + c0 * k[2]) +
           gw[2] * (-c1 * c0 * k[1] - c1 * s0 * k[2]);
  da[1] += gw[0] * w2 - gw[2] * w0;
}

// common.py:129-146: g = vector_to_pitchyaw(-(R_head (cam . [pog, 0, 1] - origin)))
template <typename T>
EVE_HD void combined_gaze(const T* origin, const T* pog_mm, const T* R, const T* cam, T* g) {
  T p3[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    p3[i] = cam[i * 4 + 0] * pog_mm[0] + cam[i * 4 + 1] * pog_mm[1] + cam[i * 4 + 3] - origin[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = -(R[i * 3 + 0] * p3[0] + R[i * 3 + 1] * p3[1] + R[i * 3 + 2] * p3[2]);
  vector_to_pitchyaw(d, g);
}
template <typename T>
EVE_HD void combined_gaze_vjp(const T* origin, const T* pog_mm, const T* R, const T* cam,
                              const T* gg, T* dpog) {
  T p3[3], d[3], gd[3] = {(T)0, (T)0, (T)0}, gp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    p3[i] = cam[i * 4 + 0] * pog_mm[0] + cam[i * 4 + 1] * pog_mm[1] + cam[i * 4 + 3] - origin[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = -(R[i * 3 + 0] * p3[0] + R[i * 3 + 1] * p3[1] + R[i * 3 + 2] * p3[2]);
  vector_to_pitchyaw_vjp(d, gg, gd);
#pragma unroll
  for (int k = 0; k < 3; ++k) gp[k] = -(R[0 * 3 + k] * gd[0] + R[1 * 3 + k] * gd[1] + R[2 * 3 + k] * gd[2]);
  dpog[0] += cam[0 * 4 + 0] * gp[0] + cam[1 * 4 + 0] * gp[1] + cam[2 * 4 + 0] * gp[2];
  dpog[1] += cam[0 * 4 + 1] * gp[0] + cam[1 * 4 + 1] * gp[1] + cam[2 * 4 + 1] * gp[2];
}

// common.py:182-218.  Intermediates kept for the VJP.
template <typename T>
struct OffsetAugMid {
  T d1[3];     // -R^T (-pitchyaw_to_vector(g))
  T a1[2];     // vector_to_pitchyaw(d1)
  T kv[3];     // kappa vector (sign-flipped when inverse)
  T d3[3];     // -R (-(rot kv))
};
template <typename T>
EVE_HD void offset_augmentation(const T* g, const T* R, const T* kappa, bool inverse, T* out,
                                OffsetAugMid<T>& m) {
  T v[3], w[3];
  pitchyaw_to_vector(g, v);
  // d0 = -v; d1 = -(R^T d0) = R^T v
#pragma unroll
  for (int k = 0; k < 3; ++k) m.d1[k] = R[0 * 3 + k] * v[0] + R[1 * 3 + k] * v[1] + R[2 * 3 + k] * v[2];
  pitchyaw_to_vector(kappa, m.kv);
  if (inverse) {
    m.kv[0] = -m.kv[0];
    m.kv[1] = -m.kv[1];
  }
  vector_to_pitchyaw(m.d1, m.a1);
  rotate_by_pitchyaw(m.a1, m.kv, w);
  // d2 = -w; d3 = -(R d2) = R w
#pragma unroll
  for (int i = 0; i < 3; ++i) m.d3[i] = R[i * 3 + 0] * w[0] + R[i * 3 + 1] * w[1] + R[i * 3 + 2] * w[2];
  vector_to_pitchyaw(m.d3, out);
}
template <typename T>
EVE_HD void offset_augmentation_vjp(const T* g, const T* R, const T* kappa, bool inverse,
                                    const T* gout, T* dg) {
  OffsetAugMid<T> m;
  T out[2];
  offset_augmentation(g, R, kappa, inverse, out, m);
  T gd3[3] = {(T)0, (T)0, (T)0}, gw[3], ga1[2] = {(T)0, (T)0}, gd1[3] = {(T)0, (T)0, (T)0}, gv[3];
  vector_to_pitchyaw_vjp(m.d3, gout, gd3);
#pragma unroll
  for (int k = 0; k < 3; ++k) gw[k] = R[0 * 3 + k] * gd3[0] + R[1 * 3 + k] * gd3[1] + R[2 * 3 + k] * gd3[2];
  rotate_by_pitchyaw_vjp(m.a1, m.kv, gw, ga1);
  vector_to_pitchyaw_vjp(m.d1, ga1, gd1);
#pragma unroll
  for (int i = 0; i < 3; ++i) gv[i] = R[i * 3 + 0] * gd1[0] + R[i * 3 + 1] * gd1[1] + R[i * 3 + 2] * gd1[2];
  pitchyaw_to_vector_vjp(g, gv, dg);
}

// losses/angular.py:29-38: acos(clamp(cos_sim(v(a), v(b)), -1+1e-8, 1-1e-8)) in degrees; both
// arguments are pitch/yaw pairs.  cos_sim normalises each vector by max(|v|, 1e-8).
template <typename T>
EVE_HD T angular_error_deg(const T* a, const T* b, T lo, T hi) {
  T va[3], vb[3];
  pitchyaw_to_vector(a, va);
  pitchyaw_to_vector(b, vb);
  T na = sqrt(va[0] * va[0] + va[1] * va[1] + va[2] * va[2]);
  T nb = sqrt(vb[0] * vb[0] + vb[1] * vb[1] + vb[2] * vb[2]);
  na = na > (T)1e-8 ? na : (T)1e-8;
  nb = nb > (T)1e-8 ? nb : (T)1e-8;
  T sim = (va[0] / na) * (vb[0] / nb) + (va[1] / na) * (vb[1] / nb) + (va[2] / na) * (vb[2] / nb);
  sim = sim < lo ? lo : (sim > hi ? hi : sim);
  return acos(sim) * (T)(180.0 / 3.14159265358979323846);
}
// da (+)= gl * d(angular_error_deg)/da
template <typename T>
EVE_HD void angular_error_deg_vjp(const T* a, const T* b, T lo, T hi, T gl, T* da) {
  T va[3], vb[3];
  pitchyaw_to_vector(a, va);
  pitchyaw_to_vector(b, vb);
  T na = sqrt(va[0] * va[0] + va[1] * va[1] + va[2] * va[2]);
  T nb = sqrt(vb[0] * vb[0] + vb[1] * vb[1] + vb[2] * vb[2]);
  na = na > (T)1e-8 ? na : (T)1e-8;
  nb = nb > (T)1e-8 ? nb : (T)1e-8;
  const T an[3] = {va[0] / na, va[1] / na, va[2] / na};
  const T bn[3] = {vb[0] / nb, vb[1] / nb, vb[2] / nb};
  const T sim = an[0] * bn[0] + an[1] * bn[1] + an[2] * bn[2];
  if (sim < lo || sim > hi) return;                      // clamp: zero gradient outside
  const T gs = -gl * (T)(180.0 / 3.14159265358979323846) / sqrt((T)1 - sim * sim);
  T gv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) gv[i] = gs * (bn[i] - sim * an[i]) / na;
  pitchyaw_to_vector_vjp(a, gv, da);
}

}  // namespace gm
}  // namespace eve
