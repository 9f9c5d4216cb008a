// Host (double precision) instantiation of eve_b200/csrc/gaze_math.cuh for tests/test_host_math.py.
#include "../../eve_b200/csrc/gaze_math.cuh"

using namespace eve::gm;

extern "C" {
void hm_combined_gaze(int n, const double* o, const double* pog, const double* R, const double* cam,
                      double* g) {
  for (int i = 0; i < n; ++i) combined_gaze(o + 3 * i, pog + 2 * i, R + 9 * i, cam + 16 * i, g + 2 * i);
}
void hm_combined_gaze_vjp(int n, const double* o, const double* pog, const double* R,
                          const double* cam, const double* gg, double* dpog) {
  for (int i = 0; i < n; ++i) {
    dpog[2 * i] = dpog[2 * i + 1] = 0.0;
    combined_gaze_vjp(o + 3 * i, pog + 2 * i, R + 9 * i, cam + 16 * i, gg + 2 * i, dpog + 2 * i);
  }
}
void hm_offset_aug(int n, const double* g, const double* R, const double* kappa, int inverse,
                   double* out) {
  for (int i = 0; i < n; ++i) {
    OffsetAugMid<double> m;
    offset_augmentation(g + 2 * i, R + 9 * i, kappa + 2 * i, inverse != 0, out + 2 * i, m);
  }
}
void hm_offset_aug_vjp(int n, const double* g, const double* R, const double* kappa, int inverse,
                       const double* gout, double* dg) {
  for (int i = 0; i < n; ++i) {
    dg[2 * i] = dg[2 * i + 1] = 0.0;
    offset_augmentation_vjp(g + 2 * i, R + 9 * i, kappa + 2 * i, inverse != 0, gout + 2 * i, dg + 2 * i);
  }
}
void hm_angular(int n, const double* a, const double* b, double lo, double hi, double* out) {
  for (int i = 0; i < n; ++i) out[i] = angular_error_deg(a + 2 * i, b + 2 * i, lo, hi);
}
void hm_angular_vjp(int n, const double* a, const double* b, double lo, double hi, const double* gl,
                    double* da) {
  for (int i = 0; i < n; ++i) {
    da[2 * i] = da[2 * i + 1] = 0.0;
    angular_error_deg_vjp(a + 2 * i, b + 2 * i, lo, hi, gl[i], da + 2 * i);
  }
}
}
