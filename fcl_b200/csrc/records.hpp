// records.hpp — host-side packing of the single-precision steering records (shared by the
// upload step and by the host build of the device math used in the CPU tests).
#pragma once
#include <cfloat>
#include <cmath>

#include "bounds_f32.cuh"

namespace fclgpu {

#ifndef FCLGPU_HD
#define FCLGPU_HD __host__ __device__ inline
#endif

FCLGPU_HD float round_up_f32(double x) {  // smallest float >= x
#ifdef __CUDA_ARCH__
  return __double2float_ru(x);
#else
  float f = (float)x;
  if ((double)f < x) f = std::nextafterf(f, FLT_MAX);
  return f;
#endif
}

// axis9 row-major, To = rectangle corner, l[2], r  ->  centred single-precision record
FCLGPU_HD void pack_rss32(const double* axis9, const double* To, const double* l, double r, RssRec32& o) {
  for (int k = 0; k < 9; ++k) o.a[k] = (float)axis9[k];
  double s = 0;
  for (int k = 0; k < 3; ++k) {
    const double c = To[k] + 0.5 * l[0] * axis9[3 * k] + 0.5 * l[1] * axis9[3 * k + 1];
    o.c[k] = (float)c;
    s += fabs(To[k]);
  }
  o.h0 = round_up_f32(0.5 * l[0]);
  o.h1 = round_up_f32(0.5 * l[1]);
  o.r = round_up_f32(r);
  o.s = round_up_f32(s + l[0] + l[1] + r);
}

FCLGPU_HD void pack_obb32(const double* axis9, const double* To, const double* ext, ObbRec32& o) {
  for (int k = 0; k < 9; ++k) o.a[k] = (float)axis9[k];
  double s = 0;
  for (int k = 0; k < 3; ++k) {
    o.c[k] = (float)To[k];
    o.e[k] = round_up_f32(ext[k]);
    s += fabs(To[k]) + ext[k];
  }
  o.s = round_up_f32(s);
}

FCLGPU_HD void pack_pose32(const double* R9, const double* T3, float* R0, float* T0, float& t_l1) {
  for (int k = 0; k < 9; ++k) R0[k] = (float)R9[k];
  double s = 0;
  for (int k = 0; k < 3; ++k) {
    T0[k] = (float)T3[k];
    s += fabs(T3[k]);
  }
  t_l1 = round_up_f32(s);
}

}  // namespace fclgpu
