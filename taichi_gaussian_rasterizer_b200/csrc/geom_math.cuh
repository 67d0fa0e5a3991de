// geom_math.cuh — per-gaussian geometry shared by the projection and tile-mapper kernels.
//
// Everything here is __host__ __device__ and written as explicit binary IEEE operations in a
// fixed order: for float the results are the bit-exactness contract for visible sets, tile
// overlap lists and sort keys (include/gs_numeric.h; this translation unit is built with
// -fmad=false and the host image with -ffp-contract=off).  The host image is only reachable
// through the gs_selftest_* exports, which let the CPU test-suite detect an operation-order
// divergence from the oracle without a GPU; no product path calls it.
//
// Restates: perspective/projection.py:50-80, taichi_lib/generic.py:95-158,216-237,418-427,
// taichi_lib/grid_query.py:9-91 (paths relative to /root/reference/taichi_splatting/).
#pragma once

#include <math.h>

#include "../../include/gs_numeric.h"

namespace gs {

template <typename T> struct Math;
template <> struct Math<float> {
  GS_HD static float exp_(float x) { return gs_expf(x); }
  GS_HD static float log_(float x) { return gs_logf(x); }
  GS_HD static float sqrt_(float x) { return sqrtf(x); }
};
template <> struct Math<double> {
  GS_HD static double exp_(double x) { return exp(x); }
  GS_HD static double log_(double x) { return log(x); }
  GS_HD static double sqrt_(double x) { return sqrt(x); }
};

template <typename T>
struct CameraConst {
  T Tcw[12];         // rows 0..2 of T_camera_world (3x4)
  T fx, fy, cx, cy;  // projection
  T w, h;            // image size
  T near_, far_;
  T blur;
  T lo_x, lo_y, hi_x, hi_y;  // jacobian clamp bounds: -size*margin, (size-1)*(1+margin)
  T alpha_threshold;
};

template <typename T>
struct Projected {
  T mean_x, mean_y, axis_x, axis_y, sigma_x, sigma_y, alpha, z;
  bool in_view;
};

template <typename T>
GS_HD T clampv(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// Intermediate values the backward pass reuses.
template <typename T>
struct ProjectState {
  T qn;             // |q|
  T q[4];           // normalised quaternion x y z w
  T s[3];           // exp(log_scale)
  T cam[3];         // point in camera
  T tu, tv;         // clamped uv
  T u, v;
  T m[2][3];        // J W RS
  T a, b, c;        // covariance (after blur)
  T sg, l1, l2;     // sqrt(gap), eigenvalues
  T nx, ny, nn;     // un-normalised eigenvector and its norm
};

template <typename T>
GS_HD Projected<T> project_one(const T* p, const T* ls, const T* q, T logit, const CameraConst<T>& C,
                               ProjectState<T>* st = nullptr) {
  typedef Math<T> M;
  Projected<T> o;
  T n2 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
  T qn = M::sqrt_(n2);
  T inv = T(1) / qn;
  T x = q[0] * inv, y = q[1] * inv, zq = q[2] * inv, wq = q[3] * inv;
  T s0 = M::exp_(ls[0]), s1 = M::exp_(ls[1]), s2 = M::exp_(ls[2]);

  T cam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    cam[i] = ((C.Tcw[i * 4 + 0] * p[0] + C.Tcw[i * 4 + 1] * p[1]) + C.Tcw[i * 4 + 2] * p[2]) + C.Tcw[i * 4 + 3];
  T z = cam[2];
  T u = (C.fx * cam[0]) / z + C.cx;
  T v = (C.fy * cam[1]) / z + C.cy;
  T tu = clampv(u, C.lo_x, C.hi_x), tv = clampv(v, C.lo_y, C.hi_y);
  T J00 = C.fx / z, J02 = -(tu - C.cx) / z, J11 = C.fy / z, J12 = -(tv - C.cy) / z;

  T x2 = x * x, y2 = y * y, z2 = zq * zq;
  T RS[3][3];
  RS[0][0] = s0 * ((T(1) - T(2) * y2) - T(2) * z2);
  RS[0][1] = s1 * (T(2) * x * y - T(2) * wq * zq);
  RS[0][2] = s2 * (T(2) * x * zq + T(2) * wq * y);
  RS[1][0] = s0 * (T(2) * x * y + T(2) * wq * zq);
  RS[1][1] = s1 * ((T(1) - T(2) * x2) - T(2) * z2);
  RS[1][2] = s2 * (T(2) * y * zq - T(2) * wq * x);
  RS[2][0] = s0 * (T(2) * x * zq - T(2) * wq * y);
  RS[2][1] = s1 * (T(2) * y * zq + T(2) * wq * x);
  RS[2][2] = s2 * ((T(1) - T(2) * x2) - T(2) * y2);

  T JW[2][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    JW[0][j] = J00 * C.Tcw[0 * 4 + j] + J02 * C.Tcw[2 * 4 + j];
    JW[1][j] = J11 * C.Tcw[1 * 4 + j] + J12 * C.Tcw[2 * 4 + j];
  }
  T m[2][3];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) m[i][j] = (JW[i][0] * RS[0][j] + JW[i][1] * RS[1][j]) + JW[i][2] * RS[2][j];
  T a = (m[0][0] * m[0][0] + m[0][1] * m[0][1]) + m[0][2] * m[0][2];
  T b = (m[0][0] * m[1][0] + m[0][1] * m[1][1]) + m[0][2] * m[1][2];
  T c = (m[1][0] * m[1][0] + m[1][1] * m[1][1]) + m[1][2] * m[1][2];
  if (C.blur > T(0)) { a = a + C.blur; c = c + C.blur; }

  T tr = a + c, det = a * c - b * b;
  T gap = tr * tr - T(4) * det;
  T sg = M::sqrt_(gap > T(0) ? gap : T(0));
  T l1 = (tr + sg) * T(0.5), l2 = (tr - sg) * T(0.5);
  T vx = a - l2, vy = b;
  T vn = M::sqrt_(vx * vx + vy * vy);
  T v1x, v1y;
  if (vn > T(0)) { T iv = T(1) / vn; v1x = vx * iv; v1y = vy * iv; }
  else { v1x = T(1); v1y = T(0); }
  T sig0 = M::sqrt_(l1), sig1 = M::sqrt_(l2);

  T alpha = T(1) / (T(1) + M::exp_(-logit));
  T gscale = M::sqrt_(T(2) * M::log_(alpha / C.alpha_threshold));
  T sx = sig0 * gscale, sy = sig1 * gscale;
  T e1x = v1x * sx, e1y = v1y * sx, e2x = -v1y * sy, e2y = v1x * sy;
  T ex = M::sqrt_(e1x * e1x + e2x * e2x), ey = M::sqrt_(e1y * e1y + e2y * e2y);
  T lox = u - ex, loy = v - ey, upx = u + ex, upy = v + ey;
  o.in_view = (z > C.near_) && (z < C.far_) && (upx > T(0)) && (upy > T(0)) && (lox < C.w) && (loy < C.h);
  o.mean_x = u; o.mean_y = v; o.axis_x = v1x; o.axis_y = v1y;
  o.sigma_x = sig0; o.sigma_y = sig1; o.alpha = alpha; o.z = z;

  if (st) {
    st->qn = qn; st->q[0] = x; st->q[1] = y; st->q[2] = zq; st->q[3] = wq;
    st->s[0] = s0; st->s[1] = s1; st->s[2] = s2;
    st->cam[0] = cam[0]; st->cam[1] = cam[1]; st->cam[2] = cam[2];
    st->tu = tu; st->tv = tv; st->u = u; st->v = v;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) st->m[i][j] = m[i][j];
    st->a = a; st->b = b; st->c = c; st->sg = sg; st->l1 = l1; st->l2 = l2;
    st->nx = vx; st->ny = vy; st->nn = vn;
  }
  return o;
}

// ---------------------------------------------------------------- OBB tile query (f32)
struct TileQuery {
  float ib00, ib01, ib10, ib11;
  float rel_x, rel_y;
  int min_x, min_y, span_x, span_y;
};

GS_HD int f2i_clamped(float x) {
  const float lim = 1073741824.0f;
  if (x > lim) x = lim;
  if (x < -lim) x = -lim;
  return (int)x;
}

GS_HD bool finitef_(float x) { return (gs_f2u(x) & 0x7f800000u) != 0x7f800000u; }

GS_HD TileQuery obb_query(float mx, float my, float ax, float ay, float s0, float s1, float alpha, int img_w,
                          int img_h, int ts, float alpha_threshold) {
  TileQuery qy;
  float gscale = sqrtf(2.0f * gs_logf(alpha / alpha_threshold));
  float sx = s0 * gscale, sy = s1 * gscale;
  float a2x = -ay, a2y = ax;
  float e1x = ax * sx, e1y = ay * sx, e2x = a2x * sy, e2y = a2y * sy;
  float ex = sqrtf(e1x * e1x + e2x * e2x), ey = sqrtf(e1y * e1y + e2y * e2y);
  float min_bx = mx - ex, min_by = my - ey, max_bx = mx + ex, max_by = my + ey;
  bool finite = finitef_(min_bx) && finitef_(min_by) && finitef_(max_bx) && finitef_(max_by);
  if (!finite) {
    qy.span_x = qy.span_y = 0; qy.min_x = qy.min_y = 0;
    qy.ib00 = qy.ib01 = qy.ib10 = qy.ib11 = qy.rel_x = qy.rel_y = 0.f;
    return qy;
  }
  qy.ib00 = ax / sx; qy.ib01 = ay / sx; qy.ib10 = a2x / sy; qy.ib11 = a2y / sy;
  int max_tx = (img_w - 1) / ts, max_ty = (img_h - 1) / ts;
  float fts = (float)ts;
  int lo_x = f2i_clamped(floorf(min_bx / fts)); lo_x = lo_x > 0 ? lo_x : 0;
  int lo_y = f2i_clamped(floorf(min_by / fts)); lo_y = lo_y > 0 ? lo_y : 0;
  int hi_x = f2i_clamped(ceilf(max_bx / fts)); hi_x = hi_x > lo_x + 1 ? hi_x : lo_x + 1; hi_x = hi_x < max_tx + 1 ? hi_x : max_tx + 1;
  int hi_y = f2i_clamped(ceilf(max_by / fts)); hi_y = hi_y > lo_y + 1 ? hi_y : lo_y + 1; hi_y = hi_y < max_ty + 1 ? hi_y : max_ty + 1;
  qy.min_x = lo_x; qy.min_y = lo_y;
  qy.span_x = hi_x - lo_x; qy.span_y = hi_y - lo_y;
  qy.rel_x = (float)(lo_x * ts) - mx; qy.rel_y = (float)(lo_y * ts) - my;
  return qy;
}

GS_HD float min4f(float a, float b, float c, float d) {
  float m = a; m = (b < m) ? b : m; m = (c < m) ? c : m; m = (d < m) ? d : m; return m;
}
GS_HD float max4f(float a, float b, float c, float d) {
  float m = a; m = (b > m) ? b : m; m = (c > m) ? c : m; m = (d > m) ? d : m; return m;
}

GS_HD bool test_tile(const TileQuery& qy, int tu, int tv, int ts) {
  float lx = qy.rel_x + (float)(tu * ts), ly = qy.rel_y + (float)(tv * ts);
  float ux = lx + (float)ts, uy = ly + (float)ts;
  float a0 = qy.ib00 * lx + qy.ib01 * ly, a1 = qy.ib00 * ux + qy.ib01 * ly;
  float a2 = qy.ib00 * ux + qy.ib01 * uy, a3 = qy.ib00 * lx + qy.ib01 * uy;
  float b0 = qy.ib10 * lx + qy.ib11 * ly, b1 = qy.ib10 * ux + qy.ib11 * ly;
  float b2 = qy.ib10 * ux + qy.ib11 * uy, b3 = qy.ib10 * lx + qy.ib11 * uy;
  bool separates = false;
  if (min4f(a0, a1, a2, a3) > 1.0f || max4f(a0, a1, a2, a3) < -1.0f) separates = true;
  if (min4f(b0, b1, b2, b3) > 1.0f || max4f(b0, b1, b2, b3) < -1.0f) separates = true;
  return !separates;
}

GS_HD int span_count(const TileQuery& qy) {
  return (qy.span_x > 0 && qy.span_y > 0) ? qy.span_x * qy.span_y : 0;
}

}  // namespace gs
