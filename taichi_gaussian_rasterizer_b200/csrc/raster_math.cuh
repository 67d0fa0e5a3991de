// raster_math.cuh — per pixel gaussian evaluation and its derivatives (device).
// Restates taichi_lib/generic.py:310-404 (gaussian_pdf, gaussian_pdf_with_grad, S_sig, S_sig_grad,
// gaussian_pdf_antialias, gaussian_pdf_antialias_with_grad) of /root/reference/taichi_splatting/.
#pragma once

#include "common.cuh"

namespace gs {

template <typename T> __device__ __forceinline__ T exp_(T x);
template <> __device__ __forceinline__ float exp_(float x) { return expf(x); }
template <> __device__ __forceinline__ double exp_(double x) { return exp(x); }

template <typename T>
struct Gauss2D {  // the packed (N,7) record: mean, axis, sigma, alpha
  T mx, my, ax, ay, sx, sy, alpha;
};

template <typename T>
__device__ __forceinline__ Gauss2D<T> load_gauss(const T* g) {
  Gauss2D<T> r;
  r.mx = g[0]; r.my = g[1]; r.ax = g[2]; r.ay = g[3]; r.sx = g[4]; r.sy = g[5]; r.alpha = g[6];
  return r;
}

template <typename T>
__device__ __forceinline__ T pdf(T px, T py, const Gauss2D<T>& g) {
  T dx = px - g.mx, dy = py - g.my;
  T tx = (dx * g.ax + dy * g.ay) / g.sx;
  T ty = (dx * -g.ay + dy * g.ax) / g.sy;
  return exp_<T>(T(-0.5) * (tx * tx + ty * ty));
}

template <typename T>
__device__ __forceinline__ T s_sig(T x, T sigma) {
  T z = x / sigma;
  return T(1) / (T(1) + exp_<T>(T(-1.6) * z - T(0.07) * z * z * z));
}

template <typename T>
__device__ __forceinline__ T pdf_aa(T px, T py, const Gauss2D<T>& g) {
  T dx = px - g.mx, dy = py - g.my;
  T tx = dx * g.ax + dy * g.ay;
  T ty = dx * -g.ay + dy * g.ax;
  T Sx1 = s_sig(tx + T(0.5), g.sx), Sx2 = s_sig(tx - T(0.5), g.sx);
  T Sy1 = s_sig(ty + T(0.5), g.sy), Sy2 = s_sig(ty - T(0.5), g.sy);
  return T(6.283185307179586) * g.sx * (Sx1 - Sx2) * g.sy * (Sy1 - Sy2);
}

template <typename T>
struct PdfGrad {
  T p, dmx, dmy, dax, day, dsx, dsy;
};

template <typename T>
__device__ __forceinline__ PdfGrad<T> pdf_grad(T px, T py, const Gauss2D<T>& g) {
  PdfGrad<T> r;
  T dx = px - g.mx, dy = py - g.my;
  T tx = (dx * g.ax + dy * g.ay) / g.sx;
  T ty = (dx * -g.ay + dy * g.ax) / g.sy;
  T tx2 = tx * tx, ty2 = ty * ty;
  T p = exp_<T>(T(-0.5) * (tx2 + ty2));
  r.p = p;
  r.dsx = tx2 * p / g.sx; r.dsy = ty2 * p / g.sy;
  T txs = tx / g.sx, tys = ty / g.sy;
  r.dax = p * (txs * -dx + tys * -dy);
  r.day = p * (txs * -dy + tys * dx);
  r.dmx = p * (txs * g.ax + tys * -g.ay);
  r.dmy = p * (txs * g.ay + tys * g.ax);
  return r;
}

template <typename T>
__device__ __forceinline__ void s_sig_grad(T x, T sigma, T& s, T& ds_dx, T& ds_dsig) {
  T z = x / sigma;
  s = T(1) / (T(1) + exp_<T>(T(-1.6) * z - T(0.07) * z * z * z));
  T d = (T(1.6) + T(0.21) * z * z) * s * (T(1) - s);
  ds_dx = d / sigma;
  ds_dsig = ds_dx * -z;
}

template <typename T>
__device__ __forceinline__ PdfGrad<T> pdf_aa_grad(T px, T py, const Gauss2D<T>& g) {
  PdfGrad<T> r;
  T dx = px - g.mx, dy = py - g.my;
  T tx = dx * g.ax + dy * g.ay;
  T ty = dx * -g.ay + dy * g.ax;
  T Sx1, dSx1, dSx1s, Sx2, dSx2, dSx2s, Sy1, dSy1, dSy1s, Sy2, dSy2, dSy2s;
  s_sig_grad(tx + T(0.5), g.sx, Sx1, dSx1, dSx1s);
  s_sig_grad(tx - T(0.5), g.sx, Sx2, dSx2, dSx2s);
  s_sig_grad(ty + T(0.5), g.sy, Sy1, dSy1, dSy1s);
  s_sig_grad(ty - T(0.5), g.sy, Sy2, dSy2, dSy2s);
  T ix = g.sx * (Sx1 - Sx2), iy = g.sy * (Sy1 - Sy2);
  const T tau = T(6.283185307179586);
  r.p = tau * ix * iy;
  T dSx = iy * g.sx * (dSx1 - dSx2);
  T dSy = ix * g.sy * (dSy1 - dSy2);
  r.dmx = tau * (dSx * -g.ax + dSy * g.ay);
  r.dmy = tau * (dSx * -g.ay + dSy * -g.ax);
  r.dsx = tau * iy * (Sx1 - Sx2 + (dSx1s - dSx2s) * g.sx);
  r.dsy = tau * ix * (Sy1 - Sy2 + (dSy1s - dSy2s) * g.sy);
  r.dax = tau * (dSx * dx + dSy * dy);
  r.day = tau * (dSx * dy + dSy * -dx);
  return r;
}

}  // namespace gs
