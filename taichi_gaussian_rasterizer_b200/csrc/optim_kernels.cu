// optim_kernels.cu — per point optimizer steps for the visible set (the consumers of the
// render path's visibility output).
//
// Replaces (paths relative to /root/reference/taichi_splatting/):
//   optim/fractional_adam.py:7-85      scalar_kernel / vector_kernel (weighted Adam)
//   optim/fractional_laprop.py:7-86    scalar_kernel / vector_kernel (weighted LaProp)
//   optim/fractional.py:108-148        weighted_step: local basis, mask_lr, point_lr
//   optim/fractional.py:150-151,186    saturate(weight) and  param[indexes] -= lr_step * saturate(weight)
//   optim/visibility_aware.py:37-48    update_visibility (running visibility, per point weight)
//   optim/visibility_aware.py:95-97    gradient rescaling  grad * grad_scale / (visibility + vis_smooth)
// The reference runs one Taichi kernel plus ~10 torch ops per parameter group (a dense zeros_like + scatter of the
// gradient, index_put of the step, einsum with the local basis, ...).  Here a group's whole update is ONE kernel
// over the visible points: gradient rescale -> (inverse basis) -> moment update -> step -> (basis) -> lr masks ->
// parameter update in place.  All tensors f32 like the reference's kernels; indexes must be unique.
#include "common.cuh"

namespace gs {

__device__ __forceinline__ float opt_lerp(float t, float a, float b) { return a * t + b * (1.0f - t); }  // generic.py lerp
__device__ __forceinline__ float opt_saturate(float x) { return 1.0f - 1.0f / expf(2.0f * x); }

struct OptScalars {
  float lr, beta1, beta2, eps, grad_scale, vis_smooth;
  int bias_correction, use_vis;
};

// first / second moment update and the raw step for one component (scalar groups) or one point (vector groups);
// `gg` is g*g for scalar groups and |g|^2 for vector groups, v the matching running average.
template <int ALGO>
__device__ __forceinline__ float opt_v(const OptScalars& s, float w, float v_old, float gg) {
  return opt_lerp(powf(s.beta2, w), v_old, gg);
}

template <int ALGO>
__device__ __forceinline__ float opt_m_and_step(const OptScalars& s, float w, float tw, float m_old, float g, float v,
                                                float* step) {
  const float b1w = powf(s.beta1, w);
  if (ALGO == 0) {  // Adam (fractional_adam.py:29-41)
    const float bias = s.bias_correction ? sqrtf(1.0f - powf(s.beta2, tw)) / (1.0f - powf(s.beta1, tw)) : 1.0f;
    const float m = opt_lerp(b1w, m_old, g);
    *step = m / fmaxf(sqrtf(v), s.eps) * bias * s.lr;
    return m;
  } else {          // LaProp (fractional_laprop.py:29-43)
    const float bias1 = s.bias_correction ? 1.0f - powf(s.beta1, tw) : 1.0f;
    const float bias2 = s.bias_correction ? 1.0f - powf(s.beta2, tw) : 1.0f;
    const float m = opt_lerp(b1w, m_old, g / fmaxf(sqrtf(v / bias2), s.eps));
    *step = m * s.lr / bias1;
    return m;
  }
}

// scalar groups: one thread per (visible point, column)
template <int ALGO>
__global__ void __launch_bounds__(256)
opt_step_scalar_kernel(OptScalars s, int64_t M, int D, const int64_t* __restrict__ indexes,
                       const float* __restrict__ weight, const float* __restrict__ visibility,
                       const float* __restrict__ grad, float* __restrict__ m_arr, float* __restrict__ v_arr,
                       const float* __restrict__ total_weight, float* __restrict__ param,
                       const float* __restrict__ mask_lr, const float* __restrict__ point_lr) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * D) return;
  const int64_t i = e / D;
  const int j = (int)(e - i * D);
  const int64_t idx = indexes[i];
  const float w = weight[i], tw = total_weight[idx];
  float g = grad[idx * D + j];
  if (s.use_vis) g = g * s.grad_scale / (visibility[i] + s.vis_smooth);
  const float v = opt_v<ALGO>(s, w, v_arr[idx * D + j], g * g);
  float step;
  const float m = opt_m_and_step<ALGO>(s, w, tw, m_arr[idx * D + j], g, v, &step);
  m_arr[idx * D + j] = m;
  v_arr[idx * D + j] = v;
  if (mask_lr) step *= mask_lr[j];
  if (point_lr) step *= point_lr[idx];
  param[idx * D + j] -= step * opt_saturate(w);
}

// vector / local_vector groups: one thread per visible point, v is the running squared NORM (one value per point).
// LOCAL > 0: the gradient is first expressed in the point's local basis (LOCAL x LOCAL, row major, inverse applied
// here) and the step mapped back with the basis (fractional.py:126-139).
template <int ALGO, int LOCAL>
__global__ void __launch_bounds__(128)
opt_step_vector_kernel(OptScalars s, int64_t M, int D, const int64_t* __restrict__ indexes,
                       const float* __restrict__ weight, const float* __restrict__ visibility,
                       const float* __restrict__ grad, float* __restrict__ m_arr, float* __restrict__ v_arr,
                       const float* __restrict__ total_weight, float* __restrict__ param,
                       const float* __restrict__ mask_lr, const float* __restrict__ point_lr,
                       const float* __restrict__ basis) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t idx = indexes[i];
  const float w = weight[i], tw = total_weight[idx];
  const float gs_ = s.use_vis ? s.grad_scale / (visibility[i] + s.vis_smooth) : 1.0f;
  const float sat = opt_saturate(w);
  const float plr = point_lr ? point_lr[idx] : 1.0f;

  if (LOCAL > 0) {
    constexpr int L = LOCAL > 0 ? LOCAL : 1;
    float B[L][L], Binv[L][L], g[L], gl[L], st[L];
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int c = 0; c < L; ++c) B[r][c] = basis[(i * L + r) * L + c];
    if (L == 2) {
      const float det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
      Binv[0][0] = B[1][1] / det; Binv[0][1] = -B[0][1] / det;
      Binv[1][0] = -B[1][0] / det; Binv[1][1] = B[0][0] / det;
    } else {
      const float c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1], c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2],
                  c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
      const float det = B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02;
      Binv[0][0] = c00 / det; Binv[1][0] = c01 / det; Binv[2][0] = c02 / det;
      Binv[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) / det;
      Binv[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) / det;
      Binv[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) / det;
      Binv[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) / det;
      Binv[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) / det;
      Binv[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) / det;
    }
#pragma unroll
    for (int c = 0; c < L; ++c) g[c] = grad[idx * L + c] * gs_;
    float gg = 0.f;
#pragma unroll
    for (int r = 0; r < L; ++r) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < L; ++c) a += Binv[r][c] * g[c];
      gl[r] = a;
      gg += a * a;
    }
    const float v = opt_v<ALGO>(s, w, v_arr[idx], gg);
    v_arr[idx] = v;
#pragma unroll
    for (int r = 0; r < L; ++r) {
      const float m = opt_m_and_step<ALGO>(s, w, tw, m_arr[idx * L + r], gl[r], v, &st[r]);
      m_arr[idx * L + r] = m;
    }
#pragma unroll
    for (int r = 0; r < L; ++r) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < L; ++c) a += B[r][c] * st[c];
      if (mask_lr) a *= mask_lr[r];
      param[idx * L + r] -= a * plr * sat;
    }
  } else {
    float gg = 0.f;
    for (int j = 0; j < D; ++j) {
      const float g = grad[idx * D + j] * gs_;
      gg += g * g;
    }
    const float v = opt_v<ALGO>(s, w, v_arr[idx], gg);
    v_arr[idx] = v;
    for (int j = 0; j < D; ++j) {
      const float g = grad[idx * D + j] * gs_;
      float step;
      const float m = opt_m_and_step<ALGO>(s, w, tw, m_arr[idx * D + j], g, v, &step);
      m_arr[idx * D + j] = m;
      if (mask_lr) step *= mask_lr[j];
      param[idx * D + j] -= step * plr * sat;
    }
  }
}

// visibility_aware.py:37-48: updated = (lerp(beta, vis^4, running^4))^(1/4) with lerp(t, a, b) = a + (b - a) t;
// running[idx] = updated; weight = vis / max(updated, eps); total_weight[idx] += weight (:88-89).
__global__ void __launch_bounds__(256)
opt_update_visibility_kernel(int64_t M, const int64_t* __restrict__ indexes, const float* __restrict__ visibility,
                             float* __restrict__ running_vis, float* __restrict__ total_weight,
                             float* __restrict__ weight_out, float beta, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t idx = indexes[i];
  const float vis = visibility[i], run = running_vis[idx];
  const float a = vis * vis * vis * vis, b = run * run * run * run;
  const float updated = powf(a + (b - a) * beta, 0.25f);
  running_vis[idx] = updated;
  const float w = vis / fmaxf(updated, eps);
  weight_out[i] = w;
  total_weight[idx] += w;
}

__global__ void __launch_bounds__(256)
opt_accumulate_weight_kernel(int64_t M, const int64_t* __restrict__ indexes, const float* __restrict__ weight,
                             float* __restrict__ total_weight) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) total_weight[indexes[i]] += weight[i];
}

}  // namespace gs

using namespace gs;

extern "C" {

int gs_opt_update_visibility(int64_t num_visible, const int64_t* indexes, const float* visibility, float* running_vis,
                             float* total_weight, float* weight_out, double vis_beta, double eps, void* stream) {
  GS_CHECK_ARG(num_visible >= 0, "gs_opt_update_visibility: negative size");
  if (num_visible == 0) return GS_OK;
  GS_CHECK_ARG(indexes && visibility && running_vis && total_weight && weight_out, "gs_opt_update_visibility: null tensor");
  opt_update_visibility_kernel<<<(unsigned)ceil_div(num_visible, 256), 256, 0, (cudaStream_t)stream>>>(
      num_visible, indexes, visibility, running_vis, total_weight, weight_out, (float)vis_beta, (float)eps);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_opt_accumulate_weight(int64_t num_visible, const int64_t* indexes, const float* weight, float* total_weight,
                             void* stream) {
  GS_CHECK_ARG(num_visible >= 0, "gs_opt_accumulate_weight: negative size");
  if (num_visible == 0) return GS_OK;
  GS_CHECK_ARG(indexes && weight && total_weight, "gs_opt_accumulate_weight: null tensor");
  opt_accumulate_weight_kernel<<<(unsigned)ceil_div(num_visible, 256), 256, 0, (cudaStream_t)stream>>>(
      num_visible, indexes, weight, total_weight);
  GS_LAUNCH_CHECK();
  return GS_OK;
}

int gs_opt_step(const GsOptParams* p, const int64_t* indexes, const float* weight, const float* visibility,
                const float* grad, float* m, float* v, const float* total_weight, float* param, const float* mask_lr,
                const float* point_lr, const float* basis, void* stream) {
  GS_CHECK_ARG(p != nullptr, "gs_opt_step: null params");
  GS_CHECK_ARG(p->algorithm == GS_OPT_ADAM || p->algorithm == GS_OPT_LAPROP, "gs_opt_step: unknown algorithm %d", p->algorithm);
  GS_CHECK_ARG(p->dims > 0 && p->num_points >= 0 && p->num_visible >= 0 && p->num_visible <= p->num_points,
               "gs_opt_step: bad sizes");
  if (p->num_visible == 0) return GS_OK;
  GS_CHECK_ARG(indexes && weight && grad && m && v && total_weight && param, "gs_opt_step: null tensor");
  const bool use_vis = p->vis_smooth >= 0.0;
  GS_CHECK_ARG(!use_vis || visibility != nullptr, "gs_opt_step: visibility scaling requested without visibility");
  OptScalars s;
  s.lr = (float)p->lr; s.beta1 = (float)p->beta1; s.beta2 = (float)p->beta2; s.eps = (float)p->eps;
  s.grad_scale = (float)p->grad_scale; s.vis_smooth = (float)p->vis_smooth;
  s.bias_correction = p->bias_correction; s.use_vis = use_vis ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = p->num_visible;
  const int D = p->dims;
#define GS_OPT_ARGS s, M, D, indexes, weight, visibility, grad, m, v, total_weight, param, mask_lr, point_lr
  if (p->group_type == GS_OPT_SCALAR) {
    const unsigned blocks = (unsigned)ceil_div(M * D, 256);
    if (p->algorithm == GS_OPT_ADAM) opt_step_scalar_kernel<0><<<blocks, 256, 0, st>>>(GS_OPT_ARGS);
    else opt_step_scalar_kernel<1><<<blocks, 256, 0, st>>>(GS_OPT_ARGS);
  } else if (p->group_type == GS_OPT_VECTOR) {
    const unsigned blocks = (unsigned)ceil_div(M, 128);
    if (p->algorithm == GS_OPT_ADAM) opt_step_vector_kernel<0, 0><<<blocks, 128, 0, st>>>(GS_OPT_ARGS, nullptr);
    else opt_step_vector_kernel<1, 0><<<blocks, 128, 0, st>>>(GS_OPT_ARGS, nullptr);
  } else if (p->group_type == GS_OPT_LOCAL_VECTOR) {
    GS_CHECK_ARG(basis != nullptr, "gs_opt_step: basis is required for local_vector groups");
    const unsigned blocks = (unsigned)ceil_div(M, 128);
    if (D == 2) {
      if (p->algorithm == GS_OPT_ADAM) opt_step_vector_kernel<0, 2><<<blocks, 128, 0, st>>>(GS_OPT_ARGS, basis);
      else opt_step_vector_kernel<1, 2><<<blocks, 128, 0, st>>>(GS_OPT_ARGS, basis);
    } else if (D == 3) {
      if (p->algorithm == GS_OPT_ADAM) opt_step_vector_kernel<0, 3><<<blocks, 128, 0, st>>>(GS_OPT_ARGS, basis);
      else opt_step_vector_kernel<1, 3><<<blocks, 128, 0, st>>>(GS_OPT_ARGS, basis);
    } else {
      GS_UNSUPPORTED("gs_opt_step: local_vector groups need 2 or 3 columns, got %d", D);
    }
  } else {
    GS_UNSUPPORTED("gs_opt_step: unknown group type %d", p->group_type);
  }
#undef GS_OPT_ARGS
  GS_LAUNCH_CHECK();
  return GS_OK;
}

}  // extern "C"
