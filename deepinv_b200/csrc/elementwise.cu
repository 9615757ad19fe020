// elementwise.cu — axpy-type step algebra, batched reductions and CG scalar updates.
//
// Replaces the elementwise bodies of
//   OptimIterator.relaxation_step / fStep / gStep algebra   deepinv/optim/optim_iterators/{optim_iterator,pgd,admm,hqs}.py
//   conjugate_gradient vector updates + batched dot          deepinv/optim/linear/conjugate_gradient.py:47-66, linear/utils.py:6-26
//   DDRM spectral-domain update                              deepinv/sampling/diffusion.py:163-222
// All of these are HBM-bound streaming kernels: 128-bit loads when pointers are 16-byte aligned,
// grid sized to a multiple of the SM count (grid-stride loops).
#include "common.cuh"

#include <algorithm>

namespace dinvk {

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(256) axpbypcz_kernel(float* __restrict__ out, const float* __restrict__ x, float a,
                                                       const float* __restrict__ y, float b,
                                                       const float* __restrict__ z, float c, long long n, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* z4 = reinterpret_cast<const float4*>(z);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (; i < n4; i += stride) {
      float4 v = x4[i];
      v.x *= a; v.y *= a; v.z *= a; v.w *= a;
      if (y) { const float4 u = y4[i]; v.x += b * u.x; v.y += b * u.y; v.z += b * u.z; v.w += b * u.w; }
      if (z) { const float4 u = z4[i]; v.x += c * u.x; v.y += c * u.y; v.z += c * u.z; v.w += c * u.w; }
      o4[i] = v;
    }
    // tail
    for (long long j = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
      float v = a * x[j];
      if (y) v += b * y[j];
      if (z) v += c * z[j];
      out[j] = v;
    }
  } else {
    for (; i < n; i += stride) {
      float v = a * x[i];
      if (y) v += b * y[i];
      if (z) v += c * z[i];
      out[i] = v;
    }
  }
}

// out[b,i] = x[b,i] + sa * s[b] * y[b,i]
__global__ void __launch_bounds__(256) batched_axpy_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                           const float* __restrict__ y, const float* __restrict__ s,
                                                           float sa, long long n_per) {
  const int b = blockIdx.y;
  const float sc = sa * __ldg(s + b);
  const long long base = (long long)b * n_per;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per; i += stride)
    out[base + i] = x[base + i] + sc * y[base + i];
}

__device__ __forceinline__ double block_reduce_sum(double v) {
  __shared__ double warp_part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double r = 0.0;
  if (warp == 0) {
    r = lane < nw ? warp_part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// stage 1: partial[b, chunk] = sum over the chunk of x*y
__global__ void __launch_bounds__(256) batched_dot_partial_kernel(double* __restrict__ partial, const float* __restrict__ x,
                                                                  const float* __restrict__ y, long long n_per, int chunks) {
  const int b = blockIdx.y, ch = blockIdx.x;
  const long long base = (long long)b * n_per;
  float acc = 0.f;
  const long long stride = (long long)chunks * blockDim.x;
  for (long long i = (long long)ch * blockDim.x + threadIdx.x; i < n_per; i += stride) acc += x[base + i] * y[base + i];
  const double r = block_reduce_sum((double)acc);
  if (threadIdx.x == 0) partial[(long long)b * chunks + ch] = r;
}
// stage 2: out[b] = sum_chunk partial[b, chunk]
__global__ void __launch_bounds__(256) batched_dot_final_kernel(float* __restrict__ out, const double* __restrict__ partial, int chunks) {
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) acc += partial[(long long)b * chunks + i];
  const double r = block_reduce_sum(acc);
  if (threadIdx.x == 0) out[b] = (float)r;
}

__global__ void cg_scalars_kernel(int mode, float* __restrict__ out0, const float* __restrict__ num,
                                  const float* __restrict__ den, float eps, const float* __restrict__ bnorm2,
                                  float tol2, int* __restrict__ done, int B) {
  // single block.  `done` is a sticky device flag (caller zero-initialises it once per solve):
  //   mode 0 (alpha): a converged solve gets alpha = 0, so iterations issued after convergence do not
  //                   move x or r — the host may poll the flag lazily without changing the result;
  //   mode 1 (beta + stopping test): done <- 1 when every sample satisfies r.r < tol^2 * b.b.
  __shared__ int s_not_done;
  if (threadIdx.x == 0) s_not_done = 0;
  __syncthreads();
  const int frozen = done ? *done : 0;
  int not_done = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float v = num[b] / (den[b] + eps);
    out0[b] = (mode == 0 && frozen) ? 0.f : v;
    if (mode == 1) {
      float bn = bnorm2[b];
      bn = bn > 0.f ? bn : 1.f;
      if (!(num[b] < bn * tol2)) not_done = 1;
    }
  }
  if (mode == 1 && not_done) atomicExch(&s_not_done, 1);
  __syncthreads();
  if (mode == 1 && threadIdx.x == 0 && done && s_not_done == 0) *done = 1;
}

// DDRM update, see header.  init != 0: first draw (diffusion.py:177-190); y_bar is normalised in place.
__global__ void __launch_bounds__(256) ddrm_update_kernel(float* __restrict__ x_bar_out, const float* __restrict__ x_bar,
                                                          const float* __restrict__ x_bar_prev, float* __restrict__ y_bar,
                                                          const float* __restrict__ mask, const float* __restrict__ noise,
                                                          long long n, long long mask_n, float sigma_t, float sigma_prev,
                                                          float sigma_noise, float eta, float etab, float c_sig, float eps, int init) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv_sqrt2 = 0.70710678118654752440f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float m = fabsf(__ldg(mask + (i % mask_n)));
    const bool cs = m > sigma_noise;
    const float nsr = cs ? sigma_noise / (m + eps) : 0.f;
    float mean, sd;
    if (init) {
      float yb = y_bar[i];
      if (cs) { yb = yb / (m + eps); y_bar[i] = yb; }
      mean = cs ? yb : 0.f;
      sd = cs ? sqrtf(sigma_t * sigma_t - nsr * nsr) : sigma_t;
    } else {
      const float xb = x_bar[i];
      if (cs && sigma_t < nsr) {
        mean = xb + c_sig * (y_bar[i] - xb) / (nsr + eps);
        sd = eta * sigma_t;
      } else if (cs) {
        mean = (1.0f - etab) * xb + etab * y_bar[i];
        sd = sqrtf(fmaxf(sigma_t * sigma_t - (nsr * etab) * (nsr * etab), 0.f));
      } else {
        mean = xb + c_sig * (x_bar_prev[i] - xb) / sigma_prev;
        sd = eta * sigma_t;
      }
    }
    x_bar_out[i] = mean + sd * noise[i] * inv_sqrt2;
  }
}

static inline int stream_grid(long long n_threads_wanted) {
  const int per = 256;
  long long blocks = (n_threads_wanted + per - 1) / per;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace dinvk

using namespace dinvk;

extern "C" int dinvk_axpbypcz(float* out, const float* x, float a, const float* y, float b, const float* z, float c,
                              int64_t n, void* stream) {
  DINVK_CHECK_ARG(out && x && n >= 0, "dinvk_axpbypcz: bad arguments");
  if (n == 0) return DINVK_OK;
  const int vec = aligned16(out) && aligned16(x) && (!y || aligned16(y)) && (!z || aligned16(z)) && n >= 4;
  const long long work = vec ? (n >> 2) : n;
  DINVK_LAUNCH(axpbypcz_kernel, dim3(stream_grid(work)), dim3(256), 0, stream, out, x, a, y, b, z, c, (long long)n, vec);
  return DINVK_POST_LAUNCH();
}

// interleaved complex (B, n) [re, im, re, im, ...] -> planar (B, 2, n): the k-space layout of raw MRI files -> the operators'
// layout (deepinv/utils/mixins.py:148-156 from_torch_complex: view_as_real + moveaxis + contiguous), one pass
__global__ void __launch_bounds__(256) interleaved_to_planar_kernel(const float2* __restrict__ in, float* __restrict__ out, long long n) {
  const long long b = blockIdx.y;
  const float2* src = in + b * n;
  float* re = out + b * 2 * n;
  float* im = re + n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float2 v = src[i];
    re[i] = v.x;
    im[i] = v.y;
  }
}

extern "C" int dinvk_interleaved_to_planar(const float* in, float* out, int B, int64_t n_per, void* stream) {
  DINVK_CHECK_ARG(in && out && B >= 0 && n_per >= 0, "dinvk_interleaved_to_planar: bad arguments");
  DINVK_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 7) == 0, "dinvk_interleaved_to_planar: input must be 8-byte aligned");
  DINVK_CHECK_ARG(B <= 65535, "dinvk_interleaved_to_planar: batch too large");
  if (B == 0 || n_per == 0) return DINVK_OK;
  int gx = stream_grid(n_per);
  gx = std::max(1, std::min(gx, std::max(1, sm_count() * 8 / B)));
  DINVK_LAUNCH(interleaved_to_planar_kernel, dim3(gx, B), dim3(256), 0, stream, reinterpret_cast<const float2*>(in), out, (long long)n_per);
  return DINVK_POST_LAUNCH();
}

extern "C" int dinvk_batched_axpy(float* out, const float* x, const float* y, const float* s, float sa, int B,
                                  int64_t n_per, void* stream) {
  DINVK_CHECK_ARG(out && x && y && s && B >= 0 && n_per >= 0, "dinvk_batched_axpy: bad arguments");
  if (B == 0 || n_per == 0) return DINVK_OK;
  int gx = stream_grid(n_per);
  gx = std::max(1, std::min(gx, std::max(1, sm_count() * 8 / B)));
  DINVK_LAUNCH(batched_axpy_kernel, dim3(gx, B), dim3(256), 0, stream, out, x, y, s, sa, (long long)n_per);
  return DINVK_POST_LAUNCH();
}

static int dot_chunks(int B, int64_t n_per) {
  long long want = (n_per + 256 * 16 - 1) / (256 * 16);
  long long cap = std::max(1, sm_count() * 8 / std::max(1, B));
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}
extern "C" size_t dinvk_batched_dot_workspace_bytes(int B, int64_t n_per) {
  return sizeof(double) * (size_t)std::max(1, B) * (size_t)dot_chunks(B, n_per) + 64;
}
extern "C" int dinvk_batched_dot(float* out, const float* x, const float* y, int B, int64_t n_per, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  DINVK_CHECK_ARG(out && x && y && B >= 0 && n_per >= 0 && workspace, "dinvk_batched_dot: bad arguments");
  if (B == 0) return DINVK_OK;
  if (workspace_bytes < dinvk_batched_dot_workspace_bytes(B, n_per))
    return set_error(DINVK_EWORKSPACE, "dinvk_batched_dot: workspace too small");
  const int chunks = dot_chunks(B, n_per);
  double* partial = reinterpret_cast<double*>(((uintptr_t)workspace + 7) & ~(uintptr_t)7);
  DINVK_LAUNCH(batched_dot_partial_kernel, dim3(chunks, B), dim3(256), 0, stream, partial, x, y, (long long)n_per, chunks);
  DINVK_LAUNCH(batched_dot_final_kernel, dim3(B), dim3(256), 0, stream, out, (const double*)partial, chunks);
  return DINVK_POST_LAUNCH();
}

extern "C" int dinvk_cg_scalars(int mode, float* out0, const float* num, const float* den, float eps, const float* bnorm2,
                                float tol2, int32_t* all_done_flag, int B, void* stream) {
  DINVK_CHECK_ARG(out0 && num && den && B >= 0 && (mode == 0 || (mode == 1 && bnorm2)), "dinvk_cg_scalars: bad arguments");
  if (B == 0) return DINVK_OK;
  DINVK_LAUNCH(cg_scalars_kernel, dim3(1), dim3(256), 0, stream, mode, out0, num, den, eps, bnorm2, tol2, (int*)all_done_flag, B);
  return DINVK_POST_LAUNCH();
}

extern "C" int dinvk_ddrm_update(float* x_bar_out, const float* x_bar, const float* x_bar_prev, float* y_bar,
                                 const float* mask, const float* noise, int64_t n, int64_t mask_n, float sigma_t,
                                 float sigma_prev, float sigma_noise, float eta, float etab, float c_sig, float eps,
                                 int init, void* stream) {
  DINVK_CHECK_ARG(x_bar_out && y_bar && mask && noise && n >= 0 && mask_n > 0, "dinvk_ddrm_update: bad arguments");
  DINVK_CHECK_ARG(init || (x_bar && x_bar_prev), "dinvk_ddrm_update: x_bar/x_bar_prev required after the first draw");
  if (n == 0) return DINVK_OK;
  DINVK_LAUNCH(ddrm_update_kernel, dim3(stream_grid(n)), dim3(256), 0, stream, x_bar_out, x_bar, x_bar_prev, y_bar, mask,
               noise, (long long)n, (long long)mask_n, sigma_t, sigma_prev, sigma_noise, eta, etab, c_sig, eps, init);
  return DINVK_POST_LAUNCH();
}
