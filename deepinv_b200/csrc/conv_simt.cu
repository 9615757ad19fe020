// conv_simt.cu — fp32 (reference-precision) denoiser convolutions on CUDA cores, NCHW.
//
// Replaces the ATen calls behind deepinv/models/drunet.py:200-210,323-433 (3x3 conv, ResBlock,
// strided 2x2 down-conv, transposed 2x2 up-conv, all bias-free) and deepinv/models/dncnn.py:116-131
// (3x3 conv + bias + ReLU).  This is the parity path (matches the fp32 reference to ~1e-6); the
// throughput path is the tcgen05 implicit GEMM in conv_tc.cu.
//
//   kind 0: 3x3, stride 1, zero padding 1 — register-tiled direct convolution: a CTA computes an
//           8x32 pixel x 32 output-channel tile, staging 8 input channels (with halo) and their
//           weights in shared memory per step; each thread owns 4 pixels x 8 channels.
//   kind 1/2: 2x2 stride-2 conv / transposed conv as a small gather-GEMM-scatter (64x32 tiles).
// Fusions: input sum (x + xadd: the U-Net skip), bias, ReLU, residual add.
#include "common.cuh"

namespace dinvk {

constexpr int C3_CC = 8;     // input channels per step
constexpr int C3_TH = 8;     // tile rows
constexpr int C3_TW = 32;    // tile cols
constexpr int C3_TC = 32;    // output channels per CTA
constexpr int C3_ROWP = 36;  // padded smem row (34 used)

__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ xadd,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ res, float* __restrict__ out,
                                                          int Cin, int Cout, int H, int W, int tiles_x, int act) {
  __shared__ __align__(16) float s_in[C3_CC][C3_TH + 2][C3_ROWP];
  __shared__ __align__(16) float s_w[C3_CC][9][C3_TC];
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = (tid >> 3) & 7, tz = tid >> 6;
  const int ty0 = (blockIdx.x / tiles_x) * C3_TH, tx0 = (blockIdx.x % tiles_x) * C3_TW;
  const int co0 = blockIdx.y * C3_TC;
  const int b = blockIdx.z;
  const long long HW = (long long)H * W;
  const float* xb = x + (long long)b * Cin * HW;
  const float* ab = xadd ? xadd + (long long)b * Cin * HW : nullptr;

  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[p][q] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += C3_CC) {
    // stage inputs (halo, zero padding) — 8 x 10 x 34 values
    for (int idx = tid; idx < C3_CC * (C3_TH + 2) * (C3_TW + 2); idx += 256) {
      const int c = idx / ((C3_TH + 2) * (C3_TW + 2));
      const int rem = idx - c * ((C3_TH + 2) * (C3_TW + 2));
      const int r = rem / (C3_TW + 2), col = rem - r * (C3_TW + 2);
      const int gy = ty0 + r - 1, gx = tx0 + col - 1;
      float v = 0.f;
      if (c0 + c < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const long long o = (long long)(c0 + c) * HW + (long long)gy * W + gx;
        v = __ldg(xb + o);
        if (ab) v += __ldg(ab + o);
      }
      s_in[c][r][col] = v;
    }
    // stage weights: w[(co, ci, ky, kx)]
    for (int idx = tid; idx < C3_CC * 9 * C3_TC; idx += 256) {
      const int co = idx / (C3_CC * 9);
      const int rem = idx - co * (C3_CC * 9);
      const int c = rem / 9, k = rem - c * 9;
      float v = 0.f;
      if (c0 + c < Cin && co0 + co < Cout) v = __ldg(w + ((long long)(co0 + co) * Cin + (c0 + c)) * 9 + k);
      s_w[c][k][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < C3_CC; ++c) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float4 i0 = *reinterpret_cast<const float4*>(&s_in[c][ty + ky][tx * 4]);
        const float2 i1 = *reinterpret_cast<const float2*>(&s_in[c][ty + ky][tx * 4 + 4]);
        const float in6[6] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c][ky * 3 + kx][tz * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c][ky * 3 + kx][tz * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[p][q] = fmaf(in6[p + kx], wv[q], acc[p][q]);
        }
      }
    }
    __syncthreads();
  }

  // epilogue: bias -> activation -> residual
  const int gy = ty0 + ty;
  if (gy < H) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int co = co0 + tz * 8 + q;
      if (co < Cout) {
        const float bv = bias ? __ldg(bias + co) : 0.f;
        const long long ob = ((long long)b * Cout + co) * HW + (long long)gy * W;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int gx = tx0 + tx * 4 + p;
          if (gx < W) {
            float v = acc[p][q] + bv;
            if (act == 1) v = fmaxf(v, 0.f);
            if (res) v += __ldg(res + ob + gx);
            out[ob + gx] = v;
          }
        }
      }
    }
  }
}

// gather-GEMM-scatter for the 2x2 stride-2 conv (mode 1) and its transpose (mode 2)
//   mode 1: M = B*Ho*Wo, N = Cout, K = 4*Cin;  A(m,k) = x[b, c, 2yo+dy, 2xo+dx], Wt(k,n) = w[n*K + k]
//   mode 2: M = B*H*W,   N = 4*Cout, K = Cin;  A(m,k) = (x+xadd)[b, k, y, x],   Wt(k,n) = w[k*N + n],
//           out[b, n/4, 2y + (n%4)/2, 2x + n%2]
constexpr int G_TM = 64, G_TN = 32, G_TK = 32;

__global__ void __launch_bounds__(256) conv2x2_f32_kernel(const float* __restrict__ x, const float* __restrict__ xadd,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ res, float* __restrict__ out,
                                                          int B, int Cin, int Cout, int H, int W, int mode, int act) {
  __shared__ float sA[G_TK][G_TM + 1];
  __shared__ float sW[G_TK][G_TN + 1];
  const int tid = threadIdx.x;
  const int Ho = H / 2, Wo = W / 2;
  const long long M = mode == 1 ? (long long)B * Ho * Wo : (long long)B * H * W;
  const int N = mode == 1 ? Cout : 4 * Cout;
  const int K = mode == 1 ? 4 * Cin : Cin;
  const long long m0 = (long long)blockIdx.x * G_TM;
  const int n0 = blockIdx.y * G_TN;
  const int tm = tid & 31, tn = tid >> 5;  // thread owns pixels {tm, tm+32} x channels {tn*4 .. tn*4+3}
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const long long HW = (long long)H * W;

  for (int k0 = 0; k0 < K; k0 += G_TK) {
    for (int idx = tid; idx < G_TK * G_TM; idx += 256) {
      const int kk = idx / G_TM, mm = idx - kk * G_TM;
      const long long m = m0 + mm;
      const int k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) {
        if (mode == 1) {
          const int xo = (int)(m % Wo), yo = (int)((m / Wo) % Ho);
          const long long bb = m / ((long long)Wo * Ho);
          const int c = k >> 2, dy = (k >> 1) & 1, dx = k & 1;
          const long long o = (bb * Cin + c) * HW + (long long)(2 * yo + dy) * W + (2 * xo + dx);
          v = __ldg(x + o);
          if (xadd) v += __ldg(xadd + o);
        } else {
          const long long bb = m / HW, p = m - bb * HW;
          const long long o = (bb * Cin + k) * HW + p;
          v = __ldg(x + o);
          if (xadd) v += __ldg(xadd + o);
        }
      }
      sA[kk][mm] = v;
    }
    for (int idx = tid; idx < G_TK * G_TN; idx += 256) {
      int kk, nn;
      if (mode == 1) { nn = idx / G_TK; kk = idx - nn * G_TK; }  // k fastest in memory
      else { kk = idx / G_TN; nn = idx - kk * G_TN; }            // n fastest in memory
      const int k = k0 + kk, n = n0 + nn;
      float v = 0.f;
      if (k < K && n < N) v = mode == 1 ? __ldg(w + (long long)n * K + k) : __ldg(w + (long long)k * N + n);
      sW[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < G_TK; ++kk) {
      const float a0 = sA[kk][tm], a1 = sA[kk][tm + 32];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float wv = sW[kk][tn * 4 + q];
        acc[0][q] = fmaf(a0, wv, acc[0][q]);
        acc[1][q] = fmaf(a1, wv, acc[1][q]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long m = m0 + tm + 32 * i;
    if (m >= M) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int n = n0 + tn * 4 + q;
      if (n >= N) continue;
      long long o;
      int co;
      if (mode == 1) {
        const int xo = (int)(m % Wo), yo = (int)((m / Wo) % Ho);
        const long long bb = m / ((long long)Wo * Ho);
        co = n;
        o = (bb * Cout + co) * ((long long)Ho * Wo) + (long long)yo * Wo + xo;
      } else {
        const long long bb = m / HW, p = m - bb * HW;
        const int y = (int)(p / W), xx = (int)(p - (long long)y * W);
        co = n >> 2;
        const int dy = (n >> 1) & 1, dx = n & 1;
        o = (bb * Cout + co) * (4 * HW) + (long long)(2 * y + dy) * (2 * W) + (2 * xx + dx);
      }
      float v = acc[i][q] + (bias ? __ldg(bias + co) : 0.f);
      if (act == 1) v = fmaxf(v, 0.f);
      if (res) v += __ldg(res + o);
      out[o] = v;
    }
  }
}

}  // namespace dinvk

using namespace dinvk;

extern "C" int dinvk_conv_f32(const float* x, const float* xadd, const float* weight, const float* bias, const float* res,
                              float* out, int B, int Cin, int Cout, int H, int W, int kind, int act, void* stream) {
  DINVK_CHECK_ARG(x && weight && out, "dinvk_conv_f32: null pointer");
  DINVK_CHECK_ARG(B >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1, "dinvk_conv_f32: bad shape");
  DINVK_CHECK_ARG(kind >= 0 && kind <= 2, "dinvk_conv_f32: kind %d", kind);
  DINVK_CHECK_ARG(act == 0 || act == 1, "dinvk_conv_f32: act %d", act);
  if (B == 0) return DINVK_OK;
  if (kind == 0) {
    const int tiles_x = ceil_div(W, C3_TW), tiles_y = ceil_div(H, C3_TH);
    DINVK_CHECK_ARG(B <= 65535 && ceil_div(Cout, C3_TC) <= 65535, "dinvk_conv_f32: grid too large");
    DINVK_LAUNCH(conv3x3_f32_kernel, dim3(tiles_x * tiles_y, ceil_div(Cout, C3_TC), B), dim3(256), 0, stream, x, xadd,
                 weight, bias, res, out, Cin, Cout, H, W, tiles_x, act);
  } else {
    if (kind == 1) DINVK_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "dinvk_conv_f32: strided conv needs even H, W");
    const long long M = kind == 1 ? (long long)B * (H / 2) * (W / 2) : (long long)B * H * W;
    const int N = kind == 1 ? Cout : 4 * Cout;
    DINVK_LAUNCH(conv2x2_f32_kernel, dim3((unsigned)ceil_div(M, G_TM), ceil_div(N, G_TN)), dim3(256), 0, stream, x, xadd,
                 weight, bias, res, out, B, Cin, Cout, H, W, kind, act);
  }
  return DINVK_POST_LAUNCH();
}
