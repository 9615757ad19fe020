// spectral_pipe320.cuh — the two-pass pipelined 2-D transform of spectral_pipe.cuh for 320 x 320 images (cfg4: single- and
// multi-coil MRI), included by spectral.cu.
//
// 320 = 16 x 20 on both axes:
//   W transform of a row:  stage 1 = 20 radix-16 butterflies on x[j + 20 r], stage 2 = 16 radix-20 butterflies
//                          (Dft<20> = 5 x 4 Cooley-Tukey in registers) -> thread j < 16 owns bins {j + 16 r'}
//   H transform:           pass 1 owns the 16 rows {b + 20 a} of an image (tile b = 0..19), does the radix-16 a-butterfly per
//                          column and the twiddle w320^(b k_lo); pass 2 reads the 20 contiguous intermediate rows
//                          ws[img][k_lo][0..19][*] (one 51 KB bulk copy) and does the radix-20 b-butterfly per column
//                          -> rows h = k_lo + 16 k_hi.
// CTAs have 320 threads: thread = (line, j) with 20 threads per line in the row stages, thread = column in the column
// stages.  A line's 20 threads straddle warps, so the stage exchanges use CTA barriers; the staged line doubles as the
// line's work row (all loads of a stage complete before its stores), which keeps a CTA at 2 x 44 KB and two CTAs per SM.
// Multi-coil sources / destinations (MultiCoilMRI.A / A_adjoint, mri.py:254-324): the coil map multiply is fused into the
// pass-1 load, the per-coil k-space goes straight to the (batch, 2, ncoil, H, W) output, the adjoint's pass 2 writes the
// interleaved per-coil images that coil_combine_kernel reduces.
#pragma once
#include "spectral_pipe.cuh"

namespace dinvk {
namespace sp320 {

using sp::bulk_g2s;
using sp::fence_async_smem;
using sp::fence_mbar_init;
using sp::gmap;
using sp::griddep_launch;
using sp::griddep_wait;
using sp::mb_arrive;
using sp::mb_expect_tx;
using sp::mb_init;
using sp::mb_wait;

constexpr int N = 320;
constexpr int HA = 16;                  // rows per pass-1 tile (radix of the a-butterfly)
constexpr int HB = 20;                  // tiles per image in pass 1 / rows per pass-2 tile (radix of the b-butterfly)
constexpr int NT = 320;                 // threads per CTA
constexpr int IMO = 336;                // offset of the imaginary row inside a staged line (floats)
constexpr int LSTR = 688;               // staged line stride (floats) >= 2 * 341 (the line reused as 341 float2 work row)
constexpr int STAGE_F = HA * LSTR;
constexpr int P1_WARPS = NT / 32;
constexpr size_t P1_SMEM = (size_t)2 * STAGE_F * 4 + 2 * 320 * 8 + 64;
constexpr int P2_STAGES = 2;
constexpr int P2_TILE_F = HB * N * 2;   // floats per pass-2 tile
constexpr size_t P2_SMEM = (size_t)P2_STAGES * P2_TILE_F * 4 + 64;

struct Params {
  int B;                                // complex images handled by this launch
  int img0;                             // index of the first one (chunked launches keep the intermediate L2-resident)
  const float* p0; const float* p1; float a0, a1;
  int src_nc, src_div;                  // planar source: coil planes per sample, image -> sample divisor
  const float2* coil; long long coil_sb; int ncoil;   // multiply the source by the coil map (A)
  int gmode; const float* g; long long gsb, gsc, gsh; float gc; const float* gcb;
  const float* q0; const float* q1; float e0, e1, e2;
  float* out; int dst_nc; float2* tout; // planar destination with dst_nc coil planes, or interleaved (coil reduction follows)
  float2* ws;
  const float2* tw;                     // exp(-2 pi i k / 320)
  int inverse, centered, g_at_load;
};

__device__ __forceinline__ long long planar_off(int img, int nc, long long HW, long long& cs) {
  cs = (long long)nc * HW;
  return (long long)(img / nc) * 2 * cs + (long long)(img % nc) * HW;
}

// issue the 32 row copies (16 rows x 2 planes, 1280 B each) of pass-1 tile (img, b); called by warp 0
__device__ __forceinline__ void issue_rows(float* dst, const Params& P, int img, int b, uint64_t* bar, int lane) {
  constexpr long long HW = (long long)N * N;
  if (lane == 0) mb_expect_tx(bar, 32 * N * 4);
  __syncwarp();
  long long cs;
  const long long base = planar_off(img / P.src_div, P.src_nc, HW, cs);
  const int plane = lane >> 4, row = lane & 15;
  bulk_g2s(dst + row * LSTR + plane * IMO, P.p0 + base + plane * cs + (long long)(b + HB * row) * N, N * 4, bar);
}

template <bool HAS_P1>
__global__ void __launch_bounds__(NT, 2) sp320_pass1(const Params P) {
  DINVK_SP_DYN_SMEM();
  float* in0 = reinterpret_cast<float*>(sp_raw);
  float2* tws = reinterpret_cast<float2*>(in0 + 2 * STAGE_F);  // [r < 20][j < 16]: sign(r) w320^(r j)
  float2* twf = tws + 320;                                     // w320^k
  uint64_t* full = reinterpret_cast<uint64_t*>(twf + 320);
  uint64_t* empty = full + 2;
  const int tid = threadIdx.x, lane = tid & 31;
  const int line = tid / 20, j = tid - line * 20;
  const int ntiles = P.B * HB;
  constexpr long long HW = (long long)N * N;

  const int t0 = blockIdx.x;
  if (tid < 32) {
    if (tid == 0) {
      mb_init(&full[0], 1); mb_init(&full[1], 1);
      mb_init(&empty[0], P1_WARPS); mb_init(&empty[1], P1_WARPS);
      fence_mbar_init();
    }
    griddep_wait();
    if (tid == 0) griddep_launch();
    __syncwarp();
    if (t0 < ntiles) issue_rows(in0, P, P.img0 + t0 / HB, t0 % HB, &full[0], lane);
  }
  {
    const int r = tid >> 4, jj = tid & 15;  // 320 entries: r < 20, jj < 16
    float2 t = __ldg(P.tw + ((r * jj) % N));
    if (P.centered && (r & 1)) { t.x = -t.x; t.y = -t.y; }
    tws[tid] = t;
    twf[tid] = __ldg(P.tw + tid);
  }
  __syncthreads();

  const float sgn_im = P.inverse ? -1.0f : 1.0f;
  const float sw = (P.centered && (tid & 1)) ? -1.0f : 1.0f;  // (-1)^k_w, column = tid

  int it = 0;
  for (int t = t0; t < ntiles; t += gridDim.x, ++it) {
    const int s = it & 1;
    const int iml = t / HB, b = t - iml * HB;
    const int img = P.img0 + iml;
    const int tn = t + gridDim.x;
    if (tid < 32 && tn < ntiles) {
      if (it >= 1) mb_wait(&empty[s ^ 1], ((it - 1) >> 1) & 1);
      fence_async_smem();
      issue_rows(in0 + (s ^ 1) * STAGE_F, P, P.img0 + tn / HB, tn % HB, &full[s ^ 1], lane);
    }
    mb_wait(&full[s], (it >> 1) & 1);
    float* stage = in0 + s * STAGE_F;
    float* sl = stage + line * LSTR;
    float2* wk = reinterpret_cast<float2*>(sl);
    const int h = b + HB * line;

    // ---- stage 1: radix 16 on x[j + 20 r] ---------------------------------------------------------------------------
    float2 v[20];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = make_float2(sl[j + 20 * r], sl[IMO + j + 20 * r]);
    if (HAS_P1) {
      long long cs;
      const float* g1 = P.p1 + planar_off(img / P.src_div, P.src_nc, HW, cs) + (long long)h * N + j;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        v[r].x = P.a0 * v[r].x + P.a1 * __ldg(g1 + 20 * r);
        v[r].y = P.a0 * v[r].y + P.a1 * __ldg(g1 + cs + 20 * r);
      }
    } else if (P.a0 != 1.0f) {
#pragma unroll
      for (int r = 0; r < 16; ++r) { v[r].x *= P.a0; v[r].y *= P.a0; }
    }
    if (P.coil) {  // A: x[b] * S[b, n]
      const float2* sp = P.coil + (long long)(img / P.ncoil) * P.coil_sb + (long long)(img % P.ncoil) * HW + (long long)h * N + j;
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = cmul(v[r], __ldg(sp + 20 * r));
    }
    if (P.g_at_load && P.gmode != DINVK_G_NONE) {
      const int mb = img / P.ncoil;
      const float* gp = P.g + (long long)mb * P.gsb + (long long)h * P.gsh + j;
      const float c = P.gcb ? __ldg(P.gcb + mb) : P.gc;
      float m0[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) m0[r] = __ldg(gp + 20 * r);
      gmap<16>(P.gmode, m0, c);
      if (P.gsc == 0) {  // same multiplier on both planes (line masks)
#pragma unroll
        for (int r = 0; r < 16; ++r) { v[r].x *= m0[r]; v[r].y *= sgn_im * m0[r]; }
      } else {
        float m1[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) m1[r] = __ldg(gp + P.gsc + 20 * r);
        gmap<16>(P.gmode, m1, c);
#pragma unroll
        for (int r = 0; r < 16; ++r) { v[r].x *= m0[r]; v[r].y *= sgn_im * m1[r]; }
      }
    } else if (P.inverse) {
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r].y = -v[r].y;
    }
    Dft<16>::run(v);
    __syncthreads();  // every thread of the CTA has read its staged inputs: the lines become work rows
    {
      float2* w = wk + 17 * j;  // autosort: out[16 j + r'] (pad16)
#pragma unroll
      for (int r = 0; r < 16; ++r) w[r] = v[r];
    }
    __syncthreads();
    // ---- stage 2: radix 20 on work[j + 16 r] * w320^(r j), threads j < 16 ---------------------------------------------
    if (j < 16) {
#pragma unroll
      for (int r = 0; r < 20; ++r) v[r] = wk[j + 17 * r];
#pragma unroll
      for (int r = 1; r < 20; ++r) v[r] = cmul(v[r], tws[r * 16 + j]);
      Dft<20>::run(v);
    }
    __syncthreads();  // stage-2 inputs consumed
    if (j < 16) {
#pragma unroll
      for (int r = 0; r < 20; ++r) wk[j + 17 * r] = v[r];  // bin j + 16 r -> pad16 index
    }
    __syncthreads();
    // ---- column stage: thread = column, radix 16 over the tile's lines ------------------------------------------------
    {
      const float2* cl = reinterpret_cast<const float2*>(stage) + tid + (tid >> 4);
#pragma unroll
      for (int a = 0; a < 16; ++a) v[a] = cl[a * (LSTR / 2)];
    }
    __syncwarp();
    if (lane == 0) mb_arrive(&empty[s]);
    Dft<16>::run(v);
    {
      const float sb = ((P.centered && (b & 1)) ? -1.0f : 1.0f) * sw;
      float2* o = P.ws + ((long long)iml * N + b) * N + tid;  // ws[img - img0][k_lo][b][w], row = k_lo * 20 + b
      o[0] = make_float2(sb * v[0].x, sb * v[0].y);
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        float2 tk = twf[(b * k) % N];
        tk.x *= sb; tk.y *= sb;
        o[(long long)k * HB * N] = cmul(v[k], tk);
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 2) sp320_pass2(const Params P) {
  DINVK_SP_DYN_SMEM();
  float* ring = reinterpret_cast<float*>(sp_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + P2_STAGES * P2_TILE_F);
  uint64_t* empty = full + P2_STAGES;
  const int tid = threadIdx.x, lane = tid & 31;
  const int ntiles = P.B * HA;  // (img, k_lo)
  constexpr long long HW = (long long)N * N;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < P2_STAGES; ++i) { mb_init(&full[i], 1); mb_init(&empty[i], NT / 32); }
    fence_mbar_init();
    griddep_wait();
    griddep_launch();
    const int t = blockIdx.x;
    if (t < ntiles) {
      mb_expect_tx(&full[0], P2_TILE_F * 4);
      bulk_g2s(ring, P.ws + (long long)t * HB * N, P2_TILE_F * 4, &full[0]);
    }
  }
  __syncthreads();
  const float s2 = 1.0f / 320.0f;
  const float sgn_im = P.inverse ? -1.0f : 1.0f;
  const bool mult = (P.gmode != DINVK_G_NONE) && !P.g_at_load;

  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it & 1;
    const int img = P.img0 + (t >> 4), klo = t & 15;
    const int tn = t + gridDim.x;
    if (tid < 32 && tn < ntiles) {
      if (it >= 1) mb_wait(&empty[s ^ 1], ((it - 1) >> 1) & 1);
      if (lane == 0) {
        fence_async_smem();
        mb_expect_tx(&full[s ^ 1], P2_TILE_F * 4);
        bulk_g2s(ring + (s ^ 1) * P2_TILE_F, P.ws + (long long)tn * HB * N, P2_TILE_F * 4, &full[s ^ 1]);
      }
      __syncwarp();
    }
    mb_wait(&full[s], (it >> 1) & 1);
    float2 u[20];
    {
      const float2* src = reinterpret_cast<const float2*>(ring + s * P2_TILE_F) + tid;
#pragma unroll
      for (int b = 0; b < 20; ++b) u[b] = src[b * N];
    }
    __syncwarp();
    if (lane == 0) mb_arrive(&empty[s]);
    Dft<20>::run(u);
    const float f = ((P.centered && (klo & 1)) ? -1.0f : 1.0f) * s2 * P.e0;
    const float fi = sgn_im * f;
    if (mult) {
      const int mb = img / P.ncoil;
      const float c = P.gcb ? __ldg(P.gcb + mb) : P.gc;
      const float* gp = P.g + (long long)mb * P.gsb + (long long)klo * P.gsh + tid;
      if (P.gsh == 0) {  // multiplier independent of the row (line masks): one value per column
        float m[2] = {__ldg(gp), __ldg(gp + P.gsc)};
        gmap<2>(P.gmode, m, c);
#pragma unroll
        for (int k = 0; k < 20; ++k) { u[k].x *= m[0]; u[k].y *= m[1]; }
      } else {
        float m0[20], m1[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) {
          const float* gk = gp + (long long)(16 * k) * P.gsh;
          m0[k] = __ldg(gk);
          m1[k] = __ldg(gk + P.gsc);
        }
        gmap<20>(P.gmode, m0, c);
        gmap<20>(P.gmode, m1, c);
#pragma unroll
        for (int k = 0; k < 20; ++k) { u[k].x *= m0[k]; u[k].y *= m1[k]; }
      }
    }
    if (P.tout) {
      float2* o = P.tout + (long long)img * HW + (long long)klo * N + tid;
#pragma unroll
      for (int k = 0; k < 20; ++k) o[(long long)k * 16 * N] = make_float2(f * u[k].x, fi * u[k].y);
    } else {
      long long cs;
      const long long obase = planar_off(img, P.dst_nc, HW, cs) + (long long)klo * N + tid;
#pragma unroll
      for (int k = 0; k < 20; ++k) {
        const long long o = obase + (long long)k * 16 * N;
        float re = f * u[k].x, im = fi * u[k].y;
        if (P.q0) { re += P.e1 * __ldg(P.q0 + o); im += P.e1 * __ldg(P.q0 + o + cs); }
        if (P.q1) { re += P.e2 * __ldg(P.q1 + o); im += P.e2 * __ldg(P.q1 + o + cs); }
        P.out[o] = re;
        P.out[o + cs] = im;
      }
    }
  }
}

}  // namespace sp320
}  // namespace dinvk
