// blur.cu — direct (im2col-free) 2-D convolution for Blur.A / Blur.A_adjoint, all five paddings.
//
// Replaces deepinv/physics/functional/convolution.py:42-164 (conv2d / conv_transpose2d: flip + F.pad +
// grouped F.conv2d, and F.conv_transpose2d + _apply_transpose_padding :641-758) for the reference's
// per-sample / per-channel filter broadcast (:761-787).
//
// Everything is one tiled correlation kernel
//     out[i,j] = sum_{u,v} ks[u,v] * in[ map(i + u + off_i), map(j + v + off_j) ]
// where ks is the filter staged in shared memory (flipped for a true convolution), `map` resolves the
// padding rule in index space (zeros / circular / replicate / reflect) while the input tile + halo is
// staged, and (off_i, off_j) place the kernel origin:
//     A,  same-size paddings : flip, off = h/2 - (h-1),           map = padding rule, out (H, W)
//     A,  valid              : flip, off = 0,                      no map,             out (H-h+1, W-w+1)
//     A^T valid              : no flip, off = -(h-1), zeros,       in (H-h+1, W-w+1),  out (H, W)
//     A^T circular/constant  : no flip, off = -h/2, circular/zeros                     out (H, W)
//     A^T replicate/reflect  : no flip, off = -(h-1), zeros -> extended (H+h-1, W+w-1) image, then the
//                              transpose of the padding folds the border strips back (fold kernel).
// A CTA computes a 64x64 output tile; each thread owns 4x4 outputs and slides a 4-wide register window
// along the filter row (one 128-bit shared load per 16 FMAs per row).  Filters up to ~128x128 fit.
#include "common.cuh"
#include "tma_tile.cuh"

#include <algorithm>
#include <cstdlib>

namespace dinvk {

constexpr int BL_T = 64;  // tile edge

struct BlurParams {
  int C, Hin, Win, Hout, Wout, h, w, wp;  // wp = w rounded up to a multiple of 4
  int FB, FC, flip, off_i, off_j, map;    // map: 0 zeros, 1 circular, 2 replicate, 3 reflect
  int PW;                                 // patch row pitch (floats)
  int shift;                              // zero taps prepended to every filter row (0..3): makes off_j a multiple of 4
  int tiles_x;
};

__device__ __forceinline__ int map_index(int a, int n, int mode, bool& ok) {
  ok = true;
  if (mode == 1) { a %= n; return a < 0 ? a + n : a; }
  if (mode == 2) return a < 0 ? 0 : (a >= n ? n - 1 : a);
  if (mode == 3) { if (a < 0) a = -a; if (a > n - 1) a = 2 * (n - 1) - a; ok = (a >= 0 && a < n); return ok ? a : 0; }
  ok = (a >= 0 && a < n);
  return ok ? a : 0;
}

__global__ void __launch_bounds__(256) blur_corr_kernel(const __grid_constant__ tt::TileMap tmap, int use_tma,
                                                        const float* __restrict__ in, const float* __restrict__ filt,
                                                        float* __restrict__ out, BlurParams P) {
  DINVK_DYN_SMEM(float, smem);
  float* ks = smem;                 // [h][wp]
  // [(BL_T + h - 1)][PW], 128-byte aligned at run time (TMA destination; the launch reserves the slack)
  float* patch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem + P.h * P.wp) + 127) & ~static_cast<uintptr_t>(127));
  const int tid = threadIdx.x;
  const int bc = blockIdx.y;
  const int b = bc / P.C, c = bc - b * P.C;
  const int i0 = (blockIdx.x / P.tiles_x) * BL_T, j0 = (blockIdx.x % P.tiles_x) * BL_T;
  const float* f = filt + ((long long)(P.FB == 1 ? 0 : b) * P.FC + (P.FC == 1 ? 0 : c)) * P.h * P.w;
  for (int idx = tid; idx < P.h * P.wp; idx += 256) {
    const int u = idx / P.wp, v = idx - u * P.wp - P.shift;   // `shift` leading zero taps: see launch_corr
    float val = 0.f;
    if (v >= 0 && v < P.w) val = P.flip ? __ldg(f + (P.h - 1 - u) * P.w + (P.w - 1 - v)) : __ldg(f + u * P.w + v);
    ks[idx] = val;
  }
  const int PH = BL_T + P.h - 1;
  const float* src = in + (long long)bc * P.Hin * P.Win;
  // The input tile + halo.  Zero padding (and every tile whose halo stays inside the image, whatever the padding rule) is ONE
  // tensor-map box load: the copy engine zero-fills out-of-range elements.  Tiles that touch the border under circular /
  // replicate / reflect padding resolve the rule in index space with the cooperative loop.
  bool by_tma = false;
#ifndef DINVK_EMUL
  if (use_tma) {
    const int r0 = i0 + P.off_i, c0 = j0 + P.off_j;
    by_tma = (P.map == 0) || (r0 >= 0 && r0 + PH <= P.Hin && c0 >= 0 && c0 + P.PW <= P.Win);
    if (by_tma) {   // (uniform over the CTA)
      __shared__ __align__(8) uint64_t tile_bar;
      tt::stage_tile(patch, &tmap, &tile_bar, c0, r0, bc, (uint32_t)(PH * P.PW * 4));
    }
  }
#endif
  if (!by_tma) {
    for (int idx = tid; idx < PH * P.PW; idx += 256) {
      const int pr = idx / P.PW, pc = idx - pr * P.PW;
      bool okr, okc;
      const int gi = map_index(i0 + pr + P.off_i, P.Hin, P.map, okr);
      const int gj = map_index(j0 + pc + P.off_j, P.Win, P.map, okc);
      patch[idx] = (okr && okc) ? __ldg(src + (long long)gi * P.Win + gj) : 0.f;
    }
  }
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
  const int nchunk = P.wp >> 2;
  for (int u = 0; u < P.h; ++u) {
    float4 lo[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) lo[r] = *reinterpret_cast<const float4*>(&patch[(ty * 4 + r + u) * P.PW + tx * 4]);
    const float4* krow = reinterpret_cast<const float4*>(&ks[u * P.wp]);
    for (int vb = 0; vb < nchunk; ++vb) {
      const float4 k4 = krow[vb];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 hi = *reinterpret_cast<const float4*>(&patch[(ty * 4 + r + u) * P.PW + tx * 4 + 4 * vb + 4]);
        const float win[8] = {lo[r].x, lo[r].y, lo[r].z, lo[r].w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a = acc[r][q];
          a = fmaf(k4.x, win[q], a);
          a = fmaf(k4.y, win[q + 1], a);
          a = fmaf(k4.z, win[q + 2], a);
          a = fmaf(k4.w, win[q + 3], a);
          acc[r][q] = a;
        }
        lo[r] = hi;
      }
    }
  }
  float* dst = out + (long long)bc * P.Hout * P.Wout;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = i0 + ty * 4 + r;
    if (gi >= P.Hout) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int gj = j0 + tx * 4 + q;
      if (gj < P.Wout) dst[(long long)gi * P.Wout + gj] = acc[r][q];
    }
  }
}

// transpose of replicate / reflect padding: x[p,q] = sum over the extended indices that the padding maps onto (p,q)
// z is the extended image (He = H + h - 1 rows starting at extended index a_min = h/2 - h + 1; same for columns)
__device__ __forceinline__ int preimages(int p, int n, int amin, int amax, int mode, int* start, int* count) {
  // returns the number of ranges; each range is [start, start+count)
  if (mode == 2) {  // replicate
    int lo = p, hi = p;
    if (p == 0) lo = amin;
    if (p == n - 1) hi = amax;
    start[0] = lo; count[0] = hi - lo + 1;
    return 1;
  }
  int k = 0;
  start[k] = p; count[k] = 1; ++k;
  if (p >= 1 && -p >= amin) { start[k] = -p; count[k] = 1; ++k; }
  if (p <= n - 2 && 2 * (n - 1) - p <= amax) { start[k] = 2 * (n - 1) - p; count[k] = 1; ++k; }
  return k;
}

__global__ void __launch_bounds__(256) blur_fold_kernel(const float* __restrict__ z, float* __restrict__ x, int H, int W, int h,
                                                        int w, int mode) {
  const int bc = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W) return;
  const int p = idx / W, q = idx - p * W;
  const int amin = h / 2 - h + 1, amax = H - 1 + h / 2, bmin = w / 2 - w + 1, bmax = W - 1 + w / 2;
  const int We = W + w - 1, He = H + h - 1;
  int rs[3], rc[3], cs[3], cc[3];
  const int nr = preimages(p, H, amin, amax, mode, rs, rc), ncol = preimages(q, W, bmin, bmax, mode, cs, cc);
  const float* zz = z + (long long)bc * He * We;
  float acc = 0.f;
  for (int a = 0; a < nr; ++a)
    for (int ra = rs[a]; ra < rs[a] + rc[a]; ++ra)
      for (int bq = 0; bq < ncol; ++bq)
        for (int cb = cs[bq]; cb < cs[bq] + cc[bq]; ++cb) acc += __ldg(zz + (long long)(ra - amin) * We + (cb - bmin));
  x[(long long)bc * H * W + idx] = acc;
}

static int run_corr(const float* in, const float* filt, float* out, int B, int C, int Hin, int Win, int Hout, int Wout,
                    int FB, int FC, int h, int w, int flip, int off_i, int off_j, int map, void* stream) {
  BlurParams P;
  // A tensor-map box has to start on a 16-byte boundary of the image row.  Tiles start at multiples of 64 columns, so the patch
  // origin j0 + off_j is aligned iff off_j is a multiple of 4: prepend shift = off_j mod 4 zero taps to the filter rows and move
  // the origin left by as much (out[j] = sum_v ks'[v] in[j + v + off_j - shift], ks'[v] = ks[v - shift]): same sums, aligned box.
  const int shift = ((off_j % 4) + 4) % 4;
  off_j -= shift;
  P.shift = shift;
  P.C = C; P.Hin = Hin; P.Win = Win; P.Hout = Hout; P.Wout = Wout; P.h = h; P.w = w; P.wp = (w + shift + 3) & ~3;
  P.FB = FB; P.FC = FC; P.flip = flip; P.off_i = off_i; P.off_j = off_j; P.map = map;
  P.PW = BL_T + P.wp + 4;  // window reads reach column tx*4 + wp + 3
  P.tiles_x = ceil_div(Wout, BL_T);
  const size_t smem = sizeof(float) * ((size_t)h * P.wp + (size_t)(BL_T + h - 1) * P.PW) + 128;
  if (smem > 200 * 1024) return set_error(DINVK_EUNSUPPORTED, "blur: filter %dx%d too large for the tiled kernel", h, w);
  int rc = allow_smem(blur_corr_kernel, smem);
  if (rc) return rc;
  const long long tiles = (long long)P.tiles_x * ceil_div(Hout, BL_T);
  if ((long long)B * C > 65535 || tiles > 2147483647LL) return set_error(DINVK_EINVAL, "blur: grid too large");
  tt::TileMap tmap = tt::TileMap();
  int use_tma = 0;
#ifndef DINVK_EMUL
  if (!getenv("DINVK_NO_TMA_STAGING")) use_tma = tt::make_map_f32(&tmap, in, Win, Hin, B * C, P.PW, BL_T + h - 1) ? 1 : 0;
#endif
  DINVK_LAUNCH(blur_corr_kernel, dim3((unsigned)tiles, B * C), dim3(256), smem, stream, tmap, use_tma, in, filt, out, P);
  return DINVK_POST_LAUNCH();
}

static int check_blur_args(const void* a, const void* f, const void* o, int B, int C, int H, int W, int FB, int FC, int h, int w,
                           int padding) {
  DINVK_CHECK_ARG(a && f && o, "blur: null pointer");
  DINVK_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1 && h >= 1 && w >= 1, "blur: bad shape");
  DINVK_CHECK_ARG((FB == 1 || FB == B) && (FC == 1 || FC == C), "blur: filter batch/channel (%d,%d) must be 1 or match (%d,%d)", FB, FC, B, C);
  DINVK_CHECK_ARG(padding >= DINVK_PAD_VALID && padding <= DINVK_PAD_CONSTANT, "blur: unknown padding %d", padding);
  DINVK_CHECK_ARG(padding != DINVK_PAD_VALID || (H >= h && W >= w), "blur: valid padding needs an image at least as large as the filter");
  DINVK_CHECK_ARG(padding != DINVK_PAD_REFLECT || (h / 2 < H && w / 2 < W), "blur: reflect padding needs pad < image size");
  return 0;
}

}  // namespace dinvk

using namespace dinvk;

extern "C" int dinvk_blur_fwd(const float* x, const float* filt, float* y, int B, int C, int H, int W, int FB, int FC, int h,
                              int w, int padding, void* stream) {
  int rc = check_blur_args(x, filt, y, B, C, H, W, FB, FC, h, w, padding);
  if (rc) return rc;
  if (B == 0) return DINVK_OK;
  if (padding == DINVK_PAD_VALID) return run_corr(x, filt, y, B, C, H, W, H - h + 1, W - w + 1, FB, FC, h, w, 1, 0, 0, 0, stream);
  const int map = padding == DINVK_PAD_CIRCULAR ? 1 : padding == DINVK_PAD_REPLICATE ? 2 : padding == DINVK_PAD_REFLECT ? 3 : 0;
  return run_corr(x, filt, y, B, C, H, W, H, W, FB, FC, h, w, 1, h / 2 - (h - 1), w / 2 - (w - 1), map, stream);
}

extern "C" size_t dinvk_blur_adj_workspace_bytes(int B, int C, int H, int W, int h, int w, int padding) {
  if (padding == DINVK_PAD_REPLICATE || padding == DINVK_PAD_REFLECT)
    return sizeof(float) * (size_t)B * C * (size_t)(H + h - 1) * (size_t)(W + w - 1) + 256;
  return 256;
}

extern "C" int dinvk_blur_adj(const float* y, const float* filt, float* x, int B, int C, int H, int W, int FB, int FC, int h,
                              int w, int padding, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_blur_args(y, filt, x, B, C, H, W, FB, FC, h, w, padding);
  if (rc) return rc;
  if (B == 0) return DINVK_OK;
  if (padding == DINVK_PAD_VALID)
    return run_corr(y, filt, x, B, C, H - h + 1, W - w + 1, H, W, FB, FC, h, w, 0, -(h - 1), -(w - 1), 0, stream);
  if (padding == DINVK_PAD_CIRCULAR || padding == DINVK_PAD_CONSTANT)
    return run_corr(y, filt, x, B, C, H, W, H, W, FB, FC, h, w, 0, -(h / 2), -(w / 2), padding == DINVK_PAD_CIRCULAR ? 1 : 0, stream);
  // replicate / reflect: zero-extended correlation on the padded domain, then fold the border strips back
  const size_t need = dinvk_blur_adj_workspace_bytes(B, C, H, W, h, w, padding);
  if (!workspace || workspace_bytes < need) return set_error(DINVK_EWORKSPACE, "dinvk_blur_adj: workspace %zu < %zu", workspace_bytes, need);
  float* z = reinterpret_cast<float*>(((uintptr_t)workspace + 127) & ~(uintptr_t)127);
  const int He = H + h - 1, We = W + w - 1;
  // extended output index a' = a - a_min with a_min = h/2 - h + 1: in index = a + u - h/2 = a' + u + (a_min - h/2) = a' + u - (h-1)
  if ((rc = run_corr(y, filt, z, B, C, H, W, He, We, FB, FC, h, w, 0, -(h - 1), -(w - 1), 0, stream))) return rc;
  DINVK_LAUNCH(blur_fold_kernel, dim3(ceil_div((long long)H * W, 256), B * C), dim3(256), 0, stream, (const float*)z, x, H, W, h, w,
               padding == DINVK_PAD_REPLICATE ? 2 : 3);
  return DINVK_POST_LAUNCH();
}
