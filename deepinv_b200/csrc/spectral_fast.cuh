// spectral_fast.cuh — power-of-two fast path of the spectral tile passes (included by spectral.cu).
//
// Same semantics as spectral_pass_kernel (PassParams), specialised so that the kernel is bound by HBM and not
// by integer address arithmetic (the generic kernel spends ~270 instructions per complex element, mostly
// 64-bit index math and divisions):
//   * transform length N = 2^LOGN, threads NTHR and lines per tile LINES are template parameters: every
//     index computation is a shift or a mask, tiles never straddle images (host guarantees H % LINES == 0);
//   * global traffic moves in 128-bit vectors: a thread owns groups of 4 consecutive w
//     (2 x LDG.128 per 4 planar complex elements);
//   * for even N the centring phases are +-1 and the orthonormal scale is real: the pre/post tables disappear
//     (sign flips at load/store; in the fused forward-multiply-inverse pass the signs cancel altogether);
//   * radix stages are compile-time unrolled (16,16[,8|4|2]).
#pragma once
#include "fft_core.cuh"

namespace dinvk {

template <int LOGN, int LOGNS, class Layout>
__device__ __forceinline__ void fast_stages(float2* buf, const float2* __restrict__ tw, int lines, int tid, int nthr, Layout L) {
  constexpr int REM = LOGN - LOGNS;
  if constexpr (REM > 0) {
    constexpr int R = REM >= 4 ? 16 : (1 << REM);
    constexpr int LOGR = REM >= 4 ? 4 : REM;
    float2 v[16];
    stage_load<R>(v, buf, tw, 1 << LOGN, 1 << LOGNS, lines, tid, nthr, L);
    __syncthreads();
    stage_store<R>(v, buf, 1 << LOGN, 1 << LOGNS, lines, tid, nthr, L);
    __syncthreads();
    fast_stages<LOGN, LOGNS + LOGR>(buf, tw, lines, tid, nthr, L);
  }
}

struct FastRowLayout {
  int ls;
  __host__ __device__ __forceinline__ int idx(int line, int n) const { return line * ls + n + (n >> 4); }
  __host__ __device__ __forceinline__ void split(int task, int per_line, int, int& line, int& j) const {
    line = task / per_line;  // per_line is a compile-time power of two after inlining
    j = task - line * per_line;
  }
};

__device__ __forceinline__ void load_mult4(const PassParams& P, int mb, int h, int w, float (&m0)[4], float (&m1)[4]) {
  const long long o = (long long)mb * P.gsb + (long long)h * P.gsh + w;
  const float4 a = __ldg(reinterpret_cast<const float4*>(P.g + o));
  m0[0] = a.x; m0[1] = a.y; m0[2] = a.z; m0[3] = a.w;
  if (P.gsc != 0) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(P.g + o + P.gsc));
    m1[0] = b.x; m1[1] = b.y; m1[2] = b.z; m1[3] = b.w;
  } else {
    m1[0] = a.x; m1[1] = a.y; m1[2] = a.z; m1[3] = a.w;
  }
}

// apply the multiplier to 4 consecutive spectral samples (same semantics as apply_g)
__device__ __forceinline__ void apply_g4(const PassParams& P, float2 (&v)[4], int img, int h, int w) {
  const int mb = img / P.ncoil;
  if (P.gmode == DINVK_G_CMUL || P.gmode == DINVK_G_CMUL_CONJ) {
    const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(P.g) + (long long)mb * P.gsb + (long long)h * P.gsh + w);
    const float4 a = __ldg(mp), b = __ldg(mp + 1);
    const float2 m[4] = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(b.x, b.y), make_float2(b.z, b.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = P.gmode == DINVK_G_CMUL ? cmul(v[i], m[i]) : cmul_conj(v[i], m[i]);
    return;
  }
  float m0[4], m1[4];
  load_mult4(P, mb, h, w, m0, m1);
  if (P.gmode == DINVK_G_SQ) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { m0[i] *= m0[i]; m1[i] *= m1[i]; }
  } else if (P.gmode == DINVK_G_INV_SQ_PLUS_C) {
    const float c = P.gcb ? __ldg(P.gcb + mb) : P.gc;
#pragma unroll
    for (int i = 0; i < 4; ++i) { m0[i] = 1.0f / (m0[i] * m0[i] + c); m1[i] = 1.0f / (m1[i] * m1[i] + c); }
  } else if (P.gmode == DINVK_G_PINV) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { m0[i] = m0[i] > 1e-5f ? 1.0f / m0[i] : 0.f; m1[i] = m1[i] > 1e-5f ? 1.0f / m1[i] : 0.f; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i].x *= m0[i]; v[i].y *= m1[i]; }
}

// COLS = false: tile = LINES full rows of one image (N = W).  COLS = true: tile = 16 columns x N = H rows.
template <int LOGN, bool COLS, int NTHR, int LINES>
__global__ void __launch_bounds__(NTHR, 1024 / NTHR) spectral_fast_kernel(const PassParams P) {
  DINVK_DYN_SMEM(float2, buf);
  constexpr int N = 1 << LOGN;
  constexpr int GROUPS = LINES * N / 4;       // 4-element groups in the tile
  constexpr int GPT = GROUPS / NTHR;          // groups per thread
  static_assert(GROUPS % NTHR == 0 && GPT >= 1, "tile / thread mismatch");
  static_assert(!COLS || LINES == 16, "column strips are 16 wide");
  const int tid = threadIdx.x;
  const long long HW = (long long)P.H * P.W;
  const float scale = rsqrtf((float)N);
  const float s_gamma = (P.centered && ((N >> 1) & 1)) ? -1.0f : 1.0f;  // exp(-i*pi*N/2) for even N
  const bool cen = P.centered != 0;

  int img, h0 = 0, c0 = 0;
  if (COLS) {
    const int strips = P.W >> 4;
    img = blockIdx.x / strips;
    c0 = (blockIdx.x - img * strips) << 4;
  } else {
    const int tiles_per_img = P.H / LINES;
    img = blockIdx.x / tiles_per_img;
    h0 = (blockIdx.x - img * tiles_per_img) * LINES;
  }
  FastRowLayout RL; RL.ls = N + (N >> 4) + 1;
  ColLayout CL; CL.lines = 16;

  // group g of this thread: (line, n0) in tile coordinates and (h, w) in image coordinates
#define DINVK_GROUP_COORDS(g)                                                             \
  int line, n0, h, w;                                                                     \
  if (COLS) { h = (g) >> 2; line = ((g) & 3) << 2; n0 = h; w = c0 + line; }               \
  else { line = (g) >> (LOGN - 2); n0 = ((g) & ((N >> 2) - 1)) << 2; h = h0 + line; w = n0; }

  // ---- LOAD ------------------------------------------------------------------------------------
  {
    long long cs_src;
    const long long base_src = P.tin ? 0 : planar_base(img / P.src_div, P.src_nc, HW, cs_src);
#pragma unroll
    for (int i = 0; i < GPT; ++i) {
      const int g = tid + i * NTHR;
      DINVK_GROUP_COORDS(g)
      float2 v[4];
      const long long off = (long long)h * P.W + w;
      if (P.tin) {
        const float4* tp = reinterpret_cast<const float4*>(P.tin + (long long)img * HW + off);
        const float4 a = tp[0], b = tp[1];
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
      } else {
        const float4 re = __ldg(reinterpret_cast<const float4*>(P.p0 + base_src + off));
        const float4 im = __ldg(reinterpret_cast<const float4*>(P.p0 + base_src + cs_src + off));
        v[0] = make_float2(P.a0 * re.x, P.a0 * im.x); v[1] = make_float2(P.a0 * re.y, P.a0 * im.y);
        v[2] = make_float2(P.a0 * re.z, P.a0 * im.z); v[3] = make_float2(P.a0 * re.w, P.a0 * im.w);
        if (P.p1) {
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(P.p1 + base_src + off));
          const float4 i1 = __ldg(reinterpret_cast<const float4*>(P.p1 + base_src + cs_src + off));
          v[0].x += P.a1 * r1.x; v[0].y += P.a1 * i1.x; v[1].x += P.a1 * r1.y; v[1].y += P.a1 * i1.y;
          v[2].x += P.a1 * r1.z; v[2].y += P.a1 * i1.z; v[3].x += P.a1 * r1.w; v[3].y += P.a1 * i1.w;
        }
      }
      if (P.coil) {
        const float4* sp = reinterpret_cast<const float4*>(P.coil + (long long)(img / P.ncoil) * P.coil_sb + (long long)(img % P.ncoil) * HW + off);
        const float4 a = __ldg(sp), b = __ldg(sp + 1);
        v[0] = cmul(v[0], make_float2(a.x, a.y)); v[1] = cmul(v[1], make_float2(a.z, a.w));
        v[2] = cmul(v[2], make_float2(b.x, b.y)); v[3] = cmul(v[3], make_float2(b.z, b.w));
      }
      if (P.g_at_load) apply_g4(P, v, img, h, w);
      if (P.dir1 != 0) {
        if (P.dir1 > 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k].y = -v[k].y;
        }
        if (cen) {  // pre-phase (-1)^n along the transform axis
          if (COLS) { if (n0 & 1) {
#pragma unroll
              for (int k = 0; k < 4; ++k) { v[k].x = -v[k].x; v[k].y = -v[k].y; } } }
          else { v[1].x = -v[1].x; v[1].y = -v[1].y; v[3].x = -v[3].x; v[3].y = -v[3].y; }
        }
      }
      if (COLS) {
        float4* sp = reinterpret_cast<float4*>(&buf[CL.idx(line, n0)]);
        sp[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        sp[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) buf[RL.idx(line, n0 + k)] = v[k];
      }
    }
  }
  __syncthreads();

  // ---- transform 1 -------------------------------------------------------------------------------
  if (P.dir1 != 0) {
    if (COLS) fast_stages<LOGN, 0>(buf, P.tw, 16, tid, NTHR, CL);
    else fast_stages<LOGN, 0>(buf, P.tw, LINES, tid, NTHR, RL);
  }

  // ---- fused middle (forward -> multiplier -> inverse) ----------------------------------------------
  if (P.dir2 != 0) {
#pragma unroll
    for (int i = 0; i < GPT; ++i) {
      const int g = tid + i * NTHR;
      DINVK_GROUP_COORDS(g)
      float2 v[4];
      if (COLS) {
        const float4* sp = reinterpret_cast<const float4*>(&buf[CL.idx(line, n0)]);
        const float4 a = sp[0], b = sp[1];
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = buf[RL.idx(line, n0 + k)];
      }
      // post-phase of transform 1 and pre-phase of transform 2: the (-1)^k factors cancel; what remains is
      // scale * gamma and the conjugations of the inverse-by-conjugation trick
      const float f = scale * s_gamma;
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[k].x *= f; v[k].y *= f; }
      if (P.dir1 > 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k].y = -v[k].y;
      }
      if (P.g_after) apply_g4(P, v, img, h, w);
      if (P.dir2 > 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k].y = -v[k].y;
      }
      if (COLS) {
        float4* sp = reinterpret_cast<float4*>(&buf[CL.idx(line, n0)]);
        sp[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        sp[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) buf[RL.idx(line, n0 + k)] = v[k];
      }
    }
    __syncthreads();
    if (COLS) fast_stages<LOGN, 0>(buf, P.tw, 16, tid, NTHR, CL);
    else fast_stages<LOGN, 0>(buf, P.tw, LINES, tid, NTHR, RL);
  }

  // ---- STORE ---------------------------------------------------------------------------------------
  {
    const int last_dir = P.dir2 != 0 ? P.dir2 : P.dir1;
    long long cs_dst;
    const long long base_dst = P.tout ? 0 : planar_base(img, P.dst_nc, HW, cs_dst);
#pragma unroll
    for (int i = 0; i < GPT; ++i) {
      const int g = tid + i * NTHR;
      DINVK_GROUP_COORDS(g)
      float2 v[4];
      if (COLS) {
        const float4* sp = reinterpret_cast<const float4*>(&buf[CL.idx(line, n0)]);
        const float4 a = sp[0], b = sp[1];
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = buf[RL.idx(line, n0 + k)];
      }
      if (last_dir != 0) {
        // post-phase: (-1)^k * gamma * scale
        const float f = scale * s_gamma;
        float sg[4] = {f, f, f, f};
        if (cen) {
          if (COLS) { if (n0 & 1) { sg[0] = -f; sg[1] = -f; sg[2] = -f; sg[3] = -f; } }
          else { sg[1] = -f; sg[3] = -f; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k].x *= sg[k]; v[k].y *= sg[k]; }
        if (last_dir > 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k].y = -v[k].y;
        }
      }
      if (P.dir2 == 0 && P.g_after) apply_g4(P, v, img, h, w);
      const long long off = (long long)h * P.W + w;
      if (P.tout) {
        float4* tp = reinterpret_cast<float4*>(P.tout + (long long)img * HW + off);
        tp[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        tp[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
      } else {
        float4 re = make_float4(P.e0 * v[0].x, P.e0 * v[1].x, P.e0 * v[2].x, P.e0 * v[3].x);
        float4 im = make_float4(P.e0 * v[0].y, P.e0 * v[1].y, P.e0 * v[2].y, P.e0 * v[3].y);
        if (P.q0) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(P.q0 + base_dst + off));
          const float4 m = __ldg(reinterpret_cast<const float4*>(P.q0 + base_dst + cs_dst + off));
          re.x += P.e1 * r.x; re.y += P.e1 * r.y; re.z += P.e1 * r.z; re.w += P.e1 * r.w;
          im.x += P.e1 * m.x; im.y += P.e1 * m.y; im.z += P.e1 * m.z; im.w += P.e1 * m.w;
        }
        if (P.q1) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(P.q1 + base_dst + off));
          const float4 m = __ldg(reinterpret_cast<const float4*>(P.q1 + base_dst + cs_dst + off));
          re.x += P.e2 * r.x; re.y += P.e2 * r.y; re.z += P.e2 * r.z; re.w += P.e2 * r.w;
          im.x += P.e2 * m.x; im.y += P.e2 * m.y; im.z += P.e2 * m.z; im.w += P.e2 * m.w;
        }
        *reinterpret_cast<float4*>(P.out + base_dst + off) = re;
        *reinterpret_cast<float4*>(P.out + base_dst + cs_dst + off) = im;
      }
    }
  }
#undef DINVK_GROUP_COORDS
}

}  // namespace dinvk
