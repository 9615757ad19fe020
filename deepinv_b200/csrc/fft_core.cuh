// fft_core.cuh — in-shared-memory mixed-radix Stockham FFT used by the spectral operators.
//
// One CTA holds a tile of `lines` independent 1-D signals of length N in shared memory and runs
// all radix stages in place (all butterflies of a stage are read into registers, one barrier, then
// written back in autosort order).  Radices 16/8/4/2/5/3 cover every size the configs use
// (256 = 16*16, 320 = 16*4*5, 1024 = 16*16*4, 2048 = 16*16*8); other sizes take the O(N^2) path
// in spectral.cu.  Only the forward kernel exp(-2*pi*i*k*n/N) is implemented: the inverse is
// conj(fft(conj(.))) and the conjugations are folded into the load/store phases by the caller.
#pragma once
#include "common.cuh"

namespace dinvk {

struct FftPlan {
  int n;
  int nstages;
  int radix[12];
};
// device-side form: radices packed 5 bits each (dynamic indexing of a kernel-parameter array would
// force the whole parameter block into local memory)
struct FftPlanDev {
  int n;
  int nstages;
  unsigned long long packed;
};
inline FftPlanDev pack_plan(const FftPlan& p) {
  FftPlanDev d;
  d.n = p.n; d.nstages = p.nstages; d.packed = 0;
  for (int s = 0; s < p.nstages; ++s) d.packed |= (unsigned long long)p.radix[s] << (5 * s);
  return d;
}

// returns true when n factors completely into the supported radices
inline bool make_fft_plan(int n, FftPlan* p) {
  p->n = n;
  p->nstages = 0;
  if (n < 1) return false;
  int m = n;
  const int cands[6] = {16, 8, 4, 2, 5, 3};
  for (int ci = 0; ci < 6; ++ci) {
    const int c = cands[ci];
    while (m % c == 0 && p->nstages < 12) {
      p->radix[p->nstages++] = c;
      m /= c;
    }
  }
  return m == 1;
}

// ---- register butterflies (forward DFT, natural order in and out) ----------------------------
template <int R>
struct Dft;

template <>
struct Dft<2> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <>
struct Dft<3> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    const float s = 0.86602540378443864676f;
    float2 t = cadd(v[1], v[2]);
    float2 d = csub(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
    float2 nn = make_float2(s * d.y, -s * d.x);  // -i*s*d
    v[0] = cadd(v[0], t);
    v[1] = cadd(m, nn);
    v[2] = csub(m, nn);
  }
};

template <>
struct Dft<4> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    float2 a0 = cadd(v[0], v[2]);
    float2 a1 = csub(v[0], v[2]);
    float2 a2 = cadd(v[1], v[3]);
    float2 d = csub(v[1], v[3]);
    float2 a3 = make_float2(d.y, -d.x);  // -i*d
    v[0] = cadd(a0, a2);
    v[1] = cadd(a1, a3);
    v[2] = csub(a0, a2);
    v[3] = csub(a1, a3);
  }
};

template <>
struct Dft<5> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
    float2 d1 = csub(v[1], v[4]), d2 = csub(v[2], v[3]);
    float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    float2 n1 = make_float2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
    float2 n2 = make_float2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
    v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    v[1] = make_float2(m1.x + n1.y, m1.y - n1.x);  // m1 - i*n1
    v[4] = make_float2(m1.x - n1.y, m1.y + n1.x);  // m1 + i*n1
    v[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
    v[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
  }
};

template <>
struct Dft<8> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    const float h = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    Dft<4>::run(e);
    Dft<4>::run(o);
    // o[k] *= W8^k
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    o[2] = make_float2(o[2].y, -o[2].x);
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = cadd(e[k], o[k]);
      v[k + 4] = csub(e[k], o[k]);
    }
  }
};

template <>
struct Dft<16> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    const float c = 0.92387953251128673848f, s = 0.38268343236508978178f, h = 0.70710678118654752440f;
    // W16^m = exp(-2*pi*i*m/16), m = 0..9
    const float wr[10] = {1.f, c, h, s, 0.f, -s, -h, -c, -1.f, -c};
    const float wi[10] = {0.f, -s, -h, -c, -1.f, -c, -h, -s, 0.f, s};
    float2 y[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      float2 t[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
      Dft<4>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < 4; ++k1) {
        const int m = n2 * k1;  // compile-time after unrolling: the trivial and the 45-degree twiddles are special-cased
        if (m == 0) y[n2][k1] = t[k1];
        else if (m == 4) y[n2][k1] = make_float2(t[k1].y, -t[k1].x);
        else if (m == 2) y[n2][k1] = make_float2(h * (t[k1].x + t[k1].y), h * (t[k1].y - t[k1].x));
        else if (m == 6) y[n2][k1] = make_float2(h * (t[k1].y - t[k1].x), -h * (t[k1].x + t[k1].y));
        else y[n2][k1] = cmul(t[k1], make_float2(wr[m], wi[m]));
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      float2 t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
      Dft<4>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
    }
  }
};


// 20-point DFT as 5 x 4 Cooley-Tukey: n = 4 n1 + n2, k = k1 + 5 k2
//   X[k1 + 5 k2] = sum_n2 w4^(n2 k2) [ w20^(n2 k1) sum_n1 x[4 n1 + n2] w5^(n1 k1) ]
template <>
struct Dft<20> {
  static __host__ __device__ __forceinline__ void run(float2* v) {
    // w20^m = exp(-2 pi i m / 20) for the products n2 * k1 that occur (m = 1,2,3,4,6,8,9,12)
    const float c1 = 0.95105651629515357212f, s1 = 0.30901699437494742410f;   // m = 1
    const float c2 = 0.80901699437494742410f, s2 = 0.58778525229247312917f;   // m = 2
    const float c3 = 0.58778525229247312917f, s3 = 0.80901699437494742410f;   // m = 3
    const float c4 = 0.30901699437494742410f, s4 = 0.95105651629515357212f;   // m = 4
    float2 T[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      float2 t[5] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2], v[16 + n2]};
      Dft<5>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < 5; ++k1) T[n2][k1] = t[k1];
    }
    // twiddles (real, imag) of w20^m: (cos, -sin)
    T[1][1] = cmul(T[1][1], make_float2(c1, -s1));
    T[1][2] = cmul(T[1][2], make_float2(c2, -s2));
    T[1][3] = cmul(T[1][3], make_float2(c3, -s3));
    T[1][4] = cmul(T[1][4], make_float2(c4, -s4));
    T[2][1] = cmul(T[2][1], make_float2(c2, -s2));
    T[2][2] = cmul(T[2][2], make_float2(c4, -s4));
    T[2][3] = cmul(T[2][3], make_float2(-c4, -s4));   // m = 6
    T[2][4] = cmul(T[2][4], make_float2(-c2, -s2));   // m = 8
    T[3][1] = cmul(T[3][1], make_float2(c3, -s3));
    T[3][2] = cmul(T[3][2], make_float2(-c4, -s4));   // m = 6
    T[3][3] = cmul(T[3][3], make_float2(-c1, -s1));   // m = 9
    T[3][4] = cmul(T[3][4], make_float2(-c2, s2));    // m = 12
#pragma unroll
    for (int k1 = 0; k1 < 5; ++k1) {
      float2 u[4] = {T[0][k1], T[1][k1], T[2][k1], T[3][k1]};
      Dft<4>::run(u);
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) v[k1 + 5 * k2] = u[k2];
    }
  }
};

// ---- shared-memory tile layouts ---------------------------------------------------------------
// ROW: each line is contiguous (padded by one float2 every 16 to break power-of-two strides);
//      consecutive threads take consecutive butterflies of the same line.
// COL: the tile is a strip [n][line] (line fastest); consecutive threads take consecutive lines.
__host__ __device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

struct RowLayout {
  int ls;  // line stride in float2
  __host__ __device__ __forceinline__ int idx(int line, int n) const { return line * ls + pad16(n); }
  __host__ __device__ __forceinline__ void split(int task, int per_line, int /*lines*/, int& line, int& j) const {
    line = task / per_line;
    j = task - line * per_line;
  }
};
struct ColLayout {
  int lines;
  __host__ __device__ __forceinline__ int idx(int line, int n) const { return n * lines + line; }
  __host__ __device__ __forceinline__ void split(int task, int /*per_line*/, int nlines, int& line, int& j) const {
    j = task / nlines;
    line = task - j * nlines;
  }
};
__host__ __device__ inline int row_line_stride(int n) { return pad16(n - 1) + 1 + (((pad16(n - 1) + 1) & 1) ? 0 : 1); }

// One Stockham stage of radix R over `lines` signals held in `buf`, split into the read+butterfly
// half and the autosort write half so that the block barrier between them sits outside the radix
// switch (one shared 16-element register file for every radix).
// Precondition: lines * N <= 16 * nthr (each thread owns at most 16/R butterflies; the stage fits a
// 64-register budget: 4 CTAs of 256 threads per SM).
// `tw` is the N-entry table exp(-2*pi*i*k/N) in global memory (L1-resident).
template <int R, class Layout>
__device__ __forceinline__ void stage_load(float2* v, const float2* buf, const float2* __restrict__ tw, int N, int Ns,
                                           int lines, int tid, int nthr, Layout L) {
  constexpr int T = 16 / R;
  const int per_line = N / R;
  const int total = lines * per_line;
  const int twstep = N / (Ns * R);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int task = tid + t * nthr;
    if (task < total) {
      int line, j;
      L.split(task, per_line, lines, line, j);
      const int k = j % Ns;
      float2 x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = buf[L.idx(line, j + r * per_line)];
      if (Ns > 1) {
#pragma unroll
        for (int r = 1; r < R; ++r) x[r] = cmul(x[r], __ldg(&tw[r * k * twstep]));
      }
      Dft<R>::run(x);
#pragma unroll
      for (int r = 0; r < R; ++r) v[t * R + r] = x[r];
    }
  }
}
template <int R, class Layout>
__device__ __forceinline__ void stage_store(const float2* v, float2* buf, int N, int Ns, int lines, int tid, int nthr,
                                            Layout L) {
  constexpr int T = 16 / R;
  const int per_line = N / R;
  const int total = lines * per_line;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int task = tid + t * nthr;
    if (task < total) {
      int line, j;
      L.split(task, per_line, lines, line, j);
      const int k = j % Ns;
      const int base = (j / Ns) * Ns * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) buf[L.idx(line, base + r * Ns)] = v[t * R + r];
    }
  }
}

// all stages of a plan; data in natural order before and after
template <class Layout>
__device__ __forceinline__ void fft_tile(float2* buf, const float2* __restrict__ tw, const FftPlanDev& plan,
                                         int lines, int tid, int nthr, Layout L) {
  int Ns = 1;
  float2 v[16];
  for (int s = 0; s < plan.nstages; ++s) {
    const int R = (int)((plan.packed >> (5 * s)) & 31ull);
    switch (R) {
      case 16: stage_load<16>(v, buf, tw, plan.n, Ns, lines, tid, nthr, L); break;
      case 8: stage_load<8>(v, buf, tw, plan.n, Ns, lines, tid, nthr, L); break;
      case 4: stage_load<4>(v, buf, tw, plan.n, Ns, lines, tid, nthr, L); break;
      case 2: stage_load<2>(v, buf, tw, plan.n, Ns, lines, tid, nthr, L); break;
      case 5: stage_load<5>(v, buf, tw, plan.n, Ns, lines, tid, nthr, L); break;
      default: stage_load<3>(v, buf, tw, plan.n, Ns, lines, tid, nthr, L); break;
    }
    __syncthreads();
    switch (R) {
      case 16: stage_store<16>(v, buf, plan.n, Ns, lines, tid, nthr, L); break;
      case 8: stage_store<8>(v, buf, plan.n, Ns, lines, tid, nthr, L); break;
      case 4: stage_store<4>(v, buf, plan.n, Ns, lines, tid, nthr, L); break;
      case 2: stage_store<2>(v, buf, plan.n, Ns, lines, tid, nthr, L); break;
      case 5: stage_store<5>(v, buf, plan.n, Ns, lines, tid, nthr, L); break;
      default: stage_store<3>(v, buf, plan.n, Ns, lines, tid, nthr, L); break;
    }
    __syncthreads();
    Ns *= R;
  }
}

}  // namespace dinvk
