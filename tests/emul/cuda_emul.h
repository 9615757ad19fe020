// cuda_emul.h — TEST-ONLY host emulation of the CUDA SIMT subset used by libdinvk's kernels.
//
// This file is test infrastructure (like oracle/): it lets `pytest -m "not gpu"` execute the very
// same kernel sources on CPU threads in the GPU-less authoring container, to check index math,
// barriers and launch geometry before spending GPU minutes.  It is never compiled into
// libdinvk.so, never imported by the deepinv_b200 package, and there is no code path from the
// product to it.  One block runs at a time; each CUDA thread is a host thread; __syncthreads is a
// real barrier; warp shuffles rendezvous through a per-warp exchange buffer.
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <pthread.h>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef void* cudaStream_t;

inline thread_local uint3 threadIdx{0, 0, 0};
inline thread_local uint3 blockIdx{0, 0, 0};
inline dim3 blockDim, gridDim;

namespace emul {

// pthread barriers are futex-based: a waiter sleeps in the kernel and the last arrival wakes the group with one syscall,
// far cheaper than a mutex + condition variable when 256 host threads share 8 cores
class Barrier {
 public:
  Barrier() = default;
  Barrier(const Barrier&) = delete;
  Barrier& operator=(const Barrier&) = delete;
  ~Barrier() { if (init_) pthread_barrier_destroy(&b_); }
  void reset(int n) {
    if (init_ && n == n_) return;
    if (init_) pthread_barrier_destroy(&b_);
    pthread_barrier_init(&b_, nullptr, (unsigned)n);
    n_ = n; init_ = true;
  }
  void wait() { pthread_barrier_wait(&b_); }
 private:
  pthread_barrier_t b_;
  int n_ = 0;
  bool init_ = false;
};

struct State {
  Barrier block_bar;
  std::vector<std::unique_ptr<Barrier>> warp_bar;
  std::vector<uint64_t> xchg;  // per-thread exchange slot for shuffles
  std::vector<unsigned char> dyn;
  int nthreads = 0;
};
inline State& state() { static State s; return s; }
inline void* dyn_smem() { return state().dyn.data(); }

inline int linear_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }

// Persistent worker pool: one host thread per CUDA thread slot, created once and reused by every launch (a 256-thread
// CTA used to cost 256 pthread_create / join per launch).  Worker t runs the kernel body for thread t of every block.
class Pool {
 public:
  static Pool& get() { static Pool p; return p; }
  void run(int nt, const std::function<void(int)>& job) {
    std::unique_lock<std::mutex> lk(m_);
    while ((int)workers_.size() < nt) {
      const int t = (int)workers_.size();
      workers_.emplace_back([this, t]() { loop(t); });
    }
    job_ = &job; nt_ = nt; pending_ = nt; ++gen_;
    cv_start_.notify_all();
    cv_done_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }
 private:
  Pool() = default;
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true; ++gen_;
    }
    cv_start_.notify_all();
    for (auto& w : workers_) w.join();
  }
  void loop(int t) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_start_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        if (t < nt_) job = job_;
      }
      if (job) {
        (*job)(t);
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::mutex m_;
  std::condition_variable cv_start_, cv_done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* job_ = nullptr;
  int nt_ = 0, pending_ = 0;
  unsigned long gen_ = 0;
  bool stop_ = false;
};

template <typename F>
void launch(dim3 grid, dim3 block, size_t smem, F body) {
  State& s = state();
  gridDim = grid; blockDim = block;
  const int nt = block.x * block.y * block.z;
  s.nthreads = nt;
  s.block_bar.reset(nt);
  const int nwarps = (nt + 31) / 32;
  while ((int)s.warp_bar.size() < nwarps) s.warp_bar.emplace_back(new Barrier());
  for (int w = 0; w < nwarps; ++w) s.warp_bar[w]->reset(std::min(32, nt - 32 * w));
  s.xchg.assign(nt, 0);
  s.dyn.assign(smem + 64, 0);
  const long nblocks = (long)grid.x * grid.y * grid.z;
  const std::function<void(int)> job = [&](int t) {
    threadIdx.x = t % block.x;
    threadIdx.y = (t / block.x) % block.y;
    threadIdx.z = t / (block.x * block.y);
    for (long b = 0; b < nblocks; ++b) {
      blockIdx.x = (unsigned)(b % grid.x);
      blockIdx.y = (unsigned)((b / grid.x) % grid.y);
      blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
      body();
      s.block_bar.wait();  // block boundary: statics / dyn smem are reused by the next block
    }
  };
  Pool::get().run(nt, job);
}

template <typename T>
inline T shfl_generic(T v, int src_lane) {
  State& s = state();
  const int tid = linear_tid(), warp = tid / 32, lane = tid % 32;
  uint64_t bits = 0;
  static_assert(sizeof(T) <= 8, "shuffle payload too large");
  std::memcpy(&bits, &v, sizeof(T));
  s.xchg[tid] = bits;
  s.warp_bar[warp]->wait();
  int src = warp * 32 + (src_lane & 31);
  if (src >= s.nthreads) src = tid;
  uint64_t r = s.xchg[src];
  s.warp_bar[warp]->wait();
  T out;
  std::memcpy(&out, &r, sizeof(T));
  (void)lane;
  return out;
}
}  // namespace emul

static inline void __syncthreads() { emul::state().block_bar.wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::state().warp_bar[emul::linear_tid() / 32]->wait(); }

template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emul::shfl_generic(v, (emul::linear_tid() % 32) ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
  int lane = emul::linear_tid() % 32;
  return emul::shfl_generic(v, lane + d < 32 ? lane + d : lane);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emul::shfl_generic(v, src); }

template <typename T> static inline T __ldg(const T* p) { return *p; }

static inline float atomicAdd(float* addr, float v) {
  static std::mutex m;
  std::lock_guard<std::mutex> g(m);
  float old = *addr; *addr = old + v; return old;
}
static inline int atomicAdd(int* addr, int v) { return __atomic_fetch_add(addr, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* addr, unsigned v) { return __atomic_fetch_add(addr, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* addr, unsigned long long v) { return __atomic_fetch_add(addr, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned* addr, unsigned v) {
  unsigned old = __atomic_load_n(addr, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(addr, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline long long __float2ll_rn(float x) { return (long long)__builtin_llrintf(x); }
static inline long long __double2ll_rn(double x) { return (long long)__builtin_llrint(x); }
static inline unsigned __float_as_uint(float x) { unsigned u; __builtin_memcpy(&u, &x, 4); return u; }
static inline float __uint_as_float(unsigned u) { float x; __builtin_memcpy(&x, &u, 4); return x; }
static inline int atomicAnd(int* addr, int v) { return __atomic_fetch_and(addr, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* addr, int v) { return __atomic_exchange_n(addr, v, __ATOMIC_SEQ_CST); }

static inline int atomicOr(int* addr, int v) { return __atomic_fetch_or(addr, v, __ATOMIC_SEQ_CST); }

// fp16 (cuda_fp16.h subset used by the split32h format): IEEE binary16 through the compiler's _Float16, round-to-nearest-even
struct uint2 { unsigned x, y; };
struct __half {
  _Float16 v;
  __half() = default;
  __half(float f) : v((_Float16)f) {}
  explicit operator float() const { return (float)v; }
};
struct __half2 { __half x, y; };
static inline __half __float2half_rn(float f) { return __half(f); }
static inline float __half2float(__half h) { return (float)h.v; }
static inline unsigned short __half_as_ushort(__half h) { unsigned short u; __builtin_memcpy(&u, &h.v, 2); return u; }
static inline __half2 __floats2half2_rn(float a, float b) { __half2 r; r.x = __half(a); r.y = __half(b); return r; }
static inline float2 __half22float2(__half2 h) { return float2{(float)h.x.v, (float)h.y.v}; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline int __float2int_rd(float a) { return (int)std::floor(a); }
static inline int __float2int_rn(float a) { return (int)std::nearbyint(a); }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }

// minimal runtime stubs so host wrappers compile
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
