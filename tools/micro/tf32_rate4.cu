// Micro-probe 4 (slab-shifted A operands) (not part of the product): issue rate of tcgen05.mma kind::tf32 / kind::f16 with compile-time patterns
// (16 MMAs per elected-thread iteration, fully unrolled: the issue loop itself costs nothing per MMA).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I deepinv_b200/csrc -o tools/micro/_bin/tf32_rate3 tools/micro/tf32_rate3.cu
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace dinvk;

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
template <bool BF16, int SBO>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  constexpr uint32_t HI = tc::desc_hi_sw128(1024);
  constexpr uint32_t HIA = tc::desc_hi_sw128(SBO);
  if constexpr (BF16) {
    tc::umma_bf16_lohi(tmem_d, a_lo, HIA, b_lo, HI, idesc, 1u);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(HIA), "r"(b_lo), "r"(HI), "r"(idesc)
        : "memory");
  }
}

// one pattern = 16 MMAs described by constexpr tables
struct Pat { int m, n[16], d[16], a[16], k[16]; };

template <int ID> __host__ __device__ constexpr Pat get_pat();
#define PAT(ID, M, ...) template <> __host__ __device__ constexpr Pat get_pat<ID>() { return Pat{M, __VA_ARGS__}; }
#define R16(x) {x, x, x, x, x, x, x, x, x, x, x, x, x, x, x, x}
#define K4 {0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6}
// A offsets in 16-byte units: one slab position = 8
// 0: kernel order [128,128,64,64] x 2 halves (half = +64), all taps aligned (shift 0)
PAT(0, 128, {128, 128, 64, 64, 128, 128, 64, 64, 128, 128, 64, 64, 128, 128, 64, 64}, {0, 0, 64, 64, 128, 128, 192, 192, 0, 0, 64, 64, 128, 128, 192, 192},
    {0, 0, 0, 0, 64, 64, 64, 64, 0, 0, 0, 0, 64, 64, 64, 64}, {0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6})
// 1: same, shifted by one position (kx = 1)
PAT(1, 128, {128, 128, 64, 64, 128, 128, 64, 64, 128, 128, 64, 64, 128, 128, 64, 64}, {0, 0, 64, 64, 128, 128, 192, 192, 0, 0, 64, 64, 128, 128, 192, 192},
    {8, 8, 8, 8, 72, 72, 72, 72, 8, 8, 8, 8, 72, 72, 72, 72}, {0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6})
// 2: shifted by two positions
PAT(2, 128, {128, 128, 64, 64, 128, 128, 64, 64, 128, 128, 64, 64, 128, 128, 64, 64}, {0, 0, 64, 64, 128, 128, 192, 192, 0, 0, 64, 64, 128, 128, 192, 192},
    {16, 16, 16, 16, 80, 80, 80, 80, 16, 16, 16, 16, 80, 80, 80, 80}, {0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6, 0, 2, 4, 6})
// 3: N=64 only, aligned / 4: N=64 shifted by one
PAT(3, 128, R16(64), R16(0), R16(0), K4)
PAT(4, 128, R16(64), R16(0), R16(8), K4)
// 5: N=128 aligned / 6: shifted
PAT(5, 128, R16(128), R16(0), R16(0), K4)
PAT(6, 128, R16(128), R16(0), R16(8), K4)
// 7: N=256 aligned / 8 shifted
PAT(7, 128, R16(256), R16(0), R16(0), K4)
PAT(8, 128, R16(256), R16(0), R16(8), K4)

template <int ID, bool BF16, int SBO>
__global__ void __launch_bounds__(64, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_smem);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (warp == 1) {
    constexpr Pat P = get_pat<ID>();
    const uint32_t a0 = tc::smem_u32(smem) >> 4;                  // A tiles: 16 KB each (2)
    const uint32_t b0 = (tc::smem_u32(smem) + 64 * 1024) >> 4;    // B tile: up to 32 KB
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      if (rep == 2) t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (tc::elect_one()) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t id = BF16 ? tc::make_idesc_bf16(P.m, P.n[i]) : idesc_tf32(P.m, P.n[i]);
            umma<BF16, SBO>(tmem + P.d[i], a0 + P.a[i] + P.k[i], b0 + P.k[i], id);
          }
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc::umma_commit(&bar);
      __syncwarp();
      tc::mbar_wait(&bar, ph);
      ph ^= 1;
      if (rep == 2) t1 = clock64();
    }
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

template <int ID, bool BF16, int SBO>
static void run(const char* tag, long long* dout, int nsm) {
  const int iters = 200;
  cudaFuncSetAttribute(rate_kernel<ID, BF16, SBO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  rate_kernel<ID, BF16, SBO><<<nsm, 64, 170 * 1024>>>(iters, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", tag, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, dout, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-64s %7.1f clk per MMA\n", tag, (double)mx / iters / 16);
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  long long* dout;
  cudaMalloc(&dout, sizeof(long long) * 256);
  run<3, false, 1024>("tf32 N=64  canonical A (SBO 1024)", dout, nsm);
  run<3, false, 3072>("tf32 N=64  slab A (SBO 3072), aligned", dout, nsm);
  run<4, false, 3072>("tf32 N=64  slab A, shifted by 1 position", dout, nsm);
  run<5, false, 3072>("tf32 N=128 slab A, aligned", dout, nsm);
  run<6, false, 3072>("tf32 N=128 slab A, shifted by 1 position", dout, nsm);
  run<7, false, 3072>("tf32 N=256 slab A, aligned", dout, nsm);
  run<8, false, 3072>("tf32 N=256 slab A, shifted by 1 position", dout, nsm);
  run<0, false, 3072>("tc32 tap [128,128,64,64] x 2 halves, aligned", dout, nsm);
  run<1, false, 3072>("tc32 tap, shifted by 1 position", dout, nsm);
  run<2, false, 3072>("tc32 tap, shifted by 2 positions", dout, nsm);
  run<3, true, 3072>("bf16 N=64  slab A, aligned", dout, nsm);
  run<4, true, 3072>("bf16 N=64  slab A, shifted by 1 position", dout, nsm);
  return 0;
}
