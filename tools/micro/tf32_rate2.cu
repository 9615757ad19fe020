// Micro-probe 2 (not part of the product): what bounds back-to-back tcgen05.mma kind::tf32 at small N — accumulator
// dependency (same TMEM columns), operand fetch, or issue?  A pattern is a short list of MMAs (N, accumulator column,
// A tile, B tile, kind) replayed many times by one issuing thread; prints cycles per pattern entry.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I deepinv_b200/csrc -o tools/micro/_bin/tf32_rate2 tools/micro/tf32_rate2.cu
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace dinvk;

struct Step { int n; int dcol; int atile; int btile; int bf16; int koff; };
struct Pattern { int len; int iters; Step s[16]; };

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, int bf16) {
  if (bf16) {
    tc::umma_bf16_lohi(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
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
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  }
}

__global__ void __launch_bounds__(64, 1) rate_kernel(const __grid_constant__ Pattern P, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_smem);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (warp == 1) {
    constexpr uint32_t HI = tc::desc_hi_sw128(1024);
    const uint32_t a0 = tc::smem_u32(smem) >> 4;                  // A tiles: 16 KB each, 4 of them
    const uint32_t b0 = (tc::smem_u32(smem) + 64 * 1024) >> 4;    // B tiles: 32 KB each, 4 of them
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      if (rep == 2) t0 = clock64();
      for (int it = 0; it < P.iters; ++it) {
        if (tc::elect_one()) {
          for (int i = 0; i < P.len; ++i) {
            const Step& s = P.s[i];
            const uint32_t id = s.bf16 ? tc::make_idesc_bf16(128, s.n) : idesc_tf32(128, s.n);
            umma(tmem + s.dcol, a0 + s.atile * 1024 + s.koff, HI, b0 + s.btile * 2048 + s.koff, HI, id, s.bf16);
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

static void run(const char* tag, Pattern p, long long* dout, int nsm) {
  p.iters = 300;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  rate_kernel<<<nsm, 64, 210 * 1024>>>(p, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", tag, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, dout, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-72s %7.1f clk per MMA  (%7.1f per pattern of %d)\n", tag, (double)mx / p.iters / p.len, (double)mx / p.iters, p.len);
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  long long* dout;
  cudaMalloc(&dout, sizeof(long long) * 256);
  // Step: n, dcol, atile, btile, bf16, koff
  run("bf16 N=64 same D, same A", {1, 0, {{64, 0, 0, 0, 1, 0}}}, dout, nsm);
  run("bf16 N=64 4 k-steps same D", {4, 0, {{64, 0, 0, 0, 1, 0}, {64, 0, 0, 0, 1, 2}, {64, 0, 0, 0, 1, 4}, {64, 0, 0, 0, 1, 6}}}, dout, nsm);
  run("bf16 N=64 round-robin 4 D", {4, 0, {{64, 0, 0, 0, 1, 0}, {64, 64, 0, 0, 1, 0}, {64, 128, 0, 0, 1, 0}, {64, 192, 0, 0, 1, 0}}}, dout, nsm);
  run("bf16 N=128 same D", {1, 0, {{128, 0, 0, 0, 1, 0}}}, dout, nsm);
  run("bf16 N=256 same D", {1, 0, {{256, 0, 0, 0, 1, 0}}}, dout, nsm);
  run("tf32 N=64 same D, same A", {1, 0, {{64, 0, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=64 4 k-steps same D", {4, 0, {{64, 0, 0, 0, 0, 0}, {64, 0, 0, 0, 0, 2}, {64, 0, 0, 0, 0, 4}, {64, 0, 0, 0, 0, 6}}}, dout, nsm);
  run("tf32 N=64 round-robin 2 D, same A", {2, 0, {{64, 0, 0, 0, 0, 0}, {64, 64, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=64 round-robin 4 D, same A", {4, 0, {{64, 0, 0, 0, 0, 0}, {64, 64, 0, 0, 0, 0}, {64, 128, 0, 0, 0, 0}, {64, 192, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=64 round-robin 4 D, 4 A tiles", {4, 0, {{64, 0, 0, 0, 0, 0}, {64, 64, 1, 0, 0, 0}, {64, 128, 2, 0, 0, 0}, {64, 192, 3, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=64 same D, 4 A tiles", {4, 0, {{64, 0, 0, 0, 0, 0}, {64, 0, 1, 0, 0, 0}, {64, 0, 2, 0, 0, 0}, {64, 0, 3, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=128 same D", {1, 0, {{128, 0, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=128 round-robin 2 D", {2, 0, {{128, 0, 0, 0, 0, 0}, {128, 128, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=128 round-robin 4 D", {4, 0, {{128, 0, 0, 0, 0, 0}, {128, 128, 0, 0, 0, 0}, {128, 256, 0, 0, 0, 0}, {128, 384, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=256 same D", {1, 0, {{256, 0, 0, 0, 0, 0}}}, dout, nsm);
  run("tf32 N=256 round-robin 2 D", {2, 0, {{256, 0, 0, 0, 0, 0}, {256, 256, 0, 0, 0, 0}}}, dout, nsm);
  // the conv_tc32 stage as written: per channel block  hi(N=128) hi(N=128) lo(N=64 -> d+64) lo(N=64 -> d+64), one accumulator
  run("tc32 stage, 1 accumulator: [128,128,64,64] x 2 blocks", {8, 0, {{128, 0, 0, 0, 0, 0}, {128, 0, 0, 0, 0, 2}, {64, 64, 0, 0, 0, 4}, {64, 64, 0, 0, 0, 6},
                                                                    {128, 0, 1, 0, 0, 0}, {128, 0, 1, 0, 0, 2}, {64, 64, 1, 0, 0, 4}, {64, 64, 1, 0, 0, 6}}}, dout, nsm);
  // same work, the two channel blocks into two accumulators, interleaved
  run("tc32 stage, 2 accumulators interleaved", {8, 0, {{128, 0, 0, 0, 0, 0}, {128, 128, 1, 0, 0, 0}, {128, 0, 0, 0, 0, 2}, {128, 128, 1, 0, 0, 2},
                                                        {64, 64, 0, 0, 0, 4}, {64, 192, 1, 0, 0, 4}, {64, 64, 0, 0, 0, 6}, {64, 192, 1, 0, 0, 6}}}, dout, nsm);
  // 4 accumulators (two pixel halves x two channel blocks)
  run("tc32, 4 accumulators interleaved (2 halves x 2 blocks)", {16, 0, {{128, 0, 0, 0, 0, 0}, {128, 128, 1, 0, 0, 0}, {128, 256, 2, 0, 0, 0}, {128, 384, 3, 0, 0, 0},
                                                                        {128, 0, 0, 0, 0, 2}, {128, 128, 1, 0, 0, 2}, {128, 256, 2, 0, 0, 2}, {128, 384, 3, 0, 0, 2},
                                                                        {64, 64, 0, 0, 0, 4}, {64, 192, 1, 0, 0, 4}, {64, 320, 2, 0, 0, 4}, {64, 448, 3, 0, 0, 4},
                                                                        {64, 64, 0, 0, 0, 6}, {64, 192, 1, 0, 0, 6}, {64, 320, 2, 0, 0, 6}, {64, 448, 3, 0, 0, 6}}}, dout, nsm);
  // separate main / corr products, all N=64 (3 MMAs per k8), 4 accumulators
  run("3 x N=64 per k8 (main, corr, corr), 2 accumulator pairs", {6, 0, {{64, 0, 0, 0, 0, 0}, {64, 128, 1, 0, 0, 0}, {64, 64, 0, 1, 0, 0}, {64, 192, 1, 1, 0, 0},
                                                                         {64, 64, 0, 0, 0, 4}, {64, 192, 1, 0, 0, 4}}}, dout, nsm);
  return 0;
}
