// Micro-benchmark (not part of the product): issue rate of tcgen05.mma kind::f16 (bf16 -> fp32, M = 128) with both operands
// in shared memory, for the operand layouts conv_tc.cu uses.  One CTA per SM, one issuing thread; the operands are zeros
// (timing does not depend on the data).  Prints cycles per MMA for:
//   N in {16, 64, 128, 256};  A canonical (SBO 1024) vs slab (SBO 2048 / 3072, start shifted by tap);  B fixed vs cycling tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I deepinv_b200/csrc -o gpurun_out/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace dinvk;

struct Args { int n; uint32_t sbo_a; int shift_taps; int cycle_b; int mh; int iters; int order; };

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Args a, long long* out) {
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
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, N);
    const uint32_t HI_A = tc::desc_hi_sw128(a.sbo_a);
    constexpr uint32_t HI_B = tc::desc_hi_sw128(1024);
    const uint32_t a0 = tc::smem_u32(smem) >> 4;
    const uint32_t b0 = (tc::smem_u32(smem) + 112 * 1024) >> 4;  // B tiles after a 112 KB A region
    const int slab_x = a.sbo_a / 128;
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {  // rep 0/1 warm, rep 2 timed
      if (rep == 2) t0 = clock64();
      for (int it = 0; it < a.iters; ++it) {
        if (tc::elect_one()) {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_lo = a0 + (a.shift_taps ? static_cast<uint32_t>(((tap / 3) * slab_x + (tap % 3)) * 8) : 0u);
            const uint32_t b_lo = b0 + (a.cycle_b ? static_cast<uint32_t>((tap % 3) * (N * 128 >> 4)) : 0u);
            if (a.order == 0) {  // tap -> half -> k
              for (int h = 0; h < a.mh; ++h) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc::umma_bf16_lohi(tmem + h * N, a_lo + 64 * h + 2 * k, HI_A, b_lo + 2 * k, HI_B, idesc, 1u);
              }
            } else {  // tap -> k -> half: consecutive MMAs share the B operand
#pragma unroll
              for (int k = 0; k < 4; ++k)
                for (int h = 0; h < a.mh; ++h)
                  tc::umma_bf16_lohi(tmem + h * N, a_lo + 64 * h + 2 * k, HI_A, b_lo + 2 * k, HI_B, idesc, 1u);
            }
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

template <int N>
static void run(const char* tag, Args a, long long* dout, int nsm) {
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  mma_rate_kernel<N><<<nsm, 128, 220 * 1024>>>(a, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", tag, cudaGetErrorString(e)); return; }
  long long h[256];
  cudaMemcpy(h, dout, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < nsm; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  const double n_mma = (double)a.iters * 9 * a.mh * 4;
  printf("%-52s N=%3d mh=%d  %7.1f clk/MMA (min over SMs %7.1f)  floor %d\n", tag, N, a.mh, mx / n_mma, mn / n_mma, 128 * N / 256);
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  long long* dout;
  cudaMalloc(&dout, sizeof(long long) * 256);
  const int it = 200;
  for (int all = 1; all < 2; ++all) {
    const int g = all ? nsm : 1;
    printf("---- %d CTA(s)\n", g);
    run<64>("canonical A (SBO 1024), fixed B", {64, 1024, 0, 0, 1, it}, dout, g);
    run<64>("slab A (SBO 2048), tap shifts, fixed B", {64, 2048, 1, 0, 1, it}, dout, g);
    run<64>("slab A (SBO 2048), tap shifts, cycling B", {64, 2048, 1, 1, 1, it}, dout, g);
    run<64>("slab A (SBO 3072), tap shifts, cycling B, 2 halves", {64, 3072, 1, 1, 2, it}, dout, g);
    run<64>("slab A (SBO 3072), 2 halves, order tap-k-half", {64, 3072, 1, 1, 2, it, 1}, dout, g);
    run<64>("slab A (SBO 5120), 4 halves, order tap-half-k", {64, 5120, 1, 1, 4, it, 0}, dout, g);
    run<64>("slab A (SBO 5120), 4 halves, order tap-k-half", {64, 5120, 1, 1, 4, it, 1}, dout, g);
    run<64>("canonical A, 2 halves same B, order k-half", {64, 1024, 0, 1, 2, it, 1}, dout, g);
    run<16>("slab A (SBO 2048), tap shifts, cycling B", {16, 2048, 1, 1, 1, it}, dout, g);
    run<128>("canonical A, fixed B", {128, 1024, 0, 0, 1, it}, dout, g);
    run<128>("slab A (SBO 2048), tap shifts, cycling B", {128, 2048, 1, 1, 1, it}, dout, g);
    run<128>("slab A (SBO 3072), tap shifts, cycling B, 2 halves", {128, 3072, 1, 1, 2, it}, dout, g);
    run<128>("slab A (SBO 3072), 2 halves, order tap-k-half", {128, 3072, 1, 1, 2, it, 1}, dout, g);
    run<256>("canonical A, fixed B", {256, 1024, 0, 0, 1, it}, dout, g);
    run<256>("slab A (SBO 3072), tap shifts, cycling B, 2 halves", {256, 3072, 1, 1, 2, it}, dout, g);
  }
  return 0;
}
