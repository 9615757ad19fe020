// tma_tile.cuh — one TMA tensor-map load of an image tile (+ halo) into shared memory, out-of-range elements zero-filled by
// the copy engine: the staging step of the Radon and Blur kernels (north star: "TMA staging into shared memory").
// Host: a 3-D map (W, H, N) over a stack of fp32 images with a (box_w, box_h, 1) box; device: mbarrier + cp.async.bulk.tensor.3d
// (SASS: UTMALDG).  Requirements checked by the callers: 16-byte aligned base and row pitch (W % 4 == 0), box_w * 4 a multiple
// of 16, both box extents <= 256, 128-byte aligned shared-memory destination.
#pragma once
#ifndef DINVK_EMUL
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>

#include "tc_ptx.cuh"

namespace dinvk {
namespace tt {

typedef CUtensorMap TileMap;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// true when the geometry can be expressed as a tensor map and the map was built
static inline bool make_map_f32(TileMap* m, const void* ptr, int W, int H, int N, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc || (reinterpret_cast<uintptr_t>(ptr) & 15) || (W & 3) || (box_w & 3) || box_w > 256 || box_h > 256 || box_w < 1 || box_h < 1) return false;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_3d(void* smem, const TileMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   tc::smem_u32(smem)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// whole-CTA helper: thread 0 arms the barrier and issues the box load; every thread returns when the tile has landed.
// `bar` is a shared 8-byte word used once per kernel (phase 0).
__device__ __forceinline__ void stage_tile(void* smem, const TileMap* m, uint64_t* bar, int x, int y, int n, uint32_t bytes) {
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(bar, bytes);
    tma_load_3d(smem, m, bar, x, y, n);
  }
  tc::mbar_wait(bar, 0);
}

}  // namespace tt
}  // namespace dinvk
#else
namespace dinvk {
namespace tt {
struct TileMap { int unused; };  // host emulation: the cooperative staging loop is used
}
}
#ifndef __grid_constant__
#define __grid_constant__
#endif
#endif
