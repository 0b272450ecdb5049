// Micro-benchmark: ceiling of TMA bulk fp32 reductions (cp.reduce.async.bulk ... add.f32, 1 KB rows) into random rows of an
// 8 MB L2-resident buffer - the scatter-add pattern of score_bwd_mma (dz).   nvcc -arch=sm_100a -O3 -o redbench redbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bulk_reduce_add_f32(float* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MODE>  // 0: TMA bulk reduce, 16 rows per trip; 1: red.global.add.v4.f32 by lanes (2 instructions per row)
__global__ void red_kernel(float* dz, int rows, int H, int trips) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stg = reinterpret_cast<float*>(sm) + (size_t)warp * 16 * H;
  for (int i = lane; i < 16 * H; i += 32) stg[i] = 1.f;
  fence_async_smem();
  __syncwarp();
  uint32_t seed = (blockIdx.x * 64 + warp) * 2654435761u + 12345u;
  for (int it = 0; it < trips; it++) {
    seed = seed * 1664525u + 1013904223u;
    const uint32_t r = ((seed >> 8) + lane * 40503u) % (uint32_t)rows;
    if (MODE == 0) {
      if (lane < 16) {
        bulk_wait_read0();
        bulk_reduce_add_f32(dz + (size_t)r * H, s_u32(stg + (size_t)lane * H), H * 4);
        bulk_commit();
      }
    } else {
      for (int rr = 0; rr < 16; rr++) {
        const uint32_t row = __shfl_sync(0xffffffffu, r, rr);
        float* d = dz + (size_t)row * H;
        for (int c = lane * 4; c < H; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(stg + rr * H + c);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    }
  }
  if (MODE == 0 && lane < 16) bulk_wait0();
}

int main() {
  const int rows = 8192, H = 256;
  float* dz;
  cudaMalloc(&dz, (size_t)rows * H * 4);
  cudaMemset(dz, 0, (size_t)rows * H * 4);
  const long long total_rows = 7424LL * 144;  // rows reduced per training step
  for (int mode = 0; mode < 2; mode++)
    for (int wpc : {5, 8, 12}) {
      const int trips = (int)(total_rows / 16 / (148 * wpc));
      const size_t smem = (size_t)wpc * 16 * H * 4;
      auto k = mode == 0 ? red_kernel<0> : red_kernel<1>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<148, wpc * 32, smem>>>(dz, rows, H, trips);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; i++) k<<<148, wpc * 32, smem>>>(dz, rows, H, trips);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 5;
      const double bytes = (double)trips * 16 * 148 * wpc * H * 4;
      printf("mode %d (%s) warps/CTA %2d: %.1f us per step-equivalent, %.2f TB/s of fp32 reductions  [%s]\n", mode,
             mode == 0 ? "TMA bulk reduce" : "red.v4.f32", wpc, ms * 1e3, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
