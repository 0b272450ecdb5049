// Micro-benchmark: ceiling of scatter-adds of 256-wide gradient rows into random rows of an 8 MB L2-resident buffer - the
// pattern of score_bwd_mma (dz).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/redbench tools/redbench.cu
//   mode 0: TMA bulk reduce  cp.reduce.async.bulk ... add.f32, one 1 KB row per instruction (what score_bwd_mma does)
//   mode 1: red.global.add.v4.f32 by lanes (2 instructions per row)
//   mode 2: TMA bulk reduce  add.noftz.bf16, one 512 B row per instruction (half the bytes, bf16 accumulation)
//   mode 3: TMA bulk STORE   cp.async.bulk.global.shared::cta, 1 KB rows (same bytes, no read-modify-write): L2 write ceiling
//   mode 4: TMA bulk reduce f32, 1 KB rows, 32 rows in flight per warp instead of 16
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bulk_reduce_add_f32(float* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_bf16(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MODE>
__global__ void red_kernel(float* dz, int rows, int H, int trips) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TILES = MODE == 4 ? 2 : 1;
  float* stg = reinterpret_cast<float*>(sm) + (size_t)warp * 16 * TILES * H;
  for (int i = lane; i < 16 * TILES * H; i += 32) stg[i] = MODE == 2 ? 0.f : 1.f;
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
    } else if (MODE == 4) {
      if (lane < 16) {
        bulk_wait_read1();
        bulk_reduce_add_f32(dz + (size_t)r * H, s_u32(stg + (size_t)((it & 1) * 16 + lane) * H), H * 4);
        bulk_commit();
      }
    } else if (MODE == 2) {
      if (lane < 16) {
        bulk_wait_read0();
        bulk_reduce_add_bf16(reinterpret_cast<unsigned char*>(dz) + (size_t)r * H * 2, s_u32(stg + (size_t)lane * H), H * 2);
        bulk_commit();
      }
    } else if (MODE == 3) {
      if (lane < 16) {
        bulk_wait_read0();
        bulk_store(dz + (size_t)r * H, s_u32(stg + (size_t)lane * H), H * 4);
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
  if (MODE != 1 && lane < 16) bulk_wait0();
}

int main() {
  const int rows = 8192, H = 256;
  float* dz;
  cudaMalloc(&dz, (size_t)rows * H * 4);
  cudaMemset(dz, 0, (size_t)rows * H * 4);
  const long long total_rows = 7424LL * 140;  // rows reduced per training step (128 negatives + 12 positives per anchor)
  const char* names[5] = {"TMA reduce f32 1KB", "red.v4.f32", "TMA reduce bf16 512B", "TMA store 1KB", "TMA reduce f32 1KB x2 in flight"};
  for (int mode = 0; mode < 5; mode++)
    for (int wpc : {5, 8, 12}) {
      const int trips = (int)(total_rows / 16 / (148 * wpc));
      const size_t smem = (size_t)wpc * 16 * H * 4 * (mode == 4 ? 2 : 1);
      void (*k)(float*, int, int, int) = mode == 0 ? red_kernel<0> : mode == 1 ? red_kernel<1> : mode == 2 ? red_kernel<2>
                                       : mode == 3 ? red_kernel<3> : red_kernel<4>;
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
      const double nrows = (double)trips * 16 * 148 * wpc;
      const double bytes = nrows * H * (mode == 2 ? 2 : 4);
      printf("mode %d (%-32s) warps/CTA %2d: %7.1f us per step-equivalent, %6.2f Grows/s, %.2f TB/s  [%s]\n", mode, names[mode], wpc,
             ms * 1e3, nrows / ms / 1e6, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
