// rec_mma.cuh - what the tensor-core recurrence kernels (gru_mma.cu, gru_mma_wide.cu, lstm_mma.cu) share: ldmatrix / mma.sync
// wrappers, the mbarrier + bulk-copy exchange over distributed shared memory, named barriers, fast gate functions and the
// cluster launch.  Included inside namespace cpcb200 { namespace { ... } } of each translation unit.
#pragma once

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- hidden-state exchange without a cluster barrier: bulk shared->shared copies into the peers' shared memory complete
// their byte count on the peer's mbarrier, the consumer waits for the byte count of one step.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
// one bulk copy local shared -> a peer's shared memory, completing `bytes` on the peer's mbarrier
__device__ __forceinline__ void bulk_s2s(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) __trap();  // a broken exchange faults instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_mbar_init_cluster() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void pair_sync(int ub) { asm volatile("bar.sync %0, 64;" ::"r"(ub + 1) : "memory"); }
// the four warps that split the k range of one 16-unit block (gru_mma_wide.cu)
__device__ __forceinline__ void quad_sync(int ub) { asm volatile("bar.sync %0, 128;" ::"r"(ub + 1) : "memory"); }
// publish barrier (id 5, all 256 threads): producers arrive without blocking, the issuing warp waits
__device__ __forceinline__ void publish_arrive() { asm volatile("bar.arrive 5, 256;" ::: "memory"); }
__device__ __forceinline__ void publish_sync() { asm volatile("bar.sync 5, 256;" ::: "memory"); }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }


template <class K>
int launch_cluster(const char* name, K kernel, int cs, int nclusters, cudaStream_t st, void** args, size_t dyn_smem = 0) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * nclusters);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = dyn_smem;
  if (cs > 8) CPC_CHECK_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  if (dyn_smem > 0) CPC_CHECK_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CPC_CHECK_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), args));
  CPC_LAUNCHED_N(name, st);
  return 0;
}

