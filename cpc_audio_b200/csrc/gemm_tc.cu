// gemm_tc.cu - tcgen05 / TMEM / TMA GEMM kernels (bf16 operands, fp32 accumulation) for the bf16 path.
//
//   gemm_nt_tc : C[m,n] = sum_k A[m,k] B[n,k] (+bias)    both operands K-major in shared memory (SWIZZLE_128B)
//                conv1-4 forward / dgrad, GRU input projection, prediction heads and their data gradients.
//   gemm_tn_tc : C[n1,n2] += sum_m A[m,n1] B[m,n2]       both operands MN-major (reduction dim = rows), split
//                over row blocks across CTAs, fp32 red.global.add epilogue.  Every weight gradient.
//
// A conv-style row view (RowView.taps / .s) is loaded through a 4-D tensor map (c, phase, group, batch) with
// dims (C, s, rows/s, nb): source row = s*group + phase, so tap j of output row t is the box at
// (c0, j % s, t + j / s, b) - no im2col, no overlapping strides.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane group = warp % 4).  One output tile per CTA, 2 CTAs per SM co-resident so
// that one CTA's epilogue overlaps the other's main loop.
#include <stdlib.h>

#include <cstring>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cpcb200 {

namespace {

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int NT_THREADS = 192;

// per-CTA cycle stamps of the last gemm_nt_tc2 launch (debug hook cpcb200_debug_gemm_timeline): 8 slots per CTA
__device__ unsigned long long g_nt2_tl[148 * 8];
// compiled in only with -DCPC_B200_TIMELINE (tools/gemm_probe.py builds that way): release builds carry no stamps
#ifdef CPC_B200_TIMELINE
#define TL_STAMP(slot) do { if (blockIdx.x < 148) g_nt2_tl[blockIdx.x * 8 + (slot)] = (unsigned long long)(clock64() - tl_t0); } while (0)
#else
#define TL_STAMP(slot) do { } while (0)
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 4-D bf16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost first; strides in ELEMENTS.
int make_map4(CUtensorMap* m, const void* base, const unsigned long long dims[4], const unsigned long long strides_el[3],
              const unsigned box[4]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(CPCB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides_el[0] * 2, strides_el[1] * 2, strides_el[2] * 2};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(CPCB200_ERR_BAD_DIMS, "tensor map base not 16-B aligned");
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CPCB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u %u %u",
                (int)r, dims[0], dims[1], dims[2], dims[3], strides_el[0], strides_el[1], strides_el[2], box[0], box[1], box[2], box[3]);
  return 0;
}

// map of a RowView: dims (cin, s, groups, nb); box (64, 1, box_rows, 1)
int make_rowview_map(CUtensorMap* m, const RowView& v, int inner, int nb, int box_rows, bool exact_rows) {
  const int cin = inner / v.taps;
  if (cin % 64 != 0) return fail(CPCB200_ERR_BAD_DIMS, "row view: %d channels per tap not a multiple of 64", cin);
  if (v.taps > 1 && v.rs != (long long)v.s * cin) return fail(CPCB200_ERR_BAD_DIMS, "row view: rs != s*cin");
  const long long group_stride = v.rs;                                   // elements between consecutive logical rows
  const long long phase_stride = v.taps > 1 ? cin : v.rs;                // elements between consecutive source rows
  const unsigned long long groups = exact_rows ? (unsigned long long)v.rpb : (unsigned long long)(v.rpb + (v.taps - 1) / v.s);
  const long long bstride = nb > 1 ? v.bs : (long long)groups * group_stride;
  unsigned long long dims[4] = {(unsigned long long)cin, (unsigned long long)v.s, groups, (unsigned long long)nb};
  unsigned long long st[3] = {(unsigned long long)phase_stride, (unsigned long long)group_stride, (unsigned long long)bstride};
  unsigned box[4] = {64, 1, (unsigned)box_rows, 1};
  return make_map4(m, v.p, dims, st, box);
}

__device__ __forceinline__ void store_out(float* p, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; i++) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void store_out(bf16* p, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; j++) h[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
    reinterpret_cast<uint4*>(p)[i] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// NT kernel
// ---------------------------------------------------------------------------------------------------------
template <int BN, int STAGES, class TO>
__global__ void __launch_bounds__(NT_THREADS) gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, int nkb,
                                                                 int chunks_per_tap, int s, int tiles_per_batch, int N,
                                                                 const float* __restrict__ bias, OutView C) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~uintptr_t(1023));
  unsigned char* smA = sm;
  unsigned char* smB = sm + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int b = blockIdx.y / tiles_per_batch;
  const int t0 = (blockIdx.y - b * tiles_per_batch) * BM;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; i++) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; kb++) {
        const int st = kb % STAGES, it = kb / STAGES;
        if (it > 0) ptx::mbar_wait(&empty[st], (it - 1) & 1);
        ptx::mbar_arrive_expect_tx(&full[st], A_BYTES + B_BYTES);
        const int tap = kb / chunks_per_tap, c0 = (kb - tap * chunks_per_tap) * BK;
        ptx::tma_load_4d(&tmA, &full[st], smA + st * A_BYTES, c0, tap % s, t0 + tap / s, b);
        ptx::tma_load_4d(&tmB, &full[st], smB + st * B_BYTES, kb * BK, n0, 0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
      for (int kb = 0; kb < nkb; kb++) {
        const int st = kb % STAGES, it = kb / STAGES;
        ptx::mbar_wait(&full[st], it & 1);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smA + st * A_BYTES), b0 = ptx::smem_u32(smB + st * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UK; k++) {
          const uint64_t ad = ptx::make_sdesc_sw128(a0 + k * UK * 2, 16, 1024);
          const uint64_t bd = ptx::make_sdesc_sw128(b0 + k * UK * 2, 16, 1024);
          ptx::umma_bf16(tmem_acc, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[st]);
      }
      ptx::umma_commit(acc_full);
    }
  } else {
    // epilogue: thread <-> accumulator row (TMEM lane), 32 columns at a time
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int t = t0 + row;
    const bool row_ok = out_row_ok(C, t, n0);
    TO* crow = static_cast<TO*>(C.p) + (long long)b * C.bs + (long long)t * C.rs + n0;
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (row_ok) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
        if (bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] += __ldg(bias + n0 + c0 + j);
        }
        if (C.relu) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        }
        store_out(crow + c0, v);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_acc, BN);
}

// ---------------------------------------------------------------------------------------------------------
// TN kernel (weight gradients): C[n1, n2] += sum over row blocks.  Operands MN-major.
//   A tile  : [64 rows][128 n1]  = 2 swizzled [64][64] blocks   (UMMA M = 128)
//   B tile  : [64 rows][BN n2]   = BN/64 swizzled [64][64] blocks (UMMA N = BN)
// ---------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(NT_THREADS) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB, int kb_total,
                                                                 int kb_per_cta, int kb_per_batch, int b_chunks_per_tap,
                                                                 int b_s, int N1, int N2, float* __restrict__ Cacc, int ldc,
                                                                 int mode, int Ci, int taps) {
  constexpr uint32_t BLK = 64 * 64 * 2;  // one swizzled [64 rows][64 ch] block
  constexpr uint32_t A_BYTES = 2 * BLK, B_BYTES = (BN / 64) * BLK;
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~uintptr_t(1023));
  unsigned char* smA = sm;
  unsigned char* smB = sm + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n20 = blockIdx.x * BN, n10 = blockIdx.y * BM;
  const int kb_beg = blockIdx.z * kb_per_cta;
  const int kb_end = min(kb_total, kb_beg + kb_per_cta);
  const int nkb = kb_end - kb_beg;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; i++) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, BN);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // B operand columns n2 = tap*Ci' + c  ->  (tap, c0) per 64-wide block
      for (int i = 0; i < nkb; i++) {
        const int kb = kb_beg + i;
        const int st = i % STAGES, it = i / STAGES;
        if (it > 0) ptx::mbar_wait(&empty[st], (it - 1) & 1);
        ptx::mbar_arrive_expect_tx(&full[st], A_BYTES + B_BYTES);
        const int bb = kb / kb_per_batch, r0 = (kb - bb * kb_per_batch) * 64;
#pragma unroll
        for (int j = 0; j < 2; j++) ptx::tma_load_4d(&tmA, &full[st], smA + st * A_BYTES + j * BLK, n10 + j * 64, 0, r0, bb);
#pragma unroll
        for (int j = 0; j < BN / 64; j++) {
          const int blk = (n20 >> 6) + j;
          const int tap = blk / b_chunks_per_tap, c0 = (blk - tap * b_chunks_per_tap) * 64;
          ptx::tma_load_4d(&tmB, &full[st], smB + st * B_BYTES + j * BLK, c0, tap % b_s, r0 + tap / b_s, bb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 1, 1);
      for (int i = 0; i < nkb; i++) {
        const int st = i % STAGES, it = i / STAGES;
        ptx::mbar_wait(&full[st], it & 1);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(smA + st * A_BYTES), b0 = ptx::smem_u32(smB + st * B_BYTES);
#pragma unroll
        for (int k = 0; k < 64 / UK; k++) {
          // MN-major SW128: LBO = stride between 64-wide MN blocks, SBO = stride between 8-row K groups
          const uint64_t ad = ptx::make_sdesc_sw128(a0 + k * UK * 128, BLK, 1024);
          const uint64_t bd = ptx::make_sdesc_sw128(b0 + k * UK * 128, BLK, 1024);
          ptx::umma_bf16(tmem_acc, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&empty[st]);
      }
      ptx::umma_commit(acc_full);
    }
  } else if (nkb > 0) {
    const int lg = warp & 3;
    const int n1 = n10 + lg * 32 + lane;
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (n1 < N1) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int n2 = n20 + c0 + j;
          if (n2 < N2) {
            long long o;
            if (mode == STORE_CONV_W) { const int tap = n2 / Ci, ci = n2 - tap * Ci; o = ((long long)n1 * Ci + ci) * taps + tap; }
            else o = (long long)n1 * ldc + n2;
            atomicAdd(Cacc + o, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_acc, BN);
}

// ---- epilogue staging (gen 2): a warp parks 32 rows x 64 bytes in shared memory (row stride 80 B: conflict-free
// 16-byte stores) and copies them out with 8 rows x 64 B per instruction instead of 32 rows x 16 B -------------
constexpr int NT2_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant: column halves)
constexpr int STG_RS = 80, STG_WARP = 32 * STG_RS;

__device__ __forceinline__ void stage_put(unsigned char* my_row, const float (&v)[32], int sub, bf16*) {  // 32 bf16 = 64 B
  (void)sub;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; j++) h[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
    reinterpret_cast<uint4*>(my_row)[i] = t;
  }
}
__device__ __forceinline__ void stage_put(unsigned char* my_row, const float (&v)[32], int sub, float*) {  // 16 fp32 = 64 B
#pragma unroll
  for (int i = 0; i < 4; i++)
    reinterpret_cast<float4*>(my_row)[i] = make_float4(v[16 * sub + 4 * i], v[16 * sub + 4 * i + 1], v[16 * sub + 4 * i + 2], v[16 * sub + 4 * i + 3]);
}
// rows whose bit is set in `ok` go to gbase + row * row_bytes (64 contiguous bytes each)
__device__ __forceinline__ void stage_flush(const unsigned char* stg, unsigned char* gbase, long long row_bytes, uint32_t ok, int lane) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int row = 8 * i + (lane >> 2), piece = lane & 3;
    if ((ok >> row) & 1u)
      *reinterpret_cast<uint4*>(gbase + (long long)row * row_bytes + piece * 16) = *reinterpret_cast<const uint4*>(stg + row * STG_RS + piece * 16);
  }
}
// same, as 16-byte fp32 reductions (split-K accumulation of the TN kernel)
__device__ __forceinline__ void stage_flush_red(const unsigned char* stg, float* gbase, long long row_el, uint32_t ok, int lane) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int row = 8 * i + (lane >> 2), piece = lane & 3;
    if ((ok >> row) & 1u) {
      const float4 v = *reinterpret_cast<const float4*>(stg + row * STG_RS + piece * 16);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gbase + (long long)row * row_el + piece * 4), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w) : "memory");
    }
  }
}
// one 32-column chunk of this thread's row -> global through the warp's staging buffer
template <class TO>
__device__ __forceinline__ void emit_chunk(unsigned char* stg, int lane, const float (&v)[32], TO* gchunk, long long rs_el, uint32_t ok) {
  constexpr int SUBS = sizeof(TO) == 2 ? 1 : 2;
#pragma unroll
  for (int sub = 0; sub < SUBS; sub++) {
    stage_put(stg + lane * STG_RS, v, sub, static_cast<TO*>(nullptr));
    __syncwarp();
    stage_flush(stg, reinterpret_cast<unsigned char*>(gchunk + sub * 16), rs_el * (long long)sizeof(TO), ok, lane);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// NT kernel, second generation: persistent, 128 x 256 tiles, the weight tile (B operand, identical for every
// M tile) is loaded ONCE per cluster and multicast by TMA to the CM CTAs of the cluster (each CTA fetches
// 256/CM rows), accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  L2->SM operand traffic per 128x256x64 block drops from 48 KB to 16 + 32/CM KB.
// ---------------------------------------------------------------------------------------------------------
template <int CM, class TO, bool CN>
__global__ void __launch_bounds__(NT2_THREADS, 1) gemm_nt_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                     const __grid_constant__ CUtensorMap tmB, int nkb,
                                                                     int chunks_per_tap, int s, int tiles_per_batch,
                                                                     int m_tiles, int n_tiles, int nb,
                                                                     const float* __restrict__ bias, OutView C, CNormEpi E,
                                                                     HeadBatch HB) {
  constexpr int BN2 = 256, STAGES = 4;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN2 * BK * 2, B_SLICE = B_BYTES / CM;
  constexpr uint16_t MASK = (uint16_t)((1u << CM) - 1);
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~uintptr_t(1023));
  unsigned char* smA = sm;
  unsigned char* smB = sm + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);
  float* cn_par = reinterpret_cast<float*>(sm + STAGES * (A_BYTES + B_BYTES) + 256);  // [3][256]: bias, gamma, beta (CN)
  float2* cn_xch = reinterpret_cast<float2*>(cn_par + 3 * BN2);                       // [tile parity][2 halves][128 rows] (mean, M2)
  unsigned char* stg_all = reinterpret_cast<unsigned char*>(cn_xch + 4 * BM);         // [8 warps][STG_WARP]

#ifdef CPC_B200_TIMELINE
  const long long tl_t0 = clock64();
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x / CM, num_clusters = gridDim.x / CM;
  const int m_groups = (m_tiles + CM - 1) / CM;
  const int total_groups = m_groups * n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; i++) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], CM); }
    for (int i = 0; i < 2; i++) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  // everything above touches no global memory: with a programmatic dependent launch it overlaps the previous kernel's tail
  pdl_wait();
  pdl_trigger();
  if (CN) {
    for (int i = threadIdx.x; i < BN2; i += NT2_THREADS) {
      cn_par[i] = bias != nullptr ? bias[i] : 0.f; cn_par[BN2 + i] = E.gam[i]; cn_par[2 * BN2 + i] = E.bet[i];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  if (threadIdx.x == 0) TL_STAMP(0);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int gid = cluster_id; gid < total_groups; gid += num_clusters) {
        const int gm = gid / n_tiles, gn = gid - gm * n_tiles;
        const int mt = gm * CM + rank;
        const int b = mt < m_tiles ? mt / tiles_per_batch : nb;  // b == nb: out of bounds -> zero fill
        const int t0 = (mt % tiles_per_batch) * BM;
        // stacked per-head weights (HeadBatch): batch b reads A batch b % a_mod and weight rows (b / w_div) * w_rows + n
        const int ba = (HB.a_mod > 0 && b < nb) ? b % HB.a_mod : (HB.a_mod > 0 ? HB.a_mod : b);
        const int n0 = gn * BN2 + (HB.w_div > 0 ? ((b < nb ? b : nb - 1) / HB.w_div) * HB.w_rows : 0);
        int c0 = 0, chunk = 0, tph = 0, tgr = 0;  // channel offset inside the tap; tap = tgr * s + tph
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int st = it % STAGES, u = it / STAGES;
          if (u > 0) ptx::mbar_wait(&empty[st], (u - 1) & 1);
          ptx::mbar_arrive_expect_tx(&full[st], A_BYTES + B_BYTES);
          ptx::tma_load_4d(&tmA, &full[st], smA + st * A_BYTES, c0, tph, t0 + tgr, ba);
          c0 += BK;
          if (++chunk == chunks_per_tap) { chunk = 0; c0 = 0; if (++tph == s) { tph = 0; tgr++; } }
          if (CM == 1) ptx::tma_load_4d(&tmB, &full[st], smB + st * B_BYTES, kb * BK, n0, 0, 0);
          else ptx::tma_load_4d_mc(&tmB, &full[st], smB + st * B_BYTES + rank * B_SLICE, kb * BK, n0 + rank * (BN2 / CM), 0, 0, MASK);
          if (it == 0) TL_STAMP(1);
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop (uniform control flow keeps the descriptors in uniform registers; a single-lane
    // loop costs ~120 issue slots per k-block in register-to-uniform moves, more than the 4 MMAs take to execute);
    // one elected lane issues.  A shared-memory descriptor is the constant part + (byte offset >> 4).
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN2, 0, 0);
    const uint64_t adesc0 = ptx::make_sdesc_sw128(ptx::smem_u32(smA), 16, 1024);
    const uint64_t bdesc0 = ptx::make_sdesc_sw128(ptx::smem_u32(smB), 16, 1024);
    const bool leader = ptx::elect_one();
    int it = 0, ti = 0;
    for (int gid = cluster_id; gid < total_groups; gid += num_clusters, ti++) {
      const int acc = ti & 1, ua = ti >> 1;
      if (ua > 0) ptx::mbar_wait(&tempty[acc], (ua - 1) & 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN2;
      for (int kb = 0; kb < nkb; kb++, it++) {
        const int st = it % STAGES, u = it / STAGES;
        ptx::mbar_wait(&full[st], u & 1);
        ptx::tc_fence_after();
        if (leader) {
          if (it == 0) TL_STAMP(2);
          const uint64_t ad = adesc0 + (uint64_t)((st * A_BYTES) >> 4), bd = bdesc0 + (uint64_t)((st * B_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BK / UK; k++)
            ptx::umma_bf16(d_tmem, ad + (uint64_t)(k * ((UK * 2) >> 4)), bd + (uint64_t)(k * ((UK * 2) >> 4)), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          if (CM == 1) ptx::umma_commit(&empty[st]);
          else ptx::umma_commit_mc(&empty[st], MASK);
        }
        __syncwarp();
      }
      if (leader) {
        ptx::umma_commit(&tfull[acc]);
        TL_STAMP(3);
      }
      __syncwarp();
    }
  } else {
    // epilogue: warp -> TMEM lane quadrant lg = warp % 4 (hardware rule), column half hf = (warp - 2) / 4
    const int lg = warp & 3, hf = (warp - 2) >> 2;
    const int row = lg * 32 + lane;
    unsigned char* stg = stg_all + (warp - 2) * STG_WARP;
    int ti = 0;
    for (int gid = cluster_id; gid < total_groups; gid += num_clusters, ti++) {
      const int gm = gid / n_tiles, gn = gid - gm * n_tiles;
      const int mt = gm * CM + rank;
      const int b = mt / tiles_per_batch;
      const int t = (mt % tiles_per_batch) * BM + row;
      const int n0 = gn * BN2;
      const bool row_ok = mt < m_tiles && out_row_ok(C, t, n0);
      const uint32_t ok = __ballot_sync(0xffffffffu, row_ok);
      // global address of (first row of this warp, first column of this warp's half)
      TO* cwarp = static_cast<TO*>(C.p) + (long long)b * C.bs + (long long)(t - lane) * C.rs + n0 + hf * (BN2 / 2);
      const int acc = ti & 1, ua = ti >> 1;
      ptx::mbar_wait(&tfull[acc], ua & 1);
      ptx::tc_fence_after();
      if (warp == 2 && lane == 0) { if (ti == 0) TL_STAMP(4); TL_STAMP(7); }
      const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * BN2 + hf * (BN2 / 2));
      if constexpr (CN) {
        // u = bf16(acc + bias) exactly as the unfused path stores it; statistics and the ReLU mask are taken from the
        // rounded values so that backward (which re-derives them from the saved u) sees the same numbers.  Each warp
        // owns 128 of the 256 channels: shifted one-pass moments per half, merged across the two warps of the quadrant.
        const float* pb = cn_par + hf * (BN2 / 2);
        float sd = 0.f, sd2 = 0.f, v0 = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN2 / 2; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(trow + c0, r);
          ptx::tmem_ld_wait();
          if (c0 == 0) v0 = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[0]) + pb[0]));
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float d = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j]) + pb[c0 + j])) - v0;
            sd += d; sd2 = fmaf(d, d, sd2);
          }
        }
        constexpr float inv_half = 1.f / (BN2 / 2);
        const float m_mine = v0 + sd * inv_half, q_mine = fmaxf(sd2 - sd * sd * inv_half, 0.f);
        float2* xch = cn_xch + (ti & 1) * 2 * BM;
        xch[hf * BM + row] = make_float2(m_mine, q_mine);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + lg) : "memory");
        const float2 oth = xch[(hf ^ 1) * BM + row];
        const float mean = 0.5f * (m_mine + oth.x);
        const float dm = m_mine - oth.x;
        const float m2 = q_mine + oth.y + dm * dm * (float)(BN2 / 4);
        const float rstd = rsqrtf(m2 * (1.f / (BN2 - 1)) + 1e-5f);
        if (hf == 0 && row_ok && E.stats != nullptr) E.stats[(long long)b * C.rpb + t] = make_float2(mean, rstd);
        bf16* ywarp = E.y != nullptr ? static_cast<bf16*>(E.y) + (long long)b * C.bs + (long long)(t - lane) * C.rs + hf * (BN2 / 2) : nullptr;
        float* zwarp = E.z != nullptr ? E.z + ((long long)b * C.rpb + (t - lane)) * BN2 + hf * (BN2 / 2) : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < BN2 / 2; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(trow + c0, r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j]) + pb[c0 + j]));
          if (E.save_u) emit_chunk<bf16>(stg, lane, v, static_cast<bf16*>(static_cast<void*>(cwarp)) + c0, C.rs, ok);
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaxf(fmaf((v[j] - mean) * rstd, pb[BN2 + c0 + j], pb[2 * BN2 + c0 + j]), 0.f);
          if (ywarp != nullptr) emit_chunk<bf16>(stg, lane, v, ywarp + c0, C.rs, ok);
          if (zwarp != nullptr) emit_chunk<float>(stg, lane, v, zwarp + c0, (long long)BN2, ok);
        }
        if (row_ok && E.y != nullptr && hf == 0) {  // zero rows around the window (the conv padding of the next layer)
          bf16* yrow = static_cast<bf16*>(E.y) + (long long)b * C.bs + (long long)t * C.rs;
          for (int pr = 1; pr <= E.pad_rows; pr++) {
            if (t == 0) {
              uint4* z4 = reinterpret_cast<uint4*>(yrow - (long long)pr * C.rs);
              for (int i = 0; i < BN2 / 8; i++) z4[i] = make_uint4(0, 0, 0, 0);
            }
            if (t == C.rpb - 1) {
              uint4* z4 = reinterpret_cast<uint4*>(yrow + (long long)pr * C.rs);
              for (int i = 0; i < BN2 / 8; i++) z4[i] = make_uint4(0, 0, 0, 0);
            }
          }
        }
      } else {
        uint32_t r[2][32];
        ptx::tmem_ld32(trow, r[0]);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          ptx::tmem_ld_wait();
          if (c + 1 < 4) ptx::tmem_ld32(trow + 32 * (c + 1), r[(c + 1) & 1]);
          const int cc = hf * (BN2 / 2) + 32 * c;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[c & 1][j]);
          if (bias != nullptr) {
            const float* bb = bias + (HB.w_div > 0 ? ((b < nb ? b : nb - 1) / HB.w_div) * HB.w_rows : 0);
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] += __ldg(bb + n0 + cc + j);
          }
          if (C.relu) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
          }
          emit_chunk<TO>(stg, lane, v, cwarp + 32 * c, C.rs, ok);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      if (warp == 2 && lane == 0) TL_STAMP(5);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 0) TL_STAMP(6);
}

constexpr size_t nt2_smem() { return (size_t)4 * (BM * BK * 2 + 256 * BK * 2) + 256 + 3 * 256 * 4 + 4 * BM * 8 + 8 * STG_WARP + 1024; }

template <int CM, class TO, bool CN = false>
int launch_nt2(const CUtensorMap& tmA, const CUtensorMap& tmB, int nkb, int cpt, int s, int tpb, int m_tiles, int n_tiles, int nb,
               const float* bias, const OutView& C, cudaStream_t st, const CNormEpi& E = CNormEpi{},
               const HeadBatch& HB = HeadBatch{}) {
  auto k = gemm_nt_tc2_kernel<CM, TO, CN>;
  const size_t smem = nt2_smem();
  CPC_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int total_groups = ((m_tiles + CM - 1) / CM) * n_tiles;
  int clusters = (148 - sm_reserve()) / CM;  // sm_reserve() > 0: another kernel runs beside this one on SMs of its own
  if (clusters > total_groups) clusters = total_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CM);
  cfg.blockDim = dim3(NT2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CM; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CPC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k, tmA, tmB, nkb, cpt, s, tpb, m_tiles, n_tiles, nb, bias, C, E, HB));
  CPC_LAUNCHED_N(CN ? "gemm_nt_cnorm_tc2" : "gemm_nt_tc2", st);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// TN kernel, second generation: 128 x 256 output tiles; the two CTAs of a cluster take the two n1 tiles of the
// same (n2 tile, row range) and share the B operand: each fetches half of its four [64 rows][64 ch] blocks and
// multicasts them.  4-stage ring, 256 TMEM columns, fp32 red.global.add epilogue.
// ---------------------------------------------------------------------------------------------------------
// Up to four independent problems share one launch ("grouped" split-K): the CTA range of the grid is cut into one
// slice per problem (TnProblem::cta_begin), so that the small weight-gradient GEMMs of a backward pass fill the SMs the
// big one leaves idle and pay one launch / prologue / tail instead of four.
struct TnProblem {
  int cta_begin, n2_tiles, kb_total, kb_per_cta, kb_per_batch, b_chunks_per_tap, b_s, N1, N2, ldc, mode, Ci, taps;
  float* Cacc;
};
struct TnGroup {
  CUtensorMap tmA[4];
  CUtensorMap tmB[4];
  TnProblem pr[4];
  int n;
};

template <int CM>
__global__ void __launch_bounds__(NT2_THREADS, 1) gemm_tn_tc2_kernel(const __grid_constant__ TnGroup G) {
  int pid = 0;
#pragma unroll
  for (int i = 1; i < 4; i++)
    if (i < G.n && (int)blockIdx.x >= G.pr[i].cta_begin) pid = i;
  const TnProblem& pr = G.pr[pid];
  const CUtensorMap* tmA = &G.tmA[pid];
  const CUtensorMap* tmB = &G.tmB[pid];
  const int kb_total = pr.kb_total, kb_per_cta = pr.kb_per_cta, kb_per_batch = pr.kb_per_batch;
  const int b_chunks_per_tap = pr.b_chunks_per_tap, b_s = pr.b_s, N1 = pr.N1, N2 = pr.N2, ldc = pr.ldc;
  const int mode = pr.mode, Ci = pr.Ci, taps = pr.taps;
  float* __restrict__ Cacc = pr.Cacc;
  constexpr int BN2 = 256, STAGES = 4;
  constexpr uint32_t BLK = 64 * 64 * 2;
  constexpr uint32_t A_BYTES = 2 * BLK, B_BYTES = 4 * BLK;
  constexpr uint16_t MASK = (uint16_t)((1u << CM) - 1);
  extern __shared__ __align__(1024) unsigned char smraw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smraw) + 1023) & ~uintptr_t(1023));
  unsigned char* smA = sm;
  unsigned char* smB = sm + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);
  unsigned char* stg_all = sm + STAGES * (A_BYTES + B_BYTES) + 256;  // [8 warps][STG_WARP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();     // == n1 tile % CM (consecutive CTAs = consecutive n1 tiles)
  const int local = (int)blockIdx.x - pr.cta_begin, n1_tiles = N1 / BM;
  const int n1t = local % n1_tiles, rest = local / n1_tiles;
  const int n2t = rest % pr.n2_tiles, split = rest / pr.n2_tiles;
  const int n20 = n2t * BN2, n10 = n1t * BM;
  const int kb_beg = split * kb_per_cta;
  const int kb_end = min(kb_total, kb_beg + kb_per_cta);
  const int nkb = kb_end - kb_beg;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(tmA);
    ptx::prefetch_tmap(tmB);
    for (int i = 0; i < STAGES; i++) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], CM); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, BN2);
    ptx::tmem_relinquish();
  }
  pdl_wait();  // the prologue above overlaps the previous kernel's tail (programmatic dependent launch)
  pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if (CM > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_acc = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; i++) {
        const int kb = kb_beg + i;
        const int st = i % STAGES, u = i / STAGES;
        if (u > 0) ptx::mbar_wait(&empty[st], (u - 1) & 1);
        ptx::mbar_arrive_expect_tx(&full[st], A_BYTES + B_BYTES);
        const int bb = kb / kb_per_batch, r0 = (kb - bb * kb_per_batch) * 64;
#pragma unroll
        for (int j = 0; j < 2; j++) ptx::tma_load_4d(tmA, &full[st], smA + st * A_BYTES + j * BLK, n10 + j * 64, 0, r0, bb);
#pragma unroll
        for (int jj = 0; jj < 4 / CM; jj++) {
          const int j = rank * (4 / CM) + jj;
          const int blk = (n20 >> 6) + j;
          const int tap = blk / b_chunks_per_tap, c0 = (blk - tap * b_chunks_per_tap) * 64;
          if (CM == 1) ptx::tma_load_4d(tmB, &full[st], smB + st * B_BYTES + j * BLK, c0, tap % b_s, r0 + tap / b_s, bb);
          else ptx::tma_load_4d_mc(tmB, &full[st], smB + st * B_BYTES + j * BLK, c0, tap % b_s, r0 + tap / b_s, bb, MASK);
        }
      }
    }
  } else if (warp == 1) {
    // whole warp in the loop, one elected issuer (see gemm_nt_tc2_kernel)
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN2, 1, 1);
    // MN-major SW128: LBO = stride between 64-wide MN blocks, SBO = stride between 8-row K groups
    const uint64_t adesc0 = ptx::make_sdesc_sw128(ptx::smem_u32(smA), BLK, 1024);
    const uint64_t bdesc0 = ptx::make_sdesc_sw128(ptx::smem_u32(smB), BLK, 1024);
    const bool leader = ptx::elect_one();
    for (int i = 0; i < nkb; i++) {
      const int st = i % STAGES, u = i / STAGES;
      ptx::mbar_wait(&full[st], u & 1);
      ptx::tc_fence_after();
      if (leader) {
        const uint64_t ad = adesc0 + (uint64_t)((st * A_BYTES) >> 4), bd = bdesc0 + (uint64_t)((st * B_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < 64 / UK; k++)
          ptx::umma_bf16(tmem_acc, ad + (uint64_t)(k * ((UK * 128) >> 4)), bd + (uint64_t)(k * ((UK * 128) >> 4)), idesc,
                         (i > 0 || k > 0) ? 1u : 0u);
        if (CM == 1) ptx::umma_commit(&empty[st]);
        else ptx::umma_commit_mc(&empty[st], MASK);
      }
      __syncwarp();
    }
    if (leader) ptx::umma_commit(acc_full);
    __syncwarp();
  } else if (nkb > 0) {
    // epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves; rows go out as 64-byte runs through the staging buffer
    const int lg = warp & 3, hf = (warp - 2) >> 2;
    const int n1 = n10 + lg * 32 + lane;
    unsigned char* stg = stg_all + (warp - 2) * STG_WARP;
    const uint32_t ok = __ballot_sync(0xffffffffu, n1 < N1);
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const uint32_t trow = tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)(hf * (BN2 / 2));
    const bool vec = mode == STORE_PLAIN && (ldc & 3) == 0 && n20 + BN2 <= N2;
    uint32_t r[2][32];
    ptx::tmem_ld32(trow, r[0]);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      ptx::tmem_ld_wait();
      if (c + 1 < 4) ptx::tmem_ld32(trow + 32 * (c + 1), r[(c + 1) & 1]);
      const int c0 = hf * (BN2 / 2) + 32 * c;
      if (vec) {  // 16-byte vector reductions, 8 rows x 64 B per instruction
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[c & 1][j]);
        float* gwarp = Cacc + (long long)(n1 - lane) * ldc + n20 + c0;
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
          stage_put(stg + lane * STG_RS, v, sub, static_cast<float*>(nullptr));
          __syncwarp();
          stage_flush_red(stg, gwarp + 16 * sub, (long long)ldc, ok, lane);
          __syncwarp();
        }
      } else if (n1 < N1) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int n2 = n20 + c0 + j;
          if (n2 < N2) {
            long long o;
            if (mode == STORE_CONV_W) { const int tap = n2 / Ci, ci = n2 - tap * Ci; o = ((long long)n1 * Ci + ci) * taps + tap; }
            else o = (long long)n1 * ldc + n2;
            atomicAdd(Cacc + o, __uint_as_float(r[c & 1][j]));
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CM > 1) ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc(tmem_acc, BN2);
}

template <int BN, int STAGES> constexpr size_t nt_smem() { return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 256 + 1024; }
template <int BN, int STAGES> constexpr size_t tn_smem() { return (size_t)STAGES * (2 * 8192 + (BN / 64) * 8192) + 256 + 1024; }

}  // namespace

// hb != NULL: the weights (and the bias) are `heads` stacked (N, Kd) matrices, batch b multiplies matrix b / hb->w_div with
// A batch b % hb->a_mod (a_mod = 0: A batch b) - the per-head products of the transformer prediction heads in ONE launch.
// Second-generation kernel only; clusters never span two heads (CM drops to 1 when the tiles of a head are odd).
int gemm_nt_tc(bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C,
               cudaStream_t st, bool* handled, const HeadBatch* hb) {
  *handled = false;
  constexpr int BN = 128, STAGES = 3;
  const int cin = Kd / A.taps;
  if (N % BN != 0 || Kd % BK != 0 || cin % 64 != 0 || (A.taps > 1 && A.rs != (long long)A.s * cin)) return 0;
  if ((A.rs % 8) != 0 || (A.bs % 8) != 0 || (reinterpret_cast<uintptr_t>(A.p) & 15) || (reinterpret_cast<uintptr_t>(Bm) & 15)) return 0;
  if ((C.rs % 8) != 0 || (C.bs % 8) != 0) return 0;
  CUtensorMap tmA, tmB;
  const int heads = (hb != nullptr && hb->w_div > 0) ? nb / hb->w_div : 1;
  CPC_TRY(make_rowview_map(&tmA, A, Kd, (hb != nullptr && hb->a_mod > 0) ? hb->a_mod : nb, BM, false));
  static const int gen = []() { const char* e = getenv("CPC_B200_GEMM_GEN"); return e ? atoi(e) : 2; }();
  static const int cm_env = []() { const char* e = getenv("CPC_B200_GEMM_CM"); return e ? atoi(e) : 2; }();
  if (C.res_w > 0 && C.res_w % BN != 0) return 0;  // an output tile must not straddle two dgrad residues
  if (gen == 2 && N % 256 == 0 && (C.res_w == 0 || C.res_w % 256 == 0)) {
    int cm = (cm_env == 1 || cm_env == 2 || cm_env == 4) ? cm_env : 2;
    const int tpb2 = (A.rpb + BM - 1) / BM;
    HeadBatch HB{};
    if (hb != nullptr) {
      HB = *hb;
      HB.w_rows = N;
      while (cm > 1 && HB.w_div > 0 && (HB.w_div * tpb2) % cm != 0) cm >>= 1;  // the CTAs of a cluster share ONE weight tile
    }
    unsigned long long dims[4] = {(unsigned long long)Kd, (unsigned long long)N * heads, 1, 1};
    unsigned long long stq[3] = {(unsigned long long)Kd, (unsigned long long)Kd * N * heads, (unsigned long long)Kd * N * heads};
    unsigned box[4] = {64, (unsigned)(256 / cm), 1, 1};
    CPC_TRY(make_map4(&tmB, Bm, dims, stq, box));
    const int m_tiles = nb * tpb2, n_tiles = N / 256;
#define NT2(CMV)                                                                                                                    \
  (out_f32 ? launch_nt2<CMV, float>(tmA, tmB, Kd / BK, cin / BK, A.s, tpb2, m_tiles, n_tiles, nb, bias, C, st, CNormEpi{}, HB)       \
           : launch_nt2<CMV, bf16>(tmA, tmB, Kd / BK, cin / BK, A.s, tpb2, m_tiles, n_tiles, nb, bias, C, st, CNormEpi{}, HB))
    if (cm == 4) CPC_TRY(NT2(4)); else if (cm == 2) CPC_TRY(NT2(2)); else CPC_TRY(NT2(1));
#undef NT2
    *handled = true;
    return 0;
  }
  if (hb != nullptr) return 0;  // (first-generation kernel: one weight matrix)
  {
    unsigned long long dims[4] = {(unsigned long long)Kd, (unsigned long long)N, 1, 1};
    unsigned long long stq[3] = {(unsigned long long)Kd, (unsigned long long)Kd * N, (unsigned long long)Kd * N};
    unsigned box[4] = {64, (unsigned)BN, 1, 1};
    CPC_TRY(make_map4(&tmB, Bm, dims, stq, box));
  }
  const int tpb = (A.rpb + BM - 1) / BM;
  dim3 grid(N / BN, nb * tpb);
  const size_t smem = nt_smem<BN, STAGES>();
  if (out_f32) {
    auto k = gemm_nt_tc_kernel<BN, STAGES, float>;
    CPC_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NT_THREADS, smem, st>>>(tmA, tmB, Kd / BK, cin / BK, A.s, tpb, N, bias, C);
  } else {
    auto k = gemm_nt_tc_kernel<BN, STAGES, bf16>;
    CPC_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NT_THREADS, smem, st>>>(tmA, tmB, Kd / BK, cin / BK, A.s, tpb, N, bias, C);
  }
  CPC_LAUNCHED_N("gemm_nt_tc", st);
  *handled = true;
  return 0;
}

int debug_gemm_timeline(unsigned long long* host_out) {
  CPC_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_nt2_tl, sizeof(unsigned long long) * 148 * 8));
  return 0;
}

int gemm_nt_cnorm_tc(int nb, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C, const CNormEpi& E,
                     cudaStream_t st, bool* handled) {
  *handled = false;
  constexpr int N = 256;
  static const bool off = []() { const char* e = getenv("CPC_B200_CNORM_FUSE"); return e && atoi(e) == 0; }();
  static const int cm_env = []() { const char* e = getenv("CPC_B200_GEMM_CM"); return e ? atoi(e) : 2; }();
  const int cin = Kd / A.taps;
  if (off || Kd % BK != 0 || cin % 64 != 0 || (A.taps > 1 && A.rs != (long long)A.s * cin)) return 0;
  if ((A.rs % 8) != 0 || (A.bs % 8) != 0 || (reinterpret_cast<uintptr_t>(A.p) & 15) || (reinterpret_cast<uintptr_t>(Bm) & 15)) return 0;
  if ((C.rs % 8) != 0 || (C.bs % 8) != 0 || C.res_w != 0 || C.relu != 0 || C.t_lo != 0 || C.t_hi != C.rpb) return 0;
  if (reinterpret_cast<uintptr_t>(C.p) & 15) return 0;
  if (E.y != nullptr && (reinterpret_cast<uintptr_t>(E.y) & 15)) return 0;
  CUtensorMap tmA, tmB;
  CPC_TRY(make_rowview_map(&tmA, A, Kd, nb, BM, false));
  const int cm = cm_env == 1 ? 1 : 2;
  unsigned long long dims[4] = {(unsigned long long)Kd, (unsigned long long)N, 1, 1};
  unsigned long long stq[3] = {(unsigned long long)Kd, (unsigned long long)Kd * N, (unsigned long long)Kd * N};
  unsigned box[4] = {64, (unsigned)(256 / cm), 1, 1};
  CPC_TRY(make_map4(&tmB, Bm, dims, stq, box));
  const int tpb2 = (A.rpb + BM - 1) / BM;
  const int m_tiles = nb * tpb2;
  if (cm == 2) CPC_TRY((launch_nt2<2, bf16, true>(tmA, tmB, Kd / BK, cin / BK, A.s, tpb2, m_tiles, 1, nb, bias, C, st, E)));
  else CPC_TRY((launch_nt2<1, bf16, true>(tmA, tmB, Kd / BK, cin / BK, A.s, tpb2, m_tiles, 1, nb, bias, C, st, E)));
  *handled = true;
  return 0;
}

// problems that fit the gen-2 TN kernel (any other shape goes through gemm_tn one by one)
static bool tn2_fits(const TnDesc& d) {
  const RowView &A = d.A, &B = d.B;
  if (A.rpb != B.rpb || A.taps != 1) return false;
  const int bcin = d.N2 / B.taps;
  if (d.N1 % 256 != 0 || d.N2 % 256 != 0 || bcin % 64 != 0 || (B.taps > 1 && B.rs != (long long)B.s * bcin)) return false;
  if ((A.rs % 8) != 0 || (A.bs % 8) != 0 || (B.rs % 8) != 0 || (B.bs % 8) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(A.p) & 15) || (reinterpret_cast<uintptr_t>(B.p) & 15)) return false;
  return true;
}

int gemm_tn_group_tc(int n, const TnDesc* d, cudaStream_t st, bool* handled) {
  *handled = false;
  static const int gen = []() { const char* e = getenv("CPC_B200_GEMM_GEN"); return e ? atoi(e) : 2; }();
  if (gen != 2 || n < 1 || n > 4) return 0;
  for (int i = 0; i < n; i++)
    if (!tn2_fits(d[i])) return 0;
  TnGroup G;
  memset(&G, 0, sizeof(G));
  G.n = n;
  long long work = 0;
  int tiles[4], kbt[4];
  for (int i = 0; i < n; i++) {
    CPC_TRY(make_rowview_map(&G.tmA[i], d[i].A, d[i].N1, d[i].nb, 64, true));   // exact row count: rows >= rpb are zero-filled by TMA
    CPC_TRY(make_rowview_map(&G.tmB[i], d[i].B, d[i].N2, d[i].nb, 64, false));
    tiles[i] = (d[i].N1 / BM) * (d[i].N2 / 256);
    kbt[i] = ((d[i].A.rpb + 63) / 64) * d[i].nb;
    work += (long long)tiles[i] * kbt[i];
  }
  // smallest per-CTA k-block count L for which all problems fit one wave of 148 CTAs
  int L = (int)((work + 147) / 148);
  if (L < 1) L = 1;
  for (;; L++) {
    int ctas = 0;
    for (int i = 0; i < n; i++) ctas += tiles[i] * ((kbt[i] + L - 1) / L);
    if (ctas <= 148) break;
    bool all_one = true;
    for (int i = 0; i < n; i++) all_one = all_one && (kbt[i] <= L);
    if (all_one) break;  // more than 148 tiles in total: several waves, no split-K
  }
  int cta = 0;
  for (int i = 0; i < n; i++) {
    TnProblem& p = G.pr[i];
    int sp = (kbt[i] + L - 1) / L;
    const int kpc = (kbt[i] + sp - 1) / sp;
    sp = (kbt[i] + kpc - 1) / kpc;
    p.cta_begin = cta;
    p.n2_tiles = d[i].N2 / 256;
    p.kb_total = kbt[i];
    p.kb_per_cta = kpc;
    p.kb_per_batch = (d[i].A.rpb + 63) / 64;
    p.b_chunks_per_tap = (d[i].N2 / d[i].B.taps) / 64;
    p.b_s = d[i].B.s;
    p.N1 = d[i].N1; p.N2 = d[i].N2; p.ldc = d[i].ldc; p.mode = d[i].mode; p.Ci = d[i].Ci; p.taps = d[i].taps;
    p.Cacc = d[i].Cacc;
    cta += tiles[i] * sp;
  }
  auto k2 = gemm_tn_tc2_kernel<2>;
  const size_t smem2 = (size_t)4 * (2 * 8192 + 4 * 8192) + 256 + 8 * STG_WARP + 1024;
  CPC_CHECK_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cta);
  cfg.blockDim = dim3(NT2_THREADS);
  cfg.dynamicSmemBytes = smem2;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CPC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k2, G));
  CPC_LAUNCHED_N("gemm_tn_tc2", st);
  *handled = true;
  return 0;
}

int gemm_tn_tc(int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc, int mode, int Ci, int taps,
               cudaStream_t st, bool* handled) {
  *handled = false;
  {
    TnDesc one{nb, N1, N2, A, B, Cacc, ldc, mode, Ci, taps};
    CPC_TRY(gemm_tn_group_tc(1, &one, st, handled));
    if (*handled) return 0;
  }
  constexpr int BN = 128, STAGES = 3;
  if (A.rpb != B.rpb || A.taps != 1) return 0;
  const int bcin = N2 / B.taps;
  if (N1 % BM != 0 || N2 % BN != 0 || bcin % 64 != 0 || (B.taps > 1 && B.rs != (long long)B.s * bcin)) return 0;
  if ((A.rs % 8) != 0 || (A.bs % 8) != 0 || (B.rs % 8) != 0 || (B.bs % 8) != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(A.p) & 15) || (reinterpret_cast<uintptr_t>(B.p) & 15)) return 0;
  CUtensorMap tmA, tmB;
  CPC_TRY(make_rowview_map(&tmA, A, N1, nb, 64, true));   // exact row count: rows >= rpb are zero-filled by TMA
  CPC_TRY(make_rowview_map(&tmB, B, N2, nb, 64, false));
  const int kb_per_batch = (A.rpb + 63) / 64;
  const int kb_total = kb_per_batch * nb;
  const int tiles = (N1 / BM) * (N2 / BN);
  int splits = (148 * 2 + tiles - 1) / tiles;
  if (splits > kb_total) splits = kb_total;
  if (splits < 1) splits = 1;
  const int kb_per_cta = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per_cta - 1) / kb_per_cta;
  dim3 grid(N2 / BN, N1 / BM, splits);
  const size_t smem = tn_smem<BN, STAGES>();
  auto k = gemm_tn_tc_kernel<BN, STAGES>;
  CPC_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, NT_THREADS, smem, st>>>(tmA, tmB, kb_total, kb_per_cta, kb_per_batch, bcin / 64, B.s, N1, N2, Cacc, ldc, mode, Ci, taps);
  CPC_LAUNCHED_N("gemm_tn_tc", st);
  *handled = true;
  return 0;
}

}  // namespace cpcb200
