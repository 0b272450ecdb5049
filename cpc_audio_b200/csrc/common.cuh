// common.cuh - shared host/device helpers for libcpc_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/cpc_b200.h"

namespace cpcb200 {

// ---- error reporting (thread-local message, negative status) ---------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

int fail(int code, const char* fmt, ...);

#define CPC_CHECK_CUDA(expr)                                                                         \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::cpcb200::fail(CPCB200_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,          \
                             cudaGetErrorString(_e));                                                \
  } while (0)

bool prof_enabled();
void prof_mark(cudaStream_t st);
void prof_note(const char* name, cudaStream_t st);

// call after every kernel launch: counts it and surfaces launch-configuration errors without syncing
#define CPC_LAUNCHED_N(name, st)                                                                             \
  do {                                                                                               \
    ::cpcb200::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
    ::cpcb200::prof_note(name, st);                                                                  \
    cudaError_t _e = cudaPeekAtLastError();                                                          \
    if (_e != cudaSuccess)                                                                           \
      return ::cpcb200::fail(CPCB200_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,      \
                             cudaGetErrorString(_e));                                                \
  } while (0)

#define CPC_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != 0) return _s;      \
  } while (0)

// ---- geometry of the encoder (cpc/model.py:83-92) --------------------------------------------------------
constexpr int kConvK[5] = {10, 8, 4, 4, 4};
constexpr int kConvS[5] = {5, 4, 2, 2, 2};
constexpr int kConvP[5] = {3, 2, 1, 1, 1};
constexpr int kPad = 2;  // zero rows stored before and after every window of a padded activation

struct Geo {
  int B, L, H, Har, K, N, nL, S, W;
  int dff = 0, nheads = 0;  // > 0: transformer prediction heads (rnnMode='transformer')
  int Lout[5];  // output length of conv i
  bool bf16;
};

int make_geo(const cpcb200_dims* d, Geo* g);

__host__ __device__ static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided buffer
struct Carver {
  char* base;
  size_t off = 0, cap;
  Carver(void* p, size_t c) : base(static_cast<char*>(p)), cap(c) {}
  template <class T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T));
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  void* take_bytes(size_t n) { return take<char>(n); }
  bool ok() const { return off <= cap; }
};

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// Every kernel of the library calls pdl_wait() before its first access to global memory - it returns once the previous
// kernel in the stream has completed and flushed - and pdl_trigger() right after it: the NEXT kernel may then be scheduled
// while this one runs, so its launch latency and prologue (barrier init, TMEM allocation, descriptor prefetch) overlap
// this kernel instead of following it.  Wait-then-trigger keeps at most two kernels in flight, so code in front of the
// wait may only race with the immediately preceding kernel.  Kernels launched through launch_k() carry the
// stream-serialization attribute that arms this; with CPC_B200_PDL=0 (or a plain <<<>>> launch) both instructions are
// no-ops and the stream order is the classic one.
bool pdl_enabled();

// one-shot hook (cpcb200_encoder_bwd_set_event): the next encoder backward ON STREAM `st` records this event once every
// parameter gradient except conv0's / batchNorm0's is final; returns nullptr when none is armed for that stream
cudaEvent_t take_grads_ready_event(cudaStream_t st);

// SMs the persistent GEMMs launched by this thread must leave free (0 by default).  The encoder backward sets it around the
// last data-gradient GEMM when an early gradient exchange has been armed: that exchange (cpcb200_peer_reduce_range) runs
// beside the GEMM on its own SMs - a GEMM CTA (221 KB of shared memory) never shares an SM with it.
int sm_reserve();
void set_sm_reserve(int n);
constexpr int kEarlyExchangeSMs = 4;  // 144 SMs = 72 clusters of 2 take the 1152 tile pairs of dgrad_1 in exactly 16 rounds

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// launch `kern` on `st`; cluster_x > 1 adds a cluster dimension
template <class... KA, class... A>
inline cudaError_t launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    n++;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
#endif

// ---- device helpers ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// load / store NV consecutive elements as fp32 (vectorised where the type allows)
template <int NV> __device__ __forceinline__ void load_vec(const float* p, float (&v)[NV]) {
  static_assert(NV % 4 == 0, "");
#pragma unroll
  for (int i = 0; i < NV / 4; i++) {
    float4 t = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
template <int NV> __device__ __forceinline__ void load_vec(const bf16* p, float (&v)[NV]) {
  static_assert(NV % 4 == 0, "");
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV / 8; i++) {
      uint4 t = reinterpret_cast<const uint4*>(p)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 4; j++) { float2 f = __bfloat1622float2(h[j]); v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV / 4; i++) {
      uint2 t = reinterpret_cast<const uint2*>(p)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 2; j++) { float2 f = __bfloat1622float2(h[j]); v[4 * i + 2 * j] = f.x; v[4 * i + 2 * j + 1] = f.y; }
    }
  }
}
template <int NV> __device__ __forceinline__ void store_vec(float* p, const float (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV / 4; i++) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <int NV> __device__ __forceinline__ void store_vec(bf16* p, const float (&v)[NV]) {
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV / 8; i++) {
      uint4 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 4; j++) h[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
      reinterpret_cast<uint4*>(p)[i] = t;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV / 4; i++) {
      uint2 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 2; j++) h[j] = __floats2bfloat162_rn(v[4 * i + 2 * j], v[4 * i + 2 * j + 1]);
      reinterpret_cast<uint2*>(p)[i] = t;
    }
  }
}
#endif  // __CUDACC__

// ---- GEMM building blocks (gemm_simt.cu / gemm_tc.cu) -----------------------------------------------------
// Row view: logical row m -> (b = m / rpb, t = m % rpb) -> element offset  b*bs + t*rs ; inner dim contiguous.
// Rows may overlap (rs < row length): that is how a strided Conv1d over a channel-last activation becomes a
// plain GEMM without im2col (DESIGN.md "conv as a strided-row view").
struct RowView {
  const void* p;
  long long bs;  // elements between batches
  long long rs;  // elements between consecutive rows
  int rpb;       // rows per batch
  // conv-style rows: the inner dim is `taps` consecutive source rows of (inner/taps) channels and consecutive
  // logical rows start `s` source rows apart (rs == s * inner/taps).  Plain matrices: taps = s = 1.
  int taps = 1;
  int s = 1;
};
struct OutView {
  void* p;
  long long bs, rs;
  int rpb;
  int t_lo, t_hi;  // only rows with t_lo <= t < t_hi are stored
  // merged transposed-conv output (dgrad): the N columns are `res_s` residues of res_w channels each; column block
  // r = n / res_w of logical row t is input row j = res_s*t + r - res_p, which exists iff 0 <= j < res_rows (the input
  // length of the layer: res_s*(rpb-1) when the window length is a multiple of 160, up to res_s-1 rows more otherwise).
  // res_w == 0: plain output.
  int res_w = 0;
  int res_p = 0;
  int res_s = 0;
  int res_rows = 0;
  int relu = 0;  // apply max(., 0) after the bias (FFN of the transformer prediction heads)
};
__host__ __device__ inline bool out_row_ok(const OutView& C, int t, int n0) {
  if (t >= C.rpb || t < C.t_lo || t >= C.t_hi) return false;
  if (C.res_w > 0) {
    const int j = C.res_s * t + n0 / C.res_w - C.res_p;
    if (j < 0 || j >= C.res_rows) return false;
  }
  return true;
}
enum StoreMode { STORE_PLAIN = 0, STORE_CONV_W = 1 };

// ChannelNorm + ReLU fused into the GEMM epilogue (N == 256: one output tile holds every channel of a frame, so the
// statistics of model.py:52-54 are a per-thread reduction over the accumulator row).  The pre-norm row u goes to the
// OutView (saved for backward), the normalised row to y (activation dtype, same row geometry as the OutView, its
// kPad zero rows around each window written here) or, for the last layer, to z (fp32, (nb, rpb, 256) dense).
struct CNormEpi {
  const float* gam;
  const float* bet;
  void* y;
  float* z;
  int pad_rows;
  float2* stats = nullptr;  // (mean, rstd) of every row, index b*rpb + t: saved for the ChannelNorm backward
  int save_u = 1;           // 0: inference - the pre-norm rows are not written (the OutView only supplies the row geometry)
};

// C[m,n] = sum_k A[m,k] * B[n,k] (+ bias[n]).  A row view (M = nb*rpb rows, inner Kd); B dense (N, Kd) ld=Kd.
// out_f32: C is fp32, else C has the activation dtype.
int gemm_nt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias,
            const OutView& C, cudaStream_t st);
// Stacked per-head weights (bf16 tensor-core path only; anything else is CPCB200_ERR_UNSUPPORTED): Bm / bias hold nb / w_div
// matrices of (N, Kd) / vectors of N, batch b multiplies matrix b / w_div with A batch b % a_mod (a_mod = 0: A batch b).
struct HeadBatch { int a_mod = 0, w_div = 0, w_rows = 0; };
int gemm_nt_heads(bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C,
                  int a_mod, int w_div, cudaStream_t st);
// bf16 tensor-core path only: u = A.B^T + bias -> C (bf16), ChannelNorm+ReLU(u) -> E.y / E.z.  *handled = false
// when the shape does not fit (N != 256, ...): the caller then runs gemm_nt + the stand-alone norm kernel.
int gemm_nt_cnorm_tc(int nb, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C, const CNormEpi& E,
                     cudaStream_t st, bool* handled);
// Cacc[n1,n2] += sum_m A[m,n1] * B[m,n2]  (fp32 atomics).  mode STORE_CONV_W: n2 = tap*Ci + ci -> Cacc[(n1*Ci + ci)*taps + tap]
int gemm_tn(bool bf16_in, int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc,
            int mode, int Ci, int taps, cudaStream_t st);

// several independent TN products in one launch (bf16 tensor-core path; falls back to one gemm_tn per problem)
struct TnDesc { int nb, N1, N2; RowView A, B; float* Cacc; int ldc, mode, Ci, taps; };
int gemm_tn_group(bool bf16_in, int n, const TnDesc* d, cudaStream_t st);

}  // namespace cpcb200
