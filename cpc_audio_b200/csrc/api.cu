// api.cu - the extern "C" surface declared in include/cpc_b200.h, argument validation, GEMM dispatch.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cstdlib>
#include <cooperative_groups.h>
#include "common.cuh"

namespace cpcb200 {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

// ---- optional per-kernel timing (bench.py's roofline leg): one CUDA event after every launch, on the
// launching stream; the time attributed to a kernel is the gap to the previous event of the same API call.
namespace {
struct ProfRec { const char* name; int prev, cur; };
std::mutex g_prof_mu;
std::atomic<int> g_prof_on{0};
std::vector<cudaEvent_t> g_prof_ev;
std::vector<ProfRec> g_prof_rec;
int g_prof_used = 0, g_prof_last = -1;
int prof_next_event(cudaStream_t st) {
  if (g_prof_used == (int)g_prof_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    g_prof_ev.push_back(e);
  }
  int i = g_prof_used++;
  cudaEventRecord(g_prof_ev[i], st);
  return i;
}
}  // namespace
void prof_mark(cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_last = prof_next_event(st);
}
void prof_note(const char* name, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int e = prof_next_event(st);
  if (e >= 0 && g_prof_last >= 0) g_prof_rec.push_back({name, g_prof_last, e});
  g_prof_last = e;
}

static std::atomic<void*> g_grads_ready_event{nullptr};
cudaEvent_t take_grads_ready_event() { return static_cast<cudaEvent_t>(g_grads_ready_event.exchange(nullptr)); }

bool pdl_enabled() {
  static const bool on = []() { const char* e = getenv("CPC_B200_PDL"); return !(e && atoi(e) == 0); }();
  return on;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int make_geo(const cpcb200_dims* d, Geo* g) {
  if (!d) return fail(CPCB200_ERR_NULL, "dims is NULL");
  if (d->B <= 0 || d->L <= 0 || d->L % 160 != 0) return fail(CPCB200_ERR_BAD_DIMS, "B=%d L=%d (L must be a positive multiple of 160)", d->B, d->L);
  if (d->H % 64 != 0 || d->H <= 0 || d->H > 512) return fail(CPCB200_ERR_BAD_DIMS, "H=%d must be a multiple of 64 in [64,512]", d->H);
  if (d->Har % 64 != 0 || d->Har <= 0 || d->Har > 512) return fail(CPCB200_ERR_BAD_DIMS, "Har=%d must be a multiple of 64 in [64,512]", d->Har);
  if (d->nLayers < 1 || d->nLayers > CPCB200_MAX_GRU_LAYERS) return fail(CPCB200_ERR_BAD_DIMS, "nLayers=%d", d->nLayers);
  if (d->dtype != CPCB200_F32 && d->dtype != CPCB200_BF16) return fail(CPCB200_ERR_UNSUPPORTED, "dtype=%d", d->dtype);
  g->B = d->B; g->L = d->L; g->H = d->H; g->Har = d->Har; g->K = d->K; g->N = d->N; g->nL = d->nLayers;
  g->S = d->L / 160;
  g->W = g->S - d->K;
  g->bf16 = d->dtype == CPCB200_BF16;
  int L = d->L;
  for (int i = 0; i < 5; i++) { L = (L + 2 * kConvP[i] - kConvK[i]) / kConvS[i] + 1; g->Lout[i] = L; }
  if (g->Lout[4] != g->S) return fail(CPCB200_ERR_BAD_DIMS, "internal: conv stack length %d != S %d", g->Lout[4], g->S);
  return 0;
}

static int check_crit(const Geo& g) {
  if (g.K < 1 || g.W < 1) return fail(CPCB200_ERR_BAD_DIMS, "K=%d leaves no anchor positions (S=%d)", g.K, g.S);
  if (g.N < 1) return fail(CPCB200_ERR_BAD_DIMS, "N=%d", g.N);
  return 0;
}

// implemented in the per-op translation units
size_t encoder_save_elems(const Geo& g);
size_t encoder_ws_bytes(const Geo& g, int backward);
int encoder_fwd(const Geo&, const float*, const cpcb200_encoder_params*, float*, void*, void*, size_t, cudaStream_t);
int encoder_bwd(const Geo&, const float*, const cpcb200_encoder_params*, const float*, const void*,
                const cpcb200_encoder_params*, void*, size_t, cudaStream_t);
size_t gru_save_bytes(const Geo& g);
size_t gru_ws_bytes(const Geo& g, int backward);
int gru_fwd(const Geo&, const float*, const float*, const cpcb200_gru_params*, float*, float*, void*, void*, size_t, cudaStream_t);
int gru_bwd(const Geo&, const float*, const float*, const cpcb200_gru_params*, const float*, const float*, const void*, float*,
            const cpcb200_gru_params*, void*, size_t, cudaStream_t);
int sample_ext_idx(const Geo&, const int64_t*, const int64_t*, int32_t*, cudaStream_t);
size_t criterion_save_bytes(const Geo& g);
size_t criterion_ws_bytes(const Geo& g, int backward);
int criterion_fwd(const Geo&, const float*, const float*, const float*, const cpcb200_thead_params*, const int*, float*, float*,
                  void*, void*, size_t, cudaStream_t);
int criterion_bwd(const Geo&, const float*, const float*, const float*, const cpcb200_thead_params*, const int*, const float*,
                  const void*, float*, float*, float*, const cpcb200_thead_params*, void*, size_t, cudaStream_t);

int gemm_nt_simt(bool, bool, int, int, int, const RowView&, const void*, const float*, const OutView&, cudaStream_t);
int gemm_tn_simt(bool, int, int, int, const RowView&, const RowView&, float*, int, int, int, int, cudaStream_t);
int debug_gemm_timeline(unsigned long long* host_out);
int gemm_nt_tc(bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C,
               cudaStream_t st, bool* handled);
int gemm_tn_tc(int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc, int mode, int Ci, int taps,
               cudaStream_t st, bool* handled);

// Dispatch: the bf16 path takes the tcgen05 kernels whenever the shape fits their tiling, else CUDA cores.
int gemm_nt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias,
            const OutView& C, cudaStream_t st) {
  if (bf16_in) {
    bool handled = false;
    CPC_TRY(gemm_nt_tc(out_f32, nb, N, Kd, A, Bm, bias, C, st, &handled));
    if (handled) return 0;
  }
  return gemm_nt_simt(bf16_in, out_f32, nb, N, Kd, A, Bm, bias, C, st);
}
int gemm_tn(bool bf16_in, int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc, int mode,
            int Ci, int taps, cudaStream_t st) {
  if (bf16_in) {
    bool handled = false;
    CPC_TRY(gemm_tn_tc(nb, N1, N2, A, B, Cacc, ldc, mode, Ci, taps, st, &handled));
    if (handled) return 0;
  }
  return gemm_tn_simt(bf16_in, nb, N1, N2, A, B, Cacc, ldc, mode, Ci, taps, st);
}

int gemm_tn_group_tc(int n, const TnDesc* d, cudaStream_t st, bool* handled);
int gemm_tn_group(bool bf16_in, int n, const TnDesc* d, cudaStream_t st) {
  static const bool off = []() { const char* e = getenv("CPC_B200_TN_GROUP"); return e && atoi(e) == 0; }();
  if (bf16_in && n > 1 && !off) {
    bool handled = false;
    CPC_TRY(gemm_tn_group_tc(n, d, st, &handled));
    if (handled) return 0;
  }
  for (int i = 0; i < n; i++)
    CPC_TRY(gemm_tn(bf16_in, d[i].nb, d[i].N1, d[i].N2, d[i].A, d[i].B, d[i].Cacc, d[i].ldc, d[i].mode, d[i].Ci, d[i].taps, st));
  return 0;
}

// torch.optim.Adam (non-amsgrad) over a flat bucket: cpc/train.py:335-337
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  pdl_wait();
  pdl_trigger();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// graph-capturable variant: step count on the device (state[0]) together with beta1^step and beta2^step as running
// float64 products (state[2..5]; double-precision pow() in the kernel costs ~90 us on this part), the last block to
// finish (ticket in state[1]) publishes the next step's values - every block has read the current ones by then
template <bool ZERO>
__global__ void adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                size_t n, float lr, float b1, float b2, float eps, float wd, int* __restrict__ state) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_bc[2];
  double* pw = reinterpret_cast<double*>(state + 2);  // beta1^(steps+1), beta2^(steps+1) once steps > 0
  double p1 = 0.0, p2 = 0.0;
  if (threadIdx.x == 0) {
    const int done = *reinterpret_cast<volatile int*>(state);
    p1 = done == 0 ? (double)b1 : *reinterpret_cast<volatile double*>(pw);
    p2 = done == 0 ? (double)b2 : *reinterpret_cast<volatile double*>(pw + 1);
    s_bc[0] = (float)(1.0 - p1);
    s_bc[1] = sqrtf((float)(1.0 - p2));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  const float step_size = lr / bc1;
  auto upd = [&](float gi, float pi, float& mi, float& vi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = fmaf(b1, mi, (1.f - b1) * gi);
    vi = fmaf(b2, vi, (1.f - b2) * gi * gi);
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    return pi - step_size * (mi / denom);
  };
  // all loads of a quad first, every store (the cleared gradient included) after the arithmetic
  const size_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 gq = g4[i];
    float4 pq = p4[i], mq = m4[i], vq = v4[i];
    pq.x = upd(gq.x, pq.x, mq.x, vq.x); pq.y = upd(gq.y, pq.y, mq.y, vq.y);
    pq.z = upd(gq.z, pq.z, mq.z, vq.z); pq.w = upd(gq.w, pq.w, mq.w, vq.w);
    m4[i] = mq; v4[i] = vq; p4[i] = pq;
    if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float mi = m[i], vi = v[i];
    const float pn = upd(g[i], p[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pn;
    if (ZERO) g[i] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(state + 1, 1);
    if (ticket == (int)gridDim.x - 1) {
      state[1] = 0;
      pw[0] = p1 * (double)b1;
      pw[1] = p2 * (double)b2;
      __threadfence();
      atomicAdd(state, 1);
    }
  }
}

// ---- all-reduce + Adam + zero_grad over peer memory (cpcb200_allreduce_adam_step) ------------------------------------
struct PeerPtrs { float* g[8]; unsigned* sig[8]; int rank, world; float* mc; };
// NVSwitch in-fabric reduction: the sum over every GPU's copy of the 16 bytes at this multicast address / broadcast store
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Slice data is moved with ordinary L2-level accesses (ld.global.cg / st.global.cg): the node barriers (release / acquire
// at system scope) order them, and sys-scoped data accesses measured ~4x slower than the link on this path.
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_relaxed_sys_v4(float* p, const float4& v) { __stcg(reinterpret_cast<float4*>(p), v); }
// node-wide barrier on signal words.  arrive: one thread per peer (block 0) tells that peer this rank reached `epoch`;
// wait: the first `world` threads of EVERY block poll this rank's own words (local memory, written by the peers) - no
// grid-wide sync is needed to release the other blocks.  Epochs only grow, so a peer that is already one barrier ahead
// still satisfies the wait.
__device__ __forceinline__ void node_arrive(const PeerPtrs& P, unsigned epoch) {
  if (blockIdx.x == 0 && (int)threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys(P.sig[threadIdx.x] + P.rank, epoch);
  }
}
__device__ __forceinline__ void node_wait(const PeerPtrs& P, unsigned epoch) {
  if ((int)threadIdx.x < P.world) {
    const unsigned* mine = P.sig[P.rank] + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > (1ll << 36)) __trap();  // ~35 s (ranks may start seconds apart while 8 processes load their
                                                   // modules): a rank that never arrives faults instead of hanging the GPU
    }
  }
  __syncthreads();
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool ZERO>
__global__ void __launch_bounds__(256) allreduce_adam_kernel(PeerPtrs P, float* __restrict__ p, float* __restrict__ m,
                                                             float* __restrict__ v, size_t n, float lr, float b1, float b2,
                                                             float eps, float wd, int* __restrict__ state) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ float s_bc[2];
  double* pw = reinterpret_cast<double*>(state + 2);
  double p1 = 0.0, p2 = 0.0;
  unsigned calls = 0;
  if (threadIdx.x == 0) {
    const int done = *reinterpret_cast<volatile int*>(state);
    p1 = done == 0 ? (double)b1 : *reinterpret_cast<volatile double*>(pw);
    p2 = done == 0 ? (double)b2 : *reinterpret_cast<volatile double*>(pw + 1);
    s_bc[0] = (float)(1.0 - p1);
    s_bc[1] = sqrtf((float)(1.0 - p2));
  }
  calls = (unsigned)*reinterpret_cast<volatile int*>(state + 6);  // node-barrier epoch base: 2 barriers per call
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  const float step_size = lr / bc1;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  float* g = P.g[P.rank];

  // phase stamps (ns since kernel start) in this rank's signal words 40..44: read by tools/peer_adam_check.py
  const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t_start = stamp ? globaltimer_ns() : 0ull;
  unsigned* dbg = P.sig[P.rank] + 40;
  // ---- every rank's gradients are final ----
  node_arrive(P, 2 * calls + 1);
  node_wait(P, 2 * calls + 1);
  if (stamp) dbg[0] = (unsigned)(globaltimer_ns() - t_start);
  // ---- two-shot all-reduce, in place: this rank owns slice `rank` of every buffer ----
  const size_t n4 = n / 4;
  const size_t per = (n4 + P.world - 1) / P.world;
  const size_t lo = per * P.rank, hi = lo + per < n4 ? lo + per : n4;
  // (peer loads cost ~2 us each: all the loads of U quads are issued before the first add)
  constexpr int U = 8;
  if (P.mc != nullptr) {
    for (size_t i0 = lo + tid; i0 < hi; i0 += U * nthr) {
      float4 acc[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + (size_t)u * nthr;
        if (i < hi) acc[u] = multimem_ld_reduce_v4(P.mc + 4 * i);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + (size_t)u * nthr;
        if (i < hi) multimem_st_v4(P.mc + 4 * i, acc[u]);
      }
    }
  } else
  for (size_t i0 = lo + tid; i0 < hi; i0 += U * nthr) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + (size_t)u * nthr;
      acc[u] = i < hi ? ld_relaxed_sys_v4(P.g[0] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int j = 1; j < P.world; j++) {
      float4 t[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + (size_t)u * nthr;
        t[u] = i < hi ? ld_relaxed_sys_v4(P.g[j] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; u++) { acc[u].x += t[u].x; acc[u].y += t[u].y; acc[u].z += t[u].z; acc[u].w += t[u].w; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + (size_t)u * nthr;
      if (i < hi)
        for (int j = 0; j < P.world; j++) st_relaxed_sys_v4(P.g[j] + 4 * i, acc[u]);
    }
  }
  if (P.rank == P.world - 1) {  // the < 4 floats past the last quad
    for (size_t i = n4 * 4 + tid; i < n; i += nthr) {
      float acc = 0.f;
      for (int j = 0; j < P.world; j++) acc += *reinterpret_cast<volatile float*>(P.g[j] + i);
      for (int j = 0; j < P.world; j++) *reinterpret_cast<volatile float*>(P.g[j] + i) = acc;
    }
  }
  __threadfence_system();  // this thread's peer stores are performed system-wide before it reports in
  if (stamp) dbg[1] = (unsigned)(globaltimer_ns() - t_start);
  grid.sync();  // every block of this rank is past its fence: block 0 may tell the peers
  if (stamp) dbg[2] = (unsigned)(globaltimer_ns() - t_start);
  // ---- every slice of the local buffer has been written by its owner ----
  node_arrive(P, 2 * calls + 2);
  node_wait(P, 2 * calls + 2);
  if (stamp) dbg[3] = (unsigned)(globaltimer_ns() - t_start);
  // ---- Adam on the full local replica, gradients read past the L1 (peers wrote them) ----
  auto upd = [&](float gi, float pi, float& mi, float& vi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = fmaf(b1, mi, (1.f - b1) * gi);
    vi = fmaf(b2, vi, (1.f - b2) * gi * gi);
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    return pi - step_size * (mi / denom);
  };
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (size_t i = tid; i < n4; i += nthr) {
    const float4 gq = __ldcg(g4 + i);
    float4 pq = p4[i], mq = m4[i], vq = v4[i];
    pq.x = upd(gq.x, pq.x, mq.x, vq.x); pq.y = upd(gq.y, pq.y, mq.y, vq.y);
    pq.z = upd(gq.z, pq.z, mq.z, vq.z); pq.w = upd(gq.w, pq.w, mq.w, vq.w);
    m4[i] = mq; v4[i] = vq; p4[i] = pq;
    if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (size_t i = n4 * 4 + tid; i < n; i += nthr) {
    float mi = m[i], vi = v[i];
    const float pn = upd(__ldcg(g + i), p[i], mi, vi);
    m[i] = mi; v[i] = vi; p[i] = pn;
    if (ZERO) g[i] = 0.f;
  }
  if (stamp) dbg[4] = (unsigned)(globaltimer_ns() - t_start);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // every block read the state before the grid.sync
    pw[0] = p1 * (double)b1;
    pw[1] = p2 * (double)b2;
    state[6] = (int)(calls + 1);
    __threadfence();
    atomicAdd(state, 1);
  }
}

}  // namespace cpcb200

using namespace cpcb200;

#define GEO_OR_RETURN(d, g) \
  Geo g;                    \
  CPC_TRY(make_geo(d, &g))
#define NOT_NULL(p) \
  if (!(p)) return fail(CPCB200_ERR_NULL, #p " is NULL")

extern "C" {

int cpcb200_version(void) { return CPCB200_VERSION; }
const char* cpcb200_last_error(void) { return g_err; }
uint64_t cpcb200_launch_count(void) { return g_launches.load(); }

int cpcb200_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_rec.clear();
  g_prof_used = 0;
  g_prof_last = -1;
  g_prof_on.store(on ? 1 : 0);
  return 0;
}
int cpcb200_prof_report(char* buf, size_t cap) {
  NOT_NULL(buf);
  CPC_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<int, double>> agg;
  for (const ProfRec& r : g_prof_rec) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[r.prev], g_prof_ev[r.cur]) != cudaSuccess) continue;
    auto& a = agg[r.name];
    a.first += 1;
    a.second += ms;
  }
  std::string out;
  char line[160];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (out.size() + 1 > cap) return fail(CPCB200_ERR_WORKSPACE, "prof_report: buffer too small (%zu needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

size_t cpcb200_encoder_save_bytes(const cpcb200_dims* d) {
  Geo g;
  if (make_geo(d, &g)) return 0;
  return encoder_save_elems(g) * (g.bf16 ? 2 : 4) + 256;
}
size_t cpcb200_encoder_ws_bytes(const cpcb200_dims* d, int backward) {
  Geo g;
  if (make_geo(d, &g)) return 0;
  return encoder_ws_bytes(g, backward);
}
int cpcb200_encoder_fwd(const cpcb200_dims* d, const float* x, const cpcb200_encoder_params* p, float* z, void* save,
                        void* ws, size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  NOT_NULL(x); NOT_NULL(p); NOT_NULL(z); NOT_NULL(save); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return encoder_fwd(g, x, p, z, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cpcb200_encoder_bwd(const cpcb200_dims* d, const float* x, const cpcb200_encoder_params* p, const float* dz,
                        const void* save, const cpcb200_encoder_params* grads, void* ws, size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  NOT_NULL(x); NOT_NULL(p); NOT_NULL(dz); NOT_NULL(save); NOT_NULL(grads); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return encoder_bwd(g, x, p, dz, save, grads, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

size_t cpcb200_gru_save_bytes(const cpcb200_dims* d) {
  Geo g;
  if (make_geo(d, &g)) return 0;
  return gru_save_bytes(g);
}
size_t cpcb200_gru_ws_bytes(const cpcb200_dims* d, int backward) {
  Geo g;
  if (make_geo(d, &g)) return 0;
  return gru_ws_bytes(g, backward);
}
int cpcb200_gru_fwd(const cpcb200_dims* d, const float* z, const float* h0, const cpcb200_gru_params* p, float* c, float* hT,
                    void* save, void* ws, size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  NOT_NULL(z); NOT_NULL(p); NOT_NULL(c); NOT_NULL(save); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return gru_fwd(g, z, h0, p, c, hT, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cpcb200_gru_bwd(const cpcb200_dims* d, const float* z, const float* h0, const cpcb200_gru_params* p, const float* c,
                    const float* dc, const void* save, float* dz, const cpcb200_gru_params* grads, void* ws, size_t ws_bytes,
                    void* stream) {
  GEO_OR_RETURN(d, g);
  NOT_NULL(z); NOT_NULL(p); NOT_NULL(c); NOT_NULL(dc); NOT_NULL(save); NOT_NULL(dz); NOT_NULL(grads); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return gru_bwd(g, z, h0, p, c, dc, save, dz, grads, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int cpcb200_sample_ext_idx(const cpcb200_dims* d, const int64_t* batch_idx, const int64_t* seq_idx, int32_t* ext, void* stream) {
  GEO_OR_RETURN(d, g);
  CPC_TRY(check_crit(g));
  NOT_NULL(batch_idx); NOT_NULL(seq_idx); NOT_NULL(ext);
  return sample_ext_idx(g, batch_idx, seq_idx, ext, static_cast<cudaStream_t>(stream));
}

size_t cpcb200_criterion_save_bytes(const cpcb200_dims* d) {
  Geo g;
  if (make_geo(d, &g) || check_crit(g)) return 0;
  return criterion_save_bytes(g);
}
size_t cpcb200_criterion_ws_bytes(const cpcb200_dims* d, int backward) {
  Geo g;
  if (make_geo(d, &g) || check_crit(g)) return 0;
  return criterion_ws_bytes(g, backward);
}
int cpcb200_criterion_fwd(const cpcb200_dims* d, const float* c, const float* z, const float* w_pred, const int32_t* ext,
                          float* losses, float* acc, void* save, void* ws, size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  CPC_TRY(check_crit(g));
  NOT_NULL(c); NOT_NULL(z); NOT_NULL(w_pred); NOT_NULL(ext); NOT_NULL(losses); NOT_NULL(acc); NOT_NULL(save); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return criterion_fwd(g, c, z, w_pred, nullptr, ext, losses, acc, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cpcb200_criterion_bwd(const cpcb200_dims* d, const float* c, const float* z, const float* w_pred, const int32_t* ext,
                          const float* dlosses, const void* save, float* dc, float* dz, float* dw_pred, void* ws,
                          size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  CPC_TRY(check_crit(g));
  NOT_NULL(c); NOT_NULL(z); NOT_NULL(w_pred); NOT_NULL(ext); NOT_NULL(dlosses); NOT_NULL(save); NOT_NULL(dc); NOT_NULL(dz);
  NOT_NULL(dw_pred); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return criterion_bwd(g, c, z, w_pred, nullptr, ext, dlosses, save, dc, dz, dw_pred, nullptr, ws, ws_bytes,
                       static_cast<cudaStream_t>(stream));
}

static int thead_geo(const cpcb200_dims* d, int dff, int nheads, Geo* g) {
  CPC_TRY(make_geo(d, g));
  CPC_TRY(check_crit(*g));
  if (dff <= 0 || dff % 64 != 0 || nheads <= 0 || g->H % nheads != 0)
    return fail(CPCB200_ERR_BAD_DIMS, "transformer heads: dff=%d (multiple of 64), nheads=%d (divides H=%d)", dff, nheads, g->H);
  if (g->H != g->Har) return fail(CPCB200_ERR_UNSUPPORTED, "transformer heads need hiddenGar == hiddenEncoder");
  if (g->W > 128) return fail(CPCB200_ERR_UNSUPPORTED, "transformer heads: W=%d > 128", g->W);
  g->dff = dff;
  g->nheads = nheads;
  return 0;
}
size_t cpcb200_criterion_t_save_bytes(const cpcb200_dims* d, int dff, int nheads) {
  Geo g;
  if (thead_geo(d, dff, nheads, &g)) return 0;
  return criterion_save_bytes(g);
}
size_t cpcb200_criterion_t_ws_bytes(const cpcb200_dims* d, int dff, int nheads, int backward) {
  Geo g;
  if (thead_geo(d, dff, nheads, &g)) return 0;
  return criterion_ws_bytes(g, backward);
}
int cpcb200_criterion_t_fwd(const cpcb200_dims* d, const float* c, const float* z, const cpcb200_thead_params* p,
                            const int32_t* ext, float* losses, float* acc, void* save, void* ws, size_t ws_bytes, void* stream) {
  NOT_NULL(p);
  Geo g;
  CPC_TRY(thead_geo(d, p->dff, p->nheads, &g));
  NOT_NULL(c); NOT_NULL(z); NOT_NULL(ext); NOT_NULL(losses); NOT_NULL(acc); NOT_NULL(save); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return criterion_fwd(g, c, z, nullptr, p, ext, losses, acc, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cpcb200_criterion_t_bwd(const cpcb200_dims* d, const float* c, const float* z, const cpcb200_thead_params* p,
                            const int32_t* ext, const float* dlosses, const void* save, float* dc, float* dz,
                            const cpcb200_thead_params* grads, void* ws, size_t ws_bytes, void* stream) {
  NOT_NULL(p); NOT_NULL(grads);
  Geo g;
  CPC_TRY(thead_geo(d, p->dff, p->nheads, &g));
  NOT_NULL(c); NOT_NULL(z); NOT_NULL(ext); NOT_NULL(dlosses); NOT_NULL(save); NOT_NULL(dc); NOT_NULL(dz); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return criterion_bwd(g, c, z, nullptr, p, ext, dlosses, save, dc, dz, nullptr, grads, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int cpcb200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  NOT_NULL(param); NOT_NULL(grad); NOT_NULL(exp_avg); NOT_NULL(exp_avg_sq);
  if (step < 1) return fail(CPCB200_ERR_BAD_DIMS, "adam: step=%d must be >= 1", step);
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CPC_CHECK_CUDA(launch_k(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, st, 1, param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                          beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2)));
  CPC_LAUNCHED_N("adam", st);
  return 0;
}

int cpcb200_encoder_bwd_set_event(void* cuda_event) {
  g_grads_ready_event.store(cuda_event);
  return 0;
}

int cpcb200_adam_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int32_t* state, int zero_grad, void* stream) {
  NOT_NULL(param); NOT_NULL(grad); NOT_NULL(exp_avg); NOT_NULL(exp_avg_sq); NOT_NULL(state);
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(CPCB200_ERR_BAD_DIMS, "adam_step_dev: buffers must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(state) & 7) return fail(CPCB200_ERR_BAD_DIMS, "adam_step_dev: state must be 8-byte aligned");
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zero_grad)
    CPC_CHECK_CUDA(launch_k(adam_dev_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, st, 1, param, grad, exp_avg, exp_avg_sq, n,
                            lr, beta1, beta2, eps, weight_decay, reinterpret_cast<int*>(state)));
  else
    CPC_CHECK_CUDA(launch_k(adam_dev_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, st, 1, param, grad, exp_avg, exp_avg_sq, n,
                            lr, beta1, beta2, eps, weight_decay, reinterpret_cast<int*>(state)));
  CPC_LAUNCHED_N("adam", st);
  return 0;
}

int cpcb200_allreduce_adam_step(const cpcb200_peers* peers, float* param, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int32_t* state, int zero_grad,
                                void* stream) {
  NOT_NULL(peers); NOT_NULL(param); NOT_NULL(exp_avg); NOT_NULL(exp_avg_sq); NOT_NULL(state);
  if (peers->world < 1 || peers->world > 8 || peers->rank < 0 || peers->rank >= peers->world)
    return fail(CPCB200_ERR_BAD_DIMS, "allreduce_adam: rank %d / world %d (1..8 GPUs of one node)", peers->rank, peers->world);
  PeerPtrs P{};
  P.rank = peers->rank; P.world = peers->world;
  P.mc = static_cast<float*>(peers->grads_mc);
  if (reinterpret_cast<uintptr_t>(P.mc) & 15) return fail(CPCB200_ERR_BAD_DIMS, "allreduce_adam: multicast pointer must be 16-byte aligned");
  for (int i = 0; i < peers->world; i++) {
    if (!peers->grads[i] || !peers->signals[i]) return fail(CPCB200_ERR_NULL, "allreduce_adam: peer %d pointer is NULL", i);
    if (reinterpret_cast<uintptr_t>(peers->grads[i]) & 15) return fail(CPCB200_ERR_BAD_DIMS, "allreduce_adam: gradient buffers must be 16-byte aligned");
    P.g[i] = static_cast<float*>(peers->grads[i]);
    P.sig[i] = static_cast<unsigned*>(peers->signals[i]);
  }
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return fail(CPCB200_ERR_BAD_DIMS, "allreduce_adam: buffers must be 16-byte aligned");
  if (reinterpret_cast<uintptr_t>(state) & 7) return fail(CPCB200_ERR_BAD_DIMS, "allreduce_adam: state must be 8-byte aligned");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  CPC_CHECK_CUDA(cudaGetDevice(&dev));
  CPC_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const void* kern = zero_grad ? reinterpret_cast<const void*>(allreduce_adam_kernel<true>)
                               : reinterpret_cast<const void*>(allreduce_adam_kernel<false>);
  int* state_i = reinterpret_cast<int*>(state);
  void* args[] = {&P, &param, &exp_avg, &exp_avg_sq, &n, &lr, &beta1, &beta2, &eps, &weight_decay, &state_i};
  CPC_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)(2 * sms)), dim3(256), args, 0, st));  // 2 CTAs per SM, all resident
  CPC_LAUNCHED_N("allreduce_adam", st);
  return 0;
}

int cpcb200_test_gemm_nt(int dtype, int M, int N, int Kd, const void* A, const void* B, const float* bias, float* C, void* stream) {
  NOT_NULL(A); NOT_NULL(B); NOT_NULL(C);
  RowView a{A, 0, (long long)Kd, M};
  OutView c{C, 0, (long long)N, M, 0, M, 0};
  return gemm_nt(dtype == CPCB200_BF16, true, 1, N, Kd, a, B, bias, c, static_cast<cudaStream_t>(stream));
}
int cpcb200_test_gemm_nt_act(int dtype, int M, int N, int Kd, const void* A, const void* B, const float* bias, void* C, void* stream) {
  NOT_NULL(A); NOT_NULL(B); NOT_NULL(C);
  RowView a{A, 0, (long long)Kd, M};
  OutView c{C, 0, (long long)N, M, 0, M, 0};
  return gemm_nt(dtype == CPCB200_BF16, false, 1, N, Kd, a, B, bias, c, static_cast<cudaStream_t>(stream));
}
int cpcb200_debug_gemm_timeline(unsigned long long* host_out) {
  NOT_NULL(host_out);
  return debug_gemm_timeline(host_out);
}
int cpcb200_test_gemm_tn(int dtype, int M, int N1, int N2, const void* A, const void* B, float* C, void* stream) {
  NOT_NULL(A); NOT_NULL(B); NOT_NULL(C);
  RowView a{A, 0, (long long)N1, M};
  RowView b{B, 0, (long long)N2, M};
  return gemm_tn(dtype == CPCB200_BF16, 1, N1, N2, a, b, C, N2, STORE_PLAIN, 0, 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
