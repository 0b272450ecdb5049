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

// one-shot 'early gradients are final' events, keyed by the stream the encoder backward will run on (one entry per
// armed stream: DataParallel threads / several processes' streams do not see each other's events)
static std::mutex g_ev_mu;
static std::map<cudaStream_t, cudaEvent_t> g_grads_ready_events;
cudaEvent_t take_grads_ready_event(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ev_mu);
  auto it = g_grads_ready_events.find(st);
  if (it == g_grads_ready_events.end()) return nullptr;
  cudaEvent_t ev = it->second;
  g_grads_ready_events.erase(it);
  return ev;
}

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed) != 0; }

static thread_local int g_sm_reserve = 0;
int sm_reserve() { return g_sm_reserve; }
void set_sm_reserve(int n) { g_sm_reserve = n < 0 ? 0 : (n > 64 ? 64 : n); }

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
  if (d->B <= 0 || d->L <= 0) return fail(CPCB200_ERR_BAD_DIMS, "B=%d L=%d", d->B, d->L);
  if (d->H % 64 != 0 || d->H <= 0 || d->H > 512) return fail(CPCB200_ERR_BAD_DIMS, "H=%d must be a multiple of 64 in [64,512]", d->H);
  if (d->Har % 64 != 0 || d->Har <= 0 || d->Har > 512) return fail(CPCB200_ERR_BAD_DIMS, "Har=%d must be a multiple of 64 in [64,512]", d->Har);
  if (d->nLayers < 1 || d->nLayers > CPCB200_MAX_GRU_LAYERS) return fail(CPCB200_ERR_BAD_DIMS, "nLayers=%d", d->nLayers);
  if (d->dtype != CPCB200_F32 && d->dtype != CPCB200_BF16) return fail(CPCB200_ERR_UNSUPPORTED, "dtype=%d", d->dtype);
  g->B = d->B; g->L = d->L; g->H = d->H; g->Har = d->Har; g->K = d->K; g->N = d->N; g->nL = d->nLayers;
  g->bf16 = d->dtype == CPCB200_BF16;
  // frames per window: the conv stack of cpc/model.py:83-92 (L/160 when L is a multiple of 160; any L that leaves at least
  // one frame after every layer is accepted - feature extraction feeds chunks of arbitrary length, feature_loader.py:228-269)
  int L = d->L;
  for (int i = 0; i < 5; i++) {
    const int num = L + 2 * kConvP[i] - kConvK[i];
    if (num < 0) return fail(CPCB200_ERR_BAD_DIMS, "L=%d samples is too short for the encoder (no output frame after conv%d)", d->L, i);
    L = num / kConvS[i] + 1;
    g->Lout[i] = L;
  }
  g->S = g->Lout[4];
  g->W = g->S - d->K;
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

size_t lstm_save_bytes(const Geo& g);
size_t lstm_ws_bytes(const Geo& g, int mode);
int lstm_fwd(const Geo&, const float*, const float*, const float*, const cpcb200_gru_params*, float*, float*, float*, void*, void*,
             size_t, cudaStream_t);
int lstm_bwd(const Geo&, const float*, const float*, const float*, const cpcb200_gru_params*, const float*, const float*, const void*,
             float*, const cpcb200_gru_params*, void*, size_t, cudaStream_t);
size_t tlayer_save_bytes(const Geo& g);
size_t tlayer_ws_bytes(const Geo& g, int backward);
int tlayer_fwd(const Geo&, const float*, const cpcb200_thead_params*, float*, void*, void*, size_t, cudaStream_t);
int tlayer_bwd(const Geo&, const float*, const cpcb200_thead_params*, const float*, const void*, float*, const cpcb200_thead_params*,
               void*, size_t, cudaStream_t);

int gather_windows(const float*, long long, const long long*, int, int, float*, const long long*, int, long long*, int*, cudaStream_t);

int gemm_nt_simt(bool, bool, int, int, int, const RowView&, const void*, const float*, const OutView&, cudaStream_t);
int gemm_tn_simt(bool, int, int, int, const RowView&, const RowView&, float*, int, int, int, int, cudaStream_t);
int debug_gemm_timeline(unsigned long long* host_out);
int gemm_nt_tc(bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C,
               cudaStream_t st, bool* handled, const HeadBatch* hb = nullptr);
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
int gemm_nt_heads(bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias, const OutView& C,
                  int a_mod, int w_div, cudaStream_t st) {
  HeadBatch hb{};
  hb.a_mod = a_mod; hb.w_div = w_div;
  bool handled = false;
  CPC_TRY(gemm_nt_tc(out_f32, nb, N, Kd, A, Bm, bias, C, st, &handled, &hb));
  if (!handled) return fail(CPCB200_ERR_UNSUPPORTED, "gemm_nt_heads: N=%d Kd=%d does not fit the 128 x 256 tensor-core tiling", N, Kd);
  return 0;
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

// words of the device-side optimizer state (16 x int32, include/cpc_b200.h)
constexpr int kStSteps = 0, kStTicket = 1, kStPow = 2, kStEpoch = 6, kStErr = 7, kStLr = 8, kStEarlyEpoch = 9, kStEarlyTicket = 10;

// graph-capturable variant: step count on the device (state[0]) together with beta1^step and beta2^step as running
// float64 products (state[2..5]; double-precision pow() in the kernel costs ~90 us on this part), the last block to
// finish (ticket in state[1]) publishes the next step's values - every block has read the current ones by then
template <bool ZERO>
__global__ void adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                size_t n, float lr, float b1, float b2, float eps, float wd, int* __restrict__ state) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_bc[3];
  double* pw = reinterpret_cast<double*>(state + 2);  // beta1^(steps+1), beta2^(steps+1) once steps > 0
  double p1 = 0.0, p2 = 0.0;
  if (threadIdx.x == 0) {
    const int done = *reinterpret_cast<volatile int*>(state);
    p1 = done == 0 ? (double)b1 : *reinterpret_cast<volatile double*>(pw);
    p2 = done == 0 ? (double)b2 : *reinterpret_cast<volatile double*>(pw + 1);
    s_bc[0] = (float)(1.0 - p1);
    s_bc[1] = sqrtf((float)(1.0 - p2));
    s_bc[2] = lr < 0.f ? __int_as_float(*reinterpret_cast<volatile int*>(state + kStLr)) : lr;  // lr < 0: device-side learning rate
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  const float step_size = s_bc[2] / bc1;
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

// ---- all-reduce + Adam + zero_grad over peer memory (cpcb200_allreduce_adam_step, cpcb200_peer_reduce_range) --------
struct PeerPtrs { float* g[8]; unsigned* sig[8]; int rank, world; float* mc; long long timeout_ns; };
struct RangeList { long long lo[4], hi[4]; int n; };  // [lo, hi) in floats; lo a multiple of 4
// signal words of a rank (>= 64 x uint32): [0..7] arrivals at the step kernel's barriers (word j is written by rank j),
// [16..23] arrivals at the early-reduce kernel's barrier, [40..44] phase stamps of the last step kernel (debug)
constexpr int kSigStep = 0, kSigEarly = 16, kSigDbg = 40, kSigAbs = 48;  // [48..55]: absolute globaltimer stamps (4 x u64)

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
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Slice data is moved with ordinary L2-level accesses (ld.global.cg / st.global.cg): the node barriers (release / acquire
// at system scope) order them, and sys-scoped data accesses measured ~4x slower than the link on this path.
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_relaxed_sys_v4(float* p, const float4& v) { __stcg(reinterpret_cast<float4*>(p), v); }
// node-wide barrier on signal words.  arrive: one thread per peer (block 0) tells that peer this rank reached `epoch`;
// wait: the first `world` threads of EVERY block poll this rank's own words (local memory, written by the peers) - no
// grid-wide sync is needed to release the other blocks.  Epochs only grow, so a peer that is already one barrier ahead
// still satisfies the wait.  A wait that exceeds P.timeout_ns returns false (block-uniform): the caller records the
// error in the optimizer state and leaves the kernel without touching the parameters - no trap, the context survives.
__device__ __forceinline__ void node_arrive(const PeerPtrs& P, int base, unsigned epoch) {
  if (blockIdx.x == 0 && (int)threadIdx.x < P.world) {
    __threadfence_system();
    st_release_sys(P.sig[threadIdx.x] + base + P.rank, epoch);
  }
}
__device__ __forceinline__ bool node_wait(const PeerPtrs& P, int base, unsigned epoch) {
  int bad = 0;
  if ((int)threadIdx.x < P.world) {
    const unsigned* mine = P.sig[P.rank] + base + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > (unsigned long long)P.timeout_ns) { bad = 1; break; }
    }
  }
  return __syncthreads_or(bad) == 0;
}

// Two-shot all-reduce of the float range [lo, hi), in place: this rank sums its 1/world slice of every rank's copy and
// writes the sum back into that slice of every copy.  With a multicast mapping the sum is ONE multimem.ld_reduce per 16
// bytes, computed inside the NVSwitch, and the write-back ONE multimem.st.
// (peer loads cost ~2 us each: all the loads of U quads are issued before the first add)
template <int U>
__device__ __forceinline__ void reduce_my_slice(const PeerPtrs& P, long long lo_f, long long hi_f, size_t tid, size_t nthr) {
  const size_t n4 = (size_t)(hi_f - lo_f) / 4, base4 = (size_t)lo_f / 4;
  const size_t per = (n4 + P.world - 1) / P.world;
  const size_t lo = base4 + per * P.rank;
  size_t hi = lo + per;
  if (hi > base4 + n4) hi = base4 + n4;
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
  } else {
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
  }
  if (P.rank == P.world - 1) {  // the < 4 floats past the last quad of the range
    for (size_t i = (size_t)lo_f + n4 * 4 + tid; i < (size_t)hi_f; i += nthr) {
      float acc = 0.f;
      for (int j = 0; j < P.world; j++) acc += *reinterpret_cast<volatile float*>(P.g[j] + i);
      for (int j = 0; j < P.world; j++) *reinterpret_cast<volatile float*>(P.g[j] + i) = acc;
    }
  }
}

// Early exchange (cpcb200_peer_reduce_range): all-reduce of the given ranges of the bucket on a side stream while the
// backward pass is still producing the remaining (late) gradients.  An ordinary (non-cooperative) launch of a few CTAs:
// no CTA ever waits for another CTA of the same grid, only for the peers' arrival words.  Completion needs no signal of
// its own: a rank enters the step kernel's first barrier only after its own early kernel has finished (stream order), so
// once every rank has arrived there every slice of every early range has been written everywhere.
// Placement: a few CTAs of 1024 threads (a whole register file each), launched as ONE cluster so that they take
// neighbouring SMs of one GPC, while the data-gradient GEMM that runs at the same time is launched on 148 - kEarlyExchangeSMs
// SMs (sm_reserve).  Measured on B200: a CTA of this kernel and a persistent GEMM CTA never share an SM (the GEMM CTA waits
// until the other one exits, whatever its size: 1 / 4 / 32 / 148 small CTAs cost +600 / +100 / +34 / +27 us per step) and the
// GEMM's static tile schedule then waits for the slowest SM - so the exchange gets SMs of its own instead.
__global__ void __launch_bounds__(1024, 1) peer_reduce_kernel(PeerPtrs P, RangeList R, int* __restrict__ state) {
  if (*reinterpret_cast<volatile int*>(state + kStErr) != 0) return;  // set by an earlier kernel only: grid-uniform
  const unsigned ep = (unsigned)*reinterpret_cast<volatile int*>(state + kStEarlyEpoch);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  unsigned long long* abs_t = reinterpret_cast<unsigned long long*>(P.sig[P.rank] + kSigAbs);  // debug: absolute times
  const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) abs_t[0] = globaltimer_ns();
  node_arrive(P, kSigEarly, ep + 1);
  const bool ok = node_wait(P, kSigEarly, ep + 1);
  if (stamp) abs_t[1] = globaltimer_ns();
  if (ok) {
    if (P.mc != nullptr) { for (int r = 0; r < R.n; r++) reduce_my_slice<8>(P, R.lo[r], R.hi[r], tid, nthr); }
    else { for (int r = 0; r < R.n; r++) reduce_my_slice<4>(P, R.lo[r], R.hi[r], tid, nthr); }
  } else if (threadIdx.x == 0) {
    atomicExch(state + kStErr, 3);
  }
  __threadfence_system();  // this thread's peer stores are performed system-wide before the kernel counts as complete
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(state + kStEarlyTicket, 1);
    if (t == (int)gridDim.x - 1) {  // last block out: every block has read the epoch
      abs_t[2] = globaltimer_ns();
      state[kStEarlyTicket] = 0;
      __threadfence();
      atomicAdd(state + kStEarlyEpoch, 1);
    }
  }
}

template <bool ZERO>
__global__ void __launch_bounds__(256) allreduce_adam_kernel(PeerPtrs P, RangeList R, float* __restrict__ p, float* __restrict__ m,
                                                             float* __restrict__ v, size_t n, float lr, float b1, float b2,
                                                             float eps, float wd, int* __restrict__ state) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  if (*reinterpret_cast<volatile int*>(state + kStErr) != 0) return;  // a previous exchange failed: grid-uniform, nothing is touched
  __shared__ float s_bc[3];
  double* pw = reinterpret_cast<double*>(state + kStPow);
  double p1 = 0.0, p2 = 0.0;
  unsigned calls = 0;
  if (threadIdx.x == 0) {
    const int done = *reinterpret_cast<volatile int*>(state);
    p1 = done == 0 ? (double)b1 : *reinterpret_cast<volatile double*>(pw);
    p2 = done == 0 ? (double)b2 : *reinterpret_cast<volatile double*>(pw + 1);
    s_bc[0] = (float)(1.0 - p1);
    s_bc[1] = sqrtf((float)(1.0 - p2));
    s_bc[2] = lr < 0.f ? __int_as_float(*reinterpret_cast<volatile int*>(state + kStLr)) : lr;
  }
  calls = (unsigned)*reinterpret_cast<volatile int*>(state + kStEpoch);  // node-barrier epoch base: 2 barriers per call
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  const float step_size = s_bc[2] / bc1;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  float* g = P.g[P.rank];

  // phase stamps (ns since kernel start) in this rank's signal words 40..44: read by tools/peer_adam_check.py
  const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t_start = stamp ? globaltimer_ns() : 0ull;
  unsigned* dbg = P.sig[P.rank] + kSigDbg;
  if (stamp) reinterpret_cast<unsigned long long*>(P.sig[P.rank] + kSigAbs)[3] = t_start;
  // ---- every rank's gradients are final (and every rank's early exchange, if any, has completed) ----
  node_arrive(P, kSigStep, 2 * calls + 1);
  const bool ok1 = node_wait(P, kSigStep, 2 * calls + 1);
  if (stamp) dbg[0] = (unsigned)(globaltimer_ns() - t_start);
  // ---- two-shot all-reduce, in place, of the ranges that have not been exchanged yet ----
  if (ok1) {
    for (int r = 0; r < R.n; r++) reduce_my_slice<8>(P, R.lo[r], R.hi[r], tid, nthr);
  } else if (threadIdx.x == 0) {
    atomicExch(state + kStErr, 1);
  }
  __threadfence_system();  // this thread's peer stores are performed system-wide before it reports in
  if (stamp) dbg[1] = (unsigned)(globaltimer_ns() - t_start);
  grid.sync();  // every block of this rank is past its fence: block 0 may tell the peers
  if (*reinterpret_cast<volatile int*>(state + kStErr) != 0) return;  // grid-uniform after the sync: no update is applied
  if (stamp) dbg[2] = (unsigned)(globaltimer_ns() - t_start);
  // ---- every slice of the local buffer has been written by its owner ----
  node_arrive(P, kSigStep, 2 * calls + 2);
  if (!node_wait(P, kSigStep, 2 * calls + 2)) {  // a peer died inside this very kernel: give up, flag, keep the context alive
    if (threadIdx.x == 0) atomicExch(state + kStErr, 2);
    return;
  }
  if (stamp) dbg[3] = (unsigned)(globaltimer_ns() - t_start);
  // ---- Adam on the full local replica, gradients read past the L1 (peers wrote them) ----
  auto upd = [&](float gi, float pi, float& mi, float& vi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = fmaf(b1, mi, (1.f - b1) * gi);
    vi = fmaf(b2, vi, (1.f - b2) * gi * gi);
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    return pi - step_size * (mi / denom);
  };
  const size_t n4 = n / 4;
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
    state[kStEpoch] = (int)(calls + 1);
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
  NOT_NULL(x); NOT_NULL(p); NOT_NULL(z); NOT_NULL(ws);  // save == NULL: inference forward
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

size_t cpcb200_lstm_save_bytes(const cpcb200_dims* d) {
  Geo g;
  if (make_geo(d, &g)) return 0;
  return lstm_save_bytes(g);
}
size_t cpcb200_lstm_ws_bytes(const cpcb200_dims* d, int mode) {
  Geo g;
  if (make_geo(d, &g)) return 0;
  return lstm_ws_bytes(g, mode);
}
int cpcb200_lstm_fwd(const cpcb200_dims* d, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p, float* out,
                     float* hT, float* cT, void* save, void* ws, size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  NOT_NULL(z); NOT_NULL(p); NOT_NULL(out); NOT_NULL(ws);  // save == NULL: inference forward
  prof_mark(static_cast<cudaStream_t>(stream));
  return lstm_fwd(g, z, h0, c0, p, out, hT, cT, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cpcb200_lstm_bwd(const cpcb200_dims* d, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p,
                     const float* out, const float* dout, const void* save, float* dz, const cpcb200_gru_params* grads, void* ws,
                     size_t ws_bytes, void* stream) {
  GEO_OR_RETURN(d, g);
  NOT_NULL(z); NOT_NULL(p); NOT_NULL(out); NOT_NULL(dout); NOT_NULL(save); NOT_NULL(dz); NOT_NULL(grads); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return lstm_bwd(g, z, h0, c0, p, out, dout, save, dz, grads, ws, ws_bytes, static_cast<cudaStream_t>(stream));
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

// transformer context network: one TransformerLayer over the S frames of every window (d->Har, d->K, d->N are ignored)
static int tlayer_geo(const cpcb200_dims* d, int dff, int nheads, Geo* g) {
  CPC_TRY(make_geo(d, g));
  if (dff <= 0 || dff % 64 != 0 || nheads <= 0 || g->H % nheads != 0)
    return fail(CPCB200_ERR_BAD_DIMS, "transformer layer: dff=%d (multiple of 64), nheads=%d (divides H=%d)", dff, nheads, g->H);
  if (g->S > 128) return fail(CPCB200_ERR_UNSUPPORTED, "transformer layer: %d frames per window > 128", g->S);
  g->Har = g->H; g->K = 1; g->N = 1; g->W = g->S;
  g->dff = dff; g->nheads = nheads;
  return 0;
}
size_t cpcb200_tlayer_save_bytes(const cpcb200_dims* d, int dff, int nheads) {
  Geo g;
  if (tlayer_geo(d, dff, nheads, &g)) return 0;
  return tlayer_save_bytes(g);
}
size_t cpcb200_tlayer_ws_bytes(const cpcb200_dims* d, int dff, int nheads, int backward) {
  Geo g;
  if (tlayer_geo(d, dff, nheads, &g)) return 0;
  return tlayer_ws_bytes(g, backward);
}
int cpcb200_tlayer_fwd(const cpcb200_dims* d, const float* x, const cpcb200_thead_params* p, float* y, void* save, void* ws,
                       size_t ws_bytes, void* stream) {
  NOT_NULL(p);
  Geo g;
  CPC_TRY(tlayer_geo(d, p->dff, p->nheads, &g));
  NOT_NULL(x); NOT_NULL(y); NOT_NULL(save); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return tlayer_fwd(g, x, p, y, save, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int cpcb200_tlayer_bwd(const cpcb200_dims* d, const float* x, const cpcb200_thead_params* p, const float* dy, const void* save,
                       float* dx, const cpcb200_thead_params* grads, void* ws, size_t ws_bytes, void* stream) {
  NOT_NULL(p); NOT_NULL(grads);
  Geo g;
  CPC_TRY(tlayer_geo(d, p->dff, p->nheads, &g));
  NOT_NULL(x); NOT_NULL(dy); NOT_NULL(save); NOT_NULL(dx); NOT_NULL(ws);
  prof_mark(static_cast<cudaStream_t>(stream));
  return tlayer_bwd(g, x, p, dy, save, dx, grads, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int cpcb200_gather_windows(const float* data, int64_t n_samples, const int64_t* starts, int B, int L, float* out,
                           const int64_t* bounds, int n_bounds, int64_t* labels, int32_t* err, void* stream) {
  NOT_NULL(data); NOT_NULL(starts); NOT_NULL(out); NOT_NULL(err);
  if (labels != nullptr && (bounds == nullptr || n_bounds < 2)) return fail(CPCB200_ERR_NULL, "gather_windows: labels need >= 2 interval bounds");
  return gather_windows(data, (long long)n_samples, reinterpret_cast<const long long*>(starts), B, L, out,
                        reinterpret_cast<const long long*>(bounds), n_bounds, reinterpret_cast<long long*>(labels), err,
                        static_cast<cudaStream_t>(stream));
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

int cpcb200_encoder_bwd_set_event(void* stream, void* cuda_event) {
  std::lock_guard<std::mutex> lk(g_ev_mu);
  if (cuda_event) g_grads_ready_events[static_cast<cudaStream_t>(stream)] = static_cast<cudaEvent_t>(cuda_event);
  else g_grads_ready_events.erase(static_cast<cudaStream_t>(stream));
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

static int fill_peers(const cpcb200_peers* peers, PeerPtrs* P) {
  if (peers->world < 1 || peers->world > 8 || peers->rank < 0 || peers->rank >= peers->world)
    return fail(CPCB200_ERR_BAD_DIMS, "peer exchange: rank %d / world %d (1..8 GPUs of one node)", peers->rank, peers->world);
  P->rank = peers->rank; P->world = peers->world;
  P->mc = static_cast<float*>(peers->grads_mc);
  P->timeout_ns = peers->timeout_ns > 0 ? peers->timeout_ns : 600ll * 1000000000ll;
  if (reinterpret_cast<uintptr_t>(P->mc) & 15) return fail(CPCB200_ERR_BAD_DIMS, "peer exchange: multicast pointer must be 16-byte aligned");
  for (int i = 0; i < peers->world; i++) {
    if (!peers->grads[i] || !peers->signals[i]) return fail(CPCB200_ERR_NULL, "peer exchange: peer %d pointer is NULL", i);
    if (reinterpret_cast<uintptr_t>(peers->grads[i]) & 15) return fail(CPCB200_ERR_BAD_DIMS, "peer exchange: gradient buffers must be 16-byte aligned");
    P->g[i] = static_cast<float*>(peers->grads[i]);
    P->sig[i] = static_cast<unsigned*>(peers->signals[i]);
  }
  return 0;
}
static int fill_ranges(const int64_t* ranges, int n_ranges, size_t n, RangeList* R) {
  if (n_ranges < 0 || n_ranges > 4) return fail(CPCB200_ERR_BAD_DIMS, "peer exchange: %d ranges (at most 4)", n_ranges);
  R->n = n_ranges;
  for (int i = 0; i < n_ranges; i++) {
    const long long lo = ranges[2 * i], hi = ranges[2 * i + 1];
    if (lo < 0 || hi < lo || (lo & 3) || (n && (size_t)hi > n))
      return fail(CPCB200_ERR_BAD_DIMS, "peer exchange: range %d = [%lld, %lld) (start must be a multiple of 4 floats)", i, lo, hi);
    R->lo[i] = lo; R->hi[i] = hi;
  }
  return 0;
}

int cpcb200_peer_reduce_range(const cpcb200_peers* peers, const int64_t* ranges, int n_ranges, int32_t* state, void* stream) {
  NOT_NULL(peers); NOT_NULL(ranges); NOT_NULL(state);
  PeerPtrs P{};
  CPC_TRY(fill_peers(peers, &P));
  RangeList R{};
  CPC_TRY(fill_ranges(ranges, n_ranges, 0, &R));
  if (n_ranges == 0) return 0;
  static const bool noop = []() { const char* e = getenv("CPC_B200_EARLY_NOOP"); return e && atoi(e) == 1; }();  // debug: fork / join only
  if (noop) return 0;
  static const int ctas = []() { const char* e = getenv("CPC_B200_EARLY_CTAS"); int v = e ? atoi(e) : kEarlyExchangeSMs; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;  // one cluster: the CTAs land on SMs of one GPC
    at[0].val.clusterDim.x = (unsigned)ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int* state_i = reinterpret_cast<int*>(state);
    CPC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, peer_reduce_kernel, P, R, state_i));
  }
  CPC_LAUNCHED_N("peer_reduce", st);
  return 0;
}

int cpcb200_allreduce_adam_step(const cpcb200_peers* peers, float* param, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int32_t* state, int zero_grad,
                                const int64_t* ranges, int n_ranges, void* stream) {
  NOT_NULL(peers); NOT_NULL(param); NOT_NULL(exp_avg); NOT_NULL(exp_avg_sq); NOT_NULL(state);
  PeerPtrs P{};
  CPC_TRY(fill_peers(peers, &P));
  RangeList R{};
  if (ranges != nullptr) CPC_TRY(fill_ranges(ranges, n_ranges, n, &R));
  else { R.n = 1; R.lo[0] = 0; R.hi[0] = (long long)n; }  // nothing exchanged yet: the whole bucket
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
  void* args[] = {&P, &R, &param, &exp_avg, &exp_avg_sq, &n, &lr, &beta1, &beta2, &eps, &weight_decay, &state_i};
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
