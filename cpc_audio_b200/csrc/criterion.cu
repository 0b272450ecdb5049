// criterion.cu - CPCUnsupersivedCriterion hot path (reference: cpc/criterion/criterion.py).
//   sample_ext_idx : criterion.py:191-199  index arithmetic on the two torch.randint draws (bit-exact, int)
//   heads          : criterion.py:106-108  K x Linear(Har->H, bias=False) as ONE GEMM  (B*W, Har) x (Har, K*H)
//   score + CE     : criterion.py:115-117, 207-217, 245-257 - the reference materialises K x (B, N+1, W, H)
//                    candidate tensors (11.8 GB at B=64); here the negatives are gathered from z on the fly,
//                    shared by the K heads, and the cross-entropy / accuracy are reduced in the same kernel.
#include "common.cuh"

namespace cpcb200 {

int gemm_nt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias,
            const OutView& C, cudaStream_t st);
int gemm_tn(bool bf16_in, int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc, int mode,
            int Ci, int taps, cudaStream_t st);
template <class T> int launch_cast(const float* src, T* dst, long long n, cudaStream_t st);
template <class T> int launch_transpose_cast(const float* src, T* dst, int R, int C, cudaStream_t st, int nb = 1);

bool score_mma_supported(int H, int K, int N);
int score_transpose_ext(const int* ext, int* ext_t, int B, int N, int W, cudaStream_t st);
int score_fwd_mma(const bf16* pred, const bf16* z, const int* ext_t, float* lossbuf, float* corrbuf, float* lsebuf, int B, int S,
                  int W, int H, int K, int N, cudaStream_t st);
int score_bwd_mma(const bf16* pred, const bf16* z, const int* ext_t, const float* lsebuf, const float* dloss, bf16* dpred,
                  float* dz, int B, int S, int W, int H, int K, int N, cudaStream_t st);

int launch_cast3_bf16(const float* s0, bf16* d0, long long n0, const float* s1, bf16* d1, long long n1, const float* s2, bf16* d2,
                      long long n2, cudaStream_t st);

size_t thead_save_bytes(const Geo& g);
size_t thead_ws_bytes(const Geo& g, int backward);
template <class T> int thead_fwd(const Geo& g, const T* cp, const cpcb200_thead_params* tp, T* pred, void* save, Carver& ws, cudaStream_t st);
template <class T> int thead_bwd(const Geo& g, const T* cp, const cpcb200_thead_params* tp, const T* dpred, const void* save, float* dc,
                                 const cpcb200_thead_params* gr, Carver& ws, cudaStream_t st);

namespace {

constexpr int KMAX = 16;

__global__ void ext_idx_kernel(const long long* __restrict__ bi, const long long* __restrict__ si, int* __restrict__ ext,
                               long long n, int W, int S) {
  pdl_wait();
  pdl_trigger();
  const unsigned nn = (unsigned)n;  // n < 2^31 (checked on the host): 32-bit index arithmetic
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    const unsigned w = i % (unsigned)W;
    const long long sraw = si[i];
    long long s;
    if (sraw >= 0 && sraw < S) {  // what torch.randint(1, S) draws: one conditional subtraction instead of a 64-bit division
      int sv = (int)sraw + (int)w;
      if (sv >= S) sv -= S;
      s = sv;
    } else {                      // any other int64 input: torch.remainder semantics
      s = (sraw + (long long)w) % S;
      if (s < 0) s += S;
    }
    ext[i] = (int)(s + bi[i] * S);
  }
}

// ---------------------------------------------------------------------------------------------------------
// score + CE forward, CUDA-core version: one CTA per anchor position p = (b, w); thread i < N owns negative i
// (dots with all K predictions), thread N + k owns the positive of step k+1.
// logits out: (B*W, K, N+1) fp32 with class 0 = positive.
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void score_fwd_kernel(const T* __restrict__ pred, const T* __restrict__ z, const int* __restrict__ ext,
                                 float* __restrict__ logits, float* __restrict__ lossbuf, float* __restrict__ corrbuf,
                                 int B, int S, int W, int H, int K, int N) {
  extern __shared__ __align__(16) float sm[];
  float* psm = sm;                    // [H][KMAX] predictions of this position, k fastest
  float* lg = sm + (size_t)H * KMAX;  // [K][N+1]
  const int p = blockIdx.x;
  const int b = p / W, w = p - b * W;
  const int tid = threadIdx.x;
  for (int i = tid; i < H * KMAX; i += blockDim.x) psm[i] = 0.f;
  __syncthreads();
  const T* pp = pred + (size_t)p * K * H;
  for (int i = tid; i < K * H; i += blockDim.x) { int k = i / H, d = i - k * H; psm[d * KMAX + k] = to_f(pp[i]); }
  __syncthreads();

  const float invH = 1.f / (float)H;
  if (tid < N + K) {
    long long row;
    if (tid < N) row = ext[((size_t)b * N + tid) * W + w];
    else row = (long long)b * S + w + (tid - N) + 1;
    const T* zr = z + row * H;
    float acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) acc[k] = 0.f;
    for (int d0 = 0; d0 < H; d0 += 4) {
      float cv[4];
      load_vec<4>(zr + d0, cv);
#pragma unroll
      for (int dd = 0; dd < 4; dd++) {
        const float4* pr = reinterpret_cast<const float4*>(psm + (d0 + dd) * KMAX);
#pragma unroll
        for (int k4 = 0; k4 < KMAX / 4; k4++) {
          float4 pv = pr[k4];
          acc[4 * k4 + 0] = fmaf(pv.x, cv[dd], acc[4 * k4 + 0]);
          acc[4 * k4 + 1] = fmaf(pv.y, cv[dd], acc[4 * k4 + 1]);
          acc[4 * k4 + 2] = fmaf(pv.z, cv[dd], acc[4 * k4 + 2]);
          acc[4 * k4 + 3] = fmaf(pv.w, cv[dd], acc[4 * k4 + 3]);
        }
      }
    }
    if (tid < N) {
#pragma unroll
      for (int k = 0; k < KMAX; k++)
        if (k < K) lg[k * (N + 1) + 1 + tid] = acc[k] * invH;
    } else {
      const int k = tid - N;
#pragma unroll
      for (int kk = 0; kk < KMAX; kk++)
        if (kk == k) lg[k * (N + 1)] = acc[kk] * invH;
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    const float* l = lg + k * (N + 1);
    float m = -INFINITY;
    for (int j = lane; j <= N; j += 32) m = fmaxf(m, l[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j <= N; j += 32) s += expf(l[j] - m);
    s = warp_sum(s);
    if (lane == 0) {
      lossbuf[(size_t)p * K + k] = (m + logf(s)) - l[0];
      corrbuf[(size_t)p * K + k] = (l[0] >= m) ? 1.f : 0.f;  // argmax == 0 (first maximum wins, as torch.max)
    }
    float* lo = logits + ((size_t)p * K + k) * (N + 1);
    for (int j = lane; j <= N; j += 32) lo[j] = l[j];
  }
}

// deterministic mean over positions: out[k] = sum_p buf[p][k] / P
__global__ void mean_over_positions_kernel(const float* __restrict__ a, const float* __restrict__ bq, float* __restrict__ oa,
                                           float* __restrict__ ob, int P, int K) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s1[1024], s2[1024];  // 1024 threads: the strided column reads are latency-bound (7 trips instead of 29)
  const int k = blockIdx.x;
  float x = 0.f, y = 0.f;
  for (int p = threadIdx.x; p < P; p += 1024) { x += a[(size_t)p * K + k]; y += bq[(size_t)p * K + k]; }
  s1[threadIdx.x] = x; s2[threadIdx.x] = y;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s1[threadIdx.x] += s1[threadIdx.x + o]; s2[threadIdx.x] += s2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { oa[k] = s1[0] / (float)P; ob[k] = s2[0] / (float)P; }
}

// ---------------------------------------------------------------------------------------------------------
// score + CE backward: one CTA per anchor position, thread per feature d.
//   G[k][j] = (softmax(logits[k])[j] - [j==0]) * dloss[k] / (B*W) / H
//   dpred[k][d]      = sum_j G[k][j] cand_j[d]
//   dz[row_j][d]    += sum_k G[k][j] pred[k][d]      (negatives shared by all k -> one atomic row per negative)
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void score_bwd_kernel(const T* __restrict__ pred, const T* __restrict__ z, const int* __restrict__ ext,
                                 const float* __restrict__ logits, const float* __restrict__ dloss, T* __restrict__ dpred,
                                 float* __restrict__ dz, int B, int S, int W, int H, int K, int N) {
  extern __shared__ __align__(16) float sm[];
  float* G = sm;                                  // [N+1][KMAX]  (k fastest)
  int* rows = reinterpret_cast<int*>(sm + (size_t)(N + 1) * KMAX);  // [N]
  const int p = blockIdx.x;
  const int b = p / W, w = p - b * W;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int i = tid; i < (N + 1) * KMAX; i += blockDim.x) G[i] = 0.f;
  for (int i = tid; i < N; i += blockDim.x) rows[i] = ext[((size_t)b * N + i) * W + w];
  __syncthreads();
  const float scale = 1.f / ((float)B * (float)W * (float)H);
  for (int k = warp; k < K; k += nwarps) {
    const float* l = logits + ((size_t)p * K + k) * (N + 1);
    float m = -INFINITY;
    for (int j = lane; j <= N; j += 32) m = fmaxf(m, l[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j <= N; j += 32) s += expf(l[j] - m);
    s = warp_sum(s);
    const float inv = 1.f / s, gk = dloss[k] * scale;
    for (int j = lane; j <= N; j += 32) G[j * KMAX + k] = (expf(l[j] - m) * inv - (j == 0 ? 1.f : 0.f)) * gk;
  }
  __syncthreads();
  for (int d = tid; d < H; d += blockDim.x) {
    float pr[KMAX], acc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) { acc[k] = 0.f; pr[k] = (k < K) ? to_f(pred[((size_t)p * K + k) * H + d]) : 0.f; }
    // positives: class 0 of head k is z[b, w+k+1]
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      if (k < K) {
        const size_t row = (size_t)b * S + w + k + 1;
        const float g0 = G[k];
        acc[k] = fmaf(g0, to_f(z[row * H + d]), acc[k]);
        atomicAdd(dz + row * H + d, g0 * pr[k]);
      }
    }
    for (int j = 0; j < N; j++) {
      const size_t row = (size_t)rows[j];
      const float cv = to_f(z[row * H + d]);
      const float4* g4 = reinterpret_cast<const float4*>(G + (j + 1) * KMAX);
      float val = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < KMAX / 4; k4++) {
        float4 gv = g4[k4];
        acc[4 * k4 + 0] = fmaf(gv.x, cv, acc[4 * k4 + 0]); val = fmaf(gv.x, pr[4 * k4 + 0], val);
        acc[4 * k4 + 1] = fmaf(gv.y, cv, acc[4 * k4 + 1]); val = fmaf(gv.y, pr[4 * k4 + 1], val);
        acc[4 * k4 + 2] = fmaf(gv.z, cv, acc[4 * k4 + 2]); val = fmaf(gv.z, pr[4 * k4 + 2], val);
        acc[4 * k4 + 3] = fmaf(gv.w, cv, acc[4 * k4 + 3]); val = fmaf(gv.w, pr[4 * k4 + 3], val);
      }
      atomicAdd(dz + row * H + d, val);
    }
#pragma unroll
    for (int k = 0; k < KMAX; k++)
      if (k < K) dpred[((size_t)p * K + k) * H + d] = from_f<T>(acc[k]);
  }
}

struct CritLayout { size_t pred, logits, lse, ext_t, cT, zT, thead, total; bool mma; };
CritLayout crit_layout(const Geo& g) {
  CritLayout l{};
  const size_t es = g.bf16 ? 2 : 4;
  const size_t P = (size_t)g.B * g.W;
  l.mma = g.bf16 && score_mma_supported(g.H, g.K, g.N);
  l.pred = 0;
  l.logits = align_up(P * g.K * g.H * es);
  if (!l.mma) {  // CUDA-core scoring keeps the logits for its backward
    l.total = l.logits + align_up(P * g.K * (g.N + 1) * 4);
  } else {       // tensor-core scoring recomputes them: only lse per (p, k) and the transposed indices are kept
    l.lse = l.logits;
    l.ext_t = l.lse + align_up(P * g.K * 4);
    l.total = l.ext_t + align_up(P * g.N * 4);
  }
  if (g.bf16) {  // bf16 copies of c and z made once in forward, reused by backward
    l.cT = l.total;
    l.zT = l.cT + align_up((size_t)g.B * g.S * g.Har * 2);
    l.total = l.zT + align_up((size_t)g.B * g.S * g.H * 2);
  }
  if (g.dff > 0) {  // transformer prediction heads keep their activations here
    l.thead = l.total;
    l.total = l.thead + align_up(thead_save_bytes(g));
  }
  return l;
}

template <class T>
int criterion_fwd_t(const Geo& g, const float* c, const float* z, const float* w_pred, const cpcb200_thead_params* tp,
                    const int* ext, float* losses, float* acc, void* save, void* wsp, size_t ws_bytes, cudaStream_t st) {
  const int B = g.B, S = g.S, W = g.W, H = g.H, Har = g.Har, K = g.K, N = g.N;
  const int P = B * W;
  constexpr bool isf = sizeof(T) == 4;
  CritLayout lay = crit_layout(g);
  char* sv = static_cast<char*>(save);
  T* pred = reinterpret_cast<T*>(sv + lay.pred);
  float* logits = reinterpret_cast<float*>(sv + lay.logits);
  Carver ws(wsp, ws_bytes);
  T* cT = isf ? nullptr : reinterpret_cast<T*>(sv + lay.cT);
  T* zT = isf ? nullptr : reinterpret_cast<T*>(sv + lay.zT);
  T* wT = ws.take<T>((isf || tp) ? 1 : (size_t)K * H * Har);
  float* lossbuf = ws.take<float>((size_t)P * K);
  float* corrbuf = ws.take<float>((size_t)P * K);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "criterion_fwd: workspace %zu < %zu", ws_bytes, ws.off);
  const T *cp, *zp, *wp;
  if (isf) { cp = reinterpret_cast<const T*>(c); zp = reinterpret_cast<const T*>(z); wp = reinterpret_cast<const T*>(w_pred); }
  else {
    CPC_TRY(launch_cast3_bf16(c, reinterpret_cast<bf16*>(cT), (long long)B * S * Har, z, reinterpret_cast<bf16*>(zT),
                              (long long)B * S * H, tp ? nullptr : w_pred, reinterpret_cast<bf16*>(wT),
                              tp ? 0 : (long long)K * H * Har, st));
    cp = cT; zp = zT; wp = wT;
  }
  if (tp) {  // transformer heads (rnnMode='transformer')
    CPC_TRY(thead_fwd<T>(g, cp, tp, pred, sv + lay.thead, ws, st));
  } else {   // linear heads: pred[(b,w), (k,h)] = sum_a c[b,w,a] * Wk[h,a]
    RowView A{cp, (long long)S * Har, (long long)Har, W};
    OutView C{pred, (long long)W * K * H, (long long)K * H, W, 0, W, 0};
    CPC_TRY(gemm_nt(g.bf16, false, B, K * H, Har, A, wp, nullptr, C, st));
  }
  // scoring + CE
  if constexpr (!isf) {
    if (lay.mma) {
      float* lse = reinterpret_cast<float*>(sv + lay.lse);
      int* ext_t = reinterpret_cast<int*>(sv + lay.ext_t);
      CPC_TRY(score_transpose_ext(ext, ext_t, B, N, W, st));
      CPC_TRY(score_fwd_mma(pred, zp, ext_t, lossbuf, corrbuf, lse, B, S, W, H, K, N, st));
      CPC_CHECK_CUDA(launch_k(mean_over_positions_kernel, dim3(K), dim3(1024), 0, st, 1, lossbuf, corrbuf, losses, acc, P, K));
      CPC_LAUNCHED_N("mean_over_positions", st);
      return 0;
    }
  }
  int threads = ((N + K + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > 1024) return fail(CPCB200_ERR_UNSUPPORTED, "criterion: N + K = %d > 1024", N + K);
  const size_t smem = ((size_t)H * KMAX + (size_t)K * (N + 1)) * sizeof(float);
  CPC_CHECK_CUDA(cudaFuncSetAttribute(score_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  score_fwd_kernel<T><<<P, threads, smem, st>>>(pred, zp, ext, logits, lossbuf, corrbuf, B, S, W, H, K, N);
  CPC_LAUNCHED_N("score_fwd", st);
  mean_over_positions_kernel<<<K, 1024, 0, st>>>(lossbuf, corrbuf, losses, acc, P, K);
  CPC_LAUNCHED_N("mean_over_positions", st);
  return 0;
}

template <class T>
int criterion_bwd_t(const Geo& g, const float* c, const float* z, const float* w_pred, const cpcb200_thead_params* tp,
                    const int* ext, const float* dlosses, const void* save, float* dc, float* dz, float* dw_pred,
                    const cpcb200_thead_params* tg, void* wsp, size_t ws_bytes, cudaStream_t st) {
  const int B = g.B, S = g.S, W = g.W, H = g.H, Har = g.Har, K = g.K, N = g.N;
  const int P = B * W;
  constexpr bool isf = sizeof(T) == 4;
  CritLayout lay = crit_layout(g);
  const char* sv = static_cast<const char*>(save);
  const T* pred = reinterpret_cast<const T*>(sv + lay.pred);
  const float* logits = reinterpret_cast<const float*>(sv + lay.logits);
  Carver ws(wsp, ws_bytes);
  const T* cT = isf ? nullptr : reinterpret_cast<const T*>(sv + lay.cT);
  const T* zT = isf ? nullptr : reinterpret_cast<const T*>(sv + lay.zT);
  T* wTt = ws.take<T>(tp ? 1 : (size_t)K * H * Har);  // transposed heads: [Har][K*H]
  T* dpred = ws.take<T>((size_t)P * K * H);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "criterion_bwd: workspace %zu < %zu", ws_bytes, ws.off);
  const T *cp, *zp;
  if (isf) { cp = reinterpret_cast<const T*>(c); zp = reinterpret_cast<const T*>(z); }
  else {
    cp = cT; zp = zT;
  }
  if (!tp) CPC_TRY(launch_transpose_cast<T>(w_pred, wTt, K * H, Har, st));
  CPC_CHECK_CUDA(cudaMemsetAsync(dz, 0, (size_t)B * S * H * sizeof(float), st));
  CPC_CHECK_CUDA(cudaMemsetAsync(dc, 0, (size_t)B * S * Har * sizeof(float), st));
  bool done = false;
  if constexpr (!isf) {
    if (lay.mma) {
      const float* lse = reinterpret_cast<const float*>(sv + lay.lse);
      const int* ext_t = reinterpret_cast<const int*>(sv + lay.ext_t);
      CPC_TRY(score_bwd_mma(pred, zp, ext_t, lse, dlosses, dpred, dz, B, S, W, H, K, N, st));
      done = true;
    }
  }
  if (!done) {
    int threads = H < 1024 ? H : 1024;
    const size_t smem = (size_t)(N + 1) * KMAX * sizeof(float) + (size_t)N * sizeof(int);
    CPC_CHECK_CUDA(cudaFuncSetAttribute(score_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    score_bwd_kernel<T><<<P, threads, smem, st>>>(pred, zp, ext, logits, dlosses, dpred, dz, B, S, W, H, K, N);
    CPC_LAUNCHED_N("score_bwd", st);
  }
  if (tp) return thead_bwd<T>(g, cp, tp, dpred, sv + lay.thead, dc, tg, ws, st);
  {  // dW[(k,h)][a] = sum_p dpred[p][(k,h)] * c[p][a]
    RowView A{dpred, (long long)W * K * H, (long long)K * H, W};
    RowView Bv{cp, (long long)S * Har, (long long)Har, W};
    CPC_TRY(gemm_tn(g.bf16, B, K * H, Har, A, Bv, dw_pred, Har, STORE_PLAIN, 0, 0, st));
  }
  {  // dc[b,w,a] = sum_(k,h) dpred[p][(k,h)] * Wk[h][a]
    RowView A{dpred, (long long)W * K * H, (long long)K * H, W};
    OutView C{dc, (long long)S * Har, (long long)Har, W, 0, W, 0};
    CPC_TRY(gemm_nt(g.bf16, true, B, Har, K * H, A, wTt, nullptr, C, st));
  }
  return 0;
}

}  // namespace

int sample_ext_idx(const Geo& g, const int64_t* bi, const int64_t* si, int32_t* ext, cudaStream_t st) {
  const long long n = (long long)g.B * g.N * g.W;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPC_CHECK_CUDA(launch_k(ext_idx_kernel, dim3(blocks), dim3(256), 0, st, 1, reinterpret_cast<const long long*>(bi),
                          reinterpret_cast<const long long*>(si), ext, n, g.W, g.S));
  CPC_LAUNCHED_N("ext_idx", st);
  return 0;
}

size_t criterion_save_bytes(const Geo& g) { return crit_layout(g).total + 256; }

size_t criterion_ws_bytes(const Geo& g, int backward) {
  const size_t es = g.bf16 ? 2 : 4;
  const size_t P = (size_t)g.B * g.W;
  size_t tot = 0;
  if (!backward) {
    tot += align_up(g.bf16 ? (size_t)g.K * g.H * g.Har * es : 4);
    tot += 2 * align_up(P * g.K * 4);
  } else {
    tot += align_up((size_t)g.K * g.H * g.Har * es);
    tot += align_up(P * g.K * g.H * es);
  }
  if (g.dff > 0) tot += thead_ws_bytes(g, backward);
  return tot + 256;
}

int criterion_fwd(const Geo& g, const float* c, const float* z, const float* w_pred, const cpcb200_thead_params* tp,
                  const int* ext, float* losses, float* acc, void* save, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.K > KMAX) return fail(CPCB200_ERR_UNSUPPORTED, "criterion: K=%d > %d", g.K, KMAX);
  if (g.bf16) return criterion_fwd_t<bf16>(g, c, z, w_pred, tp, ext, losses, acc, save, ws, ws_bytes, st);
  return criterion_fwd_t<float>(g, c, z, w_pred, tp, ext, losses, acc, save, ws, ws_bytes, st);
}
int criterion_bwd(const Geo& g, const float* c, const float* z, const float* w_pred, const cpcb200_thead_params* tp,
                  const int* ext, const float* dlosses, const void* save, float* dc, float* dz, float* dw_pred,
                  const cpcb200_thead_params* tg, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.K > KMAX) return fail(CPCB200_ERR_UNSUPPORTED, "criterion: K=%d > %d", g.K, KMAX);
  if (g.bf16) return criterion_bwd_t<bf16>(g, c, z, w_pred, tp, ext, dlosses, save, dc, dz, dw_pred, tg, ws, ws_bytes, st);
  return criterion_bwd_t<float>(g, c, z, w_pred, tp, ext, dlosses, save, dc, dz, dw_pred, tg, ws, ws_bytes, st);
}

}  // namespace cpcb200
