// thead.cu - the K one-layer transformer prediction heads of rnnMode='transformer'
// (reference: cpc/criterion/criterion.py:82-88 -> cpc/transformers.py:10-139, nLayers = 1, abspos = False).
//
// Per head k (input x = c[:, :W], D = dmodel = H = Har, nh heads of dk = D/nh, F = dff):
//   q,k,v = x Wq^T, x Wk^T, x Wv^T                                   transformers.py:60-65,76-80   (GEMMs)
//   scores[i][c] = (q_i.k_c + q_i.Krelpos[:, W-1-(i-c)]) / sqrt(dk), c <= i   38-48 (relative-position "skew")
//   a = softmax(scores) ; att = dropout(a) v                         48-49 (train mode: caller-supplied keep masks)   attn kernels
//   y1 = LN(x + att Wo^T)                                            81-83, 109                       GEMM + add_ln
//   out = LN(y1 + relu(y1 W1^T + b1) W2^T + b2)                      86-95, 110-111                   GEMMs + add_ln
// Everything dense goes through gemm_nt / gemm_tn (tcgen05 on the bf16 path); attention runs on mma.sync (attn_mma.cu) on the
// bf16 path and on the CUDA-core kernels below otherwise; LayerNorm, the ReLU mask and dropout are row / element kernels (fp32
// math, T storage).  Backward recomputes the attention probabilities.  On the bf16 path the K heads run STACKED - batched
// GEMMs over stacked per-head weights, one launch per stage for all heads (thead_fwd_stacked / thead_bwd_stacked below); the
// per-head form on stream lanes remains for fp32.
#include "common.cuh"

namespace cpcb200 {

int gemm_nt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias,
            const OutView& C, cudaStream_t st);
int gemm_tn(bool bf16_in, int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc, int mode,
            int Ci, int taps, cudaStream_t st);
template <class T> int launch_cast(const float* src, T* dst, long long n, cudaStream_t st);
template <class T> int launch_transpose_cast(const float* src, T* dst, int R, int C, cudaStream_t st, int nb = 1);
template <class T> int launch_colsum(const T* src, float* out, long long M, int N, cudaStream_t st, int nb = 1);
int gemm_tn_group(bool bf16_in, int n, const TnDesc* d, cudaStream_t st);

bool attn_mma_supported(int W, int D, int nh);
int attn_fwd_mma(const bf16* qkv, const float* krel, bf16* att, int B, int W, int D, int nh, const unsigned char* keep, float dscale,
                 cudaStream_t st, int bph = 0);
int attn_bwd_mma(const bf16* qkv, const bf16* datt, const bf16* att, const float* krel, bf16* dqkv, float* dkrel, int B, int W, int D,
                 int nh, const unsigned char* keep, float dscale, cudaStream_t st, int bph = 0);

namespace {

constexpr float kLnEps = 1e-5f;

// ---------------------------------------------------------------------------------------------------------
// attention forward: one CTA per (head, window), thread i = query row i.
// qkv: (P, 3D) rows = (b, w); att: (P, D).
// ---------------------------------------------------------------------------------------------------------
// keep != NULL (train mode, transformers.py:18,49): probability (i, c) of (window b, head h) is multiplied by
// keep[((b*nh + h)*W + i)*W + c] * dscale before it meets V - the mask torch's nn.Dropout drew for that element.
//
// Thread i owns query row i.  Every inner loop walks an index that is THE SAME for all lanes of a warp (a key c, or a
// relative-position column m), so the K / V / Krelpos rows it needs are warp-wide broadcasts read as float4 - the skew of
// transformers.py:42-47 (key c <= query i reads column m = W-1-(i-c) of Q.Krelpos) is applied by first storing the row
// QP[i][m] = q_i . Krelpos[:, m] to shared memory and then adding QP[i][W-1-i+c] in the key loop.
template <int DK>
__device__ __forceinline__ float dot_bcast(const float (&q)[DK], const float* __restrict__ row) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent chains: the loop is FMA-latency-bound at 4 warps per SM
#pragma unroll
  for (int d4 = 0; d4 < DK / 4; d4++) {
    const float4 v = reinterpret_cast<const float4*>(row)[d4];
    s0 = fmaf(q[4 * d4], v.x, s0); s1 = fmaf(q[4 * d4 + 1], v.y, s1); s2 = fmaf(q[4 * d4 + 2], v.z, s2); s3 = fmaf(q[4 * d4 + 3], v.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
template <int DK>
__device__ __forceinline__ void axpy_bcast(float (&o)[DK], float a, const float* __restrict__ row) {
#pragma unroll
  for (int d4 = 0; d4 < DK / 4; d4++) {
    const float4 v = reinterpret_cast<const float4*>(row)[d4];
    o[4 * d4] = fmaf(a, v.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(a, v.y, o[4 * d4 + 1]);
    o[4 * d4 + 2] = fmaf(a, v.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(a, v.w, o[4 * d4 + 3]);
  }
}

template <class T, int DK>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ krel,
                                                        T* __restrict__ att, int W, int D, const unsigned char* __restrict__ keep,
                                                        float dscale) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                      // [W][DK]
  float* Vs = Ks + W * DK;             // [W][DK]
  float* Rt = Vs + W * DK;             // [W][DK]   Krelpos transposed: Rt[m][d] = Krelpos[d][m]
  float* Ps = Rt + W * DK;             // [W][W+1]  QP row, then scores / probabilities
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const T* base = qkv + (size_t)b * W * 3 * D + h * DK;
  for (int i = tid; i < W * DK; i += blockDim.x) {
    const int r = i / DK, d = i - r * DK;
    Ks[i] = to_f(base[(size_t)r * 3 * D + D + d]);
    Vs[i] = to_f(base[(size_t)r * 3 * D + 2 * D + d]);
  }
  for (int i = tid; i < DK * W; i += blockDim.x) { const int d = i / W, m = i - d * W; Rt[m * DK + d] = krel[i]; }
  __syncthreads();
  const int i = tid < W ? tid : W - 1;  // (threads past the last row shadow it: warp-uniform loops, no stores)
  const bool live = tid < W;
  float q[DK];
#pragma unroll
  for (int d = 0; d < DK; d++) q[d] = to_f(base[(size_t)i * 3 * D + d]);
  const float scale = rsqrtf((float)DK);
  float* prow = Ps + (size_t)i * (W + 1);
  const int i_hi = min(W - 1, (tid | 31));  // last row of this warp
  // pass A: QP[i][m] for the columns this row needs (m >= W-1-i); all lanes walk the same m
  for (int m = W - 1 - i_hi; m < W; m++) {
    const float v = dot_bcast<DK>(q, Rt + m * DK);
    if (live && m >= W - 1 - i) prow[m] = v;
  }
  __syncwarp();
  // pass B: scaled, skewed scores of keys c <= i (in place: column c <= m = W-1-i+c, so the QP entry of c is read
  // before anything overwrites it)
  float mx = -INFINITY;
  for (int c = 0; c <= i_hi; c++) {
    const float v = dot_bcast<DK>(q, Ks + c * DK);
    if (live && c <= i) {
      const float sc = (v + prow[W - 1 - i + c]) * scale;
      prow[c] = sc;
      mx = fmaxf(mx, sc);
    }
  }
  float sum = 0.f;
  if (live) for (int c = 0; c <= i; c++) { const float e = __expf(prow[c] - mx); prow[c] = e; sum += e; }
  const float inv = live ? 1.f / sum : 0.f;
  float o[DK];
#pragma unroll
  for (int d = 0; d < DK; d++) o[d] = 0.f;
  const unsigned char* krow = keep != nullptr ? keep + (((size_t)b * gridDim.x + h) * W + i) * W : nullptr;
  for (int c = 0; c <= i_hi; c++) {
    float p = (live && c <= i) ? prow[c] * inv : 0.f;
    if (krow != nullptr && live && c <= i) p = krow[c] ? p * dscale : 0.f;
    axpy_bcast<DK>(o, p, Vs + c * DK);
  }
  if (!live) return;
  T* orow = att + ((size_t)b * W + i) * D + h * DK;
#pragma unroll
  for (int d = 0; d < DK; d++) orow[d] = from_f<T>(o[d]);
}

// attention backward: recompute P, then dq (thread = query), dk/dv (thread = key), dKrelpos (thread = column m)
template <class T, int DK>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ datt,
                                                        const float* __restrict__ krel, T* __restrict__ dqkv,
                                                        float* __restrict__ dkrel, int W, int D,
                                                        const unsigned char* __restrict__ keep, float dscale) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                      // [W][DK]
  float* Ks = Qs + W * DK;
  float* Vs = Ks + W * DK;
  float* Gs = Vs + W * DK;             // dO
  float* Rt = Gs + W * DK;             // [W][DK]  Krelpos transposed
  float* Ps = Rt + W * DK;             // [W][W+1] probabilities (dropped ones after the query pass)
  float* Ss = Ps + W * (W + 1);        // [W][W+1] QP row, then d(scaled score)
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const T* base = qkv + (size_t)b * W * 3 * D + h * DK;
  const T* gbase = datt + (size_t)b * W * D + h * DK;
  for (int i = tid; i < W * DK; i += blockDim.x) {
    const int r = i / DK, d = i - r * DK;
    Qs[i] = to_f(base[(size_t)r * 3 * D + d]);
    Ks[i] = to_f(base[(size_t)r * 3 * D + D + d]);
    Vs[i] = to_f(base[(size_t)r * 3 * D + 2 * D + d]);
    Gs[i] = to_f(gbase[(size_t)r * D + d]);
  }
  for (int i = tid; i < DK * W; i += blockDim.x) { const int d = i / W, m = i - d * W; Rt[m * DK + d] = krel[i]; }
  for (int i = tid; i < W * (W + 1); i += blockDim.x) { Ps[i] = 0.f; Ss[i] = 0.f; }
  __syncthreads();
  const float scale = rsqrtf((float)DK);
  T* dbase = dqkv + (size_t)b * W * 3 * D + h * DK;
  {
    const int i = tid < W ? tid : W - 1;
    const bool live = tid < W;
    const int i_hi = min(W - 1, (tid | 31));
    // threads past the last row shadow row W-1 (warp-uniform loops): they only READ shared memory, every write below is
    // guarded by `live`
    float* prow = Ps + (size_t)i * (W + 1);
    float* srow = Ss + (size_t)i * (W + 1);
    float q[DK], go[DK];
#pragma unroll
    for (int d = 0; d < DK; d++) { q[d] = Qs[i * DK + d]; go[d] = Gs[i * DK + d]; }
    // QP[i][m] -> srow[m] (m >= W-1-i)
    for (int m = W - 1 - i_hi; m < W; m++) {
      const float v = dot_bcast<DK>(q, Rt + m * DK);
      if (live && m >= W - 1 - i) srow[m] = v;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int c = 0; c <= i_hi; c++) {
      const float v = dot_bcast<DK>(q, Ks + c * DK);
      if (c <= i) {
        const float sc = (v + srow[W - 1 - i + c]) * scale;
        if (live) prow[c] = sc;
        mx = fmaxf(mx, sc);
      }
    }
    __syncwarp();
    float sum = 0.f;
    if (live) for (int c = 0; c <= i; c++) { const float e = __expf(prow[c] - mx); prow[c] = e; sum += e; }
    const float inv = live ? 1.f / sum : 0.f;
    const unsigned char* krow = keep != nullptr ? keep + (((size_t)b * gridDim.x + h) * W + i) * W : nullptr;
    float dsum = 0.f;
    for (int c = 0; c <= i_hi; c++) {
      float dp = dot_bcast<DK>(go, Vs + c * DK);
      if (live && c <= i) {
        const float p = prow[c] * inv;
        prow[c] = p;
        if (krow != nullptr) dp = krow[c] ? dp * dscale : 0.f;  // gradient w.r.t. the probability before dropout
        srow[c] = dp;                                           // (the QP entries of this row were consumed by the score pass)
        dsum = fmaf(dp, p, dsum);
      }
    }
    // ds in place; probabilities become the DROPPED ones (dv below needs them)
    if (live) for (int c = 0; c <= i; c++) {
      const float ds = prow[c] * (srow[c] - dsum) * scale;
      srow[c] = ds;
      if (krow != nullptr) prow[c] = krow[c] ? prow[c] * dscale : 0.f;
    }
    __syncwarp();
    // dq = sum_c ds[c] K[c]  +  sum_m ds[c = m - (W-1-i)] Krelpos[:, m]   (both loops: warp-uniform index, broadcast rows)
    float dq[DK];
#pragma unroll
    for (int d = 0; d < DK; d++) dq[d] = 0.f;
    for (int c = 0; c <= i_hi; c++) {
      const float ds = (live && c <= i) ? srow[c] : 0.f;
      axpy_bcast<DK>(dq, ds, Ks + c * DK);
    }
    for (int m = W - 1 - i_hi; m < W; m++) {
      const int c = m - (W - 1 - i);
      const float ds = (live && c >= 0) ? srow[c] : 0.f;
      axpy_bcast<DK>(dq, ds, Rt + m * DK);
    }
    if (live) {
#pragma unroll
      for (int d = 0; d < DK; d++) dbase[(size_t)i * 3 * D + d] = from_f<T>(dq[d]);
    }
  }
  __syncthreads();
  if (tid < W) {
    const int c = tid;  // key index
    float dk[DK], dv[DK];
#pragma unroll
    for (int d = 0; d < DK; d++) { dk[d] = 0.f; dv[d] = 0.f; }
    const int c_lo = tid & ~31;  // rows i >= c_lo can touch a key of this warp: warp-uniform loop, zeros above the diagonal
    for (int i = c_lo; i < W; i++) {
      // (entries above the diagonal of Ss hold left-over QP values: masked here)
      const float ds = i >= c ? Ss[(size_t)i * (W + 1) + c] : 0.f, p = i >= c ? Ps[(size_t)i * (W + 1) + c] : 0.f;
      axpy_bcast<DK>(dk, ds, Qs + i * DK);
      axpy_bcast<DK>(dv, p, Gs + i * DK);
    }
#pragma unroll
    for (int d = 0; d < DK; d++) {
      dbase[(size_t)c * 3 * D + D + d] = from_f<T>(dk[d]);
      dbase[(size_t)c * 3 * D + 2 * D + d] = from_f<T>(dv[d]);
    }
    // dKrelpos[d][m] += sum_{i >= W-1-m} dS[i][i-(W-1-m)] * q_i[d]
    const int m = tid;
    float dr[DK];
#pragma unroll
    for (int d = 0; d < DK; d++) dr[d] = 0.f;
    const int m_hi = min(W - 1, (tid | 31));
    for (int i = W - 1 - m_hi; i < W; i++) {
      const int cc = i - (W - 1 - m);
      const float ds = cc >= 0 ? Ss[(size_t)i * (W + 1) + cc] : 0.f;
      axpy_bcast<DK>(dr, ds, Qs + i * DK);
    }
#pragma unroll
    for (int d = 0; d < DK; d++) atomicAdd(dkrel + d * W + m, dr[d]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// row kernels over D channels: one warp per row, lane owns float4 chunks at channel 4*(lane + 32*i)
// ---------------------------------------------------------------------------------------------------------
template <int I, class T>
__device__ __forceinline__ void rload(const T* row, int D, int lane, float (&v)[I][4]) {
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < D) load_vec<4>(row + c, v[i]);
    else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f; }
  }
}
template <int I, class T>
__device__ __forceinline__ void rstore(T* row, int D, int lane, const float (&v)[I][4]) {
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < D) store_vec<4>(row + c, v[i]);
  }
}
// biased-variance LayerNorm statistics (torch.nn.LayerNorm, eps 1e-5)
template <int I>
__device__ __forceinline__ void ln_stats(const float (&u)[I][4], int D, int lane, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < I; i++)
    if (4 * (lane + 32 * i) < D) s += (u[i][0] + u[i][1]) + (u[i][2] + u[i][3]);
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < I; i++)
    if (4 * (lane + 32 * i) < D) {
#pragma unroll
      for (int j = 0; j < 4; j++) { const float dl = u[i][j] - mean; q = fmaf(dl, dl, q); }
    }
  rstd = rsqrtf(warp_sum(q) / (float)D + kLnEps);
}

// s = a + b ; y = LN(s).  a: rows (p / rpb, p % rpb) with strides (a_bs, a_rs); b, s dense (P, D); y rows stride y_rs.
// Stacked heads (HS.rph > 0): P counts the rows of ALL heads, head = p / rph owns gamma / beta + head * D, its rows of a start
// at a + head * a_hs (a_hs = 0: the heads share a) and its rows of y at y + head * y_hs.
struct HeadStride { int rph = 0; long long a_hs = 0, y_hs = 0; };
template <int I, class TA, class T>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const TA* __restrict__ a, long long a_bs, long long a_rs, int rpb,
                                                          const T* __restrict__ b, const float* __restrict__ gam,
                                                          const float* __restrict__ bet, T* __restrict__ s_out,
                                                          T* __restrict__ y, long long y_rs, int P, int D, HeadStride HS) {
  const int lane = threadIdx.x & 31;
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= P) return;
  long long pl = p;  // row inside its head
  if (HS.rph > 0) {
    const int head = (int)(p / HS.rph);
    pl = p - (long long)head * HS.rph;
    a += head * HS.a_hs; y += head * HS.y_hs; gam += (size_t)head * D; bet += (size_t)head * D;
  }
  const int bi = (int)(pl / rpb), ti = (int)(pl - (long long)bi * rpb);
  float va[I][4], vb[I][4];
  rload<I>(a + (long long)bi * a_bs + (long long)ti * a_rs, D, lane, va);
  rload<I>(b + p * D, D, lane, vb);
#pragma unroll
  for (int i = 0; i < I; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) va[i][j] += vb[i][j];
  rstore<I>(s_out + p * D, D, lane, va);
  // statistics on the value that was stored (bf16-rounded on the bf16 path) so that backward sees the same xhat
  rload<I>(s_out + p * D, D, lane, va);
  float mean, rstd;
  ln_stats<I>(va, D, lane, mean, rstd);
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < D) {
      const float4 g4 = *reinterpret_cast<const float4*>(gam + c), b4 = *reinterpret_cast<const float4*>(bet + c);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; j++) va[i][j] = fmaf((va[i][j] - mean) * rstd, g[j], be[j]);
    }
  }
  rstore<I>(y + pl * y_rs, D, lane, va);
}

// LayerNorm backward: dy = dya (+ dyb); ds = rstd (dxh - mean(dxh) - xh mean(dxh xh)); dgamma += dy xh; dbeta += dy
// blockIdx.y = head (stacked heads: every per-head array advances by its head stride, dya by dya_hs).  dsum != NULL: the column
// sums of ds are added there too (the bias gradient of the linear layer that produced the LayerNorm input).
template <int I, class T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ dya, long long dya_rs, const T* __restrict__ dyb,
                                                      const T* __restrict__ s, const float* __restrict__ gam,
                                                      T* __restrict__ ds, float* __restrict__ dgam, float* __restrict__ dbet,
                                                      int P, int D, long long dya_hs, float* __restrict__ dsum) {
  extern __shared__ __align__(16) float accs[];  // [3][D]
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) accs[i] = 0.f;
  {
    const size_t head = blockIdx.y;
    dya += head * dya_hs; s += head * (size_t)P * D; ds += head * (size_t)P * D;
    if (dyb != nullptr) dyb += head * (size_t)P * D;
    gam += head * D; dgam += head * D; dbet += head * D;
    if (dsum != nullptr) dsum += head * D;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  float ag[I][4], ab[I][4], as[I][4];
#pragma unroll
  for (int i = 0; i < I; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) ag[i][j] = ab[i][j] = as[i][j] = 0.f;
  for (long long p = w0; p < P; p += nw) {
    float v[I][4], d[I][4];
    rload<I>(s + p * D, D, lane, v);
    rload<I>(dya + p * dya_rs, D, lane, d);
    if (dyb != nullptr) {
      float e[I][4];
      rload<I>(dyb + p * D, D, lane, e);
#pragma unroll
      for (int i = 0; i < I; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] += e[i][j];
    }
    float mean, rstd;
    ln_stats<I>(v, D, lane, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < I; i++) {
      const int c = 4 * (lane + 32 * i);
      if (c < D) {
        const float4 g4 = *reinterpret_cast<const float4*>(gam + c);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float xh = (v[i][j] - mean) * rstd;
          ag[i][j] = fmaf(d[i][j], xh, ag[i][j]);
          ab[i][j] += d[i][j];
          const float dx = d[i][j] * g[j];
          v[i][j] = xh; d[i][j] = dx;
          s1 += dx; s2 = fmaf(dx, xh, s2);
        }
      }
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
#pragma unroll
    for (int i = 0; i < I; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) d[i][j] = rstd * (d[i][j] - s1 - v[i][j] * s2);
    rstore<I>(ds + p * D, D, lane, d);
    if (dsum != nullptr) {  // (of the values as stored: what a separate column-sum pass over ds would read)
      rload<I>(ds + p * D, D, lane, d);
#pragma unroll
      for (int i = 0; i < I; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) as[i][j] += d[i][j];
    }
  }
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < D) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        atomicAdd(&accs[c + j], ag[i][j]); atomicAdd(&accs[D + c + j], ab[i][j]);
        if (dsum != nullptr) atomicAdd(&accs[2 * D + c + j], as[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgam + i, accs[i]); atomicAdd(dbet + i, accs[D + i]);
    if (dsum != nullptr) atomicAdd(dsum + i, accs[2 * D + i]);
  }
}

// h is the FFN hidden AFTER relu (and after dropout in train mode): h > 0 <=> the unit was active and kept, and the
// gradient through a kept unit carries the dropout scale (transformers.py:92-95)
template <class T>
__global__ void relu_mask_kernel(T* __restrict__ dh, const T* __restrict__ h, long long n, float dscale) {
  if constexpr (sizeof(T) == 2) {  // bf16: 8 elements (16 B) per thread and iteration
    const long long n8 = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
      uint4 dv = reinterpret_cast<const uint4*>(dh)[i];
      const uint4 hv = reinterpret_cast<const uint4*>(h)[i];
      __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(&dv);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hv);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float2 hf = __bfloat1622float2(h2[j]);
        float2 df = __bfloat1622float2(d2[j]);
        df.x = hf.x > 0.f ? df.x * dscale : 0.f;
        df.y = hf.y > 0.f ? df.y * dscale : 0.f;
        d2[j] = __floats2bfloat162_rn(df.x, df.y);
      }
      reinterpret_cast<uint4*>(dh)[i] = dv;
    }
    for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      dh[i] = from_f<T>(to_f(h[i]) > 0.f ? to_f(dh[i]) * dscale : 0.f);
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      if (!(to_f(h[i]) > 0.f)) dh[i] = from_f<T>(0.f);
      else if (dscale != 1.f) dh[i] = from_f<T>(to_f(dh[i]) * dscale);
    }
  }
}
// train-mode dropout of the FFN hidden (transformers.py:92): h *= keep * dscale, in place
template <class T>
__global__ void dropout_apply_kernel(T* __restrict__ h, const unsigned char* __restrict__ keep, long long n, float dscale) {
  if constexpr (sizeof(T) == 2) {
    const long long n8 = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
      uint4 hv = reinterpret_cast<const uint4*>(h)[i];
      const uint2 kv = reinterpret_cast<const uint2*>(keep)[i];
      const unsigned char* kb = reinterpret_cast<const unsigned char*>(&kv);
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hv);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float2 f = __bfloat1622float2(h2[j]);
        f.x = kb[2 * j] ? f.x * dscale : 0.f;
        f.y = kb[2 * j + 1] ? f.y * dscale : 0.f;
        h2[j] = __floats2bfloat162_rn(f.x, f.y);
      }
      reinterpret_cast<uint4*>(h)[i] = hv;
    }
    for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      h[i] = keep[i] ? from_f<T>(to_f(h[i]) * dscale) : from_f<T>(0.f);
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      h[i] = keep[i] ? from_f<T>(to_f(h[i]) * dscale) : from_f<T>(0.f);
  }
}
// acc (fp32, P x D) += a + b
template <class T>
__global__ void acc_add_kernel(float* __restrict__ acc, const T* __restrict__ a, const T* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc[i] += to_f(a[i]) + to_f(b[i]);
}
// dc[b, w < W, :] = sum over the lanes' accumulators acc[l][(b, w), :]
__global__ void scatter_rows_kernel(const float* __restrict__ acc, size_t lane_stride, int nlanes, float* __restrict__ dc, int B, int S,
                                    int W, int D) {
  const long long n = (long long)B * W * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D); const long long pw = i / D; const int w = (int)(pw % W), b = (int)(pw / W);
    float v = 0.f;
    for (int l = 0; l < nlanes; l++) v += acc[(size_t)l * lane_stride + i];
    dc[((long long)b * S + w) * D + d] = v;
  }
}
// stacked heads: dc[b, w < W, :] = sum_k (a[k][(b, w), :] + b[k][(b, w), :])
template <class T>
__global__ void sum_heads_scatter_kernel(const T* __restrict__ a, const T* __restrict__ bsrc, int K, float* __restrict__ dc, int B, int S,
                                         int W, int D) {
  const long long n = (long long)B * W * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D); const long long pw = i / D; const int w = (int)(pw % W), b = (int)(pw / W);
    float v = 0.f;
    for (int k = 0; k < K; k++) v += to_f(a[(size_t)k * n + i]) + to_f(bsrc[(size_t)k * n + i]);
    dc[((long long)b * S + w) * D + d] = v;
  }
}
// dst[d][j*D + r] = w_j[r][d]  for j in {q, k, v}: the (D, 3D) transposed concatenation used by d(x) = dqkv . [Wq;Wk;Wv]
template <class T>
__global__ void concat_transpose3_kernel(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ wv,
                                         T* __restrict__ dst, int D) {
  const long long n = (long long)3 * D * D;
  wq += (size_t)blockIdx.z * D * D; wk += (size_t)blockIdx.z * D * D; wv += (size_t)blockIdx.z * D * D;  // head blockIdx.z
  dst += (size_t)blockIdx.z * 3 * D * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (3 * D)), d = (int)(i / (3 * D));
    const int j = col / D, r = col - j * D;
    const float* w = j == 0 ? wq : (j == 1 ? wk : wv);
    dst[i] = from_f<T>(w[(size_t)r * D + d]);
  }
}

inline int grid_for(long long n) { long long b = (n + 255) / 256; return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b)); }

template <class T> size_t attn_fwd_smem(int W, int DK) { return (size_t)(3 * W * DK + W * (W + 1)) * 4; }
template <class T> size_t attn_bwd_smem(int W, int DK) { return (size_t)(5 * W * DK + 2 * W * (W + 1)) * 4; }

template <class T>
int launch_attn_fwd(const T* qkv, const float* krel, T* att, int B, int W, int D, int nh, const unsigned char* keep, float dscale,
                    cudaStream_t st) {
  const int DK = D / nh;
  if constexpr (sizeof(T) == 2) {  // bf16 path, dk = 32: tensor-core attention (attn_mma.cu)
    if (attn_mma_supported(W, D, nh)) return attn_fwd_mma(qkv, krel, att, B, W, D, nh, keep, dscale, st);
  }
  const size_t smem = attn_fwd_smem<T>(W, DK);
  dim3 grid(nh, B);
#define AF(DKV)                                                                                                   \
  {                                                                                                               \
    CPC_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T, DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_fwd_kernel<T, DKV><<<grid, 128, smem, st>>>(qkv, krel, att, W, D, keep, dscale);                                        \
  }
  if (DK == 32) AF(32) else if (DK == 8) AF(8) else if (DK == 16) AF(16) else if (DK == 64) AF(64)
  else return fail(CPCB200_ERR_UNSUPPORTED, "transformer heads: dk=%d", DK);
#undef AF
  CPC_LAUNCHED_N("attn_fwd", st);
  return 0;
}
template <class T>
int launch_attn_bwd(const T* qkv, const T* datt, const T* att, const float* krel, T* dqkv, float* dkrel, int B, int W, int D, int nh,
                    const unsigned char* keep, float dscale, cudaStream_t st) {
  const int DK = D / nh;
  if constexpr (sizeof(T) == 2) {
    if (attn_mma_supported(W, D, nh)) return attn_bwd_mma(qkv, datt, att, krel, dqkv, dkrel, B, W, D, nh, keep, dscale, st);
  }
  const size_t smem = attn_bwd_smem<T>(W, DK);
  dim3 grid(nh, B);
#define AB(DKV)                                                                                                   \
  {                                                                                                               \
    CPC_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<T, DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    attn_bwd_kernel<T, DKV><<<grid, 128, smem, st>>>(qkv, datt, krel, dqkv, dkrel, W, D, keep, dscale);                          \
  }
  if (DK == 32) AB(32) else if (DK == 8) AB(8) else if (DK == 16) AB(16) else if (DK == 64) AB(64)
  else return fail(CPCB200_ERR_UNSUPPORTED, "transformer heads: dk=%d", DK);
#undef AB
  CPC_LAUNCHED_N("attn_bwd", st);
  return 0;
}

template <class TA, class T>
int launch_add_ln(const TA* a, long long a_bs, long long a_rs, int rpb, const T* b, const float* gam, const float* bet, T* s_out,
                  T* y, long long y_rs, int P, int D, cudaStream_t st, const HeadStride& HS = HeadStride{}) {
  const int I = (D + 127) / 128;
  const int blocks = (int)(((long long)P * 32 + 255) / 256);
#define AL(II) add_ln_fwd_kernel<II, TA, T><<<blocks, 256, 0, st>>>(a, a_bs, a_rs, rpb, b, gam, bet, s_out, y, y_rs, P, D, HS)
  if (I == 1) AL(1); else if (I == 2) AL(2); else if (I == 3) AL(3); else AL(4);
#undef AL
  CPC_LAUNCHED_N("add_ln_fwd", st);
  return 0;
}
template <class T>
int launch_ln_bwd(const T* dya, long long dya_rs, const T* dyb, const T* s, const float* gam, T* ds, float* dgam, float* dbet, int P,
                  int D, cudaStream_t st, int heads = 1, long long dya_hs = 0, float* dsum = nullptr) {
  const int I = (D + 127) / 128;
  int blocks = (int)(((long long)P * 32 + 255) / 256);
  const int cap = heads > 1 ? (148 * 8 + heads - 1) / heads : 148 * 4;
  if (blocks > cap) blocks = cap;
  const size_t smem = 3 * (size_t)D * 4;
#define LB(II) ln_bwd_kernel<II, T><<<dim3(blocks, heads), 256, smem, st>>>(dya, dya_rs, dyb, s, gam, ds, dgam, dbet, P, D, dya_hs, dsum)
  if (I == 1) LB(1); else if (I == 2) LB(2); else if (I == 3) LB(3); else LB(4);
#undef LB
  CPC_LAUNCHED_N("ln_bwd", st);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Head-level concurrency.  The K heads are independent until their gradients meet in dc, and one head's kernels are small
// (a (B*W) x 256 x 256 product is 29 output tiles on 148 SMs): the heads are dealt round-robin onto kLanes streams - the
// caller's stream plus library-owned side streams forked from it and joined back before the call returns, so the caller
// still sees plain stream order and a CUDA-graph capture records the lanes as parallel branches.  Every lane has its own
// scratch (converted weights, intermediates, dc accumulator).  Per-kernel profiling (prof_enabled) runs single-lane.
// ---------------------------------------------------------------------------------------------------------
constexpr int kLanes = 4;
struct Lanes {
  cudaStream_t st[kLanes];
  int n;
};
int lanes_fork(cudaStream_t main, int want, Lanes* L) {
  static thread_local cudaStream_t side[kLanes - 1] = {nullptr, nullptr, nullptr};
  static thread_local cudaEvent_t ev_fork = nullptr;
  static thread_local int dev_of = -1;
  int dev = 0;
  CPC_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev_of != dev) {  // (one set per host thread and device; a thread that changes device gets fresh ones)
    for (int i = 0; i < kLanes - 1; i++) CPC_CHECK_CUDA(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
    CPC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    dev_of = dev;
  }
  static const int env_lanes = []() { const char* e = getenv("CPC_B200_HEAD_LANES"); int v = e ? atoi(e) : kLanes; return v < 1 ? 1 : (v > kLanes ? kLanes : v); }();
  int n = want < env_lanes ? want : env_lanes;
  if (prof_enabled() || n < 1) n = 1;
  L->n = n;
  L->st[0] = main;
  if (n > 1) {
    CPC_CHECK_CUDA(cudaEventRecord(ev_fork, main));
    for (int i = 1; i < n; i++) { L->st[i] = side[i - 1]; CPC_CHECK_CUDA(cudaStreamWaitEvent(side[i - 1], ev_fork, 0)); }
  }
  return 0;
}
int lanes_join(const Lanes& L) {
  static thread_local cudaEvent_t ev_join[kLanes - 1] = {nullptr, nullptr, nullptr};
  for (int i = 1; i < L.n; i++) {
    if (ev_join[i - 1] == nullptr) CPC_CHECK_CUDA(cudaEventCreateWithFlags(&ev_join[i - 1], cudaEventDisableTiming));
    CPC_CHECK_CUDA(cudaEventRecord(ev_join[i - 1], L.st[i]));
    CPC_CHECK_CUDA(cudaStreamWaitEvent(L.st[0], ev_join[i - 1], 0));
  }
  return 0;
}

struct THeadLayout { size_t qkv, att, s1, y1, h, s2, per_k; };  // element offsets (units of T) inside one head's block
THeadLayout thead_layout(const Geo& g) {
  THeadLayout l{};
  const size_t P = (size_t)g.B * g.W, D = g.H, F = g.dff;
  size_t off = 0;
  auto take = [&](size_t n) { size_t r = off; off += (n + 127) / 128 * 128; return r; };
  l.qkv = take(P * 3 * D); l.att = take(P * D); l.s1 = take(P * D); l.y1 = take(P * D); l.h = take(P * F); l.s2 = take(P * D);
  l.per_k = off;
  return l;
}


// ---------------------------------------------------------------------------------------------------------
// Stacked heads (bf16 path): the K heads run as ONE sequence of launches.  Every activation is an array over (head, row):
// [K][P][.] - so that a per-head product is one batch of a batched GEMM whose weights are the heads' stacked matrices
// (gemm_nt_heads), a per-head weight gradient is one problem of a grouped TN launch, and the row / element kernels walk
// K*P rows with the head's parameters looked up from the row index.  ~30 launches per direction instead of ~9 per head
// and direction on four stream lanes; a (B*W) x 256 x 256 product becomes 12 x 29 = 348 output tiles instead of 29.
// CPC_B200_HEAD_STACK=0 returns to the per-head lanes (also used by the fp32 path and when a product does not fit the
// 128 x 256 tensor-core tiling).
// ---------------------------------------------------------------------------------------------------------
struct StackLayout { size_t qkv, att, s1, y1, h, s2; };  // element offsets of the [K][P][.] arrays in the save buffer
StackLayout stack_layout(const Geo& g) {
  const size_t KP = (size_t)g.K * g.B * g.W, D = g.H, F = g.dff;
  StackLayout l{};
  size_t off = 0;
  auto take = [&](size_t n) { size_t r = off; off += (n + 127) / 128 * 128; return r; };
  l.qkv = take(KP * 3 * D); l.att = take(KP * D); l.s1 = take(KP * D); l.y1 = take(KP * D); l.h = take(KP * F); l.s2 = take(KP * D);
  return l;  // (never more than K * thead_layout(g).per_k elements: the per-head layout rounds every array of every head up)
}
bool heads_stacked(const Geo& g) {
  static const bool off = []() { const char* e = getenv("CPC_B200_HEAD_STACK"); return e && atoi(e) == 0; }();
  return !off && g.bf16 && g.H % 256 == 0 && g.dff % 256 == 0 && g.H == g.Har && attn_mma_supported(g.W, g.H, g.nheads);
}

int thead_fwd_stacked(const Geo& g, const bf16* cp, const cpcb200_thead_params* tp, bf16* pred, void* save, Carver& ws, cudaStream_t st) {
  typedef bf16 T;
  const int B = g.B, S = g.S, W = g.W, D = g.H, K = g.K, F = g.dff, nh = g.nheads;
  const int P = B * W;
  const long long KP = (long long)K * P;
  const StackLayout lay = stack_layout(g);
  T* sv = static_cast<T*>(save);
  T *qkv = sv + lay.qkv, *att = sv + lay.att, *s1 = sv + lay.s1, *y1 = sv + lay.y1, *h = sv + lay.h, *s2 = sv + lay.s2;
  T* o = ws.take<T>((size_t)KP * D);
  T* f = ws.take<T>((size_t)KP * D);
  T* wall[6] = {nullptr};
  const size_t wsz[6] = {(size_t)D * D, (size_t)D * D, (size_t)D * D, (size_t)D * D, (size_t)F * D, (size_t)F * D};
  for (int j = 0; j < 6; j++) wall[j] = ws.take<T>(wsz[j] * K);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "thead_fwd: workspace too small (%zu needed)", ws.off);
  const float* src[6] = {tp->wq, tp->wk, tp->wv, tp->wo, tp->w1, tp->w2};
  for (int j = 0; j < 6; j++) CPC_TRY(launch_cast<T>(src[j], wall[j], (long long)(wsz[j] * K), st));
  const bool drop = tp->att_keep != nullptr && tp->ffn_keep != nullptr;
  // q | k | v: batch (head, window) reads window b of x and the head's matrix
  const RowView X{cp, (long long)S * D, (long long)D, W};
  for (int j = 0; j < 3; j++) {
    OutView C{qkv + (size_t)j * D, (long long)W * 3 * D, (long long)3 * D, W, 0, W, 0};
    CPC_TRY(gemm_nt_heads(false, K * B, D, D, X, wall[j], nullptr, C, B, B, st));
  }
  CPC_TRY(attn_fwd_mma(qkv, tp->krelpos, att, K * B, W, D, nh, drop ? tp->att_keep : nullptr, tp->keep_scale, st, B));
  {
    RowView A{att, (long long)P * D, (long long)D, P};
    OutView C{o, (long long)P * D, (long long)D, P, 0, P, 0};
    CPC_TRY(gemm_nt_heads(false, K, D, D, A, wall[3], nullptr, C, 0, 1, st));
  }
  {
    HeadStride HS; HS.rph = P; HS.a_hs = 0; HS.y_hs = (long long)P * D;
    CPC_TRY((launch_add_ln<T, T>(cp, (long long)S * D, (long long)D, W, o, tp->ln1_w, tp->ln1_b, s1, y1, (long long)D, (int)KP, D, st, HS)));
  }
  {
    RowView A{y1, (long long)P * D, (long long)D, P};
    OutView C{h, (long long)P * F, (long long)F, P, 0, P, 0};
    C.relu = 1;
    CPC_TRY(gemm_nt_heads(false, K, F, D, A, wall[4], tp->b1, C, 0, 1, st));
    if (drop) {
      dropout_apply_kernel<T><<<grid_for(KP * F), 256, 0, st>>>(h, tp->ffn_keep, KP * F, tp->keep_scale);
      CPC_LAUNCHED_N("dropout_apply", st);
    }
  }
  {
    RowView A{h, (long long)P * F, (long long)F, P};
    OutView C{f, (long long)P * D, (long long)D, P, 0, P, 0};
    CPC_TRY(gemm_nt_heads(false, K, D, F, A, wall[5], tp->b2, C, 0, 1, st));
  }
  {  // pred[(b, w), k, :] = LN2(y1 + f) of head k
    HeadStride HS; HS.rph = P; HS.a_hs = (long long)P * D; HS.y_hs = (long long)D;
    CPC_TRY((launch_add_ln<T, T>(y1, (long long)P * D, (long long)D, P, f, tp->ln2_w, tp->ln2_b, s2, pred, (long long)K * D, (int)KP, D, st, HS)));
  }
  return 0;
}

int thead_bwd_stacked(const Geo& g, const bf16* cp, const cpcb200_thead_params* tp, const bf16* dpred, const void* save, float* dc,
                      const cpcb200_thead_params* gr, Carver& ws, cudaStream_t st) {
  typedef bf16 T;
  const int B = g.B, S = g.S, W = g.W, D = g.H, K = g.K, F = g.dff, nh = g.nheads;
  const int P = B * W, DK = D / nh;
  const long long KP = (long long)K * P;
  const StackLayout lay = stack_layout(g);
  const T* sv = static_cast<const T*>(save);
  const T *qkv = sv + lay.qkv, *att = sv + lay.att, *s1 = sv + lay.s1, *y1 = sv + lay.y1, *h = sv + lay.h, *s2 = sv + lay.s2;
  T* wqkvT = ws.take<T>((size_t)3 * D * D * K); T* woT = ws.take<T>((size_t)D * D * K);
  T* w1T = ws.take<T>((size_t)F * D * K); T* w2T = ws.take<T>((size_t)F * D * K);
  T* ds2 = ws.take<T>((size_t)KP * D); T* dy1 = ws.take<T>((size_t)KP * D); T* ds1 = ws.take<T>((size_t)KP * D);
  T* datt = ws.take<T>((size_t)KP * D); T* dx = ws.take<T>((size_t)KP * D);
  T* dh = ws.take<T>((size_t)KP * F); T* dqkv = ws.take<T>((size_t)KP * 3 * D);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "thead_bwd: workspace too small (%zu needed)", ws.off);
  concat_transpose3_kernel<T><<<dim3(grid_for((long long)3 * D * D), 1, K), 256, 0, st>>>(tp->wq, tp->wk, tp->wv, wqkvT, D);
  CPC_LAUNCHED_N("concat_transpose3", st);
  CPC_TRY(launch_transpose_cast<T>(tp->wo, woT, D, D, st, K));
  CPC_TRY(launch_transpose_cast<T>(tp->w1, w1T, F, D, st, K));
  CPC_TRY(launch_transpose_cast<T>(tp->w2, w2T, D, F, st, K));
  const bool drop = tp->att_keep != nullptr && tp->ffn_keep != nullptr;
  const size_t PD = (size_t)P * D, PF = (size_t)P * F;
  // per-head weight gradients: four problems per grouped TN launch
  auto tn_heads = [&](const T* a, size_t a_hs, int N1, const T* b, size_t b_hs, int N2, float* dw) -> int {
    TnDesc d[4];
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int n = K - k0 < 4 ? K - k0 : 4;
      for (int i = 0; i < n; i++) {
        const int k = k0 + i;
        d[i] = TnDesc{1, N1, N2, RowView{a + k * a_hs, 0, (long long)N1, P}, RowView{b + k * b_hs, 0, (long long)N2, P},
                      dw + (size_t)k * N1 * N2, N2, STORE_PLAIN, 0, 0};
      }
      CPC_TRY(gemm_tn_group(true, n, d, st));
    }
    return 0;
  };
  // LN2 backward (+ db2 = column sums of ds2)
  CPC_TRY(launch_ln_bwd<T>(dpred, (long long)K * D, nullptr, s2, tp->ln2_w, ds2, gr->ln2_w, gr->ln2_b, P, D, st, K, (long long)D, gr->b2));
  // FFN backward
  CPC_TRY(tn_heads(ds2, PD, D, h, PF, F, gr->w2));                                            // dW2[d][f]
  {
    RowView A{ds2, (long long)PD, (long long)D, P};
    OutView C{dh, (long long)PF, (long long)F, P, 0, P, 0};
    CPC_TRY(gemm_nt_heads(false, K, F, D, A, w2T, nullptr, C, 0, 1, st));                     // dh = ds2 . W2
  }
  relu_mask_kernel<T><<<grid_for(KP * F), 256, 0, st>>>(dh, h, KP * F, drop ? tp->keep_scale : 1.f);
  CPC_LAUNCHED_N("relu_mask", st);
  CPC_TRY(launch_colsum<T>(dh, gr->b1, P, F, st, K));
  CPC_TRY(tn_heads(dh, PF, F, y1, PD, D, gr->w1));                                            // dW1[f][d]
  {
    RowView A{dh, (long long)PF, (long long)F, P};
    OutView C{dy1, (long long)PD, (long long)D, P, 0, P, 0};
    CPC_TRY(gemm_nt_heads(false, K, D, F, A, w1T, nullptr, C, 0, 1, st));                     // dy1 (FFN branch)
  }
  // LN1 backward on dy1 + ds2 (residual)
  CPC_TRY(launch_ln_bwd<T>(dy1, (long long)D, ds2, s1, tp->ln1_w, ds1, gr->ln1_w, gr->ln1_b, P, D, st, K, (long long)PD, nullptr));
  CPC_TRY(tn_heads(ds1, PD, D, att, PD, D, gr->wo));
  {
    RowView A{ds1, (long long)PD, (long long)D, P};
    OutView C{datt, (long long)PD, (long long)D, P, 0, P, 0};
    CPC_TRY(gemm_nt_heads(false, K, D, D, A, woT, nullptr, C, 0, 1, st));
  }
  CPC_TRY(attn_bwd_mma(qkv, datt, att, tp->krelpos, dqkv, gr->krelpos, K * B, W, D, nh, drop ? tp->att_keep : nullptr, tp->keep_scale, st, B));
  (void)DK;
  {  // Wq, Wk, Wv: 3 K problems, each a sum over the B windows
    const RowView X{cp, (long long)S * D, (long long)D, W};
    float* dw3[3] = {gr->wq, gr->wk, gr->wv};
    TnDesc d[4];
    int n = 0;
    for (int k = 0; k < K; k++)
      for (int j = 0; j < 3; j++) {
        d[n++] = TnDesc{B, D, D, RowView{dqkv + (size_t)k * P * 3 * D + (size_t)j * D, (long long)W * 3 * D, (long long)3 * D, W}, X,
                        dw3[j] + (size_t)k * D * D, D, STORE_PLAIN, 0, 0};
        if (n == 4 || (k == K - 1 && j == 2)) { CPC_TRY(gemm_tn_group(true, n, d, st)); n = 0; }
      }
  }
  {
    RowView A{dqkv, (long long)P * 3 * D, (long long)3 * D, P};
    OutView C{dx, (long long)PD, (long long)D, P, 0, P, 0};
    CPC_TRY(gemm_nt_heads(false, K, D, 3 * D, A, wqkvT, nullptr, C, 0, 1, st));
  }
  sum_heads_scatter_kernel<T><<<grid_for((long long)P * D), 256, 0, st>>>(dx, ds1, K, dc, B, S, W, D);
  CPC_LAUNCHED_N("sum_heads_scatter", st);
  return 0;
}

}  // namespace

size_t thead_save_bytes(const Geo& g) { return thead_layout(g).per_k * g.K * (g.bf16 ? 2 : 4) + 256; }

size_t thead_ws_bytes(const Geo& g, int backward) {
  const size_t es = g.bf16 ? 2 : 4, P = (size_t)g.B * g.W, D = g.H, F = g.dff;
  size_t t = 0, w = 0;
  const size_t K = (size_t)g.K;
  if (!backward) {
    w += 4 * align_up(D * D * es * K) + 2 * align_up(F * D * es * K);   // Wq, Wk, Wv, Wo, W1, W2 of all heads in T
    t += 2 * align_up(P * D * es);                                      // o, f
  } else {
    w += align_up(3 * D * D * es * K) + align_up(D * D * es * K) + 2 * align_up(F * D * es * K);   // transposed weights, all heads
    t += 5 * align_up(P * D * es) + align_up(P * F * es) + align_up(P * 3 * D * es);   // ds2, dy1, ds1, datt, dx, dh, dqkv
    t += align_up(P * D * 4);                                                          // dc accumulator
  }
  size_t lanes = t * kLanes + w + 1024;  // one scratch set per lane + the converted weights
  if (heads_stacked(g)) {               // stacked heads: one scratch set over all K heads (no dc accumulators)
    size_t ts = backward ? 5 * align_up(K * P * D * es) + align_up(K * P * F * es) + align_up(K * P * 3 * D * es) : 2 * align_up(K * P * D * es);
    if (ts + w + 1024 > lanes) lanes = ts + w + 1024;
  }
  return lanes;
}

template <class T>
int thead_fwd(const Geo& g, const T* cp, const cpcb200_thead_params* tp, T* pred, void* save, Carver& ws, cudaStream_t st) {
  const int B = g.B, S = g.S, W = g.W, D = g.H, K = g.K, F = g.dff, nh = g.nheads;
  const int P = B * W;
  if (g.Har != g.H) return fail(CPCB200_ERR_UNSUPPORTED, "transformer heads need hiddenGar == hiddenEncoder (criterion.py:85)");
  if (W > 128 || D % nh != 0) return fail(CPCB200_ERR_UNSUPPORTED, "transformer heads: W=%d (max 128), D=%d, heads=%d", W, D, nh);
  constexpr bool isf = sizeof(T) == 4;
  if constexpr (!isf) {
    if (heads_stacked(g)) return thead_fwd_stacked(g, cp, tp, pred, save, ws, st);
  }
  const THeadLayout lay = thead_layout(g);
  T* sv = static_cast<T*>(save);
  struct FwdScratch { T *o, *f; } scr[kLanes];
  for (int l = 0; l < kLanes; l++) { scr[l].o = ws.take<T>((size_t)P * D); scr[l].f = ws.take<T>((size_t)P * D); }
  // the six weight matrices of ALL heads in the storage type: one conversion launch per parameter (K stacked matrices)
  T* wall[6] = {nullptr};
  const size_t wsz[6] = {(size_t)D * D, (size_t)D * D, (size_t)D * D, (size_t)D * D, (size_t)F * D, (size_t)F * D};
  if (!isf) for (int j = 0; j < 6; j++) wall[j] = ws.take<T>(wsz[j] * K);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "thead_fwd: workspace too small (%zu needed)", ws.off);
  if (!isf) {
    const float* src[6] = {tp->wq, tp->wk, tp->wv, tp->wo, tp->w1, tp->w2};
    for (int j = 0; j < 6; j++) CPC_TRY(launch_cast<T>(src[j], wall[j], (long long)(wsz[j] * K), st));
  }
  const RowView X{cp, (long long)S * D, (long long)D, W};
  const cudaStream_t st_main = st;
  Lanes lanes;
  CPC_TRY(lanes_fork(st_main, K, &lanes));
  for (int k = 0; k < K; k++) {
    const FwdScratch& sc = scr[k % lanes.n];
    st = lanes.st[k % lanes.n];
    T *o = sc.o, *f = sc.f;
    T* blk = sv + (size_t)k * lay.per_k;
    T *qkv = blk + lay.qkv, *att = blk + lay.att, *s1 = blk + lay.s1, *y1 = blk + lay.y1, *h = blk + lay.h, *s2 = blk + lay.s2;
    const T *Wq, *Wk, *Wv, *Wo, *W1, *W2;
    if (isf) {
      Wq = reinterpret_cast<const T*>(tp->wq + (size_t)k * D * D); Wk = reinterpret_cast<const T*>(tp->wk + (size_t)k * D * D);
      Wv = reinterpret_cast<const T*>(tp->wv + (size_t)k * D * D); Wo = reinterpret_cast<const T*>(tp->wo + (size_t)k * D * D);
      W1 = reinterpret_cast<const T*>(tp->w1 + (size_t)k * F * D); W2 = reinterpret_cast<const T*>(tp->w2 + (size_t)k * D * F);
    } else {
      Wq = wall[0] + (size_t)k * D * D; Wk = wall[1] + (size_t)k * D * D; Wv = wall[2] + (size_t)k * D * D;
      Wo = wall[3] + (size_t)k * D * D; W1 = wall[4] + (size_t)k * F * D; W2 = wall[5] + (size_t)k * D * F;
    }
    const T* Wqkv[3] = {Wq, Wk, Wv};
    for (int j = 0; j < 3; j++) {  // q | k | v column blocks of qkv
      OutView C{qkv + (size_t)j * D, (long long)W * 3 * D, (long long)3 * D, W, 0, W, 0};
      CPC_TRY(gemm_nt(g.bf16, false, B, D, D, X, Wqkv[j], nullptr, C, st));
    }
    const bool drop = tp->att_keep != nullptr && tp->ffn_keep != nullptr;
    CPC_TRY(launch_attn_fwd<T>(qkv, tp->krelpos + (size_t)k * (D / nh) * W, att, B, W, D, nh,
                               drop ? tp->att_keep + (size_t)k * B * nh * W * W : nullptr, tp->keep_scale, st));
    {
      RowView A{att, 0, (long long)D, P};
      OutView C{o, 0, (long long)D, P, 0, P, 0};
      CPC_TRY(gemm_nt(g.bf16, false, 1, D, D, A, Wo, nullptr, C, st));
    }
    CPC_TRY((launch_add_ln<T, T>(cp, (long long)S * D, (long long)D, W, o, tp->ln1_w + (size_t)k * D, tp->ln1_b + (size_t)k * D, s1, y1,
                                 (long long)D, P, D, st)));
    {
      RowView A{y1, 0, (long long)D, P};
      OutView C{h, 0, (long long)F, P, 0, P, 0};
      C.relu = 1;
      CPC_TRY(gemm_nt(g.bf16, false, 1, F, D, A, W1, tp->b1 + (size_t)k * F, C, st));
      if (drop) {
        dropout_apply_kernel<T><<<grid_for((long long)P * F), 256, 0, st>>>(h, tp->ffn_keep + (size_t)k * P * F, (long long)P * F, tp->keep_scale);
        CPC_LAUNCHED_N("dropout_apply", st);
      }
    }
    {
      RowView A{h, 0, (long long)F, P};
      OutView C{f, 0, (long long)D, P, 0, P, 0};
      CPC_TRY(gemm_nt(g.bf16, false, 1, D, F, A, W2, tp->b2 + (size_t)k * D, C, st));
    }
    CPC_TRY((launch_add_ln<T, T>(y1, (long long)P * D, (long long)D, P, f, tp->ln2_w + (size_t)k * D, tp->ln2_b + (size_t)k * D, s2,
                                 pred + (size_t)k * D, (long long)K * D, P, D, st)));
  }
  CPC_TRY(lanes_join(lanes));
  return 0;
}

template <class T>
int thead_bwd(const Geo& g, const T* cp, const cpcb200_thead_params* tp, const T* dpred, const void* save, float* dc,
              const cpcb200_thead_params* gr, Carver& ws, cudaStream_t st) {
  const int B = g.B, S = g.S, W = g.W, D = g.H, K = g.K, F = g.dff, nh = g.nheads;
  const int P = B * W, DK = D / nh;
  if constexpr (sizeof(T) == 2) {
    if (heads_stacked(g)) return thead_bwd_stacked(g, cp, tp, dpred, save, dc, gr, ws, st);
  }
  const THeadLayout lay = thead_layout(g);
  const T* sv = static_cast<const T*>(save);
  struct BwdScratch { T *ds2, *dy1, *ds1, *datt, *dx, *dh, *dqkv; float* dcw; } scr[kLanes];
  // transposed weights of ALL heads: one launch per parameter (batched over the K stacked matrices)
  T* wqkvT_all = ws.take<T>((size_t)3 * D * D * K); T* woT_all = ws.take<T>((size_t)D * D * K);
  T* w1T_all = ws.take<T>((size_t)F * D * K); T* w2T_all = ws.take<T>((size_t)F * D * K);
  for (int l = 0; l < kLanes; l++) {
    scr[l].ds2 = ws.take<T>((size_t)P * D); scr[l].dy1 = ws.take<T>((size_t)P * D); scr[l].ds1 = ws.take<T>((size_t)P * D);
    scr[l].datt = ws.take<T>((size_t)P * D); scr[l].dx = ws.take<T>((size_t)P * D);
    scr[l].dh = ws.take<T>((size_t)P * F); scr[l].dqkv = ws.take<T>((size_t)P * 3 * D);
  }
  for (int l = 0; l < kLanes; l++) scr[l].dcw = ws.take<float>((size_t)P * D);  // contiguous: cleared with one memset
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "thead_bwd: workspace too small (%zu needed)", ws.off);
  const cudaStream_t st_main = st;
  CPC_CHECK_CUDA(cudaMemsetAsync(scr[0].dcw, 0, (size_t)((char*)scr[kLanes - 1].dcw - (char*)scr[0].dcw) + (size_t)P * D * 4, st_main));
  const RowView X{cp, (long long)S * D, (long long)D, W};
  concat_transpose3_kernel<T><<<dim3(grid_for((long long)3 * D * D), 1, K), 256, 0, st_main>>>(tp->wq, tp->wk, tp->wv, wqkvT_all, D);
  CPC_LAUNCHED_N("concat_transpose3", st_main);
  CPC_TRY(launch_transpose_cast<T>(tp->wo, woT_all, D, D, st_main, K));       // woT[a][o] = Wo[o][a]
  CPC_TRY(launch_transpose_cast<T>(tp->w1, w1T_all, F, D, st_main, K));       // [D][F]
  CPC_TRY(launch_transpose_cast<T>(tp->w2, w2T_all, D, F, st_main, K));       // [F][D]
  Lanes lanes;
  CPC_TRY(lanes_fork(st_main, K, &lanes));
  for (int k = 0; k < K; k++) {
    const BwdScratch& sc = scr[k % lanes.n];
    st = lanes.st[k % lanes.n];
    T *ds2 = sc.ds2, *dy1 = sc.dy1, *ds1 = sc.ds1, *datt = sc.datt, *dx = sc.dx, *dh = sc.dh, *dqkv = sc.dqkv;
    float* dcw = sc.dcw;
    const T* wqkvT = wqkvT_all + (size_t)k * 3 * D * D;
    const T* woT = woT_all + (size_t)k * D * D;
    const T* w1T = w1T_all + (size_t)k * F * D;
    const T* w2T = w2T_all + (size_t)k * F * D;
    const T* blk = sv + (size_t)k * lay.per_k;
    const T *qkv = blk + lay.qkv, *att = blk + lay.att, *s1 = blk + lay.s1, *y1 = blk + lay.y1, *h = blk + lay.h, *s2 = blk + lay.s2;
    const size_t oDD = (size_t)k * D * D, oFD = (size_t)k * F * D;
    // LN2 backward
    CPC_TRY(launch_ln_bwd<T>(dpred + (size_t)k * D, (long long)K * D, nullptr, s2, tp->ln2_w + (size_t)k * D, ds2,
                             gr->ln2_w + (size_t)k * D, gr->ln2_b + (size_t)k * D, P, D, st));
    // FFN backward
    CPC_TRY(launch_colsum<T>(ds2, gr->b2 + (size_t)k * D, P, D, st));
    {
      RowView A{ds2, 0, (long long)D, P}, Bv{h, 0, (long long)F, P};
      CPC_TRY(gemm_tn(g.bf16, 1, D, F, A, Bv, gr->w2 + oFD, F, STORE_PLAIN, 0, 0, st));       // dW2[d][f]
      OutView C{dh, 0, (long long)F, P, 0, P, 0};
      CPC_TRY(gemm_nt(g.bf16, false, 1, F, D, A, w2T, nullptr, C, st));                       // dh = ds2 . W2
    }
    const bool drop = tp->att_keep != nullptr && tp->ffn_keep != nullptr;
    relu_mask_kernel<T><<<grid_for((long long)P * F), 256, 0, st>>>(dh, h, (long long)P * F, drop ? tp->keep_scale : 1.f);
    CPC_LAUNCHED_N("relu_mask", st);
    CPC_TRY(launch_colsum<T>(dh, gr->b1 + (size_t)k * F, P, F, st));
    {
      RowView A{dh, 0, (long long)F, P}, Bv{y1, 0, (long long)D, P};
      CPC_TRY(gemm_tn(g.bf16, 1, F, D, A, Bv, gr->w1 + oFD, D, STORE_PLAIN, 0, 0, st));       // dW1[f][d]
      OutView C{dy1, 0, (long long)D, P, 0, P, 0};
      CPC_TRY(gemm_nt(g.bf16, false, 1, D, F, A, w1T, nullptr, C, st));                       // dy1 (FFN branch)
    }
    // LN1 backward on dy1 + ds2 (residual)
    CPC_TRY(launch_ln_bwd<T>(dy1, (long long)D, ds2, s1, tp->ln1_w + (size_t)k * D, ds1, gr->ln1_w + (size_t)k * D,
                             gr->ln1_b + (size_t)k * D, P, D, st));
    {  // Wo
      RowView A{ds1, 0, (long long)D, P}, Bv{att, 0, (long long)D, P};
      CPC_TRY(gemm_tn(g.bf16, 1, D, D, A, Bv, gr->wo + oDD, D, STORE_PLAIN, 0, 0, st));
      OutView C{datt, 0, (long long)D, P, 0, P, 0};
      CPC_TRY(gemm_nt(g.bf16, false, 1, D, D, A, woT, nullptr, C, st));
    }
    CPC_TRY(launch_attn_bwd<T>(qkv, datt, att, tp->krelpos + (size_t)k * DK * W, dqkv, gr->krelpos + (size_t)k * DK * W, B, W, D, nh,
                               drop ? tp->att_keep + (size_t)k * B * nh * W * W : nullptr, tp->keep_scale, st));
    {  // Wq, Wk, Wv and d(x)
      float* dw3[3] = {gr->wq + oDD, gr->wk + oDD, gr->wv + oDD};
      for (int j = 0; j < 3; j++) {
        RowView A{dqkv + (size_t)j * D, (long long)W * 3 * D, (long long)3 * D, W};
        CPC_TRY(gemm_tn(g.bf16, B, D, D, A, X, dw3[j], D, STORE_PLAIN, 0, 0, st));
      }
      RowView A{dqkv, 0, (long long)3 * D, P};
      OutView C{dx, 0, (long long)D, P, 0, P, 0};
      CPC_TRY(gemm_nt(g.bf16, false, 1, D, 3 * D, A, wqkvT, nullptr, C, st));
    }
    acc_add_kernel<T><<<grid_for((long long)P * D), 256, 0, st>>>(dcw, dx, ds1, (long long)P * D);
    CPC_LAUNCHED_N("acc_add", st);
  }
  CPC_TRY(lanes_join(lanes));
  st = st_main;
  scatter_rows_kernel<<<grid_for((long long)P * D), 256, 0, st>>>(scr[0].dcw, (size_t)((char*)scr[1].dcw - (char*)scr[0].dcw) / 4, kLanes, dc, B, S, W, D);
  CPC_LAUNCHED_N("scatter_rows", st);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// the same layer as a CONTEXT network (--arMode transformer: feature_loader.py:138-142 -> buildTransformerAR(hiddenEncoder,
// 1, sizeWindow // 160, abspos=False), transformers.py:129-139): one TransformerLayer over all S frames of a window.
// x (B, S, D) fp32 -> y (B, S, D) fp32; g.K == 1 and g.W == g.S here.
// ---------------------------------------------------------------------------------------------------------
namespace {
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}
}  // namespace

size_t tlayer_save_bytes(const Geo& g) {
  return thead_save_bytes(g) + (g.bf16 ? align_up((size_t)g.B * g.S * g.H * 2) : 0) + 256;
}
size_t tlayer_ws_bytes(const Geo& g, int backward) {
  return thead_ws_bytes(g, backward) + (g.bf16 ? align_up((size_t)g.B * g.S * g.H * 2) : 0) + 256;
}
int tlayer_fwd(const Geo& g, const float* x, const cpcb200_thead_params* tp, float* y, void* save, void* wsp, size_t ws_bytes,
               cudaStream_t st) {
  const long long n = (long long)g.B * g.S * g.H;
  Carver ws(wsp, ws_bytes);
  if (!g.bf16) return thead_fwd<float>(g, x, tp, y, save, ws, st);
  bf16* xT = static_cast<bf16*>(save);
  char* sv = static_cast<char*>(save) + align_up((size_t)n * 2);
  bf16* yT = ws.take<bf16>((size_t)n);
  CPC_TRY(launch_cast<bf16>(x, xT, n, st));
  CPC_TRY(thead_fwd<bf16>(g, xT, tp, yT, sv, ws, st));
  bf16_to_f32_kernel<<<grid_for(n), 256, 0, st>>>(yT, y, n);
  CPC_LAUNCHED_N("bf16_to_f32", st);
  return 0;
}
int tlayer_bwd(const Geo& g, const float* x, const cpcb200_thead_params* tp, const float* dy, const void* save, float* dx,
               const cpcb200_thead_params* gr, void* wsp, size_t ws_bytes, cudaStream_t st) {
  const long long n = (long long)g.B * g.S * g.H;
  Carver ws(wsp, ws_bytes);
  if (!g.bf16) return thead_bwd<float>(g, x, tp, dy, save, dx, gr, ws, st);
  const bf16* xT = static_cast<const bf16*>(save);
  const char* sv = static_cast<const char*>(save) + align_up((size_t)n * 2);
  bf16* dyT = ws.take<bf16>((size_t)n);
  CPC_TRY(launch_cast<bf16>(dy, dyT, n, st));
  return thead_bwd<bf16>(g, xT, tp, dyT, sv, dx, gr, ws, st);
}

template int thead_fwd<float>(const Geo&, const float*, const cpcb200_thead_params*, float*, void*, Carver&, cudaStream_t);
template int thead_fwd<bf16>(const Geo&, const bf16*, const cpcb200_thead_params*, bf16*, void*, Carver&, cudaStream_t);
template int thead_bwd<float>(const Geo&, const float*, const cpcb200_thead_params*, const float*, const void*, float*,
                              const cpcb200_thead_params*, Carver&, cudaStream_t);
template int thead_bwd<bf16>(const Geo&, const bf16*, const cpcb200_thead_params*, const bf16*, const void*, float*,
                             const cpcb200_thead_params*, Carver&, cudaStream_t);

}  // namespace cpcb200
