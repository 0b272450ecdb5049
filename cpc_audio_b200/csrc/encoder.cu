// encoder.cu - CPCEncoder forward/backward (reference: cpc/model.py:83-105, ChannelNorm model.py:50-58).
//
// Data layout in HBM: every activation is CHANNEL-LAST, (B, kPad + L_i + kPad, H) with kPad zero rows around
// each window.  A strided Conv1d(k, s, p) over such a tensor is a plain GEMM whose A operand is a view with
// OVERLAPPING rows: output row t reads the k*H contiguous elements starting at padded row (kPad - p + s*t).
// No im2col buffer exists anywhere; the transposed conv (dgrad) is s GEMMs over a 2-row overlapping view of
// the output gradient, the weight gradient is a reduction over positions of the same views.
//
// conv0 (C_in = 1, k = 10) is not a GEMM: it is HBM-bound (AI ~ 5 FLOP/B) and runs as one fused CUDA-core
// kernel conv + ChannelNorm + ReLU, one warp per output frame, channel-last vectorised stores.  Its backward
// recomputes the conv instead of saving the 268 MB pre-norm tensor.
#include "common.cuh"

namespace cpcb200 {

int gemm_nt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias,
            const OutView& C, cudaStream_t st);

bool conv0_mma_supported(int H);
bool conv0_mma_wide_fwd_supported(int H);
bool conv0_mma_wide_bwd_supported(int H, int L0);
int conv0_fwd_mma(const float* x, const float* w, const float* bias, const float* gam, const float* bet, bf16* y, int B, int L, int L0,
                  int H, cudaStream_t st);
int conv0_bwd_mma(const float* x, const float* w, const float* bias, const float* gam, const float* bet, bf16* dy, float* dw,
                  float* dbias, float* dgam, float* dbet, int B, int L, int L0, int H, cudaStream_t st);

namespace {

constexpr float kEps = 1e-5f;  // cpc/model.py:29

// Row ops: one warp per (b, t) row of H channels; lane owns float4 chunks at channel 4*(lane + 32*i), i < I.
template <int I> struct RowRegs { float v[I][4]; };

template <int I, class T>
__device__ __forceinline__ void row_load(const T* row, int H, int lane, float (&v)[I][4]) {
#pragma unroll
  for (int i = 0; i < I; i++) {
    int c = 4 * (lane + 32 * i);
    if (c < H) load_vec<4>(row + c, v[i]);
    else { v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f; }
  }
}
template <int I, class T>
__device__ __forceinline__ void row_store(T* row, int H, int lane, const float (&v)[I][4]) {
#pragma unroll
  for (int i = 0; i < I; i++) {
    int c = 4 * (lane + 32 * i);
    if (c < H) store_vec<4>(row + c, v[i]);
  }
}

// zero one padded row (all H channels) with a warp
template <class T>
__device__ __forceinline__ void zero_row(T* row, int H, int lane) {
  for (int c = 4 * lane; c < H; c += 128) { float z4[4] = {0.f, 0.f, 0.f, 0.f}; store_vec<4>(row + c, z4); }
}

// ChannelNorm statistics of one row held across a warp (two-pass, unbiased variance: model.py:52-54)
template <int I>
__device__ __forceinline__ void row_stats(const float (&u)[I][4], int H, int lane, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < I; i++)
    if (4 * (lane + 32 * i) < H) s += (u[i][0] + u[i][1]) + (u[i][2] + u[i][3]);
  mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < I; i++)
    if (4 * (lane + 32 * i) < H) {
#pragma unroll
      for (int j = 0; j < 4; j++) { float dlt = u[i][j] - mean; q = fmaf(dlt, dlt, q); }
    }
  float var = warp_sum(q) / (float)(H - 1);
  rstd = rsqrtf(var + kEps);
}

// ---------------------------------------------------------------------------------------------------------
// conv0 + ChannelNorm + ReLU  (model.py:100).  x (B, L) fp32 -> y0 (B, Lp0, H) T.
// One warp per chunk of kC0Chunk consecutive frames of one window: the lane's 4*I x 10 weights live in registers,
// the 10-sample input window slides by 5 samples per frame (5 broadcast loads), stores are channel-last vectors.
// ---------------------------------------------------------------------------------------------------------
constexpr int kC0Chunk = 32;  // frames per warp trip (the last chunk of a window is shorter when L0 % 32 != 0)

template <int I>
__device__ __forceinline__ void conv0_load_w(const float* __restrict__ w, int H, int lane, float (&wr)[I][4][10]) {
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int k = 0; k < 10; k++) wr[i][j][k] = (c < H) ? __ldg(w + (c + j) * 10 + k) : 0.f;
  }
}

template <int I>
__device__ __forceinline__ void conv0_row(const float (&wr)[I][4][10], const float* __restrict__ bsm, const float (&xs)[10],
                                          int H, int lane, float (&u)[I][4]) {
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < H) {
      float4 a = *reinterpret_cast<const float4*>(bsm + c);
      u[i][0] = a.x; u[i][1] = a.y; u[i][2] = a.z; u[i][3] = a.w;
#pragma unroll
      for (int k = 0; k < 10; k++) {
#pragma unroll
        for (int j = 0; j < 4; j++) u[i][j] = fmaf(wr[i][j][k], xs[k], u[i][j]);
      }
    } else { u[i][0] = u[i][1] = u[i][2] = u[i][3] = 0.f; }
  }
}

// sliding 10-sample window of frame t: samples 5t-3 .. 5t+6 (zero outside [0, L))
__device__ __forceinline__ void conv0_window_init(const float* __restrict__ xb, int L, int t, float (&xs)[10]) {
  const int s0 = 5 * t - 3;
#pragma unroll
  for (int j = 0; j < 10; j++) { const int s = s0 + j; xs[j] = (s >= 0 && s < L) ? __ldg(xb + s) : 0.f; }
}
__device__ __forceinline__ void conv0_window_next(const float* __restrict__ xb, int L, int t_next, float (&xs)[10]) {
#pragma unroll
  for (int j = 0; j < 5; j++) xs[j] = xs[j + 5];
  const int s0 = 5 * t_next + 2;
#pragma unroll
  for (int j = 0; j < 5; j++) { const int s = s0 + j; xs[5 + j] = (s < L) ? __ldg(xb + s) : 0.f; }
}

template <int I, class T>
__global__ void __launch_bounds__(128) conv0_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ gam,
                                                         const float* __restrict__ bet, T* __restrict__ y, int B, int L,
                                                         int L0, int H) {
  extern __shared__ __align__(16) float sm[];  // bias, gamma, beta: [3][H]
  for (int i = threadIdx.x; i < H; i += blockDim.x) { sm[i] = bias[i]; sm[H + i] = gam[i]; sm[2 * H + i] = bet[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float wr[I][4][10];
  conv0_load_w<I>(w, H, lane, wr);
  const int cpw = (L0 + kC0Chunk - 1) / kC0Chunk;  // chunks per window
  const long long Lp0 = L0 + 2 * kPad;
  for (int ch = warp; ch < B * cpw; ch += nwarps) {
    const int b = ch / cpw, t0 = (ch - b * cpw) * kC0Chunk;
    const float* xb = x + (long long)b * L;
    float xs[10];
    conv0_window_init(xb, L, t0, xs);
    T* yrow = y + ((long long)b * Lp0 + kPad + t0) * H;
    if (t0 == 0) { for (int r = 0; r < kPad; r++) zero_row(y + ((long long)b * Lp0 + r) * H, H, lane); }
    if (t0 + kC0Chunk >= L0) { for (int r = 0; r < kPad; r++) zero_row(y + ((long long)b * Lp0 + kPad + L0 + r) * H, H, lane); }
    const int nfr = min(kC0Chunk, L0 - t0);
#pragma unroll 1
    for (int tt = 0; tt < nfr; tt++) {
      float u[I][4];
      conv0_row<I>(wr, sm, xs, H, lane, u);
      if (tt + 1 < nfr) conv0_window_next(xb, L, t0 + tt + 1, xs);
      float mean, rstd;
      row_stats<I>(u, H, lane, mean, rstd);
#pragma unroll
      for (int i = 0; i < I; i++) {
        const int c = 4 * (lane + 32 * i);
        if (c < H) {
          float4 g4 = *reinterpret_cast<const float4*>(sm + H + c);
          float4 b4 = *reinterpret_cast<const float4*>(sm + 2 * H + c);
          float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int j = 0; j < 4; j++) u[i][j] = fmaxf(fmaf((u[i][j] - mean) * rstd, g[j], be[j]), 0.f);
        }
      }
      row_store<I>(yrow + (long long)tt * H, H, lane, u);
    }
  }
}

// conv0 backward, part 1: recompute u from x, ChannelNorm+ReLU backward in place (dy0 -> du0, same buffer),
// accumulate dbias0, dgamma0, dbeta0.
template <int I, class T>
__global__ void __launch_bounds__(128) conv0_bwd_du_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ gam,
                                                            const float* __restrict__ bet, T* __restrict__ dy,
                                                            float* __restrict__ dbias, float* __restrict__ dgam,
                                                            float* __restrict__ dbet, int B, int L, int L0, int H) {
  extern __shared__ __align__(16) float sm[];  // [3][H] params + [3][H] accumulators
  float* accs = sm + 3 * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) { sm[i] = bias[i]; sm[H + i] = gam[i]; sm[2 * H + i] = bet[i]; }
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) accs[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float wr[I][4][10];
  conv0_load_w<I>(w, H, lane, wr);
  float ab[I][4], ag[I][4], abe[I][4];
#pragma unroll
  for (int i = 0; i < I; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) ab[i][j] = ag[i][j] = abe[i][j] = 0.f;
  const int cpw = (L0 + kC0Chunk - 1) / kC0Chunk;
  for (int ch = warp; ch < B * cpw; ch += nwarps) {
    const int b = ch / cpw, t0 = (ch - b * cpw) * kC0Chunk;
    const float* xb = x + (long long)b * L;
    float xs[10];
    conv0_window_init(xb, L, t0, xs);
    T* drow = dy + ((long long)b * L0 + t0) * H;
    const int nfr = min(kC0Chunk, L0 - t0);
#pragma unroll 1
    for (int tt = 0; tt < nfr; tt++) {
      float u[I][4], d[I][4];
      row_load<I>(drow + (long long)tt * H, H, lane, d);
      conv0_row<I>(wr, sm, xs, H, lane, u);
      if (tt + 1 < nfr) conv0_window_next(xb, L, t0 + tt + 1, xs);
      float mean, rstd;
      row_stats<I>(u, H, lane, mean, rstd);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < I; i++) {
        const int c = 4 * (lane + 32 * i);
        if (c < H) {
          float4 g4 = *reinterpret_cast<const float4*>(sm + H + c);
          float4 b4 = *reinterpret_cast<const float4*>(sm + 2 * H + c);
          float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float xh = (u[i][j] - mean) * rstd;
            const float v = fmaf(xh, g[j], be[j]);
            const float dv = v > 0.f ? d[i][j] : 0.f;
            ag[i][j] = fmaf(dv, xh, ag[i][j]);
            abe[i][j] += dv;
            const float dx = dv * g[j];
            u[i][j] = xh; d[i][j] = dx;
            s1 += dx; s2 = fmaf(dx, xh, s2);
          }
        }
      }
      s1 = warp_sum(s1) / (float)H;
      s2 = warp_sum(s2) / (float)(H - 1);
#pragma unroll
      for (int i = 0; i < I; i++) {
        if (4 * (lane + 32 * i) < H) {
#pragma unroll
          for (int j = 0; j < 4; j++) { const float o = rstd * (d[i][j] - s1 - u[i][j] * s2); ab[i][j] += o; d[i][j] = o; }
        }
      }
      row_store<I>(drow + (long long)tt * H, H, lane, d);
    }
  }
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < H) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        atomicAdd(&accs[c + j], ab[i][j]); atomicAdd(&accs[H + c + j], ag[i][j]); atomicAdd(&accs[2 * H + c + j], abe[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dbias + i, accs[i]); atomicAdd(dgam + i, accs[H + i]); atomicAdd(dbet + i, accs[2 * H + i]);
  }
}

// conv0 backward, part 2: dW0[c][tap] += sum_{b,t} du0[b,t,c] * x[b, 5t-3+tap]
template <int I, class T>
__global__ void __launch_bounds__(128) conv0_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ du,
                                                           float* __restrict__ dw, int B, int L, int L0, int H) {
  extern __shared__ __align__(16) float accs[];  // [10][H]
  for (int i = threadIdx.x; i < 10 * H; i += blockDim.x) accs[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float aw[I][4][10];
#pragma unroll
  for (int i = 0; i < I; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int k = 0; k < 10; k++) aw[i][j][k] = 0.f;
  const int cpw = (L0 + kC0Chunk - 1) / kC0Chunk;
  for (int ch = warp; ch < B * cpw; ch += nwarps) {
    const int b = ch / cpw, t0 = (ch - b * cpw) * kC0Chunk;
    const float* xb = x + (long long)b * L;
    float xs[10];
    conv0_window_init(xb, L, t0, xs);
    const T* drow = du + ((long long)b * L0 + t0) * H;
    const int nfr = min(kC0Chunk, L0 - t0);
#pragma unroll 2
    for (int tt = 0; tt < nfr; tt++) {
      float d[I][4];
      row_load<I>(drow + (long long)tt * H, H, lane, d);
#pragma unroll
      for (int i = 0; i < I; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int k = 0; k < 10; k++) aw[i][j][k] = fmaf(d[i][j], xs[k], aw[i][j][k]);
      if (tt + 1 < nfr) conv0_window_next(xb, L, t0 + tt + 1, xs);
    }
  }
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < H) {
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 10; k++) atomicAdd(&accs[k * H + c + j], aw[i][j][k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * H; i += blockDim.x) { const int tap = i / H, c = i - tap * H; atomicAdd(dw + c * 10 + tap, accs[i]); }
}

// ---------------------------------------------------------------------------------------------------------
// ChannelNorm + ReLU over a padded activation (layers 1..4).  u (B,Lp,H) -> y (B,Lp,H) [+ z (B,Lc,H) fp32]
// ---------------------------------------------------------------------------------------------------------
template <int I, class T>
__global__ void __launch_bounds__(256) cnorm_relu_fwd_kernel(const T* __restrict__ u, const float* __restrict__ gam,
                                                              const float* __restrict__ bet, T* __restrict__ y,
                                                              float* __restrict__ zout, float2* __restrict__ stats, int B, int Lc,
                                                              int H) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long rows = (long long)B * Lc;
  if (warp >= rows) return;
  const int b = (int)(warp / Lc), t = (int)(warp - (long long)b * Lc);
  const long long prow = ((long long)b * (Lc + 2 * kPad) + kPad + t) * H;
  float v[I][4];
  row_load<I>(u + prow, H, lane, v);
  float mean, rstd;
  row_stats<I>(v, H, lane, mean, rstd);
  if (lane == 0 && stats != nullptr) stats[warp] = make_float2(mean, rstd);  // saved for backward
#pragma unroll
  for (int i = 0; i < I; i++) {
    int c = 4 * (lane + 32 * i);
    if (c < H) {
      float4 g4 = *reinterpret_cast<const float4*>(gam + c);
      float4 b4 = *reinterpret_cast<const float4*>(bet + c);
      float g[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; j++) v[i][j] = fmaxf(fmaf((v[i][j] - mean) * rstd, g[j], be[j]), 0.f);
    }
  }
  if (y) {
    row_store<I>(y + prow, H, lane, v);
    const long long w0 = (long long)b * (Lc + 2 * kPad) * H;
    if (t == 0) { for (int r = 0; r < kPad; r++) zero_row(y + w0 + (long long)r * H, H, lane); }
    if (t == Lc - 1) { for (int r = 0; r < kPad; r++) zero_row(y + w0 + (long long)(kPad + Lc + r) * H, H, lane); }
  }
  if (zout) row_store<I>(zout + warp * H, H, lane, v);
}

// backward of ChannelNorm+ReLU: dy (B,Lc,H) TD unpadded, u padded -> du padded (T); dgamma/dbeta/dbias +=
// Persistent grid; a warp takes R consecutive rows per trip: all 2R row loads are in flight before the first use and
// the four warp reductions of the R rows (sum, centred squares, s1, s2) run as R interleaved shuffle chains.  The
// per-channel sums stay in registers for the whole kernel and leave through one shared-memory reduction over the
// 8 warps + one global atomic per channel per CTA.
template <int R>
__device__ __forceinline__ void warp_sum_n(float (&x)[R]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < R; r++) x[r] += __shfl_xor_sync(0xffffffffu, x[r], off);
}

template <int I, int R, class TD, class T>
__device__ __forceinline__ void cnorm_bwd_rows(const TD* __restrict__ dy, const T* __restrict__ u,
                                               const float2* __restrict__ stats, T* __restrict__ du,
                                               long long r0l, int Lc, int H, int lane, const float (&g)[I][4],
                                               const float (&be)[I][4], float (&ag)[I][4], float (&abe)[I][4],
                                               float (&ab)[I][4], float inv_h, float inv_h1) {
  float v[R][I][4], d[R][I][4];
  long long prow[R];
  int tt[R], bb[R];
  const long long Lp = (long long)Lc + 2 * kPad;
  {
    const unsigned r0 = (unsigned)r0l;  // rows < 2^31 (checked by the host)
    int b = (int)(r0 / (unsigned)Lc), t = (int)(r0 - (unsigned)b * (unsigned)Lc);
#pragma unroll
    for (int rr = 0; rr < R; rr++) {
      bb[rr] = b; tt[rr] = t;
      prow[rr] = ((long long)b * Lp + kPad + t) * H;
      row_load<I>(u + prow[rr], H, lane, v[rr]);
      row_load<I>(dy + (r0l + rr) * H, H, lane, d[rr]);
      if (++t >= Lc) { t = 0; b++; }
    }
  }
  // (mean, rstd) of every row were saved by the forward pass (GEMM epilogue or cnorm_relu_fwd_kernel)
  float rstd[R], nmr[R];
  float s1[R], s2[R];
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    const float2 st = __ldg(stats + r0l + rr);
    rstd[rr] = st.y;
    nmr[rr] = -st.x * st.y;
    s1[rr] = s2[rr] = 0.f;
#pragma unroll
    for (int i = 0; i < I; i++) {
      if (4 * (lane + 32 * i) < H) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float xh = fmaf(v[rr][i][j], rstd[rr], nmr[rr]);
          const float dv = fmaf(xh, g[i][j], be[i][j]) > 0.f ? d[rr][i][j] : 0.f;
          ag[i][j] = fmaf(dv, xh, ag[i][j]);
          abe[i][j] += dv;
          const float dx = dv * g[i][j];
          v[rr][i][j] = xh; d[rr][i][j] = dx;
          s1[rr] += dx; s2[rr] = fmaf(dx, xh, s2[rr]);
        }
      }
    }
  }
  warp_sum_n<R>(s1);
  warp_sum_n<R>(s2);
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    const float c1 = -rstd[rr] * s1[rr] * inv_h, c2 = -rstd[rr] * s2[rr] * inv_h1;
#pragma unroll
    for (int i = 0; i < I; i++) {
      if (4 * (lane + 32 * i) < H) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float o = fmaf(v[rr][i][j], c2, fmaf(d[rr][i][j], rstd[rr], c1));
          ab[i][j] += o; d[rr][i][j] = o;
        }
      }
    }
    row_store<I>(du + prow[rr], H, lane, d[rr]);
    if (tt[rr] == 0 || tt[rr] == Lc - 1) {
      const long long w0 = (long long)bb[rr] * Lp * H;
      if (tt[rr] == 0) { for (int r2 = 0; r2 < kPad; r2++) zero_row(du + w0 + (long long)r2 * H, H, lane); }
      if (tt[rr] == Lc - 1) { for (int r2 = 0; r2 < kPad; r2++) zero_row(du + w0 + (long long)(kPad + Lc + r2) * H, H, lane); }
    }
  }
}

template <int I, class TD, class T>
__global__ void __launch_bounds__(256, (I <= 2 ? 2 : 1)) cnorm_relu_bwd_kernel(const TD* __restrict__ dy, const T* __restrict__ u,
                                                                 const float* __restrict__ gam, const float* __restrict__ bet,
                                                                 const float2* __restrict__ stats, T* __restrict__ du,
                                                                 float* __restrict__ dgam,
                                                                 float* __restrict__ dbet, float* __restrict__ dbias, int B,
                                                                 int Lc, int H) {
  pdl_wait();
  pdl_trigger();
  constexpr int kCbR = I <= 2 ? 4 : 2;
  extern __shared__ __align__(16) float red[];  // [8 warps][3][H]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows = (long long)B * Lc;
  float ag[I][4], abe[I][4], ab[I][4], g[I][4], be[I][4];
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
    if (c < H) { g4 = *reinterpret_cast<const float4*>(gam + c); b4 = *reinterpret_cast<const float4*>(bet + c); }
    g[i][0] = g4.x; g[i][1] = g4.y; g[i][2] = g4.z; g[i][3] = g4.w;
    be[i][0] = b4.x; be[i][1] = b4.y; be[i][2] = b4.z; be[i][3] = b4.w;
#pragma unroll
    for (int j = 0; j < 4; j++) ag[i][j] = abe[i][j] = ab[i][j] = 0.f;
  }
  const float inv_h = 1.f / (float)H, inv_h1 = 1.f / (float)(H - 1);
  const long long full = rows / kCbR;  // trips of kCbR rows; the < kCbR leftover rows go one at a time
  for (long long trip = warp0; trip < full; trip += nwarps)
    cnorm_bwd_rows<I, kCbR, TD, T>(dy, u, stats, du, trip * kCbR, Lc, H, lane, g, be, ag, abe, ab, inv_h, inv_h1);
  for (long long r = full * kCbR + warp0; r < rows; r += nwarps)
    cnorm_bwd_rows<I, 1, TD, T>(dy, u, stats, du, r, Lc, H, lane, g, be, ag, abe, ab, inv_h, inv_h1);
  float* mine = red + (size_t)wib * 3 * H;
#pragma unroll
  for (int i = 0; i < I; i++) {
    const int c = 4 * (lane + 32 * i);
    if (c < H) {
      *reinterpret_cast<float4*>(mine + c) = make_float4(ag[i][0], ag[i][1], ag[i][2], ag[i][3]);
      *reinterpret_cast<float4*>(mine + H + c) = make_float4(abe[i][0], abe[i][1], abe[i][2], abe[i][3]);
      *reinterpret_cast<float4*>(mine + 2 * H + c) = make_float4(ab[i][0], ab[i][1], ab[i][2], ab[i][3]);
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
    float a = 0.f;
    for (int ww = 0; ww < nw; ww++) a += red[(size_t)ww * 3 * H + i];
    const int k = i / H, c = i - k * H;
    atomicAdd((k == 0 ? dgam : k == 1 ? dbet : dbias) + c, a);
  }
}

// ---------------------------------------------------------------------------------------------------------
// small prep kernels
// ---------------------------------------------------------------------------------------------------------
// ---- all four H->H layers in one launch each (blockIdx.y = layer - 1) ----------------------------------
struct Conv4Ptrs { const float* w[4]; void* out[4]; float* acc[4]; int taps[4]; int s[4]; };

// forward GEMM weights: block (co, layer), thread ci reads its `taps` contiguous floats (coalesced 32-B pieces) and
// writes Wp[co][tap*Ci + ci] coalesced over ci
template <class T>
__global__ void prep_w_fwd_all_kernel(Conv4Ptrs P, int Ci) {
  pdl_wait();
  pdl_trigger();
  const int l = blockIdx.y, co = blockIdx.x, taps = P.taps[l];
  const float* w = P.w[l] + (size_t)co * Ci * taps;
  T* wp = static_cast<T*>(P.out[l]) + (size_t)co * Ci * taps;
  for (int ci = threadIdx.x; ci < Ci; ci += blockDim.x) {
    for (int tap = 0; tap < taps; tap++) wp[(size_t)tap * Ci + ci] = from_f<T>(w[(size_t)ci * taps + tap]);
  }
}
// dgrad GEMM weights: block (ci, layer), thread co: Wd[r][ci][half*Co + co] = W[co][ci][r + s*(1-half)]
template <class T>
__global__ void prep_w_dgrad_all_kernel(Conv4Ptrs P, int Co, int Ci) {
  pdl_wait();
  pdl_trigger();
  const int l = blockIdx.y, ci = blockIdx.x, taps = P.taps[l], s = P.s[l];
  T* wd = static_cast<T*>(P.out[l]);
  for (int co = threadIdx.x; co < Co; co += blockDim.x) {
    const float* w = P.w[l] + ((size_t)co * Ci + ci) * taps;
    for (int tap = 0; tap < taps; tap++) {
      const int r = tap % s, half = 1 - tap / s;
      wd[((size_t)r * Ci + ci) * 2 * Co + (size_t)half * Co + co] = from_f<T>(w[tap]);
    }
  }
}
// dW[co][ci][tap] += scratch[co][tap*Ci + ci]: block (co, layer).  The scratch row is read coalesced into shared memory
// (rows padded by one float: the transposed reads are conflict-free) and the parameter-layout row is updated as one
// contiguous run.
__global__ void permute_add_wgrad_all_kernel(Conv4Ptrs P, int Ci) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float prow[];  // [taps][Ci + 1]
  const int l = blockIdx.y, co = blockIdx.x, taps = P.taps[l];
  const float* sc = P.w[l] + (size_t)co * Ci * taps;
  float* dw = P.acc[l] + (size_t)co * Ci * taps;
  const int n = Ci * taps;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int tap = i / Ci, ci = i - tap * Ci;
    prow[tap * (Ci + 1) + ci] = sc[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / taps, tap = i - ci * taps;
    dw[i] += prow[tap * (Ci + 1) + ci];
  }
}

struct EncLayout {
  size_t y[4], u[5];  // element offsets into save (u[0] unused)
  size_t st[5];       // (mean, rstd) per row of layers 1..4 (float2, offsets in elements of the activation type)
  size_t total;       // elements
};
EncLayout enc_layout(const Geo& g) {
  EncLayout e{};
  size_t off = 0;
  auto take = [&](size_t n) { size_t r = off; off += (n + 127) / 128 * 128; return r; };
  for (int i = 0; i < 4; i++) e.y[i] = take((size_t)g.B * (g.Lout[i] + 2 * kPad) * g.H);
  for (int i = 1; i < 5; i++) e.u[i] = take((size_t)g.B * (g.Lout[i] + 2 * kPad) * g.H);
  for (int i = 1; i < 5; i++) e.st[i] = take((size_t)g.B * g.Lout[i] * 8 / (g.bf16 ? 2 : 4));
  e.total = off;
  return e;
}

inline int ilog_I(int H) { return (H + 127) / 128; }

template <class T>
int encoder_fwd_t(const Geo& g, const float* x, const cpcb200_encoder_params* p, float* z, void* save, void* wsp,
                  size_t ws_bytes, cudaStream_t st) {
  const int H = g.H, B = g.B;
  EncLayout e = enc_layout(g);
  Carver ws(wsp, ws_bytes);
  T* wp[5] = {nullptr};
  for (int i = 1; i < 5; i++) wp[i] = ws.take<T>((size_t)H * kConvK[i] * H);
  // training: every activation lives in `save` (read again by the backward pass).  Inference (save == NULL: no_grad
  // forward, feature extraction): y0..y3 ping-pong between two workspace buffers, the pre-norm rows and the row
  // statistics are not kept at all on the fused path (one scratch buffer on the unfused one).
  const bool infer = save == nullptr;
  T* yb[4] = {nullptr};
  T* ub[5] = {nullptr};
  float2* stb[5] = {nullptr};
  if (!infer) {
    T* sv = static_cast<T*>(save);
    for (int i = 0; i < 4; i++) yb[i] = sv + e.y[i];
    for (int i = 1; i < 5; i++) { ub[i] = sv + e.u[i]; stb[i] = reinterpret_cast<float2*>(sv + e.st[i]); }
  } else {
    T* pa = ws.take<T>((size_t)B * (g.Lout[0] + 2 * kPad) * H);
    T* pb = ws.take<T>((size_t)B * (g.Lout[1] + 2 * kPad) * H);
    T* us = ws.take<T>((size_t)B * (g.Lout[1] + 2 * kPad) * H);
    yb[0] = pa; yb[1] = pb; yb[2] = pa; yb[3] = pb;
    for (int i = 1; i < 5; i++) ub[i] = us;
  }
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "encoder_fwd: workspace %zu < %zu", ws_bytes, ws.off);

  {
    Conv4Ptrs P{};
    for (int i = 1; i < 5; i++) { P.w[i - 1] = p->conv_w[i]; P.out[i - 1] = wp[i]; P.taps[i - 1] = kConvK[i]; P.s[i - 1] = kConvS[i]; }
    CPC_CHECK_CUDA(launch_k(prep_w_fwd_all_kernel<T>, dim3(H, 4), dim3(256), 0, st, 1, P, H));
    CPC_LAUNCHED_N("prep_w_fwd_all", st);
  }
  // the zero rows around every window of y0..y3 are written by the kernels that produce the interior
  const int I = ilog_I(H);
  bool c0_done = false;
  if constexpr (sizeof(T) == 2) {
    if (conv0_mma_supported(H) || conv0_mma_wide_fwd_supported(H)) {
      CPC_TRY(conv0_fwd_mma(x, p->conv_w[0], p->conv_b[0], p->norm_w[0], p->norm_b[0], yb[0], B, g.L, g.Lout[0], H, st));
      c0_done = true;
    }
  }
  if (!c0_done) {
    const size_t smem = 3 * (size_t)H * sizeof(float);
    int blocks = (B * ((g.Lout[0] + kC0Chunk - 1) / kC0Chunk) + 3) / 4;
    if (blocks > 148 * 8) blocks = 148 * 8;
#define LAUNCH_C0(II)                                                                                             \
  conv0_fwd_kernel<II, T><<<blocks, 128, smem, st>>>(x, p->conv_w[0], p->conv_b[0], p->norm_w[0], p->norm_b[0],   \
                                                     yb[0], B, g.L, g.Lout[0], H)
    if (I == 1) LAUNCH_C0(1); else if (I == 2) LAUNCH_C0(2); else if (I == 3) LAUNCH_C0(3); else LAUNCH_C0(4);
#undef LAUNCH_C0
    CPC_LAUNCHED_N("conv0_fwd", st);
  }
  for (int i = 1; i < 5; i++) {
    const int Lin = g.Lout[i - 1], Lo = g.Lout[i];
    RowView A{yb[i - 1] + (size_t)(kPad - kConvP[i]) * H, (long long)(Lin + 2 * kPad) * H, (long long)kConvS[i] * H, Lo, kConvK[i], kConvS[i]};
    OutView C{ub[i] + (size_t)kPad * H, (long long)(Lo + 2 * kPad) * H, (long long)H, Lo, 0, Lo, 0};
    T* yo = i < 4 ? yb[i] : nullptr;
    float* zo = i == 4 ? z : nullptr;
    if (g.bf16 && H == 256) {  // ChannelNorm + ReLU inside the GEMM epilogue
      bool fused = false;
      CNormEpi E{p->norm_w[i], p->norm_b[i], yo != nullptr ? static_cast<void*>(yo + (size_t)kPad * H) : nullptr, zo, kPad, stb[i]};
      E.save_u = infer ? 0 : 1;
      CPC_TRY(gemm_nt_cnorm_tc(B, kConvK[i] * H, A, wp[i], p->conv_b[i], C, E, st, &fused));
      if (fused) continue;
    }
    CPC_TRY(gemm_nt(g.bf16, false, B, H, kConvK[i] * H, A, wp[i], p->conv_b[i], C, st));
    const long long rows = (long long)B * Lo;
    const int blocks = (int)((rows * 32 + 255) / 256);
#define LAUNCH_CN(II) CPC_CHECK_CUDA(launch_k(cnorm_relu_fwd_kernel<II, T>, dim3(blocks), dim3(256), 0, st, 1, ub[i], p->norm_w[i], p->norm_b[i], yo, zo, stb[i], B, Lo, H))
    if (I == 1) LAUNCH_CN(1); else if (I == 2) LAUNCH_CN(2); else if (I == 3) LAUNCH_CN(3); else LAUNCH_CN(4);
#undef LAUNCH_CN
    CPC_LAUNCHED_N("cnorm_relu_fwd", st);
  }
  return 0;
}

template <class T>
int encoder_bwd_t(const Geo& g, const float* x, const cpcb200_encoder_params* p, const float* dz, const void* save,
                  const cpcb200_encoder_params* gr, void* wsp, size_t ws_bytes, cudaStream_t st) {
  const int H = g.H, B = g.B;
  EncLayout e = enc_layout(g);
  const T* sv = static_cast<const T*>(save);
  Carver ws(wsp, ws_bytes);
  T* wd[5] = {nullptr};
  T* du[5] = {nullptr};
  T* dy[4] = {nullptr};
  for (int i = 1; i < 5; i++) wd[i] = ws.take<T>((size_t)kConvS[i] * H * 2 * H);
  for (int i = 1; i < 5; i++) du[i] = ws.take<T>((size_t)B * (g.Lout[i] + 2 * kPad) * H);
  for (int i = 0; i < 4; i++) dy[i] = ws.take<T>((size_t)B * g.Lout[i] * H);
  float* dwp[5] = {nullptr};
  size_t dwp_total = 0;
  for (int i = 1; i < 5; i++) { dwp[i] = ws.take<float>((size_t)H * kConvK[i] * H); dwp_total = (size_t)((char*)dwp[i] - (char*)dwp[1]) + (size_t)H * kConvK[i] * H * 4; }
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "encoder_bwd: workspace %zu < %zu", ws_bytes, ws.off);
  CPC_CHECK_CUDA(cudaMemsetAsync(dwp[1], 0, dwp_total, st));
  const int I = ilog_I(H);

  {
    Conv4Ptrs P{};
    for (int i = 1; i < 5; i++) { P.w[i - 1] = p->conv_w[i]; P.out[i - 1] = wd[i]; P.taps[i - 1] = kConvK[i]; P.s[i - 1] = kConvS[i]; }
    CPC_CHECK_CUDA(launch_k(prep_w_dgrad_all_kernel<T>, dim3(H, 4), dim3(256), 0, st, 1, P, H, H));
    CPC_LAUNCHED_N("prep_w_dgrad_all", st);
  }
  TnDesc wg[4];
  bool early_exchange = false;
  for (int i = 4; i >= 1; i--) {
    const int Lo = g.Lout[i], Lin = g.Lout[i - 1], s = kConvS[i], pp = kConvP[i];
    // ChannelNorm+ReLU backward -> du_i, dgamma_i, dbeta_i, dbias_i
    {
      const long long rows = (long long)B * Lo;
      const int kCbR = I <= 2 ? 4 : 2;
      int blocks = (int)((rows * 32 / kCbR + 255) / 256);
      if (blocks > 148 * 2) blocks = 148 * 2;
      if (blocks < 1) blocks = 1;
      const size_t smem = 8 * 3 * (size_t)H * sizeof(float);
#define LAUNCH_CB(II, TD, SRC)                                                                                       \
  CPC_CHECK_CUDA(launch_k(cnorm_relu_bwd_kernel<II, TD, T>, dim3(blocks), dim3(256), smem, st, 1, SRC, sv + e.u[i], p->norm_w[i], \
                          p->norm_b[i], reinterpret_cast<const float2*>(sv + e.st[i]), du[i], gr->norm_w[i], gr->norm_b[i],  \
                          gr->conv_b[i], B, Lo, H))
      if (i == 4) { if (I == 1) LAUNCH_CB(1, float, dz); else if (I == 2) LAUNCH_CB(2, float, dz); else if (I == 3) LAUNCH_CB(3, float, dz); else LAUNCH_CB(4, float, dz); }
      else { if (I == 1) LAUNCH_CB(1, T, dy[i]); else if (I == 2) LAUNCH_CB(2, T, dy[i]); else if (I == 3) LAUNCH_CB(3, T, dy[i]); else LAUNCH_CB(4, T, dy[i]); }
#undef LAUNCH_CB
      CPC_LAUNCHED_N("cnorm_relu_bwd", st);
    }
    // weight gradient: dW[co][ci][tap] += sum_{b,t} du[b,t,co] * y_{i-1}[b, s t - p + tap, ci].  The four products are
    // independent of the rest of the chain: they are collected here and run as ONE grouped launch after the loop.
    {
      RowView A{du[i] + (size_t)kPad * H, (long long)(Lo + 2 * kPad) * H, (long long)H, Lo};
      RowView Bv{sv + e.y[i - 1] + (size_t)(kPad - pp) * H, (long long)(Lin + 2 * kPad) * H, (long long)s * H, Lo, kConvK[i], s};
      wg[4 - i] = TnDesc{B, H, kConvK[i] * H, A, Bv, dwp[i], kConvK[i] * H, STORE_PLAIN, 0, 0};
    }
    if (i == 1) {
      // every du_i exists now: the four weight-gradient products run as ONE grouped launch BEFORE the last (largest) data
      // gradient, so that all parameter gradients except conv0's are final ~200 us before the backward pass ends - the
      // data-parallel all-reduce of those 99.9 % of the bucket overlaps dgrad_1 and the conv0 backward (GradBucket)
      CPC_TRY(gemm_tn_group(g.bf16, 4, wg, st));  // wg[0] = layer 4 ... wg[3] = layer 1
      {  // scratch (Co, k*Ci) layout -> parameter layout (Co, Ci, k), all four layers
        Conv4Ptrs P{};
        for (int l = 1; l < 5; l++) { P.w[l - 1] = dwp[l]; P.acc[l - 1] = gr->conv_w[l]; P.taps[l - 1] = kConvK[l]; P.s[l - 1] = kConvS[l]; }
        CPC_CHECK_CUDA(launch_k(permute_add_wgrad_all_kernel, dim3(H, 4), dim3(256), (size_t)8 * (H + 1) * sizeof(float), st, 1, P, H));
        CPC_LAUNCHED_N("permute_add_wgrad_all", st);
      }
      if (cudaEvent_t ev = take_grads_ready_event(st)) {
        CPC_CHECK_CUDA(cudaEventRecord(ev, st));
        early_exchange = true;  // the caller runs the gradient exchange beside dgrad_1: leave it its SMs
      }
    }
    // data gradient: input row j = s q + r - p gets [du[q-1], du[q]] . Wd[r].  All s residues in ONE GEMM with
    // N = s*H: row q of the product is the s consecutive input rows s q - p .. s q - p + s - 1.
    {
      RowView A{du[i] + (size_t)(kPad - 1) * H, (long long)(Lo + 2 * kPad) * H, (long long)H, Lo + 1, 2, 1};
      OutView C{dy[i - 1] - (long long)pp * H, (long long)Lin * H, (long long)s * H, Lo + 1, 0, Lo + 1, H, pp, s, Lin};
      // window lengths that are not multiples of 160 leave up to s-1 trailing input rows that no output frame reads
      // (zero gradient) and that the merged product does not cover: clear the buffer first (ragged shapes only)
      if (Lin > s * Lo + s - pp) CPC_CHECK_CUDA(cudaMemsetAsync(dy[i - 1], 0, (size_t)B * Lin * H * sizeof(T), st));
      if (i == 1 && early_exchange) set_sm_reserve(kEarlyExchangeSMs);
      const int rc = gemm_nt(g.bf16, false, B, s * H, 2 * H, A, wd[i], nullptr, C, st);
      set_sm_reserve(0);
      CPC_TRY(rc);
    }
  }
  bool c0_done = false;
  if constexpr (sizeof(T) == 2) {
    if (conv0_mma_supported(H) || conv0_mma_wide_bwd_supported(H, g.Lout[0])) {
      CPC_TRY(conv0_bwd_mma(x, p->conv_w[0], p->conv_b[0], p->norm_w[0], p->norm_b[0], dy[0], gr->conv_w[0], gr->conv_b[0],
                            gr->norm_w[0], gr->norm_b[0], B, g.L, g.Lout[0], H, st));
      c0_done = true;
    }
  }
  if (!c0_done) {
    int blocks = (B * ((g.Lout[0] + kC0Chunk - 1) / kC0Chunk) + 3) / 4;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const size_t smem1 = 6 * (size_t)H * sizeof(float), smem2 = 10 * (size_t)H * sizeof(float);
#define LAUNCH_C0B(II)                                                                                              \
  conv0_bwd_du_kernel<II, T><<<blocks, 128, smem1, st>>>(x, p->conv_w[0], p->conv_b[0], p->norm_w[0], p->norm_b[0], \
                                                         dy[0], gr->conv_b[0], gr->norm_w[0], gr->norm_b[0], B, g.L, \
                                                         g.Lout[0], H)
    if (I == 1) LAUNCH_C0B(1); else if (I == 2) LAUNCH_C0B(2); else if (I == 3) LAUNCH_C0B(3); else LAUNCH_C0B(4);
#undef LAUNCH_C0B
    CPC_LAUNCHED_N("conv0_bwd_du", st);
#define LAUNCH_C0W(II) conv0_wgrad_kernel<II, T><<<blocks, 128, smem2, st>>>(x, dy[0], gr->conv_w[0], B, g.L, g.Lout[0], H)
    if (I == 1) LAUNCH_C0W(1); else if (I == 2) LAUNCH_C0W(2); else if (I == 3) LAUNCH_C0W(3); else LAUNCH_C0W(4);
#undef LAUNCH_C0W
    CPC_LAUNCHED_N("conv0_wgrad", st);
  }
  return 0;
}

}  // namespace

size_t encoder_save_elems(const Geo& g) { return enc_layout(g).total; }

size_t encoder_ws_bytes(const Geo& g, int backward) {
  const size_t es = g.bf16 ? 2 : 4;
  size_t tot = 0;
  if (backward != 1) {
    for (int i = 1; i < 5; i++) tot += align_up((size_t)g.H * kConvK[i] * g.H * es);
    if (backward == 2) {  // inference forward (save == NULL): the activations ping-pong inside the workspace
      tot += align_up((size_t)g.B * (g.Lout[0] + 2 * kPad) * g.H * es);
      tot += 2 * align_up((size_t)g.B * (g.Lout[1] + 2 * kPad) * g.H * es);
    }
  } else {
    for (int i = 1; i < 5; i++) tot += align_up((size_t)kConvS[i] * g.H * 2 * g.H * es);
    for (int i = 1; i < 5; i++) tot += align_up((size_t)g.B * (g.Lout[i] + 2 * kPad) * g.H * es);
    for (int i = 0; i < 4; i++) tot += align_up((size_t)g.B * g.Lout[i] * g.H * es);
    for (int i = 1; i < 5; i++) tot += align_up((size_t)g.H * kConvK[i] * g.H * 4);
  }
  return tot + 256;
}

int encoder_fwd(const Geo& g, const float* x, const cpcb200_encoder_params* p, float* z, void* save, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  if (g.bf16) return encoder_fwd_t<bf16>(g, x, p, z, save, ws, ws_bytes, st);
  return encoder_fwd_t<float>(g, x, p, z, save, ws, ws_bytes, st);
}
int encoder_bwd(const Geo& g, const float* x, const cpcb200_encoder_params* p, const float* dz, const void* save,
                const cpcb200_encoder_params* gr, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.bf16) return encoder_bwd_t<bf16>(g, x, p, dz, save, gr, ws, ws_bytes, st);
  return encoder_bwd_t<float>(g, x, p, dz, save, gr, ws, ws_bytes, st);
}

}  // namespace cpcb200
