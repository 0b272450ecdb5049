// conv0_mma.cu - conv0 + ChannelNorm + ReLU (cpc/model.py:100) for the bf16 path, forward and backward, with the
// 10-tap convolution on mma.sync instead of 80 FFMA per lane per frame.
//
// A tile = 16 consecutive frames of one window x all H channels, owned by one warp:
//   u[16 x H] = X[16 x 16] . W0^T[16 x H]     taps 0..9 = the 10 samples of the frame, tap 10 = 1.0 (carries the
//                                             bias), taps 11..15 = 0.  X is split x = hi + lo (two bf16 MMAs) so
//                                             the waveform keeps ~16 mantissa bits; W0 is bf16 like every weight
//                                             of the tensor-core path.
// Accumulator layout (m16n8k16): thread (g, t) holds rows g, g+8 and channels 8j+2t, 8j+2t+1 of every n-tile j,
// so the ChannelNorm statistics of a frame are an in-thread sum over 2H/8 values plus two shuffles.
// Tiles are staged through shared memory so that HBM sees 512-byte rows.
#include <cstdlib>
#include "common.cuh"

namespace cpcb200 {

namespace {

constexpr float kEps = 1e-5f;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// B fragments of W0^T for all n-tiles: b0 = taps (2t, 2t+1), b1 = taps (2t+8, 2t+9) with tap 10 = bias
template <int H>
__device__ __forceinline__ void load_wfrags(const float* __restrict__ w, const float* __restrict__ bias, int g, int t,
                                            uint32_t (&wb)[H / 8][2]) {
#pragma unroll
  for (int j = 0; j < H / 8; j++) {
    const int c = 8 * j + g;
    wb[j][0] = pack_bf16(__ldg(w + c * 10 + 2 * t), __ldg(w + c * 10 + 2 * t + 1));
    float lo = 0.f, hi = 0.f;
    if (t == 0) { lo = __ldg(w + c * 10 + 8); hi = __ldg(w + c * 10 + 9); }
    else if (t == 1) { lo = __ldg(bias + c); }
    wb[j][1] = pack_bf16(lo, hi);
  }
}

// the same fragments parked in shared memory ([n-tile][lane] -> one 8-byte load per n-tile, conflict-free): keeps
// 2*H/8 registers free so that three CTAs fit an SM
template <int H>
__device__ __forceinline__ void stage_wfrags(const float* __restrict__ w, const float* __restrict__ bias, uint2* wsm) {
  for (int i = threadIdx.x; i < (H / 8) * 32; i += blockDim.x) {
    const int j = i >> 5, lane = i & 31, g = lane >> 2, t = lane & 3;
    const int c = 8 * j + g;
    uint2 v;
    v.x = pack_bf16(__ldg(w + c * 10 + 2 * t), __ldg(w + c * 10 + 2 * t + 1));
    float lo = 0.f, hi = 0.f;
    if (t == 0) { lo = __ldg(w + c * 10 + 8); hi = __ldg(w + c * 10 + 9); }
    else if (t == 1) { lo = __ldg(bias + c); }
    v.y = pack_bf16(lo, hi);
    wsm[i] = v;
  }
}

// A fragments (hi and lo parts) of the 16 frames starting at f0 of window xb
__device__ __forceinline__ void load_xfrags(const float* __restrict__ xb, int L, int f0, int g, int t, uint32_t (&ah)[4],
                                            uint32_t (&al)[4]) {
  float v[2][4];  // [row half][tap 2t, 2t+1, 2t+8, 2t+9]
#pragma unroll
  for (int hf = 0; hf < 2; hf++) {
    const int s0 = 5 * (f0 + g + 8 * hf) - 3;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int s = s0 + 2 * t + q;
      v[hf][q] = (s >= 0 && s < L) ? __ldg(xb + s) : 0.f;
    }
    v[hf][2] = v[hf][3] = 0.f;
    if (t == 0) {
#pragma unroll
      for (int q = 0; q < 2; q++) { const int s = s0 + 8 + q; v[hf][2 + q] = (s >= 0 && s < L) ? __ldg(xb + s) : 0.f; }
    } else if (t == 1) {
      v[hf][2] = 1.f;  // bias tap
    }
  }
  float h[2][4];
#pragma unroll
  for (int hf = 0; hf < 2; hf++)
#pragma unroll
    for (int q = 0; q < 4; q++) h[hf][q] = bf16_round(v[hf][q]);
  ah[0] = pack_bf16(h[0][0], h[0][1]); ah[1] = pack_bf16(h[1][0], h[1][1]);
  ah[2] = pack_bf16(h[0][2], h[0][3]); ah[3] = pack_bf16(h[1][2], h[1][3]);
  al[0] = pack_bf16(v[0][0] - h[0][0], v[0][1] - h[0][1]); al[1] = pack_bf16(v[1][0] - h[1][0], v[1][1] - h[1][1]);
  al[2] = pack_bf16(v[0][2] - h[0][2], v[0][3] - h[0][3]); al[3] = pack_bf16(v[1][2] - h[1][2], v[1][3] - h[1][3]);
}

// row statistics (model.py:52-54) of the two rows a thread holds: mean, rstd with unbiased variance
template <int NT>
__device__ __forceinline__ void tile_stats(const float (&acc)[NT][4], int H, float (&mean)[2], float (&rstd)[2]) {
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NT; j++) { s[0] += acc[j][0] + acc[j][1]; s[1] += acc[j][2] + acc[j][3]; }
#pragma unroll
  for (int hf = 0; hf < 2; hf++) {
    s[hf] += __shfl_xor_sync(0xffffffffu, s[hf], 1);
    s[hf] += __shfl_xor_sync(0xffffffffu, s[hf], 2);
    mean[hf] = s[hf] / (float)H;
  }
  float q[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NT; j++) {
    float d0 = acc[j][0] - mean[0], d1 = acc[j][1] - mean[0], d2 = acc[j][2] - mean[1], d3 = acc[j][3] - mean[1];
    q[0] = fmaf(d0, d0, q[0]); q[0] = fmaf(d1, d1, q[0]); q[1] = fmaf(d2, d2, q[1]); q[1] = fmaf(d3, d3, q[1]);
  }
#pragma unroll
  for (int hf = 0; hf < 2; hf++) {
    q[hf] += __shfl_xor_sync(0xffffffffu, q[hf], 1);
    q[hf] += __shfl_xor_sync(0xffffffffu, q[hf], 2);
    rstd[hf] = rsqrtf(q[hf] / (float)(H - 1) + kEps);
  }
}

// copy a staged [16][H] bf16 tile (row stride RS bytes) to 16 consecutive global rows of H channels
// (nrows < 16: the last tile of a window whose frame count is not a multiple of 16 - rows past the end are not stored)
template <int H>
__device__ __forceinline__ void tile_to_global(const unsigned char* st, bf16* __restrict__ dst, int lane, int nrows = 16) {
  constexpr int SEGS = H / 8, RS = 2 * H + 16;
#pragma unroll
  for (int it = 0; it < (16 * SEGS) / 32; it++) {
    const int flat = it * 32 + lane, r = flat / SEGS, sg = flat - r * SEGS;
    const uint4 v = *reinterpret_cast<const uint4*>(st + r * RS + sg * 16);
    if (r < nrows) *reinterpret_cast<uint4*>(dst + (size_t)r * H + sg * 8) = v;
  }
}
// rows past the end of the window read as zero: a zero gradient row contributes nothing to any sum of the backward pass
template <int H>
__device__ __forceinline__ void tile_from_global(unsigned char* st, const bf16* __restrict__ src, int lane, int nrows = 16) {
  constexpr int SEGS = H / 8, RS = 2 * H + 16;
#pragma unroll
  for (int it = 0; it < (16 * SEGS) / 32; it++) {
    const int flat = it * 32 + lane, r = flat / SEGS, sg = flat - r * SEGS;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < nrows) v = *reinterpret_cast<const uint4*>(src + (size_t)r * H + sg * 8);
    *reinterpret_cast<uint4*>(st + r * RS + sg * 16) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward: x (B, L) fp32 -> y0 (B, kPad + L0 + kPad, H) bf16 (pads zeroed here)
// ---------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128, 3) conv0_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ gam,
                                                             const float* __restrict__ bet, bf16* __restrict__ y, int B, int L,
                                                             int L0) {
  pdl_wait();
  pdl_trigger();
  constexpr int NT = H / 8, RS = 2 * H + 16;
  extern __shared__ __align__(16) unsigned char sm[];
  float* gb = reinterpret_cast<float*>(sm);                       // gamma[H], beta[H]
  uint2* wsm = reinterpret_cast<uint2*>(sm + 2 * H * 4);          // weight fragments [NT][32]
  unsigned char* stage = sm + 2 * H * 4 + NT * 32 * 8 + (threadIdx.x >> 5) * (16 * RS);
  for (int i = threadIdx.x; i < H; i += blockDim.x) { gb[i] = gam[i]; gb[H + i] = bet[i]; }
  stage_wfrags<H>(w, bias, wsm);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int tpw = (L0 + 15) / 16;  // tiles per window (the last one may be partial: L is not always a multiple of 160)
  const long long Lp0 = L0 + 2 * kPad;
  for (int tile = warp; tile < B * tpw; tile += nwarps) {
    const int b = tile / tpw, f0 = (tile - b * tpw) * 16;
    uint32_t ah[4], al[4];
    load_xfrags(x + (long long)b * L, L, f0, g, t, ah, al);
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const uint2 wv = wsm[j * 32 + lane];
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      mma16816(acc[j], ah, wv.x, wv.y);
      mma16816(acc[j], al, wv.x, wv.y);
    }
    float mean[2], rstd[2];
    tile_stats<NT>(acc, H, mean, rstd);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const int c = 8 * j + 2 * t;
      const float2 g2 = *reinterpret_cast<const float2*>(gb + c), b2 = *reinterpret_cast<const float2*>(gb + H + c);
      const float y0 = fmaxf(fmaf((acc[j][0] - mean[0]) * rstd[0], g2.x, b2.x), 0.f);
      const float y1 = fmaxf(fmaf((acc[j][1] - mean[0]) * rstd[0], g2.y, b2.y), 0.f);
      const float y2 = fmaxf(fmaf((acc[j][2] - mean[1]) * rstd[1], g2.x, b2.x), 0.f);
      const float y3 = fmaxf(fmaf((acc[j][3] - mean[1]) * rstd[1], g2.y, b2.y), 0.f);
      *reinterpret_cast<uint32_t*>(stage + g * RS + c * 2) = pack_bf16(y0, y1);
      *reinterpret_cast<uint32_t*>(stage + (g + 8) * RS + c * 2) = pack_bf16(y2, y3);
    }
    __syncwarp();
    bf16* dst = y + ((long long)b * Lp0 + kPad + f0) * H;
    tile_to_global<H>(stage, dst, lane, L0 - f0);
    if (f0 == 0 || f0 + 16 >= L0) {  // zero rows around the window
      bf16* pad = y + ((long long)b * Lp0 + (f0 == 0 ? 0 : kPad + L0)) * H;
      for (int i = lane; i < kPad * H / 8; i += 32) reinterpret_cast<uint4*>(pad)[i] = make_uint4(0, 0, 0, 0);
      if (f0 == 0 && f0 + 16 >= L0) {
        bf16* pad2 = y + ((long long)b * Lp0 + kPad + L0) * H;
        for (int i = lane; i < kPad * H / 8; i += 32) reinterpret_cast<uint4*>(pad2)[i] = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward, H = 512 (BASELINE config 5): one warp still owns a 16-frame tile and all channels, but 512 channels of fp32
// accumulators do not fit the register file, so the product is formed twice, 256 channels at a time: pass A only accumulates
// the row sums (sum u, sum u^2), pass B recomputes, normalises and stages.  conv0 is HBM-bound (the MMAs are ~17 % of the
// tensor pipe at H = 256), the second product is cheaper than a round trip through shared memory.
// ---------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128, 2) conv0_fwd_mma_wide_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, const float* __restrict__ gam,
                                                                  const float* __restrict__ bet, bf16* __restrict__ y, int B, int L,
                                                                  int L0) {
  pdl_wait();
  pdl_trigger();
  constexpr int NT = H / 8, NTH = NT / 2, RS = 2 * H + 16;
  extern __shared__ __align__(16) unsigned char sm[];
  float* gb = reinterpret_cast<float*>(sm);
  uint2* wsm = reinterpret_cast<uint2*>(sm + 2 * H * 4);
  unsigned char* stage = sm + 2 * H * 4 + NT * 32 * 8 + (threadIdx.x >> 5) * (16 * RS);
  for (int i = threadIdx.x; i < H; i += blockDim.x) { gb[i] = gam[i]; gb[H + i] = bet[i]; }
  stage_wfrags<H>(w, bias, wsm);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int tpw = (L0 + 15) / 16;
  const long long Lp0 = L0 + 2 * kPad;
  for (int tile = warp; tile < B * tpw; tile += nwarps) {
    const int b = tile / tpw, f0 = (tile - b * tpw) * 16;
    uint32_t ah[4], al[4];
    load_xfrags(x + (long long)b * L, L, f0, g, t, ah, al);
    float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
#pragma unroll
      for (int j = 0; j < NTH; j++) {
        const uint2 wv = wsm[(hh * NTH + j) * 32 + lane];
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(a, ah, wv.x, wv.y);
        mma16816(a, al, wv.x, wv.y);
        s[0] += a[0] + a[1]; s[1] += a[2] + a[3];
        q[0] = fmaf(a[0], a[0], q[0]); q[0] = fmaf(a[1], a[1], q[0]);
        q[1] = fmaf(a[2], a[2], q[1]); q[1] = fmaf(a[3], a[3], q[1]);
      }
    }
    float mean[2], rstd[2];
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      s[hf] += __shfl_xor_sync(0xffffffffu, s[hf], 1); s[hf] += __shfl_xor_sync(0xffffffffu, s[hf], 2);
      q[hf] += __shfl_xor_sync(0xffffffffu, q[hf], 1); q[hf] += __shfl_xor_sync(0xffffffffu, q[hf], 2);
      mean[hf] = s[hf] / (float)H;
      rstd[hf] = rsqrtf(fmaxf((q[hf] - s[hf] * mean[hf]) / (float)(H - 1), 0.f) + kEps);  // unbiased variance, model.py:52-54
    }
    __syncwarp();
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
#pragma unroll
      for (int j = 0; j < NTH; j++) {
        const uint2 wv = wsm[(hh * NTH + j) * 32 + lane];
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(a, ah, wv.x, wv.y);
        mma16816(a, al, wv.x, wv.y);
        const int c = 8 * (hh * NTH + j) + 2 * t;
        const float2 g2 = *reinterpret_cast<const float2*>(gb + c), b2 = *reinterpret_cast<const float2*>(gb + H + c);
        const float y0 = fmaxf(fmaf((a[0] - mean[0]) * rstd[0], g2.x, b2.x), 0.f);
        const float y1 = fmaxf(fmaf((a[1] - mean[0]) * rstd[0], g2.y, b2.y), 0.f);
        const float y2 = fmaxf(fmaf((a[2] - mean[1]) * rstd[1], g2.x, b2.x), 0.f);
        const float y3 = fmaxf(fmaf((a[3] - mean[1]) * rstd[1], g2.y, b2.y), 0.f);
        *reinterpret_cast<uint32_t*>(stage + g * RS + c * 2) = pack_bf16(y0, y1);
        *reinterpret_cast<uint32_t*>(stage + (g + 8) * RS + c * 2) = pack_bf16(y2, y3);
      }
    }
    __syncwarp();
    bf16* dst = y + ((long long)b * Lp0 + kPad + f0) * H;
    tile_to_global<H>(stage, dst, lane, L0 - f0);
    if (f0 == 0 || f0 + 16 >= L0) {  // zero rows around the window
      bf16* pad = y + ((long long)b * Lp0 + (f0 == 0 ? 0 : kPad + L0)) * H;
      for (int i = lane; i < kPad * H / 8; i += 32) reinterpret_cast<uint4*>(pad)[i] = make_uint4(0, 0, 0, 0);
      if (f0 == 0 && f0 + 16 >= L0) {
        bf16* pad2 = y + ((long long)b * Lp0 + kPad + L0) * H;
        for (int i = lane; i < kPad * H / 8; i += 32) reinterpret_cast<uint4*>(pad2)[i] = make_uint4(0, 0, 0, 0);
      }
    }
    __syncwarp();  // the staging tile is rewritten by the next tile's pass B
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward part 1: recompute u, ChannelNorm+ReLU backward in place (dy0 -> du0), dbias0/dgamma0/dbeta0.
// The three per-channel sums over frames are column sums of bf16 tiles: computed with MMAs against a ones
// matrix (A = tile^T via ldmatrix.trans) and accumulated in shared memory.
// ---------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128, 2) conv0_bwd_du_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, const float* __restrict__ gam,
                                                                const float* __restrict__ bet, bf16* __restrict__ dy,
                                                                float* __restrict__ dbias, float* __restrict__ dgam,
                                                                float* __restrict__ dbet, int B, int L, int L0) {
  pdl_wait();
  pdl_trigger();
  constexpr int NT = H / 8, RS = 2 * H + 16, TILE = 16 * RS;
  extern __shared__ __align__(16) unsigned char sm[];
  float* gb = reinterpret_cast<float*>(sm);           // gamma[H], beta[H]
  float* accs = gb + 2 * H;                           // [3][H]: dbias, dgamma, dbeta
  uint2* wsm = reinterpret_cast<uint2*>(sm + 5 * H * 4);  // weight fragments [NT][32]
  unsigned char* wbase = sm + 5 * H * 4 + NT * 32 * 8 + (threadIdx.x >> 5) * (2 * TILE);
  unsigned char* t_a = wbase;                         // dy on input -> dv * xhat (in place)
  unsigned char* t_b = wbase + TILE;                  // dv = dy masked by the ReLU -> du (in place)
  for (int i = threadIdx.x; i < H; i += blockDim.x) { gb[i] = gam[i]; gb[H + i] = bet[i]; }
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) accs[i] = 0.f;
  stage_wfrags<H>(w, bias, wsm);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int tpw = (L0 + 15) / 16;
  const uint32_t ones = 0x3f803f80u;  // bf16 (1, 1)
  for (int tile = warp; tile < B * tpw; tile += nwarps) {
    const int b = tile / tpw, f0 = (tile - b * tpw) * 16;
    bf16* drow = dy + ((long long)b * L0 + f0) * H;
    tile_from_global<H>(t_a, drow, lane, L0 - f0);
    uint32_t ah[4], al[4];
    load_xfrags(x + (long long)b * L, L, f0, g, t, ah, al);
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const uint2 wv = wsm[j * 32 + lane];
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      mma16816(acc[j], ah, wv.x, wv.y);
      mma16816(acc[j], al, wv.x, wv.y);
    }
    float mean[2], rstd[2];
    tile_stats<NT>(acc, H, mean, rstd);
    __syncwarp();
    // pass 1: xhat (kept in acc), dv, dv*xhat tiles, row sums s1 = sum dx, s2 = sum dx*xhat
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const int c = 8 * j + 2 * t;
      const float2 g2 = *reinterpret_cast<const float2*>(gb + c), b2 = *reinterpret_cast<const float2*>(gb + H + c);
      const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(t_a + g * RS + c * 2);
      const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(t_a + (g + 8) * RS + c * 2);
      const float dyv[4] = {__low2float(d01), __high2float(d01), __low2float(d23), __high2float(d23)};
      const float gg[4] = {g2.x, g2.y, g2.x, g2.y}, bb[4] = {b2.x, b2.y, b2.x, b2.y};
      float dv[4], dvx[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int hf = e >> 1;
        const float xh = (acc[j][e] - mean[hf]) * rstd[hf];
        dv[e] = fmaf(xh, gg[e], bb[e]) > 0.f ? dyv[e] : 0.f;
        dvx[e] = dv[e] * xh;
        const float dx = dv[e] * gg[e];
        s1[hf] += dx;
        s2[hf] = fmaf(dx, xh, s2[hf]);
        acc[j][e] = xh;
      }
      *reinterpret_cast<uint32_t*>(t_b + g * RS + c * 2) = pack_bf16(dv[0], dv[1]);
      *reinterpret_cast<uint32_t*>(t_b + (g + 8) * RS + c * 2) = pack_bf16(dv[2], dv[3]);
      *reinterpret_cast<uint32_t*>(t_a + g * RS + c * 2) = pack_bf16(dvx[0], dvx[1]);   // over the dy this thread just read
      *reinterpret_cast<uint32_t*>(t_a + (g + 8) * RS + c * 2) = pack_bf16(dvx[2], dvx[3]);
    }
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      s1[hf] += __shfl_xor_sync(0xffffffffu, s1[hf], 1); s1[hf] += __shfl_xor_sync(0xffffffffu, s1[hf], 2);
      s2[hf] += __shfl_xor_sync(0xffffffffu, s2[hf], 1); s2[hf] += __shfl_xor_sync(0xffffffffu, s2[hf], 2);
      s1[hf] /= (float)H;
      s2[hf] /= (float)(H - 1);
    }
    __syncwarp();
    // column sums over the 16 frames of dv*xhat and dv:  [16 ch x 8] = tile^T[16 ch x 16 frames] . ones
#pragma unroll
    for (int m = 0; m < H / 16; m++) {
      const int mi = lane >> 3;
      const uint32_t off = ((mi >> 1) * 8 + (lane & 7)) * RS + (m * 16 + (mi & 1) * 8) * 2;
      uint32_t a[4];
      float cs[4];
      ldsm_x4_t(a, s_u32(t_a + off));
      cs[0] = cs[1] = cs[2] = cs[3] = 0.f; mma16816(cs, a, ones, ones);
      if (t == 0) { atomicAdd(&accs[H + m * 16 + g], cs[0]); atomicAdd(&accs[H + m * 16 + g + 8], cs[2]); }
      ldsm_x4_t(a, s_u32(t_b + off));
      cs[0] = cs[1] = cs[2] = cs[3] = 0.f; mma16816(cs, a, ones, ones);
      if (t == 0) { atomicAdd(&accs[2 * H + m * 16 + g], cs[0]); atomicAdd(&accs[2 * H + m * 16 + g + 8], cs[2]); }
    }
    __syncwarp();
    // pass 2: du = rstd * (dx - s1 - xhat * s2), written in place over dv
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const int c = 8 * j + 2 * t;
      const float2 g2 = *reinterpret_cast<const float2*>(gb + c);
      const __nv_bfloat162 v01 = *reinterpret_cast<const __nv_bfloat162*>(t_b + g * RS + c * 2);
      const __nv_bfloat162 v23 = *reinterpret_cast<const __nv_bfloat162*>(t_b + (g + 8) * RS + c * 2);
      const float o0 = rstd[0] * (__low2float(v01) * g2.x - s1[0] - acc[j][0] * s2[0]);
      const float o1 = rstd[0] * (__high2float(v01) * g2.y - s1[0] - acc[j][1] * s2[0]);
      const float o2 = rstd[1] * (__low2float(v23) * g2.x - s1[1] - acc[j][2] * s2[1]);
      const float o3 = rstd[1] * (__high2float(v23) * g2.y - s1[1] - acc[j][3] * s2[1]);
      *reinterpret_cast<uint32_t*>(t_b + g * RS + c * 2) = pack_bf16(o0, o1);
      *reinterpret_cast<uint32_t*>(t_b + (g + 8) * RS + c * 2) = pack_bf16(o2, o3);
    }
    __syncwarp();
    tile_to_global<H>(t_b, drow, lane, L0 - f0);
#pragma unroll
    for (int m = 0; m < H / 16; m++) {  // column sums of du
      const int mi = lane >> 3;
      uint32_t a[4];
      float cs[4] = {0.f, 0.f, 0.f, 0.f};
      ldsm_x4_t(a, s_u32(t_b + ((mi >> 1) * 8 + (lane & 7)) * RS + (m * 16 + (mi & 1) * 8) * 2));
      mma16816(cs, a, ones, ones);
      if (t == 0) { atomicAdd(&accs[m * 16 + g], cs[0]); atomicAdd(&accs[m * 16 + g + 8], cs[2]); }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dbias + i, accs[i]); atomicAdd(dgam + i, accs[H + i]); atomicAdd(dbet + i, accs[2 * H + i]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward part 2: dW0[c][tap] += sum_frames du0[frame][c] * x[5 frame - 3 + tap]
//   D[16 ch x 8 taps] (x2 n-tiles: taps 0-7, 8-15) += du^T[16 ch x 16 frames] . X[16 frames x taps]
// accumulated in registers over all the tiles of a warp (x in bf16: weight-gradient precision).
// ---------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128) conv0_wgrad_mma_kernel(const float* __restrict__ x, const bf16* __restrict__ du,
                                                               float* __restrict__ dw, int B, int L, int L0) {
  pdl_wait();
  pdl_trigger();
  constexpr int MT = H / 16, RS = 2 * H + 16, TILE = 16 * RS;
  extern __shared__ __align__(16) unsigned char sm[];
  float* accs = reinterpret_cast<float*>(sm);  // [H][10]
  unsigned char* stage = sm + H * 10 * 4 + (threadIdx.x >> 5) * TILE;
  for (int i = threadIdx.x; i < H * 10; i += blockDim.x) accs[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float acc[MT][2][4];
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[m][n][e] = 0.f;
  const int tpw = (L0 + 15) / 16;
  for (int tile = warp; tile < B * tpw; tile += nwarps) {
    const int b = tile / tpw, f0 = (tile - b * tpw) * 16;
    tile_from_global<H>(stage, du + ((long long)b * L0 + f0) * H, lane, L0 - f0);
    // B fragments: X[k = frame][n = tap]: b0 = frames (2t, 2t+1), b1 = frames (2t+8, 2t+9); n = g (taps 0-7) / g+8
    const float* xb = x + (long long)b * L;
    uint32_t bx[2][2];
#pragma unroll
    for (int n = 0; n < 2; n++) {
      const int tap = g + 8 * n;
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int fr = f0 + 2 * t + (q & 1) + 8 * (q >> 1);
        const int s = 5 * fr - 3 + tap;
        v[q] = (tap < 10 && s >= 0 && s < L) ? __ldg(xb + s) : 0.f;
      }
      bx[n][0] = pack_bf16(v[0], v[1]);
      bx[n][1] = pack_bf16(v[2], v[3]);
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MT; m++) {
      const int mi = lane >> 3;
      uint32_t a[4];
      ldsm_x4_t(a, s_u32(stage + ((mi >> 1) * 8 + (lane & 7)) * RS + (m * 16 + (mi & 1) * 8) * 2));
      mma16816(acc[m][0], a, bx[0][0], bx[0][1]);
      mma16816(acc[m][1], a, bx[1][0], bx[1][1]);
    }
    __syncwarp();
  }
  // D fragment: rows = channels 16m + g (+8), cols = taps 8n + 2t (+1)
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int c = 16 * m + g + 8 * (e >> 1), tap = 8 * n + 2 * t + (e & 1);
        if (tap < 10) atomicAdd(&accs[c * 10 + tap], acc[m][n][e]);
      }
  __syncthreads();
  for (int i = threadIdx.x; i < H * 10; i += blockDim.x) atomicAdd(dw + i, accs[i]);
}


// ---------------------------------------------------------------------------------------------------------
// backward, second generation (L0 % 64 == 0).  A CTA owns a tile of 64 frames x H channels; warp w owns the channel
// slice [64w, 64w+64) of all 64 frames (4 m-tiles x 8 n-tiles of m16n8k16).  Compared with the warp-per-16-frames
// kernel above: (i) the three per-channel sums (dgamma, dbeta, and dW0/dbias) accumulate in REGISTERS over all the
// tiles of the CTA - no ones-matrix MMAs, no shared-memory atomics; (ii) the dy tile and the waveform samples of
// the next tile are prefetched with cp.async while the current one is processed; (iii) the weight gradient
// dW0[c][tap] += sum_f du[f][c] x[5f-3+tap] is taken from the du tile while it is still in shared memory
// (A = du^T by ldmatrix.trans, B = waveform samples; tap 10 = 1.0 yields dbias0), so du never goes to HBM.
// Row statistics cross the NW warps through two small shared arrays (sum / sum of squares, then s1 / s2).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int kFT = 64;    // frames per tile
constexpr int kXS = 336;   // floats reserved for the 5*64+5 samples of a tile
constexpr int kCW = 64;    // channels per warp.  Measured on B200 (r2k): 32 channels per warp (8 warps per CTA, 128 registers, 16 warps
                           // per SM instead of 8) is SLOWER - 0.152 vs 0.130 ms - because the per-row work (sample fragments,
                           // statistics, s1 / s2 exchanges) is replicated in every warp; 64 stays.
constexpr int kNJ = kCW / 8, kNM = kCW / 16;

template <int H>
constexpr size_t bwd2_smem() {
  return (size_t)2 * H * 4 + (size_t)(H / 8) * 32 * 8 + (size_t)2 * (H / kCW) * kFT * 8 + 2 * kXS * 4 + (size_t)2 * kFT * (2 * H + 16);
}

template <int H>
__global__ void __launch_bounds__((H / kCW) * 32, (H <= 256 ? 2 : 1))
conv0_bwd2_mma_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ gam, const float* __restrict__ bet, bf16* __restrict__ dy,
                      float* __restrict__ dw, float* __restrict__ dbias, float* __restrict__ dgam, float* __restrict__ dbet,
                      int B, int L, int L0) {
  pdl_wait();
  pdl_trigger();
  constexpr int NW = H / kCW, NTHR = NW * 32, RS = 2 * H + 16, TILE = kFT * RS, SEGS = H / 8;
  extern __shared__ __align__(16) unsigned char sm[];
  float* gb = reinterpret_cast<float*>(sm);                                    // gamma[H], beta[H]
  uint2* wsm = reinterpret_cast<uint2*>(sm + 2 * H * 4);                       // weight fragments [H/8][32]
  float2* part0 = reinterpret_cast<float2*>(sm + 2 * H * 4 + SEGS * 32 * 8);   // [NW][64] (sum u, sum u^2)
  float2* part1 = part0 + NW * kFT;                                            // [NW][64] (s1, s2)
  float* xs = reinterpret_cast<float*>(part1 + NW * kFT);                      // [2][kXS]
  unsigned char* tiles = reinterpret_cast<unsigned char*>(xs + 2 * kXS);       // [2][TILE]
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < H; i += NTHR) { gb[i] = gam[i]; gb[H + i] = bet[i]; }
  stage_wfrags<H>(w, bias, wsm);

  const int tpw = L0 / kFT, ntiles = B * tpw;
  auto issue = [&](int tile, int buf) {
    const int b = tile / tpw, f0 = (tile - b * tpw) * kFT;
    const bf16* src = dy + ((long long)b * L0 + f0) * H;
    const uint32_t dst = s_u32(tiles + buf * TILE);
    for (int i = tid; i < kFT * SEGS; i += NTHR) {
      const int r = i / SEGS, sg = i - r * SEGS;
      cp_async16(dst + r * RS + sg * 16, src + (size_t)r * H + sg * 8);
    }
    const float* xb = x + (long long)b * L;
    const int s0 = 5 * f0 - 3;
    for (int i = tid; i < 5 * kFT + 5; i += NTHR) {
      const int sidx = s0 + i;
      if (sidx >= 0 && sidx < L) cp_async4(s_u32(xs + buf * kXS + i), xb + sidx);
      else xs[buf * kXS + i] = 0.f;
    }
    cp_async_commit();
  };

  float cdg[kNJ][2], cdb[kNJ][2], wacc[kNM][2][4];
#pragma unroll
  for (int j = 0; j < kNJ; j++) cdg[j][0] = cdg[j][1] = cdb[j][0] = cdb[j][1] = 0.f;
#pragma unroll
  for (int m = 0; m < kNM; m++)
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
      for (int e = 0; e < 4; e++) wacc[m][n][e] = 0.f;

  int tile = blockIdx.x, it = 0;
  if (tile < ntiles) issue(tile, 0);
  for (; tile < ntiles; tile += gridDim.x, it++) {
    const int buf = it & 1;
    cp_async_wait_all();
    __syncthreads();  // (A) this tile is visible; every warp is done with the other buffer
    if (tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, buf ^ 1);
    unsigned char* tl = tiles + buf * TILE;
    const float* xt = xs + buf * kXS;
    const int b = tile / tpw, f0 = (tile - b * tpw) * kFT;

    // ---- u = X . W0^T for the warp's kCW channels -------------------------------------------------------
    float acc[4][kNJ][4];
    {
      uint32_t ah[4][4], al[4][4];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        float v[2][4], h[2][4];
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const int base = 5 * (16 * m + g + 8 * hf);
          v[hf][0] = xt[base + 2 * t]; v[hf][1] = xt[base + 2 * t + 1];
          v[hf][2] = t == 0 ? xt[base + 8] : (t == 1 ? 1.f : 0.f);
          v[hf][3] = t == 0 ? xt[base + 9] : 0.f;
#pragma unroll
          for (int q = 0; q < 4; q++) h[hf][q] = bf16_round(v[hf][q]);
        }
        ah[m][0] = pack_bf16(h[0][0], h[0][1]); ah[m][1] = pack_bf16(h[1][0], h[1][1]);
        ah[m][2] = pack_bf16(h[0][2], h[0][3]); ah[m][3] = pack_bf16(h[1][2], h[1][3]);
        al[m][0] = pack_bf16(v[0][0] - h[0][0], v[0][1] - h[0][1]); al[m][1] = pack_bf16(v[1][0] - h[1][0], v[1][1] - h[1][1]);
        al[m][2] = pack_bf16(v[0][2] - h[0][2], v[0][3] - h[0][3]); al[m][3] = pack_bf16(v[1][2] - h[1][2], v[1][3] - h[1][3]);
      }
#pragma unroll
      for (int j = 0; j < kNJ; j++) {
        const uint2 wv = wsm[(kNJ * wp + j) * 32 + lane];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
          mma16816(acc[m][j], ah[m], wv.x, wv.y);
          mma16816(acc[m][j], al[m], wv.x, wv.y);
        }
      }
    }
    // ---- row statistics over all H channels (model.py:52-54) ----------------------------------------------
    float rstd[4][2], nmr[4][2];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < kNJ; j++) {
          const float a0 = acc[m][j][2 * hf], a1 = acc[m][j][2 * hf + 1];
          s += a0 + a1; q = fmaf(a0, a0, q); q = fmaf(a1, a1, q);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 1); q += __shfl_xor_sync(0xffffffffu, q, 2);
        if (t == 0) part0[wp * kFT + 16 * m + g + 8 * hf] = make_float2(s, q);
      }
    __syncthreads();  // (B)
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) { const float2 p2 = part0[ww * kFT + 16 * m + g + 8 * hf]; s += p2.x; q += p2.y; }
        const float mean = s / (float)H;
        const float var = fmaxf((q - s * mean) / (float)(H - 1), 0.f);
        rstd[m][hf] = rsqrtf(var + kEps);
        nmr[m][hf] = -mean * rstd[m][hf];
      }
    // ---- pass 1: xhat (kept in acc), dgamma / dbeta partial sums, row sums s1 = sum dx, s2 = sum dx*xhat ----
    float s1[4][2], s2[4][2];
#pragma unroll
    for (int m = 0; m < 4; m++) s1[m][0] = s1[m][1] = s2[m][0] = s2[m][1] = 0.f;
#pragma unroll
    for (int j = 0; j < kNJ; j++) {
      const int c = kCW * wp + 8 * j + 2 * t;
      const float2 g2 = *reinterpret_cast<const float2*>(gb + c), b2 = *reinterpret_cast<const float2*>(gb + H + c);
#pragma unroll
      for (int m = 0; m < 4; m++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const __nv_bfloat162 d2 = *reinterpret_cast<const __nv_bfloat162*>(tl + (16 * m + g + 8 * hf) * RS + c * 2);
          const float dyv[2] = {__low2float(d2), __high2float(d2)}, gg[2] = {g2.x, g2.y}, bb[2] = {b2.x, b2.y};
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const float xh = fmaf(acc[m][j][2 * hf + q], rstd[m][hf], nmr[m][hf]);
            const float dv = fmaf(xh, gg[q], bb[q]) > 0.f ? dyv[q] : 0.f;
            cdg[j][q] = fmaf(dv, xh, cdg[j][q]);
            cdb[j][q] += dv;
            const float dx = dv * gg[q];
            s1[m][hf] += dx;
            s2[m][hf] = fmaf(dx, xh, s2[m][hf]);
            acc[m][j][2 * hf + q] = xh;
          }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        float a = s1[m][hf], c2 = s2[m][hf];
        a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
        c2 += __shfl_xor_sync(0xffffffffu, c2, 1); c2 += __shfl_xor_sync(0xffffffffu, c2, 2);
        if (t == 0) part1[wp * kFT + 16 * m + g + 8 * hf] = make_float2(a, c2);
      }
    __syncthreads();  // (C)
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        float a = 0.f, c2 = 0.f;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) { const float2 p2 = part1[ww * kFT + 16 * m + g + 8 * hf]; a += p2.x; c2 += p2.y; }
        s1[m][hf] = rstd[m][hf] * a / (float)H;          // rstd * mean(dx)
        s2[m][hf] = rstd[m][hf] * c2 / (float)(H - 1);   // rstd * sum(dx xhat) / (H-1)
      }
    // ---- pass 2: du = rstd * (dx - s1 - xhat * s2), in place over dy (this warp's columns only) -----------
#pragma unroll
    for (int j = 0; j < kNJ; j++) {
      const int c = kCW * wp + 8 * j + 2 * t;
      const float2 g2 = *reinterpret_cast<const float2*>(gb + c), b2 = *reinterpret_cast<const float2*>(gb + H + c);
      __nv_bfloat162 d2[4][2];
#pragma unroll
      for (int m = 0; m < 4; m++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) d2[m][hf] = *reinterpret_cast<const __nv_bfloat162*>(tl + (16 * m + g + 8 * hf) * RS + c * 2);
#pragma unroll
      for (int m = 0; m < 4; m++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const float dyv[2] = {__low2float(d2[m][hf]), __high2float(d2[m][hf])}, gg[2] = {g2.x, g2.y}, bb[2] = {b2.x, b2.y};
          float o[2];
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const float xh = acc[m][j][2 * hf + q];
            const float dv = fmaf(xh, gg[q], bb[q]) > 0.f ? dyv[q] : 0.f;
            o[q] = fmaf(-xh, s2[m][hf], fmaf(dv * gg[q], rstd[m][hf], -s1[m][hf]));
          }
          *reinterpret_cast<uint32_t*>(tl + (16 * m + g + 8 * hf) * RS + c * 2) = pack_bf16(o[0], o[1]);
        }
    }
    __syncwarp();
    // ---- dW0 / dbias0 from the du tile: D[16 ch x taps] += du^T[16 ch x 16 frames] . X[16 frames x taps] ----
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      uint32_t bx[2][2];
#pragma unroll
      for (int n = 0; n < 2; n++) {
        const int tap = g + 8 * n;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int fr = 16 * ks + 2 * t + (q & 1) + 8 * (q >> 1);
          v[q] = tap < 10 ? xt[5 * fr + tap] : (tap == 10 ? 1.f : 0.f);
        }
        bx[n][0] = pack_bf16(v[0], v[1]);
        bx[n][1] = pack_bf16(v[2], v[3]);
      }
#pragma unroll
      for (int m = 0; m < kNM; m++) {
        const int mi = lane >> 3;
        uint32_t a[4];
        ldsm_x4_t(a, s_u32(tl + (16 * ks + (mi >> 1) * 8 + (lane & 7)) * RS + (kCW * wp + 16 * m + (mi & 1) * 8) * 2));
        mma16816(wacc[m][0], a, bx[0][0], bx[0][1]);
        mma16816(wacc[m][1], a, bx[1][0], bx[1][1]);
      }
    }
    // du0 has no consumer besides the weight gradient taken above (conv0 is the first layer: there is no data gradient),
    // so the tile is NOT written back: the kernel's HBM traffic is one read of dy0.  Barrier (A) of the next iteration
    // orders these ldmatrix reads against the prefetch that reuses the buffer.
  }
  cp_async_wait_all();
  // ---- flush the register accumulators -------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < kNJ; j++)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      float a = cdg[j][q], c2 = cdb[j][q];
      a += __shfl_xor_sync(0xffffffffu, a, 4); a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
      c2 += __shfl_xor_sync(0xffffffffu, c2, 4); c2 += __shfl_xor_sync(0xffffffffu, c2, 8); c2 += __shfl_xor_sync(0xffffffffu, c2, 16);
      if (g == 0) { const int c = kCW * wp + 8 * j + 2 * t + q; atomicAdd(dgam + c, a); atomicAdd(dbet + c, c2); }
    }
#pragma unroll
  for (int m = 0; m < kNM; m++)
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int c = kCW * wp + 16 * m + g + 8 * (e >> 1), tap = 8 * n + 2 * t + (e & 1);
        if (tap < 10) atomicAdd(dw + c * 10 + tap, wacc[m][n][e]);
        else if (tap == 10) atomicAdd(dbias + c, wacc[m][n][e]);
      }
}

template <int H>
int launch_all_fwd(const float* x, const float* w, const float* bias, const float* gam, const float* bet, bf16* y, int B, int L,
                   int L0, cudaStream_t st) {
  const size_t smem = 2 * H * 4 + (H / 8) * 32 * 8 + 4 * 16 * (2 * H + 16);
  int blocks = (B * ((L0 + 15) / 16) + 3) / 4;
  if (blocks > 148 * 3) blocks = 148 * 3;
  CPC_CHECK_CUDA(cudaFuncSetAttribute(conv0_fwd_mma_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CPC_CHECK_CUDA(launch_k(conv0_fwd_mma_kernel<H>, dim3(blocks), dim3(128), smem, st, 1, x, w, bias, gam, bet, y, B, L, L0));
  CPC_LAUNCHED_N("conv0_fwd_mma", st);
  return 0;
}
template <int H>
int launch_all_bwd(const float* x, const float* w, const float* bias, const float* gam, const float* bet, bf16* dy, float* dw,
                   float* dbias, float* dgam, float* dbet, int B, int L, int L0, cudaStream_t st) {
  const size_t smem1 = 5 * H * 4 + (H / 8) * 32 * 8 + 4 * 2 * 16 * (2 * H + 16);
  const size_t smem2 = H * 10 * 4 + 4 * 16 * (2 * H + 16);
  static const bool gen1 = []() { const char* e = getenv("CPC_B200_CONV0_BWD_GEN"); return e && atoi(e) == 1; }();
  if (L0 % kFT == 0 && !gen1) {
    const size_t smem = bwd2_smem<H>();
    int blocks2 = B * (L0 / kFT);
    if (blocks2 > 148 * 2) blocks2 = 148 * 2;
    CPC_CHECK_CUDA(cudaFuncSetAttribute(conv0_bwd2_mma_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPC_CHECK_CUDA(launch_k(conv0_bwd2_mma_kernel<H>, dim3(blocks2), dim3((H / kCW) * 32), smem, st, 1, x, w, bias, gam, bet, dy, dw, dbias, dgam, dbet, B, L, L0));
    CPC_LAUNCHED_N("conv0_bwd2_mma", st);
    return 0;
  }
  int blocks = (B * ((L0 + 15) / 16) + 3) / 4;
  if (blocks > 148 * 3) blocks = 148 * 3;
  CPC_CHECK_CUDA(cudaFuncSetAttribute(conv0_bwd_du_mma_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  CPC_CHECK_CUDA(launch_k(conv0_bwd_du_mma_kernel<H>, dim3(blocks), dim3(128), smem1, st, 1, x, w, bias, gam, bet, dy, dbias, dgam, dbet, B, L, L0));
  CPC_LAUNCHED_N("conv0_bwd_du_mma", st);
  CPC_CHECK_CUDA(cudaFuncSetAttribute(conv0_wgrad_mma_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  CPC_CHECK_CUDA(launch_k(conv0_wgrad_mma_kernel<H>, dim3(blocks), dim3(128), smem2, st, 1, x, dy, dw, B, L, L0));
  CPC_LAUNCHED_N("conv0_wgrad_mma", st);
  return 0;
}

}  // namespace

bool conv0_mma_supported(int H) { return H == 64 || H == 128 || H == 256; }
// H = 512: forward with the two-pass kernel; backward with conv0_bwd2_mma (8 warps x 64 channels) when the frame count is a
// multiple of its 64-frame tiles, else the CUDA-core kernels of encoder.cu
bool conv0_mma_wide_fwd_supported(int H) { return H == 512; }
bool conv0_mma_wide_bwd_supported(int H, int L0) { return H == 512 && L0 % kFT == 0; }

int conv0_fwd_mma(const float* x, const float* w, const float* bias, const float* gam, const float* bet, bf16* y, int B, int L, int L0,
                  int H, cudaStream_t st) {
  if (H == 512) {
    constexpr int HW = 512;
    const size_t smem = 2 * HW * 4 + (HW / 8) * 32 * 8 + 4 * 16 * (2 * HW + 16);
    int blocks = (B * ((L0 + 15) / 16) + 3) / 4;
    if (blocks > 148 * 2) blocks = 148 * 2;
    CPC_CHECK_CUDA(cudaFuncSetAttribute(conv0_fwd_mma_wide_kernel<HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPC_CHECK_CUDA(launch_k(conv0_fwd_mma_wide_kernel<HW>, dim3(blocks), dim3(128), smem, st, 1, x, w, bias, gam, bet, y, B, L, L0));
    CPC_LAUNCHED_N("conv0_fwd_mma", st);
    return 0;
  }
  if (H == 256) return launch_all_fwd<256>(x, w, bias, gam, bet, y, B, L, L0, st);
  if (H == 128) return launch_all_fwd<128>(x, w, bias, gam, bet, y, B, L, L0, st);
  return launch_all_fwd<64>(x, w, bias, gam, bet, y, B, L, L0, st);
}
int conv0_bwd_mma(const float* x, const float* w, const float* bias, const float* gam, const float* bet, bf16* dy, float* dw,
                  float* dbias, float* dgam, float* dbet, int B, int L, int L0, int H, cudaStream_t st) {
  if (H == 512) {  // (conv0_mma_wide_bwd_supported: L0 % 64 == 0)
    constexpr int HW = 512;
    const size_t smem = bwd2_smem<HW>();
    int blocks2 = B * (L0 / kFT);
    if (blocks2 > 148) blocks2 = 148;
    CPC_CHECK_CUDA(cudaFuncSetAttribute(conv0_bwd2_mma_kernel<HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPC_CHECK_CUDA(launch_k(conv0_bwd2_mma_kernel<HW>, dim3(blocks2), dim3((HW / kCW) * 32), smem, st, 1, x, w, bias, gam, bet, dy, dw, dbias, dgam, dbet, B, L, L0));
    CPC_LAUNCHED_N("conv0_bwd2_mma", st);
    return 0;
  }
  if (H == 256) return launch_all_bwd<256>(x, w, bias, gam, bet, dy, dw, dbias, dgam, dbet, B, L, L0, st);
  if (H == 128) return launch_all_bwd<128>(x, w, bias, gam, bet, dy, dw, dbias, dgam, dbet, B, L, L0, st);
  return launch_all_bwd<64>(x, w, bias, gam, bet, dy, dw, dbias, dgam, dbet, B, L, L0, st);
}

}  // namespace cpcb200
