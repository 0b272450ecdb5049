// score_mma.cu - tensor-core scoring + InfoNCE for the bf16 path (reference: criterion.py:115-117, 207-217, 245-257).
//
// One WARP per anchor position p = (b, w), persistent over positions.  Per position the candidates are the N
// negatives z[ext[b, :, w]] (gathered row by row with cp.async, 512 B per instruction, double-buffered in chunks
// of 16 rows) followed by the K positives z[b, w+1 .. w+K] (contiguous rows).  All contractions run on
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32):
//   forward : logits[j][k] = cand_j . pred_k / H        M = candidates, N = heads (padded to 16), K = H
//             softmax / cross-entropy / argmax in the accumulator registers (reduction over candidates =
//             in-thread + 3 shuffles), nothing but (loss, correct, lse) per (p, k) leaves the SM.
//   backward: logits recomputed per chunk, G = (softmax - onehot) * dloss/(P*H) formed in registers;
//             dz[cand_j] += G^T . pred   (A operand = the logits accumulator fragments themselves) staged in
//                           shared memory and added to HBM by the TMA engine (cp.reduce.async.bulk .add.f32,
//                           one 1 KB row per instruction) instead of per-lane atomics;
//             dpred      += G . cand     accumulated over chunks in registers.
// The kernel is bound by the L2 gather (N x 512 B per position) and the scatter-add, not by the tensor pipe.
// Templates cover H = 64 / 128 / 256 columns per warp, N = 128 / 256 negatives, and a warp PAIR per position for a feature
// dim of 512 (each warp one half of every row; partial logits swapped per chunk).
#include <stdlib.h>

#include "common.cuh"

namespace cpcb200 {

namespace {

constexpr int CH = 16;    // candidate rows per gather chunk (CH/16 m-tiles); small chunks -> more warps per SM
constexpr int MPC = CH / 16;
// Template parameters of both kernels:
//   H     feature columns one WARP covers (64 / 128 / 256)
//   SPLIT warps per anchor position: 1, or 2 for feature dim 2*H = 512 (BASELINE config 5) - the two warps of a pair own the
//         low / high half of every row (their share of the contraction for the logits, their columns of dz and dpred) and swap
//         the 16 x 16 partial logits of every chunk through shared memory (one 64-thread named barrier per chunk)
//   NNEG  negative chunks per position: N = NNEG * CH negatives (8 -> 128, 16 -> 256)
constexpr int XCH = 2 * 8 * 32 * 4;  // bytes of the double-buffered partial-logit tile of one warp (SPLIT = 2)
__device__ __forceinline__ void pair_bar(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
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
// TMA-engine reduction: global[dst .. dst+bytes) += shared[src .. src+bytes)   (fp32 add)
__device__ __forceinline__ void bulk_reduce_add_f32(float* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// (P, N) <- (B, N, W): a position's N negative rows become contiguous
__global__ void transpose_ext_kernel(const int* __restrict__ ext, int* __restrict__ ext_t, int B, int N, int W) {
  pdl_wait();
  pdl_trigger();
  const long long n = (long long)B * N * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % N);
    const long long pw = i / N;
    const int w = (int)(pw % W), b = (int)(pw / W);
    ext_t[i] = ext[((long long)b * N + j) * W + w];
  }
}

template <int H> struct Cfg {
  static constexpr int RS = 2 * H + 16;        // bytes per staged row (16 B pad: conflict-free ldmatrix)
  static constexpr int SEGS = H / 8;           // 16-B segments per row
  static constexpr int KS = H / 16;            // k-steps over the feature dim
  static constexpr int PSM = 16 * RS;          // pred tile [16 heads][H]
  static constexpr int CBUF = CH * RS;         // one candidate chunk
};

// stage pred[p] (K x H bf16, contiguous) into psm[16][RS]; rows >= K stay zero
// (ld = elements between rows in global memory: H, or 2*H when a warp pair splits the row)
template <int H>
__device__ __forceinline__ void stage_pred(unsigned char* psm, const bf16* __restrict__ pp, int K, int lane, int ld) {
  constexpr int SEGS = Cfg<H>::SEGS, RS = Cfg<H>::RS;
  for (int i = lane; i < K * SEGS; i += 32) {
    const int k = i / SEGS, sg = i - k * SEGS;
    cp_async16(s_u32(psm + k * RS + sg * 16), pp + (size_t)k * ld + sg * 8);
  }
}
// gather `rows` candidate rows (row index held by lane r) into buf
template <int H>
__device__ __forceinline__ void gather_rows(unsigned char* buf, const bf16* __restrict__ z, int my_row, int rows, int lane, int ld) {
  constexpr int SEGS = Cfg<H>::SEGS, RS = Cfg<H>::RS;
#pragma unroll 4
  for (int it = 0; it < (CH * SEGS) / 32; it++) {
    const int flat = it * 32 + lane;
    const int r = flat / SEGS, sg = flat - r * SEGS;
    const int row = __shfl_sync(0xffffffffu, my_row, r);
    if (r < rows) cp_async16(s_u32(buf + r * RS + sg * 16), z + (size_t)row * ld + sg * 8);
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <int H, int SPLIT, int NNEG>
__global__ void __launch_bounds__(256) score_fwd_mma_kernel(const bf16* __restrict__ pred, const bf16* __restrict__ z,
                                                             const int* __restrict__ ext_t, float* __restrict__ lossbuf,
                                                             float* __restrict__ corrbuf, float* __restrict__ lsebuf, int B,
                                                             int S, int W, int K, int N, int warps_per_cta) {
  pdl_wait();
  pdl_trigger();
  using C = Cfg<H>;
  constexpr int NMT = NNEG * MPC;  // m-tiles of negatives; m-tile NMT holds the positives
  constexpr int HT = H * SPLIT;    // the feature dim
  constexpr int PER_WARP = C::PSM + 2 * C::CBUF + (SPLIT == 2 ? XCH : 0);
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int half = SPLIT == 2 ? (warp & 1) : 0, slot = warp / SPLIT, apc = warps_per_cta / SPLIT;
  unsigned char* psm = sm + (size_t)warp * PER_WARP;
  unsigned char* cbuf = psm + C::PSM;
  float* xmine = reinterpret_cast<float*>(cbuf + 2 * C::CBUF);
  const float* xpeer = reinterpret_cast<const float*>(sm + (size_t)(warp ^ 1) * PER_WARP + C::PSM + 2 * C::CBUF);
  for (int i = lane; i < C::PSM / 16; i += 32) reinterpret_cast<uint4*>(psm)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const int P = B * W;
  constexpr int nneg = NNEG;
  constexpr int nchunks = nneg + 1;
  const float invH = 1.f / (float)HT;
  const bf16* zh = z + half * H;
  // (a pair walks the positions together: its two warps meet at a named barrier in every chunk.  The exchange tile is double
  // buffered over ALL chunks of the walk - the chunk count per position is odd, xph carries the parity across positions -
  // so that a buffer is rewritten only after a barrier that follows the partner's read of it)
  int xph = 0;
  for (int p = blockIdx.x * apc + slot; p < P; p += gridDim.x * apc, xph ^= (nchunks & 1)) {
    const int b = p / W, w = p - b * W;
    stage_pred<H>(psm, pred + (size_t)p * K * HT + half * H, K, lane, HT);
    {
      const int row = lane < CH ? ext_t[(size_t)p * N + lane] : 0;
      gather_rows<H>(cbuf, zh, row, CH, lane, HT);
    }
    cp_async_commit();
    uint32_t bfr[C::KS][4];
    float acc[NMT + 1][2][4];
#pragma unroll
    for (int m = 0; m < NMT + 1; m++)
#pragma unroll
      for (int n = 0; n < 2; n++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[m][n][e] = 0.f;

#pragma unroll
    for (int c = 0; c < NNEG + 1; c++) {
      if (c < nchunks) {
        unsigned char* cur = cbuf + (c & 1) * C::CBUF;
        if (c + 1 < nchunks) {
          unsigned char* nxt = cbuf + ((c + 1) & 1) * C::CBUF;
          if (c + 1 < nneg) {
            const int row = lane < CH ? ext_t[(size_t)p * N + (c + 1) * CH + lane] : 0;
            gather_rows<H>(nxt, zh, row, CH, lane, HT);
          } else {
            const int row = b * S + w + 1 + (lane < K ? lane : 0);
            gather_rows<H>(nxt, zh, row, K, lane, HT);
            for (int i = lane; i < (16 - K) * (C::RS / 16); i += 32)  // rows K..15 of the positive tile: zeros
              reinterpret_cast<uint4*>(nxt + K * C::RS)[i] = make_uint4(0, 0, 0, 0);
          }
          cp_async_commit();
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        if (c == 0) {
#pragma unroll
          for (int ks = 0; ks < C::KS; ks++) {
            const int mi = lane >> 3;
            ldsm_x4(bfr[ks], s_u32(psm + ((mi >> 1) * 8 + (lane & 7)) * C::RS + (ks * 16 + (mi & 1) * 8) * 2));
          }
        }
        const int mts = (c < nneg) ? MPC : 1;
#pragma unroll
        for (int mt = 0; mt < MPC; mt++) {
          if (mt < mts) {
            const int m = (c < nneg) ? (c * MPC + mt) : NMT;
            if (m < NMT + 1) {
#pragma unroll
              for (int ks = 0; ks < C::KS; ks++) {
                uint32_t a[4];
                const int mi = lane >> 3;
                ldsm_x4(a, s_u32(cur + (mt * 16 + (mi & 1) * 8 + (lane & 7)) * C::RS + (ks * 16 + (mi >> 1) * 8) * 2));
                mma16816(acc[m][0], a, bfr[ks][0], bfr[ks][1]);
                mma16816(acc[m][1], a, bfr[ks][2], bfr[ks][3]);
              }
            }
          }
        }
        if constexpr (SPLIT == 2) {  // (MPC = 1) both warps of the pair end up with the full logits of the chunk
          const int m = (c < nneg) ? c : NMT;
          float* xo = xmine + ((c & 1) ^ xph) * 256;
#pragma unroll
          for (int n = 0; n < 2; n++)
#pragma unroll
            for (int e = 0; e < 4; e++) xo[(n * 4 + e) * 32 + lane] = acc[m][n][e];
          pair_bar(slot);
          const float* xi = xpeer + ((c & 1) ^ xph) * 256;
#pragma unroll
          for (int n = 0; n < 2; n++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[m][n][e] += xi[(n * 4 + e) * 32 + lane];
        }
        __syncwarp();
      }
    }
    // ---- softmax / CE / argmax per head column (this thread: heads 8*nt + 2*t + e) ----
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int k = nt * 8 + 2 * t + e;
        const float pos = __shfl_sync(0xffffffffu, acc[NMT][nt][2 * nt + e], (2 * t + e) * 4 + t) * invH;
        float mx = -INFINITY;
#pragma unroll
        for (int m = 0; m < NMT; m++) mx = fmaxf(mx, fmaxf(acc[m][nt][e], acc[m][nt][2 + e]));
        mx *= invH;
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        const float M = fmaxf(mx, pos);
        float sum = 0.f;
#pragma unroll
        for (int m = 0; m < NMT; m++) sum += __expf(acc[m][nt][e] * invH - M) + __expf(acc[m][nt][2 + e] * invH - M);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        sum += __expf(pos - M);
        if (g == 0 && k < K && half == 0) {
          const float lse = M + __logf(sum);
          lossbuf[(size_t)p * K + k] = lse - pos;
          corrbuf[(size_t)p * K + k] = (pos >= mx) ? 1.f : 0.f;
          lsebuf[(size_t)p * K + k] = lse;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// RED selects how the dz rows leave the SM:  0 = staged in shared memory, one TMA bulk reduction (1 KB) per row;
// 1 = red.global.add.v2.f32 straight from the accumulator fragments; 2 = red.global.add.v4.f32 after a lane-pair swap.
// Both reach the same L2 reduction ceiling (tools/redbench: 5.4 TB/s); without the 16 KB staging tile a warp needs
// 26 KB of shared memory instead of 42 KB, so 8 warps fit on an SM instead of 5.
template <int H, int RED, int SPLIT, int NNEG>
__global__ void __launch_bounds__((RED == 0 && SPLIT == 1) ? 160 : 256) score_bwd_mma_kernel(const bf16* __restrict__ pred, const bf16* __restrict__ z,
                                                            const int* __restrict__ ext_t, const float* __restrict__ lsebuf,
                                                            const float* __restrict__ dloss, bf16* __restrict__ dpred,
                                                            float* __restrict__ dz, int B, int S, int W, int K, int N,
                                                            int warps_per_cta) {
  pdl_wait();
  pdl_trigger();
  using C = Cfg<H>;
  constexpr int GRS = (CH + 8) * 2;     // bytes per row of Gs[16 heads][32 cand]
  constexpr int SRS = H * 4 + 32;                 // staging row stride: 32 B pad -> the 8-byte fragment stores of the 8 row
                                                  // groups fall into distinct banks (an unpadded 1 KB stride is an 8-way conflict)
  constexpr int STG = RED == 0 ? 16 * SRS : 0;    // staging tile [16 rows][H] fp32
  constexpr int HT = H * SPLIT;
  constexpr int PER_WARP = C::PSM + 2 * C::CBUF + 16 * GRS + STG + (SPLIT == 2 ? XCH : 0);
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int half = SPLIT == 2 ? (warp & 1) : 0, slot = warp / SPLIT, apc = warps_per_cta / SPLIT;
  unsigned char* psm = sm + (size_t)warp * PER_WARP;
  unsigned char* cbuf = psm + C::PSM;
  unsigned char* gs = cbuf + 2 * C::CBUF;
  unsigned char* stg = gs + 16 * GRS;
  float* xmine = reinterpret_cast<float*>(stg + STG);
  const float* xpeer = reinterpret_cast<const float*>(sm + (size_t)(warp ^ 1) * PER_WARP + C::PSM + 2 * C::CBUF + 16 * GRS + STG);
  for (int i = lane; i < C::PSM / 16; i += 32) reinterpret_cast<uint4*>(psm)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const int P = B * W;
  constexpr int nneg = NNEG;
  constexpr int nchunks = nneg + 1;
  const float invH = 1.f / (float)HT;
  const float gscale = 1.f / ((float)P * (float)HT);
  const bf16* zh = z + half * H;
  float* dzh = dz + half * H;

  // negative-sample rows of a position: lane l holds ext_t[p][32 j + l], j = 0..3 (N = 128); those of the NEXT position are
  // fetched while the current one is processed, so no index load sits in front of a gather
  const int pstride = gridDim.x * apc;
  int nidx[NNEG * CH / 32];
  {
    const int p0 = blockIdx.x * apc + slot;
#pragma unroll
    for (int j = 0; j < NNEG * CH / 32; j++) nidx[j] = p0 < P ? ext_t[(size_t)p0 * N + 32 * j + lane] : 0;
  }
  auto chunk_rows = [&](const int (&idx)[NNEG * CH / 32], int c) {  // lanes 0..CH-1 get the rows of negative chunk c
    const int q = (c * CH) >> 5;
    int v = idx[0];
#pragma unroll
    for (int j = 1; j < NNEG * CH / 32; j++) v = (q == j) ? idx[j] : v;
    return __shfl_sync(0xffffffffu, v, ((c * CH) & 31) + (lane & (CH - 1)));
  };

  int xph = 0;  // parity of the exchange buffer across positions (see the forward kernel)
  for (int p = blockIdx.x * apc + slot; p < P; p += pstride, xph ^= (nchunks & 1)) {
    const int b = p / W, w = p - b * W;
    stage_pred<H>(psm, pred + (size_t)p * K * HT + half * H, K, lane, HT);
    int cidx[NNEG * CH / 32];
#pragma unroll
    for (int j = 0; j < NNEG * CH / 32; j++) cidx[j] = nidx[j];
    if (p + pstride < P) {
#pragma unroll
      for (int j = 0; j < NNEG * CH / 32; j++) nidx[j] = ext_t[(size_t)(p + pstride) * N + 32 * j + lane];
    }
    int my_row = chunk_rows(cidx, 0);
    gather_rows<H>(cbuf, zh, my_row, CH, lane, HT);
    cp_async_commit();
    float lse[2][2], gk[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int k = nt * 8 + 2 * t + e;
        lse[nt][e] = k < K ? lsebuf[(size_t)p * K + k] : 0.f;
        gk[nt][e] = k < K ? dloss[k] * gscale : 0.f;
      }
    float d1[2 * C::KS][4];  // dpred accumulators: 16 heads x H
#pragma unroll
    for (int n = 0; n < 2 * C::KS; n++)
#pragma unroll
      for (int e = 0; e < 4; e++) d1[n][e] = 0.f;

    for (int c = 0; c < nchunks; c++) {
      unsigned char* cur = cbuf + (c & 1) * C::CBUF;
      const int cur_row = my_row;
      if (c + 1 < nchunks) {
        unsigned char* nxt = cbuf + ((c + 1) & 1) * C::CBUF;
        if (c + 1 < nneg) {
          my_row = chunk_rows(cidx, c + 1);
          gather_rows<H>(nxt, zh, my_row, CH, lane, HT);
        } else {
          my_row = b * S + w + 1 + (lane < K ? lane : 0);
          gather_rows<H>(nxt, zh, my_row, K, lane, HT);
          for (int i = lane; i < (16 - K) * (C::RS / 16); i += 32)
            reinterpret_cast<uint4*>(nxt + K * C::RS)[i] = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const bool is_pos = c >= nneg;
      const int mts = is_pos ? 1 : MPC;
      // ---- logits of this chunk: L[mt][nt]; even / odd k-steps accumulate separately (two independent HMMA chains) ----
      float L[MPC][2][4], L2[MPC][2][4];
#pragma unroll
      for (int mt = 0; mt < MPC; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int e = 0; e < 4; e++) L[mt][n][e] = L2[mt][n][e] = 0.f;
#pragma unroll 4
      for (int ks = 0; ks < C::KS; ks += 2) {
        uint32_t bq[4], bq2[4];
        const int mi = lane >> 3;
        ldsm_x4(bq, s_u32(psm + ((mi >> 1) * 8 + (lane & 7)) * C::RS + (ks * 16 + (mi & 1) * 8) * 2));
        ldsm_x4(bq2, s_u32(psm + ((mi >> 1) * 8 + (lane & 7)) * C::RS + ((ks + 1) * 16 + (mi & 1) * 8) * 2));
#pragma unroll
        for (int mt = 0; mt < MPC; mt++) {
          if (mt < mts) {
            uint32_t a[4], a2[4];
            ldsm_x4(a, s_u32(cur + (mt * 16 + (mi & 1) * 8 + (lane & 7)) * C::RS + (ks * 16 + (mi >> 1) * 8) * 2));
            ldsm_x4(a2, s_u32(cur + (mt * 16 + (mi & 1) * 8 + (lane & 7)) * C::RS + ((ks + 1) * 16 + (mi >> 1) * 8) * 2));
            mma16816(L[mt][0], a, bq[0], bq[1]);
            mma16816(L[mt][1], a, bq[2], bq[3]);
            mma16816(L2[mt][0], a2, bq2[0], bq2[1]);
            mma16816(L2[mt][1], a2, bq2[2], bq2[3]);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MPC; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int e = 0; e < 4; e++) L[mt][n][e] += L2[mt][n][e];
      if constexpr (SPLIT == 2) {  // (MPC = 1) the other half of the contraction comes from the partner warp
        float* xo = xmine + ((c & 1) ^ xph) * 256;
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int e = 0; e < 4; e++) xo[(n * 4 + e) * 32 + lane] = L[0][n][e];
        pair_bar(slot);
        const float* xi = xpeer + ((c & 1) ^ xph) * 256;
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int e = 0; e < 4; e++) L[0][n][e] += xi[(n * 4 + e) * 32 + lane];
      }
      // ---- G = (softmax - onehot) * dloss / (P*H); A fragments for dz and Gs[k][j] for dpred ----
      uint32_t ga[MPC][4];
#pragma unroll
      for (int mt = 0; mt < MPC; mt++) {
        float G[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int ce = e & 1;
            const int k = nt * 8 + 2 * t + ce;
            const int j = g + (e >> 1) * 8;  // row inside the m-tile
            float v = 0.f;
            if (mt < mts && k < K) {
              const float pr = __expf(L[mt][nt][e] * invH - lse[nt][ce]);
              if (!is_pos) v = pr * gk[nt][ce];
              else v = (j == k) ? (pr - 1.f) * gk[nt][ce] : 0.f;
            }
            G[nt][e] = v;
            *reinterpret_cast<bf16*>(gs + k * GRS + (mt * 16 + j) * 2) = __float2bfloat16_rn(v);
          }
        ga[mt][0] = pack_bf16(G[0][0], G[0][1]);
        ga[mt][1] = pack_bf16(G[0][2], G[0][3]);
        ga[mt][2] = pack_bf16(G[1][0], G[1][1]);
        ga[mt][3] = pack_bf16(G[1][2], G[1][3]);
      }
      __syncwarp();
      // ---- dz rows of this chunk: D2[16 j][H] = G^T . pred, added to dz (L2-resident) by reductions ----
#pragma unroll
      for (int mt = 0; mt < MPC; mt++) {
        if (mt < mts) {
          if constexpr (RED == 0) {
            if (lane < 16) bulk_wait_read0();  // the staging tile is free again
            __syncwarp();
            // the B fragments of step np + 1 are fetched before the products of step np (no LDSM latency in front of an HMMA)
            const int mi = lane >> 3;
            const uint32_t pbase = s_u32(psm + ((mi & 1) * 8 + (lane & 7)) * C::RS + ((mi >> 1) * 8) * 2);
            uint32_t bqq[2][4];
            ldsm_x4_t(bqq[0], pbase);
#pragma unroll
            for (int np = 0; np < C::KS; np++) {
              if (np + 1 < C::KS) ldsm_x4_t(bqq[(np + 1) & 1], pbase + (np + 1) * 32);
              const uint32_t (&bq)[4] = bqq[np & 1];
              float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
              mma16816(o0, ga[mt], bq[0], bq[1]);
              mma16816(o1, ga[mt], bq[2], bq[3]);
              float* r0 = reinterpret_cast<float*>(stg + (size_t)g * SRS) + np * 16 + 2 * t;
              float* r1 = reinterpret_cast<float*>(stg + (size_t)(g + 8) * SRS) + np * 16 + 2 * t;
              *reinterpret_cast<float2*>(r0) = make_float2(o0[0], o0[1]);
              *reinterpret_cast<float2*>(r1) = make_float2(o0[2], o0[3]);
              *reinterpret_cast<float2*>(r0 + 8) = make_float2(o1[0], o1[1]);
              *reinterpret_cast<float2*>(r1 + 8) = make_float2(o1[2], o1[3]);
            }
            fence_async_smem();
            __syncwarp();
            const int drow = __shfl_sync(0xffffffffu, cur_row, (mt * 16 + lane) & 31);
            if (lane < 16 && (!is_pos || lane < K)) {
              bulk_reduce_add_f32(dzh + (size_t)drow * HT, s_u32(stg + (size_t)lane * SRS), H * 4);
              bulk_commit();
            }
          } else {
            // accumulator rows g and g + 8 of the m-tile are candidates mt*16 + g and mt*16 + g + 8
            const int row_lo = __shfl_sync(0xffffffffu, cur_row, mt * 16 + g);
            const int row_hi = __shfl_sync(0xffffffffu, cur_row, mt * 16 + g + 8);
            const bool ok_lo = !is_pos || g < K, ok_hi = !is_pos || g + 8 < K;
            float* d_lo = dzh + (size_t)row_lo * HT;
            float* d_hi = dzh + (size_t)row_hi * HT;
#pragma unroll 4
            for (int np = 0; np < C::KS; np++) {
              uint32_t bq[4];
              const int mi = lane >> 3;
              ldsm_x4_t(bq, s_u32(psm + ((mi & 1) * 8 + (lane & 7)) * C::RS + (np * 16 + (mi >> 1) * 8) * 2));
              float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
              mma16816(o0, ga[mt], bq[0], bq[1]);
              mma16816(o1, ga[mt], bq[2], bq[3]);
              if constexpr (RED == 1) {
                const int cc = np * 16 + 2 * t;
                if (ok_lo) { red_add_v2(d_lo + cc, o0[0], o0[1]); red_add_v2(d_lo + cc + 8, o1[0], o1[1]); }
                if (ok_hi) { red_add_v2(d_hi + cc, o0[2], o0[3]); red_add_v2(d_hi + cc + 8, o1[2], o1[3]); }
              } else {
                // even lanes of a pair take row g, odd lanes row g + 8: four consecutive columns each
                const bool odd = t & 1;
                const float s0 = odd ? o0[0] : o0[2], s1 = odd ? o0[1] : o0[3], s2 = odd ? o1[0] : o1[2], s3 = odd ? o1[1] : o1[3];
                const float q0 = __shfl_xor_sync(0xffffffffu, s0, 1), q1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                const float q2 = __shfl_xor_sync(0xffffffffu, s2, 1), q3 = __shfl_xor_sync(0xffffffffu, s3, 1);
                const int cc = np * 16 + 2 * (t & 2);
                if (!odd) { if (ok_lo) { red_add_v4(d_lo + cc, o0[0], o0[1], q0, q1); red_add_v4(d_lo + cc + 8, o1[0], o1[1], q2, q3); } }
                else { if (ok_hi) { red_add_v4(d_hi + cc, q0, q1, o0[2], o0[3]); red_add_v4(d_hi + cc + 8, q2, q3, o1[2], o1[3]); } }
              }
            }
          }
        }
      }
      // ---- dpred += G . cand ----
#pragma unroll
      for (int ks = 0; ks < MPC; ks++) {
        if (ks < mts) {
          uint32_t a[4];
          const int mi = lane >> 3;
          ldsm_x4(a, s_u32(gs + ((mi & 1) * 8 + (lane & 7)) * GRS + (ks * 16 + (mi >> 1) * 8) * 2));
#pragma unroll
          for (int np = 0; np < C::KS; np++) {
            uint32_t bq[4];
            ldsm_x4_t(bq, s_u32(cur + (ks * 16 + (mi & 1) * 8 + (lane & 7)) * C::RS + (np * 16 + (mi >> 1) * 8) * 2));
            mma16816(d1[2 * np], a, bq[0], bq[1]);
            mma16816(d1[2 * np + 1], a, bq[2], bq[3]);
          }
        }
      }
      __syncwarp();
    }
    // ---- dpred[p][k][d] (bf16) ----
    bf16* dp = dpred + (size_t)p * K * HT + half * H;
#pragma unroll
    for (int n = 0; n < 2 * C::KS; n++) {
      const int d = n * 8 + 2 * t;
      if (g < K) *reinterpret_cast<uint32_t*>(dp + (size_t)g * HT + d) = pack_bf16(d1[n][0], d1[n][1]);
      if (g + 8 < K) *reinterpret_cast<uint32_t*>(dp + (size_t)(g + 8) * HT + d) = pack_bf16(d1[n][2], d1[n][3]);
    }
  }
  if (RED == 0 && lane < 16) bulk_wait0();
}

template <int H, int SPLIT> constexpr size_t fwd_warp_smem() { return Cfg<H>::PSM + 2 * Cfg<H>::CBUF + (SPLIT == 2 ? XCH : 0); }
template <int H, int RED, int SPLIT> constexpr size_t bwd_warp_smem() {
  return Cfg<H>::PSM + 2 * Cfg<H>::CBUF + 16 * (CH + 8) * 2 + (RED == 0 ? 16 * (H * 4 + 32) : 0) + (SPLIT == 2 ? XCH : 0);
}

template <int H, int SPLIT, int NNEG>
int launch_fwd(const bf16* pred, const bf16* z, const int* ext_t, float* lossbuf, float* corrbuf, float* lsebuf, int B, int S,
               int W, int K, int N, cudaStream_t st) {
  int wpc = (int)((216 * 1024) / fwd_warp_smem<H, SPLIT>());
  if (wpc > 8) wpc = 8;
  wpc -= wpc % SPLIT;
  const size_t smem = wpc * fwd_warp_smem<H, SPLIT>();
  auto kern = score_fwd_mma_kernel<H, SPLIT, NNEG>;
  CPC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CPC_CHECK_CUDA(launch_k(kern, dim3(148), dim3(wpc * 32), smem, st, 1, pred, z, ext_t, lossbuf, corrbuf, lsebuf, B, S, W, K, N, wpc));
  CPC_LAUNCHED_N("score_fwd_mma", st);
  return 0;
}
template <int H, int RED, int SPLIT, int NNEG>
int launch_bwd_red(const bf16* pred, const bf16* z, const int* ext_t, const float* lsebuf, const float* dloss, bf16* dpred, float* dz,
                   int B, int S, int W, int K, int N, cudaStream_t st) {
  int wpc = (int)((216 * 1024) / bwd_warp_smem<H, RED, SPLIT>());
  const int cap = (RED == 0 && SPLIT == 1) ? 5 : 8;
  if (wpc > cap) wpc = cap;
  wpc -= wpc % SPLIT;
  const size_t smem = wpc * bwd_warp_smem<H, RED, SPLIT>();
  auto kern = score_bwd_mma_kernel<H, RED, SPLIT, NNEG>;
  CPC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CPC_CHECK_CUDA(launch_k(kern, dim3(148), dim3(wpc * 32), smem, st, 1, pred, z, ext_t, lsebuf, dloss, dpred, dz, B, S, W, K, N, wpc));
  CPC_LAUNCHED_N("score_bwd_mma", st);
  return 0;
}
template <int H, int SPLIT, int NNEG>
int launch_bwd(const bf16* pred, const bf16* z, const int* ext_t, const float* lsebuf, const float* dloss, bf16* dpred, float* dz,
               int B, int S, int W, int K, int N, cudaStream_t st) {
  if constexpr (SPLIT == 1 && NNEG == 8) {  // the alternative reduction paths exist for the default shape only (experiments)
    static const int red = []() { const char* e = getenv("CPC_B200_SCORE_RED"); return e ? atoi(e) : 0; }();
    if (red == 1) return launch_bwd_red<H, 1, SPLIT, NNEG>(pred, z, ext_t, lsebuf, dloss, dpred, dz, B, S, W, K, N, st);
    if (red == 2) return launch_bwd_red<H, 2, SPLIT, NNEG>(pred, z, ext_t, lsebuf, dloss, dpred, dz, B, S, W, K, N, st);
  }
  return launch_bwd_red<H, 0, SPLIT, NNEG>(pred, z, ext_t, lsebuf, dloss, dpred, dz, B, S, W, K, N, st);
}

// experiment: H = 256 as a warp pair of 128 columns each (half the shared memory per warp -> 8 warps per SM, half the dz / dpred
// products per warp).  CPC_B200_SCORE_SPLIT=1
bool split256() {
  static const bool on = []() { const char* e = getenv("CPC_B200_SCORE_SPLIT"); return e && atoi(e) == 1; }();
  return on;
}
// (feature dim, negatives) -> instantiation
#define CPC_SCORE_DISPATCH(FN, ...)                                            \
  do {                                                                         \
    if (N == 8 * CH) {                                                         \
      if (H == 512) return FN<256, 2, 8>(__VA_ARGS__);                         \
      if (H == 256 && split256()) return FN<128, 2, 8>(__VA_ARGS__);           \
      if (H == 256) return FN<256, 1, 8>(__VA_ARGS__);                         \
      if (H == 128) return FN<128, 1, 8>(__VA_ARGS__);                         \
      return FN<64, 1, 8>(__VA_ARGS__);                                        \
    }                                                                          \
    if (H == 512) return FN<256, 2, 16>(__VA_ARGS__);                          \
    if (H == 256) return FN<256, 1, 16>(__VA_ARGS__);                          \
    if (H == 128) return FN<128, 1, 16>(__VA_ARGS__);                          \
    return FN<64, 1, 16>(__VA_ARGS__);                                         \
  } while (0)

}  // namespace

bool score_mma_supported(int H, int K, int N) {
  return (H == 64 || H == 128 || H == 256 || H == 512) && K >= 1 && K <= 16 && (N == 8 * CH || N == 16 * CH);
}

int score_transpose_ext(const int* ext, int* ext_t, int B, int N, int W, cudaStream_t st) {
  const long long n = (long long)B * N * W;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPC_CHECK_CUDA(launch_k(transpose_ext_kernel, dim3(blocks), dim3(256), 0, st, 1, ext, ext_t, B, N, W));
  CPC_LAUNCHED_N("transpose_ext", st);
  return 0;
}

int score_fwd_mma(const bf16* pred, const bf16* z, const int* ext_t, float* lossbuf, float* corrbuf, float* lsebuf, int B, int S,
                  int W, int H, int K, int N, cudaStream_t st) {
  CPC_SCORE_DISPATCH(launch_fwd, pred, z, ext_t, lossbuf, corrbuf, lsebuf, B, S, W, K, N, st);
}
int score_bwd_mma(const bf16* pred, const bf16* z, const int* ext_t, const float* lsebuf, const float* dloss, bf16* dpred,
                  float* dz, int B, int S, int W, int H, int K, int N, cudaStream_t st) {
  CPC_SCORE_DISPATCH(launch_bwd, pred, z, ext_t, lsebuf, dloss, dpred, dz, B, S, W, K, N, st);
}

}  // namespace cpcb200
