// attn_mma.cu - relative-position causal attention of one TransformerLayer (cpc/transformers.py:10-49) on mma.sync,
// forward and backward, for the bf16 path with dk = 32 (dmodel 256, 8 heads) and W <= 128 positions.
//
//   scores[i][c] = (q_i . k_c + q_i . Krelpos[:, W-1-(i-c)]) / sqrt(dk)   for c <= i      (the "skew" of transformers.py:42-47)
//   a = softmax_c(scores) ; (train mode) a~ = a * keep * 1/(1-p) ; o_i = sum_c a~[i][c] v_c
//
// One CTA per (head, window), 4 warps.  The mask is causal, so the work of a 16-row tile grows with its index: warp w owns the
// row tiles w and 7-w (rows 16w.. and 16(7-w)..) - every warp then sees the same number of keys - and, on the key side of
// backward, the key tiles w and 7-w.  Every product is an m16n8k16 bf16 MMA with fp32 accumulation:
//   QP = Q . Krelpos   -> per-warp fp32 scratch in shared memory (the skew is a row-dependent shift: it needs a round trip)
//   S  = Q . K^T (+ shifted QP), softmax in the accumulator registers, O = P . V with the accumulator-to-A-fragment reuse
// Backward recomputes S chunk by chunk (32 keys) from the row statistics, uses delta_i = dO_i . O_i (= sum_c da[i][c] a[i][c], also
// under dropout), writes dS and the dropped probabilities to shared memory for the transposed products (dK = dS^T Q,
// dV = P~^T dO) and un-skews dS in place over the QP scratch for dQ += dS~ . Krelpos^T and dKrelpos = Q^T dS~.
// The fp32 path and other head widths keep the CUDA-core kernels of thead.cu.
#include "common.cuh"

namespace cpcb200 {

namespace {

constexpr int WP = 128;   // padded positions
constexpr int DKC = 32;   // head width
constexpr int RS = 40;    // bf16 elements per shared-memory row of a [pos][dk] tile (80 B: ldmatrix rows hit distinct 16-B slots)
constexpr int QS = 132;   // fp32 elements per row of the per-warp QP / dS~ scratch
constexpr int SS = 136;   // bf16 elements per row of the dS / P~ tiles (272 B)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
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

// A fragments (16 rows x 16 k) of a row-major [row][k] bf16 tile with row stride `rs` elements
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* tile, int rs, int row0, int k0, int lane) {
  ldsm_x4(a, s_u32(tile + (row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * rs + k0 + (lane >> 4) * 8));
}
// A fragments of the TRANSPOSE of a row-major [k][row] tile: A[m][k] = tile[k0 + k][m0 + m]
__device__ __forceinline__ void load_a_t(uint32_t (&a)[4], const bf16* tile, int rs, int m0, int k0, int lane) {
  ldsm_x4_t(a, s_u32(tile + (k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * rs + m0 + ((lane >> 3) & 1) * 8));
}
// B fragments of one n-tile (8 columns n) for BOTH k-steps of a 32-deep product, from a [n][k] tile (B "col" layout):
// b[0], b[1] -> k 0..15 ; b[2], b[3] -> k 16..31
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], const bf16* tile, int n0, int lane) {
  ldsm_x4(b, s_u32(tile + (n0 + (lane & 7)) * RS + (lane >> 3) * 8));
}
// B fragments of TWO n-tiles (n0, n0+8) for one k-step (16 rows k0..k0+15) from a [k][n] tile: b[0], b[1] -> n0 ; b[2], b[3] -> n0+8
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], const bf16* tile, int rs, int k0, int n0, int lane) {
  ldsm_x4_t(b, s_u32(tile + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * rs + n0 + (lane >> 4) * 8));
}

// Q, K, V (and dO) rows of (window b, head h) -> [pos][dk] tiles; rows >= W are zero
__device__ __forceinline__ void fill_tiles(const bf16* __restrict__ qkv, const bf16* __restrict__ datt, bf16* Qs, bf16* Ks, bf16* Vs,
                                           bf16* Gs, int b, int h, int W, int D, int tid) {
  for (int idx = tid; idx < WP * 4; idx += 128) {
    const int r = idx >> 2, sg = idx & 3;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, g = q;
    if (r < W) {
      const bf16* src = qkv + ((size_t)b * W + r) * 3 * D + h * DKC + sg * 8;
      q = *reinterpret_cast<const uint4*>(src);
      k = *reinterpret_cast<const uint4*>(src + D);
      v = *reinterpret_cast<const uint4*>(src + 2 * D);
      if (Gs != nullptr) g = *reinterpret_cast<const uint4*>(datt + ((size_t)b * W + r) * D + h * DKC + sg * 8);
    }
    *reinterpret_cast<uint4*>(Qs + r * RS + sg * 8) = q;
    *reinterpret_cast<uint4*>(Ks + r * RS + sg * 8) = k;
    *reinterpret_cast<uint4*>(Vs + r * RS + sg * 8) = v;
    if (Gs != nullptr) *reinterpret_cast<uint4*>(Gs + r * RS + sg * 8) = g;
  }
}
// Rs[m][d] = Krelpos[d][m] (bf16), rows m >= W zero
__device__ __forceinline__ void fill_relpos(const float* __restrict__ krel, bf16* Rs, int W, int tid) {
  for (int idx = tid; idx < DKC * WP; idx += 128) {
    const int d = idx / WP, m = idx - d * WP;
    Rs[m * RS + d] = __float2bfloat16_rn(m < W ? krel[(size_t)d * W + m] : 0.f);
  }
}

// QP[il][m] = q_{r0+il} . Krelpos[:, m] for the 32 rows of a warp -> fp32 scratch, ZERO where the entry belongs to no key
// (m < W-1-i, or a padding row): the backward pass overwrites the live cells with dS and then reads the whole row as dS~
__device__ __forceinline__ void qp_to_scratch(const uint32_t (&aq)[2][2][4], const bf16* Rs, float* qp, const int (&row0)[2], int W,
                                              int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 16; nt++) {
    uint32_t bb[4];
    load_b_nk(bb, Rs, 8 * nt, lane);
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(acc, aq[mt][0], bb[0], bb[1]);
      mma16816(acc, aq[mt][1], bb[2], bb[3]);
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        const int il = 16 * mt + g + 8 * hf, i = row0[mt] + g + 8 * hf, m = 8 * nt + 2 * t;
        float2 v = make_float2(acc[2 * hf], acc[2 * hf + 1]);
        if (i >= W || m < W - 1 - i || m >= W) v.x = 0.f;
        if (i >= W || m + 1 < W - 1 - i || m + 1 >= W) v.y = 0.f;
        *reinterpret_cast<float2*>(qp + il * QS + m) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(const bf16* __restrict__ qkv, const float* __restrict__ krel,
                                                            bf16* __restrict__ att, int W, int D,
                                                            const unsigned char* __restrict__ keep, float dscale, int bph) {
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + WP * RS;
  bf16* Vs = Ks + WP * RS;
  bf16* Rs = Vs + WP * RS;
  float* QP = reinterpret_cast<float*>(Rs + WP * RS);  // [4 warps][32][QS]
  const int h = blockIdx.x, b = blockIdx.y, nh = gridDim.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  fill_tiles(qkv, nullptr, Qs, Ks, Vs, nullptr, b, h, W, D, tid);
  if (bph > 0) krel += (size_t)(b / bph) * DKC * W;  // blockIdx.y = (prediction head, window): stacked Krelpos
  fill_relpos(krel, Rs, W, tid);
  __syncthreads();
  const int row0[2] = {16 * warp, 16 * (7 - warp)};          // first rows of this warp's two m-tiles
  const int ntm[2] = {2 * (warp + 1), 2 * (8 - warp)};       // 8-key tiles a row tile can see: keys c <= i < row0 + 16
  uint32_t aq[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int ks = 0; ks < 2; ks++) load_a(aq[mt][ks], Qs, RS, row0[mt], 16 * ks, lane);
  float* qp = QP + warp * 32 * QS;
  qp_to_scratch(aq, Rs, qp, row0, W, lane);
  __syncwarp();
  float S[2][16][4];
#pragma unroll
  for (int nt = 0; nt < 16; nt++) {
    if (nt < ntm[1]) {
      uint32_t bb[4];
      load_b_nk(bb, Ks, 8 * nt, lane);
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        if (nt < ntm[mt]) {
          S[mt][nt][0] = S[mt][nt][1] = S[mt][nt][2] = S[mt][nt][3] = 0.f;
          mma16816(S[mt][nt], aq[mt][0], bb[0], bb[1]);
          mma16816(S[mt][nt], aq[mt][1], bb[2], bb[3]);
        }
      }
    }
  }
  const float scale = rsqrtf((float)DKC);
  float mx[2][2], inv[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const int il = 16 * mt + g + 8 * hf, i = row0[mt] + g + 8 * hf;
      float m_ = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 16; nt++) {
        if (nt < ntm[mt]) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int c = 8 * nt + 2 * t + q;
            float s = -INFINITY;
            if (i < W && c <= i) s = (S[mt][nt][2 * hf + q] + qp[il * QS + W - 1 - i + c]) * scale;
            S[mt][nt][2 * hf + q] = s;
            m_ = fmaxf(m_, s);
          }
        }
      }
      m_ = fmaxf(m_, __shfl_xor_sync(0xffffffffu, m_, 1));
      m_ = fmaxf(m_, __shfl_xor_sync(0xffffffffu, m_, 2));
      if (m_ == -INFINITY) m_ = 0.f;
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 16; nt++) {
        if (nt < ntm[mt]) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const float e = __expf(S[mt][nt][2 * hf + q] - m_);
            S[mt][nt][2 * hf + q] = e;
            sum += e;
          }
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      mx[mt][hf] = m_;
      inv[mt][hf] = sum > 0.f ? 1.f / sum : 0.f;
    }
  // probabilities (dropped in train mode)
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const int i = row0[mt] + g + 8 * hf;
      const unsigned char* krow = (keep != nullptr && i < W) ? keep + (((size_t)b * nh + h) * W + i) * W : nullptr;
#pragma unroll
      for (int nt = 0; nt < 16; nt++) {
        if (nt < ntm[mt]) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int c = 8 * nt + 2 * t + q;
            float p = S[mt][nt][2 * hf + q] * inv[mt][hf];
            if (krow != nullptr && c <= i) p = krow[c] ? p * dscale : 0.f;
            S[mt][nt][2 * hf + q] = p;
          }
        }
      }
    }
  (void)mx;
  // O = P . V
  float O[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nd = 0; nd < 4; nd++) O[mt][nd][0] = O[mt][nd][1] = O[mt][nd][2] = O[mt][nd][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 8; kk++) {
    if (2 * kk < ntm[1]) {
      uint32_t bv[2][4];
      load_b_kn(bv[0], Vs, RS, 16 * kk, 0, lane);
      load_b_kn(bv[1], Vs, RS, 16 * kk, 16, lane);
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        if (2 * kk >= ntm[mt]) continue;
        uint32_t a[4];
        a[0] = pack_bf16(S[mt][2 * kk][0], S[mt][2 * kk][1]);
        a[1] = pack_bf16(S[mt][2 * kk][2], S[mt][2 * kk][3]);
        a[2] = pack_bf16(S[mt][2 * kk + 1][0], S[mt][2 * kk + 1][1]);
        a[3] = pack_bf16(S[mt][2 * kk + 1][2], S[mt][2 * kk + 1][3]);
        mma16816(O[mt][0], a, bv[0][0], bv[0][1]);
        mma16816(O[mt][1], a, bv[0][2], bv[0][3]);
        mma16816(O[mt][2], a, bv[1][0], bv[1][1]);
        mma16816(O[mt][3], a, bv[1][2], bv[1][3]);
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const int i = row0[mt] + g + 8 * hf;
      if (i < W) {
        bf16* orow = att + ((size_t)b * W + i) * D + h * DKC;
#pragma unroll
        for (int nd = 0; nd < 4; nd++)
          *reinterpret_cast<uint32_t*>(orow + 8 * nd + 2 * t) = pack_bf16(O[mt][nd][2 * hf], O[mt][nd][2 * hf + 1]);
      }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ datt,
                                                            const bf16* __restrict__ att, const float* __restrict__ krel,
                                                            bf16* __restrict__ dqkv, float* __restrict__ dkrel, int W, int D,
                                                            const unsigned char* __restrict__ keep, float dscale, int bph) {
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + WP * RS;
  bf16* Vs = Ks + WP * RS;
  bf16* Gs = Vs + WP * RS;   // dO
  bf16* Rs = Gs + WP * RS;
  float* QP = reinterpret_cast<float*>(Rs + WP * RS);               // [4][32][QS]: QP, then dS~ (un-skewed dS), fp32
  bf16* dSs = reinterpret_cast<bf16*>(QP + 4 * 32 * QS);            // [WP][SS]  dS[i][c]
  bf16* Pds = dSs + WP * SS;                                        // [WP][SS]  dropped probabilities
  const int h = blockIdx.x, b = blockIdx.y, nh = gridDim.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  fill_tiles(qkv, datt, Qs, Ks, Vs, Gs, b, h, W, D, tid);
  if (bph > 0) { krel += (size_t)(b / bph) * DKC * W; dkrel += (size_t)(b / bph) * DKC * W; }
  fill_relpos(krel, Rs, W, tid);
  for (int idx = tid; idx < WP * SS / 8; idx += 128) {
    reinterpret_cast<uint4*>(dSs)[idx] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(Pds)[idx] = make_uint4(0, 0, 0, 0);
  }
  for (int idx = tid; idx < 4 * 32 * QS / 4; idx += 128) reinterpret_cast<float4*>(QP)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int row0[2] = {16 * warp, 16 * (7 - warp)};          // this warp's two row tiles (and, below, key tiles)
  const int ntm[2] = {2 * (warp + 1), 2 * (8 - warp)};       // 8-key tiles a row tile can see
  const int kcm[2] = {warp >> 1, (7 - warp) >> 1};           // last 32-key chunk a row tile can see
  const float scale = rsqrtf((float)DKC);
  float* qp = QP + warp * 32 * QS;
  bf16* dbase = dqkv + (size_t)b * W * 3 * D + h * DKC;
  {
    uint32_t aq[2][2][4], ag[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        load_a(aq[mt][ks], Qs, RS, row0[mt], 16 * ks, lane);
        load_a(ag[mt][ks], Gs, RS, row0[mt], 16 * ks, lane);
      }
    qp_to_scratch(aq, Rs, qp, row0, W, lane);
    __syncwarp();
    // ---- row statistics (max, 1 / sum) from a full pass over the keys; delta_i = dO_i . O_i ----
    float mx[2][2], inv[2][2], delta[2][2];
    {
      float S[2][16][4];
#pragma unroll
      for (int nt = 0; nt < 16; nt++) {
        if (nt < ntm[1]) {
          uint32_t bb[4];
          load_b_nk(bb, Ks, 8 * nt, lane);
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            if (nt < ntm[mt]) {
              S[mt][nt][0] = S[mt][nt][1] = S[mt][nt][2] = S[mt][nt][3] = 0.f;
              mma16816(S[mt][nt], aq[mt][0], bb[0], bb[1]);
              mma16816(S[mt][nt], aq[mt][1], bb[2], bb[3]);
            }
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const int il = 16 * mt + g + 8 * hf, i = row0[mt] + g + 8 * hf;
          float m_ = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 16; nt++) {
            if (nt < ntm[mt]) {
#pragma unroll
              for (int q = 0; q < 2; q++) {
                const int c = 8 * nt + 2 * t + q;
                float s = -INFINITY;
                if (i < W && c <= i) s = (S[mt][nt][2 * hf + q] + qp[il * QS + W - 1 - i + c]) * scale;
                S[mt][nt][2 * hf + q] = s;
                m_ = fmaxf(m_, s);
              }
            }
          }
          m_ = fmaxf(m_, __shfl_xor_sync(0xffffffffu, m_, 1));
          m_ = fmaxf(m_, __shfl_xor_sync(0xffffffffu, m_, 2));
          if (m_ == -INFINITY) m_ = 0.f;
          float sum = 0.f;
#pragma unroll
          for (int nt = 0; nt < 16; nt++) {
            if (nt < ntm[mt]) {
#pragma unroll
              for (int q = 0; q < 2; q++) sum += __expf(S[mt][nt][2 * hf + q] - m_);
            }
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          mx[mt][hf] = m_;
          inv[mt][hf] = sum > 0.f ? 1.f / sum : 0.f;
          // delta: lane t covers d = 8t .. 8t+7 of the row
          float dl = 0.f;
          if (i < W) {
            const uint4 ov = *reinterpret_cast<const uint4*>(att + ((size_t)b * W + i) * D + h * DKC + 8 * t);
            const uint4 gv = *reinterpret_cast<const uint4*>(Gs + i * RS + 8 * t);
            const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
            const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 a = __bfloat1622float2(o2[j]), c2 = __bfloat1622float2(g2[j]);
              dl = fmaf(a.x, c2.x, dl); dl = fmaf(a.y, c2.y, dl);
            }
          }
          dl += __shfl_xor_sync(0xffffffffu, dl, 1);
          dl += __shfl_xor_sync(0xffffffffu, dl, 2);
          delta[mt][hf] = dl;
        }
    }
    // ---- chunks of 32 keys: recompute P, dP = dO . V^T, dS; dQ += dS . K ----
    float dQ[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nd = 0; nd < 4; nd++) dQ[mt][nd][0] = dQ[mt][nd][1] = dQ[mt][nd][2] = dQ[mt][nd][3] = 0.f;
#pragma unroll 1
    for (int kc = 0; kc <= kcm[1]; kc++) {
      float S[2][4][4], dP[2][4][4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int nt = 4 * kc + j;
        uint32_t bk[4], bv[4];
        load_b_nk(bk, Ks, 8 * nt, lane);
        load_b_nk(bv, Vs, 8 * nt, lane);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          if (kc > kcm[mt]) continue;  // (warp-uniform: the low row tile sees fewer chunks)
          S[mt][j][0] = S[mt][j][1] = S[mt][j][2] = S[mt][j][3] = 0.f;
          dP[mt][j][0] = dP[mt][j][1] = dP[mt][j][2] = dP[mt][j][3] = 0.f;
          mma16816(S[mt][j], aq[mt][0], bk[0], bk[1]);
          mma16816(S[mt][j], aq[mt][1], bk[2], bk[3]);
          mma16816(dP[mt][j], ag[mt][0], bv[0], bv[1]);
          mma16816(dP[mt][j], ag[mt][1], bv[2], bv[3]);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          if (kc > kcm[mt]) continue;  // (its dS / P~ cells keep the zeros of the initial fill)
          const int il = 16 * mt + g + 8 * hf, i = row0[mt] + g + 8 * hf;
          const unsigned char* krow = (keep != nullptr && i < W) ? keep + (((size_t)b * nh + h) * W + i) * W : nullptr;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float ds2[2], pd2[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
              const int c = 32 * kc + 8 * j + 2 * t + q;
              float ds = 0.f, pd = 0.f;
              if (i < W && c <= i) {
                float* cell = qp + il * QS + W - 1 - i + c;
                const float s = (S[mt][j][2 * hf + q] + *cell) * scale;
                const float p = __expf(s - mx[mt][hf]) * inv[mt][hf];
                float dp = dP[mt][j][2 * hf + q];
                pd = p;
                if (krow != nullptr) {
                  const bool kp = krow[c] != 0;
                  dp = kp ? dp * dscale : 0.f;
                  pd = kp ? p * dscale : 0.f;
                }
                ds = p * (dp - delta[mt][hf]) * scale;
                *cell = ds;  // un-skewed dS~[i][m = W-1-i+c] over the QP entry that was just consumed
              }
              ds2[q] = ds; pd2[q] = pd;
            }
            S[mt][j][2 * hf] = ds2[0]; S[mt][j][2 * hf + 1] = ds2[1];
            *reinterpret_cast<uint32_t*>(dSs + i * SS + 32 * kc + 8 * j + 2 * t) = pack_bf16(ds2[0], ds2[1]);
            *reinterpret_cast<uint32_t*>(Pds + i * SS + 32 * kc + 8 * j + 2 * t) = pack_bf16(pd2[0], pd2[1]);
          }
        }
      // dQ += dS_chunk . K_chunk   (A from the accumulators, B = K as [k = key][n = d])
#pragma unroll
      for (int kk = 0; kk < 2; kk++) {
        uint32_t bk2[2][4];
        load_b_kn(bk2[0], Ks, RS, 32 * kc + 16 * kk, 0, lane);
        load_b_kn(bk2[1], Ks, RS, 32 * kc + 16 * kk, 16, lane);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          if (kc > kcm[mt]) continue;
          uint32_t a[4];
          a[0] = pack_bf16(S[mt][2 * kk][0], S[mt][2 * kk][1]);
          a[1] = pack_bf16(S[mt][2 * kk][2], S[mt][2 * kk][3]);
          a[2] = pack_bf16(S[mt][2 * kk + 1][0], S[mt][2 * kk + 1][1]);
          a[3] = pack_bf16(S[mt][2 * kk + 1][2], S[mt][2 * kk + 1][3]);
          mma16816(dQ[mt][0], a, bk2[0][0], bk2[0][1]);
          mma16816(dQ[mt][1], a, bk2[0][2], bk2[0][3]);
          mma16816(dQ[mt][2], a, bk2[1][0], bk2[1][1]);
          mma16816(dQ[mt][3], a, bk2[1][2], bk2[1][3]);
        }
      }
    }
    __syncwarp();
    // dQ += dS~ . Krelpos^T : A from the fp32 scratch (rows il, k = m), B = Rs as [k = m][n = d]
#pragma unroll 1
    for (int km = 0; km < 8; km++) {
      uint32_t br[2][4];
      load_b_kn(br[0], Rs, RS, 16 * km, 0, lane);
      load_b_kn(br[1], Rs, RS, 16 * km, 16, lane);
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        const float* s0 = qp + (16 * mt + g) * QS + 16 * km + 2 * t;
        const float* s1 = s0 + 8 * QS;
        uint32_t a[4];
        a[0] = pack_bf16(s0[0], s0[1]);
        a[1] = pack_bf16(s1[0], s1[1]);
        a[2] = pack_bf16(s0[8], s0[9]);
        a[3] = pack_bf16(s1[8], s1[9]);
        mma16816(dQ[mt][0], a, br[0][0], br[0][1]);
        mma16816(dQ[mt][1], a, br[0][2], br[0][3]);
        mma16816(dQ[mt][2], a, br[1][0], br[1][1]);
        mma16816(dQ[mt][3], a, br[1][2], br[1][3]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int hf = 0; hf < 2; hf++) {
        const int i = row0[mt] + g + 8 * hf;
        if (i < W) {
#pragma unroll
          for (int nd = 0; nd < 4; nd++)
            *reinterpret_cast<uint32_t*>(dbase + (size_t)i * 3 * D + 8 * nd + 2 * t) = pack_bf16(dQ[mt][nd][2 * hf], dQ[mt][nd][2 * hf + 1]);
        }
      }
  }
  __syncthreads();
  // ---- key side: warp w owns the key / column tiles w and 7-w (16 each): key tile j meets the row tiles ki >= j ----
  float dK[2][4][4], dV[2][4][4], dR[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nd = 0; nd < 4; nd++)
#pragma unroll
      for (int e = 0; e < 4; e++) dK[mt][nd][e] = dV[mt][nd][e] = dR[mt][nd][e] = 0.f;
  const int n_it = (W + 15) / 16;
#pragma unroll 1
  for (int ki = 0; ki < n_it; ki++) {
    uint32_t bq[2][4], bg[2][4];
    load_b_kn(bq[0], Qs, RS, 16 * ki, 0, lane);
    load_b_kn(bq[1], Qs, RS, 16 * ki, 16, lane);
    if (16 * ki >= row0[0]) {  // rows i >= c only: tiles entirely above the diagonal hold zeros
      load_b_kn(bg[0], Gs, RS, 16 * ki, 0, lane);
      load_b_kn(bg[1], Gs, RS, 16 * ki, 16, lane);
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        if (16 * ki < row0[mt]) continue;
        uint32_t a[4], ap[4];
        load_a_t(a, dSs, SS, row0[mt], 16 * ki, lane);
        load_a_t(ap, Pds, SS, row0[mt], 16 * ki, lane);
        mma16816(dK[mt][0], a, bq[0][0], bq[0][1]);
        mma16816(dK[mt][1], a, bq[0][2], bq[0][3]);
        mma16816(dK[mt][2], a, bq[1][0], bq[1][1]);
        mma16816(dK[mt][3], a, bq[1][2], bq[1][3]);
        mma16816(dV[mt][0], ap, bg[0][0], bg[0][1]);
        mma16816(dV[mt][1], ap, bg[0][2], bg[0][3]);
        mma16816(dV[mt][2], ap, bg[1][0], bg[1][1]);
        mma16816(dV[mt][3], ap, bg[1][2], bg[1][3]);
      }
    }
    // dKrelpos^T[m][d] += sum_i dS~[i][m] q_i[d] : A[m][k = i] from the fp32 scratch of the warp that owns row tile ki
    // (warp ki holds it as its first tile when ki < 4, warp 7-ki as its second otherwise)
    const float* sc = QP + (ki < 4 ? ki * 32 : (7 - ki) * 32 + 16) * QS;
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      const int m = row0[mt] + g;
      uint32_t a[4];
      a[0] = pack_bf16(sc[(2 * t) * QS + m], sc[(2 * t + 1) * QS + m]);
      a[1] = pack_bf16(sc[(2 * t) * QS + m + 8], sc[(2 * t + 1) * QS + m + 8]);
      a[2] = pack_bf16(sc[(2 * t + 8) * QS + m], sc[(2 * t + 9) * QS + m]);
      a[3] = pack_bf16(sc[(2 * t + 8) * QS + m + 8], sc[(2 * t + 9) * QS + m + 8]);
      mma16816(dR[mt][0], a, bq[0][0], bq[0][1]);
      mma16816(dR[mt][1], a, bq[0][2], bq[0][3]);
      mma16816(dR[mt][2], a, bq[1][0], bq[1][1]);
      mma16816(dR[mt][3], a, bq[1][2], bq[1][3]);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const int c = row0[mt] + g + 8 * hf;
      if (c < W) {
#pragma unroll
        for (int nd = 0; nd < 4; nd++) {
          *reinterpret_cast<uint32_t*>(dbase + (size_t)c * 3 * D + D + 8 * nd + 2 * t) = pack_bf16(dK[mt][nd][2 * hf], dK[mt][nd][2 * hf + 1]);
          *reinterpret_cast<uint32_t*>(dbase + (size_t)c * 3 * D + 2 * D + 8 * nd + 2 * t) = pack_bf16(dV[mt][nd][2 * hf], dV[mt][nd][2 * hf + 1]);
#pragma unroll
          for (int q = 0; q < 2; q++) atomicAdd(dkrel + (size_t)(8 * nd + 2 * t + q) * W + c, dR[mt][nd][2 * hf + q]);
        }
      }
    }
}

constexpr size_t kFwdSmem = (size_t)4 * WP * RS * 2 + (size_t)4 * 32 * QS * 4;
constexpr size_t kBwdSmem = (size_t)5 * WP * RS * 2 + (size_t)4 * 32 * QS * 4 + (size_t)2 * WP * SS * 2;

}  // namespace

bool attn_mma_supported(int W, int D, int nh) {
  static const bool off = []() { const char* e = getenv("CPC_B200_ATTN_MMA"); return e && atoi(e) == 0; }();
  return !off && W <= WP && D == nh * DKC && D % 8 == 0;
}

// bph > 0: B counts (prediction head, window) pairs, bph windows per head; krel / dkrel are the heads' stacked (dk, W) matrices
int attn_fwd_mma(const bf16* qkv, const float* krel, bf16* att, int B, int W, int D, int nh, const unsigned char* keep, float dscale,
                 cudaStream_t st, int bph) {
  CPC_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
  attn_fwd_mma_kernel<<<dim3(nh, B), 128, kFwdSmem, st>>>(qkv, krel, att, W, D, keep, dscale, bph);
  CPC_LAUNCHED_N("attn_fwd_mma", st);
  return 0;
}
int attn_bwd_mma(const bf16* qkv, const bf16* datt, const bf16* att, const float* krel, bf16* dqkv, float* dkrel, int B, int W, int D,
                 int nh, const unsigned char* keep, float dscale, cudaStream_t st, int bph) {
  CPC_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
  attn_bwd_mma_kernel<<<dim3(nh, B), 128, kBwdSmem, st>>>(qkv, datt, att, krel, dqkv, dkrel, W, D, keep, dscale, bph);
  CPC_LAUNCHED_N("attn_bwd_mma", st);
  return 0;
}

}  // namespace cpcb200
