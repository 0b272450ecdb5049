// lstm_mma.cu - tensor-core LSTM recurrence for the bf16 path (reference: cpc/model.py:171-173, 193 -> torch.nn.LSTM, the
// reference's DEFAULT context network, cpc_default_config.py:74; gate order i, f, g, o).
//
// The design of gru_mma.cu with four gate rows per hidden unit: cluster of Har/64 CTAs, CTA r owns hidden units
// [64r, 64r+64), its slice of W_hh lives in REGISTERS as mma.sync.m16n8k16 A fragments for all S steps (4 gates x Har/32
// k-steps x 4 registers = 128 registers per thread at Har = 256), h_{t-1} / d(gates)_t is the bf16 B operand in shared memory
// read with ldmatrix.trans, the two warps of a pair split the k range and swap half of their partial sums, the CTA's 64 new
// units leave as ONE bulk shared->shared copy per peer CTA that completes a byte count on the peer's mbarrier (no cluster
// barrier on the per-step path).  The cell state of a (unit, sequence) pair stays in a register of the thread that owns it.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cpcb200 {

namespace {

constexpr int HC = 64;   // hidden units per CTA
constexpr int BT = 8;    // sequences per cluster (= the n of m16n8k16)

#include "rec_mma.cuh"

// ---------------------------------------------------------------------------------------------------------
// forward.  block = 256 threads = 8 warps: warp w owns units 16*(w&3).. of the CTA's 64-unit slice and the k range
// [ (w>>2)*HAR/2, +HAR/2 ) of the product; after the pair exchange every thread finishes ONE unit for two sequences.
// gi (B, S, 4*HAR) bf16 = W_ih x + b_ih (hoisted GEMM).  gates4 == NULL: inference, nothing is saved.
// ---------------------------------------------------------------------------------------------------------
template <int HAR>
__global__ void __launch_bounds__(256, 1)
lstm_rec_fwd_mma_kernel(const bf16* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                        const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ out, bf16* __restrict__ outT,
                        uint2* __restrict__ gates4, float* __restrict__ cell, float* __restrict__ hT, float* __restrict__ cT,
                        int B, int S) {
  constexpr int KS = HAR / 16, KSH = KS / 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 3, kh = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  __shared__ __align__(128) bf16 hs[2][HAR][BT];       // h_{t-1} as the B operand: [k][sequence]
  __shared__ float part[4][2][8][32];                    // partial sums handed to the partner warp, per unit block
  __shared__ __align__(128) bf16 hstage[2][HC][BT];     // the CTA's 64 new units, staged for ONE bulk copy per destination CTA
  __shared__ __align__(8) uint64_t hbar[2];
  constexpr uint32_t kStepBytes = HAR * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_mbar_init_cluster();
  }
  // resident A fragments: gate gt, rows 16*ub + {g, g+8} of this CTA's slice, k-steps of this warp's half
  uint32_t wf[4][KSH][4];
#pragma unroll
  for (int gt = 0; gt < 4; gt++) {
    const float* r0 = w_hh + (size_t)(gt * HAR + HC * rank + 16 * ub + g) * HAR;
    const float* r1 = r0 + 8 * HAR;
#pragma unroll
    for (int ks = 0; ks < KSH; ks++) {
      const int k = (kh * KSH + ks) * 16 + 2 * t4;
      wf[gt][ks][0] = pack_bf16(__ldg(r0 + k), __ldg(r0 + k + 1));
      wf[gt][ks][1] = pack_bf16(__ldg(r1 + k), __ldg(r1 + k + 1));
      wf[gt][ks][2] = pack_bf16(__ldg(r0 + k + 8), __ldg(r0 + k + 9));
      wf[gt][ks][3] = pack_bf16(__ldg(r1 + k + 8), __ldg(r1 + k + 9));
    }
  }
  pdl_wait();
  pdl_trigger();
  const int col = HC * rank + 16 * ub + 8 * kh + g;  // the unit this thread finishes, sequences b0 + 2*t4 + {0, 1}
  float bh[4];
#pragma unroll
  for (int gt = 0; gt < 4; gt++) bh[gt] = __ldg(b_hh + gt * HAR + col);
  float hprev[2], cprev[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int bq = b0 + 2 * t4 + q;
    hprev[q] = (h0 != nullptr && bq < B) ? h0[(size_t)bq * HAR + col] : 0.f;
    cprev[q] = (c0 != nullptr && bq < B) ? c0[(size_t)bq * HAR + col] : 0.f;
  }
  for (int i = threadIdx.x; i < HAR * BT; i += blockDim.x) {
    const int k = i / BT, bq = b0 + (i - k * BT);
    hs[0][k][i - k * BT] = __float2bfloat16_rn((h0 != nullptr && bq < B) ? h0[(size_t)bq * HAR + k] : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // arm both buffers: hs[1] is produced by step 0, hs[0] by step 1
    mbar_expect_tx(&hbar[1], kStepBytes);
    mbar_expect_tx(&hbar[0], kStepBytes);
  }
  cluster.sync();
  const uint32_t hs_local = s_u32(&hs[0][0][0]), bar_local = s_u32(&hbar[0]);
  const uint32_t pub_dst = mapa_u32(hs_local + (uint32_t)(HC * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

  bool okq[2];
  const bf16* gip[2];
  size_t orow[2];  // element offset of (sequence q, t = 0, unit col) in the (B, S, HAR) arrays
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int bq = b0 + 2 * t4 + q;
    okq[q] = bq < B;
    const size_t row0 = (size_t)(okq[q] ? bq : b0) * S;
    gip[q] = gi + row0 * 4 * HAR + col;
    orow[q] = row0 * HAR + col;
  }
  bf16 gq_raw[4][2];
  auto load_gi = [&]() {
#pragma unroll
    for (int q = 0; q < 2; q++) {
#pragma unroll
      for (int gt = 0; gt < 4; gt++) gq_raw[gt][q] = gip[q][gt * HAR];
      gip[q] += 4 * HAR;
    }
  };
  load_gi();

  for (int t = 0; t < S; t++) {
    const int cur = t & 1, nxt = cur ^ 1;
    float gq[4][2];
#pragma unroll
    for (int gt = 0; gt < 4; gt++)
#pragma unroll
      for (int q = 0; q < 2; q++) gq[gt][q] = __bfloat162float(gq_raw[gt][q]);
    if (t + 1 < S) load_gi();  // in flight during this step's product and exchange
    if (t > 0) {
      mbar_wait(&hbar[cur], ((t - 1 - (cur ^ 1)) >> 1) & 1);
      if (threadIdx.x == 0 && t + 1 < S) mbar_expect_tx(&hbar[cur], kStepBytes);  // re-arm for step t+2
    }
    float acc[4][2][4];
#pragma unroll
    for (int gt = 0; gt < 4; gt++)
#pragma unroll
      for (int ch = 0; ch < 2; ch++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[gt][ch][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSH / 2; q++) {
      uint32_t bq4[4];
      ldsm_x4_t(bq4, s_u32(&hs[cur][kh * (HAR / 2) + 32 * q + lane][0]));
#pragma unroll
      for (int gt = 0; gt < 4; gt++) {
        mma16816(acc[gt][0], wf[gt][2 * q], bq4[0], bq4[1]);
        mma16816(acc[gt][1], wf[gt][2 * q + 1], bq4[2], bq4[3]);
      }
    }
    // accumulator e: row g (e = 0,1) / g+8 (e = 2,3), sequence 2*t4 + (e&1).  Keep rows g + 8*kh, hand the others over.
    float mine[4][2];
#pragma unroll
    for (int gt = 0; gt < 4; gt++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float lo = acc[gt][0][q] + acc[gt][1][q], hi = acc[gt][0][2 + q] + acc[gt][1][2 + q];
        mine[gt][q] = kh ? hi : lo;
        part[ub][kh][gt * 2 + q][lane] = kh ? lo : hi;
      }
    pair_sync(ub);
    float hn[2], cn[2];
    uint2 sv4[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const float ai = mine[0][q] + part[ub][1 - kh][q][lane];
      const float af = mine[1][q] + part[ub][1 - kh][2 + q][lane];
      const float ag = mine[2][q] + part[ub][1 - kh][4 + q][lane];
      const float ao = mine[3][q] + part[ub][1 - kh][6 + q][lane];
      const float ig = sigmoid_fast(gq[0][q] + ai + bh[0]);
      const float fg = sigmoid_fast(gq[1][q] + af + bh[1]);
      const float gg = tanh_fast(gq[2][q] + ag + bh[2]);
      const float og = sigmoid_fast(gq[3][q] + ao + bh[3]);
      cn[q] = fmaf(fg, cprev[q], ig * gg);
      hn[q] = og * tanh_fast(cn[q]);
      cprev[q] = cn[q];
      hprev[q] = hn[q];
      sv4[q] = make_uint2(pack_bf16(ig, fg), pack_bf16(gg, og));
    }
    if (t + 1 < S) {
      *reinterpret_cast<uint32_t*>(&hstage[cur][16 * ub + 8 * kh + g][2 * t4]) = pack_bf16(hn[0], hn[1]);
      fence_async_smem();
      if (warp == 0) {
        publish_sync();
        if (lane < CS) bulk_s2s(pub_dst + nxt * (HAR * BT * 2), s_u32(&hstage[cur][0][0]), HC * BT * 2, pub_bar + nxt * 8);
      } else {
        publish_arrive();
      }
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (okq[q]) {
        const size_t o = orow[q] + (size_t)t * HAR;
        out[o] = hn[q];
        if (outT != nullptr) outT[o] = __float2bfloat16_rn(hn[q]);
        if (gates4 != nullptr) { gates4[o] = sv4[q]; cell[o] = cn[q]; }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; q++) {
    if (okq[q]) {
      if (hT != nullptr) hT[(size_t)(b0 + 2 * t4 + q) * HAR + col] = hprev[q];
      if (cT != nullptr) cT[(size_t)(b0 + 2 * t4 + q) * HAR + col] = cprev[q];
    }
  }
  cluster.sync();  // nobody leaves while a peer may still be writing into it
}

// ---------------------------------------------------------------------------------------------------------
// BPTT.  Same launch shape.  Resident: A[i][gate index] = W_hh[gate index][64*rank + 16*(w&3) + i], gate-index range of
// warp w: [ (w>>2)*4HAR/2, +4HAR/2 ).  dg (B, S, 4*HAR) bf16 = gradient of the four pre-activations (it feeds BOTH hoisted
// weight-gradient GEMMs and d(input): in an LSTM the same vector multiplies W_ih and W_hh).
// ---------------------------------------------------------------------------------------------------------
template <int HAR>
__global__ void __launch_bounds__(256, 1)
lstm_rec_bwd_mma_kernel(const float* __restrict__ dout, const float* __restrict__ c0, const uint2* __restrict__ gates4,
                        const float* __restrict__ cell, const float* __restrict__ w_hh, bf16* __restrict__ dg,
                        float* __restrict__ db_ih, float* __restrict__ db_hh, int B, int S) {
  pdl_wait();
  pdl_trigger();
  constexpr int G = 4 * HAR, KS = G / 16, KSH = KS / 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 3, kh = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  // d(gates)_t as the B operand, [row][sequence]; row(gate index k) = (u / 64) * 256 + gate * 64 + u % 64 with gate = k / HAR,
  // u = k % HAR: the 4 x 64 rows a source CTA produces are contiguous, so that it delivers them with ONE bulk copy
  __shared__ __align__(128) bf16 ds[2][G][BT];
  __shared__ float part[4][2][2][32];
  __shared__ __align__(128) bf16 dstage[2][4][HC][BT];
  __shared__ __align__(8) uint64_t dbar[2];
  constexpr uint32_t kStepBytes = G * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&dbar[0], 1);
    mbar_init(&dbar[1], 1);
    fence_mbar_init_cluster();
  }
  uint32_t wf[KSH][4];
  {
    const int c0u = HC * rank + 16 * ub + g;
#pragma unroll
    for (int ks = 0; ks < KSH; ks++) {
      const int k = (kh * KSH + ks) * 16 + 2 * t4;
      wf[ks][0] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0u), __ldg(w_hh + (size_t)(k + 1) * HAR + c0u));
      wf[ks][1] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0u + 8), __ldg(w_hh + (size_t)(k + 1) * HAR + c0u + 8));
      wf[ks][2] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0u), __ldg(w_hh + (size_t)(k + 9) * HAR + c0u));
      wf[ks][3] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0u + 8), __ldg(w_hh + (size_t)(k + 9) * HAR + c0u + 8));
    }
  }
  const int col = HC * rank + 16 * ub + 8 * kh + g;
  float carry[2] = {0.f, 0.f};     // dh flowing back through W_hh
  float dccarry[2] = {0.f, 0.f};   // dc flowing back through the forget gate
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&dbar[0], kStepBytes);
    mbar_expect_tx(&dbar[1], kStepBytes);
  }
  cluster.sync();
  const uint32_t ds_local = s_u32(&ds[0][0][0]), bar_local = s_u32(&dbar[0]);
  const uint32_t pub_dst = mapa_u32(ds_local + (uint32_t)(4 * HC * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

  // finish step `it`: wait for its exchange, carry = d(gates) . W_hh[:, slice]
  auto consume = [&](int it) {
    const int tt = S - 1 - it, buf = tt & 1;
    mbar_wait(&dbar[buf], (it >> 1) & 1);
    if (threadIdx.x == 0 && it + 2 < S) mbar_expect_tx(&dbar[buf], kStepBytes);
    float acc[4][4];
#pragma unroll
    for (int a4 = 0; a4 < 4; a4++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[a4][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSH / 2; q++) {
      uint32_t bq4[4];
      const int kidx = kh * (G / 2) + 32 * q, gate = kidx / HAR, u = kidx - gate * HAR;
      ldsm_x4_t(bq4, s_u32(&ds[buf][(u / HC) * 4 * HC + gate * HC + (u % HC) + lane][0]));
      mma16816(acc[(2 * q) & 3], wf[2 * q], bq4[0], bq4[1]);
      mma16816(acc[(2 * q + 1) & 3], wf[2 * q + 1], bq4[2], bq4[3]);
    }
    float keep[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const float lo = (acc[0][q] + acc[1][q]) + (acc[2][q] + acc[3][q]);
      const float hi = (acc[0][2 + q] + acc[1][2 + q]) + (acc[2][2 + q] + acc[3][2 + q]);
      keep[q] = kh ? hi : lo;
      part[ub][kh][q][lane] = kh ? lo : hi;
    }
    pair_sync(ub);
#pragma unroll
    for (int q = 0; q < 2; q++) carry[q] = keep[q] + part[ub][1 - kh][q][lane];
  };

  bool okq[2];
  size_t orow[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int bq = b0 + 2 * t4 + q;
    okq[q] = bq < B;
    orow[q] = (size_t)(okq[q] ? bq : b0) * S * HAR + col;
  }
  // operands of a step are fetched ONE FULL STEP ahead of their use; the cell state of step t-1 is also step t-1's own c_t
  float dov_n[2], cp_n[2], cn_cur[2];
  uint2 g4_n[2];
  auto load_ops = [&](int tt) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const size_t o = orow[q] + (size_t)tt * HAR;
      dov_n[q] = dout[o];
      g4_n[q] = gates4[o];
      if (tt > 0) cp_n[q] = cell[o - HAR];
      else cp_n[q] = (c0 != nullptr && okq[q]) ? c0[(size_t)(b0 + 2 * t4 + q) * HAR + col] : 0.f;
    }
  };
#pragma unroll
  for (int q = 0; q < 2; q++) cn_cur[q] = cell[orow[q] + (size_t)(S - 1) * HAR];
  load_ops(S - 1);

  for (int it = 0; it < S; it++) {
    const int t = S - 1 - it, buf = t & 1;
    float dov[2], cp[2], cnv[2];
    uint2 g4[2];
#pragma unroll
    for (int q = 0; q < 2; q++) { dov[q] = dov_n[q]; cp[q] = cp_n[q]; g4[q] = g4_n[q]; cnv[q] = cn_cur[q]; cn_cur[q] = cp_n[q]; }
    if (t > 0) load_ops(t - 1);
    if (it > 0) consume(it - 1);
    float da[4][2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      da[0][q] = da[1][q] = da[2][q] = da[3][q] = 0.f;
      if (okq[q]) {
        const float dh = carry[q] + dov[q];
        const float ig = __uint_as_float(g4[q].x << 16), fg = __uint_as_float(g4[q].x & 0xffff0000u);
        const float gg = __uint_as_float(g4[q].y << 16), og = __uint_as_float(g4[q].y & 0xffff0000u);
        const float tc = tanh_fast(cnv[q]);
        const float dcl = fmaf(dh * og, 1.f - tc * tc, dccarry[q]);
        da[0][q] = dcl * gg * ig * (1.f - ig);
        da[1][q] = dcl * cp[q] * fg * (1.f - fg);
        da[2][q] = dcl * ig * (1.f - gg * gg);
        da[3][q] = dh * tc * og * (1.f - og);
        dccarry[q] = dcl * fg;
        sb[0] += da[0][q]; sb[1] += da[1][q]; sb[2] += da[2][q]; sb[3] += da[3][q];
      }
    }
    const int urow = 16 * ub + 8 * kh + g;
#pragma unroll
    for (int gt = 0; gt < 4; gt++) *reinterpret_cast<uint32_t*>(&dstage[buf][gt][urow][2 * t4]) = pack_bf16(da[gt][0], da[gt][1]);
    fence_async_smem();
    if (warp == 0) {  // ONE 4 KB bulk copy per destination CTA once all 8 warps have parked their rows
      publish_sync();
      if (lane < CS) bulk_s2s(pub_dst + buf * (G * BT * 2), s_u32(&dstage[buf][0][0][0]), 4 * HC * BT * 2, pub_bar + buf * 8);
    } else {
      publish_arrive();
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (okq[q]) {
        bf16* pg = dg + (orow[q] - col + (size_t)t * HAR) * 4 + col;  // row (b, t) of the (B, S, 4*HAR) array
#pragma unroll
        for (int gt = 0; gt < 4; gt++) pg[gt * HAR] = __float2bfloat16_rn(da[gt][q]);
      }
    }
  }
  consume(S - 1);
  if (db_ih != nullptr) {  // both bias vectors enter every pre-activation with coefficient 1
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float v = sb[q];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      sb[q] = v;
    }
    if (t4 == 0) {
#pragma unroll
      for (int gt = 0; gt < 4; gt++) { atomicAdd(db_ih + gt * HAR + col, sb[gt]); atomicAdd(db_hh + gt * HAR + col, sb[gt]); }
    }
  }
  cluster.sync();
}


// ---------------------------------------------------------------------------------------------------------
// Wide variant (the split of gru_mma_wide.cu): a CTA owns 32 hidden units, the cluster has Har/32 CTAs (16 at Har = 512), the k
// range of a product is split over FOUR warps, the four warps of a 16-unit block swap partial sums through shared memory (one
// 128-thread named barrier; fixed summation order, so a sequence's result does not depend on its slot in the batch tile) and
// every thread finishes ONE (unit, sequence) pair.  Half as many products per warp and step as the 64-unit kernels at
// Har = 256, and the only tensor-core form of Har = 512 (4 gates x Har/64 k-steps x 4 = 128 registers of fragments).
// ---------------------------------------------------------------------------------------------------------
constexpr int HCW = 32, NKH = 4;

template <int HAR>
__global__ void __launch_bounds__(256, 1)
lstm_rec_fwd_wide_kernel(const bf16* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                         const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ out, bf16* __restrict__ outT,
                         uint2* __restrict__ gates4, float* __restrict__ cell, float* __restrict__ hT, float* __restrict__ cT,
                         int B, int S) {
  constexpr int KSW = HAR / 16 / NKH;
  static_assert(KSW % 2 == 0, "ldmatrix.x4 covers two k-steps");
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 1, kh = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;

  __shared__ __align__(128) bf16 hs[2][HAR][BT];
  __shared__ float part[2][NKH][16][32];                 // [gate*4 + e][lane]
  __shared__ __align__(128) bf16 hstage[2][HCW][BT];
  __shared__ __align__(8) uint64_t hbar[2];
  constexpr uint32_t kStepBytes = HAR * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_mbar_init_cluster();
  }
  uint32_t wf[4][KSW][4];
#pragma unroll
  for (int gt = 0; gt < 4; gt++) {
    const float* r0 = w_hh + (size_t)(gt * HAR + HCW * rank + 16 * ub + g) * HAR;
    const float* r1 = r0 + 8 * HAR;
#pragma unroll
    for (int ks = 0; ks < KSW; ks++) {
      const int k = (kh * KSW + ks) * 16 + 2 * t4;
      wf[gt][ks][0] = pack_bf16(__ldg(r0 + k), __ldg(r0 + k + 1));
      wf[gt][ks][1] = pack_bf16(__ldg(r1 + k), __ldg(r1 + k + 1));
      wf[gt][ks][2] = pack_bf16(__ldg(r0 + k + 8), __ldg(r0 + k + 9));
      wf[gt][ks][3] = pack_bf16(__ldg(r1 + k + 8), __ldg(r1 + k + 9));
    }
  }
  pdl_wait();
  pdl_trigger();
  // this thread finishes accumulator element e = kh: unit row g + 8*(kh>>1) of the 16-block, sequence 2*t4 + (kh&1)
  const int urow = 16 * ub + 8 * (kh >> 1) + g;
  const int col = HCW * rank + urow;
  const int sq = 2 * t4 + (kh & 1);
  const int bq = b0 + sq;
  const bool ok = bq < B;
  float bh[4];
#pragma unroll
  for (int gt = 0; gt < 4; gt++) bh[gt] = __ldg(b_hh + gt * HAR + col);
  float hprev = (h0 != nullptr && ok) ? h0[(size_t)bq * HAR + col] : 0.f;
  float cprev = (c0 != nullptr && ok) ? c0[(size_t)bq * HAR + col] : 0.f;
  for (int i = threadIdx.x; i < HAR * BT; i += blockDim.x) {
    const int k = i / BT, bb = b0 + (i - k * BT);
    hs[0][k][i - k * BT] = __float2bfloat16_rn((h0 != nullptr && bb < B) ? h0[(size_t)bb * HAR + k] : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&hbar[1], kStepBytes);
    mbar_expect_tx(&hbar[0], kStepBytes);
  }
  cluster.sync();
  const uint32_t hs_local = s_u32(&hs[0][0][0]), bar_local = s_u32(&hbar[0]);
  const uint32_t pub_dst = mapa_u32(hs_local + (uint32_t)(HCW * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

  const size_t row0 = (size_t)(ok ? bq : b0) * S;
  const bf16* gip = gi + row0 * 4 * HAR + col;
  size_t o = row0 * HAR + col;  // element offset of (sequence, t, unit) in the (B, S, HAR) arrays
  bf16 gq_raw[4];
  auto load_gi = [&]() {
#pragma unroll
    for (int gt = 0; gt < 4; gt++) gq_raw[gt] = gip[gt * HAR];
    gip += 4 * HAR;
  };
  load_gi();

  for (int t = 0; t < S; t++) {
    const int cur = t & 1, nxt = cur ^ 1;
    float gq[4];
#pragma unroll
    for (int gt = 0; gt < 4; gt++) gq[gt] = __bfloat162float(gq_raw[gt]);
    if (t + 1 < S) load_gi();
    if (t > 0) {
      mbar_wait(&hbar[cur], ((t - 1 - (cur ^ 1)) >> 1) & 1);
      if (threadIdx.x == 0 && t + 1 < S) mbar_expect_tx(&hbar[cur], kStepBytes);
    }
    float acc[4][2][4];
#pragma unroll
    for (int gt = 0; gt < 4; gt++)
#pragma unroll
      for (int ch = 0; ch < 2; ch++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[gt][ch][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSW / 2; q++) {
      uint32_t bq4[4];
      ldsm_x4_t(bq4, s_u32(&hs[cur][kh * (HAR / NKH) + 32 * q + lane][0]));
#pragma unroll
      for (int gt = 0; gt < 4; gt++) {
        mma16816(acc[gt][0], wf[gt][2 * q], bq4[0], bq4[1]);
        mma16816(acc[gt][1], wf[gt][2 * q + 1], bq4[2], bq4[3]);
      }
    }
#pragma unroll
    for (int gt = 0; gt < 4; gt++)
#pragma unroll
      for (int e = 0; e < 4; e++) part[ub][kh][gt * 4 + e][lane] = acc[gt][0][e] + acc[gt][1][e];
    quad_sync(ub);
    float a4[4];
#pragma unroll
    for (int gt = 0; gt < 4; gt++)
      a4[gt] = ((part[ub][0][gt * 4 + kh][lane] + part[ub][1][gt * 4 + kh][lane]) + part[ub][2][gt * 4 + kh][lane]) +
               part[ub][3][gt * 4 + kh][lane];
    const float ig = sigmoid_fast(gq[0] + a4[0] + bh[0]);
    const float fg = sigmoid_fast(gq[1] + a4[1] + bh[1]);
    const float gg = tanh_fast(gq[2] + a4[2] + bh[2]);
    const float og = sigmoid_fast(gq[3] + a4[3] + bh[3]);
    const float cn = fmaf(fg, cprev, ig * gg);
    const float hn = og * tanh_fast(cn);
    cprev = cn;
    hprev = hn;
    if (t + 1 < S) {
      hstage[cur][urow][sq] = __float2bfloat16_rn(hn);
      fence_async_smem();
      if (warp == 0) {
        publish_sync();
        if (lane < CS) bulk_s2s(pub_dst + nxt * (HAR * BT * 2), s_u32(&hstage[cur][0][0]), HCW * BT * 2, pub_bar + nxt * 8);
      } else {
        publish_arrive();
      }
    }
    if (ok) {
      out[o] = hn;
      if (outT != nullptr) outT[o] = __float2bfloat16_rn(hn);
      if (gates4 != nullptr) { gates4[o] = make_uint2(pack_bf16(ig, fg), pack_bf16(gg, og)); cell[o] = cn; }
    }
    o += HAR;
  }
  if (ok) {
    if (hT != nullptr) hT[(size_t)bq * HAR + col] = hprev;
    if (cT != nullptr) cT[(size_t)bq * HAR + col] = cprev;
  }
  cluster.sync();
}

template <int HAR>
__global__ void __launch_bounds__(256, 1)
lstm_rec_bwd_wide_kernel(const float* __restrict__ dout, const float* __restrict__ c0, const uint2* __restrict__ gates4,
                         const float* __restrict__ cell, const float* __restrict__ w_hh, bf16* __restrict__ dg,
                         float* __restrict__ db_ih, float* __restrict__ db_hh, int B, int S) {
  pdl_wait();
  pdl_trigger();
  constexpr int G = 4 * HAR, KSW = G / 16 / NKH;
  static_assert(KSW % 2 == 0 && (G / NKH) % 32 == 0, "k range of a warp is a whole number of 32-row blocks");
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 1, kh = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;

  // d(gates)_t as the B operand, [row][sequence]; row(gate index k) = (u / 32) * 128 + gate * 32 + u % 32: the 4 x 32 rows a
  // source CTA produces are contiguous (ONE bulk copy).  2 x 4*HAR x 16 B = 64 KB at HAR = 512: dynamic shared memory
  extern __shared__ __align__(128) unsigned char dyn[];
  bf16 (*ds)[G][BT] = reinterpret_cast<bf16 (*)[G][BT]>(dyn);
  __shared__ float part[2][NKH][4][32];
  __shared__ __align__(128) bf16 dstage[2][4][HCW][BT];
  __shared__ __align__(8) uint64_t dbar[2];
  constexpr uint32_t kStepBytes = G * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&dbar[0], 1);
    mbar_init(&dbar[1], 1);
    fence_mbar_init_cluster();
  }
  uint32_t wf[KSW][4];
  {
    const int c0u = HCW * rank + 16 * ub + g;
#pragma unroll
    for (int ks = 0; ks < KSW; ks++) {
      const int k = (kh * KSW + ks) * 16 + 2 * t4;
      wf[ks][0] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0u), __ldg(w_hh + (size_t)(k + 1) * HAR + c0u));
      wf[ks][1] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0u + 8), __ldg(w_hh + (size_t)(k + 1) * HAR + c0u + 8));
      wf[ks][2] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0u), __ldg(w_hh + (size_t)(k + 9) * HAR + c0u));
      wf[ks][3] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0u + 8), __ldg(w_hh + (size_t)(k + 9) * HAR + c0u + 8));
    }
  }
  const int urow = 16 * ub + 8 * (kh >> 1) + g;
  const int col = HCW * rank + urow;
  const int sq = 2 * t4 + (kh & 1);
  const int bq = b0 + sq;
  const bool ok = bq < B;
  float carry = 0.f, dccarry = 0.f;
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&dbar[0], kStepBytes);
    mbar_expect_tx(&dbar[1], kStepBytes);
  }
  cluster.sync();
  const uint32_t ds_local = s_u32(&ds[0][0][0]), bar_local = s_u32(&dbar[0]);
  const uint32_t pub_dst = mapa_u32(ds_local + (uint32_t)(4 * HCW * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

  auto consume = [&](int it) {
    const int tt = S - 1 - it, buf = tt & 1;
    mbar_wait(&dbar[buf], (it >> 1) & 1);
    if (threadIdx.x == 0 && it + 2 < S) mbar_expect_tx(&dbar[buf], kStepBytes);
    float acc[4][4];
#pragma unroll
    for (int a4 = 0; a4 < 4; a4++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[a4][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSW / 2; q++) {
      uint32_t bq4[4];
      const int kidx = kh * (G / NKH) + 32 * q, gate = kidx / HAR, u = kidx - gate * HAR;
      ldsm_x4_t(bq4, s_u32(&ds[buf][(u / HCW) * 4 * HCW + gate * HCW + lane][0]));
      mma16816(acc[(2 * q) & 3], wf[2 * q], bq4[0], bq4[1]);
      mma16816(acc[(2 * q + 1) & 3], wf[2 * q + 1], bq4[2], bq4[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) part[ub][kh][e][lane] = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
    quad_sync(ub);
    carry = ((part[ub][0][kh][lane] + part[ub][1][kh][lane]) + part[ub][2][kh][lane]) + part[ub][3][kh][lane];
  };

  const size_t obase = (size_t)(ok ? bq : b0) * S * HAR + col;
  float dov_n, cp_n, cn_cur;
  uint2 g4_n;
  auto load_ops = [&](int tt) {
    const size_t o = obase + (size_t)tt * HAR;
    dov_n = dout[o];
    g4_n = gates4[o];
    if (tt > 0) cp_n = cell[o - HAR];
    else cp_n = (c0 != nullptr && ok) ? c0[(size_t)bq * HAR + col] : 0.f;
  };
  cn_cur = cell[obase + (size_t)(S - 1) * HAR];
  load_ops(S - 1);

  for (int it = 0; it < S; it++) {
    const int t = S - 1 - it, buf = t & 1;
    const float dov = dov_n, cp = cp_n, cnv = cn_cur;
    const uint2 g4 = g4_n;
    cn_cur = cp_n;  // the cell state of step t-1 is also step t-1's own c_t
    if (t > 0) load_ops(t - 1);
    if (it > 0) consume(it - 1);
    float da[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
      const float dh = carry + dov;
      const float ig = __uint_as_float(g4.x << 16), fg = __uint_as_float(g4.x & 0xffff0000u);
      const float gg = __uint_as_float(g4.y << 16), og = __uint_as_float(g4.y & 0xffff0000u);
      const float tc = tanh_fast(cnv);
      const float dcl = fmaf(dh * og, 1.f - tc * tc, dccarry);
      da[0] = dcl * gg * ig * (1.f - ig);
      da[1] = dcl * cp * fg * (1.f - fg);
      da[2] = dcl * ig * (1.f - gg * gg);
      da[3] = dh * tc * og * (1.f - og);
      dccarry = dcl * fg;
      sb[0] += da[0]; sb[1] += da[1]; sb[2] += da[2]; sb[3] += da[3];
    }
#pragma unroll
    for (int gt = 0; gt < 4; gt++) dstage[buf][gt][urow][sq] = __float2bfloat16_rn(da[gt]);
    fence_async_smem();
    if (warp == 0) {
      publish_sync();
      if (lane < CS) bulk_s2s(pub_dst + buf * (G * BT * 2), s_u32(&dstage[buf][0][0][0]), 4 * HCW * BT * 2, pub_bar + buf * 8);
    } else {
      publish_arrive();
    }
    if (ok) {
      bf16* pg = dg + (obase - col + (size_t)t * HAR) * 4 + col;  // row (b, t) of the (B, S, 4*HAR) array
#pragma unroll
      for (int gt = 0; gt < 4; gt++) pg[gt * HAR] = __float2bfloat16_rn(da[gt]);
    }
  }
  consume(S - 1);
  if (db_ih != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float v = sb[q];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      sb[q] = v;
    }
    if (t4 == 0) {
#pragma unroll
      for (int gt = 0; gt < 4; gt++) { atomicAdd(db_ih + gt * HAR + col, sb[gt]); atomicAdd(db_hh + gt * HAR + col, sb[gt]); }
    }
  }
  cluster.sync();
}

}  // namespace

// CPC_B200_LSTM_MMA=0: CUDA-core recurrence; CPC_B200_LSTM_WIDE=0 / 1: never / always the 32-unit split (default: Har >= 256)
bool lstm_mma_supported(int Har) {
  static const bool off = []() { const char* e = getenv("CPC_B200_LSTM_MMA"); return e && atoi(e) == 0; }();
  return !off && (Har == 64 || Har == 128 || Har == 256 || Har == 512);
}
static bool lstm_wide(int Har) {
  static const int mode = []() { const char* e = getenv("CPC_B200_LSTM_WIDE"); return e ? atoi(e) : -1; }();
  if (Har == 512) return true;
  if (Har == 64) return false;
  return mode == 1 || (mode == -1 && Har >= 256);
}

// gates4: the four saved gates (i, f, g, o) of every (b, t, unit) as ONE array of bf16 quadruples (8-byte stores)
int lstm_rec_fwd_mma(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, const float* c0, float* out, bf16* outT,
                     void* gates4, float* cell, float* hT, float* cT, int B, int S, int Har, cudaStream_t st) {
  uint2* g4 = static_cast<uint2*>(gates4);
  void* args[] = {&gi, &w_hh, &b_hh, &h0, &c0, &out, &outT, &g4, &cell, &hT, &cT, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (lstm_wide(Har)) {
    if (Har == 512) return launch_cluster("lstm_rec_fwd_wide", lstm_rec_fwd_wide_kernel<512>, 16, ncl, st, args);
    if (Har == 256) return launch_cluster("lstm_rec_fwd_wide", lstm_rec_fwd_wide_kernel<256>, 8, ncl, st, args);
    return launch_cluster("lstm_rec_fwd_wide", lstm_rec_fwd_wide_kernel<128>, 4, ncl, st, args);
  }
  if (Har == 256) return launch_cluster("lstm_rec_fwd_mma", lstm_rec_fwd_mma_kernel<256>, 4, ncl, st, args);
  if (Har == 128) return launch_cluster("lstm_rec_fwd_mma", lstm_rec_fwd_mma_kernel<128>, 2, ncl, st, args);
  return launch_cluster("lstm_rec_fwd_mma", lstm_rec_fwd_mma_kernel<64>, 1, ncl, st, args);
}
int lstm_rec_bwd_mma(const float* dout, const float* c0, const void* gates4, const float* cell, const float* w_hh, bf16* dg,
                     float* db_ih, float* db_hh, int B, int S, int Har, cudaStream_t st) {
  const uint2* g4 = static_cast<const uint2*>(gates4);
  void* args[] = {&dout, &c0, &g4, &cell, &w_hh, &dg, &db_ih, &db_hh, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (lstm_wide(Har)) {
    const size_t dyn = (size_t)2 * 4 * Har * BT * 2;
    if (Har == 512) return launch_cluster("lstm_rec_bwd_wide", lstm_rec_bwd_wide_kernel<512>, 16, ncl, st, args, dyn);
    if (Har == 256) return launch_cluster("lstm_rec_bwd_wide", lstm_rec_bwd_wide_kernel<256>, 8, ncl, st, args, dyn);
    return launch_cluster("lstm_rec_bwd_wide", lstm_rec_bwd_wide_kernel<128>, 4, ncl, st, args, dyn);
  }
  if (Har == 256) return launch_cluster("lstm_rec_bwd_mma", lstm_rec_bwd_mma_kernel<256>, 4, ncl, st, args);
  if (Har == 128) return launch_cluster("lstm_rec_bwd_mma", lstm_rec_bwd_mma_kernel<128>, 2, ncl, st, args);
  return launch_cluster("lstm_rec_bwd_mma", lstm_rec_bwd_mma_kernel<64>, 1, ncl, st, args);
}

}  // namespace cpcb200
