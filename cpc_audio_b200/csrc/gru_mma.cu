// gru_mma.cu - tensor-core GRU recurrence for the bf16 path (reference: cpc/model.py:193 -> torch.nn.GRU).
//
// Same decomposition as gru.cu (cluster of Har/64 CTAs, CTA r owns hidden units [64r, 64r+64), hidden state
// exchanged through distributed shared memory, one cluster barrier per time step) but the per-step product
// runs on mma.sync.m16n8k16 with the CTA's slice of W_hh held in REGISTERS for the whole sequence:
//   forward : warp w owns hidden units 16w..16w+15 of the slice and their three gate rows (r, z, n):
//             3 m-tiles x Har/16 k-steps of A fragments = 12*Har/16 registers per thread (192 at Har=256);
//             D[gate rows x 8 sequences] = W_slice . h_{t-1}^T, then the gate math for the 4 (unit, sequence)
//             pairs a thread's accumulators cover happens in that same thread - no cross-warp exchange.
//   backward: warp w owns 16 outputs of dh_{t-1} = dgh_t . W_hh[:, slice]  (one m-tile x 3Har/16 k-steps).
// The B operand (h or dgh, bf16, [k][8 sequences], 16 B per row) is read with ldmatrix.trans.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cpcb200 {

namespace {

constexpr int HC = 64;   // hidden units per CTA
constexpr int BT = 8;    // sequences per cluster (= the n of m16n8k16)

#include "rec_mma.cuh"

// ---------------------------------------------------------------------------------------------------------
// forward.  block = 256 threads = 8 warps: warp w owns units 16*(w&3).. of the CTA's 64-unit slice and the k
// range [ (w>>2)*HAR/2, +HAR/2 ) of the product.  The two warps of a pair swap half of their partial sums
// through shared memory (64-thread named barrier), so each finishes 8 units x 8 sequences: gate math
// (tanh.approx), one 8-byte store of the four saved gates per element, one 128-byte bulk copy per peer CTA.
// cluster = HAR/64 CTAs, grid = cluster * ceil(B/8).
// ---------------------------------------------------------------------------------------------------------
template <int HAR>
__global__ void __launch_bounds__(256, 1)
gru_rec_fwd_mma_kernel(const bf16* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                       const float* __restrict__ h0, float* __restrict__ c, bf16* __restrict__ cT, uint2* __restrict__ gates4,
                       float* __restrict__ hT, int B, int S) {
  constexpr int KS = HAR / 16, KSH = KS / 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 3, kh = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  __shared__ __align__(128) bf16 hs[2][HAR][BT];  // h_{t-1} as the B operand: [k][sequence]
  __shared__ float part[4][2][6][32];              // partial sums handed to the partner warp, per unit block
  __shared__ __align__(128) bf16 hstage[2][HC][BT];    // the CTA's 64 new units, staged for ONE bulk copy per destination CTA
  __shared__ __align__(8) uint64_t hbar[2];        // hbar[b] completes when buffer b holds a full new state
  constexpr uint32_t kStepBytes = HAR * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_mbar_init_cluster();
  }

  // resident A fragments: gate gt, rows 16*ub + {g, g+8} of this CTA's slice, k-steps of this warp's half
  uint32_t wf[3][KSH][4];
#pragma unroll
  for (int gt = 0; gt < 3; gt++) {
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
  // W_hh was loaded ahead of the dependency wait: the kernel in front of this one in gru_fwd is always the input-projection
  // GEMM, and with wait-then-trigger ordering only the immediate predecessor can still be running
  pdl_wait();
  pdl_trigger();
  // this thread finishes unit `col` (row g + 8*kh of the 16-block) for sequences b0 + 2*t4 + {0, 1}
  const int col = HC * rank + 16 * ub + 8 * kh + g;
  float bh[3];
#pragma unroll
  for (int gt = 0; gt < 3; gt++) bh[gt] = __ldg(b_hh + gt * HAR + col);
  float hprev[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int bq = b0 + 2 * t4 + q;
    hprev[q] = (h0 != nullptr && bq < B) ? h0[(size_t)bq * HAR + col] : 0.f;
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
  // lane l < CS of warp 0 delivers to CTA l: its rows of that CTA's hs[0] and that CTA's hbar[0]
  const uint32_t pub_dst = mapa_u32(hs_local + (uint32_t)(HC * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

  // per-thread streams (sequence q of this thread): pointers advance by one time step per iteration instead of being
  // rebuilt from (b, t) with 64-bit multiplies on the critical path; out-of-range sequences alias sequence b0 for loads
  // and are masked for stores
  bool okq[2];
  const bf16* gip[2];
  float* cptr[2];
  bf16* ctptr[2];
  uint2* g4ptr[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int bq = b0 + 2 * t4 + q;
    okq[q] = bq < B;
    const size_t row0 = (size_t)(okq[q] ? bq : b0) * S;
    gip[q] = gi + row0 * 3 * HAR + col;
    cptr[q] = c + row0 * HAR + col;
    ctptr[q] = cT + row0 * HAR + col;
    g4ptr[q] = gates4 + row0 * HAR + col;
  }
  bf16 gq_raw[3][2];
  auto load_gi = [&]() {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      gq_raw[0][q] = gip[q][0]; gq_raw[1][q] = gip[q][HAR]; gq_raw[2][q] = gip[q][2 * HAR];
      gip[q] += 3 * HAR;
    }
  };
  load_gi();

  for (int t = 0; t < S; t++) {
    const int cur = t & 1, nxt = cur ^ 1;
    float gq[3][2];
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
#pragma unroll
      for (int q = 0; q < 2; q++) gq[gt][q] = __bfloat162float(gq_raw[gt][q]);
    if (t + 1 < S) load_gi();  // in flight during this step's product and exchange
    if (t > 0) {
      mbar_wait(&hbar[cur], ((t - 1 - (cur ^ 1)) >> 1) & 1);
      if (threadIdx.x == 0 && t + 1 < S) mbar_expect_tx(&hbar[cur], kStepBytes);  // re-arm for step t+2
    }
    float acc[3][2][4];
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
#pragma unroll
      for (int ch = 0; ch < 2; ch++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[gt][ch][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSH / 2; q++) {
      uint32_t bq4[4];
      ldsm_x4_t(bq4, s_u32(&hs[cur][kh * (HAR / 2) + 32 * q + lane][0]));
#pragma unroll
      for (int gt = 0; gt < 3; gt++) {
        mma16816(acc[gt][0], wf[gt][2 * q], bq4[0], bq4[1]);
        mma16816(acc[gt][1], wf[gt][2 * q + 1], bq4[2], bq4[3]);
      }
    }
    // accumulator e: row g (e = 0,1) / g+8 (e = 2,3), sequence 2*t4 + (e&1).  Keep rows g + 8*kh, hand the other
    // two of every gate to the partner warp.
    float mine[3][2];
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const float lo = acc[gt][0][q] + acc[gt][1][q], hi = acc[gt][0][2 + q] + acc[gt][1][2 + q];
        const float keep = kh ? hi : lo, give = kh ? lo : hi;
        mine[gt][q] = keep;
        part[ub][kh][gt * 2 + q][lane] = give;
      }
    pair_sync(ub);
    float hn[2];
    uint2 sv4[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const float ar = mine[0][q] + part[ub][1 - kh][q][lane];
      const float au = mine[1][q] + part[ub][1 - kh][2 + q][lane];
      const float an = mine[2][q] + part[ub][1 - kh][4 + q][lane];
      const float ghn = an + bh[2];
      const float rg = sigmoid_fast(gq[0][q] + ar + bh[0]);
      const float ug = sigmoid_fast(gq[1][q] + au + bh[1]);
      const float ng = tanh_fast(gq[2][q] + rg * ghn);
      hn[q] = fmaf(ug, hprev[q] - ng, ng);  // (1-u) n + u h
      hprev[q] = hn[q];
      sv4[q] = make_uint2(pack_bf16(rg, ug), pack_bf16(ng, ghn));
    }
    // publish the CTA's 64 units FIRST (the peers wait for them): every warp parks its 8 units in the staging tile and
    // arrives on a named barrier; warp 0 waits for all of them and issues ONE 1 KB bulk copy per destination CTA
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
    // then the outputs and the gates saved for backward (nobody waits for these)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (okq[q]) {
        *cptr[q] = hn[q];
        *ctptr[q] = __float2bfloat16_rn(hn[q]);
        *g4ptr[q] = sv4[q];
      }
      cptr[q] += HAR; ctptr[q] += HAR; g4ptr[q] += HAR;
    }
  }
  if (hT != nullptr) {
#pragma unroll
    for (int q = 0; q < 2; q++)
      if (okq[q]) hT[(size_t)(b0 + 2 * t4 + q) * HAR + col] = hprev[q];
  }
  cluster.sync();  // nobody leaves while a peer may still be writing into it
}

// ---------------------------------------------------------------------------------------------------------
// BPTT.  Same launch shape.  Resident: A[i][gate index] = W_hh[gate index][64*rank + 16*(w&3) + i], gate-index
// range of warp w: [ (w>>2)*3HAR/2, +3HAR/2 ).  Each warp finishes 8 units x 8 sequences per step.
// ---------------------------------------------------------------------------------------------------------
template <int HAR>
__global__ void __launch_bounds__(256, 1)
gru_rec_bwd_mma_kernel(const float* __restrict__ dc, const float* __restrict__ c, const float* __restrict__ h0,
                       const uint2* __restrict__ gates4, const float* __restrict__ w_hh, bf16* __restrict__ dgi,
                       bf16* __restrict__ dgh, float* __restrict__ dh0, float* __restrict__ db_ih, float* __restrict__ db_hh,
                       int B, int S) {
  pdl_wait();
  pdl_trigger();
  constexpr int G = 3 * HAR, KS = G / 16, KSH = KS / 2;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 3, kh = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  // dgh_t as the B operand, [row][sequence]; row(gate index k) = (u / 64) * 192 + gate * 64 + u % 64 with gate = k / HAR,
  // u = k % HAR: the 3 x 64 rows a source CTA produces are contiguous, so that it delivers them with ONE bulk copy
  __shared__ __align__(128) bf16 ds[2][G][BT];
  __shared__ float part[4][2][2][32];
  __shared__ __align__(128) bf16 dstage[2][3][HC][BT];
  __shared__ __align__(8) uint64_t dbar[2];
  constexpr uint32_t kStepBytes = G * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&dbar[0], 1);
    mbar_init(&dbar[1], 1);
    fence_mbar_init_cluster();
  }

  uint32_t wf[KSH][4];
  {
    const int c0 = HC * rank + 16 * ub + g;
#pragma unroll
    for (int ks = 0; ks < KSH; ks++) {
      const int k = (kh * KSH + ks) * 16 + 2 * t4;
      wf[ks][0] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0), __ldg(w_hh + (size_t)(k + 1) * HAR + c0));
      wf[ks][1] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0 + 8), __ldg(w_hh + (size_t)(k + 1) * HAR + c0 + 8));
      wf[ks][2] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0), __ldg(w_hh + (size_t)(k + 9) * HAR + c0));
      wf[ks][3] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0 + 8), __ldg(w_hh + (size_t)(k + 9) * HAR + c0 + 8));
    }
  }
  const int col = HC * rank + 16 * ub + 8 * kh + g;  // the unit this thread finishes, sequences b0 + 2*t4 + {0,1}
  float carry[2] = {0.f, 0.f};
  float direct[2] = {0.f, 0.f};
  float sb[4] = {0.f, 0.f, 0.f, 0.f};  // sums over (t, sequence) of dr, du, dn, dn*r for this unit
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&dbar[0], kStepBytes);
    mbar_expect_tx(&dbar[1], kStepBytes);
  }
  cluster.sync();
  const uint32_t ds_local = s_u32(&ds[0][0][0]), bar_local = s_u32(&dbar[0]);
  const uint32_t pub_dst = mapa_u32(ds_local + (uint32_t)(3 * HC * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

  // finish step `it`: wait for its exchange, carry = direct + dgh . W_hh[:, slice]
  auto consume = [&](int it) {
    const int tt = S - 1 - it, buf = tt & 1;
    mbar_wait(&dbar[buf], (it >> 1) & 1);
    if (threadIdx.x == 0 && it + 2 < S) mbar_expect_tx(&dbar[buf], kStepBytes);
    float acc[3][4];
#pragma unroll
    for (int a3 = 0; a3 < 3; a3++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[a3][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSH / 2; q++) {
      uint32_t bq4[4];
      const int kidx = kh * (G / 2) + 32 * q, gate = kidx / HAR, u = kidx - gate * HAR;
      ldsm_x4_t(bq4, s_u32(&ds[buf][(u / HC) * 3 * HC + gate * HC + (u % HC) + lane][0]));
      mma16816(acc[q % 3], wf[2 * q], bq4[0], bq4[1]);
      mma16816(acc[(q + 1) % 3], wf[2 * q + 1], bq4[2], bq4[3]);
    }
    float keep[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const float lo = (acc[0][q] + acc[1][q]) + acc[2][q], hi = (acc[0][2 + q] + acc[1][2 + q]) + acc[2][2 + q];
      keep[q] = kh ? hi : lo;
      part[ub][kh][q][lane] = kh ? lo : hi;
    }
    pair_sync(ub);
#pragma unroll
    for (int q = 0; q < 2; q++) carry[q] = direct[q] + keep[q] + part[ub][1 - kh][q][lane];
  };

  // per-thread streams, walking backwards in time (see the forward kernel)
  bool okq[2];
  const float* dcp[2];
  const float* hpp[2];   // h_{t-1}
  const uint2* g4p[2];
  bf16* dgip[2];
  bf16* dghp[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const int bq = b0 + 2 * t4 + q;
    okq[q] = bq < B;
    const size_t last = (size_t)(okq[q] ? bq : b0) * S + (S - 1);
    dcp[q] = dc + last * HAR + col;
    hpp[q] = c + last * HAR + col - HAR;
    g4p[q] = gates4 + last * HAR + col;
    dgip[q] = dgi + last * G + col;
    dghp[q] = dgh + last * G + col;
  }

  // operands of a step (dc_t, saved gates, h_{t-1}) are fetched ONE FULL STEP ahead of their use: a load issued at the top
  // of the step that consumes it is still in flight (L2 / HBM latency) when the previous step's product has finished
  float dcv_n[2], hp_n[2];
  uint2 g4_n[2];
  auto load_ops = [&](int tt) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      dcv_n[q] = *dcp[q];
      g4_n[q] = *g4p[q];
      if (tt > 0) hp_n[q] = *hpp[q];
      else hp_n[q] = (h0 != nullptr && okq[q]) ? h0[(size_t)(b0 + 2 * t4 + q) * HAR + col] : 0.f;
      dcp[q] -= HAR; g4p[q] -= HAR; hpp[q] -= HAR;
    }
  };
  load_ops(S - 1);

  for (int it = 0; it < S; it++) {
    const int t = S - 1 - it, buf = t & 1;
    // (A) this step's operands were loaded during the previous step; start the loads of the next one
    float dcv[2], hp[2];
    uint2 g4[2];
#pragma unroll
    for (int q = 0; q < 2; q++) { dcv[q] = dcv_n[q]; hp[q] = hp_n[q]; g4[q] = g4_n[q]; }
    if (t > 0) load_ops(t - 1);
    // (B) finish the previous step
    if (it > 0) consume(it - 1);
    // (C) gate gradients of step t, publish dgh_t
    float dr[2], du[2], dnr[2], dnv[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      dr[q] = du[q] = dnr[q] = dnv[q] = direct[q] = 0.f;
      if (okq[q]) {
        const float dh = carry[q] + dcv[q];
        const float rg = __uint_as_float(g4[q].x << 16), ug = __uint_as_float(g4[q].x & 0xffff0000u);
        const float ng = __uint_as_float(g4[q].y << 16), hnv = __uint_as_float(g4[q].y & 0xffff0000u);
        const float dn = dh * (1.f - ug) * (1.f - ng * ng);
        du[q] = dh * (hp[q] - ng) * ug * (1.f - ug);
        dr[q] = dn * hnv * rg * (1.f - rg);
        dnr[q] = dn * rg;
        dnv[q] = dn;
        direct[q] = dh * ug;
        sb[0] += dr[q]; sb[1] += du[q]; sb[2] += dn; sb[3] += dnr[q];
      }
    }
    // publish dgh_t first (the peers wait for it), then write the hoisted-GEMM operands
    const int urow = 16 * ub + 8 * kh + g;
    *reinterpret_cast<uint32_t*>(&dstage[buf][0][urow][2 * t4]) = pack_bf16(dr[0], dr[1]);
    *reinterpret_cast<uint32_t*>(&dstage[buf][1][urow][2 * t4]) = pack_bf16(du[0], du[1]);
    *reinterpret_cast<uint32_t*>(&dstage[buf][2][urow][2 * t4]) = pack_bf16(dnr[0], dnr[1]);
    fence_async_smem();
    if (warp == 0) {  // ONE 3 KB bulk copy per destination CTA once all 8 warps have parked their rows
      publish_sync();
      if (lane < CS) bulk_s2s(pub_dst + buf * (G * BT * 2), s_u32(&dstage[buf][0][0][0]), 3 * HC * BT * 2, pub_bar + buf * 8);
    } else {
      publish_arrive();
    }
    const ptrdiff_t back = -(ptrdiff_t)it * G;  // the store addresses are rebuilt from the fixed bases every step
#pragma unroll
    for (int q = 0; q < 2; q++) {
      if (okq[q]) {
        bf16* pi = dgip[q] + back;
        bf16* ph = dghp[q] + back;
        pi[0] = __float2bfloat16_rn(dr[q]); pi[HAR] = __float2bfloat16_rn(du[q]); pi[2 * HAR] = __float2bfloat16_rn(dnv[q]);
        ph[0] = __float2bfloat16_rn(dr[q]); ph[HAR] = __float2bfloat16_rn(du[q]); ph[2 * HAR] = __float2bfloat16_rn(dnr[q]);
      }
    }
  }
  consume(S - 1);
  if (db_ih != nullptr) {  // bias gradients: db_ih = sum(dr, du, dn), db_hh = sum(dr, du, dn*r)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      float v = sb[q];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      sb[q] = v;
    }
    if (t4 == 0) {
      atomicAdd(db_ih + col, sb[0]); atomicAdd(db_ih + HAR + col, sb[1]); atomicAdd(db_ih + 2 * HAR + col, sb[2]);
      atomicAdd(db_hh + col, sb[0]); atomicAdd(db_hh + HAR + col, sb[1]); atomicAdd(db_hh + 2 * HAR + col, sb[3]);
    }
  }
  if (dh0 != nullptr) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int bq = b0 + 2 * t4 + q;
      if (bq < B) dh0[(size_t)bq * HAR + col] = carry[q];
    }
  }
  cluster.sync();
}

}  // namespace

bool gru_mma_supported(int Har) { return Har == 64 || Har == 128 || Har == 256; }

int gru_rec_fwd_mma(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, float* c, bf16* cT, bf16* sR, bf16* sU,
                    bf16* sN, bf16* sHN, float* hT, int B, int S, int Har, cudaStream_t st) {
  // the four gate arrays are contiguous in the save buffer: the tensor-core kernels use them as ONE array of
  // {r, u, n, W_hn h + b_hn} bf16 quadruples (8-byte loads / stores)
  if (reinterpret_cast<char*>(sU) - reinterpret_cast<char*>(sR) != (ptrdiff_t)((size_t)B * S * Har * 2) ||
      reinterpret_cast<char*>(sHN) - reinterpret_cast<char*>(sR) != (ptrdiff_t)((size_t)3 * B * S * Har * 2))
    return fail(CPCB200_ERR_BAD_DIMS, "gru mma: gate arrays are not contiguous");
  uint2* gates4 = reinterpret_cast<uint2*>(sR);
  (void)sN;
  void* args[] = {&gi, &w_hh, &b_hh, &h0, &c, &cT, &gates4, &hT, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (Har == 256) return launch_cluster("gru_rec_fwd_mma", gru_rec_fwd_mma_kernel<256>, 4, ncl, st, args);
  if (Har == 128) return launch_cluster("gru_rec_fwd_mma", gru_rec_fwd_mma_kernel<128>, 2, ncl, st, args);
  return launch_cluster("gru_rec_fwd_mma", gru_rec_fwd_mma_kernel<64>, 1, ncl, st, args);
}
int gru_rec_bwd_mma(const float* dc, const float* c, const float* h0, const bf16* sR, const bf16* sU, const bf16* sN,
                    const bf16* sHN, const float* w_hh, bf16* dgi, bf16* dgh, float* dh0, float* db_ih, float* db_hh, int B, int S,
                    int Har, cudaStream_t st) {
  const uint2* gates4 = reinterpret_cast<const uint2*>(sR);
  (void)sU; (void)sN; (void)sHN;
  void* args[] = {&dc, &c, &h0, &gates4, &w_hh, &dgi, &dgh, &dh0, &db_ih, &db_hh, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (Har == 256) return launch_cluster("gru_rec_bwd_mma", gru_rec_bwd_mma_kernel<256>, 4, ncl, st, args);
  if (Har == 128) return launch_cluster("gru_rec_bwd_mma", gru_rec_bwd_mma_kernel<128>, 2, ncl, st, args);
  return launch_cluster("gru_rec_bwd_mma", gru_rec_bwd_mma_kernel<64>, 1, ncl, st, args);
}

}  // namespace cpcb200
