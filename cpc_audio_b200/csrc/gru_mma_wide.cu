// gru_mma_wide.cu - the tensor-core GRU recurrence of gru_mma.cu for wide hidden states (Har = 512: BASELINE config 5,
// cpc/model.py:193 with --hiddenGar 512) - and, as an option, a lower-latency split of Har = 256.
//
// gru_mma.cu keeps a 64-unit slice of W_hh per CTA in registers: 12 * Har / 16 registers per thread, 192 at Har = 256 and 384 at
// Har = 512 - which does not exist.  Here a CTA owns 32 hidden units, the cluster has Har / 32 CTAs (16 at Har = 512: the
// non-portable cluster size, one cluster per GPC) and the k range of a product is split over FOUR warps:
//   forward : warp w: units 16*(w&1).. of the CTA's 32, k range [ (w>>1) * Har/4, +Har/4 ): 3 gates x Har/64 k-steps x 4 = 96
//             registers at Har = 512.  The 4 warps of a unit block swap partial sums through shared memory (one 128-thread named
//             barrier) and every thread finishes ONE (unit, sequence) pair: accumulator element e = w>>1 of its fragment.
//   backward: A[i][gate index] = W_hh[gate index][32*rank + 16*(w&1) + i], gate-index range [ (w>>1) * 3Har/4, +3Har/4 ).
// Exchange, staging and the saved-gate layout are those of gru_mma.cu (one 512 B / 1.5 KB bulk copy per peer CTA and step).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cpcb200 {

namespace {

constexpr int HCW = 32;  // hidden units per CTA
constexpr int BT = 8;    // sequences per cluster (= the n of m16n8k16)
constexpr int NKH = 4;   // warps that split the k range of one unit block

#include "rec_mma.cuh"

template <int HAR>
__global__ void __launch_bounds__(256, 1)
gru_rec_fwd_wide_kernel(const bf16* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                        const float* __restrict__ h0, float* __restrict__ c, bf16* __restrict__ cT, uint2* __restrict__ gates4,
                        float* __restrict__ hT, int B, int S) {
  constexpr int KSW = HAR / 16 / NKH;  // k-steps per warp
  static_assert(KSW % 2 == 0, "ldmatrix.x4 covers two k-steps");
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 1, kh = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;

  __shared__ __align__(128) bf16 hs[2][HAR][BT];       // h_{t-1} as the B operand: [k][sequence]
  __shared__ float part[2][NKH][12][32];                 // partial sums of the 4 warps of a unit block: [gate*4 + e][lane]
  __shared__ __align__(128) bf16 hstage[2][HCW][BT];    // the CTA's 32 new units, staged for ONE bulk copy per destination CTA
  __shared__ __align__(8) uint64_t hbar[2];
  constexpr uint32_t kStepBytes = HAR * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_mbar_init_cluster();
  }
  uint32_t wf[3][KSW][4];
#pragma unroll
  for (int gt = 0; gt < 3; gt++) {
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
  float bh[3];
#pragma unroll
  for (int gt = 0; gt < 3; gt++) bh[gt] = __ldg(b_hh + gt * HAR + col);
  float hprev = (h0 != nullptr && ok) ? h0[(size_t)bq * HAR + col] : 0.f;
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
  const bf16* gip = gi + row0 * 3 * HAR + col;
  float* cptr = c + row0 * HAR + col;
  bf16* ctptr = cT + row0 * HAR + col;
  uint2* g4ptr = gates4 + row0 * HAR + col;
  bf16 gq_raw[3];
  auto load_gi = [&]() {
    gq_raw[0] = gip[0]; gq_raw[1] = gip[HAR]; gq_raw[2] = gip[2 * HAR];
    gip += 3 * HAR;
  };
  load_gi();

  for (int t = 0; t < S; t++) {
    const int cur = t & 1, nxt = cur ^ 1;
    float gq[3];
#pragma unroll
    for (int gt = 0; gt < 3; gt++) gq[gt] = __bfloat162float(gq_raw[gt]);
    if (t + 1 < S) load_gi();
    if (t > 0) {
      mbar_wait(&hbar[cur], ((t - 1 - (cur ^ 1)) >> 1) & 1);
      if (threadIdx.x == 0 && t + 1 < S) mbar_expect_tx(&hbar[cur], kStepBytes);
    }
    float acc[3][2][4];
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
#pragma unroll
      for (int ch = 0; ch < 2; ch++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[gt][ch][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KSW / 2; q++) {
      uint32_t bq4[4];
      ldsm_x4_t(bq4, s_u32(&hs[cur][kh * (HAR / NKH) + 32 * q + lane][0]));
#pragma unroll
      for (int gt = 0; gt < 3; gt++) {
        mma16816(acc[gt][0], wf[gt][2 * q], bq4[0], bq4[1]);
        mma16816(acc[gt][1], wf[gt][2 * q + 1], bq4[2], bq4[3]);
      }
    }
    // every warp parks all four elements of its partial sums; after the barrier a thread adds the four k-quarters of ITS element
    // in the fixed order 0, 1, 2, 3: a sequence's result does not depend on the slot of the batch tile it sits in
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
#pragma unroll
      for (int e = 0; e < 4; e++) part[ub][kh][gt * 4 + e][lane] = acc[gt][0][e] + acc[gt][1][e];
    quad_sync(ub);
    float a3[3];
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
      a3[gt] = ((part[ub][0][gt * 4 + kh][lane] + part[ub][1][gt * 4 + kh][lane]) + part[ub][2][gt * 4 + kh][lane]) +
               part[ub][3][gt * 4 + kh][lane];
    const float ghn = a3[2] + bh[2];
    const float rg = sigmoid_fast(gq[0] + a3[0] + bh[0]);
    const float ug = sigmoid_fast(gq[1] + a3[1] + bh[1]);
    const float ng = tanh_fast(gq[2] + rg * ghn);
    const float hn = fmaf(ug, hprev - ng, ng);  // (1-u) n + u h
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
      *cptr = hn;
      *ctptr = __float2bfloat16_rn(hn);
      *g4ptr = make_uint2(pack_bf16(rg, ug), pack_bf16(ng, ghn));
    }
    cptr += HAR; ctptr += HAR; g4ptr += HAR;
  }
  if (hT != nullptr && ok) hT[(size_t)bq * HAR + col] = hprev;
  cluster.sync();
}

template <int HAR>
__global__ void __launch_bounds__(256, 1)
gru_rec_bwd_wide_kernel(const float* __restrict__ dc, const float* __restrict__ c, const float* __restrict__ h0,
                        const uint2* __restrict__ gates4, const float* __restrict__ w_hh, bf16* __restrict__ dgi,
                        bf16* __restrict__ dgh, float* __restrict__ dh0, float* __restrict__ db_ih, float* __restrict__ db_hh,
                        int B, int S) {
  pdl_wait();
  pdl_trigger();
  constexpr int G = 3 * HAR, KSW = G / 16 / NKH;
  static_assert(KSW % 2 == 0 && (G / NKH) % 32 == 0, "k range of a warp is a whole number of 32-row blocks");
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ub = warp & 1, kh = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;

  // dgh_t as the B operand, [row][sequence]; row(gate index k) = (u / 32) * 96 + gate * 32 + u % 32: the 3 x 32 rows a source
  // CTA produces are contiguous (ONE bulk copy).  2 x 3*HAR x 16 B = 48 KB at HAR = 512: dynamic shared memory
  extern __shared__ __align__(128) unsigned char dyn[];
  bf16 (*ds)[G][BT] = reinterpret_cast<bf16 (*)[G][BT]>(dyn);
  __shared__ float part[2][NKH][4][32];
  __shared__ __align__(128) bf16 dstage[2][3][HCW][BT];
  __shared__ __align__(8) uint64_t dbar[2];
  constexpr uint32_t kStepBytes = G * BT * 2;
  if (threadIdx.x == 0) {
    mbar_init(&dbar[0], 1);
    mbar_init(&dbar[1], 1);
    fence_mbar_init_cluster();
  }
  uint32_t wf[KSW][4];
  {
    const int c0 = HCW * rank + 16 * ub + g;
#pragma unroll
    for (int ks = 0; ks < KSW; ks++) {
      const int k = (kh * KSW + ks) * 16 + 2 * t4;
      wf[ks][0] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0), __ldg(w_hh + (size_t)(k + 1) * HAR + c0));
      wf[ks][1] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0 + 8), __ldg(w_hh + (size_t)(k + 1) * HAR + c0 + 8));
      wf[ks][2] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0), __ldg(w_hh + (size_t)(k + 9) * HAR + c0));
      wf[ks][3] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0 + 8), __ldg(w_hh + (size_t)(k + 9) * HAR + c0 + 8));
    }
  }
  const int urow = 16 * ub + 8 * (kh >> 1) + g;
  const int col = HCW * rank + urow;
  const int sq = 2 * t4 + (kh & 1);
  const int bq = b0 + sq;
  const bool ok = bq < B;
  float carry = 0.f, direct = 0.f;
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&dbar[0], kStepBytes);
    mbar_expect_tx(&dbar[1], kStepBytes);
  }
  cluster.sync();
  const uint32_t ds_local = s_u32(&ds[0][0][0]), bar_local = s_u32(&dbar[0]);
  const uint32_t pub_dst = mapa_u32(ds_local + (uint32_t)(3 * HCW * rank * BT * 2), lane < CS ? lane : 0);
  const uint32_t pub_bar = mapa_u32(bar_local, lane < CS ? lane : 0);

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
    for (int q = 0; q < KSW / 2; q++) {
      uint32_t bq4[4];
      const int kidx = kh * (G / NKH) + 32 * q, gate = kidx / HAR, u = kidx - gate * HAR;
      ldsm_x4_t(bq4, s_u32(&ds[buf][(u / HCW) * 3 * HCW + gate * HCW + lane][0]));
      mma16816(acc[q % 3], wf[2 * q], bq4[0], bq4[1]);
      mma16816(acc[(q + 1) % 3], wf[2 * q + 1], bq4[2], bq4[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) part[ub][kh][e][lane] = (acc[0][e] + acc[1][e]) + acc[2][e];
    quad_sync(ub);
    const float s = direct + (((part[ub][0][kh][lane] + part[ub][1][kh][lane]) + part[ub][2][kh][lane]) + part[ub][3][kh][lane]);
    carry = s;
  };

  const size_t last = (size_t)(ok ? bq : b0) * S + (S - 1);
  const float* dcp = dc + last * HAR + col;
  const float* hpp = c + last * HAR + col - HAR;
  const uint2* g4p = gates4 + last * HAR + col;
  bf16* dgip = dgi + last * G + col;
  bf16* dghp = dgh + last * G + col;
  float dcv_n, hp_n;
  uint2 g4_n;
  auto load_ops = [&](int tt) {
    dcv_n = *dcp;
    g4_n = *g4p;
    if (tt > 0) hp_n = *hpp;
    else hp_n = (h0 != nullptr && ok) ? h0[(size_t)bq * HAR + col] : 0.f;
    dcp -= HAR; g4p -= HAR; hpp -= HAR;
  };
  load_ops(S - 1);

  for (int it = 0; it < S; it++) {
    const int t = S - 1 - it, buf = t & 1;
    const float dcv = dcv_n, hp = hp_n;
    const uint2 g4 = g4_n;
    if (t > 0) load_ops(t - 1);
    if (it > 0) consume(it - 1);
    float dr = 0.f, du = 0.f, dnr = 0.f, dnv = 0.f;
    direct = 0.f;
    if (ok) {
      const float dh = carry + dcv;
      const float rg = __uint_as_float(g4.x << 16), ug = __uint_as_float(g4.x & 0xffff0000u);
      const float ng = __uint_as_float(g4.y << 16), hnv = __uint_as_float(g4.y & 0xffff0000u);
      const float dn = dh * (1.f - ug) * (1.f - ng * ng);
      du = dh * (hp - ng) * ug * (1.f - ug);
      dr = dn * hnv * rg * (1.f - rg);
      dnr = dn * rg;
      dnv = dn;
      direct = dh * ug;
      sb[0] += dr; sb[1] += du; sb[2] += dn; sb[3] += dnr;
    }
    dstage[buf][0][urow][sq] = __float2bfloat16_rn(dr);
    dstage[buf][1][urow][sq] = __float2bfloat16_rn(du);
    dstage[buf][2][urow][sq] = __float2bfloat16_rn(dnr);
    fence_async_smem();
    if (warp == 0) {
      publish_sync();
      if (lane < CS) bulk_s2s(pub_dst + buf * (G * BT * 2), s_u32(&dstage[buf][0][0][0]), 3 * HCW * BT * 2, pub_bar + buf * 8);
    } else {
      publish_arrive();
    }
    if (ok) {
      const ptrdiff_t back = -(ptrdiff_t)it * G;
      bf16* pi = dgip + back;
      bf16* ph = dghp + back;
      pi[0] = __float2bfloat16_rn(dr); pi[HAR] = __float2bfloat16_rn(du); pi[2 * HAR] = __float2bfloat16_rn(dnv);
      ph[0] = __float2bfloat16_rn(dr); ph[HAR] = __float2bfloat16_rn(du); ph[2 * HAR] = __float2bfloat16_rn(dnr);
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
      atomicAdd(db_ih + col, sb[0]); atomicAdd(db_ih + HAR + col, sb[1]); atomicAdd(db_ih + 2 * HAR + col, sb[2]);
      atomicAdd(db_hh + col, sb[0]); atomicAdd(db_hh + HAR + col, sb[1]); atomicAdd(db_hh + 2 * HAR + col, sb[3]);
    }
  }
  if (dh0 != nullptr && ok) dh0[(size_t)bq * HAR + col] = carry;
  cluster.sync();
}

}  // namespace

// Har = 512 only exists here; at Har = 256 the 32-unit split halves the products a warp issues per step and measures faster
// than gru_mma.cu (B = 64, S = 128: forward 0.115 -> 0.106 ms, backward 0.174 -> 0.121 ms).  CPC_B200_GRU_WIDE=0 / 1: never / always.
bool gru_wide_supported(int Har) { return Har == 512 || Har == 256 || Har == 128; }
bool gru_wide_preferred(int Har) {
  static const int mode = []() { const char* e = getenv("CPC_B200_GRU_WIDE"); return e ? atoi(e) : -1; }();
  if (mode == 0) return false;
  if (mode == 1) return gru_wide_supported(Har);
  return Har == 512 || Har == 256;
}

int gru_rec_fwd_wide(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, float* c, bf16* cT, bf16* sR,
                     float* hT, int B, int S, int Har, cudaStream_t st) {
  uint2* gates4 = reinterpret_cast<uint2*>(sR);
  void* args[] = {&gi, &w_hh, &b_hh, &h0, &c, &cT, &gates4, &hT, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (Har == 512) return launch_cluster("gru_rec_fwd_wide", gru_rec_fwd_wide_kernel<512>, 16, ncl, st, args);
  if (Har == 256) return launch_cluster("gru_rec_fwd_wide", gru_rec_fwd_wide_kernel<256>, 8, ncl, st, args);
  return launch_cluster("gru_rec_fwd_wide", gru_rec_fwd_wide_kernel<128>, 4, ncl, st, args);
}
int gru_rec_bwd_wide(const float* dc, const float* c, const float* h0, const bf16* sR, const float* w_hh, bf16* dgi, bf16* dgh,
                     float* dh0, float* db_ih, float* db_hh, int B, int S, int Har, cudaStream_t st) {
  const uint2* gates4 = reinterpret_cast<const uint2*>(sR);
  void* args[] = {&dc, &c, &h0, &gates4, &w_hh, &dgi, &dgh, &dh0, &db_ih, &db_hh, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  const size_t dyn = (size_t)2 * 3 * Har * BT * 2;
  if (Har == 512) return launch_cluster("gru_rec_bwd_wide", gru_rec_bwd_wide_kernel<512>, 16, ncl, st, args, dyn);
  if (Har == 256) return launch_cluster("gru_rec_bwd_wide", gru_rec_bwd_wide_kernel<256>, 8, ncl, st, args, dyn);
  return launch_cluster("gru_rec_bwd_wide", gru_rec_bwd_wide_kernel<128>, 4, ncl, st, args, dyn);
}

}  // namespace cpcb200
