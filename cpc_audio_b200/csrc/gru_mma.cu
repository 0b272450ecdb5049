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

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
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
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------
// forward.  block = 128 threads (4 warps), cluster = HAR/64 CTAs, grid = cluster * ceil(B/8).
// ---------------------------------------------------------------------------------------------------------
template <int HAR>
__global__ void __launch_bounds__(128, 1)
gru_rec_fwd_mma_kernel(const bf16* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                       const float* __restrict__ h0, float* __restrict__ c, bf16* __restrict__ cT, bf16* __restrict__ sR,
                       bf16* __restrict__ sU, bf16* __restrict__ sN, bf16* __restrict__ sHN, float* __restrict__ hT, int B,
                       int S) {
  constexpr int KS = HAR / 16;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;

  __shared__ __align__(128) bf16 hs[2][HAR][BT];  // h_{t-1} as the B operand: [k][sequence]

  // resident A fragments: gate gt, rows 16*warp + {g, g+8} of this CTA's slice
  uint32_t wf[3][KS][4];
#pragma unroll
  for (int gt = 0; gt < 3; gt++) {
    const float* r0 = w_hh + (size_t)(gt * HAR + HC * rank + 16 * warp + g) * HAR;
    const float* r1 = r0 + 8 * HAR;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int k = ks * 16 + 2 * t4;
      wf[gt][ks][0] = pack_bf16(__ldg(r0 + k), __ldg(r0 + k + 1));
      wf[gt][ks][1] = pack_bf16(__ldg(r1 + k), __ldg(r1 + k + 1));
      wf[gt][ks][2] = pack_bf16(__ldg(r0 + k + 8), __ldg(r0 + k + 9));
      wf[gt][ks][3] = pack_bf16(__ldg(r1 + k + 8), __ldg(r1 + k + 9));
    }
  }
  // element e of a thread: unit jj = g + 8*(e>>1), sequence bb = 2*t4 + (e&1)
  float bh[3][2];
#pragma unroll
  for (int gt = 0; gt < 3; gt++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) bh[gt][hf] = __ldg(b_hh + gt * HAR + HC * rank + 16 * warp + g + 8 * hf);
  float hprev[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const int col = HC * rank + 16 * warp + g + 8 * (e >> 1), bq = b0 + 2 * t4 + (e & 1);
    hprev[e] = (h0 != nullptr && bq < B) ? h0[(size_t)bq * HAR + col] : 0.f;
  }
  for (int i = threadIdx.x; i < HAR * BT; i += blockDim.x) {
    const int k = i / BT, bq = b0 + (i - k * BT);
    hs[0][k][i - k * BT] = __float2bfloat16_rn((h0 != nullptr && bq < B) ? h0[(size_t)bq * HAR + k] : 0.f);
  }
  cluster.sync();

  for (int t = 0; t < S; t++) {
    const int cur = t & 1, nxt = cur ^ 1;
    // prefetch the input projections of this step (consumed after the matrix product)
    float gq[3][4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int col = HC * rank + 16 * warp + g + 8 * (e >> 1), bq = b0 + 2 * t4 + (e & 1);
      if (bq < B) {
        const bf16* gp = gi + ((size_t)bq * S + t) * 3 * HAR + col;
        gq[0][e] = __bfloat162float(gp[0]); gq[1][e] = __bfloat162float(gp[HAR]); gq[2][e] = __bfloat162float(gp[2 * HAR]);
      } else { gq[0][e] = gq[1][e] = gq[2][e] = 0.f; }
    }
    float acc[3][4];
#pragma unroll
    for (int gt = 0; gt < 3; gt++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[gt][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KS / 2; q++) {
      uint32_t bq4[4];
      ldsm_x4_t(bq4, s_u32(&hs[cur][32 * q + lane][0]));
#pragma unroll
      for (int gt = 0; gt < 3; gt++) {
        mma16816(acc[gt], wf[gt][2 * q], bq4[0], bq4[1]);
        mma16816(acc[gt], wf[gt][2 * q + 1], bq4[2], bq4[3]);
      }
    }
    float hn[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int hf = e >> 1;
      const float ghn = acc[2][e] + bh[2][hf];
      const float rg = sigmoidf_(gq[0][e] + acc[0][e] + bh[0][hf]);
      const float ug = sigmoidf_(gq[1][e] + acc[1][e] + bh[1][hf]);
      const float ng = tanhf(gq[2][e] + rg * ghn);
      hn[e] = (1.f - ug) * ng + ug * hprev[e];
      hprev[e] = hn[e];
      const int col = HC * rank + 16 * warp + g + 8 * hf, bq = b0 + 2 * t4 + (e & 1);
      if (bq < B) {
        const size_t o = ((size_t)bq * S + t) * HAR + col;
        c[o] = hn[e];
        cT[o] = __float2bfloat16_rn(hn[e]);
        sR[o] = __float2bfloat16_rn(rg); sU[o] = __float2bfloat16_rn(ug); sN[o] = __float2bfloat16_rn(ng);
        sHN[o] = __float2bfloat16_rn(ghn);
        if (hT != nullptr && t == S - 1) hT[(size_t)bq * HAR + col] = hn[e];
      }
    }
    // publish the 64 new units to every CTA of the cluster: rows (unit), 2 sequences per 32-bit store
    {
      const uint32_t v0 = pack_bf16(hn[0], hn[1]), v1 = pack_bf16(hn[2], hn[3]);
      const int row0 = HC * rank + 16 * warp + g;
      for (int pr = 0; pr < CS; pr++) {
        bf16* base = cluster.map_shared_rank(&hs[nxt][0][0], pr);
        *reinterpret_cast<uint32_t*>(base + (size_t)row0 * BT + 2 * t4) = v0;
        *reinterpret_cast<uint32_t*>(base + (size_t)(row0 + 8) * BT + 2 * t4) = v1;
      }
    }
    cluster.sync();
  }
}

// ---------------------------------------------------------------------------------------------------------
// BPTT.  Same launch shape.  Resident: A[i][gate index] = W_hh[gate index][64*rank + 16*warp + i].
// ---------------------------------------------------------------------------------------------------------
template <int HAR>
__global__ void __launch_bounds__(128, 1)
gru_rec_bwd_mma_kernel(const float* __restrict__ dc, const float* __restrict__ c, const float* __restrict__ h0,
                       const bf16* __restrict__ sR, const bf16* __restrict__ sU, const bf16* __restrict__ sN,
                       const bf16* __restrict__ sHN, const float* __restrict__ w_hh, bf16* __restrict__ dgi,
                       bf16* __restrict__ dgh, float* __restrict__ dh0, int B, int S) {
  constexpr int G = 3 * HAR, KS = G / 16;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;

  __shared__ __align__(128) bf16 ds[2][G][BT];  // dgh_t as the B operand: [gate index][sequence]

  uint32_t wf[KS][4];
  {
    const int c0 = HC * rank + 16 * warp + g;
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int k = ks * 16 + 2 * t4;
      wf[ks][0] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0), __ldg(w_hh + (size_t)(k + 1) * HAR + c0));
      wf[ks][1] = pack_bf16(__ldg(w_hh + (size_t)k * HAR + c0 + 8), __ldg(w_hh + (size_t)(k + 1) * HAR + c0 + 8));
      wf[ks][2] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0), __ldg(w_hh + (size_t)(k + 9) * HAR + c0));
      wf[ks][3] = pack_bf16(__ldg(w_hh + (size_t)(k + 8) * HAR + c0 + 8), __ldg(w_hh + (size_t)(k + 9) * HAR + c0 + 8));
    }
  }
  float carry[4] = {0.f, 0.f, 0.f, 0.f};
  cluster.sync();

  for (int t = S - 1; t >= 0; t--) {
    const int buf = t & 1;
    float direct[4], dr[4], du[4], dnr[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int col = HC * rank + 16 * warp + g + 8 * (e >> 1), bq = b0 + 2 * t4 + (e & 1);
      dr[e] = du[e] = dnr[e] = direct[e] = 0.f;
      if (bq < B) {
        const size_t o = ((size_t)bq * S + t) * HAR + col;
        const float dh = carry[e] + dc[o];
        const float rg = __bfloat162float(sR[o]), ug = __bfloat162float(sU[o]), ng = __bfloat162float(sN[o]);
        const float hnv = __bfloat162float(sHN[o]);
        const float hp = t > 0 ? c[o - HAR] : (h0 != nullptr ? h0[(size_t)bq * HAR + col] : 0.f);
        const float dn = dh * (1.f - ug) * (1.f - ng * ng);
        du[e] = dh * (hp - ng) * ug * (1.f - ug);
        dr[e] = dn * hnv * rg * (1.f - rg);
        dnr[e] = dn * rg;
        direct[e] = dh * ug;
        const size_t og = ((size_t)bq * S + t) * G + col;
        dgi[og] = __float2bfloat16_rn(dr[e]); dgi[og + HAR] = __float2bfloat16_rn(du[e]); dgi[og + 2 * HAR] = __float2bfloat16_rn(dn);
        dgh[og] = __float2bfloat16_rn(dr[e]); dgh[og + HAR] = __float2bfloat16_rn(du[e]); dgh[og + 2 * HAR] = __float2bfloat16_rn(dnr[e]);
      }
    }
    {
      const int row0 = HC * rank + 16 * warp + g;
      for (int pr = 0; pr < CS; pr++) {
        bf16* base = cluster.map_shared_rank(&ds[buf][0][0], pr);
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
          const int r = row0 + 8 * hf;
          *reinterpret_cast<uint32_t*>(base + (size_t)r * BT + 2 * t4) = pack_bf16(dr[2 * hf], dr[2 * hf + 1]);
          *reinterpret_cast<uint32_t*>(base + (size_t)(HAR + r) * BT + 2 * t4) = pack_bf16(du[2 * hf], du[2 * hf + 1]);
          *reinterpret_cast<uint32_t*>(base + (size_t)(2 * HAR + r) * BT + 2 * t4) = pack_bf16(dnr[2 * hf], dnr[2 * hf + 1]);
        }
      }
    }
    cluster.sync();
    float acc[3][4];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[a][e] = 0.f;
#pragma unroll
    for (int q = 0; q < KS / 2; q++) {
      uint32_t bq4[4];
      ldsm_x4_t(bq4, s_u32(&ds[buf][32 * q + lane][0]));
      mma16816(acc[q % 3], wf[2 * q], bq4[0], bq4[1]);
      mma16816(acc[q % 3], wf[2 * q + 1], bq4[2], bq4[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; e++) carry[e] = direct[e] + (acc[0][e] + acc[1][e]) + acc[2][e];
  }
  if (dh0 != nullptr) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int col = HC * rank + 16 * warp + g + 8 * (e >> 1), bq = b0 + 2 * t4 + (e & 1);
      if (bq < B) dh0[(size_t)bq * HAR + col] = carry[e];
    }
  }
}

template <class K>
int launch_cluster(const char* name, K kernel, int cs, int nclusters, cudaStream_t st, void** args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * nclusters);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CPC_CHECK_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), args));
  CPC_LAUNCHED_N(name, st);
  return 0;
}

}  // namespace

bool gru_mma_supported(int Har) { return Har == 64 || Har == 128 || Har == 256; }

int gru_rec_fwd_mma(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, float* c, bf16* cT, bf16* sR, bf16* sU,
                    bf16* sN, bf16* sHN, float* hT, int B, int S, int Har, cudaStream_t st) {
  void* args[] = {&gi, &w_hh, &b_hh, &h0, &c, &cT, &sR, &sU, &sN, &sHN, &hT, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (Har == 256) return launch_cluster("gru_rec_fwd_mma", gru_rec_fwd_mma_kernel<256>, 4, ncl, st, args);
  if (Har == 128) return launch_cluster("gru_rec_fwd_mma", gru_rec_fwd_mma_kernel<128>, 2, ncl, st, args);
  return launch_cluster("gru_rec_fwd_mma", gru_rec_fwd_mma_kernel<64>, 1, ncl, st, args);
}
int gru_rec_bwd_mma(const float* dc, const float* c, const float* h0, const bf16* sR, const bf16* sU, const bf16* sN,
                    const bf16* sHN, const float* w_hh, bf16* dgi, bf16* dgh, float* dh0, int B, int S, int Har, cudaStream_t st) {
  void* args[] = {&dc, &c, &h0, &sR, &sU, &sN, &sHN, &w_hh, &dgi, &dgh, &dh0, &B, &S};
  const int ncl = (B + BT - 1) / BT;
  if (Har == 256) return launch_cluster("gru_rec_bwd_mma", gru_rec_bwd_mma_kernel<256>, 4, ncl, st, args);
  if (Har == 128) return launch_cluster("gru_rec_bwd_mma", gru_rec_bwd_mma_kernel<128>, 2, ncl, st, args);
  return launch_cluster("gru_rec_bwd_mma", gru_rec_bwd_mma_kernel<64>, 1, ncl, st, args);
}

}  // namespace cpcb200
