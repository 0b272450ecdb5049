// gru.cu - CPCAR GRU branch forward + BPTT (reference: cpc/model.py:175-176,185-204 -> torch.nn.GRU, gates r,z,n).
//
// Structure:  the input projection W_ih x_t (all t) and every weight gradient are hoisted out of the time loop
// as GEMMs; the recurrence itself is ONE persistent kernel per direction: a thread-block cluster of Har/64
// CTAs owns a tile of BT sequences, each CTA keeps its 192 x Har slice of W_hh resident in shared memory for
// the whole sequence, and the CTAs exchange the 64 hidden units they produce through distributed shared
// memory (one cluster barrier per time step).  BPTT mirrors it with the transposed slice (64 x 3Har).
// That is the fp32 / CUDA-core form kept in this file (and the LSTM twin further down); on the bf16 path the recurrences run on
// tensor cores with the W_hh slice in registers: gru_mma.cu, gru_mma_wide.cu, lstm_mma.cu (dispatched from here).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cpcb200 {

int gemm_nt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm, const float* bias,
            const OutView& C, cudaStream_t st);
int gemm_tn(bool bf16_in, int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc, int mode,
            int Ci, int taps, cudaStream_t st);

bool gru_mma_supported(int Har);
int gru_rec_fwd_mma(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, float* c, bf16* cT, bf16* sR, bf16* sU,
                    bf16* sN, bf16* sHN, float* hT, int B, int S, int Har, cudaStream_t st);
bool gru_wide_supported(int Har);
bool gru_wide_preferred(int Har);
int gru_rec_fwd_wide(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, float* c, bf16* cT, bf16* sR,
                     float* hT, int B, int S, int Har, cudaStream_t st);
int gru_rec_bwd_wide(const float* dc, const float* c, const float* h0, const bf16* sR, const float* w_hh, bf16* dgi, bf16* dgh,
                     float* dh0, float* db_ih, float* db_hh, int B, int S, int Har, cudaStream_t st);
bool lstm_mma_supported(int Har);
int lstm_rec_fwd_mma(const bf16* gi, const float* w_hh, const float* b_hh, const float* h0, const float* c0, float* out, bf16* outT,
                     void* gates4, float* cell, float* hT, float* cT, int B, int S, int Har, cudaStream_t st);
int lstm_rec_bwd_mma(const float* dout, const float* c0, const void* gates4, const float* cell, const float* w_hh, bf16* dg,
                     float* db_ih, float* db_hh, int B, int S, int Har, cudaStream_t st);
int gru_rec_bwd_mma(const float* dc, const float* c, const float* h0, const bf16* sR, const bf16* sU, const bf16* sN,
                    const bf16* sHN, const float* w_hh, bf16* dgi, bf16* dgh, float* dh0, float* db_ih, float* db_hh, int B, int S,
                    int Har, cudaStream_t st);

namespace {

constexpr int HC = 64;  // hidden units owned by one CTA of the cluster

template <class WT> struct WVec;
template <> struct WVec<float> { static constexpr int V = 4; };
template <> struct WVec<bf16> { static constexpr int V = 8; };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// dot products of one resident weight row against BT fp32 vectors in shared memory (k split over KS lanes)
template <class WT, int KS, int BT>
__device__ __forceinline__ void row_dot(const WT* __restrict__ wrow, const float* __restrict__ vec, int vstride, int Kd,
                                        int q, float (&acc)[BT]) {
  constexpr int V = WVec<WT>::V;
#pragma unroll
  for (int b = 0; b < BT; b++) acc[b] = 0.f;
  const int iters = Kd / (V * KS);
  for (int i = 0; i < iters; i++) {
    const int k = (i * KS + q) * V;
    float w[V];
    load_vec<V>(wrow + k, w);
#pragma unroll
    for (int b = 0; b < BT; b++) {
      float h[V];
      load_vec<V>(vec + b * vstride + k, h);
#pragma unroll
      for (int j = 0; j < V; j++) acc[b] = fmaf(w[j], h[j], acc[b]);
    }
  }
#pragma unroll
  for (int o = 1; o < KS; o <<= 1)
#pragma unroll
    for (int b = 0; b < BT; b++) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
}

// ---------------------------------------------------------------------------------------------------------
// forward recurrence.  grid = (cluster CS = Har/64) x ceil(B/BT) clusters; block = 3*HC*KS threads.
// ---------------------------------------------------------------------------------------------------------
template <class WT, class T, int BT>
__global__ void __launch_bounds__(384, 1)
gru_rec_fwd_kernel(const T* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                   const float* __restrict__ h0, float* __restrict__ c, T* __restrict__ cT, T* __restrict__ sR,
                   T* __restrict__ sU, T* __restrict__ sN, T* __restrict__ sHN, float* __restrict__ hT, int B, int S,
                   int Har) {
  constexpr int KS = 2, ROWS = 3 * HC;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int tid = threadIdx.x;
  const int wstride = Har + WVec<WT>::V * KS;

  extern __shared__ __align__(16) unsigned char smraw[];
  WT* Wsm = reinterpret_cast<WT*>(smraw);                                           // [ROWS][wstride]
  float* hsm = reinterpret_cast<float*>(smraw + align_up((size_t)ROWS * wstride * sizeof(WT), 16));  // [2][BT][Har]
  float* gsm = hsm + 2 * BT * Har;                                                  // [ROWS][BT]
  float* bsm = gsm + ROWS * BT;                                                     // [ROWS]

  // resident slice of W_hh: local row r = gate*64 + j  <->  global row gate*Har + 64*rank + j
  for (int i = tid; i < ROWS * (Har / 4); i += blockDim.x) {
    const int r = i / (Har / 4), k4 = (i - r * (Har / 4)) * 4;
    const int grow = (r / HC) * Har + HC * rank + (r % HC);
    float4 v = *reinterpret_cast<const float4*>(w_hh + (size_t)grow * Har + k4);
    WT* dst = Wsm + (size_t)r * wstride + k4;
    dst[0] = from_f<WT>(v.x); dst[1] = from_f<WT>(v.y); dst[2] = from_f<WT>(v.z); dst[3] = from_f<WT>(v.w);
  }
  for (int i = tid; i < ROWS; i += blockDim.x) bsm[i] = b_hh[(i / HC) * Har + HC * rank + (i % HC)];
  for (int i = tid; i < BT * Har; i += blockDim.x) {
    const int b = i / Har, k = i - b * Har;
    hsm[i] = (h0 != nullptr && b0 + b < B) ? h0[(size_t)(b0 + b) * Har + k] : 0.f;
  }
  cluster.sync();

  const int r = tid / KS, q = tid % KS;
  const WT* wrow = Wsm + (size_t)r * wstride;
  const bool gate_thread = tid < HC * BT;
  const int gj = tid % HC, gb = tid / HC;             // gate thread -> (hidden unit, sequence)
  const int col = HC * rank + gj;
  const bool gvalid = gate_thread && (b0 + gb < B);

  for (int t = 0; t < S; t++) {
    const float* hcur = hsm + (t & 1) * BT * Har;
    float* hnxt = hsm + ((t + 1) & 1) * BT * Har;
    float gir = 0.f, giz = 0.f, gin = 0.f;
    if (gvalid) {
      const T* g = gi + ((size_t)(b0 + gb) * S + t) * 3 * Har + col;
      gir = to_f(g[0]); giz = to_f(g[Har]); gin = to_f(g[2 * Har]);
    }
    float acc[BT];
    row_dot<WT, KS, BT>(wrow, hcur, Har, Har, q, acc);
    if (q == 0) {
#pragma unroll
      for (int b = 0; b < BT; b++) gsm[r * BT + b] = acc[b];
    }
    __syncthreads();
    if (gate_thread) {
      const float ghr = gsm[gj * BT + gb] + bsm[gj];
      const float ghz = gsm[(HC + gj) * BT + gb] + bsm[HC + gj];
      const float ghn = gsm[(2 * HC + gj) * BT + gb] + bsm[2 * HC + gj];
      const float rg = sigmoidf_(gir + ghr);
      const float ug = sigmoidf_(giz + ghz);
      const float ng = tanhf(gin + rg * ghn);
      const float hp = hcur[gb * Har + col];
      const float hn = (1.f - ug) * ng + ug * hp;
      for (int pr = 0; pr < CS; pr++) {
        float* dst = cluster.map_shared_rank(hnxt, pr);
        dst[gb * Har + col] = hn;
      }
      if (gvalid) {
        const size_t o = ((size_t)(b0 + gb) * S + t) * Har + col;
        c[o] = hn;
        if (cT) cT[o] = from_f<T>(hn);
        sR[o] = from_f<T>(rg); sU[o] = from_f<T>(ug); sN[o] = from_f<T>(ng); sHN[o] = from_f<T>(ghn);
        if (hT != nullptr && t == S - 1) hT[(size_t)(b0 + gb) * Har + col] = hn;
      }
    }
    cluster.sync();
  }
}

// ---------------------------------------------------------------------------------------------------------
// BPTT.  block = HC*KS = 512 threads.  Wt[j][g] = W_hh[g][64*rank + j] resident; per step:
//   gate grads (local 64 units) -> all-gather dgh over DSMEM -> dh_prev = dh*u + dgh . W_hh[:, slice]
// ---------------------------------------------------------------------------------------------------------
template <class WT, class T, int BT>
__global__ void __launch_bounds__(512, 1)
gru_rec_bwd_kernel(const float* __restrict__ dc, const float* __restrict__ c, const float* __restrict__ h0,
                   const T* __restrict__ sR, const T* __restrict__ sU, const T* __restrict__ sN,
                   const T* __restrict__ sHN, const float* __restrict__ w_hh, T* __restrict__ dgi, T* __restrict__ dgh,
                   float* __restrict__ dh0, int B, int S, int Har) {
  constexpr int KS = 8;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int tid = threadIdx.x;
  const int G = 3 * Har;
  const int wstride = G + WVec<WT>::V * KS;

  extern __shared__ __align__(16) unsigned char smraw[];
  WT* Wsm = reinterpret_cast<WT*>(smraw);                                             // [HC][wstride]
  float* dsm = reinterpret_cast<float*>(smraw + align_up((size_t)HC * wstride * sizeof(WT), 16));  // [2][BT][G]
  float* osm = dsm + 2 * BT * G;                                                      // [HC][BT]

  for (int i = tid; i < HC * G; i += blockDim.x) {
    const int g = i / HC, j = i - g * HC;  // consecutive threads read consecutive columns of row g
    Wsm[(size_t)j * wstride + g] = from_f<WT>(w_hh[(size_t)g * Har + HC * rank + j]);
  }
  cluster.sync();

  const int r = tid / KS, q = tid % KS;
  const WT* wrow = Wsm + (size_t)r * wstride;
  const bool gate_thread = tid < HC * BT;
  const int gj = tid % HC, gb = tid / HC;
  const int col = HC * rank + gj;
  const bool gvalid = gate_thread && (b0 + gb < B);
  float dh_carry = 0.f;

  for (int t = S - 1; t >= 0; t--) {
    float* dbuf = dsm + (t & 1) * BT * G;
    float dh_direct = 0.f;
    if (gate_thread) {
      float dr = 0.f, du = 0.f, dn = 0.f, dnr = 0.f;
      if (gvalid) {
        const size_t o = ((size_t)(b0 + gb) * S + t) * Har + col;
        const float dh = dh_carry + dc[o];
        const float rg = to_f(sR[o]), ug = to_f(sU[o]), ng = to_f(sN[o]), hn = to_f(sHN[o]);
        const float hp = t > 0 ? c[o - Har] : (h0 ? h0[(size_t)(b0 + gb) * Har + col] : 0.f);
        dn = dh * (1.f - ug) * (1.f - ng * ng);
        du = dh * (hp - ng) * ug * (1.f - ug);
        dr = dn * hn * rg * (1.f - rg);
        dnr = dn * rg;
        dh_direct = dh * ug;
        const size_t og = ((size_t)(b0 + gb) * S + t) * G + col;
        dgi[og] = from_f<T>(dr); dgi[og + Har] = from_f<T>(du); dgi[og + 2 * Har] = from_f<T>(dn);
        dgh[og] = from_f<T>(dr); dgh[og + Har] = from_f<T>(du); dgh[og + 2 * Har] = from_f<T>(dnr);
      }
      for (int pr = 0; pr < CS; pr++) {
        float* dst = cluster.map_shared_rank(dbuf, pr) + gb * G + col;
        dst[0] = dr; dst[Har] = du; dst[2 * Har] = dnr;
      }
    }
    cluster.sync();
    float acc[BT];
    row_dot<WT, KS, BT>(wrow, dbuf, G, G, q, acc);
    if (q == 0) {
#pragma unroll
      for (int b = 0; b < BT; b++) osm[r * BT + b] = acc[b];
    }
    __syncthreads();
    if (gate_thread) dh_carry = dh_direct + osm[gj * BT + gb];
    __syncthreads();
  }
  if (gvalid && dh0 != nullptr) dh0[(size_t)(b0 + gb) * Har + col] = dh_carry;
}

// ---------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------
template <class T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  pdl_wait();
  pdl_trigger();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}
__global__ void cast3_bf16_kernel(const float* __restrict__ s0, bf16* __restrict__ d0, long long n0, const float* __restrict__ s1,
                                  bf16* __restrict__ d1, long long n1, const float* __restrict__ s2, bf16* __restrict__ d2,
                                  long long n2) {
  pdl_wait();
  pdl_trigger();
  const long long n = n0 + n1 + n2;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    const float* s; bf16* d; long long j;
    if (i < n0) { s = s0; d = d0; j = i; } else if (i < n0 + n1) { s = s1; d = d1; j = i - n0; } else { s = s2; d = d2; j = i - n0 - n1; }
    const float4 v = *reinterpret_cast<const float4*>(s + j);  // every tensor length is a multiple of 4
    uint2 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(d + j) = o;
  }
}
// dst[c][r] = src[r][c]   (blockIdx.z = matrix index of a batch of equally shaped, densely stacked matrices)
template <class T>
__global__ void transpose_cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int R, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  src += (size_t)blockIdx.z * R * C;
  dst += (size_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int rr = r0 + i, cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (rr < R && cc < C) ? src[(size_t)rr * C + cc] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int cc = c0 + i, rr = r0 + threadIdx.x;
    if (cc < C && rr < R) dst[(size_t)cc * R + rr] = from_f<T>(tile[threadIdx.x][i]);
  }
}
// out[n] += sum_m src[m][n]
template <class T>
__global__ void colsum_kernel(const T* __restrict__ src, float* __restrict__ out, long long M, int N, int rows_per_block) {
  const long long m0 = (long long)blockIdx.y * rows_per_block;
  const long long m1 = min(M, m0 + rows_per_block);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  src += (size_t)blockIdx.z * M * N;  // stacked matrices: one (M, N) block and one output vector per blockIdx.z
  out += (size_t)blockIdx.z * N;
  float s = 0.f;
  for (long long m = m0; m < m1; m++) s += to_f(src[m * N + n]);
  atomicAdd(out + n, s);
}

}  // namespace

int launch_cast3_bf16(const float* s0, bf16* d0, long long n0, const float* s1, bf16* d1, long long n1, const float* s2, bf16* d2,
                      long long n2, cudaStream_t st) {
  if ((n0 | n1 | n2) & 3) return fail(CPCB200_ERR_BAD_DIMS, "cast3: lengths must be multiples of 4");
  const long long n = n0 + n1 + n2;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPC_CHECK_CUDA(launch_k(cast3_bf16_kernel, dim3(blocks), dim3(256), 0, st, 1, s0, d0, n0, s1, d1, n1, s2, d2, n2));
  CPC_LAUNCHED_N("cast3", st);
  return 0;
}
template <class T> int launch_cast(const float* src, T* dst, long long n, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPC_CHECK_CUDA(launch_k(cast_kernel<T>, dim3(blocks), dim3(256), 0, st, 1, src, dst, n));
  CPC_LAUNCHED_N("cast", st);
  return 0;
}
template int launch_cast<bf16>(const float*, bf16*, long long, cudaStream_t);
template int launch_cast<float>(const float*, float*, long long, cudaStream_t);

template <class T> int launch_transpose_cast(const float* src, T* dst, int R, int C, cudaStream_t st, int nb) {
  dim3 grid((C + 31) / 32, (R + 31) / 32, nb), block(32, 8);
  CPC_CHECK_CUDA(launch_k(transpose_cast_kernel<T>, grid, block, 0, st, 1, src, dst, R, C));
  CPC_LAUNCHED_N("transpose_cast", st);
  return 0;
}
template int launch_transpose_cast<bf16>(const float*, bf16*, int, int, cudaStream_t, int);
template int launch_transpose_cast<float>(const float*, float*, int, int, cudaStream_t, int);

template <class T> int launch_colsum(const T* src, float* out, long long M, int N, cudaStream_t st, int nb = 1) {
  const int rpb = nb > 1 ? 64 : 256;
  dim3 grid((N + 127) / 128, (unsigned)((M + rpb - 1) / rpb), nb);
  colsum_kernel<T><<<grid, 128, 0, st>>>(src, out, M, N, rpb);
  CPC_LAUNCHED_N("colsum", st);
  return 0;
}
template int launch_colsum<bf16>(const bf16*, float*, long long, int, cudaStream_t, int);
template int launch_colsum<float>(const float*, float*, long long, int, cudaStream_t, int);

namespace {

template <class T> constexpr bool isf_dummy() { return sizeof(T) == 4; }

constexpr int kBT = 2;

template <class WT> size_t rec_fwd_smem(int Har, int bt = kBT) {
  const int wstride = Har + WVec<WT>::V * 2;
  return align_up((size_t)3 * HC * wstride * sizeof(WT), 16) + (size_t)(2 * bt * Har + 3 * HC * bt + 3 * HC) * sizeof(float);
}
template <class WT> size_t rec_bwd_smem(int Har, int bt = kBT) {
  const int wstride = 3 * Har + WVec<WT>::V * 8;
  return align_up((size_t)HC * wstride * sizeof(WT), 16) + (size_t)(2 * bt * 3 * Har + HC * bt) * sizeof(float);
}

template <class K>
int launch_cluster(const char* name, K kernel, int cs, int nclusters, int threads, size_t smem, cudaStream_t st, void** args) {
  CPC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) CPC_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * nclusters);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CPC_CHECK_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), args));
  CPC_LAUNCHED_N(name, st);
  return 0;
}

struct GruLayout {
  size_t gates[CPCB200_MAX_GRU_LAYERS][4];  // byte offsets: R,U,N,HN (T)
  size_t cT[CPCB200_MAX_GRU_LAYERS];        // T copy of the layer output (bf16 path only)
  size_t cf[CPCB200_MAX_GRU_LAYERS];        // fp32 output of non-last layers
  size_t inT;                               // bf16 copy of z (layer-0 input), bf16 path only
  size_t total;
};
GruLayout gru_layout(const Geo& g) {
  GruLayout l{};
  size_t off = 0;
  const size_t es = g.bf16 ? 2 : 4;
  const size_t n = (size_t)g.B * g.S * g.Har;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes); return r; };
  for (int i = 0; i < g.nL; i++) {
    for (int k = 0; k < 4; k++) l.gates[i][k] = take(n * es);
    l.cT[i] = g.bf16 ? take(n * es) : 0;
    l.cf[i] = (i < g.nL - 1) ? take(n * 4) : 0;
  }
  l.inT = g.bf16 ? take((size_t)g.B * g.S * g.H * 2) : 0;
  l.total = off;
  return l;
}

template <class T>
int gru_fwd_t(const Geo& g, const float* z, const float* h0, const cpcb200_gru_params* p, float* c, float* hT, void* save,
              void* wsp, size_t ws_bytes, cudaStream_t st) {
  typedef T WT;
  const int B = g.B, S = g.S, Har = g.Har;
  if (Har % HC != 0 || Har / HC > 8) return fail(CPCB200_ERR_UNSUPPORTED, "gru: Har=%d needs a cluster of %d CTAs (max 8)", Har, Har / HC);
  int bt = kBT;  // sequences per cluster; falls back to 1 when the resident slice leaves no room for 2
  size_t smem = rec_fwd_smem<WT>(Har, bt);
  if (smem > 227 * 1024) { bt = 1; smem = rec_fwd_smem<WT>(Har, bt); }
  if (smem > 227 * 1024) return fail(CPCB200_ERR_UNSUPPORTED, "gru fwd: W_hh slice needs %zu B of shared memory (Har=%d, dtype %s)", smem, Har, g.bf16 ? "bf16" : "f32");
  GruLayout lay = gru_layout(g);
  char* sv = static_cast<char*>(save);
  Carver ws(wsp, ws_bytes);
  const int Hmax = g.H > Har ? g.H : Har;
  T* inT = isf_dummy<T>() ? ws.take<T>(1) : reinterpret_cast<T*>(sv + lay.inT);
  T* wih = ws.take<T>((size_t)3 * Har * Hmax);
  T* gi = ws.take<T>((size_t)B * S * 3 * Har);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "gru_fwd: workspace %zu < %zu", ws_bytes, ws.off);
  constexpr bool isf = sizeof(T) == 4;

  for (int l = 0; l < g.nL; l++) {
    const int Hin = l == 0 ? g.H : Har;
    const T* in;
    const T* w;
    if (isf) {
      in = reinterpret_cast<const T*>(l == 0 ? z : reinterpret_cast<const float*>(sv + lay.cf[l - 1]));
      w = reinterpret_cast<const T*>(p->w_ih[l]);
    } else {
      if (l == 0) {
        CPC_TRY(launch_cast3_bf16(z, reinterpret_cast<bf16*>(inT), (long long)B * S * Hin, p->w_ih[l], reinterpret_cast<bf16*>(wih),
                                  (long long)3 * Har * Hin, nullptr, nullptr, 0, st));
        in = inT;
      } else {
        in = reinterpret_cast<const T*>(sv + lay.cT[l - 1]);
        CPC_TRY(launch_cast<T>(p->w_ih[l], wih, (long long)3 * Har * Hin, st));
      }
      w = wih;
    }
    RowView A{in, 0, (long long)Hin, B * S};
    OutView C{gi, 0, (long long)3 * Har, B * S, 0, B * S, 0};
    CPC_TRY(gemm_nt(g.bf16, false, 1, 3 * Har, Hin, A, w, p->b_ih[l], C, st));

    const bool last = l == g.nL - 1;
    float* cout = last ? c : reinterpret_cast<float*>(sv + lay.cf[l]);
    T* cTo = isf ? nullptr : reinterpret_cast<T*>(sv + lay.cT[l]);
    T* sR = reinterpret_cast<T*>(sv + lay.gates[l][0]);
    T* sU = reinterpret_cast<T*>(sv + lay.gates[l][1]);
    T* sN = reinterpret_cast<T*>(sv + lay.gates[l][2]);
    T* sHN = reinterpret_cast<T*>(sv + lay.gates[l][3]);
    const float* h0l = h0 ? h0 + (size_t)l * B * Har : nullptr;
    float* hTl = hT ? hT + (size_t)l * B * Har : nullptr;
    const T* gic = gi;
    const float* whh = p->w_hh[l];
    const float* bhh = p->b_hh[l];
    int Bv = B, Sv = S, Hv = Har;
    bool done = false;
    if constexpr (!isf) {
      const bool contiguous = ((size_t)B * S * Har * 2) % 256 == 0;  // the four gate arrays form one array of quadruples
      if (contiguous && gru_wide_preferred(Har)) {  // 32 units per CTA: Har = 512 (config 5)
        CPC_TRY(gru_rec_fwd_wide(gic, whh, bhh, h0l, cout, cTo, sR, hTl, B, S, Har, st));
        done = true;
      } else if (gru_mma_supported(Har) && contiguous) {  // tensor-core recurrence, W_hh slice resident in registers
        CPC_TRY(gru_rec_fwd_mma(gic, whh, bhh, h0l, cout, cTo, sR, sU, sN, sHN, hTl, B, S, Har, st));
        done = true;
      }
    }
    if (!done) {
      void* args[] = {&gic, &whh, &bhh, &h0l, &cout, &cTo, &sR, &sU, &sN, &sHN, &hTl, &Bv, &Sv, &Hv};
      if (bt == kBT) CPC_TRY(launch_cluster("gru_rec_fwd", gru_rec_fwd_kernel<WT, T, kBT>, Har / HC, (B + kBT - 1) / kBT, 3 * HC * 2, smem, st, args));
      else CPC_TRY(launch_cluster("gru_rec_fwd", gru_rec_fwd_kernel<WT, T, 1>, Har / HC, B, 3 * HC * 2, smem, st, args));
    }
  }
  return 0;
}

template <class T>
int gru_bwd_t(const Geo& g, const float* z, const float* h0, const cpcb200_gru_params* p, const float* c, const float* dc,
              const void* save, float* dz, const cpcb200_gru_params* gr, void* wsp, size_t ws_bytes, cudaStream_t st) {
  typedef T WT;
  const int B = g.B, S = g.S, Har = g.Har, G = 3 * Har;
  if (Har % HC != 0 || Har / HC > 8) return fail(CPCB200_ERR_UNSUPPORTED, "gru: Har=%d", Har);
  int bt = kBT;
  size_t smem = rec_bwd_smem<WT>(Har, bt);
  if (smem > 227 * 1024) { bt = 1; smem = rec_bwd_smem<WT>(Har, bt); }
  if (smem > 227 * 1024) return fail(CPCB200_ERR_UNSUPPORTED, "gru bwd: W_hh slice needs %zu B of shared memory", smem);
  GruLayout lay = gru_layout(g);
  const char* sv = static_cast<const char*>(save);
  Carver ws(wsp, ws_bytes);
  const int Hmax = g.H > Har ? g.H : Har;
  const T* inT = isf_dummy<T>() ? nullptr : reinterpret_cast<const T*>(sv + lay.inT);
  T* wihT = ws.take<T>((size_t)G * Hmax);
  T* dgi = ws.take<T>((size_t)B * S * G);
  T* dgh = ws.take<T>((size_t)B * S * G);
  T* h0T = ws.take<T>((size_t)B * Har);
  float* dmid = ws.take<float>(g.nL > 1 ? (size_t)B * S * Har : 1);
  float* dmid2 = ws.take<float>(g.nL > 2 ? (size_t)B * S * Har : 1);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "gru_bwd: workspace %zu < %zu", ws_bytes, ws.off);
  constexpr bool isf = sizeof(T) == 4;

  const float* dcl = dc;
  for (int l = g.nL - 1; l >= 0; l--) {
    const int Hin = l == 0 ? g.H : Har;
    const bool last = l == g.nL - 1;
    const float* cl = last ? c : reinterpret_cast<const float*>(sv + lay.cf[l]);
    const T* sR = reinterpret_cast<const T*>(sv + lay.gates[l][0]);
    const T* sU = reinterpret_cast<const T*>(sv + lay.gates[l][1]);
    const T* sN = reinterpret_cast<const T*>(sv + lay.gates[l][2]);
    const T* sHN = reinterpret_cast<const T*>(sv + lay.gates[l][3]);
    const float* h0l = h0 ? h0 + (size_t)l * B * Har : nullptr;
    const float* whh = p->w_hh[l];
    float* dh0 = nullptr;
    int Bv = B, Sv = S, Hv = Har;
    bool done = false;
    if constexpr (!isf) {
      const bool contiguous = ((size_t)B * S * Har * 2) % 256 == 0;
      if (contiguous && gru_wide_preferred(Har)) {
        CPC_TRY(gru_rec_bwd_wide(dcl, cl, h0l, sR, whh, dgi, dgh, dh0, gr->b_ih[l], gr->b_hh[l], B, S, Har, st));
        done = true;
      } else if (gru_mma_supported(Har) && contiguous) {
        CPC_TRY(gru_rec_bwd_mma(dcl, cl, h0l, sR, sU, sN, sHN, whh, dgi, dgh, dh0, gr->b_ih[l], gr->b_hh[l], B, S, Har, st));
        done = true;
      }
    }
    if (!done) {
      void* args[] = {&dcl, &cl, &h0l, &sR, &sU, &sN, &sHN, &whh, &dgi, &dgh, &dh0, &Bv, &Sv, &Hv};
      if (bt == kBT) CPC_TRY(launch_cluster("gru_rec_bwd", gru_rec_bwd_kernel<WT, T, kBT>, Har / HC, (B + kBT - 1) / kBT, HC * 8, smem, st, args));
      else CPC_TRY(launch_cluster("gru_rec_bwd", gru_rec_bwd_kernel<WT, T, 1>, Har / HC, B, HC * 8, smem, st, args));
    }

    if (!done) {  // the tensor-core recurrence accumulates the bias gradients itself
      CPC_TRY(launch_colsum<T>(dgi, gr->b_ih[l], (long long)B * S, G, st));
      CPC_TRY(launch_colsum<T>(dgh, gr->b_hh[l], (long long)B * S, G, st));
    }

    // operands of the hoisted weight-gradient GEMMs
    const T* in;
    const T* hseq;
    if (isf) {
      in = reinterpret_cast<const T*>(l == 0 ? z : reinterpret_cast<const float*>(sv + lay.cf[l - 1]));
      hseq = reinterpret_cast<const T*>(cl);
    } else {
      if (l == 0) in = inT;
      else in = reinterpret_cast<const T*>(sv + lay.cT[l - 1]);
      hseq = reinterpret_cast<const T*>(sv + lay.cT[l]);
    }
    {  // dW_ih += dgi^T . in  and  dW_hh += sum_{t>=1} dgh_t^T . h_{t-1}: one grouped launch
      TnDesc wg[2];
      int nw = 0;
      wg[nw++] = TnDesc{1, G, Hin, RowView{dgi, 0, (long long)G, B * S}, RowView{in, 0, (long long)Hin, B * S}, gr->w_ih[l], Hin,
                        STORE_PLAIN, 0, 0};
      if (S > 1)
        wg[nw++] = TnDesc{B, G, Har, RowView{dgh + G, (long long)S * G, (long long)G, S - 1},
                          RowView{hseq, (long long)S * Har, (long long)Har, S - 1}, gr->w_hh[l], Har, STORE_PLAIN, 0, 0};
      CPC_TRY(gemm_tn_group(g.bf16, nw, wg, st));
    }
    if (h0l) {  // t = 0 term with the carried hidden state
      const T* h0p;
      if (isf) h0p = reinterpret_cast<const T*>(h0l);
      else { CPC_TRY(launch_cast<T>(h0l, h0T, (long long)B * Har, st)); h0p = h0T; }
      RowView A{dgh, (long long)S * G, (long long)G, 1};
      RowView Bv2{h0p, (long long)Har, (long long)Har, 1};
      CPC_TRY(gemm_tn(g.bf16, B, G, Har, A, Bv2, gr->w_hh[l], Har, STORE_PLAIN, 0, 0, st));
    }
    {  // d(input) = dgi . W_ih   (as NT against the transposed weights)
      CPC_TRY(launch_transpose_cast<T>(p->w_ih[l], wihT, G, Hin, st, 1));
      float* dst = l == 0 ? dz : (dcl == dmid ? dmid2 : dmid);
      RowView A{dgi, 0, (long long)G, B * S};
      OutView C{dst, 0, (long long)Hin, B * S, 0, B * S, 0};
      CPC_TRY(gemm_nt(g.bf16, true, 1, Hin, G, A, wihT, nullptr, C, st));
      dcl = dst;
    }
  }
  return 0;
}


// =========================================================================================================
// LSTM context network (cpc/model.py:171-173: nn.LSTM(batch_first=True), the reference's default --arMode,
// cpc_default_config.py:74).  torch gate order (i, f, g, o):
//   i = sigma(W_ii x + b_ii + W_hi h + b_hi)   f = sigma(..f..)   g = tanh(..g..)   o = sigma(..o..)
//   c' = f c + i g                              h' = o tanh(c')
// Same structure as the GRU: hoisted input projection / weight-gradient GEMMs, one persistent cluster kernel per
// direction with the W_hh slice resident in shared memory; a CTA owns 32 hidden units (4 x 32 gate rows), the cell
// state of a (unit, sequence) pair stays in the register of its gate thread for all S steps.
// =========================================================================================================
constexpr int LHC = 32;

template <class WT, class T, int BT>
__global__ void __launch_bounds__(256, 1)
lstm_rec_fwd_kernel(const T* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                    const float* __restrict__ h0, const float* __restrict__ c0, float* __restrict__ out, T* __restrict__ outT,
                    T* __restrict__ sI, T* __restrict__ sF, T* __restrict__ sG, T* __restrict__ sO, float* __restrict__ cell,
                    float* __restrict__ hT, float* __restrict__ cT, int B, int S, int Har) {
  constexpr int KS = 2, ROWS = 4 * LHC;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int tid = threadIdx.x;
  const int wstride = Har + WVec<WT>::V * KS;

  extern __shared__ __align__(16) unsigned char smraw[];
  WT* Wsm = reinterpret_cast<WT*>(smraw);                                                           // [ROWS][wstride]
  float* hsm = reinterpret_cast<float*>(smraw + align_up((size_t)ROWS * wstride * sizeof(WT), 16));  // [2][BT][Har]
  float* gsm = hsm + 2 * BT * Har;                                                                  // [ROWS][BT]
  float* bsm = gsm + ROWS * BT;                                                                     // [ROWS]

  // resident slice of W_hh: local row r = gate*32 + j  <->  global row gate*Har + 32*rank + j
  for (int i = tid; i < ROWS * (Har / 4); i += blockDim.x) {
    const int r = i / (Har / 4), k4 = (i - r * (Har / 4)) * 4;
    const int grow = (r / LHC) * Har + LHC * rank + (r % LHC);
    float4 v = *reinterpret_cast<const float4*>(w_hh + (size_t)grow * Har + k4);
    WT* dst = Wsm + (size_t)r * wstride + k4;
    dst[0] = from_f<WT>(v.x); dst[1] = from_f<WT>(v.y); dst[2] = from_f<WT>(v.z); dst[3] = from_f<WT>(v.w);
  }
  for (int i = tid; i < ROWS; i += blockDim.x) bsm[i] = b_hh[(i / LHC) * Har + LHC * rank + (i % LHC)];
  for (int i = tid; i < BT * Har; i += blockDim.x) {
    const int b = i / Har, k = i - b * Har;
    hsm[i] = (h0 != nullptr && b0 + b < B) ? h0[(size_t)(b0 + b) * Har + k] : 0.f;
  }
  cluster.sync();

  const int r = tid / KS, q = tid % KS;
  const WT* wrow = Wsm + (size_t)r * wstride;
  const bool gate_thread = tid < LHC * BT;
  const int gj = tid % LHC, gb = tid / LHC;  // gate thread -> (hidden unit, sequence)
  const int col = LHC * rank + gj;
  const bool gvalid = gate_thread && (b0 + gb < B);
  float cprev = (gvalid && c0 != nullptr) ? c0[(size_t)(b0 + gb) * Har + col] : 0.f;

  for (int t = 0; t < S; t++) {
    const float* hcur = hsm + (t & 1) * BT * Har;
    float* hnxt = hsm + ((t + 1) & 1) * BT * Har;
    float gin[4] = {0.f, 0.f, 0.f, 0.f};
    if (gvalid) {
      const T* g = gi + ((size_t)(b0 + gb) * S + t) * 4 * Har + col;
#pragma unroll
      for (int k = 0; k < 4; k++) gin[k] = to_f(g[(size_t)k * Har]);
    }
    float acc[BT];
    row_dot<WT, KS, BT>(wrow, hcur, Har, Har, q, acc);
    if (q == 0) {
#pragma unroll
      for (int b = 0; b < BT; b++) gsm[r * BT + b] = acc[b];
    }
    __syncthreads();
    if (gate_thread) {
      const float ai = gin[0] + gsm[gj * BT + gb] + bsm[gj];
      const float af = gin[1] + gsm[(LHC + gj) * BT + gb] + bsm[LHC + gj];
      const float ag = gin[2] + gsm[(2 * LHC + gj) * BT + gb] + bsm[2 * LHC + gj];
      const float ao = gin[3] + gsm[(3 * LHC + gj) * BT + gb] + bsm[3 * LHC + gj];
      const float ig = sigmoidf_(ai), fg = sigmoidf_(af), gg = tanhf(ag), og = sigmoidf_(ao);
      const float cn = fg * cprev + ig * gg;
      const float hn = og * tanhf(cn);
      cprev = cn;
      for (int pr = 0; pr < CS; pr++) {
        float* dst = cluster.map_shared_rank(hnxt, pr);
        dst[gb * Har + col] = hn;
      }
      if (gvalid) {
        const size_t o = ((size_t)(b0 + gb) * S + t) * Har + col;
        out[o] = hn;
        if (outT) outT[o] = from_f<T>(hn);
        if (sI != nullptr) {
          sI[o] = from_f<T>(ig); sF[o] = from_f<T>(fg); sG[o] = from_f<T>(gg); sO[o] = from_f<T>(og);
          cell[o] = cn;
        }
        if (t == S - 1) {
          if (hT != nullptr) hT[(size_t)(b0 + gb) * Har + col] = hn;
          if (cT != nullptr) cT[(size_t)(b0 + gb) * Har + col] = cn;
        }
      }
    }
    cluster.sync();
  }
}

// BPTT.  block = LHC*KS = 256 threads.  Wt[j][g] = W_hh[g][32*rank + j] resident (g over all 4*Har gate rows); per step:
//   gate gradients of the local 32 units -> all-gather over DSMEM -> dh_prev = dgates . W_hh[:, slice]
template <class WT, class T, int BT>
__global__ void __launch_bounds__(256, 1)
lstm_rec_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ c0, const T* __restrict__ sI,
                    const T* __restrict__ sF, const T* __restrict__ sG, const T* __restrict__ sO, const float* __restrict__ cell,
                    const float* __restrict__ w_hh, T* __restrict__ dg, int B, int S, int Har) {
  constexpr int KS = 8;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b0 = (blockIdx.x / CS) * BT;
  const int tid = threadIdx.x;
  const int G = 4 * Har;
  const int wstride = G + WVec<WT>::V * KS;

  extern __shared__ __align__(16) unsigned char smraw[];
  WT* Wsm = reinterpret_cast<WT*>(smraw);                                                          // [LHC][wstride]
  float* dsm = reinterpret_cast<float*>(smraw + align_up((size_t)LHC * wstride * sizeof(WT), 16));  // [2][BT][G]
  float* osm = dsm + 2 * BT * G;                                                                   // [LHC][BT]

  for (int i = tid; i < LHC * G; i += blockDim.x) {
    const int g = i / LHC, j = i - g * LHC;
    Wsm[(size_t)j * wstride + g] = from_f<WT>(w_hh[(size_t)g * Har + LHC * rank + j]);
  }
  cluster.sync();

  const int r = tid / KS, q = tid % KS;
  const WT* wrow = Wsm + (size_t)r * wstride;
  const bool gate_thread = tid < LHC * BT;
  const int gj = tid % LHC, gb = tid / LHC;
  const int col = LHC * rank + gj;
  const bool gvalid = gate_thread && (b0 + gb < B);
  float dh_carry = 0.f, dc_carry = 0.f;

  for (int t = S - 1; t >= 0; t--) {
    float* dbuf = dsm + (t & 1) * BT * G;
    if (gate_thread) {
      float da[4] = {0.f, 0.f, 0.f, 0.f};
      if (gvalid) {
        const size_t o = ((size_t)(b0 + gb) * S + t) * Har + col;
        const float dh = dh_carry + dout[o];
        const float ig = to_f(sI[o]), fg = to_f(sF[o]), gg = to_f(sG[o]), og = to_f(sO[o]);
        const float cn = cell[o];
        const float cp = t > 0 ? cell[o - Har] : (c0 ? c0[(size_t)(b0 + gb) * Har + col] : 0.f);
        const float tc = tanhf(cn);
        const float dc = dc_carry + dh * og * (1.f - tc * tc);
        da[0] = dc * gg * ig * (1.f - ig);
        da[1] = dc * cp * fg * (1.f - fg);
        da[2] = dc * ig * (1.f - gg * gg);
        da[3] = dh * tc * og * (1.f - og);
        dc_carry = dc * fg;
        const size_t og_ = ((size_t)(b0 + gb) * S + t) * G + col;
#pragma unroll
        for (int k = 0; k < 4; k++) dg[og_ + (size_t)k * Har] = from_f<T>(da[k]);
      }
      for (int pr = 0; pr < CS; pr++) {
        float* dst = cluster.map_shared_rank(dbuf, pr) + gb * G + col;
#pragma unroll
        for (int k = 0; k < 4; k++) dst[k * Har] = da[k];
      }
    }
    cluster.sync();
    float acc[BT];
    row_dot<WT, KS, BT>(wrow, dbuf, G, G, q, acc);
    if (q == 0) {
#pragma unroll
      for (int b = 0; b < BT; b++) osm[r * BT + b] = acc[b];
    }
    __syncthreads();
    if (gate_thread) dh_carry = osm[gj * BT + gb];
    __syncthreads();
  }
}

template <class WT> size_t lstm_fwd_smem(int Har) {
  const int wstride = Har + WVec<WT>::V * 2;
  return align_up((size_t)4 * LHC * wstride * sizeof(WT), 16) + (size_t)(2 * kBT * Har + 4 * LHC * kBT + 4 * LHC) * sizeof(float);
}
template <class WT> size_t lstm_bwd_smem(int Har) {
  const int wstride = 4 * Har + WVec<WT>::V * 8;
  return align_up((size_t)LHC * wstride * sizeof(WT), 16) + (size_t)(2 * kBT * 4 * Har + LHC * kBT) * sizeof(float);
}

struct LstmLayout {
  size_t gates[CPCB200_MAX_GRU_LAYERS][4];  // byte offsets: I, F, G, O (T)
  size_t cell[CPCB200_MAX_GRU_LAYERS];      // fp32 cell-state sequence
  size_t hT[CPCB200_MAX_GRU_LAYERS];        // T copy of the layer output (bf16 path only)
  size_t hf[CPCB200_MAX_GRU_LAYERS];        // fp32 output of non-last layers
  size_t inT;                               // bf16 copy of the layer-0 input, bf16 path only
  size_t total;
};
LstmLayout lstm_layout(const Geo& g) {
  LstmLayout l{};
  size_t off = 0;
  const size_t es = g.bf16 ? 2 : 4;
  const size_t n = (size_t)g.B * g.S * g.Har;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes); return r; };
  for (int i = 0; i < g.nL; i++) {
    for (int k = 0; k < 4; k++) l.gates[i][k] = take(n * es);
    l.cell[i] = take(n * 4);
    l.hT[i] = g.bf16 ? take(n * es) : 0;
    l.hf[i] = (i < g.nL - 1) ? take(n * 4) : 0;
  }
  l.inT = g.bf16 ? take((size_t)g.B * g.S * g.H * 2) : 0;
  l.total = off;
  return l;
}

template <class T>
int lstm_fwd_t(const Geo& g, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p, float* out, float* hT,
               float* cT, void* save, void* wsp, size_t ws_bytes, cudaStream_t st) {
  typedef T WT;
  const int B = g.B, S = g.S, Har = g.Har, G = 4 * Har;
  const int cs = Har / LHC;
  if (Har % LHC != 0 || cs > 16) return fail(CPCB200_ERR_UNSUPPORTED, "lstm: Har=%d needs a cluster of %d CTAs (max 16)", Har, cs);
  const size_t smem = lstm_fwd_smem<WT>(Har);
  if (smem > 227 * 1024) return fail(CPCB200_ERR_UNSUPPORTED, "lstm fwd: W_hh slice needs %zu B of shared memory (Har=%d, dtype %s)", smem, Har, g.bf16 ? "bf16" : "f32");
  constexpr bool isf = sizeof(T) == 4;
  const bool infer = save == nullptr;  // no_grad forward: gates / cell sequence are not kept
  LstmLayout lay = lstm_layout(g);
  char* sv = static_cast<char*>(save);
  Carver ws(wsp, ws_bytes);
  const int Hmax = g.H > Har ? g.H : Har;
  T* wih = ws.take<T>((size_t)G * Hmax);
  T* gi = ws.take<T>((size_t)B * S * G);
  T* inT = nullptr;
  T* hTs[2] = {nullptr, nullptr};
  float* hfs[2] = {nullptr, nullptr};
  if (!isf) inT = infer ? ws.take<T>((size_t)B * S * g.H) : reinterpret_cast<T*>(sv + lay.inT);
  if (infer && g.nL > 1) {
    for (int k = 0; k < 2; k++) { hfs[k] = ws.take<float>((size_t)B * S * Har); if (!isf) hTs[k] = ws.take<T>((size_t)B * S * Har); }
  }
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "lstm_fwd: workspace %zu < %zu", ws_bytes, ws.off);

  const float* in_f = z;
  const T* in_t = nullptr;
  for (int l = 0; l < g.nL; l++) {
    const int Hin = l == 0 ? g.H : Har;
    const T* in;
    const T* w;
    if (isf) {
      in = reinterpret_cast<const T*>(in_f);
      w = reinterpret_cast<const T*>(p->w_ih[l]);
    } else {
      if (l == 0) {
        CPC_TRY(launch_cast3_bf16(z, reinterpret_cast<bf16*>(inT), (long long)B * S * Hin, p->w_ih[l], reinterpret_cast<bf16*>(wih),
                                  (long long)G * Hin, nullptr, nullptr, 0, st));
        in = inT;
      } else {
        in = in_t;
        CPC_TRY(launch_cast<T>(p->w_ih[l], wih, (long long)G * Hin, st));
      }
      w = wih;
    }
    RowView A{in, 0, (long long)Hin, B * S};
    OutView C{gi, 0, (long long)G, B * S, 0, B * S, 0};
    CPC_TRY(gemm_nt(g.bf16, false, 1, G, Hin, A, w, p->b_ih[l], C, st));

    const bool last = l == g.nL - 1;
    float* o_f = last ? out : (infer ? hfs[l & 1] : reinterpret_cast<float*>(sv + lay.hf[l]));
    T* o_t = isf ? nullptr : (infer ? (last ? nullptr : hTs[l & 1]) : reinterpret_cast<T*>(sv + lay.hT[l]));
    T* sI = infer ? nullptr : reinterpret_cast<T*>(sv + lay.gates[l][0]);
    T* sF = infer ? nullptr : reinterpret_cast<T*>(sv + lay.gates[l][1]);
    T* sG = infer ? nullptr : reinterpret_cast<T*>(sv + lay.gates[l][2]);
    T* sO = infer ? nullptr : reinterpret_cast<T*>(sv + lay.gates[l][3]);
    float* cellp = infer ? nullptr : reinterpret_cast<float*>(sv + lay.cell[l]);
    const float* h0l = h0 ? h0 + (size_t)l * B * Har : nullptr;
    const float* c0l = c0 ? c0 + (size_t)l * B * Har : nullptr;
    float* hTl = hT ? hT + (size_t)l * B * Har : nullptr;
    float* cTl = cT ? cT + (size_t)l * B * Har : nullptr;
    const T* gic = gi;
    const float* whh = p->w_hh[l];
    const float* bhh = p->b_hh[l];
    int Bv = B, Sv = S, Hv = Har;
    void* args[] = {&gic, &whh, &bhh, &h0l, &c0l, &o_f, &o_t, &sI, &sF, &sG, &sO, &cellp, &hTl, &cTl, &Bv, &Sv, &Hv};
    if constexpr (!isf) {
      if (lstm_mma_supported(Har)) {  // tensor-core recurrence; the four gate arrays are one array of bf16 quadruples there
        CPC_TRY(lstm_rec_fwd_mma(gic, whh, bhh, h0l, c0l, o_f, o_t, sI, cellp, hTl, cTl, B, S, Har, st));
        in_f = o_f;
        in_t = o_t;
        continue;
      }
    }
    CPC_TRY(launch_cluster("lstm_rec_fwd", lstm_rec_fwd_kernel<WT, T, kBT>, cs, (B + kBT - 1) / kBT, 4 * LHC * 2, smem, st, args));
    in_f = o_f;
    in_t = o_t;
  }
  return 0;
}

template <class T>
int lstm_bwd_t(const Geo& g, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p, const float* out,
               const float* dout, const void* save, float* dz, const cpcb200_gru_params* gr, void* wsp, size_t ws_bytes,
               cudaStream_t st) {
  typedef T WT;
  const int B = g.B, S = g.S, Har = g.Har, G = 4 * Har;
  const int cs = Har / LHC;
  if (Har % LHC != 0 || cs > 16) return fail(CPCB200_ERR_UNSUPPORTED, "lstm: Har=%d", Har);
  const size_t smem = lstm_bwd_smem<WT>(Har);
  if (smem > 227 * 1024) return fail(CPCB200_ERR_UNSUPPORTED, "lstm bwd: W_hh slice needs %zu B of shared memory", smem);
  constexpr bool isf = sizeof(T) == 4;
  LstmLayout lay = lstm_layout(g);
  const char* sv = static_cast<const char*>(save);
  Carver ws(wsp, ws_bytes);
  const int Hmax = g.H > Har ? g.H : Har;
  T* wihT = ws.take<T>((size_t)G * Hmax);
  T* dg = ws.take<T>((size_t)B * S * G);
  T* h0T = ws.take<T>((size_t)B * Har);
  float* dmid = ws.take<float>(g.nL > 1 ? (size_t)B * S * Har : 1);
  float* dmid2 = ws.take<float>(g.nL > 2 ? (size_t)B * S * Har : 1);
  if (!ws.ok()) return fail(CPCB200_ERR_WORKSPACE, "lstm_bwd: workspace %zu < %zu", ws_bytes, ws.off);
  const T* inT = isf ? nullptr : reinterpret_cast<const T*>(sv + lay.inT);

  const float* dl = dout;
  for (int l = g.nL - 1; l >= 0; l--) {
    const int Hin = l == 0 ? g.H : Har;
    const bool last = l == g.nL - 1;
    const float* hl = last ? out : reinterpret_cast<const float*>(sv + lay.hf[l]);
    const T* sI = reinterpret_cast<const T*>(sv + lay.gates[l][0]);
    const T* sF = reinterpret_cast<const T*>(sv + lay.gates[l][1]);
    const T* sG = reinterpret_cast<const T*>(sv + lay.gates[l][2]);
    const T* sO = reinterpret_cast<const T*>(sv + lay.gates[l][3]);
    const float* cellp = reinterpret_cast<const float*>(sv + lay.cell[l]);
    const float* h0l = h0 ? h0 + (size_t)l * B * Har : nullptr;
    const float* c0l = c0 ? c0 + (size_t)l * B * Har : nullptr;
    const float* whh = p->w_hh[l];
    int Bv = B, Sv = S, Hv = Har;
    void* args[] = {&dl, &c0l, &sI, &sF, &sG, &sO, &cellp, &whh, &dg, &Bv, &Sv, &Hv};
    bool done = false;
    if constexpr (!isf) {
      if (lstm_mma_supported(Har)) {
        CPC_TRY(lstm_rec_bwd_mma(dl, c0l, sI, cellp, whh, dg, gr->b_ih[l], gr->b_hh[l], B, S, Har, st));
        done = true;
      }
    }
    if (!done) {
      CPC_TRY(launch_cluster("lstm_rec_bwd", lstm_rec_bwd_kernel<WT, T, kBT>, cs, (B + kBT - 1) / kBT, LHC * 8, smem, st, args));
      // both bias vectors enter every pre-activation with coefficient 1: the same column sums
      CPC_TRY(launch_colsum<T>(dg, gr->b_ih[l], (long long)B * S, G, st));
      CPC_TRY(launch_colsum<T>(dg, gr->b_hh[l], (long long)B * S, G, st));
    }
    const T* in;
    const T* hseq;
    if (isf) {
      in = reinterpret_cast<const T*>(l == 0 ? z : reinterpret_cast<const float*>(sv + lay.hf[l - 1]));
      hseq = reinterpret_cast<const T*>(hl);
    } else {
      in = l == 0 ? inT : reinterpret_cast<const T*>(sv + lay.hT[l - 1]);
      hseq = reinterpret_cast<const T*>(sv + lay.hT[l]);
    }
    {  // dW_ih += dg^T . in  and  dW_hh += sum_{t>=1} dg_t^T . h_{t-1}: one grouped launch
      TnDesc wg[2];
      int nw = 0;
      wg[nw++] = TnDesc{1, G, Hin, RowView{dg, 0, (long long)G, B * S}, RowView{in, 0, (long long)Hin, B * S}, gr->w_ih[l], Hin,
                        STORE_PLAIN, 0, 0};
      if (S > 1)
        wg[nw++] = TnDesc{B, G, Har, RowView{dg + G, (long long)S * G, (long long)G, S - 1},
                          RowView{hseq, (long long)S * Har, (long long)Har, S - 1}, gr->w_hh[l], Har, STORE_PLAIN, 0, 0};
      CPC_TRY(gemm_tn_group(g.bf16, nw, wg, st));
    }
    if (h0l) {  // t = 0 term with the carried hidden state
      const T* h0p;
      if (isf) h0p = reinterpret_cast<const T*>(h0l);
      else { CPC_TRY(launch_cast<T>(h0l, h0T, (long long)B * Har, st)); h0p = h0T; }
      RowView A{dg, (long long)S * G, (long long)G, 1};
      RowView Bv2{h0p, (long long)Har, (long long)Har, 1};
      CPC_TRY(gemm_tn(g.bf16, B, G, Har, A, Bv2, gr->w_hh[l], Har, STORE_PLAIN, 0, 0, st));
    }
    {  // d(input) = dg . W_ih   (as NT against the transposed weights)
      CPC_TRY(launch_transpose_cast<T>(p->w_ih[l], wihT, G, Hin, st, 1));
      float* dst = l == 0 ? dz : (dl == dmid ? dmid2 : dmid);
      RowView A{dg, 0, (long long)G, B * S};
      OutView C{dst, 0, (long long)Hin, B * S, 0, B * S, 0};
      CPC_TRY(gemm_nt(g.bf16, true, 1, Hin, G, A, wihT, nullptr, C, st));
      dl = dst;
    }
  }
  return 0;
}

}  // namespace

size_t gru_save_bytes(const Geo& g) { return gru_layout(g).total + 256; }

size_t gru_ws_bytes(const Geo& g, int backward) {
  const size_t es = g.bf16 ? 2 : 4;
  const int Hmax = g.H > g.Har ? g.H : g.Har;
  size_t tot = align_up((size_t)g.B * g.S * g.H * es) + align_up((size_t)3 * g.Har * Hmax * es);
  if (!backward) tot += align_up((size_t)g.B * g.S * 3 * g.Har * es);
  else {
    tot += 2 * align_up((size_t)g.B * g.S * 3 * g.Har * es) + align_up((size_t)g.B * g.Har * es);
    tot += align_up(g.nL > 1 ? (size_t)g.B * g.S * g.Har * 4 : 4) + align_up(g.nL > 2 ? (size_t)g.B * g.S * g.Har * 4 : 4);
  }
  return tot + 256;
}

int gru_fwd(const Geo& g, const float* z, const float* h0, const cpcb200_gru_params* p, float* c, float* hT, void* save,
            void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.bf16) return gru_fwd_t<bf16>(g, z, h0, p, c, hT, save, ws, ws_bytes, st);
  return gru_fwd_t<float>(g, z, h0, p, c, hT, save, ws, ws_bytes, st);
}
int gru_bwd(const Geo& g, const float* z, const float* h0, const cpcb200_gru_params* p, const float* c, const float* dc,
            const void* save, float* dz, const cpcb200_gru_params* gr, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.bf16) return gru_bwd_t<bf16>(g, z, h0, p, c, dc, save, dz, gr, ws, ws_bytes, st);
  return gru_bwd_t<float>(g, z, h0, p, c, dc, save, dz, gr, ws, ws_bytes, st);
}

}  // namespace cpcb200

namespace cpcb200 {
size_t lstm_save_bytes(const Geo& g) { return lstm_layout(g).total + 256; }
size_t lstm_ws_bytes(const Geo& g, int mode) {  // mode 0: training forward, 1: backward, 2: inference forward
  const size_t es = g.bf16 ? 2 : 4;
  const int Hmax = g.H > g.Har ? g.H : g.Har;
  const size_t G = (size_t)4 * g.Har, n = (size_t)g.B * g.S;
  size_t tot = align_up(G * Hmax * es) + align_up(n * G * es);
  if (mode == 1) tot += align_up((size_t)g.B * g.Har * es) + align_up(g.nL > 1 ? n * g.Har * 4 : 4) + align_up(g.nL > 2 ? n * g.Har * 4 : 4);
  if (mode == 2) tot += align_up(n * g.H * es) + 2 * (align_up(n * g.Har * 4) + align_up(n * g.Har * es));
  return tot + 256;
}
int lstm_fwd(const Geo& g, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p, float* out, float* hT,
             float* cT, void* save, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.bf16) return lstm_fwd_t<bf16>(g, z, h0, c0, p, out, hT, cT, save, ws, ws_bytes, st);
  return lstm_fwd_t<float>(g, z, h0, c0, p, out, hT, cT, save, ws, ws_bytes, st);
}
int lstm_bwd(const Geo& g, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p, const float* out,
             const float* dout, const void* save, float* dz, const cpcb200_gru_params* gr, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.bf16) return lstm_bwd_t<bf16>(g, z, h0, c0, p, out, dout, save, dz, gr, ws, ws_bytes, st);
  return lstm_bwd_t<float>(g, z, h0, c0, p, out, dout, save, dz, gr, ws, ws_bytes, st);
}
}  // namespace cpcb200
