// gemm_simt.cu - CUDA-core GEMM building blocks (fp32 accumulate) over row views.
// Used by the CPCB200_F32 path for every dense contraction and by the BF16 path wherever the tcgen05 kernels
// (gemm_tc.cu) do not apply.  Two shapes:
//   gemm_nt : C[m,n]   = sum_k A[m,k] B[n,k] (+bias)      conv fwd / dgrad, GRU projections, prediction heads
//   gemm_tn : C[n1,n2] += sum_m A[m,n1] B[m,n2]           every weight gradient (reduction over positions)
#include "common.cuh"

namespace cpcb200 {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <class T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]) { load_vec<4>(p, v); }

template <class TI, class TO>
__global__ void __launch_bounds__(NT) gemm_nt_kernel(int M, int N, int Kd, RowView A, const TI* __restrict__ Bm,
                                                      const float* __restrict__ bias, OutView C) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row within tile, k offset
  const TI* Ap = static_cast<const TI*>(A.p);

  // per-thread source rows
  const int am = m0 + lr;
  const bool a_ok = am < M;
  long long a_off = 0;
  if (a_ok) { int b = am / A.rpb, t = am - b * A.rpb; a_off = (long long)b * A.bs + (long long)t * A.rs; }
  const int bn = n0 + lr;
  const bool b_ok = bn < N;
  const long long b_off = (long long)bn * Kd;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kd; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok) load4(Ap + a_off + k0 + lk, av);
    if (b_ok) load4(Bm + b_off + k0 + lk, bv);
#pragma unroll
    for (int i = 0; i < 4; i++) { As[lk + i][lr] = av[i]; Bs[lk + i][lr] = bv[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= N) return;
  float bb[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) { bb[0] = bias[n]; bb[1] = bias[n + 1]; bb[2] = bias[n + 2]; bb[3] = bias[n + 3]; }
  TO* Cp = static_cast<TO*>(C.p);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int b = m / C.rpb, t = m - b * C.rpb;
    if (!out_row_ok(C, t, n)) continue;
    float o[4] = {acc[i][0] + bb[0], acc[i][1] + bb[1], acc[i][2] + bb[2], acc[i][3] + bb[3]};
    if (C.relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
    store_vec<4>(Cp + (long long)b * C.bs + (long long)t * C.rs + n, o);
  }
}

template <class TI>
__global__ void __launch_bounds__(NT) gemm_tn_kernel(int M, int N1, int N2, RowView A, RowView B, float* __restrict__ Cacc,
                                                      int ldc, int mode, int Ci, int taps, int rows_per_split) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int n10 = blockIdx.y * BM, n20 = blockIdx.x * BN;
  const int mbeg = blockIdx.z * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);
  const int lr = tid >> 4, lc = (tid & 15) * 4;  // loader: row (0..15), col offset
  const TI* Ap = static_cast<const TI*>(A.p);
  const TI* Bp = static_cast<const TI*>(B.p);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int mk = mbeg; mk < mend; mk += BK) {
    int m = mk + lr;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      int b = m / A.rpb, t = m - b * A.rpb;
      if (n10 + lc < N1) load4(Ap + (long long)b * A.bs + (long long)t * A.rs + n10 + lc, av);
      if (n20 + lc < N2) load4(Bp + (long long)b * B.bs + (long long)t * B.rs + n20 + lc, bv);
    }
    *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int n1 = n10 + ty * 4 + i;
    if (n1 >= N1) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n2 = n20 + tx * 4 + j;
      if (n2 >= N2) continue;
      long long o;
      if (mode == STORE_CONV_W) { int tap = n2 / Ci, ci = n2 - tap * Ci; o = ((long long)n1 * Ci + ci) * taps + tap; }
      else o = (long long)n1 * ldc + n2;
      atomicAdd(Cacc + o, acc[i][j]);
    }
  }
}

}  // namespace

int gemm_nt_simt(bool bf16_in, bool out_f32, int nb, int N, int Kd, const RowView& A, const void* Bm,
                 const float* bias, const OutView& C, cudaStream_t st) {
  const int M = nb * A.rpb;
  if (M <= 0 || N <= 0) return 0;
  if (Kd % BK != 0 || N % 4 != 0) return fail(CPCB200_ERR_BAD_DIMS, "gemm_nt: Kd %% 16 / N %% 4 (Kd=%d N=%d)", Kd, N);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if (!bf16_in) {
    gemm_nt_kernel<float, float><<<grid, NT, 0, st>>>(M, N, Kd, A, static_cast<const float*>(Bm), bias, C);
  } else if (out_f32) {
    gemm_nt_kernel<bf16, float><<<grid, NT, 0, st>>>(M, N, Kd, A, static_cast<const bf16*>(Bm), bias, C);
  } else {
    gemm_nt_kernel<bf16, bf16><<<grid, NT, 0, st>>>(M, N, Kd, A, static_cast<const bf16*>(Bm), bias, C);
  }
  CPC_LAUNCHED_N("gemm_nt_simt", st);
  return 0;
}

int gemm_tn_simt(bool bf16_in, int nb, int N1, int N2, const RowView& A, const RowView& B, float* Cacc, int ldc,
                 int mode, int Ci, int taps, cudaStream_t st) {
  const int M = nb * A.rpb;
  if (M <= 0) return 0;
  if (A.rpb != B.rpb) return fail(CPCB200_ERR_BAD_DIMS, "gemm_tn: row views disagree");
  if (N1 % 4 != 0 || N2 % 4 != 0) return fail(CPCB200_ERR_BAD_DIMS, "gemm_tn: N %% 4");
  const int tiles = ((N1 + BM - 1) / BM) * ((N2 + BN - 1) / BN);
  int splits = (148 * 4 + tiles - 1) / tiles;
  int max_splits = (M + 4 * BK - 1) / (4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rps = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (M + rps - 1) / rps;
  dim3 grid((N2 + BN - 1) / BN, (N1 + BM - 1) / BM, splits);
  if (!bf16_in) gemm_tn_kernel<float><<<grid, NT, 0, st>>>(M, N1, N2, A, B, Cacc, ldc, mode, Ci, taps, rps);
  else gemm_tn_kernel<bf16><<<grid, NT, 0, st>>>(M, N1, N2, A, B, Cacc, ldc, mode, Ci, taps, rps);
  CPC_LAUNCHED_N("gemm_tn_simt", st);
  return 0;
}

}  // namespace cpcb200
