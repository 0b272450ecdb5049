// feeder.cu - input side of the training step: batches are cut ON THE GPU out of an audio pack that is
// resident in HBM, instead of being sliced sample by sample on the host and copied every step.
//
// Reference: cpc/dataset.py:185-202 (AudioBatchData.__getitem__: outData = data[idx : idx + sizeWindow].view(1, -1),
// label = getSpeakerLabel(idx), :177-180 = index of the speaker interval that contains idx), collated by the DataLoader
// into (B, 1, sizeWindow) / (B) and moved with .cuda() at cpc/train.py:81.  A pack (`AudioBatchData.data`, up to
// MAX_SIZE_LOADED = 4e9 samples = 16 GB, dataset.py:28) fits in the 180 GB of HBM many times over.
//
// The kernel is a pure HBM copy (B x L x 4 bytes read + written, 10.5 MB for the default batch): window starts are
// arbitrary sample offsets, so the source is only 4-byte aligned - each thread moves 4 consecutive samples with scalar
// loads (coalesced across the warp) and one 16-byte store.
#include "common.cuh"

namespace cpcb200 {

namespace {

__global__ void __launch_bounds__(256) gather_windows_kernel(const float* __restrict__ data, long long n,
                                                             const long long* __restrict__ starts, int L, float* __restrict__ out,
                                                             const long long* __restrict__ bounds, int n_bounds,
                                                             long long* __restrict__ labels, int* __restrict__ err) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  const long long s0 = starts[b];
  if (s0 < 0 || s0 + L > n) {  // the reference would silently return a short slice (dataset.py:187-188 only prints)
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(err, 1);
    return;
  }
  const float* src = data + s0;
  float* dst = out + (long long)b * L;
  const int L4 = L >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L4; i += gridDim.x * blockDim.x) {
    const float4 v = make_float4(__ldg(src + 4 * i), __ldg(src + 4 * i + 1), __ldg(src + 4 * i + 2), __ldg(src + 4 * i + 3));
    reinterpret_cast<float4*>(dst)[i] = v;
  }
  if (blockIdx.x == 0) {
    for (int i = 4 * L4 + threadIdx.x; i < L; i += blockDim.x) dst[i] = __ldg(src + i);
    if (threadIdx.x == 0 && labels != nullptr) {
      // getSpeakerLabel (dataset.py:177-180): first boundary > idx, minus one  ==  upper_bound(bounds, idx) - 1
      int lo = 0, hi = n_bounds;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bounds[mid] > s0) hi = mid; else lo = mid + 1;
      }
      labels[b] = (long long)lo - 1;
    }
  }
}

}  // namespace

int gather_windows(const float* data, long long n, const long long* starts, int B, int L, float* out, const long long* bounds,
                   int n_bounds, long long* labels, int* err, cudaStream_t st) {
  if (B <= 0 || L <= 0 || n < L) return fail(CPCB200_ERR_BAD_DIMS, "gather_windows: B=%d L=%d n=%lld", B, L, n);
  if (reinterpret_cast<uintptr_t>(out) & 15) return fail(CPCB200_ERR_BAD_DIMS, "gather_windows: output must be 16-byte aligned");
  if ((L & 3) && B > 1) return fail(CPCB200_ERR_BAD_DIMS, "gather_windows: L=%d must be a multiple of 4 when B > 1", L);
  int bx = (L / 4 + 255) / 256;
  if (bx > 32) bx = 32;
  if (bx < 1) bx = 1;
  CPC_CHECK_CUDA(launch_k(gather_windows_kernel, dim3(bx, B), dim3(256), 0, st, 1, data, n, starts, L, out, bounds, n_bounds, labels, err));
  CPC_LAUNCHED_N("gather_windows", st);
  return 0;
}

}  // namespace cpcb200
