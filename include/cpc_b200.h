/*
 * cpc_b200.h - C ABI of libcpc_b200.so: the B200 (sm_100a) implementation of the CPC training-step hot path.
 *
 * Drop-in boundary.  The reference (facebookresearch/CPC_audio @ b98a1bd) is pure Python on
 * PyTorch: it has no FFI of its own, its "operator interface" for this path is the two torch.nn.Module
 * surfaces cpc/model.py:276-289 (CPCModel) and cpc/criterion/criterion.py:139-257 (CPCUnsupersivedCriterion).
 * Each entry point below replaces the torch-op sequence of the reference lines it cites; the Python mirror
 * of the module surfaces (cpc_audio_b200/model.py, criterion.py) binds them through ctypes.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *     the library never allocates, frees or retains device memory, and never synchronises the device.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, on the current device.
 *   - tensors crossing the ABI are fp32 (activations, parameters, gradients) or int64/int32 (indices), dense,
 *     row-major, CHANNEL-LAST: z/c are (B, S, H) exactly as CPCModel.forward returns them (model.py:287).
 *   - `dims.dtype` selects the storage type of the library-internal activations (saved buffers, workspace):
 *     CPCB200_F32 = everything fp32 on CUDA cores (tight parity); CPCB200_BF16 = bf16 storage, fp32
 *     accumulation, tensor cores (tcgen05) for the dense contractions.
 *   - `save`  = caller-allocated buffer written by *_fwd and read by the matching *_bwd (size: *_save_bytes);
 *     `ws`    = caller-allocated scratch, contents undefined after return (size: *_ws_bytes); both 256-B aligned.
 *   - return value: 0 on success, negative cpcb200_status on failure; message via cpcb200_last_error()
 *     (thread-local).  Unsupported configurations fail with CPCB200_ERR_UNSUPPORTED - there is no fallback.
 *   - re-entrant: no global mutable state besides per-device one-time kernel attributes and a launch counter.
 */
#ifndef CPC_B200_H_
#define CPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPCB200_VERSION 200

typedef enum {
  CPCB200_OK = 0,
  CPCB200_ERR_BAD_DIMS = -1,
  CPCB200_ERR_UNSUPPORTED = -2,
  CPCB200_ERR_WORKSPACE = -3,
  CPCB200_ERR_CUDA = -4,
  CPCB200_ERR_NULL = -5
} cpcb200_status;

typedef enum { CPCB200_F32 = 0, CPCB200_BF16 = 1 } cpcb200_dtype;

/* Shape of one local batch.  S = L/160 frames, W = S-K anchor positions (criterion.py:232). */
typedef struct {
  int32_t B;       /* windows in the local batch                                   */
  int32_t L;       /* samples per window, multiple of 160 (cpc_default_config.py:41) */
  int32_t H;       /* hiddenEncoder  (multiple of 64, <= 512)                      */
  int32_t Har;     /* hiddenGar      (multiple of 64, <= 512)                      */
  int32_t K;       /* nPredicts      (<= 16)                                       */
  int32_t N;       /* negativeSamplingExt                                          */
  int32_t nLayers; /* nLevelsGRU                                                   */
  int32_t dtype;   /* cpcb200_dtype                                                */
} cpcb200_dims;

/* Parameters of CPCEncoder (cpc/model.py:83-93): conv{i}.weight (H,Cin,k), conv{i}.bias (H),
 * batchNorm{i}.weight/.bias (1,H,1) viewed as (H).  The same struct carries the gradients in *_bwd. */
typedef struct {
  float* conv_w[5];
  float* conv_b[5];
  float* norm_w[5];
  float* norm_b[5];
} cpcb200_encoder_params;

/* Parameters of torch.nn.GRU as held by CPCAR.baseNet (cpc/model.py:175-176), gate order (r,z,n). */
#define CPCB200_MAX_GRU_LAYERS 4
typedef struct {
  float* w_ih[CPCB200_MAX_GRU_LAYERS]; /* (3Har, Hin)  */
  float* w_hh[CPCB200_MAX_GRU_LAYERS]; /* (3Har, Har)  */
  float* b_ih[CPCB200_MAX_GRU_LAYERS]; /* (3Har)       */
  float* b_hh[CPCB200_MAX_GRU_LAYERS]; /* (3Har)       */
} cpcb200_gru_params;

/* ---- library ------------------------------------------------------------------------------------------- */
int cpcb200_version(void);
const char* cpcb200_last_error(void);
/* number of kernels this library has launched since load (all threads); bench.py reports the delta */
uint64_t cpcb200_launch_count(void);

/* per-kernel timing for the roofline report: when enabled, a CUDA event is recorded on the launching stream
 * after every kernel; report = text lines "<kernel> <launches> <total_ms>" (synchronises the device). */
int cpcb200_prof_enable(int on);
int cpcb200_prof_report(char* buf, size_t cap);

/* ---- input side: AudioBatchData.__getitem__ + DataLoader collate + .cuda() (cpc/dataset.py:185-202,
 * cpc/train.py:81) for a pack that is RESIDENT in HBM.  data (n_samples) fp32 = AudioBatchData.data; starts (B) int64 = the
 * window start indices one batch of the sampler yields (dataset.py:361-408); out (B,1,L) fp32 (16-byte aligned);
 * labels (B) int64 or NULL = getSpeakerLabel(idx) (dataset.py:177-180) over the interval bounds `bounds` (n_bounds int64,
 * bounds[0] = 0: AudioBatchData.speakerLabel).  A start outside [0, n_samples - L] sets *err (device int32, caller-zeroed)
 * and leaves that window untouched.  All pointers are device pointers. */
int cpcb200_gather_windows(const float* data, int64_t n_samples, const int64_t* starts, int B, int L, float* out,
                           const int64_t* bounds, int n_bounds, int64_t* labels, int32_t* err, void* stream);

/* ---- CPCEncoder.forward  (cpc/model.py:99-105: 5 x [Conv1d -> ChannelNorm(model.py:50-58) -> ReLU]) -------
 * x (B,1,L) fp32  ->  z (B,S,H) fp32, channel-last (what model.py:287 obtains with .permute(0,2,1)). */
size_t cpcb200_encoder_save_bytes(const cpcb200_dims* d);
size_t cpcb200_encoder_ws_bytes(const cpcb200_dims* d, int backward);
int cpcb200_encoder_fwd(const cpcb200_dims* d, const float* x, const cpcb200_encoder_params* p, float* z,
                        void* save, void* ws, size_t ws_bytes, void* stream);
/* dz (B,S,H) fp32 -> parameter gradients ACCUMULATED (+=) into `grads` (fp32, caller zero-fills). */
int cpcb200_encoder_bwd(const cpcb200_dims* d, const float* x, const cpcb200_encoder_params* p,
                        const float* dz, const void* save, const cpcb200_encoder_params* grads, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- CPCAR.forward, GRU branch (cpc/model.py:185-204 -> torch.nn.GRU(batch_first=True)) -----------------
 * z (B,S,H) -> c (B,S,Har).  h0 (nLayers,B,Har) or NULL (= zeros, model.py:189 hidden=None);
 * hT (nLayers,B,Har) or NULL receives the final state (model.py:194-198 keepHidden). */
size_t cpcb200_gru_save_bytes(const cpcb200_dims* d);
size_t cpcb200_gru_ws_bytes(const cpcb200_dims* d, int backward);
int cpcb200_gru_fwd(const cpcb200_dims* d, const float* z, const float* h0, const cpcb200_gru_params* p,
                    float* c, float* hT, void* save, void* ws, size_t ws_bytes, void* stream);
/* dc (B,S,Har) -> dz (B,S,H) (overwritten), parameter gradients accumulated into `grads`. */
int cpcb200_gru_bwd(const cpcb200_dims* d, const float* z, const float* h0, const cpcb200_gru_params* p,
                    const float* c, const float* dc, const void* save, float* dz,
                    const cpcb200_gru_params* grads, void* ws, size_t ws_bytes, void* stream);

/* ---- CPCAR.forward, LSTM branch (cpc/model.py:171-173 -> torch.nn.LSTM(batch_first=True); --arMode LSTM is the reference's
 * default, cpc_default_config.py:74).  Same calling convention as the GRU; the parameter struct carries 4*Har gate rows
 * (torch order i, f, g, o).  h0 / c0 (nLayers,B,Har) or NULL (= zeros); hT / cT receive the final states (keepHidden).
 * mode of cpcb200_lstm_ws_bytes: 0 = training forward, 1 = backward, 2 = inference forward (save == NULL: nothing is kept). */
size_t cpcb200_lstm_save_bytes(const cpcb200_dims* d);
size_t cpcb200_lstm_ws_bytes(const cpcb200_dims* d, int mode);
int cpcb200_lstm_fwd(const cpcb200_dims* d, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p,
                     float* out, float* hT, float* cT, void* save, void* ws, size_t ws_bytes, void* stream);
int cpcb200_lstm_bwd(const cpcb200_dims* d, const float* z, const float* h0, const float* c0, const cpcb200_gru_params* p,
                     const float* out, const float* dout, const void* save, float* dz, const cpcb200_gru_params* grads,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- negative sampling index arithmetic (criterion.py:191-199) -----------------------------------------
 * batch_idx, seq_idx: the two raw torch.randint draws of criterion.py:181-189, int64, length B*N*W, flat
 * layout (B,N,W).  ext[i] = ((seq_idx[i] + i%W) mod S) + batch_idx[i]*S  as int32.  Bit-exact. */
int cpcb200_sample_ext_idx(const cpcb200_dims* d, const int64_t* batch_idx, const int64_t* seq_idx,
                           int32_t* ext, void* stream);

/* ---- PredictionNetwork linear heads + scoring + InfoNCE (criterion.py:106-117, 207-217, 245-257) -------
 * c (B,S,Har), z (B,S,H) fp32; w_pred (K,H,Har) fp32 = predictors.{k}.weight stacked; ext (B,N,W) int32.
 * losses (K), acc (K) fp32: per-step mean cross-entropy against class 0 and argmax accuracy. */
size_t cpcb200_criterion_save_bytes(const cpcb200_dims* d);
size_t cpcb200_criterion_ws_bytes(const cpcb200_dims* d, int backward);
int cpcb200_criterion_fwd(const cpcb200_dims* d, const float* c, const float* z, const float* w_pred,
                          const int32_t* ext, float* losses, float* acc, void* save, void* ws,
                          size_t ws_bytes, void* stream);
/* dlosses (K) fp32 = d(total)/d(losses[k]).  dc (B,S,Har) and dz (B,S,H) are OVERWRITTEN; dw_pred (K,H,Har) is
 * ACCUMULATED (+=) like every other parameter gradient of the library (the caller zero-fills it, or passes its
 * gradient bucket). */
int cpcb200_criterion_bwd(const cpcb200_dims* d, const float* c, const float* z, const float* w_pred,
                          const int32_t* ext, const float* dlosses, const void* save, float* dc, float* dz,
                          float* dw_pred, void* ws, size_t ws_bytes, void* stream);

/* ---- rnnMode='transformer' prediction heads (criterion.py:82-88 -> cpc/transformers.py:98-139): K one-layer
 * post-LN transformers with relative-position attention (eval() or train() semantics, see att_keep below).  Every array is
 * the per-head parameter stacked over K: wq/wk/wv/wo (K,H,H) = multihead.W{q,k,v,o}.weight, krelpos (K, H/nheads, W)
 * = multihead.Att.Krelpos, ln1_* (K,H) = ln_multihead, w1 (K,dff,H) b1 (K,dff) = ffnetwork.lin1, w2 (K,H,dff) b2 (K,H)
 * = ffnetwork.lin2, ln2_* (K,H) = ln_ffnetwork.  Requires Har == H and W <= 128.  The same struct carries the
 * gradients in _bwd (ACCUMULATED, caller zero-fills). */
typedef struct {
  float *wq, *wk, *wv, *wo, *krelpos, *ln1_w, *ln1_b, *w1, *b1, *w2, *b2, *ln2_w, *ln2_b;
  int32_t dff;     /* 2048 in the reference (transformers.py:98) */
  int32_t nheads;  /* 8 */
  /* train() mode (dropout 0.1 at transformers.py:18,49 on the attention probabilities and :92 on the FFN hidden): the keep
   * masks torch's nn.Dropout draws - generated by the caller with the same torch calls, in the same order, as the
   * reference (so the same generator state gives the same masks) - as bytes (0 = dropped, 1 = kept):
   * att_keep (K, B*nheads, W, W), ffn_keep (K, B*W, dff); kept values are multiplied by keep_scale = 1/(1-p).
   * Both NULL = eval() semantics.  Ignored in the gradient struct. */
  const uint8_t* att_keep;
  const uint8_t* ffn_keep;
  float keep_scale;
} cpcb200_thead_params;
size_t cpcb200_criterion_t_save_bytes(const cpcb200_dims* d, int dff, int nheads);
size_t cpcb200_criterion_t_ws_bytes(const cpcb200_dims* d, int dff, int nheads, int backward);
int cpcb200_criterion_t_fwd(const cpcb200_dims* d, const float* c, const float* z, const cpcb200_thead_params* p,
                            const int32_t* ext, float* losses, float* acc, void* save, void* ws, size_t ws_bytes,
                            void* stream);
int cpcb200_criterion_t_bwd(const cpcb200_dims* d, const float* c, const float* z, const cpcb200_thead_params* p,
                            const int32_t* ext, const float* dlosses, const void* save, float* dc, float* dz,
                            const cpcb200_thead_params* grads, void* ws, size_t ws_bytes, void* stream);

/* ---- transformer CONTEXT network (--arMode transformer: feature_loader.py:138-142 -> transformers.py:129-139, one
 * TransformerLayer(sizeSeq = S = L/160 <= 128, dmodel = H) over all frames of a window): x (B,S,H) -> y (B,S,H) fp32.
 * `p` holds ONE layer (the K dimension of the arrays is 1); att_keep (B*nheads, S, S) / ffn_keep (B*S, dff) as above. */
size_t cpcb200_tlayer_save_bytes(const cpcb200_dims* d, int dff, int nheads);
size_t cpcb200_tlayer_ws_bytes(const cpcb200_dims* d, int dff, int nheads, int backward);
int cpcb200_tlayer_fwd(const cpcb200_dims* d, const float* x, const cpcb200_thead_params* p, float* y, void* save, void* ws,
                       size_t ws_bytes, void* stream);
int cpcb200_tlayer_bwd(const cpcb200_dims* d, const float* x, const cpcb200_thead_params* p, const float* dy, const void* save,
                       float* dx, const cpcb200_thead_params* grads, void* ws, size_t ws_bytes, void* stream);

/* One-shot hook for overlapping the data-parallel gradient exchange with the tail of the backward pass: the NEXT
 * cpcb200_encoder_bwd call enqueued ON `stream` records `cuda_event` (a cudaEvent_t) on that stream at the point where every
 * parameter gradient of the step except those of conv0 / batchNorm0 is final (the layer 1-4 weight gradients run before
 * the last data gradient).  Pass cuda_event = NULL to disarm.  The hook is keyed by the stream: DataParallel threads /
 * several streams of one process do not see each other's events. */
int cpcb200_encoder_bwd_set_event(void* stream, void* cuda_event);

/* ---- fused Adam over a flat fp32 bucket (cpc/train.py:335-337,90-91; torch.optim.Adam semantics) -------- */
int cpcb200_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);
/* Same update with the step count kept ON THE DEVICE, so that the launch can be captured in a CUDA graph and replayed
 * (torch.optim.Adam(capturable=True) semantics): `state` = CPCB200_OPT_STATE_WORDS (16) int32 words, zero-initialised by
 * the caller and 8-byte aligned:
 *   [0] steps taken so far (incremented by the kernel after every block has read it)   [1] block ticket (scratch)
 *   [2..5] beta1^(steps+1), beta2^(steps+1) as running float64 products                [6] node-barrier epoch (peer exchange)
 *   [7] error flag of the peer exchange (0 = ok; 1/2/3 = a peer did not arrive within the timeout at the first / second
 *       barrier of the step kernel / at the early exchange: the update was NOT applied, later calls return at once)
 *   [8] learning rate as float bits, used when the `lr` ARGUMENT is negative - the host can then change the learning rate
 *       of a captured graph (lr_scheduler.step(), cpc/train.py:351-370) with one 4-byte upload
 *   [9], [10] epoch / block ticket of the early exchange; the rest is reserved.
 * zero_grad != 0 also clears `grad` (optimizer.zero_grad() of cpc/train.py:91 folded into the same pass). */
#define CPCB200_OPT_STATE_WORDS 16
int cpcb200_adam_step_dev(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int32_t* state, int zero_grad,
                          void* stream);

/* ---- data-parallel exchange fused with the optimizer (the ONE collective of the path) --------------
 * One kernel per step does  all-reduce(sum) of the flat gradient bucket over the `world` GPUs of the node  +  the Adam
 * update of cpcb200_adam_step_dev  +  zero_grad, over PEER MEMORY (NVLink / NVSwitch loads and stores), instead of an
 * NCCL all-reduce followed by an optimizer kernel:
 *   barrier (every rank's backward pass is complete) -> rank r sums slice r of all `world` gradient buffers and writes the
 *   sum back into slice r of every buffer (two-shot all-reduce, in place) -> barrier -> every rank applies Adam to its
 *   full replica from its now-reduced local buffer and clears it.
 * grads[i] = device pointer to rank i's gradient bucket (n floats, peer-mapped on this device; grads[rank] is the local
 * one), signals[i] = rank i's signal words (>= 64 x uint32, zero-initialised, peer-mapped), both typically from
 * torch.distributed._symmetric_memory.  state as in cpcb200_adam_step_dev (zero-initialised).  Every rank must call it
 * the same number of times.  The launch is cooperative (all CTAs co-resident) and graph-capturable.
 * A rank that waits longer than timeout_ns for its peers sets state[7], skips the update and returns normally (no trap):
 * the host reads state[7] when it next synchronises.
 *
 * `ranges` (n_ranges <= 4 pairs [lo, hi) in floats, lo a multiple of 4; NULL = the whole bucket) restricts the exchange to
 * the parts of the bucket that have NOT been exchanged yet: with cpcb200_peer_reduce_range the bulk of the bucket (every
 * gradient that is final before the last data-gradient GEMM, see cpcb200_encoder_bwd_set_event) is all-reduced by a small
 * kernel on a side stream WHILE the backward pass finishes, and the step kernel only exchanges the few KB of conv0 /
 * batchNorm0 gradients.  Adam always covers the full bucket. */
typedef struct {
  void* grads[8];
  void* signals[8];
  int32_t rank, world;
  void* grads_mc;  /* optional NVSwitch multicast mapping of the same gradient buckets (NULL: peer loads/stores): the
                    * reduction of a slice is then ONE multimem.ld_reduce per 16 bytes, done inside the switch, and the
                    * write-back ONE multimem.st */
  int64_t timeout_ns; /* how long a barrier may wait for the peers (<= 0: 600 s) */
} cpcb200_peers;
int cpcb200_allreduce_adam_step(const cpcb200_peers* peers, float* param, float* exp_avg, float* exp_avg_sq, size_t n,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int32_t* state,
                                int zero_grad, const int64_t* ranges, int n_ranges, void* stream);
/* Early exchange: in-place all-reduce(sum) of `ranges` of the gradient buckets over peer memory, enqueued on `stream` (a
 * side stream that waits for the event of cpcb200_encoder_bwd_set_event).  Ordinary launch of a few CTAs (CPC_B200_EARLY_CTAS,
 * default 32; 256 threads x <= 64 registers each) that co-reside with the GEMM / conv0 kernels of the backward tail.  Every rank must call it once per step,
 * before that step's cpcb200_allreduce_adam_step, which must then be given the complementary ranges. */
int cpcb200_peer_reduce_range(const cpcb200_peers* peers, const int64_t* ranges, int n_ranges, int32_t* state, void* stream);

/* ---- test hooks: the GEMM building blocks, exposed so tests can pin them against torch.matmul ----------
 * C[M,N] = A[M,Kd] * B[N,Kd]^T (+bias[N]) ; C2[N1,N2] += A[M,N1]^T * B[M,N2].  dtype as in cpcb200_dims. */
int cpcb200_test_gemm_nt(int dtype, int M, int N, int Kd, const void* A, const void* B, const float* bias,
                         float* C, void* stream);
int cpcb200_test_gemm_tn(int dtype, int M, int N1, int N2, const void* A, const void* B, float* C,
                         void* stream);
/* same product with C in the activation dtype (bf16 when dtype == CPCB200_BF16), as the conv layers store it */
int cpcb200_test_gemm_nt_act(int dtype, int M, int N, int Kd, const void* A, const void* B, const float* bias,
                             void* C, void* stream);
/* debug: per-CTA cycle stamps (148 x 8 uint64, relative to each CTA's start) of the last persistent NT GEMM launch:
 * [0] prologue done, [1] first TMA issued, [2] first operands landed, [3] last MMA committed, [4] first accumulator
 * ready, [5] last epilogue done, [6] teardown done, [7] last accumulator ready.  Synchronises the device.  The stamps are
 * only written by a library built with -DCPC_B200_TIMELINE (CPC_B200_TIMELINE=1 python -m cpc_audio_b200.build -f); a
 * release build returns zeros. */
int cpcb200_debug_gemm_timeline(unsigned long long* host_out);

#ifdef __cplusplus
}
#endif
#endif /* CPC_B200_H_ */
