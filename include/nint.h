/* nint.h -- C ABI of the B200-native Smart-NINT ConvLSTM hot path (libnint.so, sm_100a).
 *
 * The reference (smhassanerfani/nasa-niswan) has no FFI layer: its boundary for this path is the
 * Python class surface of model.py.  Each entry point below names the reference code it replaces
 * (file:line into the upstream repository).  Plain pointers and sizes only; every pointer is a
 * DEVICE pointer unless it says "host"; `stream` is a cudaStream_t passed as void*.
 * All functions return 0 on success and a non-zero code on failure; nint_last_error() returns a
 * thread-local description.  There is no CPU fallback: without a CUDA device every compute call
 * fails.
 *
 * Tensors crossing the boundary keep the reference's layouts:
 *   x      [B,T,C,H,W] fp32            (model.py:253-255)
 *   h, c   [B,Hc,H,W]  fp32            (model.py:216-217, 258-262)
 *   weight [4*Hc, C+Hc, k, k] fp32, bias [4*Hc]     (model.py:207-211; gate order i,f,g,o model.py:221)
 *   head   weight [1,Hc_last,1,1], bias [1]         (model.py:251)
 *   pred   [B,1,H,W] fp32 ; seq [B,T,H,W] fp32      (model.py:274 ; commented variant model.py:264,272)
 */
#ifndef NINT_H_
#define NINT_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NINT_MAX_LAYERS 8
#define NINT_DTYPE_BF16 0 /* bf16 operands, fp32 accumulation and cell state */
#define NINT_DTYPE_TF32 1 /* tf32 operands (fp32 storage), fp32 accumulation */

typedef struct nint_plan nint_plan;

/* Geometry of one ConvLSTM (model.py:235-251: ConvLSTM(input_channels, hidden_channels[],
 * kernel_size[], num_layers)) plus the batch/sequence shape of the calls it will serve. */
typedef struct nint_config {
  int32_t batch, seq_len, height, width; /* B, T, H, W of x (model.py:255) */
  int32_t in_channels;                   /* C */
  int32_t num_layers;                    /* L <= NINT_MAX_LAYERS */
  int32_t hidden[NINT_MAX_LAYERS];       /* Hc_l: any size in [1, 256] (model.py:207 takes any int); run internally padded
                                            to a multiple of 16 (64 above 64) with zero weights: results unchanged */
  int32_t ksize[NINT_MAX_LAYERS];        /* odd (model.py:204: padding = k // 2) */
  int32_t dtype;                         /* NINT_DTYPE_* */
  int32_t training;                      /* 1: keep gates/c/h of every step for BPTT */
  int32_t return_sequence;               /* 1: also apply the head at every t (model.py:264,272) */
  int32_t flags;                         /* NINT_FLAG_* */
} nint_config;

/* NINT_FLAG_DETERMINISTIC: gradients are bit-reproducible run to run (utils.py:77-88 asks cuDNN for the same): split-K
 * partial sums go to per-split buffers and are reduced in a fixed order instead of fp32 atomics.
 * NINT_FLAG_INPUT_GRAD: training plans also keep what nint_backward_input / nint_cell_backward need (the reference's
 * autograd yields x.grad and the gradient w.r.t. an explicit initial state, model.py:216-231). */
#define NINT_FLAG_DETERMINISTIC 1
#define NINT_FLAG_INPUT_GRAD 2
#define NINT_X_FP32 0
#define NINT_X_BF16 1

int nint_version(void);
const char* nint_last_error(void);

/* ---- plan life cycle (host-side object; owns no device memory) */
int nint_plan_create(const nint_config* cfg, nint_plan** out);
void nint_plan_destroy(nint_plan* plan);
/* device workspace the caller must provide (torch allocates it): activations, saved gates/states,
 * packed weights, gradient accumulators */
size_t nint_plan_workspace_bytes(const nint_plan* plan);
/* attaches the workspace (256-byte aligned device pointer), zero-fills it and encodes the TMA
 * tensor maps */
int nint_plan_bind(nint_plan* plan, void* workspace, size_t bytes, void* stream);

/* ---- parameters: repack the fp32 OIHW masters (state_dict names layers.{l}.conv.weight/bias,
 * conv.weight/bias; utils.py:27,39) into tensor-core operand panels.  Call after every
 * optimizer step. */
int nint_plan_set_weights(nint_plan* plan, int layer, const float* weight, const float* bias, void* stream);
int nint_plan_set_head(nint_plan* plan, const float* weight, const float* bias, void* stream);

/* ---- recurrent state.  model.py:258-262 starts every forward from zeros (the default);
 * ConvLSTMCell.forward (model.py:216-217) takes an explicit (h, c). */
int nint_plan_reset_state(nint_plan* plan, void* stream);
int nint_plan_set_state(nint_plan* plan, int layer, const float* h, const float* c, void* stream);
int nint_plan_get_state(nint_plan* plan, int layer, float* h, float* c, void* stream);

/* ---- ConvLSTM.forward (model.py:253-274): T x L fused cell steps + head.
 * pred [B,1,H,W]; seq [B,T,H,W] or NULL (requires return_sequence). */
int nint_forward(nint_plan* plan, const float* x, float* pred, float* seq, void* stream);
/* The same with x [B,T,C,H,W] given as bf16 (x_dtype = NINT_X_BF16; bf16 plans only): a host loader that stages its
 * windows in bf16 halves the host-to-device bytes of train.py:92, and the values the tensor cores see are identical
 * (the fp32 path rounds x to bf16 as it packs it). */
int nint_forward_ex(nint_plan* plan, const void* x, int x_dtype, float* pred, float* seq, void* stream);

/* ---- frame bank: the B200 form of dataset.py:551-637 (E33OMA90D_CRNN keeps the whole record in RAM and makes its
 * windows with sliding_window_view, consecutive windows sharing T-1 frames).  Here the record lives once in HBM in the
 * model's own operand layout -- bank [n_frames][H][W][c_pad] of bf16 (bf16 plans) or tf32-rounded fp32, channels-last,
 * padding lanes zero except lane `ones_lane` = 1.0 (nint_plan_input_layout) -- and a batch is B window start indices:
 * sample b reads frames win_start[b] .. win_start[b] + T - 1 straight through the TMA descriptors (no per-step copy or
 * packing pass).  win_start: device int32 [B]; it and the bank must stay valid until the backward of the step is
 * queued.  nint_pack_frames builds a bank from frames [n_frames,C,H,W] (fp32 or bf16); nint_fuse_inputs_bank builds it
 * from the raw fields (below). */
int nint_plan_input_layout(const nint_plan* plan, int* c_pad, int* ones_lane, int* elem_bytes);
int nint_pack_frames(int dtype, const void* frames, int x_dtype, long long n_frames, int channels, int height, int width,
                     int c_pad, int ones_lane, void* bank, void* stream);
int nint_forward_bank(nint_plan* plan, const void* bank, long long n_frames, const int* win_start, float* pred,
                      float* seq, void* stream);

/* ---- BPTT of the last nint_forward (replaces autograd over model.py:216-274, train.py:109).
 * dpred [B,1,H,W], dseq [B,T,H,W] or NULL.  grad_weight[l] / grad_bias[l] are host arrays of
 * device pointers with the parameters' shapes; gradients are WRITTEN (not accumulated). */
int nint_backward(nint_plan* plan, const float* dpred, const float* dseq, float* const* grad_weight,
                  float* const* grad_bias, float* grad_head_weight, float* grad_head_bias, void* stream);
/* The same in two stages, so a data-parallel caller can all-reduce one layer's gradient bucket while the next
 * layer's weight gradient is still being computed (train.py has no DDP):
 * nint_backward_bptt = head gradients + the reverse-time loop (leaves dgates of every step in the workspace),
 * nint_backward_wgrad = weight / bias gradient of ONE layer over all T steps; any layer order, each layer once. */
int nint_backward_bptt(nint_plan* plan, const float* dpred, const float* dseq, float* grad_head_weight,
                       float* grad_head_bias, void* stream);
int nint_backward_wgrad(nint_plan* plan, int layer, float* grad_weight, float* grad_bias, void* stream);
/* Gradient w.r.t. the input, dx [B,T,C,H,W] fp32 (x.grad of the reference's autograd through model.py:219-220); after
 * nint_backward_bptt, plans created with NINT_FLAG_INPUT_GRAD. */
int nint_backward_input(nint_plan* plan, float* dx, void* stream);

/* ---- ConvLSTMCell as a differentiable unit (model.py:216-231): one step from an explicit state, and its backward.
 * Plans with seq_len = 1, num_layers = 1, training = 1, NINT_FLAG_INPUT_GRAD.  nint_cell_forward = set_state + one
 * fused step + get_state; nint_cell_backward takes dL/dh', dL/dc' [B,Hc,H,W] (either may be NULL = zero) and writes
 * dx [B,C,H,W], dh, dc [B,Hc,H,W], grad_weight, grad_bias (any output may be NULL). */
int nint_cell_forward(nint_plan* plan, const float* x, const float* h, const float* c, float* h_out, float* c_out,
                      void* stream);
int nint_cell_backward(nint_plan* plan, const float* dh_out, const float* dc_out, float* dx, float* dh, float* dc,
                       float* grad_weight, float* grad_bias, void* stream);

/* ---- preprocessing fusion (north-star item 4).  Stacks the first `levels` model
 * levels of a 3-D forcing levels3d [frames,levels,H,W] with the 2-D emission field emis2d [frames,H,W] as the last
 * channel (dataset.py:526), z-scores every channel with mean/std [levels+1] (dataset.py:520-529), appends the
 * n_static already z-scored static attribute fields statics [n_static,H,W] to every frame (dataset.py:100-122,
 * 532-533; NULL / 0 for none) and adds the geophysical halo -- cyclic in longitude, reflected in latitude
 * (dataset.py:67-98, 535-536) -- giving out [frames,levels+1+n_static,padded_height,padded_width] fp32, the
 * model's input layout.  mode 0: true reflect
 * (dataset.py:38-53); mode 1: bug-compatible with the shipped RNN dataset (np.fliplr flips channels, dataset.py:96).
 * Same error conditions as dataset.py:80,98.  PARITY UNPINNED upstream for levels > 1 (no shipped code). */
int nint_fuse_inputs(const float* levels3d, const float* emis2d, const float* mean, const float* std,
                     const float* statics, int n_static, long long frames, int levels, int height, int width,
                     int padded_height, int padded_width, int mode, float* out, void* stream);
/* The same fusion written straight into the frame-bank layout (dtype = NINT_DTYPE_*): bank
 * [frames][padded_height][padded_width][c_pad], no fp32 NCHW intermediate, no second packing pass. */
int nint_fuse_inputs_bank(const float* levels3d, const float* emis2d, const float* mean, const float* std,
                          const float* statics, int n_static, long long frames, int levels, int height, int width,
                          int padded_height, int padded_width, int mode, int dtype, int c_pad, int ones_lane,
                          void* bank, void* stream);

/* ---- the rest of the training step (train.py:101-110), SURVEY.md section 8f rank 1.
 * nint_loss_mse_l1: loss = MSELoss(y, p) + L1Loss(y, p) (train.py:74-75,105) with p = pred[:, 0, y0:y1, x0:x1]
 * (train.py:102 crop; pass 0, H, 0, W for none); pred [B,1,H,W], y [B, y1-y0, x1-x0]; writes the scalar loss, and
 * d loss / d pred into dpred [B,1,H,W] (zero outside the crop; NULL = value only).  stats: 5 floats of device
 * scratch, left holding {sum (p-y)^2, sum |p-y|, sum y, sum y^2, -} (enough for an on-device R^2, train.py:114).
 * nint_adam_step: torch.optim.Adam (train.py:71; no weight decay / amsgrad) over one flat fp32 buffer; `step` counts
 * from 1; grads are multiplied by grad_scale first (1 / world_size after a sum all-reduce). */
int nint_loss_mse_l1(const float* pred, const float* y, int batch, int height, int width, int crop_y0, int crop_y1,
                     int crop_x0, int crop_x1, float* dpred, float* loss, float* stats, void* stream);
int nint_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, int step, float grad_scale, void* stream);
/* nint_loss_mse_l1 with the targets in a bank y [n_frames, y1-y0, x1-x0]: sample b's target is frame
 * win_start[b] + y_offset (dataset.py:600-601: y[seq_len - 1:] pairs window i with frame i + T - 1); an index outside
 * [0, n_frames) reads as a zero target (the input side reads zeros through TMA) instead of faulting. */
int nint_loss_mse_l1_bank(const float* pred, const float* ybank, long long n_frames, const int* win_start, int y_offset, int batch,
                          int height, int width, int crop_y0, int crop_y1, int crop_x0, int crop_x1, float* dpred,
                          float* loss, float* stats, void* stream);
/* nint_adam_step with the step count and learning rate in DEVICE memory: state = 4 floats {step, lr, -, -}; the call
 * increments state[0] itself.  Nothing in the launch depends on host values that change between steps, so a CUDA graph
 * of the whole training step can be replayed. */
int nint_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float* state,
                       float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- data-parallel step tail over NVLink peer memory (the reference has no multi-GPU code).
 * Replaces ncclAllReduce + nint_adam_step_dev by ONE kernel: cross-rank barrier, sum of every rank's gradients read
 * straight from the peers' memory in rank order (bit-identical on all ranks), Adam on this rank's parameters.
 * peer_buffers: HOST array of `world` device pointers -- the same symmetric allocation as mapped for each rank
 * ([rank] is this GPU's own).  Layout of the allocation: two gradient slots of n fp32 used alternately (the caller
 * writes step s's gradients to the slot at slot_offset_bytes, alternating by step parity) and a zero-initialised flag
 * block of 32 uint32 at flags_offset_bytes.  seq: step number, increasing from 1 on every rank.  state as in
 * nint_adam_step_dev.  Every rank must make the call for a step; a rank that never arrives traps the waiters after
 * ~20 s instead of hanging them. */
int nint_dp_allreduce_adam(const void* const* peer_buffers, long long slot_offset_bytes, long long flags_offset_bytes,
                           int rank, int world, unsigned seq, float* params, float* exp_avg, float* exp_avg_sq,
                           long long n, float* state, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---- measurement.  Kernel classes: 0 = fused gate-conv forward, 1 = dgrad + gate backward,
 * 2 = wgrad, 3 = everything else (layout packing, head, gradient unpacking); -1 = all.
 * nint_launch_count: kernels launched by this library in the calling process so far.
 * nint_plan_profile(plan, 1) brackets every later launch of the plan with CUDA events on the
 * launching stream; nint_plan_profile_read waits for them and returns, per class, the summed
 * device time in ms and the launch count since the last read (arrays of 4). */
long long nint_launch_count(int kernel_class);
int nint_plan_profile(nint_plan* plan, int enable);
int nint_plan_profile_read(nint_plan* plan, double* ms, long long* count);

/* ---- test hook: raw gate pre-activations of layer 0 at t = 0 without bias,
 * out [B,H,W,4*Hc] fp32 in kernel column order (see nint_gate_column). */
int nint_debug_raw_gates(nint_plan* plan, const float* x, float* out, void* stream);
/* debug timeline: with NINT_DEBUG_FLAGS & 8, CTA 0 of every conv launch records clock64() stamps per warp role
 * (8 roles x 1024 stamps; later launches overwrite earlier ones).  Copies them to `host`, optionally clears. */
int nint_debug_read_trace(long long* host, int n, int clear);
/* Post-mortem of a time-fused conv launch's step hand-off (a wait that lasts ~2 s traps instead of hanging the GPU, and
 * a trap destroys the context): the first call arms a host-mapped record, later calls -- also after a failed launch --
 * return {code (0 = none, 3 = hand-off timed out), block, thread, (step << 32) | image, counter value}. */
int nint_debug_fail_record(unsigned long long* out5);
/* reference gate channel n = gate*Hc + c for kernel column q of a layer with Hc hidden channels */
int nint_gate_column(int q, int hidden);
/* pixel tile chosen for a grid (host-only helper, no device needed) */
int nint_pick_tile(int height, int width, int* tile_w, int* tile_h);

#ifdef __cplusplus
}
#endif
#endif /* NINT_H_ */
