/*
 * vqw.h -- C-ABI of libvqw.so: the B200 (sm_100a) hot path of dhgrs/chainer-VQ-VAE.
 *
 * The reference has no FFI/plugin layer (it is pure Python on Chainer); its hot path is the
 * set of Chain.__call__ methods in net.py, utils.py and WaveNet/modules.py.  Each entry point
 * below replaces the library calls one of those methods makes; the reference lines are cited
 * per function (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * stub that binds them from the reference's side.
 *
 * Conventions (all entry points):
 *   - every pointer is CALLER-OWNED DEVICE memory (the library never allocates or frees);
 *   - tensors use the reference layout (B, C, T, 1) float32, contiguous, T fastest;
 *   - the call is ASYNCHRONOUS on `stream` (a cudaStream_t passed as void*);
 *   - return 0 = ok, <0 = argument error detected before launch (see vqw_last_error()),
 *     >0 = cudaError_t of the failed launch;
 *   - no global mutable state other than a thread-local error string.
 */
#ifndef VQW_H_
#define VQW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vqw_stream_t; /* cudaStream_t */

#define VQW_VERSION 100

int vqw_version(void);
const char* vqw_last_error(void);
/* number of CUDA kernels this process has launched through the library (diagnostic) */
long long vqw_launch_count(void);
/* measurement hook (bench.py's roofline leg): the next vqw_resnet_forward call in a tensor-core
 * mode records `start_event` on its stream right before the first fused block kernel and
 * `end_event` right after the last one (cudaEvent_t handles created with timing enabled), so
 * that (end - start) / n_blocks is the block kernel's launch duration inside the running step,
 * without the operand packing that precedes it.  One-shot; NULL handles clear it. */
int vqw_probe_forward_kernels(void* start_event, void* end_event);

/* ------------------------------------------------------------------------------------
 * VQ nearest-codebook lookup.  Replaces StraightThrough.forward, utils.py:176-211
 * (expand/broadcast/sub/square/sum(axis=2)/argmin(axis=1)/take/transpose).
 *   z (B,d,T) f32, W (k,d) f32  ->  idx (B,T) i32, e (B,d,T) f32 = W[idx] channel-first.
 * Distances are accumulated sequentially over d in fp32 without FMA contraction and ties
 * go to the lowest k, i.e. NumPy's semantics: indices are bit-exact.
 * Optional outputs (NULL to skip; each must be zeroed by the caller, they are accumulated):
 *   count (k) f32        per-code usage n_k
 *   zsum  (k,d) f32      per-code sum of assigned z
 *   sqerr (1) f64        sum over all elements of (z - e)^2  (loss2/loss3 numerator, net.py:90-91)
 */
int vqw_vq_forward(const float* z, const float* W, int32_t* idx, float* e, float* count,
                   float* zsum, double* sqerr, int B, int d, int T, int k, vqw_stream_t stream);

/* StraightThrough.backward for W, utils.py:222-230: gW[k,:] = sum_{idx==k} gy[b,:,t]
 * (float64 accumulation then cast to f32 like the reference's eye(k) dot).  gW is overwritten. */
int vqw_vq_backward_w(const float* gy, const int32_t* idx, float* gW, int B, int d, int T, int k,
                      vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Generic strided/dilated 1-D convolution family (fp32 SIMT).  Replaces the cuDNN/NumPy
 * convolution calls made by L.Convolution2D / L.DilatedConvolution2D in
 * net.py:12-17,34-43 (Encoder, ConditionEmbed), modules.py:127-141,151-159 (embed, proj1,
 * proj2) and the unfused pieces of the ResidualBlock backward.
 *
 *   out[b,m,t] = post( bias[m] + sum_s sum_k  w_s[m*wm_s + k*wk_s] * pre_s(in_s[b,k,ti]) )
 *   ti = (t*mul_s + shift_s) / div_s  (term dropped unless divisible and 0 <= ti < Tin_s)
 *   pre_s(v)  = v, relu(v), or v * (in_mask_s[b,k,ti] > 0)
 *   post(v)   = v (+ addend[b,m,t]) (relu) (* (out_mask[b,m,t] > 0)) (+ previous out if accumulate)
 */
#define VQW_MAX_SRC 4
typedef struct {
  const float* in;      /* (B, K, Tin) */
  const float* w;       /* weight base pointer (already offset to the tap) */
  const float* in_mask; /* optional (B, K, Tin): in is multiplied by (mask > 0) */
  int K;                /* input channels */
  int Tin;              /* input length */
  int wm, wk;           /* weight strides (in floats) along output channel m / input channel k */
  int mul, shift, div;  /* time index map */
  int relu_in;          /* apply relu to the input */
} vqw_conv_src;

typedef struct {
  int B, M, T;          /* batch, output channels, output length */
  int nsrc;
  vqw_conv_src src[VQW_MAX_SRC];
  const float* bias;    /* (M) or NULL */
  const float* addend;  /* (B, M, T) or NULL */
  const float* out_mask;/* (B, M, T) or NULL */
  int relu_out;
  int accumulate;       /* out += result instead of out = result */
  /* gate backward epilogue (modules.py:47-48 differentiated): when gate_tanh != NULL the GEMM
   * result is gz (M = Cd/2 rows) and TWO rows are written into out (B, 2M, T):
   *   out[b,m,t]   = gz * sig * (1 - tanh^2),  out[b,m+M,t] = gz * tanh * sig * (1 - sig) */
  const float* gate_tanh; /* (B, M, T) or NULL */
  const float* gate_sig;  /* (B, M, T) or NULL */
} vqw_conv_desc;

int vqw_conv_forward(const vqw_conv_desc* desc, float* out, vqw_stream_t stream);

/* Weight/bias gradient of the same family:
 *   gw[m*gm + k*gk] (+)= sum_{b,t} A(b,m,t) * Bv(b,k,ti),   gb[m] (+)= sum_{b,t} A(b,m,t)
 *   A  = a (* (a_mask > 0));  Bv = pre(in) (* in_mul);  ti as above.
 * gw/gb are ACCUMULATED with atomics (zero them first). */
typedef struct {
  int B, M, T;          /* a is (B, M, T) */
  const float* a;
  const float* a_mask;  /* optional (B, M, T) */
  const float* in;      /* (B, K, Tin) */
  const float* in_mul;  /* optional (B, K, Tin): in is multiplied elementwise (z = tanh*sig) */
  int K, Tin;
  int mul, shift, div;
  int relu_in;
  int gm, gk;           /* gw strides */
  /* all taps of a (k,1) filter in ONE launch: tap j reads the input at shift + j*tap_dshift and
   * accumulates into gw + j*tap_gw (ntaps <= 1: a single tap, the fields above as given) */
  int ntaps, tap_dshift, tap_gw;
} vqw_wgrad_desc;

int vqw_conv_wgrad(const vqw_wgrad_desc* desc, float* gw, float* gb /* or NULL */,
                   vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Fused WaveNet residual block.  Replaces ResidualBlock.__call__, modules.py:30-56
 * (dilated causal conv :40-41, condition_proj :44, split/tanh/sigmoid/mul :47-48,
 * res 1x1 + residual add :51-54, skip 1x1 :55) and ResidualNet's skip accumulation
 * (modules.py:92-95) in ONE kernel.
 *
 * Weights are in the reference's layout: conv_w (Cd,Cr,fs), cond_w (Cd,Cc), res_w (Cr,Cd/2),
 * skip_w (Cs,Cd/2), biases (Cd),(Cd),(Cr),(Cs).
 */
#define VQW_MODE_FP32 0      /* fp32 CUDA-core path, any channel count with Cd/2 <= 256 */
#define VQW_MODE_BF16X3 1    /* tcgen05, bf16 hi/lo split operands, 3 MMAs per product (~fp32) */
#define VQW_MODE_BF16 2      /* tcgen05, single bf16 pass (throughput mode, not parity) */
#define VQW_MODE_FP16 3      /* tcgen05, single IEEE-fp16 pass (11 significant bits: ~2e-4 relative per
                              * product, inside the 1e-3 parity bar), fp32 accumulation; the backward
                              * carries its gradients with a power-of-two scale chosen on the device */
#define VQW_MODE_FP16X3 4    /* tcgen05, IEEE-fp16 hi/lo split operands (22 significant bits), 3 MMAs per
                              * product like VQW_MODE_BF16X3, gradients carried with the power-of-two
                              * scale of VQW_MODE_FP16.  Because an fp16 hi plane alone already holds
                              * 11 bits, the weight-gradient GEMMs -- one contraction over time each,
                              * no accumulation of rounding over the depth of the stack -- may run on
                              * the hi planes only (VQW_WGRAD_PASSES=1|2|3) */

typedef struct {
  int B, T;
  int Cr, Cd, Cs, Cc;
  int fs, dilation;
  int skip_accumulate;   /* 0: skip = value (first block), 1: skip += value (modules.py:92-95) */
  int write_residual;    /* 0: the last block's residual is unused (modules.py:52,91) */
  int mode;
} vqw_resblock_desc;

typedef struct {
  const float *conv_w, *conv_b, *cond_w, *cond_b, *res_w, *res_b, *skip_w, *skip_b;
} vqw_resblock_weights;

/* x (B,Cr,T), cond (B,Cc,T) -> residual (B,Cr,T) [may alias nothing], skip (B,Cs,T) in/out,
 * gate_tanh / gate_sig (B,Cd/2,T) saved for backward (both NULL for inference). */
int vqw_resblock_forward(const vqw_resblock_desc* desc, const float* x, const float* cond,
                         const vqw_resblock_weights* w, float* residual, float* skip,
                         float* gate_tanh, float* gate_sig, vqw_stream_t stream);

typedef struct {
  float *conv_w, *conv_b, *cond_w, *cond_b, *res_w, *res_b, *skip_w, *skip_b;
} vqw_resblock_wgrads;

/* Backward of the block.  g_res (B,Cr,T) may be NULL (last block),
 * g_skip (B,Cs,T).  gx (B,Cr,T) overwritten; gcond (B,Cc,T) accumulated; weight grads
 * accumulated.  workspace: vqw_resblock_backward_workspace() bytes, 256-byte aligned. */
int64_t vqw_resblock_backward_workspace(const vqw_resblock_desc* desc);
int vqw_resblock_backward(const vqw_resblock_desc* desc, const float* g_res, const float* g_skip,
                          const float* x, const float* cond, const float* gate_tanh,
                          const float* gate_sig, const vqw_resblock_weights* w, float* gx,
                          float* gcond, const vqw_resblock_wgrads* gw, void* workspace,
                          vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * The whole residual stack.  Replaces ResidualNet.__call__, modules.py:89-96: n_blocks fused
 * block kernels chained on the device, skip summed in place.  This is the entry point the
 * tensor-core modes need: between blocks the activations stay in the packed bf16 hi/lo
 * (B,T,C) layout the MMA consumes, and each block's weights are packed once per call.
 *
 * Host arrays (length n_blocks): dilations, weights (structs of device pointers),
 * residuals[i] = (B,Cr,T) fp32 output of block i (the saved input of block i+1; required in
 * the tensor-core modes for i < n_blocks-1, optional -- ping-pong in the workspace -- in fp32
 * mode; the last entry is only written when keep_last_residual), gate_tanh[i] / gate_sig[i] =
 * B*(Cd/2)*T fp32 saved gate factors (arrays may be NULL for inference): opaque to the caller --
 * fp32 mode lays them out (B,Cd/2,T), the tensor-core modes time-major (B,T,Cd/2), because there
 * one thread owns one time row in the forward gate epilogue and in the backward that reads them.
 * VQW_MODE_BF16X3 only writes gate_sig[i] (gate_tanh[i] may be NULL): the backward recovers
 * tanh = z / sigmoid from the z planes in `saved`.
 *
 * Hoisted global condition (modules.py:17-18,44 with net.py:59-63): the
 * last Cg of the Cc condition channels may be CONSTANT over time -- the speaker embedding that
 * ConditionEmbed broadcasts (net.py:60-61).  With Cg > 0 (tensor-core modes only) `cond` holds
 * only the Cc-Cg time-varying channels, (B,Cc-Cg,T), and `cond_global` (B,Cg) the constant ones:
 * W_p[:, Cc-Cg:] . g is then one bias vector per (item, block) instead of Cg contraction columns
 * at every time step, and its gradient a column sum.  The backward writes gcond (B,Cc-Cg,T) and
 * ACCUMULATES g_cond_global (B,Cg).  Cg = 0: `cond` carries all Cc channels (the reference form).
 */
typedef struct {
  int B, T, Cr, Cd, Cs, Cc, fs;
  int n_blocks;
  const int* dilations;
  int mode;
  int keep_last_residual;
  int Cg;                       /* trailing time-constant condition channels, 0 = none */
  const float* cond_global;     /* (B,Cg) f32 or NULL */
  float* g_cond_global;         /* backward: (B,Cg) f32, accumulated, or NULL */
  /* backward, tensor-core modes: n_blocks cudaEvent_t handles or NULL.  block_events[i] is recorded
   * on the stream once EVERY weight / bias gradient of block i is final (blocks complete from
   * n_blocks-1 down to 0), so that a data-parallel caller can start reducing the gradients of
   * the blocks already done on another stream while the rest of the backward is still running */
  void* const* block_events;
} vqw_resnet_desc;

int64_t vqw_resnet_forward_workspace(const vqw_resnet_desc* desc);
/* Bytes of the caller-owned `saved` buffer (0 in fp32 mode).  Training in the tensor-core modes
 * passes it to the forward, which keeps there the time-major bf16 hi/lo planes of the condition
 * and of every block's input and gated activation, and hands it unchanged to the backward;
 * residuals[] (fp32) is then only needed for the entry the caller wants back (keep_last_residual).
 * saved = NULL = inference. */
int64_t vqw_resnet_saved_bytes(const vqw_resnet_desc* desc);
int vqw_resnet_forward(const vqw_resnet_desc* desc, const float* x, const float* cond,
                       const vqw_resblock_weights* weights, float* const* residuals, float* skip,
                       float* const* gate_tanh, float* const* gate_sig, void* workspace,
                       void* saved, vqw_stream_t stream);

/* Backward of the whole stack, blocks walked in reverse with the shared
 * g_skip (B,Cs,T) and the accumulated g_condition.  x = the stack input, residuals[i] /
 * gate_tanh[i] / gate_sig[i] = what vqw_resnet_forward saved (fp32 mode reads x, cond and
 * residuals[]; the tensor-core modes read `saved` instead and ignore them).  g_last_res
 * (B,Cr,T) or NULL.
 * gx (B,Cr,T) overwritten (may be NULL); gcond (B,Cc,T) and every weight gradient ACCUMULATED.
 * fp32 mode composes the CUDA-core conv family; the tensor-core modes run tcgen05 GEMMs for
 * the gate-gradient, data-gradient and weight-gradient contractions. */
int64_t vqw_resnet_backward_workspace(const vqw_resnet_desc* desc);
int vqw_resnet_backward(const vqw_resnet_desc* desc, const float* g_skip, const float* g_last_res,
                        const float* x, const float* cond, float* const* residuals,
                        float* const* gate_tanh, float* const* gate_sig,
                        const vqw_resblock_weights* weights, float* gx, float* gcond,
                        const vqw_resblock_wgrads* wgrads, void* workspace, const void* saved,
                        vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * WaveNet output head on the tensor cores.  Replaces relu -> proj1 -> relu -> proj2 of
 * WaveNet.__call__, modules.py:155-159 (two 1x1 convolutions) and their backward.
 *   skip (B,Cs,T) f32, W1 (Cs,Cs), b1 (Cs), W2 (Q,Cs), b2 (Q)  ->  y (B,Q,T) f32.
 * mode = VQW_MODE_BF16X3 / VQW_MODE_BF16 / VQW_MODE_FP16 / VQW_MODE_FP16X3; needs Cs % 256 == 0, T >= 128, T % 8 == 0 (other
 * shapes and fp32 go through vqw_conv_forward).  `saved` (vqw_head_saved_bytes) keeps the
 * bf16 planes of relu(skip) and of the hidden activation for the backward; NULL = inference.
 * Backward: gy (B,Q,T) -> gskip (B,Cs,T) overwritten; gW1, gb1, gW2, gb2 ACCUMULATED. */
typedef struct {
  int B, T, Cs, Q;
  int mode;
} vqw_head_desc;
int64_t vqw_head_workspace(const vqw_head_desc* desc);
int64_t vqw_head_saved_bytes(const vqw_head_desc* desc);
int vqw_head_forward(const vqw_head_desc* desc, const float* skip, const float* W1, const float* b1,
                     const float* W2, const float* b2, float* y, void* workspace, void* saved,
                     vqw_stream_t stream);
int vqw_head_backward(const vqw_head_desc* desc, const float* gy, const float* W1, const float* W2,
                      float* gskip, float* gW1, float* gb1, float* gW2, float* gb2, void* workspace,
                      const void* saved, vqw_stream_t stream);
/* Head and loss in one pass (modules.py:155-160 + train.py:92-95 /
 * modules.py:169-230).  Q <= 256 output channels are ONE accumulator tile, so the epilogue of the
 * proj2 GEMM sees a whole row of logits: it reduces the loss and writes d loss / d y straight into
 * the bf16 hi/lo planes the backward GEMMs consume (kept in `saved`) -- the (B,Q,T) logits and
 * their gradient never go to HBM (y_opt != NULL also stores the logits).  Exactly one target:
 *   t_labels (B,T) i32 : softmax cross entropy, mean over the labels in [0,Q) (ignore_label -1)
 *   t_values (B,T) f32 : discretised mixture of logistics on Q = 3*n_mix <= 32 outputs
 * loss = 2 doubles {loss (accumulated: zero both first), number of positions averaged over}.
 * The backward takes the upstream gradient of the scalar loss from DEVICE memory (g_loss, 1 f32).
 * Weight / bias gradients are ACCUMULATED, gskip (B,Cs,T) is overwritten.  bf16 modes only. */
int vqw_head_loss_forward(const vqw_head_desc* desc, const float* skip, const float* W1,
                          const float* b1, const float* W2, const float* b2, const int32_t* t_labels,
                          const float* t_values, int quantize, float log_scale_min, double* loss,
                          float* y_opt, void* workspace, void* saved, vqw_stream_t stream);
int vqw_head_loss_backward(const vqw_head_desc* desc, const float* g_loss, const float* W1,
                           const float* W2, float* gskip, float* gW1, float* gb1, float* gW2,
                           float* gb2, void* workspace, const void* saved, vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Causal embedding of mu-law indices.  Replaces WaveNet.__call__'s embed conv over a one-hot
 * tensor, modules.py:151-152: out[b,c,t] = b[c] + W[c,q[t-1],0] + W[c,q[t],1] (t-1<0 dropped).
 *   q (B,T) i32 in [0,Q), W (Cr,Q,2), bias (Cr) -> out (B,Cr,T). */
int vqw_embed_gather_forward(const int32_t* q, const float* W, const float* bias, float* out,
                             int B, int T, int Cr, int Q, vqw_stream_t stream);
/* gW (Cr,Q,2) and gb (Cr) accumulated. */
int vqw_embed_gather_backward(const int32_t* q, const float* gout, float* gW, float* gb, int B,
                              int T, int Cr, int Q, vqw_stream_t stream);
/* Same gradient as two K=time tcgen05 GEMMs of the gradient planes against a one-hot plane
 * (mode = VQW_MODE_BF16X3 / VQW_MODE_BF16 / VQW_MODE_FP16 / VQW_MODE_FP16X3; needs Cr % 64 == 0, T >= 128, T % 8 == 0; the
 * workspace query returns -1 for unsupported shapes).  gW, gb accumulated. */
int64_t vqw_embed_gather_backward_tc_workspace(int B, int T, int Cr, int Q);
int vqw_embed_gather_backward_tc(const int32_t* q, const float* gout, float* gW, float* gb, int B,
                                 int T, int Cr, int Q, int mode, void* workspace,
                                 vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Loss and optimiser passes over flat fp32 ranges (one HBM-bound kernel each).
 *   vqw_softmax_ce  chainer.functions.softmax_cross_entropy (train.py:95): y (B,Q,T) f32 logits,
 *                   t (B,T) i32 labels -> loss[0] (f64, ACCUMULATED: zero both first) = mean NLL
 *                   over the VALID labels, loss[1] = their count (labels outside [0,Q), e.g.
 *                   Chainer's ignore_label -1, add nothing, get a zero gradient and are not
 *                   counted: normalize=True), gy (B,Q,T) = d loss / d y (or NULL).
 *   vqw_adam_step   chainer.optimizers.Adam's rule (train.py:101): m += (1-b1)(g-m);
 *                   v += (1-b2)(g*g-v); p -= lr*m/(sqrt(v)+eps); lr already bias corrected.
 *                   Hyper-parameters are doubles: 1-beta is formed in double and rounded to
 *                   float32 once, as NumPy does with the reference's Python-float scalars.
 *   vqw_ema_update  ExponentialMovingAverage, utils.py:153-154: ema = decay*target + (1-decay)*ema.
 */
int vqw_softmax_ce(const float* y, const int32_t* t, float* gy, double* loss, int B, int Q, int T,
                   vqw_stream_t stream);
/* WaveNet.calculate_logistic_loss, modules.py:169-230 (discretised mixture of logistics), and
 * its gradient: y (B,3*n_mix,T) f32 = {logit_probs, means, log_scales}, t (B,T) f32 in [-1,1]
 * -> *loss (f64, ACCUMULATED: zero it first) = -mean logsumexp, gy (B,3*n_mix,T) or NULL. */
int vqw_mol_loss(const float* y, const float* t, float* gy, double* loss, int B, int n_mix, int T,
                 int quantize, float log_scale_min, vqw_stream_t stream);
/* Tail of ConditionEmbed.__call__, net.py:58-63: out (B,Cl+Cg,T_out) = concat(align-corners
 * linear resize of local (B,Cl,H) to T_out steps [F.resize_images, exact integer coordinates],
 * glob (B,Cg) broadcast over time).  Backward: g (B,Cl+Cg,T_out) -> g_local (B,Cl,H), g_glob
 * (B,Cg), both overwritten. */
int vqw_upsample_concat_forward(const float* local, const float* glob, float* out, int B, int Cl,
                                int Cg, int H, int T_out, vqw_stream_t stream);
int vqw_upsample_concat_backward(const float* g, float* g_local, float* g_glob, int B, int Cl, int Cg,
                                 int H, int T_out, vqw_stream_t stream);
int vqw_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                  double beta2, double eps, vqw_stream_t stream);
/* same, with the (bias-corrected) learning rate read from device memory: the form a CUDA-graph
 * replay of the training step uses, where lr changes every step but kernel arguments cannot */
int vqw_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev,
                      double beta1, double beta2, double eps, vqw_stream_t stream);
int vqw_ema_update(float* ema, const float* target, long long n, double decay, vqw_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Persistent autoregressive generation.  Replaces the sample loop of generate.py:109-145 and
 * the queue-based incremental decoder it drives (WaveNet.initialize / generate,
 * modules.py:232-255; ResidualNet.generate :102-110; ResidualBlock.push / pop :58-74) with ONE
 * cooperative kernel launch for `n_steps` samples of one utterance (n = 1, generate.py:42).
 *
 *   cond      (Cc, T_total) f32   full-rate condition (ConditionEmbed output, net.py:48-64)
 *   uniforms  (n_steps) f64       the draws numpy.random.choice would make (generate.py:139)
 *   forced    (n_steps) i32/NULL  teacher forcing: value fed back instead of the drawn sample
 *   samples   (n_steps) i32       drawn indices (generate.py:145)
 *   logits    (n_steps, Q) f32    decoder outputs per step, or NULL
 * t_start = 0 zero-initialises the dilation queues (initialize()); a later call with
 * t_start = previous t_start + n_steps continues the same utterance from the workspace.
 * set_state != 0 overrides the two most recent inputs (s1 = input of step t_start, s2 = the
 * one before; -1 = the all-zero vector of generate.py:51) before the first step.
 */
typedef struct {
  int n_blocks;
  const int* dilations;
  int fs, Cr, Cd, Cs, Cc, Q;
  int T_total, n_steps, t_start;
  int set_state, s1, s2;
  int cond_t0;   /* time index of cond's first column (0 for a whole-utterance condition) */
  /* mixture-of-logistics decoder (generate.py:116-137): embed_w is (Cr,1,2), Q = 3*nr_mix output
   * channels {logit_probs, means, log_scales}; uniforms is (n_steps, nr_mix) f64; the value fed
   * back is sum_k softmax_k * (mean_k + exp(max(log_scale_k, log_scale_min)) * logit(u_k)) / 127.5
   * clipped to [-1, 1]; samples / forced / s1 / s2 then hold float32 values as their bit
   * patterns (initial state = 0.0). */
  int use_logistic;
  float log_scale_min;
} vqw_generate_desc;

int64_t vqw_generate_workspace(const vqw_generate_desc* desc);
int vqw_generate(const vqw_generate_desc* desc, const vqw_resblock_weights* blocks,
                 const float* embed_w, const float* embed_b, const float* proj1_w,
                 const float* proj1_b, const float* proj2_w, const float* proj2_b,
                 const float* cond, const double* uniforms, const int32_t* forced,
                 int32_t* samples, float* logits, void* workspace, vqw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQW_H_ */
