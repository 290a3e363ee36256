/* smct.h — C-ABI of libsmct_b200.so: the SMC-Transformer particle-filter forward pass on B200.
 *
 * The reference (AMDonati/SMC-T-v2, /root/reference) has no FFI: its boundary is the Python/Keras
 * call surface.  Each entry point below therefore names the reference *Python* interface it
 * replaces (file:line relative to /root/reference); the Python mirror of that surface lives in
 * smc-t-v2_b200/models/ and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns 0 (SMCT_OK) or a negative error code and
 *     leaves a message retrievable with smct_last_error() (thread-local);
 *   - all data pointers are DEVICE pointers (fp32 / fp64 / int32 as documented), contiguous,
 *     owned by the caller; the library never allocates caller-visible memory — scratch comes
 *     from a caller-provided workspace whose size is queried first;
 *   - every call is asynchronous on the given CUDA stream (a cudaStream_t passed as void*);
 *   - dense kernels use the Keras layout W[in][out] (y = x·W + b);
 *   - tensors use the reference layouts: state K,V,R (B,P,S,D); predictions (B,P,S,F_y);
 *     weights (B,P,S); per-step indices (S,B,P).
 *   - there is NO CPU fallback: without a CUDA device every compute call returns SMCT_ERR_CUDA.
 */
#ifndef SMCT_H_
#define SMCT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCT_ABI_VERSION 3

enum {
  SMCT_OK = 0,
  SMCT_ERR_INVALID = -1,     /* bad argument / shape */
  SMCT_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  SMCT_ERR_WORKSPACE = -3,   /* workspace too small */
  SMCT_ERR_UNSUPPORTED = -4  /* shape outside the supported envelope */
};

enum { SMCT_WEIGHTS_GAUSSIAN = 0, SMCT_WEIGHTS_CLASSIFICATION = 1 };
enum { SMCT_NOISE_Z_CHECKOUT = 0, SMCT_NOISE_Z_PYC = 1 };
enum { SMCT_INPUT_LINEAR = 0, SMCT_INPUT_EMBEDDING = 1, SMCT_INPUT_ENCODED = 2 };
/* FUSED: persistent kernel (two CTAs per SM, mbarrier-sequenced dense part on tcgen05, bulk-copied weights);
 * FUSED_V1: the first-generation persistent kernel (one 512-thread CTA per SM; kept for F_y > 4 and as a cross-check);
 * FUSED_FFMA: first-generation structure, dense part on CUDA cores (pure fp32 cross-check);
 * STAGED: one launch per stage and timestep, any shape; STAGED_SEQ: the staged device code in ONE launch, one thread-
 * block cluster per sequence; SMALL: the reference's own small shapes (d_model 8/16/32, gaussian): one CTA per sequence,
 * a particle's rows in the registers of d_model/8 lanes */
enum { SMCT_ALGO_AUTO = 0, SMCT_ALGO_STAGED = 1, SMCT_ALGO_FUSED = 2, SMCT_ALGO_FUSED_FFMA = 3, SMCT_ALGO_STAGED_SEQ = 4,
       SMCT_ALGO_FUSED_V1 = 5, SMCT_ALGO_SMALL = 6 };

/* Problem shape + switches.  Mirrors the constructor arguments of
 * SMC_Transformer(d_model, output_size, seq_len, full_model, dff, num_heads, attn_window)
 * (src/models/SMC_Transformer/SMC_Transformer.py:21-22) and the cell state installed by
 * add_SMC_parameters / add_stop_resampling (SMC_TransformerCell.py:72-76,125-127). */
typedef struct smct_shape {
  int32_t struct_bytes;   /* = sizeof(smct_shape); checked */
  int32_t B, P, S, D, H, dff;
  int32_t F_x;            /* input features (linear) or vocabulary (embedding) */
  int32_t F_y;            /* target features (gaussian) or classes (classification) */
  int32_t full_model;     /* 1: LN1(z+x) -> FFN -> LN2 ; 0: r = z */
  int32_t attn_window;    /* -1: none */
  int32_t len_resampling; /* -1: resample at every step */
  int32_t weights_mode;   /* SMCT_WEIGHTS_* */
  int32_t noise_z_mode;   /* SMCT_NOISE_Z_* */
  int32_t input_mode;     /* SMCT_INPUT_* */
  int32_t noise_on;       /* cell.noise */
  int32_t logvar_dim;     /* 1: scalar log-variances ("uni"); D: per-dimension ("multi") */
  int32_t algo;           /* SMCT_ALGO_* (forward only) */
  float ln_eps;           /* 1e-6 (SMC_TransformerCell.py:36) */
  float sigma_obs;        /* observation VARIANCE of the Gaussian weight */
  int64_t b_global0;      /* global index of sequence 0 of this shard (RNG is keyed by it) */
} smct_shape;

/* Device pointers to the model parameters (fp32).  Names follow the reference layers:
 * wq/wk/wv/dense (self_attention_SMC.py:21-24), layernorm1/2, ffn, output_layer
 * (SMC_TransformerCell.py:36-40), embedding / projection_layer_ts (SMC_Transformer.py:56). */
typedef struct smct_params {
  const float *Wq, *bq, *Wk, *bk, *Wv, *bv, *Wz, *bz; /* (D,D),(D) */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;         /* (D) */
  const float *W1, *b1, *W2, *b2;                     /* (D,dff),(dff),(dff,D),(D) */
  const float *Wo;                                    /* (D,F_y), no bias */
  const float *W_in, *b_in;                           /* (F_x,D),(D): projection or embedding table */
  const float *logvar_k, *logvar_q, *logvar_v, *logvar_z; /* (logvar_dim) */
} smct_params;

/* Device pointers to gradient buffers, one per parameter (same names/shapes as smct_params; fp32).  smct_backward
 * ACCUMULATES into them (zero them before the first shard/chunk of a step); NULL entries are skipped. */
typedef struct smct_grads {
  float *Wq, *bq, *Wk, *bk, *Wv, *bv, *Wz, *bz;
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  float *W1, *b1, *W2, *b2;
  float *Wo;
  float *W_in, *b_in;
  float *logvar_k, *logvar_q, *logvar_v, *logvar_z;
} smct_grads;

enum { SMCT_LOSS_PYC = 0, SMCT_LOSS_CHECKOUT = 1 };

/* Inputs / outputs of one forward pass.  Optional pointers may be NULL. */
typedef struct smct_io {
  const void *x_in;        /* (B,S,F_x) f32 | (B,S) i32 tokens | (B,S,D) f32 encoded */
  const void *y;           /* (B,S,F_y) f32 (gaussian) | (B,S) i32 (classification) */
  const float *eps;        /* optional (4,S,B,P,D) injected N(0,1) in order k,q,v,z; NULL -> Philox(seed) */
  const double *u;         /* optional (S,B,P) injected U[0,1); NULL -> Philox(seed) */
  const int32_t *i_forced; /* optional (S,B,P) ancestor indices replacing the draw (teacher forcing) */
  uint64_t seed;
  float *pred;             /* (B,P,S,F_y)  output_layer(R unresampled) */
  float *pred_resampl;     /* (B,P,S,F_y)  output_layer(R resampled) */
  float *K, *V, *R;        /* (B,P,S,D)    final resampled states; 16-byte aligned (128-bit stores / bulk copies) */
  float *w;                /* (B,P,S)      filtering weights (gaussian: normalised; classification: raw) */
  int32_t *i_out;          /* optional (S,B,P) drawn ancestor indices */
  float *logw;             /* optional (S,B,P) unnormalised log-weights (gaussian) / w (classification) */
  float *noise_q, *noise_z;/* optional (B,P,S,D) */
  float *noise_K, *noise_V;/* optional (B,P,S,D) K - wk(X), V - wv(X) (SMC_Transformer.py:235-236) */
  float *attn;             /* optional (B,P,H,S,S) */
  double *loss_sums;       /* optional (8): sum n^2/sigma^2 for K,q,V,z; sum (y-pred_resampl)^2; 3 spare */
  const float *classic_w;  /* optional (B,S), smct_backward only: weight of the classic-loss term of row (b,s) relative to the
                            * uniform 1/(B*P*S) mean — the attention_mask weighting of compute_classic_loss
                            * (SMC_Transformer.py:158-163): mask[b,s] * (global B*S) / (global sum of the mask) */
} smct_io;

int smct_abi_version(void);
const char *smct_last_error(void);
const char *smct_build_info(void);
/* First 32 hex digits of the sha256 over the sources (csrc/ and this header) the binary was compiled from; the binding
 * refuses a library whose hash differs from the tree it is loaded from (smc-t-v2_b200/build.py:source_hash). */
const char *smct_source_hash(void);
/* Number of CUDA kernels this library has launched in the process (reset!=0: return and zero). */
long long smct_launch_count(int reset);

/* Workspace (bytes, device memory) needed by smct_forward / smct_cell_step for this shape. */
size_t smct_workspace_bytes(const smct_shape *shape);

/* SMC_Transformer.call(inputs, targets)  — src/models/SMC_Transformer/SMC_Transformer.py:183-248
 * (encode, zero state, S cell steps, output heads, resampled noises).                           */
int smct_forward(const smct_shape *shape, const smct_params *params, const smct_io *io,
                 void *workspace, size_t workspace_bytes, void *stream);

/* tape.gradient(loss, smc_transformer.trainable_variables) of train_step_SMC_T — src/train/train_step_functions.py:43-63
 * for the regression loss of compute_SMC_loss (SMC_Transformer.py:123-146; regression-era form recovered from the pyc):
 *   loss = inv_n * sum_{b,p,s} [ 1/2 sum_n ||noise_n||^2/sigma_n^2 + 1/2 ||y - pred_resampl||^2/Sigma_obs ]
 *   (+ the log-determinant terms when loss_variant == SMCT_LOSS_CHECKOUT; they only touch the log-variances).
 * Driven by the OUTPUTS of smct_forward on the same shape/params/inputs: io->K, io->V (final states), io->i_out
 * (ancestor indices), io->loss_sums (for the log-variance gradients), io->x_in, io->y and the same io->eps or
 * io->seed.  inv_n = 1 / (global B*P*S) so that chunks / shards of one optimisation step can be accumulated.
 * Supported: gaussian (MSE) and classification (sparse CE) observation models, 1 head, full model, scalar or
 * per-dimension log-variances (the latter need io->noise_K, noise_q, noise_V, noise_z: the forward's recorded noises),
 * noise_z_mode = SMCT_NOISE_Z_PYC (the checkout quirk noise_z = z - z_ is forward-only: its loss term involves particles
 * outside the final genealogy),
 * d_model <= 128, S*d_model <= 12800.  Gradients are ACCUMULATED into `grads`. */
size_t smct_backward_workspace_bytes(const smct_shape *shape);
int smct_backward(const smct_shape *shape, const smct_params *params, const smct_io *io, const smct_grads *grads,
                  float inv_n, int32_t loss_variant, void *workspace, size_t workspace_bytes, void *stream);

/* optimizer.apply_gradients of tf.keras.optimizers.Adam (run_SMC_T.py:18-22): one parameter tensor of n elements.
 * lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) is computed by the caller. */
int smct_adam_step(int64_t n, float *param, const float *grad, float *m, float *v, float lr_t, float beta1,
                   float beta2, float epsilon, void *stream);

/* SMC_Transf_Cell.call(inputs, states) — src/models/SMC_Transformer/SMC_TransformerCell.py:129-184
 * One timestep t on explicit states.  x (B,P,D) encoded input of this step (x_per_particle=1) or
 * (B,D) shared by the particles (0); y (B,F_y) f32 | (B) i32; eps4 (4,B,P,D) or NULL; u (B,P) or NULL.
 * K,V,R (B,P,S,D) are updated in place (slot t written, rows 0..t gathered).
 * Outputs (optional): r (B,P,D), pred (B,P,F_y), attn (B,P,H,S), noise_q/noise_z (B,P,D),
 * w (B,P), i_out (B,P).                                                                         */
int smct_cell_step(const smct_shape *shape, const smct_params *params, int32_t t,
                   const float *x, int32_t x_per_particle, const void *y,
                   const float *eps4, const double *u, const int32_t *i_forced, uint64_t seed,
                   float *K, float *V, float *R,
                   float *r, float *pred, float *attn, float *noise_q, float *noise_z,
                   float *w, int32_t *i_out,
                   void *workspace, size_t workspace_bytes, void *stream);

/* Self_Attention_SMC.call(inputs, timestep, K, V) — src/models/SMC_Transformer/self_attention_SMC.py:140-193
 * inputs (B,P,D) (already layer-normed by the caller, as in the cell); writes k_t,v_t into slot t of
 * K,V (B,P,S,D) and returns z (B,P,D), attention weights (B,P,H,S) and the recorded noises
 * noise_k/q/v/z (B,P,D) (attributes self.noise_* of the layer).  eps4 (4,B,P,D) or NULL -> Philox. */
int smct_attention_step(const smct_shape *shape, const smct_params *params, int32_t t,
                        const float *inputs, const float *eps4, uint64_t seed,
                        float *K, float *V, float *z, float *attn,
                        float *noise_k, float *noise_q, float *noise_v, float *noise_z,
                        void *workspace, size_t workspace_bytes, void *stream);

/* compute_w_regression / compute_w_classification — SMC_TransformerCell.py:78-84 (+ :206 call site).
 * gaussian: logw = -0.5*||y-pred||^2/sigma_obs (B,P); classification: softmax(pred)[y] (B,P).      */
int smct_logweights(int32_t B, int32_t P, int32_t F_y, int32_t weights_mode, float sigma_obs,
                    const float *pred, const void *y, float *logw, void *stream);

/* tf.random.categorical(w, P) call site — SMC_TransformerCell.py:166 (TF CPU inverse-CDF algorithm).
 * normalise=1: logits = logw - logsumexp_p(logw) and w_out = exp(logits) (gaussian path);
 * normalise=0: logits = logw as given (the checkout passes probabilities as logits), w_out = logw.
 * u (B,P) fp64 uniforms.  i_out (B,P).  w_out/logits_out optional.                               */
int smct_resample(int32_t B, int32_t P, int32_t normalise, const float *logw, const double *u,
                  float *w_out, float *logits_out, int32_t *i_out, void *stream);

/* resample(params, i_t) = tf.gather(params, i_t, axis=1, batch_dims=1)
 * — src/models/SMC_Transformer/transformer_utils.py:39-47.  src,dst (B,P,row_elems) distinct buffers;
 * copies the first n_elems (<= row_elems) elements of each particle row: dst[b,m,:n] = src[b,idx[b,m],:n]. */
int smct_gather(int32_t B, int32_t P, int64_t row_elems, int64_t n_elems, const float *src,
                const int32_t *idx, float *dst, void *stream);

/* tf.random.normal / uniform replacements used when eps/u are not injected (self_attention_SMC.py:95).
 * Fills out (B,P,D) with the N(0,1) stream `which` (0=k,1=q,2=v,3=z) of timestep t, resp. (B,P) fp64 U[0,1). */
int smct_fill_noise(uint64_t seed, int32_t which, int32_t t, int64_t b_global0, int32_t B, int32_t P,
                    int32_t D, float *out, void *stream);
int smct_fill_uniform(uint64_t seed, int32_t t, int64_t b_global0, int32_t B, int32_t P, double *out,
                      void *stream);

/* compute_SMC_loss reductions — SMC_Transformer.py:91-99,123-164.
 * out[0] += sum_{rows,d} n[row,d]^2 * exp(-logvar[d or 0])   (fp64 accumulation). */
int smct_sumsq_scaled(int64_t rows, int32_t D, const float *n, const float *logvar, int32_t logvar_dim,
                      double *out, void *stream);

/* SMC_Transformer.get_encoded_input — SMC_Transformer.py:166-181 (decoder is None): embedding / projection
 * times sqrt(d_model), one row per (b,s) (the particles of a sequence share it).  X (B,S,D) and/or
 * X_ln = layernorm1(X) (B,S,D); either may be NULL. */
int smct_encode(const smct_shape *shape, const smct_params *params, const void *x_in, float *X, float *X_ln,
                void *stream);

/* Self-test of the tcgen05 building blocks used by the fused forward's dense part: D (128,64) = A (128,64) · W (64,64)
 * (W in Keras layout [in][out]) computed as a three-term fp16 split (exact power-of-two scaling) on the tensor cores with fp32 TMEM
 * accumulation.  scratch: >= 16 KB + 256 B of device memory. */
int smct_tc_selftest(const float *A, const float *W, float *D, void *scratch, void *stream);

/* compute_classic_loss — SMC_Transformer.py:148-164 (sparse CE, checkout) / regression-era MSE form (pyc).
 * out[0] += sum_{b,p,s} m[b,s] * ||y[b,s]-pred[b,p,s]||^2 (gaussian; y (B,S,F_y) f32)
 *        or sum_{b,p,s} m[b,s] * CE(pred[b,p,s,:], y[b,s]) (classification; y (B,S) i32).  pred (B,P,S,F_y).
 * mask (B,S) f32 is the attention_mask of the reference's masked mean (:158-163); NULL = all ones. */
int smct_classic_loss(int32_t B, int32_t P, int32_t S, int32_t F_y, int32_t weights_mode, const float *pred,
                      const void *y, const float *mask, double *out, void *stream);

/* The training metric of train_step_SMC_T on the unresampled predictions pred (B,P,S,F_y):
 * classification — compute_categorical_cross_entropy, src/train/utils.py:3-20: sparse CE of the particle-mean softmax;
 * gaussian — regression era (pyc): squared error of the particle-mean prediction, summed over F_y.
 * out[0] += sum_{b,s} m[b,s] * metric[b,s], out[1] += sum_{b,s} m[b,s]; mask (B,S) f32 or NULL (:15-19). */
int smct_metric(int32_t B, int32_t P, int32_t S, int32_t F_y, int32_t weights_mode, const float *pred, const void *y,
                const float *mask, double *out, void *stream);

/* EM(smc_transformer, it, EM_param) — src/train/utils.py:22-51: per batch element j (in order) every scalar
 * log-variance is replaced by log((1 - it^-a) exp(logvar) + it^-a mean(noise[j]^2)) (EM_update, :48-51), for the four
 * recorded noises noise_K_resampled, noise_q, noise_V_resampled, noise_z, each (B, per_b) with per_b = P*S*D.
 * The log-variances (device scalars) are updated in place; scratch: >= 4*B doubles of device memory. */
int smct_em_update(int32_t B, int64_t per_b, const float *noise_k, const float *noise_q, const float *noise_v,
                   const float *noise_z, float *logvar_k, float *logvar_q, float *logvar_v, float *logvar_z, float it,
                   float em_param, void *scratch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SMCT_H_ */
