/* vae_tagger_b200 -- C ABI of the B200-native encode+tag hot path.
 *
 * The reference (spawner1145/vae-tagger) has no FFI: its boundary for this path is the
 * Python module surface.  This header is the C-ABI that surface is backed by
 * in this repo; every entry point names the reference interface it replaces.  The Python
 * side (vae_tagger_b200/_native.py) binds it with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; vt_last_error() then
 *     returns a thread-local, NUL-terminated description.  No exceptions cross the ABI.
 *   - plain pointers and sizes only.  "device pointer" = CUDA device memory of the
 *     context's device, caller-owned.  Workspace is owned by the context.
 *   - all work is ordered on the caller's stream (a cudaStream_t passed as void*; NULL = legacy
 *     default stream): it starts after everything already enqueued there and later work on that
 *     stream waits for it.  vt_encode may fan micro-batches out to two internal streams (joined
 *     back with events).  No host synchronisation except where stated (host variants).
 *   - one context per device / per rank; a context is not thread-safe.
 *   - tensors are fp32, PyTorch layouts (NCHW activations, OIHW conv weights, [out][in]
 *     linear weights) at the boundary; the NHWC bf16 layout used internally never leaks.
 */
#ifndef VAE_TAGGER_B200_H
#define VAE_TAGGER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_ABI_VERSION 1

typedef struct vt_ctx vt_ctx;

/* precision of the encoder contraction path */
#define VT_PREC_BF16 0 /* tcgen05 implicit GEMM, 16-bit operands (bf16 raw / fp16 bounded), fp32 accumulate (TMEM) */
#define VT_PREC_FP32 1 /* FFMA verification mode (north star "fp32 mode", rel L2 <= 1e-4) */
#define VT_PREC_F16 2  /* vt_op_* only: fp16 operands (what the encoder schedule uses for bounded operands) */

/* image input formats of vt_encode */
#define VT_IN_F32_NCHW 0 /* [B,3,H,W] fp32 already normalised to [-1,1] (modules.py:136-140) */
#define VT_IN_U8_NHWC 1  /* [B,H,W,3] uint8; (u/255-0.5)/0.5 is fused into the conv_in gather */

const char* vt_last_error(void);
int vt_abi_version(void);

int vt_ctx_create(int device, vt_ctx** out);
int vt_ctx_destroy(vt_ctx* ctx);

/* ---------------------------------------------------------------- encoder (FLUX AutoencoderKL)
 * Replaces: diffusers AutoencoderKL(**config) constructed at diffusers_vae_loader.py:8-35 and
 * its load_state_dict at :44. */
typedef struct vt_encoder_config {
    int in_channels;           /* 3 */
    int num_blocks;            /* 4 */
    int block_out_channels[8]; /* 128,256,512,512 */
    int layers_per_block;      /* 2 */
    int norm_num_groups;       /* 32 */
    int latent_channels;       /* 16 */
    int mid_block_add_attention; /* 1 */
    int has_scaling_factor;    /* hasattr(config,'scaling_factor') diffusers_vae_loader.py:81 */
    int has_shift_factor;      /* :83 */
    float scaling_factor;      /* 0.3611 */
    float shift_factor;        /* 0.1159 */
} vt_encoder_config;

int vt_encoder_configure(vt_ctx* ctx, const vt_encoder_config* cfg);
/* name = diffusers state-dict key below "encoder." (SURVEY.md Appendix B), e.g.
 * "down_blocks.0.resnets.0.conv1.weight"; data = fp32, host or device pointer, PyTorch layout */
int vt_encoder_set_param(vt_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
/* checks that every parameter is present and repacks for the kernels.  Call again after
 * changing parameters. */
int vt_encoder_finalize(vt_ctx* ctx);

typedef struct vt_encode_args {
    const void* images; /* device pointer, format in_fmt */
    int in_fmt;
    int batch, height, width; /* any size >= 8: every Downsample2D halves with floor, the latent is floor(size/8);
                               * multiples of 64 -- the bucket sizes -- tile without ragged edges at every level */
    int precision;            /* VT_PREC_* */
    int sample;               /* 0: latent_dist.mode() (diffusers_vae_loader.py:80); 1: .sample() (:74) */
    int apply_scale_shift;    /* 1: DiffusersVAEWrapper.encode semantics (mode*scale+shift, :80-84) */
    uint64_t seed;            /* sample noise stream when noise == NULL */
    const float* noise;       /* optional device [B,LC,H/8,W/8] standard normal (exact-parity sampling) */
    float* latent;            /* out, device [B,LC,H/8,W/8] fp32 NCHW; may be NULL */
    float* mean;              /* out, optional: DiagonalGaussianDistribution.mean */
    float* logvar;            /* out, optional: clamped logvar */
    int micro_batch;          /* images per internal pass; 0 = library default */
    int single_lane;          /* 1: run micro-batches back to back on the caller's stream (no overlap) */
    void* stream;
} vt_encode_args;

/* Replaces: DiffusersVAEWrapper.encode (diffusers_vae_loader.py:78-86) and
 * AutoencoderKL.encode(x).latent_dist.{mode,sample,mean,logvar} (:73-74, :79-80). */
int vt_encode(vt_ctx* ctx, const vt_encode_args* args);

/* ---------------------------------------------------------------- VAE decoder
 * Replaces: AutoencoderKL.decode(z).sample as called by DiffusersVAEWrapper.decode / .forward
 * (diffusers_vae_loader.py:72-76, :88-94).  Uses the configuration given to vt_encoder_configure.
 * name = diffusers state-dict key below "decoder.", e.g. "up_blocks.0.resnets.0.conv1.weight". */
int vt_decoder_set_param(vt_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
int vt_decoder_finalize(vt_ctx* ctx);
typedef struct vt_decode_args {
    const float* latent; /* device [B,LC,h,w] fp32 NCHW */
    int batch, lat_h, lat_w;
    int precision;         /* VT_PREC_BF16 / VT_PREC_FP32 */
    int apply_scale_shift; /* 1: DiffusersVAEWrapper.decode semantics, (z - shift) / scale first (:88-93) */
    float* image;          /* out, device [B,3,8h,8w] fp32 NCHW */
    int micro_batch;       /* images per internal pass; 0 = library default */
    void* stream;
} vt_decode_args;
int vt_decode(vt_ctx* ctx, const vt_decode_args* args);

/* ---------------------------------------------------------------- tag head
 * Replaces: AttentionClassificationDecoder / ClassificationDecoder (modules.py:303-475). */
#define VT_HEAD_ATTENTION 0 /* AttentionClassificationDecoder */
#define VT_HEAD_PLAIN 1     /* ClassificationDecoder (--no_attention) */
typedef struct vt_head_config {
    int kind;
    int latent_channels; /* 16 */
    int num_classes;
    int use_spatial_attention;
    int use_self_attention;
    int attention_heads; /* 8 */
    int use_cross_attention; /* --use_cross_attention (modules.py:388-395, :450-459): query_generator + CrossAttention */
    int plain_flat_dim;      /* VT_HEAD_PLAIN only: 0 = AdaptiveAvgPool(4,4) (use_adaptive_pooling=True, modules.py:312-314);
                              * else LC*h*w: the latent is flattened as is (use_adaptive_pooling=False, :316-317) */
} vt_head_config;
int vt_head_configure(vt_ctx* ctx, const vt_head_config* cfg);
/* name = reference state-dict key, e.g. "classifier.12.weight" */
int vt_head_set_param(vt_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
int vt_head_finalize(vt_ctx* ctx);

typedef struct vt_tag_args {
    const float* latent; /* device [B,LC,h,w] fp32 NCHW */
    int batch, lat_h, lat_w;
    float threshold;      /* infer_full.py:114 uses conf >= threshold */
    float* logits;        /* out, optional device [B,T]: decoder(latent) (modules.py:424-468) */
    float* probs;         /* out, optional device [B,T]: sigmoid(logits) in tag order */
    float* conf_sorted;   /* out, optional device [B,T]: get_confidence()[0] (modules.py:470-475) */
    int64_t* idx_sorted;  /* out, optional device [B,T]: get_confidence()[1] */
    int32_t* count;       /* out, optional device [B]: number of tags with conf >= threshold */
    void* stream;
} vt_tag_args;
int vt_tag(vt_ctx* ctx, const vt_tag_args* args);

/* ---------------------------------------------------------------- end-to-end, host buffers
 * Replaces the per-image loop of infer_full.py:95-124 for a whole batch: H2D of the images,
 * encode (mode, scale/shift), tag, D2H of the sorted confidences / indices / counts.  Host
 * pointers should be pinned; the call synchronises the stream before returning. */
typedef struct vt_infer_host_args {
    const void* images_host; /* format in_fmt */
    int in_fmt;
    int batch, height, width;
    int precision;
    float threshold;
    float* conf_sorted_host;  /* [B,T] */
    int64_t* idx_sorted_host; /* [B,T] */
    int32_t* count_host;      /* [B] */
    float* latent_host;       /* optional [B,LC,H/8,W/8] */
    int micro_batch;
    void* stream;
} vt_infer_host_args;
int vt_infer_host(vt_ctx* ctx, const vt_infer_host_args* args);

/* Same pipeline on DEVICE buffers (images already in HBM, results left in HBM): encode (mode, scale/shift) and,
 * per internal micro-batch and right behind its encoder, the tag head -- what infer_full.py:95-118 does per
 * image, for a batch, without host round trips.  No host synchronisation. */
typedef struct vt_infer_args {
    const void* images; /* device, format in_fmt */
    int in_fmt;
    int batch, height, width;
    int precision;
    float threshold;
    float* conf_sorted;  /* out, device [B,T] */
    int64_t* idx_sorted; /* out, device [B,T] */
    int32_t* count;      /* out, device [B] */
    float* latent;       /* out, device [B,LC,H/8,W/8] (required: the head reads it) */
    int micro_batch;
    int single_lane;     /* 1: micro-batches back to back on the caller's stream (per-kernel timing runs) */
    void* stream;
} vt_infer_args;
int vt_infer(vt_ctx* ctx, const vt_infer_args* args);

/* ---------------------------------------------------------------- focal loss (training step)
 * Replaces FocalLoss.forward (improved_losses.py:47-56) and its autograd backward:
 * loss_sum += sum(alpha*(1-pt)^gamma*bce); grad = grad_scale * d(sum)/d(logits). */
int vt_focal_loss(vt_ctx* ctx, const float* logits, const float* targets, int64_t n, float alpha, float gamma,
                  float grad_scale, float* loss_sum /* device, 1 float, accumulated */,
                  float* grad /* device [n], optional */, void* stream);

/* ---------------------------------------------------------------- head training step
 * Replaces, for one batch, the autograd graph of the reference training step
 * (train_decoder.py:186-195): decoder.train(); logits = decoder(latent);
 * loss = loss_fn(logits, labels); loss.backward() -- for the head configured with
 * vt_head_configure (AttentionClassificationDecoder without cross-attention, modules.py:358-468, or
 * ClassificationDecoder, modules.py:303-349).  The latent is the frozen encoder's output and gets
 * no gradient.  Parameters are read from, and gradients ACCUMULATED (+=) into, caller-owned flat
 * fp32 device buffers laid out like the reference module's parameters() -- vt_head_param_layout
 * gives the offsets; the same flat gradient buffer is what the data-parallel step all-reduces
 * (train_decoder.py:39 accelerate/DDP).  BatchNorm2d runs on batch statistics and updates the
 * running buffers in place.  nn.BCEWithLogitsLoss is focal_alpha = 1, focal_gamma = 0; ClassBalancedLoss
 * (improved_losses.py:58-72, train_decoder.py:188-189) is that plus class_weights. */
int vt_head_param_count(vt_ctx* ctx, int32_t* n_tensors, int64_t* n_floats);
/* index in [0, n_tensors): state-dict key (copied into name, NUL-terminated), offset and size in floats */
int vt_head_param_layout(vt_ctx* ctx, int32_t index, char* name, int32_t name_cap, int64_t* offset,
                         int64_t* numel);

typedef struct vt_head_train_args {
    const float* latent;  /* device [B,LC,h,w] fp32 NCHW (DiffusersVAEWrapper.encode output) */
    const float* targets; /* device [B,T] multi-hot labels */
    int batch, lat_h, lat_w;
    const float* params; /* device, flat, vt_head_param_layout order */
    float* grads;        /* device, flat, same layout, accumulated; NULL = forward + loss only */
    float* bn_running_mean;          /* device [LC/2], updated in place (feature_compress.1); may be NULL */
    float* bn_running_var;           /* device [LC/2] */
    int64_t* bn_num_batches_tracked; /* device scalar, += 1; may be NULL */
    float bn_momentum;               /* 0.1 */
    float focal_alpha, focal_gamma;  /* FocalLoss (improved_losses.py:47-56), reduction = mean */
    float loss_scale;                /* 1 / gradient_accumulation_steps: scales loss and gradients */
    int dropout;                     /* 1: Dropout layers active (module.train()); 0: identity */
    float attention_dropout;         /* p of MultiHeadSelfAttention.dropout (modules.py:64) */
    uint64_t seed;                   /* dropout mask stream of this step */
    float* loss;   /* device scalar, += loss_scale * mean loss; may be NULL */
    float* logits; /* optional out, device [B,T] */
    void* stream;
    const float* class_weights; /* optional device [T]: ClassBalancedLoss weights (improved_losses.py:66-69), each
                                 * term of the loss is multiplied by its class weight; use focal_gamma = 0 */
} vt_head_train_args;
int vt_head_train_step(vt_ctx* ctx, const vt_head_train_args* args);

/* The dropout keep-masks (0/1 floats) vt_head_train_step uses for (seed, batch): attn [B,heads,64,64],
 * cls[i] [B, width of classifier block i].  For parity tests: the oracle applies the same masks. */
int vt_head_dropout_masks(vt_ctx* ctx, int batch, float attention_dropout, uint64_t seed, float* attn,
                          float* cls0, float* cls1, float* cls2, void* stream);

/* Replaces clip_grad_norm_(max_grad_norm) + torch.optim.AdamW.step() + zero_grad
 * (train_decoder.py:197-203) on flat device buffers: g *= grad_scale (1/world after the all-reduce),
 * clipped to max_norm (<= 0: no clipping), decoupled weight decay, bias correction for `step` (1-based).
 * norm_out (device scalar, optional) receives the gradient norm before clipping. */
int vt_adamw_step(vt_ctx* ctx, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                  float max_norm, int zero_grad, float* norm_out, void* stream);

/* ---------------------------------------------------------------- image preprocessing
 * Replaces, on uint8 HWC images already on the device, what the reference does on the host through
 * Pillow before ToTensor/Normalize: SmartResize.__call__ (modules.py:142-178: crop to the bucket's aspect
 * ratio, then img.resize((W,H), Image.LANCZOS)) and transforms.Resize((res,res)) (modules.py:135, PIL
 * BILINEAR).  Bit-exact with Pillow's 8-bit ImagingResample (fixed-point taps, uint8 intermediate,
 * horizontal then vertical).  The output feeds vt_encode as VT_IN_U8_NHWC, which fuses
 * ToTensor + Normalize(0.5, 0.5) into conv_in. */
#define VT_FILTER_LANCZOS 1 /* PIL.Image.LANCZOS */
#define VT_FILTER_BILINEAR 2 /* PIL.Image.BILINEAR */
typedef struct vt_resize_args {
    const void* src;    /* device, uint8 [src_h][src_w][3], rows src_stride bytes apart */
    int src_w, src_h;
    int64_t src_stride;
    int crop_l, crop_t, crop_r, crop_b; /* img.crop((l,t,r,b)) applied first; (0,0,src_w,src_h) = none */
    void* dst;          /* device, uint8 [dst_h][dst_w][3], rows dst_stride bytes apart */
    int dst_w, dst_h;
    int64_t dst_stride;
    int filter;         /* VT_FILTER_* */
    void* stream;
} vt_resize_args;
int vt_resize_u8(vt_ctx* ctx, const vt_resize_args* args);
/* n images in one call (one bucket batch: different sources, one destination shape each); items[i].stream is
 * ignored, everything is enqueued on `stream` in order */
int vt_resize_u8_batch(vt_ctx* ctx, const vt_resize_args* items, int n, void* stream);
/* SmartResize's centre crop box for a src_w x src_h image and a dst_w x dst_h bucket (modules.py:149-172);
 * host-only helper, box4 = (left, top, right, bottom) */
int vt_smart_crop_box(int src_w, int src_h, int dst_w, int dst_h, int32_t* box4);
/* the fixed-point filter taps of one axis (host-only; for tests): returns ksize, fills bounds[out_size][2]
 * = (first input index, tap count) and kk[out_size][ksize] when the pointers are non-NULL */
int vt_resize_coefficients(int in_size, int out_size, int filter, int32_t* ksize, int32_t* bounds, int32_t* kk);

/* ---------------------------------------------------------------- accounting
 * Kernel classes: 0 implicit GEMM (tcgen05; stride-2 / 1x1 / projections / conv_out), 1 GroupNorm, 2 conv_in
 * gather (fallback), 3 softmax, 4 latent, 5 head, 6 fp32-mode contraction, 7 misc, 8 fused GroupNorm+SiLU+3x3
 * conv for 128 channels (transposed), 9 the same for 256 / 512 channels (CTA pairs), 10 fused attention,
 * 11 conv_in, 12 backward-pass contractions. */
#define VT_NUM_KERNEL_CLASSES 13
int vt_profile_enable(vt_ctx* ctx, int timing);
/* out[class][4] = {launches, milliseconds (timing mode only), flops, bytes}; synchronises the device */
int vt_profile_read(vt_ctx* ctx, double* out, int reset);

/* ---------------------------------------------------------------- single-op entry points
 * Used by the parity tests to pin each kernel family against the oracle.  Activations and
 * weights cross the boundary in fp32 PyTorch layouts and are repacked internally. */
int vt_op_conv2d(vt_ctx* ctx, const float* x /*[N,Cin,H,W]*/, const float* w /*[Cout,Cin,k,k]*/,
                 const float* bias /*[Cout] or NULL*/, const float* residual /*[N,Cout,Ho,Wo] or NULL*/,
                 const float* sc_x /*[N,Cs,Ho,Wo] or NULL*/, const float* sc_w /*[Cout,Cs,1,1] or NULL*/, int N,
                 int Cin, int H, int W, int Cout, int ksize, int stride, int Cs, int precision,
                 float* out /*[N,Cout,Ho,Wo]*/, double* stats /*[N,32,2] (sum,sumsq) or NULL*/, void* stream);
/* conv3x3(silu(group_norm_32(x))) + bias (+ residual) with the normalisation fused into the operand path;
 * x and residual are rounded to bf16 (raw storage format), weights and the normalised operand to fp16 */
int vt_op_conv3_fused(vt_ctx* ctx, const float* x /*[N,Cin,H,W]*/, const float* gamma, const float* beta,
                      const float* w /*[Cout,Cin,3,3]*/, const float* bias, const float* residual /*or NULL*/,
                      const float* sc_x /*[N,Cs,H,W] or NULL*/, const float* sc_w /*[Cout,Cs,1,1] or NULL*/, int N,
                      int Cin, int H, int W, int Cout, int Cs, float eps, int silu, float* out /*[N,Cout,H,W]*/,
                      double* stats /*[N,32,2] of out, or NULL*/, void* stream);
/* fused attention, head_dim 512: out[n,tokens,512] = softmax(scale * q k^T) v + bias_v with
 * qk = [n,tokens,1024] (q | k) and vt = [n,512,tokens] (v transposed); operands rounded to fp16 */
int vt_op_flash_attention(vt_ctx* ctx, const float* qk, const float* vt, const float* bias_v, int n, int tokens,
                          float scale, float* out, void* stream);
int vt_op_gemm_nt(vt_ctx* ctx, const float* A /*[batch,M,K]*/, const float* B /*[batch or 1,N,K]*/,
                  const float* bias, int batch, int M, int N, int K, int b_batched, float alpha, int precision,
                  float* out /*[batch,M,N]*/, void* stream);
int vt_op_group_norm(vt_ctx* ctx, const float* x /*[N,C,H,W]*/, const float* gamma, const float* beta, int N, int C,
                     int H, int W, int groups, float eps, int silu, int precision, float* out, void* stream);
int vt_op_softmax_rows(vt_ctx* ctx, const float* s, int64_t rows, int cols, int precision, float* out,
                       void* stream);

/* ---- backward building blocks of the encoder (the reference fine-tunes the VAE through autograd,
 * train_vae.py:124-186 / train_full.py:201-256; diffusers' ResnetBlock2D / Conv2d / GroupNorm).  NCHW fp32 device
 * tensors in and out like the other vt_op_* entry points; precision VT_PREC_FP32 = FFMA verification mode, anything
 * else = the 16-bit tensor-core mode (activations in the context's raw format, bf16 gradients, fp32 accumulation
 * and fp32 parameter gradients).  Null gradient pointers are skipped. */
/* y = conv2d(x, w) + b, 3x3 pad 1 or 1x1, stride 1: grad_x = d/dx, grad_w [Cout,Cin,k,k], grad_b [Cout] */
int vt_op_conv2d_backward(vt_ctx* ctx, const float* x /*[N,Cin,H,W]*/, const float* w /*[Cout,Cin,k,k]*/,
                          const float* grad_out /*[N,Cout,H,W]*/, int N, int Cin, int H, int W, int Cout, int ksize,
                          int precision, float* grad_x, float* grad_w, float* grad_b, void* stream);
/* y = act(GroupNorm32(x) * gamma + beta), act = SiLU when silu != 0 */
int vt_op_group_norm_backward(vt_ctx* ctx, const float* x /*[N,C,H,W]*/, const float* gamma, const float* beta,
                              const float* grad_y, int N, int C, int H, int W, float eps, int silu, int precision,
                              float* grad_x, float* grad_gamma, float* grad_beta, void* stream);
/* diffusers ResnetBlock2D (temb none): out = shortcut(x) + conv2(silu(norm2(conv1(silu(norm1(x)))))); shortcut =
 * identity (Cin == Cout, sc_* null) or a 1x1 conv.  Runs the block's first half forward (to rebuild conv1's output)
 * and the whole backward on the device in NHWC. */
typedef struct vt_resnet_block_params {
    const float *norm1_w, *norm1_b, *conv1_w, *conv1_b, *norm2_w, *norm2_b, *conv2_w, *conv2_b, *sc_w, *sc_b;
} vt_resnet_block_params;
typedef struct vt_resnet_block_grads {
    float *x, *norm1_w, *norm1_b, *conv1_w, *conv1_b, *norm2_w, *norm2_b, *conv2_w, *conv2_b, *sc_w, *sc_b;
} vt_resnet_block_grads;
int vt_op_resnet_block_backward(vt_ctx* ctx, const float* x /*[N,Cin,H,W]*/, const vt_resnet_block_params* params,
                                const float* grad_out /*[N,Cout,H,W]*/, int N, int Cin, int Cout, int H, int W,
                                int precision, const vt_resnet_block_grads* grads, void* stream);

/* ---- VAE fine-tuning losses of the reference (improved_losses.py), value + analytic gradient in one call
 * (SURVEY.md 8f-4).  Device pointers, fp32. */
typedef struct vt_embed_loss_args {
    const float *a, *p, *n;            /* [B][D] flattened latents: anchor / positive / negative (n null for kind 1);
                                          kind 1: the two embeddings are a and p */
    const float *labels_a, *labels_p;  /* [B][T] multi-hot labels: optional for kind 0, required for kind 1 */
    int B;
    int64_t D;
    int T;
    int kind;        /* 0: ImprovedTripletLoss (improved_losses.py:74-109), 1: ContrastiveLoss (:6-37) */
    int similarity;  /* 0: cosine, 1: euclidean */
    float margin;
    float* loss;     /* [1]: mean over the batch */
    float *grad_a, *grad_p, *grad_n;   /* d loss / d inputs [B][D], each may be null */
    void* stream;
} vt_embed_loss_args;
int vt_embed_loss(vt_ctx* ctx, const vt_embed_loss_args* args);
/* F.mse_loss(x, y) (mean) and d loss / d x */
int vt_mse_loss(vt_ctx* ctx, const float* x, const float* y, int64_t n, float* loss, float* grad_x, void* stream);
/* AdaptiveLossWeights (improved_losses.py:111-125): weights = softmax(log_w / temperature), total = sum w_i L_i;
 * d total / d L_i = weights[i], grad_log_w[j] = d total / d log_w_j */
int vt_adaptive_loss_weights(vt_ctx* ctx, const float* log_w, const float* losses, int n, float temperature,
                             float* total, float* weights, float* grad_log_w, void* stream);

/* ---- encoder training (the reference back-propagates through vae.encode in train_full.py:201-256 and
 * train_vae.py:124-186).  vt_encoder_train_forward = vt_encode on the caller's stream over the whole batch, keeping
 * every activation the backward needs inside the context; vt_encoder_backward consumes them: given d loss / d mean and
 * d loss / d logvar (NCHW fp32 [N][latent_channels][h][w], either may be null; logvar is the clamped value the forward
 * returned -- mask saturated entries yourself) it writes (accumulate = 0) or adds (accumulate != 0) the gradient of
 * EVERY encoder parameter into the fp32 device buffer bound to its name (diffusers key without the "encoder." prefix,
 * e.g. "down_blocks.0.resnets.0.conv1.weight"; the buffer has the parameter's own shape).  Image sizes must be
 * multiples of 8.  Results are bit-reproducible (no atomics). */
#define VT_MAX_TAPES 8
/* slot: which of the context's VT_MAX_TAPES tapes keeps this forward (the reference runs three forwards -- anchor,
 * positive, negative -- before one backward, train_full.py:210-212); a slot is overwritten by its next forward */
int vt_encoder_train_forward(vt_ctx* ctx, const vt_encode_args* args, int slot);
int vt_encoder_grad_bind(vt_ctx* ctx, const char* name, float* grad);
typedef struct vt_encoder_backward_args {
    const float* grad_mean;
    const float* grad_logvar;
    int slot;
    int accumulate;
    void* stream;
} vt_encoder_backward_args;
int vt_encoder_backward(vt_ctx* ctx, const vt_encoder_backward_args* args);
/* frees the activations a slot holds (they are otherwise kept for reuse by the next forward on that slot) */
int vt_encoder_tape_release(vt_ctx* ctx, int slot);

/* ---- decoder training (the reference back-propagates the reconstruction MSE through vae.decode: train_vae.py:124-186,
 * CombinedLoss improved_losses.py:278).  Same protocol as the encoder's: vt_decoder_train_forward = vt_decode on the
 * caller's stream over the whole batch with the activations kept on tape `slot`; vt_decoder_backward takes
 * d loss / d image (NCHW fp32 [N][3][H][W]), writes / adds the gradient of every decoder parameter into the buffer
 * bound to its name (diffusers key without the "decoder." prefix) and, when grad_latent is non-null, writes
 * d loss / d latent (NCHW fp32 [N][latent_channels][h][w], w.r.t. the latent the caller passed: the un-scale of
 * apply_scale_shift is included). */
int vt_decoder_train_forward(vt_ctx* ctx, const vt_decode_args* args, int slot);
int vt_decoder_grad_bind(vt_ctx* ctx, const char* name, float* grad);
typedef struct vt_decoder_backward_args {
    const float* grad_image;
    float* grad_latent;
    int slot;
    int accumulate;
    void* stream;
} vt_decoder_backward_args;
int vt_decoder_backward(vt_ctx* ctx, const vt_decoder_backward_args* args);
int vt_decoder_tape_release(vt_ctx* ctx, int slot);

#ifdef __cplusplus
}
#endif
#endif /* VAE_TAGGER_B200_H */
