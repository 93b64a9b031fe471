// Internal interface of the VAE fine-tuning losses (vt_losses.cu) used by the C-ABI layer.
#pragma once
#include "../../include/vae_tagger_b200.h"
#include "vt_internal.h"

namespace vt {

int embed_loss_chunks(int B, long long D);
size_t embed_loss_scratch_bytes(int B, long long D);
int launch_embed_loss(const vt_embed_loss_args& a, void* scratch, cudaStream_t s, Profiler* prof);
// scratch: 148*8 floats
int launch_mse_loss(const float* x, const float* y, long long n, float* loss, float* grad_x, void* scratch, cudaStream_t s,
                    Profiler* prof);
int launch_adaptive_weights(const float* log_w, const float* losses, int n, float temp, float* total, float* weights,
                            float* grad_log_w, cudaStream_t s);

}  // namespace vt
