// Internal interface of the head training step (vt_head_train.cu) used by the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/vae_tagger_b200.h"

namespace vt {

struct Profiler;

struct HeadParamEntry {
    std::string name;  // reference state-dict key
    int64_t offset;    // in floats, into the flat parameter / gradient buffer
    int64_t numel;
    std::vector<int64_t> shape;
};
// trainable parameters in the order of the reference module's parameters() (modules.py:303-422)
std::vector<HeadParamEntry> head_param_layout(const vt_head_config& h);
size_t head_train_workspace_floats(const vt_head_config& h, int B, int H, int W);
int head_train_step(const vt_head_config& h, const vt_head_train_args& a, float* ws, Profiler* pf);
int head_dropout_masks(const vt_head_config& h, int B, float attention_dropout, unsigned long long seed, float* attn,
                       float* const* cls, cudaStream_t s);
int launch_adamw(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 float wd, long long step, float grad_scale, float max_norm, double* scratch, float* norm_out,
                 int zero_grad, cudaStream_t s, Profiler* pf);

}  // namespace vt
