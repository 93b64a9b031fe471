// Internal interface of the image preprocessing kernels (vt_resize.cu).
#pragma once
#include <vector>

#include "../../include/vae_tagger_b200.h"

namespace vt {

struct Profiler;
struct ResizeCache;  // device-resident coefficient tables + the intermediate image
ResizeCache* resize_cache_create();
void resize_cache_destroy(ResizeCache*);
void resize_coefficients(int in_size, int out_size, int filter, int* ksize, std::vector<int>& bounds,
                         std::vector<int>& kk);
void smart_crop_box(int src_w, int src_h, int dst_w, int dst_h, int* box4);
int resize_u8(ResizeCache*, const vt_resize_args&, Profiler*);

}  // namespace vt
