// wvn-b200: internal interface of dense_kernels.cu (token grid <-> image resolution).
#pragma once

#include <cuda_runtime.h>

namespace wvn {

struct DenseArgs {
  int batch = 0;
  int dim = 0;                 // feature channels
  int grid_h = 0, grid_w = 0;  // token grid
  int out_h = 0, out_w = 0;    // output resolution
  float scale_y = 0.f, scale_x = 0.f;  // (grid-1)/(out-1), align_corners=True
  long long ld_out = 0;        // row pitch (elements) of interp_pixel_rows output
};

struct LogitsArgs {
  int batch = 0;
  int classes = 0;
  int grid_h = 0, grid_w = 0;
  int out_h = 0, out_w = 0;
  float scale_y = 0.f, scale_x = 0.f;  // grid/out, align_corners=False
  int npad = 0;                // rows per frame of the logits matrix (row 0 = CLS)
  long long ld = 0;            // row pitch of the logits matrix
  int col0 = 0;                // first logit column (multiple of 4)
  int col0_b = 0, classes_b = 0;  // optional second logit range resolved in the same pass
};

// tokens: [B, grid_h*grid_w, dim] fp32 (CLS already dropped).
int upsample_tokens_dense(const float* tokens, float* out_nchw, const DenseArgs& a, cudaStream_t stream);
int interp_pixel_rows(const float* tokens, void* out_bf16, const DenseArgs& a, long long pix0, long long npix,
                      cudaStream_t stream);
int logits_argmax(const float* logits, long long* seg, long long* seg_b, const LogitsArgs& a, cudaStream_t stream);
// STEGO's flip test-time augmentation: head rows of the straight pass (frames [0, B)) become the mean of themselves and
// the horizontally mirrored rows of the flipped pass (frames [B, 2B)); CLS / padding rows are zeroed.  head: [2B*npad, ld].
int flip_average(float* head, int batch, int npad, int grid, long long ld, cudaStream_t stream);

}  // namespace wvn
