// wvn-b200: internal interface of footprint_kernels.cu.
#pragma once

#include <cuda_runtime.h>

namespace wvn {

struct FootprintArgs {
  int batch = 0;
  int n_points = 0;       // polygon points per image (convex, in order)
  int h = 0, w = 0;       // mask size = the (scaled) camera's image size
  int color_batched = 0;  // colors: [batch, 3] if set, else one [3] triple for every image
};

// K: [B,4,4] scaled camera matrices; pose: [B,4,4] camera in world; points: [B,N,3] world frame.
// masks: [B,3,h,w] (NaN outside the polygon) or null; projected: [B,N,2] or null; valid: [B,N] u8 or null;
// sup: [B,3,h,w] updated in place with fmin(sup, mask * *traversability) or null.
int footprint_render(const FootprintArgs& a, const float* K, const float* pose, const float* points, const float* colors,
                     const float* traversability, float* masks, float* projected, unsigned char* valid, float* sup,
                     cudaStream_t stream);

}  // namespace wvn
