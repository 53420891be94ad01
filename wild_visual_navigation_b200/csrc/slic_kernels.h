// wvn-b200: internal interface of slic_kernels.cu.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace wvn {

struct SlicArgs {
  int batch = 0, h = 0, w = 0;
  int S = 0, nx = 0, ny = 0;   // grid interval, grid cells per axis (K = nx * ny clusters)
  long long S2 = 0, M2 = 0;    // S^2 and (compactness * 64)^2: the two weights of the integer distance
};

void slic_geometry(int h, int w, int num_components, int* S, int* nx, int* ny);
// Host lookup tables of the 8-bit sRGB -> CIELAB conversion: g256 (byte -> 12-bit linear), m9 (RGB -> XYZ / white, 12-bit
// fixed point, row major), f4096 (f(t) * 4096).
void slic_tables(int* g256, int* m9, int* f4096);
size_t slic_workspace_bytes(int batch, int h, int w, int num_components);
// img: [B,3,h,w] fp32 in [0,1]; lut_*: DEVICE copies of slic_tables(); labels: [B,h,w] int64 cluster ids in [0, nx*ny).
int slic_segment(const float* img, int batch, int h, int w, int num_components, float compactness, int iters,
                 const int* lut_g, const int* lut_m, const int* lut_f, long long* labels, void* workspace,
                 cudaStream_t stream);

}  // namespace wvn
