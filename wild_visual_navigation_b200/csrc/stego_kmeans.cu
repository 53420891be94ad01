// wvn-b200: per-image k-means of the STEGO code (run_clustering=True, the reference's default for stego
// segmentation: feature_extractor.py:47-53 -> StegoInterface -> Stego.postprocess(image_clustering=True)).
//
// [EXTERNAL-RECALLED] The clustering itself lives in the un-vendored `stego` package; what the reference fixes is the
// call (n_image_clusters clusters per image over the 90-d code, stego_interface.py:43,91-100).  Restated here — and in
// oracle/stego_head.py:image_kmeans, which the tests hold this kernel to — as plain Lloyd iterations (Euclidean, fixed
// iteration count, deterministic evenly-spaced initial centroids, empty clusters keep their centroid) over the code at
// PATCH resolution, followed by `postprocess`'s order of operations for the prediction: the code is upsampled
// (bilinear, align_corners=False) and every pixel takes its nearest centroid.  The nearest-centroid score
// <x, c_k> - |c_k|^2 / 2 is affine in x, so it commutes with the interpolation: this kernel leaves the per-patch scores
// in the head buffer's cluster-logit columns and the existing upsample+argmax kernel (dense_kernels.cu) produces the
// pixel labels — the (H, W, 90) tensor is never formed.
//
// One CTA per frame runs ALL iterations (the 3136 x 90 codes of a frame are re-read from L2 each pass; centroids, sums
// and counts live in shared memory), so the whole clustering is a single launch per batch of frames.
#include "common.cuh"
#include "host_common.h"
#include "stego_kmeans.h"

namespace wvn {

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxK = 64;
constexpr int kMaxC = 128;

__global__ void __launch_bounds__(kThreads)
stego_kmeans_kernel(float* __restrict__ rows, KmeansArgs a) {
  extern __shared__ float ksm[];
  __shared__ float half_norm[kMaxK];
  __shared__ float cnt[kMaxK];
  const int t = threadIdx.x;
  const int K = a.k, C = a.code_dim, P = a.patches;
  const int K16 = (K + 15) / 16 * 16;
  float* cent = ksm;                 // [K16][C], rows >= K stay zero (the 16-wide score blocks read them)
  float* sum = cent + K16 * C;       // [K16][C]
  float* base = rows + (static_cast<long long>(blockIdx.x) * a.npad + 1) * a.ld;  // row 0 of a frame is the CLS token
  auto code = [&](int p) { return base + static_cast<long long>(p) * a.ld + a.code_col; };
  for (int i = t; i < K16 * C; i += kThreads) cent[i] = 0.f;
  __syncthreads();

  // deterministic init: K patches evenly spaced over the frame's token sequence
  for (int i = t; i < K * C; i += kThreads) {
    const int k = i / C, c = i - k * C;
    const int p = static_cast<int>((static_cast<long long>(2 * k + 1) * P) / (2 * K));
    cent[k * C + c] = code(p)[c];
  }
  __syncthreads();

  for (int it = 0; it <= a.iters; ++it) {
    const bool last = it == a.iters;  // the last pass only writes the scores of the final centroids
    for (int i = t; i < K * C; i += kThreads) sum[i] = 0.f;
    if (t < K) {
      float n2 = 0.f;
      for (int c = 0; c < C; ++c) n2 = fmaf(cent[t * C + c], cent[t * C + c], n2);
      half_norm[t] = 0.5f * n2;
      cnt[t] = 0.f;
    }
    __syncthreads();
    for (int p = t; p < P; p += kThreads) {
      float* row = code(p);
      float best = -INFINITY;
      int arg = 0;
      for (int k0 = 0; k0 < K; k0 += 16) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int c = 0; c < C; ++c) {
          const float v = row[c];
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = fmaf(v, cent[(k0 + i) * C + c], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (k0 + i >= K) break;
          const float s = acc[i] - half_norm[k0 + i];
          if (last) row[a.logit_col - a.code_col + k0 + i] = s;
          if (s > best) { best = s; arg = k0 + i; }  // first maximum wins, like torch.argmin of the distances
        }
      }
      if (!last) {
        atomicAdd(&cnt[arg], 1.f);
        const int rot = t & 31;
        for (int i = 0; i < C; ++i) {
          int c = i + rot;  // lanes start at different channels: no same-address conflicts inside a warp
          if (c >= C) c -= C;
          atomicAdd(&sum[arg * C + c], row[c]);
        }
      }
    }
    __syncthreads();
    if (!last) {
      for (int i = t; i < K * C; i += kThreads) {
        const int k = i / C, c = i - k * C;
        if (cnt[k] > 0.f) cent[k * C + c] = sum[k * C + c] / cnt[k];  // empty cluster: keep its centroid
      }
      __syncthreads();
    }
  }
  if (a.centroids_out)
    for (int i = t; i < K * C; i += kThreads) {
      const int k = i / C, c = i - k * C;
      a.centroids_out[(static_cast<long long>(blockIdx.x) * K + k) * C + c] = cent[k * C + c];
    }
}

}  // namespace

int stego_kmeans(float* rows, const KmeansArgs& a, cudaStream_t stream) {
  WVN_REQUIRE(rows && a.batch > 0 && a.patches > 0, "kmeans: empty problem");
  WVN_REQUIRE(a.k > 0 && a.k <= kMaxK && a.code_dim > 0 && a.code_dim <= kMaxC, "kmeans: k=%d (<= %d), code_dim=%d (<= %d)",
              a.k, kMaxK, a.code_dim, kMaxC);
  WVN_REQUIRE(a.patches >= a.k && a.iters >= 0 && a.logit_col % 4 == 0 && a.logit_col + a.k <= a.ld &&
                  (a.logit_col >= a.code_col + a.code_dim || a.logit_col + a.k <= a.code_col),
              "kmeans: bad column layout / iteration count");
  const size_t smem = sizeof(float) * 2 * static_cast<size_t>((a.k + 15) / 16 * 16) * a.code_dim;
  WVN_CHECK_CUDA(cudaFuncSetAttribute(stego_kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  stego_kmeans_kernel<<<a.batch, kThreads, smem, stream>>>(rows, a);
  WVN_CHECK_LAUNCH("stego_kmeans_kernel");
  return WVN_OK;
}

}  // namespace wvn
