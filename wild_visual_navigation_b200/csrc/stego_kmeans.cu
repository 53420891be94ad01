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
// ONE launch runs all iterations for the whole batch: a thread-block CLUSTER of 8 CTAs owns a frame (8 x 392 of the
// 3136 patches at 448 px), so a batch of 32 frames fills the GPU instead of 32 SMs.  Per iteration and CTA:
//   1. one thread = one patch: the 90-d code row sits in registers, the K centroids are read as float4 broadcasts from
//      shared memory; nearest centroid -> assign[] (shared)
//   2. one warp = one patch at a time (lane = channel): the row is added to a WARP-PRIVATE copy of the K x C sums — no
//      atomics (shared-memory float atomics are CAS loops on sm_100, and large clusters are hot spots), deterministic
//   3. the warp copies are summed, the CTA's partial sums go to global memory; after one cluster barrier
//      (release / acquire) every CTA adds the 8 partials in the same order and updates its own copy of the centroids.
#include "common.cuh"
#include "host_common.h"
#include "stego_kmeans.h"

namespace wvn {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kCluster = 8;
constexpr int kMaxK = 64;
constexpr int kMaxC = 128;   // code_dim <= 128 (row registers: 128)

template <int CREG>  // code_dim rounded up to a multiple of 4, in registers
__global__ void __launch_bounds__(kThreads, 1)
stego_kmeans_kernel(float* __restrict__ rows, KmeansArgs a, float* __restrict__ partial) {
  extern __shared__ __align__(16) float ksm[];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int K = a.k, C = a.code_dim, P = a.patches;
  const int CP = (C + 3) & ~3;                 // centroid row stride (float4 reads)
  const int n_priv = a.n_priv;                 // warp-private sum copies that fit in shared memory
  float* cent = ksm;                           // [K][CP], padding columns zero
  float* half_norm = cent + K * CP;            // [K]
  float* cnt_priv = half_norm + kMaxK;         // [n_priv][K]
  float* sum_priv = cnt_priv + kWarps * kMaxK; // [n_priv][K][CP]
  __shared__ unsigned char assign[1024];       // per CTA: <= 1024 patches (8 CTAs per frame)

  const uint32_t rank = cluster_ctarank();
  const int frame = blockIdx.x / kCluster;
  const int per = (P + kCluster - 1) / kCluster;
  const int p0 = rank * per, p1 = min(P, p0 + per);
  float* base = rows + (static_cast<long long>(frame) * a.npad + 1) * a.ld;  // row 0 of a frame is the CLS token
  auto code = [&](int p) { return base + static_cast<long long>(p) * a.ld + a.code_col; };
  float* part_frame = partial + static_cast<long long>(frame) * 2 * kCluster * (K * CP + K);  // [2 buffers][8 ranks][K*CP + K]

  // deterministic init: K patches evenly spaced over the frame's token sequence (every CTA builds the same copy)
  for (int i = t; i < K * CP; i += kThreads) {
    const int k = i / CP, c = i - k * CP;
    const int p = static_cast<int>((static_cast<long long>(2 * k + 1) * P) / (2 * K));
    cent[i] = c < C ? code(p)[c] : 0.f;
  }
  __syncthreads();

  for (int it = 0; it <= a.iters; ++it) {
    const bool last = it == a.iters;  // the last pass only writes the scores of the final centroids
    if (t < K) {
      float n2 = 0.f;
      for (int c = 0; c < C; ++c) n2 = fmaf(cent[t * CP + c], cent[t * CP + c], n2);
      half_norm[t] = 0.5f * n2;
    }
    for (int i = t; i < n_priv * K * CP; i += kThreads) sum_priv[i] = 0.f;
    for (int i = t; i < n_priv * K; i += kThreads) cnt_priv[i] = 0.f;
    __syncthreads();
    // ---- 1. nearest centroid per patch (one thread = one patch, row in registers)
    for (int p = p0 + t; p < p1; p += kThreads) {
      float* row = code(p);
      float x[CREG];
#pragma unroll
      for (int c = 0; c < CREG; ++c) x[c] = c < C ? row[c] : 0.f;
      float best = -INFINITY;
      int arg = 0;
      for (int k = 0; k < K; ++k) {
        const float4* ck = reinterpret_cast<const float4*>(cent + k * CP);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < CREG / 4; ++c4) {
          if (4 * c4 < CP) {
            const float4 cv = ck[c4];
            s0 = fmaf(x[4 * c4 + 0], cv.x, s0);
            s1 = fmaf(x[4 * c4 + 1], cv.y, s1);
            s2 = fmaf(x[4 * c4 + 2], cv.z, s2);
            s3 = fmaf(x[4 * c4 + 3], cv.w, s3);
          }
        }
        const float s = ((s0 + s1) + (s2 + s3)) - half_norm[k];
        if (last) row[a.logit_col - a.code_col + k] = s;
        if (s > best) { best = s; arg = k; }  // first maximum wins, like torch.argmax
      }
      assign[p - p0] = static_cast<unsigned char>(arg);
    }
    if (last) break;
    __syncthreads();
    // ---- 2. warp-private sums: warp w < n_priv adds the rows of patches p0 + w, p0 + w + n_priv, ...
    if (warp < n_priv) {
      float* sp = sum_priv + warp * K * CP;
      float* cp = cnt_priv + warp * K;
      for (int p = p0 + warp; p < p1; p += n_priv) {
        const int k = assign[p - p0];
        const float* row = code(p);
        for (int c = lane; c < C; c += 32) sp[k * CP + c] += row[c];
        if (lane == 0) cp[k] += 1.f;
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- 3. CTA partial -> global; cluster barrier; every CTA sums the 8 partials in rank order
    float* mine = part_frame + ((it & 1) * kCluster + rank) * (K * CP + K);
    for (int i = t; i < K * CP + K; i += kThreads) {
      float s = 0.f;
      if (i < K * CP) {
        for (int w = 0; w < n_priv; ++w) s += sum_priv[w * K * CP + i];
      } else {
        for (int w = 0; w < n_priv; ++w) s += cnt_priv[w * K + (i - K * CP)];
      }
      mine[i] = s;
    }
    __threadfence();
    cluster_sync_all();
    const float* all = part_frame + (it & 1) * kCluster * (K * CP + K);
    for (int i = t; i < K * CP; i += kThreads) {
      const int k = i / CP;
      float s = 0.f, n = 0.f;
      for (int r = 0; r < kCluster; ++r) {
        s += __ldcg(all + r * (K * CP + K) + i);
        n += __ldcg(all + r * (K * CP + K) + K * CP + k);
      }
      if (n > 0.f) cent[i] = s / n;  // empty cluster: keep its centroid
    }
    __syncthreads();
  }
  // the kernel ends with a cluster barrier so that no CTA exits while a peer still reads ... (global memory only: not
  // required for correctness, but keeps the two partial buffers' reuse argument local to this launch)
  if (a.centroids_out && rank == 0) {
    __syncthreads();
    for (int i = t; i < K * C; i += kThreads) {
      const int k = i / C, c = i - k * C;
      a.centroids_out[(static_cast<long long>(frame) * K + k) * C + c] = cent[k * CP + c];
    }
  }
}

}  // namespace

size_t stego_kmeans_workspace_bytes(int batch, int k, int code_dim) {
  const int CP = (code_dim + 3) & ~3;
  return sizeof(float) * static_cast<size_t>(batch) * 2 * kCluster * (static_cast<size_t>(k) * CP + k);
}

int stego_kmeans(float* rows, const KmeansArgs& a_in, float* workspace, cudaStream_t stream) {
  KmeansArgs a = a_in;
  WVN_REQUIRE(rows && workspace && a.batch > 0 && a.patches > 0, "kmeans: empty problem");
  WVN_REQUIRE(a.k > 0 && a.k <= kMaxK && a.code_dim > 0 && a.code_dim <= kMaxC, "kmeans: k=%d (<= %d), code_dim=%d (<= %d)",
              a.k, kMaxK, a.code_dim, kMaxC);
  WVN_REQUIRE(a.patches >= a.k && a.patches <= kCluster * 1024 && a.iters >= 0 && a.logit_col % 4 == 0 &&
                  a.logit_col + a.k <= a.ld && (a.logit_col >= a.code_col + a.code_dim || a.logit_col + a.k <= a.code_col),
              "kmeans: bad geometry (patches=%d) / column layout / iteration count", a.patches);
  const int CP = (a.code_dim + 3) & ~3;
  const size_t fixed = sizeof(float) * (static_cast<size_t>(a.k) * CP + kMaxK + kWarps * kMaxK);
  const size_t per_copy = sizeof(float) * static_cast<size_t>(a.k) * CP;
  int n_priv = static_cast<int>((200 * 1024 - fixed) / per_copy);
  if (n_priv > kWarps) n_priv = kWarps;
  WVN_REQUIRE(n_priv >= 1, "kmeans: k * code_dim too large for shared memory");
  a.n_priv = n_priv;
  const size_t smem = fixed + per_copy * n_priv;
  auto kern = a.code_dim <= 96 ? stego_kmeans_kernel<96> : stego_kmeans_kernel<128>;
  WVN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.batch * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WVN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, rows, a, workspace));
  WVN_CHECK_LAUNCH("stego_kmeans_kernel");
  return WVN_OK;
}

}  // namespace wvn
