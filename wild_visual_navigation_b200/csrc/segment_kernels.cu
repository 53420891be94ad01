// wvn-b200: per-segment reductions over a segmentation map (sm_100a, HBM/L2-bound integer work).
//
// Replaces three Python per-segment loops of the reference, each of which costs one host
// sync per segment:
//   FeatureExtractor.sparsify_features  (feature_extractor.py:389-396)  per-segment feature mean
//   SegmentExtractor.centers            (segment_extractor.py:70-92)    per-segment centroid (x=col, y=row)
//   SegmentExtractor.adjacency_list     (segment_extractor.py:40-67)    4-neighbour segment graph
//   FeatureExtractor.segment_stego      (feature_extractor.py:245-246)  relabel to 0..S-1
//
// The dense (B, D, H, H) feature tensor is never formed: the mean of bilinearly upsampled
// (align_corners=True) features over a segment is a linear function of the patch tokens,
//   feat[s] = (sum_p W[s,p] * tok[p]) / count[s],   W[s,p] = sum_{pixels in s} bilinear weight of patch p,
// so one pass over the pixels accumulates W (plus counts, coordinate sums and adjacency bits)
// and a small second kernel contracts W with the token grid.
#include "common.cuh"
#include "host_common.h"
#include "segment_kernels.h"

namespace wvn {

namespace {

__device__ __forceinline__ void ac_true_coord(int dst, float scale, int in_size, int& i0, int& i1, float& w1) {
  const float s = dst * scale;
  i0 = min(static_cast<int>(s), in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  w1 = s - static_cast<float>(i0);
}

__global__ void __launch_bounds__(256)
segment_accumulate_kernel(const long long* __restrict__ seg, SegmentArgs a, unsigned long long* __restrict__ stats,
                          float* __restrict__ wseg, unsigned int* __restrict__ adj) {
  const long long total = static_cast<long long>(a.batch) * a.h * a.w;
  const int adj_words = (a.smax + 31) >> 5;
  const int P = a.grid_h * a.grid_w;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(p % a.w);
    const int y = static_cast<int>((p / a.w) % a.h);
    const long long b = p / (static_cast<long long>(a.w) * a.h);
    const long long s = seg[p];
    if (s < 0 || s >= a.smax) continue;
    unsigned long long* st = stats + (b * a.smax + s) * 3;
    atomicAdd(st + 0, 1ull);
    atomicAdd(st + 1, static_cast<unsigned long long>(x));
    atomicAdd(st + 2, static_cast<unsigned long long>(y));
    if (wseg != nullptr) {
      // dense features are (H, H): pixel (row=y, col=x) reads dense[:, y, x]; guard x < out size
      int x0, x1, y0, y1;
      float wx, wy;
      ac_true_coord(x, a.scale_x, a.grid_w, x0, x1, wx);
      ac_true_coord(y, a.scale_y, a.grid_h, y0, y1, wy);
      float* wrow = wseg + (b * a.smax + s) * P;
      atomicAdd(wrow + y0 * a.grid_w + x0, (1.f - wy) * (1.f - wx));
      atomicAdd(wrow + y0 * a.grid_w + x1, (1.f - wy) * wx);
      atomicAdd(wrow + y1 * a.grid_w + x0, wy * (1.f - wx));
      atomicAdd(wrow + y1 * a.grid_w + x1, wy * wx);
    }
    if (adj != nullptr) {
      // directed pair (left/top id -> right/bottom id), as the reference's shifted filters pair them
      if (x + 1 < a.w) {
        const long long r = seg[p + 1];
        if (r != s && r >= 0 && r < a.smax)
          atomicOr(adj + (b * a.smax + r) * adj_words + (s >> 5), 1u << (s & 31));
      }
      if (y + 1 < a.h) {
        const long long r = seg[p + a.w];
        if (r != s && r >= 0 && r < a.smax)
          atomicOr(adj + (b * a.smax + r) * adj_words + (s >> 5), 1u << (s & 31));
      }
    }
  }
}

// grid (smax, B), 128 threads; contracts the (sparse) weight row with the token grid.
__global__ void __launch_bounds__(128)
segment_pool_kernel(const float* __restrict__ wseg, const float* __restrict__ tok,
                    const unsigned long long* __restrict__ stats, float* __restrict__ feat,
                    float* __restrict__ centers, SegmentArgs a) {
  __shared__ float w_sm[128];
  const int s = blockIdx.x;
  const long long b = blockIdx.y;
  const int P = a.grid_h * a.grid_w;
  const unsigned long long* st = stats + (b * a.smax + s) * 3;
  const float cnt = static_cast<float>(st[0]);
  if (threadIdx.x == 0 && centers != nullptr) {
    // torch: nonzero(...).float().mean(0) -> (mean col, mean row); empty segment -> NaN
    centers[(b * a.smax + s) * 2 + 0] = static_cast<float>(static_cast<double>(st[1]) / static_cast<double>(st[0]));
    centers[(b * a.smax + s) * 2 + 1] = static_cast<float>(static_cast<double>(st[2]) / static_cast<double>(st[0]));
  }
  if (feat == nullptr) return;
  const float* wrow = wseg + (b * a.smax + s) * P;
  const float* tb = tok + b * P * a.dim;
  constexpr int kMaxPer = 8;  // dim <= 1024
  float acc[kMaxPer];
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) acc[i] = 0.f;
  for (int p0 = 0; p0 < P; p0 += 128) {
    __syncthreads();
    w_sm[threadIdx.x] = (p0 + threadIdx.x < P) ? wrow[p0 + threadIdx.x] : 0.f;
    __syncthreads();
    const int lim = min(128, P - p0);
    for (int q = 0; q < lim; ++q) {
      const float w = w_sm[q];
      if (w == 0.f) continue;  // block-uniform branch
      const float* tr = tb + static_cast<long long>(p0 + q) * a.dim;
#pragma unroll
      for (int i = 0; i < kMaxPer; ++i) {
        const int c = threadIdx.x + 128 * i;
        if (c < a.dim) acc[i] = fmaf(w, __ldg(tr + c), acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxPer; ++i) {
    const int c = threadIdx.x + 128 * i;
    if (c < a.dim) feat[(b * a.smax + s) * a.dim + c] = acc[i] / cnt;
  }
}

// One block per frame: walk the adjacency bitset in (right, left) order — the order
// torch.unique gives the reference's float64 keys left + right*div — and emit int64 pairs.
__global__ void __launch_bounds__(1024)
adjacency_emit_kernel(const unsigned int* __restrict__ adj, long long* __restrict__ edges, int* __restrict__ n_edges,
                      int smax, int max_edges) {
  __shared__ int row_off[1025];
  const long long b = blockIdx.x;
  const int adj_words = (smax + 31) >> 5;
  const unsigned int* ab = adj + b * smax * adj_words;
  const int r = threadIdx.x;
  int cnt = 0;
  if (r < smax)
    for (int wd = 0; wd < adj_words; ++wd) cnt += __popc(ab[r * adj_words + wd]);
  row_off[r + 1] = (r < smax) ? cnt : 0;
  if (r == 0) row_off[0] = 0;
  __syncthreads();
  if (r == 0)
    for (int i = 1; i <= 1024; ++i) row_off[i] += row_off[i - 1];
  __syncthreads();
  if (r == 0) n_edges[b] = row_off[1024];
  if (r < smax) {
    int o = row_off[r];
    long long* eb = edges + b * max_edges * 2;
    for (int wd = 0; wd < adj_words; ++wd) {
      unsigned int bits = ab[r * adj_words + wd];
      while (bits) {
        const int l = __ffs(bits) - 1;
        bits &= bits - 1;
        if (o < max_edges) {
          eb[2 * o + 0] = wd * 32 + l;  // le_idx (left / top segment)
          eb[2 * o + 1] = r;            // ri_idx (right / bottom segment)
        }
        ++o;
      }
    }
  }
}

// ---- relabel: compact the set of labels present in a frame to 0..S-1 (sorted order)
__global__ void __launch_bounds__(256)
label_presence_kernel(const long long* __restrict__ seg, int* __restrict__ present, int batch, long long pix_per_frame,
                      int num_labels) {
  const long long total = batch * pix_per_frame;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = seg[p];
    if (v >= 0 && v < num_labels) present[(p / pix_per_frame) * num_labels + v] = 1;
  }
}

__global__ void label_scan_kernel(int* __restrict__ present, int* __restrict__ counts, int num_labels) {
  // one thread per frame; num_labels is tiny (<= 1024).  present[] becomes the remap table.
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= gridDim.x * blockDim.x) return;
  int* pr = present + static_cast<long long>(b) * num_labels;
  int run = 0;
  for (int i = 0; i < num_labels; ++i) {
    const int has = pr[i];
    pr[i] = has ? run : -1;
    run += has;
  }
  counts[b] = run;
}

__global__ void __launch_bounds__(256)
label_apply_kernel(long long* __restrict__ seg, const int* __restrict__ remap, int batch, long long pix_per_frame,
                   int num_labels) {
  const long long total = batch * pix_per_frame;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = seg[p];
    if (v >= 0 && v < num_labels) seg[p] = remap[(p / pix_per_frame) * num_labels + v];
  }
}

}  // namespace

int segment_accumulate(const long long* seg, const SegmentArgs& a, unsigned long long* stats, float* wseg,
                       unsigned int* adj, cudaStream_t stream) {
  WVN_REQUIRE(a.batch > 0 && a.h > 0 && a.w > 0 && a.smax > 0 && a.smax <= 1024, "segment: bad geometry (smax=%d)",
              a.smax);
  const long long P = static_cast<long long>(a.grid_h) * a.grid_w;
  WVN_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 3 * a.batch * a.smax, stream));
  if (wseg) WVN_CHECK_CUDA(cudaMemsetAsync(wseg, 0, sizeof(float) * a.batch * a.smax * P, stream));
  if (adj) WVN_CHECK_CUDA(cudaMemsetAsync(adj, 0, sizeof(unsigned int) * a.batch * a.smax * ((a.smax + 31) >> 5), stream));
  const long long total = static_cast<long long>(a.batch) * a.h * a.w;
  long long blocks = (total + 255) / 256;
  const long long max_blocks = static_cast<long long>(sm_count()) * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  segment_accumulate_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(seg, a, stats, wseg, adj);
  WVN_CHECK_LAUNCH("segment_accumulate_kernel");
  return WVN_OK;
}

int segment_pool(const float* wseg, const float* tokens, const unsigned long long* stats, float* feat, float* centers,
                 const SegmentArgs& a, cudaStream_t stream) {
  WVN_REQUIRE(a.dim <= 1024, "segment_pool: dim %d too large", a.dim);
  dim3 grid(a.smax, a.batch);
  segment_pool_kernel<<<grid, 128, 0, stream>>>(wseg, tokens, stats, feat, centers, a);
  WVN_CHECK_LAUNCH("segment_pool_kernel");
  return WVN_OK;
}

int adjacency_emit(const unsigned int* adj, long long* edges, int* n_edges, int batch, int smax, int max_edges,
                   cudaStream_t stream) {
  WVN_REQUIRE(smax <= 1024, "adjacency_emit: smax %d > 1024", smax);
  adjacency_emit_kernel<<<batch, 1024, 0, stream>>>(adj, edges, n_edges, smax, max_edges);
  WVN_CHECK_LAUNCH("adjacency_emit_kernel");
  return WVN_OK;
}

int relabel_compact(long long* seg, int* scratch, int* counts, int batch, long long pix_per_frame, int num_labels,
                    cudaStream_t stream) {
  WVN_REQUIRE(num_labels > 0 && num_labels <= 1024, "relabel: num_labels %d unsupported", num_labels);
  WVN_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int) * batch * num_labels, stream));
  const long long total = batch * pix_per_frame;
  long long blocks = (total + 255) / 256;
  const long long max_blocks = static_cast<long long>(sm_count()) * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  label_presence_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(seg, scratch, batch, pix_per_frame, num_labels);
  WVN_CHECK_LAUNCH("label_presence_kernel");
  label_scan_kernel<<<batch, 1, 0, stream>>>(scratch, counts, num_labels);
  WVN_CHECK_LAUNCH("label_scan_kernel");
  label_apply_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(seg, scratch, batch, pix_per_frame, num_labels);
  WVN_CHECK_LAUNCH("label_apply_kernel");
  return WVN_OK;
}

}  // namespace wvn
