// wvn-b200: SLIC superpixels for FeatureExtractor(segmentation_type="slic") — replaces fast_slic's
// Slic(num_components, compactness).iterate(np.uint8(img * 255)) (feature_extractor.py:88-95, 221-225).
//
// All-integer SLIC (definition: oracle/slic.py): 8-bit sRGB -> CIELAB in 1/64 units through lookup tables, a regular
// grid of K = nx * ny centres, `iters` rounds of {nearest of the 3x3 neighbouring cells' centres under
// d = |dLab|^2 S^2 + |dxy|^2 (compactness * 64)^2 ; centre <- rounded mean of its pixels}.  Integer sums make the result
// independent of the order in which threads accumulate, so the labels are bit-identical to the numpy oracle.
// One launch per round: every block first derives the current centres from the previous round's sums (a few hundred
// integers), assigns its 32x32 pixel tile, accumulates (count, L, a, b, x, y) per centre in shared memory and flushes
// with 64-bit atomics; the sum buffers rotate over three slots so that no round ever clears a slot another block still
// reads.  Only the last round writes labels.  HBM traffic: 8 B/pixel/round (the Lab image stays L2-resident).
#include "slic_kernels.h"

#include <math.h>

#include "host_common.h"

namespace wvn {
namespace {

constexpr int kTile = 32;      // pixels per tile side
constexpr int kThreads = 256;
constexpr int kMaxK = 1024;

struct FrameWs {               // per-frame workspace layout (element offsets computed on the host)
  short4* lab;                 // [h*w]
  int* centers;                // [2][K][5]
  unsigned long long* sums;    // [3][K][6]
};

__device__ __forceinline__ FrameWs frame_ws(unsigned char* base, size_t frame_bytes, int b, int hw, int K) {
  unsigned char* p = base + frame_bytes * static_cast<size_t>(b);
  FrameWs f;
  f.lab = reinterpret_cast<short4*>(p);
  size_t off = (sizeof(short4) * static_cast<size_t>(hw) + 15) & ~static_cast<size_t>(15);
  f.sums = reinterpret_cast<unsigned long long*>(p + off);
  off += sizeof(unsigned long long) * 3 * K * 6;
  f.centers = reinterpret_cast<int*>(p + off);
  return f;
}

size_t frame_bytes_host(int hw, int K) {
  size_t off = (sizeof(short4) * static_cast<size_t>(hw) + 15) & ~static_cast<size_t>(15);
  off += sizeof(unsigned long long) * 3 * K * 6;
  off += sizeof(int) * 2 * K * 5;
  return (off + 255) & ~static_cast<size_t>(255);
}

// float image -> Lab (1/64 units), and the initial centres (the Lab colour at every grid-cell centre pixel).
__global__ void __launch_bounds__(kThreads)
slic_lab_kernel(SlicArgs a, const float* __restrict__ img, const int* __restrict__ lut_g, const int* __restrict__ lut_m,
                const int* __restrict__ lut_f, unsigned char* __restrict__ ws, size_t frame_bytes) {
  const int b = blockIdx.y;
  const int hw = a.h * a.w;
  const int K = a.nx * a.ny;
  const FrameWs f = frame_ws(ws, frame_bytes, b, hw, K);
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= hw) return;
  const float* im = img + static_cast<size_t>(b) * 3 * hw;
  int lin[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // np.uint8(img * 255): float32 product, truncation (values outside [0, 255] are clamped instead of wrapped)
    int v = static_cast<int>(__fmul_rn(im[static_cast<size_t>(c) * hw + p], 255.f));
    v = min(255, max(0, v));
    lin[c] = __ldg(lut_g + v);
  }
  int fxyz[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int t = (lut_m[3 * i] * lin[0] + lut_m[3 * i + 1] * lin[1] + lut_m[3 * i + 2] * lin[2] + 2048) >> 12;
    fxyz[i] = __ldg(lut_f + min(4095, t));
  }
  const int L = (116 * fxyz[1] - 65536) >> 6;
  const int A = (500 * (fxyz[0] - fxyz[1])) >> 6;   // arithmetic shift = floor, as numpy's >>
  const int B = (200 * (fxyz[1] - fxyz[2])) >> 6;
  f.lab[p] = make_short4(static_cast<short>(L), static_cast<short>(A), static_cast<short>(B), 0);
  const int x = p % a.w, y = p / a.w;
  const int gx = min(a.nx - 1, (x * a.nx) / a.w), gy = min(a.ny - 1, (y * a.ny) / a.h);
  if (x == ((2 * gx + 1) * a.w) / (2 * a.nx) && y == ((2 * gy + 1) * a.h) / (2 * a.ny)) {
    int* c = f.centers + (gy * a.nx + gx) * 5;      // slot 0 = the centres round 0 assigns with
    c[0] = L; c[1] = A; c[2] = B; c[3] = x; c[4] = y;
  }
}

__device__ __forceinline__ long long floor_div(long long n, long long d) {  // d > 0
  long long q = n / d;
  return (n % d < 0) ? q - 1 : q;
}

// smem: cent[K][5] int | acc[K][6] int
__global__ void __launch_bounds__(kThreads)
slic_round_kernel(SlicArgs a, int round, int last, unsigned char* __restrict__ ws, size_t frame_bytes,
                  long long* __restrict__ labels) {
  extern __shared__ int sm[];
  const int b = blockIdx.y;
  const int hw = a.h * a.w;
  const int K = a.nx * a.ny;
  int* cent = sm;
  int* acc = sm + K * 5;
  const FrameWs f = frame_ws(ws, frame_bytes, b, hw, K);
  const int t = threadIdx.x;
  // ---- the centres this round assigns with: round 0 = the grid initialisation, later = rounded means of round-1
  const int* prev = f.centers + ((round + 1) & 1) * K * 5;          // written by round - 1 (slot 0 for round 0: see below)
  int* cur = f.centers + (round & 1) * K * 5;
  const unsigned long long* sprev = f.sums + static_cast<size_t>((round + 2) % 3) * K * 6;   // (round - 1) % 3
  unsigned long long* scur = f.sums + static_cast<size_t>(round % 3) * K * 6;
  unsigned long long* snext = f.sums + static_cast<size_t>((round + 1) % 3) * K * 6;
  for (int k = t; k < K; k += kThreads) {
    int c[5];
    if (round == 0) {
      for (int j = 0; j < 5; ++j) c[j] = cur[k * 5 + j];             // slic_lab_kernel wrote slot 0
    } else {
      const long long n = static_cast<long long>(sprev[k * 6]);
      for (int j = 0; j < 5; ++j) {
        const long long s = static_cast<long long>(sprev[k * 6 + 1 + j]);   // two's complement sums
        c[j] = n > 0 ? static_cast<int>(floor_div(2 * s + n, 2 * n)) : prev[k * 5 + j];
      }
      if (blockIdx.x == 0)
        for (int j = 0; j < 5; ++j) cur[k * 5 + j] = c[j];
    }
    for (int j = 0; j < 5; ++j) cent[k * 5 + j] = c[j];
    for (int j = 0; j < 6; ++j) acc[k * 6 + j] = 0;
    if (blockIdx.x == 0)
      for (int j = 0; j < 6; ++j) snext[k * 6 + j] = 0ull;           // nobody touches slot (round + 1) % 3 this round
  }
  __syncthreads();
  // ---- assign the tile
  const int tiles_x = (a.w + kTile - 1) / kTile;
  const int tx0 = (blockIdx.x % tiles_x) * kTile, ty0 = (blockIdx.x / tiles_x) * kTile;
  for (int i = t; i < kTile * kTile; i += kThreads) {
    const int x = tx0 + (i % kTile), y = ty0 + (i / kTile);
    if (x >= a.w || y >= a.h) continue;
    const short4 px = f.lab[y * a.w + x];
    const int gx = min(a.nx - 1, (x * a.nx) / a.w), gy = min(a.ny - 1, (y * a.ny) / a.h);
    long long best = 0x7fffffffffffffffll;
    int lab = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int qx = gx + dx, qy = gy + dy;
        if (qx < 0 || qx >= a.nx || qy < 0 || qy >= a.ny) continue;
        const int k = qy * a.nx + qx;
        const int* c = cent + k * 5;
        const long long dl = px.x - c[0], da = px.y - c[1], db = px.z - c[2];
        const long long ex = x - c[3], ey = y - c[4];
        const long long d = (dl * dl + da * da + db * db) * a.S2 + (ex * ex + ey * ey) * a.M2;
        if (d < best) { best = d; lab = k; }                        // increasing k: the first minimum wins
      }
    }
    if (last) {
      labels[static_cast<size_t>(b) * hw + static_cast<size_t>(y) * a.w + x] = lab;
    } else {
      int* q = acc + lab * 6;
      atomicAdd(q, 1); atomicAdd(q + 1, px.x); atomicAdd(q + 2, px.y); atomicAdd(q + 3, px.z);
      atomicAdd(q + 4, x); atomicAdd(q + 5, y);
    }
  }
  if (last) return;
  __syncthreads();
  for (int i = t; i < K * 6; i += kThreads) {
    const int v = acc[i];
    if (v != 0) atomicAdd(scur + i, static_cast<unsigned long long>(static_cast<long long>(v)));
  }
}

}  // namespace

void slic_geometry(int h, int w, int num_components, int* S, int* nx, int* ny) {
  *S = std::max(1, static_cast<int>(nearbyint(sqrt(static_cast<double>(h) * w / static_cast<double>(num_components)))));
  *nx = std::max(1, static_cast<int>(nearbyint(static_cast<double>(w) / *S)));
  *ny = std::max(1, static_cast<int>(nearbyint(static_cast<double>(h) / *S)));
}

void slic_tables(int* g256, int* m9, int* f4096) {
  for (int v = 0; v < 256; ++v) {
    const double x = v / 255.0;
    const double lin = x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4);
    g256[v] = static_cast<int>(nearbyint(4095.0 * lin));
  }
  const double M[3][3] = {{0.4124564, 0.3575761, 0.1804375}, {0.2126729, 0.7151522, 0.0721750}, {0.0193339, 0.1191920, 0.9503041}};
  for (int i = 0; i < 3; ++i) {
    const double white = M[i][0] + M[i][1] + M[i][2];
    for (int j = 0; j < 3; ++j) m9[3 * i + j] = static_cast<int>(nearbyint(4096.0 * M[i][j] / white));
  }
  const double d = 6.0 / 29.0;
  for (int t = 0; t < 4096; ++t) {
    const double x = t / 4095.0;
    const double f = x > d * d * d ? cbrt(x) : x / (3.0 * d * d) + 4.0 / 29.0;
    f4096[t] = static_cast<int>(nearbyint(4096.0 * f));
  }
}

size_t slic_workspace_bytes(int batch, int h, int w, int num_components) {
  int S, nx, ny;
  slic_geometry(h, w, num_components, &S, &nx, &ny);
  return frame_bytes_host(h * w, nx * ny) * static_cast<size_t>(batch);
}

int slic_segment(const float* img, int batch, int h, int w, int num_components, float compactness, int iters,
                 const int* lut_g, const int* lut_m, const int* lut_f, long long* labels, void* workspace,
                 cudaStream_t stream) {
  WVN_REQUIRE(batch > 0 && h > 0 && w > 0 && num_components > 0 && iters > 0, "slic: bad sizes");
  WVN_REQUIRE(h <= 16384 && w <= 16384, "slic: image side above 16384");
  SlicArgs a;
  a.batch = batch; a.h = h; a.w = w;
  slic_geometry(h, w, num_components, &a.S, &a.nx, &a.ny);
  const int K = a.nx * a.ny;
  WVN_REQUIRE(K <= kMaxK, "slic: %d clusters, at most %d", K, kMaxK);
  a.S2 = static_cast<long long>(a.S) * a.S;
  a.M2 = static_cast<long long>(nearbyint((static_cast<double>(compactness) * 64.0) * (static_cast<double>(compactness) * 64.0)));
  const size_t fb = frame_bytes_host(h * w, K);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  // sums slot 0 must start at zero (round 0 accumulates into it); slots 1 / 2 are cleared by the rounds themselves
  for (int b = 0; b < batch; ++b) {
    const size_t off = fb * b + ((sizeof(short4) * static_cast<size_t>(h) * w + 15) & ~static_cast<size_t>(15));
    WVN_CHECK_CUDA(cudaMemsetAsync(ws + off, 0, sizeof(unsigned long long) * K * 6, stream));
  }
  dim3 g1((h * w + kThreads - 1) / kThreads, batch);
  slic_lab_kernel<<<g1, kThreads, 0, stream>>>(a, img, lut_g, lut_m, lut_f, ws, fb);
  WVN_CHECK_LAUNCH("slic_lab_kernel");
  const int tiles = ((w + kTile - 1) / kTile) * ((h + kTile - 1) / kTile);
  const size_t smem = sizeof(int) * K * 11;
  if (smem > 48 * 1024)
    WVN_CHECK_CUDA(cudaFuncSetAttribute(slic_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  for (int r = 0; r < iters; ++r) {
    slic_round_kernel<<<dim3(tiles, batch), kThreads, smem, stream>>>(a, r, r == iters - 1, ws, fb, labels);
    WVN_CHECK_LAUNCH("slic_round_kernel");
  }
  return WVN_OK;
}

}  // namespace wvn
