// wvn-b200: fused per-pixel traversability head (sm_100a).
//
// Replaces, per frame (wvn_feature_extractor_node.py:319-370, simple_mlp.py:33-39,
// confidence_generator.py:182-193):
//     x = dense_feat[0].permute(1,2,0).reshape(-1, D)          # 200704 x 384, bilinear(align_corners=True)
//     pred = SimpleMLP(x);  trav = pred[:, 0]
//     loss_reco = mse(pred[:, 1:], x).mean(1);  conf = inference_without_update(loss_reco)
//
// The per-pixel work is restructured algebraically so that neither x (384-d) nor pred (385-d) is
// ever formed per pixel — only the 256-d hidden layer is:
//   * layer 1 is affine and bilinear weights sum to 1:  W1 x + b1 = sum_k w_k (W1 t_k + b1)
//     -> G = tokens @ W1^T + b1 is computed ONCE PER TOKEN (3136 rows/frame, tcgen05 GEMM) and the
//        kernel below interpolates G (256 ch) instead of x (384 ch) followed by a 384->256 GEMM;
//   * with r = R h2 + c the reconstruction (R = W3[1:], c = b3[1:]):
//       D*loss = |r|^2 - 2 r.x + |x|^2
//       |r|^2  = h2^T (R^T R) h2 + 2 (R^T c).h2 + c.c          (32x32 quadratic form, fp32)
//       r.x    = sum_k w_k ( h2 . (R^T t_k) + c.t_k )            (U = tokens @ R, cT = tokens @ c: per token)
//       |x|^2  = bilinear form of the 2x2 neighbourhood's Gram    (5 dot products per token)
// Per pixel that leaves: interpolate 256+33 channels, one 256->32 layer on the tensor core
// (tcgen05, M=128 pixels, N=32), and ~650 fp32 FMAs — ~12x fewer FLOPs than the direct form, and
// no HBM traffic beyond the two output maps.  bf16 rounding points are the same as the unfused
// path (tokens, weights, h1); everything downstream of the layer-2 accumulator is fp32.
#include <algorithm>

#include "common.cuh"
#include "host_common.h"
#include "pixel_head.h"

namespace wvn {

namespace {

constexpr int kH1 = 256, kH2 = 32;
constexpr int kTileW = 64, kTileH = 2;            // 128 pixels per tile: lane/TMEM row i = r*64 + px
constexpr int kWinMax = 10;                        // token-window columns per tile (ratio 8: 7 + 3)
constexpr int kGvStride = 292;                     // 256 G + 32 U + cT_hi + cT_lo, padded to float4
constexpr int kThreads = 256;
constexpr uint32_t kOffA = 0;                      // [4 K-blocks][128 rows][128 B]  h1 tile (bf16, swizzled)
constexpr uint32_t kOffW2 = 65536;                 // [4 K-blocks][32 rows][128 B]
constexpr uint32_t kOffGv = kOffW2 + 16384;        // [2][kWinMax][kGvStride] fp32
constexpr uint32_t kOffN2 = kOffGv + 2 * kWinMax * kGvStride * 4;  // [2][kWinMax][2] fp32: |xv|^2, xv_c.xv_{c+1}
constexpr uint32_t kOffConst = kOffN2 + 2 * kWinMax * 2 * 4;      // packed constants (see PixelHeadConsts)
constexpr uint32_t kOffBar = kOffConst + sizeof(PixelHeadConsts);
constexpr uint32_t kSmemBytes = kOffBar + 64;

__device__ __forceinline__ void ac_true(int dst, float scale, int in_size, int& i0, float& w1) {
  const float s = dst * scale;
  i0 = min(static_cast<int>(s), in_size - 1);
  w1 = s - static_cast<float>(i0);
}

__global__ void __launch_bounds__(kThreads, 2)
pixel_head_kernel(const __grid_constant__ CUtensorMap tmap_w2, const PixelHeadArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* gv = reinterpret_cast<float*>(smem + kOffGv);
  float* n2 = reinterpret_cast<float*>(smem + kOffN2);
  PixelHeadConsts* cs = reinterpret_cast<PixelHeadConsts*>(smem + kOffConst);
  uint64_t* w2_full = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* mma_done = w2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w2_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = a.W / kTileW, tiles_y = a.H / kTileH;
  const long long tiles_per_frame = static_cast<long long>(tiles_x) * tiles_y;
  const long long num_tiles = tiles_per_frame * a.batch;
  const int P = a.gh * a.gw;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) { printf("[wvn] pixel_head: smem base not 1024B aligned\n"); __trap(); }
    mbar_init(w2_full, 1);
    mbar_init(mma_done, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(PixelHeadConsts) / 4); i += kThreads)
    reinterpret_cast<float*>(cs)[i] = reinterpret_cast<const float*>(a.consts)[i];
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(w2_full, 4 * 32 * 128);
    for (int kb = 0; kb < 4; ++kb) tma_load_2d(&tmap_w2, w2_full, smem + kOffW2 + kb * 4096, kb * 64, 0);
  }

  uint32_t mma_phase = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = static_cast<int>(tile / tiles_per_frame);
    const int trem = static_cast<int>(tile - b * tiles_per_frame);
    const int py0 = (trem / tiles_x) * kTileH, px0 = (trem % tiles_x) * kTileW;
    int cx0;
    float tmpw;
    ac_true(px0, a.sx, a.gw, cx0, tmpw);
    const float* gub = a.gu + static_cast<long long>(b) * P * a.ldg;

    // ---------------- phase A: vertical blend of the token window (G | U | cT) + |x|^2 ingredients
    {
      // all global loads of this thread are issued before any is consumed (one L2 latency per tile, not one
      // per loop trip): at most kTileH * kWinMax * 73 / 256 = 6 float4 pairs per thread
      constexpr int kMaxIt = (kTileH * kWinMax * (kGvStride / 4) + kThreads - 1) / kThreads;
      const int n_items = kTileH * a.ww * (kGvStride / 4);
      float4 g0[kMaxIt], g1[kMaxIt];
      float wys[kMaxIt];
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        const int i = threadIdx.x + it * kThreads;
        if (i < n_items) {
          const int v4 = i % (kGvStride / 4);
          const int c = (i / (kGvStride / 4)) % a.ww;
          const int r = i / ((kGvStride / 4) * a.ww);
          int y0;
          ac_true(py0 + r, a.sy, a.gh, y0, wys[it]);
          const int y1 = min(y0 + 1, a.gh - 1);
          const int tc = min(cx0 + c, a.gw - 1);
          g0[it] = __ldg(reinterpret_cast<const float4*>(gub + (static_cast<long long>(y0) * a.gw + tc) * a.ldg) + v4);
          g1[it] = __ldg(reinterpret_cast<const float4*>(gub + (static_cast<long long>(y1) * a.gw + tc) * a.ldg) + v4);
        }
      }
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        const int i = threadIdx.x + it * kThreads;
        if (i < n_items) {
          const float wy = wys[it];
          float4 o;
          o.x = (1.f - wy) * g0[it].x + wy * g1[it].x;
          o.y = (1.f - wy) * g0[it].y + wy * g1[it].y;
          o.z = (1.f - wy) * g0[it].z + wy * g1[it].z;
          o.w = (1.f - wy) * g0[it].w + wy * g1[it].w;
          reinterpret_cast<float4*>(gv)[i] = o;  // i == (r * ww + c) * (kGvStride / 4) + v4
        }
      }
    }
    if (threadIdx.x < kTileH * a.ww) {
      // xv_c = (1-wy) t[y0,c] + wy t[y1,c]:  |xv_c|^2 and xv_c . xv_{c+1} from the per-token Gram entries
      const int c = threadIdx.x % a.ww, r = threadIdx.x / a.ww;
      int y0;
      float wy;
      ac_true(py0 + r, a.sy, a.gh, y0, wy);
      const bool same_y = (y0 + 1 > a.gh - 1);
      const int y1 = same_y ? y0 : y0 + 1;
      const int tc = min(cx0 + c, a.gw - 1);
      const bool same_x = (tc + 1 > a.gw - 1);
      const int tc1 = same_x ? tc : tc + 1;
      const float* gr = a.gram + static_cast<long long>(b) * P * 5;
      auto G = [&](int y, int x, int k) { return __ldg(gr + (static_cast<long long>(y) * a.gw + x) * 5 + k); };
      const float u = 1.f - wy;
      // K(a,b) lookups: 0 self, 1 right neighbour, 2 lower neighbour, 3 lower-right, 4 (right . lower)
      const float s00 = G(y0, tc, 0), s10 = G(y1, tc, 0);
      const float v0 = same_y ? s00 : G(y0, tc, 2);
      const float nn = u * u * s00 + 2.f * u * wy * v0 + wy * wy * s10;
      float xx;
      if (same_x) {
        xx = nn;
      } else {
        const float h0 = G(y0, tc, 1), h1 = G(y1, tc, 1);
        const float d = same_y ? h0 : G(y0, tc, 3);
        const float an = same_y ? h0 : G(y0, tc, 4);
        xx = u * u * h0 + u * wy * (d + an) + wy * wy * h1;
      }
      (void)tc1;
      n2[(r * a.ww + c) * 2 + 0] = nn;
      n2[(r * a.ww + c) * 2 + 1] = xx;
    }
    __syncthreads();

    // ---------------- phase B: horizontal blend -> ReLU -> bf16 -> swizzled A tile (h1)
    {
      const int units = kTileH * (a.ww - 1);  // (row, cell) pairs; cell c spans window columns [c, c+1]
      for (int u = warp; u < units; u += kThreads / 32) {
        const int r = u / (a.ww - 1), cell = u % (a.ww - 1);
        // pixels of this row whose left token column is cx0 + cell (a contiguous run)
        unsigned long long mask = 0ull;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int x0;
          float wx;
          ac_true(px0 + lane + 32 * h, a.sx, a.gw, x0, wx);
          const unsigned int bal = __ballot_sync(0xffffffffu, x0 - cx0 == cell);
          mask |= static_cast<unsigned long long>(bal) << (32 * h);
        }
        if (mask == 0ull) continue;
        const float* g0p = gv + (r * a.ww + cell) * kGvStride + 8 * lane;
        const float4 a0 = reinterpret_cast<const float4*>(g0p)[0], a1 = reinterpret_cast<const float4*>(g0p)[1];
        const float4 b0 = reinterpret_cast<const float4*>(g0p + kGvStride)[0];
        const float4 b1 = reinterpret_cast<const float4*>(g0p + kGvStride)[1];
        const float g[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float dg[8] = {b0.x - a0.x, b0.y - a0.y, b0.z - a0.z, b0.w - a0.w,
                             b1.x - a1.x, b1.y - a1.y, b1.z - a1.z, b1.w - a1.w};
        while (mask) {
          const int px = __ffsll(static_cast<long long>(mask)) - 1;
          mask &= mask - 1;
          int x0;
          float wx;
          ac_true(px0 + px, a.sx, a.gw, x0, wx);
          float h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) h[i] = fmaxf(fmaf(wx, dg[i], g[i]), 0.f);
          const int row = r * kTileW + px;
          // channel block 8*lane..8*lane+7: K-block lane/8, 16-byte chunk lane%8 (128B swizzle)
          uint8_t* dst = smem + kOffA + (lane >> 3) * 16384 + row * 128 + (((lane & 7) ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                      pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        }
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- layer 2 on the tensor core: D[128, 32] = h1[128, 256] @ W2^T
    if (threadIdx.x == 128) {  // warp 4 lane 0: not an epilogue warp, so warps 0-3 stay convergent
      mbar_wait(w2_full, 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc_bf16(128, kH2);
#pragma unroll
      for (int ks = 0; ks < kH1 / 16; ++ks) {
        const uint64_t da = make_sw128_kmajor_desc(smem_u32(smem + kOffA + (ks >> 2) * 16384)) + 2 * (ks & 3);
        const uint64_t db = make_sw128_kmajor_desc(smem_u32(smem + kOffW2 + (ks >> 2) * 4096)) + 2 * (ks & 3);
        umma_bf16_ss(tmem_d, da, db, idesc, ks != 0);
      }
      umma_commit(mma_done);
    }

    // ---------------- epilogue: one thread per pixel (warps 0-3 <-> TMEM lanes 0-127)
    if (warp < 4) {
      const int i = threadIdx.x;  // pixel / TMEM lane
      const int r = i >> 6, px = i & 63;
      mbar_wait(mma_done, mma_phase);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32(tmem_d + (static_cast<uint32_t>(warp * 32) << 16), raw);
      tmem_ld_wait();
      float h2[kH2];
#pragma unroll
      for (int j = 0; j < kH2; ++j) h2[j] = fmaxf(__uint_as_float(raw[j]) + cs->b2[j], 0.f);
      int x0;
      float wx;
      ac_true(px0 + px, a.sx, a.gw, x0, wx);
      const int c0 = x0 - cx0;
      const bool same_x = (x0 + 1 > a.gw - 1);
      const float* u0 = gv + (r * a.ww + c0) * kGvStride + kH1;
      const float* u1 = u0 + kGvStride;  // window column c0+1 (a clamped duplicate at the right border)
      // traversability logit, |r|^2 quadratic form (upper-triangular M with doubled off-diagonals), r.x
      float t = cs->b0, q = cs->cc, cross = 0.f;
#pragma unroll
      for (int j = 0; j < kH2; ++j) {
        t = fmaf(cs->w0[j], h2[j], t);
        float acc = cs->tv[j];
#pragma unroll
        for (int k = j; k < kH2; ++k) acc = fmaf(cs->m[j * kH2 + k], h2[k], acc);
        q = fmaf(h2[j], acc, q);
        cross = fmaf(h2[j], fmaf(wx, u1[j] - u0[j], u0[j]), cross);
      }
      cross += fmaf(wx, (u1[kH2] + u1[kH2 + 1]) - (u0[kH2] + u0[kH2 + 1]), u0[kH2] + u0[kH2 + 1]);
      const float* nn = n2 + (r * a.ww + c0) * 2;
      const float ux = 1.f - wx;
      const float n_c1 = same_x ? nn[0] : nn[2];
      const float xx = same_x ? nn[0] : nn[1];
      const float gram = ux * ux * nn[0] + 2.f * ux * wx * xx + wx * wx * n_c1;
      const float loss = fmaxf(q - 2.f * cross + gram, 0.f) / static_cast<float>(a.feat);
      const float mean = __ldg(a.cg_mean), sd = __ldg(a.cg_std);
      const float shifted = mean + sd * a.std_factor;
      const float lo = fmaxf(shifted - sd, 0.f), hi = shifted + sd;
      const float xc = fminf(fmaxf(loss, lo), hi);
      const long long o = (static_cast<long long>(b) * a.H + py0 + r) * a.W + px0 + px;
      a.trav[o] = 1.f / (1.f + __expf(-t));
      a.conf[o] = 1.f - (xc - lo) / (hi - lo);
      if (a.loss_reco != nullptr) a.loss_reco[o] = loss;
      tc_fence_before();
    }
    mma_phase ^= 1;
    __syncthreads();  // A tile, Gv and the accumulator may be overwritten by the next tile
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_d, 32);
  }
}

// One warp per token: self / right / lower / lower-right / (right . lower) dot products (bf16 tokens).
__global__ void __launch_bounds__(256)
token_gram_kernel(const __nv_bfloat16* __restrict__ tok, float* __restrict__ gram, int batch, int gh, int gw, int dim) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long warps_total = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long P = static_cast<long long>(gh) * gw;
  for (long long t = warp_global; t < batch * P; t += warps_total) {
    const int x = static_cast<int>(t % gw), y = static_cast<int>((t / gw) % gh);
    const bool has_r = x + 1 < gw, has_d = y + 1 < gh;
    const __nv_bfloat16* p00 = tok + t * dim;
    const __nv_bfloat16* p01 = has_r ? p00 + dim : p00;
    const __nv_bfloat16* p10 = has_d ? p00 + static_cast<long long>(gw) * dim : p00;
    const __nv_bfloat16* p11 = p10 + (has_r ? dim : 0);
    float s = 0.f, h = 0.f, v = 0.f, d = 0.f, an = 0.f;
    for (int c = lane * 2; c < dim; c += 64) {
      const float2 a00 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p00 + c));
      const float2 a01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p01 + c));
      const float2 a10 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p10 + c));
      const float2 a11 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p11 + c));
      s = fmaf(a00.x, a00.x, fmaf(a00.y, a00.y, s));
      h = fmaf(a00.x, a01.x, fmaf(a00.y, a01.y, h));
      v = fmaf(a00.x, a10.x, fmaf(a00.y, a10.y, v));
      d = fmaf(a00.x, a11.x, fmaf(a00.y, a11.y, d));
      an = fmaf(a01.x, a10.x, fmaf(a01.y, a10.y, an));
    }
    s = warp_sum(s); h = warp_sum(h); v = warp_sum(v); d = warp_sum(d); an = warp_sum(an);
    if (lane == 0) {
      float* g = gram + t * 5;
      g[0] = s; g[1] = h; g[2] = v; g[3] = d; g[4] = an;
    }
  }
}

// Weight-only constants of the fused head, from the flat fp32 state-dict parameters (bf16-rounded R,
// to match the U = tokens @ R columns the tensor core produces from bf16 operands).
// One block of 32 x 32 threads: thread (j, k) owns M[j][k]; row 0 also produces tv / b2 / w0.
__global__ void __launch_bounds__(1024)
pixel_head_consts_kernel(const float* __restrict__ p, MlpOffsets o, int dim, PixelHeadConsts* out) {
  const int k = threadIdx.x & 31, j = threadIdx.x >> 5;
  const float* w3 = p + o.w3 + kH2;  // rows 1.. of layers.4.weight: R[d][*]
  float m = 0.f, tv = 0.f, cc = 0.f;
  for (int d = 0; d < dim; ++d) {
    const float rj = __bfloat162float(__float2bfloat16_rn(w3[static_cast<long long>(d) * kH2 + j]));
    const float rk = __bfloat162float(__float2bfloat16_rn(w3[static_cast<long long>(d) * kH2 + k]));
    const float c = p[o.b3 + 1 + d];
    m = fmaf(rj, rk, m);
    if (k == 0) tv = fmaf(rj, c, tv);
    if (threadIdx.x == 0) cc = fmaf(c, c, cc);
  }
  out->m[j * kH2 + k] = (k == j) ? m : (k > j ? 2.f * m : 0.f);
  if (k == 0) {
    out->tv[j] = 2.f * tv;
    out->b2[j] = p[o.b2 + j];
    out->w0[j] = __bfloat162float(__float2bfloat16_rn(p[o.w3 + j]));  // row 0 (bf16 like the GEMM path)
  }
  if (threadIdx.x == 0) {
    out->b0 = p[o.b3];
    out->cc = cc;
  }
}

// Wcat [320, dim_p] bf16 = [W1 ; R^T ; c_hi ; c_lo ; 0], bias [320] = [b1 ; 0]
__global__ void pixel_head_pack_kernel(const float* __restrict__ p, MlpOffsets o, int dim, int dim_p,
                                       __nv_bfloat16* __restrict__ wcat, float* __restrict__ bias) {
  const long long total = static_cast<long long>(kPixelHeadN) * dim_p;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / dim_p), d = static_cast<int>(i % dim_p);
    float v = 0.f;
    if (d < dim) {
      if (r < kH1) v = p[o.w1 + static_cast<long long>(r) * dim + d];
      else if (r < kH1 + kH2) v = p[o.w3 + static_cast<long long>(1 + d) * kH2 + (r - kH1)];
      else if (r == kH1 + kH2) v = p[o.b3 + 1 + d];
      else if (r == kH1 + kH2 + 1) {
        const float c = p[o.b3 + 1 + d];
        v = c - __bfloat162float(__float2bfloat16_rn(c));
      }
    }
    wcat[i] = __float2bfloat16_rn(v);
    if (d == 0) bias[r] = r < kH1 ? p[o.b1 + r] : 0.f;
  }
}

}  // namespace

int pixel_head_supported(int h1, int h2, int gh, int gw, int H, int W) {
  if (h1 != kH1 || h2 != kH2 || H % kTileH != 0 || W % kTileW != 0 || H < 2 || W < 2) return 0;
  const float sx = static_cast<float>(gw - 1) / static_cast<float>(W - 1);
  const int ww = static_cast<int>((kTileW - 1) * sx) + 3;
  return (ww >= 2 && ww <= kWinMax) ? ww : 0;
}

int pixel_head_pack(const float* params, const MlpShape& s, int dim_p, void* wcat_bf16, float* bias,
                    PixelHeadConsts* consts, cudaStream_t stream) {
  const MlpOffsets o = mlp_offsets(s);
  pixel_head_pack_kernel<<<128, 256, 0, stream>>>(params, o, s.dim, dim_p, reinterpret_cast<__nv_bfloat16*>(wcat_bf16), bias);
  WVN_CHECK_LAUNCH("pixel_head_pack_kernel");
  pixel_head_consts_kernel<<<1, 1024, 0, stream>>>(params, o, s.dim, consts);
  WVN_CHECK_LAUNCH("pixel_head_consts_kernel");
  return WVN_OK;
}

int token_gram(const void* tok_bf16, float* gram, int batch, int gh, int gw, int dim, cudaStream_t stream) {
  WVN_REQUIRE(dim % 64 == 0, "token_gram: dim %d must be a multiple of 64", dim);
  const long long warps = static_cast<long long>(batch) * gh * gw;
  int blocks = static_cast<int>(std::min<long long>((warps * 32 + 255) / 256, static_cast<long long>(sm_count()) * 16));
  token_gram_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(tok_bf16), gram, batch, gh, gw, dim);
  WVN_CHECK_LAUNCH("token_gram_kernel");
  return WVN_OK;
}

int pixel_head(const PixelHeadArgs& a, const void* w2_bf16, int w2_ld, cudaStream_t stream) {
  WVN_REQUIRE(a.ww >= 2 && a.ww <= kWinMax && a.W % kTileW == 0 && a.H % kTileH == 0, "pixel_head: unsupported geometry");
  WVN_REQUIRE(a.ldg >= kGvStride && a.ldg % 4 == 0, "pixel_head: ldg %lld too small", a.ldg);
  CUtensorMap tw;
  WVN_PROPAGATE(make_tmap_bf16_2d(&tw, w2_bf16, kH1, kH2, static_cast<uint64_t>(w2_ld) * 2, 64, kH2));
  static bool attr_set = false;
  if (!attr_set) {
    WVN_CHECK_CUDA(cudaFuncSetAttribute(pixel_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const long long tiles = static_cast<long long>(a.batch) * (a.H / kTileH) * (a.W / kTileW);
  int grid = static_cast<int>(std::min<long long>(tiles, static_cast<long long>(sm_count()) * 2));
  pixel_head_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tw, a);
  WVN_CHECK_LAUNCH("pixel_head_kernel");
  return WVN_OK;
}

}  // namespace wvn
