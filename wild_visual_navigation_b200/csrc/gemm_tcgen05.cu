// wvn-b200: persistent warp-specialised bf16 GEMM on tcgen05 (sm_100a).
//
//   C[M,N] = A[M,K] (bf16, row-major / K-major) x W[N,K]^T (bf16, row-major / K-major)
//
// with fp32 accumulation in tensor memory and a fused epilogue.  This one kernel family
// carries every dense contraction of the WVN hot path (SURVEY.md §2.1 K1/K3/K5/K6/K8 and
// the per-pixel traversability MLP K11): patch-embed, QKV, attention out-proj, MLP fc1/fc2,
// the STEGO head and the 384->256->32 layers of the traversability MLP.
//
// Structure (one CTA per SM, persistent over output tiles, 384 threads):
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warp 1      : MMA issuer     (single thread issues tcgen05.mma, commits to mbarriers)
//   warp 2      : TMEM allocator (2 accumulator stages so epilogue(i) overlaps mainloop(i+1))
//   warps 4..11 : epilogue       (tcgen05.ld -> bias/activation/residual -> global)
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gemm.h"
#include "host_common.h"

namespace wvn {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiThreads = 256;
constexpr int kMaxSmemBytes = 227 * 1024;
constexpr int kFixedSmemBytes = 1024 /*barriers + scratch*/ + 1024 /*align slack*/;

// Tile enumeration.  Default: tile ids run n-fastest over the whole (m, n) grid and are
// dealt round-robin to CTAs (neighbouring CTAs share the A tile through L2).  ROW_OWNER
// (used by EPI_MLP_HEAD): a CTA owns whole 128-row blocks and visits their n-chunks in
// order, so per-row reductions across n-chunks stay inside one CTA.
// B_RESIDENT (runtime, K small): a CTA is pinned to one n-block whose whole [BN, K] weight slab
// stays in shared memory; it walks the m-blocks slot, slot + S, ... (S = gridDim.x / num_n), so only
// the activation tiles stream through L2 -> smem (the L2 -> SM path, ~42 B/clk/SM, is what bounds
// these skinny-K GEMMs, not the tensor pipe).
template <bool ROW_OWNER>
struct TileIter {
  int num_m, num_n, m_blk, n_blk, lin, step;
  bool bres;
  __device__ TileIter(int nm, int nn, bool b_resident) : num_m(nm), num_n(nn), m_blk(0), n_blk(0), lin(0), step(0), bres(b_resident) {
    if (bres) { n_blk = blockIdx.x % num_n; m_blk = blockIdx.x / num_n; step = gridDim.x / num_n; }
    else if (ROW_OWNER) { m_blk = blockIdx.x; n_blk = 0; }
    else { lin = blockIdx.x; m_blk = lin / num_n; n_blk = lin % num_n; }
  }
  __device__ bool valid() const { return m_blk < num_m; }
  __device__ void next() {
    if (bres) { m_blk += step; }
    else if (ROW_OWNER) { if (++n_blk == num_n) { n_blk = 0; m_blk += gridDim.x; } }
    else { lin += gridDim.x; m_blk = lin / num_n; n_blk = lin % num_n; }
  }
};

template <int BN>
struct GemmCfg {
  static constexpr uint32_t kABytes = BM * BK * 2;
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (192 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kFixedSmemBytes;
};

// GELU(x) = x * Phi(x) with the erf form's Phi, evaluated as Phi(-|x|) = 2^p(|x|) (degree-5 minimax fit of
// log2 Phi(-t) on [0, 5.5], clamped beyond): max |error| 1.5e-6 in GELU — three orders below the bf16
// rounding of the output — at 1 MUFU + ~9 FMA/ALU ops instead of erff's ~30 (the fc1 epilogue was
// issue-bound on erff).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = fminf(fabsf(x), 5.5f);
  float p = fmaf(-0.0003865310864f, t, 0.006509808358f);
  p = fmaf(p, t, -0.05048002675f);
  p = fmaf(p, t, -0.4613505006f);
  p = fmaf(p, t, -1.150225043f);
  p = fmaf(p, t, -1.00010848f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  return x * (x >= 0.f ? 1.f - e : e);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return gelu_erf_fast(v);
  return v;
}

template <int BN, int EPI, int ACT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int kMaxStages = 8;
  const bool bres = args.b_resident != 0;
  // streaming mode: STAGES x (A, B) slots; B-resident mode: the [BN, K] slab + an A-only ring
  const int STAGES = bres ? args.a_stages : Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_region = bres ? static_cast<uint32_t>(args.K / BK) * Cfg::kBBytes : static_cast<uint32_t>(STAGES) * Cfg::kBBytes;
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + b_region;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + STAGES * Cfg::kABytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kMaxStages;      // [2]           MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * kMaxStages + 2; // [2]           epilogue -> MMA
  uint64_t* b_full = bars + 2 * kMaxStages + 4;    // [1]           resident weight slab landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (args.M + BM - 1) / BM;
  const int num_n = args.N / BN;
  const int num_k = args.K / BK;
  constexpr bool ROW_OWNER = (EPI == EPI_MLP_HEAD);
  float* row_acc = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);  // [128] EPI_MLP_HEAD scratch

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kNumEpiThreads);
    }
    mbar_init(b_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (bres) {
        const int n_blk = blockIdx.x % num_n;
        mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(num_k) * Cfg::kBBytes);
        for (int kb = 0; kb < num_k; ++kb)
          tma_load_2d(&tmap_b, b_full, smem_b + kb * Cfg::kBBytes, kb * BK, n_blk * BN);
      }
      for (TileIter<ROW_OWNER> it(num_m, num_n, bres); it.valid(); it.next()) {
        const int m_blk = it.m_blk, n_blk = it.n_blk;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], bres ? Cfg::kABytes : Cfg::kStageBytes);
          tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * BK, m_blk * BM);
          if (!bres) tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (bres) mbar_wait(b_full, 0);
      for (TileIter<ROW_OWNER> it(num_m, num_n, bres); it.valid(); it.next()) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t desc_a = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t desc_b = make_sw128_kmajor_desc(smem_u32(smem_b + (bres ? kb : stage) * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in (addr>>4) units
            umma_bf16_ss(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;     // which interleaved half of the 32-col chunks
    const int row_in_tile = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    float head_partial = 0.f;
    if (EPI == EPI_MLP_HEAD && half == 0) row_acc[row_in_tile] = 0.f;
    if (EPI == EPI_MLP_HEAD) asm volatile("bar.sync 1, 256;" ::: "memory");
    for (TileIter<ROW_OWNER> it(num_m, num_n, bres); it.valid(); it.next()) {
      const int m_blk = it.m_blk, n_blk = it.n_blk;
      const int row = m_blk * BM + row_in_tile;
      const bool row_ok = row < args.M;
      // Residual epilogue: fetch this thread's slice of the residual row BEFORE waiting for the
      // accumulator, so the global-load latency hides behind the tile's MMA time.
      constexpr bool kPreloadResid = (EPI == EPI_RESID_F32) && (BN <= 192);
      constexpr int kChunksPerThread = (BN + 63) / 64;
      float4 resid_pre[kPreloadResid ? kChunksPerThread : 1][8];
      if (kPreloadResid && row_ok) {
#pragma unroll
        for (int ci = 0; ci < kChunksPerThread; ++ci) {
          const int c0 = half * 32 + 64 * ci;
          if (c0 < BN) {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(args.out) +
                                                               static_cast<long long>(row) * args.ldo + n_blk * BN + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) resid_pre[ci][j] = src[j];
          }
        }
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();

      // Per-row destination bookkeeping
      long long out_row = row;
      int tok = 0, frame = 0;
      if (EPI == EPI_PATCH) {
        frame = row / args.tokens_in;
        tok = row - frame * args.tokens_in;
        out_row = static_cast<long long>(frame) * args.npad + 1 + tok;
      } else if (EPI == EPI_QKV) {
        frame = row / args.npad;
        tok = row - frame * args.npad;
      }

#pragma unroll
      for (int chunk_i = 0; chunk_i < kChunksPerThread; ++chunk_i) {  // the two column-halves interleave 32-col chunks
        const int c0 = half * 32 + 64 * chunk_i;
        if (c0 >= BN) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c0, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (args.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(args.bias + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if (ACT != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ACT);
        }
        if (!row_ok) {
          // out-of-range tail row: nothing to store (loads above stay warp-convergent)
        } else if (EPI == EPI_BF16) {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.out) + out_row * args.ldo + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_global_v4(dst + 8 * j, pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                         pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        } else if (EPI == EPI_F32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + out_row * args.ldo + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else if (EPI == EPI_RESID_F32) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + out_row * args.ldo + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 x = kPreloadResid ? resid_pre[chunk_i][j] : dst[j];
            x.x += v[4 * j]; x.y += v[4 * j + 1]; x.z += v[4 * j + 2]; x.w += v[4 * j + 3];
            dst[j] = x;
          }
        } else if (EPI == EPI_PATCH) {
          const float4* p4 = reinterpret_cast<const float4*>(args.pos + static_cast<long long>(1 + tok) * args.ldo + col0);
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.out) + out_row * args.ldo + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 p = __ldg(p4 + j);
            dst[j] = make_float4(v[4 * j] + p.x, v[4 * j + 1] + p.y, v[4 * j + 2] + p.z, v[4 * j + 3] + p.w);
          }
        } else if (EPI == EPI_MLP_HEAD) {
          // columns [0, feat) = reconstruction of x, column trav_col = traversability logit
          if (col0 < args.feat) {
            const uint4* x4 = reinterpret_cast<const uint4*>(
                reinterpret_cast<const __nv_bfloat16*>(args.x) + static_cast<long long>(row) * args.ldx + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 xv = __ldg(x4 + j);
              const uint32_t w[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int c = col0 + 8 * j + 2 * t;
                const float d0 = v[8 * j + 2 * t] - bf16_lo(w[t]);
                const float d1 = v[8 * j + 2 * t + 1] - bf16_hi(w[t]);
                if (c < args.feat) head_partial = fmaf(d0, d0, head_partial);
                if (c + 1 < args.feat) head_partial = fmaf(d1, d1, head_partial);
              }
            }
          } else if (col0 == args.trav_col) {
            args.trav[row] = 1.f / (1.f + __expf(-v[0]));
          }
        } else if (EPI == EPI_QKV) {
          // column -> (q|k|v, head, d); a 32-column chunk never straddles a head (dh = 64)
          const int which = col0 / args.dim;
          const int within = col0 - which * args.dim;
          const int head = within >> 6;
          const int d0 = within & 63;
          const long long bh = static_cast<long long>(frame) * args.heads + head;
          if (which < 2) {
            __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(which == 0 ? args.q : args.k);
            __nv_bfloat16* dst = base + (bh * args.npad + tok) * 64 + d0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_global_v4(dst + 8 * j, pack_bf16x2(v[8 * j + 0], v[8 * j + 1]),
                           pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                           pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          } else {
            // V is stored transposed ([b, h, d, token]) so that P·V consumes it K-major.
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.vt) + (bh * 64 + d0) * args.npad + tok;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * args.npad] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      if (EPI == EPI_MLP_HEAD && n_blk == num_n - 1) {
        // combine the two column-halves of each row, then loss_reco -> confidence
        atomicAdd(&row_acc[row_in_tile], head_partial);
        head_partial = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) {
          const float loss = row_acc[row_in_tile] / static_cast<float>(args.feat);
          row_acc[row_in_tile] = 0.f;
          if (row_ok) {
            // ConfidenceGenerator.inference_without_update (utils/confidence_generator.py:182-193)
            const float mean = __ldg(args.cg_mean), sd = __ldg(args.cg_std);
            const float shifted = mean + sd * args.cg_std_factor;
            const float lo = fmaxf(shifted - sd, 0.f);
            const float hi = shifted + sd;
            const float xc = fminf(fmaxf(loss, lo), hi);
            args.conf[row] = 1.f - (xc - lo) / (hi - lo);
            if (args.loss_reco != nullptr) args.loss_reco[row] = loss;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// $WVN_GEMM_BRES: 0/unset = streaming tiles; 1 = weight-resident mode with 192-wide tiles where the
// slab fits; 2 = weight-resident with 128-wide tiles (deeper activation ring).  Experiment knob.
int bres_env_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("WVN_GEMM_BRES");
    mode = e ? atoi(e) : 0;
    if (mode < 0 || mode > 2) mode = 0;
  }
  return mode;
}

template <int BN, int EPI, int ACT>
int launch_gemm(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, EPI, ACT>;
  static bool attr_set = false;
  if (!attr_set) {
    WVN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes));
    attr_set = true;
  }
  const int num_m = (a.M + BM - 1) / BM, num_n = a.N / BN;
  const int num_tiles = (EPI == EPI_MLP_HEAD) ? num_m : num_m * num_n;
  int grid = sm_count();
  if (a.max_ctas > 0 && a.max_ctas < grid) grid = a.max_ctas;
  if (grid > num_tiles) grid = num_tiles;
  GemmArgs launch_args = a;
  launch_args.b_resident = 0;
  uint32_t smem_bytes = Cfg::kSmemBytes;
  if (EPI != EPI_MLP_HEAD && (a.allow_b_resident || bres_env_mode() > 0)) {
    const int slab = (a.K / BK) * static_cast<int>(Cfg::kBBytes);
    const int a_stages = std::min<int>(8, (kMaxSmemBytes - kFixedSmemBytes - slab) / static_cast<int>(Cfg::kABytes));
    const int per_n = grid / num_n;
    if (a_stages >= 3 && per_n >= 1 && num_m >= 2 * per_n) {
      launch_args.b_resident = 1;
      launch_args.a_stages = a_stages;
      grid = per_n * num_n;
      smem_bytes = static_cast<uint32_t>(slab + a_stages * static_cast<int>(Cfg::kABytes) + kFixedSmemBytes);
    }
  }
  prof_begin(PROF_GEMM, stream);
  kern<<<grid, kNumThreads, smem_bytes, stream>>>(ta, tb, launch_args);
  prof_end(PROF_GEMM, stream);
  WVN_CHECK_LAUNCH("gemm_bf16_kernel");
  return WVN_OK;
}

template <int BN>
int dispatch_epi(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t s) {
  switch (a.epi) {
    case EPI_BF16:
      if (a.act == ACT_NONE) return launch_gemm<BN, EPI_BF16, ACT_NONE>(a, ta, tb, s);
      if (a.act == ACT_RELU) return launch_gemm<BN, EPI_BF16, ACT_RELU>(a, ta, tb, s);
      if (a.act == ACT_GELU) return launch_gemm<BN, EPI_BF16, ACT_GELU>(a, ta, tb, s);
      break;
    case EPI_F32:
      if (a.act == ACT_NONE) return launch_gemm<BN, EPI_F32, ACT_NONE>(a, ta, tb, s);
      break;
    case EPI_RESID_F32:
      if (a.act == ACT_NONE) return launch_gemm<BN, EPI_RESID_F32, ACT_NONE>(a, ta, tb, s);
      break;
    case EPI_PATCH:
      if (a.act == ACT_NONE) return launch_gemm<BN, EPI_PATCH, ACT_NONE>(a, ta, tb, s);
      break;
    case EPI_QKV:
      if (a.act == ACT_NONE) return launch_gemm<BN, EPI_QKV, ACT_NONE>(a, ta, tb, s);
      break;
    case EPI_MLP_HEAD:
      if (a.act == ACT_NONE) return launch_gemm<BN, EPI_MLP_HEAD, ACT_NONE>(a, ta, tb, s);
      break;
  }
  return set_error(WVN_ERR_INVALID, "gemm: unsupported epilogue/activation combination (%d, %d)", a.epi, a.act);
}

}  // namespace

int pick_block_n(int N) {
  if (N % 256 == 0) return 256;
  if (N % 224 == 0) return 224;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

int gemm_bf16(const GemmArgs& a, const void* A, long long lda, const void* W, int block_n, cudaStream_t stream) {
  WVN_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem (M=%d N=%d K=%d)", a.M, a.N, a.K);
  WVN_REQUIRE(a.K % BK == 0, "gemm: K=%d must be a multiple of %d (pad the operands)", a.K, BK);
  if (block_n == 0) {
    block_n = pick_block_n(a.N);
    // prefer a 192-wide tile whose [192, K] weight slab can stay resident in shared memory
    const int mode = a.allow_b_resident ? std::max(1, bres_env_mode()) : bres_env_mode();
    if (mode == 1 && a.epi != EPI_MLP_HEAD && a.N % 192 == 0 && (a.K / BK) * 192 * 128 + 3 * 16384 + kFixedSmemBytes <= kMaxSmemBytes)
      block_n = 192;
    if (mode == 2 && a.epi != EPI_MLP_HEAD && a.N % 128 == 0 && (a.K / BK) * 128 * 128 + 3 * 16384 + kFixedSmemBytes <= kMaxSmemBytes)
      block_n = 128;
  }
  WVN_REQUIRE(block_n == 64 || block_n == 128 || block_n == 192 || block_n == 224 || block_n == 256,
              "gemm: bad block_n %d", block_n);
  WVN_REQUIRE(a.N % block_n == 0, "gemm: N=%d must be a multiple of block_n=%d (pad the weights)", a.N, block_n);
  if (a.epi == EPI_MLP_HEAD)
    WVN_REQUIRE(a.feat > 0 && a.trav_col % 32 == 0 && a.trav_col >= a.feat && a.trav_col < a.N && a.x != nullptr &&
                    a.trav != nullptr && a.conf != nullptr && a.cg_mean != nullptr && a.cg_std != nullptr &&
                    a.ldx % 8 == 0 && a.ldx >= a.trav_col,
                "gemm: bad MLP-head epilogue arguments (feat=%d trav_col=%d N=%d)", a.feat, a.trav_col, a.N);
  if (a.epi == EPI_QKV)
    WVN_REQUIRE(a.dim % 64 == 0 && a.N == 3 * a.dim && a.heads * 64 == a.dim && a.npad % 8 == 0,
                "gemm: bad QKV epilogue geometry (dim=%d heads=%d npad=%d N=%d)", a.dim, a.heads, a.npad, a.N);
  CUtensorMap ta, tb;
  WVN_PROPAGATE(make_tmap_bf16_2d(&ta, A, a.K, a.M, static_cast<uint64_t>(lda) * 2, BK, BM));
  WVN_PROPAGATE(make_tmap_bf16_2d(&tb, W, a.K, a.N, static_cast<uint64_t>(a.K) * 2, BK, block_n));
  switch (block_n) {
    case 64: return dispatch_epi<64>(a, ta, tb, stream);
    case 128: return dispatch_epi<128>(a, ta, tb, stream);
    case 192: return dispatch_epi<192>(a, ta, tb, stream);
    case 224: return dispatch_epi<224>(a, ta, tb, stream);
    case 256: return dispatch_epi<256>(a, ta, tb, stream);
  }
  return set_error(WVN_ERR_INVALID, "gemm: unreachable");
}

}  // namespace wvn
