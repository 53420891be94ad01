// wvn-b200: persistent warp-specialised bf16 GEMM on tcgen05 (sm_100a).
//
//   C[M,N] = A[M,K] (bf16, row-major / K-major) x W[N,K]^T (bf16, row-major / K-major)
//
// with fp32 accumulation in tensor memory and a fused epilogue.  This one kernel family
// carries every dense contraction of the WVN hot path (SURVEY.md §2.1 K1/K3/K5/K6/K8 and
// the per-pixel traversability MLP K11): patch-embed, QKV, attention out-proj, MLP fc1/fc2,
// the STEGO head and the 384->256->32 layers of the traversability MLP.
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles, (4 + kEpiWarps) warps):
//   warp 0      : TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx)
//   warp 1      : MMA issuer     (single thread issues tcgen05.mma, commits to mbarriers)
//   warp 2      : TMEM allocator (2 accumulator stages so epilogue(i) overlaps mainloop(i+1))
//   warps 4..   : epilogue       (tcgen05.ld -> bias/activation/residual -> global), kEpiWarps / 4 warps per
//                 TMEM lane quarter.  Measured on B200 (profiles/): with the operands streaming at full rate the
//                 mainloop alone runs the ViT GEMMs at 1.1-1.6 PFLOP/s; the epilogue is what costs — it is a
//                 chain of dependent latencies (TMEM load, polynomial GELU, smem transpose, stores), so it is
//                 spread over many warps (thread-level parallelism hides what 2 warps per scheduler could not).
//
// Tried and dropped (git history): a CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x BN tiles, each CTA's
// half of the weight slab resident in shared memory) — bit-exact, but 5-10 % slower than this kernel at every
// ViT shape even with the epilogue removed: the mainloop is not bound by L2 -> SM operand traffic.  Also dropped:
// weight-resident single-CTA tiles, cluster multicast of the weight tile, TMA-store epilogues.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gemm.h"
#include "host_common.h"

namespace wvn {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 16;                 // multiple of 4: kEpiGroups warps share each TMEM lane quarter
// Warp roles: the sub-partition arbiter prefers the highest warp id among eligible warps, so the single-thread TMA /
// MMA-issuer warps sit ABOVE the 16 epilogue warps (see attention_tcgen05.cu; -DWVN_GEMM_AUX_FIRST restores round 1's
// layout with them at ids 0 / 1 for A/B runs).
#ifdef WVN_GEMM_AUX_FIRST
constexpr int kEpiWarp0 = 4, kWarpTma = 0, kWarpMma = 1, kWarpAlloc = 2;
#else
constexpr int kEpiWarp0 = 0, kWarpTma = kEpiWarps, kWarpMma = kEpiWarps + 1, kWarpAlloc = kEpiWarps + 2;
#endif
constexpr int kEpiGroups = kEpiWarps / 4;
constexpr int kNumEpiThreads = kEpiWarps * 32;
constexpr int kNumThreads = (4 + kEpiWarps) * 32;
constexpr int kMaxSmemBytes = 227 * 1024;
constexpr int kMaxStages = 8;

// Per-warp transpose tile of the staged epilogues: 32 rows x 32 columns of the output type.
__host__ __device__ constexpr int epi_buf_bytes(int epi) {
  return (epi == EPI_F32 || epi == EPI_RESID_F32 || epi == EPI_PATCH) ? 4096 : (epi == EPI_BF16 || epi == EPI_QKV) ? 2048 : 0;
}
__host__ __device__ constexpr int fixed_smem_bytes(int epi) {
  return 1024 /*barriers + scratch*/ + kEpiWarps * epi_buf_bytes(epi) + 1024 /*align slack*/;
}

// Tile enumeration.  Default: tile ids run n-fastest over the whole (m, n)
// grid and are dealt round-robin to CTAs (neighbouring CTAs share the A tile through L2).  ROW_OWNER
// (used by EPI_MLP_HEAD): a CTA owns whole 128-row blocks and visits their n-chunks in order, so
// per-row reductions across n-chunks stay inside one CTA.
template <bool ROW_OWNER>
struct TileIter {
  int num_m, num_n, m_blk, n_blk, lin;
  __device__ TileIter(int nm, int nn) : num_m(nm), num_n(nn), m_blk(0), n_blk(0), lin(0) {
    if (ROW_OWNER) { m_blk = blockIdx.x; n_blk = 0; }
    else { lin = blockIdx.x; m_blk = lin / num_n; n_blk = lin % num_n; }
  }
  __device__ bool valid() const { return m_blk < num_m; }
  __device__ void next() {
    if (ROW_OWNER) { if (++n_blk == num_n) { n_blk = 0; m_blk += gridDim.x; } }
    else { lin += gridDim.x; m_blk = lin / num_n; n_blk = lin % num_n; }
  }
};

template <int BN, int EPI>
struct GemmCfg {
  static constexpr uint32_t kABytes = BM * BK * 2;
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kFixed = fixed_smem_bytes(EPI);
  static constexpr int kStagesRaw = (kMaxSmemBytes - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
  static constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kFixed;
};

// GELU(x) = x * Phi(x) with the erf form's Phi.  Phi(-|x|) = 2^p(|x|) (degree-5 minimax fit of log2 Phi(-t) on
// [0, 5.5], clamped beyond) and GELU(x) = max(x, 0) - t * Phi(-t), t = min(|x|, 5.5): max |error| 1.5e-6 —
// three orders below the bf16 rounding of the output.  Two elements per call so that the polynomial runs on
// packed FFMA2: 1 MUFU + ~5.5 FMA/ALU issue slots per element instead of erff's ~30 (the fc1 epilogue is
// issue-bound).
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
  const float t0 = fminf(fabsf(x0), 5.5f), t1 = fminf(fabsf(x1), 5.5f);
  const uint64_t t = pack2(t0, t1);
  uint64_t p = fma2(pack2(-0.0003865310864f, -0.0003865310864f), t, pack2(0.006509808358f, 0.006509808358f));
  p = fma2(p, t, pack2(-0.05048002675f, -0.05048002675f));
  p = fma2(p, t, pack2(-0.4613505006f, -0.4613505006f));
  p = fma2(p, t, pack2(-1.150225043f, -1.150225043f));
  p = fma2(p, t, pack2(-1.00010848f, -1.00010848f));
  float p0, p1, e0, e1;
  unpack2(p, p0, p1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(p0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(p1));
  x0 = fmaf(-t0, e0, fmaxf(x0, 0.f));
  x1 = fmaf(-t1, e1, fmaxf(x1, 0.f));
}

// Epilogue ablations exist only in `make ablate` / `make timing` builds: $WVN_GEMM_DEBUG = 1 skips the
// global stores, 2 drops the accumulators.  The shipped library has no such switch.
#ifdef WVN_GEMM_ABLATE
#define WVN_DBG_STORE && args.debug != 1
#else
#define WVN_DBG_STORE
#endif

// One 128 x BN accumulator tile: TMEM -> registers -> bias / activation -> global memory.  Called by the kEpiWarps
// epilogue warps (ewarp & 3 must equal the hardware warp's TMEM lane quarter); the kEpiGroups warps of a lane
// quarter interleave the tile's 32-column chunks.  m_blk / n_blk locate the tile in C.
template <int BN>
struct EpiChunks {
  static constexpr int kPerThread = (BN + 32 * kEpiGroups - 1) / (32 * kEpiGroups);
};

// The bias of a warp's 32-column chunks, one column per lane (a single coalesced 128-byte load per chunk).  Issued
// BEFORE the wait for the accumulator: with ~220 KB of the SM's L1 carved out as shared memory the bias vector does not
// survive in L1 between tiles, and the 8 float4 loads per lane this replaces cost ~700 clk of L2 latency per chunk on the
// epilogue's critical path (round-2 phase timing: 1400 of 4400 clk per QKV tile).
template <int BN>
__device__ __forceinline__ void epilogue_prefetch_bias(const GemmArgs& args, const int n_blk, const int ewarp, const int lane,
                                                       float (&bias_lane)[EpiChunks<BN>::kPerThread]) {
  const int group = ewarp >> 2;
#pragma unroll
  for (int chunk_i = 0; chunk_i < EpiChunks<BN>::kPerThread; ++chunk_i) {
    const int c0 = 32 * (group + kEpiGroups * chunk_i);
    bias_lane[chunk_i] = (args.bias != nullptr && c0 < BN) ? __ldg(args.bias + n_blk * BN + c0 + lane) : 0.f;
  }
}

template <int BN, int EPI, int ACT>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& args, const uint32_t tmem_acc, const int m_blk, const int n_blk,
                                              const int ewarp, const int lane, uint8_t* epi_stage, float& head_partial,
                                              const float (&bias_lane)[EpiChunks<BN>::kPerThread]) {
  const int quarter = ewarp & 3;  // TMEM lane quarter this warp may access
  const int group = ewarp >> 2;   // which of the interleaved 32-col chunk sets
  const int row_in_tile = quarter * 32 + lane;
  const int row = m_blk * BM + row_in_tile;
  const bool row_ok = row < args.M;
  constexpr int kChunksPerThread = EpiChunks<BN>::kPerThread;

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

#ifdef WVN_GEMM_ABLATE
  if (args.debug == 2) return;
#endif
#ifdef WVN_GEMM_TIMING
  const bool etiming = args.timing != nullptr && blockIdx.x == 0 && ewarp == 0 && lane == 0;
  long long eprev = clock64();
#define WVN_ETM(i) if (etiming) { const long long tn = clock64(); args.timing[8 + i] += tn - eprev; eprev = tn; }
#else
#define WVN_ETM(i)
#endif
#pragma unroll
  for (int chunk_i = 0; chunk_i < kChunksPerThread; ++chunk_i) {
    const int c0 = 32 * (group + kEpiGroups * chunk_i);
    if (c0 >= BN) break;
    const int col0 = n_blk * BN + c0;
    uint32_t r[32];
    tmem_ld32(tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + c0, r);
    // While the accumulator load is in flight: the chunk's 32 bias values (one per lane, prefetched before the wait for
    // the accumulator) go through the warp's staging tile and come back to every lane as 8 broadcast 16-byte reads.
    constexpr bool kBiasViaSmem = epi_buf_bytes(EPI) >= 128;   // epilogues without a staging tile broadcast by shuffle
    uint4 bq[8];
    if (kBiasViaSmem && args.bias != nullptr) {
      const uint32_t bbuf = smem_u32(epi_stage) + ewarp * epi_buf_bytes(EPI);
      __syncwarp();  // the previous chunk's read-back of the staging tile is complete
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bbuf + lane * 4), "f"(bias_lane[chunk_i]) : "memory");
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) bq[j] = lds128(bbuf + 16 * j);
    }
    tmem_ld_wait();
    tmem_ld_fence32(r);
    WVN_ETM(0)
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (!kBiasViaSmem && args.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane[chunk_i], j);
    }
    if (kBiasViaSmem && args.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a0, a1, a2, a3;
        unpack2(add2(pack2(v[4 * j], v[4 * j + 1]), pack2(__uint_as_float(bq[j].x), __uint_as_float(bq[j].y))), a0, a1);
        unpack2(add2(pack2(v[4 * j + 2], v[4 * j + 3]), pack2(__uint_as_float(bq[j].z), __uint_as_float(bq[j].w))), a2, a3);
        v[4 * j] = a0; v[4 * j + 1] = a1; v[4 * j + 2] = a2; v[4 * j + 3] = a3;
      }
    }
    WVN_ETM(1)
    if (ACT == ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (ACT == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
    }
    WVN_ETM(2)
    // ---- coalesced stores: the 32x32 chunk (lane = row) is transposed through a swizzled shared-
    // memory tile so that each store instruction writes whole row segments (lanes along columns).
    // Per-thread row-wise stores were LSU-wavefront-bound — every 16-byte piece of a lane lies in a
    // different 128-byte line (measured 7.6k-19k clk/tile of epilogue against 2.3k clk of MMA) — and
    // TMA stores queue behind the mainloop's in-flight TMA loads.  The residual stream is updated
    // with vector reductions (red.global.add.v4.f32): x += acc + bias happens at L2, no load.
    constexpr bool kStagedEpi = (EPI == EPI_BF16 || EPI == EPI_F32 || EPI == EPI_RESID_F32 || EPI == EPI_QKV || EPI == EPI_PATCH);
    bool qkv_is_v = false;
    if (EPI == EPI_QKV) qkv_is_v = (col0 >= 2 * args.dim);
    if (kStagedEpi && !qkv_is_v) {
      const uint32_t buf = smem_u32(epi_stage) + ewarp * epi_buf_bytes(EPI);
      const int row_base = m_blk * BM + quarter * 32;
      __syncwarp();  // the previous chunk's read-back is complete
      if (EPI == EPI_F32 || EPI == EPI_RESID_F32 || EPI == EPI_PATCH) {
#pragma unroll
        for (int q = 0; q < 8; ++q)  // 128-byte rows: 16-byte chunk q of row r lives at q ^ (r & 7)
          sts128(buf + lane * 128 + ((q ^ (lane & 7)) << 4), __float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]),
                 __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
        __syncwarp();
        float* outp = reinterpret_cast<float*>(args.out);
        const int q = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {  // 4 rows x 128 B per instruction
          const int r = it * 4 + (lane >> 3);
          const uint4 val = lds128(buf + r * 128 + ((q ^ (r & 7)) << 4));
          if (row_base + r < args.M WVN_DBG_STORE) {
            if (EPI == EPI_PATCH) {
              // patch row -> token row (frame * npad + 1 + token) of the residual stream, plus the position embedding
              // (round 1 stored these rows lane by lane: 134 us for the patch-embed GEMM, LSU-wavefront bound)
              const int grow = row_base + r;
              const int fr = grow / args.tokens_in, tk = grow - fr * args.tokens_in;
              const float4 pe = __ldg(reinterpret_cast<const float4*>(args.pos + static_cast<long long>(1 + tk) * args.ldo + col0) + q);
              float4 o4 = make_float4(__uint_as_float(val.x) + pe.x, __uint_as_float(val.y) + pe.y,
                                      __uint_as_float(val.z) + pe.z, __uint_as_float(val.w) + pe.w);
              *reinterpret_cast<float4*>(outp + (static_cast<long long>(fr) * args.npad + 1 + tk) * args.ldo + col0 + q * 4) = o4;
              continue;
            }
            float* dst = outp + static_cast<long long>(row_base + r) * args.ldo + col0 + q * 4;
            if (EPI == EPI_RESID_F32) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(val.x)),
                           "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                           : "memory");
            } else {
              *reinterpret_cast<uint4*>(dst) = val;
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)  // 64-byte rows: chunk q of row r lives at q ^ ((r >> 1) & 3)
          sts128(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4), pack_bf16x2(v[8 * q + 0], v[8 * q + 1]),
                 pack_bf16x2(v[8 * q + 2], v[8 * q + 3]), pack_bf16x2(v[8 * q + 4], v[8 * q + 5]),
                 pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        __syncwarp();
        __nv_bfloat16* base;
        long long row_stride, first;
        if (EPI == EPI_BF16) {
          base = reinterpret_cast<__nv_bfloat16*>(args.out);
          row_stride = args.ldo;
          first = static_cast<long long>(row_base) * args.ldo + col0;
        } else {  // Q / K: [b*h, npad, 64]
          const int which = col0 / args.dim;
          const int within = col0 - which * args.dim;
          const int fr = row_base / args.npad, tk0 = row_base - fr * args.npad;
          base = reinterpret_cast<__nv_bfloat16*>(which == 0 ? args.q : args.k);
          row_stride = 64;
          first = ((static_cast<long long>(fr) * args.heads + (within >> 6)) * args.npad + tk0) * 64 + (within & 63);
        }
        const int q = lane & 3;
#pragma unroll
        for (int it = 0; it < 4; ++it) {  // 8 rows x 64 B per instruction
          const int r = it * 8 + (lane >> 2);
          const uint4 val = lds128(buf + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
          if (row_base + r < args.M WVN_DBG_STORE) *reinterpret_cast<uint4*>(base + first + r * row_stride + q * 8) = val;
        }
      }
    }
    WVN_ETM(3)
    if (!row_ok) {
      // out-of-range tail row: no per-row work below
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
    } else if (EPI == EPI_QKV && qkv_is_v) {
      // V is stored transposed ([b, h, d, token]) so that P·V consumes it K-major: lanes already run
      // along tokens, so these direct stores are 64-byte coalesced per instruction.
      const int within = col0 - 2 * args.dim;
      const long long bh = static_cast<long long>(frame) * args.heads + (within >> 6);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.vt) + (bh * 64 + (within & 63)) * args.npad + tok;
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * args.npad] = __float2bfloat16_rn(v[j]);
    }
  }
}
#ifdef WVN_GEMM_TIMING
#define WVN_TM_DECL
#define WVN_TM(i) if (timing) { const long long tn = clock64(); tm[i] += tn - tprev; tprev = tn; }
#define WVN_TM_TILE ++ntiles;
#define WVN_TM_FLUSH if (timing) { for (int i = 0; i < 3; ++i) args.timing[i] = tm[i]; args.timing[3] = ntiles; }
#else
#define WVN_TM_DECL
#define WVN_TM(i)
#define WVN_TM_TILE
#define WVN_TM_FLUSH
#endif

// ------------------------------------------------------------------------------------------------
// Single-CTA kernel: 128 x BN tiles, operands streamed through a kStages-deep TMA ring.
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI, int ACT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmArgs args) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + STAGES * Cfg::kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + STAGES * Cfg::kABytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kMaxStages;      // [2]           MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * kMaxStages + 2; // [2]           epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
  float* row_acc = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);  // [128] EPI_MLP_HEAD scratch
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(bars) + 1024;         // [8 warps][4 KB] epilogue transpose tiles

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (args.M + BM - 1) / BM;
  const int num_n = args.N / BN;
  const int num_k = args.K / BK;
  constexpr bool ROW_OWNER = (EPI == EPI_MLP_HEAD);

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kWarpAlloc) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kWarpTma) {
    // ------------------------------------------------------------------ TMA producer (whole warp in the loop, one
    // elected lane issues — see elect_one_sync in common.cuh)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (TileIter<ROW_OWNER> it(num_m, num_n); it.valid(); it.next()) {
        const int m_eff = args.reverse_m ? num_m - 1 - it.m_blk : it.m_blk;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * BK, m_eff * BM);
            tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * BK, it.n_blk * BN);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ MMA issuer (whole warp waits, one elected
    // lane issues the tcgen05.mma / commit instructions)
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
#ifdef WVN_GEMM_TIMING
      const bool timing = args.timing != nullptr && blockIdx.x == 0 && lane == 0;
      long long tm[4] = {0, 0, 0, 0}, tprev = clock64(), ntiles = 0;
#endif
      for (TileIter<ROW_OWNER> it(num_m, num_n); it.valid(); it.next()) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        WVN_TM(0)
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          WVN_TM(1)
          if (elect_one_sync()) {
            const uint64_t desc_a = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::kABytes));
            const uint64_t desc_b = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in (addr>>4) units
              umma_bf16_ss(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);                     // smem slot reusable once these MMAs retire
            if (kb == num_k - 1) umma_commit(&acc_full[acc]);   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          WVN_TM(2)
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        WVN_TM_TILE
      }
      WVN_TM_FLUSH
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue
    const int ewarp = warp - kEpiWarp0;
    const int row_in_tile = (ewarp & 3) * 32 + lane;
    const int group = ewarp >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    float head_partial = 0.f;
#ifdef WVN_GEMM_TIMING
    long long e_wait = 0, e_work = 0;
#endif
    if (EPI == EPI_MLP_HEAD && group == 0) row_acc[row_in_tile] = 0.f;
    if (EPI == EPI_MLP_HEAD) named_bar_sync(1, kNumEpiThreads);
    for (TileIter<ROW_OWNER> it(num_m, num_n); it.valid(); it.next()) {
#ifdef WVN_GEMM_TIMING
      long long et0 = clock64();
#endif
      float bias_lane[EpiChunks<BN>::kPerThread];
      epilogue_prefetch_bias<BN>(args, it.n_blk, ewarp, lane, bias_lane);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
#ifdef WVN_GEMM_TIMING
      long long et1 = clock64();
#endif
      const int m_eff = args.reverse_m ? num_m - 1 - it.m_blk : it.m_blk;
      epilogue_tile<BN, EPI, ACT>(args, tmem_base + acc * BN, m_eff, it.n_blk, ewarp, lane, epi_stage, head_partial,
                                  bias_lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#ifdef WVN_GEMM_TIMING
      e_wait += et1 - et0; e_work += clock64() - et1;
#endif

      if (EPI == EPI_MLP_HEAD && it.n_blk == num_n - 1) {
        // combine the column groups of each row, then loss_reco -> confidence
        const int row = m_eff * BM + row_in_tile;
        atomicAdd(&row_acc[row_in_tile], head_partial);
        head_partial = 0.f;
        named_bar_sync(1, kNumEpiThreads);
        if (group == 0) {
          const float loss = row_acc[row_in_tile] / static_cast<float>(args.feat);
          row_acc[row_in_tile] = 0.f;
          if (row < args.M) {
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
        named_bar_sync(1, kNumEpiThreads);
      }
    }
#ifdef WVN_GEMM_TIMING
    if (args.timing != nullptr && blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32) { args.timing[4] = e_wait; args.timing[5] = e_work; }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int EPI, int ACT>
int launch_gemm(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  static_assert(Cfg::kStages >= 2, "tile too wide for the shared-memory budget");
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
  prof_begin(PROF_GEMM, stream);
  kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, a);
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
  if (block_n == 0) block_n = pick_block_n(a.N);
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
  if (a.epi == EPI_F32 || a.epi == EPI_RESID_F32) WVN_REQUIRE(a.ldo % 4 == 0, "gemm: fp32 output pitch must be a multiple of 4");
  if (a.epi == EPI_BF16) WVN_REQUIRE(a.ldo % 8 == 0, "gemm: bf16 output pitch must be a multiple of 8");
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
