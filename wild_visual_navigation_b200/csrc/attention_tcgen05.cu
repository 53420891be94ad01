// wvn-b200: fused non-causal multi-head attention (flash-style) on tcgen05, head dim 64.
//
// Replaces the materialised `softmax(q @ k^T * scale) @ v` of the DINO ViT blocks
// (SURVEY.md §8 a3 / K4): per (frame, head) the 3137x3137 (ViT-S/8 @448) score matrix is
// never written to HBM; S lives in tensor memory, P goes through shared memory straight
// back into the tensor core, O accumulates in tensor memory.
//
// Inputs (written by the QKV GEMM epilogue, bf16):
//   Q, K : [B*H, npad, 64]   row-major (K-major for the MMA)
//   V^T  : [B*H, 64, npad]   row-major (so P·V also sees a K-major B operand)
// Output: O [B, npad, H*64] bf16 (the layout the out-projection GEMM reads).
//
// One CTA = one 128-row query tile of one (frame, head); 2 CTAs are co-resident per SM so
// the tensor core works on one CTA's MMAs while the other CTA is in its softmax phase.
//   warp 0    : TMEM alloc, then TMA producer (Q once; K / V^T tiles, 2 stages each)
//   warp 1    : MMA issuer  (S = Q K^T : 4 x UMMA 128x128x16;  O += P V : 8 x UMMA 128x64x16)
//   warps 2-5 : softmax (1 thread = 1 query row): tcgen05.ld S, online softmax with lazy
//               rescaling, P -> bf16 -> swizzled smem, final O / l epilogue.
#include <stdlib.h>

#include "attention.h"
#include "common.cuh"
#include "host_common.h"

namespace wvn {

namespace {

constexpr int kThreads = 192;
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;
constexpr int kDh = 64;
constexpr uint32_t kQBytes = kTileQ * kDh * 2;       // 16 KB
constexpr uint32_t kKBytes = kTileKV * kDh * 2;      // 16 KB
constexpr uint32_t kVBytes = kDh * kTileKV * 2;      // 16 KB (two 8 KB K-blocks)
constexpr uint32_t kPBytes = kTileQ * kTileKV * 2;   // 32 KB (two 16 KB K-blocks)
constexpr int kStages = 2;
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + kQBytes;
constexpr uint32_t kOffV = kOffK + kStages * kKBytes;
constexpr uint32_t kOffP = kOffV + kStages * kVBytes;
constexpr uint32_t kOffBar = kOffP + kPBytes;
constexpr uint32_t kSmemBytes = kOffBar + 256;
constexpr uint32_t kTmemCols = 256;  // S: [0,128)  O: [128,192)
constexpr uint32_t kColS = 0;
constexpr uint32_t kColO = 128;
constexpr uint32_t kColMail = 192;  // 6 mailbox columns: [parity 0/1][half 0/1] row maxima, then [half] row sums
constexpr float kRescaleThreshold = 8.0f;  // in log2 units (FA4-style lazy rescale)
constexpr int kDefaultSplit = 1;
constexpr int kDefaultPoly = 0;            // software-exp2 share: pairs out of every 4 pairs (see poly_exp2_pair)

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 helpers (sm_100: FFMA2 / FADD2 / FMNMX3 halve the issue slots of the softmax) ----
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// exp2 on the FMA/ALU pipes (Cody-Waite + degree-3 minimax, rel. err 7.5e-5 — well inside the bf16
// rounding of P): the MUFU unit (16 ex2/clk/SM) is the binding resource of the softmax phase, so
// POLY of every 4 element PAIRS are evaluated here instead (FA4-style software exp2 offload).
__device__ __forceinline__ void poly_exp2_pair(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x2 = pack2(fmaxf(x0, -120.f), fmaxf(x1, -120.f));      // keep the exponent arithmetic in range
  const uint64_t t2 = add2(x2, pack2(12582912.f, 12582912.f));     // 1.5*2^23: low mantissa bits = round(x)
  const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), x2);             // f = x - round(x) in [-0.5, 0.5]
  uint64_t p2 = fma2(f2, pack2(0.0551716685f, 0.0551716685f), pack2(0.2426111251f, 0.2426111251f));
  p2 = fma2(p2, f2, pack2(0.6932609677f, 0.6932609677f));
  p2 = fma2(p2, f2, pack2(0.9999280572f, 0.9999280572f));
  float p0, p1, t0, t1;
  unpack2(p2, p0, p1);
  unpack2(t2, t0, t1);
  e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));   // p * 2^round(x)
  e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

__device__ __forceinline__ float masked_exp(float s, float sl2, float mb, int col, int valid) {
  return (col < valid) ? fast_exp2(fmaf(s, sl2, -mb)) : 0.f;
}

// SPLIT = threads per query row in the softmax phase.  SPLIT == 2 gives each row to two threads of
// different warps (64 key columns each): twice the softmax warps per SM to hide the TMEM-load /
// MUFU / barrier latencies, half the registers per thread; the two halves agree on the reference
// max through a one-column mailbox in tensor memory (shared memory is full at 2 CTAs/SM).
template <int POLY, int SPLIT>
__global__ void __launch_bounds__(64 + 128 * SPLIT, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_vt, const AttnArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* s_free = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int nkv = args.npad / kTileKV;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("[wvn] attention: dynamic smem base not 1024B aligned\n");
    __trap();
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128 * SPLIT);
    mbar_init(p_full, 128 * SPLIT);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // -------------------------------------------------------------- TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_vt);
      const int row0 = bh * args.npad;
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_2d(&tmap_q, q_full, smem + kOffQ, 0, row0 + q_tile * kTileQ);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kKBytes);
        tma_load_2d(&tmap_k, &k_full[st], smem + kOffK + st * kKBytes, 0, row0 + j * kTileKV);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kVBytes);
        tma_load_2d(&tmap_vt, &v_full[st], smem + kOffV + st * kVBytes, j * kTileKV, bh * kDh);
        tma_load_2d(&tmap_vt, &v_full[st], smem + kOffV + st * kVBytes + kVBytes / 2, j * kTileKV + 64, bh * kDh);
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV);  // 128 x 128
      constexpr uint32_t idesc_o = make_idesc_bf16(kTileQ, kDh);      // 128 x 64
      const uint32_t tmem_s = tmem_base + kColS;
      const uint32_t tmem_o = tmem_base + kColO;
      const uint64_t desc_q = make_sw128_kmajor_desc(smem_u32(smem + kOffQ));

      auto issue_qk = [&](int j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[st], ph);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);  // softmax has drained S(j-1) from TMEM
        tc_fence_after();
        const uint64_t desc_k = make_sw128_kmajor_desc(smem_u32(smem + kOffK + st * kKBytes));
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) umma_bf16_ss(tmem_s, desc_q + 2 * k, desc_k + 2 * k, idesc_s, k != 0);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      };

      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_qk(j + 1);  // overlaps softmax(j)'s tail and P(j) hand-off
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(smem + kOffP);
        const uint32_t v_addr = smem_u32(smem + kOffV + st * kVBytes);
#pragma unroll
        for (int ks = 0; ks < kTileKV / 16; ++ks) {
          const uint64_t desc_p = make_sw128_kmajor_desc(p_addr + (ks >> 2) * (kPBytes / 2)) + 2 * (ks & 3);
          const uint64_t desc_v = make_sw128_kmajor_desc(v_addr + (ks >> 2) * (kVBytes / 2)) + 2 * (ks & 3);
          umma_bf16_ss(tmem_o, desc_p, desc_v, idesc_o, (j | ks) != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else if (SPLIT == 2) {
    // -------------------------------------------------------------- softmax, two threads per row
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;       // which 64 key columns of every tile (= P's K-block)
    const int row = quarter * 32 + lane;    // row inside the query tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + kColS + half * 64;
    const uint32_t tmem_o = tmem_base + lane_base + kColO + half * 32;
    const uint32_t tmem_mail = tmem_base + lane_base + kColMail;  // [parity][half] mailbox columns
    uint8_t* p_row = smem + kOffP + half * (kPBytes / 2) + row * 128;
    const int sw = row & 7;
    const float sl2 = args.scale_log2;
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int valid = args.n_valid - j * kTileKV - half * 64;  // valid columns of this half
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld32(tmem_s, sr[0]);
      tmem_ld32(tmem_s + 32, sr[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);

      const bool masked = valid < 64;
      float mx;
      if (!masked) {
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = max3(m0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
          m1 = max3(m1, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
        }
        mx = fmaxf(m0, m1);
      } else {
        mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < valid) ? __uint_as_float(sr[c][i]) : -INFINITY);
      }
      // exchange the half-row maxima with the partner thread (same row, other half) through TMEM
      const uint32_t mail = tmem_mail + (j & 1) * 2;
      tmem_st1(mail + half, __float_as_uint(mx));
      tmem_st_wait();
      tc_fence_before();
      named_bar_sync(1 + quarter, 64);
      tc_fence_after();
      mx = fmaxf(mx, __uint_as_float(tmem_ld1(mail + (half ^ 1))));
      tmem_ld_wait();

      bool waited_pv = false;
      if (j == 0) {
        m_ref = mx;
      } else {
        const float m_new = fmaxf(m_ref, mx);
        const bool need = (m_new - m_ref) * sl2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(pv_done, (j - 1) & 1);
          waited_pv = true;
          tc_fence_after();
          const float alpha = need ? fast_exp2((m_ref - m_new) * sl2) : 1.f;
          if (need) m_ref = m_new;
          l *= alpha;
          uint32_t r[32];
          tmem_ld32(tmem_o, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_o, r);
          tmem_st_wait();
          tc_fence_before();
        }
      }

      const float mb = m_ref * sl2;
      if (!masked) {
        const uint64_t sl2_2 = pack2(sl2, sl2), nmb2 = pack2(-mb, -mb);
        uint64_t la = pack2(0.f, 0.f), lb = pack2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t x2 = fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), sl2_2, nmb2);
            float e0, e1;
            if (POLY == 9) {  // timing experiment only: no exponentials at all (results are wrong)
              unpack2(x2, e0, e1);
            } else if (((i >> 1) & 3) < POLY) {
              poly_exp2_pair(x2, e0, e1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              e0 = fast_exp2(x0);
              e1 = fast_exp2(x1);
            }
            if ((i >> 1) & 1) lb = add2(lb, pack2(e0, e1)); else la = add2(la, pack2(e0, e1));
            sr[c][i >> 1] = pack_bf16x2(e0, e1);
          }
        }
        float s0, s1;
        unpack2(add2(la, lb), s0, s1);
        l += s0 + s1;
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = masked_exp(__uint_as_float(sr[c][i]), sl2, mb, c * 32 + i, valid);
            const float e1 = masked_exp(__uint_as_float(sr[c][i + 1]), sl2, mb, c * 32 + i + 1, valid);
            l += e0 + e1;
            sr[c][i >> 1] = pack_bf16x2(e0, e1);
          }
        }
      }

      if (j > 0 && !waited_pv) mbar_wait(pv_done, (j - 1) & 1);
      // this half's 64 columns are exactly K-block `half` of P: 8 swizzled 16-byte chunks
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c * 4 + q) ^ sw;
          *reinterpret_cast<uint4*>(p_row + chunk * 16) =
              make_uint4(sr[c][4 * q + 0], sr[c][4 * q + 1], sr[c][4 * q + 2], sr[c][4 * q + 3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }

    // ---- epilogue: combine the two partial row sums, O / l for this half's 32 output columns
    mbar_wait(pv_done, (nkv - 1) & 1);
    tc_fence_after();
    tmem_st1(tmem_mail + 4 + half, __float_as_uint(l));
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(1 + quarter, 64);
    tc_fence_after();
    l += __uint_as_float(tmem_ld1(tmem_mail + 4 + (half ^ 1)));
    tmem_ld_wait();
    const float inv_l = 1.f / l;
    const int b = bh / args.heads;
    const int h = bh - b * args.heads;
    const long long q_idx = static_cast<long long>(b) * args.npad + q_tile * kTileQ + row;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.out) + q_idx * args.ldo + h * kDh + half * 32;
    uint32_t r[32];
    tmem_ld32(tmem_o, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[8 * q + i]) * inv_l;
      st_global_v4(dst + 8 * q, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                   pack_bf16x2(f[6], f[7]));
    }
  } else {
    // -------------------------------------------------------------- softmax / correction / epilogue
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // row inside the query tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + kColS;
    const uint32_t tmem_o = tmem_base + lane_base + kColO;
    uint8_t* p_row = smem + kOffP + row * 128;
    const int sw = row & 7;
    const float sl2 = args.scale_log2;

    float m_ref = -INFINITY;  // running reference max (raw score units)
    float l = 0.f;            // running sum of exp2((s - m_ref) * sl2)

    for (int j = 0; j < nkv; ++j) {
      const int valid = args.n_valid - j * kTileKV;  // columns >= valid are padding tokens
      mbar_wait(s_full, j & 1);
      tc_fence_after();

      // ---- single TMEM pass: the whole 128-wide score row of this thread goes to registers
      const bool masked = valid < kTileKV;  // only the last KV tile carries padding keys
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + c * 32, sr[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);  // S(j) is in registers: QK^T(j+1) may overwrite it while we do the exps

      float mx;
      if (!masked) {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;  // FMNMX3 chains
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = max3(m0, __uint_as_float(sr[0][i]), __uint_as_float(sr[0][i + 1]));
          m1 = max3(m1, __uint_as_float(sr[1][i]), __uint_as_float(sr[1][i + 1]));
          m2 = max3(m2, __uint_as_float(sr[2][i]), __uint_as_float(sr[2][i + 1]));
          m3 = max3(m3, __uint_as_float(sr[3][i]), __uint_as_float(sr[3][i + 1]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
        mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c * 32 + i < valid) ? __uint_as_float(sr[c][i]) : -INFINITY);
      }

      // ---- reference-max update (lazy: only rescale O when the max grew by > 2^8)
      bool waited_pv = false;
      if (j == 0) {
        m_ref = mx;
      } else {
        const float m_new = fmaxf(m_ref, mx);
        const bool need = (m_new - m_ref) * sl2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(pv_done, (j - 1) & 1);  // O must be quiescent
          waited_pv = true;
          tc_fence_after();
          const float alpha = need ? fast_exp2((m_ref - m_new) * sl2) : 1.f;
          if (need) m_ref = m_new;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_o + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(tmem_o + c * 32, r);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }

      // ---- P = exp2((s - m_ref) * sl2), in place in the score registers; row sum
      const float mb = m_ref * sl2;
      if (!masked) {
        const uint64_t sl2_2 = pack2(sl2, sl2), nmb2 = pack2(-mb, -mb);
        uint64_t la = pack2(0.f, 0.f), lb = pack2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t x2 = fma2(pack2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), sl2_2, nmb2);
            float e0, e1;
            if (POLY == 9) {  // timing experiment only: no exponentials at all (results are wrong)
              unpack2(x2, e0, e1);
            } else if (((i >> 1) & 3) < POLY) {
              poly_exp2_pair(x2, e0, e1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              e0 = fast_exp2(x0);
              e1 = fast_exp2(x1);
            }
            if ((i >> 1) & 1) lb = add2(lb, pack2(e0, e1)); else la = add2(la, pack2(e0, e1));
            sr[c][i >> 1] = pack_bf16x2(e0, e1);  // packed bf16 pairs overwrite the consumed scores
          }
        }
        float s0, s1;
        unpack2(add2(la, lb), s0, s1);
        l += s0 + s1;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = masked_exp(__uint_as_float(sr[c][i]), sl2, mb, c * 32 + i, valid);
            const float e1 = masked_exp(__uint_as_float(sr[c][i + 1]), sl2, mb, c * 32 + i + 1, valid);
            l += e0 + e1;
            sr[c][i >> 1] = pack_bf16x2(e0, e1);
          }
        }
      }

      if (j > 0 && !waited_pv) mbar_wait(pv_done, (j - 1) & 1);  // P buffer free again (PV(j-1) retired)
      // ---- P -> swizzled smem: 32 columns = 4 x 16-byte chunks of K-block (c >> 1), chunk (c & 1) * 4 + q
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint8_t* blk = p_row + (c >> 1) * (kPBytes / 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ sw;
          *reinterpret_cast<uint4*>(blk + chunk * 16) =
              make_uint4(sr[c][4 * q + 0], sr[c][4 * q + 1], sr[c][4 * q + 2], sr[c][4 * q + 3]);
        }
      }
      fence_proxy_async_smem();   // P(j) visible to the tensor core (async proxy)
      mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> out[b, q, h*64 + d]
    mbar_wait(pv_done, (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const int b = bh / args.heads;
    const int h = bh - b * args.heads;
    const long long q_idx = static_cast<long long>(b) * args.npad + q_tile * kTileQ + row;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.out) + q_idx * args.ldo + h * kDh;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_o + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[8 * q + i]) * inv_l;
        st_global_v4(dst + c * 32 + 8 * q, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                     pack_bf16x2(f[6], f[7]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int attention_bf16(const AttnArgs& a, const void* q, const void* k, const void* vt, cudaStream_t stream) {
  WVN_REQUIRE(a.batch > 0 && a.heads > 0, "attention: empty problem");
  WVN_REQUIRE(a.npad % kTileKV == 0 && a.n_valid > 0 && a.n_valid <= a.npad && a.n_valid > a.npad - kTileKV,
              "attention: npad=%d must be a multiple of 128 and n_valid=%d must lie in the last tile", a.npad,
              a.n_valid);
  const long long bh = static_cast<long long>(a.batch) * a.heads;
  CUtensorMap tq, tk, tv;
  WVN_PROPAGATE(make_tmap_bf16_2d(&tq, q, kDh, bh * a.npad, kDh * 2, 64, kTileQ));
  WVN_PROPAGATE(make_tmap_bf16_2d(&tk, k, kDh, bh * a.npad, kDh * 2, 64, kTileKV));
  WVN_PROPAGATE(make_tmap_bf16_2d(&tv, vt, a.npad, bh * kDh, static_cast<uint64_t>(a.npad) * 2, 64, kDh));
  // fraction (in quarters) of the exponentials evaluated on the FMA pipe; tuned on B200, overridable
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("WVN_ATTN_POLY");
    poly = e ? atoi(e) : kDefaultPoly;
    if ((poly < 0 || poly > 3) && poly != 9) poly = kDefaultPoly;
  }
  static int split = -1;
  if (split < 0) {
    const char* e = getenv("WVN_ATTN_SPLIT");
    split = e ? atoi(e) : kDefaultSplit;
    if (split != 1 && split != 2) split = kDefaultSplit;
  }
  dim3 grid(a.npad / kTileQ, static_cast<unsigned>(bh));
  auto launch = [&](auto kern, int threads) -> int {
    WVN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    prof_begin(PROF_ATTENTION, stream);
    kern<<<grid, threads, kSmemBytes, stream>>>(tq, tk, tv, a);
    prof_end(PROF_ATTENTION, stream);
    return WVN_OK;
  };
  if (split == 2) {
    switch (poly) {
      case 0: WVN_PROPAGATE(launch(attention_kernel<0, 2>, 320)); break;
      case 1: WVN_PROPAGATE(launch(attention_kernel<1, 2>, 320)); break;
      default: WVN_PROPAGATE(launch(attention_kernel<2, 2>, 320)); break;
    }
  } else {
    switch (poly) {
      case 0: WVN_PROPAGATE(launch(attention_kernel<0, 1>, 192)); break;
      case 1: WVN_PROPAGATE(launch(attention_kernel<1, 1>, 192)); break;
      case 9: WVN_PROPAGATE(launch(attention_kernel<9, 1>, 192)); break;
      default: WVN_PROPAGATE(launch(attention_kernel<2, 1>, 192)); break;
    }
  }
  WVN_CHECK_LAUNCH("attention_kernel");
  return WVN_OK;
}

}  // namespace wvn
