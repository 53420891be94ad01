// wvn-b200: internal interface of the tcgen05 GEMM family (see gemm_tcgen05.cu).
#pragma once

#include <cuda_runtime.h>

namespace wvn {

enum GemmEpilogue : int {
  EPI_BF16 = 0,       // out(bf16)[row, col]  = act(acc + bias)
  EPI_F32 = 1,        // out(f32)[row, col]   = acc + bias
  EPI_RESID_F32 = 2,  // out(f32)[row, col] += acc + bias            (residual stream, in place)
  EPI_PATCH = 3,      // out(f32)[frame*npad + 1 + tok, col] = acc + bias + pos[1 + tok, col]
  EPI_QKV = 4,        // scatter to Q/K [b,h,npad,64] and V^T [b,h,64,npad] (bf16)
  EPI_MLP_HEAD = 5,   // last layer of the traversability MLP fused with its consumers:
                      //   cols [0,feat) reconstruct x -> loss_reco = mean((out - x)^2) -> confidence,
                      //   col feat = traversability logit -> sigmoid.  Nothing of [M,N] is stored.
};

enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  int epi = EPI_BF16, act = ACT_NONE;
  const float* bias = nullptr;  // [N] fp32 or null
  void* out = nullptr;          // primary output (see GemmEpilogue)
  long long ldo = 0;            // leading dimension of `out` (elements)
  // EPI_PATCH
  const float* pos = nullptr;   // [1 + tokens_in, ldo] positional embedding (row 0 = CLS)
  int tokens_in = 0;            // patches per frame (rows of A per frame)
  // EPI_PATCH / EPI_QKV
  int npad = 0;                 // padded tokens per frame in the activation layout
  // EPI_QKV
  int dim = 0, heads = 0;
  void* q = nullptr;
  void* k = nullptr;
  void* vt = nullptr;
  // EPI_MLP_HEAD
  int feat = 0;                 // feature dimension D: output columns [0, feat) reconstruct x
  int trav_col = 0;             // output column holding the traversability logit (multiple of 32, >= feat)
  const void* x = nullptr;      // [M, ldx] bf16 MLP input rows (reconstruction target)
  long long ldx = 0;
  float* trav = nullptr;        // [M]
  float* conf = nullptr;        // [M]
  float* loss_reco = nullptr;   // [M] optional
  const float* cg_mean = nullptr;  // device scalars of the ConfidenceGenerator
  const float* cg_std = nullptr;
  float cg_std_factor = 0.5f;
  // launch control
  int max_ctas = 0;             // 0 = one CTA per SM
  int reverse_m = 0;            // walk the m-blocks last-to-first (start on the rows the producer kernel wrote last: L2 hits)
  int debug = 0;                // debug (-DWVN_GEMM_TIMING builds): 1 = skip the global stores, 2 = skip the epilogue body
  long long* timing = nullptr;  // debug (-DWVN_GEMM_TIMING builds): phase cycle counters of CTA 0
};

int pick_block_n(int N);

// A: [M, K] bf16 with row pitch lda (elements); W: [N, K] bf16 contiguous.
int gemm_bf16(const GemmArgs& args, const void* A, long long lda, const void* W, int block_n, cudaStream_t stream);

}  // namespace wvn
