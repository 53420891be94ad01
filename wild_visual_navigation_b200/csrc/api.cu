// wvn-b200: the C ABI (include/wvn_b200.h) — handles, composite forward passes, primitives.
#include <cuda_bf16.h>

#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/wvn_b200.h"
#include "attention.h"
#include "dense_kernels.h"
#include "gemm.h"
#include "host_common.h"
#include "mlp_train.h"
#include "mlp_train_fused.h"
#include "pixel_head.h"
#include "segment_kernels.h"
#include "footprint_kernels.h"
#include "slic_kernels.h"
#include "stego_kmeans.h"
#include "vit_kernels.h"

using namespace wvn;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// fp32 rows [rows, dim] -> bf16 rows [rows, ld] (padding columns left untouched = zero)
__global__ void cast_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long rows, int dim,
                                 long long ld) {
  const long long n = rows * dim;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dim;
    const int c = static_cast<int>(i - r * dim);
    dst[r * ld + c] = __float2bfloat16_rn(src[i]);
  }
}

int cast_to_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
  if (n <= 0) return WVN_OK;
  int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 4096));
  cast_f32_to_bf16_kernel<<<blocks, 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  WVN_CHECK_LAUNCH("cast_f32_to_bf16_kernel");
  return WVN_OK;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n) {
    bytes = n;
    WVN_CHECK_CUDA(cudaMalloc(&p, n ? n : 1));
    WVN_CHECK_CUDA(cudaMemset(p, 0, n ? n : 1));
    return WVN_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
  }
};

struct Weight {
  void* p = nullptr;   // device storage (fp32 or bf16)
  long long numel = 0;
  bool bf16 = false;
  bool loaded = false;
};

}  // namespace

// ============================================================================================
// ViT handle
// ============================================================================================
struct wvn_vit {
  wvn_vit_config cfg;
  int grid = 0, P = 0, n_valid = 0, npad = 0, kpe = 0, chunk = 0;
  std::map<std::string, Weight> w;
  std::vector<DevBuf> owned;
  // workspaces (per chunk)
  DevBuf x, xn, q, k, vt, attn, hid, ape, stage;
  // per max_batch
  DevBuf tok_bf16, head_hidden;
  DevBuf qkv_f32;        // only with $WVN_VIT_PRECISE=1 at create: fp32 QKV projections of one chunk (parity-debug attention)
  bool precise = false;
  bool forwarded = false;
  int last_batch = 0;

  int add_weight(const std::string& name, long long numel, bool bf16) {
    DevBuf b;
    WVN_PROPAGATE(b.alloc(static_cast<size_t>(numel) * (bf16 ? 2 : 4)));
    owned.push_back(b);
    Weight wt;
    wt.p = b.p; wt.numel = numel; wt.bf16 = bf16;
    w[name] = wt;
    return WVN_OK;
  }
  template <class T>
  T* wp(const std::string& name) { return reinterpret_cast<T*>(w[name].p); }
};

extern "C" {

const char* wvn_last_error(void) { return last_error(); }
int wvn_version(void) { return 100; }

int wvn_check_device(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return set_error(WVN_ERR_NO_DEVICE, "no CUDA device visible");
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return set_error(WVN_ERR_NO_DEVICE, "device %d has compute capability %d.x, need 10.x (sm_100a)", dev, major);
  return WVN_OK;
}

long long wvn_launch_count(void) { return launch_count(); }
void wvn_profile_enable(int category_mask) { prof_enable(category_mask); }
int wvn_profile_collect(float* host_ms, long long* host_launches) {
  WVN_REQUIRE(host_ms && host_launches, "wvn_profile_collect: null argument");
  return prof_collect(host_ms, host_launches);
}

// -------------------------------------------------------------------------------- primitives
int wvn_gemm_bf16(const void* a, long long lda, const void* w, const float* bias, void* out, long long ldo, int m, int n,
                  int k, int out_kind, int act, int block_n, void* stream) {
  GemmArgs g;
  g.M = m; g.N = n; g.K = k;
  g.epi = out_kind == 0 ? EPI_BF16 : (out_kind == 1 ? EPI_F32 : EPI_RESID_F32);
  WVN_REQUIRE(out_kind >= 0 && out_kind <= 2, "wvn_gemm_bf16: out_kind %d", out_kind);
  g.act = act; g.bias = bias; g.out = out; g.ldo = ldo;
#ifdef WVN_GEMM_ABLATE
  {
    static int dbg = -1;  // $WVN_GEMM_DEBUG: epilogue ablations (1 = no stores, 2 = no epilogue), ablation builds only
    if (dbg < 0) { const char* e = getenv("WVN_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
    g.debug = dbg;
  }
#endif
#ifdef WVN_GEMM_TIMING
  static long long* tbuf = nullptr;  // device memory: the probes must not fault on managed pages
  if (!tbuf) cudaMalloc(&tbuf, 16 * sizeof(long long));
  cudaMemsetAsync(tbuf, 0, 16 * sizeof(long long), S(stream));
  g.timing = tbuf;
  const int rc = gemm_bf16(g, a, lda, w, block_n, S(stream));
  long long t[16];
  cudaMemcpyAsync(t, tbuf, sizeof(t), cudaMemcpyDeviceToHost, S(stream));
  cudaStreamSynchronize(S(stream));
  const long long nt = t[3] > 0 ? t[3] : 1;
  fprintf(stderr, "[gemm timing M=%d N=%d K=%d kind=%d act=%d, cycles per tile of CTA 0 (%lld tiles)] mma: wait_acc_empty %lld  "
          "wait_full(TMA) %lld  issue+commit %lld | epilogue: wait_acc_full %lld  work %lld = ldtm %lld  bias %lld  act %lld  "
          "transpose+store %lld\n",
          m, n, k, out_kind, act, nt, t[0] / nt, t[1] / nt, t[2] / nt, t[4] / nt, t[5] / nt, t[8] / nt, t[9] / nt, t[10] / nt,
          t[11] / nt);
  return rc;
#else
  return gemm_bf16(g, a, lda, w, block_n, S(stream));
#endif
}

int wvn_attention_bf16(const void* q, const void* k, const void* vt, void* out, int batch, int heads, int npad,
                       int n_valid, float scale, void* stream) {
  AttnArgs a;
  a.batch = batch; a.heads = heads; a.npad = npad; a.n_valid = n_valid;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.out = out; a.ldo = static_cast<long long>(heads) * 64;
  static long long* timing_buf = nullptr;  // $WVN_ATTN_TIMING=1: phase cycle counters, printed after each call
  static int timing_on = -1;
  if (timing_on < 0) {
    const char* e = getenv("WVN_ATTN_TIMING");
    timing_on = (e && e[0] == '1') ? 1 : 0;
    if (timing_on) cudaMallocManaged(&timing_buf, 16 * sizeof(long long));
  }
  a.timing = timing_on ? timing_buf : nullptr;
  const int rc = attention_bf16(a, q, k, vt, S(stream));
  if (timing_on && rc == 0) {
    cudaStreamSynchronize(S(stream));
    const int nkv = npad / 128;
    const char* impl = getenv("WVN_ATTN_IMPL");
    if (impl && atoi(impl) == 1)
      fprintf(stderr, "[attn v1 timing, cycles per KV tile] softmax: wait_s %lld  ldtm %lld  math %lld  wait_pv %lld  store+fence %lld | "
              "mma: issue_qk(+waits) %lld  wait_p %lld  wait_v+issue_pv %lld  loop %lld\n",
              timing_buf[0] / nkv, timing_buf[1] / nkv, timing_buf[2] / nkv, timing_buf[3] / nkv, timing_buf[4] / nkv,
              timing_buf[8] / nkv, timing_buf[9] / nkv, timing_buf[10] / nkv, timing_buf[11] / nkv);
    else if (impl == nullptr || atoi(impl) == 5)
      fprintf(stderr, "[attn v5 timing, cycles per 64-key tile, thread 0] wait_s %lld  ldtm %lld  max+rescale %lld  exp %lld  store %lld\n",
              timing_buf[0] / ((n_valid + 63) / 64), timing_buf[1] / ((n_valid + 63) / 64), timing_buf[2] / ((n_valid + 63) / 64),
              timing_buf[3] / ((n_valid + 63) / 64), timing_buf[4] / ((n_valid + 63) / 64));
    else if (atoi(impl) == 3)
      fprintf(stderr, "[attn v3 timing, cycles per KV tile, thread 0] wait_s %lld  ldtm %lld  max+mailbox %lld  rescale %lld  exp %lld  "
              "probe+wait_pv+store %lld\n", timing_buf[0] / nkv, timing_buf[1] / nkv, timing_buf[2] / nkv, timing_buf[5] / nkv,
              timing_buf[3] / nkv, timing_buf[4] / nkv);
    else
      fprintf(stderr, "[attn v2 timing, cycles per KV tile] softmax wg0: wait_s %lld  ldtm %lld  max+rescale %lld  token_wait %lld  exp %lld  "
              "wait_pv+store %lld | mma: wait_v+wait_p0 %lld  pv0+wait_sfree0+qk0 %lld  wait_p1 %lld  pv1+qk1+loop %lld\n",
              timing_buf[0] / nkv, timing_buf[1] / nkv, timing_buf[2] / nkv, timing_buf[5] / nkv, timing_buf[3] / nkv,
              timing_buf[4] / nkv, timing_buf[8] / nkv, timing_buf[9] / nkv, timing_buf[10] / nkv, timing_buf[11] / nkv);
  }
  return rc;
}

int wvn_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, long long rows, int dim,
                  float eps, void* stream) {
  LayerNormArgs a;
  a.rows = rows; a.dim = dim; a.eps = eps;
  return layernorm_rows(x, gamma, beta, out_bf16, nullptr, a, S(stream));
}

// -------------------------------------------------------------------------------- ViT
int wvn_vit_create(const wvn_vit_config* cfg, wvn_vit_t** out) {
  WVN_REQUIRE(cfg && out, "wvn_vit_create: null argument");
  WVN_PROPAGATE(wvn_check_device());
  WVN_REQUIRE(cfg->dim == 384 || cfg->dim == 768, "vit: dim %d unsupported (384, 768)", cfg->dim);
  WVN_REQUIRE(cfg->heads * 64 == cfg->dim, "vit: heads*64 must equal dim");
  WVN_REQUIRE(cfg->patch_size == 8 || cfg->patch_size == 16, "vit: patch size %d unsupported", cfg->patch_size);
  WVN_REQUIRE(cfg->image_size >= cfg->patch_size, "vit: image size too small");
  WVN_REQUIRE(cfg->mlp_dim % 64 == 0 && cfg->depth > 0 && cfg->max_batch > 0, "vit: bad mlp_dim/depth/max_batch");
  WVN_REQUIRE(cfg->head_out % 64 == 0, "vit: head_out must be a multiple of 64");
  wvn_vit* h = new wvn_vit();
  h->cfg = *cfg;
  if (h->cfg.ln_eps <= 0.f) h->cfg.ln_eps = 1e-6f;
  h->grid = cfg->image_size / cfg->patch_size;  // conv-floor semantics for non-divisible sizes
  h->P = h->grid * h->grid;
  h->n_valid = h->P + 1;
  h->npad = round_up(h->n_valid, 128);
  h->kpe = 3 * cfg->patch_size * cfg->patch_size;
  h->chunk = cfg->chunk > 0 ? cfg->chunk : 8;
  if (h->chunk > cfg->max_batch) h->chunk = cfg->max_batch;
  const int D = cfg->dim;
  int rc = WVN_OK;
  auto add = [&](const std::string& n, long long numel, bool bf) { if (rc == WVN_OK) rc = h->add_weight(n, numel, bf); };
  add("cls_token", D, false);
  add("pos_embed", static_cast<long long>(h->n_valid) * D, false);
  add("patch_embed.proj.weight", static_cast<long long>(D) * h->kpe, true);
  add("patch_embed.proj.bias", D, false);
  for (int i = 0; i < cfg->depth; ++i) {
    const std::string b = "blocks." + std::to_string(i) + ".";
    add(b + "norm1.weight", D, false);
    add(b + "norm1.bias", D, false);
    add(b + "attn.qkv.weight", 3ll * D * D, true);
    add(b + "attn.qkv.bias", 3 * D, false);
    add(b + "attn.proj.weight", static_cast<long long>(D) * D, true);
    add(b + "attn.proj.bias", D, false);
    add(b + "norm2.weight", D, false);
    add(b + "norm2.bias", D, false);
    add(b + "mlp.fc1.weight", static_cast<long long>(cfg->mlp_dim) * D, true);
    add(b + "mlp.fc1.bias", cfg->mlp_dim, false);
    add(b + "mlp.fc2.weight", static_cast<long long>(D) * cfg->mlp_dim, true);
    add(b + "mlp.fc2.bias", D, false);
  }
  add("norm.weight", D, false);
  add("norm.bias", D, false);
  if (cfg->head_out > 0) {
    add("stego.head_a.weight", static_cast<long long>(cfg->head_out) * D, true);
    add("stego.head_a.bias", cfg->head_out, false);
    add("stego.hidden.weight", static_cast<long long>(D) * D, true);
    add("stego.hidden.bias", D, false);
    add("stego.head_b.weight", static_cast<long long>(cfg->head_out) * D, true);
  }
  const size_t rows = static_cast<size_t>(h->chunk) * h->npad;
  const size_t bh = static_cast<size_t>(h->chunk) * cfg->heads;
  auto alloc = [&](DevBuf& b, size_t bytes) { if (rc == WVN_OK) rc = b.alloc(bytes); };
  alloc(h->x, rows * D * 4);
  alloc(h->xn, rows * D * 2);
  alloc(h->q, bh * h->npad * 64 * 2);
  alloc(h->k, bh * h->npad * 64 * 2);
  alloc(h->vt, bh * 64 * h->npad * 2);
  alloc(h->attn, rows * D * 2);
  alloc(h->hid, rows * cfg->mlp_dim * 2);
  alloc(h->ape, static_cast<size_t>(h->chunk) * h->P * h->kpe * 2);
  alloc(h->stage, 8u << 20);
  alloc(h->tok_bf16, static_cast<size_t>(cfg->max_batch) * h->npad * D * 2);
  if (cfg->head_out > 0) alloc(h->head_hidden, static_cast<size_t>(cfg->max_batch) * h->npad * D * 2);
  {
    // parity-debug mode: Q K^T, softmax and P V in fp32 on fp32 projections, ~40x slower attention
    const char* e = getenv("WVN_VIT_PRECISE");
    h->precise = e && atoi(e) == 1;
    if (h->precise) alloc(h->qkv_f32, rows * 3 * D * 4);
  }
  if (rc != WVN_OK) {
    wvn_vit_destroy(h);
    return rc;
  }
  *out = h;
  return WVN_OK;
}

void wvn_vit_destroy(wvn_vit_t* h) {
  if (!h) return;
  for (auto& b : h->owned) b.release();
  for (DevBuf* b : {&h->x, &h->xn, &h->q, &h->k, &h->vt, &h->attn, &h->hid, &h->ape, &h->stage, &h->tok_bf16,
                    &h->head_hidden, &h->qkv_f32})
    b->release();
  delete h;
}

int wvn_vit_npad(const wvn_vit_t* h) { return h ? h->npad : 0; }

int wvn_vit_set_weight(wvn_vit_t* h, const char* name, const float* data, long long numel) {
  WVN_REQUIRE(h && name && data, "wvn_vit_set_weight: null argument");
  auto it = h->w.find(name);
  WVN_REQUIRE(it != h->w.end(), "wvn_vit_set_weight: unknown weight '%s'", name);
  Weight& wt = it->second;
  WVN_REQUIRE(wt.numel == numel, "wvn_vit_set_weight: '%s' expects %lld elements, got %lld", name, wt.numel, numel);
  cudaPointerAttributes attr;
  bool on_device = false;
  if (cudaPointerGetAttributes(&attr, data) == cudaSuccess)
    on_device = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  else
    cudaGetLastError();
  // fp32 destination: copy straight in; bf16 destination: stage (if host) + cast on device
  if (!wt.bf16) {
    WVN_CHECK_CUDA(cudaMemcpy(wt.p, data, numel * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  } else {
    const long long chunk_elems = static_cast<long long>(h->stage.bytes / 4);
    for (long long off = 0; off < numel; off += chunk_elems) {
      const long long n = std::min(chunk_elems, numel - off);
      const float* src = data + off;
      if (!on_device) {
        WVN_CHECK_CUDA(cudaMemcpy(h->stage.p, src, n * 4, cudaMemcpyHostToDevice));
        src = reinterpret_cast<const float*>(h->stage.p);
      }
      WVN_PROPAGATE(cast_to_bf16(src, reinterpret_cast<__nv_bfloat16*>(wt.p) + off, n, 0));
      WVN_CHECK_CUDA(cudaStreamSynchronize(0));
    }
  }
  wt.loaded = true;
  return WVN_OK;
}

static int vit_check_loaded(wvn_vit* h, bool need_head) {
  for (auto& kv : h->w) {
    const bool is_head = kv.first.rfind("stego.", 0) == 0;
    if (is_head && !need_head) continue;
    if (!kv.second.loaded) return set_error(WVN_ERR_STATE, "vit: weight '%s' was never set", kv.first.c_str());
  }
  return WVN_OK;
}

constexpr int kDefaultSubAttn = 1 << 30;  // frames per (LN1, QKV, attention) pass; tuned on B200

namespace {
int vit_forward_impl(wvn_vit_t* h, const void* img, bool u8_hwc, int batch, int in_h, int in_w, int resized_h, int resized_w,
                     float* tokens_out, void* stream, bool flip_tta = false);
}

int wvn_vit_forward_tta(wvn_vit_t* h, const float* img, int batch, int in_h, int in_w, int resized_h, int resized_w,
                        float* tokens_out, void* stream) {
  return vit_forward_impl(h, img, false, batch, in_h, in_w, resized_h, resized_w, tokens_out, stream, true);
}

int wvn_vit_forward(wvn_vit_t* h, const float* img, int batch, int in_h, int in_w, int resized_h, int resized_w,
                    float* tokens_out, void* stream) {
  return vit_forward_impl(h, img, false, batch, in_h, in_w, resized_h, resized_w, tokens_out, stream);
}

int wvn_vit_forward_u8(wvn_vit_t* h, const unsigned char* img_hwc, int batch, int in_h, int in_w, int resized_h,
                       int resized_w, float* tokens_out, void* stream) {
  return vit_forward_impl(h, img_hwc, true, batch, in_h, in_w, resized_h, resized_w, tokens_out, stream);
}

namespace {
// flip_tta: `batch` source frames are run twice — frames [batch, 2*batch) of the activation layout / tokens_out are
// the backbone's output on the horizontally flipped TRANSFORMED images (Stego.get_code's second pass).
int vit_forward_impl(wvn_vit_t* h, const void* img, bool u8_hwc, int src_batch, int in_h, int in_w, int resized_h, int resized_w,
                     float* tokens_out, void* stream, bool flip_tta) {
  const int batch = flip_tta ? 2 * src_batch : src_batch;
  WVN_REQUIRE(h && img, "wvn_vit_forward: null argument");
  WVN_REQUIRE(batch > 0 && batch <= h->cfg.max_batch, "wvn_vit_forward: batch %d outside (0, %d]", batch, h->cfg.max_batch);
  WVN_REQUIRE(resized_h >= h->cfg.image_size && resized_w >= h->cfg.image_size,
              "wvn_vit_forward: resized image %dx%d smaller than the crop %d", resized_h, resized_w, h->cfg.image_size);
  WVN_PROPAGATE(vit_check_loaded(h, false));
  cudaStream_t s = S(stream);
  const wvn_vit_config& c = h->cfg;
  const int D = c.dim;

  ImagePatchArgs ia;
  ia.in_h = in_h; ia.in_w = in_w; ia.patch = c.patch_size; ia.grid_h = h->grid; ia.grid_w = h->grid;
  // torchvision CenterCrop: top = int(round((H - size) / 2.0))
  ia.crop_top = static_cast<int>(lrintf((resized_h - c.image_size) / 2.0f));
  ia.crop_left = static_cast<int>(lrintf((resized_w - c.image_size) / 2.0f));
  ia.scale_y = static_cast<float>(in_h) / static_cast<float>(resized_h);
  ia.scale_x = static_cast<float>(in_w) / static_cast<float>(resized_w);

  for (int b0 = 0; b0 < batch; b0 += h->chunk) {
    const int nb = std::min(h->chunk, batch - b0);
    const int rows = nb * h->npad;
    float* x = reinterpret_cast<float*>(h->x.p);
    ia.batch = nb;
    ia.frame0 = b0; ia.src_frames = src_batch; ia.flip_from = flip_tta ? src_batch : (1 << 30);
    WVN_PROPAGATE(image_to_patches(img, u8_hwc, h->ape.p, ia, s));
    WVN_PROPAGATE(init_token_rows(x, h->wp<float>("cls_token"), h->wp<float>("pos_embed"), nb, h->npad, h->n_valid, D, s));
    {
      GemmArgs g;
      g.M = nb * h->P; g.N = D; g.K = h->kpe; g.epi = EPI_PATCH; g.bias = h->wp<float>("patch_embed.proj.bias");
      g.out = x; g.ldo = D; g.pos = h->wp<float>("pos_embed"); g.tokens_in = h->P; g.npad = h->npad;
      WVN_PROPAGATE(gemm_bf16(g, h->ape.p, h->kpe, h->wp<void>("patch_embed.proj.weight"), 0, s));
    }
    LayerNormArgs la;
    la.rows = rows; la.dim = D; la.eps = c.ln_eps; la.npad = h->npad; la.n_valid = h->n_valid;
    // GEMMs / LayerNorms can run over sub-chunks of `sub` frames ($WVN_VIT_SUBCHUNK) to keep xn / hid / x
    // L2-resident between producer and consumer.  Measured on B200 (round 1): sub-chunking LOSES (1752 -> 1719
    // frames/s at sub = 16): what the smaller GEMMs lose to wave quantisation (the N = 384 GEMMs drop to 5.4 waves)
    // outweighs the L2 hits, which the snake order below already collects for the hottest ~100 MB — default off.
    static int sub_env = -1;
    if (sub_env < 0) {
      const char* e = getenv("WVN_VIT_SUBCHUNK");
      sub_env = e ? atoi(e) : (1 << 30);
      if (sub_env < 1) sub_env = 1 << 30;
    }
    const int sub = std::min(sub_env, nb);
    // "Snake" order: every kernel of the chain walks its rows in the opposite direction to its producer, so it
    // starts on the rows that were written last and are still in the 126 MB L2 ($WVN_VIT_SNAKE=0 disables).
    static int snake = -1;
    if (snake < 0) { const char* e = getenv("WVN_VIT_SNAKE"); snake = (e && atoi(e) == 0) ? 0 : 1; }
    int dir = 0;  // the patch-embed GEMM above ran first-to-last
    auto next_dir = [&]() { dir = snake ? dir ^ 1 : 0; return dir; };
    __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(h->xn.p);
    __nv_bfloat16* attn = reinterpret_cast<__nv_bfloat16*>(h->attn.p);
    // The attention half of a block (LN1 -> QKV -> attention) can additionally run over `sub_a` frames at a time
    // ($WVN_VIT_SUB_ATTN): Q / K / V^T of a half-chunk (118 MB at 16 frames) are consumed while still in L2.
    static int sub_attn_env = -1;
    if (sub_attn_env < 0) {
      const char* e = getenv("WVN_VIT_SUB_ATTN");
      sub_attn_env = e ? atoi(e) : kDefaultSubAttn;
      if (sub_attn_env < 1) sub_attn_env = 1 << 30;
    }
    const int sub_a = std::min(sub_attn_env, nb);   // independent of `sub`: attention wants the whole chunk (8.1 waves of CTAs)
    for (int l = 0; l < c.depth; ++l) {
      const std::string b = "blocks." + std::to_string(l) + ".";
      for (int s0 = 0; s0 < nb; s0 += sub_a) {
        const int ns = std::min(sub_a, nb - s0);
        const long long roff = static_cast<long long>(s0) * h->npad;
        LayerNormArgs ls = la;
        ls.rows = static_cast<long long>(ns) * h->npad;
        ls.reverse = next_dir();
        WVN_PROPAGATE(layernorm_rows(x + roff * D, h->wp<float>(b + "norm1.weight"), h->wp<float>(b + "norm1.bias"),
                                     xn + roff * D, nullptr, ls, s));
        if (h->precise) {
          GemmArgs g;
          float* qkv = reinterpret_cast<float*>(h->qkv_f32.p) + roff * 3 * D;
          g.M = ns * h->npad; g.N = 3 * D; g.K = D; g.epi = EPI_F32; g.bias = h->wp<float>(b + "attn.qkv.bias");
          g.out = qkv; g.ldo = 3 * D;
          WVN_PROPAGATE(gemm_bf16(g, xn + roff * D, D, h->wp<void>(b + "attn.qkv.weight"), 0, s));
          WVN_PROPAGATE(attention_f32_debug(qkv, attn + roff * D, ns, c.heads, h->npad, h->n_valid, D, 0.125f, s));
          continue;
        }
        GemmArgs g;
        g.M = ns * h->npad; g.N = 3 * D; g.K = D; g.epi = EPI_QKV; g.bias = h->wp<float>(b + "attn.qkv.bias");
        g.npad = h->npad; g.dim = D; g.heads = c.heads;
        const long long hoff = static_cast<long long>(s0) * c.heads * h->npad * 64;  // frames are outermost in q / k / vt
        g.q = reinterpret_cast<__nv_bfloat16*>(h->q.p) + hoff;
        g.k = reinterpret_cast<__nv_bfloat16*>(h->k.p) + hoff;
        g.vt = reinterpret_cast<__nv_bfloat16*>(h->vt.p) + hoff;
        g.reverse_m = next_dir();
        WVN_PROPAGATE(gemm_bf16(g, xn + roff * D, D, h->wp<void>(b + "attn.qkv.weight"), 0, s));
        AttnArgs a;
        a.batch = ns; a.heads = c.heads; a.npad = h->npad; a.n_valid = h->n_valid;
        a.scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim 64: 64^-0.5 * log2(e)
        a.out = attn + roff * D; a.ldo = D;
        a.reverse = next_dir();
        WVN_PROPAGATE(attention_bf16(a, g.q, g.k, g.vt, s));
      }
      for (int s0 = 0; s0 < nb; s0 += sub) {
        const int ns = std::min(sub, nb - s0);
        const long long roff = static_cast<long long>(s0) * h->npad;
        const int srows = ns * h->npad;
        LayerNormArgs ls = la;
        ls.rows = srows;
        {
          GemmArgs g;
          g.M = srows; g.N = D; g.K = D; g.epi = EPI_RESID_F32; g.bias = h->wp<float>(b + "attn.proj.bias");
          g.out = x + roff * D; g.ldo = D;
          g.reverse_m = next_dir();
          WVN_PROPAGATE(gemm_bf16(g, attn + roff * D, D, h->wp<void>(b + "attn.proj.weight"), 0, s));
        }
        ls.reverse = next_dir();
        WVN_PROPAGATE(layernorm_rows(x + roff * D, h->wp<float>(b + "norm2.weight"), h->wp<float>(b + "norm2.bias"),
                                     xn + roff * D, nullptr, ls, s));
        {
          GemmArgs g;
          g.M = srows; g.N = c.mlp_dim; g.K = D; g.epi = EPI_BF16; g.act = ACT_GELU;
          g.bias = h->wp<float>(b + "mlp.fc1.bias"); g.out = h->hid.p; g.ldo = c.mlp_dim;  // hid is reused per sub-chunk
          g.reverse_m = next_dir();
          WVN_PROPAGATE(gemm_bf16(g, xn + roff * D, D, h->wp<void>(b + "mlp.fc1.weight"), 0, s));
        }
        {
          GemmArgs g;
          g.M = srows; g.N = D; g.K = c.mlp_dim; g.epi = EPI_RESID_F32; g.bias = h->wp<float>(b + "mlp.fc2.bias");
          g.out = x + roff * D; g.ldo = D;
          g.reverse_m = next_dir();
          WVN_PROPAGATE(gemm_bf16(g, h->hid.p, c.mlp_dim, h->wp<void>(b + "mlp.fc2.weight"), 0, s));
        }
      }
    }
    __nv_bfloat16* tok_bf = reinterpret_cast<__nv_bfloat16*>(h->tok_bf16.p) + static_cast<long long>(b0) * h->npad * D;
    float* tok_f = tokens_out ? tokens_out + static_cast<long long>(b0) * h->P * D : nullptr;
    la.reverse = next_dir();
    WVN_PROPAGATE(layernorm_rows(x, h->wp<float>("norm.weight"), h->wp<float>("norm.bias"), tok_bf, tok_f, la, s));
  }
  h->forwarded = true;
  h->last_batch = batch;
  return WVN_OK;
}
}  // namespace

int wvn_vit_stego_head(wvn_vit_t* h, int batch, float* out, void* stream) {
  WVN_REQUIRE(h && out, "wvn_vit_stego_head: null argument");
  WVN_REQUIRE(h->cfg.head_out > 0, "wvn_vit_stego_head: handle was created without a head");
  WVN_REQUIRE(h->forwarded && batch == h->last_batch, "wvn_vit_stego_head: call wvn_vit_forward with the same batch first");
  WVN_PROPAGATE(vit_check_loaded(h, true));
  cudaStream_t s = S(stream);
  const int D = h->cfg.dim, rows = batch * h->npad, HO = h->cfg.head_out;
  GemmArgs g;
  g.M = rows; g.N = D; g.K = D; g.epi = EPI_BF16; g.act = ACT_RELU; g.bias = h->wp<float>("stego.hidden.bias");
  g.out = h->head_hidden.p; g.ldo = D;
  WVN_PROPAGATE(gemm_bf16(g, h->tok_bf16.p, D, h->wp<void>("stego.hidden.weight"), 0, s));
  GemmArgs a;
  a.M = rows; a.N = HO; a.K = D; a.epi = EPI_F32; a.bias = h->wp<float>("stego.head_a.bias"); a.out = out; a.ldo = HO;
  WVN_PROPAGATE(gemm_bf16(a, h->tok_bf16.p, D, h->wp<void>("stego.head_a.weight"), 0, s));
  GemmArgs b;
  b.M = rows; b.N = HO; b.K = D; b.epi = EPI_RESID_F32; b.bias = nullptr; b.out = out; b.ldo = HO;
  WVN_PROPAGATE(gemm_bf16(b, h->head_hidden.p, D, h->wp<void>("stego.head_b.weight"), 0, s));
  return WVN_OK;
}

// -------------------------------------------------------------------------------- dense
int wvn_upsample_dense(const float* tokens, float* out, int batch, int dim, int gh, int gw, int out_h, int out_w,
                       void* stream) {
  DenseArgs a;
  a.batch = batch; a.dim = dim; a.grid_h = gh; a.grid_w = gw; a.out_h = out_h; a.out_w = out_w;
  a.scale_y = out_h > 1 ? static_cast<float>(gh - 1) / static_cast<float>(out_h - 1) : 0.f;
  a.scale_x = out_w > 1 ? static_cast<float>(gw - 1) / static_cast<float>(out_w - 1) : 0.f;
  return upsample_tokens_dense(tokens, out, a, S(stream));
}

int wvn_logits_argmax(const float* logits, long long ld, int col0, int classes, int col0_b, int classes_b, int batch,
                      int npad, int gh, int gw, int out_h, int out_w, long long* seg, long long* seg_b, void* stream) {
  LogitsArgs a;
  a.batch = batch; a.classes = classes; a.grid_h = gh; a.grid_w = gw; a.out_h = out_h; a.out_w = out_w;
  a.scale_y = static_cast<float>(gh) / static_cast<float>(out_h);
  a.scale_x = static_cast<float>(gw) / static_cast<float>(out_w);
  a.npad = npad; a.ld = ld; a.col0 = col0; a.col0_b = col0_b; a.classes_b = classes_b;
  return logits_argmax(logits, seg, seg_b, a, S(stream));
}

int wvn_flip_average(float* head, int batch, int npad, int grid, long long ld, void* stream) {
  return flip_average(head, batch, npad, grid, ld, S(stream));
}

size_t wvn_stego_kmeans_workspace_bytes(int batch, int k, int code_dim) { return stego_kmeans_workspace_bytes(batch, k, code_dim); }

int wvn_stego_kmeans(float* rows, long long ld, int batch, int npad, int patches, int code_col, int code_dim, int logit_col,
                     int k, int iters, float* centroids_out, void* workspace, void* stream) {
  KmeansArgs a;
  a.batch = batch; a.npad = npad; a.patches = patches; a.ld = ld; a.code_col = code_col; a.code_dim = code_dim;
  a.logit_col = logit_col; a.k = k; a.iters = iters; a.centroids_out = centroids_out;
  return stego_kmeans(rows, a, reinterpret_cast<float*>(workspace), S(stream));
}

// -------------------------------------------------------------------------------- segments
static void seg_ws_layout(int batch, int smax, int gh, int gw, size_t& off_w, size_t& off_adj, size_t& total) {
  const size_t stats = static_cast<size_t>(batch) * smax * 3 * sizeof(unsigned long long);
  off_w = (stats + 255) / 256 * 256;
  const size_t wbytes = static_cast<size_t>(batch) * smax * gh * gw * sizeof(float);
  off_adj = off_w + (wbytes + 255) / 256 * 256;
  total = off_adj + static_cast<size_t>(batch) * smax * ((smax + 31) / 32) * sizeof(unsigned int);
}

size_t wvn_segment_workspace_bytes(int batch, int smax, int gh, int gw) {
  size_t a, b, t;
  seg_ws_layout(batch, smax, gh, gw, a, b, t);
  return t;
}

int wvn_segment_reduce(const long long* seg, int batch, int h, int w, int smax, const float* tokens, int gh, int gw,
                       int dim, float* feat, float* centers, long long* edges, int* n_edges, int max_edges,
                       void* workspace, void* stream) {
  WVN_REQUIRE(seg && workspace, "wvn_segment_reduce: null argument");
  WVN_REQUIRE(feat == nullptr || tokens != nullptr, "wvn_segment_reduce: feat requested without tokens");
  WVN_REQUIRE((edges == nullptr) == (n_edges == nullptr), "wvn_segment_reduce: edges and n_edges go together");
  // dense features exist on (h, h) only (dino_interface.py:87-88): pixels with x >= h have no feature — the reference
  // raises an index error there, so do we
  WVN_REQUIRE(feat == nullptr || w <= h, "wvn_segment_reduce: segment map %dx%d is wider than the (h, h) feature map", h, w);
  size_t off_w, off_adj, total;
  seg_ws_layout(batch, smax, gh, gw, off_w, off_adj, total);
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned long long* stats = reinterpret_cast<unsigned long long*>(ws);
  float* wseg = feat ? reinterpret_cast<float*>(ws + off_w) : nullptr;
  unsigned int* adj = edges ? reinterpret_cast<unsigned int*>(ws + off_adj) : nullptr;
  SegmentArgs a;
  a.batch = batch; a.h = h; a.w = w; a.smax = smax; a.grid_h = gh; a.grid_w = gw; a.dim = dim;
  // dense features are upsampled to (h, h) in the reference (dino_interface.py:87-88)
  a.scale_y = h > 1 ? static_cast<float>(gh - 1) / static_cast<float>(h - 1) : 0.f;
  a.scale_x = h > 1 ? static_cast<float>(gw - 1) / static_cast<float>(h - 1) : 0.f;
  WVN_PROPAGATE(segment_accumulate(seg, a, stats, wseg, adj, S(stream)));
  if (feat || centers) WVN_PROPAGATE(segment_pool(wseg, tokens, stats, feat, centers, a, S(stream)));
  if (edges) WVN_PROPAGATE(adjacency_emit(adj, edges, n_edges, batch, smax, max_edges, S(stream)));
  return WVN_OK;
}

int wvn_segment_relabel(long long* seg, int batch, long long pix_per_frame, int num_labels, int* scratch, int* counts,
                        void* stream) {
  return relabel_compact(seg, scratch, counts, batch, pix_per_frame, num_labels, S(stream));
}

int wvn_supervision_pool(const long long* seg, const float* mask, int batch, int channels, int h, int w, int smax,
                         float* y, unsigned char* y_valid, float* count_ws, void* stream) {
  WVN_REQUIRE(seg && mask && y && y_valid && count_ws, "wvn_supervision_pool: null argument");
  return supervision_pool(seg, mask, batch, channels, h, w, smax, y, y_valid, count_ws, S(stream));
}

void wvn_slic_tables(int* g256, int* m9, int* f4096) { slic_tables(g256, m9, f4096); }

int wvn_slic_geometry(int h, int w, int num_components, int* grid_interval, int* nx, int* ny) {
  WVN_REQUIRE(h > 0 && w > 0 && num_components > 0 && grid_interval && nx && ny, "wvn_slic_geometry: bad argument");
  slic_geometry(h, w, num_components, grid_interval, nx, ny);
  return WVN_OK;
}

size_t wvn_slic_workspace_bytes(int batch, int h, int w, int num_components) {
  return slic_workspace_bytes(batch, h, w, num_components);
}

int wvn_slic(const float* img, int batch, int h, int w, int num_components, float compactness, int iters,
             const int* lut_g, const int* lut_m, const int* lut_f, long long* labels, void* workspace, void* stream) {
  WVN_REQUIRE(img && lut_g && lut_m && lut_f && labels && workspace, "wvn_slic: null argument");
  return slic_segment(img, batch, h, w, num_components, compactness, iters, lut_g, lut_m, lut_f, labels, workspace,
                      S(stream));
}

int wvn_project_and_render(const float* K, const float* pose_camera_in_world, const float* points, const float* colors,
                           int color_batched, int batch, int n_points, int h, int w, const float* traversability,
                           float* masks, float* projected, unsigned char* valid, float* supervision_inout, void* stream) {
  WVN_REQUIRE(K && pose_camera_in_world && points, "wvn_project_and_render: null argument");
  WVN_REQUIRE(colors || (!masks && !supervision_inout), "wvn_project_and_render: colors are required to render");
  FootprintArgs a;
  a.batch = batch; a.n_points = n_points; a.h = h; a.w = w; a.color_batched = color_batched;
  return footprint_render(a, K, pose_camera_in_world, points, colors, traversability, masks, projected, valid,
                          supervision_inout, S(stream));
}

}  // extern "C"

// ============================================================================================
// Traversability MLP inference handle
// ============================================================================================
struct wvn_mlp_infer {
  int dim, h1, h2;
  int dim_p, h1_p, h2_p, n3, n3_p, bn3, trav_col;
  int chunk_rows;
  DevBuf w1, b1, w2, b2, w3, b3;  // bf16 weights (padded / permuted), fp32 biases
  DevBuf x, a1, a2;               // bf16 activations of one chunk
  // fused per-pixel head (pixel_head.cu): per-token GEMM operands + workspaces for kFusedFrames frames
  DevBuf wcat, bias_cat, head_consts, tok_bf16, gu, gram;
  int fused_tokens = 0;           // token rows the fused workspaces are sized for (grown on demand)
  int force_unfused = 0;          // debugging / A-B knob ($WVN_PIXEL_HEAD=unfused)
  bool loaded = false;
};

static constexpr int kFusedFrames = 8;

namespace {

// Pack the flat fp32 state-dict parameters into the padded bf16 operands of the three GEMMs.
// Layer 3 rows are permuted: reconstruction rows first (so output column j reconstructs x[j]),
// the traversability row at column trav_col.
__global__ void pack_mlp_kernel(const float* __restrict__ p, MlpOffsets o, int dim, int h1, int h2, int dim_p, int h1_p,
                                int h2_p, int n3_p, int trav_col, __nv_bfloat16* w1, float* b1, __nv_bfloat16* w2,
                                float* b2, __nv_bfloat16* w3, float* b3) {
  const long long n1 = static_cast<long long>(h1_p) * dim_p, n2 = static_cast<long long>(h2_p) * h1_p,
                  n3 = static_cast<long long>(n3_p) * h2_p;
  const long long total = n1 + n2 + n3 + h1_p + h2_p + n3_p;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long j = i;
    if (j < n1) {
      const int r = static_cast<int>(j / dim_p), c = static_cast<int>(j % dim_p);
      w1[j] = __float2bfloat16_rn((r < h1 && c < dim) ? p[o.w1 + static_cast<long long>(r) * dim + c] : 0.f);
      continue;
    }
    j -= n1;
    if (j < n2) {
      const int r = static_cast<int>(j / h1_p), c = static_cast<int>(j % h1_p);
      w2[j] = __float2bfloat16_rn((r < h2 && c < h1) ? p[o.w2 + static_cast<long long>(r) * h1 + c] : 0.f);
      continue;
    }
    j -= n2;
    if (j < n3) {
      const int r = static_cast<int>(j / h2_p), c = static_cast<int>(j % h2_p);
      int src = -1;
      if (r < dim) src = 1 + r; else if (r == trav_col) src = 0;
      w3[j] = __float2bfloat16_rn((src >= 0 && c < h2) ? p[o.w3 + static_cast<long long>(src) * h2 + c] : 0.f);
      continue;
    }
    j -= n3;
    if (j < h1_p) { b1[j] = j < h1 ? p[o.b1 + j] : 0.f; continue; }
    j -= h1_p;
    if (j < h2_p) { b2[j] = j < h2 ? p[o.b2 + j] : 0.f; continue; }
    j -= h2_p;
    {
      int src = -1;
      if (j < dim) src = 1 + static_cast<int>(j); else if (j == trav_col) src = 0;
      b3[j] = src >= 0 ? p[o.b3 + src] : 0.f;
    }
  }
}

int mlp_infer_chunk(wvn_mlp_infer* h, long long rows, long long row0, const float* cg_mean, const float* cg_std,
                    float std_factor, float* trav, float* conf, cudaStream_t s) {
  GemmArgs g1;
  g1.M = static_cast<int>(rows); g1.N = h->h1_p; g1.K = h->dim_p; g1.epi = EPI_BF16; g1.act = ACT_RELU;
  g1.bias = reinterpret_cast<float*>(h->b1.p); g1.out = h->a1.p; g1.ldo = h->h1_p;
  WVN_PROPAGATE(gemm_bf16(g1, h->x.p, h->dim_p, h->w1.p, 0, s));
  GemmArgs g2;
  g2.M = static_cast<int>(rows); g2.N = h->h2_p; g2.K = h->h1_p; g2.epi = EPI_BF16; g2.act = ACT_RELU;
  g2.bias = reinterpret_cast<float*>(h->b2.p); g2.out = h->a2.p; g2.ldo = h->h2_p;
  WVN_PROPAGATE(gemm_bf16(g2, h->a1.p, h->h1_p, h->w2.p, 0, s));
  GemmArgs g3;
  g3.M = static_cast<int>(rows); g3.N = h->n3_p; g3.K = h->h2_p; g3.epi = EPI_MLP_HEAD;
  g3.bias = reinterpret_cast<float*>(h->b3.p); g3.feat = h->dim; g3.trav_col = h->trav_col; g3.x = h->x.p;
  g3.ldx = h->dim_p; g3.trav = trav + row0; g3.conf = conf + row0; g3.cg_mean = cg_mean; g3.cg_std = cg_std;
  g3.cg_std_factor = std_factor;
  WVN_PROPAGATE(gemm_bf16(g3, h->a2.p, h->h2_p, h->w3.p, h->bn3, s));
  return WVN_OK;
}

}  // namespace

extern "C" {

int wvn_mlp_infer_create(int dim, int h1, int h2, int chunk_rows, wvn_mlp_infer_t** out) {
  WVN_REQUIRE(out && dim > 0 && h1 > 0 && h2 > 0, "wvn_mlp_infer_create: bad arguments");
  WVN_PROPAGATE(wvn_check_device());
  wvn_mlp_infer* h = new wvn_mlp_infer();
  h->dim = dim; h->h1 = h1; h->h2 = h2;
  h->dim_p = round_up(dim, 64); h->h1_p = round_up(h1, 64); h->h2_p = round_up(h2, 64);
  h->trav_col = round_up(dim, 32);
  h->n3 = h->trav_col + 1;
  // pick the layer-3 tile width with the least padding (ties -> wider tile)
  int best_bn = 64, best_n = round_up(h->n3, 64);
  for (int bn : {128, 192, 224, 256}) {
    const int n = round_up(h->n3, bn);
    if (n <= best_n) { best_n = n; best_bn = bn; }
  }
  h->bn3 = best_bn; h->n3_p = best_n;
  h->chunk_rows = chunk_rows > 0 ? round_up(chunk_rows, 128) : 148 * 128 * 3;
  int rc = WVN_OK;
  auto alloc = [&](DevBuf& b, size_t bytes) { if (rc == WVN_OK) rc = b.alloc(bytes); };
  alloc(h->w1, static_cast<size_t>(h->h1_p) * h->dim_p * 2);
  alloc(h->b1, static_cast<size_t>(h->h1_p) * 4);
  alloc(h->w2, static_cast<size_t>(h->h2_p) * h->h1_p * 2);
  alloc(h->b2, static_cast<size_t>(h->h2_p) * 4);
  alloc(h->w3, static_cast<size_t>(h->n3_p) * h->h2_p * 2);
  alloc(h->b3, static_cast<size_t>(h->n3_p) * 4);
  alloc(h->x, static_cast<size_t>(h->chunk_rows) * h->dim_p * 2);
  alloc(h->a1, static_cast<size_t>(h->chunk_rows) * h->h1_p * 2);
  alloc(h->a2, static_cast<size_t>(h->chunk_rows) * h->h2_p * 2);
  alloc(h->wcat, static_cast<size_t>(kPixelHeadN) * h->dim_p * 2);
  alloc(h->bias_cat, static_cast<size_t>(kPixelHeadN) * 4);
  alloc(h->head_consts, sizeof(PixelHeadConsts));
  {
    const char* e = getenv("WVN_PIXEL_HEAD");
    h->force_unfused = (e && std::string(e) == "unfused") ? 1 : 0;
  }
  if (rc != WVN_OK) {
    wvn_mlp_infer_destroy(h);
    return rc;
  }
  *out = h;
  return WVN_OK;
}

void wvn_mlp_infer_destroy(wvn_mlp_infer_t* h) {
  if (!h) return;
  for (DevBuf* b : {&h->w1, &h->b1, &h->w2, &h->b2, &h->w3, &h->b3, &h->x, &h->a1, &h->a2, &h->wcat, &h->bias_cat,
                    &h->head_consts, &h->tok_bf16, &h->gu, &h->gram})
    b->release();
  delete h;
}

int wvn_mlp_infer_reserve(wvn_mlp_infer_t* h, int tokens_per_frame) {
  WVN_REQUIRE(h && tokens_per_frame > 0, "wvn_mlp_infer_reserve: bad arguments");
  const int P = tokens_per_frame;
  if (h->fused_tokens >= kFusedFrames * P) return WVN_OK;
  for (DevBuf* b : {&h->tok_bf16, &h->gu, &h->gram}) b->release();
  WVN_PROPAGATE(h->tok_bf16.alloc(static_cast<size_t>(kFusedFrames) * P * h->dim_p * 2));
  WVN_PROPAGATE(h->gu.alloc(static_cast<size_t>(kFusedFrames) * P * kPixelHeadN * 4));
  WVN_PROPAGATE(h->gram.alloc(static_cast<size_t>(kFusedFrames) * P * 5 * 4));
  h->fused_tokens = kFusedFrames * P;
  return WVN_OK;
}

int wvn_mlp_infer_set_params(wvn_mlp_infer_t* h, const float* params, void* stream) {
  WVN_REQUIRE(h && params, "wvn_mlp_infer_set_params: null argument");
  MlpShape sh;
  sh.dim = h->dim; sh.h1 = h->h1; sh.h2 = h->h2;
  pack_mlp_kernel<<<256, 256, 0, S(stream)>>>(
      params, mlp_offsets(sh), h->dim, h->h1, h->h2, h->dim_p, h->h1_p, h->h2_p, h->n3_p, h->trav_col,
      reinterpret_cast<__nv_bfloat16*>(h->w1.p), reinterpret_cast<float*>(h->b1.p),
      reinterpret_cast<__nv_bfloat16*>(h->w2.p), reinterpret_cast<float*>(h->b2.p),
      reinterpret_cast<__nv_bfloat16*>(h->w3.p), reinterpret_cast<float*>(h->b3.p));
  WVN_CHECK_LAUNCH("pack_mlp_kernel");
  if (h->h1 == 256 && h->h2 == 32)
    WVN_PROPAGATE(pixel_head_pack(params, sh, h->dim_p, h->wcat.p, reinterpret_cast<float*>(h->bias_cat.p),
                                  reinterpret_cast<PixelHeadConsts*>(h->head_consts.p), S(stream)));
  h->loaded = true;
  return WVN_OK;
}

// Fused per-pixel head over frames [b0, b0 + nb): per-token GEMM (G | U | cT) + token Gram + one pixel kernel.
// tok_bf16: the frames' bf16 tokens, frame_rows rows per frame with the patch tokens starting at row row0.
static int pixels_fused_chunk(wvn_mlp_infer_t* h, const void* tok_bf16, long long frame_rows, int row0, int b0, int nb,
                              int gh, int gw, int out_h, int out_w, int ww, const float* cg_mean, const float* cg_std,
                              float std_factor, float* trav, float* conf, cudaStream_t s) {
  const long long rows = static_cast<long long>(nb) * frame_rows;
  GemmArgs g;
  g.M = static_cast<int>(rows); g.N = kPixelHeadN; g.K = h->dim_p; g.epi = EPI_F32;
  g.bias = reinterpret_cast<float*>(h->bias_cat.p); g.out = h->gu.p; g.ldo = kPixelHeadN;
  WVN_PROPAGATE(gemm_bf16(g, tok_bf16, h->dim_p, h->wcat.p, 64, s));
  WVN_PROPAGATE(token_gram(tok_bf16, reinterpret_cast<float*>(h->gram.p), nb, gh, gw, h->dim_p, frame_rows, row0, s));
  PixelHeadArgs a;
  a.gu = reinterpret_cast<float*>(h->gu.p); a.ldg = kPixelHeadN; a.gram = reinterpret_cast<float*>(h->gram.p);
  a.consts = reinterpret_cast<PixelHeadConsts*>(h->head_consts.p);
  a.cg_mean = cg_mean; a.cg_std = cg_std; a.std_factor = std_factor;
  a.trav = trav + static_cast<long long>(b0) * out_h * out_w;
  a.conf = conf + static_cast<long long>(b0) * out_h * out_w;
  a.batch = nb; a.gh = gh; a.gw = gw; a.H = out_h; a.W = out_w;
  a.sy = static_cast<float>(gh - 1) / static_cast<float>(out_h - 1);
  a.sx = static_cast<float>(gw - 1) / static_cast<float>(out_w - 1);
  a.ww = ww; a.feat = h->dim;
  a.frame_rows = frame_rows; a.row0 = row0;
#ifdef WVN_GEMM_TIMING
  static long long* ptb = nullptr;
  if (!ptb) cudaMalloc(&ptb, 16 * sizeof(long long));
  a.timing = ptb;
#endif
  WVN_PROPAGATE(pixel_head(a, h->w2.p, h->h1_p, s));
#ifdef WVN_GEMM_TIMING
  {
    long long t[16];
    cudaMemcpyAsync(t, ptb, sizeof(t), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    const long long n = t[5] > 0 ? t[5] : 1;
    fprintf(stderr, "[pixel_head timing, cycles per tile of CTA 0 (%lld tiles)] phaseA+sync %lld  phaseB+sync %lld  mma_wait %lld  "
            "epilogue %lld  end_sync %lld | producer: tables+bar %lld  copy %lld  pix/gram %lld\n", n, t[0] / n, t[1] / n, t[2] / n, t[3] / n,
            t[4] / n, t[8] / n, t[9] / n, t[10] / n);
  }
#endif
  return WVN_OK;
}

int wvn_mlp_infer_pixels_vit(wvn_mlp_infer_t* h, wvn_vit_t* vit, int batch, int out_h, int out_w, const float* cg_mean,
                             const float* cg_std, float std_factor, float* trav, float* conf, void* stream) {
  WVN_REQUIRE(h && vit && trav && conf && cg_mean && cg_std, "wvn_mlp_infer_pixels_vit: null argument");
  if (!h->loaded) return set_error(WVN_ERR_STATE, "wvn_mlp_infer_pixels_vit: parameters were never set");
  if (!vit->forwarded || batch > vit->last_batch)
    return set_error(WVN_ERR_STATE, "wvn_mlp_infer_pixels_vit: the backbone holds the tokens of %d frames, %d asked",
                     vit->forwarded ? vit->last_batch : 0, batch);
  WVN_REQUIRE(h->dim == vit->cfg.dim && h->dim_p == vit->cfg.dim, "wvn_mlp_infer_pixels_vit: the MLP takes %d-d features, "
              "the backbone's tokens are %d-d", h->dim, vit->cfg.dim);
  const int g = vit->grid;
  const int ww = pixel_head_supported(h->h1, h->h2, g, g, out_h, out_w);
  WVN_REQUIRE(ww > 0, "wvn_mlp_infer_pixels_vit: geometry outside the fused per-pixel head (use wvn_mlp_infer_pixels)");
  if (h->fused_tokens < kFusedFrames * vit->npad) WVN_PROPAGATE(wvn_mlp_infer_reserve(h, vit->npad));
  for (int b0 = 0; b0 < batch; b0 += kFusedFrames) {
    const int nb = std::min(kFusedFrames, batch - b0);
    const __nv_bfloat16* tok = reinterpret_cast<const __nv_bfloat16*>(vit->tok_bf16.p) +
                               static_cast<long long>(b0) * vit->npad * vit->cfg.dim;
    WVN_PROPAGATE(pixels_fused_chunk(h, tok, vit->npad, 1, b0, nb, g, g, out_h, out_w, ww, cg_mean, cg_std, std_factor, trav,
                                     conf, S(stream)));
  }
  return WVN_OK;
}

int wvn_mlp_infer_pixels(wvn_mlp_infer_t* h, const float* tokens, int batch, int gh, int gw, int out_h, int out_w,
                         const float* cg_mean, const float* cg_std, float std_factor, float* trav, float* conf,
                         void* stream) {
  WVN_REQUIRE(h && tokens && trav && conf && cg_mean && cg_std, "wvn_mlp_infer_pixels: null argument");
  if (!h->loaded) return set_error(WVN_ERR_STATE, "wvn_mlp_infer_pixels: parameters were never set");
  cudaStream_t s = S(stream);
  // any feature width works (the 90-d STEGO code is zero-padded to 128 columns in the bf16 operands)
  const int ww = h->force_unfused ? 0 : pixel_head_supported(h->h1, h->h2, gh, gw, out_h, out_w);
  if (ww > 0) {
    // ---- fused path: per-token GEMM (G | U | cT) + token Gram, then one kernel per chunk of frames
    const int P = gh * gw;
    // workspaces are sized by wvn_mlp_infer_reserve (called by the owner right after create); a larger token grid
    // than reserved grows them here once
    if (h->fused_tokens < kFusedFrames * P) WVN_PROPAGATE(wvn_mlp_infer_reserve(h, P));
    for (int b0 = 0; b0 < batch; b0 += kFusedFrames) {
      const int nb = std::min(kFusedFrames, batch - b0);
      const long long rows = static_cast<long long>(nb) * P;
      const long long elems = rows * h->dim;
      int blocks = static_cast<int>(std::min<long long>((elems + 255) / 256, 8192));
      cast_rows_kernel<<<blocks, 256, 0, s>>>(tokens + static_cast<long long>(b0) * P * h->dim,
                                             reinterpret_cast<__nv_bfloat16*>(h->tok_bf16.p), rows, h->dim, h->dim_p);
      WVN_CHECK_LAUNCH("cast_rows_kernel");
      WVN_PROPAGATE(pixels_fused_chunk(h, h->tok_bf16.p, P, 0, b0, nb, gh, gw, out_h, out_w, ww, cg_mean, cg_std,
                                       std_factor, trav, conf, s));
    }
    return WVN_OK;
  }
  DenseArgs d;
  d.batch = batch; d.dim = h->dim; d.grid_h = gh; d.grid_w = gw; d.out_h = out_h; d.out_w = out_w;
  d.scale_y = out_h > 1 ? static_cast<float>(gh - 1) / static_cast<float>(out_h - 1) : 0.f;
  d.scale_x = out_w > 1 ? static_cast<float>(gw - 1) / static_cast<float>(out_w - 1) : 0.f;
  d.ld_out = h->dim_p;
  const long long total = static_cast<long long>(batch) * out_h * out_w;
  for (long long p0 = 0; p0 < total; p0 += h->chunk_rows) {
    const long long n = std::min<long long>(h->chunk_rows, total - p0);
    WVN_PROPAGATE(interp_pixel_rows(tokens, h->x.p, d, p0, n, s));
    WVN_PROPAGATE(mlp_infer_chunk(h, n, p0, cg_mean, cg_std, std_factor, trav, conf, s));
  }
  return WVN_OK;
}

int wvn_mlp_infer_rows(wvn_mlp_infer_t* h, const float* x, long long rows, const float* cg_mean, const float* cg_std,
                       float std_factor, float* trav, float* conf, void* stream) {
  WVN_REQUIRE(h && x && trav && conf && cg_mean && cg_std, "wvn_mlp_infer_rows: null argument");
  if (!h->loaded) return set_error(WVN_ERR_STATE, "wvn_mlp_infer_rows: parameters were never set");
  cudaStream_t s = S(stream);
  for (long long r0 = 0; r0 < rows; r0 += h->chunk_rows) {
    const long long n = std::min<long long>(h->chunk_rows, rows - r0);
    const long long elems = n * h->dim;
    int blocks = static_cast<int>(std::min<long long>((elems + 255) / 256, 8192));
    cast_rows_kernel<<<blocks, 256, 0, s>>>(x + r0 * h->dim, reinterpret_cast<__nv_bfloat16*>(h->x.p), n, h->dim,
                                           h->dim_p);
    WVN_CHECK_LAUNCH("cast_rows_kernel");
    WVN_PROPAGATE(mlp_infer_chunk(h, n, r0, cg_mean, cg_std, std_factor, trav, conf, s));
  }
  return WVN_OK;
}

// -------------------------------------------------------------------------------- training
static MlpShape shape_of(int dim, int h1, int h2) {
  MlpShape s;
  s.dim = dim; s.h1 = h1; s.h2 = h2;
  return s;
}
static LossCfg loss_of(const wvn_train_config* c) {
  LossCfg l;
  l.w_trav = c->w_trav; l.w_reco = c->w_reco; l.std_factor = c->std_factor; l.anomaly_balanced = c->anomaly_balanced;
  return l;
}

size_t wvn_mlp_param_count(int dim, int h1, int h2) { return mlp_param_count(shape_of(dim, h1, h2)); }
size_t wvn_mlp_train_workspace_bytes(int dim, int h1, int h2, int max_rows) {
  return mlp_train_workspace_floats(shape_of(dim, h1, h2), max_rows) * sizeof(float);
}
size_t wvn_mlp_train_scalars_bytes(void) { return sizeof(TrainScalars); }

int wvn_mlp_train_forward_stats(int dim, int h1, int h2, const float* params, const float* x, const float* y,
                                const unsigned char* y_valid, int rows, int max_rows, void* workspace, void* scalars,
                                void* stream) {
  WVN_REQUIRE(params && x && y && y_valid && workspace && scalars, "wvn_mlp_train_forward_stats: null argument");
  return mlp_train_forward_stats(shape_of(dim, h1, h2), params, x, y, y_valid, rows, max_rows,
                                 reinterpret_cast<float*>(workspace), reinterpret_cast<TrainScalars*>(scalars), S(stream));
}

int wvn_mlp_train_backward(int dim, int h1, int h2, const float* params, const float* x, const float* y,
                           const unsigned char* y_valid, int rows, int max_rows, long long n_total,
                           const wvn_train_config* cfg, void* workspace, void* scalars, float* cg_mean, float* cg_std,
                           float* grads, float* confidence_out, void* stream) {
  WVN_REQUIRE(params && x && y && y_valid && workspace && scalars && cfg && grads && confidence_out,
              "wvn_mlp_train_backward: null argument");
  return mlp_train_backward(shape_of(dim, h1, h2), params, x, y, y_valid, rows, max_rows, n_total, loss_of(cfg),
                            reinterpret_cast<float*>(workspace), reinterpret_cast<TrainScalars*>(scalars), cg_mean,
                            cg_std, grads, confidence_out, S(stream));
}

int wvn_mlp_train_apply(int dim, int h1, int h2, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                        long long* step_counter, long long n_total, const wvn_train_config* cfg, void* scalars,
                        void* stream) {
  WVN_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_counter && cfg && scalars,
              "wvn_mlp_train_apply: null argument");
  const long long n = static_cast<long long>(mlp_param_count(shape_of(dim, h1, h2)));
  WVN_PROPAGATE(mlp_train_finalize(reinterpret_cast<TrainScalars*>(scalars), grads, n, n_total, loss_of(cfg), S(stream)));
  AdamCfg a;
  a.lr = cfg->lr; a.beta1 = cfg->beta1; a.beta2 = cfg->beta2; a.eps = cfg->eps;
  return mlp_adam_step(params, grads, exp_avg, exp_avg_sq, n, a, step_counter, S(stream));
}

__global__ void read_metrics_kernel(const TrainScalars* sc, float* out) {
  out[0] = sc->loss_total; out[1] = sc->loss_trav; out[2] = sc->loss_reco; out[3] = sc->loss_trav_conf;
  out[4] = sc->mean; out[5] = sc->std;
}

int wvn_mlp_train_read_metrics(const void* scalars, float* metrics_out, void* stream) {
  WVN_REQUIRE(scalars && metrics_out, "wvn_mlp_train_read_metrics: null argument");
  read_metrics_kernel<<<1, 1, 0, S(stream)>>>(reinterpret_cast<const TrainScalars*>(scalars), metrics_out);
  WVN_CHECK_LAUNCH("read_metrics_kernel");
  return WVN_OK;
}

int wvn_mlp_forward_f32(int dim, int h1, int h2, const float* params, const float* x, int rows, float* h1_buf,
                        float* h2_buf, float* out, void* stream) {
  WVN_REQUIRE(params && x && h1_buf && h2_buf && out && rows > 0, "wvn_mlp_forward_f32: bad argument");
  return mlp_forward_f32(shape_of(dim, h1, h2), params, x, rows, h1_buf, h2_buf, out, S(stream));
}


// -------------------------------------------------------------------------------- fused train step
struct wvn_mlp_trainer {
  FusedTrainer* impl = nullptr;
};

size_t wvn_mlp_trainer_scalars_bytes(void) { return sizeof(FusedScalars); }

int wvn_mlp_trainer_create(int dim, int h1, int h2, int max_rows, const wvn_train_config* cfg, void* scalars, float* grads,
                           wvn_mlp_trainer_t** out) {
  WVN_REQUIRE(cfg && out, "wvn_mlp_trainer_create: null argument");
  WVN_PROPAGATE(wvn_check_device());
  AdamCfg a;
  a.lr = cfg->lr; a.beta1 = cfg->beta1; a.beta2 = cfg->beta2; a.eps = cfg->eps;
  FusedTrainer* impl = nullptr;
  WVN_PROPAGATE(fused_trainer_create(shape_of(dim, h1, h2), max_rows, loss_of(cfg), a, scalars, grads, &impl));
  wvn_mlp_trainer* t = new wvn_mlp_trainer();
  t->impl = impl;
  *out = t;
  return WVN_OK;
}

void wvn_mlp_trainer_destroy(wvn_mlp_trainer_t* t) {
  if (!t) return;
  fused_trainer_destroy(t->impl);
  delete t;
}

int wvn_comm_unique_id(void* id128) { return fused_comm_unique_id(id128); }

int wvn_mlp_trainer_init_comm(wvn_mlp_trainer_t* t, const void* id128, int rank, int world) {
  WVN_REQUIRE(t, "wvn_mlp_trainer_init_comm: null trainer");
  return fused_trainer_init_comm(t->impl, id128, rank, world);
}

int wvn_mlp_trainer_set_confidence(wvn_mlp_trainer_t* t, int method, float* var, double* running_n, double* running_sum,
                                   double* running_sum_of_squares, float kf_proc_cov, float kf_meas_cov) {
  WVN_REQUIRE(t, "wvn_mlp_trainer_set_confidence: null trainer");
  return fused_trainer_set_confidence(t->impl, method, var, running_n, running_sum, running_sum_of_squares, kf_proc_cov,
                                      kf_meas_cov);
}

int wvn_mlp_train_step(wvn_mlp_trainer_t* t, float* params, float* exp_avg, float* exp_avg_sq, long long* step_counter,
                       const float* x, int groups, int rows_per_group, const int* n_rows, const float* y,
                       const unsigned char* y_valid, float* cg_mean, float* cg_std, float* confidence_out,
                       float* metrics_out, int phase_mask, void* stream) {
  WVN_REQUIRE(t, "wvn_mlp_train_step: null trainer");
  return fused_train_step(t->impl, params, exp_avg, exp_avg_sq, step_counter, x, groups, rows_per_group, n_rows, y, y_valid,
                          cg_mean, cg_std, confidence_out, metrics_out, phase_mask, S(stream));
}

}  // extern "C"
