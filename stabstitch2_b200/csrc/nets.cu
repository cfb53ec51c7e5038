// Weight packing (state-dict -> kernel layouts, eval-mode BatchNorm folded) and the forward
// graphs of SpatialNet, TemporalNet and SmoothNet.
//
// Reference behaviour restated (paths under Full_model_inference/Codes/):
//   spatial_network.py:144-331   SpatialNet.__init__/forward
//   temporal_network.py:62-147   TemporalNet.__init__/forward
//   smooth_network.py:44-157     SmoothNet / MotionPrediction
// State-dict key names: SURVEY.md Appendix B (verified by strict load into the reference).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <utility>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------
static const HostTensor* find(ss2_ctx* ctx, int net, const std::string& key) {
  auto it = ctx->host_weights[net].find(key);
  return it == ctx->host_weights[net].end() ? nullptr : &it->second;
}

static int upload(ss2_ctx* ctx, const std::vector<float>& h, float** d) {
  void* p = nullptr;
  SS2_CUDA(ctx, cudaMalloc(&p, h.size() * sizeof(float)));
  (ctx->packing_net >= 0 ? ctx->owned_net[ctx->packing_net] : ctx->owned).push_back(p);
  SS2_CUDA(ctx, cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  *d = (float*)p;
  return SS2_OK;
}


// conv weight [Cout,Cin,(KD,)KH,KW] (+ optional BN prefix, + optional bias key) -> ConvLayer.
// `flatten_hw` > 0 packs a Linear that consumes an NCHW-flattened [C, flatten_hw] map whose
// activations we hold as NHWC: input index (p*C + c) <- reference column (c*flatten_hw + p).
static int pack_conv(ss2_ctx* ctx, int net, const std::string& wkey, const std::string& bnpfx,
                     const std::string& bkey, int stride, int pad, int pad_d, int flatten_hw, ConvLayer* L) {
  const HostTensor* w = find(ctx, net, wkey);
  if (!w) return ss2_fail(ctx, SS2_ERR_MISSING_KEY, "missing state-dict key '%s'", wkey.c_str());
  const int nd = (int)w->shape.size();
  if (nd != 2 && nd != 4 && nd != 5) return ss2_fail(ctx, SS2_ERR_INVALID, "'%s': unexpected rank %d", wkey.c_str(), nd);
  const int Cout = (int)w->shape[0];
  int Cin = (int)w->shape[1];
  int KD = 1, KH = 1, KW = 1;
  if (nd == 4) { KH = (int)w->shape[2]; KW = (int)w->shape[3]; }
  if (nd == 5) { KD = (int)w->shape[2]; KH = (int)w->shape[3]; KW = (int)w->shape[4]; }
  std::vector<double> scale(Cout, 1.0), shift(Cout, 0.0);
  bool has_bias = false;
  if (!bnpfx.empty()) {
    const HostTensor *g = find(ctx, net, bnpfx + ".weight"), *b = find(ctx, net, bnpfx + ".bias"),
                     *m = find(ctx, net, bnpfx + ".running_mean"), *v = find(ctx, net, bnpfx + ".running_var");
    if (!g || !b || !m || !v)
      return ss2_fail(ctx, SS2_ERR_MISSING_KEY, "missing BatchNorm tensors under '%s'", bnpfx.c_str());
    for (int o = 0; o < Cout; ++o) {
      // F.batch_norm(eval): (x - mean) / sqrt(var + eps) * gamma + beta, eps = 1e-5
      const double s = (double)g->data[o] / sqrt((double)v->data[o] + 1e-5);
      scale[o] = s;
      shift[o] = (double)b->data[o] - (double)m->data[o] * s;
    }
    has_bias = true;
  }
  if (!bkey.empty()) {
    const HostTensor* b = find(ctx, net, bkey);
    if (!b) return ss2_fail(ctx, SS2_ERR_MISSING_KEY, "missing state-dict key '%s'", bkey.c_str());
    for (int o = 0; o < Cout; ++o) shift[o] += (double)b->data[o];
    has_bias = true;
  }
  L->Cin = Cin; L->Cout = Cout;
  L->CinP = (Cin + 3) / 4 * 4;
  // channel-padded producers: cost volumes are emitted with 128 / 64 channels
  if (Cin == 121) L->CinP = 128;
  if (Cin == 49) L->CinP = 64;
  L->CoutP = (Cout + 63) / 64 * 64;
  L->KD = KD; L->KH = KH; L->KW = KW;
  L->sd = 1; L->sh = L->sw = stride;
  L->pd = pad_d; L->ph = L->pw = pad;
  const int taps = KD * KH * KW;
  std::vector<float> wp((size_t)taps * L->CinP * L->CoutP, 0.f);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t) {
        const double v = (double)w->data[((size_t)o * Cin + c) * taps + t] * scale[o];
        size_t krow;
        if (flatten_hw > 0) {
          const int ch = c / flatten_hw, p = c % flatten_hw;  // reference column c = ch*hw + p
          krow = (size_t)p * (Cin / flatten_hw) + ch;
        } else {
          krow = (size_t)t * L->CinP + c;
        }
        wp[krow * L->CoutP + o] = (float)v;
      }
  SS2_TRY(upload(ctx, wp, &L->w));
  // tensor-core path: K-major [CoutP][Ktot] split into hi = rna_tf32(w), lo = w - hi
  L->wk_hi = L->wk_lo = nullptr;
  if (flatten_hw == 0 && (L->CinP % 32) == 0 && (Cout % 4) == 0) {
    const size_t Ktot = (size_t)taps * L->CinP;
    std::vector<float> hi((size_t)L->CoutP * Ktot, 0.f), lo((size_t)L->CoutP * Ktot, 0.f);
    for (size_t k = 0; k < Ktot; ++k)
      for (int o = 0; o < L->CoutP; ++o) {
        const float v = wp[k * L->CoutP + o];
        uint32_t b;
        memcpy(&b, &v, 4);
        b = (b + 0x1000u) & 0xFFFFE000u;  // round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32
        float h;
        memcpy(&h, &b, 4);
        if (!isfinite(h)) h = v;
        hi[(size_t)o * Ktot + k] = h;
        lo[(size_t)o * Ktot + k] = v - h;
      }
    SS2_TRY(upload(ctx, hi, &L->wk_hi));
    SS2_TRY(upload(ctx, lo, &L->wk_lo));
    // layers whose input can arrive as fp16 split planes (conv_dc.cu, conv_tc.cu): the same matrix as h16 = fp16(w),
    // l16 = fp16((w - h16) * 2048), two fp16 values per uploaded float
    L->wk_h16 = L->wk_l16 = nullptr;
    float wmax = 0.f;
    for (float v : wp) wmax = fmaxf(wmax, fabsf(v));
    if ((L->CinP % 64) == 0 && (Cout % 64) == 0 && wmax <= 65504.0f) {
      const size_t nel = (size_t)L->CoutP * Ktot;   // even: CinP % 64 == 0
      std::vector<float> ph(nel / 2), pl(nel / 2);
      __half* h16 = reinterpret_cast<__half*>(ph.data());
      __half* l16 = reinterpret_cast<__half*>(pl.data());
      for (size_t k = 0; k < Ktot; ++k)
        for (int o = 0; o < L->CoutP; ++o) {
          const float v = wp[k * L->CoutP + o];
          const float vc = v > 65504.0f ? 65504.0f : (v < -65504.0f ? -65504.0f : v);
          const __half hh = __float2half_rn(vc);
          h16[(size_t)o * Ktot + k] = hh;
          l16[(size_t)o * Ktot + k] = __float2half_rn((v - __half2float(hh)) * 2048.0f);
        }
      float *dh = nullptr, *dl = nullptr;
      SS2_TRY(upload(ctx, ph, &dh));
      SS2_TRY(upload(ctx, pl, &dl));
      L->wk_h16 = reinterpret_cast<__half*>(dh);
      L->wk_l16 = reinterpret_cast<__half*>(dl);
    }
  }
  L->stem_k32 = false;
  if (flatten_hw == 0 && KD == 1 && KH == 7 && KW == 7 && Cin == 3 && stride == 2 && pad == 3 && (Cout % 4) == 0) {
    const size_t Ktot = 7 * 32;
    std::vector<float> hi((size_t)L->CoutP * Ktot, 0.f), lo((size_t)L->CoutP * Ktot, 0.f);
    for (int o = 0; o < Cout; ++o)
      for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
          for (int c = 0; c < 3; ++c) {
            const float v = wp[((size_t)(kh * 7 + kw) * L->CinP + c) * L->CoutP + o];
            uint32_t b;
            memcpy(&b, &v, 4);
            b = (b + 0x1000u) & 0xFFFFE000u;
            float h;
            memcpy(&h, &b, 4);
            if (!isfinite(h)) h = v;
            hi[(size_t)o * Ktot + kh * 32 + kw * 4 + c] = h;
            lo[(size_t)o * Ktot + kh * 32 + kw * 4 + c] = v - h;
          }
    SS2_TRY(upload(ctx, hi, &L->wk_hi));
    SS2_TRY(upload(ctx, lo, &L->wk_lo));
    L->stem_k32 = true;
    if (L->CoutP == 64) {
      // direct kernel (conv_stem.cu): 25 k-steps of 8 = two NHWC4 pixels each.  Steps 3 kh + {0, 1, 2}: filter columns
      // (0, 2), (4, 6), (1, 3) of filter row kh; steps 21 + t: column 5 of rows 2t and 2t + 1; step 24: column 5 of row 6
      // and a zero pixel.
      std::vector<float> shi(hi.size(), 0.f), slo(lo.size(), 0.f);
      auto put = [&](int o, int step, int half, int kh, int kw) {
        for (int c = 0; c < 3; ++c) {
          const size_t from = (size_t)o * Ktot + kh * 32 + kw * 4 + c;
          const size_t to = (size_t)o * Ktot + step * 8 + half * 4 + c;
          shi[to] = hi[from];
          slo[to] = lo[from];
        }
      };
      for (int o = 0; o < Cout; ++o) {
        for (int kh = 0; kh < 7; ++kh) {
          put(o, 3 * kh + 0, 0, kh, 0); put(o, 3 * kh + 0, 1, kh, 2);
          put(o, 3 * kh + 1, 0, kh, 4); put(o, 3 * kh + 1, 1, kh, 6);
          put(o, 3 * kh + 2, 0, kh, 1); put(o, 3 * kh + 2, 1, kh, 3);
        }
        for (int t = 0; t < 3; ++t) { put(o, 21 + t, 0, 2 * t, 5); put(o, 21 + t, 1, 2 * t + 1, 5); }
        put(o, 24, 0, 6, 5);
      }
      SS2_TRY(upload(ctx, shi, &L->ws_hi));
      SS2_TRY(upload(ctx, slo, &L->ws_lo));
    }
  }
  L->bias = nullptr;
  if (has_bias) {
    std::vector<float> bp(L->CoutP, 0.f);
    for (int o = 0; o < Cout; ++o) bp[o] = (float)shift[o];
    SS2_TRY(upload(ctx, bp, &L->bias));
  }
  return SS2_OK;
}

static int pack_block(ss2_ctx* ctx, int net, const std::string& pfx, int stride, ResBlock* b) {
  SS2_TRY(pack_conv(ctx, net, pfx + ".conv1.weight", pfx + ".bn1", "", stride, 1, 0, 0, &b->c1));
  SS2_TRY(pack_conv(ctx, net, pfx + ".conv2.weight", pfx + ".bn2", "", 1, 1, 0, 0, &b->c2));
  b->has_down = find(ctx, net, pfx + ".downsample.0.weight") != nullptr;
  if (b->has_down)
    SS2_TRY(pack_conv(ctx, net, pfx + ".downsample.0.weight", pfx + ".downsample.1", "", stride, 0, 0, 0, &b->down));
  return SS2_OK;
}

static int pack_backbone(ss2_ctx* ctx, int net, bool with_stage2, Backbone* bb) {
  const std::string p1 = "feature_extractor_stage1", p2 = "feature_extractor_stage2";
  SS2_TRY(pack_conv(ctx, net, p1 + ".0.weight", p1 + ".1", "", 2, 3, 0, 0, &bb->stem));
  SS2_TRY(pack_block(ctx, net, p1 + ".4.0", 1, &bb->l1[0]));
  SS2_TRY(pack_block(ctx, net, p1 + ".4.1", 1, &bb->l1[1]));
  SS2_TRY(pack_block(ctx, net, p1 + ".5.0", 2, &bb->l2[0]));
  SS2_TRY(pack_block(ctx, net, p1 + ".5.1", 1, &bb->l2[1]));
  if (with_stage2) {
    SS2_TRY(pack_block(ctx, net, p2 + ".0.0", 2, &bb->l3[0]));
    SS2_TRY(pack_block(ctx, net, p2 + ".0.1", 1, &bb->l3[1]));
  }
  return SS2_OK;
}

static int pack_regressor(ss2_ctx* ctx, int net, const std::string& part1, const std::string& part2, int nconv,
                          int final_hw, Regressor* r) {
  static const int ids6[] = {0, 2, 5, 7, 10, 12}, ids8[] = {0, 2, 5, 7, 10, 12, 15, 17};
  const int* ids = nconv == 6 ? ids6 : ids8;
  r->convs.resize(nconv);
  r->pool_after.assign(nconv, 0);
  for (int i = 0; i < nconv; ++i) {
    SS2_TRY(pack_conv(ctx, net, part1 + "." + std::to_string(ids[i]) + ".weight", "", "", 1, 1, 0, 0, &r->convs[i]));
    r->pool_after[i] = (i % 2 == 1);
  }
  SS2_TRY(pack_conv(ctx, net, part2 + ".0.weight", "", part2 + ".0.bias", 1, 0, 0, final_hw, &r->fc[0]));
  SS2_TRY(pack_conv(ctx, net, part2 + ".2.weight", "", part2 + ".2.bias", 1, 0, 0, 0, &r->fc[1]));
  SS2_TRY(pack_conv(ctx, net, part2 + ".4.weight", "", part2 + ".4.bias", 1, 0, 0, 0, &r->fc[2]));
  return SS2_OK;
}

static int upload_key(ss2_ctx* ctx, int net, const std::string& key, float** d) {
  const HostTensor* t = find(ctx, net, key);
  if (!t) return ss2_fail(ctx, SS2_ERR_MISSING_KEY, "missing state-dict key '%s'", key.c_str());
  return upload(ctx, t->data, d);
}

static int finalize_weights_impl(ss2_ctx* ctx, int net_id);

extern "C" int ss2_finalize_weights(ss2_ctx* ctx, int net_id) {
  if (!ctx) return SS2_ERR_INVALID;
  if (net_id < 0 || net_id > 2) return ss2_fail(ctx, SS2_ERR_INVALID, "unknown net id %d", net_id);
  SS2_CUDA(ctx, cudaSetDevice(ctx->device));
  // a re-finalize replaces the network's packed weights: wait for launches that still read the old set, free it
  if (!ctx->owned_net[net_id].empty()) {
    SS2_CUDA(ctx, cudaDeviceSynchronize());
    for (void* p : ctx->owned_net[net_id]) cudaFree(p);
    ctx->owned_net[net_id].clear();
  }
  ctx->packing_net = net_id;
  const int rc = finalize_weights_impl(ctx, net_id);
  ctx->packing_net = -1;
  return rc;
}

static int finalize_weights_impl(ss2_ctx* ctx, int net_id) {
  if (net_id == SS2_NET_SPATIAL) {
    SpatialWeights& s = ctx->spatial;
    s.ready = false;
    SS2_TRY(pack_backbone(ctx, net_id, true, &s.bb));
    SS2_TRY(pack_regressor(ctx, net_id, "regressNet1_part1", "regressNet1_part2", 6, 2 * 3, &s.r1));
    SS2_TRY(pack_regressor(ctx, net_id, "regressNet2_part1_ref", "regressNet2_part2_ref", 8, 2 * 3, &s.r2_ref));
    SS2_TRY(pack_regressor(ctx, net_id, "regressNet2_part1_tgt", "regressNet2_part2_tgt", 8, 2 * 3, &s.r2_tgt));
    s.ready = true;
  } else if (net_id == SS2_NET_TEMPORAL) {
    TemporalWeights& t = ctx->temporal;
    t.ready = false;
    SS2_TRY(pack_backbone(ctx, net_id, false, &t.bb));  // stage2 is in the checkpoint but unused
    SS2_TRY(pack_regressor(ctx, net_id, "regressNet2_part1", "regressNet2_part2", 8, 2 * 3, &t.r2));
    t.ready = true;
  } else if (net_id == SS2_NET_SMOOTH) {
    SmoothWeights& m = ctx->smooth;
    m.ready = false;
    const std::string p = "MotionPre.";
    SS2_TRY(upload_key(ctx, net_id, p + "embedding1.0.weight", &m.emb1_w));
    SS2_TRY(upload_key(ctx, net_id, p + "embedding1.0.bias", &m.emb1_b));
    SS2_TRY(upload_key(ctx, net_id, p + "embedding3.0.weight", &m.emb3_w));
    SS2_TRY(upload_key(ctx, net_id, p + "embedding3.0.bias", &m.emb3_b));
    for (int i = 0; i < 3; ++i) {
      const std::string k = p + "MotionConv3D." + std::to_string(2 * i);
      SS2_TRY(pack_conv(ctx, net_id, k + ".weight", "", k + ".bias", 1, 1, 2, 0, &m.conv3d[i]));
    }
    SS2_TRY(upload_key(ctx, net_id, p + "decoding.0.weight", &m.dec_w));
    SS2_TRY(upload_key(ctx, net_id, p + "decoding.0.bias", &m.dec_b));
    m.ready = true;
  } else {
    return ss2_fail(ctx, SS2_ERR_INVALID, "unknown net id %d", net_id);
  }
  ctx->host_weights[net_id].clear();
  SS2_CUDA(ctx, cudaDeviceSynchronize());
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// forward graphs.  Activations are NHWC fp32 carved from the context arena.
// ------------------------------------------------------------------------------------------
#define ARENA(ptr, T, n)                                                                        \
  T* ptr = arena_alloc<T>(ctx, (size_t)(n));                                                    \
  if (!ptr) return ss2_fail(ctx, SS2_ERR_OOM, "workspace arena exhausted at %s:%d", __FILE__, __LINE__)

// activation with (when the tensor-core path is on) its tf32 hi/lo split planes
#define ARENA_ACT(ref, n)                                                                       \
  ActRef ref;                                                                                   \
  {                                                                                             \
    const size_t n__ = ((size_t)(n) + 63) / 64 * 64;                                            \
    ref.v = arena_alloc<float>(ctx, (ctx->use_tc ? 3 : 1) * n__);                               \
    if (!ref.v) return ss2_fail(ctx, SS2_ERR_OOM, "workspace arena exhausted at %s:%d", __FILE__, __LINE__); \
    if (ctx->use_tc) { ref.hi = ref.v + n__; ref.lo = ref.v + 2 * n__; }                        \
  }

// does conv_launch take a tensor-core kernel for this layer (they read only the hi/lo planes of their input)?
static bool runs_on_tc(ss2_ctx* ctx, const ConvLayer& L, bool in_has_split, int H, int W) {
  return ctx->use_tc && in_has_split && ((ctx->use_dc && conv_dc_eligible(L, 1, H, W)) || conv_tc_eligible(L));
}

// does the direct 3x3 kernel take this layer with fp16 split planes as its input (kind::f16 MMAs, conv_dc.cu)?
static bool dc16_ok(ss2_ctx* ctx, const ConvLayer& L, int H, int W) {
  return (ctx->use_f16 & 1) && ctx->use_tc && ctx->use_dc && ctx->tc_passes != 1 && L.wk_h16 != nullptr && conv_dc_eligible(L, 1, H, W);
}
// ... or the implicit-GEMM kernel (conv_tc.cu: stride-2 entries, 1x1 shortcuts, Conv3d, maps too small for the direct kernel)?
static bool tc16_ok(ss2_ctx* ctx, const ConvLayer& L) {
  return (ctx->use_f16 & 2) && ctx->use_tc && ctx->tc_passes != 1 && L.wk_h16 != nullptr && conv_tc_eligible(L);
}
// the kernel conv_launch picks for this layer reads fp16 planes
static bool layer16_ok(ss2_ctx* ctx, const ConvLayer& L, int H, int W) {
  if (ctx->use_tc && ctx->use_dc && conv_dc_eligible(L, 1, H, W)) return dc16_ok(ctx, L, H, W);
  return tc16_ok(ctx, L);
}
// a block whose input may arrive as plain values + fp16 split planes: every layer that reads the input's planes takes them
static bool block_takes_f16(ss2_ctx* ctx, const ResBlock& b, int H, int W) {
  return layer16_ok(ctx, b.c1, H, W) && (!b.has_down || layer16_ok(ctx, b.down, H, W));
}

// activation with plain values + fp16 split planes (8 bytes per value instead of 12)
#define ARENA_ACT16(ref, n)                                                                     \
  ActRef ref;                                                                                   \
  {                                                                                             \
    const size_t n__ = ((size_t)(n) + 63) / 64 * 64;                                            \
    ref.v = arena_alloc<float>(ctx, 2 * n__);                                                   \
    if (!ref.v) return ss2_fail(ctx, SS2_ERR_OOM, "workspace arena exhausted at %s:%d", __FILE__, __LINE__); \
    ref.h16 = reinterpret_cast<__half*>(ref.v + n__); ref.l16 = ref.h16 + n__;                  \
    ref.flag = ctx->d_range_flag;                                                               \
  }
// the one or the other
#define ARENA_ACT_SEL(ref, n, f16)                                                              \
  ActRef ref;                                                                                   \
  {                                                                                             \
    const size_t n__ = ((size_t)(n) + 63) / 64 * 64;                                            \
    const bool f__ = (f16);                                                                     \
    ref.v = arena_alloc<float>(ctx, (f__ ? 2 : (ctx->use_tc ? 3 : 1)) * n__);                   \
    if (!ref.v) return ss2_fail(ctx, SS2_ERR_OOM, "workspace arena exhausted at %s:%d", __FILE__, __LINE__); \
    if (f__) { ref.h16 = reinterpret_cast<__half*>(ref.v + n__); ref.l16 = ref.h16 + n__; ref.flag = ctx->d_range_flag; } \
    else if (ctx->use_tc) { ref.hi = ref.v + n__; ref.lo = ref.v + 2 * n__; }                   \
  }

// `next`: the block that consumes this block's output (null: the output's planes are TF32, for any consumer)
static int run_block(ss2_ctx* ctx, const ResBlock& b, const ActRef& x, int NB, int H, int W, ActRef* out, int* Ho,
                     int* Wo, cudaStream_t st, const ResBlock* next = nullptr) {
  int d, h, w;
  conv_out_dims(b.c1, 1, H, W, &d, &h, &w);
  const bool c1_f16 = x.hi == nullptr && x.h16 != nullptr;   // (the caller asked the producer for fp16 planes: block_takes_f16)
  if (c1_f16 && !block_takes_f16(ctx, b, H, W)) return ss2_fail(ctx, SS2_ERR_INVALID, "run_block: fp16 planes for a block that cannot read them");
  const bool c1_tc = c1_f16 || runs_on_tc(ctx, b.c1, x.hi != nullptr, H, W);
  // conv1's output feeds conv2 only.  When both run on the tensor cores the plain values are never read (v = hi + lo
  // exactly), so only the split planes are written: a third less store traffic for these layers.  When conv2 is a direct
  // 3x3 layer the planes are fp16 (4 bytes per value instead of 8) and conv2 runs kind::f16 MMAs: half the chunks, stages
  // and MMA instructions of the TF32 pair.
  ActRef t1;
  if (c1_tc && layer16_ok(ctx, b.c2, h, w)) {
    const size_t n1 = ((size_t)NB * h * w * b.c1.Cout + 63) / 64 * 64;
    ARENA(t1h, __half, 2 * n1);
    t1.h16 = t1h; t1.l16 = t1h + n1;
  } else if (c1_tc && runs_on_tc(ctx, b.c2, true, h, w)) {
    const size_t n1 = ((size_t)NB * h * w * b.c1.Cout + 63) / 64 * 64;
    ARENA(t1p, float, 2 * n1);
    t1.hi = t1p; t1.lo = t1p + n1;
  } else {
    ARENA_ACT(t1f, (size_t)NB * h * w * b.c1.Cout);
    t1 = t1f;
  }
  SS2_TRY(conv_launch(ctx, b.c1, x, NB, 1, H, W, t1, nullptr, 1, st));
  const float* idt = x.v;
  if (b.has_down) {
    ARENA_ACT(t2, (size_t)NB * h * w * b.down.Cout);
    ActRef t2v;
    t2v.v = t2.v;  // only the plain values of the shortcut are consumed (as the residual)
    SS2_TRY(conv_launch(ctx, b.down, x, NB, 1, H, W, t2v, nullptr, 0, st));
    idt = t2.v;
  }
  ActRef t3;
  if (next && block_takes_f16(ctx, *next, h, w) && (t1.h16 || t1.hi)) {
    ARENA_ACT16(t3h, (size_t)NB * h * w * b.c2.Cout);
    t3 = t3h;
  } else {
    ARENA_ACT(t3f, (size_t)NB * h * w * b.c2.Cout);
    t3 = t3f;
  }
  SS2_TRY(conv_launch(ctx, b.c2, t1, NB, 1, h, w, t3, idt, 1, st));
  *out = t3; *Ho = h; *Wo = w;
  return SS2_OK;
}

// stem as two kernels (implicit-GEMM tensor-core or SIMT 7x7 conv, then the pooling kernel): shapes the direct kernel
// does not take, and SS2_TC_STEM < 2
static int run_stem_pool(ss2_ctx* ctx, const Backbone& bb, const float* x_nchw, int NB, int H, int W, ActRef* out, int* Ho,
                         int* Wo, cudaStream_t st) {
  int d, h, w;
  conv_out_dims(bb.stem, 1, H, W, &d, &h, &w);
  ARENA(s, float, (size_t)NB * h * w * 64);
  ActRef xin, sout;
  sout.v = s;
  if (ctx->use_tc && ctx->use_tc_stem && bb.stem.stem_k32) {
    // tensor-core stem: zero-padded NHWC4 split planes, one filter row (8 pixels x 4 channels) per 32-wide K chunk
    const int Hp = H + 6, Wp = (2 * (w - 1) + 8 + 3) / 4 * 4;
    ARENA(xh, float, (size_t)NB * Hp * Wp * 4);
    ARENA(xl, float, (size_t)NB * Hp * Wp * 4);
    SS2_TRY(nchw_to_nhwc4_pad_split_launch(ctx, x_nchw, NB, H, W, 3, Hp, Wp, xh, xl, st));
    SS2_TRY(conv_tc_stem_launch(ctx, bb.stem, xh, xl, NB, H, W, Hp, Wp, sout, 1, st));
  } else {
    ARENA(x, float, (size_t)NB * H * W * 4);
    SS2_TRY(nchw_to_nhwc4_launch(ctx, x_nchw, NB, 3, H, W, x, st));
    xin.v = x;
    SS2_TRY(conv_launch(ctx, bb.stem, xin, NB, 1, H, W, sout, nullptr, 1, st));
  }
  const int hp = (h + 2 - 3) / 2 + 1, wp = (w + 2 - 3) / 2 + 1;
  ARENA_ACT(p, (size_t)NB * hp * wp * 64);
  SS2_TRY(maxpool_launch(ctx, s, NB, h, w, 64, 3, 2, 1, p, st));
  *out = p; *Ho = hp; *Wo = wp;
  return SS2_OK;
}

// x: NCHW [NB,3,H,W] -> f64 [NB,H/8,W/8,128] (and f32 [NB,.,.,256] when stage2)
static int run_backbone(ss2_ctx* ctx, const Backbone& bb, const float* x_nchw, int NB, int H, int W, bool stage2,
                        float** f64, int* h64, int* w64, float** f32, int* h32, int* w32, cudaStream_t st) {
  int d, h, w;
  conv_out_dims(bb.stem, 1, H, W, &d, &h, &w);
  ActRef cur;
  if (ctx->use_tc && ctx->use_tc_stem >= 2 && conv_stem_direct_eligible(bb.stem, H, W)) {
    // direct tensor-core stem with the max-pool in its epilogue: the conv map is never written
    ARENA(xw, float, conv_stem_direct_workspace_floats(NB, H));
    h = H / 4; w = W / 4;
    if (block_takes_f16(ctx, bb.l1[0], h, w)) {
      ARENA_ACT16(p, (size_t)NB * h * w * 64);
      SS2_TRY(conv_stem_pool_launch(ctx, bb.stem, x_nchw, NB, H, W, xw, p, st));
      cur = p;
    } else {
      ARENA_ACT(p, (size_t)NB * h * w * 64);
      SS2_TRY(conv_stem_pool_launch(ctx, bb.stem, x_nchw, NB, H, W, xw, p, st));
      cur = p;
    }
  } else {
    SS2_TRY(run_stem_pool(ctx, bb, x_nchw, NB, H, W, &cur, &h, &w, st));
  }
  SS2_TRY(run_block(ctx, bb.l1[0], cur, NB, h, w, &cur, &h, &w, st, &bb.l1[1]));
  SS2_TRY(run_block(ctx, bb.l1[1], cur, NB, h, w, &cur, &h, &w, st, &bb.l2[0]));
  SS2_TRY(run_block(ctx, bb.l2[0], cur, NB, h, w, &cur, &h, &w, st, &bb.l2[1]));
  SS2_TRY(run_block(ctx, bb.l2[1], cur, NB, h, w, &cur, &h, &w, st, stage2 ? &bb.l3[0] : nullptr));   // f64 is read through .v
  *f64 = cur.v; *h64 = h; *w64 = w;
  if (stage2) {
    SS2_TRY(run_block(ctx, bb.l3[0], cur, NB, h, w, &cur, &h, &w, st, &bb.l3[1]));
    SS2_TRY(run_block(ctx, bb.l3[1], cur, NB, h, w, &cur, &h, &w, st));
    *f32 = cur.v; *h32 = h; *w32 = w;
  }
  return SS2_OK;
}

// x [NB,H,W,CinP] -> out [NB, fc[2].Cout]
static int run_regressor(ss2_ctx* ctx, const Regressor& r, const ActRef& x, int NB, int H, int W, float* out,
                         cudaStream_t st) {
  ActRef cur = x;
  int h = H, w = W;
  for (size_t i = 0; i < r.convs.size(); ++i) {
    const ConvLayer& L = r.convs[i];
    const bool last = i + 1 == r.convs.size();
    const int hn = r.pool_after[i] ? h / 2 : h, wn = r.pool_after[i] ? w / 2 : w;   // the next layer's input size
    const bool next16 = !last && layer16_ok(ctx, r.convs[i + 1], hn, wn);           // it reads fp16 split planes
    // this layer runs a tcgen05 kernel (their epilogues write whichever planes are asked for, plain values optional)
    const bool tc = runs_on_tc(ctx, L, cur.hi != nullptr, h, w) || (cur.h16 != nullptr && layer16_ok(ctx, L, h, w));
    const size_t nel = (size_t)NB * h * w * L.Cout;
    ActRef t;
    if (tc && (r.pool_after[i] || last)) {
      // read by the pooling kernel / the Linear head through its plain values only: no split planes are written
      ARENA(tv, float, nel);
      t.v = tv;
    } else if (tc && next16) {
      // read by the next direct 3x3 layer through its fp16 planes only
      const size_t n1 = (nel + 63) / 64 * 64;
      ARENA(th, __half, 2 * n1);
      t.h16 = th; t.l16 = th + n1; t.flag = ctx->d_range_flag;
    } else {
      ARENA_ACT(tf, nel);
      t = tf;
    }
    SS2_TRY(conv_launch(ctx, L, cur, NB, 1, h, w, t, nullptr, 1, st));
    cur = t;
    if (r.pool_after[i]) {
      ARENA_ACT_SEL(q, (size_t)NB * hn * wn * L.Cout, next16);
      SS2_TRY(maxpool_launch(ctx, cur.v, NB, h, w, L.Cout, 2, 2, 0, q, st));
      cur = q; h = hn; w = wn;
    }
  }
  const int feat = h * w * r.convs.back().Cout;
  if (feat != r.fc[0].Cin) return ss2_fail(ctx, SS2_ERR_INVALID, "regressor: %d features, Linear expects %d", feat, r.fc[0].Cin);
  ARENA(a, float, (size_t)NB * r.fc[0].Cout);
  ARENA(b, float, (size_t)NB * r.fc[1].Cout);
  ActRef xa, xb, xo, xi;
  xi.v = cur.v; xa.v = a; xb.v = b; xo.v = out;
  SS2_TRY(conv_launch(ctx, r.fc[0], xi, NB, 1, 1, 1, xa, nullptr, 1, st));
  SS2_TRY(conv_launch(ctx, r.fc[1], xa, NB, 1, 1, 1, xb, nullptr, 1, st));
  SS2_TRY(conv_launch(ctx, r.fc[2], xb, NB, 1, 1, 1, xo, nullptr, 0, st));
  return SS2_OK;
}

#define NET_IMG_H 360
#define NET_IMG_W 480
static const size_t kBytesPerImageBackbone = (size_t)3 * 48 << 20;   // generous bound (x3: tf32 hi/lo planes), see DESIGN.md
static const size_t kBytesPerPairHead = (size_t)3 * 40 << 20;

static int spatial_chunk(ss2_ctx* ctx, const float* img1, const float* img2, int bs, int H, int W, float* o1,
                         float* oref, float* otgt, cudaStream_t st) {
  const SpatialWeights& S = ctx->spatial;
  ctx->arena.reset();
  // both views through the shared backbone as one batch of 2*bs images: [view1 | view2]
  ARENA(both, float, (size_t)2 * bs * 3 * H * W);
  SS2_CUDA(ctx, cudaMemcpyAsync(both, img1, (size_t)bs * 3 * H * W * 4, cudaMemcpyDeviceToDevice, st));
  SS2_CUDA(ctx, cudaMemcpyAsync(both + (size_t)bs * 3 * H * W, img2, (size_t)bs * 3 * H * W * 4,
                                cudaMemcpyDeviceToDevice, st));
  float *f64, *f32;
  int h64, w64, h32, w32;
  SS2_TRY(run_backbone(ctx, S.bb, both, 2 * bs, H, W, true, &f64, &h64, &w64, &f32, &h32, &w32, st));
  // stage 1: global correlation -> 4-point offsets
  ARENA(flow, float, (size_t)bs * h32 * w32 * 4);
  SS2_TRY(ccl_launch(ctx, f32, f32 + (size_t)bs * h32 * w32 * 256, bs, h32, w32, 256, flow, st));
  ActRef flow_ref;
  flow_ref.v = flow;
  SS2_TRY(run_regressor(ctx, S.r1, flow_ref, bs, h32, w32, o1, st));
  // homography split on the middle plane, warp both 1/8-scale feature maps
  ARENA(theta, float, (size_t)2 * bs * 9);
  SS2_TRY(spatial_split_launch(ctx, o1, bs, H, W, theta, theta + (size_t)bs * 9, st));
  ARENA(warped, float, (size_t)2 * bs * h64 * w64 * 128);
  SS2_TRY(homo_warp_nhwc_launch(ctx, f64, theta, 2 * bs, 128, h64, w64, warped, st));
  // stage 2: two local cost volumes -> two mesh regressors
  const size_t half = (size_t)bs * h64 * w64 * 128;
  ARENA_ACT_SEL(cv_ref, half, layer16_ok(ctx, S.r2_ref.convs[0], h64, w64));   // fp16 planes when the regressor's first layer reads them
  ARENA_ACT_SEL(cv_tgt, half, layer16_ok(ctx, S.r2_tgt.convs[0], h64, w64));
  SS2_TRY(cost_volume_launch(ctx, warped, warped + half, bs, h64, w64, 128, 5, 128, cv_ref, st));
  SS2_TRY(cost_volume_launch(ctx, warped + half, warped, bs, h64, w64, 128, 5, 128, cv_tgt, st));
  // The two mesh regressors are independent stacks of ~20 small launches each (a few dozen CTAs for 148 SMs, bound by
  // the length of one CTA's K loop): the target branch runs on a side stream next to the reference branch.  Their
  // workspace comes from the same arena (distinct regions); the caller's stream waits for the side stream before
  // anything else is enqueued, so the next arena reset is ordered after both.
  if (ctx->use_side && !ctx->prof[SS2_PROF_CONV].enabled) {
    if (!ctx->s_side) {
      SS2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking));
      SS2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      SS2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    SS2_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    SS2_CUDA(ctx, cudaStreamWaitEvent(ctx->s_side, ctx->ev_fork, 0));
    int rc = run_regressor(ctx, S.r2_ref, cv_ref, bs, h64, w64, oref, st);
    if (rc == SS2_OK) rc = run_regressor(ctx, S.r2_tgt, cv_tgt, bs, h64, w64, otgt, ctx->s_side);
    SS2_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->s_side));   // also after a failure: never leave the side stream detached
    SS2_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
    return rc;
  }
  SS2_TRY(run_regressor(ctx, S.r2_ref, cv_ref, bs, h64, w64, oref, st));
  SS2_TRY(run_regressor(ctx, S.r2_tgt, cv_tgt, bs, h64, w64, otgt, st));
  return SS2_OK;
}

// frames per launch train (SS2_SPATIAL_CHUNK / SS2_TEMPORAL_CHUNK override them for sweeps)
static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e && atoi(e) > 0 ? atoi(e) : dflt; }
// (measured, 32-pair step: 8 pairs per chunk 12.86 ms, 16 -> 11.30 ms, 32 -> 10.66 ms: the regressor stacks behind the
// backbone are a few dozen CTAs per launch, so their cost is per chunk, not per frame)
#define SPATIAL_CHUNK env_int("SS2_SPATIAL_CHUNK", 32)
#define TEMPORAL_CHUNK env_int("SS2_TEMPORAL_CHUNK", 32)
#define SMOOTH_CHUNK 256

extern "C" int ss2_spatial_forward(ss2_ctx* ctx, const float* d_img1, const float* d_img2, int bs,
                                   float* d_offset1, float* d_offset2_ref, float* d_offset2_tgt, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!ctx->spatial.ready) return ss2_fail(ctx, SS2_ERR_NO_WEIGHTS, "SpatialNet weights not finalized");
  if (bs < 0 || (bs > 0 && (!d_img1 || !d_img2 || !d_offset1 || !d_offset2_ref || !d_offset2_tgt)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_spatial_forward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = NET_IMG_H, W = NET_IMG_W;
  const int chunk = bs < SPATIAL_CHUNK ? bs : SPATIAL_CHUNK;
  SS2_TRY(ss2_workspace_enter(ctx, st));
  SS2_TRY(ss2_ensure_arena(ctx, (size_t)chunk * (2 * kBytesPerImageBackbone + kBytesPerPairHead) + ((size_t)64 << 20)));
  for (int b0 = 0; b0 < bs; b0 += chunk) {
    const int nb = bs - b0 < chunk ? bs - b0 : chunk;
    const size_t io = (size_t)b0 * 3 * H * W;
    SS2_TRY(spatial_chunk(ctx, d_img1 + io, d_img2 + io, nb, H, W, d_offset1 + (size_t)b0 * 8,
                          d_offset2_ref + (size_t)b0 * 126, d_offset2_tgt + (size_t)b0 * 126, st));
  }
  return SS2_OK;
}

extern "C" int ss2_spatial_tail(ss2_ctx* ctx, const float* d_offset1, const float* d_offset2_ref,
                                const float* d_offset2_tgt, int bs, int img_h, int img_w, float* d_motion1,
                                float* d_motion2, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  return spatial_tail_launch(ctx, d_offset1, d_offset2_ref, d_offset2_tgt, bs, img_h, img_w, d_motion1, d_motion2,
                             (cudaStream_t)stream);
}

extern "C" int ss2_build_spatial(ss2_ctx* ctx, const float* d_img1, const float* d_img2, int bs, float* d_motion1,
                                 float* d_motion2, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (bs <= 0) return bs == 0 ? SS2_OK : ss2_fail(ctx, SS2_ERR_INVALID, "negative batch");
  // the three offset vectors live in a small persistent scratch (not the arena: chunks reset it)
  float* off = nullptr;
  SS2_CUDA(ctx, cudaMallocAsync((void**)&off, (size_t)bs * (8 + 126 + 126) * sizeof(float), (cudaStream_t)stream));
  float *o1 = off, *oref = off + (size_t)bs * 8, *otgt = oref + (size_t)bs * 126;
  int rc = ss2_spatial_forward(ctx, d_img1, d_img2, bs, o1, oref, otgt, stream);
  if (rc == SS2_OK)
    rc = spatial_tail_launch(ctx, o1, oref, otgt, bs, NET_IMG_H, NET_IMG_W, d_motion1, d_motion2, (cudaStream_t)stream);
  cudaFreeAsync(off, (cudaStream_t)stream);
  return rc;
}

// frames [n,3,360,480] of one view -> motions [n,7,9,2]; motions[0] = 0
extern "C" int ss2_build_temporal(ss2_ctx* ctx, const float* d_frames, int n, float* d_motions, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!ctx->temporal.ready) return ss2_fail(ctx, SS2_ERR_NO_WEIGHTS, "TemporalNet weights not finalized");
  if (n < 0 || (n > 0 && (!d_frames || !d_motions))) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_build_temporal: bad arguments");
  if (n == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const TemporalWeights& T = ctx->temporal;
  const int H = NET_IMG_H, W = NET_IMG_W;
  SS2_CUDA(ctx, cudaMemsetAsync(d_motions, 0, (size_t)126 * sizeof(float), st));
  // n-1 motions; a chunk computes up to TEMPORAL_CHUNK motions from one more image (features of the frame before
  // the chunk are recomputed: 1-frame halo, so chunks stay independent).  Chunks are BALANCED: a temporal shard of
  // a stream hands over F+1 = 33 frames, which is one chunk of 33 images - not 32 + a second launch train over 2.
  const int nmot = n - 1;
  const int nchunks = nmot > 0 ? cdiv(nmot, TEMPORAL_CHUNK) : 0;
  const int per = nchunks > 0 ? cdiv(nmot, nchunks) : 0;   // motions per chunk (<= TEMPORAL_CHUNK)
  SS2_TRY(ss2_workspace_enter(ctx, st));
  SS2_TRY(ss2_ensure_arena(ctx, (size_t)(per + 1) * (kBytesPerImageBackbone + kBytesPerPairHead) + ((size_t)64 << 20)));
  for (int f0 = 1; f0 < n; f0 += per) {
    const int f1 = (f0 + per < n) ? f0 + per : n;   // motions f0..f1-1
    const int nimg = f1 - f0 + 1;                    // frames f0-1 .. f1-1
    ctx->arena.reset();
    float *f64, *f32 = nullptr;
    int h64, w64, h32, w32;
    SS2_TRY(run_backbone(ctx, T.bb, d_frames + (size_t)(f0 - 1) * 3 * H * W, nimg, H, W, false, &f64, &h64, &w64,
                         &f32, &h32, &w32, st));
    const int nm = nimg - 1;
    ARENA_ACT_SEL(cv, (size_t)nm * h64 * w64 * 64, layer16_ok(ctx, T.r2.convs[0], h64, w64));
    SS2_TRY(cost_volume_launch(ctx, f64, f64 + (size_t)h64 * w64 * 128, nm, h64, w64, 128, 3, 64, cv, st));
    SS2_TRY(run_regressor(ctx, T.r2, cv, nm, h64, w64, d_motions + (size_t)f0 * 126, st));
  }
  return SS2_OK;
}

// Both views of a stream through TemporalNet as ONE batch per chunk ([view a frames | view b frames] through the shared
// backbone, one regressor pass over both views' cost volumes): the regressor stack is a train of ~20 launches of a
// few dozen CTAs each, so its cost is per launch train, not per frame.  Results are bit-identical to two
// ss2_build_temporal calls (every kernel's per-image / per-row arithmetic is independent of the batch).
extern "C" int ss2_build_temporal_pair(ss2_ctx* ctx, const float* d_frames_a, const float* d_frames_b, int n,
                                       float* d_motions_a, float* d_motions_b, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!ctx->temporal.ready) return ss2_fail(ctx, SS2_ERR_NO_WEIGHTS, "TemporalNet weights not finalized");
  if (n < 0 || (n > 0 && (!d_frames_a || !d_frames_b || !d_motions_a || !d_motions_b)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_build_temporal_pair: bad arguments");
  if (n == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const TemporalWeights& T = ctx->temporal;
  const int H = NET_IMG_H, W = NET_IMG_W;
  SS2_CUDA(ctx, cudaMemsetAsync(d_motions_a, 0, (size_t)126 * sizeof(float), st));
  SS2_CUDA(ctx, cudaMemsetAsync(d_motions_b, 0, (size_t)126 * sizeof(float), st));
  const int nmot = n - 1;
  const int nchunks = nmot > 0 ? cdiv(nmot, TEMPORAL_CHUNK) : 0;
  const int per = nchunks > 0 ? cdiv(nmot, nchunks) : 0;   // motions per view per chunk (balanced, see ss2_build_temporal)
  SS2_TRY(ss2_workspace_enter(ctx, st));
  SS2_TRY(ss2_ensure_arena(ctx, (size_t)2 * (per + 1) * (kBytesPerImageBackbone + kBytesPerPairHead) + ((size_t)64 << 20)));
  const size_t ipx = (size_t)3 * H * W;
  for (int f0 = 1; f0 < n; f0 += per) {
    const int f1 = (f0 + per < n) ? f0 + per : n;   // motions f0..f1-1
    const int nimg = f1 - f0 + 1, nm = nimg - 1;     // frames f0-1 .. f1-1 of each view
    ctx->arena.reset();
    ARENA(both, float, (size_t)2 * nimg * ipx);
    SS2_CUDA(ctx, cudaMemcpyAsync(both, d_frames_a + (size_t)(f0 - 1) * ipx, (size_t)nimg * ipx * 4, cudaMemcpyDeviceToDevice, st));
    SS2_CUDA(ctx, cudaMemcpyAsync(both + (size_t)nimg * ipx, d_frames_b + (size_t)(f0 - 1) * ipx, (size_t)nimg * ipx * 4,
                                  cudaMemcpyDeviceToDevice, st));
    float *f64, *f32 = nullptr;
    int h64, w64, h32, w32;
    SS2_TRY(run_backbone(ctx, T.bb, both, 2 * nimg, H, W, false, &f64, &h64, &w64, &f32, &h32, &w32, st));
    const size_t fpx = (size_t)h64 * w64 * 128, cpx = (size_t)h64 * w64 * 64;
    ARENA_ACT_SEL(cv, (size_t)2 * nm * cpx, layer16_ok(ctx, T.r2.convs[0], h64, w64));
    ActRef cvb = cv;   // second view's half of the cost-volume batch
    cvb.v += (size_t)nm * cpx;
    if (cvb.hi) { cvb.hi += (size_t)nm * cpx; cvb.lo += (size_t)nm * cpx; }
    if (cvb.h16) { cvb.h16 += (size_t)nm * cpx; cvb.l16 += (size_t)nm * cpx; }
    SS2_TRY(cost_volume_launch(ctx, f64, f64 + fpx, nm, h64, w64, 128, 3, 64, cv, st));
    SS2_TRY(cost_volume_launch(ctx, f64 + (size_t)nimg * fpx, f64 + (size_t)(nimg + 1) * fpx, nm, h64, w64, 128, 3, 64, cvb, st));
    ARENA(mot, float, (size_t)2 * nm * 126);
    SS2_TRY(run_regressor(ctx, T.r2, cv, 2 * nm, h64, w64, mot, st));
    SS2_CUDA(ctx, cudaMemcpyAsync(d_motions_a + (size_t)f0 * 126, mot, (size_t)nm * 126 * 4, cudaMemcpyDeviceToDevice, st));
    SS2_CUDA(ctx, cudaMemcpyAsync(d_motions_b + (size_t)f0 * 126, mot + (size_t)nm * 126, (size_t)nm * 126 * 4, cudaMemcpyDeviceToDevice, st));
  }
  return SS2_OK;
}

// SpatialNet over n frame pairs and TemporalNet over both views (halo + n frames each; the first `halo` frames of
// d_lr1 / d_lr2 only feed TemporalNet: the frame before a temporal shard) as ONE call: the two networks are independent,
// and behind each backbone sits a train of ~20-40 launches of a few dozen CTAs, so TemporalNet runs on a stream of its own
// (workspace arena of its own) next to SpatialNet and the trains overlap.  Same kernels on the same data as the two
// separate calls: bit-identical outputs.  d_tm1 / d_tm2 [halo+n,7,9,2] (row 0 zero), d_sm1 / d_sm2 [n,7,9,2].
extern "C" int ss2_build_spatial_temporal(ss2_ctx* ctx, const float* d_lr1, const float* d_lr2, int n, int halo, float* d_sm1,
                                          float* d_sm2, float* d_tm1, float* d_tm2, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || halo < 0 || (n > 0 && (!d_lr1 || !d_lr2 || !d_sm1 || !d_sm2 || !d_tm1 || !d_tm2)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_build_spatial_temporal: bad arguments");
  if (n == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ipx = (size_t)3 * NET_IMG_H * NET_IMG_W;
  const float *sp1 = d_lr1 + (size_t)halo * ipx, *sp2 = d_lr2 + (size_t)halo * ipx;
  if (!ctx->use_net_overlap || ctx->prof[SS2_PROF_CONV].enabled || ctx->ws_nested) {
    SS2_TRY(ss2_build_spatial(ctx, sp1, sp2, n, d_sm1, d_sm2, stream));
    return ss2_build_temporal_pair(ctx, d_lr1, d_lr2, halo + n, d_tm1, d_tm2, stream);
  }
  SS2_TRY(ss2_workspace_enter(ctx, st));
  if (!ctx->s_net) {
    SS2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_net, cudaStreamNonBlocking));
    SS2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_nfork, cudaEventDisableTiming));
    SS2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_njoin, cudaEventDisableTiming));
  }
  SS2_CUDA(ctx, cudaEventRecord(ctx->ev_nfork, st));
  SS2_CUDA(ctx, cudaStreamWaitEvent(ctx->s_net, ctx->ev_nfork, 0));
  ctx->ws_nested = true;
  int rc = ss2_build_spatial(ctx, sp1, sp2, n, d_sm1, d_sm2, stream);
  std::swap(ctx->arena, ctx->arena_alt);
  const int rc_t = ss2_build_temporal_pair(ctx, d_lr1, d_lr2, halo + n, d_tm1, d_tm2, ctx->s_net);
  std::swap(ctx->arena, ctx->arena_alt);
  ctx->ws_nested = false;
  SS2_CUDA(ctx, cudaEventRecord(ctx->ev_njoin, ctx->s_net));   // also after a failure: never leave the side stream detached
  SS2_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_njoin, 0));
  return rc != SS2_OK ? rc : rc_t;
}

extern "C" int ss2_build_smooth(ss2_ctx* ctx, const float* d_ts1, const float* d_ts2, const float* d_sm1,
                                const float* d_sm2, int nwin, int zero_first, float* op1, float* sp1, float* om1, float* smm1,
                                float* op2, float* sp2, float* om2, float* smm2, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!ctx->smooth.ready) return ss2_fail(ctx, SS2_ERR_NO_WEIGHTS, "SmoothNet weights not finalized");
  if (nwin < 0 || (nwin > 0 && (!d_ts1 || !d_ts2 || !d_sm1 || !d_sm2)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_build_smooth: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const SmoothWeights& M = ctx->smooth;
  const size_t per_win = (size_t)SS2_WINDOW * SS2_NPT * 128 * sizeof(float);
  const int chunk = nwin < SMOOTH_CHUNK ? nwin : SMOOTH_CHUNK;
  SS2_TRY(ss2_workspace_enter(ctx, st));
  SS2_TRY(ss2_ensure_arena(ctx, (size_t)chunk * per_win * 8 + ((size_t)16 << 20)));
  for (int w0 = 0; w0 < nwin; w0 += chunk) {
    const int nw = nwin - w0 < chunk ? nwin - w0 : chunk;
    ctx->arena.reset();
    const size_t hid = (size_t)nw * SS2_WINDOW * SS2_NPT * 128;
    const bool s16 = tc16_ok(ctx, M.conv3d[0]) && tc16_ok(ctx, M.conv3d[1]) && tc16_ok(ctx, M.conv3d[2]);   // Conv3d on fp16 planes
    ARENA_ACT_SEL(h0, hid, s16);
    ARENA_ACT_SEL(h1, hid, s16);
    ARENA(p1, float, (size_t)nw * SS2_WINDOW * SS2_NPT * 2);
    ARENA(p2, float, (size_t)nw * SS2_WINDOW * SS2_NPT * 2);
    const size_t fo = (size_t)w0 * SS2_NPT * 2;  // frame offset of the first window of the chunk
    SS2_TRY(smooth_embed_launch(ctx, M, d_ts1 + fo, d_ts2 + fo, d_sm1 + fo, d_sm2 + fo, nw, zero_first, h0, p1, p2, st));
    SS2_TRY(conv_launch(ctx, M.conv3d[0], h0, nw, SS2_WINDOW, SS2_GRID_H + 1, SS2_GRID_W + 1, h1, nullptr, 1, st));
    SS2_TRY(conv_launch(ctx, M.conv3d[1], h1, nw, SS2_WINDOW, SS2_GRID_H + 1, SS2_GRID_W + 1, h0, nullptr, 1, st));
    SS2_TRY(conv_launch(ctx, M.conv3d[2], h0, nw, SS2_WINDOW, SS2_GRID_H + 1, SS2_GRID_W + 1, h1, nullptr, 1, st));
    const size_t oo = (size_t)w0 * SS2_WINDOW * SS2_NPT * 2;
    auto at = [&](float* p) { return p ? p + oo : nullptr; };
    SS2_TRY(smooth_decode_launch(ctx, M, h1.v, d_sm1 + fo, d_sm2 + fo, p1, p2, nw, at(op1), at(sp1), at(om1), at(smm1),
                                 at(op2), at(sp2), at(om2), at(smm2), st));
  }
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// generic NHWC / NDHWC convolution primitive (the layer type every network above is built of),
// exposed for unit tests and reuse.  Synchronous: packs the filter on every call.
// ------------------------------------------------------------------------------------------
__global__ void split_f16_kernel(const float* __restrict__ in, size_t n, __half* __restrict__ h, __half* __restrict__ l) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f16_split(in[i], &h[i], &l[i]);
}
__global__ void merge_f16_kernel(const __half* __restrict__ h, const __half* __restrict__ l, size_t n, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __half2float(h[i]) + __half2float(l[i]) * (1.0f / SS2_F16_LO_SCALE);
}
__global__ void split_tf32_kernel(const float* __restrict__ in, size_t n, float* __restrict__ hi, float* __restrict__ lo) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float h, l;
    tf32_split(in[i], &h, &l);
    hi[i] = h;
    lo[i] = l;
  }
}

extern "C" int ss2_conv_nhwc(ss2_ctx* ctx, const float* d_in, int B, int D, int H, int W, int Cin, const float* h_weight,
                             const int64_t* wshape, int wndim, const float* h_bias, int stride, int pad, int pad_d,
                             int relu, const float* d_residual, int use_tc, float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!d_in || !h_weight || !wshape || !d_out || (wndim != 4 && wndim != 5) || B <= 0 || D <= 0 || H <= 0 || W <= 0 ||
      wshape[1] != Cin || (Cin & 3))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_conv_nhwc: bad arguments (Cin must be a multiple of 4 and match the filter)");
  cudaStream_t st = (cudaStream_t)stream;
  SS2_TRY(ss2_workspace_enter(ctx, st));
  HostTensor wt, bt;
  wt.shape.assign(wshape, wshape + wndim);
  wt.data.assign(h_weight, h_weight + wt.numel());
  ctx->host_weights[0]["__conv.weight"] = wt;
  if (h_bias) {
    bt.shape = {wshape[0]};
    bt.data.assign(h_bias, h_bias + wshape[0]);
    ctx->host_weights[0]["__conv.bias"] = bt;
  }
  ConvLayer L;
  const size_t owned0 = ctx->owned.size();
  int rc = pack_conv(ctx, 0, "__conv.weight", "", h_bias ? "__conv.bias" : "", stride, pad, pad_d, 0, &L);
  ctx->host_weights[0].erase("__conv.weight");
  ctx->host_weights[0].erase("__conv.bias");
  float* split = nullptr;
  if (rc == SS2_OK) {
    ActRef in, out;
    in.v = const_cast<float*>(d_in);
    out.v = d_out;
    const size_t n = (size_t)B * D * H * W * Cin;
    const int saved = ctx->use_tc;
    ctx->use_tc = use_tc;
    if (use_tc) {
      if (!conv_tc_eligible(L)) rc = ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "ss2_conv_nhwc: layer not eligible for the tensor-core path");
      if (rc == SS2_OK && cudaMalloc((void**)&split, 2 * n * sizeof(float)) != cudaSuccess) rc = ss2_fail(ctx, SS2_ERR_OOM, "ss2_conv_nhwc: split buffer");
      if (rc == SS2_OK) {
        split_tf32_kernel<<<592, 256, 0, st>>>(d_in, n, split, split + n);
        in.hi = split;
        in.lo = split + n;
      }
    }
    // SS2_CONV_TEST_F16 (tests): 1 = the input arrives as fp16 split planes only (kind::f16 kernel), 2 = the output leaves as
    // fp16 split planes only and is merged back into d_out, 3 = both
    const int t16 = use_tc && getenv("SS2_CONV_TEST_F16") ? atoi(getenv("SS2_CONV_TEST_F16")) : 0;
    __half *in16 = nullptr, *out16 = nullptr;
    size_t on16 = 0;
    if (rc == SS2_OK && (t16 & 1)) {
      if (cudaMalloc((void**)&in16, 2 * n * sizeof(__half)) != cudaSuccess) rc = ss2_fail(ctx, SS2_ERR_OOM, "ss2_conv_nhwc: fp16 split buffer");
      else {
        split_f16_kernel<<<592, 256, 0, st>>>(d_in, n, in16, in16 + n);
        in.v = nullptr; in.hi = in.lo = nullptr;
        in.h16 = in16; in.l16 = in16 + n;
      }
    }
    if (rc == SS2_OK && (t16 & 2)) {
      int od, oh, ow;
      conv_out_dims(L, D, H, W, &od, &oh, &ow);
      on16 = (size_t)B * od * oh * ow * L.Cout;
      if (cudaMalloc((void**)&out16, 2 * on16 * sizeof(__half)) != cudaSuccess) rc = ss2_fail(ctx, SS2_ERR_OOM, "ss2_conv_nhwc: fp16 output buffer");
      else { out.v = nullptr; out.h16 = out16; out.l16 = out16 + on16; }
    }
    // SS2_CONV_TEST_SPLIT=1 (profiles/conv_bench.py): also write the hi/lo planes like a layer inside the networks does
    float* osplit = nullptr;
    if (rc == SS2_OK && use_tc && getenv("SS2_CONV_TEST_SPLIT") && atoi(getenv("SS2_CONV_TEST_SPLIT"))) {
      int od, oh, ow;
      conv_out_dims(L, D, H, W, &od, &oh, &ow);
      const size_t on = (size_t)B * od * oh * ow * L.Cout;
      if (cudaMalloc((void**)&osplit, 2 * on * sizeof(float)) == cudaSuccess) { out.hi = osplit; out.lo = osplit + on; }
    }
    if (rc == SS2_OK) rc = conv_launch(ctx, L, in, B, D, H, W, out, d_residual, relu, st);
    if (rc == SS2_OK && out16) merge_f16_kernel<<<592, 256, 0, st>>>(out16, out16 + on16, on16, d_out);
    ctx->use_tc = saved;
    cudaStreamSynchronize(st);
    if (osplit) cudaFree(osplit);
    if (in16) cudaFree(in16);
    if (out16) cudaFree(out16);
  }
  cudaStreamSynchronize(st);
  if (split) cudaFree(split);
  while (ctx->owned.size() > owned0) { cudaFree(ctx->owned.back()); ctx->owned.pop_back(); }
  return rc;
}

// Test / reuse entry for the network stem (conv 7x7 s2 p3 + bias + ReLU + max-pool 3x3 s2 p1), see include/ss2.h
extern "C" int ss2_stem_pool(ss2_ctx* ctx, const float* d_x_nchw, int B, int H, int W, const float* h_weight,
                             const float* h_bias, int variant, float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!d_x_nchw || !h_weight || !d_out || B <= 0 || H < 8 || W < 8 || variant < 0 || variant > 2)
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stem_pool: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  SS2_TRY(ss2_workspace_enter(ctx, st));
  HostTensor wt, bt;
  wt.shape = {64, 3, 7, 7};
  wt.data.assign(h_weight, h_weight + wt.numel());
  ctx->host_weights[0]["__stem.weight"] = wt;
  if (h_bias) {
    bt.shape = {64};
    bt.data.assign(h_bias, h_bias + 64);
    ctx->host_weights[0]["__stem.bias"] = bt;
  }
  Backbone bb;
  const size_t owned0 = ctx->owned.size();
  int rc = pack_conv(ctx, 0, "__stem.weight", "", h_bias ? "__stem.bias" : "", 2, 3, 0, 0, &bb.stem);
  ctx->host_weights[0].erase("__stem.weight");
  ctx->host_weights[0].erase("__stem.bias");
  const int saved_tc = ctx->use_tc, saved_stem = ctx->use_tc_stem;
  ctx->use_tc = variant > 0;
  ctx->use_tc_stem = variant;
  auto run = [&]() -> int {
    const int ho = ((H + 6 - 7) / 2 + 1 + 2 - 3) / 2 + 1, wo = ((W + 6 - 7) / 2 + 1 + 2 - 3) / 2 + 1;
    SS2_TRY(ss2_ensure_arena(ctx, (size_t)B * ((size_t)(H + 6) * (W + 16) * 8 + (size_t)(H / 2 + 3) * 8 * 1024 +
                                              (size_t)(H / 2 + 1) * (W / 2 + 1) * 64 + (size_t)3 * ho * wo * 64 + 4096) * sizeof(float) +
                                      ((size_t)4 << 20)));
    ctx->arena.reset();
    ActRef o;
    int h, w;
    if (variant == 2) {
      if (!conv_stem_direct_eligible(bb.stem, H, W))
        return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "ss2_stem_pool: the direct kernel takes W = 480, H % 4 == 0");
      ARENA(xw, float, conv_stem_direct_workspace_floats(B, H));
      ARENA_ACT(p, (size_t)B * (H / 4) * (W / 4) * 64);
      SS2_TRY(conv_stem_pool_launch(ctx, bb.stem, d_x_nchw, B, H, W, xw, p, st));
      o = p; h = H / 4; w = W / 4;
    } else {
      SS2_TRY(run_stem_pool(ctx, bb, d_x_nchw, B, H, W, &o, &h, &w, st));
    }
    SS2_CUDA(ctx, cudaMemcpyAsync(d_out, o.v, (size_t)B * h * w * 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SS2_OK;
  };
  if (rc == SS2_OK) rc = run();
  ctx->use_tc = saved_tc;
  ctx->use_tc_stem = saved_stem;
  cudaStreamSynchronize(st);
  while (ctx->owned.size() > owned0) { cudaFree(ctx->owned.back()); ctx->owned.pop_back(); }
  return rc;
}
