// Shared declarations for libss2.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/ss2.h"

#define SS2_NPT_PAD 64
#define SS2_NSYS 66  // 63 control points + affine part

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// One convolution / linear layer, packed for the implicit-GEMM kernels:
//   w [K = KD*KH*KW*CinP][CoutP] (CoutP = Cout rounded up to 64), bias [CoutP]
struct ConvLayer {
  float* w = nullptr;
  float* bias = nullptr;  // may be null
  int Cin = 0, CinP = 0, Cout = 0, CoutP = 0;
  int KD = 1, KH = 1, KW = 1;
  int sd = 1, sh = 1, sw = 1;  // strides
  int pd = 0, ph = 0, pw = 0;  // paddings
  // tcgen05 path: K-major filter matrices [CoutP][KD*KH*KW*CinP], hi = rna_tf32(w), lo = w - hi
  float* wk_hi = nullptr;
  float* wk_lo = nullptr;
  // stride-1 3x3 layers with CinP % 64 == 0: the same matrices as fp16 planes for the kind::f16 direct kernel (conv_dc.cu):
  // h16 = fp16(w), l16 = fp16((w - h16) * 2048)
  __half* wk_h16 = nullptr;
  __half* wk_l16 = nullptr;
  // 7x7 stride-2 stem with 3 input channels: wk_hi/lo are [CoutP][7 * 32], k = kh * 32 + kw * 4 + c (one filter row =
  // 8 NHWC4 pixels, the eighth and the fourth channel are zero); consumed by conv_tc_stem_launch
  bool stem_k32 = false;
  // the same stem for the direct kernel (conv_stem.cu): [64][7 * 32] = 25 k-steps of two pixels (+ 3 zero steps), order
  // in nets.cu pack_conv
  float* ws_hi = nullptr;
  float* ws_lo = nullptr;
};

// An activation tensor: plain fp32 values and (for the tensor-core layers that consume it) the
// split hi = rna_tf32(v), lo = v - hi.  hi/lo may be null.
// h16 / l16: the split as fp16 planes for a consumer that runs kind::f16 MMAs (conv_dc.cu): h16 = fp16(v),
// l16 = fp16((v - h16) * 2048); 11 + 11 significand bits like the TF32 pair.
struct ActRef {
  float* v = nullptr;
  float* hi = nullptr;
  float* lo = nullptr;
  __half* h16 = nullptr;
  __half* l16 = nullptr;
  int* flag = nullptr;   // fp16 range flag of the context (set with h16 / l16), raised by store_split* on overflow
};

struct ResBlock {
  ConvLayer c1, c2, down;
  bool has_down = false;
};

struct Backbone {  // ResNet-18 up to layer3
  ConvLayer stem;
  ResBlock l1[2], l2[2], l3[2];
};

struct Regressor {
  std::vector<ConvLayer> convs;
  std::vector<int> pool_after;  // 1 if a 2x2 max-pool follows conv i
  ConvLayer fc[3];
};

struct SpatialWeights {
  bool ready = false;
  Backbone bb;
  Regressor r1, r2_ref, r2_tgt;
};
struct TemporalWeights {
  bool ready = false;
  Backbone bb;
  Regressor r2;
};
struct SmoothWeights {
  bool ready = false;
  float* emb1_w = nullptr;  // [32][2]
  float* emb1_b = nullptr;
  float* emb3_w = nullptr;
  float* emb3_b = nullptr;
  ConvLayer conv3d[3];
  float* dec_w = nullptr;  // [4][128]
  float* dec_b = nullptr;
};

// Bump allocator over one cudaMalloc'd slab; reset at the start of every entry point.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    if (a + bytes > cap) return nullptr;
    off = a + bytes;
    return base + a;
  }
  void reset() { off = 0; }
};

struct ProfClass {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;   // start/stop pairs
  size_t used = 0;
  double work = 0.0;               // bytes or flops, accumulated by the launcher
};

struct ss2_ctx {
  int device = 0;
  std::string err;
  int64_t launches = 0;
  std::map<std::string, HostTensor> host_weights[3];
  std::vector<void*> owned;  // device allocations freed at destroy
  std::vector<void*> owned_net[3];  // packed weights per network: freed when that network is re-finalized
  int packing_net = -1;             // network whose weights upload() is packing (-1: ctx->owned)
  SpatialWeights spatial;
  TemporalWeights temporal;
  SmoothWeights smooth;
  Arena arena;
  std::map<std::string, std::pair<void*, size_t>> stream_bufs;  // named persistent device buffers
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_hr = nullptr, ev_chunk[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  ProfClass prof[SS2_PROF_COUNT];
  // the arena and the named buffers are shared workspace: when a call arrives on another stream than the previous
  // one, it first waits (event) for everything the previous stream was given (ss2_workspace_enter)
  cudaStream_t ws_stream = nullptr;
  bool ws_used = false;
  cudaEvent_t ws_ev = nullptr;
  // side stream of independent network branches (SpatialNet's two mesh regressors): fork / join events (nets.cu)
  cudaStream_t s_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int use_side = 1;     // SS2_SIDE_STREAM=0: everything on the caller's stream
  // SpatialNet and TemporalNet of a chunk are independent until the tsmotion step: TemporalNet runs on a stream of its
  // own with a workspace arena of its own next to SpatialNet (ss2_stream_meshes, ss2_build_spatial_temporal)
  Arena arena_alt;
  cudaStream_t s_net = nullptr;
  cudaEvent_t ev_nfork = nullptr, ev_njoin = nullptr;
  int use_net_overlap = 1;   // SS2_NET_OVERLAP=0: one network after the other on the caller's stream
  bool ws_nested = false;    // inside such a fork: the inner entry points must not re-serialise the streams
  void* host_slots = nullptr;  // HostSlot[HOST_SLOTS] of the host-buffer pipeline (stream.cu), created on first use
  int use_tc = 1;  // tcgen05 implicit-GEMM path for eligible layers
  bool lag_tables_ready = false;
  bool gauss_ready = false;  // linear.cu: 21-tap Gaussian in constant memory
  int tc_passes = 3;  // 3 = split-TF32 (fp32-grade), 1 = plain TF32
  int use_tc_stem = 2;  // SS2_TC_STEM: 2 = direct tensor-core stem with the max-pool fused (conv_stem.cu), 1 = implicit-GEMM
                        // tensor-core stem + pool kernel, 0 = exact-fp32 SIMT stem + pool kernel
  // SS2_F16=0: TF32 split planes everywhere.  Default: the stride-1 3x3 layers of the ResNet bodies read fp16 split planes
  // and run kind::f16 MMAs (conv_dc.cu); their producers' epilogues write those planes.  fp16 holds |v| <= 65504: an
  // epilogue that meets a larger value raises *range_flag (mapped pinned host memory); the next entry point fails loudly.
  int use_f16 = 3;   // bit 0: the direct 3x3 kernel, bit 1: the implicit-GEMM kernel (stride-2 entries, shortcuts, Conv3d, CCL)
  int* h_range_flag = nullptr;   // host view
  int* d_range_flag = nullptr;   // device view of the same word
  int use_dc = 1;     // direct 3x3 kernel (conv_dc.cu) for eligible layers; SS2_CONV_DC=0 disables
};

int ss2_fail(ss2_ctx* ctx, int code, const char* fmt, ...);
int ss2_workspace_enter(ss2_ctx* ctx, cudaStream_t st);
void ss2_host_slots_free(ss2_ctx* ctx);  // stream.cu: streams / events of the host pipeline
int ss2_ensure_arena(ss2_ctx* ctx, size_t bytes);
// profiling brackets: no-ops unless the class is enabled
void ss2_prof_begin(ss2_ctx* ctx, int which, cudaStream_t st);
void ss2_prof_end(ss2_ctx* ctx, int which, cudaStream_t st, double work);

#define SS2_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return ss2_fail(ctx, SS2_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,      \
                      cudaGetErrorString(e__));                                          \
  } while (0)

#define SS2_LAUNCH_CHECK(ctx)                                                            \
  do {                                                                                   \
    (ctx)->launches++;                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return ss2_fail(ctx, SS2_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__,         \
                      cudaGetErrorString(e__));                                          \
  } while (0)

#define SS2_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != SS2_OK) return rc__; \
  } while (0)

template <typename T>
static inline T* arena_alloc(ss2_ctx* ctx, size_t n) {
  return reinterpret_cast<T*>(ctx->arena.alloc(n * sizeof(T)));
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// hi = rna_tf32(v), lo = v - hi (exact); v == hi + lo
__device__ __forceinline__ void tf32_split(float v, float* hi, float* lo) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  *hi = __uint_as_float(r);
  *lo = v - *hi;
}
// fp16 split of the kind::f16 path: h = fp16(v), l = fp16((v - h) * 2^11): v = h + l / 2048 to 2^-22 relative (to 2^-36
// absolute below fp16's normal range, where the scaled l recovers what the subnormal h loses).  |v| > 65504 does not fit
// (h = inf): the writers below raise the context's range flag for such a value and the next entry point fails loudly.
#define SS2_F16_LO_SCALE 2048.0f
__device__ __forceinline__ void f16_split(float v, __half* h, __half* l) {
  const __half hh = __float2half_rn(v);
  *h = hh;
  *l = __float2half_rn((v - __half2float(hh)) * SS2_F16_LO_SCALE);
}
// four consecutive channels of the two fp16 planes (8 bytes each), packed conversions (cvt.rn.f16x2.f32)
__device__ __forceinline__ void store_f16_planes4(__half* h16, __half* l16, size_t o, const float (&v)[4], int* range_flag) {
  if (fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))) > 65504.0f) *range_flag = 1;   // (NaN compares false)
  const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn((v[0] - f01.x) * SS2_F16_LO_SCALE, (v[1] - f01.y) * SS2_F16_LO_SCALE);
  const __half2 l23 = __floats2half2_rn((v[2] - f23.x) * SS2_F16_LO_SCALE, (v[3] - f23.y) * SS2_F16_LO_SCALE);
  *reinterpret_cast<uint2*>(h16 + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  *reinterpret_cast<uint2*>(l16 + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}
__device__ __forceinline__ void store_split4(const ActRef& o, size_t idx, float4 v) {
  *reinterpret_cast<float4*>(o.v + idx) = v;
  if (o.hi) {
    float4 h, l;
    tf32_split(v.x, &h.x, &l.x); tf32_split(v.y, &h.y, &l.y); tf32_split(v.z, &h.z, &l.z); tf32_split(v.w, &h.w, &l.w);
    *reinterpret_cast<float4*>(o.hi + idx) = h;
    if (o.lo) *reinterpret_cast<float4*>(o.lo + idx) = l;
  }
  if (o.h16) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
    store_f16_planes4(o.h16, o.l16, idx, vv, o.flag);
  }
}
__device__ __forceinline__ void store_split1(const ActRef& o, size_t idx, float v) {
  o.v[idx] = v;
  if (o.hi) {
    float h, l;
    tf32_split(v, &h, &l);
    o.hi[idx] = h;
    if (o.lo) o.lo[idx] = l;
  }
  if (o.h16) {
    if (fabsf(v) > 65504.0f) *o.flag = 1;
    f16_split(v, &o.h16[idx], &o.l16[idx]);
  }
}
#endif

// ---- kernels' host launchers (defined in the .cu files) ------------------------------------
// tps.cu
int tps_solve_launch(ss2_ctx* ctx, const float* d_source, const float* d_target, int bn, float* d_T,
                     cudaStream_t st);
int tps_point_launch(ss2_ctx* ctx, const float* d_point, const float* d_source, const float* d_T, int bn,
                     float* d_out, cudaStream_t st);
int tps_solve_aux_launch(ss2_ctx* ctx, const float* d_source, const float* d_target, int bn, float* d_T, float* d_aux,
                         float half_w, float half_h, int Ho, int Wo, cudaStream_t st);
size_t tps_lattice_workspace_floats(int bn, int Ho, int Wo);
bool tps_lattice_supported(int Ho, int Wo);
// d_aux [bn][8] + d_nodes (tps_lattice_workspace_floats) are needed by tps == SS2_TPS_LATTICE
int tps_warp_launch(ss2_ctx* ctx, const float* d_U, const float* d_source, const float* d_T, int bn, int C,
                    int H, int W, int Ho, int Wo, int mode, int tps, float* d_out, cudaStream_t st,
                    const float* d_aux, float* d_nodes);
int tps_warp_blend_launch(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_source,
                          const float* d_T, int nframes, int H, int W, int Ho, int Wo, int mode, int tps,
                          float* d_out, cudaStream_t st, const float* d_aux, float* d_nodes, unsigned char* d_out8 = nullptr);
// one stream-ordered allocation holding T [bn][2][66], aux [bn][8] and the lattice nodes
struct TpsScratch {
  float* base = nullptr;
  float *T = nullptr, *aux = nullptr, *nodes = nullptr;
};
int tps_scratch_alloc(ss2_ctx* ctx, int bn, int Ho, int Wo, int tps, size_t extra_floats, TpsScratch* s, cudaStream_t st);
int tps_solve_for_warp(ss2_ctx* ctx, const float* d_source, const float* d_target, int bn, int H, int W, int Ho, int Wo,
                       int mode, int tps, const TpsScratch& s, cudaStream_t st);
// geom.cu
int dlt_launch(ss2_ctx* ctx, const float* d_src, const float* d_dst, int bs, float* d_H, cudaStream_t st);
int homo_warp_nchw_launch(ss2_ctx* ctx, const float* d_U, const float* d_theta, int bn, int C, int H, int W,
                          int Ho, int Wo, float* d_out, cudaStream_t st);
int homo_warp_nhwc_launch(ss2_ctx* ctx, const float* d_U, const float* d_theta, int bn, int C, int H, int W,
                          float* d_out, cudaStream_t st);
int spatial_split_launch(ss2_ctx* ctx, const float* d_offset1, int bs, int img_h, int img_w,
                         float* d_theta_ref, float* d_theta_tgt, cudaStream_t st);
int spatial_tail_launch(ss2_ctx* ctx, const float* d_o1, const float* d_oref, const float* d_otgt, int bs,
                        int img_h, int img_w, float* d_m1, float* d_m2, cudaStream_t st);
int tsmotion_prep_launch(ss2_ctx* ctx, const float* d_smotion, const float* d_tmotion, int n, int first,
                         const float* d_prev, float* d_smesh, float* d_point, float* d_source,
                         float* d_target, cudaStream_t st);
int tsmotion_finish_launch(ss2_ctx* ctx, const float* d_moved, const float* d_smesh, int n, int first,
                           float* d_tsmotion, cudaStream_t st);
int canvas_minmax_launch(ss2_ctx* ctx, const float* d_mesh1, const float* d_mesh2, int n, int img_h,
                         int img_w, float* d_minmax, cudaStream_t st);
int stable_meshes_launch(ss2_ctx* ctx, const float* d_mesh1, const float* d_mesh2, int n, int img_h,
                         int img_w, float xmin, float ymin, float out_w, float out_h, float* d_source,
                         float* d_target, cudaStream_t st);
// geom.cu: three-view glue (test_online_tra_threeview.py:345-505)
int three_view_align_launch(ss2_ctx* ctx, const float* w12m1, const float* w12m2, const float* w23m1, const float* w23m2,
                            int n, int img_h, int img_w, float* work, float* pt12, float* src12, float* pt23, float* src23,
                            float* tgt, float* mid_c, float* canvas1, cudaStream_t st);
int three_view_canvas_launch(ss2_ctx* ctx, const float* moved12, const float* moved23, const float* mid_c,
                             const float* canvas1, int n, float* m1_c, float* m3_c, float* canvas2, cudaStream_t st);
int three_view_sources_launch(ss2_ctx* ctx, const float* m1_c, const float* mid_c, const float* m3_c, int n, int img_h,
                              int img_w, float xmin, float ymin, float out_w, float out_h, float* source, float* target,
                              cudaStream_t st);
int blend3_avg_launch(ss2_ctx* ctx, const float* w1, const float* w2, const float* w3, size_t count, float* out,
                      cudaStream_t st);
// geom.cu: N-view middle-plane chain (config 5)
int nview_align_launch(ss2_ctx* ctx, const float* const* d_pairs, int nviews, int n, int img_h, int img_w, float* shifted,
                       float* mids, float* minmax, cudaStream_t st);
int nview_operands_launch(ss2_ctx* ctx, const float* shifted, const float* mids, int nviews, int n, float xmin, float ymin,
                          float ow, float oh, float* pt0, float* src0, float* tgt0, float* pt1, float* src1, float* tgt1,
                          float* meshes_out, cudaStream_t st);
int nview_finish_launch(ss2_ctx* ctx, const float* moved0, const float* moved1, int nviews, int n, float ow, float oh,
                        float* meshes_out, float* minmax, cudaStream_t st);
int nview_sources_launch(ss2_ctx* ctx, const float* meshes, int nviews, int n, int img_h, int img_w, float xmin, float ymin,
                         float out_w, float out_h, float* source, float* target, cudaStream_t st);
// tps.cu: fused N-view (2..4) resample + sequential AVERAGE blend; imgs[v] [nframes,3,H,W], source [nframes][V][63][2]
int tps_warp_blend_n_launch(ss2_ctx* ctx, const float* const* d_imgs, int nviews, const float* d_source, const float* d_T,
                            int nframes, int H, int W, int Ho, int Wo, int mode, int tps, float* d_out, cudaStream_t st,
                            const float* d_aux, float* d_nodes);
// conv_tc.cu: cuTensorMapEncodeTiled (driver entry point), null if unavailable
void* ss2_tensormap_encode_fn();
// conv.cu
int conv_launch(ss2_ctx* ctx, const ConvLayer& L, const ActRef& in, int B, int D, int H, int W, const ActRef& out,
                const float* d_residual, int relu, cudaStream_t st, int groups = 1, size_t w_group_stride = 0);
bool conv_tc_eligible(const ConvLayer& L);
int conv_tc_corr_rows(int W);
int conv_tc_corr_launch(ss2_ctx* ctx, const ActRef& n1, const ActRef& n2, int B, int H, int W, int C, float* d_match,
                        int ldo, cudaStream_t st);
int conv_tc_launch(ss2_ctx* ctx, const ConvLayer& L, const ActRef& in, int B, int D, int H, int W, const ActRef& out,
                   const float* d_residual, int relu, cudaStream_t st);
// conv_dc.cu: direct 3x3 stride-1 convolution (one staged input tile, nine shifted tap descriptors)
bool conv_dc_eligible(const ConvLayer& L, int D, int H, int W);
int conv_dc_launch(ss2_ctx* ctx, const ConvLayer& L, const ActRef& in, int B, int H, int W, const ActRef& out,
                   const float* d_residual, int relu, cudaStream_t st);
void conv_out_dims(const ConvLayer& L, int D, int H, int W, int* Do, int* Ho, int* Wo);
int maxpool_launch(ss2_ctx* ctx, const float* d_in, int B, int H, int W, int C, int k, int s, int p,
                   const ActRef& out, cudaStream_t st);
int nchw_to_nhwc4_pad_split_launch(ss2_ctx* ctx, const float* d_in, int B, int H, int W, int pad, int Hp, int Wp,
                                   float* d_hi, float* d_lo, cudaStream_t st);
// conv_tc.cu: 7x7 stride-2 stem on the tensor cores from the padded split planes above; out [B, Ho, Wo, 64] (ReLU)
int conv_tc_stem_launch(ss2_ctx* ctx, const ConvLayer& L, const float* d_hi, const float* d_lo, int B, int H, int W,
                        int Hp, int Wp, const ActRef& out, int relu, cudaStream_t st);
// conv_stem.cu: direct 7x7 stride-2 stem + ReLU + 3x3 stride-2 max-pool in one kernel; x NCHW [B,3,H,W] -> out [B,H/4,W/4,64]
bool conv_stem_direct_eligible(const ConvLayer& L, int H, int W);
size_t conv_stem_direct_workspace_floats(int B, int H);
int conv_stem_pool_launch(ss2_ctx* ctx, const ConvLayer& L, const float* d_x_nchw, int B, int H, int W, float* d_work,
                          const ActRef& out, cudaStream_t st);
int nchw_to_nhwc4_launch(ss2_ctx* ctx, const float* d_in, int B, int C, int H, int W, float* d_out,
                         cudaStream_t st);
// corr.cu
int cost_volume_launch(ss2_ctx* ctx, const float* d_x1, const float* d_x2, int B, int H, int W, int C, int sr,
                       int CP, const ActRef& out, cudaStream_t st);
int ccl_launch(ss2_ctx* ctx, const float* d_f1, const float* d_f2, int B, int H, int W, int C, float* d_flow,
               cudaStream_t st);
// edges.cu: uint8 host edges on the device (cv2-exact resize, fp32 <-> uint8 layout conversions)
int load_frames_u8_launch(ss2_ctx* ctx, const unsigned char* d_u8, int n, int H, int W, float* d_hr, float* d_lr,
                          cudaStream_t st);
int frames_to_u8_launch(ss2_ctx* ctx, const float* d_frames, int n, int Ho, int Wo, unsigned char* d_out, cudaStream_t st);
// smooth.cu
int smooth_embed_launch(ss2_ctx* ctx, const SmoothWeights& sw, const float* ts1, const float* ts2,
                        const float* sm1, const float* sm2, int nwin, int zero_first, const ActRef& hidden, float* d_path1,
                        float* d_path2, cudaStream_t st);
int smooth_decode_launch(ss2_ctx* ctx, const SmoothWeights& sw, const float* d_hidden, const float* sm1,
                         const float* sm2, const float* path1, const float* path2, int nwin, float* op1,
                         float* sp1, float* om1, float* smm1, float* op2, float* sp2, float* om2, float* smm2,
                         cudaStream_t st);
