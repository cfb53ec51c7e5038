// LINEAR fusion: the reference driver's linear_blender
// (Full_model_inference/Codes/test_online_tra.py:34-58) and the LINEAR branch of get_stable_sqe (:143-150).
//
//   centres   c_k = mean (row, col) over the pixels of mask k                                   (:36-40)
//   ramp      over the overlap, the projection of (pixel - c_1) on (c_2 - c_1), scaled to [0, 1) (:44-50)
//   mask1     clamp(blur(ref_only + (1 - ramp) * m1) * m1 + ref_only, 0, 1), blur = 21x21 sigma-20 Gaussian,
//             reflect padding (torchvision GaussianBlur)                                         (:35,52)
//   out       ref * mask1 + tgt * (1 - mask1) * m2                                               (:55-56)
//
// Semantics of the masks.  The reference takes torch.nonzero of the WARPED mask channel, whose values outside the
// image are rounding residues of four clamped taps (about half of them non-zero): its centres include ~12 % residue
// pixels (tests/golden/linear.npz: the column of a centre moves by 24 px on the 184x436 test canvas).  This path is
// residue free: a mask is the set of pixels with warped value > 0.5 (the lattice resampler returns exactly 0 outside
// and 1 +- 1 ulp inside).  Measured distance to the reference's own LINEAR frames: tests/test_gpu_parity.py
// (test_linear_stream_vs_reference_golden); on exact 0/1 masks the two semantics coincide and the kernels are checked
// against the reference's linear_blender output at 1e-4.
//
// Four launches per chunk of frames (grid z = frame): centre sums (exact integer atomics), projection min / max,
// horizontal blur, vertical blur + blend.  Deterministic: integer sums and min / max do not depend on the order.
#include <math.h>

#include "common.cuh"

#define LB_TAPS 21
#define LB_HALF 10

struct LinearAcc {           // per frame
  unsigned long long cnt1, sr1, sc1, cnt2, sr2, sc2;
  unsigned int pmin, pmax;   // order-preserving keys of the float projection
  unsigned int pad[2];
};

__constant__ float c_gauss[LB_TAPS];

__device__ __forceinline__ unsigned int float_key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void linear_init_kernel(LinearAcc* acc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    LinearAcc a;
    a.cnt1 = a.sr1 = a.sc1 = a.cnt2 = a.sr2 = a.sc2 = 0ull;
    a.pmin = 0xffffffffu; a.pmax = 0u;
    a.pad[0] = a.pad[1] = 0u;
    acc[i] = a;
  }
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// centre sums: one CTA per canvas row
__global__ void linear_centroid_kernel(const float* __restrict__ m1, const float* __restrict__ m2, size_t mstride, int Wo,
                                       LinearAcc* __restrict__ acc) {
  const int f = blockIdx.z, r = blockIdx.x;
  const float* a = m1 + (size_t)f * mstride + (size_t)r * Wo;
  const float* b = m2 + (size_t)f * mstride + (size_t)r * Wo;
  unsigned long long n1 = 0, c1 = 0, n2 = 0, c2 = 0;
  for (int c = threadIdx.x; c < Wo; c += blockDim.x) {
    if (a[c] > 0.5f) { n1 += 1; c1 += (unsigned)c; }
    if (b[c] > 0.5f) { n2 += 1; c2 += (unsigned)c; }
  }
  n1 = warp_sum_u64(n1); c1 = warp_sum_u64(c1); n2 = warp_sum_u64(n2); c2 = warp_sum_u64(c2);
  if ((threadIdx.x & 31) == 0) {
    LinearAcc* A = acc + f;
    if (n1) { atomicAdd(&A->cnt1, n1); atomicAdd(&A->sr1, n1 * (unsigned long long)r); atomicAdd(&A->sc1, c1); }
    if (n2) { atomicAdd(&A->cnt2, n2); atomicAdd(&A->sr2, n2 * (unsigned long long)r); atomicAdd(&A->sc2, c2); }
  }
}

struct LinearGeo { float c1r, c1c, vr, vc; };
__device__ __forceinline__ LinearGeo linear_geo(const LinearAcc& A) {
  LinearGeo g;
  // mean over an empty set is NaN in the reference; an empty mask simply never enters the ramp here
  const double n1 = A.cnt1 ? (double)A.cnt1 : 1.0, n2 = A.cnt2 ? (double)A.cnt2 : 1.0;
  g.c1r = (float)((double)A.sr1 / n1);
  g.c1c = (float)((double)A.sc1 / n1);
  g.vr = __fsub_rn((float)((double)A.sr2 / n2), g.c1r);
  g.vc = __fsub_rn((float)((double)A.sc2 / n2), g.c1c);
  return g;
}
// (r - center1[0]) * vec[0] + (c - center1[1]) * vec[1] in the reference's fp32 operation order (:49)
__device__ __forceinline__ float linear_proj(const LinearGeo& g, int r, int c) {
  return __fadd_rn(__fmul_rn(__fsub_rn((float)r, g.c1r), g.vr), __fmul_rn(__fsub_rn((float)c, g.c1c), g.vc));
}

__global__ void linear_proj_minmax_kernel(const float* __restrict__ m1, const float* __restrict__ m2, size_t mstride,
                                          int Wo, LinearAcc* __restrict__ acc) {
  const int f = blockIdx.z, r = blockIdx.x;
  const LinearGeo g = linear_geo(acc[f]);
  const float* a = m1 + (size_t)f * mstride + (size_t)r * Wo;
  const float* b = m2 + (size_t)f * mstride + (size_t)r * Wo;
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (int c = threadIdx.x; c < Wo; c += blockDim.x) {
    if (a[c] > 0.5f && b[c] > 0.5f) {
      const unsigned int k = float_key(linear_proj(g, r, c));
      lo = min(lo, k); hi = max(hi, k);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(&acc[f].pmin, lo);
    atomicMax(&acc[f].pmax, hi);
  }
}

__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// horizontal pass over A = ref_only + (1 - ramp) * m1, one CTA per row; tmp [n][Ho][Wo]
__global__ void linear_hblur_kernel(const float* __restrict__ m1, const float* __restrict__ m2, size_t mstride, int Wo,
                                    const LinearAcc* __restrict__ acc, float* __restrict__ tmp) {
  extern __shared__ float row[];   // A of this canvas row
  const int f = blockIdx.z, r = blockIdx.x, Ho = gridDim.x;
  const LinearAcc A = acc[f];
  const LinearGeo g = linear_geo(A);
  const bool has = A.pmin <= A.pmax;
  const float pmin = has ? key_float(A.pmin) : 0.f, pmax = has ? key_float(A.pmax) : 0.f;
  const float den = __fadd_rn(__fsub_rn(pmax, pmin), 1e-3f);
  const float* a = m1 + (size_t)f * mstride + (size_t)r * Wo;
  const float* b = m2 + (size_t)f * mstride + (size_t)r * Wo;
  for (int c = threadIdx.x; c < Wo; c += blockDim.x) {
    const float x1 = a[c] > 0.5f ? 1.f : 0.f, x2 = b[c] > 0.5f ? 1.f : 0.f;
    const float ovl = x1 * x2;
    const float ramp = ovl > 0.f ? __fdiv_rn(__fsub_rn(linear_proj(g, r, c), pmin), den) : 0.f;
    row[c] = __fadd_rn(__fsub_rn(x1, ovl), __fmul_rn(__fsub_rn(1.f, ramp), x1));
  }
  __syncthreads();
  float* o = tmp + ((size_t)f * Ho + r) * Wo;
  for (int c = threadIdx.x; c < Wo; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LB_TAPS; ++k) s = fmaf(c_gauss[k], row[reflect(c + k - LB_HALF, Wo)], s);
    o[c] = s;
  }
}

// vertical pass + mask1 / mask2 + blend; 32 x 8 pixel tiles
__global__ void linear_vblur_blend_kernel(const float* __restrict__ ref, const float* __restrict__ tgt, size_t istride,
                                          const float* __restrict__ m1, const float* __restrict__ m2, size_t mstride,
                                          const float* __restrict__ tmp, int Ho, int Wo, float* __restrict__ out,
                                          float* __restrict__ mask1_out) {
  const int f = blockIdx.z, c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
  if (c >= Wo || r >= Ho) return;
  const float* t = tmp + (size_t)f * Ho * Wo + c;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LB_TAPS; ++k) s = fmaf(c_gauss[k], t[(size_t)reflect(r + k - LB_HALF, Ho) * Wo], s);
  const size_t p = (size_t)r * Wo + c, plane = (size_t)Ho * Wo;
  const float x1 = m1[(size_t)f * mstride + p] > 0.5f ? 1.f : 0.f, x2 = m2[(size_t)f * mstride + p] > 0.5f ? 1.f : 0.f;
  const float ref_only = __fsub_rn(x1, x1 * x2);
  const float k1 = fminf(fmaxf(__fadd_rn(__fmul_rn(s, x1), ref_only), 0.f), 1.f);
  const float k2 = __fmul_rn(__fsub_rn(1.f, k1), x2);
  if (mask1_out) mask1_out[(size_t)f * plane + p] = k1;
  if (out) {
    const float* a = ref + (size_t)f * istride + p;
    const float* b = tgt + (size_t)f * istride + p;
    float* o = out + (size_t)f * 3 * plane + p;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) o[ch * plane] = __fadd_rn(__fmul_rn(a[ch * plane], k1), __fmul_rn(b[ch * plane], k2));
  }
}

// [nf] x (img1 | img2) [3,H,W] -> [nf][2][4][H][W] with a ones plane appended (test_online_tra.py:144-146)
__global__ void linear_pack4_kernel(const float* __restrict__ hr1, const float* __restrict__ hr2, size_t plane, size_t total,
                                    float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i % plane, q = i / plane;       // q = (frame * 2 + view) * 4 + channel
    const int ch = (int)(q & 3), v = (int)((q >> 2) & 1);
    const size_t f = q >> 3;
    out[i] = ch == 3 ? 1.0f : (v == 0 ? hr1 : hr2)[(f * 3 + ch) * plane + p];
  }
}

static int linear_tables(ss2_ctx* ctx) {
  if (ctx->gauss_ready) return SS2_OK;
  // torchvision _get_gaussian_kernel1d: x = linspace(-10, 10, 21), pdf = exp(-0.5 (x / sigma)^2), / sum (fp32)
  float k[LB_TAPS], sum = 0.f;
  for (int i = 0; i < LB_TAPS; ++i) {
    const float x = (float)(i - LB_HALF) / 20.0f;
    k[i] = expf(-0.5f * x * x);
    sum += k[i];
  }
  for (int i = 0; i < LB_TAPS; ++i) k[i] /= sum;
  SS2_CUDA(ctx, cudaMemcpyToSymbol(c_gauss, k, sizeof(k)));
  ctx->gauss_ready = true;
  return SS2_OK;
}

int linear_blend_launch(ss2_ctx* ctx, const float* d_ref, const float* d_tgt, size_t img_stride, const float* d_ref_m,
                        const float* d_tgt_m, size_t mask_stride, int n, int Ho, int Wo, float* d_out, float* d_mask1,
                        cudaStream_t st) {
  if (n <= 0) return SS2_OK;
  if (Ho <= LB_HALF || Wo <= LB_HALF)
    return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "LINEAR fusion needs a canvas larger than the blur's reflect padding (%d), got %dx%d",
                    LB_HALF, Ho, Wo);
  SS2_TRY(linear_tables(ctx));
  char* buf = nullptr;
  const size_t acc_bytes = ((size_t)n * sizeof(LinearAcc) + 255) / 256 * 256;
  SS2_CUDA(ctx, cudaMallocAsync((void**)&buf, acc_bytes + (size_t)n * Ho * Wo * sizeof(float), st));
  LinearAcc* acc = (LinearAcc*)buf;
  float* tmp = (float*)(buf + acc_bytes);
  linear_init_kernel<<<cdiv(n, 128), 128, 0, st>>>(acc, n);
  SS2_LAUNCH_CHECK(ctx);
  linear_centroid_kernel<<<dim3(Ho, 1, n), 256, 0, st>>>(d_ref_m, d_tgt_m, mask_stride, Wo, acc);
  SS2_LAUNCH_CHECK(ctx);
  linear_proj_minmax_kernel<<<dim3(Ho, 1, n), 256, 0, st>>>(d_ref_m, d_tgt_m, mask_stride, Wo, acc);
  SS2_LAUNCH_CHECK(ctx);
  linear_hblur_kernel<<<dim3(Ho, 1, n), 256, (size_t)Wo * sizeof(float), st>>>(d_ref_m, d_tgt_m, mask_stride, Wo, acc, tmp);
  SS2_LAUNCH_CHECK(ctx);
  linear_vblur_blend_kernel<<<dim3(cdiv(Wo, 32), cdiv(Ho, 8), n), dim3(32, 8), 0, st>>>(d_ref, d_tgt, img_stride, d_ref_m, d_tgt_m,
                                                                                       mask_stride, tmp, Ho, Wo, d_out, d_mask1);
  SS2_LAUNCH_CHECK(ctx);
  cudaFreeAsync(buf, st);
  return SS2_OK;
}

extern "C" int ss2_linear_blend(ss2_ctx* ctx, const float* d_ref, const float* d_tgt, int64_t img_stride, const float* d_ref_m,
                                const float* d_tgt_m, int64_t mask_stride, int n, int Ho, int Wo, float* d_out,
                                float* d_mask1, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || Ho <= 0 || Wo <= 0 || img_stride < 0 || mask_stride < 0 ||
      (n > 0 && (!d_ref_m || !d_tgt_m || (!d_out && !d_mask1) || (d_out && (!d_ref || !d_tgt)))))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_linear_blend: bad arguments");
  if ((size_t)Wo * sizeof(float) > 48 * 1024) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "ss2_linear_blend: canvas wider than 12288");
  return linear_blend_launch(ctx, d_ref, d_tgt, (size_t)img_stride, d_ref_m, d_tgt_m, (size_t)mask_stride, n, Ho, Wo, d_out,
                             d_mask1, (cudaStream_t)stream);
}

// get_stable_sqe with fusion_mode == 'LINEAR' for n frames given the global canvas (same arguments as ss2_stable_frames)
extern "C" int ss2_stable_frames_linear(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1,
                                        const float* d_mesh2, int n, int H, int W, const float* h_minmax, int mode, int tps,
                                        float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || H <= 0 || W <= 0 || !h_minmax || (n > 0 && (!d_hr1 || !d_hr2 || !d_mesh1 || !d_mesh2)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames_linear: bad arguments");
  int Ho, Wo;
  ss2_canvas_size(h_minmax, &Ho, &Wo);
  if (Ho < 0 || Wo < 0) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames_linear: negative canvas");
  if (n == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  if (!d_out) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames_linear: null output");
  if ((size_t)Wo * sizeof(float) > 48 * 1024) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "canvas wider than 12288");
  cudaStream_t st = (cudaStream_t)stream;
  const float out_w = h_minmax[1] - h_minmax[0], out_h = h_minmax[3] - h_minmax[2];
  if (tps != SS2_TPS_LATTICE || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  const int chunk = n < 4 ? n : 4;   // frames per pass: 8 source + 8 canvas planes of temporaries each
  const size_t m = (size_t)n * 2 * SS2_NPT * 2, iplane = (size_t)H * W, oplane = (size_t)Ho * Wo;
  TpsScratch sc;
  SS2_TRY(tps_scratch_alloc(ctx, 2 * chunk, Ho, Wo, tps, 2 * m + (size_t)chunk * 8 * (iplane + oplane) + 128, &sc, st));
  float* source = sc.nodes + (tps == SS2_TPS_LATTICE ? (tps_lattice_workspace_floats(2 * chunk, Ho, Wo) + 63) / 64 * 64 : 0);
  float* target = source + m;
  float* in4 = target + m;
  in4 += (64 - ((size_t)(in4 - sc.base) & 63)) & 63;
  float* w4 = in4 + (size_t)chunk * 8 * iplane;
  w4 += (64 - ((size_t)(w4 - sc.base) & 63)) & 63;
  int rc = stable_meshes_launch(ctx, d_mesh1, d_mesh2, n, H, W, h_minmax[0], h_minmax[2], out_w, out_h, source, target, st);
  for (int k0 = 0; k0 < n && rc == SS2_OK; k0 += chunk) {
    const int nk = n - k0 < chunk ? n - k0 : chunk;
    const size_t total = (size_t)nk * 8 * iplane;
    linear_pack4_kernel<<<(int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, st>>>(
        d_hr1 + (size_t)k0 * 3 * iplane, d_hr2 + (size_t)k0 * 3 * iplane, iplane, total, in4);
    SS2_LAUNCH_CHECK(ctx);
    const float* src = source + (size_t)k0 * 2 * SS2_NPT * 2;
    const float* tgt = target + (size_t)k0 * 2 * SS2_NPT * 2;
    rc = tps_solve_for_warp(ctx, src, tgt, 2 * nk, H, W, Ho, Wo, mode, tps, sc, st);
    if (rc == SS2_OK) rc = tps_warp_launch(ctx, in4, src, sc.T, 2 * nk, 4, H, W, Ho, Wo, mode, tps, w4, st, sc.aux, sc.nodes);
    // per frame: view 1 = planes 0..3 (image, mask), view 2 = planes 4..7
    if (rc == SS2_OK)
      rc = linear_blend_launch(ctx, w4, w4 + 4 * oplane, 8 * oplane, w4 + 3 * oplane, w4 + 7 * oplane, 8 * oplane, nk, Ho, Wo,
                               d_out + (size_t)k0 * 3 * oplane, nullptr, st);
  }
  cudaFreeAsync(sc.base, st);
  return rc;
}

// [nf] frames of ONE view [3,H,W] -> [nf][4][H][W] with a ones plane appended
__global__ void linear_pack4_single_kernel(const float* __restrict__ hr, size_t plane, size_t total, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i % plane, q = i / plane;       // q = frame * 4 + channel
    const int ch = (int)(q & 3);
    out[i] = ch == 3 ? 1.0f : hr[((q >> 2) * 3 + ch) * plane + p];
  }
}

// mask12 = mask1 + mask2 - mask1 * mask2 (test_online_tra_threeview.py:501) on thresholded masks
__global__ void linear_union_mask_kernel(const float* __restrict__ m1, const float* __restrict__ m2, size_t mstride, size_t plane,
                                         size_t total, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t f = i / plane, p = i - f * plane;
    const float a = m1[f * mstride + p] > 0.5f ? 1.f : 0.f, b = m2[f * mstride + p] > 0.5f ? 1.f : 0.f;
    out[i] = a + b - a * b;
  }
}

// three-image warp + LINEAR fusion (test_online_tra_threeview.py:492-503): same arguments as ss2_three_view_frames
extern "C" int ss2_three_view_frames_linear(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_img3,
                                            const float* d_mesh1, const float* d_middle, const float* d_mesh3, int n, int H,
                                            int W, const float* h_canvas, int mode, int tps, float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || H <= 0 || W <= 0 || !h_canvas || (mode != SS2_MODE_NORMAL && mode != SS2_MODE_FAST) ||
      (n > 0 && (!d_img1 || !d_img2 || !d_img3 || !d_mesh1 || !d_middle || !d_mesh3)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_frames_linear: bad arguments");
  const float out_w = h_canvas[2], out_h = h_canvas[3];
  const int Ho = (int)out_h, Wo = (int)out_w;
  if (Ho < 0 || Wo < 0) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_frames_linear: negative canvas");
  if (n == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  if (!d_out) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_frames_linear: null output");
  if ((size_t)Wo * sizeof(float) > 48 * 1024) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "canvas wider than 12288");
  cudaStream_t st = (cudaStream_t)stream;
  if (tps != SS2_TPS_LATTICE || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  const int chunk = n < 2 ? n : 2;
  const size_t m = (size_t)n * SS2_NPT * 2, iplane = (size_t)H * W, oplane = (size_t)Ho * Wo;
  TpsScratch sc;
  // sources / targets (6 m) | one view's packed input (chunk * 4 planes) | three views' warps (3 * chunk * 4 planes) |
  // fused (1,2) (chunk * 3 planes) | mask12 (chunk planes)
  SS2_TRY(tps_scratch_alloc(ctx, chunk, Ho, Wo, tps, 6 * m + (size_t)chunk * (4 * iplane + 16 * oplane) + 256, &sc, st));
  float* source = sc.nodes + (tps == SS2_TPS_LATTICE ? (tps_lattice_workspace_floats(chunk, Ho, Wo) + 63) / 64 * 64 : 0);
  float* target = source + 3 * m;
  auto align = [&](float* p) { return p + ((64 - ((size_t)(p - sc.base) & 63)) & 63); };
  float* in4 = align(target + 3 * m);
  float* w4 = align(in4 + (size_t)chunk * 4 * iplane);
  float* f12 = align(w4 + (size_t)3 * chunk * 4 * oplane);
  float* m12 = align(f12 + (size_t)chunk * 3 * oplane);
  int rc = three_view_sources_launch(ctx, d_mesh1, d_middle, d_mesh3, n, H, W, h_canvas[0], h_canvas[1], out_w, out_h,
                                     source, target, st);
  const float* imgs[3] = {d_img1, d_img2, d_img3};
  for (int k0 = 0; k0 < n && rc == SS2_OK; k0 += chunk) {
    const int nk = n - k0 < chunk ? n - k0 : chunk;
    for (int v = 0; v < 3 && rc == SS2_OK; ++v) {
      const size_t total = (size_t)nk * 4 * iplane;
      linear_pack4_single_kernel<<<(int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, st>>>(
          imgs[v] + (size_t)k0 * 3 * iplane, iplane, total, in4);
      SS2_LAUNCH_CHECK(ctx);
      const float* src = source + ((size_t)v * n + k0) * SS2_NPT * 2;
      const float* tgt = target + ((size_t)v * n + k0) * SS2_NPT * 2;
      rc = tps_solve_for_warp(ctx, src, tgt, nk, H, W, Ho, Wo, mode, tps, sc, st);
      if (rc == SS2_OK)
        rc = tps_warp_launch(ctx, in4, src, sc.T, nk, 4, H, W, Ho, Wo, mode, tps, w4 + (size_t)v * chunk * 4 * oplane, st, sc.aux,
                             sc.nodes);
    }
    const float *w1 = w4, *w2 = w4 + (size_t)chunk * 4 * oplane, *w3 = w4 + (size_t)2 * chunk * 4 * oplane;
    if (rc == SS2_OK)   // img12 = linear_blender(warp1, warp2, mask1, mask2)
      rc = linear_blend_launch(ctx, w1, w2, 4 * oplane, w1 + 3 * oplane, w2 + 3 * oplane, 4 * oplane, nk, Ho, Wo, f12, nullptr, st);
    if (rc == SS2_OK) {
      const size_t total = (size_t)nk * oplane;
      linear_union_mask_kernel<<<(int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, st>>>(
          w1 + 3 * oplane, w2 + 3 * oplane, 4 * oplane, oplane, total, m12);
      SS2_LAUNCH_CHECK(ctx);
      // fusion = linear_blender(img12, warp3, mask12, mask3); image strides differ (3 vs 4 planes): two calls' worth of
      // arguments are expressed through the per-argument base pointers, so blend frame by frame
      for (int k = 0; k < nk && rc == SS2_OK; ++k)
        rc = linear_blend_launch(ctx, f12 + (size_t)k * 3 * oplane, w3 + (size_t)k * 4 * oplane, 0, m12 + (size_t)k * oplane,
                                 w3 + (size_t)k * 4 * oplane + 3 * oplane, 0, 1, Ho, Wo, d_out + (size_t)(k0 + k) * 3 * oplane,
                                 nullptr, st);
    }
  }
  cudaFreeAsync(sc.base, st);
  return rc;
}
