// Mesh algebra: 4-point DLT, the bidirectional (middle-plane) homography split, H2Mesh, the
// homography feature warp, tsmotion glue and the canvas reductions.  All of these are tiny,
// latency-only kernels; one thread (or one warp) per batch element, batched over frames.
//
// Reference behaviour restated (paths under Full_model_inference/Codes/):
//   utils/torch_DLT.py:17-45                  tensor_DLT
//   spatial_network.py:20-36                  H2Mesh (applies H^-1)
//   spatial_network.py:73-93, 291-313         homography split at full / 1/8 scale
//   utils/torch_homo_transform.py:128-180     homography backward warp
//   test_online_tra.py:61-91, 309-347         norm/recover mesh, tsmotion preparation
//   test_online_tra.py:103-136                canvas min/max, per-frame mesh normalisation
#include "common.cuh"

// The reference solves these small systems with fp32 LU (torch.inverse).  We solve them in
// fp64 and round once: the difference to the reference is then the reference's own fp32
// round-off only, instead of the sum of two independent fp32 error terms.
__device__ void solve8(double A[8][9]) {
  for (int k = 0; k < 8; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
    for (int r = k + 1; r < 8; ++r)
      if (fabs(A[r][k]) > best) { best = fabs(A[r][k]); p = r; }
    if (p != k)
      for (int c = 0; c < 9; ++c) { double t = A[k][c]; A[k][c] = A[p][c]; A[p][c] = t; }
    double inv = 1.0 / A[k][k];
    for (int r = 0; r < 8; ++r) {
      if (r == k) continue;
      double f = A[r][k] * inv;
      for (int c = k; c < 9; ++c) A[r][c] -= f * A[k][c];
    }
  }
  for (int k = 0; k < 8; ++k) A[k][8] /= A[k][k];
}

// H (row-major 3x3, fp32-rounded like the reference's output) from 4 correspondences
__device__ void dlt4(const float* src, const float* dst, float* H) {
  double A[8][9];
  for (int p = 0; p < 4; ++p) {
    const float x = src[2 * p], y = src[2 * p + 1], u = dst[2 * p], v = dst[2 * p + 1];
    double* r0 = A[2 * p];
    double* r1 = A[2 * p + 1];
    r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0;
    r0[6] = -(double)__fmul_rn(u, x); r0[7] = -(double)__fmul_rn(u, y); r0[8] = u;
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1;
    r1[6] = -(double)__fmul_rn(v, x); r1[7] = -(double)__fmul_rn(v, y); r1[8] = v;
  }
  solve8(A);
  for (int i = 0; i < 8; ++i) H[i] = (float)A[i][8];
  H[8] = 1.0f;
}

__device__ void inv3(const float* M, float* out) {
  const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double id = 1.0 / det;
  out[0] = (float)(A * id); out[1] = (float)(-(b * i - c * h) * id); out[2] = (float)((b * f - c * e) * id);
  out[3] = (float)(B * id); out[4] = (float)((a * i - c * g) * id); out[5] = (float)(-(a * f - c * d) * id);
  out[6] = (float)(C * id); out[7] = (float)(-(a * h - b * g) * id); out[8] = (float)((a * e - b * d) * id);
}

__device__ void mm3(const float* A, const float* B, float* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      C[r * 3 + c] = fmaf(A[r * 3 + 2], B[6 + c], fmaf(A[r * 3 + 1], B[3 + c], A[r * 3] * B[c]));
}

__global__ void dlt_kernel(const float* __restrict__ src, const float* __restrict__ dst, int bs,
                           float* __restrict__ H) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  dlt4(src + b * 8, dst + b * 8, H + b * 9);
}

int dlt_launch(ss2_ctx* ctx, const float* d_src, const float* d_dst, int bs, float* d_H, cudaStream_t st) {
  if (bs <= 0) return SS2_OK;
  dlt_kernel<<<cdiv(bs, 32), 32, 0, st>>>(d_src, d_dst, bs, d_H);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// H, H_tgt, H_ref = H^-1 H_tgt from the 8 regressed corner offsets, at coordinates / scale
__device__ void homo_split(const float* off, float img_h, float img_w, float scale, float* Href, float* Htgt) {
  float src[8] = {0.f, 0.f, img_w, 0.f, 0.f, img_h, img_w, img_h};
  float dst[8], dstt[8], s8[8];
  for (int i = 0; i < 8; ++i) {
    dst[i] = __fadd_rn(src[i], off[i]);
    dstt[i] = __fadd_rn(src[i], off[i] / 2.0f);
  }
  if (scale != 1.0f)
    for (int i = 0; i < 8; ++i) { s8[i] = src[i] / scale; dst[i] = dst[i] / scale; dstt[i] = dstt[i] / scale; }
  else
    for (int i = 0; i < 8; ++i) s8[i] = src[i];
  float Hm[9], Hi[9];
  dlt4(s8, dst, Hm);
  dlt4(s8, dstt, Htgt);
  inv3(Hm, Hi);
  mm3(Hi, Htgt, Href);
}

// spatial_network.py:291-313: normalised homographies that warp the 1/8-scale feature maps
__global__ void spatial_split_kernel(const float* __restrict__ offset1, int bs, float img_h, float img_w,
                                     float* __restrict__ theta_ref, float* __restrict__ theta_tgt) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  float Href[9], Htgt[9];
  homo_split(offset1 + b * 8, img_h, img_w, 8.0f, Href, Htgt);
  const float w8 = img_w / 8.0f, h8 = img_h / 8.0f;
  const float M[9] = {w8 / 2.0f, 0.f, w8 / 2.0f, 0.f, h8 / 2.0f, h8 / 2.0f, 0.f, 0.f, 1.f};
  float Mi[9], t[9];
  inv3(M, Mi);
  mm3(Mi, Href, t);
  mm3(t, M, theta_ref + b * 9);
  mm3(Mi, Htgt, t);
  mm3(t, M, theta_tgt + b * 9);
}

int spatial_split_launch(ss2_ctx* ctx, const float* d_offset1, int bs, int img_h, int img_w,
                         float* d_theta_ref, float* d_theta_tgt, cudaStream_t st) {
  if (bs <= 0) return SS2_OK;
  spatial_split_kernel<<<cdiv(bs, 32), 32, 0, st>>>(d_offset1, bs, (float)img_h, (float)img_w, d_theta_ref,
                                                    d_theta_tgt);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// torch.linspace(0, end, n)[i]
__device__ __forceinline__ float lin0(int i, int n, float end) {
  const float step = end / (float)(n - 1);
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(end, __fmul_rn(step, (float)(n - 1 - i)));
}

// build_SpatialNet tail, spatial_network.py:68-115: one block of 64 threads per pair
__global__ void spatial_tail_kernel(const float* __restrict__ o1, const float* __restrict__ oref,
                                    const float* __restrict__ otgt, float img_h, float img_w,
                                    float* __restrict__ m1, float* __restrict__ m2) {
  __shared__ float Hi[2][9];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    float Href[9], Htgt[9];
    homo_split(o1 + b * 8, img_h, img_w, 1.0f, Href, Htgt);
    inv3(Href, Hi[0]);
    inv3(Htgt, Hi[1]);
  }
  __syncthreads();
  if (tid >= SS2_NPT) return;
  const int gi = tid / (SS2_GRID_W + 1), gj = tid % (SS2_GRID_W + 1);
  const float x = lin0(gj, SS2_GRID_W + 1, img_w), y = lin0(gi, SS2_GRID_H + 1, img_h);
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const float* h = Hi[v];
    const float qx = fmaf(h[1], y, h[0] * x) + h[2];
    const float qy = fmaf(h[4], y, h[3] * x) + h[5];
    const float qz = fmaf(h[7], y, h[6] * x) + h[8];
    const float* mo = (v == 0 ? oref : otgt) + (size_t)b * 126 + tid * 2;
    float* out = (v == 0 ? m1 : m2) + (size_t)b * 126 + tid * 2;
    // (ini_mesh + motion) - rigid
    out[0] = __fsub_rn(__fadd_rn(__fdiv_rn(qx, qz), mo[0]), x);
    out[1] = __fsub_rn(__fadd_rn(__fdiv_rn(qy, qz), mo[1]), y);
  }
}

int spatial_tail_launch(ss2_ctx* ctx, const float* d_o1, const float* d_oref, const float* d_otgt, int bs,
                        int img_h, int img_w, float* d_m1, float* d_m2, cudaStream_t st) {
  if (bs <= 0) return SS2_OK;
  spatial_tail_kernel<<<bs, 64, 0, st>>>(d_o1, d_oref, d_otgt, (float)img_h, (float)img_w, d_m1, d_m2);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// homography backward warp.  Coordinates follow torch_homo_transform.py:147-175, sampling is
// the shared _interpolate (clamped taps, weights from clamped coordinates, unfused).
// ------------------------------------------------------------------------------------------
struct Taps {
  int ia, ib, ic, id;
  float wa, wb, wc, wd;
};

__device__ __forceinline__ float lin11g(int i, int n) {
  const float step = n > 1 ? 2.0f / (float)(n - 1) : 0.0f;
  return (i < n / 2) ? __fadd_rn(-1.0f, __fmul_rn(step, (float)i))
                     : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ Taps homo_taps(const float* th, int row, int col, int Ho, int Wo, int H, int W) {
  const float xt = lin11g(col, Wo), yt = lin11g(row, Ho);
  float xs = fmaf(th[1], yt, th[0] * xt) + th[2];
  float ys = fmaf(th[4], yt, th[3] * xt) + th[5];
  float ts = fmaf(th[7], yt, th[6] * xt) + th[8];
  ts = __fadd_rn(ts, (fabsf(ts) >= 1e-7f) ? 0.0f : 1e-6f);
  xs = __fdiv_rn(xs, ts);
  ys = __fdiv_rn(ys, ts);
  const float x = __fmul_rn(__fmul_rn(__fadd_rn(xs, 1.0f), (float)W), 0.5f);
  const float y = __fmul_rn(__fmul_rn(__fadd_rn(ys, 1.0f), (float)H), 0.5f);
  int x0 = (int)fminf(fmaxf(floorf(x), -2.0e9f), 2.0e9f);
  int y0 = (int)fminf(fmaxf(floorf(y), -2.0e9f), 2.0e9f);
  int x1 = min(max(x0 + 1, 0), W - 1), y1 = min(max(y0 + 1, 0), H - 1);
  x0 = min(max(x0, 0), W - 1);
  y0 = min(max(y0, 0), H - 1);
  Taps t;
  t.ia = y0 * W + x0; t.ib = y1 * W + x0; t.ic = y0 * W + x1; t.id = y1 * W + x1;
  t.wa = __fmul_rn(__fsub_rn((float)x1, x), __fsub_rn((float)y1, y));
  t.wb = __fmul_rn(__fsub_rn((float)x1, x), __fsub_rn(y, (float)y0));
  t.wc = __fmul_rn(__fsub_rn(x, (float)x0), __fsub_rn((float)y1, y));
  t.wd = __fmul_rn(__fsub_rn(x, (float)x0), __fsub_rn(y, (float)y0));
  return t;
}

__device__ __forceinline__ float tap_sum(const Taps& t, float Ia, float Ib, float Ic, float Id) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.wa, Ia), __fmul_rn(t.wb, Ib)), __fmul_rn(t.wc, Ic)),
                   __fmul_rn(t.wd, Id));
}

// NCHW in / NCHW out: the generic utils.torch_homo_transform.transformer entry point
__global__ void homo_warp_nchw_kernel(const float* __restrict__ U, const float* __restrict__ theta, int C, int H,
                                      int W, int Ho, int Wo, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y * blockDim.y + threadIdx.y;
  if (col >= Wo || row >= Ho) return;
  const Taps t = homo_taps(theta + b * 9, row, col, Ho, Wo, H, W);
  const size_t plane = (size_t)H * W, oplane = (size_t)Ho * Wo;
  for (int c = 0; c < C; ++c) {
    const float* p = U + ((size_t)b * C + c) * plane;
    out[((size_t)b * C + c) * oplane + (size_t)row * Wo + col] =
        tap_sum(t, __ldg(p + t.ia), __ldg(p + t.ib), __ldg(p + t.ic), __ldg(p + t.id));
  }
}

int homo_warp_nchw_launch(ss2_ctx* ctx, const float* d_U, const float* d_theta, int bn, int C, int H, int W,
                          int Ho, int Wo, float* d_out, cudaStream_t st) {
  if (bn <= 0) return SS2_OK;
  dim3 block(32, 4), grid(cdiv(Wo, 32), cdiv(Ho, 4), bn);
  homo_warp_nchw_kernel<<<grid, block, 0, st>>>(d_U, d_theta, C, H, W, Ho, Wo, d_out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// NHWC in / NHWC out, same size (the feature-map warp inside SpatialNet): one warp per output
// pixel, lanes stride the channels as float4 -> fully coalesced 512 B rows for C = 128.
__global__ void homo_warp_nhwc_kernel(const float* __restrict__ U, const float* __restrict__ theta, int C, int H,
                                      int W, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (pix >= H * W) return;
  const Taps t = homo_taps(theta + b * 9, pix / W, pix % W, H, W, H, W);
  const float4* base = reinterpret_cast<const float4*>(U + (size_t)b * H * W * C);
  float4* o = reinterpret_cast<float4*>(out + ((size_t)b * H * W + pix) * C);
  const int c4 = C / 4;
  for (int c = lane; c < c4; c += 32) {
    const float4 a = __ldg(base + (size_t)t.ia * c4 + c), bb = __ldg(base + (size_t)t.ib * c4 + c);
    const float4 cc = __ldg(base + (size_t)t.ic * c4 + c), d = __ldg(base + (size_t)t.id * c4 + c);
    float4 r;
    r.x = tap_sum(t, a.x, bb.x, cc.x, d.x);
    r.y = tap_sum(t, a.y, bb.y, cc.y, d.y);
    r.z = tap_sum(t, a.z, bb.z, cc.z, d.z);
    r.w = tap_sum(t, a.w, bb.w, cc.w, d.w);
    o[c] = r;
  }
}

int homo_warp_nhwc_launch(ss2_ctx* ctx, const float* d_U, const float* d_theta, int bn, int C, int H, int W,
                          float* d_out, cudaStream_t st) {
  if (bn <= 0) return SS2_OK;
  dim3 grid(cdiv(H * W, 8), bn);
  homo_warp_nhwc_kernel<<<grid, 256, 0, st>>>(d_U, d_theta, C, H, W, d_out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// tsmotion preparation (test_online_tra.py:309-347), one view, frames batched.
// prep: smesh_k = rigid + smotion_k; and for k with a predecessor the three normalised point
// sets of the TPS-point call: point = norm(rigid + tmotion_k), source = norm(rigid),
// target = norm(rigid + smotion_{k-1}).  finish: tsmotion_k = recover(moved) - smesh_k.
// ------------------------------------------------------------------------------------------
#define NET_H 360.0f
#define NET_W 480.0f

__device__ __forceinline__ float norm1(float v, float extent) {  // v*2/extent - 1
  return __fsub_rn(__fdiv_rn(__fmul_rn(v, 2.0f), extent), 1.0f);
}

__global__ void tsmotion_prep_kernel(const float* __restrict__ smotion, const float* __restrict__ tmotion, int n,
                                     int first, const float* __restrict__ prev, float* __restrict__ smesh,
                                     float* __restrict__ point, float* __restrict__ source,
                                     float* __restrict__ target) {
  const int k = blockIdx.x, tid = threadIdx.x;
  if (tid >= SS2_NPT) return;
  const int gi = tid / (SS2_GRID_W + 1), gj = tid % (SS2_GRID_W + 1);
  const float rx = lin0(gj, SS2_GRID_W + 1, NET_W), ry = lin0(gi, SS2_GRID_H + 1, NET_H);
  const size_t o = ((size_t)k * SS2_NPT + tid) * 2;
  smesh[o] = __fadd_rn(rx, smotion[o]);
  smesh[o + 1] = __fadd_rn(ry, smotion[o + 1]);
  const float* pm = (k > 0) ? smotion + o - SS2_NPT * 2 : (first ? nullptr : prev + tid * 2);
  float px = rx, py = ry, tx = rx, ty = ry;
  if (pm) {
    px = __fadd_rn(rx, pm[0]); py = __fadd_rn(ry, pm[1]);
    tx = __fadd_rn(rx, tmotion[o]); ty = __fadd_rn(ry, tmotion[o + 1]);
  }
  point[o] = norm1(tx, NET_W); point[o + 1] = norm1(ty, NET_H);
  source[o] = norm1(rx, NET_W); source[o + 1] = norm1(ry, NET_H);
  target[o] = norm1(px, NET_W); target[o + 1] = norm1(py, NET_H);
}

__global__ void tsmotion_finish_kernel(const float* __restrict__ moved, const float* __restrict__ smesh, int n,
                                       int first, float* __restrict__ tsmotion) {
  const int k = blockIdx.x, tid = threadIdx.x;
  if (tid >= SS2_NPT) return;
  const size_t o = ((size_t)k * SS2_NPT + tid) * 2;
  if (k == 0 && first) {
    tsmotion[o] = 0.0f;
    tsmotion[o + 1] = 0.0f;
    return;
  }
  // recover_mesh: (n+1)*extent/2
  const float x = __fdiv_rn(__fmul_rn(__fadd_rn(moved[o], 1.0f), NET_W), 2.0f);
  const float y = __fdiv_rn(__fmul_rn(__fadd_rn(moved[o + 1], 1.0f), NET_H), 2.0f);
  tsmotion[o] = __fsub_rn(x, smesh[o]);
  tsmotion[o + 1] = __fsub_rn(y, smesh[o + 1]);
}

int tsmotion_prep_launch(ss2_ctx* ctx, const float* d_smotion, const float* d_tmotion, int n, int first,
                         const float* d_prev, float* d_smesh, float* d_point, float* d_source,
                         float* d_target, cudaStream_t st) {
  if (n <= 0) return SS2_OK;
  tsmotion_prep_kernel<<<n, 64, 0, st>>>(d_smotion, d_tmotion, n, first, d_prev, d_smesh, d_point, d_source,
                                         d_target);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

int tsmotion_finish_launch(ss2_ctx* ctx, const float* d_moved, const float* d_smesh, int n, int first,
                           float* d_tsmotion, cudaStream_t st) {
  if (n <= 0) return SS2_OK;
  tsmotion_finish_kernel<<<n, 64, 0, st>>>(d_moved, d_smesh, n, first, d_tsmotion);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// canvas: min/max of the hr-rescaled smooth meshes of both views over all frames
// (test_online_tra.py:103-117).  Single CTA; meshes are KB-sized.
// ------------------------------------------------------------------------------------------
__global__ void canvas_minmax_kernel(const float* __restrict__ m1, const float* __restrict__ m2, int npts,
                                     float img_h, float img_w, float* __restrict__ out) {
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int i = threadIdx.x; i < 2 * npts; i += blockDim.x) {
    const float* p = (i < npts) ? m1 + (size_t)i * 2 : m2 + (size_t)(i - npts) * 2;
    const float x = __fdiv_rn(__fmul_rn(p[0], img_w), 480.0f);
    const float y = __fdiv_rn(__fmul_rn(p[1], img_h), 360.0f);
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x);
    ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
  __shared__ float red[4][32];
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  if (l == 0) { red[0][w] = xmin; red[1][w] = xmax; red[2][w] = ymin; red[3][w] = ymax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x / 32;
    for (int i = 1; i < nw; ++i) {
      xmin = fminf(xmin, red[0][i]); xmax = fmaxf(xmax, red[1][i]);
      ymin = fminf(ymin, red[2][i]); ymax = fmaxf(ymax, red[3][i]);
    }
    out[0] = xmin; out[1] = xmax; out[2] = ymin; out[3] = ymax;
  }
}

int canvas_minmax_launch(ss2_ctx* ctx, const float* d_mesh1, const float* d_mesh2, int n, int img_h,
                         int img_w, float* d_minmax, cudaStream_t st) {
  canvas_minmax_kernel<<<1, 256, 0, st>>>(d_mesh1, d_mesh2, n * SS2_NPT, (float)img_h, (float)img_w, d_minmax);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// per-frame normalised canvas meshes (source) and the normalised rigid mesh (target) for both
// views, laid out [n][2][63][2] as the fused resampler wants them (test_online_tra.py:100-136)
__global__ void stable_meshes_kernel(const float* __restrict__ m1, const float* __restrict__ m2, float img_h,
                                     float img_w, float xmin, float ymin, float out_w, float out_h,
                                     float* __restrict__ source, float* __restrict__ target) {
  const int k = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
  if (tid >= SS2_NPT) return;
  const float* p = (v == 0 ? m1 : m2) + ((size_t)k * SS2_NPT + tid) * 2;
  const float x = __fdiv_rn(__fmul_rn(p[0], img_w), 480.0f);
  const float y = __fdiv_rn(__fmul_rn(p[1], img_h), 360.0f);
  const size_t o = (((size_t)k * 2 + v) * SS2_NPT + tid) * 2;
  source[o] = norm1(__fsub_rn(x, xmin), out_w);
  source[o + 1] = norm1(__fsub_rn(y, ymin), out_h);
  const int gi = tid / (SS2_GRID_W + 1), gj = tid % (SS2_GRID_W + 1);
  target[o] = norm1(lin0(gj, SS2_GRID_W + 1, img_w), img_w);
  target[o + 1] = norm1(lin0(gi, SS2_GRID_H + 1, img_h), img_h);
}

int stable_meshes_launch(ss2_ctx* ctx, const float* d_mesh1, const float* d_mesh2, int n, int img_h,
                         int img_w, float xmin, float ymin, float out_w, float out_h, float* d_source,
                         float* d_target, cudaStream_t st) {
  if (n <= 0) return SS2_OK;
  stable_meshes_kernel<<<dim3(n, 2), 64, 0, st>>>(d_mesh1, d_mesh2, (float)img_h, (float)img_w, xmin, ymin,
                                                  out_w, out_h, d_source, d_target);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
