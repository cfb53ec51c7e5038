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

// ------------------------------------------------------------------------------------------
// Three-view glue (test_online_tra_threeview.py:345-455): two stitched pairs (1,2) and (2,3) share their
// middle view; pair (2,3) is shifted onto pair (1,2) by the per-frame mean vertex offset of the shared view, the
// middle plane is the average of the two instances of that view, and the outer meshes are moved through the TPS
// that maps their pair's instance of the shared view onto the middle plane.
// All three kernels are single-CTA (a stream has a few thousand vertices) and keep the reference's fp32
// operation order.  Block-wide min/max helper first.
// ------------------------------------------------------------------------------------------
__device__ void block_minmax4(float& xmin, float& xmax, float& ymin, float& ymax, float (*red)[32]) {
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();
  if (l == 0) { red[0][w] = xmin; red[1][w] = xmax; red[2][w] = ymin; red[3][w] = ymax; }
  __syncthreads();
  const int nw = blockDim.x / 32;
  xmin = red[0][0]; xmax = red[1][0]; ymin = red[2][0]; ymax = red[3][0];
  for (int i = 1; i < nw; ++i) {
    xmin = fminf(xmin, red[0][i]); xmax = fmaxf(xmax, red[1][i]);
    ymin = fminf(ymin, red[2][i]); ymax = fmaxf(ymax, red[3][i]);
  }
}

// work: 5 arrays [n][63][2] (a1, a2, b1, b2, mid in hr pixels); outputs: the five normalised TPS-point operands,
// mid_c (middle mesh in provisional-canvas pixels), canvas1 = (wmin, hmin, out_w, out_h)
__global__ void __launch_bounds__(256)
three_view_align_kernel(const float* __restrict__ w12m1, const float* __restrict__ w12m2,
                        const float* __restrict__ w23m1, const float* __restrict__ w23m2, int n, float img_h,
                        float img_w, float* __restrict__ work, float* __restrict__ pt12, float* __restrict__ src12,
                        float* __restrict__ pt23, float* __restrict__ src23, float* __restrict__ tgt,
                        float* __restrict__ mid_c, float* __restrict__ canvas1) {
  __shared__ float red[4][32];
  const size_t m = (size_t)n * SS2_NPT * 2;
  float *a1 = work, *a2 = work + m, *b1 = work + 2 * m, *b2 = work + 3 * m, *mid = work + 4 * m;
  // rescale to the frame resolution (:347-351)
  for (size_t i = threadIdx.x; i < m; i += blockDim.x) {
    const bool isx = (i & 1) == 0;
    const float s = isx ? img_w : img_h, d = isx ? 480.0f : 360.0f;
    a1[i] = __fdiv_rn(__fmul_rn(w12m1[i], s), d);
    a2[i] = __fdiv_rn(__fmul_rn(w12m2[i], s), d);
    b1[i] = __fdiv_rn(__fmul_rn(w23m1[i], s), d);
    b2[i] = __fdiv_rn(__fmul_rn(w23m2[i], s), d);
  }
  __syncthreads();
  // per-frame mean offset of the shared view (:354-360), shift pair (2,3), middle plane (:363); one warp per frame
  for (int k = threadIdx.x / 32; k < n; k += blockDim.x / 32) {
    const int lane = threadIdx.x % 32;
    float sx = 0.f, sy = 0.f;
    for (int p = lane; p < SS2_NPT; p += 32) {
      const size_t o = ((size_t)k * SS2_NPT + p) * 2;
      sx += a2[o] - b1[o];
      sy += a2[o + 1] - b1[o + 1];
    }
    for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    const float ox = sx / (float)SS2_NPT, oy = sy / (float)SS2_NPT;
    for (int p = lane; p < SS2_NPT; p += 32) {
      const size_t o = ((size_t)k * SS2_NPT + p) * 2;
      b1[o] += ox; b1[o + 1] += oy;
      b2[o] += ox; b2[o + 1] += oy;
      mid[o] = (a2[o] + b1[o]) / 2.0f;
      mid[o + 1] = (a2[o + 1] + b1[o + 1]) / 2.0f;
    }
  }
  __syncthreads();
  // provisional canvas over the four meshes (:366-399)
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (size_t i = threadIdx.x; i < 4 * (m / 2); i += blockDim.x) {
    const float x = work[2 * i], y = work[2 * i + 1];
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x);
    ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
  block_minmax4(xmin, xmax, ymin, ymax, red);
  const float ow = xmax - xmin, oh = ymax - ymin;
  if (threadIdx.x == 0) { canvas1[0] = xmin; canvas1[1] = ymin; canvas1[2] = ow; canvas1[3] = oh; }
  // translate into the canvas and normalise (:406-420)
  for (size_t i = threadIdx.x; i < m; i += blockDim.x) {
    const bool isx = (i & 1) == 0;
    const float mn = isx ? xmin : ymin, ext = isx ? ow : oh;
    pt12[i] = norm1(a1[i] - mn, ext);
    src12[i] = norm1(a2[i] - mn, ext);
    pt23[i] = norm1(b2[i] - mn, ext);
    src23[i] = norm1(b1[i] - mn, ext);
    const float mc = mid[i] - mn;
    mid_c[i] = mc;
    tgt[i] = norm1(mc, ext);
  }
}

// moved12 / moved23: TPS-point outputs (normalised) -> recovered meshes m1_c, m3_c (provisional-canvas pixels,
// :421-424) and the new canvas over (m1_c, mid_c, m3_c) (:436-455): canvas2 = (wmin, hmin, out_w, out_h)
__global__ void __launch_bounds__(256)
three_view_canvas_kernel(const float* __restrict__ moved12, const float* __restrict__ moved23,
                         const float* __restrict__ mid_c, const float* __restrict__ canvas1, int n,
                         float* __restrict__ m1_c, float* __restrict__ m3_c, float* __restrict__ canvas2) {
  __shared__ float red[4][32];
  const size_t m = (size_t)n * SS2_NPT * 2;
  const float ow = canvas1[2], oh = canvas1[3];
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (size_t i = threadIdx.x; i < m / 2; i += blockDim.x) {
    // recover_mesh: (n + 1) * extent / 2
    const float x1 = __fdiv_rn(__fmul_rn(__fadd_rn(moved12[2 * i], 1.0f), ow), 2.0f);
    const float y1 = __fdiv_rn(__fmul_rn(__fadd_rn(moved12[2 * i + 1], 1.0f), oh), 2.0f);
    const float x3 = __fdiv_rn(__fmul_rn(__fadd_rn(moved23[2 * i], 1.0f), ow), 2.0f);
    const float y3 = __fdiv_rn(__fmul_rn(__fadd_rn(moved23[2 * i + 1], 1.0f), oh), 2.0f);
    m1_c[2 * i] = x1; m1_c[2 * i + 1] = y1;
    m3_c[2 * i] = x3; m3_c[2 * i + 1] = y3;
    const float x2 = mid_c[2 * i], y2 = mid_c[2 * i + 1];
    xmin = fminf(xmin, fminf(x1, fminf(x2, x3))); xmax = fmaxf(xmax, fmaxf(x1, fmaxf(x2, x3)));
    ymin = fminf(ymin, fminf(y1, fminf(y2, y3))); ymax = fmaxf(ymax, fmaxf(y1, fmaxf(y2, y3)));
  }
  block_minmax4(xmin, xmax, ymin, ymax, red);
  if (threadIdx.x == 0) { canvas2[0] = xmin; canvas2[1] = ymin; canvas2[2] = xmax - xmin; canvas2[3] = ymax - ymin; }
}

// per view v (grid.y): normalised canvas mesh (source) and normalised rigid mesh (target), [n][63][2] each (:470-486)
__global__ void three_view_sources_kernel(const float* __restrict__ m1_c, const float* __restrict__ mid_c,
                                          const float* __restrict__ m3_c, int n, float img_h, float img_w, float xmin,
                                          float ymin, float out_w, float out_h, float* __restrict__ source,
                                          float* __restrict__ target) {
  const int k = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
  if (tid >= SS2_NPT) return;
  const float* p = (v == 0 ? m1_c : v == 1 ? mid_c : m3_c) + ((size_t)k * SS2_NPT + tid) * 2;
  const size_t o = (((size_t)v * n + k) * SS2_NPT + tid) * 2;
  source[o] = norm1(__fsub_rn(p[0], xmin), out_w);
  source[o + 1] = norm1(__fsub_rn(p[1], ymin), out_h);
  const int gi = tid / (SS2_GRID_W + 1), gj = tid % (SS2_GRID_W + 1);
  target[o] = norm1(lin0(gj, SS2_GRID_W + 1, img_w), img_w);
  target[o + 1] = norm1(lin0(gi, SS2_GRID_H + 1, img_h), img_h);
}

// AVERAGE fusion of three warped views: fuse(1,2) then fuse(12,3) (:489-490), reference operation order
__global__ void blend3_avg_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
                                  size_t count, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const float a = w1[i], b = w2[i], c = w3[i];
    const float s = __fadd_rn(__fadd_rn(a, b), 1e-6f);
    const float f12 = __fadd_rn(__fmul_rn(a, __fdiv_rn(a, s)), __fmul_rn(b, __fdiv_rn(b, s)));
    const float t = __fadd_rn(__fadd_rn(f12, c), 1e-6f);
    __stcs(out + i, __fadd_rn(__fmul_rn(f12, __fdiv_rn(f12, t)), __fmul_rn(c, __fdiv_rn(c, t))));
  }
}

int three_view_align_launch(ss2_ctx* ctx, const float* w12m1, const float* w12m2, const float* w23m1, const float* w23m2,
                            int n, int img_h, int img_w, float* work, float* pt12, float* src12, float* pt23, float* src23,
                            float* tgt, float* mid_c, float* canvas1, cudaStream_t st) {
  three_view_align_kernel<<<1, 256, 0, st>>>(w12m1, w12m2, w23m1, w23m2, n, (float)img_h, (float)img_w, work, pt12, src12,
                                             pt23, src23, tgt, mid_c, canvas1);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
int three_view_canvas_launch(ss2_ctx* ctx, const float* moved12, const float* moved23, const float* mid_c,
                             const float* canvas1, int n, float* m1_c, float* m3_c, float* canvas2, cudaStream_t st) {
  three_view_canvas_kernel<<<1, 256, 0, st>>>(moved12, moved23, mid_c, canvas1, n, m1_c, m3_c, canvas2);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
int three_view_sources_launch(ss2_ctx* ctx, const float* m1_c, const float* mid_c, const float* m3_c, int n, int img_h,
                              int img_w, float xmin, float ymin, float out_w, float out_h, float* source, float* target,
                              cudaStream_t st) {
  three_view_sources_kernel<<<dim3(n, 3), 64, 0, st>>>(m1_c, mid_c, m3_c, n, (float)img_h, (float)img_w, xmin, ymin, out_w,
                                                      out_h, source, target);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
int blend3_avg_launch(ss2_ctx* ctx, const float* w1, const float* w2, const float* w3, size_t count, float* out,
                      cudaStream_t st) {
  if (count == 0) return SS2_OK;
  const int blocks = (int)((count + 255) / 256 < 148 * 16 ? (count + 255) / 256 : 148 * 16);
  blend3_avg_kernel<<<blocks, 256, 0, st>>>(w1, w2, w3, count, out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// N-view middle-plane chain (BASELINE.json config 5; generalises the three-view glue above and reduces to it for
// N = 3, tests/test_gpu_parity.py::test_nview_equals_three_view).  Pairs (1,2), (2,3), .., (N-1,N): meshB of pair p and
// meshA of pair p+1 are two instances of one physical view.
//   align:  rescale to hr; every pair p >= 2 shifted by the per-frame mean vertex offset that brings its meshA onto
//           the (already shifted) meshB of pair p-1; middle plane of every shared view; min/max of all shifted meshes
//   remap:  translate / normalise by the (global) provisional canvas; the two OUTER views follow their pair's instance
//           of the neighbouring shared view through the TPS onto its middle plane; min/max of the N final meshes
// The two min/max results leave the kernels so that a temporally sharded stream can all-reduce them (the TPS
// normalisation depends on the provisional canvas of ALL frames).  Single CTA each, reference fp32 operation order.
// ------------------------------------------------------------------------------------------
#define NVIEW_MAX 8
struct NviewPtrs { const float* p[2 * (NVIEW_MAX - 1)]; };

__global__ void __launch_bounds__(256)
nview_align_kernel(NviewPtrs in, int nviews, int n, float img_h, float img_w, float* __restrict__ shifted,
                   float* __restrict__ mids, float* __restrict__ minmax) {
  __shared__ float red[4][32];
  const size_t m = (size_t)n * SS2_NPT * 2;
  const int npair = nviews - 1;
  for (int q = 0; q < 2 * npair; ++q)
    for (size_t i = threadIdx.x; i < m; i += blockDim.x) {
      const bool isx = (i & 1) == 0;
      shifted[q * m + i] = __fdiv_rn(__fmul_rn(in.p[q][i], isx ? img_w : img_h), isx ? 480.0f : 360.0f);
    }
  __syncthreads();
  // chain: pair p follows pair p-1 (sequential in p, one warp per frame)
  for (int p = 1; p < npair; ++p) {
    const float* prevB = shifted + (size_t)(2 * (p - 1) + 1) * m;
    float* A = shifted + (size_t)(2 * p) * m;
    float* B = A + m;
    float* mid = mids + (size_t)(p - 1) * m;
    for (int k = threadIdx.x / 32; k < n; k += blockDim.x / 32) {
      const int lane = threadIdx.x % 32;
      float sx = 0.f, sy = 0.f;
      for (int v = lane; v < SS2_NPT; v += 32) {
        const size_t o = ((size_t)k * SS2_NPT + v) * 2;
        sx += prevB[o] - A[o];
        sy += prevB[o + 1] - A[o + 1];
      }
      for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
      const float ox = sx / (float)SS2_NPT, oy = sy / (float)SS2_NPT;
      for (int v = lane; v < SS2_NPT; v += 32) {
        const size_t o = ((size_t)k * SS2_NPT + v) * 2;
        A[o] += ox; A[o + 1] += oy;
        B[o] += ox; B[o + 1] += oy;
        mid[o] = (prevB[o] + A[o]) / 2.0f;
        mid[o + 1] = (prevB[o + 1] + A[o + 1]) / 2.0f;
      }
    }
    __syncthreads();
  }
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (size_t i = threadIdx.x; i < (size_t)npair * m; i += blockDim.x) {   // i over (x, y) pairs of all 2*npair meshes
    const float x = shifted[2 * i], y = shifted[2 * i + 1];
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x);
    ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
  block_minmax4(xmin, xmax, ymin, ymax, red);
  if (threadIdx.x == 0) { minmax[0] = xmin; minmax[1] = xmax; minmax[2] = ymin; minmax[3] = ymax; }
}

// TPS-point operands of the two outer views (normalised by the provisional canvas) and the shared views' final meshes
// (provisional-canvas pixels) written straight into meshes_out[1..N-2]
__global__ void __launch_bounds__(256)
nview_operands_kernel(const float* __restrict__ shifted, const float* __restrict__ mids, int nviews, int n, float xmin,
                      float ymin, float ow, float oh, float* __restrict__ pt0, float* __restrict__ src0,
                      float* __restrict__ tgt0, float* __restrict__ pt1, float* __restrict__ src1, float* __restrict__ tgt1,
                      float* __restrict__ meshes_out) {
  const size_t m = (size_t)n * SS2_NPT * 2;
  const int npair = nviews - 1;
  const float *A0 = shifted, *B0 = shifted + m, *AL = shifted + (size_t)(2 * (npair - 1)) * m, *BL = AL + m;
  const float *mid0 = mids, *midL = mids + (size_t)(npair - 2) * m;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
    const bool isx = (i & 1) == 0;
    const float mn = isx ? xmin : ymin, ext = isx ? ow : oh;
    pt0[i] = norm1(A0[i] - mn, ext);
    src0[i] = norm1(B0[i] - mn, ext);
    tgt0[i] = norm1(mid0[i] - mn, ext);
    pt1[i] = norm1(BL[i] - mn, ext);
    src1[i] = norm1(AL[i] - mn, ext);
    tgt1[i] = norm1(midL[i] - mn, ext);
    for (int k = 0; k < npair - 1; ++k) meshes_out[(size_t)(k + 1) * m + i] = mids[(size_t)k * m + i] - mn;
  }
}

// recover the outer views' moved meshes into meshes_out[0], [N-1]; min/max over all N final meshes
__global__ void __launch_bounds__(256)
nview_finish_kernel(const float* __restrict__ moved0, const float* __restrict__ moved1, int nviews, int n, float ow, float oh,
                    float* __restrict__ meshes_out, float* __restrict__ minmax) {
  __shared__ float red[4][32];
  const size_t m = (size_t)n * SS2_NPT * 2;
  float* first = meshes_out;
  float* last = meshes_out + (size_t)(nviews - 1) * m;
  for (size_t i = threadIdx.x; i < m; i += blockDim.x) {
    const float ext = (i & 1) == 0 ? ow : oh;
    first[i] = __fdiv_rn(__fmul_rn(__fadd_rn(moved0[i], 1.0f), ext), 2.0f);   // recover_mesh: (n + 1) * extent / 2
    last[i] = __fdiv_rn(__fmul_rn(__fadd_rn(moved1[i], 1.0f), ext), 2.0f);
  }
  __syncthreads();
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (size_t i = threadIdx.x; i < (size_t)nviews * (m / 2); i += blockDim.x) {
    const float x = meshes_out[2 * i], y = meshes_out[2 * i + 1];
    xmin = fminf(xmin, x); xmax = fmaxf(xmax, x);
    ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
  }
  block_minmax4(xmin, xmax, ymin, ymax, red);
  if (threadIdx.x == 0) { minmax[0] = xmin; minmax[1] = xmax; minmax[2] = ymin; minmax[3] = ymax; }
}

// per view v (grid.y): normalised canvas mesh (source) and normalised rigid mesh (target), laid out [n][V][63][2] as the
// fused N-view resampler wants them
__global__ void nview_sources_kernel(const float* __restrict__ meshes, int nviews, int n, float img_h, float img_w, float xmin,
                                     float ymin, float out_w, float out_h, float* __restrict__ source,
                                     float* __restrict__ target) {
  const int k = blockIdx.x, v = blockIdx.y, tid = threadIdx.x;
  if (tid >= SS2_NPT) return;
  const float* p = meshes + (((size_t)v * n + k) * SS2_NPT + tid) * 2;
  const size_t o = (((size_t)k * nviews + v) * SS2_NPT + tid) * 2;
  source[o] = norm1(__fsub_rn(p[0], xmin), out_w);
  source[o + 1] = norm1(__fsub_rn(p[1], ymin), out_h);
  const int gi = tid / (SS2_GRID_W + 1), gj = tid % (SS2_GRID_W + 1);
  target[o] = norm1(lin0(gj, SS2_GRID_W + 1, img_w), img_w);
  target[o + 1] = norm1(lin0(gi, SS2_GRID_H + 1, img_h), img_h);
}

int nview_align_launch(ss2_ctx* ctx, const float* const* d_pairs, int nviews, int n, int img_h, int img_w, float* shifted,
                       float* mids, float* minmax, cudaStream_t st) {
  NviewPtrs P;
  for (int q = 0; q < 2 * (nviews - 1); ++q) P.p[q] = d_pairs[q];
  nview_align_kernel<<<1, 256, 0, st>>>(P, nviews, n, (float)img_h, (float)img_w, shifted, mids, minmax);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
int nview_operands_launch(ss2_ctx* ctx, const float* shifted, const float* mids, int nviews, int n, float xmin, float ymin,
                          float ow, float oh, float* pt0, float* src0, float* tgt0, float* pt1, float* src1, float* tgt1,
                          float* meshes_out, cudaStream_t st) {
  const size_t m = (size_t)n * SS2_NPT * 2;
  nview_operands_kernel<<<(int)((m + 255) / 256 < 64 ? (m + 255) / 256 : 64), 256, 0, st>>>(shifted, mids, nviews, n, xmin, ymin, ow, oh,
                                                                                        pt0, src0, tgt0, pt1, src1, tgt1, meshes_out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
int nview_finish_launch(ss2_ctx* ctx, const float* moved0, const float* moved1, int nviews, int n, float ow, float oh,
                        float* meshes_out, float* minmax, cudaStream_t st) {
  nview_finish_kernel<<<1, 256, 0, st>>>(moved0, moved1, nviews, n, ow, oh, meshes_out, minmax);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
int nview_sources_launch(ss2_ctx* ctx, const float* meshes, int nviews, int n, int img_h, int img_w, float xmin, float ymin,
                         float out_w, float out_h, float* source, float* target, cudaStream_t st) {
  nview_sources_kernel<<<dim3(n, nviews), 64, 0, st>>>(meshes, nviews, n, (float)img_h, (float)img_w, xmin, ymin, out_w, out_h,
                                                       source, target);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
