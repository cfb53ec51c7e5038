// Convolution family as implicit GEMM over NHWC / NDHWC activations (fp32 SIMT path).
//
//   out[m][n] = act( sum_k A[m][k] * Wp[k][n] + bias[n] + residual[m][n] )
//   m = (b, do, ho, wo) output position, k = (kd, kh, kw, cin), n = cout
//
// One kernel covers every dense layer of the three networks (reference paths under
// Full_model_inference/Codes/): the ResNet-18 stem / BasicBlocks (spatial_network.py:123-139,
// eval-mode BN folded into Wp/bias at pack time), the regressor 3x3 stacks and their Linear
// heads (spatial_network.py:147-259, temporal_network.py:65-104; a Linear is a 1x1 conv on a
// 1x1 map), the CCL correlation (per-sample filters, `groups`) and SmoothNet's Conv3d
// (smooth_network.py:124-131).  This is the exact-fp32 path; conv_tc.cu holds the tcgen05
// tensor-core version used for the large layers.
#include <algorithm>

#include "common.cuh"

void conv_out_dims(const ConvLayer& L, int D, int H, int W, int* Do, int* Ho, int* Wo) {
  *Do = (D + 2 * L.pd - L.KD) / L.sd + 1;
  *Ho = (H + 2 * L.ph - L.KH) / L.sh + 1;
  *Wo = (W + 2 * L.pw - L.KW) / L.sw + 1;
}

struct ConvParams {
  const float* in;
  const float* w;
  const float* bias;
  const float* residual;
  float* out;
  float* out_hi;  // optional tf32 split of the output (for a tensor-core consumer)
  float* out_lo;
  int B, D, H, W, Cin;       // Cin = padded input channels (multiple of 4)
  int Do, Ho, Wo, Cout, CoutP;
  int KD, KH, KW, sd, sh, sw, pd, ph, pw;
  int M, K;
  int relu;
  size_t in_group_stride, w_group_stride, out_group_stride;  // blockIdx.z = group
};

#define BK 16

template <int BM, int BN>
__global__ void __launch_bounds__(256)
conv_igemm_kernel(ConvParams P) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float* in = P.in + blockIdx.z * P.in_group_stride;
  const float* wgt = P.w + blockIdx.z * P.w_group_stride;
  float* out = P.out + blockIdx.z * P.out_group_stride;
  const ActRef oref = {out, P.out_hi ? P.out_hi + blockIdx.z * P.out_group_stride : nullptr,
                       P.out_lo ? P.out_lo + blockIdx.z * P.out_group_stride : nullptr};
  const float* residual = P.residual ? P.residual + blockIdx.z * P.out_group_stride : nullptr;

  // A loader: each thread owns rows (tid/4 + 64*i) and the 4-wide k group (tid%4)
  constexpr int AROWS = BM / 64;
  const int a_kg = tid % 4;
  int a_b[AROWS], a_d[AROWS], a_h[AROWS], a_w[AROWS];
  bool a_ok[AROWS];
#pragma unroll
  for (int i = 0; i < AROWS; ++i) {
    int m = m0 + tid / 4 + 64 * i;
    a_ok[i] = m < P.M;
    int mm = a_ok[i] ? m : 0;
    int wo = mm % P.Wo; mm /= P.Wo;
    int ho = mm % P.Ho; mm /= P.Ho;
    int d_o = mm % P.Do; mm /= P.Do;
    a_b[i] = mm;
    a_d[i] = d_o * P.sd - P.pd;
    a_h[i] = ho * P.sh - P.ph;
    a_w[i] = wo * P.sw - P.pw;
  }
  // B loader: BK x BN floats = BK*BN/4 float4; 256 threads
  constexpr int BVEC = BK * BN / 4 / 256;  // float4 per thread
  float4 a_reg[AROWS];
  float4 b_reg[BVEC > 0 ? BVEC : 1];

  auto load_tile = [&](int kt) {
    const int k = kt * BK + a_kg * 4;
    int tap = k / P.Cin, c = k - tap * P.Cin;
    int kw = tap % P.KW; tap /= P.KW;
    int kh = tap % P.KH; tap /= P.KH;
    int kd = tap;
#pragma unroll
    for (int i = 0; i < AROWS; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int d = a_d[i] + kd, h = a_h[i] + kh, w = a_w[i] + kw;
      if (a_ok[i] && k < P.K && (unsigned)d < (unsigned)P.D && (unsigned)h < (unsigned)P.H &&
          (unsigned)w < (unsigned)P.W) {
        const size_t off = ((((size_t)a_b[i] * P.D + d) * P.H + h) * P.W + w) * P.Cin + c;
        v = __ldg(reinterpret_cast<const float4*>(in + off));
      }
      a_reg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < BVEC; ++i) {
      const int e = tid + i * 256;
      const int kr = e / (BN / 4), nc = (e % (BN / 4)) * 4;
      const int kk = kt * BK + kr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < P.K) v = __ldg(reinterpret_cast<const float4*>(wgt + (size_t)kk * P.CoutP + n0 + nc));
      b_reg[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < AROWS; ++i) {
      const int r = tid / 4 + 64 * i;
      As[buf][a_kg * 4 + 0][r] = a_reg[i].x;
      As[buf][a_kg * 4 + 1][r] = a_reg[i].y;
      As[buf][a_kg * 4 + 2][r] = a_reg[i].z;
      As[buf][a_kg * 4 + 3][r] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < BVEC; ++i) {
      const int e = tid + i * 256;
      const int kr = e / (BN / 4), nc = (e % (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kr][nc]) = b_reg[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nkt = (P.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nkt) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue: bias, residual, ReLU; NHWC store (row m, contiguous couts)
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= P.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const int n = n0 + tx * TN + j;
      if (n >= P.Cout) continue;
      float v[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
      if (P.bias) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += __ldg(P.bias + n + q);  // bias is CoutP long
      }
      if (n + 3 < P.Cout && (P.Cout & 3) == 0) {
        if (residual) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(residual + (size_t)m * P.Cout + n));
          v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        if (P.relu) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
        }
        store_split4(oref, (size_t)m * P.Cout + n, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (n + q < P.Cout) {
            float x = v[q];
            if (residual) x += residual[(size_t)m * P.Cout + n + q];
            if (P.relu) x = fmaxf(x, 0.f);
            store_split1(oref, (size_t)m * P.Cout + n + q, x);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Linear layers on a handful of rows (the regressor heads: M = frames of the chunk <= 32,
// K up to 1536, N up to 1024).  The work is reading the weight matrix once: a CTA owns 32 output
// columns, its 8 warps split K, every lane streams one weight column (128 B coalesced per k) and
// keeps M accumulators; x is staged in shared memory in K chunks and read as broadcast float4.
// ------------------------------------------------------------------------------------------
#define FC_MAXM 32
#define FC_KC 128

template <int M>
__global__ void __launch_bounds__(256)
linear_smallm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int Mreal,
                     int K, int CoutP, int Cout, int relu, float* __restrict__ out, int kper, float* __restrict__ part) {
  __shared__ __align__(16) float xs[M][FC_KC];
  __shared__ float red[8][M][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.f;
  // split K: blockIdx.y owns k in [kbeg, kend) and writes raw partial sums (bias / ReLU in linear_reduce_kernel)
  const int kbeg = blockIdx.y * kper, kend = min(K, kbeg + kper);
  for (int k0 = kbeg; k0 < kend; k0 += FC_KC) {
    const int kc = min(FC_KC, kend - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < M * FC_KC; e += 256) {
      const int m = e / FC_KC, k = e % FC_KC;
      xs[m][k] = (k < kc && m < Mreal) ? __ldg(x + (size_t)m * K + k0 + k) : 0.f;
    }
    __syncthreads();
    // warp `warp` takes k = warp*4 + 32*j .. +3 of the chunk
    for (int kk = warp * 4; kk < kc; kk += 32) {
      float wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) wv[q] = (kk + q < kc) ? __ldg(w + (size_t)(k0 + kk + q) * CoutP + n) : 0.f;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[m][kk]);
        acc[m] = fmaf(xv.w, wv[3], fmaf(xv.z, wv[2], fmaf(xv.y, wv[1], fmaf(xv.x, wv[0], acc[m]))));
      }
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) red[warp][m][lane] = acc[m];
  __syncthreads();
  for (int e = threadIdx.x; e < M * 32; e += 256) {
    const int m = e / 32, c = e % 32, col = blockIdx.x * 32 + c;
    if (col < Cout && m < Mreal) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) v += red[q][m][c];
      if (part) {
        part[((size_t)blockIdx.y * M + m) * CoutP + col] = v;
      } else {
        if (bias) v += bias[col];
        if (relu) v = fmaxf(v, 0.f);
        out[(size_t)m * Cout + col] = v;
      }
    }
  }
}

// sums the K-split partials in split order (deterministic), then bias and ReLU
__global__ void linear_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, int nsplit, int Mpad,
                                     int Mreal, int CoutP, int Cout, int relu, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mreal * Cout) return;
  const int m = i / Cout, col = i - m * Cout;
  float v = 0.f;
  for (int s = 0; s < nsplit; ++s) v += part[((size_t)s * Mpad + m) * CoutP + col];
  if (bias) v += bias[col];
  if (relu) v = fmaxf(v, 0.f);
  out[i] = v;
}

static bool linear_smallm_try(ss2_ctx* ctx, const ConvLayer& L, const float* x, int M, float* out, int relu, cudaStream_t st) {
  const int K = L.CinP;
  if (M > 128) return false;
  if (M > 32) {
    // row blocks of 32 through the same kernel: a row's arithmetic (and so its bits) does not depend on how many rows
    // the call has (the two-view TemporalNet batch against the per-view calls of a temporal shard)
    for (int m0 = 0; m0 < M; m0 += 32)
      if (!linear_smallm_try(ctx, L, x + (size_t)m0 * K, M - m0 < 32 ? M - m0 : 32, out + (size_t)m0 * L.Cout, relu, st)) return false;
    return true;
  }
  // few output-column CTAs (Cout / 32): split K so that the weight matrix streams through ~one wave of CTAs
  const int ncol = L.CoutP / 32;
  int nsplit = 1;
  if (ncol < 96 && K >= 4 * FC_KC) {
    nsplit = (148 + ncol - 1) / ncol;
    const int maxsplit = K / (2 * FC_KC);
    if (nsplit > maxsplit) nsplit = maxsplit;
    if (nsplit > 32) nsplit = 32;
    if (nsplit < 1) nsplit = 1;
  }
  int kper = ((K + nsplit - 1) / nsplit + FC_KC - 1) / FC_KC * FC_KC;
  nsplit = (K + kper - 1) / kper;
  const int Mpad = M <= 4 ? 4 : M <= 8 ? 8 : M <= 16 ? 16 : 32;
  float* part = nullptr;
  if (nsplit > 1) {
    part = arena_alloc<float>(ctx, (size_t)nsplit * Mpad * L.CoutP);
    if (!part) { nsplit = 1; kper = K; }
  }
  const dim3 grid(ncol, nsplit);
#define FC_CASE(MM)                                                                                                    \
  if (Mpad == MM) linear_smallm_kernel<MM><<<grid, 256, 0, st>>>(x, L.w, L.bias, M, K, L.CoutP, L.Cout, relu, out, kper, part);
  FC_CASE(4) FC_CASE(8) FC_CASE(16) FC_CASE(32)
#undef FC_CASE
  if (part) {
    ctx->launches++;
    linear_reduce_kernel<<<cdiv(M * L.Cout, 256), 256, 0, st>>>(part, L.bias, nsplit, Mpad, M, L.CoutP, L.Cout, relu, out);
  }
  return true;
}

int conv_launch(ss2_ctx* ctx, const ConvLayer& L, const ActRef& in, int B, int D, int H, int W, const ActRef& out,
                const float* d_residual, int relu, cudaStream_t st, int groups, size_t w_group_stride) {
  ConvParams P;
  P.in = in.v; P.w = L.w; P.bias = L.bias; P.residual = d_residual; P.out = out.v;
  P.out_hi = out.hi; P.out_lo = out.lo;
  P.B = B; P.D = D; P.H = H; P.W = W; P.Cin = L.CinP;
  conv_out_dims(L, D, H, W, &P.Do, &P.Ho, &P.Wo);
  P.Cout = L.Cout; P.CoutP = L.CoutP;
  P.KD = L.KD; P.KH = L.KH; P.KW = L.KW;
  P.sd = L.sd; P.sh = L.sh; P.sw = L.sw; P.pd = L.pd; P.ph = L.ph; P.pw = L.pw;
  P.M = B * P.Do * P.Ho * P.Wo;
  P.K = L.KD * L.KH * L.KW * L.CinP;
  P.relu = relu;
  P.in_group_stride = (size_t)B * D * H * W * L.CinP;
  P.w_group_stride = w_group_stride;
  P.out_group_stride = (size_t)P.M * L.Cout;
  if (P.M <= 0) return SS2_OK;
  if ((L.CinP & 3) || (L.CoutP & 63)) return ss2_fail(ctx, SS2_ERR_INVALID, "conv: unpadded layer");
  if (groups == 1 && ctx->use_tc && ctx->use_dc && (in.hi || in.h16) && conv_dc_eligible(L, D, H, W))
    return conv_dc_launch(ctx, L, in, B, H, W, out, d_residual, relu, st);
  if (groups == 1 && ctx->use_tc && (in.hi || in.h16) && conv_tc_eligible(L))
    return conv_tc_launch(ctx, L, in, B, D, H, W, out, d_residual, relu, st);
  if (!in.v) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv: split planes without plain values are read by the tcgen05 kernels only");
  if (groups == 1 && L.KD * L.KH * L.KW == 1 && D * H * W == 1 && !d_residual && !out.hi && L.sh == 1 && L.ph == 0) {
    ss2_prof_begin(ctx, SS2_PROF_CONV, st);
    const bool done = linear_smallm_try(ctx, L, in.v, B, out.v, relu, st);
    if (done) {
      ss2_prof_end(ctx, SS2_PROF_CONV, st, 2.0 * B * (double)L.Cout * L.Cin);
      SS2_LAUNCH_CHECK(ctx);
      return SS2_OK;
    }
    ss2_prof_end(ctx, SS2_PROF_CONV, st, 0.0);
  }
  ss2_prof_begin(ctx, SS2_PROF_CONV, st);
  if (P.M >= 128 * 148) {
    dim3 grid(cdiv(P.M, 128), L.CoutP / 64, groups);
    conv_igemm_kernel<128, 64><<<grid, 256, 0, st>>>(P);
  } else {
    dim3 grid(cdiv(P.M, 64), L.CoutP / 64, groups);
    conv_igemm_kernel<64, 64><<<grid, 256, 0, st>>>(P);
  }
  // algorithmic flops: 2 * M * Cout * (taps * real Cin)
  ss2_prof_end(ctx, SS2_PROF_CONV, st, 2.0 * P.M * (double)L.Cout * L.KD * L.KH * L.KW * L.Cin * groups);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// max pooling, NHWC, float4 over channels (C % 4 == 0); floor mode, -inf padding
// ------------------------------------------------------------------------------------------
__global__ void maxpool_nhwc_kernel(const float4* __restrict__ in, int B, int H, int W, int C4, int k, int s, int p,
                                    int Ho, int Wo, ActRef out) {
  const size_t total = (size_t)B * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = i;
    const int c = t % C4; t /= C4;
    const int wo = t % Wo; t /= Wo;
    const int ho = t % Ho; t /= Ho;
    const int b = (int)t;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dh = 0; dh < k; ++dh) {
      const int h = ho * s - p + dh;
      if ((unsigned)h >= (unsigned)H) continue;
      for (int dw = 0; dw < k; ++dw) {
        const int w = wo * s - p + dw;
        if ((unsigned)w >= (unsigned)W) continue;
        const float4 v = __ldg(in + (((size_t)b * H + h) * W + w) * C4 + c);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    store_split4(out, i * 4, m);
  }
}

int maxpool_launch(ss2_ctx* ctx, const float* d_in, int B, int H, int W, int C, int k, int s, int p,
                   const ActRef& out, cudaStream_t st) {
  const int Ho = (H + 2 * p - k) / s + 1, Wo = (W + 2 * p - k) / s + 1;
  const size_t total = (size_t)B * Ho * Wo * (C / 4);
  if (total == 0) return SS2_OK;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  maxpool_nhwc_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(d_in), B, H, W, C / 4, k, s, p, Ho, Wo, out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// NCHW (C <= 4) -> NHWC with 4 channels (zero padded): the network input layout change
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ in, int B, int C, int HW, float4* __restrict__ out) {
  const size_t total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / HW, p = i % HW;
    const float* src = in + b * C * HW + p;
    float4 v;
    v.x = __ldg(src);
    v.y = C > 1 ? __ldg(src + HW) : 0.f;
    v.z = C > 2 ? __ldg(src + 2 * (size_t)HW) : 0.f;
    v.w = C > 3 ? __ldg(src + 3 * (size_t)HW) : 0.f;
    out[i] = v;
  }
}

// NCHW [B,3,H,W] -> zero-padded NHWC4 split planes [B][H + 2*pad][Wp][4] (hi = rna_tf32(v), lo = v - hi): the A
// operand of the tensor-core stem (conv_tc_stem_launch): the 7x7 stride-2 padding is materialised so that one output
// pixel's filter row is 32 contiguous floats
__global__ void nchw_to_nhwc4_pad_split_kernel(const float* __restrict__ in, int B, int H, int W, int pad, int Hp, int Wp,
                                               float4* __restrict__ hi, float4* __restrict__ lo) {
  const size_t total = (size_t)B * Hp * Wp, HW = (size_t)H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xp = (int)(i % Wp), yp = (int)((i / Wp) % Hp);
    const size_t b = i / ((size_t)Wp * Hp);
    const int x = xp - pad, y = yp - pad;
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const float* src = in + b * 3 * HW + (size_t)y * W + x;
      tf32_split(__ldg(src), &h.x, &l.x);
      tf32_split(__ldg(src + HW), &h.y, &l.y);
      tf32_split(__ldg(src + 2 * HW), &h.z, &l.z);
    }
    hi[i] = h;
    lo[i] = l;
  }
}

int nchw_to_nhwc4_pad_split_launch(ss2_ctx* ctx, const float* d_in, int B, int H, int W, int pad, int Hp, int Wp,
                                   float* d_hi, float* d_lo, cudaStream_t st) {
  const size_t total = (size_t)B * Hp * Wp;
  if (total == 0) return SS2_OK;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  nchw_to_nhwc4_pad_split_kernel<<<blocks, 256, 0, st>>>(d_in, B, H, W, pad, Hp, Wp, reinterpret_cast<float4*>(d_hi),
                                                         reinterpret_cast<float4*>(d_lo));
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

int nchw_to_nhwc4_launch(ss2_ctx* ctx, const float* d_in, int B, int C, int H, int W, float* d_out,
                         cudaStream_t st) {
  const size_t total = (size_t)B * H * W;
  if (total == 0) return SS2_OK;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  nchw_to_nhwc4_kernel<<<blocks, 256, 0, st>>>(d_in, B, C, H * W, reinterpret_cast<float4*>(d_out));
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
