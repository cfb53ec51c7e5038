// Correlation layers: the local cost volume (warp-shuffle channel reduction) and the global
// contextual correlation layer (CCL) of SpatialNet.
//
// Reference behaviour restated (paths under Full_model_inference/Codes/):
//   spatial_network.py:333-358 / temporal_network.py:149-174   cost_volume (norm=False)
//   spatial_network.py:369-425                                 CCL
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// cost volume: cv[d] = leaky_relu_0.1( mean_c x1[c] * x2[c, y+j-sr, x+i-sr] ), d = j*(2sr+1)+i
// NHWC, C = 128 (one float4 per lane).  One warp per output pixel: x1's channel vector stays
// in registers, each displacement is one coalesced 512 B read of x2, 4 FMAs and a 5-step
// butterfly; lane (d mod 32) keeps result d so the stores are coalesced 128 B rows.
// Output has CP channels (>= (2sr+1)^2, zero filled) so the next conv sees a padded Cin.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cost_volume_kernel(const float4* __restrict__ x1, const float4* __restrict__ x2, int H, int W, int C4, int sr, int CP,
                   ActRef out) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix % W;
  const size_t img = (size_t)b * H * W;
  const bool act = lane < C4;  // C <= 128: one float4 per lane
  const float4 a = act ? __ldg(x1 + (img + pix) * C4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv_c = 1.0f / (float)(4 * C4);
  const int k = 2 * sr + 1, nd = k * k;
  const size_t o = (img + pix) * CP;
  for (int d0 = 0; d0 < CP; d0 += 32) {
    float keep = 0.f;
    for (int dd = 0; dd < 32; ++dd) {
      const int d = d0 + dd;
      if (d >= nd) break;
      const int j = d / k, i = d % k;
      const int y2 = y + j - sr, x2c = x + i - sr;
      float s = 0.f;
      if ((unsigned)y2 < (unsigned)H && (unsigned)x2c < (unsigned)W) {  // warp-uniform branch
        const float4 v = act ? __ldg(x2 + (img + (size_t)y2 * W + x2c) * C4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        s = fmaf(a.w, v.w, fmaf(a.z, v.z, fmaf(a.y, v.y, a.x * v.x)));
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
        s *= inv_c;
        s = s > 0.f ? s : 0.1f * s;
      }
      if (lane == dd) keep = s;
    }
    if (d0 + lane < CP) store_split1(out, o + d0 + lane, keep);
  }
}

// ------------------------------------------------------------------------------------------
// Blocked cost volume for C = 128 (the production shape).  A CTA owns an 8x8 pixel tile and stages x1 (64 px) and the
// x2 halo ((8+2sr)^2 px) in shared memory 32 channels at a time, double buffered with cp.async (zero fill outside
// the image = the reference's F.pad).  A thread owns ONE ROW of 8 pixels x ONE displacement row j x all (2sr+1)
// horizontal displacements (88 accumulators for sr = 5) for half of the channels: per 4 channels it loads the 8 x1
// vectors of its pixel row and the 8+2sr x2 vectors of halo row py + j once and issues 8 x (2sr+1) x 4 FMAs from
// them - 13.5 FMAs per LDS.128, so the kernel is bound by the FMA pipe and no longer by shared-memory bandwidth.
// (The previous organisation - a pixel PAIR per thread - fed 7 FMAs per LDS.128 with 4 wavefronts each and ran at
// 160 us per launch for 18 us of FMA work.)  Lanes of a warp differ in (py, j), i.e. read different rows: the row
// pitches are padded to 4 banks modulo 32 so that 8 consecutive rows hit 8 different bank groups.
// The two channel halves meet in shared memory, which also turns the result into coalesced NHWC rows (+ tf32 split).
// ------------------------------------------------------------------------------------------
#define CVT 8          // tile edge
#define CV_CK 32       // channels per stage
#define CV_LD 32       // floats per staged pixel (all lanes of a load read the same pixel column: no per-pixel padding)
#define CV_X1_PITCH (CVT * CV_LD + 4)   // floats per x1 tile row: 260 = 4 (mod 32)

__device__ __forceinline__ void cv_cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}

template <int SR> struct CvShape {
  static constexpr int KD = 2 * SR + 1, HALO = CVT + 2 * SR;
  // floats per x2 halo row, padded to 4 (mod 32)
  static constexpr int X2_PITCH = HALO * CV_LD + ((4 - (HALO * CV_LD) % 32) + 32) % 32;
  static constexpr int STAGE = CVT * CV_X1_PITCH + HALO * X2_PITCH;   // floats per buffer
  static constexpr int HALF_THREADS = ((CVT * KD + 31) / 32) * 32;   // each channel half starts on a warp boundary
  static constexpr int THREADS = 2 * HALF_THREADS;
  static constexpr int SO_PITCH = CVT * 128 + 4;                     // floats per tile row of the result buffer (4 mod 32)
};

template <int SR>
__global__ void __launch_bounds__(CvShape<SR>::THREADS)
cost_volume_tiled_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int H, int W, int CP, ActRef out) {
  using S = CvShape<SR>;
  constexpr int KD = S::KD, HALO = S::HALO, NT = S::THREADS;
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.z, ty0 = blockIdx.y * CVT, tx0 = blockIdx.x * CVT;
  const int tid = threadIdx.x;
  const int idx = tid % S::HALF_THREADS, half = tid / S::HALF_THREADS;
  const int py = idx % CVT, j = idx / CVT;
  const bool active = idx < CVT * KD;   // the last lanes of each half's last warp idle
  const size_t img = (size_t)b * H * W;
  float acc[CVT][KD];
#pragma unroll
  for (int p = 0; p < CVT; ++p)
#pragma unroll
    for (int i = 0; i < KD; ++i) acc[p][i] = 0.f;
  // stage x1 tile and x2 halo of channel chunk c0 into buffer `buf` (zero outside the image), 8 float4 per pixel
  auto stage = [&](int buf, int c0) {
    float* s1 = sm + buf * S::STAGE;
    float* s2 = s1 + CVT * CV_X1_PITCH;
    for (int e = tid; e < CVT * CVT * (CV_CK / 4); e += NT) {
      const int p = e / (CV_CK / 4), q = e % (CV_CK / 4);
      const int y = ty0 + p / CVT, x = tx0 + p % CVT;
      const bool ok = y < H && x < W;
      cv_cp_async16(s1 + (p / CVT) * CV_X1_PITCH + (p % CVT) * CV_LD + q * 4, ok ? x1 + (img + (size_t)y * W + x) * 128 + c0 + q * 4 : x1, ok);
    }
    for (int e = tid; e < HALO * HALO * (CV_CK / 4); e += NT) {
      const int p = e / (CV_CK / 4), q = e % (CV_CK / 4);
      const int y = ty0 - SR + p / HALO, x = tx0 - SR + p % HALO;
      const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
      cv_cp_async16(s2 + (p / HALO) * S::X2_PITCH + (p % HALO) * CV_LD + q * 4, ok ? x2 + (img + (size_t)y * W + x) * 128 + c0 + q * 4 : x2, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  for (int ck = 0; ck < 128 / CV_CK; ++ck) {
    if (ck + 1 < 128 / CV_CK) {
      stage((ck + 1) & 1, (ck + 1) * CV_CK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      const float* s1 = sm + (ck & 1) * S::STAGE + py * CV_X1_PITCH;
      const float* s2 = sm + (ck & 1) * S::STAGE + CVT * CV_X1_PITCH + (py + j) * S::X2_PITCH;
      for (int qq = 0; qq < CV_CK / 8; ++qq) {
        const int q = 2 * qq + half;
        float4 a[CVT];
#pragma unroll
        for (int p = 0; p < CVT; ++p) a[p] = *reinterpret_cast<const float4*>(s1 + p * CV_LD + q * 4);
#pragma unroll
        for (int c = 0; c < HALO; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(s2 + c * CV_LD + q * 4);
#pragma unroll
          for (int p = 0; p < CVT; ++p) {
            if (c - p >= 0 && c - p < KD)   // compile time
              acc[p][c - p] = fmaf(a[p].w, v.w, fmaf(a[p].z, v.z, fmaf(a[p].y, v.y, fmaf(a[p].x, v.x, acc[p][c - p]))));
          }
        }
      }
    }
    __syncthreads();   // all reads of this buffer are done before the chunk after next is staged into it
  }
  // the two channel halves go to two shared-memory buffers [8 rows][8 px][CP] (row pitch 4 mod 32: the lanes of a warp
  // differ in py and j) and are added on the way out as coalesced NHWC rows
  float* so = sm + half * (CVT * S::SO_PITCH);
  if (active) {
#pragma unroll
    for (int p = 0; p < CVT; ++p)
#pragma unroll
      for (int i = 0; i < KD; ++i) so[py * S::SO_PITCH + p * CP + j * KD + i] = acc[p][i];
  }
  __syncthreads();
  for (int e = tid; e < CVT * CVT * (CP / 4); e += NT) {
    const int p = e / (CP / 4), q = e % (CP / 4);
    const int y = ty0 + p / CVT, x = tx0 + p % CVT;
    if (y < H && x < W) {
      const float* s0 = sm + (p / CVT) * S::SO_PITCH + (p % CVT) * CP + q * 4;
      const float4 u0 = *reinterpret_cast<const float4*>(s0);
      const float4 u1 = *reinterpret_cast<const float4*>(s0 + CVT * S::SO_PITCH);
      float v[4] = {u0.x + u1.x, u0.y + u1.y, u0.z + u1.z, u0.w + u1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] *= (1.0f / 128.0f);
        v[k] = v[k] > 0.f ? v[k] : 0.1f * v[k];
        if (q * 4 + k >= KD * KD) v[k] = 0.f;   // padded channels of the next convolution
      }
      store_split4(out, (img + (size_t)y * W + x) * CP + q * 4, make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

int cost_volume_launch(ss2_ctx* ctx, const float* d_x1, const float* d_x2, int B, int H, int W, int C, int sr,
                       int CP, const ActRef& out, cudaStream_t st) {
  if (C <= 0 || C > 128 || (C & 3)) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "cost_volume: C must be a multiple of 4, <= 128 (got %d)", C);
  if ((2 * sr + 1) * (2 * sr + 1) > CP || (CP & 31)) return ss2_fail(ctx, SS2_ERR_INVALID, "cost_volume: CP must be a multiple of 32 and >= (2sr+1)^2");
  if (B <= 0) return SS2_OK;
  if (C == 128 && (sr == 5 || sr == 3) && CP <= 128 && (CP & 3) == 0) {
    const size_t smem = (size_t)2 * (sr == 5 ? CvShape<5>::STAGE : CvShape<3>::STAGE) * sizeof(float);   // two channel-chunk buffers
    dim3 g(cdiv(W, CVT), cdiv(H, CVT), B);
    static bool attr_dev[16] = {false};  // per device: function attributes live in the device's context
    bool& attr = attr_dev[ctx->device & 15];
    if (!attr) {
      SS2_CUDA(ctx, cudaFuncSetAttribute(cost_volume_tiled_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      SS2_CUDA(ctx, cudaFuncSetAttribute(cost_volume_tiled_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
      attr = true;
    }
    if (sr == 5) cost_volume_tiled_kernel<5><<<g, CvShape<5>::THREADS, smem, st>>>(d_x1, d_x2, H, W, CP, out);
    else cost_volume_tiled_kernel<3><<<g, CvShape<3>::THREADS, smem, st>>>(d_x1, d_x2, H, W, CP, out);
    SS2_LAUNCH_CHECK(ctx);
    return SS2_OK;
  }
  dim3 grid(cdiv(H * W, 8), B);
  cost_volume_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(d_x1),
                                          reinterpret_cast<const float4*>(d_x2), H, W, C / 4, sr, CP, out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// CCL
//   1. L2-normalise both feature maps over C (F.normalize, eps 1e-12)
//   2. every 3x3 patch of n2 becomes a correlation filter: Wp[b][(tap,c)][k], k = h2*W + w2
//   3. match[b][p][k] = conv3x3(n1[b], Wp[b])           -> the implicit-GEMM conv kernel
//   4. softmax_k(10*match) expectation of (k%W - w, k//W - h) -> flow (flow_w, flow_h, 0, 0)
// ------------------------------------------------------------------------------------------
__global__ void l2norm_nhwc_kernel(const float* __restrict__ in, int npix, int C, ActRef out) {
  const int pix = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (pix >= npix) return;
  const float* p = in + (size_t)pix * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float v = p[c]; s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < C; c += 32) store_split1(out, (size_t)pix * C + c, p[c] * inv);
}

__global__ void ccl_filters_kernel(const float* __restrict__ n2, int H, int W, int C, int KP,
                                   float* __restrict__ Wp) {
  // grid: (cdiv(KP,128), 9*C, B) ; thread -> k
  const int b = blockIdx.z, row = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= KP) return;
  const int tap = row / C, c = row % C;
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  float v = 0.f;
  if (k < H * W) {
    const int h = k / W + dy, w = k % W + dx;
    if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) v = __ldg(n2 + (((size_t)b * H + h) * W + w) * C + c);
  }
  Wp[((size_t)b * 9 * C + row) * KP + k] = v;
}

__global__ void ccl_softmax_flow_kernel(const float* __restrict__ match, int H, int W, int ld, float4* __restrict__ flow) {
  const int HW = H * W;
  const int b = blockIdx.y;
  const int pix = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (pix >= HW) return;
  const float* m = match + ((size_t)b * HW + pix) * ld;
  float mx = -INFINITY;
  for (int k = lane; k < HW; k += 32) mx = fmaxf(mx, m[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float ph = (float)(pix / W), pw = (float)(pix % W);
  float se = 0.f, sh = 0.f, sw = 0.f;
  for (int k = lane; k < HW; k += 32) {
    const float e = expf(10.0f * m[k] - 10.0f * mx);
    se += e;
    sh = fmaf(e, (float)(k / W) - ph, sh);
    sw = fmaf(e, (float)(k % W) - pw, sw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  if (lane == 0) flow[(size_t)b * HW + pix] = make_float4(sw / se, sh / se, 0.f, 0.f);
}

int ccl_launch(ss2_ctx* ctx, const float* d_f1, const float* d_f2, int B, int H, int W, int C, float* d_flow,
               cudaStream_t st) {
  if (B <= 0) return SS2_OK;
  const int HW = H * W;
  const int KP = (HW + 63) / 64 * 64;
  const bool tc = ctx->use_tc && (C % 32) == 0 && W <= 64 && conv_tc_corr_rows(W) >= 1;
  const size_t nact = ((size_t)B * HW * C + 63) / 64 * 64;
  const bool f16 = tc && (ctx->use_f16 & 2) && ctx->tc_passes != 1 && (C % 64) == 0;   // fp16 split planes, kind::f16 correlation
  ActRef n1, n2;
  n1.v = arena_alloc<float>(ctx, (f16 ? 2 : tc ? 3 : 1) * nact);
  n2.v = arena_alloc<float>(ctx, (f16 ? 2 : tc ? 3 : 1) * nact);
  float* match = arena_alloc<float>(ctx, (size_t)B * HW * KP);
  if (!n1.v || !n2.v || !match) return ss2_fail(ctx, SS2_ERR_OOM, "ccl: workspace arena exhausted");
  if (f16) {
    n1.h16 = reinterpret_cast<__half*>(n1.v + nact); n1.l16 = n1.h16 + nact; n1.flag = ctx->d_range_flag;
    n2.h16 = reinterpret_cast<__half*>(n2.v + nact); n2.l16 = n2.h16 + nact; n2.flag = ctx->d_range_flag;
  } else if (tc) { n1.hi = n1.v + nact; n1.lo = n1.v + 2 * nact; n2.hi = n2.v + nact; n2.lo = n2.v + 2 * nact; }
  l2norm_nhwc_kernel<<<cdiv(B * HW, 8), 256, 0, st>>>(d_f1, B * HW, C, n1);
  SS2_LAUNCH_CHECK(ctx);
  l2norm_nhwc_kernel<<<cdiv(B * HW, 8), 256, 0, st>>>(d_f2, B * HW, C, n2);
  SS2_LAUNCH_CHECK(ctx);
  int ld = HW;
  if (tc) {
    ld = KP;
    SS2_TRY(conv_tc_corr_launch(ctx, n1, n2, B, H, W, C, match, ld, st));
  } else {
    float* Wp = arena_alloc<float>(ctx, (size_t)B * 9 * C * KP);
    if (!Wp) return ss2_fail(ctx, SS2_ERR_OOM, "ccl: workspace arena exhausted");
    ccl_filters_kernel<<<dim3(cdiv(KP, 128), 9 * C, B), 128, 0, st>>>(n2.v, H, W, C, KP, Wp);
    SS2_LAUNCH_CHECK(ctx);
    ConvLayer L;
    L.w = Wp; L.bias = nullptr;
    L.Cin = L.CinP = C; L.Cout = HW; L.CoutP = KP;
    L.KH = L.KW = 3; L.ph = L.pw = 1;
    ActRef a_in, a_out;
    a_in.v = n1.v; a_out.v = match;
    SS2_TRY(conv_launch(ctx, L, a_in, 1, 1, H, W, a_out, nullptr, 0, st, B, (size_t)9 * C * KP));
  }
  ccl_softmax_flow_kernel<<<dim3(cdiv(HW, 8), B), 256, 0, st>>>(match, H, W, ld, reinterpret_cast<float4*>(d_flow));
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
