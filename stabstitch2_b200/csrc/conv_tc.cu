// tcgen05 tensor-core implicit-GEMM convolution for sm_100a (NHWC / NDHWC activations).
//
//   out[m][n] = act( sum_k A[m][k] * Wt[n][k] + bias[n] + residual[m][n] )
//   m = (b, do, ho, wo) output position, k = (kd, kh, kw, cin), n = cout
//
// Data path
//   * A (im2col rows) never exists in memory: for one filter tap and one 32-channel slice the
//     rows of an output tile (TN images x TT x TH x TW positions, <= 128 rows) are ONE 5-D TMA
//     box of the activation tensor {C, W, H, D, B} (element strides = conv strides, out-of-range
//     coordinates zero-filled by the TMA unit = the conv padding), landing in shared memory as a
//     K-major SWIZZLE_128B tile - exactly the layout the UMMA shared-memory descriptor reads.
//   * B (weights) is a 2-D TMA box {32 k, 64 cout} of the K-major packed filter matrix.
//   * D accumulates in TMEM (128 lanes x 64 fp32 columns); one elected thread issues
//     tcgen05.mma.kind::tf32 (M=128, N=64, K=8), tcgen05.commit releases the smem stage.
//   * fp32-grade results ("3xTF32"): activations and weights are stored as hi = rna_tf32(v) and
//     lo = v - hi; D += Ahi*Bhi + Alo*Bhi + Ahi*Blo (error ~2^-21 relative instead of TF32's
//     2^-11).  NPASS=1 runs plain TF32 (what cuDNN does for the reference's GPU convolutions).
//   * F16 (the default behind the stem): the same three products on fp16 split planes, h16 = fp16(v) and
//     l16 = fp16((v - h16) * 2048), as kind::f16 MMAs (M=128, N=64, K=16: a 128-byte row holds 64 channels); the two
//     correction products carry the factor 2048 and accumulate in D2, the epilogue takes D1 + D2 / 2048 (common.cuh).
//   * epilogue: 4 warps tcgen05.ld their 32 TMEM lanes, add bias / residual, ReLU, and store the
//     row as fp32 plus its (hi, lo) split for the next tensor-core layer.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane
// quarter = warp_id % 4), warp 2 also owns the TMEM allocation.
//
// Replaces cuDNN under nn.Conv2d / nn.Conv3d of spatial_network.py:147-259, temporal_network.py:
// 65-104, smooth_network.py:124-131 and torchvision's BasicBlocks (spatial_network.py:123-139).
#include "tc_common.cuh"

#define TC_BM 128
#define TC_BN 64
#define TC_BK 32
#define TC_THREADS 192
// pipeline depth: 2 x 48 KB (split-TF32) / 4 x 24 KB (plain TF32) keeps two CTAs per SM resident,
// so one CTA's epilogue and pipeline fill overlap the other's main loop
#ifndef TC_STAGES3
#define TC_STAGES3 2
#endif
#ifndef TC_STAGES1
#define TC_STAGES1 4
#endif
#define TC_A_BYTES (TC_BM * TC_BK * 4)  // 16 KB
#define TC_B_BYTES (TC_BN * TC_BK * 4)  // 8 KB

struct TcParams {
  const float* bias;      // [CoutP] or null
  const float* residual;  // plain fp32 [M][Cout] or null
  float* out_v;           // plain fp32 [M][Cout]
  float* out_hi;          // split planes for the next tensor-core layer (may be null)
  float* out_lo;
  __half* out_h16;        // fp16 split planes for a kind::f16 consumer (conv_dc.cu; may be null)
  __half* out_l16;
  int* range_flag;     // raised when a value does not fit fp16 (common.cuh)
  int B, Do, Ho, Wo, Cout;
  int TN, TT, TH, TW;      // tile extents (images, depth, rows, cols); rows = TN*TT*TH*TW <= 128
  int nN, nT, nH, nW;      // tile counts per dimension
  int KD, KH, KW, nchunk;  // filter taps and CinP/32
  int sd, sh, sw, pd, ph, pw;
  int CinP;
  int relu;
  // correlation mode (CCL): the B operand is the im2col of a second activation tensor
  // (spatial_network.py:369-425: every 3x3 patch of f2 is a filter); N tile = bTH x bTW of its pixels
  int b_act, bTH, bTW;
  int ldo;   // output row stride in floats (== Cout for convolutions)
  int ncol;  // output columns per N tile (64 for convolutions)
};


// F16 (with NPASS == 3): the operands are fp16 split planes and the MMAs kind::f16, exactly as in conv_dc.cu: a 128-byte
// row is 64 channels, K = 16 per instruction, A_l x B_h accumulates into the 2048-scaled D2, the epilogue takes D1 + D2 / 2048.
template <int NPASS, int STAGES, bool F16 = false>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, TcParams P) {
  constexpr int NOPER = NPASS == 3 ? 2 : 1;  // operand planes per matrix (hi [, lo])
  constexpr int STAGE_BYTES = NOPER * (TC_A_BYTES + TC_B_BYTES);
  constexpr int BKC = F16 ? 2 * TC_BK : TC_BK;   // channels per chunk (one 128-byte row)
  // split TF32 in TWO MMAs per k-step (see conv_dc.cu): B_hi | B_lo of a stage are adjacent, so A_hi x [B_hi | B_lo] is
  // one N = 128 MMA into accumulator columns [D1 | D2] and A_lo x B_hi an N = 64 one into D1; the epilogue adds D1 + D2
  constexpr int TC_ACC_COLS = NPASS == 3 ? 2 * TC_BN : TC_BN;
  __shared__ float4 ep_stage[4][32 * 4];   // per epilogue warp: 32 rows x 16 columns, 16-byte slots XOR-swizzled
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int tix = blockIdx.x;
  const int tw = tix % P.nW; tix /= P.nW;
  const int th = tix % P.nH; tix /= P.nH;
  const int tt = tix % P.nT; tix /= P.nT;
  const int tn = tix;
  const int w0 = tw * P.TW, h0 = th * P.TH, t0 = tt * P.TT, n0 = tn * P.TN;
  const int cout0 = blockIdx.y * P.ncol;
  const int rows = P.TN * P.TT * P.TH * P.TW;
  const int ntaps = P.KD * P.KH * P.KW;
  const int niter = ntaps * P.nchunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TC_ACC_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (dc_elect_one()) {
      const uint32_t b_bytes = P.b_act ? (uint32_t)(P.bTH * P.bTW) * TC_BK * 4 : (uint32_t)TC_B_BYTES;
      const uint32_t tx_bytes = (uint32_t)NOPER * ((uint32_t)rows * TC_BK * 4 + b_bytes);
      const int bh0 = blockIdx.y * P.bTH;
      int it = 0;
      for (int tap = 0; tap < ntaps; ++tap) {
        const int kw = tap % P.KW, kh = (tap / P.KW) % P.KH, kd = tap / (P.KW * P.KH);
        const int cw = w0 * P.sw + kw - P.pw, ch = h0 * P.sh + kh - P.ph, cd = t0 * P.sd + kd - P.pd;
        for (int ck = 0; ck < P.nchunk; ++ck, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
          uint8_t* st = smem + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], tx_bytes);
          tma_load_5d(st, &tmA_hi, &full_bar[s], ck * BKC, cw, ch, cd, n0);
          if (P.b_act) tma_load_5d(st + NOPER * TC_A_BYTES, &tmB_hi, &full_bar[s], ck * BKC, kw - P.pw, bh0 + kh - P.ph, 0, n0);
          else tma_load_2d(st + NOPER * TC_A_BYTES, &tmB_hi, &full_bar[s], tap * P.CinP + ck * BKC, cout0);
          if (NPASS == 3) {
            tma_load_5d(st + TC_A_BYTES, &tmA_lo, &full_bar[s], ck * BKC, cw, ch, cd, n0);
            if (P.b_act) tma_load_5d(st + NOPER * TC_A_BYTES + TC_B_BYTES, &tmB_lo, &full_bar[s], ck * BKC, kw - P.pw, bh0 + kh - P.ph, 0, n0);
            else tma_load_2d(st + NOPER * TC_A_BYTES + TC_B_BYTES, &tmB_lo, &full_bar[s], tap * P.CinP + ck * BKC, cout0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected thread; descriptors advanced with 32-bit adds, see tc_common.cuh) =====
    if (dc_elect_one()) {
      constexpr uint32_t FMT = F16 ? 0u : ((2u << 7) | (2u << 10));   // kind::f16: A/B format 0 = fp16
      const uint32_t idesc = (1u << 4) | FMT | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | FMT | ((uint32_t)(2 * TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int it = 0; it < niter; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = dc_desc_lo(smem_u32(smem + (size_t)s * STAGE_BYTES));
        const uint32_t b_hi = a_hi + ((NOPER * TC_A_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          if (NPASS == 3 && F16) {
            dc_mma_f16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, DC_DESC_HI, idesc2, (it > 0 || k > 0) ? 1u : 0u);
            dc_mma_f16(tmem_base + TC_BN, a_hi + (TC_A_BYTES >> 4) + 2 * k, b_hi + 2 * k, DC_DESC_HI, idesc, 1u);   // scaled: into D2
          } else if (NPASS == 3) {
            dc_mma(tmem_base, a_hi + 2 * k, b_hi + 2 * k, DC_DESC_HI, idesc2, (it > 0 || k > 0) ? 1u : 0u);
            dc_mma(tmem_base, a_hi + (TC_A_BYTES >> 4) + 2 * k, b_hi + 2 * k, DC_DESC_HI, idesc, 1u);
          } else {
            dc_mma(tmem_base, a_hi + 2 * k, b_hi + 2 * k, DC_DESC_HI, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above have read it
      }
      umma_commit(&accum_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter q = warp % 4 =====
    // Coalesced store mapping as in conv_dc.cu: 16 accumulator columns at a time are transposed through a swizzled 2 KB
    // buffer, then lane l owns 16 bytes (l & 3) of rows (l >> 2) + 8 i.  These warps idle during the main loop, so the
    // residual of the first 32 columns is loaded before the wait for the accumulator, the second half during the first.
    const int q = warp & 3;
    float4* stg = ep_stage[q];
    size_t mo[4];
    bool vv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = q * 32 + (lane >> 2) + 8 * i;
      int rr = r;
      const int wl = rr % P.TW; rr /= P.TW;
      const int hl = rr % P.TH; rr /= P.TH;
      const int tl = rr % P.TT; rr /= P.TT;
      const int nl = rr;
      const int ow = w0 + wl, oh = h0 + hl, ot = t0 + tl, on = n0 + nl;
      vv[i] = r < rows && ow < P.Wo && oh < P.Ho && ot < P.Do && on < P.B;
      mo[i] = ((((size_t)on * P.Do + ot) * P.Ho + oh) * P.Wo + ow) * P.ldo;
    }
    auto col_ok = [&](int half, int sub) {
      const int cl = half * 32 + sub * 16 + (lane & 3) * 4;
      return cout0 + cl < P.Cout && cl < P.ncol;
    };
    auto residual_issue = [&](int half, float4 (&rs)[8]) {
#pragma unroll
      for (int sub = 0; sub < 2; ++sub)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          rs[sub * 4 + i] = (vv[i] && col_ok(half, sub))
                                ? __ldg(reinterpret_cast<const float4*>(P.residual + mo[i] + cout0 + half * 32 + sub * 16 + (lane & 3) * 4))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 rs[8];
    if (P.residual) residual_issue(0, rs);
    mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int half = 0; half < TC_BN / 32; ++half) {
      float4 rc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) rc[e] = rs[e];
      if (P.residual && half + 1 < TC_BN / 32) residual_issue(half + 1, rs);
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + half * 32, acc);  // warp-collective
      if (NPASS == 3) {   // D1 + D2
        uint32_t acc2[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TC_BN + half * 32, acc2);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          acc[j] = __float_as_uint(F16 ? fmaf(__uint_as_float(acc2[j]), 1.0f / SS2_F16_LO_SCALE, __uint_as_float(acc[j]))
                                       : __uint_as_float(acc[j]) + __uint_as_float(acc2[j]));
      }
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          stg[lane * 4 + (jj ^ ((lane >> 1) & 3))] =
              make_float4(__uint_as_float(acc[sub * 16 + jj * 4]), __uint_as_float(acc[sub * 16 + jj * 4 + 1]),
                          __uint_as_float(acc[sub * 16 + jj * 4 + 2]), __uint_as_float(acc[sub * 16 + jj * 4 + 3]));
        __syncwarp();
        const bool cok = col_ok(half, sub);
        const int c = cout0 + half * 32 + sub * 16 + (lane & 3) * 4;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.bias && cok) bb = __ldg(reinterpret_cast<const float4*>(P.bias + c));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = (lane >> 2) + 8 * i;
          const float4 a = stg[row * 4 + ((lane & 3) ^ ((row >> 1) & 3))];
          if (vv[i] && cok) {
            const size_t o = mo[i] + c;
            float v[4] = {a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w};
            if (P.residual) {
              const float4 r4 = rc[sub * 4 + i];
              v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
            }
            if (P.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (P.out_v) *reinterpret_cast<float4*>(P.out_v + o) = make_float4(v[0], v[1], v[2], v[3]);
            if (P.out_hi) {
              float hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) { hi[e] = rna_tf32(v[e]); lo[e] = v[e] - hi[e]; }
              *reinterpret_cast<float4*>(P.out_hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
              if (P.out_lo) *reinterpret_cast<float4*>(P.out_lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (P.out_h16) store_f16_planes4(P.out_h16, P.out_l16, o, v, P.range_flag);
          }
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_ACC_COLS));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* ss2_tensormap_encode_fn();
static EncodeTiledFn encode_fn() { return reinterpret_cast<EncodeTiledFn>(ss2_tensormap_encode_fn()); }
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda symbol dependency); shared with tps.cu
void* ss2_tensormap_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return reinterpret_cast<void*>(fn);
}

// activation tensor {C, W, H, D, B} (fp32, innermost first) with a box of one output tile
int make_act_map(ss2_ctx* ctx, CUtensorMap* map, const void* base, int C, int W, int H, int D, int B, int bw,
                        int bh, int bd, int bn, int sw, int sh, int sd, bool f16) {
  const cuuint64_t es = f16 ? 2 : 4;   // a 128-byte inner box either way: 32 fp32 or 64 fp16 channels
  EncodeTiledFn fn = encode_fn();
  if (!fn) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es, (cuuint64_t)D * H * W * C * es};
  cuuint32_t box[5] = {(cuuint32_t)(f16 ? 2 * TC_BK : TC_BK), (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, (cuuint32_t)bn};
  cuuint32_t est[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)sd, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, strides, box, est,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return ss2_fail(ctx, SS2_ERR_CUDA, "cuTensorMapEncodeTiled(activations C=%d W=%d H=%d D=%d B=%d box %d,%d,%d,%d stride %d,%d,%d) = %d",
                    C, W, H, D, B, bw, bh, bd, bn, sw, sh, sd, (int)r);
  return SS2_OK;
}

int make_weight_map(ss2_ctx* ctx, CUtensorMap* map, const void* base, int Ktot, int CoutP, bool f16) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)CoutP};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(f16 ? 2 * TC_BK : TC_BK), TC_BN};
  cuuint32_t est[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, est,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ss2_fail(ctx, SS2_ERR_CUDA, "cuTensorMapEncodeTiled(weights K=%d N=%d) = %d", Ktot, CoutP, (int)r);
  return SS2_OK;
}

// K-major hi/lo filter matrices [CoutP][Ktot] are built at pack time (nets.cu: pack_conv)
bool conv_tc_eligible(const ConvLayer& L) {
  return L.wk_hi != nullptr && (L.CinP % TC_BK) == 0 && (L.CoutP % TC_BN) == 0 && (L.Cout % 4) == 0;
}

static void tile_shape(int B, int Do, int Ho, int Wo, int* TN, int* TT, int* TH, int* TW) {
  int tw = Wo;
  if (tw > TC_BM) { const int parts = (Wo + TC_BM - 1) / TC_BM; tw = (Wo + parts - 1) / parts; }
  int th = TC_BM / tw; if (th > Ho) th = Ho; if (th < 1) th = 1;
  int tt = TC_BM / (tw * th); if (tt > Do) tt = Do; if (tt < 1) tt = 1;
  int tn = TC_BM / (tw * th * tt); if (tn > B) tn = B; if (tn < 1) tn = 1;
  *TN = tn; *TT = tt; *TH = th; *TW = tw;
}

int conv_tc_launch(ss2_ctx* ctx, const ConvLayer& L, const ActRef& in, int B, int D, int H, int W, const ActRef& out,
                   const float* d_residual, int relu, cudaStream_t st) {
  TcParams P;
  conv_out_dims(L, D, H, W, &P.Do, &P.Ho, &P.Wo);
  P.B = B; P.Cout = L.Cout; P.CinP = L.CinP;
  const bool f16 = in.h16 != nullptr && in.hi == nullptr;   // fp16 split planes in: kind::f16 MMAs
  if (f16 && (!in.l16 || !L.wk_h16 || !L.wk_l16 || (L.CinP % 64) != 0 || ctx->tc_passes == 1))
    return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc: fp16 planes given to a layer without fp16 filter planes");
  P.KD = L.KD; P.KH = L.KH; P.KW = L.KW; P.nchunk = L.CinP / (f16 ? 2 * TC_BK : TC_BK);
  P.sd = L.sd; P.sh = L.sh; P.sw = L.sw; P.pd = L.pd; P.ph = L.ph; P.pw = L.pw;
  P.relu = relu; P.bias = L.bias; P.residual = d_residual;
  P.b_act = 0; P.bTH = P.bTW = 0; P.ldo = L.Cout; P.ncol = TC_BN;
  P.out_v = out.v; P.out_hi = out.hi; P.out_lo = out.lo;
  P.out_h16 = out.h16; P.out_l16 = out.h16 ? out.l16 : nullptr; P.range_flag = ctx->d_range_flag;
  tile_shape(B, P.Do, P.Ho, P.Wo, &P.TN, &P.TT, &P.TH, &P.TW);
  P.nW = cdiv(P.Wo, P.TW); P.nH = cdiv(P.Ho, P.TH); P.nT = cdiv(P.Do, P.TT); P.nN = cdiv(B, P.TN);
  const int npass = f16 ? 3 : (ctx->tc_passes == 1 || !in.lo || !L.wk_lo) ? 1 : 3;
  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
  const int bw = (P.TW - 1) * L.sw + 1, bh = (P.TH - 1) * L.sh + 1, bd = (P.TT - 1) * L.sd + 1;
  if (bw > 256 || bh > 256 || bd > 256) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc: TMA box too large");
  if (f16) {
    SS2_TRY(make_act_map(ctx, &mA_hi, in.h16, L.CinP, W, H, D, B, bw, bh, bd, P.TN, L.sw, L.sh, L.sd, true));
    SS2_TRY(make_act_map(ctx, &mA_lo, in.l16, L.CinP, W, H, D, B, bw, bh, bd, P.TN, L.sw, L.sh, L.sd, true));
    SS2_TRY(make_weight_map(ctx, &mB_hi, L.wk_h16, L.KD * L.KH * L.KW * L.CinP, L.CoutP, true));
    SS2_TRY(make_weight_map(ctx, &mB_lo, L.wk_l16, L.KD * L.KH * L.KW * L.CinP, L.CoutP, true));
  } else {
    SS2_TRY(make_act_map(ctx, &mA_hi, in.hi, L.CinP, W, H, D, B, bw, bh, bd, P.TN, L.sw, L.sh, L.sd));
    SS2_TRY(make_weight_map(ctx, &mB_hi, L.wk_hi, L.KD * L.KH * L.KW * L.CinP, L.CoutP));
    if (npass == 3) {
      SS2_TRY(make_act_map(ctx, &mA_lo, in.lo, L.CinP, W, H, D, B, bw, bh, bd, P.TN, L.sw, L.sh, L.sd));
      SS2_TRY(make_weight_map(ctx, &mB_lo, L.wk_lo, L.KD * L.KH * L.KW * L.CinP, L.CoutP));
    } else {
      mA_lo = mA_hi; mB_lo = mB_hi;
    }
  }
  dim3 grid(P.nW * P.nH * P.nT * P.nN, L.CoutP / TC_BN);
  const double flops = 2.0 * B * P.Do * P.Ho * P.Wo * (double)L.Cout * L.KD * L.KH * L.KW * L.Cin;
  ss2_prof_begin(ctx, SS2_PROF_CONV, st);
  if (f16) {
    constexpr int STAGES = TC_STAGES3;
    const size_t smem = (size_t)STAGES * 2 * (TC_A_BYTES + TC_B_BYTES) + 1024;
    static bool attr16_dev[16] = {false};
    bool& attr16 = attr16_dev[ctx->device & 15];
    if (!attr16) { SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<3, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr16 = true; }
    conv_tc_kernel<3, STAGES, true><<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  } else if (npass == 3) {
    constexpr int STAGES = TC_STAGES3;
    const size_t smem = (size_t)STAGES * 2 * (TC_A_BYTES + TC_B_BYTES) + 1024;
    static bool attr3_dev[16] = {false};  // per device: function attributes live in the device's context
    bool& attr3 = attr3_dev[ctx->device & 15];
    if (!attr3) { SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<3, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr3 = true; }
    conv_tc_kernel<3, STAGES><<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  } else {
    constexpr int STAGES = TC_STAGES1;
    const size_t smem = (size_t)STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024;
    static bool attr1_dev[16] = {false};
    bool& attr1 = attr1_dev[ctx->device & 15];
    if (!attr1) { SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<1, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr1 = true; }
    conv_tc_kernel<1, STAGES><<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  }
  ss2_prof_end(ctx, SS2_PROF_CONV, st, flops);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// 7x7 stride-2 stem (3 input channels) as an implicit GEMM with K = 7 filter rows x 32: the A row of output pixel
// (ho, wo) for filter row kh is the 32 contiguous floats of the padded NHWC4 input starting at pixel (2 ho + kh, 2 wo)
// (7 taps x 4 channels + one pixel that meets zero weights).  The tensor map describes exactly that overlapping view:
// {32 floats, Wo pixels 32 B apart, Hp rows, 1, B}; the row stride 2 is the box's traversal stride.
// ------------------------------------------------------------------------------------------
int conv_tc_stem_launch(ss2_ctx* ctx, const ConvLayer& L, const float* d_hi, const float* d_lo, int B, int H, int W,
                        int Hp, int Wp, const ActRef& out, int relu, cudaStream_t st) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  TcParams P;
  conv_out_dims(L, 1, H, W, &P.Do, &P.Ho, &P.Wo);
  P.B = B; P.Cout = L.Cout; P.CinP = 32;
  P.KD = 1; P.KH = 7; P.KW = 1; P.nchunk = 1;
  P.sd = 1; P.sh = 2; P.sw = 1; P.pd = 0; P.ph = 0; P.pw = 0;  // padding and the column stride live in the tensor map
  P.relu = relu; P.bias = L.bias; P.residual = nullptr;
  P.b_act = 0; P.bTH = P.bTW = 0; P.ldo = L.Cout; P.ncol = TC_BN;
  P.out_v = out.v; P.out_hi = out.hi; P.out_lo = out.lo;
  P.out_h16 = out.h16; P.out_l16 = out.h16 ? out.l16 : nullptr; P.range_flag = ctx->d_range_flag;
  tile_shape(B, 1, P.Ho, P.Wo, &P.TN, &P.TT, &P.TH, &P.TW);
  P.nW = cdiv(P.Wo, P.TW); P.nH = cdiv(P.Ho, P.TH); P.nT = 1; P.nN = cdiv(B, P.TN);
  const int npass = (ctx->tc_passes == 1 || !d_lo || !L.wk_lo) ? 1 : 3;
  const int bh = (P.TH - 1) * 2 + 1;
  if (bh > 256 || P.TW > 256) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc_stem: TMA box too large");
  CUtensorMap mA[2], mB[2];
  for (int pl = 0; pl < (npass == 3 ? 2 : 1); ++pl) {
    cuuint64_t dims[5] = {32, (cuuint64_t)P.Wo, (cuuint64_t)Hp, 1, (cuuint64_t)B};
    cuuint64_t strides[4] = {32, (cuuint64_t)Wp * 16, (cuuint64_t)Hp * Wp * 16, (cuuint64_t)Hp * Wp * 16};
    cuuint32_t box[5] = {32, (cuuint32_t)P.TW, (cuuint32_t)bh, 1, (cuuint32_t)P.TN};
    cuuint32_t est[5] = {1, 1, 2, 1, 1};
    CUresult r = fn(&mA[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(pl == 0 ? d_hi : d_lo), dims, strides, box,
                    est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ss2_fail(ctx, SS2_ERR_CUDA, "cuTensorMapEncodeTiled(stem input) = %d", (int)r);
    SS2_TRY(make_weight_map(ctx, &mB[pl], pl == 0 ? L.wk_hi : L.wk_lo, 7 * 32, L.CoutP));
  }
  if (npass != 3) { mA[1] = mA[0]; mB[1] = mB[0]; }
  dim3 grid(P.nW * P.nH * P.nT * P.nN, L.CoutP / TC_BN);
  const double flops = 2.0 * B * P.Ho * P.Wo * (double)L.Cout * 49 * 3;
  ss2_prof_begin(ctx, SS2_PROF_CONV, st);
  if (npass == 3) {
    constexpr int STAGES = TC_STAGES3;
    const size_t smem = (size_t)STAGES * 2 * (TC_A_BYTES + TC_B_BYTES) + 1024;
    SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<3, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<3, STAGES><<<grid, TC_THREADS, smem, st>>>(mA[0], mA[1], mB[0], mB[1], P);
  } else {
    constexpr int STAGES = TC_STAGES1;
    const size_t smem = (size_t)STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024;
    SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<1, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<1, STAGES><<<grid, TC_THREADS, smem, st>>>(mA[0], mA[1], mB[0], mB[1], P);
  }
  ss2_prof_end(ctx, SS2_PROF_CONV, st, flops);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// CCL correlation on the tensor cores: match[b][p][k] = sum_{tap,c} n1pad[b][p+tap][c] * n2pad[b][k+tap][c]
// (spatial_network.py:369-425).  Both operands are im2col views of activation tensors.
// match rows have stride ldo >= H*W floats.
// ------------------------------------------------------------------------------------------
// rows of the second feature map per N tile: the largest count with rows*W <= 64 and rows*W % 4 == 0
// (the epilogue stores float4 columns), 0 if there is none
int conv_tc_corr_rows(int W) {
  for (int r = TC_BN / W; r >= 1; --r)
    if ((r * W) % 4 == 0) return r;
  return 0;
}

int conv_tc_corr_launch(ss2_ctx* ctx, const ActRef& n1, const ActRef& n2, int B, int H, int W, int C, float* d_match,
                        int ldo, cudaStream_t st) {
  if ((C % TC_BK) != 0 || W > 64 || (ldo & 3)) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc_corr: unsupported shape");
  const bool f16 = n1.h16 != nullptr && n1.hi == nullptr;   // both maps as fp16 split planes
  if (f16 && (!n1.l16 || !n2.h16 || !n2.l16 || (C % 64) != 0 || ctx->tc_passes == 1))
    return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc_corr: incomplete fp16 planes");
  TcParams P;
  P.B = B; P.Do = 1; P.Ho = H; P.Wo = W; P.Cout = H * W; P.CinP = C;
  P.KD = 1; P.KH = 3; P.KW = 3; P.nchunk = C / (f16 ? 2 * TC_BK : TC_BK);
  P.sd = P.sh = P.sw = 1; P.pd = 0; P.ph = P.pw = 1;
  P.relu = 0; P.bias = nullptr; P.residual = nullptr;
  P.out_v = d_match; P.out_hi = nullptr; P.out_lo = nullptr;
  P.out_h16 = P.out_l16 = nullptr; P.range_flag = nullptr;
  tile_shape(1, 1, H, W, &P.TN, &P.TT, &P.TH, &P.TW);  // one sample per tile
  P.nW = cdiv(W, P.TW); P.nH = cdiv(H, P.TH); P.nT = 1; P.nN = B;
  P.b_act = 1; P.bTW = W; P.bTH = conv_tc_corr_rows(W);
  if (P.bTH < 1) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc_corr: no aligned N tile for W=%d", W);
  P.ncol = P.bTH * P.bTW; P.ldo = ldo;
  if (P.TW != W || P.nW != 1) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_tc_corr: feature map too wide");
  const int npass = f16 ? 3 : (ctx->tc_passes == 1 || !n1.lo || !n2.lo) ? 1 : 3;
  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
  if (f16) {
    SS2_TRY(make_act_map(ctx, &mA_hi, n1.h16, C, W, H, 1, B, P.TW, P.TH, 1, 1, 1, 1, 1, true));
    SS2_TRY(make_act_map(ctx, &mB_hi, n2.h16, C, W, H, 1, B, P.bTW, P.bTH, 1, 1, 1, 1, 1, true));
    SS2_TRY(make_act_map(ctx, &mA_lo, n1.l16, C, W, H, 1, B, P.TW, P.TH, 1, 1, 1, 1, 1, true));
    SS2_TRY(make_act_map(ctx, &mB_lo, n2.l16, C, W, H, 1, B, P.bTW, P.bTH, 1, 1, 1, 1, 1, true));
  } else {
    SS2_TRY(make_act_map(ctx, &mA_hi, n1.hi, C, W, H, 1, B, P.TW, P.TH, 1, 1, 1, 1, 1));
    SS2_TRY(make_act_map(ctx, &mB_hi, n2.hi, C, W, H, 1, B, P.bTW, P.bTH, 1, 1, 1, 1, 1));
    if (npass == 3) {
      SS2_TRY(make_act_map(ctx, &mA_lo, n1.lo, C, W, H, 1, B, P.TW, P.TH, 1, 1, 1, 1, 1));
      SS2_TRY(make_act_map(ctx, &mB_lo, n2.lo, C, W, H, 1, B, P.bTW, P.bTH, 1, 1, 1, 1, 1));
    } else {
      mA_lo = mA_hi; mB_lo = mB_hi;
    }
  }
  dim3 grid(P.nH * B, cdiv(H, P.bTH));
  ss2_prof_begin(ctx, SS2_PROF_CONV, st);
  if (f16) {
    constexpr int STAGES = TC_STAGES3;
    const size_t smem = (size_t)STAGES * 2 * (TC_A_BYTES + TC_B_BYTES) + 1024;
    SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<3, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<3, STAGES, true><<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  } else if (npass == 3) {
    constexpr int STAGES = TC_STAGES3;
    const size_t smem = (size_t)STAGES * 2 * (TC_A_BYTES + TC_B_BYTES) + 1024;
    SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<3, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<3, STAGES><<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  } else {
    constexpr int STAGES = TC_STAGES1;
    const size_t smem = (size_t)STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024;
    SS2_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<1, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_tc_kernel<1, STAGES><<<grid, TC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  }
  ss2_prof_end(ctx, SS2_PROF_CONV, st, 2.0 * B * (double)H * W * H * W * 9 * C);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
