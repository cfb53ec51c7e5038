// ResNet stem on the tensor cores as a DIRECT convolution with the max-pool fused into the epilogue (sm_100a):
// conv 7x7 stride 2 pad 3 (3 -> 64, BatchNorm folded) + ReLU + max-pool 3x3 stride 2 pad 1
// (reference: spatial_network.py:123-131 / temporal_network.py:65-73 `feature_extractor_stage1[0..3]`).
//
// Why: the implicit-GEMM stem (conv_tc_stem_launch) stages one TMA box per (filter row, 128-pixel tile): every input
// pixel crosses L2 -> shared memory 28 times and the kernel ran exactly at that rate (2.5 GB per 32-image launch,
// 493 us), wrote the 180x240x64 map with its TF32 split planes (1.06 GB) and a second kernel pooled it (146 us).
//
// Here the input is converted once into FOUR parity sub-images (row parity x column parity of the zero-padded
// image), NHWC4, split into hi/lo TF32 planes.  In a sub-image the stride-2 window of an output pixel slides by ONE
// pixel = 16 bytes per output pixel, which is exactly the row pitch of the un-swizzled K-major UMMA layout (core
// matrix = 8 rows x 16 B): with SBO = 128 B and LBO = 16 B the descriptor reads row m, k at 16 (m + k / 4) + 4 (k % 4)
// bytes, i.e. pixels m + j, m + j + 1 of one staged sub-row - an overlapping ("Toeplitz") A operand straight from the
// staged image row, no im2col copy anywhere.  The 49 filter taps are 25 k-steps of K = 8 (two pixels): per filter row
// even columns (0,2), (4,6) and odd columns (1,3); the left-over column 5 pairs up across two filter rows (LBO = the
// distance of their sub-rows in the staged row pair).
//   * a tile is one conv row of one image half (M = 128 pixels, N = 64 channels); a persistent CTA walks down a band of
//     conv rows, so each input row pair is staged once per band (ring of row pairs, one 16.5 KB TMA box each);
//   * the whole filter (7 rows x [64][32], hi/lo: 112 KB) stays resident in shared memory;
//   * split TF32 in two MMAs per k-step (A_hi x [B_hi | B_lo] with N = 128 into [D1 | D2], A_lo x B_hi with N = 64 into
//     D1; the epilogue adds D1 + D2) and four TMEM accumulator stages: MMAs of later rows overlap the epilogue;
//   * epilogue: thread = one pixel column with 64 channels; the vertical 3-max lives in registers across the rows of
//     the band, the horizontal 3-max takes the neighbours with warp shuffles (+ a 64-float hand-over between warps),
//     then bias + ReLU (monotone, so they commute with the max), the TF32 split, and the pooled pixel is stored with
//     its three planes.  The 180x240 map never exists in memory.
#include "tc_common.cuh"

#define ST_THREADS 192
#define ST_PX 132                         // staged pixels per sub-row: 128 + 3 + 1
#define ST_SEG (ST_PX * 16)               // bytes of one staged sub-row (one plane, one row parity, one column parity)
#define ST_SLOT (8 * ST_SEG)              // one row pair: [plane 2][row parity 2][column parity 2] sub-rows
#define ST_NA 6                           // row pairs in the ring (4 in use + 2 in flight)
#define ST_NACC 4                         // TMEM accumulator stages
#define ST_BCHUNK (64 * 128)              // one filter row of one plane: [64 cout][32 k] SWIZZLE_128B
#define ST_BPLANE (7 * ST_BCHUNK)
#define ST_WS 256                         // sub-columns per sub-row in global memory

struct StemParams {
  const float* bias;
  float* out_v;
  float* out_hi;
  float* out_lo;
  __half* out_h16;   // fp16 split planes for a kind::f16 consumer (conv_dc.cu; may be null)
  __half* out_l16;
  int* range_flag;     // raised when a value does not fit fp16 (common.cuh)
  int B, Hc, PH, PW;    // images, conv rows, pooled rows / cols
  int NQ;               // row pairs per image = Hc + 3
  int bp, nbands;       // pooled rows per band, bands per image half
  int units;            // B * 2 * nbands
  int npass;
};

// NCHW [B,3,H,W] -> D[b][q][plane*4 + rp*2 + cp][xs][4]: D = padded(2q + rp, 2xs + cp), padded = image shifted by the
// conv padding 3 with zeros around, 4th channel zero; plane 0 = rna_tf32(v), plane 1 = v - plane 0
__global__ void stem_parity_split_kernel(const float* __restrict__ in, int B, int H, int W, int NQ, float4* __restrict__ out) {
  const size_t total = (size_t)B * NQ * 4 * ST_WS, HW = (size_t)H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xs = (int)(i % ST_WS), par = (int)((i / ST_WS) & 3);
    const size_t bq = i / (4 * ST_WS);
    const int q = (int)(bq % NQ);
    const size_t b = bq / NQ;
    const int y = 2 * q + (par >> 1) - 3, x = 2 * xs + (par & 1) - 3;
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const float* src = in + b * 3 * HW + (size_t)y * W + x;
      tf32_split(__ldg(src), &h.x, &l.x);
      tf32_split(__ldg(src + HW), &h.y, &l.y);
      tf32_split(__ldg(src + 2 * HW), &h.z, &l.z);
    }
    const size_t o = (bq * 8 + par) * ST_WS + xs;
    out[o] = h;
    out[o + (size_t)4 * ST_WS] = l;
  }
}

__device__ __forceinline__ void st_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  // A: K-major, no swizzle, LBO 16 B (in a_lo), SBO 128 B; B: K-major SWIZZLE_128B, SBO 1024 B (see tc_common.cuh)
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"((uint32_t)(128 >> 4) | (1u << 14)), "r"(DC_DESC_HI), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// unit u -> image, half, pooled-row band -> conv rows [oy0, oy1]
struct StemUnit { int n, half, p0, p1, oy0, oy1; };
__device__ __forceinline__ StemUnit stem_unit(const StemParams& P, int u) {
  StemUnit s;
  const int band = u % P.nbands, nh = u / P.nbands;
  s.half = nh & 1; s.n = nh >> 1;
  s.p0 = band * P.bp;
  s.p1 = min(s.p0 + P.bp, P.PH);
  s.oy0 = s.p0 > 0 ? 2 * s.p0 - 1 : 0;
  s.oy1 = 2 * s.p1 - 1;
  return s;
}

__global__ void __launch_bounds__(ST_THREADS, 1)
conv_stem_pool_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB_hi,
                      const __grid_constant__ CUtensorMap tmB_lo, StemParams P) {
  extern __shared__ __align__(1024) uint8_t st_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t a_full[ST_NA], a_empty[ST_NA], b_full, acc_full[ST_NACC], acc_empty[ST_NACC];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[64];
  __shared__ __align__(16) float xchg[2][4][64];
  __shared__ float4 pstage[4][16 * 4];      // per epilogue warp: 16 pooled pixels x 16 channels, swizzled
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* b_smem = smem;                       // [kh][plane][64][32] SWIZZLE_128B: hi | lo of a filter row are ONE N = 128 operand
  uint8_t* a_ring = smem + 2 * ST_BPLANE;       // ST_NA row-pair slots
  const int nplanes = P.npass == 3 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST_NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(&b_full, 1);
    for (int s = 0; s < ST_NACC; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = P.bias ? P.bias[threadIdx.x] : 0.f;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(ST_NACC * 128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer: the filter once, then the row pairs of every band =====
    if (dc_elect_one()) {
      mbar_expect_tx(&b_full, (uint32_t)(nplanes * ST_BPLANE));
      for (int kh = 0; kh < 7; ++kh) {
        tma_load_2d(b_smem + 2 * kh * ST_BCHUNK, &tmB_hi, &b_full, kh * 32, 0);
        if (nplanes == 2) tma_load_2d(b_smem + (2 * kh + 1) * ST_BCHUNK, &tmB_lo, &b_full, kh * 32, 0);
      }
      int ia = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
        const StemUnit s = stem_unit(P, u);
        const int ox0 = s.half ? P.PW - 1 : 0;
        for (int q = s.oy0; q <= s.oy1 + 3; ++q, ++ia) {
          const int sl = ia % ST_NA;
          if (ia >= ST_NA) mbar_wait(&a_empty[sl], ((ia / ST_NA) - 1) & 1);
          mbar_expect_tx(&a_full[sl], ST_SLOT);
          tma_load_3d(a_ring + (size_t)sl * ST_SLOT, &tmA, &a_full[sl], 0, ox0, (s.n * P.NQ + q) * 8);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (dc_elect_one()) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // split TF32 in two MMAs per k-step: A_hi x [B_hi | B_lo] (N = 128) into [D1 | D2], A_lo x B_hi (N = 64) into D1
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t accw = nplanes == 2 ? 128u : 64u;
      const uint32_t b_lo0 = dc_desc_lo(smem_u32(b_smem));
      const uint32_t a_base = smem_u32(a_ring);
      mbar_wait(&b_full, 0);
      int ia0 = 0, waited = 0, it = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
        const StemUnit s = stem_unit(P, u);
        const int npairs = s.oy1 + 3 - s.oy0 + 1;
        for (int oy = s.oy0; oy <= s.oy1; ++oy, ++it) {
          const int as = it % ST_NACC;
          if (it >= ST_NACC) mbar_wait(&acc_empty[as], ((it / ST_NACC) - 1) & 1);
          const int g0 = ia0 + (oy - s.oy0);                 // ring index of row pair oy
          for (; waited <= g0 + 3; ++waited) mbar_wait(&a_full[waited % ST_NA], (waited / ST_NA) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + (uint32_t)as * accw;
          // 25 k-steps of K = 8 (two pixels).  Steps 3 kh + {0, 1, 2}: pixels (0, 1), (2, 3) of the even-column sub-row and
          // (0, 1) of the odd-column sub-row of filter row kh, the two pixels 16 bytes apart (LBO).  The seven left-over
          // odd-column pixels (filter column 5) pair up ACROSS filter rows: rows 2t and 2t + 1 sit in the same row-pair
          // slot, two sub-rows apart, so LBO = 2 sub-rows makes them one k-step (the last one pairs with a zero weight).
#pragma unroll
          for (int step = 0; step < 25; ++step) {
            int kh, seg, px;
            uint32_t lbo = 1u;                                   // 16 B
            if (step < 21) { kh = step / 3; seg = (kh & 1) * 2 + (step % 3 == 2 ? 1 : 0); px = (step % 3 == 1) ? 2 : 0; }
            else if (step < 24) { kh = 2 * (step - 21); seg = 1; px = 2; lbo = (uint32_t)(2 * ST_SEG >> 4); }
            else { kh = 6; seg = 1; px = 2; }
            const int sl = (g0 + (kh >> 1)) % ST_NA;
            const uint32_t da = (((a_base + (uint32_t)sl * ST_SLOT + (uint32_t)seg * ST_SEG + (uint32_t)px * 16u) & 0x3FFFFu) >> 4) | (lbo << 16);
            const uint32_t db = b_lo0 + (uint32_t)(2 * (step >> 2) * ST_BCHUNK >> 4) + (uint32_t)((step & 3) * 2);
            if (nplanes == 2) {
              st_mma(d_tmem, da, db, idesc2, step ? 1u : 0u);
              st_mma(d_tmem, da + (uint32_t)(4 * ST_SEG >> 4), db, idesc, 1u);
            } else {
              st_mma(d_tmem, da, db, idesc, step ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[g0 % ST_NA]);     // row pair oy is not read by later rows
          umma_commit(&acc_full[as]);
        }
        for (int j = 1; j <= 3; ++j) umma_commit(&a_empty[(ia0 + (s.oy1 - s.oy0) + j) % ST_NA]);   // tail pairs of the band
        ia0 += npairs;
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter q = warp % 4 =====
    const int q = warp & 3;
    const int L = q * 32 + lane;                   // pixel column within the tile
    int it = 0, emits = 0;
    float acc[64];
    for (int u = blockIdx.x; u < P.units; u += gridDim.x) {
      const StemUnit s = stem_unit(P, u);
      const int ox0 = s.half ? P.PW - 1 : 0;
      const int x = ox0 + L;
      const bool centre = !(x & 1);
      for (int oy = s.oy0; oy <= s.oy1; ++oy, ++it) {
        const int as = it % ST_NACC;
        mbar_wait(&acc_full[as], (it / ST_NACC) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t c0[32], c1[32];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * (nplanes == 2 ? 128u : 64u);
        tmem_ld32(ta, c0);
        tmem_ld32(ta + 32, c1);
        if (nplanes == 2) {   // D1 + D2
          uint32_t e0[32];
          tmem_ld32(ta + 64, e0);
#pragma unroll
          for (int j = 0; j < 32; ++j) c0[j] = __float_as_uint(__uint_as_float(c0[j]) + __uint_as_float(e0[j]));
          tmem_ld32(ta + 96, e0);
#pragma unroll
          for (int j = 0; j < 32; ++j) c1[j] = __float_as_uint(__uint_as_float(c1[j]) + __uint_as_float(e0[j]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
        if (oy == s.oy0) {
          // first row of the band: the row above pooled row p0 (carried, nothing emitted) or image row 0
#pragma unroll
          for (int j = 0; j < 32; ++j) { acc[j] = __uint_as_float(c0[j]); acc[32 + j] = __uint_as_float(c1[j]); }
          if (oy & 1) continue;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc[j] = fmaxf(acc[j], __uint_as_float(c0[j]));
            acc[32 + j] = fmaxf(acc[32 + j], __uint_as_float(c1[j]));
          }
        }
        if (!(oy & 1)) continue;
        // ---- pooled row pr = (oy - 1) / 2 is complete in the vertical direction: horizontal 3-max, bias, ReLU, store ----
        const int pr = (oy - 1) >> 1;
        float* xb = &xchg[emits & 1][0][0];
        ++emits;
        // centres on even lanes (left half): lane 0 needs the last pixel of the previous warp; centres on odd lanes
        // (right half): lane 31 needs the first pixel of the next warp
        if (lane == (s.half ? 0 : 31)) {
#pragma unroll
          for (int j = 0; j < 64; j += 4)
            *reinterpret_cast<float4*>(xb + q * 64 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool take_l = !s.half && lane == 0, take_r = s.half && lane == 31;
        const float* nb = xb + (take_l ? (q > 0 ? q - 1 : 0) : (q < 3 ? q + 1 : 3)) * 64;
        const bool edge_l = take_l && q == 0;          // x = -1: outside the image
        // store mapping: the 16 pooled pixels of this warp x 16 channels go through a swizzled 1 KB buffer, then lane l
        // owns 16 bytes (l & 3) of pooled pixels (l >> 2) + 8 i: 8 lines per store instruction instead of 16 half-used ones
        size_t so[2];
        bool sv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int pcc = (ox0 + q * 32 + 2 * ((lane >> 2) + 8 * i) + s.half) >> 1;
          sv[i] = s.half ? (pcc >= P.PW / 2 && pcc < P.PW) : (pcc < P.PW / 2);
          so[i] = (((size_t)s.n * P.PH + pr) * P.PW + pcc) * 64;
        }
        float4* pst = pstage[q];
        const int pp = lane >> 1;                      // pooled pixel of a centre lane within the warp
#pragma unroll
        for (int j = 0; j < 64; j += 16) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float m[4];
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (take_l || take_r) e = *reinterpret_cast<const float4*>(nb + j + jj * 4);
            const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a = acc[j + jj * 4 + k];
              float l = __shfl_up_sync(0xffffffffu, a, 1);
              float r = __shfl_down_sync(0xffffffffu, a, 1);
              if (take_l) l = edge_l ? a : ev[k];
              if (take_r) r = ev[k];
              m[k] = fmaxf(fmaxf(l, r), a);
            }
            if (centre) pst[pp * 4 + (jj ^ ((pp >> 1) & 3))] = make_float4(m[0], m[1], m[2], m[3]);
          }
          __syncwarp();
          const int c = j + (lane & 3) * 4;
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + c);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int row = (lane >> 2) + 8 * i;
            const float4 a = pst[row * 4 + ((lane & 3) ^ ((row >> 1) & 3))];
            if (sv[i]) {
              const size_t o = so[i] + c;
              float v[4] = {fmaxf(a.x + bb.x, 0.f), fmaxf(a.y + bb.y, 0.f), fmaxf(a.z + bb.z, 0.f), fmaxf(a.w + bb.w, 0.f)};
              *reinterpret_cast<float4*>(P.out_v + o) = make_float4(v[0], v[1], v[2], v[3]);
              if (P.out_hi) {
                float hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { hi[k] = rna_tf32(v[k]); lo[k] = v[k] - hi[k]; }
                *reinterpret_cast<float4*>(P.out_hi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                if (P.out_lo) *reinterpret_cast<float4*>(P.out_lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
              }
              if (P.out_h16) store_f16_planes4(P.out_h16, P.out_l16, o, v, P.range_flag);
            }
          }
          __syncwarp();
        }
        // the carried row for the next pooled row is this odd conv row itself
#pragma unroll
        for (int j = 0; j < 32; ++j) { acc[j] = __uint_as_float(c0[j]); acc[32 + j] = __uint_as_float(c1[j]); }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(ST_NACC * 128));
  }
}

// the direct kernel is built for the networks' 360x480 input (conv map 180x240, pooled 90x120): two 128-pixel tiles
// cover a conv row, H a multiple of 4 keeps conv and pooled rows paired
bool conv_stem_direct_eligible(const ConvLayer& L, int H, int W) {
  return L.ws_hi != nullptr && L.ws_lo != nullptr && L.Cout == 64 && L.CoutP == 64 && W == 480 && H >= 8 && (H % 4) == 0;
}

size_t conv_stem_direct_workspace_floats(int B, int H) { return (size_t)B * (H / 2 + 3) * 8 * ST_WS * 4; }

// x NCHW [B,3,H,W] -> out [B,H/4,W/4,64] (+ split planes when out.hi/lo are set); d_work: conv_stem_direct_workspace_floats
int conv_stem_pool_launch(ss2_ctx* ctx, const ConvLayer& L, const float* d_x_nchw, int B, int H, int W, float* d_work,
                          const ActRef& out, cudaStream_t st) {
  if (B <= 0) return SS2_OK;
  if (!conv_stem_direct_eligible(L, H, W)) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_stem_pool: shape not eligible");
  StemParams P;
  P.bias = L.bias;
  P.out_v = out.v; P.out_hi = out.hi; P.out_lo = out.lo;
  P.out_h16 = out.h16; P.out_l16 = out.h16 ? out.l16 : nullptr; P.range_flag = ctx->d_range_flag;
  P.B = B; P.Hc = H / 2; P.PH = H / 4; P.PW = W / 4;
  P.NQ = P.Hc + 3;
  P.npass = ctx->tc_passes == 1 ? 1 : 3;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
  // pooled rows per band: the band size whose slowest CTA has the fewest conv rows (a band of bp pooled rows costs
  // 2 bp + 1 conv rows: one row above the band is recomputed)
  long best = -1;
  P.bp = P.PH;
  for (int bp = 3; bp <= P.PH; ++bp) {
    const int nb = cdiv(P.PH, bp);
    const long units = (long)B * 2 * nb;
    const long cost = cdiv(units, (long)nsm) * (2 * bp + 1);
    if (best < 0 || cost < best) { best = cost; P.bp = bp; }
  }
  P.nbands = cdiv(P.PH, P.bp);
  P.units = B * 2 * P.nbands;
  {
    const size_t total = (size_t)B * P.NQ * 4 * ST_WS;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)nsm * 16);
    stem_parity_split_kernel<<<blocks, 256, 0, st>>>(d_x_nchw, B, H, W, P.NQ, reinterpret_cast<float4*>(d_work));
    SS2_LAUNCH_CHECK(ctx);
  }
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ss2_tensormap_encode_fn());
  if (!fn) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  CUtensorMap mA, mB_hi, mB_lo;
  {
    cuuint64_t dims[3] = {4, ST_WS, (cuuint64_t)B * P.NQ * 8};
    cuuint64_t strides[2] = {16, (cuuint64_t)ST_WS * 16};
    cuuint32_t box[3] = {4, ST_PX, 8};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = fn(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_work, dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ss2_fail(ctx, SS2_ERR_CUDA, "cuTensorMapEncodeTiled(stem parity planes) = %d", (int)r);
  }
  SS2_TRY(make_weight_map(ctx, &mB_hi, L.ws_hi, 7 * 32, 64));
  SS2_TRY(make_weight_map(ctx, &mB_lo, L.ws_lo, 7 * 32, 64));
  const size_t smem = (size_t)2 * ST_BPLANE + (size_t)ST_NA * ST_SLOT + 1024;
  static bool attr_dev[16] = {false};
  if (!attr_dev[ctx->device & 15]) {
    SS2_CUDA(ctx, cudaFuncSetAttribute(conv_stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev[ctx->device & 15] = true;
  }
  const int grid = P.units < nsm ? P.units : nsm;
  const double flops = 2.0 * B * P.Hc * (W / 2) * 64.0 * 49 * 3;
  ss2_prof_begin(ctx, SS2_PROF_CONV, st);
  conv_stem_pool_kernel<<<grid, ST_THREADS, smem, st>>>(mA, mB_hi, mB_lo, P);
  ss2_prof_end(ctx, SS2_PROF_CONV, st, flops);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
