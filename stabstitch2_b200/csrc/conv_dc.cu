// tcgen05 DIRECT convolution for the stride-1 3x3 layers (sm_100a): the nine filter taps read ONE
// staged input tile at nine row offsets instead of nine im2col copies.
//
// conv_tc.cu's implicit GEMM moves, per 128-row output tile, one TMA box per (tap, 32-channel chunk): the same
// input pixels cross L2 -> shared memory nine times, and on B200 the ResNet layers ran exactly at the L2 -> SM
// bandwidth (layer1: 2.5 GB per launch at ~8 TB/s = the measured 315 us).  Here:
//   * the output tile is TH full image rows and the GEMM row index is m = h * P + w', P = W + 2, i.e. the two
//     halo columns are carried as junk rows (never stored).  The staged input tile is the (TH + 2) x P pixels
//     (h0 - 1 .., -1 ..) of one 32-channel chunk, ONE TMA box with zero fill = the padding, landing as P * (TH + 2)
//     K-major 128-byte rows (SWIZZLE_128B).  For tap (kh, kw) the A operand of output row m is staged row
//     m + kh * P + kw: a constant offset, so the tap's UMMA descriptor is the tile's descriptor advanced by
//     (kh * P + kw) * 128 bytes (the swizzle is a function of the absolute shared-memory address for both the
//     TMA write and the MMA read, so any 128-byte row is a valid start).  A bytes per tile drop 9x (per chunk: one
//     box instead of nine).
//   * weights stream through their own ring (one [64 cout][32 k] hi/lo stage per (chunk, tap)).
//   * persistent CTAs (one per SM) with two TMEM accumulators: the epilogue of tile i overlaps the MMAs of
//     tile i + 1.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue.
// Numerics are those of conv_tc.cu: three exact products of 11-bit split operands per fp32 product, fp32 accumulation in
// TMEM; the operands are TF32 split planes (kind::tf32, K = 8 per MMA) or, template flag F16 - the default behind the stem -
// fp16 split planes with a 2^11-scaled low half (kind::f16, K = 16 per MMA: half the chunks, stages and instructions).
#include "tc_common.cuh"

#define DC_THREADS 192
#define DC_BK 32
#define DC_BOX_N 64                      // rows of one weight TMA box

struct DcParams {
  const float* bias;
  const float* residual;
  float* out_v;
  float* out_hi;
  float* out_lo;
  __half* out_h16;     // fp16 split planes for a kind::f16 consumer (may be null)
  __half* out_l16;
  int* range_flag;     // raised when a value does not fit fp16 (common.cuh)
  int B, H, W, Cout;
  int P, TH, TW;       // padded pitch TW + 2, output rows / columns per tile
  int tiles_h, tiles_w;  // ceil(H / TH), ceil(W / TW)
  int n_mtiles, n_ntiles;
  int nchunk, CinP;
  int a_rows;          // staged rows per plane (allocation; >= 128 + 2 P + 2, multiple of 8)
  int NA, NB;          // ring depths
  int relu;
  int npass;           // 3 = split TF32, 1 = plain TF32
  int dbg;             // timing experiments (SS2_DC_DBG): 1 = no TMA traffic (MMAs on stale smem), 2 = no MMAs, 4 = no stores
};

// K-major SWIZZLE_128B descriptor whose start sits on ANY 128-byte row of the 1024-byte swizzle atom.  Measured on
// B200 (tests/probe_dc.py): the tensor core applies the 128B swizzle to the absolute shared-memory address, exactly
// like the TMA unit that wrote the tile, so the descriptor's base-offset field stays 0 (filling it with the row
// phase of the start address gives wrong results for every tap whose offset is not a multiple of 8 rows).
__device__ __forceinline__ uint64_t dc_desc_sw128(uint32_t smem_addr) { return umma_desc_sw128(smem_addr); }

// tile index -> image, first row / column, N tile
__device__ __forceinline__ void dc_tile(const DcParams& P, int t, int* n, int* h0, int* w0, int* nt) {
  const int mt = t / P.n_ntiles;
  *nt = t - mt * P.n_ntiles;
  const int tw = mt % P.tiles_w, nh = mt / P.tiles_w;
  *n = nh / P.tiles_h;
  *h0 = (nh - *n * P.tiles_h) * P.TH;
  *w0 = tw * P.TW;
}
// epilogue mapping (see the kernel): lane l stores 16 bytes (l & 3) of rows (l >> 2) + 8 i of its warp's 32 rows
__device__ __forceinline__ bool dc_row(const DcParams& P, int n, int h0, int w0, int r, size_t* off) {
  const int hl = r / P.P, wl = r - hl * P.P;
  const int oh = h0 + hl, ow = w0 + wl;
  *off = (((size_t)n * P.H + oh) * P.W + ow) * P.Cout;
  return hl < P.TH && wl < P.TW && ow < P.W && oh < P.H;
}
// residual values of one epilogue step (128-row block `blk`, 32-column half `half`) of tile t, in the store mapping
template <int BN>
__device__ __forceinline__ void dc_residual_issue(const DcParams& P, int t, int blk, int half, int q, int lane, float4 (&rs)[8]) {
  int n, h0, w0, nt;
  dc_tile(P, t, &n, &h0, &w0, &nt);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    size_t off;
    const bool ok = dc_row(P, n, h0, w0, blk * 128 + q * 32 + (lane >> 2) + 8 * i, &off);
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int c = nt * BN + half * 32 + sub * 16 + (lane & 3) * 4;
      rs[sub * 4 + i] = (ok && c < P.Cout) ? __ldg(reinterpret_cast<const float4*>(P.residual + off + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// MT = 128-row GEMM blocks per tile (1 or 2).  Measured on B200 the kernel is bound by the bytes the TMA unit can
// deliver into one SM (~22 B/clk: 110-130 cycles per MMA for N = 64 and N = 128 alike, tensor pipe ~30% active), not
// by MMA issue or accumulator dependencies, so a tile of two blocks (2 TH image rows) shares every weight stage between
// two MMAs and its TH + 2 input rows between 2 TH output rows: ~45% fewer staged bytes per output row.
// F16: the operands are fp16 split planes (h16 = fp16(v), l16 = fp16((v - h16) * 2048), common.cuh) and the MMAs are
// kind::f16: K = 16 per instruction at twice the TF32 rate, so a 128-byte K-major row is 64 channels and a layer has half
// the chunks, stages and MMA instructions.  The same three exact products: A_h x [B_h | B_l] lands as [D1 | D2] with D2
// carrying the factor 2048 of B_l, A_l x B_h (factor 2048 of A_l) accumulates into D2 as well, the epilogue takes
// D1 + D2 / 2048.  Everything else (tile, rings, barriers, byte counts) is unchanged.
template <int BN, int MT, bool F16 = false>
__global__ void __launch_bounds__(DC_THREADS, 1)
conv_dc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, DcParams P) {
  constexpr int DC_BN = BN;
  constexpr int DC_B_BYTES = BN * DC_BK * 4;  // per plane
  constexpr int BKC = F16 ? 2 * DC_BK : DC_BK;   // channels per chunk (one 128-byte row)
  // Split TF32 as TWO MMAs per k-step instead of three: the weight planes of a stage are adjacent in shared memory,
  // so A_hi x [B_hi | B_lo] is ONE MMA with N = 2 BN into accumulator columns [D1 | D2], and A_lo x B_hi a second one
  // with N = BN into D1; the epilogue adds D1 + D2.  A tcgen05 TF32 MMA of 128 x N x 8 takes ~43 + N / 2 cycles on
  // B200 (operand reads), so 2 BN + BN columns in two instructions cost 182 cycles against 225 for three (BN = 64).
  constexpr int DC_TMEM_COLS = 2 * MT * 2 * BN > 512 ? 512 : 2 * MT * 2 * BN;
  extern __shared__ __align__(1024) uint8_t dc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dc_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t a_full[4], a_empty[4], b_full[8], b_empty[8], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 ep_stage[4][32 * 4];   // per epilogue warp: 32 rows x 16 columns, 16-byte slots XOR-swizzled
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nplanes = P.npass == 3 ? 2 : 1;
  const uint32_t a_plane = (uint32_t)P.a_rows * 128u;
  const uint32_t a_stage = a_plane * nplanes;
  const uint32_t b_stage = (uint32_t)DC_B_BYTES * nplanes;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)P.NA * a_stage;
  const int total = P.n_mtiles * P.n_ntiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < P.NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(DC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (dc_elect_one() && !(P.dbg & 1)) {
      const uint32_t a_bytes = (uint32_t)nplanes * (uint32_t)(P.P * (P.TH + 2)) * 128u;
      int ia = 0, ib = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int mt = t / P.n_ntiles, nt = t - mt * P.n_ntiles;
        const int tw = mt % P.tiles_w, nh = mt / P.tiles_w;
        const int n = nh / P.tiles_h, h0 = (nh - n * P.tiles_h) * P.TH, w0 = tw * P.TW;
        for (int ck = 0; ck < P.nchunk; ++ck) {
          const int sa = ia % P.NA;
          if (ia >= P.NA) mbar_wait(&a_empty[sa], ((ia / P.NA) - 1) & 1);
          uint8_t* ad = a_ring + (size_t)sa * a_stage;
          mbar_expect_tx(&a_full[sa], a_bytes);
          tma_load_5d(ad, &tmA_hi, &a_full[sa], ck * BKC, w0 - 1, h0 - 1, 0, n);
          if (nplanes == 2) tma_load_5d(ad + a_plane, &tmA_lo, &a_full[sa], ck * BKC, w0 - 1, h0 - 1, 0, n);
          ++ia;
          for (int tap = 0; tap < 9; ++tap) {
            const int sb = ib % P.NB;
            if (ib >= P.NB) mbar_wait(&b_empty[sb], ((ib / P.NB) - 1) & 1);
            uint8_t* bd = b_ring + (size_t)sb * b_stage;
            mbar_expect_tx(&b_full[sb], b_stage);
#pragma unroll
            for (int j = 0; j < BN / DC_BOX_N; ++j) {
              tma_load_2d(bd + j * (DC_BOX_N * DC_BK * 4), &tmB_hi, &b_full[sb], tap * P.CinP + ck * BKC, nt * DC_BN + j * DC_BOX_N);
              if (nplanes == 2)
                tma_load_2d(bd + DC_B_BYTES + j * (DC_BOX_N * DC_BK * 4), &tmB_lo, &b_full[sb], tap * P.CinP + ck * BKC,
                            nt * DC_BN + j * DC_BOX_N);
            }
            ++ib;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (dc_elect_one()) {
      // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
      // (kind::f16: A/B format 0 = fp16)
      constexpr uint32_t FMT = F16 ? 0u : ((2u << 7) | (2u << 10));
      const uint32_t idesc = (1u << 4) | FMT | ((uint32_t)(DC_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | FMT | ((uint32_t)(2 * DC_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_plane16 = a_plane >> 4;
      const uint32_t accw = (uint32_t)(nplanes == 2 ? 2 * DC_BN : DC_BN);   // accumulator columns per 128-row block
      int ia = 0, ib = 0, it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int as = it & 1;
        if (it >= 2) mbar_wait(&acc_empty[as], ((it >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)as * MT * accw;
        for (int ck = 0; ck < P.nchunk; ++ck) {
          const int sa = ia % P.NA;
          if (!(P.dbg & 1)) mbar_wait(&a_full[sa], (ia / P.NA) & 1);
          const uint32_t a_lo0 = dc_desc_lo(smem_u32(a_ring + (size_t)sa * a_stage));
          uint32_t tap_off = 0;  // (kh * P + kw) * 128 / 16
          for (int tap = 0; tap < 9; ++tap) {
            const int sb = ib % P.NB;
            if (!(P.dbg & 1)) mbar_wait(&b_full[sb], (ib / P.NB) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t da = a_lo0 + tap_off;
            const uint32_t db = dc_desc_lo(smem_u32(b_ring + (size_t)sb * b_stage));
            const uint32_t first = (ck > 0 || tap > 0) ? 1u : 0u;  // accumulate flag of the first MMA into each accumulator
            constexpr uint32_t BLK = 128 * 128 / 16;                   // descriptor units between the tile's row blocks
            if (P.dbg & 2) {
            } else if (nplanes == 2) {
              const uint32_t dal = da + a_plane16;
#pragma unroll
              for (int k = 0; k < DC_BK / 8; ++k) {
#pragma unroll
                for (int b = 0; b < MT; ++b) {
                  if (F16) dc_mma_f16(d_tmem + b * accw, da + b * BLK + 2 * k, db + 2 * k, DC_DESC_HI, idesc2, k == 0 ? first : 1u);
                  else dc_mma(d_tmem + b * accw, da + b * BLK + 2 * k, db + 2 * k, DC_DESC_HI, idesc2, k == 0 ? first : 1u);
                }
#pragma unroll
                for (int b = 0; b < MT; ++b) {
                  if (F16) dc_mma_f16(d_tmem + b * accw + DC_BN, dal + b * BLK + 2 * k, db + 2 * k, DC_DESC_HI, idesc, 1u);   // scaled: into D2
                  else dc_mma(d_tmem + b * accw, dal + b * BLK + 2 * k, db + 2 * k, DC_DESC_HI, idesc, 1u);
                }
              }
            } else {
#pragma unroll
              for (int k = 0; k < DC_BK / 8; ++k)
#pragma unroll
                for (int b = 0; b < MT; ++b) dc_mma(d_tmem + b * accw, da + b * BLK + 2 * k, db + 2 * k, DC_DESC_HI, idesc, k == 0 ? first : 1u);
            }
            umma_commit(&b_empty[sb]);
            ++ib;
            tap_off += (tap % 3 == 2) ? (uint32_t)(P.P - 2) * 8u : 8u;
          }
          umma_commit(&a_empty[sa]);
          ++ia;
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter q = warp % 4 =====
    // Stores go out in a COALESCED mapping: the accumulator arrives one row per thread (tcgen05.ld 32x32b), and a
    // thread-per-row float4 store touches 32 different 128-byte lines per instruction - 32 wavefronts of the L1 /
    // shared-memory datapath that the MMA operand reads also run on (measured: + 33 us for the two split planes on a
    // 162 us kernel).  16 columns at a time are transposed through a swizzled 2 KB buffer; then lane l owns 16 bytes
    // (l & 3) of rows (l >> 2) + 8 i: 8 lines per instruction.  The residual of a step (one 128-row block x 32 columns)
    // is loaded one step AHEAD - for the first step of a tile before the wait for its accumulator - because a load
    // issued where it is used exposes a DRAM round trip per step (32 per tile: + 40-57 us per launch).
    const int q = warp & 3;
    constexpr int NH = DC_BN / 32, NSTEP = MT * NH;
    float4* stg = ep_stage[q];
    float4 rs[8];
    if (P.residual && (int)blockIdx.x < total) dc_residual_issue<BN>(P, blockIdx.x, 0, 0, q, lane, rs);
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int as = it & 1;
      int n, h0, w0, nt;
      dc_tile(P, t, &n, &h0, &w0, &nt);
      mbar_wait(&acc_full[as], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int cout0 = nt * DC_BN;
#pragma unroll
      for (int step = 0; step < NSTEP; ++step) {
        const int blk = step / NH, half = step % NH;
        float4 rc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) rc[e] = rs[e];
        if (P.residual) {   // next step's residual: this tile's next step, or the first step of this CTA's next tile
          if (step + 1 < NSTEP) dc_residual_issue<BN>(P, t, (step + 1) / NH, (step + 1) % NH, q, lane, rs);
          else if (t + (int)gridDim.x < total) dc_residual_issue<BN>(P, t + gridDim.x, 0, 0, q, lane, rs);
        }
        size_t mo[4];
        bool vv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) vv[i] = dc_row(P, n, h0, w0, blk * 128 + q * 32 + (lane >> 2) + 8 * i, &mo[i]) && !(P.dbg & 4);
        uint32_t acc[32];
        const uint32_t accw = (uint32_t)(nplanes == 2 ? 2 * DC_BN : DC_BN);
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MT + blk) * accw + (uint32_t)(half * 32);
        tmem_ld32(tcol, acc);  // warp-collective
        if (nplanes == 2) {    // D1 + D2 (see the MMA issuer)
          uint32_t acc2[32];
          tmem_ld32(tcol + DC_BN, acc2);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            acc[j] = __float_as_uint(F16 ? fmaf(__uint_as_float(acc2[j]), 1.0f / SS2_F16_LO_SCALE, __uint_as_float(acc[j]))
                                         : __uint_as_float(acc[j]) + __uint_as_float(acc2[j]));
        }
        if (step == NSTEP - 1) {
          // this warp's accumulator quarter is in registers: hand the TMEM stage back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[as])) : "memory");
        }
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            stg[lane * 4 + (jj ^ ((lane >> 1) & 3))] =
                make_float4(__uint_as_float(acc[sub * 16 + jj * 4]), __uint_as_float(acc[sub * 16 + jj * 4 + 1]),
                            __uint_as_float(acc[sub * 16 + jj * 4 + 2]), __uint_as_float(acc[sub * 16 + jj * 4 + 3]));
          __syncwarp();
          const int c = cout0 + half * 32 + sub * 16 + (lane & 3) * 4;
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.bias && c < P.Cout) bb = __ldg(reinterpret_cast<const float4*>(P.bias + c));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = (lane >> 2) + 8 * i;
            const float4 a = stg[row * 4 + ((lane & 3) ^ ((row >> 1) & 3))];
            if (vv[i] && c < P.Cout) {
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(DC_TMEM_COLS));
  }
}

// stride-1 3x3 2-D layers with 32-channel chunks, Cout a multiple of 64 and a row that fits one tile
bool conv_dc_eligible(const ConvLayer& L, int D, int H, int W) {
  return L.wk_hi != nullptr && L.KD == 1 && L.KH == 3 && L.KW == 3 && L.sh == 1 && L.sw == 1 && L.ph == 1 && L.pw == 1 &&
         L.pd == 0 && D == 1 && (L.CinP % DC_BK) == 0 && (L.CoutP % 64) == 0 && (L.Cout % 64) == 0 && W + 2 <= 128 &&
         W >= 4 && H >= 1;
}

int conv_dc_launch(ss2_ctx* ctx, const ConvLayer& L, const ActRef& in, int B, int H, int W, const ActRef& out,
                   const float* d_residual, int relu, cudaStream_t st) {
  DcParams P;
  P.bias = L.bias; P.residual = d_residual;
  P.out_v = out.v; P.out_hi = out.hi; P.out_lo = out.lo;
  P.out_h16 = out.h16; P.out_l16 = out.h16 ? out.l16 : nullptr; P.range_flag = ctx->d_range_flag;
  const bool f16 = in.h16 != nullptr && in.hi == nullptr;   // fp16 split planes in: kind::f16 MMAs
  if (f16 && (!in.l16 || !L.wk_h16 || !L.wk_l16 || (L.CinP % 64) != 0 || ctx->tc_passes == 1))
    return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_dc: fp16 planes given to a layer without fp16 filter planes");
  P.B = B; P.H = H; P.W = W; P.Cout = L.Cout;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
  // N tile: 128 output channels per MMA where the layer has them (half the tcgen05.mma instructions per FLOP and the
  // staged input tile is read once for 128 channels), else 64
  int BN = (L.CoutP % 128) == 0 ? 128 : 64;
  P.nchunk = L.CinP / (f16 ? 2 * DC_BK : DC_BK); P.CinP = L.CinP;
  P.relu = relu;
  { const char* e = getenv("SS2_DC_DBG"); P.dbg = e ? atoi(e) : 0; }
  P.npass = f16 ? 3 : (ctx->tc_passes == 1 || !in.lo || !L.wk_lo) ? 1 : 3;
  const int nplanes = P.npass == 3 ? 2 : 1;
  const size_t budget = 227 * 1024 - 2048 - 8192;   // static shared memory: barriers + the epilogue's 8 KB staging buffers
  size_t b_stage = 0;
  size_t a_stage = 0;
  int MT = 2;
  // tile = TH rows x TW columns of the image (pitch TW + 2) in MT 128-row GEMM blocks
  auto plan = [&](int tiles_w, int mt_blocks) -> bool {
    P.tiles_w = tiles_w;
    P.TW = cdiv(W, tiles_w);
    P.P = P.TW + 2;
    if (P.P > 128) return false;
    P.TH = 128 * mt_blocks / P.P; if (P.TH > H) P.TH = H;
    if (P.TH < 1) return false;
    P.a_rows = (128 * mt_blocks + 2 * P.P + 2 + 7) / 8 * 8;
    a_stage = (size_t)P.a_rows * 128 * nplanes;
    // two input stages: the next chunk's (fp16 planes, 64-channel layers: the next TILE's) box loads under the current MMAs
    P.NA = (P.nchunk >= 2 || f16) ? 2 : 1;
    if ((size_t)P.NA * a_stage + 2 * b_stage + 1024 > budget) P.NA = 1;
    if ((size_t)P.NA * a_stage + 2 * b_stage + 1024 > budget) return false;
    P.NB = (int)((budget - 1024 - (size_t)P.NA * a_stage) / b_stage);
    if (P.NB > 8) P.NB = 8;
    P.tiles_h = cdiv(H, P.TH);
    P.n_mtiles = B * P.tiles_h * P.tiles_w;
    return true;
  };
  // small maps (the regressor stacks: a few dozen tiles for 148 SMs): 64-channel N tiles double the CTAs, and a
  // 128x64x8 MMA takes ~0.7x the time of a 128x128x8 one
  int small_thr = nsm / 2;
  { const char* e = getenv("SS2_DC_SMALL"); if (e) small_thr = atoi(e); }
  int force_tw = 0, force_mt = 0;
  { const char* e = getenv("SS2_DC_PLAN"); if (e) sscanf(e, "%d,%d", &force_tw, &force_mt); }
  for (int attempt = 0; attempt < 2; ++attempt) {
    P.n_ntiles = L.CoutP / BN;
    b_stage = (size_t)BN * DC_BK * 4 * nplanes;
    // Candidates: 1-4 column tiles x one or two 128-row blocks per tile.  Model of a launch (cycles per CTA, fitted to the
    // measurements in profiles/r02_conv_dc_plan_sweep.jsonl): rounds x [MMA pairs x cost per pair (junk rows of a tile
    // are paid like real ones) + staged bytes / rate], with penalties for a single input stage (loads and MMAs
    // alternate) and for fewer than three weight stages.
    double best = -1.0;
    int best_tw = 0, best_mt = 0;
    for (int mtb = 2; mtb >= 1; --mtb) {
      if (mtb == 2 && nplanes == 2 && BN == 128) continue;   // [D1 | D2] of two blocks x two stages exceed the 512 TMEM columns
      for (int tw = 1; tw <= 4; ++tw) {
        if (force_tw > 0 && (tw != force_tw || mtb != force_mt)) continue;
        if (!plan(tw, mtb)) continue;
        if (tw > 1 && P.TW < 8) continue;
        const double pair = nplanes == 2 ? (BN == 64 ? 201.0 : 300.0) : (BN == 64 ? 75.0 : 110.0);
        const double mma = (double)P.nchunk * 9 * 4 * mtb * pair;
        const double bytes = (double)P.nchunk * ((double)(P.TH + 2) * P.P * 128 * nplanes + 9.0 * b_stage);
        const long ntl = (long)P.n_mtiles * P.n_ntiles;
        double c = (double)((ntl + nsm - 1) / nsm) * (mma + bytes / 40.0);
        if (P.NA < 2 && (P.nchunk >= 2 || f16)) c *= 1.15;
        if (P.NB < 3) c *= 1.03;
        if (best < 0 || c < best) { best = c; best_tw = tw; best_mt = mtb; }
      }
    }
    if (best < 0) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "conv_dc: tile does not fit shared memory");
    MT = best_mt;
    plan(best_tw, best_mt);
    if (BN == 128 && (long)P.n_mtiles * P.n_ntiles <= small_thr) { BN = 64; continue; }
    break;
  }
  const size_t smem = (size_t)P.NA * a_stage + (size_t)P.NB * b_stage + 1024;
  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
  if (f16) {
    SS2_TRY(make_act_map(ctx, &mA_hi, in.h16, L.CinP, W, H, 1, B, P.P, P.TH + 2, 1, 1, 1, 1, 1, true));
    SS2_TRY(make_act_map(ctx, &mA_lo, in.l16, L.CinP, W, H, 1, B, P.P, P.TH + 2, 1, 1, 1, 1, 1, true));
    SS2_TRY(make_weight_map(ctx, &mB_hi, L.wk_h16, 9 * L.CinP, L.CoutP, true));
    SS2_TRY(make_weight_map(ctx, &mB_lo, L.wk_l16, 9 * L.CinP, L.CoutP, true));
  } else {
    SS2_TRY(make_act_map(ctx, &mA_hi, in.hi, L.CinP, W, H, 1, B, P.P, P.TH + 2, 1, 1, 1, 1, 1));
    SS2_TRY(make_weight_map(ctx, &mB_hi, L.wk_hi, 9 * L.CinP, L.CoutP));
    if (P.npass == 3) {
      SS2_TRY(make_act_map(ctx, &mA_lo, in.lo, L.CinP, W, H, 1, B, P.P, P.TH + 2, 1, 1, 1, 1, 1));
      SS2_TRY(make_weight_map(ctx, &mB_lo, L.wk_lo, 9 * L.CinP, L.CoutP));
    } else {
      mA_lo = mA_hi; mB_lo = mB_hi;
    }
  }
  static size_t attr_smem_dev[16][8] = {{0}};  // per device: function attributes live in the device's context
  size_t* attr_smem = attr_smem_dev[ctx->device & 15];
  const int variant = (BN == 128 ? 2 : 0) + (MT == 2 ? 1 : 0) + (f16 ? 4 : 0);
  if (smem > attr_smem[variant]) {
    const void* fn = variant == 0 ? (const void*)conv_dc_kernel<64, 1> : variant == 1 ? (const void*)conv_dc_kernel<64, 2>
                   : variant == 2 ? (const void*)conv_dc_kernel<128, 1> : variant == 3 ? (const void*)conv_dc_kernel<128, 2>
                   : variant == 4 ? (const void*)conv_dc_kernel<64, 1, true> : variant == 5 ? (const void*)conv_dc_kernel<64, 2, true>
                   : (const void*)conv_dc_kernel<128, 1, true>;
    SS2_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[variant] = smem;
  }
  const int total = P.n_mtiles * P.n_ntiles;
  const int grid = total < nsm ? total : nsm;
  const double flops = 2.0 * B * H * W * (double)L.Cout * 9 * L.Cin;
  ss2_prof_begin(ctx, SS2_PROF_CONV, st);
  if (variant == 0) conv_dc_kernel<64, 1><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  else if (variant == 1) conv_dc_kernel<64, 2><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  else if (variant == 2) conv_dc_kernel<128, 1><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  else if (variant == 3) conv_dc_kernel<128, 2><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  else if (variant == 4) conv_dc_kernel<64, 1, true><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  else if (variant == 5) conv_dc_kernel<64, 2, true><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  else conv_dc_kernel<128, 1, true><<<grid, DC_THREADS, smem, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, P);
  ss2_prof_end(ctx, SS2_PROF_CONV, st, flops);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
