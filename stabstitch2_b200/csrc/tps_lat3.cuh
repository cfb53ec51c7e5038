// tps_warp_lat3_kernel: the production fused TPS resample + AVERAGE blend (NORMAL sampling, two views, three
// colour planes, source size known at compile time).  Included by tps.cu after the lattice helpers.
//
// Same mathematics as tps_warp_lattice_kernel (quintic lattice interpolation of the far field + exact near
// field, exactly-zero out-of-image samples), re-organised around what the round-1 ncu capture showed: that
// kernel ran at 84 % of the L1 data pipe's wavefront rate and 71 % issue activity, i.e. it was bound by LSU
// wavefronts (6 LDS.128 of the shared y-contracted table + 24 unaligned LDG per pixel row) as much as by
// instruction count.  Here
//   * the lattice never goes through shared memory: a thread owns one canvas column and keeps the x-contracted
//     values of the six node rows around its current lattice cell in registers (two views x (x, y) = 12 packed
//     pairs).  Moving down one cell shifts the window and x-contracts ONE new node row (6 LDG.128 of a few
//     hundred bytes per warp).  The y contraction of a canvas row is then 12 FFMA2 whose weights are
//     warp-uniform (they depend on the row only) and come from the constant bank through the uniform datapath;
//   * bilinear taps are kept in registers from one canvas row to the next: walking down a column at scale ~1
//     the bottom tap row of pixel r is the top tap row of pixel r+1, so a row normally loads 6 values per view
//     instead of 12 (per-lane predicates; lanes whose column drifted or whose source row jumped reload all);
//   * coordinates of the two views travel as packed pairs ((x_v0, x_v1), (y_v0, y_v1)): predictor, floor
//     (round-towards-minus-infinity add of 1.5*2^23) and fraction are packed FADD2/FFMA2; the bilinear
//     interpolation is the lerp form, packed over colour planes 0/1 and over the tap rows for plane 2;
//   * no block barrier after the prologue.
#pragma once

#ifndef L3_NCELL
#define L3_NCELL 8   // lattice cell rows per CTA: the tile is 128 x (L3_NCELL*SY) canvas pixels
#endif
#ifndef L3_MINB
#define L3_MINB 4
#endif
#ifndef L3_PF_ROWS
#define L3_PF_ROWS 2  // source rows ahead of the bottom tap row pulled towards the SM (0: no prefetch)
#endif
#ifndef L3_PF_L1
#define L3_PF_L1 0    // 1: prefetch into L1, 0: into L2 only
#endif
#define L3_THREADS 128
#define L3_MAGIC 12582912.0f          // 1.5 * 2^23: v + MAGIC (rounded down) = MAGIC + floor(v) for |v| < 2^22
#define L3_MAGIC_BITS 0x4b400000
#define L3_INVALID 0x80000000u

struct LagrangePairs { float2 w[16][LAT_TAPS]; };   // (w, w): quintic weights duplicated for packed FMAs
__constant__ LagrangePairs c_lagp[4];               // spacing 6, 8, 12, 16 (lag_idx)

__device__ __forceinline__ u64 fsub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fadd2_rm(u64 a, u64 b) {
  u64 r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// base + off*4 with a SIGNED 32-bit element offset (one IMAD.WIDE)
__device__ __forceinline__ const float* f32_at_s(const float* base, int off) {
  unsigned long long a;
  asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(a) : "r"(off), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<const float*>(a);
}

// six taps of one source row (x0, x1 of the three colour planes) at byte offset ROW from p, only in lanes with pred;
// p addresses the MIDDLE plane, so that the +-plane offsets of a 1080p frame (8.3 MB) still fit the load immediates
template <int PLANE_B, int ROW_B>
__device__ __forceinline__ void l3_load_row(float (&x0)[3], float (&x1)[3], const float* p, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %7, 0;\n\t"
      "@q ld.global.nc.f32 %0, [%6+%8];\n\t"
      "@q ld.global.nc.f32 %1, [%6+%9];\n\t"
      "@q ld.global.nc.f32 %2, [%6+%10];\n\t"
      "@q ld.global.nc.f32 %3, [%6+%11];\n\t"
      "@q ld.global.nc.f32 %4, [%6+%12];\n\t"
      "@q ld.global.nc.f32 %5, [%6+%13];\n\t}"
      : "+f"(x0[0]), "+f"(x1[0]), "+f"(x0[1]), "+f"(x1[1]), "+f"(x0[2]), "+f"(x1[2])
      : "l"(p), "r"((unsigned)pred), "n"(ROW_B - PLANE_B), "n"(ROW_B + 4 - PLANE_B), "n"(ROW_B), "n"(ROW_B + 4),
        "n"(PLANE_B + ROW_B), "n"(PLANE_B + ROW_B + 4));
}

// the three planes' lines of a source row further down, towards L2 (or L1)
template <int PLANE_B, int ROW_B>
__device__ __forceinline__ void l3_prefetch(const float* p, bool pred) {
#if L3_PF_L1
#define L3_PF_OP "prefetch.global.L1"
#else
#define L3_PF_OP "prefetch.global.L2"
#endif
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t"
      "@q " L3_PF_OP " [%0+%2];\n\t"
      "@q " L3_PF_OP " [%0+%3];\n\t"
      "@q " L3_PF_OP " [%0+%4];\n\t}"
      :: "l"(p), "r"((unsigned)pred), "n"(ROW_B - PLANE_B), "n"(ROW_B), "n"(ROW_B + PLANE_B));
#undef L3_PF_OP
}

// x contraction of one lattice node row for this thread's column: 6 nodes x (x_v0, x_v1, y_v0, y_v1)
__device__ __forceinline__ void l3_xcontract(const float4* __restrict__ nd, const u64 (&lx)[LAT_TAPS], u64& gx, u64& gy) {
  float4 q[LAT_TAPS];
#pragma unroll
  for (int a = 0; a < LAT_TAPS; ++a) q[a] = __ldg(nd + a);
  gx = fmul2(pk2(q[0].x, q[0].y), lx[0]);
  gy = fmul2(pk2(q[0].z, q[0].w), lx[0]);
#pragma unroll
  for (int a = 1; a < LAT_TAPS; ++a) {
    gx = ffma2(pk2(q[a].x, q[a].y), lx[a], gx);
    gy = ffma2(pk2(q[a].z, q[a].w), lx[a], gy);
  }
}

// P.nodes here is [n][ny][nx] float4 = (x_v0, x_v1, y_v0, y_v1) residual source pixel coordinates (tps_nodes_kernel<2, 1>)
template <int SX, int SY, int IW, int IH>
__global__ void __launch_bounds__(L3_THREADS, L3_MINB)
tps_warp_lat3_kernel(WarpParams P) {
  constexpr int V = 2;
  constexpr int TILE_ROWS = L3_NCELL * SY;
  constexpr int PLANE_B = IW * IH * 4, ROW_B = IW * 4;
  static_assert(SY <= 16, "row-in-cell index addresses a 16-row weight table");
  static_assert((size_t)IW * IH * 4 + (size_t)IW * 4 + 8 < (1u << 23), "tap offsets must fit the load immediates");
  __shared__ float4 near_c[V * SS2_NPT];   // (cx, cy, -, -) of the control points whose disc touches this tile
  __shared__ float4 near_w[V * SS2_NPT];   // (wx_v0, wx_v1, wy_v0, wy_v1) * ln2, in source pixels (one view's pair is 0)
  __shared__ int warp_cnt[L3_THREADS / 32];
  __shared__ float s_pred[V][6];
  const int n = blockIdx.z, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int col0 = blockIdx.x * L3_THREADS, row00 = blockIdx.y * TILE_ROWS;
  // ---- near list (deterministic order: view 0's points, then view 1's)
  {
    float4 ec = make_float4(0.f, 0.f, 0.f, 0.f), ew = ec;
    bool hit = false;
    const int pv = tid / SS2_NPT, pi = tid - pv * SS2_NPT;
    if (tid < V * SS2_NPT) {
      const float x_lo = fmaf(P.stepx, (float)col0, -1.0f), x_hi = fmaf(P.stepx, (float)min(col0 + L3_THREADS - 1, P.Wo - 1), -1.0f);
      const float y_lo = fmaf(P.stepy, (float)row00, -1.0f), y_hi = fmaf(P.stepy, (float)min(row00 + TILE_ROWS - 1, P.Ho - 1), -1.0f);
      const float2 c = *reinterpret_cast<const float2*>(P.source + ((size_t)(n * V + pv) * SS2_NPT + pi) * 2);
      const float* t = P.T + (size_t)(n * V + pv) * 2 * SS2_NSYS;
      const float ddx = fmaxf(fmaxf(x_lo - c.x, c.x - x_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
      hit = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
      if (hit) {
        const float wx = t[3 + pi] * (P.half_w * LN2F), wy = t[SS2_NSYS + 3 + pi] * (P.half_h * LN2F);
        ec = make_float4(c.x, c.y, 0.f, 0.f);
        ew = pv == 0 ? make_float4(wx, 0.f, wy, 0.f) : make_float4(0.f, wx, 0.f, wy);
      }
    }
    const unsigned m_all = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[wid] = __popc(m_all);
    if (tid < V * 6) s_pred[tid / 6][tid % 6] = P.aux[(size_t)(n * V + tid / 6) * 8 + tid % 6];
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int w = 0; w < L3_THREADS / 32; ++w)
      if (w < wid) off += warp_cnt[w];
    if (hit) {
      const int slot = off + __popc(m_all & ((1u << lane) - 1u));
      near_c[slot] = ec;
      near_w[slot] = ew;
    }
    __syncthreads();
  }
  int n_all = 0;
#pragma unroll
  for (int w = 0; w < L3_THREADS / 32; ++w) n_all += warp_cnt[w];
  const bool cull = n_all <= 32;  // one candidate per lane; longer lists (never seen) are evaluated in full

  // ---- per-column constants
  const int col = min(col0 + tid, P.Wo - 1);
  const bool active = col0 + tid < P.Wo;
  const int cxi = col / SX, rx = col - cxi * SX;
  u64 lx[LAT_TAPS];
#pragma unroll
  for (int a = 0; a < LAT_TAPS; ++a) {
    const float w = g_lag[lag_idx(SX)].w[rx][a];
    lx[a] = pk2(w, w);
  }
  const float colf = (float)col;
  const int row_end = min(row00 + TILE_ROWS, P.Ho);
  // Tile-local source coordinates: per view an integer origin (X0, Y0) = the predictor at the tile centre.  The affine
  // predictor px = pc[0]*col + pc[1]*row + pc[2] (tps_solve_kernel) minus the origin is folded into the x-contracted
  // node rows (Lagrange weights reproduce functions linear in the row), so the y contraction yields ORIGIN-RELATIVE
  // source coordinates of magnitude ~100: every later rounding is at the 1e-5 px level, and no per-row predictor work.
  float orgx[V], orgy[V];
  {
    const float ccol = (float)min(col0 + L3_THREADS / 2, P.Wo - 1), crow = (float)((row00 + row_end) >> 1);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      orgx[v] = fminf(fmaxf(floorf(fmaf(s_pred[v][0], ccol, fmaf(s_pred[v][1], crow, s_pred[v][2]))), -1048576.f), 1048576.f);
      orgy[v] = fminf(fmaxf(floorf(fmaf(s_pred[v][3], ccol, fmaf(s_pred[v][4], crow, s_pred[v][5]))), -1048576.f), 1048576.f);
    }
  }
  // predictor at this column and canvas row 0, origin-relative, packed over the views
  const u64 pcolX = pk2(fmaf(s_pred[0][0], colf, s_pred[0][2] - orgx[0]), fmaf(s_pred[1][0], colf, s_pred[1][2] - orgx[1]));
  const u64 pcolY = pk2(fmaf(s_pred[0][3], colf, s_pred[0][5] - orgy[0]), fmaf(s_pred[1][3], colf, s_pred[1][5] - orgy[1]));
  const u64 prowX = pk2(s_pred[0][1], s_pred[1][1]), prowY = pk2(s_pred[0][4], s_pred[1][4]);
  // in-image window of the origin-relative coordinates: 0 <= x < W-1, 0 <= y < H-1 (no tap clamps: plain bilinear)
  float lox[V], hix[V], loy[V], hiy[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    lox[v] = -orgx[v]; hix[v] = (float)(IW - 1) - orgx[v];
    loy[v] = -orgy[v]; hiy[v] = (float)(IH - 1) - orgy[v];
  }
  const float xt = fmaf(P.stepx, colf, -1.0f);
  const float wx_lo = fmaf(P.stepx, (float)min(col0 + wid * 32, P.Wo - 1), -1.0f);
  const float wx_hi = fmaf(P.stepx, (float)min(col0 + wid * 32 + 31, P.Wo - 1), -1.0f);
  const unsigned oplane = (unsigned)(P.Ho * P.Wo);
  // per view: middle plane of this frame at the tile origin (taps are addressed with signed origin-relative offsets)
  const float* imgv[V];
#pragma unroll
  for (int v = 0; v < V; ++v)
    imgv[v] = P.img[v] + (size_t)n * 3 * IW * IH + (size_t)IW * IH + ((long long)orgy[v] * IW + (long long)orgx[v]);
  // one running output pointer per colour plane (advanced by a canvas row per pixel row)
  float* po0 = P.out + (size_t)n * 3 * oplane + (size_t)row00 * P.Wo + col;
  float* po1 = po0 + oplane;
  float* po2 = po1 + oplane;

  // ---- node window: slots 0..5 <-> lattice rows cyi-2..cyi+3 of the current cell row cyi; slot b of cell row cyi
  // belongs to canvas row (cyi + b - 2) * SY
  const int cy0 = blockIdx.y * L3_NCELL;
  const float4* ndp = reinterpret_cast<const float4*>(P.nodes) + ((size_t)n * P.ny + cy0) * P.nx + cxi;
  u64 gx[LAT_TAPS], gy[LAT_TAPS];
  gx[0] = gy[0] = 0ull;
  float noderow = (float)((cy0 - LAT_LO) * SY);   // canvas row of the node row loaded next
#pragma unroll
  for (int b = 1; b < LAT_TAPS; ++b) {
    l3_xcontract(ndp, lx, gx[b], gy[b]);
    const u64 nr2 = pk2(noderow, noderow);
    gx[b] = fadd2(gx[b], ffma2(prowX, nr2, pcolX));
    gy[b] = fadd2(gy[b], ffma2(prowY, nr2, pcolY));
    ndp += P.nx;
    noderow += (float)SY;
  }

  // ---- tap registers (top / bottom source row, columns x0 / x1, three planes) and the offset they belong to
  float t0[V][3], t1[V][3], b0[V][3], b1[V][3];
  int prev_off[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    prev_off[v] = (int)L3_INVALID;
#pragma unroll
    for (int c = 0; c < 3; ++c) t0[v][c] = t1[v][c] = b0[v][c] = b1[v][c] = 0.f;
  }

  // ---- software pipeline over the tile's rows.  Iteration rc:
  //   A  interpolate row rc-2 from the tap registers (its loads were issued in the previous iteration),
  //   B  issue the tap loads of row rc-1 (both views back to back),
  //      blend + store row rc-2,
  //   C  evaluate the field of row rc (y contraction, near field, floor / fraction, in-image flags)
  // so the loads of a row are in flight during the blend/stores of the row before and the whole field evaluation of
  // the row after.  State "a" belongs to the row A handles, "b" to the row B handles.
  u64 FXa = 0ull, FYa = 0ull, FXb = 0ull, FYb = 0ull;
  int offb[V] = {0, 0};
  unsigned fla = 0u, flb = 0u;   // bit v: sample of view v inside its image; bit 2+v: some lane of the warp has bit v
  unsigned cand = 0u;
  int r = 0;                     // row within the lattice cell of row rc
#pragma unroll 1
  for (int rc = row00; rc < row_end + 2; ++rc) {
    // ---- A: bilinear interpolation of row rc-2 (lerp form: planes 0/1 packed, plane 2 packed over (top, bottom))
    u64 o01[V];
    float o2[V];
    const bool doA = rc - 2 >= row00;
    float fxa[V], fya[V];
    upk2(FXa, fxa[0], fxa[1]); upk2(FYa, fya[0], fya[1]);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      o01[v] = 0ull;
      o2[v] = 0.f;
      if (doA && (fla & (4u << v))) {
        const bool inb = (fla >> v) & 1u;
        const u64 fx2 = pk2(fxa[v], fxa[v]), fy2 = pk2(fya[v], fya[v]);
        const u64 A = pk2(t0[v][0], t0[v][1]), Cc = pk2(t1[v][0], t1[v][1]);
        const u64 B = pk2(b0[v][0], b0[v][1]), D = pk2(b1[v][0], b1[v][1]);
        const u64 top = ffma2(fx2, fsub2(Cc, A), A), bot = ffma2(fx2, fsub2(D, B), B);
        const u64 r01 = ffma2(fy2, fsub2(bot, top), top);
        const u64 L = pk2(t0[v][2], b0[v][2]), R = pk2(t1[v][2], b1[v][2]);
        float tp2, bt2;
        upk2(ffma2(fx2, fsub2(R, L), L), tp2, bt2);
        const float r2 = fmaf(fya[v], bt2 - tp2, tp2);
        o01[v] = inb ? r01 : 0ull;
        o2[v] = inb ? r2 : 0.f;
      }
    }
    // ---- B: tap loads of row rc-1.  Walking down a column the bottom tap row of the previous pixel is normally the
    // top tap row of this one (offset + one source row): those lanes move bottom -> top and load only the new bottom row
    if (rc - 1 >= row00 && rc - 1 < row_end) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if (flb & (4u << v)) {
          const bool inb = (flb >> v) & 1u;
          const int off = offb[v];
          const int d = off - prev_off[v];
          const bool same = d == 0, shift = d == IW;
          prev_off[v] = inb ? off : (int)L3_INVALID;
          const float* p = f32_at_s(imgv[v], off);
          if (shift) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { t0[v][c] = b0[v][c]; t1[v][c] = b1[v][c]; }
          }
          const bool needB = inb && !same, needT = needB && !shift;
          l3_load_row<PLANE_B, ROW_B>(b0[v], b1[v], p, needB);
          if (__any_sync(0xffffffffu, needT)) l3_load_row<PLANE_B, 0>(t0[v], t1[v], p, needT);
#if L3_PF_ROWS > 0
          l3_prefetch<PLANE_B, (1 + L3_PF_ROWS) * ROW_B>(p, needB && off < (IH - 2 - L3_PF_ROWS) * IW - (int)orgy[v] * IW);
#endif
        } else {
          prev_off[v] = (int)L3_INVALID;
        }
      }
    }
    // ---- AVERAGE blend (a*a + b*b) / (a + b + 1e-6) of row rc-2, streaming stores
    if (doA) {
      if (active) {
        float s0, s1, q0, q1;
        upk2(fadd2(fadd2(o01[0], o01[1]), pk2(1e-6f, 1e-6f)), s0, s1);
        upk2(ffma2(o01[1], o01[1], fmul2(o01[0], o01[0])), q0, q1);
        __stcs(po0, q0 * rcp_approx(s0));
        __stcs(po1, q1 * rcp_approx(s1));
        __stcs(po2, blend_avg_fast(o2[0], o2[1]));
      }
      po0 += P.Wo; po1 += P.Wo; po2 += P.Wo;
    }
    FXa = FXb; FYa = FYb; fla = flb;
    flb = 0u;
    // ---- C: field of row rc
    if (rc < row_end) {
      if (r == 0) {
        // new lattice cell row: shift the window down one node row and x-contract the new last row (+ predictor)
#pragma unroll
        for (int b = 0; b < LAT_TAPS - 1; ++b) { gx[b] = gx[b + 1]; gy[b] = gy[b + 1]; }
        l3_xcontract(ndp, lx, gx[LAT_TAPS - 1], gy[LAT_TAPS - 1]);
        const u64 nr2 = pk2(noderow, noderow);
        gx[LAT_TAPS - 1] = fadd2(gx[LAT_TAPS - 1], ffma2(prowX, nr2, pcolX));
        gy[LAT_TAPS - 1] = fadd2(gy[LAT_TAPS - 1], ffma2(prowY, nr2, pcolY));
        ndp += P.nx;
        noderow += (float)SY;
        // per-warp culling of the near list against this warp's 32 x SY pixel block
        cand = 0u;
        if (n_all > 0) {
          if (cull) {
            bool keep = false;
            if (lane < n_all) {
              const float4 c = near_c[lane];
              const float y_lo = fmaf(P.stepy, (float)rc, -1.0f), y_hi = fmaf(P.stepy, (float)min(rc + SY - 1, P.Ho - 1), -1.0f);
              const float ddx = fmaxf(fmaxf(wx_lo - c.x, c.x - wx_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
              keep = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
            }
            cand = __ballot_sync(0xffffffffu, keep);
          } else {
            cand = 0xffffffffu;
          }
        }
      }
      // y contraction with warp-uniform weight pairs (constant bank, uniform datapath)
      const u64* ly = reinterpret_cast<const u64*>(&c_lagp[lag_idx(SY)].w[r][0]);
      u64 PX = fmul2(gx[0], ly[0]), PY = fmul2(gy[0], ly[0]);
#pragma unroll
      for (int b = 1; b < LAT_TAPS; ++b) { PX = ffma2(gx[b], ly[b], PX); PY = ffma2(gy[b], ly[b], PY); }
      // near-field corrections (branch-free inside: s is clamped to R2, where psi vanishes)
      if (cand != 0u) {
        const float yt = fmaf(P.stepy, (float)rc, -1.0f);
        unsigned m = cull ? cand : 0u;
        int kk = 0;
#pragma unroll 1
        while (cull ? (m != 0u) : (kk < n_all)) {
          int k;
          if (cull) { k = __ffs(m) - 1; m &= m - 1; } else { k = kk++; }
          const float4 c = near_c[k];
          const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(&near_w[k]);
          const float dx = xt - c.x, dy = yt - c.y;
          const float s = fminf(fmaf(dy, dy, dx * dx), P.R2);
          const float psi = fmaf(s, lg2_approx(s + 1e-6f), -blend_poly(s, P.R2, P.q0, P.q1, P.q2, P.q3));
          const u64 psi2 = pk2(psi, psi);
          PX = ffma2(w.x, psi2, PX);
          PY = ffma2(w.y, psi2, PY);
        }
      }
      // floor / fraction of both views' coordinates, in-image flags, tap offsets
      const u64 mg = pk2(L3_MAGIC, L3_MAGIC);
      const u64 TXm = fadd2_rm(PX, mg), TYm = fadd2_rm(PY, mg);
      FXb = fsub2(PX, fsub2(TXm, mg));
      FYb = fsub2(PY, fsub2(TYm, mg));
      float pxv[V], pyv[V], txv[V], tyv[V];
      upk2(PX, pxv[0], pxv[1]); upk2(PY, pyv[0], pyv[1]);
      upk2(TXm, txv[0], txv[1]); upk2(TYm, tyv[0], tyv[1]);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const bool inb = pxv[v] >= lox[v] && pxv[v] < hix[v] && pyv[v] >= loy[v] && pyv[v] < hiy[v];
        const int xi = __float_as_int(txv[v]) - L3_MAGIC_BITS, yi = __float_as_int(tyv[v]) - L3_MAGIC_BITS;
        offb[v] = yi * IW + xi;
        flb |= (inb ? 1u : 0u) << v;
        flb |= (__any_sync(0xffffffffu, inb) ? 4u : 0u) << v;
      }
      r = r + 1 == SY ? 0 : r + 1;
    }
  }
}
