// tps_warp_lat4_kernel: the production fused TPS resample + AVERAGE blend (NORMAL sampling, two views, three
// colour planes, source size known at compile time).  Included by tps.cu after tps_lat3.cuh (shares its helpers).
//
// Same mathematics as tps_warp_lattice_kernel / tps_warp_lat3_kernel.  What the round-2 ncu captures showed:
// the round-1 kernel ran at 84 % of the L1 data pipe's wavefront rate (6 LDS.128 of the shared y-contracted
// table + 24 unaligned LDG per pixel row); keeping the lattice window and the taps of BOTH views in one thread's
// registers (lat3) removed those wavefronts but needs 126 registers -> 16 warps per SM, and was latency bound.
// Here a WARP handles ONE view of a 32-column strip:
//   * CTA = 8 warps = 4 column strips x 2 views over a 128 x (L4_NCELL*SY) canvas tile.  Per thread: the
//     x-contracted lattice window of its view (6 packed (x, y) pairs), the tap registers of its view (12) -> about
//     half the state, twice the resident warps, and the two views' loads are in flight from different warps;
//   * per lattice cell (SY rows) both warps of a strip write their interpolated pixels to shared memory, meet at a
//     64-thread named barrier, and each blends + stores half of the rows;
//   * everything else as in lat3: y contraction with warp-uniform weight pairs from the constant bank, origin-
//     relative coordinates with the affine predictor folded into the window, floor by a round-down magic add,
//     bottom -> top tap reuse from one canvas row to the next, lerp-form bilinear interpolation on packed pairs.
#pragma once

#ifndef L4_NCELL
#define L4_NCELL 8
#endif
#ifndef L4_MINB
#define L4_MINB 3
#endif
#ifndef L4_PF_ROWS
#define L4_PF_ROWS 2
#endif
#define L4_THREADS 256
#define L4_COLS 128

__device__ __forceinline__ void l4_pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// x contraction of one lattice node row for this thread's column and view: 6 nodes x (x, y); nodes are
// [..][nx][2 views] float2, so consecutive nodes of a view are 16 bytes apart
__device__ __forceinline__ u64 l4_xcontract(const float2* __restrict__ nd, const float (&lx)[LAT_TAPS]) {
  float2 q[LAT_TAPS];
#pragma unroll
  for (int a = 0; a < LAT_TAPS; ++a) q[a] = __ldg(nd + 2 * a);
  u64 g = fmul2(pk2(q[0].x, q[0].y), pk2(lx[0], lx[0]));
#pragma unroll
  for (int a = 1; a < LAT_TAPS; ++a) g = ffma2(pk2(q[a].x, q[a].y), pk2(lx[a], lx[a]), g);
  return g;
}

// P.nodes: [n][ny][nx][2] float2 (x, y) residual source pixel coordinates (tps_nodes_kernel<2, 0>)
template <int SX, int SY, int IW, int IH>
__global__ void __launch_bounds__(L4_THREADS, L4_MINB)
tps_warp_lat4_kernel(WarpParams P) {
  constexpr int V = 2;
  constexpr int TILE_ROWS = L4_NCELL * SY;
  constexpr int PLANE_B = IW * IH * 4, ROW_B = IW * 4;
  constexpr int HALF = SY / 2;
  static_assert((size_t)IW * IH * 4 + (size_t)IW * 4 + 8 < (1u << 23), "tap offsets must fit the load immediates");
  static_assert(SY % 2 == 0 && SY <= 16, "rows of a cell are split between the two warps of a strip");
  __shared__ float4 near_c[V][SS2_NPT];   // (cx, cy, wx*ln2, wy*ln2): control points whose disc touches this tile
  __shared__ int near_cnt[V][2];
  __shared__ float s_pred[V][6];
  __shared__ float xbuf[V][SY][3][L4_COLS];   // interpolated pixels of the current cell, per view
  const int n = blockIdx.z, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int v = wid >> 2, cw = wid & 3;       // this warp's view and column strip
  const int col0 = blockIdx.x * L4_COLS, row00 = blockIdx.y * TILE_ROWS;
  const int row_end = min(row00 + TILE_ROWS, P.Ho);
  // ---- near list of each view, built by the first two warps of that view (63 control points)
  {
    float4 ent = make_float4(0.f, 0.f, 0.f, 0.f);
    bool hit = false;
    const int pi = cw * 32 + lane;
    if (cw < 2 && pi < SS2_NPT) {
      const float x_lo = fmaf(P.stepx, (float)col0, -1.0f), x_hi = fmaf(P.stepx, (float)min(col0 + L4_COLS - 1, P.Wo - 1), -1.0f);
      const float y_lo = fmaf(P.stepy, (float)row00, -1.0f), y_hi = fmaf(P.stepy, (float)(row_end - 1), -1.0f);
      const float2 c = *reinterpret_cast<const float2*>(P.source + ((size_t)(n * V + v) * SS2_NPT + pi) * 2);
      const float* t = P.T + (size_t)(n * V + v) * 2 * SS2_NSYS;
      const float ddx = fmaxf(fmaxf(x_lo - c.x, c.x - x_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
      hit = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
      if (hit) ent = make_float4(c.x, c.y, t[3 + pi] * (P.half_w * LN2F), t[SS2_NSYS + 3 + pi] * (P.half_h * LN2F));
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (cw < 2 && lane == 0) near_cnt[v][cw] = __popc(m);
    if (tid < V * 6) s_pred[tid / 6][tid % 6] = P.aux[(size_t)(n * V + tid / 6) * 8 + tid % 6];
    __syncthreads();
    if (hit) near_c[v][(cw == 1 ? near_cnt[v][0] : 0) + __popc(m & ((1u << lane) - 1u))] = ent;
    __syncthreads();
  }
  const int n_all = near_cnt[v][0] + near_cnt[v][1];
  const bool cull = n_all <= 32;  // one candidate per lane; longer lists (never seen) are evaluated in full

  // ---- per-column constants
  const int colt = cw * 32 + lane;            // column within the tile
  const int col = min(col0 + colt, P.Wo - 1);
  const bool active = col0 + colt < P.Wo;
  const int cxi = col / SX, rx = col - cxi * SX;
  float lx[LAT_TAPS];
#pragma unroll
  for (int a = 0; a < LAT_TAPS; ++a) lx[a] = g_lag[lag_idx(SX)].w[rx][a];
  const float colf = (float)col;
  // tile-local source coordinates: integer origin = the affine predictor at the tile centre (see lat3)
  float orgx, orgy;
  {
    const float ccol = (float)min(col0 + L4_COLS / 2, P.Wo - 1), crow = (float)((row00 + row_end) >> 1);
    orgx = fminf(fmaxf(floorf(fmaf(s_pred[v][0], ccol, fmaf(s_pred[v][1], crow, s_pred[v][2]))), -1048576.f), 1048576.f);
    orgy = fminf(fmaxf(floorf(fmaf(s_pred[v][3], ccol, fmaf(s_pred[v][4], crow, s_pred[v][5]))), -1048576.f), 1048576.f);
  }
  const u64 pcol = pk2(fmaf(s_pred[v][0], colf, s_pred[v][2] - orgx), fmaf(s_pred[v][3], colf, s_pred[v][5] - orgy));
  const u64 prow = pk2(s_pred[v][1], s_pred[v][4]);
  // in-image window of the origin-relative coordinates: 0 <= x < W-1, 0 <= y < H-1 (no tap clamps: plain bilinear)
  const float lox = -orgx, hix = (float)(IW - 1) - orgx, loy = -orgy, hiy = (float)(IH - 1) - orgy;
  const int pf_limit = (IH - 3 - L4_PF_ROWS - (int)orgy) * IW;   // prefetched rows stay inside the frame
  const float xt = fmaf(P.stepx, colf, -1.0f);
  const float wx_lo = fmaf(P.stepx, (float)min(col0 + cw * 32, P.Wo - 1), -1.0f);
  const float wx_hi = fmaf(P.stepx, (float)min(col0 + cw * 32 + 31, P.Wo - 1), -1.0f);
  const unsigned oplane = (unsigned)(P.Ho * P.Wo);
  // middle plane of this frame at the tile origin (taps are addressed with signed origin-relative offsets)
  const float* img = P.img[v] + (size_t)n * 3 * IW * IH + (size_t)IW * IH + ((long long)orgy * IW + (long long)orgx);
  float* const outp = P.out + (size_t)n * 3 * oplane;

  // ---- node window: slots 0..5 <-> lattice rows cyi-2..cyi+3 of the current cell row; + predictor of the node's canvas row
  const int cy0 = blockIdx.y * L4_NCELL;
  const float2* ndp = P.nodes + (((size_t)n * P.ny + cy0) * P.nx + cxi) * V + v;
  const size_t nd_stride = (size_t)P.nx * V;
  u64 g[LAT_TAPS];
  g[0] = 0ull;
  float noderow = (float)((cy0 - LAT_LO) * SY);
#pragma unroll
  for (int b = 1; b < LAT_TAPS; ++b) {
    g[b] = fadd2(l4_xcontract(ndp, lx), ffma2(prow, pk2(noderow, noderow), pcol));
    ndp += nd_stride;
    noderow += (float)SY;
  }
  // ---- tap registers (top / bottom source row, columns x0 / x1, three planes) and the offset they belong to
  float t0[3], t1[3], b0[3], b1[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) t0[c] = t1[c] = b0[c] = b1[c] = 0.f;
  int prev_off = (int)L3_INVALID;

#pragma unroll 1
  for (int cell = 0; cell < L4_NCELL; ++cell) {
    const int row0 = row00 + cell * SY;
    if (row0 >= row_end) break;  // CTA-uniform
    // shift the window down one lattice row and x-contract the new last row
#pragma unroll
    for (int b = 0; b < LAT_TAPS - 1; ++b) g[b] = g[b + 1];
    g[LAT_TAPS - 1] = fadd2(l4_xcontract(ndp, lx), ffma2(prow, pk2(noderow, noderow), pcol));
    ndp += nd_stride;
    noderow += (float)SY;
    // per-warp culling of this view's near list against the warp's 32 x SY pixel block
    unsigned cand = 0u;
    if (n_all > 0) {
      if (cull) {
        bool keep = false;
        if (lane < n_all) {
          const float4 c = near_c[v][lane];
          const float y_lo = fmaf(P.stepy, (float)row0, -1.0f), y_hi = fmaf(P.stepy, (float)min(row0 + SY - 1, P.Ho - 1), -1.0f);
          const float ddx = fmaxf(fmaxf(wx_lo - c.x, c.x - wx_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
          keep = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
        }
        cand = __ballot_sync(0xffffffffu, keep);
      } else {
        cand = 0xffffffffu;
      }
    }
#pragma unroll
    for (int r = 0; r < SY; ++r) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      if (row0 + r < row_end) {  // CTA-uniform
        // ---- field: y contraction with warp-uniform weight pairs (constant bank, uniform datapath)
        const u64* ly = reinterpret_cast<const u64*>(&c_lagp[lag_idx(SY)].w[r][0]);
        u64 Pc = fmul2(g[0], ly[0]);
#pragma unroll
        for (int b = 1; b < LAT_TAPS; ++b) Pc = ffma2(g[b], ly[b], Pc);
        // ---- near-field corrections (branch-free inside: s is clamped to R2, where psi vanishes)
        if (cand != 0u) {
          const float yt = fmaf(P.stepy, (float)(row0 + r), -1.0f);
          unsigned m = cull ? cand : 0u;
          int kk = 0;
#pragma unroll 1
          while (cull ? (m != 0u) : (kk < n_all)) {
            int k;
            if (cull) { k = __ffs(m) - 1; m &= m - 1; } else { k = kk++; }
            const float4 c = near_c[v][k];
            const float dx = xt - c.x, dy = yt - c.y;
            const float s = fminf(fmaf(dy, dy, dx * dx), P.R2);
            const float psi = fmaf(s, lg2_approx(s + 1e-6f), -blend_poly(s, P.R2, P.q0, P.q1, P.q2, P.q3));
            Pc = ffma2(pk2(c.z, c.w), pk2(psi, psi), Pc);
          }
        }
        // ---- floor / fraction, in-image test
        const u64 mg = pk2(L3_MAGIC, L3_MAGIC);
        const u64 Tm = fadd2_rm(Pc, mg);
        const u64 Fr = fsub2(Pc, fsub2(Tm, mg));
        float px, py, fx, fy, tx, ty;
        upk2(Pc, px, py); upk2(Fr, fx, fy); upk2(Tm, tx, ty);
        const bool inb = px >= lox && px < hix && py >= loy && py < hiy;
        if (__any_sync(0xffffffffu, inb)) {
          const int xi = __float_as_int(tx) - L3_MAGIC_BITS, yi = __float_as_int(ty) - L3_MAGIC_BITS;
          const int off = yi * IW + xi;
          const int d = off - prev_off;
          const bool same = d == 0, shift = d == IW;
          prev_off = inb ? off : (int)L3_INVALID;
          const float* p = f32_at_s(img, off);
          // walking down a column the bottom tap row of the previous pixel is normally the top tap row of this one
          if (shift) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { t0[c] = b0[c]; t1[c] = b1[c]; }
          }
          const bool needB = inb && !same, needT = needB && !shift;
          l3_load_row<PLANE_B, ROW_B>(b0, b1, p, needB);
          if (__any_sync(0xffffffffu, needT)) l3_load_row<PLANE_B, 0>(t0, t1, p, needT);
#if L4_PF_ROWS > 0
          l3_prefetch<PLANE_B, (1 + L4_PF_ROWS) * ROW_B>(p, needB && off < pf_limit);
#endif
          // bilinear interpolation, lerp form: planes 0/1 packed, plane 2 packed over (top, bottom)
          const u64 fx2 = pk2(fx, fx), fy2 = pk2(fy, fy);
          const u64 A = pk2(t0[0], t0[1]), Cc = pk2(t1[0], t1[1]);
          const u64 B = pk2(b0[0], b0[1]), D = pk2(b1[0], b1[1]);
          const u64 top = ffma2(fx2, fsub2(Cc, A), A), bot = ffma2(fx2, fsub2(D, B), B);
          float r0, r1, tp2, bt2;
          upk2(ffma2(fy2, fsub2(bot, top), top), r0, r1);
          const u64 L = pk2(t0[2], b0[2]), R = pk2(t1[2], b1[2]);
          upk2(ffma2(fx2, fsub2(R, L), L), tp2, bt2);
          const float r2 = fmaf(fy, bt2 - tp2, tp2);
          o0 = inb ? r0 : 0.f;
          o1 = inb ? r1 : 0.f;
          o2 = inb ? r2 : 0.f;
        } else {
          prev_off = (int)L3_INVALID;
        }
      }
      xbuf[v][r][0][colt] = o0;
      xbuf[v][r][1][colt] = o1;
      xbuf[v][r][2][colt] = o2;
    }
    // ---- both views of this strip's cell are in shared memory: each warp blends + stores half of the rows
    l4_pair_barrier(1 + cw);
    if (active) {
#pragma unroll
      for (int rr = 0; rr < HALF; ++rr) {
        const int r = v * HALF + rr, row = row0 + r;
        if (row < row_end) {
          const float a0 = xbuf[0][r][0][colt], a1 = xbuf[0][r][1][colt], a2 = xbuf[0][r][2][colt];
          const float c0 = xbuf[1][r][0][colt], c1 = xbuf[1][r][1][colt], c2 = xbuf[1][r][2][colt];
          // AVERAGE blend (a*a + b*b) / (a + b + 1e-6), streaming stores
          float s0, s1, q0, q1;
          upk2(fadd2(fadd2(pk2(a0, a1), pk2(c0, c1)), pk2(1e-6f, 1e-6f)), s0, s1);
          upk2(ffma2(pk2(c0, c1), pk2(c0, c1), fmul2(pk2(a0, a1), pk2(a0, a1))), q0, q1);
          const unsigned opix = (unsigned)(row * P.Wo + col);
          __stcs(const_cast<float*>(f32_at(outp, opix)), q0 * rcp_approx(s0));
          __stcs(const_cast<float*>(f32_at(outp, opix + oplane)), q1 * rcp_approx(s1));
          __stcs(const_cast<float*>(f32_at(outp, opix + 2u * oplane)), blend_avg_fast(a2, c2));
        }
      }
    }
    l4_pair_barrier(1 + cw);   // the strip's buffer may be overwritten
  }
}
