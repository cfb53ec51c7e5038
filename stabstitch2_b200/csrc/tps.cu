// Thin-plate-spline kernels: system solve, point evaluation, dense resampler, and the fused
// resample + AVERAGE-blend kernel (the "warp kernel" of BASELINE.json's metric).
//
// Reference behaviour restated (paths under Full_model_inference/Codes/):
//   utils/torch_tps_transform.py:168-226  _solve_system  -> tps_solve_kernel
//   utils/torch_tps_transform.py:108-149  _meshgrid + T x grid -> tps field evaluation
//   utils/torch_tps_transform.py:30-106   _interpolate (NORMAL) -> sample_normal
//   utils/torch_tps_transform.py:158-162  F.grid_sample(align_corners=True) (FAST) -> sample_fast
//   utils/torch_tps_transform_point.py    -> tps_point_kernel
//   test_online_tra.py:142                AVERAGE fusion -> blend_avg
#include <cuda.h>

#include "common.cuh"

#define LN2F 0.69314718055994530942f

// ------------------------------------------------------------------------------------------
// 66x66 fp64 solve, one CTA per system.  The reference inverts W explicitly in fp64 and
// multiplies by the targets; solving W T = tp by Gauss-Jordan with partial pivoting in fp64
// gives the same fp32-rounded coefficients (fp64 noise is ~1e-9 of an fp32 ulp here).
//
// Latency-oriented (64 systems per 32-frame chunk: the chip is never full, so the time is the
// length of one system's dependency chain on ONE SM: 66 pivot steps).  Measured: a step costs what
// the busiest warp's INSTRUCTION STREAM costs (~7 cycles per dependent instruction; DFMA issues at
// 59 lanes per clock per SM on B200, the fp64 pipe is not the limit), so the layout minimises the
// instructions of the warp on the critical path.  The augmented matrix [W | tp] (66 x 68) lives in
// REGISTERS, rows across LANES and columns across WARPS: thread (w, l) owns rows l + 32 i (i < 3)
// and columns w + NW j (NW = 8 warps: j < 9).  With that layout
//   * the pivot row's elements a warp needs are its OWN columns of row p: one shuffle from lane
//     p % 32, nothing travels between warps; the row slot of p is resolved by ONE three-way branch
//     per step;
//   * column k+1 (the multipliers of the next step and the candidates of its pivot search) sits
//     in ONE warp, all 66 rows: that warp alone eliminates that column first, searches (redux.sync
//     max over a monotone key, the reciprocal of every lane's own best candidate computed while the
//     reduction is in flight) and publishes pivot row, reciprocal and the column through shared
//     memory; no atomics, no work in the other warps;
//   * ONE named barrier per step, placed right after the search: a warp updates its remaining
//     columns AFTER the barrier, i.e. the searcher's bulk work overlaps the next step, in which another
//     warp searches;
//   * columns <= k are dead (never read again): the phase loop below skips their slots.
// Rows are never swapped (the pivot row of step k is remembered).  The arithmetic per element is the
// same sequence of fused multiply-adds as in the previous layouts: T is bit-identical to theirs.
// History (64 systems): matrix in shared memory 97 us; registers with rows across warps, an
// atomicMax search in every warp and two barriers per step 53 us; this layout 24 us (a variant
// without any barrier - warps spinning on a step counter in shared memory - measured the same and
// is flagged by racecheck, so the barrier stays).
// A further warp computes the affine predictor of the lattice resampler and leaves.
// ------------------------------------------------------------------------------------------
#ifndef SOLVE_NW
#define SOLVE_NW 8                       // solver warps per system = column stride
#endif
#define SOLVE_THREADS (SOLVE_NW * 32 + 32)
#define SOLVE_RS 3                       // row slots per thread: lane, lane + 32, lane + 64
#define SOLVE_CS ((SS2_NSYS + 2 + SOLVE_NW - 1) / SOLVE_NW)   // column slots per thread: warp + SOLVE_NW j
#define SOLVE_RING 32                    // >= 2 SOLVE_NW
#define AUG (SS2_NSYS + 2)

__device__ __forceinline__ unsigned solve_key(double v, int row) {
  // monotone in |v| (sign stripped, exponent + 13 mantissa bits), row in the low 7 bits
  const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu;
  return ((hi >> 6) << 7) | (unsigned)row;
}

__device__ __forceinline__ void solve_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SOLVE_NW * 32) : "memory"); }

// 1 / x to fp64 rounding noise: MUFU.RCP64H (2^-23) + two Newton steps, a chain of one MUFU and four DFMAs
// (the IEEE-exact __drcp_rn adds fix-up code to a chain that sits on the critical path of every pivot step)
__device__ __forceinline__ double solve_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

struct SolveShared {
  double colk[SOLVE_RING][AUG];   // published columns (the multipliers of step k = column k of every row), ring by k % 32
  double pinv[SS2_NSYS];          // 1 / pivot of step k
  double rhs[SS2_NSYS][2];
  int perm[SS2_NSYS];             // pivot row of step k
  float sx[SS2_NPT_PAD], sy[SS2_NPT_PAD];
};

// Executed by the warp that owns column cn (values v[i] of rows lane + 32 i, already eliminated up to
// step cn - 1): pick the pivot of step cn among the rows not used yet and publish the column (= the multipliers of
// step cn, with a zero in the pivot row, which is not eliminated), the pivot row and the reciprocal pivot.
__device__ __forceinline__ void solve_search(const double v0, const double v1, const double v2, unsigned used, int lane, int cn,
                                             SolveShared& sh) {
  const unsigned k0 = (used & 1u) ? 0u : solve_key(v0, lane);
  const unsigned k1 = (used & 2u) ? 0u : solve_key(v1, lane + 32);
  const unsigned k2 = (lane + 64 < SS2_NSYS && !(used & 4u)) ? solve_key(v2, lane + 64) : 0u;
  const unsigned mine = max(k0, max(k1, k2));
  const unsigned best = __reduce_max_sync(0xffffffffu, mine);
  // every lane inverts its own best candidate while the reduction is in flight; the winner publishes
  const double inv = solve_rcp(mine == k0 ? v0 : (mine == k1 ? v1 : v2));
  double* col = sh.colk[cn & (SOLVE_RING - 1)];
  col[lane] = v0;
  col[lane + 32] = v1;
  if (lane + 64 < SS2_NSYS) col[lane + 64] = v2;
  if (mine == best && (best != 0u || lane == 0)) {
    if (best != 0u) col[best & 127u] = 0.0;   // the pivot row is not eliminated (after this lane's own store above: same warp, in order)
    sh.perm[cn] = (int)(best & 127u);
    sh.pinv[cn] = inv;
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS)
tps_solve_kernel(const float* __restrict__ source, const float* __restrict__ target, float* __restrict__ Tout,
                 float* __restrict__ aux, float half_w, float half_h, float kx, float ky) {
  __shared__ SolveShared sh;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* src = source + (size_t)b * SS2_NPT * 2;
  const float* tgt = target + (size_t)b * SS2_NPT * 2;
  if (tid < SS2_NPT) {
    sh.sx[tid] = src[2 * tid];
    sh.sy[tid] = src[2 * tid + 1];
  }
  __syncthreads();
  if (wid == SOLVE_NW) {
    // Affine predictor for the lattice resampler: least-squares fit target ~ a*sx + b*sy + c over
    // the 63 control points, expressed in source PIXEL units as a function of the canvas pixel
    // index (col,row): px = aux[0]*col + aux[1]*row + aux[2], py = aux[3]*col + aux[4]*row + aux[5].
    // Interpolation reproduces affine functions exactly, so any affine predictor is
    // mathematically neutral; it only keeps the interpolated residuals small (fp32 rounding).
    // Partial sums per lane in control-point order, fixed shuffle tree (deterministic).
    if (!aux) return;
    double S[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) S[q] = 0.0;
    for (int i = lane; i < SS2_NPT; i += 32) {
      const double x = src[2 * i], y = src[2 * i + 1], u = tgt[2 * i], v = tgt[2 * i + 1];
      S[0] += x * x; S[1] += x * y; S[2] += x; S[3] += y * y; S[4] += y; S[5] += 1.0;
      S[6] += x * u; S[7] += y * u; S[8] += u;
      S[9] += x * v; S[10] += y * v; S[11] += v;
    }
#pragma unroll
    for (int q = 0; q < 12; ++q)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) S[q] += __shfl_xor_sync(0xffffffffu, S[q], o);
    if (lane == 0) {
      const double* Bx = S + 6;
      const double* By = S + 9;
      // normal equations [[S0 S1 S2],[S1 S3 S4],[S2 S4 S5]] p = B, Cramer's rule
      const double m00 = S[3] * S[5] - S[4] * S[4], m01 = S[1] * S[5] - S[4] * S[2], m02 = S[1] * S[4] - S[3] * S[2];
      const double det = S[0] * m00 - S[1] * m01 + S[2] * m02;
      double sol[2][3] = {{0, 0, 0}, {0, 0, 0}};
      if (fabs(det) > 1e-30) {
        const double inv[3][3] = {
            {m00 / det, -m01 / det, m02 / det},
            {-m01 / det, (S[0] * S[5] - S[2] * S[2]) / det, -(S[0] * S[4] - S[1] * S[2]) / det},
            {m02 / det, -(S[0] * S[4] - S[1] * S[2]) / det, (S[0] * S[3] - S[1] * S[1]) / det}};
        for (int r = 0; r < 3; ++r) {
          sol[0][r] = inv[r][0] * Bx[0] + inv[r][1] * Bx[1] + inv[r][2] * Bx[2];
          sol[1][r] = inv[r][0] * By[0] + inv[r][1] * By[1] + inv[r][2] * By[2];
        }
      }
      float* o = aux + (size_t)b * 8;
      // s = -1 + k*idx  ->  pix = half * (a*kx*col + b*ky*row + (c + 1 - a - b))
      o[0] = (float)(half_w * sol[0][0] * kx); o[1] = (float)(half_w * sol[0][1] * ky);
      o[2] = (float)(half_w * (sol[0][2] + 1.0 - sol[0][0] - sol[0][1]));
      o[3] = (float)(half_h * sol[1][0] * kx); o[4] = (float)(half_h * sol[1][1] * ky);
      o[5] = (float)(half_h * (sol[1][2] + 1.0 - sol[1][0] - sol[1][1]));
      o[6] = 0.f; o[7] = 0.f;
    }
    return;
  }
  // assemble this thread's elements of [W | tp] (fp32 arithmetic for K exactly like the reference, then widened)
  double a[SOLVE_RS][SOLVE_CS];
#pragma unroll
  for (int i = 0; i < SOLVE_RS; ++i) {
    const int r = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < SOLVE_CS; ++j) {
      const int c = wid + SOLVE_NW * j;
      double v = 0.0;
      if (r < SS2_NSYS && c < AUG) {
        if (r < SS2_NPT) {
          if (c == 0) v = 1.0;
          else if (c == 1) v = (double)sh.sx[r];
          else if (c == 2) v = (double)sh.sy[r];
          else if (c < SS2_NSYS) {
            const int q = c - 3;
            const float dx = __fsub_rn(sh.sx[r], sh.sx[q]);
            const float dy = __fsub_rn(sh.sy[r], sh.sy[q]);
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            v = (double)__fmul_rn(d2, logf(__fadd_rn(d2, 1e-6f)));
          } else {
            v = (double)tgt[2 * r + (c - SS2_NSYS)];
          }
        } else if (c >= 3 && c < SS2_NSYS) {
          const int q = r - SS2_NPT, pt = c - 3;  // 0: ones, 1: x, 2: y
          v = q == 0 ? 1.0 : (q == 1 ? (double)sh.sx[pt] : (double)sh.sy[pt]);
        }
      }
      a[i][j] = v;
    }
  }
  unsigned used = 0u;  // bit i: this lane's row lane + 32 i has been a pivot row (the same in every warp)
  if (wid == 0) solve_search(a[0][0], a[1][0], a[2][0], used, lane, 0, sh);
  solve_bar();
  // Step k eliminates column k; the warp that owns column cn = k + 1 (warp cn % NW, slot cn / NW) searches next.
  // Phase jn = cn / NW is unrolled: within it the slots below jn are dead (columns <= k are never read again), the slots
  // above are live, slot jn is live in the warps >= cn % NW: one dynamic predicate per step and no work on dead columns.
  // pr = the pivot row's element of this column (row p = lane pl, row slot PS: the slot is resolved ONCE per step by
  // the three-way branch below, the per-warp instruction stream being what bounds a step)
#define SOLVE_ELIM(J, PS)                                                                     \
    {                                                                                         \
      const double pr = __shfl_sync(0xffffffffu, a[PS][J], pl);                               \
      a[0][J] = fma(-g0, pr, a[0][J]); a[1][J] = fma(-g1, pr, a[1][J]); a[2][J] = fma(-g2, pr, a[2][J]); \
    }
#define SOLVE_STEP(PS)                                                                        \
    {                                                                                         \
      if (wid >= kk) SOLVE_ELIM(jn, PS)                                                       \
      if (wid == kk && cn < SS2_NSYS) solve_search(a[0][jn], a[1][jn], a[2][jn], used, lane, cn, sh); \
      solve_bar();                                                                            \
      _Pragma("unroll") for (int j = jn + 1; j < SOLVE_CS; ++j) SOLVE_ELIM(j, PS)             \
    }
#pragma unroll
  for (int jn = 0; jn < SOLVE_CS; ++jn) {
    const int kk1 = SS2_NSYS + 1 - jn * SOLVE_NW < SOLVE_NW ? SS2_NSYS + 1 - jn * SOLVE_NW : SOLVE_NW;   // cn <= 66
#pragma unroll 1
    for (int kk = (jn == 0 ? 1 : 0); kk < kk1; ++kk) {
      const int cn = jn * SOLVE_NW + kk, k = cn - 1;
      const int p = sh.perm[k];
      const double inv = sh.pinv[k];
      const double* col = sh.colk[k & (SOLVE_RING - 1)];
      // multipliers of this lane's rows, already zero in the pivot row; padding rows ride along with zero
      const double g0 = col[lane] * inv, g1 = col[lane + 32] * inv, g2 = lane + 64 < SS2_NSYS ? col[lane + 64] * inv : 0.0;
      const int ps = p >> 5, pl = p & 31;
      used |= (lane == pl ? 1u : 0u) << ps;   // the pivot row is not a candidate any more
      if (ps == 0) SOLVE_STEP(0) else if (ps == 1) SOLVE_STEP(1) else SOLVE_STEP(2)   // one barrier inside, CTA-uniform branch
    }
  }
#undef SOLVE_STEP
#undef SOLVE_ELIM
  // right-hand sides: columns 66, 67 (warps 66 % NW and 67 % NW, slot 66 / NW: the same slot, NW is a power of two >= 4)
  if (wid == (SS2_NSYS & (SOLVE_NW - 1)) || wid == ((SS2_NSYS + 1) & (SOLVE_NW - 1))) {
    const int c = wid - (SS2_NSYS & (SOLVE_NW - 1));
#pragma unroll
    for (int i = 0; i < SOLVE_RS; ++i)
      if (lane + 32 * i < SS2_NSYS) sh.rhs[lane + 32 * i][c] = a[i][SS2_NSYS / SOLVE_NW];
  }
  solve_bar();
  for (int e = tid; e < 2 * SS2_NSYS; e += SOLVE_NW * 32) {
    const int c = e / SS2_NSYS, j = e % SS2_NSYS;
    Tout[(size_t)b * 2 * SS2_NSYS + c * SS2_NSYS + j] = (float)(sh.rhs[sh.perm[j]][c] * sh.pinv[j]);
  }
}

int tps_solve_launch(ss2_ctx* ctx, const float* d_source, const float* d_target, int bn, float* d_T,
                     cudaStream_t st) {
  if (bn <= 0) return SS2_OK;
  tps_solve_kernel<<<bn, SOLVE_THREADS, 0, st>>>(d_source, d_target, d_T, nullptr, 0.f, 0.f, 0.f, 0.f);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// solve + affine predictor (aux [bn][8]) for the lattice resampler: (half_w, half_h) convert the
// normalised source coordinate to pixels, (Wo, Ho) is the canvas the dense grid spans
int tps_solve_aux_launch(ss2_ctx* ctx, const float* d_source, const float* d_target, int bn, float* d_T, float* d_aux,
                         float half_w, float half_h, int Ho, int Wo, cudaStream_t st) {
  if (bn <= 0) return SS2_OK;
  const float kx = Wo > 1 ? (float)(2.0 / (Wo - 1)) : 0.f, ky = Ho > 1 ? (float)(2.0 / (Ho - 1)) : 0.f;
  tps_solve_kernel<<<bn, SOLVE_THREADS, 0, st>>>(d_source, d_target, d_T, d_aux, half_w, half_h, kx, ky);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// point evaluation (63 points per system), fp32 with full-precision logf
// ------------------------------------------------------------------------------------------
__global__ void tps_point_kernel(const float* __restrict__ point, const float* __restrict__ source,
                                 const float* __restrict__ T, float* __restrict__ out) {
  __shared__ float sx[SS2_NPT], sy[SS2_NPT], tx[SS2_NSYS], ty[SS2_NSYS];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < SS2_NPT) {
    sx[tid] = source[((size_t)b * SS2_NPT + tid) * 2];
    sy[tid] = source[((size_t)b * SS2_NPT + tid) * 2 + 1];
  }
  for (int i = tid; i < SS2_NSYS; i += blockDim.x) {
    tx[i] = T[(size_t)b * 2 * SS2_NSYS + i];
    ty[i] = T[(size_t)b * 2 * SS2_NSYS + SS2_NSYS + i];
  }
  __syncthreads();
  if (tid >= SS2_NPT) return;
  const float x = point[((size_t)b * SS2_NPT + tid) * 2];
  const float y = point[((size_t)b * SS2_NPT + tid) * 2 + 1];
  float ax = tx[0] + tx[1] * x + tx[2] * y;
  float ay = ty[0] + ty[1] * x + ty[2] * y;
  for (int i = 0; i < SS2_NPT; ++i) {
    float dx = x - sx[i], dy = y - sy[i];
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    float r = __fmul_rn(d2, logf(__fadd_rn(d2, 1e-6f)));
    ax = fmaf(tx[3 + i], r, ax);
    ay = fmaf(ty[3 + i], r, ay);
  }
  out[((size_t)b * SS2_NPT + tid) * 2] = ax;
  out[((size_t)b * SS2_NPT + tid) * 2 + 1] = ay;
}

int tps_point_launch(ss2_ctx* ctx, const float* d_point, const float* d_source, const float* d_T, int bn,
                     float* d_out, cudaStream_t st) {
  if (bn <= 0) return SS2_OK;
  tps_point_kernel<<<bn, 96, 0, st>>>(d_point, d_source, d_T, d_out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// ------------------------------------------------------------------------------------------
// dense resampler
// ------------------------------------------------------------------------------------------
// torch.linspace(-1, 1, n)[i]: symmetric two-sided formula of ATen's range factories.
__device__ __forceinline__ float lin11(int i, int n, float step) {
  return (i < n / 2) ? __fadd_rn(-1.0f, __fmul_rn(step, (float)i))
                     : __fsub_rn(1.0f, __fmul_rn(step, (float)(n - 1 - i)));
}

// _interpolate (NORMAL): clamped 4-tap gather with weights taken from the CLAMPED integer
// coordinates, separate multiplies and adds in the reference's order, so that out-of-image
// samples cancel to the same kind of rounding residue the reference produces.
template <int C>
__device__ __forceinline__ void sample_normal(const float* __restrict__ img, int H, int W, float x, float y,
                                              float (&out)[C]) {
  const float fx = floorf(x), fy = floorf(y);
  // float->int conversion saturates like torch's .int() for finite values
  int x0 = (int)fminf(fmaxf(fx, -2.0e9f), 2.0e9f);
  int y0 = (int)fminf(fmaxf(fy, -2.0e9f), 2.0e9f);
  int x1 = min(max(x0 + 1, 0), W - 1);
  int y1 = min(max(y0 + 1, 0), H - 1);
  x0 = min(max(x0, 0), W - 1);
  y0 = min(max(y0, 0), H - 1);
  const float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;
  const float wa = __fmul_rn(__fsub_rn(x1f, x), __fsub_rn(y1f, y));
  const float wb = __fmul_rn(__fsub_rn(x1f, x), __fsub_rn(y, y0f));
  const float wc = __fmul_rn(__fsub_rn(x, x0f), __fsub_rn(y1f, y));
  const float wd = __fmul_rn(__fsub_rn(x, x0f), __fsub_rn(y, y0f));
  const int ia = y0 * W + x0, ib = y1 * W + x0, ic = y0 * W + x1, id = y1 * W + x1;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* p = img + c * plane;
    const float Ia = __ldg(p + ia), Ib = __ldg(p + ib), Ic = __ldg(p + ic), Id = __ldg(p + id);
    out[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, Ia), __fmul_rn(wb, Ib)), __fmul_rn(wc, Ic)),
                       __fmul_rn(wd, Id));
  }
}

// F.grid_sample(bilinear, padding_mode='zeros', align_corners=True)
template <int C>
__device__ __forceinline__ void sample_fast(const float* __restrict__ img, int H, int W, float x, float y,
                                            float (&out)[C]) {
  const float fx = floorf(x), fy = floorf(y);
  const int x0 = (int)fminf(fmaxf(fx, -2.0e9f), 2.0e9f), y0 = (int)fminf(fmaxf(fy, -2.0e9f), 2.0e9f);
  const int x1 = x0 + 1, y1 = y0 + 1;
  const float nw = (fx + 1.0f - x) * (fy + 1.0f - y);
  const float ne = (x - fx) * (fy + 1.0f - y);
  const float sw = (fx + 1.0f - x) * (y - fy);
  const float se = (x - fx) * (y - fy);
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float* p = img + c * plane;
    float acc = 0.0f;
    if (vy0 && vx0) acc += __ldg(p + y0 * W + x0) * nw;
    if (vy0 && vx1) acc += __ldg(p + y0 * W + x1) * ne;
    if (vy1 && vx0) acc += __ldg(p + y1 * W + x0) * sw;
    if (vy1 && vx1) acc += __ldg(p + y1 * W + x1) * se;
    out[c] = acc;
  }
}

__device__ __forceinline__ float blend_avg(float a, float b) {
  // a*(a/(a+b+1e-6)) + b*(b/(a+b+1e-6)), test_online_tra.py:142 (no contraction)
  const float s = __fadd_rn(__fadd_rn(a, b), 1e-6f);
  return __fadd_rn(__fmul_rn(a, __fdiv_rn(a, s)), __fmul_rn(b, __fdiv_rn(b, s)));
}

// ------------------------------------------------------------------------------------------
// EXACT field: all 63 radial terms per pixel in the reference's arithmetic (fp32, accurate
// logf, unfused d2), accumulated in control-point order.  This is the validation mode and the
// generic utils.torch_tps_transform.transformer path; the production path is the lattice
// resampler below.
// Tile geometry: a CTA of TX x TY threads covers TX x (TY*RPT) canvas pixels; each thread owns
// RPT pixels of one column (rows r, r+TY, ...), so dx and dx^2 are shared between them.
// ------------------------------------------------------------------------------------------
#define TX 32
#define TY 4
#define RPT 2
#define TILE_H (TY * RPT)

struct WarpParams {
  const float* img[4];   // per view (up to 4): base of [n][C][H][W]
  const float* source;   // [n][V][63][2]
  const float* T;        // [n][V][2][66]
  float* out;            // BLEND: [n][C][Ho][Wo]; else [n*V][C][Ho][Wo]
  unsigned char* out8;   // fused pair kernel with the uint8 back end folded in: [n][Ho][Wo][3] (else unused)
  int H, W, Ho, Wo;
  float stepx, stepy;
  // lattice mode
  const float* aux;      // [n][V][8] affine predictor
  const float2* nodes;   // [n][ny][nx][V] residual source pixel coordinates at the lattice nodes
  int nx, ny;
  float R2, p0, p1, p2, p3;  // near radius^2 (normalised units) and the blending cubic P(s)
  float q0, q1, q2, q3;      // P / ln2
  float half_w, half_h;      // normalised -> pixel scale of the source image
};

// V = views evaluated per pixel (2 for the fused blend, 1 for the generic transformer).
template <int V, int C, int MODE, bool BLEND>
__global__ void __launch_bounds__(TX* TY)
tps_warp_exact_kernel(WarpParams P) {
  __shared__ float4 cp[V][SS2_NPT_PAD];  // (px, py, wx, wy)
  __shared__ float aff[V][6];
  const int n = blockIdx.z;
  const int tid = threadIdx.y * TX + threadIdx.x;
  for (int i = tid; i < V * SS2_NPT_PAD; i += TX * TY) {
    const int v = i / SS2_NPT_PAD, j = i % SS2_NPT_PAD;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < SS2_NPT) {
      const float* s = P.source + ((size_t)(n * V + v) * SS2_NPT + j) * 2;
      const float* t = P.T + (size_t)(n * V + v) * 2 * SS2_NSYS;
      c = make_float4(s[0], s[1], t[3 + j], t[SS2_NSYS + 3 + j]);
    }
    cp[v][j] = c;
  }
  if (tid < V * 6) {
    const int v = tid / 6, k = tid % 6;
    aff[v][k] = P.T[(size_t)(n * V + v) * 2 * SS2_NSYS + (k / 3) * SS2_NSYS + (k % 3)];
  }
  __syncthreads();
  const int col = blockIdx.x * TX + threadIdx.x;
  const int row0 = blockIdx.y * TILE_H + threadIdx.y;
  if (col >= P.Wo) return;
  const float xt = lin11(col, P.Wo, P.stepx);
  float yt[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) yt[r] = lin11(min(row0 + r * TY, P.Ho - 1), P.Ho, P.stepy);

  float res[RPT][V][C];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float ax[RPT], ay[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      ax[r] = fmaf(aff[v][2], yt[r], fmaf(aff[v][1], xt, aff[v][0]));
      ay[r] = fmaf(aff[v][5], yt[r], fmaf(aff[v][4], xt, aff[v][3]));
    }
#pragma unroll 3
    for (int i = 0; i < SS2_NPT; ++i) {
      const float4 c = cp[v][i];
      const float dx = __fsub_rn(xt, c.x);
      const float dx2 = __fmul_rn(dx, dx);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float dy = __fsub_rn(yt[r], c.y);
        const float d2 = __fadd_rn(dx2, __fmul_rn(dy, dy));
        const float rr = __fmul_rn(d2, logf(__fadd_rn(d2, 1e-6f)));
        ax[r] = fmaf(c.z, rr, ax[r]);
        ay[r] = fmaf(c.w, rr, ay[r]);
      }
    }
    const float* img = P.img[BLEND ? v : 0] + (size_t)(BLEND ? n : n * V + v) * C * P.H * P.W;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      if (MODE == SS2_MODE_NORMAL) {
        const float x = __fmul_rn(__fmul_rn(__fadd_rn(ax[r], 1.0f), (float)P.W), 0.5f);
        const float y = __fmul_rn(__fmul_rn(__fadd_rn(ay[r], 1.0f), (float)P.H), 0.5f);
        sample_normal<C>(img, P.H, P.W, x, y, res[r][v]);
      } else {
        const float x = (ax[r] + 1.0f) * 0.5f * (float)(P.W - 1);
        const float y = (ay[r] + 1.0f) * 0.5f * (float)(P.H - 1);
        sample_fast<C>(img, P.H, P.W, x, y, res[r][v]);
      }
    }
  }
  const size_t plane = (size_t)P.Ho * P.Wo;
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int row = row0 + r * TY;
    if (row >= P.Ho) continue;
    if (BLEND) {
      float* o = P.out + (size_t)n * C * plane + (size_t)row * P.Wo + col;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        // fuse(..fuse(fuse(1,2),3)..,V) (test_online_tra_threeview.py:489-490 for V = 3); V = 1: the view itself
        float f = res[r][0][c];
#pragma unroll
        for (int v = 1; v < V; ++v) f = blend_avg(f, res[r][v][c]);
        __stcs(o + c * plane, f);
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float* o = P.out + (size_t)(n * V + v) * C * plane + (size_t)row * P.Wo + col;
#pragma unroll
        for (int c = 0; c < C; ++c) __stcs(o + c * plane, res[r][v][c]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// LATTICE resampler (production path).
//
// The TPS field f(p) = affine(p) + sum_i w_i phi(|p-c_i|^2), phi(s) = s log(s + 1e-6), is split
//     phi = phi_far + psi,   phi_far(s) = phi(s) for s >= R^2, = P(s) for s < R^2,
// with P the cubic Taylor polynomial of phi at s = R^2 (phi_far is C^3 with derivatives of the
// size phi has at distance R), and psi = phi - P compactly supported in the disc s < R^2.
//   * tps_nodes_kernel evaluates affine + sum_i w_i phi_far exactly (accurate logf, fp64
//     accumulation) on a lattice with SX x SY pixel spacing and stores it as residual source
//     PIXEL coordinates w.r.t. the affine predictor of tps_solve_kernel (small magnitudes, so
//     the fp32 interpolation below rounds at the 1e-6 px level).
//   * tps_warp_lattice_kernel interpolates the lattice with tensor-product QUINTIC Lagrange
//     weights (6x6 nodes): the y contraction is done once per (row, node column) of the CTA's
//     tile into shared memory, the x contraction is 6 FMAs per coordinate per pixel; it then
//     adds sum w_i psi for the few control points whose disc touches the tile (1 MUFU lg2
//     each), samples, blends and stores.
// Interpolation error ~ c h^6 |f_far^(6)|: with (SX,SY,R) of lattice_config() the source
// coordinate is as close to the fp64 arbiter as the reference's own fp32 evaluation
// (tests/test_gpu_parity.py::test_fullsize_frame_vs_oracle_and_arbiter).
// Out-of-image samples are exactly 0 here (the reference's clamped taps cancel to a rounding
// residue of a few 1e-3 grey levels; DESIGN.md "OOB residue").
// ------------------------------------------------------------------------------------------
#define LAT_THREADS 128
#define LAT_TAPS 6
#define LAT_LO 2  // lattice index of array slot 0 is -LAT_LO

__device__ __forceinline__ float blend_poly(float s, float R2, float p0, float p1, float p2, float p3) {
  const float u = s - R2;
  return fmaf(u, fmaf(u, fmaf(u, p3, p2), p1), p0);
}

// Natural logarithm of a positive NORMAL float (here 1e-6 <= x < 16), <= 0.86 ulp (checked against fp64 over 3e6
// samples): x = m * 2^e with m in [2/3, 4/3), log m = f - f^2/2 + f^3 Q(f), f = m - 1, Q a degree-7 minimax fit.
// Inline and branch-free, ~16 instructions; libdevice's logf carries denormal / inf / nan handling the lattice does
// not need.
__device__ __forceinline__ float log_pos_normal(float x) {
  const int i = __float_as_int(x) - 0x3f2aaaab;
  const int e = i >> 23;
  const float f = __int_as_float(__float_as_int(x) - (e << 23)) - 1.0f;
  float q = -0.13346314430236816f;
  q = fmaf(q, f, 0.1415630429983139f);
  q = fmaf(q, f, -0.1207403615117073f);
  q = fmaf(q, f, 0.13967302441596985f);
  q = fmaf(q, f, -0.16688990592956543f);
  q = fmaf(q, f, 0.20013076066970825f);
  q = fmaf(q, f, -0.24999591708183289f);
  q = fmaf(q, f, 0.3333316147327423f);
  const float f2 = f * f;
  float r = fmaf(q, f2 * f, -0.5f * f2);
  r += f;
  return fmaf((float)e, 0.693147182f, r);
}

// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two lanes of work per issued instruction
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// log_pos_normal of both halves of a packed pair: the same operations in the same order on each half (bit-identical
// to the scalar function), the polynomial through packed FMAs
__device__ __forceinline__ u64 log_pos_normal2(u64 x) {
  float x0, x1;
  upk2(x, x0, x1);
  const int i0 = __float_as_int(x0) - 0x3f2aaaab, i1 = __float_as_int(x1) - 0x3f2aaaab;
  const int e0 = i0 >> 23, e1 = i1 >> 23;
  const u64 f = fadd2(pk2(__int_as_float(__float_as_int(x0) - (e0 << 23)), __int_as_float(__float_as_int(x1) - (e1 << 23))), pk2(-1.0f, -1.0f));
#define P2(c) pk2(c, c)
  u64 q = P2(-0.13346314430236816f);
  q = ffma2(q, f, P2(0.1415630429983139f));
  q = ffma2(q, f, P2(-0.1207403615117073f));
  q = ffma2(q, f, P2(0.13967302441596985f));
  q = ffma2(q, f, P2(-0.16688990592956543f));
  q = ffma2(q, f, P2(0.20013076066970825f));
  q = ffma2(q, f, P2(-0.24999591708183289f));
  q = ffma2(q, f, P2(0.3333316147327423f));
  const u64 f2 = fmul2(f, f);
  u64 r = ffma2(q, fmul2(f2, f), fmul2(P2(-0.5f), f2));
  r = fadd2(r, f);
  return ffma2(pk2((float)e0, (float)e1), P2(0.693147182f), r);
#undef P2
}

// one thread per (node, view): grid (node blocks, V, frames); nodes [n][ny][nx][V] float2 (x, y).
// Two control points per iteration through packed fp32 (same operations per point, in control-point order, as the
// scalar form: results are bit-identical to it); the 64th point is padding with zero weights.
template <int V>
__global__ void __launch_bounds__(128)
tps_nodes_kernel(WarpParams P, int SX, int SY, double kxd, double kyd) {
  __shared__ __align__(16) float4 cneg[SS2_NPT_PAD / 2];   // (-cx[2i], -cx[2i+1], -cy[2i], -cy[2i+1])
  __shared__ __align__(16) double2 cw[SS2_NPT_PAD];        // (wx, wy)
  __shared__ double aff[6];
  __shared__ float pred[6];
  const int n = blockIdx.z, v = blockIdx.y, tid = threadIdx.x;
  const float* t = P.T + (size_t)(n * V + v) * 2 * SS2_NSYS;
  if (tid < SS2_NPT_PAD) {
    float* cn = reinterpret_cast<float*>(cneg) + (tid >> 1) * 4 + (tid & 1);
    if (tid < SS2_NPT) {
      const float* s = P.source + ((size_t)(n * V + v) * SS2_NPT + tid) * 2;
      cn[0] = -s[0]; cn[2] = -s[1];
      cw[tid] = make_double2((double)t[3 + tid], (double)t[SS2_NSYS + 3 + tid]);
    } else {
      cn[0] = -3.0f; cn[2] = -3.0f;   // outside the normalised canvas: d2 > 0, weight 0
      cw[tid] = make_double2(0.0, 0.0);
    }
  } else if (tid < SS2_NPT_PAD + 6) {
    const int k = tid - SS2_NPT_PAD;
    aff[k] = (double)t[(k / 3) * SS2_NSYS + (k % 3)];
    pred[k] = P.aux[(size_t)(n * V + v) * 8 + k];
  }
  __syncthreads();
  const int node = blockIdx.x * blockDim.x + tid;
  if (node >= P.nx * P.ny) return;
  const int jy = node / P.nx, jx = node % P.nx;
  const int col = (jx - LAT_LO) * SX, row = (jy - LAT_LO) * SY;
  // normalised canvas coordinate of the node; (kxd, kyd) = 2 / (Wo - 1), 2 / (Ho - 1) from the host (no fp64 division here)
  const double xd = fma(kxd, (double)col, -1.0), yd = fma(kyd, (double)row, -1.0);
  const float xt = (float)xd, yt = (float)yd;
  double ax = aff[0] + aff[1] * xd + aff[2] * yd;
  double ay = aff[3] + aff[4] * xd + aff[5] * yd;
  const u64 xt2 = pk2(xt, xt), yt2 = pk2(yt, yt), eps2 = pk2(1e-6f, 1e-6f);
  const u64 nR2 = pk2(-P.R2, -P.R2), p0 = pk2(P.p0, P.p0), p1 = pk2(P.p1, P.p1), p2 = pk2(P.p2, P.p2), p3 = pk2(P.p3, P.p3);
#pragma unroll 4
  for (int i = 0; i < SS2_NPT_PAD / 2; ++i) {
    const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(&cneg[i]);
    const u64 dx = fadd2(xt2, c.x), dy = fadd2(yt2, c.y);
    const u64 d2 = ffma2(dy, dy, fmul2(dx, dx));
    // branch-free: both the far term d2 * log(d2 + 1e-6) and the blending cubic, then a select
    const u64 fl = fmul2(d2, log_pos_normal2(fadd2(d2, eps2)));
    const u64 u = fadd2(d2, nR2);
    const u64 fp = ffma2(u, ffma2(u, ffma2(u, p3, p2), p1), p0);
    float d20, d21, fl0, fl1, fp0, fp1;
    upk2(d2, d20, d21); upk2(fl, fl0, fl1); upk2(fp, fp0, fp1);
    const float f0 = d20 >= P.R2 ? fl0 : fp0, f1 = d21 >= P.R2 ? fl1 : fp1;
    const double2 w0 = cw[2 * i], w1 = cw[2 * i + 1];
    ax = fma(w0.x, (double)f0, ax);
    ay = fma(w0.y, (double)f0, ay);
    ax = fma(w1.x, (double)f1, ax);
    ay = fma(w1.y, (double)f1, ay);
  }
  const double px = (ax + 1.0) * (double)P.half_w - ((double)pred[0] * col + (double)pred[1] * row + (double)pred[2]);
  const double py = (ay + 1.0) * (double)P.half_h - ((double)pred[3] * col + (double)pred[4] * row + (double)pred[5]);
  const_cast<float2*>(P.nodes)[((size_t)n * P.ny * P.nx + node) * V + v] = make_float2((float)px, (float)py);
}

// quintic Lagrange weights for nodes at -2..3 and t = k/S, k = 0..S-1
struct LagrangeTable { float w[16][8]; };
__constant__ LagrangeTable c_lag[4];  // spacing 6, 8, 12, 16
__host__ __device__ constexpr int lag_idx(int S) { return S == 6 ? 0 : S == 8 ? 1 : S == 12 ? 2 : 3; }

// base + off*4 as ONE IMAD.WIDE.U32 (the compiler's generic 64-bit indexing costs 4-5 instructions)
__device__ __forceinline__ const float* f32_at(const float* base, unsigned off) {
  unsigned long long a;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(a) : "r"(off), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<const float*>(a);
}
__device__ __forceinline__ float lg2_approx(float x) {  // x is a normal positive number here
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// a*(a/s) + b*(b/s), s = a+b+1e-6 (test_online_tra.py:142) as (a*a + b*b) * (1/s); MUFU.RCP is
// accurate to ~1 ulp, i.e. <= 3e-5 grey levels here
__device__ __forceinline__ float blend_avg_fast(float a, float b) {
  const float s = (a + b) + 1e-6f;
  return fmaf(b, b, a * a) * rcp_approx(s);
}

// node-column slots a tile of LAT_THREADS pixel columns can touch
template <int SX> struct LatCols { static constexpr int value = (LAT_THREADS + SX - 1) / SX + LAT_TAPS; };

// IW, IH > 0: source image size known at compile time (all 12 tap addresses of a view become
// ONE 64-bit pointer + immediate offsets); 0 = runtime size.
#ifndef LAT_MINB
#define LAT_MINB 8
#endif
#ifndef LAT_RPI
#define LAT_RPI 1  // canvas rows whose taps are in flight together (2-3 measured slower: registers)
#endif
#ifndef LAT_TPI
#define LAT_TPI LAT_RPI  // canvas rows whose taps are in flight together (divides LAT_RPI)
#endif
#ifndef LAT_PREFETCH
#define LAT_PREFETCH 2  // source rows ahead pulled towards L1 (0/undefined: off)
#endif
#ifndef LAT_NCELL
#define LAT_NCELL 4
#endif
// LAT_NCELL: lattice cell rows per CTA: the tile is LAT_THREADS x (LAT_NCELL*SY) canvas pixels

// V = views evaluated per pixel: 1 (generic transformer), 2 (the production pair kernel), 3 / 4 (N-view fusion, config 5:
// one pass over the canvas reads every source once instead of warping each view to a temporary)
// U8: the driver's uint8 back end (frame.astype(uint8), test_online_tra.py:152,414) folded into the store of the fused pair
// kernel: the fp32 canvas (the largest buffer of the path) is neither written nor read back by a conversion pass.
template <int V, int C, int MODE, bool BLEND, int SX, int SY, int IW, int IH, bool U8 = false>
__global__ void __launch_bounds__(LAT_THREADS, V <= 2 ? LAT_MINB : (V == 3 ? 5 : 4))
tps_warp_lattice_kernel(WarpParams P) {
  constexpr int NCOL = LatCols<SX>::value;
  __shared__ float4 near_list[V * SS2_NPT];  // (cx, cy, wx_px*ln2, wy_px*ln2), sorted by view
  __shared__ int warp_cnt[LAT_THREADS / 32][2];
  __shared__ int view_end[V];                // V > 2: near_list entries of view v end at view_end[v]
  __shared__ float s_pred[V][6];
  __shared__ __align__(16) float2 ysm[SY][NCOL][V];  // y-contracted residuals of the current cell row
  const int n = blockIdx.z, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int col0 = blockIdx.x * LAT_THREADS, row00 = blockIdx.y * (LAT_NCELL * SY);
  const int jx0 = col0 / SX;  // node slot of the first cell (slot s <-> lattice column s - LAT_LO)
  const int W = IW > 0 ? IW : P.W, H = IH > 0 ? IH : P.H;
  // ---- near list: control points whose disc s < R2 touches this tile (deterministic order)
  int n0 = 0, n_all = 0;
  if (V <= 2) {
    float4 ent = make_float4(0.f, 0.f, 0.f, 0.f);
    bool hit = false;
    const int pv = tid / SS2_NPT, pi = tid - pv * SS2_NPT;
    if (tid < V * SS2_NPT) {
      const float x_lo = fmaf(P.stepx, (float)col0, -1.0f), x_hi = fmaf(P.stepx, (float)min(col0 + LAT_THREADS - 1, P.Wo - 1), -1.0f);
      const float y_lo = fmaf(P.stepy, (float)row00, -1.0f), y_hi = fmaf(P.stepy, (float)min(row00 + LAT_NCELL * SY - 1, P.Ho - 1), -1.0f);
      const float2 c = *reinterpret_cast<const float2*>(P.source + ((size_t)(n * V + pv) * SS2_NPT + pi) * 2);
      const float* t = P.T + (size_t)(n * V + pv) * 2 * SS2_NSYS;
      const float ddx = fmaxf(fmaxf(x_lo - c.x, c.x - x_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
      hit = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
      if (hit) ent = make_float4(c.x, c.y, t[3 + pi] * (P.half_w * LN2F), t[SS2_NSYS + 3 + pi] * (P.half_h * LN2F));
    }
    const unsigned m_all = __ballot_sync(0xffffffffu, hit), m_v0 = __ballot_sync(0xffffffffu, hit && pv == 0);
    if (lane == 0) { warp_cnt[wid][0] = __popc(m_all); warp_cnt[wid][1] = __popc(m_v0); }
    if (tid < V * 6) s_pred[tid / 6][tid % 6] = P.aux[(size_t)(n * V + tid / 6) * 8 + tid % 6];
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int w = 0; w < LAT_THREADS / 32; ++w)
      if (w < wid) off += warp_cnt[w][0];
    if (hit) near_list[off + __popc(m_all & ((1u << lane) - 1u))] = ent;
#pragma unroll
    for (int w = 0; w < LAT_THREADS / 32; ++w) { n_all += warp_cnt[w][0]; n0 += warp_cnt[w][1]; }
  } else {
    // one view after the other (63 points: warps 0 and 1), entries appended view by view
    if (tid < V * 6) s_pred[tid / 6][tid % 6] = P.aux[(size_t)(n * V + tid / 6) * 8 + tid % 6];
    for (int pv = 0; pv < V; ++pv) {
      float4 ent = make_float4(0.f, 0.f, 0.f, 0.f);
      bool hit = false;
      if (tid < SS2_NPT) {
        const float x_lo = fmaf(P.stepx, (float)col0, -1.0f), x_hi = fmaf(P.stepx, (float)min(col0 + LAT_THREADS - 1, P.Wo - 1), -1.0f);
        const float y_lo = fmaf(P.stepy, (float)row00, -1.0f), y_hi = fmaf(P.stepy, (float)min(row00 + LAT_NCELL * SY - 1, P.Ho - 1), -1.0f);
        const float2 c = *reinterpret_cast<const float2*>(P.source + ((size_t)(n * V + pv) * SS2_NPT + tid) * 2);
        const float* t = P.T + (size_t)(n * V + pv) * 2 * SS2_NSYS;
        const float ddx = fmaxf(fmaxf(x_lo - c.x, c.x - x_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
        hit = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
        if (hit) ent = make_float4(c.x, c.y, t[3 + tid] * (P.half_w * LN2F), t[SS2_NSYS + 3 + tid] * (P.half_h * LN2F));
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && wid < 2) warp_cnt[wid][0] = __popc(m);
      __syncthreads();
      const int c0 = warp_cnt[0][0], c1 = warp_cnt[1][0];
      if (hit) near_list[n_all + (wid == 1 ? c0 : 0) + __popc(m & ((1u << lane) - 1u))] = ent;
      n_all += c0 + c1;
      if (tid == 0) view_end[pv] = n_all;
      __syncthreads();
    }
  }
  // ---- per-column constants
  const int col = min(col0 + tid, P.Wo - 1);
  const bool active = col0 + tid < P.Wo;
  const int cxi = col / SX, rx = col - cxi * SX, js = cxi - jx0;  // this column's first node slot in ysm
  float lx[LAT_TAPS];
#pragma unroll
  for (int a = 0; a < LAT_TAPS; ++a) lx[a] = c_lag[lag_idx(SX)].w[rx][a];
  const float colf = (float)col;
  float pcol[V][2], prow[V][2];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    pcol[v][0] = fmaf(s_pred[v][0], colf, s_pred[v][2]);
    pcol[v][1] = fmaf(s_pred[v][3], colf, s_pred[v][5]);
    prow[v][0] = s_pred[v][1];
    prow[v][1] = s_pred[v][4];
  }
  const float xt = fmaf(P.stepx, colf, -1.0f);
  // x extent of this warp's 32 columns, for the per-warp candidate culling
  const float wx_lo = fmaf(P.stepx, (float)min(col0 + wid * 32, P.Wo - 1), -1.0f);
  const float wx_hi = fmaf(P.stepx, (float)min(col0 + wid * 32 + 31, P.Wo - 1), -1.0f);
  const unsigned W1 = (unsigned)(W - 1), H1 = (unsigned)(H - 1);
  const unsigned oplane = (unsigned)(P.Ho * P.Wo);
  const unsigned iplane = (unsigned)(H * W);
  // per-view frame bases (64-bit once); everything below indexes them with 32-bit offsets
  const float* imgv[V];
  const float* outv[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    imgv[v] = P.img[BLEND ? v : 0] + (size_t)(BLEND ? n : n * V + v) * C * iplane;
    outv[v] = P.out + (size_t)(BLEND ? n : n * V + v) * C * oplane;
  }
  const bool cull = n_all <= 32;  // one candidate per lane; longer lists (never seen) are not culled

  for (int cell = 0; cell < LAT_NCELL; ++cell) {
    const int row0 = row00 + cell * SY;
    if (row0 >= P.Ho) break;  // CTA-uniform
    const int cyi = blockIdx.y * LAT_NCELL + cell;
    __syncthreads();  // previous cell's readers of ysm are done (also orders near_list on the first pass)
    // ---- y contraction of the tile's node columns: ysm[r][j] = sum_b Ly[r][b] * node[cyi + b][jx0 + j]
    for (int item = tid; item < SY * NCOL; item += LAT_THREADS) {
      const int r = item / NCOL, j = item - r * NCOL;
      if (jx0 + j < P.nx) {
        const float2* nd = P.nodes + (((size_t)n * P.ny + cyi) * P.nx + (jx0 + j)) * V;
        float acc[V][2];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v][0] = acc[v][1] = 0.f;
#pragma unroll
        for (int b = 0; b < LAT_TAPS; ++b) {
          const float w = c_lag[lag_idx(SY)].w[r][b];
          if (V == 2) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(nd + (size_t)b * P.nx * V));
            acc[0][0] = fmaf(w, q.x, acc[0][0]); acc[0][1] = fmaf(w, q.y, acc[0][1]);
            acc[V - 1][0] = fmaf(w, q.z, acc[V - 1][0]); acc[V - 1][1] = fmaf(w, q.w, acc[V - 1][1]);
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float2 q = __ldg(nd + (size_t)b * P.nx * V + v);
              acc[v][0] = fmaf(w, q.x, acc[v][0]); acc[v][1] = fmaf(w, q.y, acc[v][1]);
            }
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) ysm[r][j][v] = make_float2(acc[v][0], acc[v][1]);
      }
    }
    __syncthreads();
    // ---- per-warp culling of the near list against this warp's 32 x SY pixel block
    unsigned cand = 0;
    if (n_all > 0) {
      if (cull) {
        bool keep = false;
        if (lane < n_all) {
          const float4 c = near_list[lane];
          const float y_lo = fmaf(P.stepy, (float)row0, -1.0f), y_hi = fmaf(P.stepy, (float)min(row0 + SY - 1, P.Ho - 1), -1.0f);
          const float ddx = fmaxf(fmaxf(wx_lo - c.x, c.x - wx_hi), 0.f), ddy = fmaxf(fmaxf(y_lo - c.y, c.y - y_hi), 0.f);
          keep = fmaf(ddx, ddx, ddy * ddy) < P.R2 * 1.0001f + 1e-12f;
        }
        cand = __ballot_sync(0xffffffffu, keep);
      } else {
        cand = 0xffffffffu;
      }
    }
    // LAT_RPI canvas rows per iteration: coordinates of all of them first, then ALL their taps
    // (2 views x 3 channels x 4 taps per row) in flight together, then the FMAs / blend / stores
#pragma unroll
    for (int r0 = 0; r0 < SY; r0 += LAT_RPI) {
      if (row0 + r0 >= P.Ho) break;
      float px[LAT_RPI][V], py[LAT_RPI][V];
#pragma unroll
      for (int i = 0; i < LAT_RPI; ++i) {
        const int r = r0 + i < SY ? r0 + i : SY - 1;
        const float rowf = (float)(row0 + r);
        float ax[V], ay[V];
        if (V == 2) {
          // packed: (x, y) of a view per FFMA2, the Lagrange weight broadcast to both halves
          const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(&ysm[r][js][0]);
          u64 acc0 = fmul2(q0.x, pk2(lx[0], lx[0])), acc1 = fmul2(q0.y, pk2(lx[0], lx[0]));
#pragma unroll
          for (int a = 1; a < LAT_TAPS; ++a) {
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(&ysm[r][js + a][0]);
            acc0 = ffma2(q.x, pk2(lx[a], lx[a]), acc0);
            acc1 = ffma2(q.y, pk2(lx[a], lx[a]), acc1);
          }
          upk2(acc0, ax[0], ay[0]);
          upk2(acc1, ax[V - 1], ay[V - 1]);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) ax[v] = ay[v] = 0.f;
#pragma unroll
          for (int a = 0; a < LAT_TAPS; ++a) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float2 q = ysm[r][js + a][v];
              ax[v] = fmaf(lx[a], q.x, ax[v]); ay[v] = fmaf(lx[a], q.y, ay[v]);
            }
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          px[i][v] = ax[v] + fmaf(prow[v][0], rowf, pcol[v][0]);
          py[i][v] = ay[v] + fmaf(prow[v][1], rowf, pcol[v][1]);
        }
      }
      // near-field corrections (branch-free: s is clamped to R2, where psi vanishes)
      if (cand != 0u) {
        float yt[LAT_RPI];
#pragma unroll
        for (int i = 0; i < LAT_RPI; ++i) yt[i] = fmaf(P.stepy, (float)(row0 + min(r0 + i, SY - 1)), -1.0f);
        unsigned m = cull ? cand : 0u;
        int kk = 0;
#pragma unroll 1
        while (cull ? (m != 0u) : (kk < n_all)) {
          int k;
          if (cull) { k = __ffs(m) - 1; m &= m - 1; } else { k = kk++; }
          const float4 c = near_list[k];
          const float dx = xt - c.x, dxx = dx * dx;
          const bool v0 = V == 1 || k < n0;
          int vk = 0;   // V > 2: the view this control point belongs to
          if (V > 2) {
#pragma unroll
            for (int v = 0; v < V - 1; ++v) vk += k >= view_end[v] ? 1 : 0;
          }
#pragma unroll
          for (int i = 0; i < LAT_RPI; ++i) {
            const float dy = yt[i] - c.y;
            const float s = fminf(fmaf(dy, dy, dxx), P.R2);
            // psi/ln2 = s*lg2(s+eps) - P(s)/ln2 (q0..q3 = P/ln2; the weights carry the ln2)
            const float psi = fmaf(s, lg2_approx(s + 1e-6f), -blend_poly(s, P.R2, P.q0, P.q1, P.q2, P.q3));
            if (V <= 2) {
              if (v0) { px[i][0] = fmaf(c.z, psi, px[i][0]); py[i][0] = fmaf(c.w, psi, py[i][0]); }
              else { px[i][V - 1] = fmaf(c.z, psi, px[i][V - 1]); py[i][V - 1] = fmaf(c.w, psi, py[i][V - 1]); }
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v)
                if (vk == v) { px[i][v] = fmaf(c.z, psi, px[i][v]); py[i][v] = fmaf(c.w, psi, py[i][v]); }
            }
          }
        }
      }
      // taps, blend and stores of LAT_TPI rows at a time (the coordinates above are computed for LAT_RPI rows at once)
#pragma unroll
      for (int ib = 0; ib < LAT_RPI; ib += LAT_TPI) {
        float res[LAT_RPI][V][C];
        if (MODE == SS2_MODE_NORMAL) {
          // Phase A: tap weights and addresses (branch-free; an out-of-image sample gets zero
          // weights and reads the frame's first pixel).  Phase B: all loads back to back (a view no
          // lane of the warp sees is skipped).  Phase C: the FMAs.
          float wq[LAT_RPI][V][4];
          const float* tp[LAT_RPI][V];
          bool anyv[LAT_RPI][V];
#pragma unroll
          for (int i = ib; i < ib + LAT_TPI; ++i) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const int xi = __float2int_rd(px[i][v]), yi = __float2int_rd(py[i][v]);
              const bool inb = (unsigned)xi < W1 && (unsigned)yi < H1;  // no tap is clamped: plain bilinear == _interpolate
              const float fx0 = px[i][v] - (float)xi, fy = py[i][v] - (float)yi;
              const float fx = inb ? fx0 : 0.0f, gx = inb ? 1.0f - fx0 : 0.0f, gy = 1.0f - fy;
              wq[i][v][0] = gx * gy; wq[i][v][1] = gx * fy; wq[i][v][2] = fx * gy; wq[i][v][3] = fx * fy;
              tp[i][v] = f32_at(imgv[v], inb ? (unsigned)(yi * W + xi) : 0u);
              anyv[i][v] = __any_sync(0xffffffffu, inb);
#if LAT_PREFETCH > 0
              // the next iteration samples (about) LAT_TPI source rows further down: pull them towards L1
              if (i == ib + LAT_TPI - 1 && anyv[i][v] && inb && (unsigned)(yi + LAT_PREFETCH) < H1) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                  asm volatile("prefetch.global.L1 [%0];" ::"l"(f32_at(tp[i][v], c * iplane + LAT_PREFETCH * W)));
              }
#endif
            }
          }
          float tap[LAT_RPI][V][C][4];
#pragma unroll
          for (int i = ib; i < ib + LAT_TPI; ++i) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              if (anyv[i][v]) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  if (IW > 0 && IH > 0 && (size_t)(C - 1) * IW * IH * 4 + (size_t)IW * 4 + 4 < (1u << 23)) {
                    const float* p = tp[i][v] + c * (IW * IH);  // compile-time offsets off ONE pointer
                    tap[i][v][c][0] = __ldg(p); tap[i][v][c][2] = __ldg(p + 1);
                    tap[i][v][c][1] = __ldg(p + IW); tap[i][v][c][3] = __ldg(p + IW + 1);
                  } else {
                    const float* p0 = c == 0 ? tp[i][v] : f32_at(tp[i][v], c * iplane);
                    const float* p1 = f32_at(tp[i][v], c * iplane + W);
                    tap[i][v][c][0] = __ldg(p0); tap[i][v][c][2] = __ldg(p0 + 1);
                    tap[i][v][c][1] = __ldg(p1); tap[i][v][c][3] = __ldg(p1 + 1);
                  }
                }
              }
            }
          }
#pragma unroll
          for (int i = ib; i < ib + LAT_TPI; ++i)
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
              for (int c = 0; c < C; ++c) {
                res[i][v][c] = 0.0f;
                if (anyv[i][v]) {
                  // (x0, x1) pairs of the upper and the lower source row through packed FMAs
                  float lo, hi;
                  upk2(ffma2(pk2(tap[i][v][c][1], tap[i][v][c][3]), pk2(wq[i][v][1], wq[i][v][3]),
                             fmul2(pk2(tap[i][v][c][0], tap[i][v][c][2]), pk2(wq[i][v][0], wq[i][v][2]))), lo, hi);
                  res[i][v][c] = lo + hi;
                }
              }
        } else {
#pragma unroll
          for (int i = ib; i < ib + LAT_TPI; ++i)
#pragma unroll
            for (int v = 0; v < V; ++v) sample_fast<C>(imgv[v], H, W, px[i][v], py[i][v], res[i][v]);
        }
#pragma unroll
        for (int i = ib; i < ib + LAT_TPI; ++i) {
          const int row = row0 + r0 + i;
          if (active && r0 + i < SY && row < P.Ho) {
            const unsigned opix = (unsigned)(row * P.Wo + col);
            if (BLEND && V > 2) {
              // fuse(..fuse(fuse(1,2),3)..,V): the reference's sequential AVERAGE fusion (test_online_tra_threeview.py:489-490)
#pragma unroll
              for (int c = 0; c < C; ++c) {
                float f = blend_avg_fast(res[i][0][c], res[i][1][c]);
#pragma unroll
                for (int v = 2; v < V; ++v) f = blend_avg_fast(f, res[i][v][c]);
                __stcs(const_cast<float*>(f32_at(outv[0], opix + c * oplane)), f);
              }
            } else if (BLEND && C == 3) {
              const u64 A = pk2(res[i][0][0], res[i][0][1]), B = pk2(res[i][V - 1][0], res[i][V - 1][1]);
              float s0, s1, q0, q1;
              upk2(fadd2(fadd2(A, B), pk2(1e-6f, 1e-6f)), s0, s1);
              upk2(ffma2(B, B, fmul2(A, A)), q0, q1);
              const float f0 = q0 * rcp_approx(s0), f1 = q1 * rcp_approx(s1), f2 = blend_avg_fast(res[i][0][2], res[i][V - 1][2]);
              if (U8) {
                // 0 <= f < 256 here (non-negative taps and weights), where numpy's astype(uint8) is plain truncation
                unsigned char* o8 = P.out8 + ((size_t)n * oplane + opix) * 3;
                __stcs(o8, (unsigned char)__float2int_rz(f0));
                __stcs(o8 + 1, (unsigned char)__float2int_rz(f1));
                __stcs(o8 + 2, (unsigned char)__float2int_rz(f2));
              } else {
                __stcs(const_cast<float*>(f32_at(outv[0], opix)), f0);
                __stcs(const_cast<float*>(f32_at(outv[0], opix + oplane)), f1);
                __stcs(const_cast<float*>(f32_at(outv[0], opix + 2 * oplane)), f2);
              }
            } else if (BLEND) {
#pragma unroll
              for (int c = 0; c < C; ++c)
                __stcs(const_cast<float*>(f32_at(outv[0], opix + c * oplane)), blend_avg_fast(res[i][0][c], res[i][V - 1][c]));
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v) {
#pragma unroll
                for (int c = 0; c < C; ++c) __stcs(const_cast<float*>(f32_at(outv[v], opix + c * oplane)), res[i][v][c]);
              }
            }
          }
        }
      }
    }
  }
}

static inline float linstep(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

// lattice configuration: spacing (SX, SY) in canvas pixels and near radius R (normalised)
struct LatticeConfig { int SX, SY; float R; bool ok; };

// The interpolation error is governed by the lattice step in NORMALISED canvas units,
// hx = SX*2/(Wo-1), hy = SY*2/(Ho-1): take the coarsest instantiated spacing with h <= 0.0185
// (12x6 on a 741x1748 canvas: hx 0.0137, hy 0.0162), near radius R = 4.3*h.  Canvases too small
// for that (h would exceed the bound even at the finest spacing) use the EXACT evaluation.
static LatticeConfig lattice_config(int Ho, int Wo) {
  LatticeConfig c;
  const double hmax = 0.0185, ux = 2.0 / (Wo > 1 ? Wo - 1 : 1), uy = 2.0 / (Ho > 1 ? Ho - 1 : 1);
  c.SX = 16 * ux <= hmax ? 16 : (12 * ux <= hmax ? 12 : 8);
  c.SY = 8 * uy <= hmax ? 8 : 6;
  const char* e = getenv("SS2_TPS_S");
  if (e && atoi(e) > 100) { c.SX = atoi(e) / 10; c.SY = atoi(e) % 10; }   // e.g. 168, 126, 88
  const double h = c.SX * ux > c.SY * uy ? c.SX * ux : c.SY * uy;
  c.ok = h <= hmax * 1.0001 || e;
  c.R = (float)(4.3 * h);
  e = getenv("SS2_TPS_R");
  if (e) c.R = (float)atof(e);
  return c;
}

static void lagrange_table(int S, LagrangeTable* t) {
  const double xs[LAT_TAPS] = {-2, -1, 0, 1, 2, 3};
  for (int k = 0; k < 16; ++k) {
    const double u = k < S ? (double)k / S : 0.0;
    for (int j = 0; j < 8; ++j) {
      double w = 0.0;
      if (j < LAT_TAPS) {
        w = 1.0;
        for (int m = 0; m < LAT_TAPS; ++m)
          if (m != j) w *= (u - xs[m]) / (xs[j] - xs[m]);
      }
      t->w[k][j] = (float)w;
    }
  }
}

bool tps_lattice_supported(int Ho, int Wo) { return Ho > 0 && Wo > 0 && lattice_config(Ho, Wo).ok; }

// workspace of the lattice path for `bn` (frame, view) systems on a Ho x Wo canvas
size_t tps_lattice_workspace_floats(int bn, int Ho, int Wo) {
  const LatticeConfig c = lattice_config(Ho, Wo);
  const size_t nx = (size_t)(Wo - 1) / c.SX + LAT_TAPS, ny = (size_t)(Ho - 1) / c.SY + LAT_TAPS;
  return (size_t)bn * nx * ny * 2;
}

// blending cubic of phi(s) = s log(s + eps) at s = R2 (third-order Taylor polynomial), coefficients divided by `div`
static void blend_cubic(double R2, double div, float* p) {
  const double e = 1e-6, L = log(R2 + e);
  p[0] = (float)(R2 * L / div);
  p[1] = (float)((L + R2 / (R2 + e)) / div);
  p[2] = (float)(0.5 * (1.0 / (R2 + e) + e / ((R2 + e) * (R2 + e))) / div);
  p[3] = (float)((-1.0 / ((R2 + e) * (R2 + e)) - 2.0 * e / ((R2 + e) * (R2 + e) * (R2 + e))) / 6.0 / div);
}

template <int V, int C, bool BLEND, bool U8 = false>
static int lattice_launch(ss2_ctx* ctx, WarpParams P, int nframes, int mode, float* d_nodes, cudaStream_t st) {
  const LatticeConfig cfg = lattice_config(P.Ho, P.Wo);
  P.nx = (P.Wo - 1) / cfg.SX + LAT_TAPS;
  P.ny = (P.Ho - 1) / cfg.SY + LAT_TAPS;
  P.nodes = reinterpret_cast<const float2*>(d_nodes);
  const double R2 = (double)cfg.R * cfg.R;
  P.R2 = (float)R2;
  float pc[4];
  blend_cubic(R2, 1.0, pc);
  P.p0 = pc[0]; P.p1 = pc[1]; P.p2 = pc[2]; P.p3 = pc[3];
  blend_cubic(R2, 0.69314718055994530942, pc);
  P.q0 = pc[0]; P.q1 = pc[1]; P.q2 = pc[2]; P.q3 = pc[3];
  if (!ctx->lag_tables_ready) {
    // one-time upload of the interpolation weight tables of this context's device
    LagrangeTable t[4];
    const int S[4] = {6, 8, 12, 16};
    for (int i = 0; i < 4; ++i) lagrange_table(S[i], &t[i]);
    SS2_CUDA(ctx, cudaMemcpyToSymbol(c_lag, t, sizeof(t)));
    ctx->lag_tables_ready = true;
  }
  const double kxd = P.Wo > 1 ? 2.0 / (double)(P.Wo - 1) : 0.0, kyd = P.Ho > 1 ? 2.0 / (double)(P.Ho - 1) : 0.0;
  tps_nodes_kernel<V><<<dim3(cdiv(P.nx * P.ny, 128), V, nframes), 128, 0, st>>>(P, cfg.SX, cfg.SY, kxd, kyd);
  SS2_LAUNCH_CHECK(ctx);
  dim3 grid(cdiv(P.Wo, LAT_THREADS), cdiv(P.Ho, cfg.SY * LAT_NCELL), nframes);
  // source-size specialisations of the production (fused, NORMAL) kernel: 720p and 1080p
  const int spec = (BLEND && V == 2 && mode == SS2_MODE_NORMAL) ? ((P.W == 1280 && P.H == 720) ? 1 : (P.W == 1920 && P.H == 1080) ? 2 : 0) : 0;
#define LAT_CASE(SXV, SYV)                                                                                   \
  if (cfg.SX == SXV && cfg.SY == SYV) {                                                                      \
    if (spec == 1) tps_warp_lattice_kernel<V, C, SS2_MODE_NORMAL, BLEND, SXV, SYV, (BLEND && V == 2) ? 1280 : 0, (BLEND && V == 2) ? 720 : 0, U8><<<grid, LAT_THREADS, 0, st>>>(P); \
    else if (spec == 2) tps_warp_lattice_kernel<V, C, SS2_MODE_NORMAL, BLEND, SXV, SYV, (BLEND && V == 2) ? 1920 : 0, (BLEND && V == 2) ? 1080 : 0, U8><<<grid, LAT_THREADS, 0, st>>>(P); \
    else if (mode == SS2_MODE_NORMAL) tps_warp_lattice_kernel<V, C, SS2_MODE_NORMAL, BLEND, SXV, SYV, 0, 0, U8><<<grid, LAT_THREADS, 0, st>>>(P); \
    else tps_warp_lattice_kernel<V, C, SS2_MODE_FAST, BLEND, SXV, SYV, 0, 0, U8><<<grid, LAT_THREADS, 0, st>>>(P);  \
  }
  LAT_CASE(16, 8) LAT_CASE(16, 6) LAT_CASE(12, 8) LAT_CASE(12, 6) LAT_CASE(8, 8) LAT_CASE(8, 6)
#undef LAT_CASE
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

int tps_warp_launch(ss2_ctx* ctx, const float* d_U, const float* d_source, const float* d_T, int bn, int C,
                    int H, int W, int Ho, int Wo, int mode, int tps, float* d_out, cudaStream_t st,
                    const float* d_aux, float* d_nodes) {
  if (bn <= 0 || Ho <= 0 || Wo <= 0) return SS2_OK;
  WarpParams P;
  P.img[0] = d_U; P.img[1] = d_U; P.img[2] = d_U; P.img[3] = d_U;
  P.source = d_source; P.T = d_T; P.out = d_out; P.out8 = nullptr;
  P.H = H; P.W = W; P.Ho = Ho; P.Wo = Wo;
  P.stepx = linstep(Wo); P.stepy = linstep(Ho);
  P.aux = d_aux;
  P.half_w = mode == SS2_MODE_NORMAL ? 0.5f * W : 0.5f * (W - 1);
  P.half_h = mode == SS2_MODE_NORMAL ? 0.5f * H : 0.5f * (H - 1);
  if (tps == SS2_TPS_LATTICE && d_aux && d_nodes && C == 3 && tps_lattice_supported(Ho, Wo))
    return lattice_launch<1, 3, false>(ctx, P, bn, mode, d_nodes, st);
  if (tps == SS2_TPS_LATTICE && d_aux && d_nodes && C == 4 && tps_lattice_supported(Ho, Wo))   // image + LINEAR's mask
    return lattice_launch<1, 4, false>(ctx, P, bn, mode, d_nodes, st);
  dim3 grid(cdiv(Wo, TX), cdiv(Ho, TILE_H), bn), block(TX, TY);
#define WARP_CASE(CC)                                                                              \
  if (C == CC) {                                                                                   \
    if (mode == SS2_MODE_NORMAL) tps_warp_exact_kernel<1, CC, SS2_MODE_NORMAL, false><<<grid, block, 0, st>>>(P); \
    else tps_warp_exact_kernel<1, CC, SS2_MODE_FAST, false><<<grid, block, 0, st>>>(P);            \
    SS2_LAUNCH_CHECK(ctx);                                                                         \
    return SS2_OK;                                                                                 \
  }
  WARP_CASE(1) WARP_CASE(2) WARP_CASE(3) WARP_CASE(4) WARP_CASE(6)   // 6: image + three mask planes of the metric path
#undef WARP_CASE
  // other channel counts: one plane at a time
  for (int c = 0; c < C; ++c) {
    for (int b = 0; b < bn; ++b) {
      WarpParams Q = P;
      Q.img[0] = Q.img[1] = d_U + ((size_t)b * C + c) * H * W;
      Q.source = d_source + (size_t)b * SS2_NPT * 2;
      Q.T = d_T + (size_t)b * 2 * SS2_NSYS;
      Q.out = d_out + ((size_t)b * C + c) * Ho * Wo;
      dim3 g1(grid.x, grid.y, 1);
      if (mode == SS2_MODE_NORMAL) tps_warp_exact_kernel<1, 1, SS2_MODE_NORMAL, false><<<g1, block, 0, st>>>(Q);
      else tps_warp_exact_kernel<1, 1, SS2_MODE_FAST, false><<<g1, block, 0, st>>>(Q);
      SS2_LAUNCH_CHECK(ctx);
    }
  }
  return SS2_OK;
}

int tps_warp_blend_launch(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_source,
                          const float* d_T, int nframes, int H, int W, int Ho, int Wo, int mode, int tps,
                          float* d_out, cudaStream_t st, const float* d_aux, float* d_nodes, unsigned char* d_out8) {
  if (nframes <= 0 || Ho <= 0 || Wo <= 0) return SS2_OK;
  WarpParams P;
  P.img[0] = d_img1; P.img[1] = d_img2; P.img[2] = d_img2; P.img[3] = d_img2;
  P.source = d_source; P.T = d_T; P.out = d_out; P.out8 = nullptr;
  P.H = H; P.W = W; P.Ho = Ho; P.Wo = Wo;
  P.stepx = linstep(Wo); P.stepy = linstep(Ho);
  P.aux = d_aux;
  P.half_w = mode == SS2_MODE_NORMAL ? 0.5f * W : 0.5f * (W - 1);
  P.half_h = mode == SS2_MODE_NORMAL ? 0.5f * H : 0.5f * (H - 1);
  // (the SS2_PROF_WARP bracket is set by the callers in api.cu, around canvas meshes + solves + nodes + this)
  P.out8 = d_out8;
  if (tps == SS2_TPS_LATTICE && d_aux && d_nodes && tps_lattice_supported(Ho, Wo))
    return d_out8 ? lattice_launch<2, 3, true, true>(ctx, P, nframes, mode, d_nodes, st)
                  : lattice_launch<2, 3, true>(ctx, P, nframes, mode, d_nodes, st);
  if (d_out8) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "uint8 output is fused into the lattice resampler only");
  dim3 grid(cdiv(Wo, TX), cdiv(Ho, TILE_H), nframes), block(TX, TY);
  if (mode == SS2_MODE_NORMAL) tps_warp_exact_kernel<2, 3, SS2_MODE_NORMAL, true><<<grid, block, 0, st>>>(P);
  else tps_warp_exact_kernel<2, 3, SS2_MODE_FAST, true><<<grid, block, 0, st>>>(P);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

int tps_warp_blend_n_launch(ss2_ctx* ctx, const float* const* d_imgs, int nviews, const float* d_source, const float* d_T,
                            int nframes, int H, int W, int Ho, int Wo, int mode, int tps, float* d_out, cudaStream_t st,
                            const float* d_aux, float* d_nodes) {
  if (nframes <= 0 || Ho <= 0 || Wo <= 0) return SS2_OK;
  if (nviews == 2)
    return tps_warp_blend_launch(ctx, d_imgs[0], d_imgs[1], d_source, d_T, nframes, H, W, Ho, Wo, mode, tps, d_out, st, d_aux, d_nodes, nullptr);
  if (nviews != 3 && nviews != 4) return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "fused N-view resampler: 2 <= N <= 4 (got %d)", nviews);
  WarpParams P;
  for (int v = 0; v < 4; ++v) P.img[v] = d_imgs[v < nviews ? v : 0];
  P.source = d_source; P.T = d_T; P.out = d_out; P.out8 = nullptr;
  P.H = H; P.W = W; P.Ho = Ho; P.Wo = Wo;
  P.stepx = linstep(Wo); P.stepy = linstep(Ho);
  P.aux = d_aux;
  P.half_w = mode == SS2_MODE_NORMAL ? 0.5f * W : 0.5f * (W - 1);
  P.half_h = mode == SS2_MODE_NORMAL ? 0.5f * H : 0.5f * (H - 1);
  if (tps == SS2_TPS_LATTICE && d_aux && d_nodes && tps_lattice_supported(Ho, Wo))
    return nviews == 3 ? lattice_launch<3, 3, true>(ctx, P, nframes, mode, d_nodes, st)
                       : lattice_launch<4, 3, true>(ctx, P, nframes, mode, d_nodes, st);
  dim3 grid(cdiv(Wo, TX), cdiv(Ho, TILE_H), nframes), block(TX, TY);
  if (nviews == 3) {
    if (mode == SS2_MODE_NORMAL) tps_warp_exact_kernel<3, 3, SS2_MODE_NORMAL, true><<<grid, block, 0, st>>>(P);
    else tps_warp_exact_kernel<3, 3, SS2_MODE_FAST, true><<<grid, block, 0, st>>>(P);
  } else {
    if (mode == SS2_MODE_NORMAL) tps_warp_exact_kernel<4, 3, SS2_MODE_NORMAL, true><<<grid, block, 0, st>>>(P);
    else tps_warp_exact_kernel<4, 3, SS2_MODE_FAST, true><<<grid, block, 0, st>>>(P);
  }
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
