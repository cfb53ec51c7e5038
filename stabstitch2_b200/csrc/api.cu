// C ABI entry points of libss2.so that are not network forwards (see include/ss2.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

int ss2_fail(ss2_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

int ss2_ensure_arena(ss2_ctx* ctx, size_t bytes) {
  if (ctx->arena.cap >= bytes) return SS2_OK;
  // growing is rare (first call / larger batch): drain the device, then replace the slab
  SS2_CUDA(ctx, cudaDeviceSynchronize());
  if (ctx->arena.base) SS2_CUDA(ctx, cudaFree(ctx->arena.base));
  ctx->arena.base = nullptr;
  ctx->arena.cap = 0;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return ss2_fail(ctx, SS2_ERR_OOM, "cudaMalloc(%zu) for the workspace arena: %s", bytes, cudaGetErrorString(e));
  ctx->arena.base = (char*)p;
  ctx->arena.cap = bytes;
  ctx->arena.off = 0;
  return SS2_OK;
}

int ss2_workspace_enter(ss2_ctx* ctx, cudaStream_t st) {
  if (ctx->ws_nested) return SS2_OK;   // forked by an outer entry point, which entered on the caller's stream and joins it
  if (ctx->h_range_flag && *(volatile int*)ctx->h_range_flag) {
    *ctx->h_range_flag = 0;
    return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "an activation of an earlier call exceeded the fp16 range (|v| > 65504) of the fp16 "
                    "split planes: its results are invalid; run with SS2_F16=0 (TF32 split planes)");
  }
  if (ctx->ws_used && st != ctx->ws_stream) {
    if (!ctx->ws_ev) SS2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ws_ev, cudaEventDisableTiming));
    SS2_CUDA(ctx, cudaEventRecord(ctx->ws_ev, ctx->ws_stream));
    SS2_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ws_ev, 0));
  }
  ctx->ws_stream = st;
  ctx->ws_used = true;
  return SS2_OK;
}

void ss2_prof_begin(ss2_ctx* ctx, int which, cudaStream_t st) {
  ProfClass& p = ctx->prof[which];
  if (!p.enabled) return;
  if (p.used + 2 > p.pool.size()) {
    for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); p.pool.push_back(e); }
  }
  cudaEventRecord(p.pool[p.used], st);
}

void ss2_prof_end(ss2_ctx* ctx, int which, cudaStream_t st, double work) {
  ProfClass& p = ctx->prof[which];
  if (!p.enabled) return;
  cudaEventRecord(p.pool[p.used + 1], st);
  p.used += 2;
  p.work += work;
}

int tps_scratch_alloc(ss2_ctx* ctx, int bn, int Ho, int Wo, int tps, size_t extra_floats, TpsScratch* s, cudaStream_t st) {
  const size_t nT = (size_t)bn * 2 * SS2_NSYS, nA = (size_t)bn * 8;
  const size_t nN = tps == SS2_TPS_LATTICE ? tps_lattice_workspace_floats(bn, Ho, Wo) : 0;
  const size_t pad = 64;  // keep every part 256-byte aligned
  auto up = [&](size_t v) { return (v + pad - 1) / pad * pad; };
  SS2_CUDA(ctx, cudaMallocAsync((void**)&s->base, (up(nT) + up(nA) + up(nN) + extra_floats) * sizeof(float), st));
  s->T = s->base;
  s->aux = s->T + up(nT);
  s->nodes = s->aux + up(nA);
  return SS2_OK;
}

int tps_solve_for_warp(ss2_ctx* ctx, const float* d_source, const float* d_target, int bn, int H, int W, int Ho, int Wo,
                       int mode, int tps, const TpsScratch& s, cudaStream_t st) {
  if (tps != SS2_TPS_LATTICE) return tps_solve_launch(ctx, d_source, d_target, bn, s.T, st);
  const float hw = mode == SS2_MODE_NORMAL ? 0.5f * W : 0.5f * (W - 1);
  const float hh = mode == SS2_MODE_NORMAL ? 0.5f * H : 0.5f * (H - 1);
  return tps_solve_aux_launch(ctx, d_source, d_target, bn, s.T, s.aux, hw, hh, Ho, Wo, st);
}

extern "C" {

int ss2_profile_enable(ss2_ctx* ctx, int which, int enable) {
  if (!ctx || which < 0 || which >= SS2_PROF_COUNT) return SS2_ERR_INVALID;
  ctx->prof[which].enabled = enable != 0;
  ctx->prof[which].used = 0;
  ctx->prof[which].work = 0.0;
  return SS2_OK;
}

int ss2_profile_read(ss2_ctx* ctx, int which, double* total_ms, int64_t* launches, double* work) {
  if (!ctx || which < 0 || which >= SS2_PROF_COUNT) return SS2_ERR_INVALID;
  ProfClass& p = ctx->prof[which];
  double tot = 0.0;
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    SS2_CUDA(ctx, cudaEventSynchronize(p.pool[i + 1]));
    float ms = 0.f;
    SS2_CUDA(ctx, cudaEventElapsedTime(&ms, p.pool[i], p.pool[i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)(p.used / 2);
  if (work) *work = p.work;
  return SS2_OK;
}

const char* ss2_version(void) { return "ss2 0.1 (sm_100a)"; }

int ss2_create(int device, ss2_ctx** out) {
  if (!out) return SS2_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SS2_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SS2_ERR_CUDA;
  ss2_ctx* c = new ss2_ctx();
  c->device = device;
  {
    // keep the stream-ordered allocator's memory cached across synchronisation points: the small
    // per-call scratch (TPS coefficients, lattice nodes) would otherwise be returned to the OS at
    // every host sync and re-mapped on the next call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  const char* env = getenv("SS2_USE_TC");
  if (env) c->use_tc = atoi(env);
  env = getenv("SS2_TC_PASSES");
  if (env) c->tc_passes = atoi(env) == 1 ? 1 : 3;
  env = getenv("SS2_TC_STEM");
  if (env) c->use_tc_stem = atoi(env);
  env = getenv("SS2_SIDE_STREAM");
  if (env) c->use_side = atoi(env);
  env = getenv("SS2_NET_OVERLAP");
  if (env) c->use_net_overlap = atoi(env);
  env = getenv("SS2_F16");
  if (env) c->use_f16 = atoi(env);
  if (cudaHostAlloc((void**)&c->h_range_flag, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&c->d_range_flag, c->h_range_flag, 0) != cudaSuccess) {
    cudaGetLastError();
    c->h_range_flag = c->d_range_flag = nullptr;
    c->use_f16 = 0;   // no way to report a range overflow: stay on the TF32 planes
  } else {
    *c->h_range_flag = 0;
  }
  env = getenv("SS2_CONV_DC");
  if (env) c->use_dc = atoi(env);
  *out = c;
  return SS2_OK;
}

void ss2_destroy(ss2_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ss2_host_slots_free(ctx);
  if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
  if (ctx->ws_ev) cudaEventDestroy(ctx->ws_ev);
  if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (void* p : ctx->owned) cudaFree(p);
  for (auto& v : ctx->owned_net) for (void* p : v) cudaFree(p);
  if (ctx->arena.base) cudaFree(ctx->arena.base);
  if (ctx->arena_alt.base) cudaFree(ctx->arena_alt.base);
  if (ctx->h_range_flag) cudaFreeHost(ctx->h_range_flag);
  if (ctx->s_net) cudaStreamDestroy(ctx->s_net);
  if (ctx->ev_nfork) cudaEventDestroy(ctx->ev_nfork);
  if (ctx->ev_njoin) cudaEventDestroy(ctx->ev_njoin);
  for (auto& kv : ctx->stream_bufs) cudaFree(kv.second.first);
  for (auto& pc : ctx->prof) for (cudaEvent_t e : pc.pool) cudaEventDestroy(e);
  delete ctx;
}

const char* ss2_last_error(ss2_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int64_t ss2_launch_count(ss2_ctx* ctx, int reset) {
  if (!ctx) return -1;
  int64_t n = ctx->launches;
  if (reset) ctx->launches = 0;
  return n;
}

int ss2_load_tensor(ss2_ctx* ctx, int net_id, const char* key, const float* h_data, const int64_t* shape, int ndim) {
  if (!ctx) return SS2_ERR_INVALID;
  if (net_id < 0 || net_id > 2 || !key || !h_data || ndim < 0 || ndim > 8 || (ndim > 0 && !shape))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_load_tensor: bad arguments");
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  if (n < 0) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_load_tensor: negative extent in '%s'", key);
  t.data.assign(h_data, h_data + n);
  ctx->host_weights[net_id][key] = std::move(t);
  return SS2_OK;
}

int ss2_dlt(ss2_ctx* ctx, const float* d_src, const float* d_dst, int bs, float* d_H, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (bs < 0 || (bs > 0 && (!d_src || !d_dst || !d_H))) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_dlt: bad arguments");
  return dlt_launch(ctx, d_src, d_dst, bs, d_H, (cudaStream_t)stream);
}

int ss2_homo_warp(ss2_ctx* ctx, const float* d_U, const float* d_theta, int bn, int C, int H, int W, int Ho, int Wo,
                  float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (bn < 0 || C <= 0 || H <= 0 || W <= 0 || Ho < 0 || Wo < 0 || (bn > 0 && (!d_U || !d_theta || !d_out)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_homo_warp: bad arguments");
  return homo_warp_nchw_launch(ctx, d_U, d_theta, bn, C, H, W, Ho, Wo, d_out, (cudaStream_t)stream);
}

int ss2_tps_point(ss2_ctx* ctx, const float* d_point, const float* d_source, const float* d_target, int bn,
                  float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (bn < 0 || (bn > 0 && (!d_point || !d_source || !d_target || !d_out)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_tps_point: bad arguments");
  if (bn == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* T = nullptr;
  SS2_CUDA(ctx, cudaMallocAsync((void**)&T, (size_t)bn * 2 * SS2_NSYS * sizeof(float), st));
  int rc = tps_solve_launch(ctx, d_source, d_target, bn, T, st);
  if (rc == SS2_OK) rc = tps_point_launch(ctx, d_point, d_source, T, bn, d_out, st);
  cudaFreeAsync(T, st);
  return rc;
}

int ss2_tps_warp(ss2_ctx* ctx, const float* d_U, const float* d_source, const float* d_target, int bn, int C, int H,
                 int W, int Ho, int Wo, int mode, int tps, float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (bn < 0 || C <= 0 || H <= 0 || W <= 0 || Ho < 0 || Wo < 0 || (mode != SS2_MODE_NORMAL && mode != SS2_MODE_FAST) ||
      (bn > 0 && (!d_U || !d_source || !d_target || (!d_out && Ho * Wo > 0))))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_tps_warp: bad arguments");
  if (bn == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (tps != SS2_TPS_LATTICE || (C != 3 && C != 4) || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  TpsScratch sc;
  SS2_TRY(tps_scratch_alloc(ctx, bn, Ho, Wo, tps, 0, &sc, st));
  int rc = tps_solve_for_warp(ctx, d_source, d_target, bn, H, W, Ho, Wo, mode, tps, sc, st);
  if (rc == SS2_OK) rc = tps_warp_launch(ctx, d_U, d_source, sc.T, bn, C, H, W, Ho, Wo, mode, tps, d_out, st, sc.aux, sc.nodes);
  cudaFreeAsync(sc.base, st);
  return rc;
}

int ss2_tps_warp_blend_avg(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_source,
                           const float* d_target, int nframes, int H, int W, int Ho, int Wo, int mode, int tps,
                           float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (nframes < 0 || H <= 0 || W <= 0 || Ho < 0 || Wo < 0 || (mode != SS2_MODE_NORMAL && mode != SS2_MODE_FAST) ||
      (nframes > 0 && (!d_img1 || !d_img2 || !d_source || !d_target || (!d_out && Ho * Wo > 0))))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_tps_warp_blend_avg: bad arguments");
  if (nframes == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (tps != SS2_TPS_LATTICE || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  TpsScratch sc;
  SS2_TRY(tps_scratch_alloc(ctx, 2 * nframes, Ho, Wo, tps, 0, &sc, st));
  // SS2_PROF_WARP brackets the WHOLE resampling of the chunk: TPS solves + lattice nodes + resample/blend kernel;
  // work = algorithmic bytes (both source frames read once, the fused frame written once)
  ss2_prof_begin(ctx, SS2_PROF_WARP, st);
  int rc = tps_solve_for_warp(ctx, d_source, d_target, 2 * nframes, H, W, Ho, Wo, mode, tps, sc, st);
  if (rc == SS2_OK)
    rc = tps_warp_blend_launch(ctx, d_img1, d_img2, d_source, sc.T, nframes, H, W, Ho, Wo, mode, tps, d_out, st, sc.aux, sc.nodes);
  ss2_prof_end(ctx, SS2_PROF_WARP, st, (double)nframes * (2.0 * 3 * H * W + 3.0 * Ho * Wo) * 4.0);
  cudaFreeAsync(sc.base, st);
  return rc;
}

int ss2_three_view_meshes(ss2_ctx* ctx, const float* d_w12m1, const float* d_w12m2, const float* d_w23m1,
                          const float* d_w23m2, int n, int img_h, int img_w, float* d_mesh1, float* d_middle,
                          float* d_mesh3, float* d_canvas, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n <= 0 || img_h <= 0 || img_w <= 0 || !d_w12m1 || !d_w12m2 || !d_w23m1 || !d_w23m2 || !d_mesh1 || !d_middle ||
      !d_mesh3 || !d_canvas)
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_meshes: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)n * SS2_NPT * 2;
  float* buf = nullptr;
  // work (5 m) | pt12 src12 pt23 src23 tgt (5 m) | moved12 moved23 (2 m) | canvas1 (64) | T (2 systems sets)
  SS2_CUDA(ctx, cudaMallocAsync((void**)&buf, (12 * m + 64 + (size_t)2 * n * 2 * SS2_NSYS) * sizeof(float), st));
  float *work = buf, *pt12 = buf + 5 * m, *src12 = pt12 + m, *pt23 = src12 + m, *src23 = pt23 + m, *tgt = src23 + m;
  float *moved12 = tgt + m, *moved23 = moved12 + m, *canvas1 = moved23 + m, *T = canvas1 + 64;
  int rc = three_view_align_launch(ctx, d_w12m1, d_w12m2, d_w23m1, d_w23m2, n, img_h, img_w, work, pt12, src12, pt23,
                                   src23, tgt, d_middle, canvas1, st);
  // view 1 through the TPS (shared view of pair (1,2) -> middle), view 3 through (shared view of pair (2,3) -> middle)
  if (rc == SS2_OK) rc = tps_solve_launch(ctx, src12, tgt, n, T, st);
  if (rc == SS2_OK) rc = tps_point_launch(ctx, pt12, src12, T, n, moved12, st);
  if (rc == SS2_OK) rc = tps_solve_launch(ctx, src23, tgt, n, T + (size_t)n * 2 * SS2_NSYS, st);
  if (rc == SS2_OK) rc = tps_point_launch(ctx, pt23, src23, T + (size_t)n * 2 * SS2_NSYS, n, moved23, st);
  if (rc == SS2_OK) rc = three_view_canvas_launch(ctx, moved12, moved23, d_middle, canvas1, n, d_mesh1, d_mesh3, d_canvas, st);
  cudaFreeAsync(buf, st);
  return rc;
}

int ss2_three_view_frames(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_img3,
                          const float* d_mesh1, const float* d_middle, const float* d_mesh3, int n, int H, int W,
                          const float* h_canvas, int mode, int tps, float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || H <= 0 || W <= 0 || !h_canvas || (mode != SS2_MODE_NORMAL && mode != SS2_MODE_FAST) ||
      (n > 0 && (!d_img1 || !d_img2 || !d_img3 || !d_mesh1 || !d_middle || !d_mesh3)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_frames: bad arguments");
  const float out_w = h_canvas[2], out_h = h_canvas[3];
  const int Ho = (int)out_h, Wo = (int)out_w;
  if (Ho < 0 || Wo < 0) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_frames: negative canvas");
  if (n == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  if (!d_out) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_three_view_frames: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (tps != SS2_TPS_LATTICE || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  const int chunk = n < 4 ? n : 4;  // frames warped per pass: 3 x chunk temporaries of the canvas size
  const size_t m = (size_t)n * SS2_NPT * 2, plane3 = (size_t)3 * Ho * Wo;
  TpsScratch sc;
  SS2_TRY(tps_scratch_alloc(ctx, chunk, Ho, Wo, tps, 6 * m + 3 * (size_t)chunk * plane3 + 64, &sc, st));
  float* source = sc.nodes + (tps == SS2_TPS_LATTICE ? (tps_lattice_workspace_floats(chunk, Ho, Wo) + 63) / 64 * 64 : 0);
  float* target = source + 3 * m;
  float* tmp = target + 3 * m;
  tmp += (64 - ((size_t)(tmp - sc.base) & 63)) & 63;  // keep the image temporaries 256-byte aligned
  int rc = three_view_sources_launch(ctx, d_mesh1, d_middle, d_mesh3, n, H, W, h_canvas[0], h_canvas[1], out_w, out_h,
                                     source, target, st);
  const float* imgs[3] = {d_img1, d_img2, d_img3};
  for (int k0 = 0; k0 < n && rc == SS2_OK; k0 += chunk) {
    const int nk = n - k0 < chunk ? n - k0 : chunk;
    for (int v = 0; v < 3 && rc == SS2_OK; ++v) {
      const float* src = source + ((size_t)v * n + k0) * SS2_NPT * 2;
      const float* tgt = target + ((size_t)v * n + k0) * SS2_NPT * 2;
      rc = tps_solve_for_warp(ctx, src, tgt, nk, H, W, Ho, Wo, mode, tps, sc, st);
      if (rc == SS2_OK)
        rc = tps_warp_launch(ctx, imgs[v] + (size_t)k0 * 3 * H * W, src, sc.T, nk, 3, H, W, Ho, Wo, mode, tps,
                             tmp + (size_t)v * chunk * plane3, st, sc.aux, sc.nodes);
    }
    if (rc == SS2_OK)
      rc = blend3_avg_launch(ctx, tmp, tmp + (size_t)chunk * plane3, tmp + 2 * (size_t)chunk * plane3, (size_t)nk * plane3,
                             d_out + (size_t)k0 * plane3, st);
  }
  cudaFreeAsync(sc.base, st);
  return rc;
}

// ---- N views (config 5): chain of pairs (1,2), (2,3), .., (N-1,N) ------------------------------------------------
int ss2_nview_align(ss2_ctx* ctx, const float* const* h_pair_meshes, int nviews, int n, int img_h, int img_w,
                    float* d_shifted, float* d_mids, float* d_minmax1, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (nviews < 3 || nviews > 8 || n <= 0 || img_h <= 0 || img_w <= 0 || !h_pair_meshes || !d_shifted || !d_mids || !d_minmax1)
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_align: bad arguments (3 <= views <= 8)");
  for (int q = 0; q < 2 * (nviews - 1); ++q)
    if (!h_pair_meshes[q]) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_align: null mesh pointer %d", q);
  return nview_align_launch(ctx, h_pair_meshes, nviews, n, img_h, img_w, d_shifted, d_mids, d_minmax1, (cudaStream_t)stream);
}

int ss2_nview_remap(ss2_ctx* ctx, int nviews, int n, const float* d_shifted, const float* d_mids, const float* h_minmax1,
                    float* d_meshes, float* d_minmax2, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (nviews < 3 || nviews > 8 || n <= 0 || !d_shifted || !d_mids || !h_minmax1 || !d_meshes || !d_minmax2)
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_remap: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const float xmin = h_minmax1[0], ymin = h_minmax1[2], ow = h_minmax1[1] - h_minmax1[0], oh = h_minmax1[3] - h_minmax1[2];
  const size_t m = (size_t)n * SS2_NPT * 2;
  float* buf = nullptr;
  // pt0 src0 tgt0 pt1 src1 tgt1 (6 m) | moved0 moved1 (2 m) | T (2 sets)
  SS2_CUDA(ctx, cudaMallocAsync((void**)&buf, (8 * m + (size_t)2 * n * 2 * SS2_NSYS) * sizeof(float), st));
  float *pt0 = buf, *src0 = buf + m, *tgt0 = buf + 2 * m, *pt1 = buf + 3 * m, *src1 = buf + 4 * m, *tgt1 = buf + 5 * m;
  float *moved0 = buf + 6 * m, *moved1 = buf + 7 * m, *T = buf + 8 * m;
  int rc = nview_operands_launch(ctx, d_shifted, d_mids, nviews, n, xmin, ymin, ow, oh, pt0, src0, tgt0, pt1, src1, tgt1, d_meshes, st);
  if (rc == SS2_OK) rc = tps_solve_launch(ctx, src0, tgt0, n, T, st);
  if (rc == SS2_OK) rc = tps_point_launch(ctx, pt0, src0, T, n, moved0, st);
  if (rc == SS2_OK) rc = tps_solve_launch(ctx, src1, tgt1, n, T + (size_t)n * 2 * SS2_NSYS, st);
  if (rc == SS2_OK) rc = tps_point_launch(ctx, pt1, src1, T + (size_t)n * 2 * SS2_NSYS, n, moved1, st);
  if (rc == SS2_OK) rc = nview_finish_launch(ctx, moved0, moved1, nviews, n, ow, oh, d_meshes, d_minmax2, st);
  cudaFreeAsync(buf, st);
  return rc;
}

int ss2_nview_frames(ss2_ctx* ctx, const float* const* h_imgs, const float* d_meshes, int nviews, int n, int H, int W,
                     const float* h_minmax2, int mode, int tps, float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (nviews < 2 || nviews > 4 || n < 0 || H <= 0 || W <= 0 || !h_imgs || !h_minmax2 ||
      (mode != SS2_MODE_NORMAL && mode != SS2_MODE_FAST) || (n > 0 && !d_meshes))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_frames: bad arguments (2 <= views <= 4 in one fused pass)");
  for (int v = 0; v < nviews; ++v)
    if (n > 0 && !h_imgs[v]) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_frames: null image pointer %d", v);
  const float xmin = h_minmax2[0], ymin = h_minmax2[2], out_w = h_minmax2[1] - h_minmax2[0], out_h = h_minmax2[3] - h_minmax2[2];
  const int Ho = (int)out_h, Wo = (int)out_w;
  if (Ho < 0 || Wo < 0) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_frames: negative canvas");
  if (n == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  if (!d_out) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_nview_frames: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (tps != SS2_TPS_LATTICE || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  const int chunk = n < 8 ? n : 8;
  const size_t m = (size_t)n * nviews * SS2_NPT * 2, iplane3 = (size_t)3 * H * W, oplane3 = (size_t)3 * Ho * Wo;
  TpsScratch sc;
  SS2_TRY(tps_scratch_alloc(ctx, nviews * chunk, Ho, Wo, tps, 2 * m + 64, &sc, st));
  float* source = sc.nodes + (tps == SS2_TPS_LATTICE ? (tps_lattice_workspace_floats(nviews * chunk, Ho, Wo) + 63) / 64 * 64 : 0);
  float* target = source + m;
  int rc = nview_sources_launch(ctx, d_meshes, nviews, n, H, W, xmin, ymin, out_w, out_h, source, target, st);
  for (int k0 = 0; k0 < n && rc == SS2_OK; k0 += chunk) {
    const int nk = n - k0 < chunk ? n - k0 : chunk;
    const float* src = source + (size_t)k0 * nviews * SS2_NPT * 2;
    const float* tgt = target + (size_t)k0 * nviews * SS2_NPT * 2;
    const float* imgs[4];
    for (int v = 0; v < nviews; ++v) imgs[v] = h_imgs[v] + (size_t)k0 * iplane3;
    ss2_prof_begin(ctx, SS2_PROF_WARP, st);
    rc = tps_solve_for_warp(ctx, src, tgt, nviews * nk, H, W, Ho, Wo, mode, tps, sc, st);
    if (rc == SS2_OK)
      rc = tps_warp_blend_n_launch(ctx, imgs, nviews, src, sc.T, nk, H, W, Ho, Wo, mode, tps, d_out + (size_t)k0 * oplane3, st,
                                   sc.aux, sc.nodes);
    ss2_prof_end(ctx, SS2_PROF_WARP, st, (double)nk * ((double)nviews * 3 * H * W + 3.0 * Ho * Wo) * 4.0);
  }
  cudaFreeAsync(sc.base, st);
  return rc;
}

int ss2_cost_volume_nhwc(ss2_ctx* ctx, const float* d_x1, const float* d_x2, int B, int H, int W, int C, int sr, int CP,
                         float* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (B < 0 || H <= 0 || W <= 0 || sr < 0 || (B > 0 && (!d_x1 || !d_x2 || !d_out)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_cost_volume_nhwc: bad arguments");
  ActRef o;
  o.v = d_out;
  return cost_volume_launch(ctx, d_x1, d_x2, B, H, W, C, sr, CP, o, (cudaStream_t)stream);
}

int ss2_ccl_nhwc(ss2_ctx* ctx, const float* d_f1, const float* d_f2, int B, int H, int W, int C, float* d_flow,
                 void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (B < 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || (B > 0 && (!d_f1 || !d_f2 || !d_flow)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_ccl_nhwc: bad arguments (C must be a multiple of 4)");
  const size_t hw = (size_t)H * W, kp = (hw + 63) / 64 * 64;
  SS2_TRY(ss2_workspace_enter(ctx, (cudaStream_t)stream));
  SS2_TRY(ss2_ensure_arena(ctx, (size_t)B * (6 * hw * C + 9 * C * kp + hw * kp) * sizeof(float) + (1 << 20)));
  ctx->arena.reset();
  return ccl_launch(ctx, d_f1, d_f2, B, H, W, C, d_flow, (cudaStream_t)stream);
}

int ss2_tsmotion(ss2_ctx* ctx, const float* d_smotion, const float* d_tmotion, int n, int first_is_stream_start,
                 const float* d_smotion_prev, float* d_smesh, float* d_tsmotion, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || (n > 0 && (!d_smotion || !d_tmotion || !d_smesh || !d_tsmotion)) ||
      (n > 0 && !first_is_stream_start && !d_smotion_prev))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_tsmotion: bad arguments");
  if (n == 0) return SS2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* buf = nullptr;
  const size_t m = (size_t)n * SS2_NPT * 2;
  SS2_CUDA(ctx, cudaMallocAsync((void**)&buf, (4 * m + (size_t)n * 2 * SS2_NSYS) * sizeof(float), st));
  float *point = buf, *source = buf + m, *target = buf + 2 * m, *moved = buf + 3 * m, *T = buf + 4 * m;
  int rc = tsmotion_prep_launch(ctx, d_smotion, d_tmotion, n, first_is_stream_start, d_smotion_prev, d_smesh, point,
                                source, target, st);
  if (rc == SS2_OK) rc = tps_solve_launch(ctx, source, target, n, T, st);
  if (rc == SS2_OK) rc = tps_point_launch(ctx, point, source, T, n, moved, st);
  if (rc == SS2_OK) rc = tsmotion_finish_launch(ctx, moved, d_smesh, n, first_is_stream_start, d_tsmotion, st);
  cudaFreeAsync(buf, st);
  return rc;
}

int ss2_canvas_minmax(ss2_ctx* ctx, const float* d_mesh1, const float* d_mesh2, int n, int img_h, int img_w,
                      float* d_minmax, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n <= 0 || !d_mesh1 || !d_mesh2 || !d_minmax) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_canvas_minmax: bad arguments");
  return canvas_minmax_launch(ctx, d_mesh1, d_mesh2, n, img_h, img_w, d_minmax, (cudaStream_t)stream);
}

int ss2_canvas_size(const float* h_minmax, int* out_h, int* out_w) {
  if (!h_minmax || !out_h || !out_w) return SS2_ERR_INVALID;
  const float ow = h_minmax[1] - h_minmax[0], oh = h_minmax[3] - h_minmax[2];
  *out_w = (int)ow;  // torch .int(): truncation
  *out_h = (int)oh;
  return SS2_OK;
}

static int stable_frames_impl(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1, const float* d_mesh2,
                              int n, int H, int W, const float* h_minmax, int mode, int tps, float* d_out, unsigned char* d_out8,
                              void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || H <= 0 || W <= 0 || !h_minmax || (n > 0 && (!d_hr1 || !d_hr2 || !d_mesh1 || !d_mesh2)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames: bad arguments");
  int Ho, Wo;
  ss2_canvas_size(h_minmax, &Ho, &Wo);
  if (Ho < 0 || Wo < 0) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames: negative canvas");
  if (n == 0 || Ho == 0 || Wo == 0) return SS2_OK;
  if (!d_out && !d_out8) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const float out_w = h_minmax[1] - h_minmax[0], out_h = h_minmax[3] - h_minmax[2];
  if (tps != SS2_TPS_LATTICE || !tps_lattice_supported(Ho, Wo)) tps = SS2_TPS_EXACT;
  if (d_out8 && tps != SS2_TPS_LATTICE)
    return ss2_fail(ctx, SS2_ERR_UNSUPPORTED, "ss2_stable_frames_u8: the uint8 store is fused into the lattice resampler only "
                    "(canvas %dx%d / tps mode): use ss2_stable_frames + ss2_frames_to_u8", Ho, Wo);
  const size_t m = (size_t)n * 2 * SS2_NPT * 2;
  TpsScratch sc;
  SS2_TRY(tps_scratch_alloc(ctx, 2 * n, Ho, Wo, tps, 2 * m, &sc, st));
  float *source = sc.nodes + (tps == SS2_TPS_LATTICE ? (tps_lattice_workspace_floats(2 * n, Ho, Wo) + 63) / 64 * 64 : 0);
  float* target = source + m;
  // SS2_PROF_WARP brackets everything K14 needs per chunk: canvas-normalised meshes, the fp64 TPS
  // solves, the lattice nodes and the fused resample + blend kernel
  ss2_prof_begin(ctx, SS2_PROF_WARP, st);
  int rc = stable_meshes_launch(ctx, d_mesh1, d_mesh2, n, H, W, h_minmax[0], h_minmax[2], out_w, out_h, source, target, st);
  if (rc == SS2_OK) rc = tps_solve_for_warp(ctx, source, target, 2 * n, H, W, Ho, Wo, mode, tps, sc, st);
  if (rc == SS2_OK) rc = tps_warp_blend_launch(ctx, d_hr1, d_hr2, source, sc.T, n, H, W, Ho, Wo, mode, tps, d_out, st, sc.aux, sc.nodes, d_out8);
  // algorithmic bytes of the bracket: both sources once + the fused frame once (fp32 canvas, or uint8 when the back end is fused)
  ss2_prof_end(ctx, SS2_PROF_WARP, st, (double)n * (2.0 * 3 * H * W * 4.0 + 3.0 * Ho * Wo * (d_out8 ? 1.0 : 4.0)));
  cudaFreeAsync(sc.base, st);
  return rc;
}

int ss2_stable_frames(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1, const float* d_mesh2,
                      int n, int H, int W, const float* h_minmax, int mode, int tps, float* d_out, void* stream) {
  return stable_frames_impl(ctx, d_hr1, d_hr2, d_mesh1, d_mesh2, n, H, W, h_minmax, mode, tps, d_out, nullptr, stream);
}

int ss2_stable_frames_u8(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1, const float* d_mesh2,
                         int n, int H, int W, const float* h_minmax, int mode, int tps, uint8_t* d_out, void* stream) {
  if (ctx && n > 0 && !d_out) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stable_frames_u8: null output");
  return stable_frames_impl(ctx, d_hr1, d_hr2, d_mesh1, d_mesh2, n, H, W, h_minmax, mode, tps, nullptr, d_out, stream);
}

}  // extern "C"
