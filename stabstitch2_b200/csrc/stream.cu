// Whole-stream entry points: everything the reference's test() does between frame loading
// and video writing (Full_model_inference/Codes/test_online_tra.py:284-399, AVERAGE fusion),
// as stream-ordered launches with a single 16-byte D2H (the data-dependent canvas size).
#include "common.cuh"

// smooth meshes of a stream from per-window outputs (test_online_tra.py:378-392):
//   with_head:  S[k] = win0[k] for k < 7, S[k] = win_{k-6}[6] for k >= 7  (nwin+6 frames)
//   otherwise:  S[i] = win_i[6]                                            (nwin frames)
__global__ void assemble_smooth_kernel(const float* __restrict__ win, int nwin, int with_head, float* __restrict__ out) {
  const int k = blockIdx.x, tid = threadIdx.x;
  if (tid >= SS2_NPT * 2) return;
  int w, t;
  if (with_head) {
    if (k < SS2_WINDOW) { w = 0; t = k; } else { w = k - (SS2_WINDOW - 1); t = SS2_WINDOW - 1; }
  } else {
    w = k; t = SS2_WINDOW - 1;
  }
  out[(size_t)k * SS2_NPT * 2 + tid] = win[((size_t)w * SS2_WINDOW + t) * SS2_NPT * 2 + tid];
}

static int named_buf(ss2_ctx* ctx, const char* name, size_t bytes, float** out) {
  auto& e = ctx->stream_bufs[name];
  if (e.second < bytes) {
    if (e.first) { SS2_CUDA(ctx, cudaDeviceSynchronize()); SS2_CUDA(ctx, cudaFree(e.first)); e.first = nullptr; e.second = 0; }
    void* p = nullptr;
    cudaError_t err = cudaMalloc(&p, bytes);
    if (err != cudaSuccess) return ss2_fail(ctx, SS2_ERR_OOM, "cudaMalloc(%zu) for '%s': %s", bytes, name, cudaGetErrorString(err));
    e.first = p; e.second = bytes;
  }
  *out = (float*)e.first;
  return SS2_OK;
}

extern "C" int ss2_assemble_smooth(ss2_ctx* ctx, const float* d_win_smooth, int nwin, int with_head, float* d_out,
                                   void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (nwin <= 0 || !d_win_smooth || !d_out) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_assemble_smooth: bad arguments");
  const int nf = with_head ? nwin + SS2_WINDOW - 1 : nwin;
  assemble_smooth_kernel<<<nf, 128, 0, (cudaStream_t)stream>>>(d_win_smooth, nwin, with_head, d_out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// lr1, lr2 [n,3,360,480] device -> smooth meshes [n,7,9,2] x2 (+ optional raw outputs)
extern "C" int ss2_stream_meshes(ss2_ctx* ctx, const float* d_lr1, const float* d_lr2, int n, float* d_smooth1,
                                 float* d_smooth2, float* d_smotion1, float* d_smotion2, float* d_tmotion1,
                                 float* d_tmotion2, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < SS2_WINDOW) return ss2_fail(ctx, SS2_ERR_INVALID, "a stream needs at least %d frames (got %d)", SS2_WINDOW, n);
  if (!d_lr1 || !d_lr2 || !d_smooth1 || !d_smooth2) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stream_meshes: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)n * SS2_NPT * 2;
  const int nwin = n - (SS2_WINDOW - 1);
  const size_t wm = (size_t)nwin * SS2_WINDOW * SS2_NPT * 2;
  float* buf;
  SS2_TRY(ss2_workspace_enter(ctx, st));
  SS2_TRY(named_buf(ctx, "stream_meshes", (8 * m + 2 * wm) * sizeof(float), &buf));
  float *sm1 = buf, *sm2 = buf + m, *tm1 = buf + 2 * m, *tm2 = buf + 3 * m;
  float *mesh1 = buf + 4 * m, *mesh2 = buf + 5 * m, *ts1 = buf + 6 * m, *ts2 = buf + 7 * m;
  float *w1 = buf + 8 * m, *w2 = w1 + wm;
  SS2_TRY(ss2_build_spatial_temporal(ctx, d_lr1, d_lr2, n, 0, sm1, sm2, tm1, tm2, stream));
  SS2_TRY(ss2_tsmotion(ctx, sm1, tm1, n, 1, nullptr, mesh1, ts1, stream));
  SS2_TRY(ss2_tsmotion(ctx, sm2, tm2, n, 1, nullptr, mesh2, ts2, stream));
  SS2_TRY(ss2_build_smooth(ctx, ts1, ts2, mesh1, mesh2, nwin, 1, nullptr, nullptr, nullptr, w1, nullptr, nullptr, nullptr,
                           w2, stream));
  SS2_TRY(ss2_assemble_smooth(ctx, w1, nwin, 1, d_smooth1, stream));
  SS2_TRY(ss2_assemble_smooth(ctx, w2, nwin, 1, d_smooth2, stream));
  const size_t mb = m * sizeof(float);
  if (d_smotion1) SS2_CUDA(ctx, cudaMemcpyAsync(d_smotion1, sm1, mb, cudaMemcpyDeviceToDevice, st));
  if (d_smotion2) SS2_CUDA(ctx, cudaMemcpyAsync(d_smotion2, sm2, mb, cudaMemcpyDeviceToDevice, st));
  if (d_tmotion1) SS2_CUDA(ctx, cudaMemcpyAsync(d_tmotion1, tm1, mb, cudaMemcpyDeviceToDevice, st));
  if (d_tmotion2) SS2_CUDA(ctx, cudaMemcpyAsync(d_tmotion2, tm2, mb, cudaMemcpyDeviceToDevice, st));
  return SS2_OK;
}

#define WARP_CHUNK 8
#define HOST_SLOTS 3

// per-slot streams / events of the host-buffer pipeline
struct HostSlot {
  cudaStream_t s_copy = nullptr;   // host -> device (inputs)
  cudaStream_t s_d2h = nullptr;    // device -> host (frames): a stream of its own, so that the NEXT chunk's inputs
                                   // are not queued behind this chunk's output copy
  cudaEvent_t ev_hr = nullptr, ev_chunk[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr}, ev_done = nullptr;
  cudaEvent_t ev_lr = nullptr;     // network inputs on the device
  cudaEvent_t ev_warp = nullptr;   // resampling of the slot's chunk done: its device input buffers may be overwritten
  bool busy = false, warped = false;
  // inputs already on their way (ss2_stitch_stream_host_prefetch)
  const void* pre[4] = {nullptr, nullptr, nullptr, nullptr};
  int pre_n = 0, pre_h = 0, pre_w = 0, pre_u8 = 0;
  // a submitted, not yet finished chunk (ss2_stitch_stream_host*_submit / _finish)
  bool submitted = false;
  int sub_n = 0, sub_h = 0, sub_w = 0, sub_u8 = 0;
  float *sub_hr1 = nullptr, *sub_hr2 = nullptr, *sub_s1 = nullptr, *sub_s2 = nullptr;
  cudaEvent_t ev_canvas = nullptr;
  float* h_mm = nullptr;   // pinned: canvas min/max of the submitted chunk
};
// the slots belong to the context (distinct contexts are independent, also on one device)
static HostSlot* ctx_slots(ss2_ctx* ctx) {
  if (!ctx->host_slots) ctx->host_slots = new HostSlot[HOST_SLOTS];
  return static_cast<HostSlot*>(ctx->host_slots);
}

void ss2_host_slots_free(ss2_ctx* ctx) {
  if (!ctx->host_slots) return;
  HostSlot* hs = static_cast<HostSlot*>(ctx->host_slots);
  for (int i = 0; i < HOST_SLOTS; ++i) {
    HostSlot& h = hs[i];
    if (!h.s_copy) continue;
    cudaStreamDestroy(h.s_copy);
    cudaStreamDestroy(h.s_d2h);
    if (h.h_mm) cudaFreeHost(h.h_mm);
    cudaEvent_t evs[] = {h.ev_hr, h.ev_lr, h.ev_warp, h.ev_done, h.ev_canvas, h.ev_chunk[0], h.ev_chunk[1], h.ev_d2h[0], h.ev_d2h[1]};
    for (cudaEvent_t e : evs)
      if (e) cudaEventDestroy(e);
  }
  delete[] hs;
  ctx->host_slots = nullptr;
}

static int slot_init(ss2_ctx* ctx, HostSlot& h) {
  if (h.s_copy) return SS2_OK;
  SS2_CUDA(ctx, cudaStreamCreateWithFlags(&h.s_copy, cudaStreamNonBlocking));
  SS2_CUDA(ctx, cudaStreamCreateWithFlags(&h.s_d2h, cudaStreamNonBlocking));
  SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_hr, cudaEventDisableTiming));
  SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_lr, cudaEventDisableTiming));
  SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_warp, cudaEventDisableTiming));
  SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_done, cudaEventDisableTiming));
  SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_canvas, cudaEventDisableTiming));
  SS2_CUDA(ctx, cudaMallocHost((void**)&h.h_mm, 4 * sizeof(float)));
  for (int i = 0; i < 2; ++i) SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_chunk[i], cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) SS2_CUDA(ctx, cudaEventCreateWithFlags(&h.ev_d2h[i], cudaEventDisableTiming));
  return SS2_OK;
}

// device input buffers of a slot + the H2D copies.  fp32 interface: network inputs first, then the much larger hr frames.
// uint8 interface (u8 != 0): only the decoded BGR frames travel (h_a, h_b [n,H,W,3]); hr / lr are made on the device.
static int slot_upload(ss2_ctx* ctx, HostSlot& hs, int slot, int u8, const void* h_lr1, const void* h_lr2, const void* h_a,
                       const void* h_b, int n, int H, int W, float** lr1, float** lr2, float** hr1, float** hr2,
                       unsigned char** ua, unsigned char** ub, bool enqueue) {
  const size_t lrb = (size_t)n * 3 * 360 * 480 * sizeof(float), hrb = (size_t)n * 3 * H * W * sizeof(float);
  const size_t u8b = (size_t)n * 3 * H * W;
  char nm[32];
  auto name = [&](const char* base) { snprintf(nm, sizeof(nm), "%s.%d", base, slot); return nm; };
  SS2_TRY(named_buf(ctx, name("lr1"), lrb, lr1));
  SS2_TRY(named_buf(ctx, name("lr2"), lrb, lr2));
  SS2_TRY(named_buf(ctx, name("hr1"), hrb, hr1));
  SS2_TRY(named_buf(ctx, name("hr2"), hrb, hr2));
  *ua = *ub = nullptr;
  if (u8) {
    float *pa, *pb;
    SS2_TRY(named_buf(ctx, name("u8a"), u8b, &pa));
    SS2_TRY(named_buf(ctx, name("u8b"), u8b, &pb));
    *ua = (unsigned char*)pa; *ub = (unsigned char*)pb;
  }
  if (!enqueue) return SS2_OK;
  cudaStream_t sx = hs.s_copy;
  if (hs.warped) SS2_CUDA(ctx, cudaStreamWaitEvent(sx, hs.ev_warp, 0));  // the previous chunk of this slot still reads them
  if (u8) {
    SS2_CUDA(ctx, cudaMemcpyAsync(*ua, h_a, u8b, cudaMemcpyHostToDevice, sx));
    SS2_CUDA(ctx, cudaMemcpyAsync(*ub, h_b, u8b, cudaMemcpyHostToDevice, sx));
    SS2_CUDA(ctx, cudaEventRecord(hs.ev_lr, sx));
    SS2_CUDA(ctx, cudaEventRecord(hs.ev_hr, sx));
    return SS2_OK;
  }
  SS2_CUDA(ctx, cudaMemcpyAsync(*lr1, h_lr1, lrb, cudaMemcpyHostToDevice, sx));
  SS2_CUDA(ctx, cudaMemcpyAsync(*lr2, h_lr2, lrb, cudaMemcpyHostToDevice, sx));
  SS2_CUDA(ctx, cudaEventRecord(hs.ev_lr, sx));
  SS2_CUDA(ctx, cudaMemcpyAsync(*hr1, h_a, hrb, cudaMemcpyHostToDevice, sx));
  SS2_CUDA(ctx, cudaMemcpyAsync(*hr2, h_b, hrb, cudaMemcpyHostToDevice, sx));
  SS2_CUDA(ctx, cudaEventRecord(hs.ev_hr, sx));
  return SS2_OK;
}

static int prefetch_impl(ss2_ctx* ctx, int slot, int u8, const void* h_lr1, const void* h_lr2, const void* h_a,
                         const void* h_b, int n, int H, int W) {
  if (!ctx) return SS2_ERR_INVALID;
  if (slot < 0 || slot >= HOST_SLOTS) return ss2_fail(ctx, SS2_ERR_INVALID, "bad slot");
  if ((!u8 && (!h_lr1 || !h_lr2)) || !h_a || !h_b || n < SS2_WINDOW || H <= 0 || W <= 0)
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stitch_stream_host_prefetch: bad arguments");
  SS2_CUDA(ctx, cudaSetDevice(ctx->device));
  HostSlot& hs = ctx_slots(ctx)[slot];
  SS2_TRY(slot_init(ctx, hs));
  float *lr1, *lr2, *hr1, *hr2;
  unsigned char *ua, *ub;
  SS2_TRY(slot_upload(ctx, hs, slot, u8, h_lr1, h_lr2, h_a, h_b, n, H, W, &lr1, &lr2, &hr1, &hr2, &ua, &ub, true));
  hs.pre[0] = h_lr1; hs.pre[1] = h_lr2; hs.pre[2] = h_a; hs.pre[3] = h_b;
  hs.pre_n = n; hs.pre_h = H; hs.pre_w = W; hs.pre_u8 = u8;
  return SS2_OK;
}

// Starts the host -> device copies of a slot's NEXT chunk and returns at once; the following
// ss2_stitch_stream_host_async on that slot with the same pointers and sizes uses them instead of copying again.
// Call it before the _async of the chunk in flight: the upload then runs underneath that chunk's networks
// instead of after the host has waited for them (the canvas size is a data-dependent host read).
extern "C" int ss2_stitch_stream_host_prefetch(ss2_ctx* ctx, int slot, const float* h_lr1, const float* h_lr2,
                                               const float* h_hr1, const float* h_hr2, int n, int H, int W) {
  return prefetch_impl(ctx, slot, 0, h_lr1, h_lr2, h_hr1, h_hr2, n, H, W);
}

extern "C" int ss2_stitch_stream_host_u8_prefetch(ss2_ctx* ctx, int slot, const uint8_t* h_bgr1, const uint8_t* h_bgr2,
                                                  int n, int H, int W) {
  return prefetch_impl(ctx, slot, 1, nullptr, nullptr, h_bgr1, h_bgr2, n, H, W);
}

extern "C" int ss2_stitch_stream_host_wait(ss2_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= HOST_SLOTS) return SS2_ERR_INVALID;
  HostSlot& h = ctx_slots(ctx)[slot];
  if (!h.busy) return SS2_OK;
  SS2_CUDA(ctx, cudaEventSynchronize(h.ev_done));
  h.busy = false;
  return SS2_OK;
}

// One chunk through the host pipeline in two halves; u8 != 0: uint8 frames in ([n,H,W,3] BGR) and out ([n,Ho,Wo,3]).
//   submit: H2D of the inputs (unless prefetched), device front end, the three networks, canvas min/max and its
//           16-byte D2H - everything is only ENQUEUED, the call does not block;
//   finish: waits for the canvas (the output shape is data dependent), enqueues resample + blend (+ astype(uint8))
//           and the D2H of the frames.
// A caller that submits chunk k+1 BEFORE it finishes chunk k keeps the GPU busy while the host waits for chunk k's
// canvas (stream order on the compute stream: nets(k), nets(k+1), warp(k), nets(k+2), ..).
static int submit_impl(ss2_ctx* ctx, int slot, int u8, const void* h_lr1, const void* h_lr2, const void* h_a, const void* h_b,
                       int n, int H, int W) {
  if (!ctx) return SS2_ERR_INVALID;
  if (slot < 0 || slot >= HOST_SLOTS) return ss2_fail(ctx, SS2_ERR_INVALID, "bad slot");
  if (n < SS2_WINDOW) return ss2_fail(ctx, SS2_ERR_INVALID, "a stream needs at least %d frames (got %d)", SS2_WINDOW, n);
  if ((!u8 && (!h_lr1 || !h_lr2)) || !h_a || !h_b || H <= 1 || W <= 1)
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stitch_stream_host: bad arguments");
  SS2_CUDA(ctx, cudaSetDevice(ctx->device));
  HostSlot& hs = ctx_slots(ctx)[slot];
  SS2_TRY(slot_init(ctx, hs));
  if (hs.submitted) return ss2_fail(ctx, SS2_ERR_INVALID, "slot %d: a submitted chunk has not been finished", slot);
  SS2_TRY(ss2_stitch_stream_host_wait(ctx, slot));  // the slot's buffers must be free
  if (!ctx->s_compute) SS2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
  cudaStream_t sc = ctx->s_compute;
  const size_t mb = (size_t)n * SS2_NPT * 2 * sizeof(float);
  float *lr1, *lr2, *hr1, *hr2, *small;
  unsigned char *ua, *ub;
  char nm[32];
  auto name = [&](const char* base) { snprintf(nm, sizeof(nm), "%s.%d", base, slot); return nm; };
  const bool prefetched = hs.pre[0] == h_lr1 && hs.pre[1] == h_lr2 && hs.pre[2] == h_a && hs.pre[3] == h_b &&
                          hs.pre_n == n && hs.pre_h == H && hs.pre_w == W && hs.pre_u8 == u8;
  SS2_TRY(slot_upload(ctx, hs, slot, u8, h_lr1, h_lr2, h_a, h_b, n, H, W, &lr1, &lr2, &hr1, &hr2, &ua, &ub, !prefetched));
  hs.pre[0] = hs.pre[1] = hs.pre[2] = hs.pre[3] = nullptr;
  SS2_TRY(named_buf(ctx, name("small"), 2 * mb + 64, &small));
  float *S1 = small, *S2 = small + (size_t)n * SS2_NPT * 2, *mm = S2 + (size_t)n * SS2_NPT * 2;
  // the compute stream only waits for the network inputs before it starts the networks
  SS2_CUDA(ctx, cudaStreamWaitEvent(sc, hs.ev_lr, 0));
  if (u8) {
    // device front end (test_online_tra.py:252-264): fp32 planar hr frames and the cv2-resized network inputs
    SS2_TRY(load_frames_u8_launch(ctx, ua, n, H, W, hr1, lr1, sc));
    SS2_TRY(load_frames_u8_launch(ctx, ub, n, H, W, hr2, lr2, sc));
  }
  SS2_TRY(ss2_stream_meshes(ctx, lr1, lr2, n, S1, S2, nullptr, nullptr, nullptr, nullptr, sc));
  SS2_TRY(canvas_minmax_launch(ctx, S1, S2, n, H, W, mm, sc));
  SS2_CUDA(ctx, cudaMemcpyAsync(hs.h_mm, mm, 4 * sizeof(float), cudaMemcpyDeviceToHost, sc));
  SS2_CUDA(ctx, cudaEventRecord(hs.ev_canvas, sc));
  hs.submitted = true;
  hs.sub_n = n; hs.sub_h = H; hs.sub_w = W; hs.sub_u8 = u8;
  hs.sub_hr1 = hr1; hs.sub_hr2 = hr2; hs.sub_s1 = S1; hs.sub_s2 = S2;
  return SS2_OK;
}

static int finish_impl(ss2_ctx* ctx, int slot, int mode, int tps, void* h_out, int64_t out_capacity, int* out_h, int* out_w,
                       float* h_smooth_mesh1, float* h_smooth_mesh2) {
  if (!ctx) return SS2_ERR_INVALID;
  if (slot < 0 || slot >= HOST_SLOTS) return ss2_fail(ctx, SS2_ERR_INVALID, "bad slot");
  if (!h_out || !out_h || !out_w) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stitch_stream_host: bad arguments");
  HostSlot& hs = ctx_slots(ctx)[slot];
  if (!hs.submitted) return ss2_fail(ctx, SS2_ERR_INVALID, "slot %d: nothing submitted", slot);
  SS2_CUDA(ctx, cudaSetDevice(ctx->device));
  hs.submitted = false;
  const int n = hs.sub_n, H = hs.sub_h, W = hs.sub_w, u8 = hs.sub_u8;
  float *hr1 = hs.sub_hr1, *hr2 = hs.sub_hr2, *S1 = hs.sub_s1, *S2 = hs.sub_s2;
  cudaStream_t sc = ctx->s_compute, sx = hs.s_d2h;
  const size_t mb = (size_t)n * SS2_NPT * 2 * sizeof(float);
  char nm[32];
  auto name = [&](const char* base) { snprintf(nm, sizeof(nm), "%s.%d", base, slot); return nm; };
  SS2_CUDA(ctx, cudaEventSynchronize(hs.ev_canvas));  // the canvas size is data dependent (16-byte read)
  const float* h_mm = hs.h_mm;
  int Ho, Wo;
  ss2_canvas_size(h_mm, &Ho, &Wo);
  *out_h = Ho; *out_w = Wo;
  if (Ho <= 0 || Wo <= 0) return ss2_fail(ctx, SS2_ERR_INVALID, "degenerate canvas %dx%d", Ho, Wo);
  const size_t fpx = (size_t)3 * Ho * Wo;
  if ((int64_t)(fpx * n) > out_capacity)
    return ss2_fail(ctx, SS2_ERR_INVALID, "output needs %zu elements, capacity %lld", fpx * n, (long long)out_capacity);
  if (h_smooth_mesh1 || h_smooth_mesh2) {   // the meshes are complete once the canvas event has fired
    SS2_CUDA(ctx, cudaStreamWaitEvent(sx, hs.ev_canvas, 0));
    if (h_smooth_mesh1) SS2_CUDA(ctx, cudaMemcpyAsync(h_smooth_mesh1, S1, mb, cudaMemcpyDeviceToHost, sx));
    if (h_smooth_mesh2) SS2_CUDA(ctx, cudaMemcpyAsync(h_smooth_mesh2, S2, mb, cudaMemcpyDeviceToHost, sx));
  }
  // The whole chunk is resampled into a device buffer of its own (HBM is plentiful), so the compute
  // stream is free for the next chunk's networks while the copy stream drains the frames to the host.
  // uint8 interface: the resampler stores uint8 HWC itself (ss2_stable_frames_u8) where the lattice resampler runs; the
  // fp32 canvas + conversion pass is the fallback for EXACT mode and small canvases.
  const bool fused_u8 = u8 && tps == SS2_TPS_LATTICE && tps_lattice_supported(Ho, Wo);
  float *obuf = nullptr, *obuf8f = nullptr;
  if (!fused_u8) SS2_TRY(named_buf(ctx, name("out_frames"), (size_t)n * fpx * sizeof(float), &obuf));
  if (u8) SS2_TRY(named_buf(ctx, name("out_u8"), (size_t)n * fpx, &obuf8f));
  unsigned char* obuf8 = (unsigned char*)obuf8f;
  SS2_CUDA(ctx, cudaStreamWaitEvent(sc, hs.ev_hr, 0));
  for (int f0 = 0; f0 < n; f0 += WARP_CHUNK) {
    const int nf = n - f0 < WARP_CHUNK ? n - f0 : WARP_CHUNK;
    float* dst = fused_u8 ? nullptr : obuf + (size_t)f0 * fpx;
    if (fused_u8) {
      SS2_TRY(ss2_stable_frames_u8(ctx, hr1 + (size_t)f0 * 3 * H * W, hr2 + (size_t)f0 * 3 * H * W, S1 + (size_t)f0 * SS2_NPT * 2,
                                   S2 + (size_t)f0 * SS2_NPT * 2, nf, H, W, h_mm, mode, tps, obuf8 + (size_t)f0 * fpx, sc));
    } else {
      SS2_TRY(ss2_stable_frames(ctx, hr1 + (size_t)f0 * 3 * H * W, hr2 + (size_t)f0 * 3 * H * W, S1 + (size_t)f0 * SS2_NPT * 2,
                                S2 + (size_t)f0 * SS2_NPT * 2, nf, H, W, h_mm, mode, tps, dst, sc));
      if (u8) SS2_TRY(frames_to_u8_launch(ctx, dst, nf, Ho, Wo, obuf8 + (size_t)f0 * fpx, sc));   // :152,414 astype(uint8)
    }
    SS2_CUDA(ctx, cudaEventRecord(hs.ev_chunk[0], sc));
    SS2_CUDA(ctx, cudaStreamWaitEvent(sx, hs.ev_chunk[0], 0));
    if (u8)
      SS2_CUDA(ctx, cudaMemcpyAsync((unsigned char*)h_out + (size_t)f0 * fpx, obuf8 + (size_t)f0 * fpx, (size_t)nf * fpx,
                                    cudaMemcpyDeviceToHost, sx));
    else
      SS2_CUDA(ctx, cudaMemcpyAsync((float*)h_out + (size_t)f0 * fpx, dst, (size_t)nf * fpx * sizeof(float),
                                    cudaMemcpyDeviceToHost, sx));
  }
  SS2_CUDA(ctx, cudaEventRecord(hs.ev_warp, sc));
  hs.warped = true;
  SS2_CUDA(ctx, cudaEventRecord(hs.ev_done, sx));
  hs.busy = true;
  return SS2_OK;
}

static int async_impl(ss2_ctx* ctx, int slot, int u8, const void* h_lr1, const void* h_lr2, const void* h_a, const void* h_b,
                      int n, int H, int W, int mode, int tps, void* h_out, int64_t out_capacity, int* out_h, int* out_w,
                      float* h_smooth_mesh1, float* h_smooth_mesh2) {
  if (!ctx) return SS2_ERR_INVALID;
  if (!h_out || !out_h || !out_w) return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_stitch_stream_host: bad arguments");
  SS2_TRY(submit_impl(ctx, slot, u8, h_lr1, h_lr2, h_a, h_b, n, H, W));
  return finish_impl(ctx, slot, mode, tps, h_out, out_capacity, out_h, out_w, h_smooth_mesh1, h_smooth_mesh2);
}

extern "C" int ss2_stitch_stream_host_submit(ss2_ctx* ctx, int slot, const float* h_lr1, const float* h_lr2,
                                             const float* h_hr1, const float* h_hr2, int n, int H, int W) {
  return submit_impl(ctx, slot, 0, h_lr1, h_lr2, h_hr1, h_hr2, n, H, W);
}
extern "C" int ss2_stitch_stream_host_u8_submit(ss2_ctx* ctx, int slot, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n,
                                                int H, int W) {
  return submit_impl(ctx, slot, 1, nullptr, nullptr, h_bgr1, h_bgr2, n, H, W);
}
extern "C" int ss2_stitch_stream_host_finish(ss2_ctx* ctx, int slot, int mode, int tps, void* h_out, int64_t out_capacity,
                                             int* out_h, int* out_w, float* h_smooth_mesh1, float* h_smooth_mesh2) {
  return finish_impl(ctx, slot, mode, tps, h_out, out_capacity, out_h, out_w, h_smooth_mesh1, h_smooth_mesh2);
}

extern "C" int ss2_stitch_stream_host_async(ss2_ctx* ctx, int slot, const float* h_lr1, const float* h_lr2,
                                            const float* h_hr1, const float* h_hr2, int n, int H, int W, int mode, int tps,
                                            float* h_out, int64_t out_capacity, int* out_h, int* out_w,
                                            float* h_smooth_mesh1, float* h_smooth_mesh2) {
  return async_impl(ctx, slot, 0, h_lr1, h_lr2, h_hr1, h_hr2, n, H, W, mode, tps, h_out, out_capacity, out_h, out_w,
                    h_smooth_mesh1, h_smooth_mesh2);
}

extern "C" int ss2_stitch_stream_host_u8_async(ss2_ctx* ctx, int slot, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n,
                                               int H, int W, int mode, int tps, uint8_t* h_out, int64_t out_capacity,
                                               int* out_h, int* out_w, float* h_smooth_mesh1, float* h_smooth_mesh2) {
  return async_impl(ctx, slot, 1, nullptr, nullptr, h_bgr1, h_bgr2, n, H, W, mode, tps, h_out, out_capacity, out_h, out_w,
                    h_smooth_mesh1, h_smooth_mesh2);
}

extern "C" int ss2_stitch_stream_host(ss2_ctx* ctx, const float* h_lr1, const float* h_lr2, const float* h_hr1,
                                      const float* h_hr2, int n, int H, int W, int mode, int tps, float* h_out,
                                      int64_t out_capacity, int* out_h, int* out_w, float* h_smooth_mesh1,
                                      float* h_smooth_mesh2) {
  SS2_TRY(ss2_stitch_stream_host_async(ctx, 0, h_lr1, h_lr2, h_hr1, h_hr2, n, H, W, mode, tps, h_out, out_capacity, out_h,
                                       out_w, h_smooth_mesh1, h_smooth_mesh2));
  return ss2_stitch_stream_host_wait(ctx, 0);
}

extern "C" int ss2_stitch_stream_host_u8(ss2_ctx* ctx, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n, int H, int W,
                                         int mode, int tps, uint8_t* h_out, int64_t out_capacity, int* out_h, int* out_w,
                                         float* h_smooth_mesh1, float* h_smooth_mesh2) {
  SS2_TRY(ss2_stitch_stream_host_u8_async(ctx, 0, h_bgr1, h_bgr2, n, H, W, mode, tps, h_out, out_capacity, out_h, out_w,
                                          h_smooth_mesh1, h_smooth_mesh2));
  return ss2_stitch_stream_host_wait(ctx, 0);
}
