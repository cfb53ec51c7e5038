/*
 * ss2.h - C ABI of libss2.so: the B200-native (sm_100a) StabStitch++ inference hot path.
 *
 * The reference (nie-lang/StabStitch2) has no FFI: its boundary is the set of Python names
 * its drivers import (Full_model_inference/Codes/test_online_tra.py:7-9,14-15,21-23).  The
 * Python shim modules in stabstitch2_b200/ keep those names and call the entry points
 * below through ctypes; each entry point cites the reference function it replaces
 * (paths relative to /root/reference/Full_model_inference/Codes/).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - every function returns 0 on success or a negative ss2_status; ss2_last_error(ctx)
 *    gives the message.  Nothing exits or throws across the ABI.
 *  - pointers named d_* are DEVICE pointers (caller-allocated, caller-owned); h_* are HOST
 *    pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - all entry points are asynchronous and stream-ordered unless documented otherwise.
 *  - a context is bound to one device and is NOT thread-safe; distinct contexts are
 *    independent (one per process/GPU in the multi-GPU layout).  The network entry points share
 *    the context's workspace: a call arriving on another stream than the previous one first
 *    waits (event) for the work the previous stream was given.
 *  - images are fp32 NCHW exactly like the reference tensors; meshes are fp32 [..,7,9,2]
 *    with last dim (x, y) (grid_res.py:3-4 -> 63 control points).
 */
#ifndef SS2_H
#define SS2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ss2_ctx ss2_ctx;

typedef enum {
  SS2_OK = 0,
  SS2_ERR_INVALID = -1,      /* bad argument / shape */
  SS2_ERR_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
  SS2_ERR_NO_WEIGHTS = -3,   /* forward called before ss2_finalize_weights */
  SS2_ERR_MISSING_KEY = -4,  /* a state-dict key the network needs was never loaded */
  SS2_ERR_OOM = -5,
  SS2_ERR_UNSUPPORTED = -6
} ss2_status;

enum { SS2_NET_SPATIAL = 0, SS2_NET_TEMPORAL = 1, SS2_NET_SMOOTH = 2 };
enum { SS2_MODE_NORMAL = 0, SS2_MODE_FAST = 1 };   /* torch_tps_transform.py:151-162 */
/* evaluation precision of the dense TPS field inside the resampler */
enum { SS2_TPS_EXACT = 0,    /* all 63 radial terms per pixel (the reference's arithmetic) */
       SS2_TPS_LATTICE = 1   /* far field interpolated on a per-tile lattice, near field exact */ };

#define SS2_GRID_H 6
#define SS2_GRID_W 8
#define SS2_NPT 63
#define SS2_WINDOW 7

/* ---- lifetime ------------------------------------------------------------------------ */
/* Numerics of the convolutions (read from the environment here): fp32-grade by default - every fp32 product is three exact
 * tensor-core products of 11-bit split operands accumulated in fp32; behind the ResNet stem the halves are fp16 planes
 * (low half scaled by 2^11), which hold |v| <= 65504: an activation beyond that raises a flag and the NEXT entry point of
 * the context returns SS2_ERR_UNSUPPORTED naming the cause (once).  SS2_F16=0: TF32 split planes everywhere (no range limit);
 * SS2_TC_PASSES=1: plain TF32 like cuDNN's default; SS2_USE_TC=0: exact fp32 SIMT kernels. */
int ss2_create(int device, ss2_ctx** out);
void ss2_destroy(ss2_ctx* ctx);
const char* ss2_last_error(ss2_ctx* ctx);
const char* ss2_version(void);
/* number of kernel launches issued through this context since creation / last reset */
int64_t ss2_launch_count(ss2_ctx* ctx, int reset);

/* ---- in-library kernel timing (bench.py's roofline leg) -------------------------------- */
/* While enabled, the launches of kernel class `which` are bracketed by CUDA events recorded on
 * the launching stream.  ss2_profile_read synchronises those events and returns the summed
 * duration (ms) and the number of launches since the last ss2_profile_enable. */
enum { SS2_PROF_WARP = 0,   /* the whole resampling of a chunk: canvas meshes + TPS solves + lattice nodes + fused
                               resample/blend kernel (everything SURVEY.md K14 needs) */
       SS2_PROF_CONV = 1,   /* implicit-GEMM convolution / linear kernels */
       SS2_PROF_COUNT = 2 };
int ss2_profile_enable(ss2_ctx* ctx, int which, int enable);
int ss2_profile_read(ss2_ctx* ctx, int which, double* total_ms, int64_t* launches, double* work);

/* ---- weights: replaces nn.Module.load_state_dict (test_online_tra.py:178-190) -------- */
/* Copies one fp32 state-dict tensor (host memory) under its reference key name. */
int ss2_load_tensor(ss2_ctx* ctx, int net_id, const char* key, const float* h_data,
                    const int64_t* shape, int ndim);
/* Folds eval-mode BatchNorm into the convolutions, re-packs to the kernels' layouts and
 * uploads.  Synchronous.  Fails with SS2_ERR_MISSING_KEY naming the first absent key. */
int ss2_finalize_weights(ss2_ctx* ctx, int net_id);

/* ---- geometry --------------------------------------------------------------------------*/
/* utils/torch_DLT.py:17 tensor_DLT: src,dst [bs,4,2] -> H [bs,3,3] */
int ss2_dlt(ss2_ctx* ctx, const float* d_src, const float* d_dst, int bs, float* d_H, void* stream);
/* utils/torch_homo_transform.py:6 transformer: U [bn,C,H,W], theta [bn,3,3] -> [bn,C,Ho,Wo] */
int ss2_homo_warp(ss2_ctx* ctx, const float* d_U, const float* d_theta, int bn, int C, int H, int W,
                  int Ho, int Wo, float* d_out, void* stream);
/* utils/torch_tps_transform_point.py:6 transformer: point,source,target [bn,63,2] -> [bn,63,2] */
int ss2_tps_point(ss2_ctx* ctx, const float* d_point, const float* d_source, const float* d_target,
                  int bn, float* d_out, void* stream);
/* utils/torch_tps_transform.py:7 transformer: U [bn,C,H,W], source,target [bn,63,2]
 * -> [bn,C,Ho,Wo]; mode = SS2_MODE_NORMAL | SS2_MODE_FAST; tps = SS2_TPS_EXACT | _LATTICE */
int ss2_tps_warp(ss2_ctx* ctx, const float* d_U, const float* d_source, const float* d_target,
                 int bn, int C, int H, int W, int Ho, int Wo, int mode, int tps, float* d_out,
                 void* stream);
/* Fused resampler + AVERAGE blend: the body of the get_stable_sqe loop,
 * test_online_tra.py:140-142.  img1,img2 [3,H,W] fp32; source [2,63,2] (view 1, view 2
 * normalised canvas meshes); target [2,63,2]; out [3,Ho,Wo].  One launch for the TPS
 * solves + one for warp/gather/blend.  `nframes` frames are processed in one call:
 * img pointers are [nframes,3,H,W], source/target [nframes,2,63,2], out [nframes,3,Ho,Wo]. */
int ss2_tps_warp_blend_avg(ss2_ctx* ctx, const float* d_img1, const float* d_img2,
                           const float* d_source, const float* d_target, int nframes, int H, int W,
                           int Ho, int Wo, int mode, int tps, float* d_out, void* stream);

/* ---- correlation layers (NHWC activations) ---------------------------------------------- */
/* SpatialNet.cost_volume / TemporalNet.cost_volume (spatial_network.py:333, norm=False):
 * x1,x2 [B,H,W,128] -> out [B,H,W,CP], channels d = j*(2sr+1)+i, CP >= (2sr+1)^2, rest 0 */
int ss2_cost_volume_nhwc(ss2_ctx* ctx, const float* d_x1, const float* d_x2, int B, int H, int W, int C,
                         int sr, int CP, float* d_out, void* stream);
/* SpatialNet.CCL (spatial_network.py:369): f1,f2 [B,H,W,C] -> flow [B,H,W,4] = (flow_w, flow_h, 0, 0) */
int ss2_ccl_nhwc(ss2_ctx* ctx, const float* d_f1, const float* d_f2, int B, int H, int W, int C,
                 float* d_flow, void* stream);

/* ---- network stem (test / reuse entry) ------------------------------------------------------ */
/* `feature_extractor_stage1[0..3]` of SpatialNet / TemporalNet (spatial_network.py:123-131,
 * temporal_network.py:65-73) with the BatchNorm already folded by the caller: Conv2d(3,64,7,stride 2,pad 3) + bias +
 * ReLU + MaxPool2d(3, stride 2, pad 1).  d_x_nchw [B,3,H,W] -> d_out [B,Hp,Wp,64] (NHWC), Hp = ((H-1)/2)/2 + 1 etc.
 * h_weight HOST [64,3,7,7], h_bias HOST [64] or NULL.  variant 2 = the production kernel (direct tcgen05 convolution
 * with the pool fused, needs W = 480 and H % 4 == 0), 1 = implicit-GEMM tensor-core stem + pooling kernel,
 * 0 = exact-fp32 SIMT stem + pooling kernel.  Synchronous (packs the filter on every call). */
int ss2_stem_pool(ss2_ctx* ctx, const float* d_x_nchw, int B, int H, int W, const float* h_weight, const float* h_bias,
                  int variant, float* d_out, void* stream);

/* ---- convolution primitive --------------------------------------------------------------- */
/* nn.Conv2d / nn.Conv3d (bias optional, + optional residual add and ReLU) on NHWC / NDHWC fp32
 * activations - the layer type all three networks are built from (spatial_network.py:147-259,
 * temporal_network.py:65-104, smooth_network.py:124-131).  d_in [B,D,H,W,Cin] (Cin % 4 == 0),
 * h_weight HOST [Cout,Cin,(KD,)KH,KW] (reference layout), h_bias HOST [Cout] or NULL,
 * d_residual [B,Do,Ho,Wo,Cout] or NULL -> d_out [B,Do,Ho,Wo,Cout].  use_tc != 0 runs the tcgen05
 * tensor-core kernel (needs Cin % 32 == 0 and Cout % 4 == 0), 0 the exact-fp32 SIMT kernel.
 * Synchronous (packs the filter on every call): a test / reuse entry, not the hot path. */
int ss2_conv_nhwc(ss2_ctx* ctx, const float* d_in, int B, int D, int H, int W, int Cin, const float* h_weight,
                  const int64_t* wshape, int wndim, const float* h_bias, int stride, int pad, int pad_d, int relu,
                  const float* d_residual, int use_tc, float* d_out, void* stream);

/* ---- networks ------------------------------------------------------------------------- */
/* SpatialNet.forward, spatial_network.py:276: img1,img2 [bs,3,360,480] ->
 * offset_1 [bs,8], offset_2_ref [bs,126], offset_2_tgt [bs,126] */
int ss2_spatial_forward(ss2_ctx* ctx, const float* d_img1, const float* d_img2, int bs,
                        float* d_offset1, float* d_offset2_ref, float* d_offset2_tgt, void* stream);
/* build_SpatialNet, spatial_network.py:63: -> motion1, motion2 [bs,7,9,2] */
int ss2_build_spatial(ss2_ctx* ctx, const float* d_img1, const float* d_img2, int bs,
                      float* d_motion1, float* d_motion2, void* stream);
/* the post-network tail of build_SpatialNet (spatial_network.py:68-115) on its own */
int ss2_spatial_tail(ss2_ctx* ctx, const float* d_offset1, const float* d_offset2_ref,
                     const float* d_offset2_tgt, int bs, int img_h, int img_w, float* d_motion1,
                     float* d_motion2, void* stream);
/* build_TemporalNet, temporal_network.py:23: frames [n,3,360,480] (consecutive frames of one
 * view) -> motions [n,7,9,2]; motions[0] = 0, motions[k] = motion of frame k w.r.t. k-1 */
int ss2_build_temporal(ss2_ctx* ctx, const float* d_frames, int n, float* d_motions, void* stream);
/* Both views of a stream in one batch per chunk (the whole-stream calls use this): same results, bit for bit, as two
 * ss2_build_temporal calls; frames_a/b [n,3,360,480] -> motions_a/b [n,7,9,2]. */
int ss2_build_temporal_pair(ss2_ctx* ctx, const float* d_frames_a, const float* d_frames_b, int n, float* d_motions_a,
                            float* d_motions_b, void* stream);
/* build_SpatialNet over n frame pairs AND build_TemporalNet over both views in one call (the two networks are independent:
 * TemporalNet runs on a stream of its own next to SpatialNet, fork / join inside the library; bit-identical to the two
 * separate calls).  d_lr1, d_lr2 [halo+n,3,360,480]: the first `halo` frames (0, or 1 = the frame before a temporal shard)
 * feed TemporalNet only.  d_sm1, d_sm2 [n,7,9,2] = SpatialNet's motion1 / motion2; d_tm1, d_tm2 [halo+n,7,9,2] = TemporalNet's
 * motion lists of the two views (row 0 zero).  SS2_NET_OVERLAP=0 runs one network after the other. */
int ss2_build_spatial_temporal(ss2_ctx* ctx, const float* d_lr1, const float* d_lr2, int n, int halo, float* d_sm1,
                               float* d_sm2, float* d_tm1, float* d_tm2, void* stream);
/* tsmotion preparation, test_online_tra.py:309-347, one view: smotion,tmotion [n,7,9,2] ->
 * smesh, tsmotion [n,7,9,2].  `first_is_stream_start` != 0 makes tsmotion[0] = 0 (k == 0
 * branch); otherwise smotion_prev [7,9,2] (frame before the chunk) must be given. */
int ss2_tsmotion(ss2_ctx* ctx, const float* d_smotion, const float* d_tmotion, int n,
                 int first_is_stream_start, const float* d_smotion_prev, float* d_smesh,
                 float* d_tsmotion, void* stream);
/* build_SmoothNet over `nwin` consecutive sliding windows, smooth_network.py:23 +
 * test_online_tra.py:359-392.  smesh*, tsmotion* [nwin+6,7,9,2] (frame-major); window w uses
 * frames w..w+6; zero_first != 0 zeroes the tsmotion of each window's first frame (what the
 * reference's driver does before the call, test_online_tra.py:361-365).  Outputs, each [nwin,7,7,9,2]:
 * ori_path, smooth_path, ori_mesh, smooth_mesh for view 1 then view 2 (any may be NULL). */
int ss2_build_smooth(ss2_ctx* ctx, const float* d_tsmotion1, const float* d_tsmotion2,
                     const float* d_smesh1, const float* d_smesh2, int nwin, int zero_first,
                     float* d_ori_path1, float* d_smooth_path1, float* d_ori_mesh1, float* d_smooth_mesh1,
                     float* d_ori_path2, float* d_smooth_path2, float* d_ori_mesh2, float* d_smooth_mesh2,
                     void* stream);

/* ---- canvas (get_stable_sqe, test_online_tra.py:103-120) -------------------------------- */
/* smooth meshes [n,7,9,2] @480x360 for both views -> d_minmax[4] = {xmin, xmax, ymin, ymax} of
 * the hr-rescaled meshes (x*img_w/480, y*img_h/360).  Multi-GPU: all-reduce these 4 floats
 * (min on 0,2; max on 1,3) before ss2_stable_frames. */
int ss2_canvas_minmax(ss2_ctx* ctx, const float* d_mesh1, const float* d_mesh2, int n, int img_h,
                      int img_w, float* d_minmax, void* stream);
/* The whole get_stable_sqe loop (AVERAGE fusion) for n frames, given the global canvas:
 * hr1,hr2 [n,3,H,W] fp32 0..255; smooth meshes [n,7,9,2] @480x360; minmax[4] on the HOST.
 * out [n,3,Ho,Wo] with Ho=(int)(ymax-ymin), Wo=(int)(xmax-xmin) (fp32 subtraction, then
 * truncation, as test_online_tra.py:119-120,140). */
int ss2_stable_frames(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1,
                      const float* d_mesh2, int n, int H, int W, const float* h_minmax, int mode,
                      int tps, float* d_out, void* stream);
/* The same with the driver's uint8 back end (frame.astype(uint8), test_online_tra.py:152,414) fused into the store of
 * the resampler: out [n,Ho,Wo,3] uint8 (HWC like the reference's numpy frames), bit-identical to ss2_stable_frames +
 * ss2_frames_to_u8; the fp32 canvas is never written.  Lattice resampler only: SS2_ERR_UNSUPPORTED for tps ==
 * SS2_TPS_EXACT and for canvases too small for the lattice (use the two calls there). */
int ss2_stable_frames_u8(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1,
                         const float* d_mesh2, int n, int H, int W, const float* h_minmax, int mode,
                         int tps, uint8_t* d_out, void* stream);
/* canvas size helper (host arithmetic identical to the kernels') */
int ss2_canvas_size(const float* h_minmax, int* out_h, int* out_w);

/* ---- three views (test_online_tra_threeview.py:345-505, AVERAGE fusion) ------------------- */
/* Middle-plane alignment of two stitched pairs that share their middle view (:345-455).
 * d_w12m1, d_w12m2: smooth meshes [n,7,9,2] @480x360 of pair (1,2); d_w23m1, d_w23m2: of pair (2,3);
 * w12m2 and w23m1 are the two instances of the shared view.  Outputs [n,7,9,2] in provisional-canvas
 * pixels: d_mesh1 (view 1 moved onto the middle plane), d_middle, d_mesh3; d_canvas[4] (device) =
 * {width_min, height_min, out_width, out_height} of the NEW canvas over the three output meshes. */
int ss2_three_view_meshes(ss2_ctx* ctx, const float* d_w12m1, const float* d_w12m2, const float* d_w23m1,
                          const float* d_w23m2, int n, int img_h, int img_w, float* d_mesh1, float* d_middle,
                          float* d_mesh3, float* d_canvas, void* stream);
/* The three-image warp + fusion loop (:461-490) for n frames: images [n,3,H,W] fp32 0..255 per view,
 * the three meshes of ss2_three_view_meshes, h_canvas[4] on the HOST; out [n,3,Ho,Wo] with
 * Ho = (int)out_height, Wo = (int)out_width; fuse(1,2) then fuse(12,3) in the reference's operation order. */
int ss2_three_view_frames(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_img3,
                          const float* d_mesh1, const float* d_middle, const float* d_mesh3, int n, int H, int W,
                          const float* h_canvas, int mode, int tps, float* d_out, void* stream);

/* ---- N views (BASELINE.json config 5; generalises the three-view glue, identical to it for N = 3) ----------- */
/* Chain of stitched pairs (1,2), (2,3), .., (N-1,N): h_pair_meshes is a HOST array of 2*(N-1) DEVICE pointers
 * (meshA, meshB of pair 1, meshA, meshB of pair 2, ..), each [n,7,9,2] @480x360; meshB of pair p and meshA of pair
 * p+1 are two instances of one physical view.  Step 1 (align): pairs chained by the per-frame mean offset of their
 * shared views (:354-360), middle plane of every shared view (:363).  Outputs: d_shifted [2(N-1)][n,7,9,2] and
 * d_mids [N-2][n,7,9,2] in hr pixels, d_minmax1[4] = (xmin,xmax,ymin,ymax) of the shifted meshes = the provisional
 * canvas (:366-399).  A temporally sharded stream all-reduces d_minmax1 (min on 0,2; max on 1,3) before step 2. */
int ss2_nview_align(ss2_ctx* ctx, const float* const* h_pair_meshes, int nviews, int n, int img_h, int img_w,
                    float* d_shifted, float* d_mids, float* d_minmax1, void* stream);
/* Step 2 (remap): the two outer views follow their pair's instance of the neighbouring shared view through the TPS
 * onto its middle plane (:411-427).  h_minmax1: the (global) provisional canvas on the HOST.  d_meshes [N][n,7,9,2]:
 * the final meshes in provisional-canvas pixels; d_minmax2[4] = (xmin,xmax,ymin,ymax) over them = the new canvas
 * (:433-455), to be all-reduced like d_minmax1. */
int ss2_nview_remap(ss2_ctx* ctx, int nviews, int n, const float* d_shifted, const float* d_mids, const float* h_minmax1,
                    float* d_meshes, float* d_minmax2, void* stream);
/* Step 3: the N-image warp + sequential AVERAGE fusion fuse(..fuse(fuse(1,2),3)..,N) (:461-490) for n frames in ONE
 * fused pass over the canvas (2 <= N <= 4): h_imgs HOST array of N DEVICE pointers [n,3,H,W]; out [n,3,Ho,Wo] with
 * Ho = (int)(ymax-ymin), Wo = (int)(xmax-xmin) of h_minmax2. */
int ss2_nview_frames(ss2_ctx* ctx, const float* const* h_imgs, const float* d_meshes, int nviews, int n, int H, int W,
                     const float* h_minmax2, int mode, int tps, float* d_out, void* stream);

/* ---- whole stream, device resident ------------------------------------------------------- */
/* smooth meshes of a stream from per-window SmoothNet outputs (test_online_tra.py:378-392):
 * d_win_smooth [nwin,7,7,9,2].  with_head != 0: first window contributes its 7 meshes, every
 * later one its last (nwin+6 frames out); with_head == 0: only the last mesh of each window
 * (nwin frames out) - the form a non-first rank of a temporally sharded stream needs. */
int ss2_assemble_smooth(ss2_ctx* ctx, const float* d_win_smooth, int nwin, int with_head, float* d_out,
                        void* stream);
/* SpatialNet + TemporalNet x2 + tsmotion + SmoothNet windows for one stream of n >= 7 frames
 * (test_online_tra.py:284-392): lr1,lr2 [n,3,360,480] -> smooth meshes [n,7,9,2] per view.
 * Optional raw outputs (may be NULL): smotion, tmotion [n,7,9,2] per view. */
int ss2_stream_meshes(ss2_ctx* ctx, const float* d_lr1, const float* d_lr2, int n, float* d_smooth1,
                      float* d_smooth2, float* d_smotion1, float* d_smotion2, float* d_tmotion1,
                      float* d_tmotion2, void* stream);

/* ---- whole stream through HOST buffers (the e2e call) ------------------------------------ */
/* One video chunk, everything the reference's test() does between frame loading and video
 * writing (test_online_tra.py:284-399, AVERAGE fusion): H2D of the inputs, SpatialNet,
 * TemporalNet x2, tsmotion, SmoothNet windows, canvas, resample+blend, D2H of the frames.
 *   h_lr1,h_lr2 [n,3,360,480] in [-1,1];  h_hr1,h_hr2 [n,3,H,W] 0..255 (pinned memory advised)
 *   h_out: capacity `out_capacity` floats; receives n frames [3,Ho,Wo]; *out_h,*out_w set.
 *   h_smooth_mesh1/2 (optional, [n,7,9,2]) receive the smoothed meshes.
 * Synchronous (returns after the D2H finished). */
int ss2_stitch_stream_host(ss2_ctx* ctx, const float* h_lr1, const float* h_lr2, const float* h_hr1,
                           const float* h_hr2, int n, int H, int W, int mode, int tps, float* h_out,
                           int64_t out_capacity, int* out_h, int* out_w, float* h_smooth_mesh1,
                           float* h_smooth_mesh2);

/* Pipelined form: _async returns once the resample+blend launches and the D2H copies of this
 * chunk are ENQUEUED (it still blocks for the networks, because the canvas size is data
 * dependent); ss2_stitch_stream_host_wait(slot) blocks until h_out is complete.  Alternating two of
 * the slots (0, 1, 2) a caller overlaps the D2H of chunk k with the H2D + networks of chunk k+1.
 * Host buffers of a slot must stay untouched until its wait returns. */
int ss2_stitch_stream_host_async(ss2_ctx* ctx, int slot, const float* h_lr1, const float* h_lr2, const float* h_hr1,
                                 const float* h_hr2, int n, int H, int W, int mode, int tps, float* h_out,
                                 int64_t out_capacity, int* out_h, int* out_w, float* h_smooth_mesh1,
                                 float* h_smooth_mesh2);
int ss2_stitch_stream_host_wait(ss2_ctx* ctx, int slot);
/* Starts the host -> device copies of the chunk that the NEXT ss2_stitch_stream_host_async on `slot` will process
 * (same pointers and sizes) and returns at once: issued before the _async of the chunk in flight, the upload runs
 * underneath that chunk's networks.  The host buffers must stay valid and unchanged until that _async returns. */
int ss2_stitch_stream_host_prefetch(ss2_ctx* ctx, int slot, const float* h_lr1, const float* h_lr2,
                                    const float* h_hr1, const float* h_hr2, int n, int H, int W);

/* ---- LINEAR fusion (test_online_tra.py:34-58,143-150) ------------------- */
/* linear_blender: ref,tgt [n,3,Ho,Wo] warped images, ref_m,tgt_m [n,1,Ho,Wo] warped masks -> out [n,3,Ho,Wo]
 * (and/or mask1 [n,1,Ho,Wo], the reference's mask=True result).  img_stride / mask_stride: elements between
 * consecutive frames of the image / mask arguments (3*Ho*Wo and Ho*Wo for contiguous tensors).  A mask is the set
 * of pixels with value > 0.5: residue-free semantics (the reference's torch.nonzero also counts the rounding
 * residues of out-of-image samples; DESIGN.md has the measured difference).  Needs Ho, Wo > 10 (reflect padding). */
int ss2_linear_blend(ss2_ctx* ctx, const float* d_ref, const float* d_tgt, int64_t img_stride, const float* d_ref_m,
                     const float* d_tgt_m, int64_t mask_stride, int n, int Ho, int Wo, float* d_out, float* d_mask1,
                     void* stream);
/* The get_stable_sqe loop with fusion_mode == 'LINEAR' (same arguments as ss2_stable_frames): a ones channel rides
 * through the resampler as the mask (:144-147), then linear_blender per frame. */
int ss2_stable_frames_linear(ss2_ctx* ctx, const float* d_hr1, const float* d_hr2, const float* d_mesh1,
                             const float* d_mesh2, int n, int H, int W, const float* h_minmax, int mode, int tps,
                             float* d_out, void* stream);

/* three-image warp + LINEAR fusion (test_online_tra_threeview.py:492-503): arguments of ss2_three_view_frames;
 * fuse(1,2) with linear_blender, mask12 = mask1 + mask2 - mask1*mask2, then fuse(12,3). */
int ss2_three_view_frames_linear(ss2_ctx* ctx, const float* d_img1, const float* d_img2, const float* d_img3,
                                 const float* d_mesh1, const float* d_middle, const float* d_mesh3, int n, int H, int W,
                                 const float* h_canvas, int mode, int tps, float* d_out, void* stream);

/* ---- metric path (test_metric_ssd.py) ---------------------------------------- */
/* Whole-stream original / smoothed path of one view from the per-window SmoothNet outputs (:417-436):
 * d_win_ori_path, d_win_smooth_path [nwin,7,7,9,2] (ss2_build_smooth) -> [nwin+6,7,9,2] each. */
int ss2_assemble_paths(ss2_ctx* ctx, const float* d_win_ori_path, const float* d_win_smooth_path, int nwin,
                       float* d_ori_path, float* d_smooth_path, void* stream);
/* d_out[0] = stability score of d_path [n,7,9,2] (:455-466; n >= 7), d_out[1] = distortion score of d_mesh [n,7,9,2]
 * (max over frames of inter_grid_loss + intra_grid_loss, :37-88,470-479).  Either input may be NULL. */
int ss2_metric_scores(ss2_ctx* ctx, const float* d_path, const float* d_mesh, int n, float* d_out, void* stream);
/* PSNR / SSIM of two warped views inside their overlap (:513-518): d_warp1, d_warp2 [n,6,H,W] (image planes 0..2,
 * warped ones planes 3..5, the C = 6 output of ss2_tps_warp) -> d_psnr[n], d_ssim[n]; skimage 0.15 compare_psnr /
 * compare_ssim(multichannel=True) semantics (fp64 accumulation). */
int ss2_metric_psnr_ssim(ss2_ctx* ctx, const float* d_warp1, const float* d_warp2, int n, int H, int W, float* d_psnr,
                         float* d_ssim, void* stream);

/* ---- uint8 host edges (test_online_tra.py:252-264,152,414) ------------- */
/* What the reference's driver does to a decoded frame before the networks, on the device:
 * d_u8 [n,H,W,3] uint8 (cv2.imread layout, BGR) -> d_hr [n,3,H,W] fp32 0..255 (astype(float32) + transpose) and
 * d_lr [n,3,360,480] fp32 = cv2.resize(img,(480,360)) (INTER_LINEAR, OpenCV's 11-bit fixed-point scheme, bit-exact)
 * /127.5 - 1.  Either output may be NULL. */
int ss2_load_frames_u8(ss2_ctx* ctx, const uint8_t* d_u8, int n, int H, int W, float* d_hr, float* d_lr, void* stream);
/* fused frames [n,3,Ho,Wo] fp32 -> [n,Ho,Wo,3] uint8: transpose(1,2,0) + numpy astype(uint8) (truncation toward zero,
 * wrap modulo 256), what the driver hands to cv2.VideoWriter (:414). */
int ss2_frames_to_u8(ss2_ctx* ctx, const float* d_frames, int n, int Ho, int Wo, uint8_t* d_out, void* stream);
/* The e2e calls with the uint8 interface: h_bgr1,h_bgr2 [n,H,W,3] uint8 HOST (pinned advised) -> h_out receives n
 * frames [Ho,Wo,3] uint8 (capacity in bytes).  Same slots, pipelining rules and wait call
 * (ss2_stitch_stream_host_wait) as the fp32 entry points; 4.4x fewer bytes over PCIe than the fp32 interface. */
int ss2_stitch_stream_host_u8(ss2_ctx* ctx, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n, int H, int W, int mode,
                              int tps, uint8_t* h_out, int64_t out_capacity, int* out_h, int* out_w,
                              float* h_smooth_mesh1, float* h_smooth_mesh2);
int ss2_stitch_stream_host_u8_async(ss2_ctx* ctx, int slot, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n, int H,
                                    int W, int mode, int tps, uint8_t* h_out, int64_t out_capacity, int* out_h,
                                    int* out_w, float* h_smooth_mesh1, float* h_smooth_mesh2);
int ss2_stitch_stream_host_u8_prefetch(ss2_ctx* ctx, int slot, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n,
                                       int H, int W);

/* Two-phase form of the _async calls, for full overlap: _submit only ENQUEUES a chunk's uploads (unless
 * prefetched), front end and networks and returns at once; _finish waits for that chunk's canvas, enqueues its
 * resample + blend and frame downloads and returns the output shape (arguments as in the _async call of the same
 * interface; h_out uint8 after a _u8_submit, float otherwise).  Submitting chunk k+1 before finishing chunk k keeps
 * the GPU busy while the host waits for chunk k's data-dependent canvas.  A _submit blocks until the slot's previous
 * downloads are complete, so rotate all THREE slots in this form (with two, the submit of chunk k+1 would wait for
 * the downloads of chunk k-1 and the networks of chunk k+1 would start late). */
int ss2_stitch_stream_host_submit(ss2_ctx* ctx, int slot, const float* h_lr1, const float* h_lr2, const float* h_hr1,
                                  const float* h_hr2, int n, int H, int W);
int ss2_stitch_stream_host_u8_submit(ss2_ctx* ctx, int slot, const uint8_t* h_bgr1, const uint8_t* h_bgr2, int n, int H,
                                     int W);
int ss2_stitch_stream_host_finish(ss2_ctx* ctx, int slot, int mode, int tps, void* h_out, int64_t out_capacity,
                                  int* out_h, int* out_w, float* h_smooth_mesh1, float* h_smooth_mesh2);

#ifdef __cplusplus
}
#endif
#endif /* SS2_H */
