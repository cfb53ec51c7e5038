// Host edges of the path on the device: what the reference's driver does to a decoded
// video frame before the networks and to a fused frame before the video writer.
//
// Reference behaviour restated (paths under Full_model_inference/Codes/):
//   test_online_tra.py:252-264  img = cv2.imread(..)                       uint8 [H,W,3] (BGR)
//                               hr  = img.astype(float32).transpose(2,0,1)  fp32 [3,H,W] 0..255
//                               lr  = cv2.resize(img, (480, 360))           uint8, INTER_LINEAR
//                               lr  = lr.astype(float32).transpose(2,0,1) / 127.5 - 1.0
//   test_online_tra.py:152,414  fused.cpu().numpy().transpose(1,2,0) ... .astype(np.uint8)
// cv2.resize lives in OpenCV (third party, 4.x; not vendored by the reference): INTER_LINEAR on uint8 is a
// fixed-point scheme (11-bit coefficients; horizontal pass in int, vertical pass
// (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2).  oracle/host_edges.py restates it and is pinned
// bit-exactly against cv2 for the reference's shapes; these kernels are pinned bit-exactly against that oracle
// (tests/test_gpu_parity.py::test_host_edges_u8_bit_exact).
#include <math.h>

#include "common.cuh"

#define LR_H 360
#define LR_W 480
#define COEF_BITS 11

// uint8 [n,H,W,3] -> fp32 [n,3,H,W]
__global__ void u8_to_planar_kernel(const unsigned char* __restrict__ in, size_t npix_frame, size_t total,
                                    float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t f = i / npix_frame, p = i - f * npix_frame;
    const unsigned char* s = in + i * 3;
    float* o = out + f * 3 * npix_frame + p;
    o[0] = (float)s[0];
    o[npix_frame] = (float)s[1];
    o[2 * npix_frame] = (float)s[2];
  }
}

// per destination index: first source index (clamped) and the two 11-bit weights, x table then y table
struct ResizeTab { int idx; short w0, w1; };

// uint8 [n,H,W,3] -> fp32 [n,3,360,480] = cv2.resize(INTER_LINEAR) then /127.5 - 1
__global__ void resize_u8_lr_kernel(const unsigned char* __restrict__ in, int H, int W, const ResizeTab* __restrict__ tab,
                                    float* __restrict__ out) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, f = blockIdx.z;
  if (dx >= LR_W) return;
  const ResizeTab tx = tab[dx], ty = tab[LR_W + dy];
  const int x0 = tx.idx, x1 = min(x0 + 1, W - 1), y0 = ty.idx, y1 = min(y0 + 1, H - 1);
  const unsigned char* r0 = in + ((size_t)f * H + y0) * W * 3;
  const unsigned char* r1 = in + ((size_t)f * H + y1) * W * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = (int)r0[x0 * 3 + c] * tx.w0 + (int)r0[x1 * 3 + c] * tx.w1;   // horizontal pass
    const int h1 = (int)r1[x0 * 3 + c] * tx.w0 + (int)r1[x1 * 3 + c] * tx.w1;
    const int v = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;   // vertical pass
    const float u = (float)(v < 0 ? 0 : (v > 255 ? 255 : v));                     // saturate_cast<uchar>
    out[(((size_t)f * 3 + c) * LR_H + dy) * LR_W + dx] = __fsub_rn(__fdiv_rn(u, 127.5f), 1.0f);
  }
}

// fp32 [n,3,Ho,Wo] -> uint8 [n,Ho,Wo,3]: numpy astype(uint8) = C conversion through a signed integer (truncation toward
// zero, wrap modulo 256)
__global__ void planar_to_u8_kernel(const float* __restrict__ in, size_t npix_frame, size_t total,
                                    unsigned char* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t f = i / npix_frame, p = i - f * npix_frame;
    const float* s = in + f * 3 * npix_frame + p;
    unsigned char* o = out + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldcs(s + c * npix_frame);
      const long long t = (v == v && fabsf(v) < 9.0e18f) ? (long long)truncf(v) : 0ll;
      o[c] = (unsigned char)(t & 0xFF);
    }
  }
}

// OpenCV's coefficient computation for INTER_LINEAR (resize.cpp): fx = (float)((d + 0.5) * scale - 0.5); sx = floor(fx);
// fx -= sx; clamped at both ends; weights cvRound((1 - fx) * 2048), cvRound(fx * 2048) (round half to even)
static void linear_coeffs(int ssize, int dsize, ResizeTab* t) {
  const double scale = (double)ssize / dsize;
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    t[d].idx = s;
    t[d].w0 = (short)lrintf((1.0f - f) * (float)(1 << COEF_BITS));
    t[d].w1 = (short)lrintf(f * (float)(1 << COEF_BITS));
  }
}

static int resize_table(ss2_ctx* ctx, int H, int W, const ResizeTab** d_tab) {
  char name[48];
  snprintf(name, sizeof(name), "resize_tab.%dx%d", H, W);
  auto it = ctx->stream_bufs.find(name);
  if (it != ctx->stream_bufs.end()) { *d_tab = (const ResizeTab*)it->second.first; return SS2_OK; }
  std::vector<ResizeTab> h(LR_W + LR_H);
  linear_coeffs(W, LR_W, h.data());
  linear_coeffs(H, LR_H, h.data() + LR_W);
  void* p = nullptr;
  SS2_CUDA(ctx, cudaMalloc(&p, h.size() * sizeof(ResizeTab)));
  SS2_CUDA(ctx, cudaMemcpy(p, h.data(), h.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
  ctx->stream_bufs[name] = std::make_pair(p, h.size() * sizeof(ResizeTab));
  *d_tab = (const ResizeTab*)p;
  return SS2_OK;
}

int load_frames_u8_launch(ss2_ctx* ctx, const unsigned char* d_u8, int n, int H, int W, float* d_hr, float* d_lr,
                          cudaStream_t st) {
  if (n <= 0) return SS2_OK;
  const size_t npix = (size_t)H * W, total = npix * n;
  if (d_hr) {
    const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    u8_to_planar_kernel<<<grid, 256, 0, st>>>(d_u8, npix, total, d_hr);
    SS2_LAUNCH_CHECK(ctx);
  }
  if (d_lr) {
    const ResizeTab* tab;
    SS2_TRY(resize_table(ctx, H, W, &tab));
    resize_u8_lr_kernel<<<dim3(cdiv(LR_W, 128), LR_H, n), 128, 0, st>>>(d_u8, H, W, tab, d_lr);
    SS2_LAUNCH_CHECK(ctx);
  }
  return SS2_OK;
}

int frames_to_u8_launch(ss2_ctx* ctx, const float* d_frames, int n, int Ho, int Wo, unsigned char* d_out, cudaStream_t st) {
  if (n <= 0 || Ho <= 0 || Wo <= 0) return SS2_OK;
  const size_t npix = (size_t)Ho * Wo, total = npix * n;
  const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
  planar_to_u8_kernel<<<grid, 256, 0, st>>>(d_frames, npix, total, d_out);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

extern "C" int ss2_load_frames_u8(ss2_ctx* ctx, const uint8_t* d_u8, int n, int H, int W, float* d_hr, float* d_lr,
                                  void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || H < 2 || W < 2 || (n > 0 && (!d_u8 || (!d_hr && !d_lr))))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_load_frames_u8: bad arguments");
  return load_frames_u8_launch(ctx, d_u8, n, H, W, d_hr, d_lr, (cudaStream_t)stream);
}

extern "C" int ss2_frames_to_u8(ss2_ctx* ctx, const float* d_frames, int n, int Ho, int Wo, uint8_t* d_out, void* stream) {
  if (!ctx) return SS2_ERR_INVALID;
  if (n < 0 || Ho < 0 || Wo < 0 || (n > 0 && Ho > 0 && Wo > 0 && (!d_frames || !d_out)))
    return ss2_fail(ctx, SS2_ERR_INVALID, "ss2_frames_to_u8: bad arguments");
  return frames_to_u8_launch(ctx, d_frames, n, Ho, Wo, d_out, (cudaStream_t)stream);
}
