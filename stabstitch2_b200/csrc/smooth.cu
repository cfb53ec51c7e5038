// SmoothNet glue kernels: path accumulation + Linear(2->32) embeddings (the A operand of the
// first Conv3d) and the Linear(128->4) decoding fused with the mesh/path updates.  The three
// Conv3d(128,128,(5,3,3)) layers run on the implicit-GEMM conv kernels (conv.cu / conv_tc.cu).
//
// Reference behaviour restated (paths under Full_model_inference/Codes/):
//   smooth_network.py:64-101   SmoothNet.forward (prefix-sum paths, [bs,T,7,9,2] layout)
//   smooth_network.py:139-157  MotionPrediction.forward (embedding1/3, concat order, decoding)
//   smooth_network.py:23-41    build_SmoothNet (smooth_path = path + d, smooth_mesh = mesh - d)
//   test_online_tra.py:359-366 window slicing, tsmotion of the window's first frame zeroed
#include "common.cuh"

#define SW_T SS2_WINDOW

// grid (nwin, T); 128 threads = hidden channels [emb1(mesh1) | emb3(path1) | emb1(mesh2) | emb3(path2)]
__global__ void __launch_bounds__(128)
smooth_embed_kernel(const float* __restrict__ e1w, const float* __restrict__ e1b, const float* __restrict__ e3w,
                    const float* __restrict__ e3b, const float* __restrict__ ts1, const float* __restrict__ ts2,
                    const float* __restrict__ sm1, const float* __restrict__ sm2, int zero_first,
                    ActRef hidden, float* __restrict__ path1, float* __restrict__ path2) {
  __shared__ float in[4][SS2_NPT][2];  // mesh1, path1, mesh2, path2 for this (window, t)
  const int w = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
  for (int e = tid; e < 2 * SS2_NPT * 2; e += 128) {
    const int v = e / (SS2_NPT * 2), r = e % (SS2_NPT * 2);
    const float* ts = v == 0 ? ts1 : ts2;
    const float* sm = v == 0 ? sm1 : sm2;
    // path: element 0 is tsmotion*0, then running fp32 sum in frame order
    float acc = ts[(size_t)w * SS2_NPT * 2 + r];
    if (zero_first) acc *= 0.0f;
    for (int q = 1; q <= t; ++q) acc = __fadd_rn(acc, ts[(size_t)(w + q) * SS2_NPT * 2 + r]);
    in[2 * v + 1][r / 2][r % 2] = acc;
    in[2 * v][r / 2][r % 2] = sm[(size_t)(w + t) * SS2_NPT * 2 + r];
    (v == 0 ? path1 : path2)[((size_t)w * SW_T + t) * SS2_NPT * 2 + r] = acc;
  }
  __syncthreads();
  const int grp = tid / 32, j = tid % 32;
  const float* W = (grp & 1) ? e3w : e1w;
  const float* Bv = (grp & 1) ? e3b : e1b;
  const float w0 = W[2 * j], w1 = W[2 * j + 1], bb = Bv[j];
  const size_t h = ((size_t)w * SW_T + t) * SS2_NPT * 128 + tid;
  for (int p = 0; p < SS2_NPT; ++p) {
    const float v = __fadd_rn(fmaf(in[grp][p][1], w1, in[grp][p][0] * w0), bb);
    store_split1(hidden, h + (size_t)p * 128, fmaxf(v, 0.f));
  }
}

int smooth_embed_launch(ss2_ctx* ctx, const SmoothWeights& sw, const float* ts1, const float* ts2,
                        const float* sm1, const float* sm2, int nwin, int zero_first, const ActRef& hidden, float* d_path1,
                        float* d_path2, cudaStream_t st) {
  if (nwin <= 0) return SS2_OK;
  smooth_embed_kernel<<<dim3(nwin, SW_T), 128, 0, st>>>(sw.emb1_w, sw.emb1_b, sw.emb3_w, sw.emb3_b, ts1, ts2, sm1,
                                                       sm2, zero_first, hidden, d_path1, d_path2);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}

// one warp per (window, t, vertex): 128-channel dot products with the 4 decoding rows
__global__ void __launch_bounds__(256)
smooth_decode_kernel(const float* __restrict__ dw, const float* __restrict__ db, const float* __restrict__ hidden,
                     const float* __restrict__ sm1, const float* __restrict__ sm2, const float* __restrict__ path1,
                     const float* __restrict__ path2, int total, float* op1, float* sp1, float* om1, float* smm1,
                     float* op2, float* sp2, float* om2, float* smm2) {
  const int pos = blockIdx.x * 8 + threadIdx.x / 32;  // (w*T + t)*63 + p
  const int lane = threadIdx.x & 31;
  if (pos >= total) return;
  const float4 h = __ldg(reinterpret_cast<const float4*>(hidden + (size_t)pos * 128) + lane);
  float d[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(dw + o * 128) + lane);
    float s = fmaf(h.w, wv.w, fmaf(h.z, wv.z, fmaf(h.y, wv.y, h.x * wv.x)));
#pragma unroll
    for (int q = 16; q > 0; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
    d[o] = s + db[o];
  }
  if (lane < 4) {
    const int v = lane / 2, c = lane % 2;
    const int p = pos % SS2_NPT, wt = pos / SS2_NPT, t = wt % SW_T, w = wt / SW_T;
    const float mesh = (v == 0 ? sm1 : sm2)[((size_t)(w + t) * SS2_NPT + p) * 2 + c];
    const float path = (v == 0 ? path1 : path2)[(size_t)pos * 2 + c];
    const float delta = d[lane];
    float* a_op = v == 0 ? op1 : op2;
    float* a_sp = v == 0 ? sp1 : sp2;
    float* a_om = v == 0 ? om1 : om2;
    float* a_sm = v == 0 ? smm1 : smm2;
    const size_t o = (size_t)pos * 2 + c;
    if (a_op) a_op[o] = path;
    if (a_sp) a_sp[o] = __fadd_rn(path, delta);
    if (a_om) a_om[o] = mesh;
    if (a_sm) a_sm[o] = __fsub_rn(mesh, delta);
  }
}

int smooth_decode_launch(ss2_ctx* ctx, const SmoothWeights& sw, const float* d_hidden, const float* sm1,
                         const float* sm2, const float* path1, const float* path2, int nwin, float* op1,
                         float* sp1, float* om1, float* smm1, float* op2, float* sp2, float* om2, float* smm2,
                         cudaStream_t st) {
  if (nwin <= 0) return SS2_OK;
  const int total = nwin * SW_T * SS2_NPT;
  smooth_decode_kernel<<<cdiv(total, 8), 256, 0, st>>>(sw.dec_w, sw.dec_b, d_hidden, sm1, sm2, path1, path2, total,
                                                      op1, sp1, om1, smm1, op2, sp2, om2, smm2);
  SS2_LAUNCH_CHECK(ctx);
  return SS2_OK;
}
