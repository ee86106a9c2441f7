// mapping.cu -- the GANformer mapping network (z -> ws), forward and backward wrt z, as one kernel each.
//
// Restates reference training/networks.py MappingNetwork.forward :894-942 for the GANformer-default configuration (k = 16 local + 1 global
// latents of 32 dims, resnet MLPs of 4 blocks (:154-221), latent-to-latent self-attention before every local block (TransformerLayer
// :748-822 with integration "add", no norm, no k-means, one head), positional maps, equalised learning rate lrmul = 0.01):
//   local  x0 = z[:16] * rsqrt(mean_{16x32} z^2 + 1e-8)
//          per block l: q = Wq x + (bq + Pq pos), k = Wk x + (bk + Pk pos), v = Wv x + bv, A = softmax(q k^T / sqrt(32) + maskbias),
//                       xs = x + Wm (A v) + bm ; h0 = lrelu(W0 xs + b0) sqrt2 ; x <- lrelu(W1 h0 + b1 + x)
//          out = lrelu(Wo x + bo) sqrt2
//   global the same four resnet blocks + output layer on z[16] * rsqrt(mean_{32} z^2 + 1e-8), without attention
//   ws[b, t, l, :] = out[t] for every synthesis layer l (num_broadcast, :932)
// The projection loop differentiates the loss wrt z, so this sits on the hot path twice per step; in PyTorch it is ~400 tiny kernels
// (1.4 ms of a 26 ms step at 8 images).  Here: one CTA per (image, local | global), thread = (token, channel), every 32x32 weight
// staged in shared memory (row stride 33: conflict-free in both the W x and the W^T dy orientation); the backward kernel first
// recomputes the forward into shared memory (70 KB), so nothing is saved between the two launches.  Latency-bound by design (0.6 MFLOP).
#include "common.cuh"

namespace mgf {
namespace {

constexpr int D = 32, T = 16, NB = 4;              // latent width, local latents, resnet blocks
constexpr int WSZ = D * D, WS = D + 1;             // weight size, padded shared-memory row stride
constexpr float SQRT2 = 1.41421356237f, SLOPE = 0.2f;
// packed parameter layout (floats); every W is [out][in] with its runtime gain folded in, biases with theirs
constexpr int L_WQ = 0, L_WK = WSZ, L_WV = 2 * WSZ, L_WM = 3 * WSZ, L_W0 = 4 * WSZ, L_W1 = 5 * WSZ;
constexpr int L_CQ = 6 * WSZ, L_CK = L_CQ + T * D, L_BV = L_CK + T * D, L_BM = L_BV + D, L_B0 = L_BM + D, L_B1 = L_B0 + D;
constexpr int L_SIZE = L_B1 + D;                                   // one local block
constexpr int LO_W = NB * L_SIZE, LO_B = LO_W + WSZ, LOCAL_SIZE = LO_B + D;
constexpr int G_W0 = 0, G_W1 = WSZ, G_B0 = 2 * WSZ, G_B1 = G_B0 + D, G_SIZE = G_B1 + D;
constexpr int GO_W = NB * G_SIZE, GO_B = GO_W + WSZ, GLOBAL_SIZE = GO_B + D;
constexpr int PARAM_FLOATS = LOCAL_SIZE + GLOBAL_SIZE;

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : v * SLOPE; }
__device__ __forceinline__ float dlrelu(float out) { return out > 0.f ? 1.f : SLOPE; }     // sign(out) == sign(pre-activation)

// stage n 32x32 weights (consecutive in `src` with stride `stride` floats) into sW[n][32][33]
__device__ __forceinline__ void stage_w(float* sW, const float* src, int n, int stride) {
  for (int i = threadIdx.x; i < n * WSZ; i += blockDim.x) {
    const int m = i / WSZ, r = i - m * WSZ;
    sW[m * D * WS + (r >> 5) * WS + (r & 31)] = __ldg(src + (long long)m * stride + r);
  }
}
__device__ __forceinline__ float fc_f(const float* sW, const float* xrow, int c) {        // sum_k W[c][k] x[k]
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < D; k++) a = fmaf(sW[c * WS + k], xrow[k], a);
  return a;
}
__device__ __forceinline__ float fc_b(const float* sW, const float* dyrow, int k) {       // sum_c W[c][k] dy[c]
  float a = 0.f;
#pragma unroll
  for (int c = 0; c < D; c++) a = fmaf(sW[c * WS + k], dyrow[c], a);
  return a;
}
__device__ __forceinline__ float block_sum(float v, float* red) {          // all threads get the sum; red: 32 floats
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += red[i];
  return s;
}

// per-block saved tensors of the local path (floats): Q, K, V [16][32], P [16][16], H0, XO [16][32]
constexpr int S_Q = 0, S_K = T * D, S_V = 2 * T * D, S_P = 3 * T * D, S_H0 = S_P + T * T, S_XO = S_H0 + T * D, S_SIZE = S_XO + T * D;

// Shared-memory plan of the local CTA (floats):
//   sW[3][32][33] | X [16][32] (running activation) | A, Bf, Cf [16][32] scratch | red[32] | mb[16] | saved[NB][S_SIZE] + OUT[16][32] (backward only)
constexpr int SM_W = 0, SM_X = 3 * D * WS, SM_A = SM_X + T * D, SM_B = SM_A + T * D, SM_C = SM_B + T * D, SM_RED = SM_C + T * D, SM_MB = SM_RED + 32;
constexpr int SM_SAVE = SM_MB + T, SM_FWD_FLOATS = SM_SAVE + S_SIZE, SM_BWD_FLOATS = SM_SAVE + NB * S_SIZE + T * D;

// forward of the local path; thread = (token t, channel c), 512 threads.  If `keep`, block l's tensors go to sv + l * S_SIZE, else every
// block reuses sv.  Leaves the output in `outp` (shared) and returns the normalisation factor r.
__device__ float local_forward(const float* P_, const float* z, const float* mbias, float* sm, float* sv, bool keep, float* outp) {
  const int t = threadIdx.x >> 5, c = threadIdx.x & 31;
  float* sW = sm + SM_W; float* X = sm + SM_X; float* A = sm + SM_A; float* Bf = sm + SM_B; float* red = sm + SM_RED; float* mb = sm + SM_MB;
  const float zv = z[t * D + c];
  if (threadIdx.x < T) mb[threadIdx.x] = mbias[threadIdx.x];
  const float r = rsqrtf(block_sum(zv * zv, red) / (float)(T * D) + 1e-8f);
  X[t * D + c] = zv * r;
  for (int l = 0; l < NB; l++) {
    const float* L = P_ + l * L_SIZE;
    float* S = keep ? sv + l * S_SIZE : sv;
    __syncthreads();
    stage_w(sW, L + L_WQ, 3, WSZ);
    __syncthreads();
    const float xin = X[t * D + c];
    S[S_Q + t * D + c] = fc_f(sW, X + t * D, c) + L[L_CQ + t * D + c];
    S[S_K + t * D + c] = fc_f(sW + D * WS, X + t * D, c) + L[L_CK + t * D + c];
    S[S_V + t * D + c] = fc_f(sW + 2 * D * WS, X + t * D, c) + L[L_BV + c];
    __syncthreads();
    stage_w(sW, L + L_WM, 3, WSZ);                 // Wm, W0, W1 for the rest of the block (the Q/K/V weights are done)
    if (threadIdx.x < T * T) {                     // scores + softmax: thread = (row tt, key u), 16-lane groups
      const int tt = threadIdx.x >> 4, u = threadIdx.x & 15;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < D; k++) s = fmaf(S[S_Q + tt * D + k], S[S_K + u * D + k], s);
      s = s * 0.17677669529663687f + mb[u];
      float mx = s;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = __expf(s - mx);
      float den = e;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
      S[S_P + tt * T + u] = e / den;
    }
    __syncthreads();
    {
      float a = 0.f;
#pragma unroll
      for (int u = 0; u < T; u++) a = fmaf(S[S_P + t * T + u], S[S_V + u * D + c], a);
      A[t * D + c] = a;                            // ctl = A v
    }
    __syncthreads();
    Bf[t * D + c] = xin + fc_f(sW, A + t * D, c) + L[L_BM + c];             // xs
    __syncthreads();
    S[S_H0 + t * D + c] = lrelu(fc_f(sW + D * WS, Bf + t * D, c) + L[L_B0 + c]) * SQRT2;
    __syncthreads();
    const float xo = lrelu(fc_f(sW + 2 * D * WS, S + S_H0 + t * D, c) + L[L_B1 + c] + xin);
    S[S_XO + t * D + c] = xo;
    X[t * D + c] = xo;                             // own element only: no hazard with other threads' reads of X (they read before the last sync)
  }
  __syncthreads();
  stage_w(sW, P_ + LO_W, 1, WSZ);
  __syncthreads();
  outp[t * D + c] = lrelu(fc_f(sW, X + t * D, c) + P_[LO_B + c]) * SQRT2;
  return r;
}

// global path: one warp (lane = channel).  saved: H0[NB][32], XO[NB][32] if sv != nullptr.  Weights are read straight from global memory
// (one warp, 9 matrices: L1 keeps the rows).
__device__ __forceinline__ float gfc_f(const float* W, const float* xs, int c) { float a = 0.f; for (int k = 0; k < D; k++) a = fmaf(__ldg(W + c * D + k), xs[k], a); return a; }
__device__ __forceinline__ float gfc_b(const float* W, const float* dys, int k) { float a = 0.f; for (int c = 0; c < D; c++) a = fmaf(__ldg(W + c * D + k), dys[c], a); return a; }

__device__ float global_forward(const float* G_, const float* zg, float* xs, float* hs, float* sv, float* outv) {
  const int c = threadIdx.x & 31;
  const float zv = zg[c];
  const float r = rsqrtf(warp_sum(zv * zv) / (float)D + 1e-8f);
  float x = zv * r;
  for (int l = 0; l < NB; l++) {
    const float* L = G_ + l * G_SIZE;
    xs[c] = x; __syncwarp();
    const float h0 = lrelu(gfc_f(L + G_W0, xs, c) + L[G_B0 + c]) * SQRT2;
    hs[c] = h0; __syncwarp();
    const float xo = lrelu(gfc_f(L + G_W1, hs, c) + L[G_B1 + c] + x);
    if (sv) { sv[l * 2 * D + c] = h0; sv[l * 2 * D + D + c] = xo; }
    x = xo; __syncwarp();
  }
  xs[c] = x; __syncwarp();
  *outv = lrelu(gfc_f(G_ + GO_W, xs, c) + G_[GO_B + c]) * SQRT2;
  __syncwarp();
  return r;
}

// grid (B, 2): y = 0 local path (512 threads), y = 1 global path (first warp)
__global__ void __launch_bounds__(512) mapping_fwd_kernel(const float* __restrict__ z, const float* __restrict__ params, const float* __restrict__ maskbias,
                                                          float* __restrict__ ws, int num_ws) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x;
  const float* zb = z + (long long)b * (T + 1) * D;
  float* wb = ws + (long long)b * (T + 1) * num_ws * D;
  if (blockIdx.y == 0) {
    float* outp = sm + SM_C;
    local_forward(params, zb, maskbias + b * T, sm, sm + SM_SAVE, false, outp);
    __syncthreads();
    const int t = threadIdx.x >> 5, c = threadIdx.x & 31;
    const float v = outp[t * D + c];
    for (int l = 0; l < num_ws; l++) wb[((long long)t * num_ws + l) * D + c] = v;
  } else if (threadIdx.x < 32) {
    float outv;
    global_forward(params + LOCAL_SIZE, zb + T * D, sm, sm + D, nullptr, &outv);
    for (int l = 0; l < num_ws; l++) wb[((long long)T * num_ws + l) * D + threadIdx.x] = outv;
  }
}

__global__ void __launch_bounds__(512) mapping_bwd_kernel(const float* __restrict__ z, const float* __restrict__ params, const float* __restrict__ maskbias,
                                                          const float* __restrict__ dws, float* __restrict__ dz, int num_ws) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x;
  const float* zb = z + (long long)b * (T + 1) * D;
  const float* gb = dws + (long long)b * (T + 1) * num_ws * D;
  float* dzb = dz + (long long)b * (T + 1) * D;
  if (blockIdx.y == 0) {
    const int t = threadIdx.x >> 5, c = threadIdx.x & 31;
    float* sW = sm + SM_W; float* A = sm + SM_A; float* Bf = sm + SM_B; float* Cf = sm + SM_C; float* red = sm + SM_RED;
    float* saved = sm + SM_SAVE; float* OUT = saved + NB * S_SIZE;
    const float r = local_forward(params, zb, maskbias + b * T, sm, saved, true, OUT);
    float g = 0.f;                                        // d(loss)/d(out[t][c]) = sum over the broadcast layers
    for (int l = 0; l < num_ws; l++) g += gb[((long long)t * num_ws + l) * D + c];
    __syncthreads();
    A[t * D + c] = g * SQRT2 * dlrelu(OUT[t * D + c]);   // sW still holds Wo
    __syncthreads();
    float dx = fc_b(sW, A + t * D, c);                    // gradient wrt the running activation x (output of block NB-1)
    for (int l = NB - 1; l >= 0; l--) {
      const float* L = params + l * L_SIZE;
      const float* S = saved + l * S_SIZE;
      __syncthreads();
      stage_w(sW, L + L_WM, 3, WSZ);                      // Wm, W0, W1
      const float dsum = dx * dlrelu(S[S_XO + t * D + c]);                 // through x_out = lrelu(h1 + x_in)
      float dxin = dsum;
      A[t * D + c] = dsum;
      __syncthreads();
      Bf[t * D + c] = fc_b(sW + 2 * D * WS, A + t * D, c) * SQRT2 * dlrelu(S[S_H0 + t * D + c]);     // d(pre-activation of fc0)
      __syncthreads();
      const float dxs = fc_b(sW + D * WS, Bf + t * D, c);                  // gradient wrt xs = x_in + Wm ctl + bm
      dxin += dxs;
      A[t * D + c] = dxs;
      __syncthreads();
      Cf[t * D + c] = fc_b(sW, A + t * D, c);                              // dctl
      __syncthreads();
      stage_w(sW, L + L_WQ, 3, WSZ);                      // Wq, Wk, Wv for the projections' backward
      if (threadIdx.x < T * T) {                          // dP, softmax backward -> dS (stored in Bf[0..255], pre-scaled by 1/sqrt(32))
        const int tt = threadIdx.x >> 4, u = threadIdx.x & 15;
        float dp = 0.f;
#pragma unroll
        for (int k = 0; k < D; k++) dp = fmaf(Cf[tt * D + k], S[S_V + u * D + k], dp);
        const float p = S[S_P + tt * T + u];
        float dot = p * dp;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        Bf[tt * T + u] = p * (dp - dot) * 0.17677669529663687f;
      }
      __syncthreads();
      float dq = 0.f, dk = 0.f, dv = 0.f;
#pragma unroll
      for (int u = 0; u < T; u++) {
        dq = fmaf(Bf[t * T + u], S[S_K + u * D + c], dq);                  // dq[t][c] = sum_u dS[t][u] k[u][c]
        dk = fmaf(Bf[u * T + t], S[S_Q + u * D + c], dk);                  // dk[t][c] = sum_u dS[u][t] q[u][c]
        dv = fmaf(S[S_P + u * T + t], Cf[u * D + c], dv);                  // dv[t][c] = sum_u P[u][t] dctl[u][c]
      }
      __syncthreads();
      A[t * D + c] = dq; Bf[t * D + c] = dk; Cf[t * D + c] = dv;           // (Bf's dS and Cf's dctl were fully consumed before the sync)
      __syncthreads();
      dxin += fc_b(sW, A + t * D, c) + fc_b(sW + D * WS, Bf + t * D, c) + fc_b(sW + 2 * D * WS, Cf + t * D, c);
      dx = dxin;
    }
    // normalisation: x0 = z r, r = rsqrt(mean z^2 + eps)  ->  dz = r dx0 - z r^3 / n * sum(dx0 z)
    const float zv = zb[t * D + c];
    const float dot = block_sum(dx * zv, red);
    dzb[t * D + c] = r * dx - zv * r * r * r * dot / (float)(T * D);
  } else if (threadIdx.x < 32) {
    const int c = threadIdx.x;
    const float* G_ = params + LOCAL_SIZE;
    float* xs = sm; float* hs = sm + D; float* sv = sm + 2 * D;           // sv: NB x (H0, XO)
    float outv;
    const float r = global_forward(G_, zb + T * D, xs, hs, sv, &outv);
    float g = 0.f;
    for (int l = 0; l < num_ws; l++) g += gb[((long long)T * num_ws + l) * D + c];
    xs[c] = g * SQRT2 * dlrelu(outv); __syncwarp();
    float dx = gfc_b(G_ + GO_W, xs, c); __syncwarp();
    for (int l = NB - 1; l >= 0; l--) {
      const float* L = G_ + l * G_SIZE;
      const float dsum = dx * dlrelu(sv[l * 2 * D + D + c]);
      xs[c] = dsum; __syncwarp();
      const float dh0 = gfc_b(L + G_W1, xs, c) * SQRT2 * dlrelu(sv[l * 2 * D + c]);
      hs[c] = dh0; __syncwarp();
      dx = dsum + gfc_b(L + G_W0, hs, c); __syncwarp();
    }
    const float zv = zb[T * D + c];
    const float dot = warp_sum(dx * zv);
    dzb[T * D + c] = r * dx - zv * r * r * r * dot / (float)D;
  }
}

}  // namespace
}  // namespace mgf

using namespace mgf;

extern "C" int mgf_mapping_param_floats(void) { return PARAM_FLOATS; }

// z [B,17,32] fp32, params = packed weights (mgf_mapping_param_floats floats, layout in mapping.cu / engine.pack_mapping),
// maskbias [B,16] = (1 - mask) * -10000; ws [B,17,num_ws,32] fp32 (every layer slot written).
extern "C" int mgf_mapping_fwd(const float* z, const float* params, const float* maskbias, float* ws, int B, int num_ws, void* stream) {
  if (!z || !params || !maskbias || !ws) MGF_FAIL(MGF_E_BADARG, "mapping_fwd: null tensor");
  if (B <= 0 || num_ws <= 0) MGF_FAIL(MGF_E_SHAPE, "mapping_fwd: empty batch");
  mapping_fwd_kernel<<<dim3(B, 2), 512, SM_FWD_FLOATS * sizeof(float), (cudaStream_t)stream>>>(z, params, maskbias, ws, num_ws);
  MGF_CHECK_LAUNCH("mapping_fwd");
  return 0;
}

// dz [B,17,32] = d(loss)/dz given dws [B,17,num_ws,32]; recomputes the forward internally (nothing is saved by mgf_mapping_fwd).
extern "C" int mgf_mapping_bwd(const float* z, const float* params, const float* maskbias, const float* dws, float* dz, int B, int num_ws, void* stream) {
  if (!z || !params || !maskbias || !dws || !dz) MGF_FAIL(MGF_E_BADARG, "mapping_bwd: null tensor");
  if (B <= 0 || num_ws <= 0) MGF_FAIL(MGF_E_SHAPE, "mapping_bwd: empty batch");
  const int smem = SM_BWD_FLOATS * (int)sizeof(float);
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(mapping_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cfg = true; }
  mapping_bwd_kernel<<<dim3(B, 2), 512, smem, (cudaStream_t)stream>>>(z, params, maskbias, dws, dz, num_ws);
  MGF_CHECK_LAUNCH("mapping_bwd");
  return 0;
}
