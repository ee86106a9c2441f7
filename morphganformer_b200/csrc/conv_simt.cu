// conv_simt.cu -- exact-fp32 direct convolution (forward, input gradient / transposed conv, weight gradient)
// for the conv2d_gradfix surface (reference torch_utils/ops/conv2d_gradfix.py:27-35, :96-157 call ATen/cuDNN).
// Implicit GEMM on CUDA cores with FFMA accumulation (no TF32 rounding) so the fp32 parity path can meet the
// 1e-4 image tolerance; the throughput path is the tcgen05 kernel in conv_tc.cu.
// 64x64x16 block tiles, 256 threads, 4x4 register micro-tiles, im2col indices decoded on the fly,
// arbitrary stride / padding / dilation / groups.
#include "common.cuh"

namespace mgf {

struct ConvP {
  const float *a, *b, *bias; float* out;
  int N, IC, H, W, OC, HO, WO, KH, KW, sh, sw, ph, pw, dh, dw, groups;
  int ICg, OCg; long long M, NC, K; int ksplit; long long kchunk;
};

constexpr int BM = 64, BN = 64, BK = 16;

// MODE 0: forward   M=OCg  cols=(n,oy,ox)   K=(ic,ky,kx)   a=w  b=x   out=y
// MODE 1: dgrad     M=ICg  cols=(n,iy,ix)   K=(oc,ky,kx)   a=w  b=dy  out=dx
// MODE 2: wgrad     M=OCg  cols=(ic,ky,kx)  K=(n,oy,ox)    a=dy b=x   out=dw (atomic)
template <int MODE>
__device__ __forceinline__ float load_a(const ConvP& p, int g, long long m, long long k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (MODE == 0) return p.a[((long long)g * p.OCg + m) * p.K + k];
  if (MODE == 1) {
    const int kk = p.KH * p.KW; const long long oc = k / kk; const int r = (int)(k % kk);
    return p.a[(((long long)g * p.OCg + oc) * p.ICg + m) * kk + r];
  }
  const long long hw = (long long)p.HO * p.WO; const long long n = k / hw; const long long r = k % hw;
  return p.a[(n * p.OC + (long long)g * p.OCg + m) * hw + r];
}

template <int MODE>
__device__ __forceinline__ float load_b(const ConvP& p, int g, long long k, long long col) {
  if (col >= p.NC || k >= p.K) return 0.f;
  if (MODE == 0) {
    const int kk = p.KH * p.KW; const int ic = (int)(k / kk); const int r = (int)(k % kk); const int ky = r / p.KW, kx = r % p.KW;
    const long long hw = (long long)p.HO * p.WO; const long long n = col / hw; const int q = (int)(col % hw);
    const int oy = q / p.WO, ox = q % p.WO;
    const int iy = oy * p.sh - p.ph + ky * p.dh, ix = ox * p.sw - p.pw + kx * p.dw;
    if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) return 0.f;
    return p.b[((n * p.IC + (long long)g * p.ICg + ic) * p.H + iy) * p.W + ix];
  }
  if (MODE == 1) {
    const int kk = p.KH * p.KW; const int oc = (int)(k / kk); const int r = (int)(k % kk); const int ky = r / p.KW, kx = r % p.KW;
    const long long hw = (long long)p.H * p.W; const long long n = col / hw; const int q = (int)(col % hw);
    const int iy = q / p.W, ix = q % p.W;
    const int ty = iy + p.ph - ky * p.dh, tx = ix + p.pw - kx * p.dw;
    if (ty < 0 || tx < 0 || (ty % p.sh) || (tx % p.sw)) return 0.f;
    const int oy = ty / p.sh, ox = tx / p.sw;
    if (oy >= p.HO || ox >= p.WO) return 0.f;
    return p.b[((n * p.OC + (long long)g * p.OCg + oc) * p.HO + oy) * p.WO + ox];
  }
  const int kk = p.KH * p.KW; const int ic = (int)(col / kk); const int r = (int)(col % kk); const int ky = r / p.KW, kx = r % p.KW;
  const long long hw = (long long)p.HO * p.WO; const long long n = k / hw; const int q = (int)(k % hw);
  const int oy = q / p.WO, ox = q % p.WO;
  const int iy = oy * p.sh - p.ph + ky * p.dh, ix = ox * p.sw - p.pw + kx * p.dw;
  if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) return 0.f;
  return p.b[((n * p.IC + (long long)g * p.ICg + ic) * p.H + iy) * p.W + ix];
}

template <int MODE>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int g = blockIdx.z / p.ksplit, ks = blockIdx.z % p.ksplit;
  const long long m0 = (long long)blockIdx.y * BM, c0 = (long long)blockIdx.x * BN;
  const long long kbeg = (long long)ks * p.kchunk;
  long long kend = kbeg + p.kchunk; if (kend > p.K) kend = p.K;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    // A tile: for MODE 0 consecutive k is contiguous; for MODE 1/2 consecutive m (MODE 2: consecutive k) -- pick mapping per mode
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = tid + i * 256;
      int kk, mm;
      if (MODE == 0) { kk = e % BK; mm = e / BK; } else if (MODE == 1) { kk = e % BK; mm = e / BK; } else { kk = e % BK; mm = e / BK; }
      const long long k = k0 + kk;
      As[kk][mm] = (k < kend) ? load_a<MODE>(p, g, m0 + mm, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = tid + i * 256;
      int kk, cc;
      if (MODE == 2) { kk = e % BK; cc = e / BK; } else { cc = e % BN; kk = e / BN; }
      const long long k = k0 + kk;
      Bs[kk][cc] = (k < kend) ? load_b<MODE>(p, g, k, c0 + cc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const long long m = m0 + tm + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const long long col = c0 + tn + j;
      if (col >= p.NC) continue;
      if (MODE == 0) {
        const long long hw = (long long)p.HO * p.WO; const long long n = col / hw, q = col % hw;
        const long long ch = (long long)g * p.OCg + m;
        p.out[(n * p.OC + ch) * hw + q] = acc[i][j] + (p.bias ? p.bias[ch] : 0.f);
      } else if (MODE == 1) {
        const long long hw = (long long)p.H * p.W; const long long n = col / hw, q = col % hw;
        p.out[(n * p.IC + (long long)g * p.ICg + m) * hw + q] = acc[i][j];
      } else {
        float* dst = &p.out[((long long)g * p.OCg + m) * p.NC + col];
        if (p.ksplit == 1) *dst = acc[i][j]; else atomicAdd(dst, acc[i][j]);
      }
    }
  }
}

static int check_shape(const mgf_conv_shape* s, const char* who) {
  if (!s) MGF_FAIL(MGF_E_BADARG, "%s: null shape", who);
  if (s->N < 0 || s->IC < 1 || s->OC < 1 || s->H < 1 || s->W < 1 || s->HO < 1 || s->WO < 1 || s->KH < 1 || s->KW < 1)
    MGF_FAIL(MGF_E_SHAPE, "%s: bad sizes", who);
  if (s->groups < 1 || s->IC % s->groups || s->OC % s->groups) MGF_FAIL(MGF_E_SHAPE, "%s: channels not divisible by groups", who);
  if (s->stride_h < 1 || s->stride_w < 1 || s->dil_h < 1 || s->dil_w < 1 || s->pad_h < 0 || s->pad_w < 0) MGF_FAIL(MGF_E_BADARG, "%s: bad stride/dilation/padding", who);
  return 0;
}
static ConvP make_p(const mgf_conv_shape* s) {
  ConvP p{};
  p.N = s->N; p.IC = s->IC; p.H = s->H; p.W = s->W; p.OC = s->OC; p.HO = s->HO; p.WO = s->WO; p.KH = s->KH; p.KW = s->KW;
  p.sh = s->stride_h; p.sw = s->stride_w; p.ph = s->pad_h; p.pw = s->pad_w; p.dh = s->dil_h; p.dw = s->dil_w; p.groups = s->groups;
  p.ICg = s->IC / s->groups; p.OCg = s->OC / s->groups; p.ksplit = 1;
  return p;
}
template <int MODE>
static int launch(ConvP& p, cudaStream_t st, const char* who) {
  const long long gx = (p.NC + BN - 1) / BN, gy = (p.M + BM - 1) / BM, gz = (long long)p.groups * p.ksplit;
  if (gx > 0x7fffffffLL || gy > 65535 || gz > 65535) MGF_FAIL(MGF_E_SHAPE, "%s: grid too large (%lld,%lld,%lld)", who, gx, gy, gz);
  conv_simt_kernel<MODE><<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), 256, 0, st>>>(p);
  MGF_CHECK_LAUNCH(who);
  return 0;
}
}  // namespace mgf

extern "C" int mgf_conv2d_fwd_f32(const float* x, const float* w, const float* bias, float* y, const mgf_conv_shape* s, void* stream) {
  using namespace mgf;
  if (int e = check_shape(s, "conv2d_fwd")) return e;
  if (!x || !w || !y) MGF_FAIL(MGF_E_BADARG, "conv2d_fwd: null tensor");
  if (s->N == 0) return 0;
  ConvP p = make_p(s); p.a = w; p.b = x; p.bias = bias; p.out = y;
  p.M = p.OCg; p.NC = (long long)s->N * s->HO * s->WO; p.K = (long long)p.ICg * s->KH * s->KW; p.kchunk = p.K;
  return launch<0>(p, (cudaStream_t)stream, "conv2d_fwd");
}
extern "C" int mgf_conv2d_dgrad_f32(const float* dy, const float* w, float* dx, const mgf_conv_shape* s, void* stream) {
  using namespace mgf;
  if (int e = check_shape(s, "conv2d_dgrad")) return e;
  if (!dy || !w || !dx) MGF_FAIL(MGF_E_BADARG, "conv2d_dgrad: null tensor");
  if (s->N == 0) return 0;
  ConvP p = make_p(s); p.a = w; p.b = dy; p.bias = nullptr; p.out = dx;
  p.M = p.ICg; p.NC = (long long)s->N * s->H * s->W; p.K = (long long)p.OCg * s->KH * s->KW; p.kchunk = p.K;
  return launch<1>(p, (cudaStream_t)stream, "conv2d_dgrad");
}
extern "C" int mgf_conv2d_wgrad_f32(const float* dy, const float* x, float* dw, const mgf_conv_shape* s, void* stream) {
  using namespace mgf;
  if (int e = check_shape(s, "conv2d_wgrad")) return e;
  if (!dy || !x || !dw) MGF_FAIL(MGF_E_BADARG, "conv2d_wgrad: null tensor");
  ConvP p = make_p(s); p.a = dy; p.b = x; p.bias = nullptr; p.out = dw;
  p.M = p.OCg; p.NC = (long long)p.ICg * s->KH * s->KW; p.K = (long long)s->N * s->HO * s->WO;
  if (p.K == 0) return 0;
  // split K so that the grid fills the machine; partial sums are combined with float atomics (dw pre-zeroed by caller)
  const long long tiles = ((p.NC + BN - 1) / BN) * ((p.M + BM - 1) / BM) * p.groups;
  long long want = (4LL * num_sms() + tiles - 1) / tiles;
  long long maxsplit = (p.K + 255) / 256; if (want > maxsplit) want = maxsplit;
  if (want < 1) want = 1;
  if (want * p.groups > 65535) want = 65535 / p.groups; if (want < 1) want = 1;
  p.ksplit = (int)want; p.kchunk = ((p.K + want - 1) / want + BK - 1) / BK * BK;
  return launch<2>(p, (cudaStream_t)stream, "conv2d_wgrad");
}
