// lpips.cu -- HBM-bound kernels of the projection loss (reference lpips/networks_basic.py:64-101, lpips/__init__.py:44-46,
// torch.nn.MSELoss at 1024_example_percept_MSE.py:143,224) and the fused latent update (torch.optim.Adam + latent_noise,
// 1024_example_percept_MSE.py:117,134-135,153).  The VGG16 convolutions themselves run on the tcgen05 kernel (conv_tc.cu).
//   prep      : ScalingLayer + fp32 NCHW -> bf16 NHWC 3x3 im2col (27 -> 32 channels) for conv1_1 as a K=32 GEMM, + MSE partial sums
//   prep_bwd  : col2im of the conv1_1 input gradient, / scale, + MSE gradient  -> d(img) fp32 NCHW
//   maxpool   : 2x2 NHWC forward / backward (backward also adds the LPIPS tap gradient and applies the ReLU mask)
//   head      : channel-unit-normalise, squared difference to the cached target features, 1x1 `lin`, spatial mean (fwd + bwd)
//   adam      : Adam with coupled L2 on the [B,17,32] latents, lr / noise schedule read from device arrays (graph-capturable)
#include "common.cuh"

namespace mgf {

__constant__ float c_shift[3] = {-.030f, -.088f, -.188f};
__constant__ float c_scale[3] = {.458f, .448f, .450f};

// one thread per pixel: reads the 3x3x3 neighbourhood (L1/L2 cached), writes 64 bytes
__global__ void __launch_bounds__(256) lpips_prep_kernel(const float* img, const float* target, __nv_bfloat16* col, float* mse, int R, int do_col, bool f16) {
  const int b = blockIdx.y;
  const long long HW = (long long)R * R;
  float local = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(p / R), x = (int)(p % R);
    const float* ib = img + (long long)b * 3 * HW;
    if (target) {
#pragma unroll
      for (int c = 0; c < 3; c++) { const float d = ib[c * HW + p] - target[((long long)b * 3 + c) * HW + p]; local = fmaf(d, d, local); }
    }
    if (do_col) {
      float v[32];
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        const bool in = yy >= 0 && yy < R && xx >= 0 && xx < R;
#pragma unroll
        for (int c = 0; c < 3; c++) v[t * 3 + c] = in ? (__ldg(ib + c * HW + (long long)yy * R + xx) - c_shift[c]) / c_scale[c] : 0.f;
      }
#pragma unroll
      for (int j = 27; j < 32; j++) v[j] = 0.f;
      uint4* op = reinterpret_cast<uint4*>(col + ((long long)b * HW + p) * 32);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint4 u;
        u.x = pack16(v[q * 8], v[q * 8 + 1], f16); u.y = pack16(v[q * 8 + 2], v[q * 8 + 3], f16);
        u.z = pack16(v[q * 8 + 4], v[q * 8 + 5], f16); u.w = pack16(v[q * 8 + 6], v[q * 8 + 7], f16);
        op[q] = u;
      }
    }
  }
  if (target) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.f) atomicAdd(&mse[b], local);
  }
}

// dimg[b,c,Y,X] = (1/scale_c) * sum_t dcol[b, Y-dy_t, X-dx_t, t*3+c] + mcoef * (img - target)
__global__ void __launch_bounds__(256) lpips_prep_bwd_kernel(const __nv_bfloat16* dcol, const float* img, const float* target, float mcoef,
                                                             float* dimg, int R) {
  const int b = blockIdx.y;
  const long long HW = (long long)R * R;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const int Y = (int)(p / R), X = (int)(p % R);
    float g[3] = {0.f, 0.f, 0.f};
    if (dcol) {
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const int yy = Y - (t / 3 - 1), xx = X - (t % 3 - 1);
        if (yy >= 0 && yy < R && xx >= 0 && xx < R) {
          const __nv_bfloat16* cp = dcol + ((long long)b * HW + (long long)yy * R + xx) * 32 + t * 3;
#pragma unroll
          for (int c = 0; c < 3; c++) g[c] += __bfloat162float(cp[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const long long o = ((long long)b * 3 + c) * HW + p;
      float v = g[c] / c_scale[c];
      if (target) v = fmaf(mcoef, img[o] - target[o], v);
      dimg[o] = v;
    }
  }
}

template <bool f16>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int H, int W, int C) {
  const int b = blockIdx.y, vecs = C / 8, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)Ho * Wo * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vecs); long long t = i / vecs;
    const int xo = (int)(t % Wo), yo = (int)(t / Wo);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; e++) m[e] = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C + cv * 8));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = unpack16(w4[e], f16); m[e * 2] = fmaxf(m[e * 2], f.x); m[e * 2 + 1] = fmaxf(m[e * 2 + 1], f.y); }
    }
    uint4 o;
    o.x = pack16(m[0], m[1], f16); o.y = pack16(m[2], m[3], f16); o.z = pack16(m[4], m[5], f16); o.w = pack16(m[6], m[7], f16);
    *reinterpret_cast<uint4*>(y + (((long long)b * Ho + yo) * Wo + xo) * C + cv * 8) = o;
  }
}

// dx[b,y,x,c] = ((first arg-max of the 2x2 window ? dy : 0) + extra) * (x > 0)      (x is a post-ReLU tensor)
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const __nv_bfloat16* x, const __nv_bfloat16* dy, const __nv_bfloat16* extra,
                                                           __nv_bfloat16* dx, int H, int W, int C, bool f16) {
  const int b = blockIdx.y, vecs = C / 8, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)Ho * Wo * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vecs); long long t = i / vecs;
    const int xo = (int)(t % Wo), yo = (int)(t / Wo);
    float v[4][8], g[8];
    {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)b * Ho + yo) * Wo + xo) * C + cv * 8));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = unpack_bf16(w4[e]); g[e * 2] = f.x; g[e * 2 + 1] = f.y; }
    }
    long long off[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      off[k] = (((long long)b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C + cv * 8;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off[k]));
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = unpack16(w4[e], f16); v[k][e * 2] = f.x; v[k][e * 2 + 1] = f.y; }
    }
    int arg[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      int a = 0; float m = v[0][e];
#pragma unroll
      for (int k = 1; k < 4; k++) if (v[k][e] > m) { m = v[k][e]; a = k; }
      arg[e] = a;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; e++) o[e] = (arg[e] == k) ? g[e] : 0.f;
      if (extra) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(extra + off[k]));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; e++) { const float2 f = unpack_bf16(w4[e]); o[e * 2] += f.x; o[e * 2 + 1] += f.y; }
      }
#pragma unroll
      for (int e = 0; e < 8; e++) o[e] = v[k][e] > 0.f ? o[e] : 0.f;
      uint4 u;
      u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + off[k]) = u;
    }
  }
}

// mode 0: n1 = f/(|f|+eps) (target caching).  mode 1: val[b] += (1/HW) sum_c lin_c (f_c/(|f|+eps) - n1_c)^2.
// mode 2: backward, df = coef[b]/HW * (g/(r+eps) - (g.f) f / (r (r+eps)^2)), g = 2 lin (n0 - n1); optional relu mask (df *= f > 0).
// A group of LPP = min(32, C/8) lanes owns one pixel (VPL = C/(8*LPP) 16-byte vectors per lane, kept in registers), so a warp
// covers 32/LPP pixels and every lane is busy for all VGG widths (64..512); reductions are xor-shuffles inside the group.
template <int MODE, int VPL, bool f16>
__global__ void __launch_bounds__(256) lpips_head_kernel(const __nv_bfloat16* __restrict__ f, const __nv_bfloat16* __restrict__ n1, const float* __restrict__ lin,
                                                         const float* __restrict__ coef, __nv_bfloat16* __restrict__ outp, float* val, long long HW, int C, int relu_mask) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpp = (C / 8) / VPL;            // lanes per pixel: 8, 16 or 32
  const int ppw = 32 / lpp;                 // pixels per warp
  const int sub = lane / lpp, ll = lane % lpp;
  const float eps = 1e-10f;
  float vsum = 0.f;
  const float cf = (MODE == 2) ? coef[b] / (float)HW : 0.f;
  float lw[VPL][8];
  if (MODE != 0) {
#pragma unroll
    for (int q = 0; q < VPL; q++)
#pragma unroll
      for (int e = 0; e < 8; e++) lw[q][e] = lin[(q * lpp + ll) * 8 + e];
  }
  for (long long p0 = ((long long)blockIdx.x * 8 + warp) * ppw; p0 < HW; p0 += (long long)gridDim.x * 8 * ppw) {
    const long long p = p0 + sub;
    const bool ok = p < HW;
    const long long row = ((long long)b * HW + (ok ? p : 0)) * C;
    float x[VPL][8], t[VPL][8];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; q++) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(f + row) + q * lpp + ll);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 v = unpack16(w4[e], f16); x[q][e * 2] = v.x; x[q][e * 2 + 1] = v.y; ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); }
      if (MODE != 0) {
        const uint4 un = __ldg(reinterpret_cast<const uint4*>(n1 + row) + q * lpp + ll);
        const uint32_t n4[4] = {un.x, un.y, un.z, un.w};
#pragma unroll
        for (int e = 0; e < 4; e++) { const float2 v = unpack16(n4[e], f16); t[q][e * 2] = v.x; t[q][e * 2 + 1] = v.y; }
      }
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float r = sqrtf(ss), inv = 1.f / (r + eps);
    if (MODE == 0) {
      if (ok) {
#pragma unroll
        for (int q = 0; q < VPL; q++) {
          uint4 u;
          u.x = pack16(x[q][0] * inv, x[q][1] * inv, f16); u.y = pack16(x[q][2] * inv, x[q][3] * inv, f16);
          u.z = pack16(x[q][4] * inv, x[q][5] * inv, f16); u.w = pack16(x[q][6] * inv, x[q][7] * inv, f16);
          reinterpret_cast<uint4*>(outp + row)[q * lpp + ll] = u;
        }
      }
      continue;
    }
    float dot = 0.f, v1 = 0.f;
#pragma unroll
    for (int q = 0; q < VPL; q++)
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const float dd = x[q][e] * inv - t[q][e];
        if (MODE == 1) v1 = fmaf(lw[q][e] * dd, dd, v1);
        else dot = fmaf(2.f * lw[q][e] * dd, x[q][e], dot);
      }
    if (MODE == 1) { if (ok) vsum += v1; continue; }
    for (int o = lpp >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float k2 = (r > 0.f) ? dot * inv * inv / r : 0.f;
    if (ok) {
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          g[e] = cf * (2.f * lw[q][e] * (x[q][e] * inv - t[q][e]) * inv - k2 * x[q][e]);
          if (relu_mask) g[e] = x[q][e] > 0.f ? g[e] : 0.f;
        }
        uint4 u;
        u.x = pack_bf16(g[0], g[1]); u.y = pack_bf16(g[2], g[3]); u.z = pack_bf16(g[4], g[5]); u.w = pack_bf16(g[6], g[7]);
        reinterpret_cast<uint4*>(outp + row)[q * lpp + ll] = u;
      }
    }
  }
  if (MODE == 1) {
    vsum = warp_sum(vsum);
    if (lane == 0) atomicAdd(&val[b], vsum / (float)HW);
  }
}

// Fused forward of an LPIPS tap that is followed by a max-pool: val[b] += (1/HW) sum_p sum_c lin_c (f_c/(|f|+eps) - n1_c)^2 over the four
// pixels of every 2x2 window AND y = max over the window, in one pass over f (the separate head + pool kernels read f twice).
// Same lane-group-per-window layout as the backward kernel below.
template <int VPL, bool f16>
__global__ void __launch_bounds__(256, VPL == 1 ? 3 : 2) lpips_tap_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ n1, const float* __restrict__ lin,
                                                                 __nv_bfloat16* __restrict__ y, float* val, float2* __restrict__ stats, int H, int W, int C) {
  // stats (optional) [B,H,W] = (|f|, g . f) per pixel with g = 2 lin (f inv - n1): the two channel reductions the backward kernel needs; they cost
  // this pass one FFMA per element and 8 bytes per pixel, and take the whole first pass (3 FMAs per element + shuffles) out of the backward kernel
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpp = (C / 8) / VPL, wpw = 32 / lpp;
  const int sub = lane / lpp, ll = lane % lpp;
  const int Ho = H / 2, Wo = W / 2;
  const long long nwin = (long long)Ho * Wo, HW = (long long)H * W;
  const float eps = 1e-10f;
  float lw[VPL][8];
#pragma unroll
  for (int q = 0; q < VPL; q++)
#pragma unroll
    for (int e = 0; e < 8; e++) lw[q][e] = lin[(q * lpp + ll) * 8 + e];
  float vsum = 0.f;
  for (long long w0 = ((long long)blockIdx.x * 8 + warp) * wpw; w0 < nwin; w0 += (long long)gridDim.x * 8 * wpw) {
    const long long wi = w0 + sub;
    const bool ok = wi < nwin;
    const int xo = (int)((ok ? wi : 0) % Wo), yo = (int)((ok ? wi : 0) / Wo);
    uint4 xr[4][VPL], nr[4][VPL];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long row = (((long long)b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C;
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        xr[k][q] = __ldg(reinterpret_cast<const uint4*>(x + row) + q * lpp + ll);
        nr[k][q] = __ldg(reinterpret_cast<const uint4*>(n1 + row) + q * lpp + ll);
      }
    }
    float mx[VPL][8];
#pragma unroll
    for (int q = 0; q < VPL; q++)
#pragma unroll
      for (int e = 0; e < 8; e++) mx[q][e] = -3.0e38f;
    float wsum = 0.f;
    // pass 1 over the packed registers: |f|^2 of the four pixels and the running max; the four shuffle chains are interleaved (in-order issue:
    // one pixel at a time left each chain's latency exposed).  Pass 2: the distance in its direct form (f inv - n1)^2, which stays accurate
    // when the projection converges and the two terms nearly cancel, and g . f for the backward kernel.
    float ss[4], inv[4], ds[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      ss[k] = 0.f;
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        const uint32_t w4[4] = {xr[k][q].x, xr[k][q].y, xr[k][q].z, xr[k][q].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 v = unpack16(w4[e], f16);
          mx[q][e * 2] = fmaxf(mx[q][e * 2], v.x); mx[q][e * 2 + 1] = fmaxf(mx[q][e * 2 + 1], v.y);
          ss[k] = fmaf(v.x, v.x, ss[k]); ss[k] = fmaf(v.y, v.y, ss[k]);
        }
      }
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 4; k++) ss[k] += __shfl_xor_sync(0xffffffffu, ss[k], o);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      ss[k] = sqrtf(ss[k]);
      inv[k] = 1.f / (ss[k] + eps);
      ds[k] = 0.f;                          // sum_c lin (f inv - n1) f  =  inv sa - sb
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        const uint32_t w4[4] = {xr[k][q].x, xr[k][q].y, xr[k][q].z, xr[k][q].w}, n4[4] = {nr[k][q].x, nr[k][q].y, nr[k][q].z, nr[k][q].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 v = unpack16(w4[e], f16), t = unpack16(n4[e], f16);
          const float d0 = fmaf(v.x, inv[k], -t.x), d1 = fmaf(v.y, inv[k], -t.y);
          const float l0 = lw[q][e * 2] * d0, l1 = lw[q][e * 2 + 1] * d1;
          wsum = fmaf(l0, d0, wsum); wsum = fmaf(l1, d1, wsum);
          ds[k] = fmaf(l0, v.x, ds[k]); ds[k] = fmaf(l1, v.y, ds[k]);
        }
      }
    }
    if (stats) {
      for (int o = lpp >> 1; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) ds[k] += __shfl_xor_sync(0xffffffffu, ds[k], o);
      }
      if (ok && ll < 4) {                   // lane k of the group writes pixel k (select, not a dynamic index: keeps ss / ds in registers)
        const float r_ = ll == 0 ? ss[0] : ll == 1 ? ss[1] : ll == 2 ? ss[2] : ss[3];
        const float d_ = ll == 0 ? ds[0] : ll == 1 ? ds[1] : ll == 2 ? ds[2] : ds[3];
        stats[((long long)b * H + 2 * yo + (ll >> 1)) * W + 2 * xo + (ll & 1)] = make_float2(r_, 2.f * d_);
      }
    }
    if (ok) {
      vsum += wsum;
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        uint4 o;
        o.x = pack16(mx[q][0], mx[q][1], f16); o.y = pack16(mx[q][2], mx[q][3], f16); o.z = pack16(mx[q][4], mx[q][5], f16); o.w = pack16(mx[q][6], mx[q][7], f16);
        reinterpret_cast<uint4*>(y + (((long long)b * Ho + yo) * Wo + xo) * C)[q * lpp + ll] = o;
      }
    }
  }
  vsum = warp_sum(vsum);
  __shared__ float red[8];
  if (lane == 0) red[warp] = vsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i];
    if (t != 0.f) atomicAdd(&val[b], t / (float)HW);
  }
}

// packed 16-bit pair helpers (storage type chosen at compile time): element-wise max, equality mask and "> 0" mask (0xFFFF per half)
template <bool F16> __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  if (F16) { const __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b)); return *reinterpret_cast<const uint32_t*>(&r); }
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b)); return *reinterpret_cast<const uint32_t*>(&r);
}
template <bool F16> __device__ __forceinline__ uint32_t eqmask2(uint32_t a, uint32_t b) {
  if (F16) return __heq2_mask(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return __heq2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
}
template <bool F16> __device__ __forceinline__ uint32_t gtzmask2(uint32_t a) {
  if (F16) return __hgt2_mask(*reinterpret_cast<__half2*>(&a), __float2half2_rn(0.f));
  return __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&a), __float2bfloat162_rn(0.f));
}

// Fused backward of an LPIPS tap that is followed by a max-pool (relu1_2, relu2_2, relu3_3, relu4_3):
//   dx = ( route(dy through the 2x2 max-pool) + d(head)/dx ) * (x > 0)
// replaces lpips_head<2> + maxpool2_bwd for those taps: x and n1 are read once, the head gradient never touches HBM
// (7 tensor passes -> 3.25).  A group of LPP lanes owns one 2x2 window (4 pixels), channel vectors stay in registers.
template <int VPL, bool f16, bool STATS>      // f16: compile-time forward dtype (the run-time flag cost a select per unpacked pair in an issue-bound kernel)
__global__ void __launch_bounds__(256, VPL == 1 ? 3 : 2) lpips_tap_pool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ n1,
                                                                 const float* __restrict__ lin, const float* __restrict__ coef,
                                                                 const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                                 const float2* __restrict__ stats, int H, int W, int C) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpp = (C / 8) / VPL, wpw = 32 / lpp;                // lanes per window, windows per warp
  const int sub = lane / lpp, ll = lane % lpp;
  const int Ho = H / 2, Wo = W / 2;
  const long long nwin = (long long)Ho * Wo, HW = (long long)H * W;
  const float eps = 1e-10f, cf = coef[b] / (float)HW;
  float lw[VPL][8];
#pragma unroll
  for (int q = 0; q < VPL; q++)
#pragma unroll
    for (int e = 0; e < 8; e++) lw[q][e] = lin[(q * lpp + ll) * 8 + e];
  for (long long w0 = ((long long)blockIdx.x * 8 + warp) * wpw; w0 < nwin; w0 += (long long)gridDim.x * 8 * wpw) {
    const long long wi = w0 + sub;
    const bool ok = wi < nwin;
    const int xo = (int)((ok ? wi : 0) % Wo), yo = (int)((ok ? wi : 0) / Wo);
    // the 4 pixels of the window stay PACKED in registers (16-bit pairs) and are unpacked on use: 32 registers instead of 128
    uint4 xr[4][VPL], nr[4][VPL];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long row = (((long long)b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C;
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        xr[k][q] = __ldg(reinterpret_cast<const uint4*>(x + row) + q * lpp + ll);
        nr[k][q] = __ldg(reinterpret_cast<const uint4*>(n1 + row) + q * lpp + ll);
      }
    }
    // per pixel: inv = 1 / (|f| + eps) and k2 = (g . f) inv^2 / |f| with g = 2 lin (f inv - n1).  The two channel reductions come from the
    // forward kernel (stats) when it ran; otherwise one pass over the packed registers: |f|^2, sum lin f^2, sum lin f n1 and one shuffle round
    float inv[4], k2[4];
    if (STATS) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 st = __ldg(stats + ((long long)b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1));
        inv[k] = 1.f / (st.x + eps);
        k2[k] = (st.x > 0.f) ? st.y * inv[k] * inv[k] / st.x : 0.f;
      }
    } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float ss = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
      for (int q = 0; q < VPL; q++) {
        const uint32_t w4[4] = {xr[k][q].x, xr[k][q].y, xr[k][q].z, xr[k][q].w}, n4[4] = {nr[k][q].x, nr[k][q].y, nr[k][q].z, nr[k][q].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 v = unpack16(w4[e], f16), t = unpack16(n4[e], f16);
          ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss);
          sa = fmaf(lw[q][e * 2] * v.x, v.x, sa); sa = fmaf(lw[q][e * 2 + 1] * v.y, v.y, sa);
          sb = fmaf(lw[q][e * 2] * v.x, t.x, sb); sb = fmaf(lw[q][e * 2 + 1] * v.y, t.y, sb);
        }
      }
      for (int o = lpp >> 1; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o); sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o);
      }
      const float r = sqrtf(ss);
      inv[k] = 1.f / (r + eps);
      const float dot = 2.f * (inv[k] * sa - sb);                // g . f  with g = 2 lin (f inv - n1)
      k2[k] = (r > 0.f) ? dot * inv[k] * inv[k] / r : 0.f;
    }
    }
    if (!ok) continue;
#pragma unroll
    for (int q = 0; q < VPL; q++) {
      // max-pool routing and ReLU masks on the PACKED 16-bit pairs (comparisons are exact in the storage type): sel[k][e] = 0xFFFF per half
      // where pixel k is the FIRST maximum of the 2x2 window (torch's max_pool2d backward), pos[k][e] = 0xFFFF where x > 0.  Two elements per
      // instruction and no fp32 copy of the four pixels (the scalar version spent 9 compare/select instructions per element on the arg-max
      // alone and kept 32 more registers live)
      uint32_t sel[4][4], pos[4][4], gsel[4];
      {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)b * Ho + yo) * Wo + xo) * C) + q * lpp + ll);
        gsel[0] = u.x; gsel[1] = u.y; gsel[2] = u.z; gsel[3] = u.w;
      }
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const uint32_t w0 = (&xr[0][q].x)[e], w1 = (&xr[1][q].x)[e], w2 = (&xr[2][q].x)[e], w3 = (&xr[3][q].x)[e];
        const uint32_t m = max2<f16>(max2<f16>(w0, w1), max2<f16>(w2, w3));
        const uint32_t e0 = eqmask2<f16>(w0, m), e1 = eqmask2<f16>(w1, m), e2 = eqmask2<f16>(w2, m);
        sel[0][e] = e0; sel[1][e] = e1 & ~e0; sel[2][e] = e2 & ~(e0 | e1); sel[3][e] = ~(e0 | e1 | e2);
        pos[0][e] = gtzmask2<f16>(w0); pos[1][e] = gtzmask2<f16>(w1); pos[2][e] = gtzmask2<f16>(w2); pos[3][e] = gtzmask2<f16>(w3);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t w4[4] = {xr[k][q].x, xr[k][q].y, xr[k][q].z, xr[k][q].w}, n4[4] = {nr[k][q].x, nr[k][q].y, nr[k][q].z, nr[k][q].w};
        uint32_t ow[4];
        // head gradient h = cf (2 lin (x inv - t) inv - k2 x) = x * (ca lin - ck) - t * (cb lin): two FMAs per element
        const float ca = 2.f * cf * inv[k] * inv[k], cb = 2.f * cf * inv[k], ck = cf * k2[k];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 xx = unpack16(w4[e], f16), t = unpack16(n4[e], f16);
          const float2 gr = unpack_bf16(gsel[e] & sel[k][e]);          // routed pooled gradient (bf16 pair) or zero
          const float l0 = lw[q][e * 2], l1 = lw[q][e * 2 + 1];
          const float v0 = fmaf(xx.x, fmaf(ca, l0, -ck), fmaf(-(cb * l0), t.x, gr.x));
          const float v1 = fmaf(xx.y, fmaf(ca, l1, -ck), fmaf(-(cb * l1), t.y, gr.y));
          ow[e] = pack_bf16(v0, v1) & pos[k][e];                       // ReLU mask of the tapped activation
        }
        const long long row = (((long long)b * H + 2 * yo + (k >> 1)) * W + 2 * xo + (k & 1)) * C;
        reinterpret_cast<uint4*>(dx + row)[q * lpp + ll] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
  }
}

// Adam (coupled L2, torch.optim.Adam semantics) on n latent floats, then the next step's noisy latent.
// sched[step] = {lr, noise_strength_next}; step read from a device counter (so the whole step can live in a CUDA graph).
__global__ void __launch_bounds__(256) adam_noise_kernel(float* latent, const float* grad, float* m, float* v, const float* noise_all,
                                                         int noise_rows, float* latent_n, const float* sched, const int* step_ptr,
                                                         float beta1, float beta2, float eps, float wd, long long n) {
  const int step = *step_ptr;                 // 0-based index of the step being applied
  // sched has noise_rows rows (one per scheduled step): a launch past the schedule's end re-uses its last row instead of reading
  // beyond the tensor (the host API refuses to step past the end; this bounds a replayed CUDA graph as well)
  const int srow = (noise_rows > 0 && step >= noise_rows) ? noise_rows - 1 : step;
  const float lr = sched[2 * srow], ns = sched[2 * srow + 1];
  const float t = (float)(step + 1);
  const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  const float* noise_next = (noise_all && step + 1 < noise_rows) ? noise_all + (long long)(step + 1) * n : nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float p = latent[i];
    const float g = grad[i] + wd * p;
    const float mi = beta1 * m[i] + (1.f - beta1) * g;
    const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p -= (lr / bc1) * (mi / denom);
    latent[i] = p;
    if (latent_n) latent_n[i] = p + (noise_next ? noise_next[i] * ns : 0.f);
  }
}
__global__ void step_inc_kernel(int* step_ptr) { if (threadIdx.x == 0 && blockIdx.x == 0) *step_ptr += 1; }

static inline unsigned gridp(long long work, int waves = 8) {
  long long blocks = (work + 255) / 256; const long long cap = (long long)num_sms() * waves;
  if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1; return (unsigned)blocks;
}
}  // namespace mgf

using namespace mgf;

extern "C" int mgf_lpips_prep(const float* img, const float* target, void* col, float* mse, int B, int R, void* stream) {
  if (!img || (!col && !target) || (target && !mse)) MGF_FAIL(MGF_E_BADARG, "lpips_prep: null tensor");
  dim3 grid(gridp((long long)R * R), B);
  lpips_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, target, (__nv_bfloat16*)col, mse, R, col != nullptr, fwd_f16());
  MGF_CHECK_LAUNCH("lpips_prep");
  return 0;
}
extern "C" int mgf_lpips_prep_bwd(const void* dcol, const float* img, const float* target, float mcoef, float* dimg, int B, int R, void* stream) {
  if (!dimg || (target && !img)) MGF_FAIL(MGF_E_BADARG, "lpips_prep_bwd: null tensor");
  dim3 grid(gridp((long long)R * R), B);
  lpips_prep_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, img, target, mcoef, dimg, R);
  MGF_CHECK_LAUNCH("lpips_prep_bwd");
  return 0;
}
extern "C" int mgf_maxpool2_fwd(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y) MGF_FAIL(MGF_E_BADARG, "maxpool2_fwd: null tensor");
  if (C % 8 || H % 2 || W % 2) MGF_FAIL(MGF_E_SHAPE, "maxpool2_fwd: C%%8, H%%2, W%%2 must be 0");
  dim3 grid(gridp((long long)(H / 2) * (W / 2) * (C / 8)), B);
  if (fwd_f16()) maxpool2_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, H, W, C);
  else maxpool2_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, H, W, C);
  MGF_CHECK_LAUNCH("maxpool2_fwd");
  return 0;
}
extern "C" int mgf_maxpool2_bwd(const void* x, const void* dy, const void* extra, void* dx, int B, int H, int W, int C, void* stream) {
  if (!x || !dy || !dx) MGF_FAIL(MGF_E_BADARG, "maxpool2_bwd: null tensor");
  if (C % 8 || H % 2 || W % 2) MGF_FAIL(MGF_E_SHAPE, "maxpool2_bwd: C%%8, H%%2, W%%2 must be 0");
  dim3 grid(gridp((long long)(H / 2) * (W / 2) * (C / 8)), B);
  maxpool2_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)extra, (__nv_bfloat16*)dx, H, W, C, fwd_f16());
  MGF_CHECK_LAUNCH("maxpool2_bwd");
  return 0;
}
extern "C" int mgf_lpips_head(int mode, const void* f, const void* n1, const float* lin, const float* coef, void* out, float* val,
                              int relu_mask, int B, int64_t HW, int C, void* stream) {
  if (!f || (mode != 0 && (!n1 || !lin)) || (mode != 1 && !out) || (mode == 1 && !val) || (mode == 2 && !coef)) MGF_FAIL(MGF_E_BADARG, "lpips_head: null tensor");
  if (!(C == 64 || C == 128 || C == 256 || C == 512)) MGF_FAIL(MGF_E_SHAPE, "lpips_head: C=%d must be 64, 128, 256 or 512", C);
  if (mode < 0 || mode > 2) MGF_FAIL(MGF_E_BADARG, "lpips_head: mode %d", mode);
  const int vpl = C == 512 ? 2 : 1, lpp = (C / 8) / vpl, ppw = 32 / lpp;
  long long blocks = (HW + 8 * ppw - 1) / (8 * ppw); const long long cap = (long long)num_sms() * 8; if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, B);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16 *fp = (const __nv_bfloat16*)f, *np = (const __nv_bfloat16*)n1;
  __nv_bfloat16* op = (__nv_bfloat16*)out;
#define MGF_HEAD(M, V) do { if (fwd_f16()) lpips_head_kernel<M, V, true><<<grid, 256, 0, st>>>(fp, np, lin, coef, op, val, HW, C, relu_mask); \
                            else lpips_head_kernel<M, V, false><<<grid, 256, 0, st>>>(fp, np, lin, coef, op, val, HW, C, relu_mask); } while (0)
  if (vpl == 1) { if (mode == 0) MGF_HEAD(0, 1); else if (mode == 1) MGF_HEAD(1, 1); else MGF_HEAD(2, 1); }
  else { if (mode == 0) MGF_HEAD(0, 2); else if (mode == 1) MGF_HEAD(1, 2); else MGF_HEAD(2, 2); }
#undef MGF_HEAD
  MGF_CHECK_LAUNCH("lpips_head");
  return 0;
}
extern "C" int mgf_lpips_tap_pool_fwd(const void* x, const void* n1, const float* lin, void* y, float* val, void* stats, int B, int H, int W, int C, void* stream) {
  if (!x || !n1 || !lin || !y || !val) MGF_FAIL(MGF_E_BADARG, "lpips_tap_pool_fwd: null tensor");
  if (!(C == 64 || C == 128 || C == 256 || C == 512) || H % 2 || W % 2) MGF_FAIL(MGF_E_SHAPE, "lpips_tap_pool_fwd: C in {64,128,256,512}, even H/W");
  const int vpl = C == 512 ? 2 : 1, lpp = (C / 8) / vpl, wpw = 32 / lpp;
  const long long nwin = (long long)(H / 2) * (W / 2);
  // about two waves of resident CTAs over all images (3 or 2 CTAs per SM): every warp then walks several windows, so the per-warp set-up (the
  // lin weights) and the per-CTA reduction are amortised on the small taps too (relu3_3 / relu4_3 ran 1-2 windows per warp)
  long long blocks = (nwin + 8 * wpw - 1) / (8 * wpw);
  long long cap = ((long long)num_sms() * (vpl == 1 ? 6 : 4) + B - 1) / B; if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, B);
  cudaStream_t st = (cudaStream_t)stream;
#define MGF_TPF(V, F) lpips_tap_pool_fwd_kernel<V, F><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)n1, lin, (__nv_bfloat16*)y, val, (float2*)stats, H, W, C)
  if (fwd_f16()) { if (vpl == 1) MGF_TPF(1, true); else MGF_TPF(2, true); }
  else { if (vpl == 1) MGF_TPF(1, false); else MGF_TPF(2, false); }
#undef MGF_TPF
  MGF_CHECK_LAUNCH("lpips_tap_pool_fwd");
  return 0;
}
extern "C" int mgf_lpips_tap_pool_bwd(const void* x, const void* n1, const float* lin, const float* coef, const void* dy, void* dx,
                                      const void* stats, int B, int H, int W, int C, void* stream) {
  if (!x || !n1 || !lin || !coef || !dy || !dx) MGF_FAIL(MGF_E_BADARG, "lpips_tap_pool_bwd: null tensor");
  if (!(C == 64 || C == 128 || C == 256 || C == 512) || H % 2 || W % 2) MGF_FAIL(MGF_E_SHAPE, "lpips_tap_pool_bwd: C in {64,128,256,512}, even H/W");
  const int vpl = C == 512 ? 2 : 1, lpp = (C / 8) / vpl, wpw = 32 / lpp;
  const long long nwin = (long long)(H / 2) * (W / 2);
  long long blocks = (nwin + 8 * wpw - 1) / (8 * wpw);
  long long cap = (long long)num_sms() * 8;
  if (vpl == 2) { cap = ((long long)num_sms() * 4 + B - 1) / B; if (cap < 1) cap = 1; }      // 512-channel tap: two resident CTAs per SM, about two waves
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, B);
  cudaStream_t st = (cudaStream_t)stream;
#define MGF_TPB(V, F) do { if (stats) lpips_tap_pool_bwd_kernel<V, F, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)n1, lin, coef, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, (const float2*)stats, H, W, C); \
    else lpips_tap_pool_bwd_kernel<V, F, false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)n1, lin, coef, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, nullptr, H, W, C); } while (0)
  if (fwd_f16()) { if (vpl == 1) MGF_TPB(1, true); else MGF_TPB(2, true); }
  else { if (vpl == 1) MGF_TPB(1, false); else MGF_TPB(2, false); }
#undef MGF_TPB
  MGF_CHECK_LAUNCH("lpips_tap_pool_bwd");
  return 0;
}
extern "C" int mgf_adam_noise_step(float* latent, const float* grad, float* m, float* v, const float* noise_all, int noise_rows, float* latent_n,
                                   const float* sched, int* step_ptr, float beta1, float beta2, float eps, float weight_decay, int64_t n, void* stream) {
  if (!latent || !grad || !m || !v || !sched || !step_ptr) MGF_FAIL(MGF_E_BADARG, "adam_noise_step: null tensor");
  adam_noise_kernel<<<gridp(n, 1), 256, 0, (cudaStream_t)stream>>>(latent, grad, m, v, noise_all, noise_rows, latent_n, sched, step_ptr, beta1, beta2, eps, weight_decay, n);
  MGF_CHECK_LAUNCH("adam_noise_step");
  step_inc_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_ptr);
  MGF_CHECK_LAUNCH("adam_noise_step(step)");
  return 0;
}
