// engine_kernels.cu -- the small / HBM-bound kernels of the bf16 synthesis engine (everything around the tcgen05 convs):
//   style affine + demodulation coefficients (networks.py:1022, :288-293) and their backward,
//   per-sample weight modulation into bf16 GEMM operands (the "fused_modconv" weights, networks.py:288-293),
//   tiny batched GEMM for the attention value/modulation fold, ToRGB forward/backward (networks.py:1054-1065),
//   leaky-ReLU backward with the demodulation-gradient reduction, NHWC FIR up-sampling (+ residual add) and its adjoint
//   (resnet skip path, networks.py:245-250, upfirdn2d.py:300-335).
// Activations are NHWC bf16; all math in fp32.
#include "common.cuh"

namespace mgf {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) { t = warp_sum(t); if (l == 0) red[0] = t; }
  __syncthreads();
  return red[0];
}

// ---- styles s[b,i] = ((wg[b,:] . A[i,:]) * again + abias[i]) * sgain       grid (ceil(Cin/256), B)
__global__ void __launch_bounds__(256) style_s_kernel(const float* wg, long long wg_stride, const float* A, const float* abias,
                                                      float again, float sgain, float* s_out, int Cin, int wdim) {
  __shared__ float w[64];
  const int b = blockIdx.y;
  for (int j = threadIdx.x; j < wdim; j += blockDim.x) w[j] = wg[(long long)b * wg_stride + j];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cin) return;
  float acc = 0.f;
  for (int j = 0; j < wdim; j++) acc = fmaf(w[j], A[i * wdim + j], acc);
  s_out[(long long)b * Cin + i] = (acc * again + abias[i]) * sgain;
}
// ---- demodulation d[b,o] = rsqrt(sum_i s[b,i]^2 Wsq[o,i] + 1e-8): one warp per (b,o), coalesced Wsq rows.  grid (ceil(O/8), B)
__global__ void __launch_bounds__(256) style_d_kernel(const float* s, const float* Wsq, float* d_out, int Cin, int O) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= O) return;
  const float* sr = s + (long long)b * Cin; const float* wr = Wsq + (long long)o * Cin;
  float acc = 0.f;
  for (int i = lane; i < Cin; i += 32) { const float v = sr[i]; acc = fmaf(v * v, wr[i], acc); }
  acc = warp_sum(acc);
  if (lane == 0) d_out[(long long)b * O + o] = rsqrtf(acc + 1e-8f);
}

// backward: ds_total[i] = ds[i] - s[i] * sum_o R[o] d[o]^2 Wsq[o,i]   (R[o] = sum_p dy*y, so dL/dd = R/d and dd/ds_i = -d^3 s_i Wsq)
//           dwg[j] += sgain * again * sum_i ds_total[i] A[i,j]
// block = (32 channels i, sample b): 8 warps split the o range, lane = channel (coalesced Wsq columns); partial dwg by atomics.
__global__ void __launch_bounds__(256) style_bwd_kernel(const float* ds, const float* R, const float* s, const float* d, const float* Wsq,
                                                        const float* A, float again, float sgain, float* dwg, long long dwg_stride,
                                                        int Cin, int O, int wdim) {
  __shared__ float part[8][32];
  __shared__ float dst[32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (R && i < Cin) {
    for (int o = warp; o < O; o += 8) {
      const float dv = d[(long long)b * O + o];
      acc = fmaf(R[(long long)b * O + o] * dv * dv, Wsq[(long long)o * Cin + i], acc);
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
    if (i < Cin) {
      t = ds ? ds[(long long)b * Cin + i] : 0.f;
      if (R) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) a += part[w][lane];
        t -= a * s[(long long)b * Cin + i];
      }
    }
    dst[lane] = t;
  }
  __syncthreads();
  // dwg[j] partial over this block's 32 channels: thread j (< wdim)
  for (int j = threadIdx.x; j < wdim; j += blockDim.x) {
    float a = 0.f;
#pragma unroll 8
    for (int l = 0; l < 32; l++) { const int ii = blockIdx.x * 32 + l; if (ii < Cin) a = fmaf(dst[l], A[ii * wdim + j], a); }
    atomicAdd(&dwg[(long long)b * dwg_stride + j], a * again * sgain);
  }
}

// ---- per-sample operand weights: out[b][t][n][k] = base[t][n][k] * rs[b][n % nmod] * cs[b][k]   (bf16)
__global__ void __launch_bounds__(256) modulate_kernel(const float* base, const float* rs, int nmod, const float* cs,
                                                       __nv_bfloat16* out, long long TN, int K, int NT, bool of16, unsigned int* ovf) {
  const long long b = blockIdx.y;
  float mx = 0.f;
  const long long per = TN * K;
  const int k4 = K >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < TN * k4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / k4; const int k = (int)(i % k4) * 4;
    const float4 w = *reinterpret_cast<const float4*>(base + row * K + k);
    float r = 1.f;
    if (rs) r = rs[b * nmod + (int)((row % NT) % nmod)];
    float4 c = make_float4(1.f, 1.f, 1.f, 1.f);
    if (cs) c = *reinterpret_cast<const float4*>(cs + b * K + k);
    uint2 o;
    const float o0 = w.x * r * c.x, o1 = w.y * r * c.y, o2 = w.z * r * c.z, o3 = w.w * r * c.w;
    mx = ovf_max(ovf_max(ovf_max(ovf_max(mx, o0), o1), o2), o3);
    o.x = pack16(o0, o1, of16);
    o.y = pack16(o2, o3, of16);
    *reinterpret_cast<uint2*>(out + b * per + row * K + k) = o;
  }
  ovf_commit(ovf, mx);
}

// ---- per-(sample, channel) scaling of an NHWC 16-bit tensor: out[b,p,c] = x[b,p,c] * sc[b,c]   (activation-side modulation for
// the low-resolution layers, where per-sample weight tensors would dwarf the activations)
__global__ void __launch_bounds__(256) scale_channels_kernel(const __nv_bfloat16* x, const float* sc, __nv_bfloat16* out, long long HW, int C, bool f16, unsigned int* ovf) {
  const int b = blockIdx.y, vecs = C / 8;
  float mx = 0.f;
  const long long total = HW * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vecs);
    const long long off = (long long)b * HW * C + i * 8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + off));
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + (long long)b * C + cv * 8)), s1 = __ldg(reinterpret_cast<const float4*>(sc + (long long)b * C + cv * 8 + 4));
    const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
    const float2 a = unpack16(w4[0], f16), bq = unpack16(w4[1], f16), c = unpack16(w4[2], f16), d = unpack16(w4[3], f16);
    uint4 o;
    const float v[8] = {a.x * s0.x, a.y * s0.y, bq.x * s0.z, bq.y * s0.w, c.x * s1.x, c.y * s1.y, d.x * s1.z, d.y * s1.w};
#pragma unroll
    for (int j = 0; j < 8; j++) mx = ovf_max(mx, v[j]);
    o.x = pack16(v[0], v[1], f16); o.y = pack16(v[2], v[3], f16);
    o.z = pack16(v[4], v[5], f16); o.w = pack16(v[6], v[7], f16);
    *reinterpret_cast<uint4*>(out + off) = o;
  }
  ovf_commit(ovf, mx);
}

// ---- tiny batched GEMM: out[b,m,n] = sum_k A[b,m,k] * Bm[n,k] (+ bias[n]); A strided (sAb, sAm), optional accumulate into out
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* A, long long sAb, long long sAm, const float* Bm, const float* bias,
                                                         float* out, long long sOb, long long sOm, int M, int N, int K, int accumulate) {
  const int b = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * N; e += gridDim.x * blockDim.x) {
    const int m = e / N, n = e % N;
    const float* a = A + (long long)b * sAb + (long long)m * sAm;
    const float* w = Bm + (long long)n * K;
    float acc = bias ? bias[n] : 0.f;
    for (int k = 0; k < K; k++) acc = fmaf(a[k], w[k], acc);
    float* o = out + (long long)b * sOb + (long long)m * sOm + n;
    *o = accumulate ? (*o + acc) : acc;
  }
}

// long-K variant: one warp per output element, lanes stride over k
__global__ void __launch_bounds__(256) small_gemm_warp_kernel(const float* A, long long sAb, long long sAm, const float* Bm, const float* bias,
                                                              float* out, long long sOb, long long sOm, int M, int N, int K, int accumulate) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (e >= M * N) return;
  const int m = e / N, n = e % N;
  const float* a = A + (long long)b * sAb + (long long)m * sAm;
  const float* w = Bm + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(a[k], w[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += bias ? bias[n] : 0.f;
    float* o = out + (long long)b * sOb + (long long)m * sOm + n;
    *o = accumulate ? (*o + acc) : acc;
  }
}

// ---- ToRGB: img[b,c,p] = sum_o y[b,p,o] * wrgb[c,o] * s[b,o] + bias[c]    (y NHWC bf16 -> img NCHW fp32, 3 channels)
template <int C>
__global__ void __launch_bounds__(256) torgb_fwd_kernel(const __nv_bfloat16* y, const float* wrgb, const float* s, const float* bias,
                                                        float* img, long long HW, bool yf16) {
  __shared__ float ws[3 * C];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) ws[i] = wrgb[i] * s[(long long)b * C + (i % C)];
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const uint4* yp = reinterpret_cast<const uint4*>(y + ((long long)b * HW + p) * C);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int q = 0; q < C / 8; q++) {
      const uint4 u = __ldg(yp + q);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float2 f = unpack16(w4[e], yf16);
        const int o = q * 8 + e * 2;
        a0 = fmaf(f.x, ws[o], a0); a0 = fmaf(f.y, ws[o + 1], a0);
        a1 = fmaf(f.x, ws[C + o], a1); a1 = fmaf(f.y, ws[C + o + 1], a1);
        a2 = fmaf(f.x, ws[2 * C + o], a2); a2 = fmaf(f.y, ws[2 * C + o + 1], a2);
      }
    }
    float* ip = img + (long long)b * 3 * HW + p;
    ip[0] = a0 + bias[0]; ip[HW] = a1 + bias[1]; ip[2 * HW] = a2 + bias[2];
  }
}

// backward: dy[b,p,o] = sum_c dimg[b,c,p] wrgb[c,o] s[b,o];  ds[b,o] += sum_{p,c} dimg wrgb[c,o] y[b,p,o];  R[b,o] += sum_p dy*y
// thread = (pixel lane, 8-channel vector); per-CTA shared accumulators, then one global atomic per channel per CTA.
template <bool yf16>
__global__ void __launch_bounds__(256) torgb_bwd_kernel(const float* __restrict__ dimg, const __nv_bfloat16* __restrict__ y, const float* __restrict__ wrgb,
                                                        const float* __restrict__ s, __nv_bfloat16* __restrict__ dy, float* ds, float* R, long long HW, int C, int pix_per_cta) {
  extern __shared__ float sm[];
  float* acc_ds = sm;        // [C]
  float* acc_R = sm + C;     // [C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lanes_c = C / 8, tp = 256 / lanes_c;
  const int cv = threadIdx.x % lanes_c, pl = threadIdx.x / lanes_c;
  float w0[8], w1[8], w2[8], sv[8], lds[8], lR[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int o = cv * 8 + e;
    w0[e] = wrgb[o]; w1[e] = wrgb[C + o]; w2[e] = wrgb[2 * C + o]; sv[e] = s[(long long)b * C + o]; lds[e] = 0.f; lR[e] = 0.f;
  }
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  long long pe = p0 + pix_per_cta; if (pe > HW) pe = HW;
  // four pixels per trip, all their loads issued before the first use: one load round trip per pixel kept this kernel at 0.4 of the HBM
  // roofline (long_scoreboard was its only stall reason)
  constexpr int UP = 4;
  for (long long p = p0 + pl; p < pe; p += (long long)UP * tp) {
    float g0[UP], g1[UP], g2[UP]; uint4 u[UP];
#pragma unroll
    for (int k = 0; k < UP; k++) {
      const long long q = p + (long long)k * tp;
      const bool ok = q < pe;
      const float* ip = dimg + (long long)b * 3 * HW + (ok ? q : p);
      g0[k] = __ldg(ip); g1[k] = __ldg(ip + HW); g2[k] = __ldg(ip + 2 * HW);
      u[k] = __ldg(reinterpret_cast<const uint4*>(y + ((long long)b * HW + (ok ? q : p)) * C + cv * 8));
    }
#pragma unroll
    for (int k = 0; k < UP; k++) {
      const long long q = p + (long long)k * tp;
      if (q < pe) {
        const uint32_t w4[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 f = unpack16(w4[e], yf16);
          const float t0 = g0[k] * w0[e * 2] + g1[k] * w1[e * 2] + g2[k] * w2[e * 2];
          const float t1 = g0[k] * w0[e * 2 + 1] + g1[k] * w1[e * 2 + 1] + g2[k] * w2[e * 2 + 1];
          const float d0 = t0 * sv[e * 2], d1 = t1 * sv[e * 2 + 1];
          lds[e * 2] += t0 * f.x; lds[e * 2 + 1] += t1 * f.y;
          lR[e * 2] += d0 * f.x; lR[e * 2 + 1] += d1 * f.y;
          o4[e] = pack_bf16(d0, d1);
        }
        *reinterpret_cast<uint4*>(dy + ((long long)b * HW + q) * C + cv * 8) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; e++) { atomicAdd(&acc_ds[cv * 8 + e], lds[e]); atomicAdd(&acc_R[cv * 8 + e], lR[e]); }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) { atomicAdd(&ds[(long long)b * C + i], acc_ds[i]); atomicAdd(&R[(long long)b * C + i], acc_R[i]); }
}

// ---- leaky-ReLU backward with demod-gradient reduction.  z = lrelu(y + noise*ns + bias) * gain (what the conv epilogue stored).
// dy = dz * gain * (z > 0 ? 1 : alpha)  (mode 0) or dy = dz (mode 1: dz is already the pre-activation gradient);
// R[b,o] += sum_p dy * y with y = lrelu^-1(z/gain) - noise*ns - bias.   One CTA per (pixel chunk, sample); C <= 512.
template <bool zf16>
__global__ void __launch_bounds__(256, 4) act_bwd_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ dy, float* R,
                                                      const float* __restrict__ noise, const float* __restrict__ nstr, const float* __restrict__ bias,
                                                      float alpha, float gain, int mode, long long HW, int C, int pix_per_cta, long long nbs) {
  extern __shared__ float racc[];   // [C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C; i += blockDim.x) racc[i] = 0.f;
  __syncthreads();
  const int vecs = C / 8;                       // uint4 per pixel
  const int lanes_c = vecs < 256 ? vecs : 256;  // threads along channels
  const int tp = 256 / lanes_c;                 // pixels processed concurrently
  const int cv = threadIdx.x % lanes_c, pl = threadIdx.x / lanes_c;
  const float ns = (noise && nstr) ? *nstr : 0.f;
  const long long p0 = (long long)blockIdx.x * pix_per_cta;
  float bsv[8], lr[8];
#pragma unroll
  for (int e = 0; e < 8; e++) { bsv[e] = bias ? bias[cv * 8 + e] : 0.f; lr[e] = 0.f; }
  const float inv_gain = 1.f / gain, inv_alpha = 1.f / alpha;
  if (pl < tp) {
    long long pe = p0 + pix_per_cta; if (pe > HW) pe = HW;
    // two pixels per trip with all their loads issued first (long_scoreboard was the only stall reason at 5.0-5.3 TB/s)
    constexpr int UP = 2;
    for (long long p = p0 + pl; p < pe; p += (long long)UP * tp) {
      uint4 uzv[UP], udv[UP]; float nzv[UP];
#pragma unroll
      for (int k = 0; k < UP; k++) {
        const long long q = p + (long long)k * tp;
        const long long qq = q < pe ? q : p;
        const long long off = ((long long)b * HW + qq) * C + cv * 8;
        uzv[k] = __ldg(reinterpret_cast<const uint4*>(z + off));
        udv[k] = __ldg(reinterpret_cast<const uint4*>(dz + off));
        nzv[k] = noise ? __ldg(noise + b * nbs + qq) * ns : 0.f;
      }
#pragma unroll
      for (int k = 0; k < UP; k++) {
        const long long q = p + (long long)k * tp;
        if (q < pe) {
          const long long off = ((long long)b * HW + q) * C + cv * 8;
          const float nz = nzv[k];
          const uint32_t z4[4] = {uzv[k].x, uzv[k].y, uzv[k].z, uzv[k].w}, d4[4] = {udv[k].x, udv[k].y, udv[k].z, udv[k].w};
          uint32_t o4[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float2 zf = unpack16(z4[e], zf16), df = unpack_bf16(d4[e]);
            float g0 = df.x, g1 = df.y;
            if (mode == 0) { g0 *= gain * (zf.x > 0.f ? 1.f : alpha); g1 *= gain * (zf.y > 0.f ? 1.f : alpha); }
            const float u0 = zf.x * inv_gain, u1 = zf.y * inv_gain;
            const float y0 = (u0 > 0.f ? u0 : u0 * inv_alpha) - nz - bsv[e * 2], y1 = (u1 > 0.f ? u1 : u1 * inv_alpha) - nz - bsv[e * 2 + 1];
            lr[e * 2] += g0 * y0; lr[e * 2 + 1] += g1 * y1;
            o4[e] = pack_bf16(g0, g1);
          }
          if (dy) *reinterpret_cast<uint4*>(dy + off) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; e++) atomicAdd(&racc[cv * 8 + e], lr[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&R[(long long)b * C + i], racc[i]);
}


static inline unsigned grid_for(long long work, int per_block = 256, int waves = 8) {
  long long blocks = (work + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace mgf

using namespace mgf;

extern "C" int mgf_style_fwd(const float* wg, int64_t wg_stride, const float* A, const float* abias, float again, float sgain,
                             const float* Wsq, float* s_out, float* d_out, int B, int Cin, int O, int wdim, void* stream) {
  if (!wg || !A || !abias || !s_out || (Wsq && !d_out)) MGF_FAIL(MGF_E_BADARG, "style_fwd: null tensor");
  if (wdim > 64) MGF_FAIL(MGF_E_SHAPE, "style_fwd: w_dim %d > 64", wdim);
  if (B <= 0) return 0;
  style_s_kernel<<<dim3((Cin + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(wg, wg_stride, A, abias, again, sgain, s_out, Cin, wdim);
  MGF_CHECK_LAUNCH("style_fwd(s)");
  if (Wsq) {
    style_d_kernel<<<dim3((O + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(s_out, Wsq, d_out, Cin, O);
    MGF_CHECK_LAUNCH("style_fwd(d)");
  }
  return 0;
}

extern "C" int mgf_style_bwd(const float* ds, const float* R, const float* s, const float* d, const float* Wsq, const float* A,
                             float again, float sgain, float* dwg, int64_t dwg_stride, int B, int Cin, int O, int wdim, void* stream) {
  if (!A || !dwg || (R && (!s || !d || !Wsq))) MGF_FAIL(MGF_E_BADARG, "style_bwd: null tensor");
  if (R && sgain != 1.f) MGF_FAIL(MGF_E_UNSUP, "style_bwd: demodulation with a style gain is not a case of the generator");
  if (B <= 0) return 0;
  style_bwd_kernel<<<dim3((Cin + 31) / 32, B), 256, 0, (cudaStream_t)stream>>>(ds, R, s, d, Wsq, A, again, sgain, dwg, dwg_stride, Cin, O, wdim);
  MGF_CHECK_LAUNCH("style_bwd");
  return 0;
}

extern "C" int mgf_modulate_weights(const float* base, const float* rs, int nmod, const float* cs, void* out, int out_fwd,
                                    int B, int64_t T, int64_t NT, int64_t K, void* stream) {
  if (!base || !out) MGF_FAIL(MGF_E_BADARG, "modulate_weights: null tensor");
  if (K % 4) MGF_FAIL(MGF_E_SHAPE, "modulate_weights: K must be a multiple of 4");
  if (B <= 0) return 0;
  const long long TN = T * NT;
  dim3 grid(grid_for(TN * (K / 4), 256, 4), B);
  modulate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(base, rs, nmod > 0 ? nmod : 1, cs, (__nv_bfloat16*)out, TN, (int)K, (int)NT, out_fwd && fwd_f16(),
                                                          (out_fwd && fwd_f16()) ? overflow_flag() : nullptr);
  MGF_CHECK_LAUNCH("modulate_weights");
  return 0;
}

extern "C" int mgf_small_gemm(const float* A, int64_t sAb, int64_t sAm, const float* Bm, const float* bias, float* out,
                              int64_t sOb, int64_t sOm, int B, int M, int N, int K, int accumulate, void* stream) {
  if (!A || !Bm || !out) MGF_FAIL(MGF_E_BADARG, "small_gemm: null tensor");
  if (B <= 0 || M * N == 0) return 0;
  if (K >= 128) {
    dim3 grid((M * N + 7) / 8, B);
    small_gemm_warp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, sAb, sAm, Bm, bias, out, sOb, sOm, M, N, K, accumulate);
  } else {
    dim3 grid((M * N + 255) / 256, B);
    small_gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, sAb, sAm, Bm, bias, out, sOb, sOm, M, N, K, accumulate);
  }
  MGF_CHECK_LAUNCH("small_gemm");
  return 0;
}

extern "C" int mgf_torgb_fwd(const void* y, const float* wrgb, const float* s, const float* bias, float* img, int B, int64_t HW, int C, void* stream) {
  if (!y || !wrgb || !s || !bias || !img) MGF_FAIL(MGF_E_BADARG, "torgb_fwd: null tensor");
  dim3 grid(grid_for(HW, 256, 8), B);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 32) torgb_fwd_kernel<32><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, wrgb, s, bias, img, HW, fwd_f16());
  else if (C == 64) torgb_fwd_kernel<64><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, wrgb, s, bias, img, HW, fwd_f16());
  else if (C == 128) torgb_fwd_kernel<128><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, wrgb, s, bias, img, HW, fwd_f16());
  else if (C == 256) torgb_fwd_kernel<256><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, wrgb, s, bias, img, HW, fwd_f16());   // 64^2 .. 128^2 generators
  else if (C == 512) torgb_fwd_kernel<512><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, wrgb, s, bias, img, HW, fwd_f16());
  else MGF_FAIL(MGF_E_UNSUP, "torgb_fwd: C=%d not in {32,64,128,256,512}", C);
  MGF_CHECK_LAUNCH("torgb_fwd");
  return 0;
}

extern "C" int mgf_torgb_bwd(const float* dimg, const void* y, const float* wrgb, const float* s, void* dy, float* ds, float* R,
                             int B, int64_t HW, int C, void* stream) {
  if (!dimg || !y || !wrgb || !s || !dy || !ds || !R) MGF_FAIL(MGF_E_BADARG, "torgb_bwd: null tensor");
  if (C % 8 || C / 8 > 256 || 256 % (C / 8)) MGF_FAIL(MGF_E_SHAPE, "torgb_bwd: C/8 must divide 256");
  long long ppc = (HW * B + (long long)num_sms() * 8 - 1) / ((long long)num_sms() * 8);
  if (ppc < 16) ppc = 16;
  dim3 grid((unsigned)((HW + ppc - 1) / ppc), B);
  if (fwd_f16()) torgb_bwd_kernel<true><<<grid, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(dimg, (const __nv_bfloat16*)y, wrgb, s, (__nv_bfloat16*)dy, ds, R, HW, C, (int)ppc);
  else torgb_bwd_kernel<false><<<grid, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(dimg, (const __nv_bfloat16*)y, wrgb, s, (__nv_bfloat16*)dy, ds, R, HW, C, (int)ppc);
  MGF_CHECK_LAUNCH("torgb_bwd");
  return 0;
}

extern "C" int mgf_act_bwd(const void* dz, const void* z, void* dy, float* R, const float* noise, const float* nstr, const float* bias,
                           float alpha, float gain, int mode, int B, int64_t HW, int C, int64_t noise_bstride, void* stream) {
  if (!dz || !z || !R) MGF_FAIL(MGF_E_BADARG, "act_bwd: null tensor");
  if (C % 8 || C > 4096) MGF_FAIL(MGF_E_SHAPE, "act_bwd: C must be a multiple of 8");
  if ((C / 8) < 256 && 256 % (C / 8)) MGF_FAIL(MGF_E_SHAPE, "act_bwd: C/8 must divide 256");
  if ((C / 8) > 256) MGF_FAIL(MGF_E_SHAPE, "act_bwd: C too large");
  // enough pixels per CTA that the per-CTA atomics stay cheap, enough CTAs to fill the machine
  long long ppc = (HW * B + (long long)num_sms() * 8 - 1) / ((long long)num_sms() * 8);
  if (ppc < 16) ppc = 16;
  dim3 grid((unsigned)((HW + ppc - 1) / ppc), B);
  if (fwd_f16()) act_bwd_kernel<true><<<grid, 256, C * sizeof(float), (cudaStream_t)stream>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)z, (__nv_bfloat16*)dy, R,
                                                                        noise, nstr, bias, alpha, gain, mode, HW, C, (int)ppc, (long long)noise_bstride);
  else act_bwd_kernel<false><<<grid, 256, C * sizeof(float), (cudaStream_t)stream>>>((const __nv_bfloat16*)dz, (const __nv_bfloat16*)z, (__nv_bfloat16*)dy, R,
                                                                        noise, nstr, bias, alpha, gain, mode, HW, C, (int)ppc, (long long)noise_bstride);
  MGF_CHECK_LAUNCH("act_bwd");
  return 0;
}

static int log2_exact(int v) { int s = 0; while ((1 << s) < v) s++; return (1 << s) == v ? s : -1; }




extern "C" int mgf_scale_channels(const void* x, const float* sc, void* out, int is_fwd, int B, int64_t HW, int C, void* stream) {
  if (!x || !sc || !out) MGF_FAIL(MGF_E_BADARG, "scale_channels: null tensor");
  if (C % 8) MGF_FAIL(MGF_E_SHAPE, "scale_channels: C must be a multiple of 8");
  if (B <= 0 || HW <= 0) return 0;
  dim3 grid(grid_for(HW * (C / 8), 256, 8), B);
  scale_channels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, sc, (__nv_bfloat16*)out, HW, C, is_fwd && fwd_f16(),
                                                                      (is_fwd && fwd_f16()) ? overflow_flag() : nullptr);
  MGF_CHECK_LAUNCH("scale_channels");
  return 0;
}
