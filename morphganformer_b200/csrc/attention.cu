// attention.cu -- fused duplex (bipartite) attention layer of the GANformer synthesis network, forward and backward,
// for the default configuration (reference training/networks.py:748-822 TransformerLayer.forward with kmeans=True,
// parametric centroids, 1 head, integration="mul", norm="layer"; followed by the layer tail :1036-1040 noise + bias_act).
//
// Algebra (exact in real arithmetic; everything input-independent is folded once on the host because weights are frozen):
//   scores S[f,t] = X[f,:] . Kf[t,:] + Sc[f,t] + maskbias[b,t]      Kf = (att_weight_q * centroid_q) Wq / sqrt(C)   [16,C]
//                                                                   Sc = (bq . awc_q + P[f,:] . awc_p) / sqrt(C)     [HW,16]
//   A = softmax_t(S);  ctl[f,:] = A[f,:] VM[b] + bm                 VM[b] = (Y_b Wv^T + bv) Wm^T                      [16,C]
//   xn = X * rsqrt(mean_c X^2 + 1e-8);  u = xn * (1 + ctl) + noise*ns + bias;  out = lrelu(u) * gain
// so one pass over X does the work of ~40 eager ops and ~12 HBM round trips of the reference.  The dead K-projection / QK^T /
// centroid-assignment work of the reference is never executed.
//
// Mapping: one warp per pixel, lane l owns channel vectors {l, l+32} (8 bf16 = 16 bytes each) -> fully coalesced
// 512-byte rows; Kf and VM[b] live in shared memory; warp-shuffle all-reduces for the 16 scores.  HBM-bound:
// algorithmic bytes = 2 * B*HW*C*2 forward (read X, write out), 3x that backward (read X, dz; write dX).
#include "common.cuh"

namespace mgf {

constexpr int T16 = 16;

// Shared-memory layout of the [16, C] coefficient tables: a lane owns 8 consecutive channels (one 16-byte bf16 vector of X) and reads
// them as two float4; storing the two halves in separate planes makes both reads 16-byte-strided across lanes (conflict-free)
// instead of 32-byte-strided (2-way bank conflict on every LDS.128: 42 M conflicts per launch in the first ncu capture).
__device__ __forceinline__ int perm_idx(int t, int c, int C) { return t * C + ((c >> 2) & 1) * (C >> 1) + (c >> 3) * 4 + (c & 3); }

struct AttnP {
  const __nv_bfloat16* X; const float* Kf; const float* Sc; const float* mb; const float* VM; const float* bm;
  const float* noise; const float* nstr; const float* bias; float gain, alpha;
  __nv_bfloat16* out; float* probs;
  // backward
  const __nv_bfloat16* dz; __nv_bfloat16* dX; float* dVM; float* R;
  long long HW; int C; int pix_per_cta; bool f16;
};

template <int VPL>
__device__ __forceinline__ void load_row(const __nv_bfloat16* p, int C, int lane, float (&v)[VPL][8], bool f16) {
#pragma unroll
  for (int q = 0; q < VPL; q++) {
    const int vi = q * 32 + lane;
    if (vi * 8 < C) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + vi);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; e++) { const float2 f = unpack16(w4[e], f16); v[q][e * 2] = f.x; v[q][e * 2 + 1] = f.y; }
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) v[q][e] = 0.f;
    }
  }
}
template <int VPL>
__device__ __forceinline__ void store_row(__nv_bfloat16* p, int C, int lane, const float (&v)[VPL][8], bool f16) {
#pragma unroll
  for (int q = 0; q < VPL; q++) {
    const int vi = q * 32 + lane;
    if (vi * 8 < C) {
      uint4 u;
      u.x = pack16(v[q][0], v[q][1], f16); u.y = pack16(v[q][2], v[q][3], f16); u.z = pack16(v[q][4], v[q][5], f16); u.w = pack16(v[q][6], v[q][7], f16);
      reinterpret_cast<uint4*>(p)[vi] = u;
    }
  }
}

// scores + softmax + normalisation factor for one pixel (all lanes end with the same A[16], rn)
template <int VPL>
__device__ __forceinline__ void pixel_probs(const float (&x)[VPL][8], const float* sK, int C, int lane, const float* sc_row, const float* mb_row,
                                            float (&A)[T16], float& rn) {
  float S[T16];
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < T16; t++) S[t] = 0.f;
#pragma unroll
  for (int q = 0; q < VPL; q++) {
    const int c0 = (q * 32 + lane) * 8;
    if (c0 < C) {
#pragma unroll
      for (int e = 0; e < 8; e++) ss = fmaf(x[q][e], x[q][e], ss);
#pragma unroll
      for (int t = 0; t < T16; t++) {
        const float4 k0 = *reinterpret_cast<const float4*>(sK + t * C + (c0 >> 1)), k1 = *reinterpret_cast<const float4*>(sK + t * C + (C >> 1) + (c0 >> 1));
        float a = S[t];
        a = fmaf(x[q][0], k0.x, a); a = fmaf(x[q][1], k0.y, a); a = fmaf(x[q][2], k0.z, a); a = fmaf(x[q][3], k0.w, a);
        a = fmaf(x[q][4], k1.x, a); a = fmaf(x[q][5], k1.y, a); a = fmaf(x[q][6], k1.z, a); a = fmaf(x[q][7], k1.w, a);
        S[t] = a;
      }
    }
  }
  ss = warp_sum(ss);
  float mx = -3.0e38f;
#pragma unroll
  for (int t = 0; t < T16; t++) { S[t] = warp_sum(S[t]) + sc_row[t] + mb_row[t]; mx = fmaxf(mx, S[t]); }
  float den = 0.f;
#pragma unroll
  for (int t = 0; t < T16; t++) { A[t] = __expf(S[t] - mx); den += A[t]; }
  const float inv = 1.f / den;
#pragma unroll
  for (int t = 0; t < T16; t++) A[t] *= inv;
  rn = rsqrtf(ss / (float)C + 1e-8f);
}

template <int VPL>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnP p) {
  extern __shared__ __align__(16) float smf[];
  const int C = p.C;
  float* sK = smf; float* sV = sK + T16 * C; float* sb = sV + T16 * C; float* sm1 = sb + C;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < T16 * C; i += blockDim.x) { const int pi = perm_idx(i / C, i % C, C); sK[pi] = p.Kf[i]; sV[pi] = p.VM[(long long)b * T16 * C + i]; }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sb[i] = p.bias ? p.bias[i] : 0.f; sm1[i] = 1.f + p.bm[i]; }
  __syncthreads();
  const float ns = (p.noise && p.nstr) ? *p.nstr : 0.f;
  const long long p0 = (long long)blockIdx.x * p.pix_per_cta;
  for (long long f = p0 + warp; f < p0 + p.pix_per_cta && f < p.HW; f += 8) {
    float x[VPL][8];
    load_row<VPL>(p.X + ((long long)b * p.HW + f) * C, C, lane, x, p.f16);
    float A[T16], rn;
    pixel_probs<VPL>(x, sK, C, lane, p.Sc + f * T16, p.mb + b * T16, A, rn);
    if (p.probs && lane < T16) p.probs[((long long)b * p.HW + f) * T16 + lane] = A[lane];
    const float nz = p.noise ? p.noise[f] * ns : 0.f;
    float o[VPL][8];
#pragma unroll
    for (int q = 0; q < VPL; q++) {
      const int c0 = (q * 32 + lane) * 8;
      if (c0 < C) {
        float ctl[8];
#pragma unroll
        for (int e = 0; e < 8; e++) ctl[e] = sm1[c0 + e];
#pragma unroll
        for (int t = 0; t < T16; t++) {
          const float4 v0 = *reinterpret_cast<const float4*>(sV + t * C + (c0 >> 1)), v1 = *reinterpret_cast<const float4*>(sV + t * C + (C >> 1) + (c0 >> 1));
          ctl[0] = fmaf(A[t], v0.x, ctl[0]); ctl[1] = fmaf(A[t], v0.y, ctl[1]); ctl[2] = fmaf(A[t], v0.z, ctl[2]); ctl[3] = fmaf(A[t], v0.w, ctl[3]);
          ctl[4] = fmaf(A[t], v1.x, ctl[4]); ctl[5] = fmaf(A[t], v1.y, ctl[5]); ctl[6] = fmaf(A[t], v1.z, ctl[6]); ctl[7] = fmaf(A[t], v1.w, ctl[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const float u = x[q][e] * rn * ctl[e] + nz + sb[c0 + e];
          o[q][e] = (u > 0.f ? u : u * p.alpha) * p.gain;
        }
      }
    }
    store_row<VPL>(p.out + ((long long)b * p.HW + f) * C, C, lane, o, p.f16);
  }
}

// Backward.  Phase 1 (warp per pixel, 32 pixels per round): recompute the forward, du = dz*gain*lrelu'(u),
// dctl = du*xn, dxn = du*(1+ctl), dA = dctl . VM, dS = A*(dA - sum A dA), dX = dS Kf + rn*dxn - X*rn^3/C*sum(dxn X);
// write dX, stash A and dctl (bf16) in shared memory, accumulate R[c] += dX*X.
// Phase 2 (whole CTA): dVM[t,c] += sum over the 32 pixels of A[p,t]*dctl[p,c], thread = (t, 32-channel group) register tile.
template <int VPL>
__global__ void __launch_bounds__(256, VPL == 1 ? 2 : 1) attn_bwd_kernel(AttnP p) {
  extern __shared__ __align__(16) float smf[];
  const int C = p.C;
  float* sK = smf; float* sV = sK + T16 * C; float* sb = sV + T16 * C; float* sm1 = sb + C;
  float* sR = sm1 + C;                       // [C]
  float* sA = sR + C;                        // [32][16]
  __nv_bfloat16* sD = reinterpret_cast<__nv_bfloat16*>(sA + 32 * T16);   // [32][C] bf16
  float* sDVM = reinterpret_cast<float*>(sD + 32 * C);                  // [16][C] fp32 dVM partial sums, each entry owned by one thread
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < T16 * C; i += blockDim.x) { const int pi = perm_idx(i / C, i % C, C); sK[pi] = p.Kf[i]; sV[pi] = p.VM[(long long)b * T16 * C + i]; }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sb[i] = p.bias ? p.bias[i] : 0.f; sm1[i] = 1.f + p.bm[i]; sR[i] = 0.f; }
  __syncthreads();
  const float ns = (p.noise && p.nstr) ? *p.nstr : 0.f;
  const long long p0 = (long long)blockIdx.x * p.pix_per_cta;
  long long pend = p0 + p.pix_per_cta; if (pend > p.HW) pend = p.HW;
  // phase-2 ownership: t2 = tid % 16, channel group g2 = tid / 16 covering channels [g2*CG, g2*CG+CG), CG = C/16
  const int t2 = threadIdx.x & 15, g2 = threadIdx.x >> 4, CG = C / 16;
  for (int j = 0; j < CG; j++) sDVM[t2 * C + g2 * CG + j] = 0.f;         // own entries only: no synchronisation needed
  float rloc[VPL][8];
#pragma unroll
  for (int q = 0; q < VPL; q++)
#pragma unroll
    for (int e = 0; e < 8; e++) rloc[q][e] = 0.f;

  for (long long base = p0; base < pend; base += 32) {
    // ---------------- phase 1: 4 pixels per warp
    for (int k = 0; k < 4; k++) {
      const int slot = warp * 4 + k;
      const long long f = base + slot;
      if (f < pend) {
        float x[VPL][8], g[VPL][8];
        load_row<VPL>(p.X + ((long long)b * p.HW + f) * C, C, lane, x, p.f16);
        load_row<VPL>(p.dz + ((long long)b * p.HW + f) * C, C, lane, g, false);
        float A[T16], rn;
        pixel_probs<VPL>(x, sK, C, lane, p.Sc + f * T16, p.mb + b * T16, A, rn);
        const float nz = p.noise ? p.noise[f] * ns : 0.f;
        float dA[T16];
#pragma unroll
        for (int t = 0; t < T16; t++) dA[t] = 0.f;
        float dxn[VPL][8];
        float sdot = 0.f;   // sum_c dxn * X
#pragma unroll
        for (int q = 0; q < VPL; q++) {
          const int c0 = (q * 32 + lane) * 8;
          if (c0 < C) {
            float ctl[8];
#pragma unroll
            for (int e = 0; e < 8; e++) ctl[e] = sm1[c0 + e];
#pragma unroll
            for (int t = 0; t < T16; t++) {
              const float4 v0 = *reinterpret_cast<const float4*>(sV + t * C + (c0 >> 1)), v1 = *reinterpret_cast<const float4*>(sV + t * C + (C >> 1) + (c0 >> 1));
              ctl[0] = fmaf(A[t], v0.x, ctl[0]); ctl[1] = fmaf(A[t], v0.y, ctl[1]); ctl[2] = fmaf(A[t], v0.z, ctl[2]); ctl[3] = fmaf(A[t], v0.w, ctl[3]);
              ctl[4] = fmaf(A[t], v1.x, ctl[4]); ctl[5] = fmaf(A[t], v1.y, ctl[5]); ctl[6] = fmaf(A[t], v1.z, ctl[6]); ctl[7] = fmaf(A[t], v1.w, ctl[7]);
            }
            float dctl[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const float xn = x[q][e] * rn;
              const float u = xn * ctl[e] + nz + sb[c0 + e];
              const float du = g[q][e] * p.gain * (u > 0.f ? 1.f : p.alpha);
              dctl[e] = du * xn;
              dxn[q][e] = du * ctl[e];
              sdot = fmaf(dxn[q][e], x[q][e], sdot);
            }
#pragma unroll
            for (int t = 0; t < T16; t++) {
              const float4 v0 = *reinterpret_cast<const float4*>(sV + t * C + (c0 >> 1)), v1 = *reinterpret_cast<const float4*>(sV + t * C + (C >> 1) + (c0 >> 1));
              float a = dA[t];
              a = fmaf(dctl[0], v0.x, a); a = fmaf(dctl[1], v0.y, a); a = fmaf(dctl[2], v0.z, a); a = fmaf(dctl[3], v0.w, a);
              a = fmaf(dctl[4], v1.x, a); a = fmaf(dctl[5], v1.y, a); a = fmaf(dctl[6], v1.z, a); a = fmaf(dctl[7], v1.w, a);
              dA[t] = a;
            }
            uint4 u4;
            u4.x = pack_bf16(dctl[0], dctl[1]); u4.y = pack_bf16(dctl[2], dctl[3]); u4.z = pack_bf16(dctl[4], dctl[5]); u4.w = pack_bf16(dctl[6], dctl[7]);
            *reinterpret_cast<uint4*>(sD + (long long)slot * C + c0) = u4;
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) dxn[q][e] = 0.f;
          }
        }
        sdot = warp_sum(sdot);
        float adot = 0.f;
#pragma unroll
        for (int t = 0; t < T16; t++) { dA[t] = warp_sum(dA[t]); adot = fmaf(A[t], dA[t], adot); }
        float dS[T16];
#pragma unroll
        for (int t = 0; t < T16; t++) dS[t] = A[t] * (dA[t] - adot);
        if (lane < T16) sA[slot * T16 + lane] = A[lane];
        const float k3 = rn * rn * rn * sdot / (float)C;
        float dx[VPL][8];
#pragma unroll
        for (int q = 0; q < VPL; q++) {
          const int c0 = (q * 32 + lane) * 8;
          if (c0 < C) {
            float a8[8];
#pragma unroll
            for (int e = 0; e < 8; e++) a8[e] = rn * dxn[q][e] - x[q][e] * k3;
#pragma unroll
            for (int t = 0; t < T16; t++) {
              const float4 k0 = *reinterpret_cast<const float4*>(sK + t * C + (c0 >> 1)), k1 = *reinterpret_cast<const float4*>(sK + t * C + (C >> 1) + (c0 >> 1));
              a8[0] = fmaf(dS[t], k0.x, a8[0]); a8[1] = fmaf(dS[t], k0.y, a8[1]); a8[2] = fmaf(dS[t], k0.z, a8[2]); a8[3] = fmaf(dS[t], k0.w, a8[3]);
              a8[4] = fmaf(dS[t], k1.x, a8[4]); a8[5] = fmaf(dS[t], k1.y, a8[5]); a8[6] = fmaf(dS[t], k1.z, a8[6]); a8[7] = fmaf(dS[t], k1.w, a8[7]);
            }
#pragma unroll
            for (int e = 0; e < 8; e++) { dx[q][e] = a8[e]; rloc[q][e] = fmaf(a8[e], x[q][e], rloc[q][e]); }
          }
        }
        store_row<VPL>(p.dX + ((long long)b * p.HW + f) * C, C, lane, dx, false);
      } else {
        // keep phase 2 uniform: empty slots contribute zeros
        if (lane < T16) sA[slot * T16 + lane] = 0.f;
      }
    }
    __syncthreads();
    // ---------------- phase 2: dVM partial sums, thread (t2, g2) owns channels g2*CG .. +CG (CG <= 32)
    {
      const int npix = (int)((pend - base) < 32 ? (pend - base) : 32);
      float acc[32];                                     // live only in this phase (keeps phase 1 under 128 registers)
      float* own = sDVM + t2 * C + g2 * CG;
#pragma unroll
      for (int j = 0; j < 32; j++) acc[j] = (j < CG) ? own[j] : 0.f;
      for (int pp = 0; pp < npix; pp++) {
        const float a = sA[pp * T16 + t2];
        const __nv_bfloat16* drow = sD + (long long)pp * C + g2 * CG;
        if (CG >= 8) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < CG) {
              const uint4 u = *reinterpret_cast<const uint4*>(drow + j);
              const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; e++) { const float2 f2 = unpack_bf16(w4[e]); acc[j + e * 2] = fmaf(a, f2.x, acc[j + e * 2]); acc[j + e * 2 + 1] = fmaf(a, f2.y, acc[j + e * 2 + 1]); }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) if (j < CG) acc[j] = fmaf(a, __bfloat162float(drow[j]), acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; j++) if (j < CG) own[j] = acc[j];
    }
    __syncthreads();
  }
  // flush
  for (int j = 0; j < CG; j++) atomicAdd(&p.dVM[((long long)b * T16 + t2) * C + g2 * CG + j], sDVM[t2 * C + g2 * CG + j]);
#pragma unroll
  for (int q = 0; q < VPL; q++) {
    const int c0 = (q * 32 + lane) * 8;
    if (c0 < C) {
#pragma unroll
      for (int e = 0; e < 8; e++) atomicAdd(&sR[c0 + e], rloc[q][e]);
    }
  }
  __syncthreads();
  if (p.R) for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&p.R[(long long)b * C + i], sR[i]);
}

static int attn_smem_fwd(int C) { return (2 * T16 * C + 2 * C) * 4; }
static int attn_smem_bwd(int C) { return (2 * T16 * C + 3 * C + 32 * T16) * 4 + 32 * C * 2 + T16 * C * 4; }

}  // namespace mgf

using namespace mgf;

static int attn_check(int C, const char* who) {
  if (C % 128 != 0 && !(C == 32 || C == 64)) MGF_FAIL(MGF_E_SHAPE, "%s: C=%d must be 32, 64 or a multiple of 128", who, C);
  if (C > 512) MGF_FAIL(MGF_E_SHAPE, "%s: C=%d > 512", who, C);
  return 0;
}

extern "C" int mgf_attn_fwd_simt(const void* X, const float* Kf, const float* Sc, const float* maskbias, const float* VM, const float* bm,
                            const float* noise, const float* nstr, const float* bias, float gain, float alpha,
                            void* out, float* probs, int B, int64_t HW, int C, void* stream) {
  if (!X || !Kf || !Sc || !maskbias || !VM || !bm || !out) MGF_FAIL(MGF_E_BADARG, "attn_fwd: null tensor");
  if (int e = attn_check(C, "attn_fwd")) return e;
  AttnP p{}; p.X = (const __nv_bfloat16*)X; p.Kf = Kf; p.Sc = Sc; p.mb = maskbias; p.VM = VM; p.bm = bm; p.noise = noise; p.nstr = nstr; p.bias = bias;
  p.gain = gain; p.alpha = alpha; p.out = (__nv_bfloat16*)out; p.probs = probs; p.HW = HW; p.C = C; p.f16 = fwd_f16();
  long long ppc = (HW * B + (long long)num_sms() * 4 - 1) / ((long long)num_sms() * 4);
  if (ppc < 8) ppc = 8;
  ppc = (ppc + 7) / 8 * 8; p.pix_per_cta = (int)ppc;
  dim3 grid((unsigned)((HW + ppc - 1) / ppc), B);
  const int smem = attn_smem_fwd(C);
  cudaStream_t st = (cudaStream_t)stream;
  static bool cfg_done = false;
  if (!cfg_done) {   // opt in once to > 48 KB dynamic shared memory (largest case C = 512), outside any stream capture
    cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_fwd(512));
    cudaFuncSetAttribute(attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_fwd(256));
    cfg_done = true;
  }
  if (C > 256) attn_fwd_kernel<2><<<grid, 256, smem, st>>>(p);
  else attn_fwd_kernel<1><<<grid, 256, smem, st>>>(p);
  MGF_CHECK_LAUNCH("attn_fwd");
  return 0;
}

extern "C" int mgf_attn_bwd_simt(const void* X, const void* dz, const float* Kf, const float* Sc, const float* maskbias, const float* VM, const float* bm,
                            const float* noise, const float* nstr, const float* bias, float gain, float alpha,
                            void* dX, float* dVM, float* R, int B, int64_t HW, int C, void* stream) {
  if (!X || !dz || !Kf || !Sc || !maskbias || !VM || !bm || !dX || !dVM) MGF_FAIL(MGF_E_BADARG, "attn_bwd: null tensor");
  if (int e = attn_check(C, "attn_bwd")) return e;
  if (C % 16) MGF_FAIL(MGF_E_SHAPE, "attn_bwd: C must be a multiple of 16");
  AttnP p{}; p.X = (const __nv_bfloat16*)X; p.dz = (const __nv_bfloat16*)dz; p.Kf = Kf; p.Sc = Sc; p.mb = maskbias; p.VM = VM; p.bm = bm;
  p.noise = noise; p.nstr = nstr; p.bias = bias; p.gain = gain; p.alpha = alpha; p.dX = (__nv_bfloat16*)dX; p.dVM = dVM; p.R = R; p.HW = HW; p.C = C; p.f16 = fwd_f16();
  long long ppc = (HW * B + (long long)num_sms() * 2 - 1) / ((long long)num_sms() * 2);
  if (ppc < 32) ppc = 32;
  ppc = (ppc + 31) / 32 * 32; p.pix_per_cta = (int)ppc;
  dim3 grid((unsigned)((HW + ppc - 1) / ppc), B);
  const int smem = attn_smem_bwd(C);
  cudaStream_t st = (cudaStream_t)stream;
  static bool cfg_done = false;
  if (!cfg_done) {
    cudaFuncSetAttribute(attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bwd(512));
    cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bwd(256));
    cfg_done = true;
  }
  if (C > 256) attn_bwd_kernel<2><<<grid, 256, smem, st>>>(p);
  else attn_bwd_kernel<1><<<grid, 256, smem, st>>>(p);
  MGF_CHECK_LAUNCH("attn_bwd");
  return 0;
}
