// attention_mma.cu -- tensor-core version of the fused duplex-attention layer (forward and backward).
//
// Same algebra as attention.cu (reference training/networks.py:748-822 TransformerLayer.forward, default configuration, plus the
// layer tail :1036-1040), but the four contractions of a pixel tile run on the tensor cores with mma.sync.m16n8k16:
//   forward   S[16px,16] = X[16px,C] Kf^T          (K = C;  Kf split into hi + lo 16-bit parts -> fp32-accurate coefficients)
//             ctl[16px,C] = A[16px,16] VM[16,C]    (K = 16; probabilities re-used straight from the score accumulators)
//   backward  dA[16px,16] = dctl[16px,C] VM^T      (K = C)
//             dX_att[16px,C] = dS[16px,16] Kf      (K = 16)
//             dVM[16,C] += A^T[16,px] dctl[px,C]   (K = pixels, CTA-wide: operands staged in shared memory, ldmatrix.trans)
// The contraction width of the layer is only 16 latents, so tcgen05 (M = 128 tiles, TMEM round trip) buys nothing here: the layer
// is HBM-bound once the 2 x 16 x C multiply-adds per element leave the FMA pipe (the SIMT version spent its time re-reading 32 KB of
// coefficients from shared memory per pixel).  A warp owns 16 pixels; lane (g = lane/4, t = lane%4) owns rows g and g+8.
//
// Channel permutation: the MMA K index of the score product and the N index of the ctl product are both free to permute, so a lane
// loads / stores whole 16-byte vectors: vector v = j*4 + t of a row holds channels [8v, 8v+8); its four 32-bit words feed
// (a0|a1, a2|a3) of two consecutive k-steps, and the matching B fragments are the same vector of row n of the coefficient table.
#include "common.cuh"

namespace mgf {
namespace {

constexpr int NT = 16;   // latents (keys)

template <bool F16>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <bool F16> __device__ __forceinline__ uint32_t pk(float a, float b) {
  return F16 ? pack_f16_sat(a, b) : pack_bf16(a, b);
}
template <bool F16> __device__ __forceinline__ float2 upk(uint32_t u) {
  if (F16) { __half2 h = *reinterpret_cast<__half2*>(&u); return __half22float2(h); }
  return unpack_bf16(u);
}
template <bool F16> __device__ __forceinline__ uint16_t cv16(float v) {
  if (F16) { __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h);
}
template <bool F16> __device__ __forceinline__ float cvf(uint16_t u) {
  if (F16) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }
  __nv_bfloat16 h = *reinterpret_cast<__nv_bfloat16*>(&u); return __bfloat162float(h);
}
__device__ __forceinline__ float quad_sum(float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); return v; }
__device__ __forceinline__ float quad_max(float v) { v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2)); return v; }

struct AttnMP {
  const void* X; const float* Kf; const float* Sc; const float* mb; const float* VM; const float* bm;
  const float* noise; const float* nstr; const float* bias; float gain, alpha;
  void* out; float* probs;
  const __nv_bfloat16* dz; __nv_bfloat16* dX; float* dVM; float* R;
  const void* tabK; const void* tabV;   // pre-built coefficient tables (mgf_attn_tables): shared-memory images, copied instead of rebuilt per CTA
  const float* dmask;          // attention dropout (training mode): [B,HW,16] fp32 keep-mask * scale, multiplies the probabilities after the softmax (nullptr: eval)
  long long HW; int C; int pix_per_cta; long long nbs;      // nbs: elements between per-sample noise planes (0 = shared plane)
};

// Row-major 16-bit coefficient table [16][C] with a padded row (2C + 64 bytes): the 16-byte reads of a quarter-warp (two rows g,
// four vectors t) fall into 8 distinct bank groups.
__host__ __device__ inline int krow(int C) { return C + 32; }
// N-permuted table for the K = 16 products: entry (ntile = j*4+m, lane) holds the coefficients of latents {2t, 2t+1, 2t+8, 2t+9}
// for channel chan(j, m, g) = (j*4 + g/2)*8 + 2m + g%2 (so that output column 2t+e of n-tile m is channel (j*4+t)*8 + 2m + e).
__device__ __forceinline__ int perm_chan(int nt, int g) { return ((nt >> 2) * 4 + (g >> 1)) * 8 + 2 * (nt & 3) + (g & 1); }

// Table fills are the whole cost of the low-resolution layers (a 4x4 .. 32x32 grid is a handful of warp tiles per CTA): float4 loads,
// 8-byte shared stores and 4-way unrolling keep many loads in flight instead of one dependent load per iteration.
template <bool F16>
__device__ __forceinline__ void fill_rowtable(uint16_t* hi, uint16_t* lo, const float* __restrict__ src, int C) {
  const int KS = krow(C), C4 = C >> 2;
#pragma unroll 4
  for (int i = threadIdx.x; i < NT * C4; i += blockDim.x) {
    const int r = i / C4, c = (i - r * C4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    const uint16_t h0 = cv16<F16>(v.x), h1 = cv16<F16>(v.y), h2 = cv16<F16>(v.z), h3 = cv16<F16>(v.w);
    *reinterpret_cast<uint2*>(hi + r * KS + c) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
    if (lo) {
      const uint16_t l0 = cv16<F16>(v.x - cvf<F16>(h0)), l1 = cv16<F16>(v.y - cvf<F16>(h1)), l2 = cv16<F16>(v.z - cvf<F16>(h2)), l3 = cv16<F16>(v.w - cvf<F16>(h3));
      *reinterpret_cast<uint2*>(lo + r * KS + c) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
    }
  }
}
template <bool F16>
__device__ __forceinline__ void fill_permtable(uint2* dst, const float* __restrict__ src, int C) {      // src [16][C] fp32
#pragma unroll 4
  for (int i = threadIdx.x; i < (C / 8) * 32; i += blockDim.x) {
    const int nt = i >> 5, ln = i & 31, g = ln >> 2, t = ln & 3;
    const int ch = perm_chan(nt, g);
    const float a = __ldg(src + (2 * t) * C + ch), b = __ldg(src + (2 * t + 1) * C + ch), c = __ldg(src + (2 * t + 8) * C + ch), d = __ldg(src + (2 * t + 9) * C + ch);
    uint2 u;
    u.x = pk<F16>(a, b);
    u.y = pk<F16>(c, d);
    dst[i] = u;
  }
}

// ---- pre-built tables.  Building the 16-bit coefficient tables from the fp32 constants cost every CTA of every launch ~40 us (strided
// scalar loads + conversions: the whole run time of the 4^2 .. 16^2 layers and a third of the 64^2 ones).  mgf_attn_tables writes the
// same shared-memory images to global memory once -- Kf tables when the weights are folded, VM tables per step on the side stream --
// and the kernels copy them with 16-byte vectors.
//   tabK: [Khi rows][Klo rows][Kp perm (bf16)]      tabV (per sample): [Vr rows (bf16)][Vp perm (fp16)]
__host__ __device__ inline size_t rows_bytes(int C) { return (size_t)NT * krow(C) * 2; }
__host__ __device__ inline size_t perm_bytes(int C) { return (size_t)(C / 8) * 32 * 8; }
__host__ __device__ inline size_t tabk_bytes(int C) { return 2 * rows_bytes(C) + perm_bytes(C); }
__host__ __device__ inline size_t tabv_bytes(int C) { return rows_bytes(C) + perm_bytes(C); }
// Table copies are asynchronous bulk copies (cp.async.bulk global -> shared, completion counted on an mbarrier): one thread posts them all,
// they run beside the rest of the CTA prologue, and every thread then waits on the barrier.  (A per-thread 16-byte copy loop cost ~1 us
// of L2 latency per 4 KB trip -- 13 trips for the forward tables, 21 for the backward ones: most of the run time of the 4^2 .. 32^2 layers.)
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sm_u32(dst)), "l"(src), "r"(bytes), "r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void tbar_wait(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(sm_u32(bar)) : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(256) attn_tables_kernel(const float* __restrict__ Kf, const float* __restrict__ VM, unsigned char* tabK, unsigned char* tabV, int C) {
  // block 0 .. B-1: the VM tables of sample blockIdx.x (if VM); the last block: the Kf tables (if Kf).  The fill_* helpers index their
  // destination exactly as they index shared memory, so the images are identical to what the kernels used to build in place.
  if (VM && (int)blockIdx.x < (int)gridDim.x - (Kf ? 1 : 0)) {
    unsigned char* tv = tabV + (size_t)blockIdx.x * tabv_bytes(C);
    const float* vm = VM + (size_t)blockIdx.x * NT * C;
    fill_rowtable<false>(reinterpret_cast<uint16_t*>(tv), nullptr, vm, C);
    fill_permtable<true>(reinterpret_cast<uint2*>(tv + rows_bytes(C)), vm, C);
  } else if (Kf) {
    fill_rowtable<F16>(reinterpret_cast<uint16_t*>(tabK), reinterpret_cast<uint16_t*>(tabK + rows_bytes(C)), Kf, C);
    fill_permtable<false>(reinterpret_cast<uint2*>(tabK + 2 * rows_bytes(C)), Kf, C);
  }
}

// scores + softmax for one 16-pixel tile.  On return P[0..1] = A[row g][2t, 2t+1], P[2..3] = A[row g+8][2t, 2t+1],
// P[4..5] = A[row g][8+2t, 9+2t], P[6..7] = A[row g+8][8+2t, 9+2t]; rn0 / rn1 = rsqrt(mean_c x^2 + 1e-8) of the two rows.
template <bool F16, int C32>
__device__ __forceinline__ void tile_probs(const uint4* __restrict__ x0, const uint4* __restrict__ x1, const uint16_t* sKhi, const uint16_t* sKlo,
                                           int C, int g, int t, const float* sc0, const float* sc1, const float* mbr,
                                           float (&P)[8], float& rn0, float& rn1) {
  const int KS = krow(C);
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  float ss0 = 0.f, ss1 = 0.f;
  constexpr int UP = C32 < 4 ? C32 : 4;      // partial unrolling: four independent vector-load pairs in flight, a loop body the instruction cache keeps
#pragma unroll (UP)
  for (int j = 0; j < C32; j++) {
    const uint4 a = __ldg(x0 + j * 4 + t), b = __ldg(x1 + j * 4 + t);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float2 fa = upk<F16>(aw[e]), fb = upk<F16>(bw[e]);
      ss0 = fmaf(fa.x, fa.x, ss0); ss0 = fmaf(fa.y, fa.y, ss0);
      ss1 = fmaf(fb.x, fb.x, ss1); ss1 = fmaf(fb.y, fb.y, ss1);
    }
    const int off = (j * 4 + t) * 8;
    const uint4 h0 = *reinterpret_cast<const uint4*>(sKhi + g * KS + off), h1 = *reinterpret_cast<const uint4*>(sKhi + (g + 8) * KS + off);
    const uint4 l0 = *reinterpret_cast<const uint4*>(sKlo + g * KS + off), l1 = *reinterpret_cast<const uint4*>(sKlo + (g + 8) * KS + off);
    mma16816<F16>(s0, a.x, b.x, a.y, b.y, h0.x, h0.y); mma16816<F16>(s0, a.z, b.z, a.w, b.w, h0.z, h0.w);
    mma16816<F16>(s1, a.x, b.x, a.y, b.y, h1.x, h1.y); mma16816<F16>(s1, a.z, b.z, a.w, b.w, h1.z, h1.w);
    mma16816<F16>(s0, a.x, b.x, a.y, b.y, l0.x, l0.y); mma16816<F16>(s0, a.z, b.z, a.w, b.w, l0.z, l0.w);
    mma16816<F16>(s1, a.x, b.x, a.y, b.y, l1.x, l1.y); mma16816<F16>(s1, a.z, b.z, a.w, b.w, l1.z, l1.w);
  }
  ss0 = quad_sum(ss0); ss1 = quad_sum(ss1);
  rn0 = rsqrtf(ss0 / (float)C + 1e-8f); rn1 = rsqrtf(ss1 / (float)C + 1e-8f);
  const float2 ca0 = *reinterpret_cast<const float2*>(sc0 + 2 * t), ca1 = *reinterpret_cast<const float2*>(sc0 + 8 + 2 * t);
  const float2 cb0 = *reinterpret_cast<const float2*>(sc1 + 2 * t), cb1 = *reinterpret_cast<const float2*>(sc1 + 8 + 2 * t);
  const float2 m0 = *reinterpret_cast<const float2*>(mbr + 2 * t), m1 = *reinterpret_cast<const float2*>(mbr + 8 + 2 * t);
  P[0] = s0[0] + ca0.x + m0.x; P[1] = s0[1] + ca0.y + m0.y; P[4] = s1[0] + ca1.x + m1.x; P[5] = s1[1] + ca1.y + m1.y;
  P[2] = s0[2] + cb0.x + m0.x; P[3] = s0[3] + cb0.y + m0.y; P[6] = s1[2] + cb1.x + m1.x; P[7] = s1[3] + cb1.y + m1.y;
  float mx0 = quad_max(fmaxf(fmaxf(P[0], P[1]), fmaxf(P[4], P[5])));
  float mx1 = quad_max(fmaxf(fmaxf(P[2], P[3]), fmaxf(P[6], P[7])));
  P[0] = __expf(P[0] - mx0); P[1] = __expf(P[1] - mx0); P[4] = __expf(P[4] - mx0); P[5] = __expf(P[5] - mx0);
  P[2] = __expf(P[2] - mx1); P[3] = __expf(P[3] - mx1); P[6] = __expf(P[6] - mx1); P[7] = __expf(P[7] - mx1);
  const float i0 = 1.f / quad_sum(P[0] + P[1] + P[4] + P[5]), i1 = 1.f / quad_sum(P[2] + P[3] + P[6] + P[7]);
  P[0] *= i0; P[1] *= i0; P[4] *= i0; P[5] *= i0; P[2] *= i1; P[3] *= i1; P[6] *= i1; P[7] *= i1;
}

// attention dropout (reference networks.py:505-513: probs = dropout(probs) over cells, then over whole 'to' columns; both masks and their
// 1/(1-p) scales arrive pre-multiplied in dmask): values of the two rows in the layout of P
__device__ __forceinline__ void load_dmask(const float* dm0, const float* dm1, int t, float (&M)[8]) {
  const float2 a0 = *reinterpret_cast<const float2*>(dm0 + 2 * t), a1 = *reinterpret_cast<const float2*>(dm0 + 8 + 2 * t);
  const float2 b0 = *reinterpret_cast<const float2*>(dm1 + 2 * t), b1 = *reinterpret_cast<const float2*>(dm1 + 8 + 2 * t);
  M[0] = a0.x; M[1] = a0.y; M[4] = a1.x; M[5] = a1.y; M[2] = b0.x; M[3] = b0.y; M[6] = b1.x; M[7] = b1.y;
}

// Forward shared memory: sKhi, sKlo [16][C+32] 16-bit | sVp uint2 [C/8][32] (fp16) | sm1 [C] fp32 (1 + bm) | sb [C] fp32 (bias)
static int fwd_smem(int C) { return 2 * NT * krow(C) * 2 + (C / 8) * 32 * 8 + 2 * C * 4; }

template <bool F16, int C32>
__global__ void __launch_bounds__(256, 2) attn_fwd_mma_kernel(AttnMP p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int C = p.C, KS = krow(C);
  uint16_t* sKhi = reinterpret_cast<uint16_t*>(smraw);
  uint16_t* sKlo = sKhi + NT * KS;
  uint2* sVp = reinterpret_cast<uint2*>(sKlo + NT * KS);
  float* sm1 = reinterpret_cast<float*>(sVp + (C / 8) * 32);
  float* sb = sm1 + C;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  __shared__ __align__(8) uint64_t tbar;
  const bool bulk = p.tabK || p.tabV;
  if (bulk) {
    if (threadIdx.x == 0) {
      tbar_init(&tbar);
      tbar_expect(&tbar, (uint32_t)((p.tabK ? 2 * rows_bytes(C) : 0) + (p.tabV ? perm_bytes(C) : 0)));
      if (p.tabK) bulk_copy(sKhi, p.tabK, (uint32_t)(2 * rows_bytes(C)), &tbar);
      if (p.tabV) bulk_copy(sVp, reinterpret_cast<const unsigned char*>(p.tabV) + (size_t)b * tabv_bytes(C) + rows_bytes(C), (uint32_t)perm_bytes(C), &tbar);
    }
  }
  if (!p.tabK) fill_rowtable<F16>(sKhi, sKlo, p.Kf, C);
  if (!p.tabV) fill_permtable<true>(sVp, p.VM + (long long)b * NT * C, C);
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sb[i] = p.bias ? p.bias[i] : 0.f; sm1[i] = 1.f + p.bm[i]; }
  __syncthreads();                     // the barrier word is initialised (and the fill_* / sb / sm1 writes are visible) past this point
  if (bulk) tbar_wait(&tbar);
  const float ns = (p.noise && p.nstr) ? *p.nstr : 0.f;
  const long long p0 = (long long)blockIdx.x * p.pix_per_cta;
  long long pend = p0 + p.pix_per_cta; if (pend > p.HW) pend = p.HW;
  const float* mbr = p.mb + b * NT;
  for (long long f0 = p0 + warp * 16; f0 < pend; f0 += 8 * 16) {
    const long long r0 = f0 + g, r1 = f0 + g + 8;
    const bool v0 = r0 < pend, v1 = r1 < pend;
    const long long q0 = v0 ? r0 : pend - 1, q1 = v1 ? r1 : pend - 1;            // clamp: masked rows compute on a valid row, never store
    const uint4* x0 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.X) + ((long long)b * p.HW + q0) * C);
    const uint4* x1 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.X) + ((long long)b * p.HW + q1) * C);
    float P[8], rn0, rn1;
    tile_probs<F16, C32>(x0, x1, sKhi, sKlo, C, g, t, p.Sc + q0 * NT, p.Sc + q1 * NT, mbr, P, rn0, rn1);
    if (p.dmask) {
      float M[8];
      load_dmask(p.dmask + ((long long)b * p.HW + q0) * NT, p.dmask + ((long long)b * p.HW + q1) * NT, t, M);
#pragma unroll
      for (int i = 0; i < 8; i++) P[i] *= M[i];
    }
    if (p.probs) {
      float* pr0 = p.probs + ((long long)b * p.HW + q0) * NT; float* pr1 = p.probs + ((long long)b * p.HW + q1) * NT;
      if (v0) { *reinterpret_cast<float2*>(pr0 + 2 * t) = make_float2(P[0], P[1]); *reinterpret_cast<float2*>(pr0 + 8 + 2 * t) = make_float2(P[4], P[5]); }
      if (v1) { *reinterpret_cast<float2*>(pr1 + 2 * t) = make_float2(P[2], P[3]); *reinterpret_cast<float2*>(pr1 + 8 + 2 * t) = make_float2(P[6], P[7]); }
    }
    const uint32_t pa0 = pk<true>(P[0], P[1]), pa1 = pk<true>(P[2], P[3]), pa2 = pk<true>(P[4], P[5]), pa3 = pk<true>(P[6], P[7]);
    const float nz0 = p.noise ? p.noise[b * p.nbs + q0] * ns : 0.f, nz1 = p.noise ? p.noise[b * p.nbs + q1] * ns : 0.f;
    uint4* o0 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.HW + q0) * C);
    uint4* o1 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.HW + q1) * C);
#pragma unroll
    for (int j = 0; j < C32; j++) {
      const uint4 a = __ldg(x0 + j * 4 + t), bq = __ldg(x1 + j * 4 + t);           // second touch of the tile: L1 hit
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
      const int ch = (j * 4 + t) * 8;
      const float4 m1a = *reinterpret_cast<const float4*>(sm1 + ch), m1b = *reinterpret_cast<const float4*>(sm1 + ch + 4);
      const float4 ba = *reinterpret_cast<const float4*>(sb + ch), bb = *reinterpret_cast<const float4*>(sb + ch + 4);
      const float m1v[8] = {m1a.x, m1a.y, m1a.z, m1a.w, m1b.x, m1b.y, m1b.z, m1b.w};
      const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
      uint32_t ow0[4], ow1[4];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        float acc[4] = {m1v[2 * m], m1v[2 * m + 1], m1v[2 * m], m1v[2 * m + 1]};        // 1 + bm, then += A VM
        const uint2 vb = sVp[(j * 4 + m) * 32 + lane];
        mma16816<true>(acc, pa0, pa1, pa2, pa3, vb.x, vb.y);
        const float2 xa = upk<F16>(aw[m]), xb = upk<F16>(bw[m]);
        float u0 = xa.x * rn0 * acc[0] + nz0 + bv[2 * m], u1 = xa.y * rn0 * acc[1] + nz0 + bv[2 * m + 1];
        float u2 = xb.x * rn1 * acc[2] + nz1 + bv[2 * m], u3 = xb.y * rn1 * acc[3] + nz1 + bv[2 * m + 1];
        u0 = (u0 > 0.f ? u0 : u0 * p.alpha) * p.gain; u1 = (u1 > 0.f ? u1 : u1 * p.alpha) * p.gain;
        u2 = (u2 > 0.f ? u2 : u2 * p.alpha) * p.gain; u3 = (u3 > 0.f ? u3 : u3 * p.alpha) * p.gain;
        ow0[m] = pk<F16>(u0, u1); ow1[m] = pk<F16>(u2, u3);
      }
      if (v0) o0[j * 4 + t] = make_uint4(ow0[0], ow0[1], ow0[2], ow0[3]);
      if (v1) o1[j * 4 + t] = make_uint4(ow1[0], ow1[1], ow1[2], ow1[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------ backward
// Shared memory: sKhi, sKlo [16][C+32] (forward dtype) | sVr [16][C+32] bf16 (VM rows, for dA) | sVp uint2 [C/8][32] fp16 (ctl) |
// sKp uint2 [C/8][32] bf16 (Kf, N-permuted, for dS Kf) | sm1, sb, sR [C] fp32 |
// sP [128][24] bf16 (probabilities of the round, 48-byte rows) | sD [128][C/NH + 8] bf16 (dctl of the round, one channel group)
// NH = number of channel groups the dVM phase is split into (2 from 256 channels: keeps the staging buffer at 34 / 67 KB).
constexpr int SPS = 24;
__host__ __device__ inline int nh_of(int C) { return C >= 256 ? 2 : 1; }
__host__ __device__ inline int drow(int C) { return C / nh_of(C) + 8; }
static int bwd_smem(int C) { return 3 * NT * krow(C) * 2 + 2 * (C / 8) * 32 * 8 + 3 * C * 4 + 128 * SPS * 2 + 128 * drow(C) * 2; }

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* ptr) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <bool F16, int C32>
// up to 256 channels two CTAs share an SM (88 KB of shared memory each at C = 256, 128 registers): the kernel is latency-bound, a second CTA doubles
// the loads in flight
__global__ void __launch_bounds__(256, (C32 <= 8) ? 2 : 1) attn_bwd_mma_kernel(AttnMP p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int NH = (C32 >= 8) ? 2 : 1;
  constexpr int JH = C32 / NH;                             // 32-channel chunks per channel group
  constexpr int NTG = JH * 4;                              // n-tiles (8 channels) per channel group
  constexpr int NTW = (NTG + 7) / 8;                       // n-tiles per warp in the dVM phase (1, 1, 2, 4, 3, 4 for C32 = 1..16)
  const int C = p.C, KS = krow(C), DS = drow(C);
  uint16_t* sKhi = reinterpret_cast<uint16_t*>(smraw);
  uint16_t* sKlo = sKhi + NT * KS;
  uint16_t* sVr = sKlo + NT * KS;
  uint2* sVp = reinterpret_cast<uint2*>(sVr + NT * KS);
  uint2* sKp = sVp + (C / 8) * 32;
  float* sm1 = reinterpret_cast<float*>(sKp + (C / 8) * 32);
  float* sb = sm1 + C;
  float* sR = sb + C;
  uint16_t* sP = reinterpret_cast<uint16_t*>(sR + C);
  uint16_t* sD = sP + 128 * SPS;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const float* VMb = p.VM + (long long)b * NT * C;
  __shared__ __align__(8) uint64_t tbar;
  const bool bulk = p.tabK || p.tabV;
  if (bulk && threadIdx.x == 0) {
    tbar_init(&tbar);
    tbar_expect(&tbar, (uint32_t)((p.tabK ? tabk_bytes(C) : 0) + (p.tabV ? tabv_bytes(C) : 0)));
    if (p.tabK) {
      bulk_copy(sKhi, p.tabK, (uint32_t)(2 * rows_bytes(C)), &tbar);
      bulk_copy(sKp, reinterpret_cast<const unsigned char*>(p.tabK) + 2 * rows_bytes(C), (uint32_t)perm_bytes(C), &tbar);
    }
    if (p.tabV) bulk_copy(sVr, reinterpret_cast<const unsigned char*>(p.tabV) + (size_t)b * tabv_bytes(C), (uint32_t)tabv_bytes(C), &tbar);   // sVr | sVp are contiguous, like the image
  }
  if (!p.tabK) {
    fill_rowtable<F16>(sKhi, sKlo, p.Kf, C);
    fill_permtable<false>(sKp, p.Kf, C);
  }
  if (!p.tabV) {
    fill_rowtable<false>(sVr, nullptr, VMb, C);
    fill_permtable<true>(sVp, VMb, C);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sb[i] = p.bias ? p.bias[i] : 0.f; sm1[i] = 1.f + p.bm[i]; sR[i] = 0.f; }
  __syncthreads();                     // barrier word initialised; fill_* / sb / sm1 / sR visible
  if (bulk) tbar_wait(&tbar);
  const float ns = (p.noise && p.nstr) ? *p.nstr : 0.f;
  const long long p0 = (long long)blockIdx.x * p.pix_per_cta;
  long long pend = p0 + p.pix_per_cta; if (pend > p.HW) pend = p.HW;
  const float* mbr = p.mb + b * NT;
  float dvm[NH][NTW][4];
#pragma unroll
  for (int h = 0; h < NH; h++)
#pragma unroll
    for (int i = 0; i < NTW; i++) { dvm[h][i][0] = dvm[h][i][1] = dvm[h][i][2] = dvm[h][i][3] = 0.f; }
  // The channel loops below are unrolled by UJ only: fully unrolled, the C = 512 kernel was 12 K instructions (190 KB) of straight-line code
  // that every round streamed through the instruction cache -- `no_instruction` was its top stall reason (profiles/r01 attention table).
  constexpr int UJ = C32 < 2 ? C32 : 2;

  for (long long base = p0; base < pend; base += 128) {
    const long long f0 = base + warp * 16;
    const bool active = f0 < pend;                        // warp-uniform
    uint16_t* myP = sP + warp * 16 * SPS;
    uint16_t* myD = sD + (long long)warp * 16 * DS;
    const long long r0 = f0 + g, r1 = f0 + g + 8;
    const bool v0 = r0 < pend, v1 = r1 < pend;
    const long long q0 = v0 ? r0 : pend - 1, q1 = v1 ? r1 : pend - 1;   // clamp: masked rows compute on a valid row with zero gradient
    const uint4* x0 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.X) + ((long long)b * p.HW + q0) * C);
    const uint4* x1 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.X) + ((long long)b * p.HW + q1) * C);
    const uint4* g0 = reinterpret_cast<const uint4*>(p.dz + ((long long)b * p.HW + q0) * C);
    const uint4* g1 = reinterpret_cast<const uint4*>(p.dz + ((long long)b * p.HW + q1) * C);
    float P[8], M[8], rn0 = 0.f, rn1 = 0.f;      // P: softmax probabilities; M: dropout mask * scale (ones in eval mode); A = P * M feeds ctl and dVM
#pragma unroll
    for (int i = 0; i < 8; i++) M[i] = 1.f;
    uint32_t pa0 = 0, pa1 = 0, pa2 = 0, pa3 = 0;
    float nz0 = 0.f, nz1 = 0.f;
    const float gm0 = v0 ? p.gain : 0.f, gm1 = v1 ? p.gain : 0.f;
    float dA0[4] = {0.f, 0.f, 0.f, 0.f}, dA1[4] = {0.f, 0.f, 0.f, 0.f};
    float sd0 = 0.f, sd1 = 0.f;
    if (active) {
      tile_probs<F16, C32>(x0, x1, sKhi, sKlo, C, g, t, p.Sc + q0 * NT, p.Sc + q1 * NT, mbr, P, rn0, rn1);
      if (p.dmask) load_dmask(p.dmask + ((long long)b * p.HW + q0) * NT, p.dmask + ((long long)b * p.HW + q1) * NT, t, M);
      float A[8];
#pragma unroll
      for (int i = 0; i < 8; i++) A[i] = P[i] * M[i];
      pa0 = pk<true>(A[0], A[1]); pa1 = pk<true>(A[2], A[3]); pa2 = pk<true>(A[4], A[5]); pa3 = pk<true>(A[6], A[7]);
      if (p.noise) { nz0 = p.noise[b * p.nbs + q0] * ns; nz1 = p.noise[b * p.nbs + q1] * ns; }
      // (dropped) probabilities of the tile (bf16) for the dVM phase
      *reinterpret_cast<uint32_t*>(myP + g * SPS + 2 * t) = pack_bf16(A[0], A[1]); *reinterpret_cast<uint32_t*>(myP + g * SPS + 8 + 2 * t) = pack_bf16(A[4], A[5]);
      *reinterpret_cast<uint32_t*>(myP + (g + 8) * SPS + 2 * t) = pack_bf16(A[2], A[3]); *reinterpret_cast<uint32_t*>(myP + (g + 8) * SPS + 8 + 2 * t) = pack_bf16(A[6], A[7]);
    }
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (active) {
        // ---- pass B: ctl, du, dctl -> dA (tensor cores), sdot = sum_c dxn * x; dctl staged for the dVM phase
#pragma unroll (UJ)
        for (int jj = 0; jj < JH; jj++) {
          const int j = h * JH + jj;
          const uint4 a = __ldg(x0 + j * 4 + t), bq = __ldg(x1 + j * 4 + t);
          const uint4 ga = __ldg(g0 + j * 4 + t), gb = __ldg(g1 + j * 4 + t);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
          const uint32_t gaw[4] = {ga.x, ga.y, ga.z, ga.w}, gbw[4] = {gb.x, gb.y, gb.z, gb.w};
          const int ch = (j * 4 + t) * 8;
          const float4 m1a = *reinterpret_cast<const float4*>(sm1 + ch), m1b = *reinterpret_cast<const float4*>(sm1 + ch + 4);
          const float4 ba = *reinterpret_cast<const float4*>(sb + ch), bb = *reinterpret_cast<const float4*>(sb + ch + 4);
          const float m1v[8] = {m1a.x, m1a.y, m1a.z, m1a.w, m1b.x, m1b.y, m1b.z, m1b.w};
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          uint32_t dc0[4], dc1[4];
#pragma unroll
          for (int m = 0; m < 4; m++) {
            float acc[4] = {m1v[2 * m], m1v[2 * m + 1], m1v[2 * m], m1v[2 * m + 1]};
            const uint2 vb = sVp[(j * 4 + m) * 32 + lane];
            mma16816<true>(acc, pa0, pa1, pa2, pa3, vb.x, vb.y);
            const float2 xa = upk<F16>(aw[m]), xb = upk<F16>(bw[m]);
            const float2 da = unpack_bf16(gaw[m]), db = unpack_bf16(gbw[m]);
            const float xn0 = xa.x * rn0, xn1 = xa.y * rn0, xn2 = xb.x * rn1, xn3 = xb.y * rn1;
            const float du0 = da.x * gm0 * ((xn0 * acc[0] + nz0 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
            const float du1 = da.y * gm0 * ((xn1 * acc[1] + nz0 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
            const float du2 = db.x * gm1 * ((xn2 * acc[2] + nz1 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
            const float du3 = db.y * gm1 * ((xn3 * acc[3] + nz1 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
            dc0[m] = pack_bf16(du0 * xn0, du1 * xn1); dc1[m] = pack_bf16(du2 * xn2, du3 * xn3);
            sd0 = fmaf(du0 * acc[0], xa.x, sd0); sd0 = fmaf(du1 * acc[1], xa.y, sd0);
            sd1 = fmaf(du2 * acc[2], xb.x, sd1); sd1 = fmaf(du3 * acc[3], xb.y, sd1);
          }
          const int off = (j * 4 + t) * 8;
          const uint4 w0 = *reinterpret_cast<const uint4*>(sVr + g * KS + off), w1 = *reinterpret_cast<const uint4*>(sVr + (g + 8) * KS + off);
          mma16816<false>(dA0, dc0[0], dc1[0], dc0[1], dc1[1], w0.x, w0.y); mma16816<false>(dA0, dc0[2], dc1[2], dc0[3], dc1[3], w0.z, w0.w);
          mma16816<false>(dA1, dc0[0], dc1[0], dc0[1], dc1[1], w1.x, w1.y); mma16816<false>(dA1, dc0[2], dc1[2], dc0[3], dc1[3], w1.z, w1.w);
          const int soff = (jj * 4 + t) * 8;
          *reinterpret_cast<uint4*>(myD + g * DS + soff) = make_uint4(dc0[0], dc0[1], dc0[2], dc0[3]);
          *reinterpret_cast<uint4*>(myD + (g + 8) * DS + soff) = make_uint4(dc1[0], dc1[1], dc1[2], dc1[3]);
        }
      }
      __syncthreads();
      // ---- CTA-wide phase: dVM[16, group h] += P^T[16, 128px] dctl[128px, group h]; warp w owns n-tiles {w, w+8, ...} of the group
      {
        const int lr = lane & 7, lq = lane >> 3;
        const int i2 = (NTW > 1) ? (lq >> 1) : 0;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {                  // 16 pixels per k-step (= the tile of warp ks)
          if (base + ks * 16 < pend) {                    // CTA-uniform
            uint32_t af[4];
            // matrices (px 0-7, t 0-7), (px 0-7, t 8-15), (px 8-15, t 0-7), (px 8-15, t 8-15), transposed on load -> a0..a3 of P^T
            ldsm_x4_t(af, sP + (ks * 16 + (lq >> 1) * 8 + lr) * SPS + (lq & 1) * 8);
#pragma unroll
            for (int i = 0; i < NTW; i += 2) {
              const int nt_l = warp + (i + i2) * 8;         // this lane's row address: n-tile i (lanes 0-15) or i+1 (lanes 16-31)
              uint32_t bf[4];
              ldsm_x4_t(bf, sD + (long long)(ks * 16 + (lq & 1) * 8 + lr) * DS + (nt_l < NTG ? nt_l : 0) * 8);
              if (warp + i * 8 < NTG) mma16816<false>(dvm[h][i], af[0], af[1], af[2], af[3], bf[0], bf[1]);
              if (i + 1 < NTW && warp + (i + 1) * 8 < NTG) mma16816<false>(dvm[h][i + 1 < NTW ? i + 1 : i], af[0], af[1], af[2], af[3], bf[2], bf[3]);
            }
          }
        }
      }
      __syncthreads();
    }
    if (active) {
      sd0 = quad_sum(sd0); sd1 = quad_sum(sd1);
      // dA is the gradient wrt the dropped probabilities: dP = dA * M, then the softmax backward dS = P * (dP - sum_t P dP);
      // the accumulator layout of dA equals the layout of P
      dA0[0] *= M[0]; dA0[1] *= M[1]; dA0[2] *= M[2]; dA0[3] *= M[3]; dA1[0] *= M[4]; dA1[1] *= M[5]; dA1[2] *= M[6]; dA1[3] *= M[7];
      const float ad0 = quad_sum(P[0] * dA0[0] + P[1] * dA0[1] + P[4] * dA1[0] + P[5] * dA1[1]);
      const float ad1 = quad_sum(P[2] * dA0[2] + P[3] * dA0[3] + P[6] * dA1[2] + P[7] * dA1[3]);
      const uint32_t sa0 = pack_bf16(P[0] * (dA0[0] - ad0), P[1] * (dA0[1] - ad0)), sa1 = pack_bf16(P[2] * (dA0[2] - ad1), P[3] * (dA0[3] - ad1));
      const uint32_t sa2 = pack_bf16(P[4] * (dA1[0] - ad0), P[5] * (dA1[1] - ad0)), sa3 = pack_bf16(P[6] * (dA1[2] - ad1), P[7] * (dA1[3] - ad1));
      const float k30 = rn0 * rn0 * rn0 * sd0 / (float)C, k31 = rn1 * rn1 * rn1 * sd1 / (float)C;
      // ---- pass C: dX = dS Kf + rn * dxn - x * k3 ; R[c] += dX * x
      uint4* o0 = reinterpret_cast<uint4*>(p.dX + ((long long)b * p.HW + q0) * C);
      uint4* o1 = reinterpret_cast<uint4*>(p.dX + ((long long)b * p.HW + q1) * C);
#pragma unroll (UJ)
      for (int j = 0; j < C32; j++) {
        const uint4 a = __ldg(x0 + j * 4 + t), bq = __ldg(x1 + j * 4 + t);
        const uint4 ga = __ldg(g0 + j * 4 + t), gb = __ldg(g1 + j * 4 + t);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
        const uint32_t gaw[4] = {ga.x, ga.y, ga.z, ga.w}, gbw[4] = {gb.x, gb.y, gb.z, gb.w};
        const int ch = (j * 4 + t) * 8;
        const float4 m1a = *reinterpret_cast<const float4*>(sm1 + ch), m1b = *reinterpret_cast<const float4*>(sm1 + ch + 4);
        const float4 ba = *reinterpret_cast<const float4*>(sb + ch), bb = *reinterpret_cast<const float4*>(sb + ch + 4);
        const float m1v[8] = {m1a.x, m1a.y, m1a.z, m1a.w, m1b.x, m1b.y, m1b.z, m1b.w};
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        uint32_t ow0[4], ow1[4];
        float rr[8];                                   // dX * x summed over this lane's two rows, 8 channels of vector (j, t)
#pragma unroll
        for (int m = 0; m < 4; m++) {
          float acc[4] = {m1v[2 * m], m1v[2 * m + 1], m1v[2 * m], m1v[2 * m + 1]};
          const uint2 vb = sVp[(j * 4 + m) * 32 + lane];
          mma16816<true>(acc, pa0, pa1, pa2, pa3, vb.x, vb.y);
          float dx[4] = {0.f, 0.f, 0.f, 0.f};
          const uint2 kb = sKp[(j * 4 + m) * 32 + lane];
          mma16816<false>(dx, sa0, sa1, sa2, sa3, kb.x, kb.y);
          const float2 xa = upk<F16>(aw[m]), xb = upk<F16>(bw[m]);
          const float2 da = unpack_bf16(gaw[m]), db = unpack_bf16(gbw[m]);
          const float du0 = da.x * gm0 * ((xa.x * rn0 * acc[0] + nz0 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
          const float du1 = da.y * gm0 * ((xa.y * rn0 * acc[1] + nz0 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
          const float du2 = db.x * gm1 * ((xb.x * rn1 * acc[2] + nz1 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
          const float du3 = db.y * gm1 * ((xb.y * rn1 * acc[3] + nz1 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
          const float d0 = dx[0] + rn0 * du0 * acc[0] - xa.x * k30, d1 = dx[1] + rn0 * du1 * acc[1] - xa.y * k30;
          const float d2 = dx[2] + rn1 * du2 * acc[2] - xb.x * k31, d3 = dx[3] + rn1 * du3 * acc[3] - xb.y * k31;
          ow0[m] = pack_bf16(d0, d1); ow1[m] = pack_bf16(d2, d3);
          rr[2 * m] = (v0 ? d0 * xa.x : 0.f) + (v1 ? d2 * xb.x : 0.f);
          rr[2 * m + 1] = (v0 ? d1 * xa.y : 0.f) + (v1 ? d3 * xb.y : 0.f);
        }
        if (v0) o0[j * 4 + t] = make_uint4(ow0[0], ow0[1], ow0[2], ow0[3]);
        if (v1) o1[j * 4 + t] = make_uint4(ow1[0], ow1[1], ow1[2], ow1[3]);
        // reduce rr over the 8 row groups g (lanes with equal t) by a transpose-reduce: lane g ends with the sum of channel ch + g
        float h4[4], h2[2];
        {
          const bool up = (g & 4) != 0;
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float send = up ? rr[e] : rr[e + 4], keep = up ? rr[e + 4] : rr[e];
            h4[e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool up = (g & 2) != 0;
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const float send = up ? h4[e] : h4[e + 2], keep = up ? h4[e + 2] : h4[e];
            h2[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool up = (g & 1) != 0;
          const float send = up ? h2[0] : h2[1], keep = up ? h2[1] : h2[0];
          // lane (g, t) now holds the 16-row sum of channel ch + g: one conflict-free shared-memory reduction per lane (bank = 8 t + g)
          atomicAdd(&sR[ch + g], keep + __shfl_xor_sync(0xffffffffu, send, 4));
        }
      }
    }
  }
  // ---- flush dVM: accumulator rows = latents g / g+8, columns = channels 2t, 2t+1 of the n-tile
#pragma unroll
  for (int h = 0; h < NH; h++)
#pragma unroll
    for (int i = 0; i < NTW; i++) {
      const int nt = warp + i * 8;
      if (nt < NTG) {
        float* d = p.dVM + (long long)b * NT * C + (h * NTG + nt) * 8 + 2 * t;
        atomicAdd(d + g * C, dvm[h][i][0]); atomicAdd(d + g * C + 1, dvm[h][i][1]);
        atomicAdd(d + (g + 8) * C, dvm[h][i][2]); atomicAdd(d + (g + 8) * C + 1, dvm[h][i][3]);
      }
    }
  __syncthreads();
  if (p.R) for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&p.R[(long long)b * C + i], sR[i]);
}

// ------------------------------------------------------------------------------------------- backward, low-resolution layers
// At 4^2 .. 32^2 a launch has only 8 .. 512 pixel tiles: one warp per 16-pixel x 512-channel tile is a 12 K-instruction dependent chain
// (27-38 us per launch whatever the grid size, profiles/r02b_membound_ncu.md).  Here WPT warps share a tile, each owning C / WPT channels:
// the score / norm partial sums and the dA / sdot partial sums are exchanged through shared memory (every warp adds the parts in the same order, so
// all of them continue with bit-identical probabilities), everything else is per channel.  A CTA of 8 warps holds 8 / WPT tiles per round.
template <int C32, int WPT>
struct SplitCfg {
  static constexpr int TPC = 8 / WPT;                  // tiles per CTA round
  static constexpr int JW = C32 / WPT;                 // 32-channel chunks per warp
  static constexpr int ROWS = TPC * 16;
  static_assert(C32 % WPT == 0 && 8 % WPT == 0, "split-channel attention: WPT must divide the chunk count and the warp count");
};
template <int C32, int WPT>
static int bwd_split_smem(int C) {
  using S = SplitCfg<C32, WPT>;
  return 3 * NT * krow(C) * 2 + 2 * (C / 8) * 32 * 8 + 3 * C * 4 + S::ROWS * SPS * 2 + S::ROWS * (C + 8) * 2 + 2 * 8 * 32 * 10 * 4;
}

__device__ __forceinline__ void tile_sync(int tile, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(tile + 1), "r"(nthreads) : "memory"); }

template <bool F16, int C32, int WPT>
__global__ void __launch_bounds__(256, 1) attn_bwd_split_kernel(AttnMP p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  using S = SplitCfg<C32, WPT>;
  constexpr int TPC = S::TPC, JW = S::JW, ROWS = S::ROWS;
  constexpr int NTG = C32 * 4, NTW = NTG / 8;              // n-tiles (8 channels) in all, per warp in the dVM phase
  constexpr int UJ = JW < 2 ? JW : 2;
  const int C = p.C, KS = krow(C), DS = C + 8;
  uint16_t* sKhi = reinterpret_cast<uint16_t*>(smraw);
  uint16_t* sKlo = sKhi + NT * KS;
  uint16_t* sVr = sKlo + NT * KS;
  uint2* sVp = reinterpret_cast<uint2*>(sVr + NT * KS);
  uint2* sKp = sVp + (C / 8) * 32;
  float* sm1 = reinterpret_cast<float*>(sKp + (C / 8) * 32);
  float* sb = sm1 + C;
  float* sR = sb + C;
  uint16_t* sP = reinterpret_cast<uint16_t*>(sR + C);
  uint16_t* sD = sP + ROWS * SPS;
  float* sRedA = reinterpret_cast<float*>(sD + ROWS * DS);  // [8 warps][32 lanes][10]: score / norm partial sums
  float* sRedB = sRedA + 8 * 32 * 10;                       // same shape: dA / sdot partial sums
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int tl = warp / WPT, part = warp % WPT, j0 = part * JW;
  const float* VMb = p.VM + (long long)b * NT * C;
  __shared__ __align__(8) uint64_t tbar;
  const bool bulk = p.tabK || p.tabV;
  if (bulk && threadIdx.x == 0) {
    tbar_init(&tbar);
    tbar_expect(&tbar, (uint32_t)((p.tabK ? tabk_bytes(C) : 0) + (p.tabV ? tabv_bytes(C) : 0)));
    if (p.tabK) {
      bulk_copy(sKhi, p.tabK, (uint32_t)(2 * rows_bytes(C)), &tbar);
      bulk_copy(sKp, reinterpret_cast<const unsigned char*>(p.tabK) + 2 * rows_bytes(C), (uint32_t)perm_bytes(C), &tbar);
    }
    if (p.tabV) bulk_copy(sVr, reinterpret_cast<const unsigned char*>(p.tabV) + (size_t)b * tabv_bytes(C), (uint32_t)tabv_bytes(C), &tbar);
  }
  if (!p.tabK) { fill_rowtable<F16>(sKhi, sKlo, p.Kf, C); fill_permtable<false>(sKp, p.Kf, C); }
  if (!p.tabV) { fill_rowtable<false>(sVr, nullptr, VMb, C); fill_permtable<true>(sVp, VMb, C); }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sb[i] = p.bias ? p.bias[i] : 0.f; sm1[i] = 1.f + p.bm[i]; sR[i] = 0.f; }
  __syncthreads();
  if (bulk) tbar_wait(&tbar);
  const float ns = (p.noise && p.nstr) ? *p.nstr : 0.f;
  const long long p0 = (long long)blockIdx.x * p.pix_per_cta;
  long long pend = p0 + p.pix_per_cta; if (pend > p.HW) pend = p.HW;
  const float* mbr = p.mb + b * NT;
  float dvm[NTW][4];
#pragma unroll
  for (int i = 0; i < NTW; i++) { dvm[i][0] = dvm[i][1] = dvm[i][2] = dvm[i][3] = 0.f; }

  for (long long base = p0; base < pend; base += ROWS) {
    const long long f0 = base + tl * 16;
    const bool active = f0 < pend;                        // uniform over the WPT warps of a tile
    uint16_t* myP = sP + tl * 16 * SPS;
    uint16_t* myD = sD + (long long)tl * 16 * DS;
    const long long r0 = f0 + g, r1 = f0 + g + 8;
    const bool v0 = r0 < pend, v1 = r1 < pend;
    const long long q0 = v0 ? r0 : pend - 1, q1 = v1 ? r1 : pend - 1;
    const uint4* x0 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.X) + ((long long)b * p.HW + q0) * C);
    const uint4* x1 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.X) + ((long long)b * p.HW + q1) * C);
    const uint4* g0 = reinterpret_cast<const uint4*>(p.dz + ((long long)b * p.HW + q0) * C);
    const uint4* g1 = reinterpret_cast<const uint4*>(p.dz + ((long long)b * p.HW + q1) * C);
    float P[8], M[8], rn0 = 0.f, rn1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { M[i] = 1.f; P[i] = 0.f; }
    uint32_t pa0 = 0, pa1 = 0, pa2 = 0, pa3 = 0;
    float nz0 = 0.f, nz1 = 0.f;
    const float gm0 = v0 ? p.gain : 0.f, gm1 = v1 ? p.gain : 0.f;
    if (active) {
      // ---- partial scores and squared norms over this warp's channels
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, ss0 = 0.f, ss1 = 0.f;
#pragma unroll (UJ)
      for (int jj = 0; jj < JW; jj++) {
        const int j = j0 + jj;
        const uint4 a = __ldg(x0 + j * 4 + t), bq = __ldg(x1 + j * 4 + t);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 fa = upk<F16>(aw[e]), fb = upk<F16>(bw[e]);
          ss0 = fmaf(fa.x, fa.x, ss0); ss0 = fmaf(fa.y, fa.y, ss0);
          ss1 = fmaf(fb.x, fb.x, ss1); ss1 = fmaf(fb.y, fb.y, ss1);
        }
        const int off = (j * 4 + t) * 8;
        const uint4 h0 = *reinterpret_cast<const uint4*>(sKhi + g * KS + off), h1 = *reinterpret_cast<const uint4*>(sKhi + (g + 8) * KS + off);
        const uint4 l0 = *reinterpret_cast<const uint4*>(sKlo + g * KS + off), l1 = *reinterpret_cast<const uint4*>(sKlo + (g + 8) * KS + off);
        mma16816<F16>(s0, a.x, bq.x, a.y, bq.y, h0.x, h0.y); mma16816<F16>(s0, a.z, bq.z, a.w, bq.w, h0.z, h0.w);
        mma16816<F16>(s1, a.x, bq.x, a.y, bq.y, h1.x, h1.y); mma16816<F16>(s1, a.z, bq.z, a.w, bq.w, h1.z, h1.w);
        mma16816<F16>(s0, a.x, bq.x, a.y, bq.y, l0.x, l0.y); mma16816<F16>(s0, a.z, bq.z, a.w, bq.w, l0.z, l0.w);
        mma16816<F16>(s1, a.x, bq.x, a.y, bq.y, l1.x, l1.y); mma16816<F16>(s1, a.z, bq.z, a.w, bq.w, l1.z, l1.w);
      }
      ss0 = quad_sum(ss0); ss1 = quad_sum(ss1);
      float* mine = sRedA + (warp * 32 + lane) * 10;
      mine[0] = s0[0]; mine[1] = s0[1]; mine[2] = s0[2]; mine[3] = s0[3]; mine[4] = s1[0]; mine[5] = s1[1]; mine[6] = s1[2]; mine[7] = s1[3];
      mine[8] = ss0; mine[9] = ss1;
      tile_sync(tl, WPT * 32);
      float tot[10];
#pragma unroll
      for (int i = 0; i < 10; i++) tot[i] = 0.f;
#pragma unroll
      for (int w = 0; w < WPT; w++) {
        const float* o = sRedA + ((tl * WPT + w) * 32 + lane) * 10;
#pragma unroll
        for (int i = 0; i < 10; i++) tot[i] += o[i];
      }
      rn0 = rsqrtf(tot[8] / (float)C + 1e-8f); rn1 = rsqrtf(tot[9] / (float)C + 1e-8f);
      const float* sc0 = p.Sc + q0 * NT; const float* sc1 = p.Sc + q1 * NT;
      const float2 ca0 = *reinterpret_cast<const float2*>(sc0 + 2 * t), ca1 = *reinterpret_cast<const float2*>(sc0 + 8 + 2 * t);
      const float2 cb0 = *reinterpret_cast<const float2*>(sc1 + 2 * t), cb1 = *reinterpret_cast<const float2*>(sc1 + 8 + 2 * t);
      const float2 m0 = *reinterpret_cast<const float2*>(mbr + 2 * t), m1 = *reinterpret_cast<const float2*>(mbr + 8 + 2 * t);
      P[0] = tot[0] + ca0.x + m0.x; P[1] = tot[1] + ca0.y + m0.y; P[4] = tot[4] + ca1.x + m1.x; P[5] = tot[5] + ca1.y + m1.y;
      P[2] = tot[2] + cb0.x + m0.x; P[3] = tot[3] + cb0.y + m0.y; P[6] = tot[6] + cb1.x + m1.x; P[7] = tot[7] + cb1.y + m1.y;
      const float mx0 = quad_max(fmaxf(fmaxf(P[0], P[1]), fmaxf(P[4], P[5])));
      const float mx1 = quad_max(fmaxf(fmaxf(P[2], P[3]), fmaxf(P[6], P[7])));
      P[0] = __expf(P[0] - mx0); P[1] = __expf(P[1] - mx0); P[4] = __expf(P[4] - mx0); P[5] = __expf(P[5] - mx0);
      P[2] = __expf(P[2] - mx1); P[3] = __expf(P[3] - mx1); P[6] = __expf(P[6] - mx1); P[7] = __expf(P[7] - mx1);
      const float i0 = 1.f / quad_sum(P[0] + P[1] + P[4] + P[5]), i1 = 1.f / quad_sum(P[2] + P[3] + P[6] + P[7]);
      P[0] *= i0; P[1] *= i0; P[4] *= i0; P[5] *= i0; P[2] *= i1; P[3] *= i1; P[6] *= i1; P[7] *= i1;
      if (p.dmask) load_dmask(p.dmask + ((long long)b * p.HW + q0) * NT, p.dmask + ((long long)b * p.HW + q1) * NT, t, M);
      float A[8];
#pragma unroll
      for (int i = 0; i < 8; i++) A[i] = P[i] * M[i];
      pa0 = pk<true>(A[0], A[1]); pa1 = pk<true>(A[2], A[3]); pa2 = pk<true>(A[4], A[5]); pa3 = pk<true>(A[6], A[7]);
      if (p.noise) { nz0 = p.noise[b * p.nbs + q0] * ns; nz1 = p.noise[b * p.nbs + q1] * ns; }
      if (part == 0) {
        *reinterpret_cast<uint32_t*>(myP + g * SPS + 2 * t) = pack_bf16(A[0], A[1]); *reinterpret_cast<uint32_t*>(myP + g * SPS + 8 + 2 * t) = pack_bf16(A[4], A[5]);
        *reinterpret_cast<uint32_t*>(myP + (g + 8) * SPS + 2 * t) = pack_bf16(A[2], A[3]); *reinterpret_cast<uint32_t*>(myP + (g + 8) * SPS + 8 + 2 * t) = pack_bf16(A[6], A[7]);
      }
    }
    float dA0[4] = {0.f, 0.f, 0.f, 0.f}, dA1[4] = {0.f, 0.f, 0.f, 0.f}, sd0 = 0.f, sd1 = 0.f;
    if (active) {
      // ---- pass B over this warp's channels: ctl, du, dctl -> partial dA and sdot; dctl staged for the dVM phase
#pragma unroll (UJ)
      for (int jj = 0; jj < JW; jj++) {
        const int j = j0 + jj;
        const uint4 a = __ldg(x0 + j * 4 + t), bq = __ldg(x1 + j * 4 + t);
        const uint4 ga = __ldg(g0 + j * 4 + t), gb = __ldg(g1 + j * 4 + t);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
        const uint32_t gaw[4] = {ga.x, ga.y, ga.z, ga.w}, gbw[4] = {gb.x, gb.y, gb.z, gb.w};
        const int ch = (j * 4 + t) * 8;
        const float4 m1a = *reinterpret_cast<const float4*>(sm1 + ch), m1b = *reinterpret_cast<const float4*>(sm1 + ch + 4);
        const float4 ba = *reinterpret_cast<const float4*>(sb + ch), bb = *reinterpret_cast<const float4*>(sb + ch + 4);
        const float m1v[8] = {m1a.x, m1a.y, m1a.z, m1a.w, m1b.x, m1b.y, m1b.z, m1b.w};
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        uint32_t dc0[4], dc1[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          float acc[4] = {m1v[2 * m], m1v[2 * m + 1], m1v[2 * m], m1v[2 * m + 1]};
          const uint2 vb = sVp[(j * 4 + m) * 32 + lane];
          mma16816<true>(acc, pa0, pa1, pa2, pa3, vb.x, vb.y);
          const float2 xa = upk<F16>(aw[m]), xb = upk<F16>(bw[m]);
          const float2 da = unpack_bf16(gaw[m]), db = unpack_bf16(gbw[m]);
          const float xn0 = xa.x * rn0, xn1 = xa.y * rn0, xn2 = xb.x * rn1, xn3 = xb.y * rn1;
          const float du0 = da.x * gm0 * ((xn0 * acc[0] + nz0 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
          const float du1 = da.y * gm0 * ((xn1 * acc[1] + nz0 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
          const float du2 = db.x * gm1 * ((xn2 * acc[2] + nz1 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
          const float du3 = db.y * gm1 * ((xn3 * acc[3] + nz1 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
          dc0[m] = pack_bf16(du0 * xn0, du1 * xn1); dc1[m] = pack_bf16(du2 * xn2, du3 * xn3);
          sd0 = fmaf(du0 * acc[0], xa.x, sd0); sd0 = fmaf(du1 * acc[1], xa.y, sd0);
          sd1 = fmaf(du2 * acc[2], xb.x, sd1); sd1 = fmaf(du3 * acc[3], xb.y, sd1);
        }
        const int off = (j * 4 + t) * 8;
        const uint4 w0 = *reinterpret_cast<const uint4*>(sVr + g * KS + off), w1 = *reinterpret_cast<const uint4*>(sVr + (g + 8) * KS + off);
        mma16816<false>(dA0, dc0[0], dc1[0], dc0[1], dc1[1], w0.x, w0.y); mma16816<false>(dA0, dc0[2], dc1[2], dc0[3], dc1[3], w0.z, w0.w);
        mma16816<false>(dA1, dc0[0], dc1[0], dc0[1], dc1[1], w1.x, w1.y); mma16816<false>(dA1, dc0[2], dc1[2], dc0[3], dc1[3], w1.z, w1.w);
        *reinterpret_cast<uint4*>(myD + g * DS + off) = make_uint4(dc0[0], dc0[1], dc0[2], dc0[3]);
        *reinterpret_cast<uint4*>(myD + (g + 8) * DS + off) = make_uint4(dc1[0], dc1[1], dc1[2], dc1[3]);
      }
      sd0 = quad_sum(sd0); sd1 = quad_sum(sd1);
      float* mine = sRedB + (warp * 32 + lane) * 10;
      mine[0] = dA0[0]; mine[1] = dA0[1]; mine[2] = dA0[2]; mine[3] = dA0[3]; mine[4] = dA1[0]; mine[5] = dA1[1]; mine[6] = dA1[2]; mine[7] = dA1[3];
      mine[8] = sd0; mine[9] = sd1;
    }
    __syncthreads();                                     // dctl / probabilities of all tiles staged, partial sums visible
    // ---- CTA-wide phase: dVM[16, C] += P^T[16, ROWS px] dctl[ROWS px, C]; warp w owns n-tiles {w, w+8, ...}
    {
      const int lr = lane & 7, lq = lane >> 3;
      const int i2 = lq >> 1;
#pragma unroll
      for (int ks = 0; ks < TPC; ks++) {
        if (base + ks * 16 < pend) {                      // CTA-uniform
          uint32_t af[4];
          ldsm_x4_t(af, sP + (ks * 16 + (lq >> 1) * 8 + lr) * SPS + (lq & 1) * 8);
#pragma unroll
          for (int i = 0; i < NTW; i += 2) {
            const int nt_l = warp + (i + i2) * 8;
            uint32_t bf[4];
            ldsm_x4_t(bf, sD + (long long)(ks * 16 + (lq & 1) * 8 + lr) * DS + nt_l * 8);
            mma16816<false>(dvm[i], af[0], af[1], af[2], af[3], bf[0], bf[1]);
            mma16816<false>(dvm[i + 1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
          }
        }
      }
    }
    if (active) {
      float tot[10];
#pragma unroll
      for (int i = 0; i < 10; i++) tot[i] = 0.f;
#pragma unroll
      for (int w = 0; w < WPT; w++) {
        const float* o = sRedB + ((tl * WPT + w) * 32 + lane) * 10;
#pragma unroll
        for (int i = 0; i < 10; i++) tot[i] += o[i];
      }
      // dP = dA * M; dS = P * (dP - sum_t P dP)
      const float d00 = tot[0] * M[0], d01 = tot[1] * M[1], d02 = tot[2] * M[2], d03 = tot[3] * M[3];
      const float d10 = tot[4] * M[4], d11 = tot[5] * M[5], d12 = tot[6] * M[6], d13 = tot[7] * M[7];
      const float ad0 = quad_sum(P[0] * d00 + P[1] * d01 + P[4] * d10 + P[5] * d11);
      const float ad1 = quad_sum(P[2] * d02 + P[3] * d03 + P[6] * d12 + P[7] * d13);
      const uint32_t sa0 = pack_bf16(P[0] * (d00 - ad0), P[1] * (d01 - ad0)), sa1 = pack_bf16(P[2] * (d02 - ad1), P[3] * (d03 - ad1));
      const uint32_t sa2 = pack_bf16(P[4] * (d10 - ad0), P[5] * (d11 - ad0)), sa3 = pack_bf16(P[6] * (d12 - ad1), P[7] * (d13 - ad1));
      const float k30 = rn0 * rn0 * rn0 * tot[8] / (float)C, k31 = rn1 * rn1 * rn1 * tot[9] / (float)C;
      // ---- pass C over this warp's channels: dX = dS Kf + rn * dxn - x * k3 ; R[c] += dX * x
      uint4* o0 = reinterpret_cast<uint4*>(p.dX + ((long long)b * p.HW + q0) * C);
      uint4* o1 = reinterpret_cast<uint4*>(p.dX + ((long long)b * p.HW + q1) * C);
#pragma unroll (UJ)
      for (int jj = 0; jj < JW; jj++) {
        const int j = j0 + jj;
        const uint4 a = __ldg(x0 + j * 4 + t), bq = __ldg(x1 + j * 4 + t);
        const uint4 ga = __ldg(g0 + j * 4 + t), gb = __ldg(g1 + j * 4 + t);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w};
        const uint32_t gaw[4] = {ga.x, ga.y, ga.z, ga.w}, gbw[4] = {gb.x, gb.y, gb.z, gb.w};
        const int ch = (j * 4 + t) * 8;
        const float4 m1a = *reinterpret_cast<const float4*>(sm1 + ch), m1b = *reinterpret_cast<const float4*>(sm1 + ch + 4);
        const float4 ba = *reinterpret_cast<const float4*>(sb + ch), bb = *reinterpret_cast<const float4*>(sb + ch + 4);
        const float m1v[8] = {m1a.x, m1a.y, m1a.z, m1a.w, m1b.x, m1b.y, m1b.z, m1b.w};
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        uint32_t ow0[4], ow1[4];
        float rr[8];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          float acc[4] = {m1v[2 * m], m1v[2 * m + 1], m1v[2 * m], m1v[2 * m + 1]};
          const uint2 vb = sVp[(j * 4 + m) * 32 + lane];
          mma16816<true>(acc, pa0, pa1, pa2, pa3, vb.x, vb.y);
          float dx[4] = {0.f, 0.f, 0.f, 0.f};
          const uint2 kb = sKp[(j * 4 + m) * 32 + lane];
          mma16816<false>(dx, sa0, sa1, sa2, sa3, kb.x, kb.y);
          const float2 xa = upk<F16>(aw[m]), xb = upk<F16>(bw[m]);
          const float2 da = unpack_bf16(gaw[m]), db = unpack_bf16(gbw[m]);
          const float du0 = da.x * gm0 * ((xa.x * rn0 * acc[0] + nz0 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
          const float du1 = da.y * gm0 * ((xa.y * rn0 * acc[1] + nz0 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
          const float du2 = db.x * gm1 * ((xb.x * rn1 * acc[2] + nz1 + bv[2 * m]) > 0.f ? 1.f : p.alpha);
          const float du3 = db.y * gm1 * ((xb.y * rn1 * acc[3] + nz1 + bv[2 * m + 1]) > 0.f ? 1.f : p.alpha);
          const float d0 = dx[0] + rn0 * du0 * acc[0] - xa.x * k30, d1 = dx[1] + rn0 * du1 * acc[1] - xa.y * k30;
          const float d2 = dx[2] + rn1 * du2 * acc[2] - xb.x * k31, d3 = dx[3] + rn1 * du3 * acc[3] - xb.y * k31;
          ow0[m] = pack_bf16(d0, d1); ow1[m] = pack_bf16(d2, d3);
          rr[2 * m] = (v0 ? d0 * xa.x : 0.f) + (v1 ? d2 * xb.x : 0.f);
          rr[2 * m + 1] = (v0 ? d1 * xa.y : 0.f) + (v1 ? d3 * xb.y : 0.f);
        }
        if (v0) o0[j * 4 + t] = make_uint4(ow0[0], ow0[1], ow0[2], ow0[3]);
        if (v1) o1[j * 4 + t] = make_uint4(ow1[0], ow1[1], ow1[2], ow1[3]);
        float h4[4], h2[2];
        {
          const bool up = (g & 4) != 0;
#pragma unroll
          for (int e = 0; e < 4; e++) { const float send = up ? rr[e] : rr[e + 4], keep = up ? rr[e + 4] : rr[e]; h4[e] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
        }
        {
          const bool up = (g & 2) != 0;
#pragma unroll
          for (int e = 0; e < 2; e++) { const float send = up ? h4[e] : h4[e + 2], keep = up ? h4[e + 2] : h4[e]; h2[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
        }
        {
          const bool up = (g & 1) != 0;
          const float send = up ? h2[0] : h2[1], keep = up ? h2[1] : h2[0];
          atomicAdd(&sR[ch + g], keep + __shfl_xor_sync(0xffffffffu, send, 4));
        }
      }
    }
    __syncthreads();                                     // staging buffers and partial sums are rewritten by the next round
  }
#pragma unroll
  for (int i = 0; i < NTW; i++) {
    const int nt = warp + i * 8;
    float* d = p.dVM + (long long)b * NT * C + nt * 8 + 2 * t;
    atomicAdd(d + g * C, dvm[i][0]); atomicAdd(d + g * C + 1, dvm[i][1]);
    atomicAdd(d + (g + 8) * C, dvm[i][2]); atomicAdd(d + (g + 8) * C + 1, dvm[i][3]);
  }
  __syncthreads();
  if (p.R) for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&p.R[(long long)b * C + i], sR[i]);
}

template <bool F16, int WPT>
int launch_bwd_split(const AttnMP& p, dim3 grid, cudaStream_t st) {
  const int smem = bwd_split_smem<16, WPT>(p.C);
  static bool done = false;
  if (!done) { cudaFuncSetAttribute(attn_bwd_split_kernel<F16, 16, WPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); done = true; }
  attn_bwd_split_kernel<F16, 16, WPT><<<grid, 256, smem, st>>>(p);
  return 0;
}

template <bool F16>
int launch_fwd(const AttnMP& p, dim3 grid, int smem, cudaStream_t st) {
  switch (p.C / 32) {
#define MGF_CASE(n) case n: { static bool done = false; if (!done) { cudaFuncSetAttribute(attn_fwd_mma_kernel<F16, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(32 * n)); done = true; } \
                             attn_fwd_mma_kernel<F16, n><<<grid, 256, smem, st>>>(p); return 0; }
    MGF_CASE(1) MGF_CASE(2) MGF_CASE(4) MGF_CASE(8) MGF_CASE(12) MGF_CASE(16)
#undef MGF_CASE
  }
  return -1;
}
template <bool F16>
int launch_bwd(const AttnMP& p, dim3 grid, int smem, cudaStream_t st) {
  switch (p.C / 32) {
#define MGF_CASE(n) case n: { static bool done = false; if (!done) { cudaFuncSetAttribute(attn_bwd_mma_kernel<F16, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(32 * n)); done = true; } \
                             attn_bwd_mma_kernel<F16, n><<<grid, 256, smem, st>>>(p); return 0; }
    MGF_CASE(1) MGF_CASE(2) MGF_CASE(4) MGF_CASE(8) MGF_CASE(12) MGF_CASE(16)
#undef MGF_CASE
  }
  return -1;
}

// pixels per CTA: the images share the SMs (blockIdx.y = image), a CTA takes whole 16-pixel warp tiles and the grid is one wave
int pix_per_cta(long long HW, int B, int ctas_per_sm) {
  long long per_img = ((long long)num_sms() * ctas_per_sm) / B;
  if (per_img < 1) per_img = 1;
  long long ppc = (HW + per_img - 1) / per_img;
  ppc = (ppc + 15) / 16 * 16;
  if (ppc < 128) ppc = HW < 128 ? (HW + 15) / 16 * 16 : 128;
  return (int)ppc;
}

bool g_attn_split = true;     // split-channel backward kernel for the low-resolution layers (A/B switch: mgf_attn_set_split)

int check_c(int C, const char* who) {
  const int n = C / 32;
  if (C % 32 != 0 || !(n == 1 || n == 2 || n == 4 || n == 8 || n == 12 || n == 16))
    MGF_FAIL(MGF_E_SHAPE, "%s: C=%d must be one of 32, 64, 128, 256, 384, 512", who, C);
  return 0;
}

}  // namespace
}  // namespace mgf

using namespace mgf;

extern "C" int mgf_attn_fwd(const void* X, const float* Kf, const float* Sc, const float* maskbias, const float* VM, const float* bm,
                            const float* noise, const float* nstr, const float* bias, float gain, float alpha,
                            void* out, float* probs, const float* dmask, const void* tabK, const void* tabV,
                            int B, int64_t HW, int C, int64_t noise_bstride, void* stream) {
  if (!X || !Kf || !Sc || !maskbias || !VM || !bm || !out) MGF_FAIL(MGF_E_BADARG, "attn_fwd: null tensor");
  if (B <= 0 || HW <= 0) MGF_FAIL(MGF_E_SHAPE, "attn_fwd: empty batch or grid");
  if (int e = check_c(C, "attn_fwd")) return e;
  AttnMP p{}; p.X = X; p.Kf = Kf; p.Sc = Sc; p.mb = maskbias; p.VM = VM; p.bm = bm; p.noise = noise; p.nstr = nstr; p.bias = bias;
  p.gain = gain; p.alpha = alpha; p.out = out; p.probs = probs; p.dmask = dmask; p.tabK = tabK; p.tabV = tabV; p.HW = HW; p.C = C; p.nbs = noise_bstride;
  p.pix_per_cta = pix_per_cta(HW, B, 2);      // one wave of 2 CTAs per SM
  dim3 grid((unsigned)((HW + p.pix_per_cta - 1) / p.pix_per_cta), B);
  const int rc = fwd_f16() ? launch_fwd<true>(p, grid, fwd_smem(C), (cudaStream_t)stream) : launch_fwd<false>(p, grid, fwd_smem(C), (cudaStream_t)stream);
  if (rc) MGF_FAIL(MGF_E_SHAPE, "attn_fwd: unsupported C=%d", C);
  MGF_CHECK_LAUNCH("attn_fwd");
  return 0;
}

extern "C" int mgf_attn_bwd(const void* X, const void* dz, const float* Kf, const float* Sc, const float* maskbias, const float* VM, const float* bm,
                            const float* noise, const float* nstr, const float* bias, float gain, float alpha,
                            void* dX, float* dVM, float* R, const float* dmask, const void* tabK, const void* tabV,
                            int B, int64_t HW, int C, int64_t noise_bstride, void* stream) {
  if (!X || !dz || !Kf || !Sc || !maskbias || !VM || !bm || !dX || !dVM) MGF_FAIL(MGF_E_BADARG, "attn_bwd: null tensor");
  if (B <= 0 || HW <= 0) MGF_FAIL(MGF_E_SHAPE, "attn_bwd: empty batch or grid");
  if (int e = check_c(C, "attn_bwd")) return e;
  AttnMP p{}; p.X = X; p.dz = (const __nv_bfloat16*)dz; p.Kf = Kf; p.Sc = Sc; p.mb = maskbias; p.VM = VM; p.bm = bm;
  p.noise = noise; p.nstr = nstr; p.bias = bias; p.gain = gain; p.alpha = alpha; p.dX = (__nv_bfloat16*)dX; p.dVM = dVM; p.R = R; p.dmask = dmask; p.tabK = tabK; p.tabV = tabV; p.HW = HW; p.C = C; p.nbs = noise_bstride;
  // low-resolution 512-channel layers: several warps per pixel tile (attn_bwd_split_kernel) while the tiles of the launch do not fill the SMs
  if (C == 512 && g_attn_split) {
    const long long tiles = (HW + 15) / 16 * B;
    const int wpt = tiles <= num_sms() ? 8 : (tiles / 2 <= num_sms() ? 4 : (tiles / 4 <= num_sms() ? 2 : 0));
    if (wpt) {
      p.pix_per_cta = (8 / wpt) * 16;
      dim3 grid((unsigned)((HW + p.pix_per_cta - 1) / p.pix_per_cta), B);
      const bool f16 = fwd_f16();
      if (wpt == 8) { if (f16) launch_bwd_split<true, 8>(p, grid, (cudaStream_t)stream); else launch_bwd_split<false, 8>(p, grid, (cudaStream_t)stream); }
      else if (wpt == 4) { if (f16) launch_bwd_split<true, 4>(p, grid, (cudaStream_t)stream); else launch_bwd_split<false, 4>(p, grid, (cudaStream_t)stream); }
      else { if (f16) launch_bwd_split<true, 2>(p, grid, (cudaStream_t)stream); else launch_bwd_split<false, 2>(p, grid, (cudaStream_t)stream); }
      MGF_CHECK_LAUNCH("attn_bwd(split)");
      return 0;
    }
  }
  p.pix_per_cta = pix_per_cta(HW, B, C <= 256 ? 2 : 1);      // one wave: two CTAs per SM up to 256 channels, one above
  dim3 grid((unsigned)((HW + p.pix_per_cta - 1) / p.pix_per_cta), B);
  const int rc = fwd_f16() ? launch_bwd<true>(p, grid, bwd_smem(C), (cudaStream_t)stream) : launch_bwd<false>(p, grid, bwd_smem(C), (cudaStream_t)stream);
  if (rc) MGF_FAIL(MGF_E_SHAPE, "attn_bwd: unsupported C=%d", C);
  MGF_CHECK_LAUNCH("attn_bwd");
  return 0;
}

extern "C" int64_t mgf_attn_table_bytes(int which, int C) {
  if (C < 32 || C % 32) return -1;
  return (int64_t)(which == 0 ? mgf::tabk_bytes(C) : mgf::tabv_bytes(C));
}

extern "C" int mgf_attn_tables(const float* Kf, const float* VM, void* tabK, void* tabV, int B, int C, void* stream) {
  if ((!Kf && !VM) || (Kf && !tabK) || (VM && !tabV)) MGF_FAIL(MGF_E_BADARG, "attn_tables: null tensor");
  if (int e = check_c(C, "attn_tables")) return e;
  if (VM && B <= 0) MGF_FAIL(MGF_E_SHAPE, "attn_tables: empty batch");
  if (((uintptr_t)tabK | (uintptr_t)tabV) & 15) MGF_FAIL(MGF_E_ALIGN, "attn_tables: tables must be 16-byte aligned");
  const unsigned grid = (VM ? B : 0) + (Kf ? 1 : 0);
  if (fwd_f16()) attn_tables_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(Kf, VM, (unsigned char*)tabK, (unsigned char*)tabV, C);
  else attn_tables_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(Kf, VM, (unsigned char*)tabK, (unsigned char*)tabV, C);
  MGF_CHECK_LAUNCH("attn_tables");
  return 0;
}

/* A/B switch: 0 keeps the one-warp-per-tile backward kernel for every layer */
extern "C" int mgf_attn_set_split(int enabled) { mgf::g_attn_split = enabled != 0; return 0; }
