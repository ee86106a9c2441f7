// pointwise.cu -- 1x1 convolution with shared weights on NHWC 16-bit tensors: out[p, n] = sum_k x[p, k] W[n, k].
//
// Replaces the resnet-skip 1x1 Conv2dLayer (reference training/networks.py:245-250 with kernel_size 1, called from SynthesisBlock.forward
// :1157-1160) and its input gradient at the resolutions where the layer is pure HBM traffic (>= 128^2: 32..256 channels).  On the tcgen05
// implicit-GEMM kernel these launches are bound by the per-tile epilogue of a one-tap, K <= 256 tile (two MMAs' worth of work per 16 KB of
// output: 2.1 TB/s of output, scripts/bench_halo.py), so they run here on mma.sync.m16n8k16 with every warp streaming pixels:
// a warp owns 16 pixels; lane (g = lane / 4, t = lane % 4) owns rows g and g + 8 and loads / stores whole 16-byte vectors through the
// same channel permutations as the attention kernels (attention_mma.cu): vector v = j*4 + t of a row holds channels [8v, 8v + 8), its
// four words feed the A fragments of two k-steps, and output column 2t + e of n-tile (jj*4 + m) is channel (jj*4 + t)*8 + 2m + e, so a
// lane ends up with 8 consecutive output channels per 32-channel group.  The weights sit in shared memory (rows padded by 64 bytes:
// the 16-byte reads of a quarter-warp fall into distinct banks).
#include "common.cuh"

namespace mgf {
namespace {

template <bool F16>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ int perm_chan(int nt, int g) { return ((nt >> 2) * 4 + (g >> 1)) * 8 + 2 * (nt & 3) + (g & 1); }
__host__ __device__ inline int wrow(int K) { return K + 32; }      // padded weight row (elements)

// K32 = K / 32 (compile time: the activation vectors of a pixel pair stay in registers across the output-channel groups)
template <bool F16, int K32>
__global__ void __launch_bounds__(256) pointwise_kernel(const uint16_t* __restrict__ X, const uint16_t* __restrict__ W, uint16_t* __restrict__ out,
                                                        long long P, int N, unsigned int* ovf) {
  extern __shared__ __align__(16) uint16_t sW[];            // [N][K + 32]
  constexpr int K = 32 * K32;
  const int KS = wrow(K);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // weight staging, eight 16-byte loads in flight per thread (a one-load-per-trip loop cost 16 dependent L2 round trips at K = 256, N = 128)
  {
    const int nvec = N * (K / 8);
    for (int i0 = threadIdx.x; i0 < nvec; i0 += 8 * blockDim.x) {
      uint4 w[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { const int i = i0 + u * blockDim.x; if (i < nvec) w[u] = __ldg(reinterpret_cast<const uint4*>(W) + i); }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = i0 + u * blockDim.x;
        if (i < nvec) { const int n = i / (K / 8), v = i - n * (K / 8); *reinterpret_cast<uint4*>(sW + n * KS + v * 8) = w[u]; }
      }
    }
  }
  __syncthreads();
  float mx = 0.f;
  const long long ntile = (P + 15) / 16;
  for (long long tile = (long long)blockIdx.x * 8 + warp; tile < ntile; tile += (long long)gridDim.x * 8) {
    const long long r0 = tile * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < P, v1 = r1 < P;
    const uint4* x0 = reinterpret_cast<const uint4*>(X + (v0 ? r0 : P - 1) * K);
    const uint4* x1 = reinterpret_cast<const uint4*>(X + (v1 ? r1 : P - 1) * K);
    uint4 a[K32], b[K32];
#pragma unroll
    for (int j = 0; j < K32; j++) { a[j] = __ldg(x0 + j * 4 + t); b[j] = __ldg(x1 + j * 4 + t); }
    uint4* o0 = reinterpret_cast<uint4*>(out + r0 * N);
    uint4* o1 = reinterpret_cast<uint4*>(out + r1 * N);
#pragma unroll 1
    for (int jj = 0; jj < N / 32; jj++) {                  // 32 output channels per trip: four n-tiles
      float acc[4][4];
#pragma unroll
      for (int m = 0; m < 4; m++) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }
#pragma unroll
      for (int j = 0; j < K32; j++) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const uint4 w = *reinterpret_cast<const uint4*>(sW + perm_chan(jj * 4 + m, g) * KS + (j * 4 + t) * 8);
          mma16816<F16>(acc[m], a[j].x, b[j].x, a[j].y, b[j].y, w.x, w.y);
          mma16816<F16>(acc[m], a[j].z, b[j].z, a[j].w, b[j].w, w.z, w.w);
        }
      }
      uint32_t w0[4], w1[4];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        if (F16) { mx = ovf_max(mx, acc[m][0]); mx = ovf_max(mx, acc[m][1]); mx = ovf_max(mx, acc[m][2]); mx = ovf_max(mx, acc[m][3]); }
        w0[m] = pack16(acc[m][0], acc[m][1], F16); w1[m] = pack16(acc[m][2], acc[m][3], F16);
      }
      if (v0) o0[jj * 4 + t] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
      if (v1) o1[jj * 4 + t] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    }
  }
  if (F16) ovf_commit(ovf, mx);
}

template <bool F16>
int launch_pw(const void* x, const void* w, void* out, long long P, int K, int N, cudaStream_t st) {
  const int smem = N * wrow(K) * 2;
  // every CTA stages the whole weight matrix first (up to 72 KB): at least four pixel tiles per warp, so that on the small grids the
  // staging traffic stays below the activation traffic
  long long blocks = (P / 16 + 31) / 32;
  const long long cap = (long long)num_sms() * 6;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  unsigned int* ovf = F16 ? overflow_flag() : nullptr;
  switch (K / 32) {
#define MGF_CASE(n) case n: { static bool done = false; \
      if (!done) { cudaFuncSetAttribute(pointwise_kernel<F16, n>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); done = true; } \
      pointwise_kernel<F16, n><<<(unsigned)blocks, 256, smem, st>>>((const uint16_t*)x, (const uint16_t*)w, (uint16_t*)out, P, N, ovf); return 0; }
    MGF_CASE(1) MGF_CASE(2) MGF_CASE(4) MGF_CASE(8)
#undef MGF_CASE
  }
  return -1;
}

}  // namespace
}  // namespace mgf

using namespace mgf;

// x [P, K], w [N, K] (the layout of a [1, 1, N, K] mgf_conv_tc weight tensor), out [P, N]; all three 16-bit with channel stride 1.
// is_fwd != 0: forward tensors (bf16 or fp16 per mgf_set_forward_dtype, fp16 stores tracked by the overflow flag); 0: bf16 gradients.
// Supported: K in {32, 64, 128, 256}, N a multiple of 32 with the padded weights within 96 KB of shared memory; MGF_E_UNSUP otherwise
// (the caller then uses mgf_conv_tc).
extern "C" int mgf_pointwise(const void* x, const void* w, void* out, int64_t P, int K, int N, int is_fwd, void* stream) {
  if (!x || !w || !out) MGF_FAIL(MGF_E_BADARG, "pointwise: null tensor");
  if (P <= 0) MGF_FAIL(MGF_E_SHAPE, "pointwise: empty input");
  if (!(K == 32 || K == 64 || K == 128 || K == 256) || N < 32 || N % 32 || (long long)N * wrow(K) * 2 > 96 * 1024)
    MGF_FAIL(MGF_E_UNSUP, "pointwise: K=%d N=%d outside the supported shapes", K, N);
  if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) MGF_FAIL(MGF_E_ALIGN, "pointwise: tensors must be 16-byte aligned");
  const bool f16 = is_fwd && fwd_f16();
  const int rc = f16 ? launch_pw<true>(x, w, out, P, K, N, (cudaStream_t)stream) : launch_pw<false>(x, w, out, P, K, N, (cudaStream_t)stream);
  if (rc) MGF_FAIL(MGF_E_UNSUP, "pointwise: no kernel for K=%d", K);
  MGF_CHECK_LAUNCH("pointwise");
  return 0;
}

extern "C" int mgf_pointwise_supported(int K, int N) {
  return ((K == 32 || K == 64 || K == 128 || K == 256) && N >= 32 && N % 32 == 0 && (long long)N * mgf::wrow(K) * 2 <= 96 * 1024) ? 1 : 0;
}
