// upfirdn2d.cu -- pad -> zero-insert upsample -> 2-D FIR -> decimate -> gain, any up/down/filter/padding.
// Semantics: reference _upfirdn2d_ref (torch_utils/ops/upfirdn2d.py:161-200) / upfirdn2d.cu:21-192.
// sm_100a design: one CTA produces a 64x16 output tile of one (n,c) plane from an input patch staged in
// shared memory (zero-filled outside the image, so padding/cropping costs nothing), only the filter taps
// that hit non-zero upsampled samples are visited, each thread writes 4 consecutive pixels (16-byte
// stores when aligned).  Planes that are not W-contiguous (channels_last) take the gather kernel.
// HBM-bound: algorithmic bytes = (inW*inH + outW*outH) * C * N * sizeof(T).
#include "common.cuh"

namespace mgf {

struct UpfirdnParams {
  const void* x; const float* f; void* y;
  long long inW, inH, inC, inN, isW, isH, isC, isN;
  long long outW, outH, osW, osH, osC, osN;
  int fW, fH; long long fsW, fsH;
  int upx, upy, downx, downy, padx0, pady0, flip; float gain;
  int tileInW, tileInH;
};

constexpr int TW = 64, TH = 16, VX = 4;

__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ int ceildiv(int a, int b) { return -floordiv(-a, b); }

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_tile_kernel(UpfirdnParams p) {
  typedef typename Acc<T>::type S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S* sf = reinterpret_cast<S*>(smem_raw);                 // [fH][fW], already flipped + gain
  S* sx = sf + ((p.fW * p.fH + 3) & ~3);                    // [tileInH][tileInW]
  const int tilesX = (int)((p.outW + TW - 1) / TW);
  const int tilesY = (int)((p.outH + TH - 1) / TH);
  const int tile = blockIdx.x % (tilesX * tilesY);
  const int tx = tile % tilesX, ty = tile / tilesX;
  const long long plane = blockIdx.x / (tilesX * tilesY);   // n * C + c
  const long long n = plane / p.inC, c = plane % p.inC;
  const int ox0 = tx * TW, oy0 = ty * TH;
  // first input sample needed by this tile
  const int ix0 = ceildiv(ox0 * p.downx - p.padx0, p.upx);
  const int iy0 = ceildiv(oy0 * p.downy - p.pady0, p.upy);
  const T* xp = (const T*)p.x + n * p.isN + c * p.isC;
  for (int i = threadIdx.x; i < p.fW * p.fH; i += blockDim.x) {
    int fy = i / p.fW, fx = i % p.fW;
    int sy = p.flip ? fy : p.fH - 1 - fy, sxx = p.flip ? fx : p.fW - 1 - fx;
    sf[i] = (S)p.f[sy * p.fsH + sxx * p.fsW] * (S)p.gain;
  }
  for (int i = threadIdx.x; i < p.tileInW * p.tileInH; i += blockDim.x) {
    int ry = i / p.tileInW, rx = i % p.tileInW;
    long long iy = iy0 + ry, ix = ix0 + rx;
    S v = (S)0;
    if (ix >= 0 && ix < p.inW && iy >= 0 && iy < p.inH) v = (S)Cvt<T>::to(xp[iy * p.isH + ix * p.isW]);
    sx[i] = v;
  }
  __syncthreads();
  const int lx = (threadIdx.x % (TW / VX)) * VX, ly = threadIdx.x / (TW / VX);
  const int oy = oy0 + ly;
  if (oy >= p.outH) return;
  S acc[VX];
#pragma unroll
  for (int j = 0; j < VX; j++) acc[j] = (S)0;
  // taps in y: u-row = oy*downy + fy - pady0 must be a multiple of upy
  const int by = oy * p.downy - p.pady0;
  int fy0 = ((-by) % p.upy + p.upy) % p.upy;
  for (int fy = fy0; fy < p.fH; fy += p.upy) {
    const int ry = (by + fy) / p.upy - iy0;   // exact division
    const S* row = sx + ry * p.tileInW;
    const S* frow = sf + fy * p.fW;
#pragma unroll
    for (int j = 0; j < VX; j++) {
      const int bx = (ox0 + lx + j) * p.downx - p.padx0;
      int fx0 = ((-bx) % p.upx + p.upx) % p.upx;
      S a = acc[j];
      for (int fx = fx0; fx < p.fW; fx += p.upx) a += row[(bx + fx) / p.upx - ix0] * frow[fx];
      acc[j] = a;
    }
  }
  T* yp = (T*)p.y + n * p.osN + c * p.osC + (long long)oy * p.osH;
  const int ox = ox0 + lx;
  const bool vec_ok = (p.osW == 1) && (ox + VX <= p.outW) && ((((uintptr_t)(yp + ox)) & (sizeof(T) * VX - 1)) == 0);
  if (vec_ok) {
    struct alignas(sizeof(T) * VX) Pack { T v[VX]; } pk;
#pragma unroll
    for (int j = 0; j < VX; j++) pk.v[j] = Cvt<T>::from(acc[j]);
    *reinterpret_cast<Pack*>(yp + ox) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < VX; j++) if (ox + j < p.outW) yp[(long long)(ox + j) * p.osW] = Cvt<T>::from(acc[j]);
  }
}

// gather kernel for arbitrary strides (channels_last etc.): one output element per thread, fastest index = C.
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_gather_kernel(UpfirdnParams p) {
  typedef typename Acc<T>::type S;
  const long long total = p.outW * p.outH * p.inC * p.inN;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const long long c = t % p.inC; t /= p.inC;
    const int ox = (int)(t % p.outW); t /= p.outW;
    const int oy = (int)(t % p.outH); const long long n = t / p.outH;
    const T* xp = (const T*)p.x + n * p.isN + c * p.isC;
    const int bx = ox * p.downx - p.padx0, by = oy * p.downy - p.pady0;
    const int fx0 = ((-bx) % p.upx + p.upx) % p.upx, fy0 = ((-by) % p.upy + p.upy) % p.upy;
    S acc = (S)0;
    for (int fy = fy0; fy < p.fH; fy += p.upy) {
      const int iy = (by + fy) / p.upy;
      if (iy < 0 || iy >= p.inH) continue;
      for (int fx = fx0; fx < p.fW; fx += p.upx) {
        const int ix = (bx + fx) / p.upx;
        if (ix < 0 || ix >= p.inW) continue;
        const int sy = p.flip ? fy : p.fH - 1 - fy, sxx = p.flip ? fx : p.fW - 1 - fx;
        acc += (S)Cvt<T>::to(xp[iy * p.isH + ix * p.isW]) * (S)p.f[sy * p.fsH + sxx * p.fsW];
      }
    }
    ((T*)p.y)[n * p.osN + c * p.osC + (long long)oy * p.osH + (long long)ox * p.osW] = Cvt<T>::from(acc * (S)p.gain);
  }
}

template <class T>
static int launch_upfirdn(UpfirdnParams& p, cudaStream_t st) {
  typedef typename Acc<T>::type S;
  p.tileInW = ((TW - 1) * p.downx + p.fW - 1) / p.upx + 2;
  p.tileInH = ((TH - 1) * p.downy + p.fH - 1) / p.upy + 2;
  size_t smem = (((size_t)p.fW * p.fH + 3) & ~(size_t)3) * sizeof(S) + (size_t)p.tileInW * p.tileInH * sizeof(S);
  const long long planes = p.inC * p.inN;
  const long long tiles = ((p.outW + TW - 1) / TW) * ((p.outH + TH - 1) / TH);
  if (p.isW == 1 && smem <= 48 * 1024 && tiles * planes <= 0x7fffffffLL) {
    upfirdn2d_tile_kernel<T><<<(unsigned)(tiles * planes), 256, smem, st>>>(p);
  } else {
    upfirdn2d_gather_kernel<T><<<num_sms() * 8, 256, 0, st>>>(p);
  }
  return 0;
}
}  // namespace mgf

extern "C" int mgf_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                             const int64_t inSize[4], const int64_t inStride[4],
                             const int32_t fSize[2], const int64_t fStride[2],
                             const int64_t outSize[4], const int64_t outStride[4],
                             const int32_t up[2], const int32_t down[2], const int32_t pad0[2],
                             int flip, float gain, void* stream) {
  using namespace mgf;
  if (!x || !f || !y) MGF_FAIL(MGF_E_BADARG, "upfirdn2d: null tensor");
  if (up[0] < 1 || up[1] < 1 || down[0] < 1 || down[1] < 1) MGF_FAIL(MGF_E_BADARG, "upfirdn2d: up/down must be >= 1");
  if (fSize[0] < 1 || fSize[1] < 1) MGF_FAIL(MGF_E_SHAPE, "upfirdn2d: empty filter");
  for (int i = 0; i < 4; i++) if (inSize[i] < 0 || outSize[i] < 0) MGF_FAIL(MGF_E_SHAPE, "upfirdn2d: negative size");
  if (outSize[2] != inSize[2] || outSize[3] != inSize[3]) MGF_FAIL(MGF_E_SHAPE, "upfirdn2d: channel/batch mismatch");
  if (outSize[0] < 1 || outSize[1] < 1) MGF_FAIL(MGF_E_SHAPE, "upfirdn2d: output size must be >= 1 (got %lld x %lld)", (long long)outSize[0], (long long)outSize[1]);
  if (inSize[0] * inSize[1] >= (1LL << 31) ) MGF_FAIL(MGF_E_SHAPE, "upfirdn2d: plane too large");
  if (inSize[2] * inSize[3] == 0) return 0;
  UpfirdnParams p;
  p.x = x; p.f = f; p.y = y;
  p.inW = inSize[0]; p.inH = inSize[1]; p.inC = inSize[2]; p.inN = inSize[3];
  p.isW = inStride[0]; p.isH = inStride[1]; p.isC = inStride[2]; p.isN = inStride[3];
  p.outW = outSize[0]; p.outH = outSize[1];
  p.osW = outStride[0]; p.osH = outStride[1]; p.osC = outStride[2]; p.osN = outStride[3];
  p.fW = fSize[0]; p.fH = fSize[1]; p.fsW = fStride[0]; p.fsH = fStride[1];
  p.upx = up[0]; p.upy = up[1]; p.downx = down[0]; p.downy = down[1]; p.padx0 = pad0[0]; p.pady0 = pad0[1];
  p.flip = flip; p.gain = gain; p.tileInW = p.tileInH = 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MGF_F32: launch_upfirdn<float>(p, st); break;
    case MGF_BF16: launch_upfirdn<__nv_bfloat16>(p, st); break;
    case MGF_F16: launch_upfirdn<__half>(p, st); break;
    case MGF_F64: launch_upfirdn<double>(p, st); break;
    default: MGF_FAIL(MGF_E_DTYPE, "upfirdn2d: unsupported dtype %d", dtype);
  }
  MGF_CHECK_LAUNCH("upfirdn2d");
  return 0;
}
