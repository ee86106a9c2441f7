// elementwise.cu -- small HBM-bound helpers for the host mirror (fma = a*b+c with the broadcasts networks.py:322 uses).
#include "common.cuh"
namespace mgf {
template <class T>
__global__ void __launch_bounds__(256) fma_kernel(const T* a, const T* b, const T* c, T* out, long long N, long long C, long long HW, int bmode, int cmode) {
  const long long total = N * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long nc = i / HW, hw = i % HW;
    float av = (float)Cvt<T>::to(a[i]);
    float bv = (float)Cvt<T>::to(bmode == 1 ? b[nc] : b[i]);
    float cv = cmode == 0 ? 0.f : (float)Cvt<T>::to(cmode == 2 ? c[hw] : c[i]);
    out[i] = Cvt<T>::from(fmaf(av, bv, cv));
  }
}
}  // namespace mgf
extern "C" int mgf_fma(const void* a, const void* b, const void* c, void* out, int dtype,
                       int64_t N, int64_t C, int64_t HW, int bmode, int cmode, void* stream) {
  using namespace mgf;
  if (!a || !b || !out || (cmode && !c)) MGF_FAIL(MGF_E_BADARG, "fma: null tensor");
  const long long total = N * C * HW; if (total == 0) return 0;
  long long blocks = (total + 255) / 256; const long long cap = (long long)num_sms() * 16; if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MGF_F32: fma_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)a, (const float*)b, (const float*)c, (float*)out, N, C, HW, bmode, cmode); break;
    case MGF_BF16: fma_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)c, (__nv_bfloat16*)out, N, C, HW, bmode, cmode); break;
    default: MGF_FAIL(MGF_E_DTYPE, "fma: unsupported dtype %d", dtype);
  }
  MGF_CHECK_LAUNCH("fma");
  return 0;
}
