// bias_act.cu -- fused bias + activation + gain + clamp, forward and 1st/2nd-order gradient forms.
// Semantics follow the reference op (torch_utils/ops/bias_act.cu:15-139, bias_act.py:15-25); the kernel is a
// fresh sm_100a design: 16-byte vector loads/stores, grid sized in waves of the SM count, activation and
// gradient order resolved at compile time, fp32 internal math for 16-bit types.
// HBM-bound: algorithmic bytes = (#present input tensors + 1) * sizeX * sizeof(T).
#include "common.cuh"

namespace mgf {

struct BiasActParams {
  const void *x, *b, *xref, *yref, *dy; void* y;
  float alpha, gain, clamp; long long sizeX, sizeB, stepB;
};

template <int A, int G, class S>
__device__ __forceinline__ S act_eval(S x, S xref, S yy, S alpha) {
  const S one = (S)1, two = (S)2, range = (S)80, half_range = (S)40;
  const S selu_s = (S)1.0507009873554804934193349852946, selu_a = (S)1.6732632423543772848170429916717;
  if (A == 1) return (G <= 1) ? x : (S)0;
  if (A == 2) { if (G == 0) return x > 0 ? x : (S)0; if (G == 1) return yy > 0 ? x : (S)0; return (S)0; }
  if (A == 3) { if (G == 0) return x > 0 ? x : x * alpha; if (G == 1) return yy > 0 ? x : x * alpha; return (S)0; }
  if (A == 4) {
    if (G == 0) { S c = exp(x), d = one / c; return x < -range ? -one : (x > range ? one : (c - d) / (c + d)); }
    if (G == 1) return x * (one - yy * yy);
    return x * (one - yy * yy) * (-two * yy);
  }
  if (A == 5) {
    if (G == 0) return x < -range ? (S)0 : one / (exp(-x) + one);
    if (G == 1) return x * yy * (one - yy);
    return x * yy * (one - yy) * (one - two * yy);
  }
  if (A == 6) {
    if (G == 0) return x >= 0 ? x : exp(x) - one;
    if (G == 1) return yy >= 0 ? x : x * (yy + one);
    return yy >= 0 ? (S)0 : x * (yy + one);
  }
  if (A == 7) {
    if (G == 0) return x >= 0 ? selu_s * x : (selu_s * selu_a) * (exp(x) - one);
    if (G == 1) return yy >= 0 ? x * selu_s : x * (yy + selu_s * selu_a);
    return yy >= 0 ? (S)0 : x * (yy + selu_s * selu_a);
  }
  if (A == 8) {
    if (G == 0) return x > range ? x : log(exp(x) + one);
    if (G == 1) return x * (one - exp(-yy));
    S c = exp(-yy); return x * c * (one - c);
  }
  if (A == 9) {
    if (G == 0) return x < -range ? (S)0 : x / (exp(-x) + one);
    S c = exp(xref), d = c + one;
    if (G == 1) return xref > half_range ? x : x * c * (xref + d) / (d * d);
    return xref > half_range ? (S)0 : x * c * (xref * (two - d) + two * d) / (d * d * d);
  }
  return (S)0;
}

template <class T, int A, int G>
__device__ __forceinline__ T bias_act_one(T xv, typename Acc<T>::type b, T xrefv, T yrefv, T dyv, bool has_dy,
                                          typename Acc<T>::type alpha, typename Acc<T>::type gain, typename Acc<T>::type clamp) {
  typedef typename Acc<T>::type S;
  S x = (S)Cvt<T>::to(xv), xref = (S)Cvt<T>::to(xrefv), yref = (S)Cvt<T>::to(yrefv);
  S dy = has_dy ? (S)Cvt<T>::to(dyv) : (S)1;
  S yy = (gain != 0) ? yref / gain : (S)0;
  if (G == 0) x += b; else xref += b;
  S y = act_eval<A, G, S>(x, xref, yy, alpha);
  if (A == 9 && G > 0) yref = (xref < (S)-80) ? (S)0 : xref / (exp(-xref) + (S)1) * gain;
  y *= gain * dy;
  if (clamp >= 0) {
    if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0 ? clamp : -clamp);
    else y = (yref > -clamp && yref < clamp) ? y : (S)0;
  }
  return Cvt<T>::from(y);
}

template <class T, int A, int G>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActParams p) {
  typedef typename Acc<T>::type S;
  constexpr int V = Vec16<T>::N;
  const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
  const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = (const T*)p.xref;
  const T* yr = (const T*)p.yref; const T* dy = (const T*)p.dy; T* y = (T*)p.y;
  const long long nvec = p.sizeX / V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool bias_per_vec = (p.stepB % V) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vec16<T> vx = reinterpret_cast<const Vec16<T>*>(x)[i];
    Vec16<T> vxr, vyr, vdy, vo;
    if (xr) vxr = reinterpret_cast<const Vec16<T>*>(xr)[i];
    if (yr) vyr = reinterpret_cast<const Vec16<T>*>(yr)[i];
    if (dy) vdy = reinterpret_cast<const Vec16<T>*>(dy)[i];
    const long long e0 = i * V;
    S bv = (S)0;
    if (b && bias_per_vec) bv = (S)Cvt<T>::to(b[(e0 / p.stepB) % p.sizeB]);
#pragma unroll
    for (int j = 0; j < V; j++) {
      if (b && !bias_per_vec) bv = (S)Cvt<T>::to(b[((e0 + j) / p.stepB) % p.sizeB]);
      T z = Cvt<T>::from((S)0);
      vo.v[j] = bias_act_one<T, A, G>(vx.v[j], bv, xr ? vxr.v[j] : z, yr ? vyr.v[j] : z, dy ? vdy.v[j] : z, dy != nullptr, alpha, gain, clamp);
    }
    reinterpret_cast<Vec16<T>*>(y)[i] = vo;
  }
  // tail (sizeX not a multiple of the vector width)
  for (long long e = nvec * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < p.sizeX; e += stride) {
    S bv = b ? (S)Cvt<T>::to(b[(e / p.stepB) % p.sizeB]) : (S)0;
    T z = Cvt<T>::from((S)0);
    y[e] = bias_act_one<T, A, G>(x[e], bv, xr ? xr[e] : z, yr ? yr[e] : z, dy ? dy[e] : z, dy != nullptr, alpha, gain, clamp);
  }
}

template <class T, int A>
static void* pick_grad(int grad) {
  switch (grad) {
    case 0: return (void*)bias_act_kernel<T, A, 0>;
    case 1: return (void*)bias_act_kernel<T, A, 1>;
    default: return (void*)bias_act_kernel<T, A, 2>;
  }
}
template <class T>
static void* pick_act(int act, int grad) {
  switch (act) {
    case 1: return pick_grad<T, 1>(grad); case 2: return pick_grad<T, 2>(grad); case 3: return pick_grad<T, 3>(grad);
    case 4: return pick_grad<T, 4>(grad); case 5: return pick_grad<T, 5>(grad); case 6: return pick_grad<T, 6>(grad);
    case 7: return pick_grad<T, 7>(grad); case 8: return pick_grad<T, 8>(grad); case 9: return pick_grad<T, 9>(grad);
  }
  return nullptr;
}
}  // namespace mgf

extern "C" int mgf_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                            int dtype, int grad, int act, float alpha, float gain, float clamp,
                            int64_t sizeX, int64_t sizeB, int64_t stepB, void* stream) {
  using namespace mgf;
  if (!x || !y) MGF_FAIL(MGF_E_BADARG, "bias_act: x and y must be non-null");
  if (grad < 0 || grad > 2) MGF_FAIL(MGF_E_BADARG, "bias_act: grad must be 0, 1 or 2 (got %d)", grad);
  if (act < 1 || act > 9) MGF_FAIL(MGF_E_BADARG, "bias_act: act index %d outside 1..9", act);
  if (sizeX < 0 || (b && (sizeB <= 0 || stepB <= 0))) MGF_FAIL(MGF_E_SHAPE, "bias_act: bad sizes");
  if (sizeX == 0) return 0;
  const void* ptrs[6] = {x, b ? x : x, xref, yref, dy, y};
  for (int i = 0; i < 6; i++) if (ptrs[i] && ((uintptr_t)ptrs[i] & 15)) MGF_FAIL(MGF_E_ALIGN, "bias_act: tensors must be 16-byte aligned");
  void* fn = nullptr; int esize = 4;
  switch (dtype) {
    case MGF_F32: fn = pick_act<float>(act, grad); esize = 4; break;
    case MGF_BF16: fn = pick_act<__nv_bfloat16>(act, grad); esize = 2; break;
    case MGF_F16: fn = pick_act<__half>(act, grad); esize = 2; break;
    case MGF_F64: fn = pick_act<double>(act, grad); esize = 8; break;
    default: MGF_FAIL(MGF_E_DTYPE, "bias_act: unsupported dtype %d", dtype);
  }
  BiasActParams p{x, b, xref, yref, dy, y, alpha, gain, clamp, (long long)sizeX, (long long)(b ? sizeB : 1), (long long)(b ? stepB : 1)};
  const long long nvec = (sizeX + (16 / esize) - 1) / (16 / esize);
  long long blocks = (nvec + 255) / 256;
  const long long cap = (long long)num_sms() * 16;   // 16 waves' worth of 256-thread CTAs per SM at most; grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  void* args[] = {&p};
  cudaError_t e = cudaLaunchKernel(fn, dim3((unsigned)blocks), dim3(256), args, 0, (cudaStream_t)stream);
  if (e != cudaSuccess) MGF_FAIL((int)e, "bias_act: %s", cudaGetErrorString(e));
  MGF_CHECK_LAUNCH("bias_act");
  return 0;
}
