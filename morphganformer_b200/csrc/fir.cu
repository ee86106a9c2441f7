// fir.cu -- the [1,3,3,1] FIR passes of the synthesis engine on NHWC 16-bit tensors, as register sliding windows.
//
// All of these are HBM-bound (upfirdn2d): every kernel here reads each input element from DRAM once and writes each
// output once.  A thread owns one (column, 8-channel vector) and walks DOWN a strip of rows: per input row it loads the few
// horizontally neighbouring pixels (16-byte vectors; the neighbours of the same warp hit L1), forms the horizontal partial sum in fp32
// registers and feeds a ring of vertical accumulators -- no shared memory, no re-reads of rows, 3-4 loads per output instead of 7-16.
//   fir4_kernel       same-resolution 4x4 FIR with zero padding.  Two users:
//                       * second stage of the up-convolution (reference conv2d_resample.py:117-134: conv_transpose2d(stride 2) ->
//                         upfirdn2d(pad 1, gain 4)) with the layer tail fused: + noise * strength, + bias, leaky-ReLU * gain
//                         (networks.py:1036-1040);
//                       * first stage of that layer's input gradient (the adjoint FIR onto the (2h+1)^2 grid).
//   upfir2_add_kernel resnet skip: FIR x2 up-sampling of the 1x1-conv output + residual add (networks.py:245-250, :1157-1160).
//   upfir2_bwd_kernel its adjoint.
#include "common.cuh"

namespace mgf {

// NV-channel (8 or 4) vector load / store: the 4-channel variants halve the registers of a thread (accumulator rings, loaded taps), which
// doubles the resident warps of these latency-bound walkers (mgf_fir_set_mode).
template <bool F16, int NV>
__device__ __forceinline__ void ldv(const uint16_t* p, float (&v)[NV]) {
  if (NV == 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < NV / 2; e++) { const float2 f = unpack16(w[e], F16); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  } else {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const uint32_t w[2] = {u.x, u.y};
#pragma unroll
    for (int e = 0; e < NV / 2; e++) { const float2 f = unpack16(w[e], F16); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  }
}
template <bool F16, int NV>
__device__ __forceinline__ void stv(uint16_t* p, const float (&v)[NV]) {
  if (NV == 8) {
    uint4 o;
    o.x = pack16(v[0], v[1], F16); o.y = pack16(v[2], v[3], F16); o.z = pack16(v[NV - 4], v[NV - 3], F16); o.w = pack16(v[NV - 2], v[NV - 1], F16);
    *reinterpret_cast<uint4*>(p) = o;
  } else {
    uint2 o;
    o.x = pack16(v[0], v[1], F16); o.y = pack16(v[2], v[3], F16);
    *reinterpret_cast<uint2*>(p) = o;
  }
}
// A/B switch (mgf_fir_set_mode): bit 0 fir4, bit 1 upfir2_add, bit 2 upfir2_bwd run with 4 channels per thread instead of 8.
// Measured (8 images, scripts/bench_fir_modes.py, profiles/r02e_fir_modes.txt): upfir2_add is latency-bound and gains from the doubled warp count
// (1024^2 x 32 ch: 0.258 -> 0.228 ms = 5.3 TB/s, 512^2 x 64: 0.136 -> 0.120, equal below 256^2); fir4 is issue-bound and loses (0.295 -> 0.341 ms:
// the per-thread overhead instructions double per byte); upfir2_bwd is a wash (0.153 -> 0.169, 0.090 -> 0.089).  Default: bit 1 only.
static int g_fir_mode = 2;

struct Fir4Args {
  const uint16_t* in; uint16_t* out;
  float fh[4], fv[4];            // horizontal taps (gain folded in) and vertical taps: out[Y,X] = sum_{t,u} fv[t] fh[u] in[Y+off+t, X+off+u]
  int off;                       // -1: second stage of the up-convolution; -2: its adjoint onto the padded gradient grid
  int Hi, Wi, Ho, Wo, Hv, Wv;    // input size, allocated output size, valid output size (rows/columns beyond it are written as zeros)
  int vshift, rows;              // channels / 8 = 1 << vshift; output rows per strip (multiple of 3)
  const float* noise; const float* nstr; long long noise_bstride; const float* bias; int act; float alpha, gain;
  unsigned int* ovf;
};

template <bool IN_F16, bool OUT_F16, bool EPI, int NV>
__global__ void __launch_bounds__(256, (NV == 4) ? 6 : 0) fir4_kernel(const Fir4Args a) {
  const int lv = a.vshift + (NV == 4 ? 1 : 0);          // log2(vectors per pixel)
  const int vecs = 1 << lv;
  const int b = blockIdx.z, Y0 = blockIdx.y * a.rows;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (a.Wo << lv)) return;
  const int cv = i & (vecs - 1), X = i >> lv;
  const int sh = a.vshift + 3;                          // log2(channels)
  // four running row pointers (one per horizontal tap), advanced by one input row per step: the first version recomputed 64-bit
  // addresses per load (~70 of its ~200 instructions per output were address arithmetic in an issue-bound kernel).
  // Out-of-range taps: zero weight on a clamped (valid) address, so the loads stay unconditional.
  const long long in_row = (long long)a.Wi << sh;
  const int r0 = Y0 + a.off;
  const uint16_t* base = a.in + (((long long)b * a.Hi + r0) * a.Wi << sh) + cv * NV;       // row r0 may lie outside the image: only dereferenced when valid
  const float fh0 = a.fh[0], fh1 = a.fh[1], fh2 = a.fh[2], fh3 = a.fh[3];
  const float fv0 = a.fv[0], fv1 = a.fv[1], fv2 = a.fv[2], fv3 = a.fv[3];
  const int x0 = X + a.off, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  const bool colv = X < a.Wv;
  const bool k0 = colv && x0 >= 0 && x0 < a.Wi, k1 = colv && x1 >= 0 && x1 < a.Wi, k2 = colv && x2 >= 0 && x2 < a.Wi, k3 = colv && x3 >= 0 && x3 < a.Wi;
  const float w0 = k0 ? fh0 : 0.f, w1 = k1 ? fh1 : 0.f, w2 = k2 ? fh2 : 0.f, w3 = k3 ? fh3 : 0.f;
  const uint16_t* p0 = base + ((long long)(k0 ? x0 : 0) << sh);
  const uint16_t* p1 = base + ((long long)(k1 ? x1 : 0) << sh);
  const uint16_t* p2 = base + ((long long)(k2 ? x2 : 0) << sh);
  const uint16_t* p3 = base + ((long long)(k3 ? x3 : 0) << sh);
  uint16_t* op = a.out + ((((long long)b * a.Ho + Y0) * a.Wo + X) << sh) + cv * NV;       // output row Y0, advanced per emitted row
  const long long out_row = (long long)a.Wo << sh;
  float bias[NV];
  float nstr = 0.f;
  if (EPI) {
#pragma unroll
    for (int e = 0; e < NV; e++) bias[e] = a.bias ? __ldg(a.bias + cv * NV + e) * a.gain : 0.f;     // act gain folded: lrelu(v) g = lrelu(v g)
    if (a.noise) nstr = (a.nstr ? __ldg(a.nstr) : 1.f) * a.gain;
  }
  const float* nz = (EPI && a.noise) ? a.noise + (long long)b * a.noise_bstride + (long long)Y0 * a.Wo + X : nullptr;
  const float og = EPI ? a.gain : 1.f;       // folded into the last vertical tap and the accumulators' hand-over
  float mx = 0.f;
  float P[NV], Q[NV], S[NV];
#pragma unroll
  for (int e = 0; e < NV; e++) P[e] = Q[e] = S[e] = 0.f;
  int r = r0, Yo = Y0 - 3;

  // one input row: horizontal sum h, then out = OLD + fv3*h is complete (output row Yo), MID += fv2*h, YOUNG += fv1*h, OLD = fv0*h
#define FIR4_STEP(OLD, MID, YOUNG)                                                                                      \
  {                                                                                                                      \
    float h[NV];                                                                                                          \
    if (r >= 0 && r < a.Hi) {                                                                                            \
      float q0[NV], q1[NV], q2[NV], q3[NV];                                                                                  \
      ldv<IN_F16, NV>(p0, q0); ldv<IN_F16, NV>(p1, q1); ldv<IN_F16, NV>(p2, q2); ldv<IN_F16, NV>(p3, q3);                                \
      _Pragma("unroll") for (int e = 0; e < NV; e++) h[e] = fmaf(w3, q3[e], fmaf(w2, q2[e], fmaf(w1, q1[e], w0 * q0[e]))); \
    } else {                                                                                                             \
      _Pragma("unroll") for (int e = 0; e < NV; e++) h[e] = 0.f;                                                          \
    }                                                                                                                    \
    p0 += in_row; p1 += in_row; p2 += in_row; p3 += in_row; r++;                                                         \
    if (Yo >= Y0 && Yo < a.Ho) {                                                                                         \
      float o[NV];                                                                                                        \
      _Pragma("unroll") for (int e = 0; e < NV; e++) o[e] = fmaf(fv3, h[e], OLD[e]);                                      \
      if (Yo >= a.Hv) { _Pragma("unroll") for (int e = 0; e < NV; e++) o[e] = 0.f; }                                      \
      if (EPI) {                                                                                                         \
        const float n = nz ? __ldg(nz) * nstr : 0.f;                                                                     \
        _Pragma("unroll") for (int e = 0; e < NV; e++) {                                                                  \
          const float v = fmaf(o[e], og, n + bias[e]);                                                                   \
          o[e] = (a.act == 1) ? fmaxf(v, v * a.alpha) : v;                                                               \
        }                                                                                                                \
      }                                                                                                                  \
      if (OUT_F16) { _Pragma("unroll") for (int e = 0; e < NV; e++) mx = ovf_max(mx, o[e]); }                        \
      stv<OUT_F16, NV>(op, o);                                                                                               \
      op += out_row; if (EPI && nz) nz += a.Wo;                                                                          \
    }                                                                                                                    \
    Yo++;                                                                                                                \
    _Pragma("unroll") for (int e = 0; e < NV; e++) {                                                                      \
      MID[e] = fmaf(fv2, h[e], MID[e]); YOUNG[e] = fmaf(fv1, h[e], YOUNG[e]); OLD[e] = fv0 * h[e];                       \
    }                                                                                                                    \
  }
  for (int k = 0; k < a.rows + 3; k += 3) {
    if (Yo >= a.Ho) break;
    FIR4_STEP(P, Q, S)
    FIR4_STEP(Q, S, P)
    FIR4_STEP(S, P, Q)
  }
#undef FIR4_STEP
  if (OUT_F16) ovf_commit(a.ovf, mx);
}

// ---- resnet skip: out[b,Y,X,c] = add[b,Y,X,c] + g * sum_{fy,fx} fk[fy] fk[fx] v[b,(Y+fy-2)/2,(X+fx-2)/2,c] over the taps with even index
// (reference Conv2dLayer.forward :245-250 -> conv2d_resample 1x1-up branch -> upfirdn2d(up=2, pad=[2,1,2,1], gain=4) -> bias_act gain).
// Polyphase form: rows 2m, 2m+1 and columns 2n, 2n+1 of the output read the 3x3 low-resolution neighbourhood of (m, n):
//   hE[m] = f0 v[m,n-1] + f2 v[m,n]    hO[m] = f1 v[m,n] + f3 v[m,n+1]
//   out[2m] = f0 h[m-1] + f2 h[m]      out[2m+1] = f1 h[m] + f3 h[m+1]
// A thread owns (output column, 8-channel vector) and walks down the low-resolution rows: 2 loads of v, 2 loads of add, 2 stores per step.
struct Upfir2Args {
  const uint16_t* v; const uint16_t* add; uint16_t* out;
  float fh[4], fv[4];            // horizontal taps with the gain folded in, vertical taps
  int h, w, vshift, rows;        // low-resolution size, channels / 8 = 1 << vshift, low-resolution rows per strip
  unsigned int* ovf;
};

template <bool F16, int NV>
__global__ void __launch_bounds__(256, (NV == 4) ? 6 : 0) upfir2_add_kernel2(const Upfir2Args a) {
  // thread = (output column X, 8-channel vector), walking down the low-resolution rows of a strip: per row 2 loads of v (the two
  // low-resolution pixels column X reads), then two output rows (one add load + one store each).  One column per thread keeps the
  // kernel at ~70 registers (the first version carried both columns of a low-resolution pixel: 126 registers, 25 % occupancy,
  // latency-bound at 3.9 TB/s).
  const int lv = a.vshift + (NV == 4 ? 1 : 0);          // log2(vectors per pixel)
  const int vecs = 1 << lv;
  const int b = blockIdx.z, m0 = blockIdx.y * a.rows;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((2 * a.w) << lv)) return;
  const int cv = i & (vecs - 1), X = i >> lv, n = X >> 1;
  const int sh = a.vshift + 3;
  // X even: f0 v[n-1] + f2 v[n];  X odd: f1 v[n] + f3 v[n+1]
  const bool odd = X & 1;
  const int na = odd ? n : n - 1, nb = odd ? n + 1 : n;
  const float wa = (na >= 0) ? (odd ? a.fh[1] : a.fh[0]) : 0.f, wb = (nb < a.w) ? (odd ? a.fh[3] : a.fh[2]) : 0.f;
  const int oa = (na >= 0 ? na : 0) << sh, ob_ = (nb < a.w ? nb : a.w - 1) << sh;
  // running pointers (one add per row instead of 64-bit index arithmetic per access)
  const long long vrow = (long long)a.w << sh;
  const uint16_t* pa = a.v + ((((long long)b * a.h + (m0 - 1)) * a.w) << sh) + cv * NV + oa;      // row m0 - 1 may lie outside: dereferenced only when valid
  const uint16_t* pb = pa - oa + ob_;
  const long long W2 = 2LL * a.w;
  const long long obase = ((((long long)b * 2 * a.h + 2 * m0) * W2 + X) << sh) + cv * NV;
  const uint16_t* ab = a.add ? a.add + obase : nullptr;
  uint16_t* ob = a.out + obase;
  const long long orow = W2 << sh;
  float mx = 0.f;
  float h[3][NV];                             // ring of horizontal sums: rows m-1, m, m+1
  int rin = m0 - 1;                          // low-resolution row the pointers stand on

  auto hrow = [&](float (&h_)[NV]) {          // horizontal sum of row `rin`, then step down
    if (rin >= 0 && rin < a.h) {
      float p[NV], q[NV];
      ldv<F16, NV>(pa, p); ldv<F16, NV>(pb, q);
#pragma unroll
      for (int e = 0; e < NV; e++) h_[e] = fmaf(wa, p[e], wb * q[e]);
    } else {
#pragma unroll
      for (int e = 0; e < NV; e++) h_[e] = 0.f;
    }
    pa += vrow; pb += vrow; rin++;
  };
  auto emit = [&](const float (&ta)[NV], float ca, const float (&tb)[NV], float cb) {      // next output row
    float o[NV];
#pragma unroll
    for (int e = 0; e < NV; e++) o[e] = fmaf(ca, ta[e], cb * tb[e]);
    if (ab) {
      float x0[NV];
      ldv<F16, NV>(ab, x0);
      ab += orow;
#pragma unroll
      for (int e = 0; e < NV; e++) o[e] += x0[e];
    }
    if (F16) {
#pragma unroll
      for (int e = 0; e < NV; e++) mx = ovf_max(mx, o[e]);
    }
    stv<F16, NV>(ob, o);
    ob += orow;
  };
  const float fv0 = a.fv[0], fv1 = a.fv[1], fv2 = a.fv[2], fv3 = a.fv[3];
  hrow(h[0]);
  hrow(h[1]);
#define UPFIR2_STEP(PREV, CUR, NEXT, m)                                        \
  if ((m) < a.h && (m) < m0 + a.rows) {                                        \
    hrow(h[NEXT]);                                                             \
    emit(h[PREV], fv0, h[CUR], fv2);                                           \
    emit(h[CUR], fv1, h[NEXT], fv3);                                           \
  }
  for (int m = m0; m < m0 + a.rows && m < a.h; m += 3) {
    UPFIR2_STEP(0, 1, 2, m)
    UPFIR2_STEP(1, 2, 0, m + 1)
    UPFIR2_STEP(2, 0, 1, m + 2)
  }
#undef UPFIR2_STEP
  if (F16) ovf_commit(a.ovf, mx);
}

// ---- adjoint of the skip FIR (without the add): dv[b,m,n,c] = g * sum_{fy,fx} fk[fy] fk[fx] dout[b, 2m+2-fy, 2n+2-fx, c]   (bf16 gradients)
// H[Y] = sum_fx fk[fx] dout[Y, 2n+2-fx] (4 loads per row); dv[m] = f3 H[2m-1] + f2 H[2m] + f1 H[2m+1] + f0 H[2m+2]: two new rows per step.
struct Upfir2BwdArgs {
  const uint16_t* dout; uint16_t* dv;
  float fh[4], fv[4];
  int h, w, vshift, rows;
};

template <int NV>
__global__ void __launch_bounds__(256, (NV == 4) ? 5 : 0) upfir2_bwd_kernel2(const Upfir2BwdArgs a) {
  const int lv = a.vshift + (NV == 4 ? 1 : 0);          // log2(vectors per pixel)
  const int vecs = 1 << lv;
  const int b = blockIdx.z, m0 = blockIdx.y * a.rows;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (a.w << lv)) return;
  const int cv = i & (vecs - 1), n = i >> lv;
  const int sh = a.vshift + 3;
  const int H2 = 2 * a.h, W2 = 2 * a.w;
  float wx[4]; int xo[4];
#pragma unroll
  for (int fx = 0; fx < 4; fx++) {
    const int X = 2 * n + 2 - fx;
    const bool ok = X >= 0 && X < W2;
    wx[fx] = ok ? a.fh[fx] : 0.f; xo[fx] = (ok ? X : 0) << sh;
  }
  const long long drow = (long long)W2 << sh;
  int Yin = 2 * m0 - 1;                      // dout row the pointers stand on (may start at -1: dereferenced only when valid)
  const uint16_t* base = a.dout + ((((long long)b * H2 + Yin) * W2) << sh) + cv * NV;
  const uint16_t* p0 = base + xo[0];
  const uint16_t* p1 = base + xo[1];
  const uint16_t* p2 = base + xo[2];
  const uint16_t* p3 = base + xo[3];
  uint16_t* op = a.dv + ((((long long)b * a.h + m0) * a.w + n) << sh) + cv * NV;
  const long long orow = (long long)a.w << sh;
  const float fv0 = a.fv[0], fv1 = a.fv[1], fv2 = a.fv[2], fv3 = a.fv[3];
  float Ha[NV], Hb[NV], Hc[NV], Hd[NV];          // H[2m-1], H[2m], H[2m+1], H[2m+2]
  auto hrow = [&](float (&h_)[NV]) {          // horizontal sum of dout row Yin, then step down
    if (Yin >= 0 && Yin < H2) {
      float q0[NV], q1[NV], q2[NV], q3[NV];
      ldv<false, NV>(p0, q0); ldv<false, NV>(p1, q1); ldv<false, NV>(p2, q2); ldv<false, NV>(p3, q3);
#pragma unroll
      for (int e = 0; e < NV; e++) h_[e] = fmaf(wx[3], q3[e], fmaf(wx[2], q2[e], fmaf(wx[1], q1[e], wx[0] * q0[e])));
    } else {
#pragma unroll
      for (int e = 0; e < NV; e++) h_[e] = 0.f;
    }
    p0 += drow; p1 += drow; p2 += drow; p3 += drow; Yin++;
  };
  hrow(Ha);
  hrow(Hb);
  // two steps per iteration so that the ring is renamed instead of moved: (Ha, Hb, Hc, Hd) -> (Hc, Hd, Ha, Hb)
  for (int m = m0; m < m0 + a.rows && m < a.h; m += 2) {
    hrow(Hc);
    hrow(Hd);
    {
      float o[NV];
#pragma unroll
      for (int e = 0; e < NV; e++) o[e] = fmaf(fv3, Ha[e], fmaf(fv2, Hb[e], fmaf(fv1, Hc[e], fv0 * Hd[e])));
      stv<false, NV>(op, o);
      op += orow;
    }
    if (m + 1 < m0 + a.rows && m + 1 < a.h) {
      hrow(Ha);
      hrow(Hb);
      float o[NV];
#pragma unroll
      for (int e = 0; e < NV; e++) o[e] = fmaf(fv3, Hc[e], fmaf(fv2, Hd[e], fmaf(fv1, Ha[e], fv0 * Hb[e])));
      stv<false, NV>(op, o);
      op += orow;
    }
  }
}

static inline int log2_exact_(int v) { int s = 0; while ((1 << s) < v) s++; return (1 << s) == v ? s : -1; }

// rows per strip: long strips amortise the 3-row warm-up, but small images still have to fill 148 SMs
static inline int strip_rows(int out_rows, long long ctas_per_row_strip, int unit, int max_rows) {
  int rows = max_rows;
  while (rows > unit && ((out_rows + rows - 1) / rows) * ctas_per_row_strip < 2LL * num_sms()) rows -= unit;
  return rows;
}
}  // namespace mgf

using namespace mgf;

extern "C" int mgf_fir_set_mode(int bits) { g_fir_mode = bits; return 0; }
extern "C" int mgf_fir_get_mode() { return g_fir_mode; }

extern "C" int mgf_fir4(const void* in, void* out, const float* fk4, float gain, int off, int B, int Hi, int Wi, int Ho, int Wo, int Hv, int Wv,
                        int C, int in_fwd, int out_fwd, const float* noise, const float* nstr, int64_t noise_bstride, const float* bias, int act,
                        float alpha, float act_gain, void* stream) {
  if (!in || !out || !fk4) MGF_FAIL(MGF_E_BADARG, "fir4: null tensor");
  const int vs = (C % 8) ? -1 : log2_exact_(C / 8);
  if (vs < 0) MGF_FAIL(MGF_E_SHAPE, "fir4: C/8 must be a power of two (C = %d)", C);
  if (B <= 0 || B > 65535 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || Hv > Ho || Wv > Wo) MGF_FAIL(MGF_E_SHAPE, "fir4: bad sizes");
  if (off != -1 && off != -2) MGF_FAIL(MGF_E_BADARG, "fir4: off must be -1 (up-convolution second stage) or -2 (its adjoint)");
  Fir4Args a;
  a.in = (const uint16_t*)in; a.out = (uint16_t*)out;
  for (int t = 0; t < 4; t++) { a.fh[t] = fk4[t] * gain; a.fv[t] = fk4[t]; }
  a.off = off; a.Hi = Hi; a.Wi = Wi; a.Ho = Ho; a.Wo = Wo; a.Hv = Hv; a.Wv = Wv; a.vshift = vs;
  const bool epi = noise || bias || act || act_gain != 1.f;
  a.noise = noise; a.nstr = nstr; a.noise_bstride = noise_bstride; a.bias = bias; a.act = act; a.alpha = alpha; a.gain = act_gain;
  const bool if16 = in_fwd && fwd_f16(), of16 = out_fwd && fwd_f16();
  a.ovf = of16 ? overflow_flag() : nullptr;
  const bool half = g_fir_mode & 1;                      // 4 channels per thread
  const int items = Wo << (vs + (half ? 1 : 0));
  const int bx = (items + 255) / 256;
  a.rows = strip_rows(Ho, (long long)bx * B, 3, 30);
  dim3 grid(bx, (Ho + a.rows - 1) / a.rows, B);
  cudaStream_t st = (cudaStream_t)stream;
#define MGF_FIR4(I, O, E) do { if (half) fir4_kernel<I, O, E, 4><<<grid, 256, 0, st>>>(a); else fir4_kernel<I, O, E, 8><<<grid, 256, 0, st>>>(a); } while (0)
  if (if16 && of16) { if (epi) MGF_FIR4(true, true, true); else MGF_FIR4(true, true, false); }
  else if (!if16 && !of16) { if (epi) MGF_FIR4(false, false, true); else MGF_FIR4(false, false, false); }
  else MGF_FAIL(MGF_E_UNSUP, "fir4: input and output must have the same 16-bit type");
#undef MGF_FIR4
  MGF_CHECK_LAUNCH("fir4");
  return 0;
}

extern "C" int mgf_fir4_pad(const void* dy, void* g, const float* fk4, float gain, int B, int H, int W, int C, void* stream) {
  // g [B,H+2,W+2,C] = FIR of dy [B,H,W,C] onto the (H+1) x (W+1) grid, zero padding 2, last row / column zero (bf16 gradients)
  const float flipped[4] = {fk4[3], fk4[2], fk4[1], fk4[0]};
  return mgf_fir4(dy, g, flipped, gain, -2, B, H, W, H + 2, W + 2, H + 1, W + 1, C, 0, 0, nullptr, nullptr, 0, nullptr, 0, 0.f, 1.f, stream);
}

extern "C" int mgf_upfir2_add(const void* v, const void* add, void* out, const float* fk4, float gain, int B, int h, int w, int C, void* stream) {
  if (!v || !out || !fk4) MGF_FAIL(MGF_E_BADARG, "upfir2_add: null tensor");
  const int vs = (C % 8) ? -1 : log2_exact_(C / 8);
  if (vs < 0) MGF_FAIL(MGF_E_SHAPE, "upfir2_add: C/8 must be a power of two");
  if (B <= 0 || B > 65535 || h <= 0 || w <= 0) MGF_FAIL(MGF_E_SHAPE, "upfir2_add: bad image size");
  Upfir2Args a;
  a.v = (const uint16_t*)v; a.add = (const uint16_t*)add; a.out = (uint16_t*)out;
  for (int t = 0; t < 4; t++) { a.fh[t] = fk4[t] * gain; a.fv[t] = fk4[t]; }
  a.h = h; a.w = w; a.vshift = vs;
  a.ovf = fwd_f16() ? overflow_flag() : nullptr;
  const bool half = g_fir_mode & 2;                      // 4 channels per thread
  const int bx = (((2 * w) << (vs + (half ? 1 : 0))) + 255) / 256;
  a.rows = strip_rows(h, (long long)bx * B, 3, 15);
  dim3 grid(bx, (h + a.rows - 1) / a.rows, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (fwd_f16()) { if (half) upfir2_add_kernel2<true, 4><<<grid, 256, 0, st>>>(a); else upfir2_add_kernel2<true, 8><<<grid, 256, 0, st>>>(a); }
  else { if (half) upfir2_add_kernel2<false, 4><<<grid, 256, 0, st>>>(a); else upfir2_add_kernel2<false, 8><<<grid, 256, 0, st>>>(a); }
  MGF_CHECK_LAUNCH("upfir2_add");
  return 0;
}

extern "C" int mgf_upfir2_bwd(const void* dout, void* dv, const float* fk4, float gain, int B, int h, int w, int C, void* stream) {
  if (!dout || !dv || !fk4) MGF_FAIL(MGF_E_BADARG, "upfir2_bwd: null tensor");
  const int vs = (C % 8) ? -1 : log2_exact_(C / 8);
  if (vs < 0) MGF_FAIL(MGF_E_SHAPE, "upfir2_bwd: C/8 must be a power of two");
  if (B <= 0 || B > 65535 || h <= 0 || w <= 0) MGF_FAIL(MGF_E_SHAPE, "upfir2_bwd: bad image size");
  Upfir2BwdArgs a;
  a.dout = (const uint16_t*)dout; a.dv = (uint16_t*)dv;
  for (int t = 0; t < 4; t++) { a.fh[t] = fk4[t] * gain; a.fv[t] = fk4[t]; }
  a.h = h; a.w = w; a.vshift = vs;
  const bool half = g_fir_mode & 4;                      // 4 channels per thread
  const int bx = ((w << (vs + (half ? 1 : 0))) + 255) / 256;
  a.rows = strip_rows(h, (long long)bx * B, 1, 16);
  dim3 grid(bx, (h + a.rows - 1) / a.rows, B);
  if (half) upfir2_bwd_kernel2<4><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else upfir2_bwd_kernel2<8><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  MGF_CHECK_LAUNCH("upfir2_bwd");
  return 0;
}
