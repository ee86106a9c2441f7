// vgg_first.cu -- first layer of the LPIPS-VGG16 trunk fused with the LPIPS input stage, forward and backward.
//
// Replaces, per projection step (reference lpips/networks_basic.py:94-101 ScalingLayer + pretrained_networks.py:97-135 conv1_1 + ReLU,
// and the MSE term of 1024_example_percept_MSE.py:143):
//   forward   img fp32 NCHW -> (x - shift)/scale -> conv3x3(3 -> 64) + bias + ReLU -> h0 [B,R,R,64] 16-bit NHWC,  mse[b] += |img - target|^2
//   backward  g = d(loss)/d(conv1_1 pre-activation) [B,R,R,64] bf16 -> d(img) fp32 NCHW  (+ the MSE gradient)
// The first build ran these as im2col (a [B,R,R,32] 16-bit buffer, 537 MB per 8 images at 1024^2) + a K = 32 GEMM on the tcgen05
// conv kernel + col2im: four kernels moving 4.7 GB.  Here the 27-wide patch is gathered from a shared-memory image tile straight into
// mma.sync.m16n8k16 A fragments (forward), and the 64 -> 27 product of the backward lands in a shared-memory tile that the col2im
// gather reads in place, so the only HBM traffic is the image, the 64-channel tensor once, and the image gradient.
// K = 27 (forward) / N = 27 (backward) is far too small for a 128-row tcgen05 tile to pay; the layer is HBM-bound.
#include "common.cuh"

namespace mgf {
namespace {

__constant__ float v_shift[3] = {-.030f, -.088f, -.188f};
__constant__ float v_scale[3] = {.458f, .448f, .450f};

template <bool F16>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 16-bit pack without the fp16 range clamp of common.cuh: the scaled image is within +-3 and a 27-term dot product of it cannot leave
// the fp16 range for any sane weights; ReLU is applied on the packed pair
template <bool F16> __device__ __forceinline__ uint32_t pk_nc(float a, float b) {
  if (F16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  return pack_bf16(a, b);
}
template <bool F16> __device__ __forceinline__ uint32_t pk_relu(float a, float b) {
  if (F16) { __half2 h = __hmax2(__floats2half2_rn(a, b), __float2half2_rn(0.f)); return *reinterpret_cast<uint32_t*>(&h); }
  __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(a, b), __float2bfloat162_rn(0.f)); return *reinterpret_cast<uint32_t*>(&h);
}
// output column 2t+e of n-tile nt = jj*4+m is channel (jj*4+t)*8 + 2m + e, so a lane ends up with 8 consecutive channels per jj;
// the B fragment of column n = g therefore belongs to channel perm_chan(nt, g)
__device__ __forceinline__ int perm_chan(int nt, int g) { return ((nt >> 2) * 4 + (g >> 1)) * 8 + 2 * (nt & 3) + (g & 1); }

// ------------------------------------------------------------------------------------------------------------------- forward
constexpr int FTH = 8, FTW = 64, FSW = FTW + 2 + 2, FPL = (FTH + 2) * FSW;     // tile 8 x 64 pixels, padded row 68, plane 680 floats

template <bool F16>
__global__ void __launch_bounds__(256, 2) vgg_conv1_fwd_kernel(const float* __restrict__ img, const float* __restrict__ target, float* mse,
                                                            const float* __restrict__ W, const float* __restrict__ bias, uint16_t* out, int R, int ntiles) {
  __shared__ float sx[3 * FPL + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const long long HW = (long long)R * R;
  const int tiles_x = (R + FTW - 1) / FTW, tiles_y = (R + FTH - 1) / FTH;
  // ---- weights as B fragments (k = patch index (ky*3+kx)*3+c, zero above 26), bias for this lane's 16 channels
  uint32_t bw[2][8][2];
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      // patch columns 27..31 do not exist: their weights are forced to zero here, so the A fragment may hold any finite value there
      const float* wr = W + perm_chan(nt, g) * 32;
      const int k0 = 16 * s + 2 * t;
      bw[s][nt][0] = pack16(k0 < 27 ? wr[k0] : 0.f, k0 + 1 < 27 ? wr[k0 + 1] : 0.f, F16);
      bw[s][nt][1] = pack16(k0 + 8 < 27 ? wr[k0 + 8] : 0.f, k0 + 9 < 27 ? wr[k0 + 9] : 0.f, F16);
    }
  __shared__ float sbias[64];            // bias stays in shared memory (the prefetch registers need the room)
  if (threadIdx.x < 64) sbias[threadIdx.x] = bias[threadIdx.x];
  // patch offsets of this lane's 8 k indices: {2t, 2t+1, 2t+8, 2t+9} + 16 s
  int koff[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int k = 16 * (i >> 2) + 2 * t + 8 * ((i >> 1) & 1) + (i & 1);
    const int tap = k / 3, c = k - tap * 3;
    koff[i] = (k < 27) ? c * FPL + (tap / 3) * FSW + (tap % 3) : 0;      // k >= 27: any finite tile value (its weight is zero)
  }
  float mse_local = 0.f;
  int mse_b = -1;
  // The haloed image tile (3 channels x 10 rows x 66 columns = 1980 floats, FPT per thread) and the target values of its interior are
  // fetched into registers ONE TILE AHEAD: all of a thread's loads are in flight together and their HBM latency is covered by the MMA
  // phase of the current tile.  (Fetching in place -- a row loop of dependent load -> store trips per warp, then a second loop for the MSE --
  // serialised ~7 DRAM latencies per tile: the kernel ran at 0.3 of the HBM roofline, profiles/r01e_secondary_kernels_ncu.md.)
  constexpr int FNE = 3 * (FTH + 2) * (FTW + 2), FPT = (FNE + 255) / 256;
  float vi[FPT], vt[FPT];
  auto fetch = [&](long long tile) {
    const int b = (int)(tile / (tiles_x * tiles_y)), tr = (int)(tile % (tiles_x * tiles_y));
    const int Y0 = (tr / tiles_x) * FTH, X0 = (tr % tiles_x) * FTW;
    const float* ib = img + (long long)b * 3 * HW;
    const float* tb = target ? target + (long long)b * 3 * HW : nullptr;
#pragma unroll
    for (int k = 0; k < FPT; k++) {
      const int e = threadIdx.x + 256 * k;
      const int c = e / ((FTH + 2) * (FTW + 2)), rem = e - c * ((FTH + 2) * (FTW + 2));
      const int hy = rem / (FTW + 2), hx = rem - hy * (FTW + 2);
      const int y = Y0 - 1 + hy, x = X0 - 1 + hx;
      const bool in = e < FNE && y >= 0 && y < R && x >= 0 && x < R;
      const bool interior = in && hy >= 1 && hy <= FTH && hx >= 1 && hx <= FTW;
      const long long o = c * HW + (long long)y * R + x;
      vi[k] = in ? __ldg(ib + o) : 0.f;
      vt[k] = (interior && tb) ? __ldg(tb + o) : 0.f;           // two independent predicated loads: nothing here may wait on vi[k] (the
                                                                // consumer decides whether the pair enters the MSE sum)
    }
  };
  if ((long long)blockIdx.x < (long long)ntiles) fetch(blockIdx.x);
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < (long long)ntiles; tile += gridDim.x) {
  const int b = (int)(tile / (tiles_x * tiles_y)), tr = (int)(tile % (tiles_x * tiles_y));
  const int Y0 = (tr / tiles_x) * FTH, X0 = (tr % tiles_x) * FTW;
  if (target && b != mse_b) {          // flush the MSE partial sum when the image changes (tiles of one image are contiguous)
    if (mse_b >= 0) { const float tot = warp_sum(mse_local); if (lane == 0 && tot != 0.f) atomicAdd(&mse[mse_b], tot); }
    mse_local = 0.f; mse_b = b;
  }
  // ---- registers -> scaled tile in shared memory (zero outside the image = the conv's zero padding of the SCALED input) + MSE partial sum
#pragma unroll
  for (int k = 0; k < FPT; k++) {
    const int e = threadIdx.x + 256 * k;
    const int c = e / ((FTH + 2) * (FTW + 2)), rem = e - c * ((FTH + 2) * (FTW + 2));
    const int hy = rem / (FTW + 2), hx = rem - hy * (FTW + 2);
    const int y = Y0 - 1 + hy, x = X0 - 1 + hx;
    if (e < FNE) {
      const bool in = y >= 0 && y < R && x >= 0 && x < R;
      sx[c * FPL + hy * FSW + hx] = in ? (vi[k] - v_shift[c]) * (1.f / v_scale[c]) : 0.f;
      const bool interior = in && hy >= 1 && hy <= FTH && hx >= 1 && hx <= FTW;
      const float d = (interior && target) ? vi[k] - vt[k] : 0.f;
      mse_local = fmaf(d, d, mse_local);
    }
  }
  __syncthreads();
  if (tile + gridDim.x < (long long)ntiles) fetch(tile + gridDim.x);      // in flight during the MMA phase below
  const int y = Y0 + warp;
#pragma unroll 1
  for (int cg = 0; cg < FTW / 16; cg++) {
    const int xl0 = cg * 16 + g, xl1 = xl0 + 8;
    if (y >= R || X0 + cg * 16 >= R) break;
    const float* p0 = sx + warp * FSW + xl0;           // patch origin (halo coordinates: pixel (y, x) sits at [y - Y0 + 1][x - X0 + 1])
    const float* p1 = sx + warp * FSW + xl1;
    float v0[8], v1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { v0[i] = p0[koff[i]]; v1[i] = p1[koff[i]]; }
    uint32_t a[2][4];
#pragma unroll
    for (int s = 0; s < 2; s++) {
      a[s][0] = pk_nc<F16>(v0[4 * s], v0[4 * s + 1]); a[s][1] = pk_nc<F16>(v1[4 * s], v1[4 * s + 1]);
      a[s][2] = pk_nc<F16>(v0[4 * s + 2], v0[4 * s + 3]); a[s][3] = pk_nc<F16>(v1[4 * s + 2], v1[4 * s + 3]);
    }
    const bool ok0 = X0 + xl0 < R, ok1 = X0 + xl1 < R;
    uint16_t* o0 = out + (((long long)b * R + y) * R + X0 + xl0) * 64;
    uint16_t* o1 = out + (((long long)b * R + y) * R + X0 + xl1) * 64;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      uint32_t w0[4], w1[4];
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const float2 bq = *reinterpret_cast<const float2*>(sbias + (jj * 4 + t) * 8 + 2 * m);
        float acc[4] = {bq.x, bq.y, bq.x, bq.y};
        mma16816<F16>(acc, a[0][0], a[0][1], a[0][2], a[0][3], bw[0][jj * 4 + m][0], bw[0][jj * 4 + m][1]);
        mma16816<F16>(acc, a[1][0], a[1][1], a[1][2], a[1][3], bw[1][jj * 4 + m][0], bw[1][jj * 4 + m][1]);
        w0[m] = pk_relu<F16>(acc[0], acc[1]); w1[m] = pk_relu<F16>(acc[2], acc[3]);
      }
      if (ok0) *reinterpret_cast<uint4*>(o0 + (jj * 4 + t) * 8) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
      if (ok1) *reinterpret_cast<uint4*>(o1 + (jj * 4 + t) * 8) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    }
  }
  __syncthreads();                     // the tile buffer is reused by the next tile
  }
  if (target && mse_b >= 0) { const float tot = warp_sum(mse_local); if (lane == 0 && tot != 0.f) atomicAdd(&mse[mse_b], tot); }
}

// ------------------------------------------------------------------------------------------------------------------ backward
constexpr int BTH = 8, BTW = 32, BHW = BTW + 2, BNP = (BTH + 2) * BHW;       // 8 x 32 output pixels, 340 haloed pixels
constexpr int BMT = (BNP + 15) / 16, BDS = 37;   // 22 M-tiles; dcol row stride (floats): odd -> conflict-free gather, 5g+2t -> near conflict-free scatter

__global__ void __launch_bounds__(256, 2) vgg_conv1_bwd_kernel(const uint16_t* __restrict__ gy, const float* __restrict__ W, const float* __restrict__ img,
                                                            const float* __restrict__ target, float mcoef, float* dimg, int R, int ntiles) {
  extern __shared__ __align__(16) float sd[];             // [BMT * 16][BDS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const long long HW = (long long)R * R;
  const int tiles_x = (R + BTW - 1) / BTW, tiles_y = (R + BTH - 1) / BTH;
  // B fragments: k = output channel o of conv1_1 (K-permuted like the A vectors), n = patch index
  uint32_t bw[4][4][2];
#pragma unroll
  for (int jj = 0; jj < 2; jj++)
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int ch = (jj * 4 + t) * 8 + 4 * h, n = nt * 8 + g;
        bw[jj * 2 + h][nt][0] = pack_bf16(W[ch * 32 + n], W[(ch + 1) * 32 + n]);
        bw[jj * 2 + h][nt][1] = pack_bf16(W[(ch + 2) * 32 + n], W[(ch + 3) * 32 + n]);
      }
  // The gradient vectors of the tile's 340 haloed pixels (3 M-tiles of 16 pixels per warp, 4 x 16 bytes per lane each) and the image / target
  // values of this thread's output pixel are fetched into registers for the NEXT tile right after the MMAs of the current one have consumed
  // them: one exposed DRAM latency per tile instead of four (three dependent M-tile rounds + the MSE-gradient loads), and that one is
  // covered by the col2im phase and the other resident CTA.
  constexpr int MPW = (BMT + 7) / 8;          // M-tiles per warp
  uint4 pva[MPW][2], pvb[MPW][2];
  float pim[3] = {0.f, 0.f, 0.f}, ptg[3] = {0.f, 0.f, 0.f};
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  auto fetch = [&](long long tile) {
    const int b = (int)(tile / (tiles_x * tiles_y)), tr = (int)(tile % (tiles_x * tiles_y));
    const int Y0 = (tr / tiles_x) * BTH, X0 = (tr % tiles_x) * BTW;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < MPW; i++) {
      const int mt = warp + 8 * i;
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      const int y0 = Y0 - 1 + i0 / BHW, x0 = X0 - 1 + i0 % BHW, y1 = Y0 - 1 + i1 / BHW, x1 = X0 - 1 + i1 % BHW;
      const bool ok0 = mt < BMT && i0 < BNP && y0 >= 0 && y0 < R && x0 >= 0 && x0 < R, ok1 = mt < BMT && i1 < BNP && y1 >= 0 && y1 < R && x1 >= 0 && x1 < R;
      const uint4* r0 = reinterpret_cast<const uint4*>(gy + (((long long)b * R + (ok0 ? y0 : 0)) * R + (ok0 ? x0 : 0)) * 64);
      const uint4* r1 = reinterpret_cast<const uint4*>(gy + (((long long)b * R + (ok1 ? y1 : 0)) * R + (ok1 ? x1 : 0)) * 64);
#pragma unroll
      for (int jj = 0; jj < 2; jj++) { pva[i][jj] = ok0 ? __ldg(r0 + jj * 4 + t) : z; pvb[i][jj] = ok1 ? __ldg(r1 + jj * 4 + t) : z; }
    }
    const int Y = Y0 + ty, X = X0 + tx;
    if (target && Y < R && X < R) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const long long o = ((long long)b * 3 + c) * HW + (long long)Y * R + X;
        pim[c] = __ldg(img + o); ptg[c] = __ldg(target + o);
      }
    }
  };
  if ((long long)blockIdx.x < (long long)ntiles) fetch(blockIdx.x);
#pragma unroll 1
  for (long long tile = blockIdx.x; tile < (long long)ntiles; tile += gridDim.x) {
  const int b = (int)(tile / (tiles_x * tiles_y)), tr = (int)(tile % (tiles_x * tiles_y));
  const int Y0 = (tr / tiles_x) * BTH, X0 = (tr % tiles_x) * BTW;
#pragma unroll
  for (int i = 0; i < MPW; i++) {
    const int mt = warp + 8 * i;
    if (mt < BMT) {
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; nt++) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int jj = 0; jj < 2; jj++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          mma16816<false>(acc[nt], pva[i][jj].x, pvb[i][jj].x, pva[i][jj].y, pvb[i][jj].y, bw[jj * 2][nt][0], bw[jj * 2][nt][1]);
          mma16816<false>(acc[nt], pva[i][jj].z, pvb[i][jj].z, pva[i][jj].w, pvb[i][jj].w, bw[jj * 2 + 1][nt][0], bw[jj * 2 + 1][nt][1]);
        }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        sd[i0 * BDS + nt * 8 + 2 * t] = acc[nt][0]; sd[i0 * BDS + nt * 8 + 2 * t + 1] = acc[nt][1];
        sd[i1 * BDS + nt * 8 + 2 * t] = acc[nt][2]; sd[i1 * BDS + nt * 8 + 2 * t + 1] = acc[nt][3];
      }
    }
  }
  const float cim[3] = {pim[0], pim[1], pim[2]}, ctg[3] = {ptg[0], ptg[1], ptg[2]};
  if (tile + gridDim.x < (long long)ntiles) fetch(tile + gridDim.x);
  __syncthreads();
  // ---- col2im gather: d(xs)[c](Y, X) = sum_taps dcol[(Y - (ky-1), X - (kx-1))][tap*3 + c]
  const int Y = Y0 + ty, X = X0 + tx;
  if (Y < R && X < R) {
    float gs[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; tap++) {
      const int slot = (ty + 2 - tap / 3) * BHW + (tx + 2 - tap % 3);
#pragma unroll
      for (int c = 0; c < 3; c++) gs[c] += sd[slot * BDS + tap * 3 + c];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const long long o = ((long long)b * 3 + c) * HW + (long long)Y * R + X;
      float v = gs[c] / v_scale[c];
      if (target) v = fmaf(mcoef, cim[c] - ctg[c], v);
      dimg[o] = v;
    }
  }
  __syncthreads();                     // sd is rewritten by the next tile
  }
}

}  // namespace
}  // namespace mgf

using namespace mgf;

// img [B,3,R,R] fp32 -> out [B,R,R,64] (forward 16-bit type) = relu(conv3x3((img - shift)/scale, W) + bias); W [64][32] fp32 with
// column (ky*3+kx)*3+c (columns 27..31 ignored).  target/mse optional: mse[b] += sum (img - target)^2.
extern "C" int mgf_vgg_conv1_fwd(const float* img, const float* target, float* mse, const float* W, const float* bias, void* out,
                                 int B, int R, void* stream) {
  if (!img || !W || !bias || !out || (target && !mse)) MGF_FAIL(MGF_E_BADARG, "vgg_conv1_fwd: null tensor");
  if (B <= 0 || R <= 0) MGF_FAIL(MGF_E_SHAPE, "vgg_conv1_fwd: empty input");
  const long long nt = (long long)((R + FTW - 1) / FTW) * ((R + FTH - 1) / FTH) * B;
  if (nt > 0x7fffffffLL) MGF_FAIL(MGF_E_SHAPE, "vgg_conv1_fwd: too many tiles");
  const long long cap = (long long)num_sms() * 2;          // persistent: weights / offsets are set up once per CTA
  const unsigned grid = (unsigned)(nt < cap ? nt : cap);
  if (fwd_f16()) vgg_conv1_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(img, target, mse, W, bias, (uint16_t*)out, R, (int)nt);
  else vgg_conv1_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(img, target, mse, W, bias, (uint16_t*)out, R, (int)nt);
  MGF_CHECK_LAUNCH("vgg_conv1_fwd");
  return 0;
}

// gy [B,R,R,64] bf16 = gradient wrt the conv1_1 pre-activation -> dimg [B,3,R,R] fp32 = conv^T(gy, W)/scale + mcoef * (img - target)
extern "C" int mgf_vgg_conv1_bwd(const void* gy, const float* W, const float* img, const float* target, float mcoef, float* dimg,
                                 int B, int R, void* stream) {
  if (!gy || !W || !dimg || (target && !img)) MGF_FAIL(MGF_E_BADARG, "vgg_conv1_bwd: null tensor");
  if (B <= 0 || R <= 0) MGF_FAIL(MGF_E_SHAPE, "vgg_conv1_bwd: empty input");
  const long long nt = (long long)((R + BTW - 1) / BTW) * ((R + BTH - 1) / BTH) * B;
  if (nt > 0x7fffffffLL) MGF_FAIL(MGF_E_SHAPE, "vgg_conv1_bwd: too many tiles");
  const int smem = BMT * 16 * BDS * (int)sizeof(float);
  static bool cfg = false;
  if (!cfg) { cudaFuncSetAttribute(vgg_conv1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); cfg = true; }
  const long long cap = (long long)num_sms() * 2;
  const unsigned grid = (unsigned)(nt < cap ? nt : cap);
  vgg_conv1_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const uint16_t*)gy, W, img, target, mcoef, dimg, R, (int)nt);
  MGF_CHECK_LAUNCH("vgg_conv1_bwd");
  return 0;
}
