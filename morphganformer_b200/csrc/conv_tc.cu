// conv_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate).
//
// One kernel covers every dense contraction of the hot path:
//   * 3x3 stride-1 modulated conv (per-sample weight tiles with the style and demodulation folded in),
//   * the up-sampling conv (reference: conv_transpose2d stride 2 -> (2H+1)^2 -> 4x4 FIR, conv2d_resample.py:117-134)
//     as four output phases of 3x3 taps on the low-res grid (the FIR is folded into the per-phase weights, so the
//     (2H+1)^2 intermediate never exists),
//   * the resnet skip (1x1 conv + FIR up2) in the same form, the VGG16 convs of LPIPS, and every input gradient
//     (dgrad = the same contraction with transposed weights; for the up conv 36 taps over four strided phase views).
//
// GEMM view: D[128 pixels, BN out-channels] += A[128 pixels, BK in-channels] * B[BN, BK]^T, looped over
// (channel chunk, tap).  A tiles are TMA boxes {BK, TW, TH, TB} of an NHWC activation tensor shifted by the tap
// offset -- out-of-bounds rows/cols are zero-filled by TMA, which is exactly the conv zero padding.  B tiles are
// TMA boxes {BK, BN, 1} of the [G][T][NT][K] weight tensor.  Both land in shared memory in the 128B (or 64B for
// 32-channel layers) swizzled K-major layout that tcgen05.mma reads through shared-memory descriptors.
//
// CTA = 10 warps, persistent over output tiles:
//   warp 8 (one lane): TMA producer, STAGES-deep mbarrier ring (up to 20 stages so small tiles keep enough bytes in flight)
//   warp 9 (one lane): tcgen05.mma issuer, accumulators double-buffered in TMEM (2 x BN columns)
//   warps 0-3 / 4-7  : two epilogue groups, group g drains accumulator stage g (alternate tiles) -- tcgen05.ld the
//                      accumulator, then the fused per-layer tail:
//                      * per-(sample, channel) scale           (dgrad: the style s[b,i])
//                      * d(styles) partial sums                (sum over pixels of acc * X, warp-shuffle transpose-reduce + atomics)
//                      * + noise * strength, + bias, leaky-ReLU/ReLU * gain     (networks.py:1036-1040, bias_act)
//                      * + residual tensor                     (resnet add, networks.py:1160)
//                      * activation-gradient mask by the saved output X   (backward of the previous layer's bias_act)
//                      and 16-byte bf16 stores.  The epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>
#include <string.h>

namespace mgf {
namespace tc {

constexpr int MAX_TAPS = 40;
constexpr int MAX_AMAPS = 4;
constexpr int SMEM_BUDGET = 192 * 1024;     // operand stages; + 2 x 16 KB epilogue staging + barriers stays under 227 KB
constexpr int STG_BYTES = 128 * 128;        // one epilogue group's output staging tile: 128 pixels x 64 channels x 2 B
constexpr int RACC = 256;                   // floats per epilogue group for the per-CTA d(style) accumulation (BN <= 256)

struct Tap { int8_t amap, dy, dx, pad; int32_t wz; };

struct alignas(64) Params {
  CUtensorMap amap[MAX_AMAPS];
  CUtensorMap bmap;
  CUtensorMap omap[4];          // output tensor, one map per output phase (TMA store of the staged tile)
  CUtensorMap xmap;             // saved activation X (dgrad epilogues), same geometry as omap[0]: TMA-loaded into the staging tile
  Tap taps[MAX_TAPS];
  int ntaps, kchunks;
  int NB, GH, GW, TB, TH, TW, tilesB, tilesH, tilesW, rows;
  int NT, Cout, n_tiles, per_sample, w_T;
  void* out; long long OH, OW, OC; int osy, osx; int ofy[4], ofx[4];
  const float* scale_n; float* reduce_out; const __nv_bfloat16* X;
  const float* noise; const float* noise_strength; const float* bias;
  int act; float alpha, gain;
  const __nv_bfloat16* add;
  int actgrad; float ag_alpha, ag_gain;
  uint32_t tx_bytes;
  int total_tiles;
  uint32_t idesc; int out_f16, x_f16, add_f16;
  int tiles_hw;      // halo kernel: tiles per image (tilesW * tilesH)
  long long noise_bstride;       // elements between per-sample noise planes (0: one plane for all samples)
  int superpix;                  // output rows are PAIRS of pixels (32+32 channels): noise differs between the two 32-column halves
  unsigned int* ovf;             // fp16 overflow flag word (nullptr unless the output is an fp16 forward tensor)
  int ph_taps;                   // per-phase tap lists: phase ph owns taps [ph_tap0[ph], ph_tap0[ph+1]); weights are shared by the phases
  int ph_tap0[5]; int ph_tiles;  // ph_tiles = tiles of one phase (tile order is phase-major so the persistent CTAs stay balanced)
  int x_tma; uint32_t x_bytes;   // X tile arrives by TMA (one 64/32-channel group per tile) instead of per-thread strided loads
  int dbg;                       // experiment switches (scripts/bench_halo.py; 0 in production): 1 = epilogue only drains TMEM, 2 = no MMAs, 4 = no activation loads, 8 = no output store, 16 = no staging either
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }
// one lane of a fully converged warp (keeps the surrounding loop warp-uniform, so the compiler stays on the uniform datapath
// instead of wrapping every descriptor move in an ELECT/R2UR.BROADCAST retry loop, as it does inside an `if (lane == 0)` branch)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- CTA-pair (cta_group::2) wrappers: two SMs of one TPC execute one M = 256 MMA, each supplying its own 128 rows of A and HALF of B
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address: "the same location in CTA 0 of the pair"
__device__ __forceinline__ void tma_load_4d_2cta(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  // executed by both CTAs: the data lands in the issuing CTA's shared memory, the transaction bytes are counted on CTA 0's barrier
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar) {      // arrives on `bar` in BOTH CTAs once the MMAs issued so far have completed
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {    // arrive on the barrier at the same offset in CTA 0 of the pair
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major shared-memory operand descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30)
// (=1, unused for swizzled K-major), SBO>>4 [32,46) = 8 rows * row bytes, version=1 [46,48), layout [61,64) (2 = 128B, 4 = 64B swizzle).
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  constexpr uint64_t row_bytes = BK * 2;
  constexpr uint64_t sbo = 8 * row_bytes;
  constexpr uint64_t layout = (BK == 64) ? 2 : 4;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}

template <int BN, int BK, int NG_ = 2>
struct Cfg {
  static constexpr int NG = NG_;                          // epilogue groups = accumulator stages in TMEM
  static constexpr int THREADS = 64 + 128 * NG;
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTG = (BN <= 64) ? 2 : 1;        // staging tiles per epilogue group (double-buffered when a tile is one group)
  static constexpr int STAGES_RAW = (SMEM_BUDGET - (NG * NSTG - 2) * STG_BYTES - (NG - 2) * 2048) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 20 ? 20 : STAGES_RAW;
  static constexpr int BAR_BYTES = (2 * STAGES + 4 * NG) * 8 + 16;
  static constexpr int SMEM = STAGES * STAGE + NG * NSTG * STG_BYTES + NG * RACC * 4 + BAR_BYTES;
  static constexpr uint32_t TMEM_COLS = NG * BN <= 32 ? 32 : NG * BN <= 64 ? 64 : NG * BN <= 128 ? 128 : NG * BN <= 256 ? 256 : 512;
  static_assert(NG * BN <= 512, "accumulator stages must fit the 512 TMEM columns");
  static_assert(SMEM <= 227 * 1024, "shared-memory carve-up exceeds the 227 KB a CTA may opt into");
  // instruction descriptor (InstrDescriptor): D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major A/B, N>>3 at 17, M>>4 at 24
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};

struct TileCoord { int n0, x0, y0, b0; };
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int tile, int BN) {
  TileCoord t;
  int nb, m;
  if (p.ph_taps) {       // phase-major order: n0 = ph * Cout + (column block inside the phase)
    const int ph = tile / p.ph_tiles, rem = tile % p.ph_tiles, nbp = p.Cout / BN;
    nb = ph * nbp + rem % nbp; m = rem / nbp;
  } else {
    nb = tile % p.n_tiles; m = tile / p.n_tiles;
  }
  t.n0 = nb * BN;
  t.x0 = (m % p.tilesW) * p.TW; m /= p.tilesW;
  t.y0 = (m % p.tilesH) * p.TH; m /= p.tilesH;
  t.b0 = m * p.TB;
  return t;
}

// Fused layer tail for one 128 x BN accumulator tile (called by all four warps of an epilogue group after the tfull wait).
// Row `valid`/(x, y, b) identify this thread's pixel; tacc = TMEM address of the tile (lane quarter already applied).
template <int BN, int NSTG, int GW32>
__device__ __forceinline__ void epilogue_tile(const Params& p, const TileCoord& t, int x, int y, int b, bool valid, uint32_t tacc, int lane,
                                              float nz0, float nz1, float nz2, float nz3,
                                              uint8_t* stg, int group, int r, float* racc, uint64_t* xbar, uint32_t xphase, uint64_t* tempty_bar,
                                              bool tempty_on_cta0 = false) {
      // A tile may span several output phases (BN > Cout, e.g. the four parities of an up-convolution in one 128- or 256-column tile, so
      // the activation tiles are fetched once for all of them): phase and channel offset are per 32-column chunk; a staged group of
      // GW32 * 32 columns never straddles two phases (host check).  nz0..nz3: noise of this thread's pixel in the tile's phases
      // (super-pixel rows: nz0 / nz1 = left / right pixel of the pair).
      const int phase_lo = t.n0 / p.Cout;
      float ovf_mx = 0.f;
      if (p.dbg & 1) {                 // experiment: drain the accumulator and hand it back, nothing else (which stage bounds the tile rate?)
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) { uint32_t raw[32]; tmem_ld32(tacc + (uint32_t)(c * 32), raw); }
        tc_fence_before();
        if (tempty_on_cta0) mbar_arrive_cta0(tempty_bar); else mbar_arrive(tempty_bar);
        if (p.X && (p.reduce_out || p.actgrad) && p.x_tma) mbar_wait(xbar, xphase);
        return;
      }
      // phase / channel offset of the current 32-column chunk, advanced incrementally (a division per chunk showed up in an issue-bound tail)
      int phase_idx = phase_lo, co = t.n0 - phase_lo * p.Cout - 32;
#pragma unroll 1
      for (int c = 0; c < BN / 32; c++) {
        co += 32;
        if (co >= p.Cout) { co -= p.Cout; phase_idx++; }
        // element offset of this chunk's 32 output channels of this pixel (only the per-thread X / residual loads need it)
        const long long obase = (p.add || (p.X && !p.x_tma))
            ? ((((long long)b * p.OH + ((long long)y * p.osy + p.ofy[phase_idx])) * p.OW + ((long long)x * p.osx + p.ofx[phase_idx])) * p.OC + co) : 0;
        uint32_t raw[32];
        tmem_ld32(tacc + (uint32_t)(c * 32), raw);
        if (c == BN / 32 - 1) {          // the accumulator now lives in registers: hand the TMEM stage back to the MMA issuer at once
          tc_fence_before();
          if (tempty_on_cta0) mbar_arrive_cta0(tempty_bar); else mbar_arrive(tempty_bar);
        }
        // rows outside the grid are never stored (the TMA store clips them); their accumulators may hold anything (NaN included), so they are
        // zeroed before they can reach the column reductions -- a rarely taken branch instead of 32 selects per chunk for every thread
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(raw[j]);
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = 0.f;
        }
        float xv[32];
        const bool needX = (p.X != nullptr) && (p.reduce_out != nullptr || p.actgrad);
        constexpr int GX32 = GW32;
        if (needX && p.x_tma) {
          // the X tile of this (single-group) output tile was TMA-loaded into the staging buffer while the MMAs were running:
          // read this thread's row (same swizzle as the store path); the outputs later overwrite exactly the slots read here.
          if (c == 0) mbar_wait(xbar, xphase);
          const uint8_t* row = stg + r * (GX32 * 64);
          uint32_t xw[16];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int j = (c % GX32) * 4 + q;
            const int pos = (GX32 == 2) ? (j ^ (r & 7)) : (j ^ ((r >> 1) & 3));
            const uint4 u = *reinterpret_cast<const uint4*>(row + pos * 16);
            xw[q * 4] = u.x; xw[q * 4 + 1] = u.y; xw[q * 4 + 2] = u.z; xw[q * 4 + 3] = u.w;
          }
          // warp-uniform branch on the storage type instead of both conversions and a select per pair
          if (p.x_f16) {
#pragma unroll
            for (int e = 0; e < 16; e++) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xw[e])); xv[2 * e] = f.x; xv[2 * e + 1] = f.y; }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e++) { const float2 f = unpack_bf16(xw[e]); xv[2 * e] = f.x; xv[2 * e + 1] = f.y; }
          }
          if (!valid) {          // rows past the tile's last pixel hold stale staging data
#pragma unroll
            for (int j = 0; j < 32; j++) xv[j] = 0.f;
          }
        } else if (needX) {
          if (valid) {
            const uint4* xp = reinterpret_cast<const uint4*>(p.X + obase);
            uint32_t xw[16];
#pragma unroll
            for (int q = 0; q < 4; q++) { const uint4 u = __ldg(xp + q); xw[q * 4] = u.x; xw[q * 4 + 1] = u.y; xw[q * 4 + 2] = u.z; xw[q * 4 + 3] = u.w; }
            if (p.x_f16) {
#pragma unroll
              for (int e = 0; e < 16; e++) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xw[e])); xv[2 * e] = f.x; xv[2 * e + 1] = f.y; }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++) { const float2 f = unpack_bf16(xw[e]); xv[2 * e] = f.x; xv[2 * e + 1] = f.y; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) xv[j] = 0.f;
          }
        }
        if (p.reduce_out) {
          // column sums over the 128 rows of acc * X: warp transpose-reduce (31 shuffles), then one atomic per column per warp.
          // (Keeping per-thread partial sums in registers across the tiles of a (sample, N-block) would take this off the per-tile path,
          // but BN = 64 needs 64 more registers and 10 warps cap a thread at 168: ptxas spilled, and setmaxnreg did not lift its limit.)
          float s[32];
#pragma unroll
          for (int j = 0; j < 32; j++) s[j] = v[j] * xv[j];
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool hi = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; i++) {
              const float send = hi ? s[i] : s[i + o];
              const float keep = hi ? s[i + o] : s[i];
              s[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          // lane L now holds the sum of column L over this warp's 32 rows (all rows of a tile share one sample: TB == 1).
          // Accumulate per CTA in shared memory; global atomics only when the CTA moves to another (sample, N-block): at 1024^2
          // that is 148 x BN global atomics per launch instead of 8.4 M onto the same few hundred addresses.
          atomicAdd(racc + c * 32 + lane, s[0]);
        }
        if (p.scale_n) {
          const float4* sp = reinterpret_cast<const float4*>(p.scale_n + (long long)(valid ? b : 0) * p.NT + t.n0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; j++) { const float4 f = __ldg(sp + j); v[4 * j] *= f.x; v[4 * j + 1] *= f.y; v[4 * j + 2] *= f.z; v[4 * j + 3] *= f.w; }
        }
        // layer tail (networks.py:1036-1040): (v + noise + bias) -> leaky-ReLU -> * gain, computed as t = v g + (bias g + noise g) and
        // max(t, alpha t): 4 instructions per element instead of 6 (gain > 0 and 0 <= alpha <= 1, host-checked)
        float nzg = 0.f;
        if (p.noise) {
          const int k = p.superpix ? (c & 1) : phase_idx - phase_lo;  // super-pixel rows: odd 32-column chunk = right pixel of the pair
          nzg = (k == 0 ? nz0 : k == 1 ? nz1 : k == 2 ? nz2 : nz3) * p.gain;
        }
        if (p.bias) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + co);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 f = __ldg(bp + j);
            v[4 * j] = fmaf(v[4 * j], p.gain, fmaf(f.x, p.gain, nzg)); v[4 * j + 1] = fmaf(v[4 * j + 1], p.gain, fmaf(f.y, p.gain, nzg));
            v[4 * j + 2] = fmaf(v[4 * j + 2], p.gain, fmaf(f.z, p.gain, nzg)); v[4 * j + 3] = fmaf(v[4 * j + 3], p.gain, fmaf(f.w, p.gain, nzg));
          }
        } else if (p.noise || p.gain != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaf(v[j], p.gain, nzg);
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], v[j] * p.alpha);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.add && valid) {
          const uint4* ap = reinterpret_cast<const uint4*>(p.add + obase);
#pragma unroll
          uint32_t aw[16];
#pragma unroll
          for (int q = 0; q < 4; q++) { const uint4 u = __ldg(ap + q); aw[q * 4] = u.x; aw[q * 4 + 1] = u.y; aw[q * 4 + 2] = u.z; aw[q * 4 + 3] = u.w; }
          if (p.add_f16) {
#pragma unroll
            for (int e = 0; e < 16; e++) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&aw[e])); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e++) { const float2 f = unpack_bf16(aw[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
          }
        }
        if (p.actgrad) {
          const float gpos = p.ag_gain, gneg = p.ag_alpha * p.ag_gain;
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] *= (xv[j] > 0.f ? gpos : gneg);
        }
        // stage this thread's 32 outputs (64 B) into the swizzled shared-memory tile; a 64-channel group (or the whole tile for
        // BN = 32) then leaves with ONE TMA store: fully coalesced, asynchronous, clipped at the image border by the hardware.
        // GW32 = 32-column chunks per staged group (2: 128-byte rows, 1: 64-byte rows).  Tiles wider than one group with GW32 == 1
        // (32-channel phases inside a 128-column tile) alternate between the two 8 KB halves of the staging tile.
        const int h = c % GW32;
        constexpr bool SPLIT = (GW32 == 1 && BN > 32);
        uint8_t* stg_c = SPLIT ? stg + (c & 1) * (STG_BYTES / 2) : stg;
        if (p.dbg & 16) continue;
        if (h == 0 && (!p.x_tma || SPLIT)) {               // the previous store must have finished reading the staging tile
          if (r == 0) { if (SPLIT) tma_store_wait_read<1>(); else tma_store_wait_read<NSTG - 1>(); }   // (x_tma: the caller already did this before loading X into it)
          group_sync(group);
        }
        if (p.ovf) {
#pragma unroll
          for (int j = 0; j < 32; j++) ovf_mx = ovf_max(ovf_mx, v[j]);
        }
        {
          uint8_t* row = stg_c + r * (GW32 * 64);
          uint32_t pw[16];
          if (p.out_f16) {
#pragma unroll
            for (int e = 0; e < 16; e++) pw[e] = pack_f16_sat(v[2 * e], v[2 * e + 1]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; e++) pw[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint4 u = make_uint4(pw[q * 4], pw[q * 4 + 1], pw[q * 4 + 2], pw[q * 4 + 3]);
            const int j = h * 4 + q;
            const int pos = (GW32 == 2) ? (j ^ (r & 7)) : (j ^ ((r >> 1) & 3));     // 128B / 64B swizzle, as the tensor map expects
            *reinterpret_cast<uint4*>(row + pos * 16) = u;
          }
        }
        if (h == GW32 - 1) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          group_sync(group);
          if (r == 0 && !(p.dbg & 8)) tma_store_4d(&p.omap[phase_idx], stg_c, co - h * 32, t.x0, t.y0, t.b0);
        }
      }
      ovf_commit(p.ovf, valid ? ovf_mx : 0.f);      // rows outside the tile hold whatever the accumulator had
}

template <int BN>
__device__ __forceinline__ void flush_reduce(const Params& p, float* racc, int key, int group, int r) {
  group_sync(group);
  const long long b0 = key / p.n_tiles; const int n0 = (key % p.n_tiles) * BN;
  for (int j = r; j < BN; j += 128) {
    const float v = racc[j];
    if (v != 0.f) atomicAdd(p.reduce_out + b0 * p.NT + n0 + j, v);
    racc[j] = 0.f;
  }
  group_sync(group);
}

template <int BN, int BK, bool G32 = false, int NG_ = 2>     // G32: staged output groups are 32 columns wide (32-channel phases inside a wider tile); NG_: epilogue groups
__global__ void __launch_bounds__(64 + 128 * NG_, 1) conv_tc_kernel(const __grid_constant__ Params p) {
  using C = Cfg<BN, BK, NG_>;
  constexpr int NG = NG_, W_PROD = 4 * NG_, W_MMA = 4 * NG_ + 1;
  constexpr int GW32 = (BN >= 64 && !G32) ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];                      // swizzled tiles need 1024-byte aligned bases
  uint8_t* stg_base = smem + C::STAGES * C::STAGE;                       // 2 x 16 KB epilogue staging, 1024-byte aligned
  float* racc_base = reinterpret_cast<float*>(stg_base + NG * C::NSTG * STG_BYTES); // NG x 256 floats: per-group d(style) partial sums
  uint64_t* full = reinterpret_cast<uint64_t*>(racc_base + NG * RACC);
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + NG;
  uint64_t* xbar = tempty + NG;                // [group][staging buffer]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2 * NG);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < NG; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    for (int s = 0; s < 2 * NG; s++) mbar_init(&xbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W_PROD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int iters = p.kchunks * p.ntaps;

  if (warp == W_PROD) {     // ------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile, BN);
      const int wbase = p.per_sample ? t.b0 * p.w_T : 0;
      int tp0 = 0, tp1 = p.ntaps, wn0 = t.n0;
      if (p.ph_taps) { const int ph = t.n0 / p.Cout; tp0 = p.ph_tap0[ph]; tp1 = p.ph_tap0[ph + 1]; wn0 = t.n0 - ph * p.Cout; }
      for (int kc = 0; kc < p.kchunks; kc++) {
        for (int tp = tp0; tp < tp1; tp++) {
          const Tap tap = p.taps[tp];
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], p.tx_bytes);
            uint8_t* sa = smem + stage * C::STAGE;
            tma_load_4d(&p.amap[tap.amap], &full[stage], sa, kc * BK, t.x0 + tap.dx, t.y0 + tap.dy, t.b0);
            tma_load_3d(&p.bmap, &full[stage], sa + C::A_BYTES, kc * BK, wn0, wbase + tap.wz);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {   // ------------------------------------------------- MMA issuer (whole warp loops, one lane issues)
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
      int iters_t = iters;
      if (p.ph_taps) { const int ph = tile / p.ph_tiles; iters_t = p.kchunks * (p.ph_tap0[ph + 1] - p.ph_tap0[ph]); }
      for (int it = 0; it < iters_t; it++) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE);
          const uint64_t adesc = make_desc<BK>(sa), bdesc = make_desc<BK>(sa + C::A_BYTES);
          if (it > 0) {
#pragma unroll
            for (int k = 0; k < BK / 16; k++) tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; k++) tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, k > 0 ? 1u : 0u);
          }
          tc_commit(&empty[stage]);           // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&tfull[as]); // accumulator complete -> epilogue
      __syncwarp();
      if (++as == NG) { as = 0; aphase ^= 1; }
    }
  } else {                 // ------------------------------------------------ epilogue: two groups of 4 warps, group g owns accumulator stage g
    const int as = warp >> 2; uint32_t aphase = 0;
    const int wq = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = wq * 32 + lane;            // accumulator row == TMEM lane == pixel index inside the A box
    const int tx = r % p.TW, ty = (r / p.TW) % p.TH, tb = r / (p.TW * p.TH);
    const float nstr = (p.noise && p.noise_strength) ? *p.noise_strength : 1.f;
    float* racc = racc_base + as * RACC;
    int red_key = -1;
    if (p.reduce_out) { for (int j = r; j < RACC; j += 128) racc[j] = 0.f; group_sync(as); }
    uint32_t xph0 = 0, xph1 = 0; bool x_first = true;
    for (int tile = blockIdx.x + as * gridDim.x; tile < p.total_tiles; tile += NG * gridDim.x) {
      const TileCoord t = decode_tile(p, tile, BN);
      const int x = t.x0 + tx, y = t.y0 + ty, b = t.b0 + tb;
      const bool valid = (r < p.rows) && x < p.GW && y < p.GH && b < p.NB;
      if (p.reduce_out) {
        const int key = t.b0 * p.n_tiles + t.n0 / BN;
        if (key != red_key) {
          if (red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
          red_key = key;
        }
      }
      const int buf = (C::NSTG == 2) ? (int)aphase : 0;
      uint8_t* stg = stg_base + (as * C::NSTG + buf) * STG_BYTES;   // alternate staging tiles per tile
      if (p.x_tma && r == 0) {
        // saved-activation (X) tiles arrive by TMA in the staging buffer their outputs will leave from (NSTG == 2, host guarantee).
        // The tile of the group's NEXT turn is requested now, a whole epilogue ahead, so its DRAM latency is off the critical path
        // (requesting it at the start of its own turn left ~2000 cycles of latency exposed per tile: the kernel was epilogue-bound).
        if (x_first) { mbar_arrive_expect_tx(&xbar[as * 2 + buf], p.x_bytes); tma_load_4d(&p.xmap, &xbar[as * 2 + buf], stg, t.n0, t.x0, t.y0, t.b0); }
        const int next = tile + NG * gridDim.x;
        if (next < p.total_tiles) {
          const TileCoord tn = decode_tile(p, next, BN);
          tma_store_wait_read<0>();             // the previous turn's TMA store has finished reading the other staging buffer
          mbar_arrive_expect_tx(&xbar[as * 2 + (buf ^ 1)], p.x_bytes);
          tma_load_4d(&p.xmap, &xbar[as * 2 + (buf ^ 1)], stg_base + (as * C::NSTG + (buf ^ 1)) * STG_BYTES, tn.n0, tn.x0, tn.y0, tn.b0);
        }
      }
      x_first = false;
      float nz[4] = {0.f, 0.f, 0.f, 0.f};     // noise value(s) of this thread's pixel, fetched before the accumulator wait
      if (p.noise && valid) {
        if (p.superpix) {                     // row = pixel pair (2x, 2x+1) of a [H, 2*GW] noise plane
          const float2 n2 = __ldg(reinterpret_cast<const float2*>(p.noise + (long long)b * p.noise_bstride + (long long)y * (2 * p.OW) + 2 * x));
          nz[0] = n2.x * nstr; nz[1] = n2.y * nstr;
        } else {
          const int ph0 = t.n0 / p.Cout;      // one value per output phase the tile covers (BN > Cout: several)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int ph = ph0 + k;
            if (k == 0 || (k * p.Cout < BN && ph < 4))
              nz[k] = __ldg(p.noise + (long long)b * p.noise_bstride + ((long long)y * p.osy + p.ofy[ph]) * p.OW + (long long)x * p.osx + p.ofx[ph]) * nstr;
          }
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, C::NSTG, GW32>(p, t, x, y, b, valid, tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BN), lane, nz[0], nz[1], nz[2], nz[3],
                                 stg, as, r, racc, &xbar[as * 2 + buf], buf ? xph1 : xph0, &tempty[as]);
      if (buf) xph1 ^= 1; else xph0 ^= 1;
      aphase ^= 1;
    }
    if (p.reduce_out && red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
    if (r == 0) tma_store_wait_all();       // the group's last TMA store must complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_PROD) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}


// ================================================================================================================
// CTA-pair variant (tcgen05.mma.cta_group::2) for the wide tiles (BN = 128 / 256).  Two CTAs of a cluster -- two SMs of one TPC -- work on
// two neighbouring 128-pixel tiles of the same N block: one M = 256 MMA per k-step, issued by CTA 0, reads each CTA's own A tile and HALF
// of the weight tile from each CTA's shared memory, and writes each CTA's 128 accumulator rows into its own TMEM.  Per SM the weight tile
// is filled and read at half the rate: with the 128 B/clk shared-memory port model of DESIGN.md section 4, BN = 128 goes from 174 to
// 130 B/clk per MMA-cycle (port limit 74 % -> ~98 %) and BN = 256 from 130 to 87.
// Protocol: every CTA's TMA producer fills its own stage and counts the bytes on CTA 0's `full` barrier (cta_group::2 loads); CTA 0's MMA warp
// frees a stage / publishes an accumulator in BOTH CTAs with multicast commits; the epilogue groups of both CTAs drain their own TMEM
// and arrive on CTA 0's `tempty` barrier (256 arrivals).
template <int BN, int BK, int NG_ = 2>
struct Cfg2 {
  static constexpr int NG = NG_;                         // epilogue groups = accumulator stages (three for BN = 128 launches with a fused tail)
  static constexpr int THREADS = 64 + 128 * NG;
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;      // this CTA's half of the weight tile
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTG = 1;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - (NG - 2) * (STG_BYTES + 2048)) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 12 ? 12 : STAGES_RAW;
  static constexpr int BAR_BYTES = (2 * STAGES + 4 * NG) * 8 + 16;
  static constexpr int SMEM = STAGES * STAGE + NG * NSTG * STG_BYTES + NG * RACC * 4 + BAR_BYTES;
  static constexpr uint32_t TMEM_COLS = NG * BN <= 256 ? 256 : 512;
  static_assert(NG * BN <= 512 && SMEM <= 227 * 1024, "pair kernel configuration");
};

__device__ __forceinline__ TileCoord decode_pair(const Params& p, int pt, int rank, int BN) {
  // pair tiles are ordered like tiles (N block fastest); a pair = M tiles 2 mp and 2 mp + 1 of one N block.  An M index past the end decodes to a
  // sample index >= NB: TMA zero-fills its loads and clips its stores, and every row is invalid
  TileCoord t;
  const int nb = pt % p.n_tiles; int m = 2 * (pt / p.n_tiles) + rank;
  t.n0 = nb * BN;
  t.x0 = (m % p.tilesW) * p.TW; m /= p.tilesW;
  t.y0 = (m % p.tilesH) * p.TH; m /= p.tilesH;
  t.b0 = m * p.TB;
  return t;
}

template <int BN, int BK, int NG_ = 2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * NG_, 1) conv_tc2_kernel(const __grid_constant__ Params p) {
  using C = Cfg2<BN, BK, NG_>;
  constexpr int NG = NG_, W_PROD = 4 * NG_, W_MMA = 4 * NG_ + 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stg_base = smem + C::STAGES * C::STAGE;
  float* racc_base = reinterpret_cast<float*>(stg_base + NG * C::NSTG * STG_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(racc_base + NG * RACC);    // used in CTA 0 only
  uint64_t* empty = full + C::STAGES;
  uint64_t* tfull = empty + C::STAGES;
  uint64_t* tempty = tfull + NG;                                           // used in CTA 0 only
  uint64_t* xbar = tempty + NG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2 * NG);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < NG; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W_PROD) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();                              // barriers of both CTAs initialised and visible before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int iters = p.kchunks * p.ntaps;
  const int pair_tiles = p.total_tiles;        // host: number of PAIR tiles

  if (warp == W_PROD) {     // ------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0; uint32_t phase = 0;
    for (int pt = cid; pt < pair_tiles; pt += nclusters) {
      const TileCoord t = decode_pair(p, pt, rank, BN);
      const TileCoord t0 = decode_pair(p, pt, 0, BN);
      const int wbase = p.per_sample ? t0.b0 * p.w_T : 0;      // both tiles of a pair read the weights of CTA 0's tile (host: same sample)
      for (int kc = 0; kc < p.kchunks; kc++) {
        for (int tp = 0; tp < p.ntaps; tp++) {
          const Tap tap = p.taps[tp];
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * p.tx_bytes);
            uint8_t* sa = smem + stage * C::STAGE;
            tma_load_4d_2cta(&p.amap[tap.amap], &full[stage], sa, kc * BK, t.x0 + tap.dx, t.y0 + tap.dy, t.b0);
            tma_load_3d_2cta(&p.bmap, &full[stage], sa + C::A_BYTES, kc * BK, t.n0 + rank * (BN / 2), wbase + tap.wz);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {   // ------------------------------------------------- MMA issuer (CTA 0 only)
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int pt = cid; pt < pair_tiles; pt += nclusters) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int it = 0; it < iters; it++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * C::STAGE);
            const uint64_t adesc = make_desc<BK>(sa), bdesc = make_desc<BK>(sa + C::A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; k++) tc_mma_bf16_2cta(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
            tc_commit_2cta(&empty[stage]);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit_2cta(&tfull[as]);
        __syncwarp();
        if (++as == NG) { as = 0; aphase ^= 1; }
      }
    }
  } else {                 // ------------------------------------------------ epilogue (both CTAs): group g drains accumulator stage g
    const int as = warp >> 2; uint32_t aphase = 0;
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const int tx = r % p.TW, ty = (r / p.TW) % p.TH, tb = r / (p.TW * p.TH);
    const float nstr = (p.noise && p.noise_strength) ? *p.noise_strength : 1.f;
    float* racc = racc_base + as * RACC;
    int red_key = -1;
    if (p.reduce_out) { for (int j = r; j < RACC; j += 128) racc[j] = 0.f; group_sync(as); }
    for (int pt = cid + as * nclusters; pt < pair_tiles; pt += NG * nclusters) {
      const TileCoord t = decode_pair(p, pt, rank, BN);
      const int x = t.x0 + tx, y = t.y0 + ty, b = t.b0 + tb;
      const bool valid = (r < p.rows) && x < p.GW && y < p.GH && b < p.NB;
      if (p.reduce_out) {
        const int key = t.b0 * p.n_tiles + t.n0 / BN;
        if (key != red_key) {
          if (red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
          red_key = key;
        }
      }
      uint8_t* stg = stg_base + as * C::NSTG * STG_BYTES;
      float nz[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.noise && valid) {
        const int ph0 = t.n0 / p.Cout;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int ph = ph0 + k;
          if (k == 0 || (k * p.Cout < BN && ph < 4))
            nz[k] = __ldg(p.noise + (long long)b * p.noise_bstride + ((long long)y * p.osy + p.ofy[ph]) * p.OW + (long long)x * p.osx + p.ofx[ph]) * nstr;
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, C::NSTG, 2>(p, t, x, y, b, valid, tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BN), lane, nz[0], nz[1], nz[2], nz[3],
                                    stg, as, r, racc, &xbar[0], 0u, &tempty[as], rank != 0);
      aphase ^= 1;
    }
    if (p.reduce_out && red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
    if (r == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync();                              // neither CTA may leave (or free its TMEM) while the pair's MMAs / remote arrivals are in flight
  if (warp == W_PROD) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ================================================================================================================
// Halo variant for narrow layers (C = 64 or 128, 3x3 taps, large images).  The generic kernel above re-loads a shifted A tile
// from L2 for each of the 9 taps and re-fetches the weight tile for every output tile; for C <= 128 that L2->shared-memory
// traffic, not HBM or the tensor pipe, is the limiter (profiles/r01_conv_tc_ncu_full.md).  Here
//   * ONE haloed activation tile {64 ch, 16 px, 18 rows} is loaded per channel chunk (the output tile is 8 px x 16 rows; the
//     16-pixel pitch makes every 8-row core-matrix group start on a 1024-byte swizzle-atom boundary) and the 9 taps are just 9
//     shared-memory descriptors into it: start = tile + (dy*16+dx)*128 B, SBO = 2048 B (swizzle phase follows the address bits),
//   * the 3x3 weights of the current (sample, N-block) stay resident in shared memory across tiles (KC*9 tiles of BN x 64),
//     re-loaded only when the CTA moves to another sample / N-block.
// L2->SM traffic per 128-pixel tile drops from 9*(16+BN/8) KB to 36 KB per channel chunk.
constexpr int HALO_PITCH = 10;      // haloed tile row pitch in pixels: 8 output pixels + 1 halo pixel each side (no padding pixels)
template <int BN, int KC, int BK, int NSTG_, int NG_>
struct HaloCfg {
  static constexpr int A_BOX = 18 * HALO_PITCH * BK * 2;          // bytes one TMA box delivers: 18 rows x 10 px x (128 B | 64 B)
  static constexpr int A_STAGE = ((A_BOX + 1023) / 1024) * 1024;  // stages start on swizzle-atom boundaries
  static constexpr int B_TILE = BN * BK * 2;
  static constexpr int B_BYTES = KC * 9 * B_TILE;
  static constexpr int SMEM_MAX = 227 * 1024;
  static constexpr int NSTG = NSTG_;
  static constexpr int NG = NG_;                                   // epilogue groups = accumulator stages in TMEM
  static constexpr int THREADS = 64 + 128 * NG;
  static constexpr int NS_RAW = (SMEM_MAX - B_BYTES - NG * NSTG * STG_BYTES - NG * RACC * 4 - 1024) / A_STAGE;
  static constexpr int NS = NS_RAW > 8 ? 8 : NS_RAW;
  static constexpr int SMEM = B_BYTES + NS * A_STAGE + NG * NSTG * STG_BYTES + NG * RACC * 4 + 512;
  static constexpr uint32_t TMEM_COLS = NG * BN <= 32 ? 32 : NG * BN <= 64 ? 64 : NG * BN <= 128 ? 128 : NG * BN <= 256 ? 256 : 512;
  static_assert(NG >= 2 && NG <= 4 && NG * BN <= 512, "accumulator stages must fit the 512 TMEM columns");
  static_assert(NS >= 2, "halo kernel needs at least two activation stages");
};

template <int BK>
__device__ __forceinline__ uint64_t make_desc_halo(uint32_t saddr) {
  // K-major, 128B swizzle, rows of a core-matrix group 128 B apart, groups (8 px = one tile row) one haloed row (HALO_PITCH px) apart.
  // base_offset stays 0: measured on B200 the swizzle XOR is taken from the shared-memory ADDRESS bits, exactly like the TMA write
  // side, so neither a start address shifted by dx*128 B inside the atom nor a group stride that is not a multiple of the 1024-byte
  // atom needs a phase correction (setting base_offset = dx, as a literal reading of the descriptor format suggests, corrupts the
  // dx != -1 taps).  The 10-pixel pitch keeps the tile at 180 px (23 KB) instead of 288 px (36 KB): more stages in flight, 37 % less
  // L2 -> shared-memory traffic.
  constexpr uint64_t sbo = HALO_PITCH * BK * 2;
  constexpr uint64_t layout = (BK == 64) ? 2 : 4;  // 128B / 64B swizzle
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}

struct HaloTile { int n0, x0, y0, b0, key; };
__device__ __forceinline__ HaloTile decode_halo(const Params& p, int tile, int BN) {
  HaloTile t;
  const int m = tile % p.tiles_hw; t.key = tile / p.tiles_hw;      // key = b * n_tiles + n_blk: consecutive tiles share the weights
  t.n0 = (t.key % p.n_tiles) * BN; t.b0 = t.key / p.n_tiles;
  t.x0 = (m % p.tilesW) * 8; t.y0 = (m / p.tilesW) * 16;
  return t;
}

// NG epilogue groups of four warps drain NG accumulator stages in turn.  The epilogue of a 64-column tile is a dependent chain of ~700-1400
// instructions per thread (scripts/bench_halo.py: with two groups the noise+bias / reduce+X tails took 0.37-0.44 ms per launch on their own, the
// MMAs 0.28 ms), and a group is one warp per scheduler: a third group adds the latency hiding the chain lacks.
template <int BN, int KC, int BK, int NSTG_, int NG_>
__global__ void __launch_bounds__(64 + 128 * NG_, 1) conv_halo_kernel(const __grid_constant__ Params p) {
  using C = HaloCfg<BN, KC, BK, NSTG_, NG_>;
  constexpr int NG = NG_;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + C::B_BYTES;
  uint8_t* stg_base = sA + C::NS * C::A_STAGE;
  float* racc_base = reinterpret_cast<float*>(stg_base + NG * C::NSTG * STG_BYTES);
  uint64_t* afull = reinterpret_cast<uint64_t*>(racc_base + NG * RACC);
  uint64_t* aempty = afull + C::NS;
  uint64_t* bfull = aempty + C::NS;
  uint64_t* tfull = bfull + 1;
  uint64_t* tempty = tfull + NG;
  uint64_t* xbar = tempty + NG;                // [group][staging buffer]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2 * NG);
  constexpr int W_PROD = 4 * NG, W_MMA = 4 * NG + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NS; s++) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    mbar_init(bfull, 1);
    for (int s = 0; s < NG; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    for (int s = 0; s < 2 * NG; s++) mbar_init(&xbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W_PROD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_PROD) {     // ------------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0; int cur_key = -1; int last_stage = -1; uint32_t last_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const HaloTile t = decode_halo(p, tile, BN);
      if (t.key != cur_key) {
        // the resident weights are about to be overwritten: every MMA issued so far must have finished reading them.
        // MMAs complete in order, so waiting for the consumption of the most recently issued activation stage is enough.
        if (last_stage >= 0) mbar_wait(&aempty[last_stage], last_phase);
        if (elect_one()) {
          mbar_arrive_expect_tx(bfull, (uint32_t)C::B_BYTES);
          const int wbase = p.per_sample ? t.b0 * p.w_T : 0;
          for (int kc = 0; kc < KC; kc++)
            for (int tp = 0; tp < 9; tp++)
              tma_load_3d(&p.bmap, bfull, sB + (kc * 9 + tp) * C::B_TILE, kc * BK, t.n0, wbase + p.taps[tp].wz);
        }
        __syncwarp();
        cur_key = t.key;
      }
      for (int kc = 0; kc < KC; kc++) {
        mbar_wait(&aempty[stage], phase ^ 1);
        if (elect_one()) {
          if (p.dbg & 4) mbar_arrive(&afull[stage]);
          else {
            mbar_arrive_expect_tx(&afull[stage], (uint32_t)C::A_BOX);
            tma_load_4d(&p.amap[0], &afull[stage], sA + stage * C::A_STAGE, kc * BK, t.x0 - 1, t.y0 - 1, t.b0);
          }
        }
        __syncwarp();
        last_stage = stage; last_phase = phase;
        if (++stage == C::NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_MMA) {   // ------------------------------------------------- MMA issuer
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0; int cur_key = -1; uint32_t bphase = 0;
    const uint32_t sB_addr = smem_u32(sB);
    // tap -> byte offset of its shifted view inside the haloed tile (warp-uniform, hoisted out of the tile loop)
    uint32_t tap_off[9];
#pragma unroll
    for (int tp = 0; tp < 9; tp++) tap_off[tp] = (uint32_t)(((p.taps[tp].dy + 1) * HALO_PITCH + p.taps[tp].dx + 1) * (BK * 2));
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const HaloTile t = decode_halo(p, tile, BN);
      if (t.key != cur_key) { mbar_wait(bfull, bphase); bphase ^= 1; cur_key = t.key; }
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
      for (int kc = 0; kc < KC; kc++) {
        mbar_wait(&afull[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(sA + stage * C::A_STAGE);
          if (!(p.dbg & 2))
#pragma unroll
          for (int tp = 0; tp < 9; tp++) {
            const uint64_t adesc = make_desc_halo<BK>(sa + tap_off[tp]);
            const uint64_t bdesc = make_desc<BK>(sB_addr + (uint32_t)((kc * 9 + tp) * C::B_TILE));
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kc > 0 || tp > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&aempty[stage]);
        }
        __syncwarp();
        if (++stage == C::NS) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) tc_commit(&tfull[as]);
      __syncwarp();
      if (++as == NG) { as = 0; aphase ^= 1; }
    }
  } else {                 // ------------------------------------------------ epilogue groups
    const int as = warp >> 2; uint32_t aphase = 0;
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const int tx = r & 7, ty = r >> 3;
    const float nstr = (p.noise && p.noise_strength) ? *p.noise_strength : 1.f;
    float* racc = racc_base + as * RACC;
    int red_key = -1;
    if (p.reduce_out) { for (int j = r; j < RACC; j += 128) racc[j] = 0.f; group_sync(as); }
    uint32_t xph0 = 0, xph1 = 0; bool x_first = true;
    for (int tile = blockIdx.x + as * gridDim.x; tile < p.total_tiles; tile += NG * gridDim.x) {
      const HaloTile h = decode_halo(p, tile, BN);
      TileCoord t; t.n0 = h.n0; t.x0 = h.x0; t.y0 = h.y0; t.b0 = h.b0;
      const int x = t.x0 + tx, y = t.y0 + ty, b = t.b0;
      const bool valid = x < p.GW && y < p.GH;
      if (p.reduce_out && h.key != red_key) {
        if (red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
        red_key = h.key;
      }
      const int buf = (C::NSTG == 2) ? (int)aphase : 0;
      uint8_t* stg = stg_base + (as * C::NSTG + buf) * STG_BYTES;   // alternate staging tiles per tile
      if (p.x_tma && r == 0) {                // saved-activation tiles: requested one turn ahead (see conv_tc_kernel); NSTG == 2 (host guarantee)
        if (x_first) { mbar_arrive_expect_tx(&xbar[as * 2 + buf], p.x_bytes); tma_load_4d(&p.xmap, &xbar[as * 2 + buf], stg, t.n0, t.x0, t.y0, t.b0); }
        const int next = tile + NG * gridDim.x;
        if (next < p.total_tiles) {
          const HaloTile hn = decode_halo(p, next, BN);
          tma_store_wait_read<0>();
          mbar_arrive_expect_tx(&xbar[as * 2 + (buf ^ 1)], p.x_bytes);
          tma_load_4d(&p.xmap, &xbar[as * 2 + (buf ^ 1)], stg_base + (as * C::NSTG + (buf ^ 1)) * STG_BYTES, hn.n0, hn.x0, hn.y0, hn.b0);
        }
      }
      x_first = false;
      float nz[4] = {0.f, 0.f, 0.f, 0.f};     // noise value(s) of this thread's pixel, fetched before the accumulator wait
      if (p.noise && valid) {
        if (p.superpix) {                     // row = pixel pair (2x, 2x+1) of a [H, 2*GW] noise plane
          const float2 n2 = __ldg(reinterpret_cast<const float2*>(p.noise + (long long)b * p.noise_bstride + (long long)y * (2 * p.OW) + 2 * x));
          nz[0] = n2.x * nstr; nz[1] = n2.y * nstr;
        } else {
          const int ph0 = t.n0 / p.Cout;      // one value per output phase the tile covers (BN > Cout: several)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int ph = ph0 + k;
            if (k == 0 || (k * p.Cout < BN && ph < 4))
              nz[k] = __ldg(p.noise + (long long)b * p.noise_bstride + ((long long)y * p.osy + p.ofy[ph]) * p.OW + (long long)x * p.osx + p.ofx[ph]) * nstr;
          }
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, C::NSTG, (BN >= 64 ? 2 : 1)>(p, t, x, y, b, valid, tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BN), lane, nz[0], nz[1], nz[2], nz[3],
                                 stg, as, r, racc, &xbar[as * 2 + buf], buf ? xph1 : xph0, &tempty[as]);
      if (buf) xph1 ^= 1; else xph0 ^= 1;
      aphase ^= 1;
    }
    if (p.reduce_out && red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
    if (r == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_PROD) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ================================================================================================================
// CTA-pair halo kernel (tcgen05.mma.cta_group::2): two CTAs of a cluster work on two neighbouring 8 x 16 tiles of the same (sample, N block);
// CTA 0 issues one M = 256 MMA per (tap, k-step) that reads each CTA's own haloed activation tile and HALF of the resident weight tile
// (BN / 2 rows) from each CTA's shared memory.  The N = 64 halo kernel is bound by the 128 B/clk shared-memory port (stage-isolation runs,
// scripts/bench_halo.py: the 36 MMAs of a tile take 2260 cycles against a 1150-cycle tensor floor; operand reads 6 KB per MMA); in pair mode
// the weight reads per SM halve (5 KB per MMA) and the resident weights shrink to 37 KB, which buys back activation stages.
// Protocol as in conv_tc2_kernel: TMA completions of both CTAs are counted on CTA 0's barriers, CTA 0's MMA warp frees stages / publishes
// accumulators in both CTAs with multicast commits, the epilogue groups of both CTAs arrive on CTA 0's `tempty`.
template <int BN, int BK, int NSTG_, int NG_>
struct Halo2Cfg {
  static constexpr int A_BOX = 18 * HALO_PITCH * BK * 2;
  static constexpr int A_STAGE = ((A_BOX + 1023) / 1024) * 1024;
  static constexpr int B_TILE = (BN / 2) * BK * 2;                 // this CTA's half of one tap's weight tile
  static constexpr int B_BYTES = 9 * B_TILE;
  static constexpr int SMEM_MAX = 227 * 1024;
  static constexpr int NSTG = NSTG_;
  static constexpr int NG = NG_;
  static constexpr int THREADS = 64 + 128 * NG;
  static constexpr int NS_RAW = (SMEM_MAX - B_BYTES - NG * NSTG * STG_BYTES - NG * RACC * 4 - 1024) / A_STAGE;
  static constexpr int NS = NS_RAW > 6 ? 6 : NS_RAW;
  static constexpr int SMEM = B_BYTES + NS * A_STAGE + NG * NSTG * STG_BYTES + NG * RACC * 4 + 512;
  static constexpr uint32_t TMEM_COLS = NG * BN <= 32 ? 32 : NG * BN <= 64 ? 64 : NG * BN <= 128 ? 128 : NG * BN <= 256 ? 256 : 512;
  static_assert(NS >= 2 && NG * BN <= 512 && BN % 16 == 0, "pair halo kernel configuration");
};

template <int BN, int BK, int NSTG_, int NG_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * NG_, 1) conv_halo2_kernel(const __grid_constant__ Params p) {
  using C = Halo2Cfg<BN, BK, NSTG_, NG_>;
  constexpr int NG = NG_;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sA = smem + C::B_BYTES;
  uint8_t* stg_base = sA + C::NS * C::A_STAGE;
  float* racc_base = reinterpret_cast<float*>(stg_base + NG * C::NSTG * STG_BYTES);
  uint64_t* afull = reinterpret_cast<uint64_t*>(racc_base + NG * RACC);      // used in CTA 0 only
  uint64_t* aempty = afull + C::NS;
  uint64_t* bfull = aempty + C::NS;                                          // used in CTA 0 only
  uint64_t* tfull = bfull + 1;
  uint64_t* tempty = tfull + NG;                                             // used in CTA 0 only
  uint64_t* xbar = tempty + NG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2 * NG);
  constexpr int W_PROD = 4 * NG, W_MMA = 4 * NG + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int pair_tiles = p.total_tiles >> 1;          // host: tiles per (sample, N block) are even, so the two tiles of a pair share the weights

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NS; s++) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    mbar_init(bfull, 1);
    for (int s = 0; s < NG; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 256); }
    for (int s = 0; s < 2 * NG; s++) mbar_init(&xbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == W_PROD) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_PROD) {     // ------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0; uint32_t phase = 0; int cur_key = -1; int last_stage = -1; uint32_t last_phase = 0;
    for (int pt = cid; pt < pair_tiles; pt += nclusters) {
      const HaloTile t = decode_halo(p, 2 * pt + rank, BN);
      if (t.key != cur_key) {
        if (last_stage >= 0) mbar_wait(&aempty[last_stage], last_phase);      // every MMA issued so far has read the resident weights
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(bfull, 2u * (uint32_t)C::B_BYTES);
          const int wbase = p.per_sample ? t.b0 * p.w_T : 0;
          for (int tp = 0; tp < 9; tp++)
            tma_load_3d_2cta(&p.bmap, bfull, sB + tp * C::B_TILE, 0, t.n0 + rank * (BN / 2), wbase + p.taps[tp].wz);
        }
        __syncwarp();
        cur_key = t.key;
      }
      mbar_wait(&aempty[stage], phase ^ 1);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&afull[stage], 2u * (uint32_t)C::A_BOX);
        tma_load_4d_2cta(&p.amap[0], &afull[stage], sA + stage * C::A_STAGE, 0, t.x0 - 1, t.y0 - 1, t.b0);
      }
      __syncwarp();
      last_stage = stage; last_phase = phase;
      if (++stage == C::NS) { stage = 0; phase ^= 1; }
    }
  } else if (warp == W_MMA) {   // ------------------------------------------------- MMA issuer (CTA 0 only)
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0; int cur_key = -1; uint32_t bphase = 0;
      const uint32_t sB_addr = smem_u32(sB);
      uint32_t tap_off[9];
#pragma unroll
      for (int tp = 0; tp < 9; tp++) tap_off[tp] = (uint32_t)(((p.taps[tp].dy + 1) * HALO_PITCH + p.taps[tp].dx + 1) * (BK * 2));
      for (int pt = cid; pt < pair_tiles; pt += nclusters) {
        const HaloTile t = decode_halo(p, 2 * pt, BN);
        if (t.key != cur_key) { mbar_wait(bfull, bphase); bphase ^= 1; cur_key = t.key; }
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        mbar_wait(&afull[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(sA + stage * C::A_STAGE);
#pragma unroll
          for (int tp = 0; tp < 9; tp++) {
            const uint64_t adesc = make_desc_halo<BK>(sa + tap_off[tp]);
            const uint64_t bdesc = make_desc<BK>(sB_addr + (uint32_t)(tp * C::B_TILE));
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              tc_mma_bf16_2cta(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (tp > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_2cta(&aempty[stage]);
        }
        __syncwarp();
        if (++stage == C::NS) { stage = 0; phase ^= 1; }
        if (elect_one()) tc_commit_2cta(&tfull[as]);
        __syncwarp();
        if (++as == NG) { as = 0; aphase ^= 1; }
      }
    }
  } else {                 // ------------------------------------------------ epilogue groups (both CTAs)
    const int as = warp >> 2; uint32_t aphase = 0;
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const int tx = r & 7, ty = r >> 3;
    const float nstr = (p.noise && p.noise_strength) ? *p.noise_strength : 1.f;
    float* racc = racc_base + as * RACC;
    int red_key = -1;
    if (p.reduce_out) { for (int j = r; j < RACC; j += 128) racc[j] = 0.f; group_sync(as); }
    uint32_t xph0 = 0, xph1 = 0; bool x_first = true;
    for (int pt = cid + as * nclusters; pt < pair_tiles; pt += NG * nclusters) {
      const HaloTile h = decode_halo(p, 2 * pt + rank, BN);
      TileCoord t; t.n0 = h.n0; t.x0 = h.x0; t.y0 = h.y0; t.b0 = h.b0;
      const int x = t.x0 + tx, y = t.y0 + ty, b = t.b0;
      const bool valid = x < p.GW && y < p.GH;
      if (p.reduce_out && h.key != red_key) {
        if (red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
        red_key = h.key;
      }
      const int buf = (C::NSTG == 2) ? (int)aphase : 0;
      uint8_t* stg = stg_base + (as * C::NSTG + buf) * STG_BYTES;
      if (p.x_tma && r == 0) {                // saved-activation tiles: requested one turn ahead; NSTG == 2 (host guarantee)
        if (x_first) { mbar_arrive_expect_tx(&xbar[as * 2 + buf], p.x_bytes); tma_load_4d(&p.xmap, &xbar[as * 2 + buf], stg, t.n0, t.x0, t.y0, t.b0); }
        const int next = pt + NG * nclusters;
        if (next < pair_tiles) {
          const HaloTile hn = decode_halo(p, 2 * next + rank, BN);
          tma_store_wait_read<0>();
          mbar_arrive_expect_tx(&xbar[as * 2 + (buf ^ 1)], p.x_bytes);
          tma_load_4d(&p.xmap, &xbar[as * 2 + (buf ^ 1)], stg_base + (as * C::NSTG + (buf ^ 1)) * STG_BYTES, hn.n0, hn.x0, hn.y0, hn.b0);
        }
      }
      x_first = false;
      float nz[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.noise && valid) {
        if (p.superpix) {
          const float2 n2 = __ldg(reinterpret_cast<const float2*>(p.noise + (long long)b * p.noise_bstride + (long long)y * (2 * p.OW) + 2 * x));
          nz[0] = n2.x * nstr; nz[1] = n2.y * nstr;
        } else {
          const int ph0 = t.n0 / p.Cout;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int ph = ph0 + k;
            if (k == 0 || (k * p.Cout < BN && ph < 4))
              nz[k] = __ldg(p.noise + (long long)b * p.noise_bstride + ((long long)y * p.osy + p.ofy[ph]) * p.OW + (long long)x * p.osx + p.ofx[ph]) * nstr;
          }
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      epilogue_tile<BN, C::NSTG, (BN >= 64 ? 2 : 1)>(p, t, x, y, b, valid, tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BN), lane, nz[0], nz[1], nz[2], nz[3],
                                 stg, as, r, racc, &xbar[as * 2 + buf], buf ? xph1 : xph0, &tempty[as], rank != 0);
      if (buf) xph1 ^= 1; else xph0 ^= 1;
      aphase ^= 1;
    }
    if (p.reduce_out && red_key >= 0) flush_reduce<BN>(p, racc, red_key, as, r);
    if (r == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync();                              // neither CTA may leave (or free its TMEM) while the pair's MMAs / remote arrivals are in flight
  if (warp == W_PROD) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  }
  return fn;
}

static int encode(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, int bk) {
  EncodeFn fn = get_encode();
  if (!fn) MGF_FAIL(MGF_E_DRIVER, "conv_tc: cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) MGF_FAIL(MGF_E_DRIVER, "conv_tc: cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=%llu,%llu,%llu box=%u,%u,%u",
                                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1], box[2]);
  return 0;
}


// output tensor maps for the TMA-store epilogue: one per output phase, box = {64 or 32 channels, TW, TH, TB} of the phase grid
static int encode_out_maps(Params& p, const mgf_conv_tc_desc* d, int BN, int TW, int TH, int TB, bool g32 = false) {
  const int gwd = (BN >= 64 && !g32) ? 64 : 32;
  if (((uintptr_t)d->out & 15) || (d->OC * 2) % 16) MGF_FAIL(MGF_E_ALIGN, "conv_tc: output must be 16-byte aligned with OC %% 8 == 0");
  for (int ph = 0; ph < d->phases; ph++) {
    const long long ofy = d->ofy[ph], ofx = d->ofx[ph];
    if (ofy < 0 || ofx < 0 || ofy >= d->osy || ofx >= d->osx) MGF_FAIL(MGF_E_BADARG, "conv_tc: phase offset outside the output stride");
    const char* base = (const char*)d->out + ((ofy * d->OW + ofx) * d->OC) * 2;
    cuuint64_t dims[4] = {(cuuint64_t)d->OC, (cuuint64_t)((d->OW - ofx + d->osx - 1) / d->osx), (cuuint64_t)((d->OH - ofy + d->osy - 1) / d->osy), (cuuint64_t)d->NB};
    cuuint64_t strides[3] = {(cuuint64_t)d->osx * d->OC * 2, (cuuint64_t)d->osy * d->OW * d->OC * 2, (cuuint64_t)d->OH * d->OW * d->OC * 2};
    cuuint32_t box[4] = {(cuuint32_t)gwd, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    if (int e = encode(&p.omap[ph], base, 4, dims, strides, box, gwd)) return e;
  }
  return 0;
}

// X (saved activation) tile by TMA when the tile is a single staging group (BN <= 64) and the output is not phase-interleaved
static int encode_x_map(Params& p, const mgf_conv_tc_desc* d, int BN, int TW, int TH, int TB, int rows) {
  p.x_tma = 0; p.x_bytes = 0;
  const bool needX = d->X && (d->reduce_out || d->actgrad);
  if (!needX || BN > 64 || d->phases != 1 || d->osx != 1 || d->osy != 1) return 0;
  if ((uintptr_t)d->X & 15) MGF_FAIL(MGF_E_ALIGN, "conv_tc: X must be 16-byte aligned");
  const int gwd = BN >= 64 ? 64 : 32;
  cuuint64_t dims[4] = {(cuuint64_t)d->OC, (cuuint64_t)d->OW, (cuuint64_t)d->OH, (cuuint64_t)d->NB};
  cuuint64_t strides[3] = {(cuuint64_t)d->OC * 2, (cuuint64_t)d->OW * d->OC * 2, (cuuint64_t)d->OH * d->OW * d->OC * 2};
  cuuint32_t box[4] = {(cuuint32_t)gwd, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
  if (int e = encode(&p.xmap, d->X, 4, dims, strides, box, gwd)) return e;
  p.x_tma = 1; p.x_bytes = (uint32_t)(rows * gwd * 2);
  return 0;
}

template <int BN, int BK, bool G32 = false, int NG = 2>
static int launch(const Params& p, int grid, cudaStream_t st) {
  using C = Cfg<BN, BK, NG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, BK, G32, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) MGF_FAIL((int)e, "conv_tc: cannot set %d bytes of dynamic shared memory: %s", C::SMEM, cudaGetErrorString(e));
    configured = true;
  }
  conv_tc_kernel<BN, BK, G32, NG><<<grid, C::THREADS, C::SMEM, st>>>(p);
  MGF_CHECK_LAUNCH("conv_tc");
  return 0;
}

template <int BN, int BK, int NG = 2>
static int launch2(const Params& p, int grid, cudaStream_t st) {
  using C = Cfg2<BN, BK, NG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN, BK, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) MGF_FAIL((int)e, "conv_tc(2cta): cannot set %d bytes of dynamic shared memory: %s", C::SMEM, cudaGetErrorString(e));
    configured = true;
  }
  conv_tc2_kernel<BN, BK, NG><<<grid, C::THREADS, C::SMEM, st>>>(p);      // __cluster_dims__(2,1,1): grid is even
  MGF_CHECK_LAUNCH("conv_tc(2cta)");
  return 0;
}

template <int BN, int KC, int BK, int NSTG_, int NG_>
static int launch_halo(const Params& p, int grid, cudaStream_t st) {
  using C = HaloCfg<BN, KC, BK, NSTG_, NG_>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, KC, BK, NSTG_, NG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) MGF_FAIL((int)e, "conv_tc(halo): cannot set %d bytes of dynamic shared memory: %s", C::SMEM, cudaGetErrorString(e));
    configured = true;
  }
  conv_halo_kernel<BN, KC, BK, NSTG_, NG_><<<grid, C::THREADS, C::SMEM, st>>>(p);
  MGF_CHECK_LAUNCH("conv_tc(halo)");
  return 0;
}

template <int BN, int BK, int NSTG_, int NG_>
static int launch_halo2(const Params& p, int grid, cudaStream_t st) {
  using C = Halo2Cfg<BN, BK, NSTG_, NG_>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo2_kernel<BN, BK, NSTG_, NG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) MGF_FAIL((int)e, "conv_tc(halo pair): cannot set %d bytes of dynamic shared memory: %s", C::SMEM, cudaGetErrorString(e));
    configured = true;
  }
  conv_halo2_kernel<BN, BK, NSTG_, NG_><<<grid, C::THREADS, C::SMEM, st>>>(p);      // __cluster_dims__(2,1,1): grid is even
  MGF_CHECK_LAUNCH("conv_tc(halo pair)");
  return 0;
}

// CTA-pair halo kernel for the 64-channel layers: OFF by default (A/B switch: mgf_conv_tc_set_halo bit 6 enables it).  Measured on B200
// (scripts/bench_halo.py, 8 images): it is correct (tests/test_conv_tc_gpu.py runs both) but SLOWER than the one-CTA kernel -- 1024^2 32-channel
// forward 0.372 vs 0.333 ms, dgrad + reduction 0.402 vs 0.359, VGG conv1_2 0.669 vs 0.632: halving the weight reads (6 -> 5 KB per MMA) buys less
// than the pair costs (every tile waits for the slower CTA's activation tile, accumulator hand-over and tail through cluster-scope barriers).
static bool g_halo_pair = false;
static bool g_halo_enabled = true;
static bool g_halo_phases = false;  // halo kernel also for multi-phase (up-convolution) launches; default: those run as one wide tile per pixel block
static bool g_cg2_enabled = true;     // CTA-pair (cta_group::2) kernel for the wide tiles (A/B switch: mgf_conv_tc_set_halo bit 4 disables it)
static int g_halo_groups = 3;      // epilogue groups of the halo kernel (A/B switch: mgf_conv_tc_set_halo bit 5 selects 2)
static int g_dbg = 0;             // experiment switches copied into Params::dbg (mgf_conv_tc_set_halo bits 8..10)
static int g_halo_nstg = 1;        // epilogue staging tiles per group in the halo kernel (A/B switch: mgf_conv_tc_set_halo(1 | 2 << 1))

}  // namespace tc
}  // namespace mgf

extern "C" int mgf_conv_tc_set_halo(int mode) {
  // bit 0: halo kernel on/off; bits 1..2: staging tiles per epilogue group (0 = default 1, else 1 or 2); bit 3: halo kernel for multi-phase launches too; bit 4: disable the CTA-pair kernel; bit 5: two epilogue groups in the halo kernel (default three); bit 6: enable the CTA-pair halo kernel (default off: slower); bits 8..13: experiment switches
  mgf::tc::g_halo_enabled = (mode & 1) != 0;
  const int n = (mode >> 1) & 3;
  mgf::tc::g_halo_nstg = (n == 2) ? 2 : 1;
  mgf::tc::g_halo_phases = (mode & 8) != 0;
  mgf::tc::g_cg2_enabled = (mode & 16) == 0;
  mgf::tc::g_dbg = (mode >> 8) & 63;
  mgf::tc::g_halo_groups = (mode & 32) ? 2 : 3;
  mgf::tc::g_halo_pair = (mode & 64) != 0;
  return 0;
}

extern "C" int mgf_conv_tc(const mgf_conv_tc_desc* d, void* stream) {
  using namespace mgf;
  using namespace mgf::tc;
  if (!d) MGF_FAIL(MGF_E_BADARG, "conv_tc: null descriptor");
  if (d->n_a < 1 || d->n_a > MAX_AMAPS) MGF_FAIL(MGF_E_BADARG, "conv_tc: n_a=%d outside 1..%d", d->n_a, MAX_AMAPS);
  if (d->ntaps < 1 || d->ntaps > MAX_TAPS) MGF_FAIL(MGF_E_BADARG, "conv_tc: ntaps=%d outside 1..%d", d->ntaps, MAX_TAPS);
  if (!d->w || !d->out) MGF_FAIL(MGF_E_BADARG, "conv_tc: null weight/output");
  if (!(d->gain > 0.f) || (d->act == 1 && !(d->alpha >= 0.f && d->alpha <= 1.f))) MGF_FAIL(MGF_E_BADARG, "conv_tc: the fused tail needs gain > 0 and 0 <= alpha <= 1");
  const long long Cc = d->a[0].C;
  if (Cc % 32 != 0 || Cc < 32) MGF_FAIL(MGF_E_SHAPE, "conv_tc: input channels (%lld) must be a multiple of 32", Cc);
  if (d->w_K != Cc) MGF_FAIL(MGF_E_SHAPE, "conv_tc: weight K (%lld) != activation channels (%lld)", (long long)d->w_K, Cc);
  const int BK = (Cc % 64 == 0) ? 64 : 32;
  if (d->phases < 1 || d->phases > 4 || d->Cout < 32 || d->Cout % 32) MGF_FAIL(MGF_E_SHAPE, "conv_tc: Cout (%d) must be a multiple of 32, phases 1..4", d->Cout);
  const long long NT = (long long)d->phases * d->Cout;
  const bool ph_taps = d->phase_ntaps[0] > 0;
  if (ph_taps) {
    int tot = 0;
    for (int i = 0; i < d->phases; i++) { if (d->phase_ntaps[i] < 1) MGF_FAIL(MGF_E_BADARG, "conv_tc: phase %d has no taps", i); tot += d->phase_ntaps[i]; }
    if (tot != d->ntaps) MGF_FAIL(MGF_E_BADARG, "conv_tc: phase tap lists hold %d taps, ntaps = %d", tot, d->ntaps);
    if (d->reduce_out || d->X || d->superpix) MGF_FAIL(MGF_E_UNSUP, "conv_tc: per-phase tap lists are a forward-only form (no reduce_out / X / superpix)");
    if (d->w_NT != d->Cout) MGF_FAIL(MGF_E_SHAPE, "conv_tc: per-phase tap lists share one [T][Cout][K] weight tensor (w_NT %lld != Cout %d)", (long long)d->w_NT, d->Cout);
  } else if (d->w_NT != NT) MGF_FAIL(MGF_E_SHAPE, "conv_tc: weight NT (%lld) != phases*Cout (%lld)", (long long)d->w_NT, NT);
  if (d->OC < d->Cout || d->OC % 8) MGF_FAIL(MGF_E_SHAPE, "conv_tc: bad output channel stride");
  int BN = d->bn;
  // a tile either lies inside one phase (BN divides Cout) or covers whole phases (Cout divides BN, BN divides phases * Cout): the latter
  // fetches every activation tile once for all the phases it covers (the four parities of an up-convolution in one tile)
  auto bn_ok = [&](int bn) { return (d->Cout % bn == 0) || (!ph_taps && !d->superpix && bn % d->Cout == 0 && NT % bn == 0); };
  if (BN == 0) BN = bn_ok(256) ? 256 : bn_ok(128) ? 128 : bn_ok(64) ? 64 : 32;
  if (!(BN == 32 || BN == 64 || BN == 128 || BN == 256) || !bn_ok(BN)) MGF_FAIL(MGF_E_SHAPE, "conv_tc: BN=%d does not fit Cout=%d x %d phases", BN, d->Cout, d->phases);
  if ((d->reduce_out || d->X || d->add) && d->bn == 0) {        // epilogues that read per-pixel tensors stay inside one phase
    BN = (d->Cout % 256 == 0) ? 256 : (d->Cout % 128 == 0) ? 128 : (d->Cout % 64 == 0) ? 64 : 32;
  }
  if (d->GW < 1 || d->GH < 1 || d->NB < 1) MGF_FAIL(MGF_E_SHAPE, "conv_tc: empty grid");
  const int per_sample = d->w_G > 1 ? 1 : 0;
  if (per_sample && d->w_G != d->NB) MGF_FAIL(MGF_E_SHAPE, "conv_tc: per-sample weights need G == NB");
  if (d->reduce_out && !per_sample && d->NB > 1 && !d->reduce_per_sample) MGF_FAIL(MGF_E_UNSUP, "conv_tc: the per-sample reduction needs reduce_per_sample (one sample per tile)");

  Params p;
  memset(&p, 0, sizeof(p));
  p.dbg = g_dbg;
  // ---- halo variant: 3x3 taps in {-1,0,1}^2 on one activation map, C = 64 or 128, images large enough that a CTA keeps the
  // resident weights for many tiles
  // (measured: a win for C = 64 -- 0.95 -> 0.68 ms on 64->64 @1024^2 x8 -- but not for C = 128, where the generic kernel's
  // 128-wide N tile beats two 64-wide halo passes; scripts/bench_conv_tc.py)
  bool halo = g_halo_enabled && !ph_taps && (d->phases == 1 || g_halo_phases) && d->n_a == 1 && d->ntaps == 9 && (Cc == 64 || Cc == 32 || (Cc == 128 && d->Cout == 64)) && d->bn == 0 && d->GW >= 16 &&
              (long long)d->GH * d->GW >= 256LL * 256 && (per_sample || d->w_G == 1);
  if (halo) {
    int seen = 0;
    for (int i = 0; i < 9; i++) {
      const int dy = d->taps[i].dy, dx = d->taps[i].dx;
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || d->taps[i].amap != 0) { halo = false; break; }
      seen |= 1 << ((dy + 1) * 3 + dx + 1);
    }
    if (seen != 0x1FF) halo = false;
  }
  if (halo) {
    // 128 -> 64 channels (the input gradient of VGG conv2_1): two 64-channel chunks per tile, 147 KB of resident weights, two activation
    // stages, one staging tile per group (so the saved activation, if any, is read per thread instead of by TMA)
    const int KC = (Cc == 128) ? 2 : 1;
    const int HBK = (Cc == 128) ? 64 : (int)Cc;   // 64 (128B-swizzled rows) or 32 (64B-swizzled rows)
    int HBN = (d->Cout % 64 == 0) ? 64 : 32;
    p.TW = 8; p.TH = 16; p.TB = 1; p.rows = 128;
    p.NB = d->NB; p.GH = d->GH; p.GW = d->GW;
    p.tilesW = (d->GW + 7) / 8; p.tilesH = (d->GH + 15) / 16; p.tilesB = d->NB; p.tiles_hw = p.tilesW * p.tilesH;
    p.NT = (int)NT; p.Cout = d->Cout; p.n_tiles = (int)(NT / HBN); p.per_sample = per_sample; p.w_T = (int)d->w_T;
    const long long total = (long long)p.tiles_hw * d->NB * p.n_tiles;
    if (total > 0x7fffffffLL) MGF_FAIL(MGF_E_SHAPE, "conv_tc: too many tiles");
    p.total_tiles = (int)total; p.ntaps = 9; p.kchunks = KC;
    for (int i = 0; i < 9; i++) { p.taps[i].amap = 0; p.taps[i].dy = d->taps[i].dy; p.taps[i].dx = d->taps[i].dx; p.taps[i].wz = d->taps[i].wz;
      if (d->taps[i].wz < 0 || d->taps[i].wz >= d->w_T) MGF_FAIL(MGF_E_BADARG, "conv_tc: tap %d out of range", i); }
    const mgf_tc_act& a = d->a[0];
    if (!a.ptr || ((uintptr_t)a.ptr & 15) || (a.sW * 2) % 16 || (a.sH * 2) % 16 || (a.sN * 2) % 16) MGF_FAIL(MGF_E_ALIGN, "conv_tc: activation must be 16-byte aligned/strided");
    {
      cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
      cuuint64_t strides[3] = {(cuuint64_t)a.sW * 2, (cuuint64_t)a.sH * 2, (cuuint64_t)a.sN * 2};
      cuuint32_t box[4] = {(cuuint32_t)HBK, (cuuint32_t)HALO_PITCH, 18, 1};
      if (int e = encode(&p.amap[0], a.ptr, 4, dims, strides, box, HBK)) return e;
    }
    {
      if ((uintptr_t)d->w & 15) MGF_FAIL(MGF_E_ALIGN, "conv_tc: weights must be 16-byte aligned");
      cuuint64_t dims[3] = {(cuuint64_t)d->w_K, (cuuint64_t)d->w_NT, (cuuint64_t)(d->w_G * d->w_T)};
      cuuint64_t strides[2] = {(cuuint64_t)d->w_K * 2, (cuuint64_t)d->w_K * d->w_NT * 2};
      cuuint32_t box[3] = {(cuuint32_t)HBK, (cuuint32_t)HBN, 1};
      if (int e = encode(&p.bmap, d->w, 3, dims, strides, box, HBK)) return e;
    }
    p.out = d->out; p.OH = d->OH; p.OW = d->OW; p.OC = d->OC; p.osy = d->osy; p.osx = d->osx;
    for (int i = 0; i < 4; i++) { p.ofy[i] = d->ofy[i]; p.ofx[i] = d->ofx[i]; }
    p.scale_n = d->scale_n; p.reduce_out = d->reduce_out; p.X = (const __nv_bfloat16*)d->X;
    p.noise = d->noise; p.noise_strength = d->noise_strength; p.bias = d->bias;
    p.act = d->act; p.alpha = d->alpha; p.gain = d->gain; p.add = (const __nv_bfloat16*)d->add;
    p.actgrad = d->actgrad; p.ag_alpha = d->ag_alpha; p.ag_gain = d->ag_gain; p.superpix = d->superpix; p.noise_bstride = d->noise_bstride;
    if ((p.reduce_out || p.actgrad) && !p.X) MGF_FAIL(MGF_E_BADARG, "conv_tc: reduce/actgrad need X");
    const bool f16 = fwd_f16();
    const uint32_t fmt = (f16 && d->ab_fwd) ? 0u : 1u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(HBN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.out_f16 = f16 && d->out_fwd; p.x_f16 = f16 && d->x_fwd; p.add_f16 = f16 && d->add_fwd;
    p.ovf = p.out_f16 ? overflow_flag() : nullptr;
    if (int e = encode_out_maps(p, d, HBN, 8, 16, 1)) return e;
    if (KC == 1) { if (int e = encode_x_map(p, d, HBN, 8, 16, 1, 128)) return e; }
    int grid = num_sms(); if (grid > p.total_tiles) grid = p.total_tiles;
    cudaStream_t st = (cudaStream_t)stream;
    if (KC == 2) return launch_halo<64, 2, 64, 1, 2>(p, grid, st);
    // CTA-pair kernel: 64 -> 64 channel tiles whose pair shares the weights (an even number of tiles per sample and N block)
    if (g_halo_pair && HBN == 64 && HBK == 64 && p.tiles_hw % 2 == 0 && p.total_tiles >= 2 * num_sms() && !p.dbg) {
      Params q = p;
      q.idesc = (p.idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
      cuuint64_t dims[3] = {(cuuint64_t)d->w_K, (cuuint64_t)d->w_NT, (cuuint64_t)(d->w_G * d->w_T)};
      cuuint64_t strides[2] = {(cuuint64_t)d->w_K * 2, (cuuint64_t)d->w_K * d->w_NT * 2};
      cuuint32_t box[3] = {(cuuint32_t)HBK, (cuuint32_t)(HBN / 2), 1};
      if (int e = encode(&q.bmap, d->w, 3, dims, strides, box, HBK)) return e;
      const int g2 = (num_sms() / 2) * 2;
      const bool tail3 = g_halo_groups == 3 && (p.noise || p.bias || p.act || p.X || p.reduce_out || p.add || p.scale_n);
      if (p.x_tma || g_halo_nstg == 2) return tail3 ? launch_halo2<64, 64, 2, 3>(q, g2, st) : launch_halo2<64, 64, 2, 2>(q, g2, st);
      return tail3 ? launch_halo2<64, 64, 1, 3>(q, g2, st) : launch_halo2<64, 64, 1, 2>(q, g2, st);
    }
    const bool nstg2 = g_halo_nstg == 2 || p.x_tma;      // X tiles are prefetched into the second staging buffer
    // a third epilogue group pays when the per-tile tail is long (measured, scripts/bench_halo.py at 1024x512x64: noise+bias+lrelu 0.40 -> 0.34 ms,
    // reduce+X 0.43 -> 0.41, VGG bias+ReLU 0.67 -> 0.65); a bare convert-and-store tail is MMA-bound with two (0.31 vs 0.32 ms)
    const bool tail_work = p.noise || p.bias || p.act || p.X || p.reduce_out || p.add || p.scale_n;
#define MGF_HALO_CASE(bn, bk) if (HBN == bn && HBK == bk) { \
      if (g_halo_groups == 3 && tail_work) return nstg2 ? launch_halo<bn, 1, bk, 2, 3>(p, grid, st) : launch_halo<bn, 1, bk, 1, 3>(p, grid, st); \
      return nstg2 ? launch_halo<bn, 1, bk, 2, 2>(p, grid, st) : launch_halo<bn, 1, bk, 1, 2>(p, grid, st); }
    MGF_HALO_CASE(64, 64) MGF_HALO_CASE(32, 64) MGF_HALO_CASE(64, 32) MGF_HALO_CASE(32, 32)
#undef MGF_HALO_CASE
    MGF_FAIL(MGF_E_UNSUP, "conv_tc: no halo kernel for BN=%d KC=%d", HBN, KC);
  }
  // tile shape: TW x TH x TB pixels = at most 128 rows
  int TW = 1; while (TW * 2 <= d->GW && TW < 16) TW *= 2;
  int TH = 1; while (TH * 2 <= d->GH && TW * TH * 2 <= 128) TH *= 2;
  int TB = 1;
  if (!per_sample && !d->reduce_per_sample) { while (TW * TH * TB * 2 <= 128 && TB * 2 <= d->NB) TB *= 2; }
  p.TW = TW; p.TH = TH; p.TB = TB; p.rows = TW * TH * TB;
  p.NB = d->NB; p.GH = d->GH; p.GW = d->GW;
  p.tilesW = (d->GW + TW - 1) / TW; p.tilesH = (d->GH + TH - 1) / TH; p.tilesB = (d->NB + TB - 1) / TB;
  if (d->bn == 0) {
    // tiny images: a 128 x 256 tile leaves most of the 148 SMs idle while a few CTAs stream the whole weight tensor through their
    // own L2 port (measured 0.12 ms for a 4x4 layer).  Narrow the N tile until the grid covers the machine.
    const long long mt = (long long)p.tilesW * p.tilesH * p.tilesB;
    while (BN > 32 && mt * (NT / BN) < num_sms() / 2 && bn_ok(BN >> 1)) BN >>= 1;
  }
  const bool g32 = BN >= 64 && (d->Cout % 64) != 0;      // phases of 32 channels inside a wider tile: 32-column staged groups
  if (BN > d->Cout && (d->reduce_out || d->X || d->add)) MGF_FAIL(MGF_E_UNSUP, "conv_tc: tiles that span phases are a forward-only form (no reduce_out / X / add)");
  p.NT = (int)NT; p.Cout = d->Cout; p.n_tiles = (int)(NT / BN); p.per_sample = per_sample; p.w_T = (int)d->w_T;
  const long long total = (long long)p.tilesW * p.tilesH * p.tilesB * p.n_tiles;
  if (total > 0x7fffffffLL) MGF_FAIL(MGF_E_SHAPE, "conv_tc: too many tiles");
  p.total_tiles = (int)total;
  p.ntaps = d->ntaps; p.kchunks = (int)(Cc / BK);
  p.ph_taps = ph_taps ? 1 : 0;
  if (ph_taps) {
    p.ph_tap0[0] = 0;
    for (int i = 0; i < d->phases; i++) p.ph_tap0[i + 1] = p.ph_tap0[i] + d->phase_ntaps[i];
    p.ph_tiles = p.total_tiles / d->phases;
  }
  for (int i = 0; i < d->ntaps; i++) {
    if (d->taps[i].amap < 0 || d->taps[i].amap >= d->n_a || d->taps[i].wz < 0 || d->taps[i].wz >= d->w_T)
      MGF_FAIL(MGF_E_BADARG, "conv_tc: tap %d out of range", i);
    p.taps[i].amap = d->taps[i].amap; p.taps[i].dy = d->taps[i].dy; p.taps[i].dx = d->taps[i].dx; p.taps[i].wz = d->taps[i].wz;
  }
  for (int i = 0; i < d->n_a; i++) {
    const mgf_tc_act& a = d->a[i];
    if (!a.ptr || a.C != Cc) MGF_FAIL(MGF_E_SHAPE, "conv_tc: activation %d: null or channel mismatch", i);
    if (((uintptr_t)a.ptr & 15) || (a.sW * 2) % 16 || (a.sH * 2) % 16 || (a.sN * 2) % 16) MGF_FAIL(MGF_E_ALIGN, "conv_tc: activation %d must be 16-byte aligned/strided", i);
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {(cuuint64_t)a.sW * 2, (cuuint64_t)a.sH * 2, (cuuint64_t)a.sN * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    if (int e = encode(&p.amap[i], a.ptr, 4, dims, strides, box, BK)) return e;
  }
  {
    if ((uintptr_t)d->w & 15) MGF_FAIL(MGF_E_ALIGN, "conv_tc: weights must be 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)d->w_K, (cuuint64_t)d->w_NT, (cuuint64_t)(d->w_G * d->w_T)};
    cuuint64_t strides[2] = {(cuuint64_t)d->w_K * 2, (cuuint64_t)d->w_K * d->w_NT * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    if (int e = encode(&p.bmap, d->w, 3, dims, strides, box, BK)) return e;
  }
  p.out = d->out; p.OH = d->OH; p.OW = d->OW; p.OC = d->OC; p.osy = d->osy; p.osx = d->osx;
  for (int i = 0; i < 4; i++) { p.ofy[i] = d->ofy[i]; p.ofx[i] = d->ofx[i]; }
  p.scale_n = d->scale_n; p.reduce_out = d->reduce_out; p.X = (const __nv_bfloat16*)d->X;
  p.noise = d->noise; p.noise_strength = d->noise_strength; p.bias = d->bias;
  p.act = d->act; p.alpha = d->alpha; p.gain = d->gain; p.add = (const __nv_bfloat16*)d->add;
  p.actgrad = d->actgrad; p.ag_alpha = d->ag_alpha; p.ag_gain = d->ag_gain; p.superpix = d->superpix; p.noise_bstride = d->noise_bstride;
  if (p.superpix && (d->phases != 1 || d->Cout != 64)) MGF_FAIL(MGF_E_UNSUP, "conv_tc: superpix needs phases == 1 and Cout == 64 (two 32-channel pixels)");
  if ((p.reduce_out || p.actgrad) && !p.X) MGF_FAIL(MGF_E_BADARG, "conv_tc: reduce/actgrad need X");
  p.tx_bytes = (uint32_t)((p.rows * BK + BN * BK) * 2);
  {
    const bool f16 = fwd_f16();
    const uint32_t fmt = (f16 && d->ab_fwd) ? 0u : 1u;      // InstrDescriptor a_format/b_format: 0 = F16, 1 = BF16
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    p.out_f16 = f16 && d->out_fwd; p.x_f16 = f16 && d->x_fwd; p.add_f16 = f16 && d->add_fwd;
    p.ovf = p.out_f16 ? overflow_flag() : nullptr;
  }
  if (int e = encode_out_maps(p, d, BN, TW, TH, TB, g32)) return e;
  if (int e = encode_x_map(p, d, BN, TW, TH, TB, p.rows)) return e;
  int grid = num_sms(); if (grid > p.total_tiles) grid = p.total_tiles;
  cudaStream_t st = (cudaStream_t)stream;
  // CTA-pair kernel: wide full tiles, both tiles of a pair under the same weights (shared weights, or an even number of tiles per sample)
  const long long mtiles = (long long)p.tilesW * p.tilesH * p.tilesB;
  if (g_cg2_enabled && BN >= 128 && !g32 && !ph_taps && p.rows == 128 && !p.x_tma && mtiles >= 2 &&
      (!per_sample || ((long long)p.tilesW * p.tilesH) % 2 == 0)) {
    Params q = p;
    const long long pairs = ((mtiles + 1) / 2) * q.n_tiles;
    q.total_tiles = (int)pairs;
    q.idesc = (p.idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
    // each CTA loads half of the weight tile
    cuuint64_t dims[3] = {(cuuint64_t)d->w_K, (cuuint64_t)d->w_NT, (cuuint64_t)(d->w_G * d->w_T)};
    cuuint64_t strides[2] = {(cuuint64_t)d->w_K * 2, (cuuint64_t)d->w_K * d->w_NT * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(BN / 2), 1};
    if (int e = encode(&q.bmap, d->w, 3, dims, strides, box, BK)) return e;
    q.tx_bytes = (uint32_t)((128 * BK + (BN / 2) * BK) * 2);
    int g2 = (num_sms() / 2) * 2; if (g2 > 2 * pairs) g2 = (int)(2 * pairs);
    const bool tail3 = g_halo_groups == 3 && (p.noise || p.bias || p.act || p.X || p.reduce_out || p.add || p.scale_n);
    if (BN == 128 && BK == 64 && tail3) return launch2<128, 64, 3>(q, g2, st);
    if (BN == 128 && BK == 32 && tail3) return launch2<128, 32, 3>(q, g2, st);
    if (BN == 256 && BK == 64) return launch2<256, 64>(q, g2, st);
    if (BN == 128 && BK == 64) return launch2<128, 64>(q, g2, st);
    if (BN == 256 && BK == 32) return launch2<256, 32>(q, g2, st);
    if (BN == 128 && BK == 32) return launch2<128, 32>(q, g2, st);
  }
  // narrow tiles with a long per-tile tail: three epilogue groups (see conv_halo_kernel); wide tiles keep two (TMEM holds two 256-column stages)
  const bool tail_work = p.noise || p.bias || p.act || p.X || p.reduce_out || p.add || p.scale_n;
  if (g_halo_groups == 3 && tail_work && BN <= 128) {
    if (g32) {
      if (BN == 128 && BK == 64) return launch<128, 64, true, 3>(p, grid, st);
      if (BN == 128 && BK == 32) return launch<128, 32, true, 3>(p, grid, st);
      if (BN == 64 && BK == 64) return launch<64, 64, true, 3>(p, grid, st);
      if (BN == 64 && BK == 32) return launch<64, 32, true, 3>(p, grid, st);
    } else {
      if (BN == 128 && BK == 64) return launch<128, 64, false, 3>(p, grid, st);
      if (BN == 64 && BK == 64) return launch<64, 64, false, 3>(p, grid, st);
      if (BN == 32 && BK == 64) return launch<32, 64, false, 3>(p, grid, st);
      if (BN == 128 && BK == 32) return launch<128, 32, false, 3>(p, grid, st);
      if (BN == 64 && BK == 32) return launch<64, 32, false, 3>(p, grid, st);
      if (BN == 32 && BK == 32) return launch<32, 32, false, 3>(p, grid, st);
    }
  }
  if (g32) {
    if (BN == 128 && BK == 64) return launch<128, 64, true>(p, grid, st);
    if (BN == 128 && BK == 32) return launch<128, 32, true>(p, grid, st);
    if (BN == 64 && BK == 64) return launch<64, 64, true>(p, grid, st);
    if (BN == 64 && BK == 32) return launch<64, 32, true>(p, grid, st);
    MGF_FAIL(MGF_E_UNSUP, "conv_tc: no 32-column-group kernel for BN=%d BK=%d", BN, BK);
  }
#define MGF_TC_CASE(bn, bk) if (BN == bn && BK == bk) return launch<bn, bk>(p, grid, st);
  MGF_TC_CASE(256, 64) MGF_TC_CASE(128, 64) MGF_TC_CASE(64, 64) MGF_TC_CASE(32, 64)
  MGF_TC_CASE(256, 32) MGF_TC_CASE(128, 32) MGF_TC_CASE(64, 32) MGF_TC_CASE(32, 32)
#undef MGF_TC_CASE
  MGF_FAIL(MGF_E_UNSUP, "conv_tc: no kernel for BN=%d BK=%d", BN, BK);
}
