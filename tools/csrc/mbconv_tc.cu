// Fused MobileNetV2 inverted-residual block, all three convolutions on the tensor cores (sm_100a, eval mode,
// BatchNorm folded, bf16 activations):
//
//     y = project_1x1( relu6( dw3x3_s( relu6( expand_1x1(x) ) ) ) )  [+ x]
//
// (tv:models/mobilenetv2.py:38-62, reached from the reference through unet.py:15-19,34-42.)  Same contract as
// b200seg_mbconv (mbconv.cu); there the depthwise stencil runs on the CUDA cores and bounds the kernel (bf16->f32
// conversions + FMAs at ~2 IPC).  Here the depthwise 3x3 is 9 taps x 4 channel groups of tcgen05.mma
// (M = 128 pixels, N = 16, K = 16) with a 16x16 DIAGONAL weight block as the B operand; the tap shift is an offset
// of the A descriptor's start address inside the expanded tile in shared memory (UMMA's 128-byte swizzle is a
// function of the absolute smem address, so a start address advanced by whole 128-byte pixel rows stays
// consistent -- measured for conv_tc's HALO mode).  The CUDA cores only move accumulators: TMEM -> +bias, ReLU6 ->
// bf16 -> smem, twice per 64-channel chunk.
//
// Addressing that makes every tap a contiguous run of 128 A rows:
//   stride 1: the expanded tile is stored row-major over the halo tile (IW = 18 columns); output pixel (oh, ow) is
//             accumulator row r = oh*18 + ow (2 dead columns per row, TH = 7 rows -> 126 of 128 rows), and tap
//             (dh, dw) reads rows r + dh*18 + dw.
//   stride 2: the expanded tile is de-interleaved into 4 parity planes (ih&1, iw&1), each row-major with 17
//             columns; r = oh*17 + ow and tap (dh, dw) reads plane (dh&1, dw&1) at rows r + (dh>>1)*17 + (dw>>1).
// The project GEMM and the output epilogue use the same row index r.
//
// Roles: warp 0 TMA producer (halo tile of x once per tile; per chunk We[64][Cin] and Wp[Cout][64]); warp 1 issues every
// tcgen05.mma; warps 2..9 do the two accumulator hand-offs and the output, and write the 9 x 64 diagonal entries of the
// depthwise B tiles of each chunk (the tiles are zeroed once; streaming them by TMA cost 18 KB per chunk, which left room
// for a single weight stage and serialised expand(c+1) behind project(c)).
#include "common.cuh"

// experiment, built only into tools/_build/libb200seg_tools.so (tools/build_tools.py); not part of the product ABI
extern "C" int b200seg_mbconv_tc(const void* x, const void* w_exp, const float* b_exp, const float* w_dw, const float* b_dw,
                                 const void* w_proj, const float* b_proj, int residual, void* y, int B, int H, int W, int Cin,
                                 int Ce, int Cout, int stride, int flags, b200seg_stream_t s);

namespace b200 {

namespace {

constexpr int MT_THREADS = 320;
constexpr int MT_CWARPS = 8;
constexpr int WD_BYTES = 9 * 2048;      // 9 taps x [16 rows][64 k] bf16

struct MtArgs {
  const __nv_bfloat16* x;    // [B][H][W][Cin]
  __nv_bfloat16* y;          // [B][Ho][Wo][Cout]
  const float* b_exp;        // [ce_chunks*64]
  const float* w_dw;         // [9][ce_chunks*64] f32 (rounded to bf16 when written into the B tiles)
  const float* b_dw;         // [ce_chunks*64]
  const float* b_proj;       // [cout_pad]
  int B, H, W, Ho, Wo, Cin, Ce, Cout;
  int residual;
  int kcn;                   // 64-channel chunks of Cin
  int ce_chunks;             // 64-channel chunks of Ce
  int cout_pad, n_proj, proj_n;
  int nbuf_e, nbuf_d, nws;
  int x_chunk_stride, x_bytes, we_bytes, wp_bytes;
  int tiles_w, tiles_h, total_tiles;
  int tmem_cols;
};

template <int S, int TH_>
struct GeoT {
  static constexpr int TW = 16, TH = TH_;
  static constexpr int IW = (TW - 1) * S + 3, IH = (TH - 1) * S + 3;
  static constexpr int NHALO = IW * IH;
  static constexpr int MX = (NHALO + 127) / 128;
  static constexpr int PW = S == 1 ? IW : (IW + 1) / 2;          // plane width (accumulator row pitch)
  static constexpr int PH = S == 1 ? IH : (IH + 1) / 2;
  static constexpr int PSR = (PW * PH + 7) / 8 * 8;              // plane stride in rows (keeps the swizzle phase)
  static constexpr int NPL = S * S;
  static constexpr int E_BYTES = (NPL * PSR * 128 + 1023) / 1024 * 1024;
};

__device__ __forceinline__ bool mt_try_wait_park(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mt_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mt_try_wait_park(bar, parity)) {
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
      printf("b200seg: mbconv_tc mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x,
             parity);
      __trap();
    }
  }
}

struct RingT {
  int i;
  uint32_t ph;
  __device__ __forceinline__ void next(int n) {
    if (++i == n) { i = 0; ph ^= 1u; }
  }
};

__device__ __forceinline__ uint32_t relu6_pack2(float lo, float hi) {
  uint32_t r;
  asm("{\n\t.reg .b32 t;\n\t"
      "cvt.rn.relu.bf16x2.f32 t, %2, %1;\n\t"
      "min.bf16x2 %0, t, %3;\n\t}"
      : "=r"(r)
      : "f"(lo), "f"(hi), "r"(0x40C040C0u));
  return r;
}

// 32 accumulator columns of one TMEM lane -> +bias, ReLU6 -> bf16 -> 64 bytes of a 128B-swizzled smem row
__device__ __forceinline__ void acc32_to_smem(const uint32_t taddr, const float (&bias)[32], uint8_t* row_ptr,
                                              const int row_idx, const int half, const bool write, const bool zero) {
  uint32_t v[2][16];
  tmem_ld16(taddr, v[0]);
  tmem_ld16(taddr + 16u, v[1]);
  tmem_ld_wait();
  if (write) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        pk[i] = relu6_pack2(__uint_as_float(v[cc][2 * i]) + bias[cc * 16 + 2 * i],
                            __uint_as_float(v[cc][2 * i + 1]) + bias[cc * 16 + 2 * i + 1]);
      if (zero) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
      }
      const int j = half * 4 + cc * 2;
      *reinterpret_cast<uint4*>(row_ptr + ((j ^ (row_idx & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(row_ptr + (((j + 1) ^ (row_idx & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
}

__device__ __forceinline__ void load_bias32(const float* p, float (&b)[32]) {
  const float4* bp = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b4 = __ldg(bp + i);
    b[4 * i] = b4.x; b[4 * i + 1] = b4.y; b[4 * i + 2] = b4.z; b[4 * i + 3] = b4.w;
  }
}

template <int MINB, int S, int TH>
__global__ void __launch_bounds__(MT_THREADS, MINB)
mbconv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                 const __grid_constant__ CUtensorMap tmWp, const MtArgs a) {
  using G = GeoT<S, TH>;
  constexpr int MX = G::MX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- shared memory carve-up (every operand tile 1024-byte aligned).  M tiles / tap windows may run past the
  // rows that were written (into the next region): those accumulator rows are never read back. ----
  uint8_t* sX = smem;                                         // kcn chunks x [halo px][64] bf16
  uint8_t* sE = sX + a.kcn * a.x_chunk_stride;                // expanded tile (planes), 64 channels
  uint8_t* sD = sE + G::E_BYTES;                              // nbuf_d x [128][64] bf16
  uint8_t* sWd = sD + a.nbuf_d * 16384;                       // 9 diagonal tap blocks [16 rows][64 k] (written in place)
  uint8_t* sW = sWd + WD_BYTES;                               // nws x (We chunk | Wp chunk)
  const int w_stage = a.we_bytes + a.wp_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + a.nws * w_stage);
  uint64_t* x_full = bars;           // [1] TMA
  uint64_t* x_empty = bars + 1;      // [1] commit (last expand of the tile)
  uint64_t* w_full = bars + 2;       // [2] TMA
  uint64_t* w_empty = bars + 4;      // [2] commit (project)
  uint64_t* e_full = bars + 6;       // [2] commit (expand)          accumulator E ready
  uint64_t* e_empty = bars + 8;      // [2] 8 warps                  accumulator E drained
  uint64_t* es_full = bars + 10;     // [1] 8 warps                  smem E written
  uint64_t* dw_done = bars + 11;     // [1] commit (depthwise)       accumulator DW ready, smem E free
  uint64_t* dwa_empty = bars + 12;   // [1] 8 warps                  accumulator DW drained
  uint64_t* d_full = bars + 13;      // [2] 8 warps                  smem D written
  uint64_t* d_empty = bars + 15;     // [2] commit (project)
  uint64_t* p_full = bars + 17;      // [1] commit (last project)
  uint64_t* p_empty = bars + 18;     // [1] 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  for (int i = threadIdx.x; i < WD_BYTES / 16; i += MT_THREADS) reinterpret_cast<uint4*>(sWd)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWe);
    tma_prefetch_desc(&tmWp);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
      mbar_init(&e_full[i], 1);
      mbar_init(&e_empty[i], MT_CWARPS);
      mbar_init(&d_full[i], MT_CWARPS);
      mbar_init(&d_empty[i], 1);
    }
    mbar_init(es_full, MT_CWARPS);
    mbar_init(dw_done, 1);
    mbar_init(dwa_empty, MT_CWARPS);
    mbar_init(p_full, 1);
    mbar_init(p_empty, MT_CWARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t dwcol = (uint32_t)(a.nbuf_e * MX * 64);      // depthwise accumulator (64 columns)
  const uint32_t pcol0 = dwcol + 64u;                         // project accumulator
  const int nc = a.ce_chunks;

  if (warp == 0) {
    // ================= TMA producer =================
    uint32_t tph = 0;
    RingT wr = {0, 0u};
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, tph ^= 1u) {
      int t = tile;
      const int w0 = (t % a.tiles_w) * G::TW; t /= a.tiles_w;
      const int h0 = (t % a.tiles_h) * G::TH;
      const int bb = t / a.tiles_h;
      mt_wait(x_empty, tph ^ 1u, 10);
      if (lane == 0) {
        mbar_arrive_expect_tx(x_full, (uint32_t)(a.kcn * a.x_bytes));
        for (int kc = 0; kc < a.kcn; ++kc)
          tma_load_4d(sX + kc * a.x_chunk_stride, &tmX, x_full, kc * 64, w0 * S - 1, h0 * S - 1, bb);
      }
      __syncwarp();
      for (int c = 0; c < nc; ++c) {
        mt_wait(&w_empty[wr.i], wr.ph ^ 1u, 11);
        if (lane == 0) {
          uint8_t* st = sW + wr.i * w_stage;
          mbar_arrive_expect_tx(&w_full[wr.i], (uint32_t)w_stage);
          for (int kc = 0; kc < a.kcn; ++kc) tma_load_3d(st + kc * 8192, &tmWe, &w_full[wr.i], kc * 64, 0, c * 64);
          for (int j = 0; j < a.n_proj; ++j)
            tma_load_3d(st + a.we_bytes + j * a.proj_n * 128, &tmWp, &w_full[wr.i], c * 64, 0, j * a.proj_n);
        }
        __syncwarp();
        wr.next(a.nws);
      }
    }
  } else if (warp == 1) {
    // ================= tcgen05 issuer =================
    const uint32_t idesc_e = umma_idesc_bf16(128, 64);
    const uint32_t idesc_d = umma_idesc_bf16(128, 16);
    const uint32_t idesc_p = umma_idesc_bf16(128, a.proj_n);
    const uint32_t desc_hi = (uint32_t)(umma_desc_k128(0) >> 32);
    const uint32_t x_lo0 = (uint32_t)umma_desc_k128(smem_u32(sX));
    const uint32_t e_lo0 = (uint32_t)umma_desc_k128(smem_u32(sE));
    const uint32_t d_lo0 = (uint32_t)umma_desc_k128(smem_u32(sD));
    const uint32_t w_lo0 = (uint32_t)umma_desc_k128(smem_u32(sW));
    const uint32_t wd_lo0 = (uint32_t)umma_desc_k128(smem_u32(sWd));
    const uint32_t x_chunk16 = (uint32_t)a.x_chunk_stride >> 4;
    const uint32_t w_stage16 = (uint32_t)w_stage >> 4, we16 = (uint32_t)a.we_bytes >> 4;
    uint32_t tph = 0, uph = 0;                       // tile phase, unit phase (es_full / dwa_empty flip every chunk)
    RingT we_r = {0, 0u}, e_r = {0, 0u};            // expand side (runs one chunk ahead)
    RingT wc_r = {0, 0u}, d_r = {0, 0u};            // depthwise / project side

    auto expand = [&](const int c) {
      mt_wait(&w_full[we_r.i], we_r.ph, 23);
      mt_wait(&e_empty[e_r.i], e_r.ph ^ 1u, 24);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t bl0 = w_lo0 + (uint32_t)we_r.i * w_stage16;
        for (int m = 0; m < MX; ++m) {
          const uint32_t tacc = tmem_base + (uint32_t)((e_r.i * MX + m) * 64);
          uint32_t first = 0u;
          for (int kc = 0; kc < a.kcn; ++kc) {
            const int ks = (min(64, a.Cin - kc * 64) + 15) >> 4;
            const uint32_t al = x_lo0 + (uint32_t)kc * x_chunk16 + (uint32_t)m * (16384u >> 4);
            const uint32_t bl = bl0 + (uint32_t)kc * (8192u >> 4);
            for (int k = 0; k < ks; ++k) {
              umma_bf16_lohi(tacc, al + 2u * k, bl + 2u * k, desc_hi, idesc_e, first);
              first = 1u;
            }
          }
        }
        umma_commit(&e_full[e_r.i]);
        if (c == nc - 1) umma_commit(x_empty);
      }
      __syncwarp();
      e_r.next(a.nbuf_e);
      we_r.next(a.nws);
    };
    auto depthwise = [&](const int c) {
      mt_wait(es_full, uph, 25);                    // smem E of this chunk written
      mt_wait(dwa_empty, uph ^ 1u, 26);             // previous chunk's depthwise accumulator drained
      tc_fence_after();
      if (elect_one()) {
        const int ks = (min(64, a.Ce - c * 64) + 15) >> 4;      // 16-channel groups that exist in this chunk
        const uint32_t bl0 = wd_lo0;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh)
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const int tap = dh * 3 + dw;
            const int row_off = S == 1 ? dh * G::PW + dw
                                       : ((dh & 1) * 2 + (dw & 1)) * G::PSR + (dh >> 1) * G::PW + (dw >> 1);
            const uint32_t al = e_lo0 + (uint32_t)row_off * 8u;
            const uint32_t bl = bl0 + (uint32_t)tap * (2048u >> 4);
            for (int kk = 0; kk < ks; ++kk)
              umma_bf16_lohi(tmem_base + dwcol + (uint32_t)(kk * 16), al + 2u * kk, bl + 2u * kk, desc_hi, idesc_d,
                             tap > 0 ? 1u : 0u);
          }
        umma_commit(dw_done);
      }
      __syncwarp();
    };
    auto project = [&](const int c) {
      mt_wait(&d_full[d_r.i], d_r.ph, 20);
      if (c == 0) mt_wait(p_empty, tph ^ 1u, 21);   // previous tile's output has been read
      tc_fence_after();
      if (elect_one()) {
        const int ks = (min(64, a.Ce - c * 64) + 15) >> 4;
        const uint32_t al = d_lo0 + (uint32_t)d_r.i * (16384u >> 4);
        const uint32_t bl = w_lo0 + (uint32_t)wc_r.i * w_stage16 + we16;
        for (int j = 0; j < a.n_proj; ++j)
          for (int k = 0; k < ks; ++k)
            umma_bf16_lohi(tmem_base + pcol0 + (uint32_t)(j * a.proj_n), al + 2u * k,
                           bl + (uint32_t)(j * a.proj_n * 8) + 2u * k, desc_hi, idesc_p, (c > 0 || k > 0) ? 1u : 0u);
        umma_commit(&d_empty[d_r.i]);
        umma_commit(&w_empty[wc_r.i]);
        if (c == nc - 1) umma_commit(p_full);
      }
      __syncwarp();
      d_r.next(a.nbuf_d);
      wc_r.next(a.nws);
    };

    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, tph ^= 1u) {
      mt_wait(x_full, tph, 22);
      expand(0);
      for (int c = 0; c < nc; ++c) {
        depthwise(c);
        if (c + 1 < nc && a.nws > 1) expand(c + 1);      // tensor pipe keeps working while the CUDA cores drain chunk c
        project(c);
        if (c + 1 < nc && a.nws == 1) expand(c + 1);     // single weight stage: chunk c must release it first
        uph ^= 1u;
      }
    }
  } else {
    // ================= compute warps =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;       // which 32 of the chunk's 64 columns this warp moves
    uint32_t tph = 0, uph = 0;
    RingT e_r = {0, 0u}, d_r = {0, 0u};
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, tph ^= 1u) {
      int t = tile;
      const int w0 = (t % a.tiles_w) * G::TW; t /= a.tiles_w;
      const int h0 = (t % a.tiles_h) * G::TH;
      const int bb = t / a.tiles_h;
      for (int c = 0; c < nc; ++c, uph ^= 1u) {
        float bias[32];
        // ---- diagonal entries of this chunk's 9 depthwise B tiles (the previous chunk's depthwise MMAs are complete:
        // every thread passed dw_done before its last D hand-off) ----
        {
          const int ct = threadIdx.x - 64;
          const int cstride = a.ce_chunks * 64;
          for (int idx = ct; idx < 9 * 64; idx += MT_CWARPS * 32) {
            const int tap = idx >> 6, ch = idx & 63;
            const int row = ch & 15, j = ch >> 3;
            const __nv_bfloat16 v = __float2bfloat16_rn(__ldg(a.w_dw + tap * cstride + c * 64 + ch));
            *reinterpret_cast<__nv_bfloat16*>(sWd + tap * 2048 + row * 128 + ((j ^ (row & 7)) << 4) + (ch & 7) * 2) = v;
          }
        }
        // ---- accumulator E -> smem E (zero outside the image: the depthwise conv pads the EXPANDED tensor) ----
        load_bias32(a.b_exp + c * 64 + half * 32, bias);
        mt_wait(&e_full[e_r.i], e_r.ph, 30);
        tc_fence_after();
#pragma unroll
        for (int m = 0; m < MX; ++m) {
          const int p = m * 128 + q * 32 + lane;
          const int ih = p / G::IW, iw = p - ih * G::IW;
          const int gh = h0 * S - 1 + ih, gw = w0 * S - 1 + iw;
          const bool inside = gh >= 0 && gh < a.H && gw >= 0 && gw < a.W;
          const int er = S == 1 ? p : ((ih & 1) * 2 + (iw & 1)) * G::PSR + (ih >> 1) * G::PW + (iw >> 1);
          acc32_to_smem(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((e_r.i * MX + m) * 64 + half * 32), bias,
                        sE + er * 128, er, half, p < G::NHALO, !inside);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&e_empty[e_r.i]);
          mbar_arrive(es_full);
        }
        e_r.next(a.nbuf_e);
        // ---- accumulator DW -> smem D ----
        load_bias32(a.b_dw + c * 64 + half * 32, bias);
        mt_wait(dw_done, uph, 31);
        mt_wait(&d_empty[d_r.i], d_r.ph ^ 1u, 32);
        tc_fence_after();
        {
          const int r = q * 32 + lane;
          acc32_to_smem(tmem_base + ((uint32_t)(q * 32) << 16) + dwcol + (uint32_t)(half * 32), bias,
                        sD + d_r.i * 16384 + r * 128, r, half, true, false);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(dwa_empty);
          mbar_arrive(&d_full[d_r.i]);
        }
        d_r.next(a.nbuf_d);
      }
      // ---- accumulator P -> global ----
      mt_wait(p_full, tph, 33);
      tc_fence_after();
      {
        const int r = q * 32 + lane;
        const int oh = r / G::PW, ow = r - oh * G::PW;
        const int gh = h0 + oh, gw = w0 + ow;
        const bool ok = oh < G::TH && ow < G::TW && gh < a.Ho && gw < a.Wo;
        const long long pix = ((long long)bb * a.Ho + gh) * a.Wo + gw;
        __nv_bfloat16* yp = a.y + pix * a.Cout;
        const __nv_bfloat16* rp = a.x + pix * a.Cin;        // residual: stride 1 and Cin == Cout
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + pcol0;
        for (int ch = half; ch < (a.cout_pad >> 4); ch += 2) {
          const int c0 = ch * 16;
          uint32_t v[16];
          tmem_ld16(trow + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.b_proj + c0 + i));
            f[i] = __uint_as_float(v[i]) + b4.x;
            f[i + 1] = __uint_as_float(v[i + 1]) + b4.y;
            f[i + 2] = __uint_as_float(v[i + 2]) + b4.z;
            f[i + 3] = __uint_as_float(v[i + 3]) + b4.w;
          }
          const bool ok0 = ok && c0 < a.Cout, ok1 = ok && c0 + 8 < a.Cout;     // Cout % 8 == 0
          if (a.residual) {
            if (ok0) {
              const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(rp + c0));
              f[0] += bf16lo(t4.x); f[1] += bf16hi(t4.x); f[2] += bf16lo(t4.y); f[3] += bf16hi(t4.y);
              f[4] += bf16lo(t4.z); f[5] += bf16hi(t4.z); f[6] += bf16lo(t4.w); f[7] += bf16hi(t4.w);
            }
            if (ok1) {
              const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(rp + c0 + 8));
              f[8] += bf16lo(t4.x); f[9] += bf16hi(t4.x); f[10] += bf16lo(t4.y); f[11] += bf16hi(t4.y);
              f[12] += bf16lo(t4.z); f[13] += bf16hi(t4.z); f[14] += bf16lo(t4.w); f[15] += bf16hi(t4.w);
            }
          }
          if (ok0)
            *reinterpret_cast<uint4*>(yp + c0) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                            pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          if (ok1)
            *reinterpret_cast<uint4*>(yp + c0 + 8) = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                                                                pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

template <int S, int TH>
int launch_mt(MtArgs a, const void* x, const void* w_exp, const void* w_proj, int flags, cudaStream_t st) {
  using G = GeoT<S, TH>;
  a.x_bytes = G::NHALO * 128;
  a.x_chunk_stride = (a.x_bytes + 1023) / 1024 * 1024;
  a.tiles_w = (a.Wo + G::TW - 1) / G::TW;
  a.tiles_h = (a.Ho + G::TH - 1) / G::TH;
  const long long tiles = (long long)a.B * a.tiles_h * a.tiles_w;
  B200_REQUIRE(tiles < (1LL << 31), "mbconv_tc: too many tiles");
  a.total_tiles = (int)tiles;

  const int cap = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/;
  const int stage = a.we_bytes + a.wp_bytes;
  auto smem_for = [&](int nd, int nw) { return a.kcn * a.x_chunk_stride + G::E_BYTES + nd * 16384 + WD_BYTES + nw * stage; };
  // two CTAs per SM when both fit (<= 256 TMEM columns, half the shared memory): one CTA's accumulator hand-offs
  // overlap the other's MMAs
  const int half_cap = cap / 2 - 1024;
  int per_sm = 1;
  a.nbuf_e = 1; a.nbuf_d = 1; a.nws = 2;
  if (G::MX * 64 + 64 + a.cout_pad <= 256) {
    if (smem_for(1, 2) <= half_cap) per_sm = 2;
    else if (smem_for(1, 1) <= half_cap) { per_sm = 2; a.nws = 1; }
  }
  if (((flags >> 6) & 3) == 1) per_sm = 1;
  if (per_sm == 1) {
    a.nbuf_e = (2 * G::MX * 64 + 64 + a.cout_pad <= 512) ? 2 : 1;
    a.nbuf_d = 2; a.nws = 2;
    if (smem_for(a.nbuf_d, a.nws) > cap) a.nbuf_d = 1;
    if (smem_for(a.nbuf_d, a.nws) > cap) a.nws = 1;
  }
  if (flags & 3) a.nbuf_e = flags & 3;
  if ((flags >> 2) & 3) a.nbuf_d = (flags >> 2) & 3;
  if ((flags >> 4) & 3) a.nws = (flags >> 4) & 3;
  B200_REQUIRE(a.nbuf_e <= 2 && a.nbuf_d <= 2 && a.nws <= 2, "mbconv_tc: at most 2 buffers per ring");
  if (smem_for(a.nbuf_d, a.nws) > cap) a.nbuf_d = 1;          // requested depths are upper bounds
  if (smem_for(a.nbuf_d, a.nws) > cap) a.nws = 1;
  if (a.nbuf_e * G::MX * 64 + 64 + a.cout_pad > 512) a.nbuf_e = 1;
  const int smem_used = smem_for(a.nbuf_d, a.nws);
  B200_REQUIRE(smem_used <= cap, "mbconv_tc: Cin=%d Ce=%d Cout=%d stride=%d needs %d B of shared memory", a.Cin, a.Ce,
               a.Cout, S, smem_used);
  const int tmem_need = a.nbuf_e * G::MX * 64 + 64 + a.cout_pad;
  B200_REQUIRE(tmem_need <= 512, "mbconv_tc: %d TMEM columns needed", tmem_need);
  a.tmem_cols = 32;
  while (a.tmem_cols < tmem_need) a.tmem_cols <<= 1;
  int smem = smem_used + 1024 + 256;
  if (a.tmem_cols > 256 || smem > cap / 2) per_sm = 1;
  if (per_sm == 1 && smem < 116 * 1024) smem = 116 * 1024;      // a >256-column CTA must own the SM

  CUtensorMap tmX, tmWe, tmWp;
  {
    uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.Cin * 2 * a.W, (uint64_t)a.Cin * 2 * a.W * a.H};
    uint32_t box[4] = {64, (uint32_t)G::IW, (uint32_t)G::IH, 1};
    int rc = make_tmap_bf16(&tmX, x, 4, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a.Cin, 1, (uint64_t)a.Ce};
    uint64_t str[2] = {(uint64_t)a.Cin * 2, (uint64_t)a.Cin * 2};
    uint32_t box[3] = {64, 1, 64};
    int rc = make_tmap_bf16(&tmWe, w_exp, 3, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a.Ce, 1, (uint64_t)a.Cout};
    uint64_t str[2] = {(uint64_t)a.Ce * 2, (uint64_t)a.Ce * 2};
    uint32_t box[3] = {64, 1, (uint32_t)a.proj_n};
    int rc = make_tmap_bf16(&tmWp, w_proj, 3, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mbconv_tc_kernel<1, S, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mbconv_tc_kernel<2, S, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "mbconv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  long long grid = (long long)sm_count() * per_sm;
  if ((flags >> 8) & 0xff) grid = (long long)((flags >> 8) & 0xff) * 4;
  if (grid > a.total_tiles) grid = a.total_tiles;
  if (per_sm == 2)
    mbconv_tc_kernel<2, S, TH><<<(unsigned)grid, MT_THREADS, (size_t)smem, st>>>(tmX, tmWe, tmWp, a);
  else
    mbconv_tc_kernel<1, S, TH><<<(unsigned)grid, MT_THREADS, (size_t)smem, st>>>(tmX, tmWe, tmWp, a);
  return check_launch("mbconv_tc");
}

}  // namespace

}  // namespace b200

using namespace b200;

// flags: bits 0-1 expand accumulator buffers (0 = auto), bits 2-3 D buffers, bits 4-5 weight stages,
//        bits 6-7 CTAs per SM (1 = force one), bits 8-15 grid/4, bits 16-17 tile rows (1 = 7, 2 = 4)
extern "C" int b200seg_mbconv_tc(const void* x, const void* w_exp, const float* b_exp, const float* w_dw,
                                 const float* b_dw, const void* w_proj, const float* b_proj, int residual, void* y,
                                 int B, int H, int W, int Cin, int Ce, int Cout, int stride, int flags,
                                 b200seg_stream_t s) {
  B200_REQUIRE(x && w_exp && b_exp && w_dw && b_dw && w_proj && b_proj && y, "mbconv_tc: null pointer");
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "mbconv_tc: empty tensor");
  B200_REQUIRE(stride == 1 || stride == 2, "mbconv_tc: stride=%d (1 or 2)", stride);
  B200_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cin <= 192, "mbconv_tc: Cin=%d must be a multiple of 8, <= 192", Cin);
  B200_REQUIRE(Ce > 0 && Ce % 8 == 0, "mbconv_tc: Ce=%d must be a multiple of 8", Ce);
  B200_REQUIRE(Cout > 0 && Cout % 8 == 0 && Cout <= 320, "mbconv_tc: Cout=%d must be a multiple of 8, <= 320", Cout);
  B200_REQUIRE(!residual || (stride == 1 && Cin == Cout), "mbconv_tc: residual needs stride 1 and Cin == Cout");
  MtArgs a;
  a.x = (const __nv_bfloat16*)x; a.y = (__nv_bfloat16*)y;
  a.b_exp = b_exp; a.w_dw = w_dw; a.b_dw = b_dw; a.b_proj = b_proj;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Ce = Ce; a.Cout = Cout; a.residual = residual;
  a.Ho = (H - 1) / stride + 1; a.Wo = (W - 1) / stride + 1;          // k=3, pad=1
  a.kcn = (Cin + 63) / 64;
  a.ce_chunks = (Ce + 63) / 64;
  a.cout_pad = (Cout + 15) & ~15;
  a.n_proj = a.cout_pad > 256 ? 2 : 1;
  a.proj_n = a.cout_pad / a.n_proj;
  B200_REQUIRE(a.proj_n % 16 == 0, "mbconv_tc: Cout=%d cannot be split into UMMA N tiles", Cout);
  a.we_bytes = a.kcn * 8192;
  a.wp_bytes = a.cout_pad * 128;
  // 7-row tiles fill 126 of the 128 accumulator rows; maps of <= 8 rows (or a forced flag) use 4-row tiles
  int th = a.Ho <= 8 ? 4 : 7;
  if (stride == 2 && a.kcn > 1) th = 4;                 // the 15x33 halo tile of a 7-row stride-2 tile x 2 chunks does not fit
  if ((flags >> 16) & 3) th = ((flags >> 16) & 3) == 1 ? 7 : 4;
  cudaStream_t st = (cudaStream_t)s;
  if (stride == 1) return th == 7 ? launch_mt<1, 7>(a, x, w_exp, w_proj, flags, st)
                                  : launch_mt<1, 4>(a, x, w_exp, w_proj, flags, st);
  return th == 7 ? launch_mt<2, 7>(a, x, w_exp, w_proj, flags, st)
                 : launch_mt<2, 4>(a, x, w_exp, w_proj, flags, st);
}
