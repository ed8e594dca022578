// Fused MobileNetV2 inverted-residual block for sm_100a (eval mode, BatchNorm folded, bf16 activations):
//
//     y = project_1x1( relu6( dw3x3_s( relu6( expand_1x1(x) ) ) ) )  [+ x]
//
// replaces the three ConvBNActivation launches of torchvision's InvertedResidual (tv:models/mobilenetv2.py:38-62,
// reached from the reference through unet.py:15-19,34-42).  The 6x-expanded activation -- 70 % of the unfused
// forward's HBM bytes -- never leaves the SM: HBM traffic of a block is its input tile (+halo) and its output.
//
// One CTA = one output tile (8x16 pixels at stride 1, 4x16 at stride 2), persistent over tiles.  The expanded
// channels Ce are processed in chunks of 64:
//   warp 0      TMA producer: the input halo tile X[(TH-1)s+3][(TW-1)s+3][Cin] once per tile (image borders are
//               TMA out-of-bounds zero fill), then per chunk the expand weights We[64][Cin] and the project
//               weights Wp[Cout][64] into a small stage ring.
//   warp 1      tcgen05 issuer.  expand(c):  E_acc[halo px, 64] = X . We^T   (M = 128 x MX, N = 64, K = Cin)
//                                project(c): P_acc[128 px, Cout] += D_c . Wp^T (M = 128, N = Cout, K = 64)
//               expand(c+1) is issued before project(c) so the tensor pipe works while the CUDA cores do chunk c.
//   warps 2..9  (256 threads) per chunk: E_acc (TMEM) -> +bias, ReLU6, zero outside the image (the depthwise conv
//               pads the EXPANDED tensor with zeros) -> bf16 -> smem E;  depthwise 3x3 from smem E in registers
//               (thread = 8 channels x 4|2 adjacent outputs, fp32 accumulate) -> +bias, ReLU6 -> bf16 -> smem D in
//               the 128B-swizzled K-major layout the project MMA reads.  After the last chunk: P_acc -> +bias
//               (+ residual) -> bf16 -> global.
// All hand-offs are mbarriers (TMA complete_tx, tcgen05.commit, or one arrival per compute warp).
#include "common.cuh"

namespace b200 {

namespace {

constexpr int MB_THREADS = 320;
// Row pitch of the expanded tile in shared memory.  Only CUDA cores touch it (E-phase stores: lane = pixel row; stencil
// loads: 8 lanes = the 8 16-byte channel groups of one pixel), so instead of the 128-byte XOR swizzle the rows are padded by
// 16 bytes: consecutive rows start in different bank groups (conflict-free stores) and every stencil load is base + immediate
// (the 18 swizzled addresses per thread were hoisted out of the chunk loop and spilled under the 96-register cap).
constexpr int MB_E_PITCH = 144;
constexpr int MB_CWARPS = 8;

struct MbArgs {
  const __nv_bfloat16* x;    // [B][H][W][Cin]
  __nv_bfloat16* y;          // [B][Ho][Wo][Cout]
  const float* b_exp;        // [ce_chunks*64]
  const __nv_bfloat16* w_dw; // [9][ce_chunks*64] bf16 taps (mixed-precision FMA: f32 += bf16 * bf16, no unpack)
  const float* b_dw;         // [ce_chunks*64]
  const float* b_proj;       // [cout_pad]
  int B, H, W, Ho, Wo, Cin, Ce, Cout;
  int stride, residual;
  int TH, TW, IH, IW;        // output tile, input halo tile
  int MX;                    // 128-row M tiles covering IH*IW
  int kcn;                   // 64-channel chunks of Cin
  int ce_chunks;             // 64-channel chunks of Ce
  int cout_pad, n_proj, proj_n;   // project N: n_proj MMAs of proj_n columns
  int nbuf_e, nbuf_d, nws;
  int last_nv;               // valid channels of the last Ce chunk handled exactly: 16 | 32, else 64 (= full-width path)
  int x_chunk_stride, x_bytes, e_bytes, we_bytes, wp_bytes;
  int tiles_w, tiles_h;
  long long total_tiles;
  int tmem_cols;
};

// Wait that parks the warp in hardware (try_wait with a suspend-time hint) instead of spinning: in the first
// version ~27 % of all issued instructions were wait-loop instructions of the producer / issuer warps competing
// with the compute warps of their scheduler.  2 s watchdog as everywhere else.
__device__ __forceinline__ bool mb_try_wait_park(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mb_try_wait_park(bar, parity)) {
    if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
      printf("b200seg: mbconv mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

struct Ring {          // index + phase of an n-deep mbarrier ring, advanced without div/mod
  int i;
  uint32_t ph;
  __device__ __forceinline__ void next(int n) {
    if (++i == n) { i = 0; ph ^= 1u; }
  }
};

// bf16x2( min(max(x, 0), 6) ): the ReLU rides on the conversion, the upper clamp is one packed min
// (rounding is monotonic and 6.0 is a bf16 number, so this equals rounding the fp32 clamp).
__device__ __forceinline__ uint32_t relu6_pack(float lo, float hi) {
  uint32_t r;
  asm("{\n\t.reg .b32 t;\n\t"
      "cvt.rn.relu.bf16x2.f32 t, %2, %1;\n\t"
      "min.bf16x2 %0, t, %3;\n\t}"
      : "=r"(r)
      : "f"(lo), "f"(hi), "r"(0x40C040C0u));
  return r;
}

template <int S, int TH_>
struct Geo {
  static constexpr int TW = 16, TH = TH_;
  static constexpr int IW = (TW - 1) * S + 3, IH = (TH - 1) * S + 3;
  static constexpr int NHALO = IW * IH;
  static constexpr int MX = (NHALO + 127) / 128;
};

// GSH = log2(valid 8-channel groups of the chunk): a narrow last chunk (Ce % 64 = 32 | 16) spreads its pixels over 2x | 4x
// as many threads instead of computing 64 - nv dead channels (Ce = 96 and 144 -- the three most expensive blocks -- would
// otherwise pay for 128 and 192 channels).
template <int S, int TH, int GSH>
__device__ __forceinline__ void dw_chunk(const uint8_t* __restrict__ sE, uint8_t* __restrict__ sD, const MbArgs& a,
                                         const int chunk, const int ct) {
  constexpr int NPT = 256 >> GSH;                 // threads along the pixel axis
  constexpr int PXT = (TH * 16) / NPT > 0 ? (TH * 16) / NPT : 1;   // adjacent outputs per thread
  constexpr int TPR = 16 / PXT;                   // threads per output row
  constexpr int NCOL = (PXT - 1) * S + 3;         // input columns they touch
  constexpr int IW = Geo<S, TH>::IW, TW = Geo<S, TH>::TW;
  const int g = ct & ((1 << GSH) - 1), pt = ct >> GSH;
  if (pt * PXT >= TH * 16) return;                // more threads than outputs (4-row tile, 16-channel chunk)
  const int orow = pt / TPR;
  const int ocol0 = (pt % TPR) * PXT;
  const int cbase = chunk * 64 + g * 8;
  const int cstride = a.ce_chunks * 64;
  float acc[PXT][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.b_dw + cbase));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.b_dw + cbase + 4));
#pragma unroll
    for (int o = 0; o < PXT; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {
    uint4 wt[3];
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) wt[dw] = __ldg(reinterpret_cast<const uint4*>(a.w_dw + (dh * 3 + dw) * cstride + cbase));
    const int prow0 = (orow * S + dh) * IW + ocol0 * S;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
      const int pr = prow0 + j;
      const uint4 t = *reinterpret_cast<const uint4*>(sE + pr * MB_E_PITCH + (g << 4));
#pragma unroll
      for (int o = 0; o < PXT; ++o) {
        const int dw = j - o * S;                 // compile-time after unrolling
        if (dw >= 0 && dw < 3) {
          const uint4 wv = wt[dw];
          float* ac = acc[o];
          ac[0] = fma_bf16_lo(t.x, wv.x, ac[0]); ac[1] = fma_bf16_hi(t.x, wv.x, ac[1]);
          ac[2] = fma_bf16_lo(t.y, wv.y, ac[2]); ac[3] = fma_bf16_hi(t.y, wv.y, ac[3]);
          ac[4] = fma_bf16_lo(t.z, wv.z, ac[4]); ac[5] = fma_bf16_hi(t.z, wv.z, ac[5]);
          ac[6] = fma_bf16_lo(t.w, wv.w, ac[6]); ac[7] = fma_bf16_hi(t.w, wv.w, ac[7]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < PXT; ++o) {
    const int r = orow * TW + ocol0 + o;          // row of the project A operand
    *reinterpret_cast<uint4*>(sD + r * 128 + ((g ^ (r & 7)) << 4)) =
        make_uint4(relu6_pack(acc[o][0], acc[o][1]), relu6_pack(acc[o][2], acc[o][3]),
                   relu6_pack(acc[o][4], acc[o][5]), relu6_pack(acc[o][6], acc[o][7]));
  }
}

// E_acc (TMEM) -> ReLU6 -> bf16 -> smem E for one chunk.  Warp (q, half) converts lane quarter q of every M tile, columns
// [half * NV/2, +NV/2) of the chunk's NV valid channels.  The expand bias is already in the accumulator (it enters the
// MMA as an extra k-step against an all-ones A tile), so a value costs half a cvt.relu + half a packed min.  The loads
// are software-pipelined: the tcgen05.ld of stage i+1 is in flight while stage i is converted.
template <int S, int TH, int NV>
__device__ __forceinline__ void e_chunk(uint8_t* __restrict__ sE, const uint32_t tcol, const int q, const int half,
                                        const int lane, const int gh0, const int gw0, const int H, const int W) {
  using G = Geo<S, TH>;
  constexpr int CW = NV / 2;                      // columns per warp
  constexpr int LW = CW >= 16 ? 16 : 8;           // columns per tcgen05.ld
  constexpr int NL = CW / LW;
  constexpr int NST = G::MX * NL;
  uint32_t v[2][16];
  auto load = [&](const int st, uint32_t (&r)[16]) {
    const uint32_t ad = tcol + (uint32_t)((st / NL) * 64 + half * CW + (st % NL) * LW);
    if (LW == 16) tmem_ld16(ad, r); else tmem_ld8(ad, r);
  };
  load(0, v[0]);
#pragma unroll
  for (int st = 0; st < NST; ++st) {
    tmem_ld_wait();
    if (st + 1 < NST) load(st + 1, v[(st + 1) & 1]);
    const uint32_t(&r)[16] = v[st & 1];
    const int p = (st / NL) * 128 + q * 32 + lane;
    if (p < G::NHALO) {
      const int ih = p / G::IW, iw = p - ih * G::IW;
      const int gh = gh0 + ih, gw = gw0 + iw;
      const bool inside = gh >= 0 && gh < H && gw >= 0 && gw < W;      // the depthwise conv pads E with zeros
      uint32_t pk[LW / 2];
#pragma unroll
      for (int i = 0; i < LW / 2; ++i)
        pk[i] = inside ? relu6_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])) : 0u;
      uint8_t* erow = sE + p * MB_E_PITCH;
      const int j = (half * CW + (st % NL) * LW) >> 3;
      *reinterpret_cast<uint4*>(erow + (j << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (LW == 16)
        *reinterpret_cast<uint4*>(erow + ((j + 1) << 4)) = make_uint4(pk[LW / 2 - 4], pk[LW / 2 - 3], pk[LW / 2 - 2], pk[LW / 2 - 1]);
    }
  }
}

constexpr int MB_BIAS_TILE = 2048;    // [64 expanded channels][16 k] bf16, no swizzle: 8 row groups x (2 core matrices of 128 B)
constexpr int MB_ONES_BYTES = 256;    // all-ones A operand of the bias k-step: every core matrix aliases these bytes (SBO = 0)

template <int MINB, int S, int TH>
__global__ void __launch_bounds__(MB_THREADS, MINB)
mbconv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
              const __grid_constant__ CUtensorMap tmWp, const MbArgs a) {
  using G = Geo<S, TH>;
  constexpr int MX = G::MX;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- shared memory carve-up (every operand tile 1024-byte aligned) ----
  // The last 128-row M tile of an X chunk extends past the rows TMA fills (into the next chunk / into E): those
  // accumulator rows are never read back, only the addresses have to stay inside the allocation.
  uint8_t* sX = smem;                                                        // kcn chunks, rows = halo pixels
  uint8_t* sE = sX + a.kcn * a.x_chunk_stride;                               // expanded tile, 64 channels
  uint8_t* sD = sE + a.e_bytes;                                              // nbuf_d x [128][64] bf16
  uint8_t* sW = sD + a.nbuf_d * 16384;                                       // nws x (We chunk | Wp chunk)
  const int w_stage = a.we_bytes + a.wp_bytes + MB_BIAS_TILE;                // + the chunk's expand-bias tile (built in place)
  uint8_t* sOnes = sW + a.nws * w_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + MB_ONES_BYTES);
  uint64_t* x_full = bars;          // [1]
  uint64_t* x_empty = bars + 1;     // [1]
  uint64_t* w_full = bars + 2;      // [2]  expand side of a weight stage: We chunk + bias tile
  uint64_t* w_empty = bars + 4;     // [2]
  uint64_t* e_full = bars + 6;      // [2]
  uint64_t* e_empty = bars + 8;     // [2]
  uint64_t* d_full = bars + 10;     // [2]
  uint64_t* d_empty = bars + 12;    // [2]
  uint64_t* p_full = bars + 14;     // [1]
  uint64_t* p_empty = bars + 15;    // [1]
  uint64_t* wp_full = bars + 16;    // [2]  project side of a weight stage (Wp chunk): its own barriers, so that the next
  uint64_t* wp_empty = bars + 18;   // [2]  chunk's We never waits for this chunk's project MMA (single-stage configurations)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWe);
    tma_prefetch_desc(&tmWp);
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
      mbar_init(&wp_full[i], 1);
      mbar_init(&wp_empty[i], 1);
      mbar_init(&e_full[i], 1);
      mbar_init(&e_empty[i], MB_CWARPS);
      mbar_init(&d_full[i], MB_CWARPS);
      mbar_init(&d_empty[i], 1);
    }
    mbar_init(p_full, 1);
    mbar_init(p_empty, MB_CWARPS);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  if (warp == 2) {                     // all-ones A tile (bf16 1.0) and the zero half (k = 8..15) of every bias tile
    for (int i = lane; i < MB_ONES_BYTES / 4; i += 32) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    for (int st = 0; st < a.nws; ++st) {
      uint32_t* bt = reinterpret_cast<uint32_t*>(sW + st * w_stage + a.we_bytes + a.wp_bytes);
      for (int i = lane; i < MB_BIAS_TILE / 4; i += 32) bt[i] = 0u;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t pcol0 = (uint32_t)(a.nbuf_e * MX * 64);     // first column of the project accumulator
  const int nc = a.ce_chunks;
  const int total_tiles = (int)a.total_tiles;

  if (warp == 0) {
    // ================= TMA producer =================
    uint32_t tph = 0;
    Ring wr = {0, 0u};
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tph ^= 1u) {
      int t = tile;
      const int w0 = (t % a.tiles_w) * G::TW; t /= a.tiles_w;
      const int h0 = (t % a.tiles_h) * G::TH;
      const int bb = t / a.tiles_h;
      mb_wait(x_empty, tph ^ 1u, 10);
      if (lane == 0) {
        mbar_arrive_expect_tx(x_full, (uint32_t)(a.kcn * a.x_bytes));
        for (int kc = 0; kc < a.kcn; ++kc)
          tma_load_4d(sX + kc * a.x_chunk_stride, &tmX, x_full, kc * 64, w0 * S - 1, h0 * S - 1, bb);
      }
      __syncwarp();
      for (int c = 0; c < nc; ++c) {
        mb_wait(&w_empty[wr.i], wr.ph ^ 1u, 11);
        {
          // bias tile of the chunk: row n (expanded channel) holds {hi, lo, 0...} with hi + lo = b_exp[n] to 16 mantissa
          // bits; as B operand against the all-ones A tile it adds the bias inside the expand MMA
          uint8_t* bt = sW + wr.i * w_stage + a.we_bytes + a.wp_bytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = lane + 32 * h;
            const float b = __ldg(a.b_exp + c * 64 + n);
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            *reinterpret_cast<uint32_t*>(bt + (n >> 3) * 256 + (n & 7) * 16) =
                (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
          }
          fence_proxy_async_smem();
          __syncwarp();
        }
        if (lane == 0) {
          uint8_t* st = sW + wr.i * w_stage;
          mbar_arrive_expect_tx(&w_full[wr.i], (uint32_t)a.we_bytes);
          for (int kc = 0; kc < a.kcn; ++kc) tma_load_3d(st + kc * 8192, &tmWe, &w_full[wr.i], kc * 64, 0, c * 64);
        }
        mb_wait(&wp_empty[wr.i], wr.ph ^ 1u, 12);
        if (lane == 0) {
          uint8_t* st = sW + wr.i * w_stage;
          mbar_arrive_expect_tx(&wp_full[wr.i], (uint32_t)a.wp_bytes);
          for (int j = 0; j < a.n_proj; ++j)
            tma_load_3d(st + a.we_bytes + j * a.proj_n * 128, &tmWp, &wp_full[wr.i], c * 64, 0, j * a.proj_n);
        }
        __syncwarp();
        wr.next(a.nws);
      }
    }
  } else if (warp == 1) {
    // ================= tcgen05 issuer =================
    const uint32_t idesc_e = umma_idesc_bf16(128, 64);
    const uint32_t idesc_el = umma_idesc_bf16(128, a.last_nv);        // last chunk: only its valid channels
    const uint64_t ones_desc = umma_desc_nosw(smem_u32(sOnes), 128, 0);
    const uint32_t idesc_p = umma_idesc_bf16(128, a.proj_n);
    const uint32_t desc_hi = (uint32_t)(umma_desc_k128(0) >> 32);
    const uint32_t x_lo0 = (uint32_t)umma_desc_k128(smem_u32(sX));
    const uint32_t d_lo0 = (uint32_t)umma_desc_k128(smem_u32(sD));
    const uint32_t w_lo0 = (uint32_t)umma_desc_k128(smem_u32(sW));
    const uint32_t x_chunk16 = (uint32_t)a.x_chunk_stride >> 4;
    const uint32_t w_stage16 = (uint32_t)w_stage >> 4, we16 = (uint32_t)a.we_bytes >> 4;
    uint32_t tph = 0, pe_ph = 0;
    int pend = -1;                            // chunk whose project MMA is still to be issued (may belong to the previous tile)
    Ring we_r = {0, 0u}, e_r = {0, 0u};       // expand side: weight stage, accumulator buffer
    Ring wp_r = {0, 0u}, d_r = {0, 0u};       // project side (runs one chunk behind): weight stage, D buffer

    auto project = [&](const int c) {
      mb_wait(&wp_full[wp_r.i], wp_r.ph, 25);
      mb_wait(&d_full[d_r.i], d_r.ph, 20);
      if (c == 0) { mb_wait(p_empty, pe_ph ^ 1u, 21); pe_ph ^= 1u; }     // previous tile's output has been read
      tc_fence_after();
      if (elect_one()) {
        const int ks = (min(64, a.Ce - c * 64) + 15) >> 4;
        const uint32_t al = d_lo0 + (uint32_t)d_r.i * (16384u >> 4);
        const uint32_t bl = w_lo0 + (uint32_t)wp_r.i * w_stage16 + we16;
        for (int j = 0; j < a.n_proj; ++j)
          for (int k = 0; k < ks; ++k)
            umma_bf16_lohi(tmem_base + pcol0 + (uint32_t)(j * a.proj_n), al + 2u * k,
                           bl + (uint32_t)(j * a.proj_n * 8) + 2u * k, desc_hi, idesc_p, (c > 0 || k > 0) ? 1u : 0u);
        umma_commit(&d_empty[d_r.i]);
        umma_commit(&wp_empty[wp_r.i]);
        if (c == nc - 1) umma_commit(p_full);
      }
      __syncwarp();
      d_r.next(a.nbuf_d);
      wp_r.next(a.nws);
    };

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tph ^= 1u) {
      mb_wait(x_full, tph, 22);
      for (int c = 0; c < nc; ++c) {
        mb_wait(&w_full[we_r.i], we_r.ph, 23);
        mb_wait(&e_empty[e_r.i], e_r.ph ^ 1u, 24);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bl0 = w_lo0 + (uint32_t)we_r.i * w_stage16;
          const uint32_t id_e = c == nc - 1 ? idesc_el : idesc_e;
          const uint64_t bias_desc = umma_desc_nosw(smem_u32(sW) + (uint32_t)(we_r.i * w_stage + a.we_bytes + a.wp_bytes), 128, 256);
          for (int m = 0; m < MX; ++m) {
            const uint32_t tacc = tmem_base + (uint32_t)((e_r.i * MX + m) * 64);
            umma_bf16(tacc, ones_desc, bias_desc, id_e, 0u);           // E_acc = 1 . bias^T
            for (int kc = 0; kc < a.kcn; ++kc) {
              const int ks = (min(64, a.Cin - kc * 64) + 15) >> 4;
              const uint32_t al = x_lo0 + (uint32_t)kc * x_chunk16 + (uint32_t)m * (16384u >> 4);
              const uint32_t bl = bl0 + (uint32_t)kc * (8192u >> 4);
              for (int k = 0; k < ks; ++k) umma_bf16_lohi(tacc, al + 2u * k, bl + 2u * k, desc_hi, id_e, 1u);
            }
          }
          umma_commit(&e_full[e_r.i]);
          umma_commit(&w_empty[we_r.i]);
          if (c == nc - 1) umma_commit(x_empty);
        }
        __syncwarp();
        e_r.next(a.nbuf_e);
        we_r.next(a.nws);
        // the expand of the NEXT chunk -- also across a tile boundary -- is always in flight before the project of this one
        // is issued, so the E hand-off of the next chunk/tile never waits for project -> commit -> expand in series
        if (pend >= 0) project(pend);
        pend = c;
      }
    }
    if (pend >= 0) project(pend);
  } else {
    // ================= compute warps =================
    const int ct = threadIdx.x - 64;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;       // which 32 of the chunk's 64 columns this warp converts
    // ---- P_acc -> global (+bias, +residual) for the tile (bb, h0, w0) ----
    auto p_epilogue = [&](const int bb, const int h0, const int w0, const uint32_t parity) {
      mb_wait(p_full, parity, 32);
      tc_fence_after();

        const int r = q * 32 + lane;
        const int oh = r / G::TW, ow = r - oh * G::TW;
        const int gh = h0 + oh, gw = w0 + ow;
        const bool ok = r < G::TH * G::TW && gh < a.Ho && gw < a.Wo;
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_empty);
    };
    uint32_t tph = 0;
    int pbb = 0, ph0 = 0, pw0 = 0;
    bool have_prev = false;
    Ring e_r = {0, 0u}, d_r = {0, 0u};
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tph ^= 1u) {
      int t = tile;
      const int w0 = (t % a.tiles_w) * G::TW; t /= a.tiles_w;
      const int h0 = (t % a.tiles_h) * G::TH;
      const int bb = t / a.tiles_h;
      for (int c = 0; c < nc; ++c) {
        // ---- E_acc -> smem E ----
        const bool narrow = c == nc - 1 && a.last_nv != 64;
        mb_wait(&e_full[e_r.i], e_r.ph, 30);
        tc_fence_after();
        {
          const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(e_r.i * MX * 64);
          if (!narrow) e_chunk<S, TH, 64>(sE, tcol, q, half, lane, h0 * S - 1, w0 * S - 1, a.H, a.W);
          else if (a.last_nv == 32) e_chunk<S, TH, 32>(sE, tcol, q, half, lane, h0 * S - 1, w0 * S - 1, a.H, a.W);
          else e_chunk<S, TH, 16>(sE, tcol, q, half, lane, h0 * S - 1, w0 * S - 1, a.H, a.W);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_empty[e_r.i]);
        e_r.next(a.nbuf_e);
        // the previous tile's output leaves here: its last project MMA had the whole E hand-off above to complete
        if (c == 0 && have_prev) p_epilogue(pbb, ph0, pw0, tph ^ 1u);
        named_bar_sync(1, MB_CWARPS * 32);                 // E complete
        // ---- depthwise 3x3: smem E -> smem D ----
        mb_wait(&d_empty[d_r.i], d_r.ph ^ 1u, 31);
        if (!narrow) dw_chunk<S, TH, 3>(sE, sD + d_r.i * 16384, a, c, ct);
        else if (a.last_nv == 32) dw_chunk<S, TH, 2>(sE, sD + d_r.i * 16384, a, c, ct);
        else dw_chunk<S, TH, 1>(sE, sD + d_r.i * 16384, a, c, ct);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_full[d_r.i]);
        d_r.next(a.nbuf_d);
        named_bar_sync(2, MB_CWARPS * 32);                 // every read of E done before the next chunk overwrites it
      }
      pbb = bb; ph0 = h0; pw0 = w0; have_prev = true;
    }
    if (have_prev) p_epilogue(pbb, ph0, pw0, tph ^ 1u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

}  // namespace b200

using namespace b200;

// flags: bits 0-1 expand accumulator buffers (0 = auto), bits 2-3 D buffers, bits 4-5 weight stages,
//        bits 6-7 CTAs per SM (1 = force one), bits 8-15 grid/4, bits 16-17 stride-1 tile rows (1 = 8, 2 = 4)
extern "C" int b200seg_mbconv(const void* x, const void* w_exp, const float* b_exp, const void* w_dw,
                              const float* b_dw, const void* w_proj, const float* b_proj, int residual, void* y,
                              int B, int H, int W, int Cin, int Ce, int Cout, int stride, int flags,
                              b200seg_stream_t s) {
  B200_REQUIRE(x && w_exp && b_exp && w_dw && b_dw && w_proj && b_proj && y, "mbconv: null pointer");
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "mbconv: empty tensor");
  B200_REQUIRE(stride == 1 || stride == 2, "mbconv: stride=%d (1 or 2)", stride);
  B200_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cin <= 192, "mbconv: Cin=%d must be a multiple of 8, <= 192", Cin);
  B200_REQUIRE(Ce > 0 && Ce % 8 == 0, "mbconv: Ce=%d must be a multiple of 8", Ce);
  B200_REQUIRE(Cout > 0 && Cout % 8 == 0 && Cout <= 320, "mbconv: Cout=%d must be a multiple of 8, <= 320", Cout);
  B200_REQUIRE(!residual || (stride == 1 && Cin == Cout), "mbconv: residual needs stride 1 and Cin == Cout");
  MbArgs a;
  a.x = (const __nv_bfloat16*)x; a.y = (__nv_bfloat16*)y;
  a.b_exp = b_exp; a.w_dw = reinterpret_cast<const __nv_bfloat16*>(w_dw); a.b_dw = b_dw; a.b_proj = b_proj;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Ce = Ce; a.Cout = Cout; a.stride = stride; a.residual = residual;
  a.Ho = (H - 1) / stride + 1; a.Wo = (W - 1) / stride + 1;          // k=3, pad=1
  a.TW = 16; a.TH = stride == 1 ? 8 : 4;
  // small maps: 4-row tiles double the tile count (more CTAs busy, finer balance, half the smem -> 2 CTAs/SM)
  if (stride == 1 && (long long)B * ((a.Ho + 7) / 8) * ((a.Wo + 15) / 16) < (long long)sm_count()) a.TH = 4;
  if (stride == 1 && ((flags >> 16) & 3)) a.TH = ((flags >> 16) & 3) == 1 ? 8 : 4;
  a.IH = (a.TH - 1) * stride + 3; a.IW = (a.TW - 1) * stride + 3;
  a.MX = (a.IH * a.IW + 127) / 128;
  a.kcn = (Cin + 63) / 64;
  a.ce_chunks = (Ce + 63) / 64;
  a.cout_pad = (Cout + 15) & ~15;
  a.n_proj = a.cout_pad > 256 ? 2 : 1;
  a.proj_n = a.cout_pad / a.n_proj;
  B200_REQUIRE(a.proj_n % 16 == 0, "mbconv: Cout=%d cannot be split into UMMA N tiles", Cout);
  a.x_bytes = a.IH * a.IW * 128;
  a.x_chunk_stride = round_up(a.x_bytes, 1024);
  a.e_bytes = round_up(a.IH * a.IW * MB_E_PITCH, 1024);
  a.last_nv = (Ce % 64 == 16 || Ce % 64 == 32) ? Ce % 64 : 64;
  a.we_bytes = a.kcn * 8192;
  a.wp_bytes = a.cout_pad * 128;
  a.tiles_w = (a.Wo + a.TW - 1) / a.TW;
  a.tiles_h = (a.Ho + a.TH - 1) / a.TH;
  a.total_tiles = (long long)B * a.tiles_h * a.tiles_w;

  const int cap = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - MB_ONES_BYTES;
  const int x_region = a.kcn * a.x_chunk_stride;      // the over-read of the last M tile lands in E (see kernel)
  auto smem_for = [&](int nd, int nw) { return x_region + a.e_bytes + nd * 16384 + nw * (a.we_bytes + a.wp_bytes + MB_BIAS_TILE); };
  // Two CTAs per SM (each single-buffered, <= 256 TMEM columns, <= half the shared memory) hide one CTA's phase
  // barriers and TMEM/LDS latency behind the other; otherwise one CTA per SM with double-buffered rings.
  const int half_cap = cap / 2 - 1024;
  int per_sm = 1;
  a.nbuf_e = 1; a.nbuf_d = 1; a.nws = 2;
  if (a.MX * 64 + a.cout_pad <= 256) {
    if (smem_for(1, 2) <= half_cap) per_sm = 2;
    else if (smem_for(1, 1) <= half_cap) { per_sm = 2; a.nws = 1; }
  }
  if (((flags >> 6) & 3) == 1) per_sm = 1;
  if (per_sm == 1) {
    a.nbuf_e = (2 * a.MX * 64 + a.cout_pad <= 512) ? 2 : 1;
    a.nbuf_d = 2; a.nws = 2;
    if (smem_for(a.nbuf_d, a.nws) > cap) a.nbuf_d = 1;
    if (smem_for(a.nbuf_d, a.nws) > cap) a.nws = 1;
  }
  if (flags & 3) a.nbuf_e = flags & 3;
  if ((flags >> 2) & 3) a.nbuf_d = (flags >> 2) & 3;
  if ((flags >> 4) & 3) a.nws = (flags >> 4) & 3;
  B200_REQUIRE(a.nbuf_e <= 2 && a.nbuf_d <= 2 && a.nws <= 2, "mbconv: at most 2 buffers per ring");
  const int smem_used = smem_for(a.nbuf_d, a.nws);
  B200_REQUIRE(smem_used <= cap, "mbconv: Cin=%d Ce=%d Cout=%d needs %d B of shared memory", Cin, Ce, Cout, smem_used);
  const int tmem_need = a.nbuf_e * a.MX * 64 + a.cout_pad;
  B200_REQUIRE(tmem_need <= 512, "mbconv: %d TMEM columns needed", tmem_need);
  a.tmem_cols = 32;
  while (a.tmem_cols < tmem_need) a.tmem_cols <<= 1;
  int smem = smem_used + 1024 + 256 + MB_ONES_BYTES;
  if (a.tmem_cols > 256 || smem > cap / 2) per_sm = 1;
  // a CTA that owns more than half of the SM's 512 TMEM columns must not share the SM (the second CTA's
  // tcgen05.alloc would block until the first exits)
  if (per_sm == 1 && smem < 116 * 1024) smem = 116 * 1024;

  CUtensorMap tmX, tmWe, tmWp;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * W, (uint64_t)Cin * 2 * W * H};
    uint32_t box[4] = {64, (uint32_t)a.IW, (uint32_t)a.IH, 1};
    int rc = make_tmap_bf16(&tmX, x, 4, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, 1, (uint64_t)Ce};
    uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2};
    uint32_t box[3] = {64, 1, 64};
    int rc = make_tmap_bf16(&tmWe, w_exp, 3, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Ce, 1, (uint64_t)Cout};
    uint64_t str[2] = {(uint64_t)Ce * 2, (uint64_t)Ce * 2};
    uint32_t box[3] = {64, 1, (uint32_t)a.proj_n};
    int rc = make_tmap_bf16(&tmWp, w_proj, 3, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaSuccess;
    const void* fns[6] = {(const void*)mbconv_kernel<1, 1, 8>, (const void*)mbconv_kernel<1, 1, 4>,
                          (const void*)mbconv_kernel<1, 2, 4>, (const void*)mbconv_kernel<2, 1, 8>,
                          (const void*)mbconv_kernel<2, 1, 4>, (const void*)mbconv_kernel<2, 2, 4>};
    for (int i = 0; i < 6 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "mbconv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  long long grid = (long long)sm_count() * per_sm;
  if ((flags >> 8) & 0xff) grid = (long long)((flags >> 8) & 0xff) * 4;
  if (grid > a.total_tiles) grid = a.total_tiles;
  const dim3 gd((unsigned)grid), bd(MB_THREADS);
  cudaStream_t st = (cudaStream_t)s;
#define MB_LAUNCH(MINB, S, TH) mbconv_kernel<MINB, S, TH><<<gd, bd, (size_t)smem, st>>>(tmX, tmWe, tmWp, a)
  if (per_sm == 2) {
    if (stride == 2) MB_LAUNCH(2, 2, 4);
    else if (a.TH == 8) MB_LAUNCH(2, 1, 8);
    else MB_LAUNCH(2, 1, 4);
  } else {
    if (stride == 2) MB_LAUNCH(1, 2, 4);
    else if (a.TH == 8) MB_LAUNCH(1, 1, 8);
    else MB_LAUNCH(1, 1, 4);
  }
#undef MB_LAUNCH
  return check_launch("mbconv");
}
