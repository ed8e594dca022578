// conv_tc.cu -- dense 1x1 / 3x3 convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 pixels, N couts] (fp32, TMEM) = sum over (tap, cin-chunk) A[128, 64] * B[N, 64]^T
//
// * A tiles are fetched by TMA straight out of the NHWC activation tensor: a 4-D tiled tensor map
//   (C, W, H, B) with box {64, BW, BH, 1}; the 3x3 taps are the same box shifted by (dw, dh), and
//   the zero padding, the ragged image border and the Cin tail are all produced by TMA's
//   out-of-bounds zero fill -- there is no im2col buffer and no predicated gather.
//   The box lands in shared memory as BW*BH rows of 128 bytes with the 128-byte swizzle, which is
//   exactly the canonical K-major UMMA operand layout.
// * B tiles come from the packed weights bf16 [Cout][taps][Cin] (3-D map, box {64, 1, N}).
// * One elected thread issues tcgen05.mma (M=128, N=block_n, K=16 per instruction); accumulators
//   live in TMEM; a ring of `stages` smem slots is recycled through tcgen05.commit -> mbarrier.
// * Epilogue: 4 warps read TMEM (32 lanes each), add the folded-BN shift, apply ReLU/ReLU6, add the
//   residual tile (TMA-loaded into the staging buffer up front), convert to bf16, write the
//   swizzled staging tile, and one thread TMA-stores it (the store clips the image border and the
//   Cout tail).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).
#include "common.cuh"

namespace b200 {

struct ConvTcArgs {
  const float* bias;
  const void* res;   // raw pointers, used only by the direct-store epilogue
  void* y;
  int taps, Cin, Cout;
  int W, H, B;       // geometry the tile scheduler walks (1x1: W = #pixels, H = B = 1)
  int BW, BH;        // M tile = BH rows x BW columns of pixels, BW*BH <= 128
  int tiles_w, tiles_h;
  int block_n, n_tiles;
  int k_chunks;      // ceil(Cin / 64)
  int stages;
  int act;
  int has_res;
  int direct_store;
  int tmem_cols;
};

constexpr int TC_THREADS = 192;
constexpr int A_STAGE_BYTES = 128 * 128;  // 128 rows x 64 bf16

__global__ void __launch_bounds__(TC_THREADS, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_stage_bytes = a.block_n * 128;
  const int n_boxes = (a.block_n + 63) >> 6;
  uint8_t* sA = smem;
  uint8_t* sB = sA + a.stages * A_STAGE_BYTES;
  uint8_t* sOut = sB + a.stages * b_stage_bytes;
  float* sBias = reinterpret_cast<float*>(sOut + n_boxes * A_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + 256);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;
  uint64_t* res_full = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);

  // ---- tile coordinates ----
  const int n_tile = blockIdx.x % a.n_tiles;
  int mt = blockIdx.x / a.n_tiles;
  const int tw = mt % a.tiles_w; mt /= a.tiles_w;
  const int th = mt % a.tiles_h;
  const int bb = mt / a.tiles_h;
  const int w0 = tw * a.BW, h0 = th * a.BH, n0 = n_tile * a.block_n;
  const int rows = a.BW * a.BH;
  const int num_kb = a.taps * a.k_chunks;

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!a.direct_store) tma_prefetch_desc(&tmC);
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(res_full, 1);
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

  if (warp == 0) {
    // ================= TMA producer =================
    if (a.has_res && !a.direct_store) {
      if (lane == 0) {
        int nb = 0;
        for (int bx = 0; bx < n_boxes; ++bx)
          if (n0 + bx * 64 < a.Cout) ++nb;
        mbar_arrive_expect_tx(res_full, (uint32_t)(nb * rows * 128));
        for (int bx = 0; bx < n_boxes; ++bx)
          if (n0 + bx * 64 < a.Cout) tma_load_4d(sOut + bx * A_STAGE_BYTES, &tmR, res_full, n0 + bx * 64, w0, h0, bb);
      }
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % a.stages;
      const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
      mbar_wait(&empty[s], ph ^ 1u, 1);
      if (lane == 0) {
        const int tap = kb / a.k_chunks, chunk = kb - tap * a.k_chunks;
        int dw = 0, dh = 0;
        if (a.taps == 9) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
        mbar_arrive_expect_tx(&full[s], (uint32_t)(rows * 128 + b_stage_bytes));
        tma_load_4d(sA + s * A_STAGE_BYTES, &tmA, &full[s], chunk * 64, w0 + dw, h0 + dh, bb);
        tma_load_3d(sB + s * b_stage_bytes, &tmB, &full[s], chunk * 64, tap, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = umma_idesc_bf16(128, a.block_n);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % a.stages;
      const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
      mbar_wait(&full[s], ph, 2);
      tc_fence_after();
      if (lane == 0) {
        const int chunk = kb % a.k_chunks;
        const int kvalid = min(64, a.Cin - chunk * 64);
        const int ksteps = (kvalid + 15) >> 4;
        const uint64_t adesc = umma_desc_k128(smem_u32(sA + s * A_STAGE_BYTES));
        const uint64_t bdesc = umma_desc_k128(smem_u32(sB + s * b_stage_bytes));
        for (int k = 0; k < ksteps; ++k) {
          // advancing 16 bf16 = 32 bytes along K inside the swizzled row: +2 in 16-byte units
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);                      // slot reusable once these MMAs retire
        if (kb == num_kb - 1) umma_commit(acc_full); // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int et = threadIdx.x - 64;       // 0..127
    const int q = warp & 3;                // TMEM lane quadrant this warp may read
    const int r = q * 32 + lane;           // accumulator row = pixel within the tile
    for (int i = et; i < a.block_n; i += 128) sBias[i] = (a.bias && n0 + i < a.Cout) ? a.bias[n0 + i] : 0.f;
    named_bar_sync(1, 128);
    mbar_wait(acc_full, 0, 3);
    tc_fence_after();
    if (a.has_res && !a.direct_store) mbar_wait(res_full, 0, 4);

    const int hl = r / a.BW, wl = r - hl * a.BW;
    const bool row_ok = r < rows && (h0 + hl) < a.H && (w0 + wl) < a.W;
    const long long pix = ((long long)bb * a.H + (h0 + hl)) * a.W + (w0 + wl);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);

    for (int c0 = 0; c0 < a.block_n; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)c0, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = apply_act_rt(__uint_as_float(v[i]) + sBias[c0 + i], a.act);
      if (!a.direct_store) {
        const int bx = c0 >> 6;
        const int j0 = (c0 & 63) >> 3;
        uint8_t* rowp = sOut + bx * A_STAGE_BYTES + r * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint4* p = reinterpret_cast<uint4*>(rowp + (((j0 + hh) ^ (r & 7)) << 4));
          float* g = f + hh * 8;
          if (a.has_res) {
            const uint4 t = *p;
            g[0] += bf16lo(t.x); g[1] += bf16hi(t.x); g[2] += bf16lo(t.y); g[3] += bf16hi(t.y);
            g[4] += bf16lo(t.z); g[5] += bf16hi(t.z); g[6] += bf16lo(t.w); g[7] += bf16hi(t.w);
          }
          uint4 o;
          o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
          o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
          *p = o;
        }
      } else if (row_ok) {
        __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + pix * a.Cout + n0 + c0;
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.res) + pix * a.Cout + n0 + c0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (n0 + c0 + hh * 8 < a.Cout) {   // Cout % 8 == 0: whole 16-byte groups are in or out
            float* g = f + hh * 8;
            if (a.has_res) {
              const uint4 t = *reinterpret_cast<const uint4*>(rp + hh * 8);
              g[0] += bf16lo(t.x); g[1] += bf16hi(t.x); g[2] += bf16lo(t.y); g[3] += bf16hi(t.y);
              g[4] += bf16lo(t.z); g[5] += bf16hi(t.z); g[6] += bf16lo(t.w); g[7] += bf16hi(t.w);
            }
            uint4 o;
            o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
            o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
            *reinterpret_cast<uint4*>(yp + hh * 8) = o;
          }
        }
      }
    }
    if (!a.direct_store) {
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the TMA engine
      named_bar_sync(1, 128);
      if (et == 0) {
        for (int bx = 0; bx < n_boxes; ++bx)
          if (n0 + bx * 64 < a.Cout) tma_store_4d(&tmC, sOut + bx * A_STAGE_BYTES, n0 + bx * 64, w0, h0, bb);
        tma_store_commit();
        tma_store_wait_all();
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

static int pick_tile(int W, int H, int* BW, int* BH) {
  // maximise the fraction of the 128 accumulator rows that hold real pixels
  double best = -1.0;
  int bw_best = 1, bh_best = 1;
  const int wmax = W < 128 ? W : 128;
  for (int bw = wmax; bw >= 1; --bw) {
    int bh = 128 / bw;
    if (bh > H) bh = H;
    if (bh < 1) bh = 1;
    const long long tiles = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh);
    const double eff = (double)W * H / ((double)tiles * 128.0);
    if (eff > best + 1e-9) { best = eff; bw_best = bw; bh_best = bh; }
    if (bw <= 8 && best > 0) break;
  }
  *BW = bw_best; *BH = bh_best;
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200seg_conv_tc(const void* x, const void* w, const float* bias, const void* res, void* y, int B,
                               int H, int W, int Cin, int Cout, int taps, int act, int flags,
                               b200seg_stream_t s) {
  B200_REQUIRE(taps == 1 || taps == 9, "conv_tc: taps=%d (1 or 9)", taps);
  B200_REQUIRE(Cin > 0 && Cin % 8 == 0, "conv_tc: Cin=%d must be a positive multiple of 8", Cin);
  B200_REQUIRE(Cout > 0 && Cout % 8 == 0, "conv_tc: Cout=%d must be a positive multiple of 8", Cout);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "conv_tc: empty tensor");
  B200_REQUIRE(x && w && y, "conv_tc: null pointer");

  ConvTcArgs a;
  a.bias = bias; a.res = res; a.y = y;
  a.taps = taps; a.Cin = Cin; a.Cout = Cout;
  if (taps == 1) {   // pointwise: pixels are one flat axis
    const long long M = (long long)B * H * W;
    B200_REQUIRE(M < (1ll << 31), "conv_tc: too many pixels");
    a.W = (int)M; a.H = 1; a.B = 1;
  } else {
    a.W = W; a.H = H; a.B = B;
  }
  pick_tile(a.W, a.H, &a.BW, &a.BH);
  a.tiles_w = (a.W + a.BW - 1) / a.BW;
  a.tiles_h = (a.H + a.BH - 1) / a.BH;
  // N tiling: one tile if Cout <= 256 (rounded to the UMMA granule 16); otherwise tiles that are a
  // multiple of 64 wide so that a 64-channel store box never spills into a neighbour tile.
  if (Cout <= 256) {
    a.block_n = (Cout + 15) & ~15; a.n_tiles = 1;
  } else {
    int nt = (Cout + 255) / 256;
    int bn = (((Cout + nt - 1) / nt) + 63) & ~63;
    a.block_n = bn; a.n_tiles = (Cout + bn - 1) / bn;
  }
  a.k_chunks = (Cin + 63) / 64;
  a.act = act;
  a.has_res = res != nullptr;
  a.direct_store = flags & 1;
  a.tmem_cols = 32;
  while (a.tmem_cols < a.block_n) a.tmem_cols <<= 1;

  const int num_kb = taps * a.k_chunks;
  const int n_boxes = (a.block_n + 63) / 64;
  const int per_stage = A_STAGE_BYTES + a.block_n * 128;
  const int fixed = n_boxes * A_STAGE_BYTES + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*align slack*/;
  int budget = (num_kb <= 4) ? 100 * 1024 : 200 * 1024;   // short-K layers: keep >=2 CTAs per SM
  int stages = (budget - fixed) / per_stage;
  if (stages > num_kb) stages = num_kb;
  if (stages > 8) stages = 8;
  if (stages < 1) stages = 1;
  a.stages = stages;
  const int smem = fixed + stages * per_stage;
  B200_REQUIRE(smem <= 227 * 1024, "conv_tc: smem %d too large", smem);

  CUtensorMap tmA, tmB, tmC, tmR;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * a.W, (uint64_t)Cin * 2 * a.W * a.H};
    uint32_t box[4] = {64, (uint32_t)a.BW, (uint32_t)a.BH, 1};
    int rc = make_tmap_bf16(&tmA, x, 4, dims, str, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)taps, (uint64_t)Cout};
    uint64_t str[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * taps};
    uint32_t box[3] = {64, 1, (uint32_t)a.block_n};
    int rc = make_tmap_bf16(&tmB, w, 3, dims, str, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)Cout * 2 * a.W, (uint64_t)Cout * 2 * a.W * a.H};
    uint32_t box[4] = {64, (uint32_t)a.BW, (uint32_t)a.BH, 1};
    int rc = make_tmap_bf16(&tmC, y, 4, dims, str, box, 1);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmR, res ? res : y, 4, dims, str, box, 1);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const long long grid = (long long)a.tiles_w * a.tiles_h * a.B * a.n_tiles;
  B200_REQUIRE(grid < (1ll << 31), "conv_tc: grid too large");
  conv_tc_kernel<<<(unsigned)grid, TC_THREADS, smem, (cudaStream_t)s>>>(tmA, tmB, tmC, tmR, a);
  return check_launch("conv_tc");
}
