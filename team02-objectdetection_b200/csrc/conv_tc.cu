// conv_tc.cu -- dense 1x1 / 3x3 convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 pixels, N couts] (fp32, TMEM) = sum over (tap, cin-chunk) A[128, 64] * B[N, 64]^T
//
// Persistent, warp-specialised kernel: one CTA per SM slot loops over output tiles; three
// pipelines run concurrently
//     TMA producer --(smem ring: full/empty)--> MMA issuer --(TMEM x2: acc_full/acc_empty)--> epilogue
//     epilogue --(staging smem x2)--> TMA store (bulk async group, drained one tile later)
// so the loads of tile i+1, the MMAs of tile i and the stores of tile i-1 overlap.
//
// A operand (activations, NHWC bf16) -- two addressing modes, both pure TMA:
//   TAP   : a 4-D tiled tensor map (C, W, H, B), box {64, BW, BH, 1}; each 3x3 tap is the same box
//           shifted by (dw, dh).  Zero padding, ragged borders and the Cin tail are TMA OOB zero fill.
//           1x1 convs use it with the pixels flattened to one axis (box {64,128,1,1}).
//   HALO  : (3x3, tiles = 128 consecutive pixels of one image row) one box {64, 130, 1, 1} per input
//           row h-1, h, h+1 is loaded ONCE per cin-chunk; the three horizontal taps are the same
//           shared-memory tile read through UMMA descriptors whose start address is advanced by
//           dw*128 bytes (descriptor base_offset = dw keeps the 128-byte swizzle phase right).
//           L2->smem traffic for A drops from 9x to 3 x 130/128.
// B operand (weights bf16 [Cout][taps][Cin]): 3-D map, box {64,1,N}; when all taps/chunks of the N
// tile fit in shared memory they are loaded once per CTA ("resident") instead of once per tile.
//
// Epilogue: TMEM -> registers (tcgen05.ld 32x32b.x16) -> +bias -> ReLU/ReLU6 -> +residual -> bf16 ->
// 128B-swizzled staging tile -> TMA store (clips the image border and the Cout tail).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).
#include "common.cuh"

namespace b200 {

struct ConvTcArgs {
  const float* bias;
  const void* res;   // residual, read straight from global in the epilogue (or NULL)
  void* y;           // only for the direct-store debug epilogue
  int taps, Cin, Cout;
  int W, H, B;       // geometry the tile scheduler walks (1x1: W = #pixels, H = B = 1)
  int BW, BH;        // M tile = BH rows x BW columns of pixels, BW*BH <= 128
  int tiles_w, tiles_h;
  int block_n, n_tiles;
  int k_chunks;      // ceil(Cin / 64)
  int stages;
  int act;
  int direct_store;
  int tmem_cols;
  int halo;          // A addressing mode (see above)
  int b_resident;    // weights loaded once per CTA
  int n_sbuf;        // staging buffers (1 or 2)
  int halo_bo;       // HALO: put the swizzle phase of the shifted start address into descriptor.base_offset
  int a_stage_bytes, b_stage_bytes;
  long long total_tiles;
};

constexpr int TC_THREADS = 192;
constexpr int TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int HALO_BYTES = 17 * 1024;   // 130 rows x 128 B rounded up to the 1024-B swizzle atom

__device__ __forceinline__ uint64_t umma_desc_k128_off(uint32_t smem_addr) {
  // start address not 1024-aligned: base_offset = (addr >> 7) & 7 (PTX ISA, tcgen05 matrix descriptor)
  return umma_desc_k128(smem_addr) | ((uint64_t)((smem_addr >> 7) & 7u) << 49);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const ConvTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_boxes = (a.block_n + 63) >> 6;
  const int kb_per_tile = a.halo ? a.k_chunks * 3 : a.k_chunks * a.taps;
  const int b_tiles_resident = a.taps * a.k_chunks;
  const int b_tile_bytes = a.block_n * 128;

  uint8_t* sA = smem;
  uint8_t* sB = sA + a.stages * a.a_stage_bytes;
  uint8_t* sOut = sB + (a.b_resident ? b_tiles_resident * b_tile_bytes : a.stages * a.b_stage_bytes);
  float* sBias = reinterpret_cast<float*>(sOut + a.n_sbuf * n_boxes * TILE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + 256);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;      // [2]
  uint64_t* acc_empty = acc_full + 2;  // [2]
  uint64_t* b_full = acc_empty + 2;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!a.direct_store) tma_prefetch_desc(&tmC);
    for (int i = 0; i < a.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    mbar_init(b_full, 1);
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
  const int rows = a.BW * a.BH;

  if (warp == 0) {
    // ================= TMA producer =================
    if (a.b_resident && lane == 0) {
      mbar_arrive_expect_tx(b_full, (uint32_t)(b_tiles_resident * b_tile_bytes));
      for (int t = 0; t < a.taps; ++t)
        for (int c = 0; c < a.k_chunks; ++c)
          tma_load_3d(sB + (t * a.k_chunks + c) * b_tile_bytes, &tmB, b_full, c * 64, t, 0);
    }
    long long kb_glob = 0;
    for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int n_tile = (int)(tile % a.n_tiles);
      long long mt = tile / a.n_tiles;
      const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
      const int th = (int)(mt % a.tiles_h);
      const int bb = (int)(mt / a.tiles_h);
      const int w0 = tw * a.BW, h0 = th * a.BH, n0 = n_tile * a.block_n;
      for (int kb = 0; kb < kb_per_tile; ++kb, ++kb_glob) {
        const int s = (int)(kb_glob % a.stages);
        const uint32_t ph = (uint32_t)(kb_glob / a.stages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u, 1);
        if (lane == 0) {
          if (a.halo) {
            const int chunk = kb / 3, dh = kb - chunk * 3;
            const uint32_t bytes = 130u * 128u + (a.b_resident ? 0u : 3u * (uint32_t)b_tile_bytes);
            mbar_arrive_expect_tx(&full[s], bytes);
            tma_load_4d(sA + s * a.a_stage_bytes, &tmA, &full[s], chunk * 64, w0 - 1, h0 + dh - 1, bb);
            if (!a.b_resident)
              for (int dw = 0; dw < 3; ++dw)
                tma_load_3d(sB + s * a.b_stage_bytes + dw * b_tile_bytes, &tmB, &full[s], chunk * 64, dh * 3 + dw, n0);
          } else {
            const int tap = kb / a.k_chunks, chunk = kb - tap * a.k_chunks;
            int dw = 0, dh = 0;
            if (a.taps == 9) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
            mbar_arrive_expect_tx(&full[s], (uint32_t)(rows * 128 + (a.b_resident ? 0 : b_tile_bytes)));
            tma_load_4d(sA + s * a.a_stage_bytes, &tmA, &full[s], chunk * 64, w0 + dw, h0 + dh, bb);
            if (!a.b_resident) tma_load_3d(sB + s * a.b_stage_bytes, &tmB, &full[s], chunk * 64, tap, n0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const uint32_t idesc = umma_idesc_bf16(128, a.block_n);
    if (a.b_resident) mbar_wait(b_full, 0, 5);
    long long kb_glob = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&acc_empty[ab], aph ^ 1u, 6);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(ab * a.block_n);
      for (int kb = 0; kb < kb_per_tile; ++kb, ++kb_glob) {
        const int s = (int)(kb_glob % a.stages);
        const uint32_t ph = (uint32_t)(kb_glob / a.stages) & 1u;
        mbar_wait(&full[s], ph, 2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(sA + s * a.a_stage_bytes);
          if (a.halo) {
            const int chunk = kb / 3, dh = kb - chunk * 3;
            const int ksteps = (min(64, a.Cin - chunk * 64) + 15) >> 4;
            for (int dw = 0; dw < 3; ++dw) {
              const int tap = dh * 3 + dw;
              const uint64_t adesc = a.halo_bo ? umma_desc_k128_off(a_addr + (uint32_t)dw * 128u)
                                                 : umma_desc_k128(a_addr + (uint32_t)dw * 128u);
              const uint32_t b_addr = a.b_resident ? smem_u32(sB + (tap * a.k_chunks + chunk) * b_tile_bytes)
                                                   : smem_u32(sB + s * a.b_stage_bytes + dw * b_tile_bytes);
              const uint64_t bdesc = umma_desc_k128(b_addr);
              for (int k = 0; k < ksteps; ++k)
                umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                          (kb > 0 || dw > 0 || k > 0) ? 1u : 0u);
            }
          } else {
            const int tap = kb / a.k_chunks, chunk = kb - tap * a.k_chunks;
            const int ksteps = (min(64, a.Cin - chunk * 64) + 15) >> 4;
            const uint64_t adesc = umma_desc_k128(a_addr);
            const uint32_t b_addr = a.b_resident ? smem_u32(sB + (tap * a.k_chunks + chunk) * b_tile_bytes)
                                                 : smem_u32(sB + s * a.b_stage_bytes);
            const uint64_t bdesc = umma_desc_k128(b_addr);
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (kb == kb_per_tile - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hl = r / a.BW, wl = r - hl * a.BW;
    int it = 0;
    int cur_n_tile = -1;
    for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      const int n_tile = (int)(tile % a.n_tiles);
      long long mt = tile / a.n_tiles;
      const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
      const int th = (int)(mt % a.tiles_h);
      const int bb = (int)(mt / a.tiles_h);
      const int w0 = tw * a.BW, h0 = th * a.BH, n0 = n_tile * a.block_n;
      const int ab = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      uint8_t* stg = sOut + (a.n_sbuf == 2 ? (it & 1) : 0) * n_boxes * TILE_BYTES;

      if (n_tile != cur_n_tile) {   // (re)stage the bias slice of this N tile
        named_bar_sync(2, 128);     // nobody is still reading the old slice
        for (int i = et; i < a.block_n; i += 128) sBias[i] = (a.bias && n0 + i < a.Cout) ? a.bias[n0 + i] : 0.f;
        cur_n_tile = n_tile;
        named_bar_sync(2, 128);
      }
      mbar_wait(&acc_full[ab], aph, 3);
      tc_fence_after();

      const bool row_ok = r < rows && (h0 + hl) < a.H && (w0 + wl) < a.W;
      const long long pix = ((long long)bb * a.H + (h0 + hl)) * a.W + (w0 + wl);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * a.block_n);
      const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.res) + pix * a.Cout + n0;
      __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(a.y) + pix * a.Cout + n0;

      for (int c0 = 0; c0 < a.block_n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = apply_act_rt(__uint_as_float(v[i]) + sBias[c0 + i], a.act);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float* g = f + hh * 8;
          const bool col_ok = n0 + c0 + hh * 8 < a.Cout;   // Cout % 8 == 0: 16-byte groups are all-in or all-out
          if (a.res != nullptr && row_ok && col_ok) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(rp + c0 + hh * 8));
            g[0] += bf16lo(t.x); g[1] += bf16hi(t.x); g[2] += bf16lo(t.y); g[3] += bf16hi(t.y);
            g[4] += bf16lo(t.z); g[5] += bf16hi(t.z); g[6] += bf16lo(t.w); g[7] += bf16hi(t.w);
          }
          uint4 o;
          o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
          o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
          if (!a.direct_store) {
            const int bx = c0 >> 6;
            const int j = ((c0 & 63) >> 3) + hh;
            *reinterpret_cast<uint4*>(stg + bx * TILE_BYTES + r * 128 + ((j ^ (r & 7)) << 4)) = o;
          } else if (row_ok && col_ok) {
            *reinterpret_cast<uint4*>(yp + c0 + hh * 8) = o;
          }
        }
      }
      // accumulator fully read -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
      if (!a.direct_store) {
        fence_proxy_async_smem();
        // the staging buffer the NEXT tile will write must no longer be read by an in-flight store
        if (et == 0) {
          if (a.n_sbuf == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        named_bar_sync(1, 128);
        if (et == 0) {
          for (int bx = 0; bx < n_boxes; ++bx)
            if (n0 + bx * 64 < a.Cout) tma_store_4d(&tmC, stg + bx * TILE_BYTES, n0 + bx * 64, w0, h0, bb);
          tma_store_commit();
          if (a.n_sbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        if (a.n_sbuf == 1) named_bar_sync(1, 128);
      }
    }
    if (!a.direct_store && et == 0) tma_store_wait_all();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

static void pick_tile(int W, int H, int* BW, int* BH) {
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
}

}  // namespace b200

using namespace b200;

// flags: bit0 direct-store epilogue (debug); bit1 forbid HALO addressing; bit2 forbid resident weights;
//        bit3 single staging buffer; bit4 HALO descriptors WITH base_offset (experiment: wrong on B200); bits 8..15: force grid size = value * 4 CTAs (0 = auto)
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
  a.halo = (taps == 9 && W >= 96 && !(flags & 2)) ? 1 : 0;
  if (taps == 1) {   // pointwise: pixels are one flat axis
    const long long M = (long long)B * H * W;
    B200_REQUIRE(M < (1ll << 31), "conv_tc: too many pixels");
    a.W = (int)M; a.H = 1; a.B = 1;
  } else {
    a.W = W; a.H = H; a.B = B;
  }
  if (a.halo) { a.BW = W < 128 ? W : 128; a.BH = 1; }
  else pick_tile(a.W, a.H, &a.BW, &a.BH);
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
  a.direct_store = flags & 1;
  a.halo_bo = (flags & 16) ? 1 : 0;   // measured on B200: the swizzle uses absolute smem address bits, base_offset must stay 0
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.block_n) a.tmem_cols <<= 1;   // two accumulators

  const int n_boxes = (a.block_n + 63) / 64;
  const int b_tile = a.block_n * 128;
  const int smem_cap = 227 * 1024 - 1024 /*align slack*/ - 1024 /*bias*/ - 256 /*barriers*/;
  a.a_stage_bytes = a.halo ? HALO_BYTES : TILE_BYTES;
  const int kb_per_tile = a.halo ? a.k_chunks * 3 : a.k_chunks * taps;
  const int b_all = taps * a.k_chunks * b_tile;
  a.n_sbuf = (flags & 8) ? 1 : 2;
  if (a.n_sbuf * n_boxes * TILE_BYTES > 96 * 1024) a.n_sbuf = 1;
  const int out_bytes = a.n_sbuf * n_boxes * TILE_BYTES;
  a.b_resident = (a.n_tiles == 1 && !(flags & 4) && b_all <= 80 * 1024 && b_all + out_bytes + 2 * a.a_stage_bytes <= smem_cap) ? 1 : 0;
  a.b_stage_bytes = a.b_resident ? 0 : (a.halo ? 3 * b_tile : b_tile);
  const int per_stage = a.a_stage_bytes + a.b_stage_bytes;
  const int fixed = out_bytes + (a.b_resident ? b_all : 0);
  // short pipelines leave room for a second CTA per SM (more tiles in flight for the HBM-bound layers)
  int want = kb_per_tile <= 3 ? 4 : 6;
  int stages = (smem_cap - fixed) / per_stage;
  if (stages > want) stages = want;
  if (stages > 8) stages = 8;
  B200_REQUIRE(stages >= 1, "conv_tc: tile does not fit in shared memory (Cin=%d Cout=%d taps=%d)", Cin, Cout, taps);
  a.stages = stages;
  const int smem = fixed + stages * per_stage + 1024 + 1024 + 256;
  B200_REQUIRE(smem <= 227 * 1024, "conv_tc: smem %d too large", smem);
  a.total_tiles = (long long)a.tiles_w * a.tiles_h * a.B * a.n_tiles;

  CUtensorMap tmA, tmB, tmC;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * a.W, (uint64_t)Cin * 2 * a.W * a.H};
    uint32_t box[4] = {64, (uint32_t)(a.halo ? 130 : a.BW), (uint32_t)a.BH, 1};
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
  }
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  // persistent grid: as many CTAs as fit (smem and TMEM: 512 columns per SM)
  int per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm > 512 / a.tmem_cols) per_sm = 512 / a.tmem_cols;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sm_count() * per_sm;
  if ((flags >> 8) & 0xff) grid = (long long)((flags >> 8) & 0xff) * 4;
  if (grid > a.total_tiles) grid = a.total_tiles;
  conv_tc_kernel<<<(unsigned)grid, TC_THREADS, smem, (cudaStream_t)s>>>(tmA, tmB, tmC, a);
  return check_launch("conv_tc");
}
