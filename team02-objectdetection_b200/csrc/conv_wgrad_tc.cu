// conv_wgrad_tc.cu -- dense weight gradient on the 5th-gen tensor cores (bf16 in, fp32 accumulate).
//
//   dW[co][tap][ci] = sum over pixels p of  dz[p][co] * x[p shifted by tap][ci]
//
// is a GEMM whose reduction axis is the PIXEL axis:  D[M = 128 couts, N <= 256 cins] += A^T B with
// A = dz tile [128 pixels][couts], B = x tile [128 pixels][cins].  Both operands are exactly the NHWC boxes
// the forward kernel loads (same 4-D tensor maps, same 128-byte swizzle; the 3x3 tap is again a shifted box
// with TMA zero fill as the padding) -- only the UMMA view changes: the tiles are read as MN-major operands
// (M/N = channels contiguous, K = pixel rows), 16 pixels per tcgen05.mma, 8 MMAs per 128-pixel tile.
//
// One CTA owns one (cout tile, cin tile, tap) and a slice of the pixel tiles; it streams its pixel tiles
// through a TMA/mbarrier ring, accumulates in ONE TMEM tile, and finally adds the tile into dW with
// red.global.add.v4.f32 (the pixel axis is split over CTAs to fill the GPU).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#include <stdlib.h>
#include "common.cuh"

namespace b200 {

struct WgradArgs {
  float* dw;
  int taps, Cin, Cout;
  int W, H, B;             // pixel geometry walked by the tile scheduler (1x1: flat)
  int BW, BH, tiles_w, tiles_h;
  int n_boxes;             // 64-channel cin boxes per N tile (N = 64 * n_boxes <= 256)
  int n_tiles, m_tiles;
  int stages;
  int rowmode;             // 3x3 on rows >= 128 px: one CTA owns a tap ROW (dh) -- the dz tile and one 130-pixel x tile
                           // feed the three horizontal taps (B descriptor start advanced by one 128-byte pixel row)
  long long pix_tiles, tiles_per_split;
  int tmem_cols;
  int exact_n;
  int m64;                 // Cout <= 64: M = 64 MMAs (one dz box; accumulator rows on TMEM lanes 32*(r/16) + r%16)
};

constexpr int WG_THREADS = 192;
constexpr int WG_BOX = 128 * 128;     // one 128-pixel x 64-channel box
constexpr int WG_XBOX = 17 * 1024;    // row mode: one 130-pixel x 64-channel box (16640 B), padded to the swizzle atom

// MN-major operand, 128-byte swizzle: 64 channels (128 B) contiguous, next 64-channel group LBO bytes away,
// 8 pixel rows per swizzle atom, next atom SBO = 1024 bytes away (cute::UMMA::make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX,
                     const WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xbox = a.rowmode ? WG_XBOX : WG_BOX;
  const int stage_bytes = 2 * WG_BOX + a.n_boxes * xbox;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  // work decomposition: blockIdx.x = ((split * taps + tap) * n_tiles + n_tile) * m_tiles + m_tile
  int bid = blockIdx.x;
  const int m_tile = bid % a.m_tiles; bid /= a.m_tiles;
  const int n_tile = bid % a.n_tiles; bid /= a.n_tiles;
  const int tap_items = a.rowmode ? 3 : a.taps;
  const int tap = bid % tap_items;           // row mode: tap = dh + 1
  const int split = bid / tap_items;
  const long long t_begin = (long long)split * a.tiles_per_split;
  const long long t_end = min(t_begin + a.tiles_per_split, a.pix_tiles);
  const int co0 = m_tile * 128, ci0 = n_tile * a.n_boxes * 64;
  int boxes = a.n_boxes;                       // cin boxes that really exist in this N tile
  while (boxes > 1 && ci0 + (boxes - 1) * 64 >= a.Cin) --boxes;
  const int a_boxes = (!a.m64 && co0 + 64 < a.Cout) ? 2 : 1;
  // MMA N: the cin that really exist in this tile, rounded to the instruction granularity (16 at M = 128) -- not the
  // 64-channel box count: a 32-channel layer issued N = 64 MMAs whose upper half multiplied zeros
  int umma_n = boxes * 64;
  if (a.exact_n) { const int real = (a.Cin - ci0 + 15) & ~15; if (real < umma_n) umma_n = real; }
  int dh = 0, dwv = 0;
  if (a.rowmode) { dh = tap - 1; dwv = -1; }
  else if (a.taps == 9) { dh = tap / 3 - 1; dwv = tap % 3 - 1; }
  const int rows = a.BW * a.BH;

  // pixel rows >= BW*BH of every stage are never written by TMA but ARE read by the MMA (the pixel axis is the
  // reduction axis): they must be zero, not stale.  Boxes that are skipped entirely are zeroed too.
  for (int i = threadIdx.x; i < a.stages * stage_bytes / 16; i += WG_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
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
    int s = 0;
    uint32_t ph = 0;
    const uint32_t tx = (uint32_t)(a_boxes * rows * 128 + boxes * (a.rowmode ? 130 : rows) * 128);
    for (long long t = t_begin; t < t_end; ++t) {
      long long mt = t;
      const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
      const int th = (int)(mt % a.tiles_h);
      const int bb = (int)(mt / a.tiles_h);
      const int w0 = tw * a.BW, h0 = th * a.BH;
      mbar_wait(&empty[s], ph ^ 1u, 11);
      if (lane == 0) {
        uint8_t* st = smem + s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], tx);
        for (int j = 0; j < a_boxes; ++j) tma_load_4d(st + j * WG_BOX, &tmD, &full[s], co0 + j * 64, w0, h0, bb);
        for (int j = 0; j < boxes; ++j)
          tma_load_4d(st + 2 * WG_BOX + j * xbox, &tmX, &full[s], ci0 + j * 64, w0 + dwv, h0 + dh, bb);
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // instruction descriptor: bf16 x bf16 -> f32, A and B both MN-major (bits 15, 16), M = 128, N = umma_n
    // With <= 64 output channels M = 64: the MN-major A operand is read from shared memory by every MMA, and at M = 128 three
    // quarters of those reads (and of the tensor-pipe rows) were the zero half / zero box of the dz tile -- the kernel ran at
    // ~82 cycles per N = 64 MMA against a 32-cycle issue floor (profiles/r02_wgrad_kbench.txt).
    const uint32_t idesc = umma_idesc_bf16(a.m64 ? 64 : 128, umma_n) | (1u << 15) | (1u << 16);
    const uint32_t desc_hi = (uint32_t)(umma_desc_mn128(0, WG_BOX) >> 32);
    const uint32_t lbo_bits = (uint32_t)((WG_BOX >> 4) & 0x3FFF) << 16;
    const uint32_t lbo_bits_x = (uint32_t)((xbox >> 4) & 0x3FFF) << 16;
    const int n_acc = a.rowmode ? 3 : 1;
    int s = 0;
    uint32_t ph = 0;
    uint32_t first = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[s], ph, 12);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = ((smem_u32(smem + s * stage_bytes) & 0x3FFFF) >> 4) | lbo_bits;
        const uint32_t b_lo = ((smem_u32(smem + s * stage_bytes + 2 * WG_BOX) & 0x3FFFF) >> 4) | lbo_bits_x;
        for (int j = 0; j < n_acc; ++j) {      // row mode: tap dw = j - 1 reads the x tile one pixel row (128 B) further
#pragma unroll
          for (int k = 0; k < 8; ++k)          // 16 pixel rows = 2048 bytes per MMA
            umma_bf16_lohi(tmem_base + (uint32_t)(j * umma_n), a_lo + 128u * k, b_lo + 8u * j + 128u * k, desc_hi, idesc,
                           (first | (uint32_t)(k > 0)));
        }
        first = 1u;
        umma_commit(&empty[s]);
        if (t == t_end - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
  } else if (t_begin < t_end) {
    // ================= epilogue: TMEM -> red.global.add =================
    const int q = warp & 3;
    // M = 128: accumulator row r on TMEM lane r.  M = 64: rows 16q .. 16q+15 on the first 16 lanes of warp q's partition.
    const int co = a.m64 ? (lane < 16 ? co0 + q * 16 + lane : a.Cout) : co0 + q * 32 + lane;
    mbar_wait(acc_full, 0, 13);
    tc_fence_after();
    const int n_acc = a.rowmode ? 3 : 1;
    for (int j = 0; j < n_acc; ++j) {
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * umma_n);
    float* drow = a.dw + ((long long)co * a.taps + (a.rowmode ? tap * 3 + j : tap)) * a.Cin + ci0;
    for (int c0 = 0; c0 < umma_n; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)c0, v);
      tmem_ld_wait();
      if (co < a.Cout) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          if (ci0 + c0 + i < a.Cin) {            // Cin % 8 == 0: groups of 4 are all-in or all-out
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + i), "f"(__uint_as_float(v[i])),
                         "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                         : "memory");
          }
        }
      }
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

}  // namespace b200

using namespace b200;

extern "C" int b200seg_conv_wgrad_tc(const void* x, const void* dz, float* dw, int B, int H, int W, int Cin, int Cout,
                                     int taps, b200seg_stream_t s) {
  B200_REQUIRE(taps == 1 || taps == 9, "conv_wgrad_tc: taps=%d", taps);
  B200_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cout > 0 && Cout % 8 == 0, "conv_wgrad_tc: Cin=%d Cout=%d must be multiples of 8", Cin, Cout);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && x && dz && dw, "conv_wgrad_tc: bad arguments");
  WgradArgs a;
  a.dw = dw; a.taps = taps; a.Cin = Cin; a.Cout = Cout;
  if (taps == 1) {
    const long long M = (long long)B * H * W;
    B200_REQUIRE(M < (1ll << 31), "conv_wgrad_tc: too many pixels");
    a.W = (int)M; a.H = 1; a.B = 1;
  } else {
    a.W = W; a.H = H; a.B = B;
  }
  // pixel tile: BW x BH <= 128 (rows beyond BW*BH stay zero in smem)
  {
    double best = -1.0;
    int bwb = 1, bhb = 1;
    const int wmax = a.W < 128 ? a.W : 128;
    for (int bw = wmax; bw >= 1; --bw) {
      int bh = 128 / bw;
      if (bh > a.H) bh = a.H;
      if (bh < 1) bh = 1;
      const long long tiles = (long long)((a.W + bw - 1) / bw) * ((a.H + bh - 1) / bh);
      const double eff = (double)a.W * a.H / ((double)tiles * 128.0);
      if (eff > best + 1e-9) { best = eff; bwb = bw; bhb = bh; }
      if (bw <= 8 && best > 0) break;
    }
    a.BW = bwb; a.BH = bhb;
  }
  {
    static int row_ok = -1;
    if (row_ok < 0) { const char* e = getenv("B200SEG_WGRAD_ROW"); row_ok = (e && e[0] == '0') ? 0 : 1; }
    a.rowmode = (taps == 9 && W >= 128 && row_ok) ? 1 : 0;
  }
  if (a.rowmode) { a.BW = 128; a.BH = 1; }
  {
    static int exact = -1;
    if (exact < 0) { const char* e = getenv("B200SEG_WGRAD_EXACT_N"); exact = (e && e[0] == '0') ? 0 : 1; }
    a.exact_n = exact;
  }
  a.tiles_w = (a.W + a.BW - 1) / a.BW;
  a.tiles_h = (a.H + a.BH - 1) / a.BH;
  a.pix_tiles = (long long)a.tiles_w * a.tiles_h * a.B;
  const int cin_boxes = (Cin + 63) / 64;
  const int max_boxes = a.rowmode ? 2 : 4;           // row mode keeps three accumulators: 3 * N <= 512 TMEM columns
  a.n_boxes = cin_boxes < max_boxes ? cin_boxes : max_boxes;
  a.n_tiles = (cin_boxes + a.n_boxes - 1) / a.n_boxes;
  a.m_tiles = (Cout + 127) / 128;
  {
    static int m64 = -1;
    if (m64 < 0) { const char* e = getenv("B200SEG_WGRAD_M64"); m64 = (e && e[0] == '0') ? 0 : 1; }
    a.m64 = (m64 && Cout <= 64) ? 1 : 0;
  }
  a.tmem_cols = 32;
  while (a.tmem_cols < (a.rowmode ? 3 : 1) * a.n_boxes * 64) a.tmem_cols <<= 1;
  const int stage_bytes = 2 * WG_BOX + a.n_boxes * (a.rowmode ? WG_XBOX : WG_BOX);
  int stages = (227 * 1024 - 2048) / stage_bytes;
  if (stages > 4) stages = 4;
  {
    static int cap = -1;          // B200SEG_WGRAD_STAGES: cap of the load ring (a smaller footprint lets main-stream CTAs co-reside)
    if (cap < 0) { const char* e = getenv("B200SEG_WGRAD_STAGES"); cap = e ? atoi(e) : 0; }
    if (cap > 0 && stages > cap) stages = cap;
  }
  B200_REQUIRE(stages >= 1, "conv_wgrad_tc: stage does not fit");
  a.stages = stages;
  const int smem = stages * stage_bytes + 1024 + 256;
  const long long mn = (long long)a.m_tiles * a.n_tiles * (a.rowmode ? 3 : taps);
  long long splits = ((long long)sm_count() * 2 + mn - 1) / mn;
  if (splits > a.pix_tiles) splits = a.pix_tiles;
  if (splits < 1) splits = 1;
  a.tiles_per_split = (a.pix_tiles + splits - 1) / splits;
  splits = (a.pix_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  const long long grid = mn * splits;
  B200_REQUIRE(grid < (1ll << 31), "conv_wgrad_tc: grid too large");

  CUtensorMap tmD, tmX;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)Cout * 2 * a.W, (uint64_t)Cout * 2 * a.W * a.H};
    uint32_t box[4] = {64, (uint32_t)a.BW, (uint32_t)a.BH, 1};
    int rc = make_tmap_bf16(&tmD, dz, 4, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Cin * 2 * a.W, (uint64_t)Cin * 2 * a.W * a.H};
    uint32_t box[4] = {64, (uint32_t)(a.rowmode ? 130 : a.BW), (uint32_t)a.BH, 1};
    int rc = make_tmap_bf16(&tmX, x, 4, dims, str, box, 1, nullptr);
    if (rc) return rc;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[64] = {false};
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  conv_wgrad_tc_kernel<<<(unsigned)grid, WG_THREADS, smem, (cudaStream_t)s>>>(tmD, tmX, a);
  return check_launch("conv_wgrad_tc");
}
