// stem_mb1.cu -- the first two stages of the MobileNetV2 encoder in ONE kernel (eval mode, BatchNorm folded, bf16 storage):
//
//     y = project_1x1( relu6( dw3x3( relu6( stem_3x3_s2(x) ) ) ) )
//
// = features.0 (ConvBNReLU 3->32, stride 2) followed by features.1 (InvertedResidual with expand ratio 1: depthwise 3x3 +
// ReLU6, linear 1x1 32->16; tv:models/mobilenetv2.py:42-57, reached through unet.py:15,34).  Unfused these are three
// launches that write and re-read the two largest 32-channel maps of the encoder (B x 128 x 256 x 32 bf16, twice):
// 0.57 GB of HBM traffic at the bench shape for 0.17 GB of input + output.  Here a CTA owns an 8 x 32 tile of the
// half-resolution map and keeps both intermediates in shared memory:
//   1. the input patch (3 x 21 x 72 fp32, zero outside the image) is staged with coalesced 16-byte loads of the NCHW planes;
//   2. stem: im2col fragments gathered from the patch, mma.sync.m16n8k16 (K = 27 -> 32, N = 32) over the 10 x 34 halo
//      pixels the depthwise stencil needs, + bias, ReLU6, zero outside the map (the depthwise conv pads ITS input with
//      zeros) -> bf16 tile E in shared memory;
//   3. depthwise 3x3 on the mixed-precision FMA (f32 += bf16 * bf16; thread = 8 channels x 4 adjacent pixels) + bias, ReLU6
//      -> bf16 tile D in shared memory (the A operand of the projection);
//   4. projection 32 -> 16 with ldmatrix + mma.sync, + bias -> staged -> 16-byte NHWC stores (1 KB contiguous per tile row).
// Same rounding points as the three-kernel path (each intermediate rounded to bf16 once).  K = 27/32 and N = 16/32 are far
// below what a TMA/tcgen05 pipeline needs, and the planar stride-2 input has no TMA box: warp-level MMA is the right tool.
#include "common.cuh"

namespace b200 {

namespace {

constexpr int SM_TH = 8, SM_TW = 32;                 // output tile (half-resolution pixels)
constexpr int SM_HH = SM_TH + 2, SM_HW = SM_TW + 2;  // stem outputs incl. the depthwise halo
constexpr int SM_NHALO = SM_HH * SM_HW;              // 340
constexpr int SM_PR = 2 * SM_HH + 1;                 // 21 input rows
// input patch columns: the 69 columns 2*w0-3 .. 2*w0+65, starting LEAD = one 16-byte vector left of column 2*w0 so that
// every vector load is aligned and entirely inside or outside the image (4 f32 / 8 bf16 elements per vector)
template <typename TI> struct PatchGeo {
  static constexpr int VL = 16 / (int)sizeof(TI);
  static constexpr int LEAD = VL;
  static constexpr int PC = (LEAD - 3 + 69 + VL - 1) / VL * VL;      // 72 (f32) | 80 (bf16)
  static constexpr int BYTES = (3 * (2 * (8 + 2) + 1) * PC * (int)sizeof(TI) + 127) / 128 * 128;   // one raw patch buffer
};
constexpr int SM_CS = 32, SM_CO = 16;                // stem / projection output channels
constexpr int SM_EP = 40;                            // bf16 pitch of E and D rows: 20 words -> conflict-free fragment accesses
constexpr int SM_OP = 24;                            // bf16 pitch of the staged output rows
constexpr int SM_E_BYTES = SM_NHALO * SM_EP * 2;               // 27200
constexpr int SM_D_BYTES = SM_TH * SM_TW * SM_EP * 2;          // 20480
constexpr int SM_W_BYTES = SM_CS * SM_EP * 2;                   // stem weights as bf16 [32 n][32 k] (k = c*9+kh*3+kw), pitch 40
template <typename TI> constexpr int sm_smem() { return 2 * PatchGeo<TI>::BYTES + SM_E_BYTES + SM_D_BYTES + SM_W_BYTES; }

// 16-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros (pixels outside the image)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float patch_f32(const float* p, int i) { return p[i]; }
__device__ __forceinline__ float patch_f32(const __nv_bfloat16* p, int i) { return __bfloat162float(p[i]); }

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t relu6_pack2(float lo, float hi) {
  uint32_t r;
  asm("{\n\t.reg .b32 t;\n\tcvt.rn.relu.bf16x2.f32 t, %2, %1;\n\tmin.bf16x2 %0, t, %3;\n\t}"
      : "=r"(r) : "f"(lo), "f"(hi), "r"(0x40C040C0u));
  return r;
}

template <typename TI>
__global__ void __launch_bounds__(256, 3)
stem_mb1_kernel(const TI* __restrict__ x, const float* __restrict__ w_stem, const float* __restrict__ b_stem,
                const __nv_bfloat16* __restrict__ w_dw, const float* __restrict__ b_dw,
                const __nv_bfloat16* __restrict__ w_pw, const float* __restrict__ b_pw, __nv_bfloat16* __restrict__ y,
                int B, int H, int W, int Ho, int Wo, int tiles_h, int tiles_w) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int SM_PC = PatchGeo<TI>::PC, LEAD = PatchGeo<TI>::LEAD;
  // two raw input patches [3][21][PC] in the input type: the patch of the NEXT tile streams in (cp.async) while this tile computes
  uint8_t* sE = smem + 2 * PatchGeo<TI>::BYTES;                            // [340][40] bf16
  uint8_t* sD = sE + SM_E_BYTES;                                           // [256][40] bf16
  uint8_t* sO = sE;                                                        // [256][24] bf16 staged output (E is dead by then)
  uint8_t* sW = sD + SM_D_BYTES;                                           // stem weights, bf16 [32][40]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int VL = PatchGeo<TI>::VL;

  // ---- once per CTA: stem weights f32 [kh][kw][c][32] -> bf16 [n][k], k = c*9 + kh*3 + kw (27 -> 32, zero padded) ----
  for (int i = tid; i < SM_CS * 32; i += 256) {
    const int n = i >> 5, k = i & 31;
    const int c = k / 9, r9 = k - c * 9;
    reinterpret_cast<__nv_bfloat16*>(sW)[n * SM_EP + k] = __float2bfloat16_rn(k < 27 ? __ldg(w_stem + (r9 * 3 + c) * SM_CS + n) : 0.f);
  }
  // gather offsets of this thread's 8 im2col elements into the patch (same for every pixel)
  int koff[2][2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = ks * 16 + h * 8 + 2 * t + e;
        const int c = k / 9, r9 = k - c * 9;
        koff[ks][h][e] = (k < 27) ? (c * SM_PR + r9 / 3) * SM_PC + r9 % 3 + LEAD - 3 : -1;
      }
  const int g4 = tid & 3, pt = tid >> 2;
  const int orow = pt >> 3, ocol0 = (pt & 7) * 4;
  const long long HW = (long long)H * W;
  const int total = B * tiles_h * tiles_w;

  auto issue_patch = [&](const int tile, const int buf) {
    int q = tile;
    const int w0 = (q % tiles_w) * SM_TW; q /= tiles_w;
    const int h0 = (q % tiles_h) * SM_TH;
    const int b = q / tiles_h;
    const TI* xb = x + (long long)b * 3 * HW;
    const int hi0 = 2 * h0 - 3, wi0 = 2 * w0 - LEAD;        // input coordinates of patch[.][0][0]
    TI* dst = reinterpret_cast<TI*>(smem + buf * PatchGeo<TI>::BYTES);
    for (int i = tid; i < 3 * SM_PR * (SM_PC / VL); i += 256) {
      const int rowi = i / (SM_PC / VL), v = i - rowi * (SM_PC / VL);
      const int c = rowi / SM_PR, pr = rowi - c * SM_PR;
      const int hi = hi0 + pr, wi = wi0 + v * VL;
      const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;   // W % VL == 0 and wi % VL == 0: never straddles the border
      cp_async16(dst + rowi * SM_PC + v * VL, ok ? xb + (long long)c * HW + (long long)hi * W + wi : x, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  int buf = 0;
  if ((int)blockIdx.x < total) issue_patch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
    int q = tile;
    const int w0 = (q % tiles_w) * SM_TW; q /= tiles_w;
    const int h0 = (q % tiles_h) * SM_TH;
    const int b = q / tiles_h;
    const TI* patch = reinterpret_cast<const TI*>(smem + buf * PatchGeo<TI>::BYTES);
    // ---- 1. input patch: prefetch the next tile's, wait for this tile's ----
    __syncthreads();                                        // the other buffer's readers (previous tile's stem phase) and the
                                                            // previous tile's staged output (aliases E) are done
    if (tile + (int)gridDim.x < total) {
      issue_patch(tile + gridDim.x, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    // ---- 2. stem on the halo pixels -> E ----
    uint32_t bs[2][4][2];                                   // B fragments: W[n = j*8+g][k = ks*16 + h*8 + 2t, +1]
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          bs[ks][j][h] = reinterpret_cast<const uint32_t*>(sW + (j * 8 + g) * (SM_EP * 2))[ks * 8 + h * 4 + t];
    for (int mt = warp; mt < (SM_NHALO + 15) / 16; mt += 8) {
      int poff[2];
      bool keep[2];
      int rows[2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = mt * 16 + g + rr * 8;
        const int rc = r < SM_NHALO ? r : SM_NHALO - 1;
        const int hy = rc / SM_HW, hx = rc - hy * SM_HW;
        poff[rr] = 2 * hy * SM_PC + 2 * hx;
        const int gh = h0 - 1 + hy, gw = w0 - 1 + hx;
        keep[rr] = gh >= 0 && gh < Ho && gw >= 0 && gw < Wo;
        rows[rr] = r;
      }
      uint32_t afr[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) v[e] = koff[ks][h][e] >= 0 ? patch_f32(patch, koff[ks][h][e] + poff[rr]) : 0.f;
            afr[ks][h * 2 + rr] = pack_bf16x2(v[0], v[1]);
          }
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b0 = __ldg(b_stem + j * 8 + 2 * t), b1 = __ldg(b_stem + j * 8 + 2 * t + 1);
        acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(acc[j], afr[ks], bs[ks][j][0], bs[ks][j][1]);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
        if (rows[rr] < SM_NHALO) {
          uint32_t* er = reinterpret_cast<uint32_t*>(sE + rows[rr] * (SM_EP * 2));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            er[j * 4 + t] = keep[rr] ? relu6_pack2(acc[j][2 * rr], acc[j][2 * rr + 1]) : 0u;
        }
    }
    __syncthreads();
    // ---- 3. depthwise 3x3 + ReLU6: E -> D ----
    {
      float acc[4][8];
      {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b_dw + g4 * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(b_dw + g4 * 8 + 4));
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
          acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
        }
      }
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        uint4 wt[3];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) wt[dw] = __ldg(reinterpret_cast<const uint4*>(w_dw + (dh * 3 + dw) * SM_CS + g4 * 8));
        const uint8_t* ep = sE + ((orow + dh) * SM_HW + ocol0) * (SM_EP * 2) + g4 * 16;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint4 tv = *reinterpret_cast<const uint4*>(ep + j * (SM_EP * 2));
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int dw = j - o;
            if (dw >= 0 && dw < 3) {
              const uint4 wv = wt[dw];
              float* ac = acc[o];
              ac[0] = fma_bf16_lo(tv.x, wv.x, ac[0]); ac[1] = fma_bf16_hi(tv.x, wv.x, ac[1]);
              ac[2] = fma_bf16_lo(tv.y, wv.y, ac[2]); ac[3] = fma_bf16_hi(tv.y, wv.y, ac[3]);
              ac[4] = fma_bf16_lo(tv.z, wv.z, ac[4]); ac[5] = fma_bf16_hi(tv.z, wv.z, ac[5]);
              ac[6] = fma_bf16_lo(tv.w, wv.w, ac[6]); ac[7] = fma_bf16_hi(tv.w, wv.w, ac[7]);
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 4; ++o)
        *reinterpret_cast<uint4*>(sD + (orow * SM_TW + ocol0 + o) * (SM_EP * 2) + g4 * 16) =
            make_uint4(relu6_pack2(acc[o][0], acc[o][1]), relu6_pack2(acc[o][2], acc[o][3]),
                       relu6_pack2(acc[o][4], acc[o][5]), relu6_pack2(acc[o][6], acc[o][7]));
    }
    __syncthreads();
    // ---- 4. projection 32 -> 16 (+bias, linear) -> staged output ----
    // B fragments: W[n = j*8+g][k = ks*16 + 2t (+8)], k-major rows of 32 (re-read per tile from L1: 8 registers fewer
    // live across the stem and depthwise phases)
    uint32_t bp[2][2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          bp[ks][j][h] = __ldg(reinterpret_cast<const uint32_t*>(w_pw + (j * 8 + g) * SM_CS + ks * 16 + h * 8 + 2 * t));
#pragma unroll
    for (int m2 = 0; m2 < 2; ++m2) {
      const int px0 = (warp * 2 + m2) * 16;
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float b0 = __ldg(b_pw + j * 8 + 2 * t), b1 = __ldg(b_pw + j * 8 + 2 * t + 1);
        acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
        ldmatrix_x4(a, smem_u32(sD + (px0 + (lane & 15)) * (SM_EP * 2) + (ks * 16 + (lane >> 4) * 8) * 2));
#pragma unroll
        for (int j = 0; j < 2; ++j) mma16816(acc[j], a, bp[ks][j][0], bp[ks][j][1]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        reinterpret_cast<uint32_t*>(sO + (px0 + g) * (SM_OP * 2))[j * 4 + t] = pack_bf16x2(acc[j][0], acc[j][1]);
        reinterpret_cast<uint32_t*>(sO + (px0 + g + 8) * (SM_OP * 2))[j * 4 + t] = pack_bf16x2(acc[j][2], acc[j][3]);
      }
    }
    __syncthreads();
    // ---- 5. coalesced NHWC stores: 32 bytes per pixel, 1 KB per tile row ----
#pragma unroll
    for (int i = tid; i < SM_TH * SM_TW * 2; i += 256) {
      const int px = i >> 1, hv = i & 1;
      const int gh = h0 + (px >> 5), gw = w0 + (px & 31);
      if (gh < Ho && gw < Wo)
        *reinterpret_cast<uint4*>(y + (((long long)b * Ho + gh) * Wo + gw) * SM_CO + hv * 8) =
            *reinterpret_cast<const uint4*>(sO + px * (SM_OP * 2) + hv * 16);
    }
  }
}

}  // namespace

}  // namespace b200

using namespace b200;

// 1 if b200seg_stem_mb1 supports the configuration (the engine asks before replacing the three steps)
extern "C" int b200seg_stem_mb1_supported(int x_dtype, int H, int W, int Cstem, int Cout) {
  const int vl = x_dtype == B200SEG_BF16 ? 8 : 4;
  return (x_dtype == B200SEG_F32 || x_dtype == B200SEG_BF16) && Cstem == SM_CS && Cout == SM_CO && H >= 2 && W >= 2 &&
         H % 2 == 0 && W % 2 == 0 && W % vl == 0;
}

// x NCHW [B,3,H,W] (f32 | bf16); w_stem f32 [3][3][3][32] + b_stem f32 [32] (features.0, BN folded, ReLU6);
// w_dw bf16 [9][32] + b_dw f32 [32] (features.1.conv.0, ReLU6); w_pw bf16 [16][32] + b_pw f32 [16] (features.1.conv.1, linear);
// y NHWC bf16 [B, H/2, W/2, 16].
extern "C" int b200seg_stem_mb1(const void* x, int x_dtype, const float* w_stem, const float* b_stem, const void* w_dw,
                                const float* b_dw, const void* w_pw, const float* b_pw, void* y, int B, int H, int W,
                                b200seg_stream_t s) {
  B200_REQUIRE(x && w_stem && b_stem && w_dw && b_dw && w_pw && b_pw && y && B > 0, "stem_mb1: bad arguments");
  B200_REQUIRE(b200seg_stem_mb1_supported(x_dtype, H, W, SM_CS, SM_CO), "stem_mb1: unsupported input H=%d W=%d dtype=%d", H, W, x_dtype);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "stem_mb1: input must be 16-byte aligned");
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_h = (Ho + SM_TH - 1) / SM_TH, tiles_w = (Wo + SM_TW - 1) / SM_TW;
  const long long total = (long long)B * tiles_h * tiles_w;
  B200_REQUIRE(total < (1LL << 31), "stem_mb1: too many tiles");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(stem_mb1_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_smem<float>());
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem_mb1_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_smem<__nv_bfloat16>());
    if (e != cudaSuccess) return set_error((int)e, "stem_mb1: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  long long grid = (long long)sm_count() * 3;
  if (grid > total) grid = total;
  cudaStream_t st = (cudaStream_t)s;
  if (x_dtype == B200SEG_F32)
    stem_mb1_kernel<float><<<(unsigned)grid, 256, sm_smem<float>(), st>>>(
        (const float*)x, w_stem, b_stem, (const __nv_bfloat16*)w_dw, b_dw, (const __nv_bfloat16*)w_pw, b_pw, (__nv_bfloat16*)y, B, H,
        W, Ho, Wo, tiles_h, tiles_w);
  else
    stem_mb1_kernel<__nv_bfloat16><<<(unsigned)grid, 256, sm_smem<__nv_bfloat16>(), st>>>(
        (const __nv_bfloat16*)x, w_stem, b_stem, (const __nv_bfloat16*)w_dw, b_dw, (const __nv_bfloat16*)w_pw, b_pw,
        (__nv_bfloat16*)y, B, H, W, Ho, Wo, tiles_h, tiles_w);
  return check_launch("stem_mb1");
}
