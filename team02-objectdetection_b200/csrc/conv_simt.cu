// conv_simt.cu -- dense 1x1 / 3x3 convolution as an implicit GEMM on the FP32 SIMT pipes.
//
// This is the fp32-exact path (BASELINE config 1 asks for 1e-4 relative logits; tensor-core
// bf16/tf32 products cannot give that, SURVEY finding 10c).  NHWC activations, weights
// f32 [Cout][taps][Cin], fp32 accumulation, fused bias + activation + residual epilogue.
// Tiling: 64 pixels x 64 output channels per CTA, K chunks of 16, 4x4 outputs per thread.
#include "common.cuh"

namespace b200 {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 const T* __restrict__ res, T* __restrict__ y, int B, int H, int W, int Cin, int Cout, int taps,
                 int act) {
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ float Bs[SM_BK][SM_BN + 4];
  const long long M = (long long)B * H * W;
  const int K = taps * Cin;
  const long long m0 = (long long)blockIdx.x * SM_BM;
  const int n0 = blockIdx.y * SM_BN;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;

  // loader mapping: 4 consecutive k for one row
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const long long am = m0 + lrow;
  int ab = 0, ah = 0, aw = 0;
  const bool arow_ok = am < M;
  if (arow_ok) {
    aw = (int)(am % W);
    long long t = am / W;
    ah = (int)(t % H);
    ab = (int)(t / H);
  }
  const int bn = n0 + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SM_BK) {
    const int k = k0 + lk;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < K) {
      const int tap = k / Cin, cin = k - tap * Cin;
      if (arow_ok) {
        int hi = ah, wi = aw;
        if (taps == 9) { hi += tap / 3 - 1; wi += tap % 3 - 1; }
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
          const T* xp = x + (((long long)ab * H + hi) * W + wi) * Cin + cin;
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = to_f32<T>(xp[i]);
        }
      }
      if (bn < Cout) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(w + (long long)bn * K + k));
        bv[0] = t4.x; bv[1] = t4.y; bv[2] = t4.z; bv[3] = t4.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lrow] = av[i];
      Bs[lk + i][lrow] = bv[i];
    }
    __syncthreads();
    // blocked summation: the 16 products of this K chunk are summed first and only then added to the running
    // total, so the rounding error grows with sqrt(K/16) instead of sqrt(K) (K reaches 12096 in up1.conv.0)
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < SM_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Cout) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
      v = apply_act_rt(v, act);
      if (res) v += to_f32<T>(res[m * Cout + n]);
      y[m * Cout + n] = from_f32<T>(v);
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200seg_conv_simt(const void* x, const float* w, const float* b, const void* res, void* y,
                                 int dtype, int B, int H, int W, int Cin, int Cout, int taps, int act,
                                 b200seg_stream_t s) {
  B200_REQUIRE(taps == 1 || taps == 9, "conv_simt: taps=%d", taps);
  B200_REQUIRE(Cin % 4 == 0 && Cin > 0 && Cout > 0, "conv_simt: Cin=%d must be a multiple of 4", Cin);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "conv_simt: empty tensor");
  B200_REQUIRE(dtype == B200SEG_F32 || dtype == B200SEG_BF16, "conv_simt: bad dtype");
  const long long M = (long long)B * H * W;
  dim3 grid((unsigned)((M + SM_BM - 1) / SM_BM), (unsigned)((Cout + SM_BN - 1) / SM_BN));
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == B200SEG_F32)
    conv_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)x, w, b, (const float*)res, (float*)y, B, H, W, Cin,
                                                  Cout, taps, act);
  else
    conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, w, b, (const __nv_bfloat16*)res,
                                                          (__nv_bfloat16*)y, B, H, W, Cin, Cout, taps, act);
  return check_launch("conv_simt");
}
