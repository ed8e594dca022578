// Frame pre-processing of the inference script on the GPU (inference.py:28-46):
//   cv2.resize(frame, (W, H)) -> BGR2RGB -> ToTensor (/255) -> Normalize(mean, std) -> [B,3,H,W]
// for a batch of uint8 HWC BGR frames.  The resize is OpenCV's fixed-point INTER_LINEAR restated bit for bit
// (oracle/preprocess_oracle.py explains the algorithm and is checked against cv2 itself): 11-bit coefficients from
// fx = (float)((dx + 0.5) * scale - 0.5), horizontally a clamped tap zeroes its fraction, vertically the two source rows
// are clamped individually; the vertical pass is ((b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2) >> 2.
// One thread = one output pixel (3 channels): it derives its own coefficients (no tables, no host work), so the entry
// point neither allocates nor synchronises.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ void lin_coeff(int d, double scale, int ssize, bool zero_clamped, int* s_out, int* a0, int* a1) {
  // separate multiply and subtract (no fma contraction): this is how the host library evaluates it
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (zero_clamped) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  }
  *s_out = s;
  *a0 = __float2int_rn(__fsub_rn(1.f, f) * 2048.f);      // cvRound: round half to even
  *a1 = __float2int_rn(f * 2048.f);
}

template <typename TO>
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ frames, TO* __restrict__ out, uint8_t* __restrict__ rgb, int B, int Hs,
                     int Ws, int H, int W, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  if (idx >= total) return;
  const int dx = (int)(idx % W);
  const long long t = idx / W;
  const int dy = (int)(t % H);
  const int b = (int)(t / H);
  int sx, ax0, ax1, sy, by0, by1;
  lin_coeff(dx, (double)Ws / (double)W, Ws, true, &sx, &ax0, &ax1);
  lin_coeff(dy, (double)Hs / (double)H, Hs, false, &sy, &by0, &by1);
  const int x1 = min(sx + 1, Ws - 1);
  const int y0 = min(max(sy, 0), Hs - 1), y1 = min(max(sy + 1, 0), Hs - 1);
  const uint8_t* base = frames + (long long)b * Hs * Ws * 3;
  const uint8_t* r0 = base + (long long)y0 * Ws * 3;
  const uint8_t* r1 = base + (long long)y1 * Ws * 3;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  const long long plane = (long long)H * W;
#pragma unroll
  for (int k = 0; k < 3; ++k) {           // k = RGB channel = BGR channel 2 - k
    const int c = 2 - k;
    const int S0 = (int)r0[sx * 3 + c] * ax0 + (int)r0[x1 * 3 + c] * ax1;
    const int S1 = (int)r1[sx * 3 + c] * ax0 + (int)r1[x1 * 3 + c] * ax1;
    int v = (((by0 * (S0 >> 4)) >> 16) + ((by1 * (S1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    if (rgb) rgb[idx * 3 + k] = (uint8_t)v;
    const float f = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), mean[k]), stdv[k]);   // ToTensor, Normalize
    out[((long long)b * 3 + k) * plane + (long long)dy * W + dx] = from_f32<TO>(f);
  }
}

// Source already at the network size (the bench / a camera configured for it): the fixed-point bilinear reduces to the
// identity (coefficients 2048/0), so only the channel swap, /255 and the normalisation remain.  One thread = 4 pixels:
// three 4-byte loads, one 8/16-byte store per colour plane.
template <typename TO>
__global__ void __launch_bounds__(256)
preprocess_u8_same_kernel(const uint8_t* __restrict__ frames, TO* __restrict__ out, uint8_t* __restrict__ rgb, long long n4,
                          long long plane, float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 pixels
  if (i >= n4) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(frames) + i * 3;
  const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
  uint8_t px[12];
#pragma unroll
  for (int j = 0; j < 4; ++j) { px[j] = (w0 >> (8 * j)) & 255; px[4 + j] = (w1 >> (8 * j)) & 255; px[8 + j] = (w2 >> (8 * j)) & 255; }
  const long long p0 = i * 4;                       // first pixel (flat over B*H*W)
  const long long b = p0 / plane, r = p0 - b * plane;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      f[j] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[j * 3 + 2 - k], 255.f), mean[k]), stdv[k]);
    TO* dst = out + (b * 3 + k) * plane + r;
    if (sizeof(TO) == 2) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
  if (rgb) {
    uint8_t o[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[j * 3] = px[j * 3 + 2]; o[j * 3 + 1] = px[j * 3 + 1]; o[j * 3 + 2] = px[j * 3]; }
    uint32_t* d = reinterpret_cast<uint32_t*>(rgb) + i * 3;
#pragma unroll
    for (int q = 0; q < 3; ++q) d[q] = o[4 * q] | (o[4 * q + 1] << 8) | (o[4 * q + 2] << 16) | ((uint32_t)o[4 * q + 3] << 24);
  }
}


// dataset label remap on the GPU (BDD100KDataset.py:23-35,66-69 and the other datasets' class_map loops):
// out[i] = lut[in[i]] as int64 (the dtype nn.CrossEntropyLoss wants, `.long()` at BDD100KDataset.py:75), 16 labels per thread.
__global__ void __launch_bounds__(256)
remap_labels_kernel(const uint8_t* __restrict__ in, int64_t* __restrict__ out, const uint8_t* __restrict__ lut, long long n) {
  __shared__ uint8_t sl[256];
  sl[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 16;
  if (i0 + 16 <= n && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i0));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      longlong2 a = make_longlong2(sl[w[q] & 255u], sl[(w[q] >> 8) & 255u]);
      longlong2 b = make_longlong2(sl[(w[q] >> 16) & 255u], sl[w[q] >> 24]);
      *reinterpret_cast<longlong2*>(out + i0 + 4 * q) = a;
      *reinterpret_cast<longlong2*>(out + i0 + 4 * q + 2) = b;
    }
  } else {
    for (long long i = i0; i < n && i < i0 + 16; ++i) out[i] = sl[in[i]];
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200seg_preprocess_u8(const uint8_t* frames, int B, int Hs, int Ws, void* out, int out_dtype,
                                     uint8_t* rgb, int H, int W, float mean0, float mean1, float mean2, float std0,
                                     float std1, float std2, b200seg_stream_t s) {
  B200_REQUIRE(frames && out, "preprocess_u8: null pointer");
  B200_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "preprocess_u8: empty tensor");
  B200_REQUIRE(std0 != 0.f && std1 != 0.f && std2 != 0.f, "preprocess_u8: zero std");
  const long long total = (long long)B * H * W;
  B200_REQUIRE((total + 255) / 256 < (1ll << 31), "preprocess_u8: too many pixels");
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)s;
  if (out_dtype != B200SEG_F32 && out_dtype != B200SEG_BF16) return set_error(-1, "preprocess_u8: bad out dtype %d", out_dtype);
  const long long plane = (long long)H * W;
  if (Hs == H && Ws == W && plane % 4 == 0 && ((uintptr_t)frames & 3) == 0 && ((uintptr_t)out & 15) == 0 &&
      (!rgb || ((uintptr_t)rgb & 3) == 0)) {
    const long long n4 = total / 4;
    const unsigned g4 = (unsigned)((n4 + 255) / 256);
    if (out_dtype == B200SEG_F32)
      preprocess_u8_same_kernel<float><<<g4, 256, 0, st>>>(frames, (float*)out, rgb, n4, plane, mean0, mean1, mean2, std0, std1, std2);
    else
      preprocess_u8_same_kernel<__nv_bfloat16><<<g4, 256, 0, st>>>(frames, (__nv_bfloat16*)out, rgb, n4, plane, mean0, mean1, mean2, std0, std1, std2);
    return check_launch("preprocess_u8");
  }
  if (out_dtype == B200SEG_F32)
    preprocess_u8_kernel<float><<<grid, 256, 0, st>>>(frames, (float*)out, rgb, B, Hs, Ws, H, W, mean0, mean1, mean2, std0, std1, std2);
  else if (out_dtype == B200SEG_BF16)
    preprocess_u8_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(frames, (__nv_bfloat16*)out, rgb, B, Hs, Ws, H, W, mean0, mean1, mean2, std0, std1, std2);
  else
    return set_error(-1, "preprocess_u8: bad out dtype %d", out_dtype);
  return check_launch("preprocess_u8");
}

// out[i] = lut[in[i]] (uint8 labels -> int64 targets through a 256-entry table).
extern "C" int b200seg_remap_labels(const uint8_t* in, int64_t* out, const uint8_t* lut256, long long n, b200seg_stream_t s) {
  B200_REQUIRE(in && out && lut256 && n > 0, "remap_labels: bad arguments");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "remap_labels: output must be 16-byte aligned");
  const long long blocks = (n + 4095) / 4096;
  B200_REQUIRE(blocks < (1ll << 31), "remap_labels: too many labels");
  remap_labels_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(in, out, lut256, n);
  return check_launch("remap_labels");
}
