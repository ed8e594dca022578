// softmax_ce.cu -- per-pixel softmax cross-entropy fused with its gradient (main.py:99, train.py:37-38).
//
// logits are the model's public output: NCHW f32.  One thread owns one pixel: it reads its C class
// planes (each plane access is contiguous across the warp), keeps them in registers, computes the
// max-subtracted log-sum-exp once, and -- in the same pass -- writes (softmax - onehot) * scale into
// the gradient planes.  The loss is reduced with warp shuffles, then one atomicAdd per CTA.
#include "common.cuh"

namespace b200 {

template <int CMAX>
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, float* __restrict__ loss_sum,
                  float* __restrict__ dlogits, float grad_scale, const float* __restrict__ counts, int B, int C,
                  long long HW) {
  const long long total = (long long)B * HW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (idx < total) {
    const int b = (int)(idx / HW);
    const long long pix = idx - (long long)b * HW;
    const float* lp = logits + (long long)b * C * HW + pix;
    float v[CMAX];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      v[c] = (c < C) ? __ldg(lp + (long long)c * HW) : -INFINITY;
      m = fmaxf(m, v[c]);
    }
    const long long t = target[idx];
    const bool valid = t >= 0 && t < C;     // ignore_index (-100) contributes nothing
    float sum = 0.f, zt = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      const float z = v[c] - m;
      if (c == (int)t) zt = z;
      v[c] = (c < C) ? expf(z) : 0.f;
      sum += v[c];
    }
    const float inv = 1.f / sum;
    if (valid) loss = logf(sum) - zt;
    if (dlogits) {
      // 'mean' reduction divides by the number of non-ignored targets (counts[0], from b200seg_ce_count)
      const float gs = counts ? grad_scale / fmaxf(__ldg(counts), 1.f) : grad_scale;
      float* gp = dlogits + (long long)b * C * HW + pix;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) gp[(long long)c * HW] = valid ? (v[c] * inv - (c == (int)t ? 1.f : 0.f)) * gs : 0.f;
    }
  }
  // block reduction of the loss
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = wsum[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(loss_sum, t);
  }
}

int launch_softmax_ce_generic(const float* logits, const int64_t* target, float* loss_sum, float* dlogits, float grad_scale,
                              const float* counts, int B, int C, long long HW, cudaStream_t st);   // generic_ops.cu


// x *= s[0] unless s[0] == 1 (the gradient autograd hands to loss.backward() is ones: the whole pass over the 168 MB of
// d(logits) is then a few hundred CTAs that read one word and exit).
__global__ void __launch_bounds__(256)
scale_unless_one_kernel(float* __restrict__ x, long long n4, const float* __restrict__ s) {
  const float f = __ldg(s);
  if (f == 1.f) return;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    x4[i] = v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200seg_softmax_ce(const float* logits, const int64_t* target, float* loss_sum, float* dlogits,
                                  float grad_scale, const float* counts, int B, int C, int H, int W, b200seg_stream_t s) {
  B200_REQUIRE(C >= 1, "softmax_ce: C=%d", C);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "softmax_ce: empty tensor");
  B200_REQUIRE(logits && target && loss_sum, "softmax_ce: null pointer");
  const long long HW = (long long)H * W;
  const long long total = (long long)B * HW;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)s;
  if (C <= 16)
    softmax_ce_kernel<16><<<grid, 256, 0, st>>>(logits, target, loss_sum, dlogits, grad_scale, counts, B, C, HW);
  else if (C <= 32)
    softmax_ce_kernel<32><<<grid, 256, 0, st>>>(logits, target, loss_sum, dlogits, grad_scale, counts, B, C, HW);
  else
    return launch_softmax_ce_generic(logits, target, loss_sum, dlogits, grad_scale, counts, B, C, HW, st);
  return check_launch("softmax_ce");
}

// In-place x[i] *= s[0] for a device scalar s; returns immediately on the device when s[0] == 1.  n % 4 == 0, x 16-byte aligned.
extern "C" int b200seg_scale_unless_one(float* x, long long n, const float* s, b200seg_stream_t st) {
  B200_REQUIRE(x && s && n > 0 && n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "scale_unless_one: bad arguments");
  b200::scale_unless_one_kernel<<<b200::sm_count() * 4, 256, 0, (cudaStream_t)st>>>(x, n / 4, s);
  return b200::check_launch("scale_unless_one");
}
