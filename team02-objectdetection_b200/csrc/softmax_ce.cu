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

// Four consecutive pixels per thread (HW % 4 == 0): the class planes are read and the gradient planes written with 16-byte
// accesses, all C plane loads of the thread issued before the first use -- a quarter of the memory instructions of the
// one-pixel kernel above, which ran at 3.4 TB/s on the 370 MB of the B = 32 step.
template <int CMAX>
__global__ void __launch_bounds__(256)
softmax_ce_x4_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, float* __restrict__ loss_sum,
                     float* __restrict__ dlogits, float grad_scale, const float* __restrict__ counts, int B, int C,
                     unsigned HW4) {
  const unsigned total = (unsigned)B * HW4;             // groups of 4 pixels (launcher: < 2^31)
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (idx < total) {
    const unsigned b = idx / HW4, g = idx - b * HW4;
    const float4* lp = reinterpret_cast<const float4*>(logits) + (size_t)b * C * HW4 + g;
    float4 v[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      v[c] = (c < C) ? __ldg(lp + (size_t)c * HW4) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const longlong2 t01 = __ldg(reinterpret_cast<const longlong2*>(target) + (size_t)idx * 2);
    const longlong2 t23 = __ldg(reinterpret_cast<const longlong2*>(target) + (size_t)idx * 2 + 1);
    const long long tt[4] = {t01.x, t01.y, t23.x, t23.y};
    const float gs = (dlogits && counts) ? grad_scale / fmaxf(__ldg(counts), 1.f) : grad_scale;
    float inv[4];
    bool valid[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) m = fmaxf(m, (&v[c].x)[p]);
      float sum = 0.f, zt = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        const float z = (&v[c].x)[p] - m;
        if (c == (int)tt[p]) zt = z;
        const float e = (c < C) ? expf(z) : 0.f;
        (&v[c].x)[p] = e;
        sum += e;
      }
      valid[p] = tt[p] >= 0 && tt[p] < C;
      inv[p] = 1.f / sum;
      if (valid[p]) loss += logf(sum) - zt;
    }
    if (dlogits) {
      float4* gp = reinterpret_cast<float4*>(dlogits) + (size_t)b * C * HW4 + g;
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          float4 o;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            (&o.x)[p] = valid[p] ? ((&v[c].x)[p] * inv[p] - (c == (int)tt[p] ? 1.f : 0.f)) * gs : 0.f;
          gp[(size_t)c * HW4] = o;
        }
      }
    }
  }
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
  const bool aligned = HW % 4 == 0 && total / 4 < (1LL << 31) && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(target) & 15) == 0 && (!dlogits || (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0);
  if (C <= 16 && aligned) {
    const unsigned g4 = (unsigned)((total / 4 + 255) / 256);
    if (C <= 10)          // the path's 10 classes: no dead class slots in registers
      softmax_ce_x4_kernel<10><<<g4, 256, 0, st>>>(logits, target, loss_sum, dlogits, grad_scale, counts, B, C, (unsigned)(HW / 4));
    else
      softmax_ce_x4_kernel<16><<<g4, 256, 0, st>>>(logits, target, loss_sum, dlogits, grad_scale, counts, B, C, (unsigned)(HW / 4));
    return check_launch("softmax_ce");
  }
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
