// generic_ops.cu -- the uncommon shapes of the public surface (any class count), plain HBM kernels.
//
// The fast kernels (hbm_ops.cu, softmax_ce.cu, train_ops.cu) keep <= 16 logit channels in registers, which covers the
// reference's 10-class model.  `outconv(in_ch, out_ch)` (unet.py:108-121) takes any out_ch, so the same operators exist
// here for C > 16: one thread per output pixel, a loop over the classes, NHWC logits with pixel pitch `ldc`.
//   upsample2x_ac_generic   final_upsample (unet.py:30,49) -> NCHW logits or the uint8 argmax mask
//   final_bwd_generic       its adjoint
//   nhwc_argmax             argmax over NHWC logits at input resolution (plain UNet's predict_mask, inference.py:64)
//   ce_count                number of non-ignored targets + number of out-of-range labels (nn.CrossEntropyLoss 'mean')
#include "common.cuh"

namespace b200 {

template <typename T, typename TO>
__global__ void __launch_bounds__(256)
upsample2x_ac_generic_kernel(const T* __restrict__ lg, int ldc, TO* __restrict__ out, uint8_t* __restrict__ mask, int B,
                             int h, int w, int C) {
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)B * Ho * Wo;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo);
  long long p = idx / Wo;
  const int ho = (int)(p % Ho);
  const int b = (int)(p / Ho);
  const float sch = (Ho > 1) ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float scw = (Wo > 1) ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float sy = sch * ho, sx = scw * wo;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - y0, hy = 1.f - ly, lx = sx - x0, hx = 1.f - lx;
  const T* p00 = lg + (((long long)b * h + y0) * w + x0) * ldc;
  const T* p01 = lg + (((long long)b * h + y0) * w + x1) * ldc;
  const T* p10 = lg + (((long long)b * h + y1) * w + x0) * ldc;
  const T* p11 = lg + (((long long)b * h + y1) * w + x1) * ldc;
  int best = 0;
  float bv = -INFINITY;
  for (int c = 0; c < C; ++c) {
    // same association as the fast kernel: vertical blend of each column first, then horizontal
    const float l = hy * to_f32<T>(p00[c]) + ly * to_f32<T>(p10[c]);
    const float r = hy * to_f32<T>(p01[c]) + ly * to_f32<T>(p11[c]);
    const float v = hx * l + lx * r;
    if (mask) {
      if (v > bv) { bv = v; best = c; }      // first maximum wins, like torch.max
    } else {
      out[(((long long)b * C + c) * Ho + ho) * Wo + wo] = from_f32<TO>(v);
    }
  }
  if (mask) mask[idx] = (uint8_t)best;
}

__device__ __forceinline__ float bil_w_ac(int r, int i, int n_in, float scale) {
  const float s = scale * r;
  const int i0 = (int)s;
  const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  const float l = s - i0;
  return (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
}

// dout NCHW f32 [B,C,2h,2w] -> dlogits NHWC [B,h,w,ldc] (channels >= C zeroed); gather form, no atomics
template <typename T>
__global__ void __launch_bounds__(256)
final_bwd_generic_kernel(const float* __restrict__ dout, T* __restrict__ dl, int B, int h, int w, int C, int ldc) {
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)B * h * w;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % w);
  long long p = idx / w;
  const int i = (int)(p % h);
  const int b = (int)(p / h);
  const float sch = (Ho > 1) ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float scw = (Wo > 1) ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const int r_lo = max(2 * i - 2, 0), r_hi = min(2 * i + 3, Ho - 1);
  const int q_lo = max(2 * j - 2, 0), q_hi = min(2 * j + 3, Wo - 1);
  T* op = dl + idx * ldc;
  for (int c = 0; c < ldc; ++c) {
    float acc = 0.f;
    if (c < C) {
      for (int r = r_lo; r <= r_hi; ++r) {
        const float wy = bil_w_ac(r, i, h, sch);
        if (wy == 0.f) continue;
        for (int q = q_lo; q <= q_hi; ++q) {
          const float wx = bil_w_ac(q, j, w, scw);
          if (wx != 0.f) acc = fmaf(wy * wx, __ldg(dout + (((long long)b * C + c) * Ho + r) * Wo + q), acc);
        }
      }
    }
    op[c] = from_f32<T>(acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_argmax_kernel(const T* __restrict__ x, int ldc, uint8_t* __restrict__ mask, long long P, int C) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const T* p = x + idx * ldc;
  int best = 0;
  float bv = to_f32<T>(p[0]);
  for (int c = 1; c < C; ++c) {
    const float v = to_f32<T>(p[c]);
    if (v > bv) { bv = v; best = c; }
  }
  mask[idx] = (uint8_t)best;
}

// counts[0] += #targets in [0,C) ; counts[1] += #targets that are neither in range nor ignore_index
__global__ void __launch_bounds__(256)
ce_count_kernel(const int64_t* __restrict__ target, float* __restrict__ counts, long long N, int C, long long ignore_index) {
  float valid = 0.f, bad = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const long long t = target[i];
    if (t >= 0 && t < C) valid += 1.f;
    else if (t != ignore_index) bad += 1.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    valid += __shfl_xor_sync(0xffffffffu, valid, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  __shared__ float sv[8], sb[8];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = valid; sb[threadIdx.x >> 5] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { v += sv[i]; b += sb[i]; }
    if (v != 0.f) atomicAdd(counts, v);          // integer-valued partial sums: exact in f32 up to 2^24 per add
    if (b != 0.f) atomicAdd(counts + 1, b);
  }
}

// Any class count (C > 32): three passes over the class planes of one pixel (max, sum, gradient), nothing kept in registers.
__global__ void __launch_bounds__(256)
softmax_ce_generic_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, float* __restrict__ loss_sum,
                          float* __restrict__ dlogits, float grad_scale, const float* __restrict__ counts, int B, int C,
                          long long HW) {
  const long long total = (long long)B * HW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (idx < total) {
    const int b = (int)(idx / HW);
    const long long pix = idx - (long long)b * HW;
    const float* lp = logits + (long long)b * C * HW + pix;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(lp + (long long)c * HW));
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(__ldg(lp + (long long)c * HW) - m);
    const long long t = target[idx];
    const bool valid = t >= 0 && t < C;
    if (valid) loss = logf(sum) - (__ldg(lp + t * HW) - m);
    if (dlogits) {
      const float gs = counts ? grad_scale / fmaxf(counts[0], 1.f) : grad_scale;
      const float inv = 1.f / sum;
      float* gp = dlogits + (long long)b * C * HW + pix;
      for (int c = 0; c < C; ++c)
        gp[(long long)c * HW] = valid ? (expf(__ldg(lp + (long long)c * HW) - m) * inv - (c == (int)t ? 1.f : 0.f)) * gs : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += wsum[i];
    atomicAdd(loss_sum, t);
  }
}

int launch_softmax_ce_generic(const float* logits, const int64_t* target, float* loss_sum, float* dlogits, float grad_scale,
                              const float* counts, int B, int C, long long HW, cudaStream_t st) {
  const unsigned grid = (unsigned)(((long long)B * HW + 255) / 256);
  softmax_ce_generic_kernel<<<grid, 256, 0, st>>>(logits, target, loss_sum, dlogits, grad_scale, counts, B, C, HW);
  return check_launch("softmax_ce");
}

}  // namespace b200

using namespace b200;
typedef __nv_bfloat16 bf16;

extern "C" {

int b200seg_upsample2x_ac_generic(const void* logits, int dtype, int ldc, void* out, int out_dtype, uint8_t* mask, int B,
                                  int h, int w, int C, b200seg_stream_t s) {
  B200_REQUIRE(C >= 1 && ldc >= C && B > 0 && h > 0 && w > 0, "upsample2x_ac_generic: bad shape (C=%d ldc=%d)", C, ldc);
  B200_REQUIRE((out != nullptr) != (mask != nullptr), "upsample2x_ac_generic: exactly one of out / mask");
  B200_REQUIRE(mask == nullptr || C <= 256, "upsample2x_ac_generic: a uint8 mask holds at most 256 classes");
  const long long total = (long long)B * 4 * h * w;
  const unsigned g = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)s;
#define GL(T, TO) upsample2x_ac_generic_kernel<T, TO><<<g, 256, 0, st>>>((const T*)logits, ldc, (TO*)out, mask, B, h, w, C)
  if (dtype == B200SEG_BF16 && (mask || out_dtype == B200SEG_F32)) GL(bf16, float);
  else if (dtype == B200SEG_BF16 && out_dtype == B200SEG_BF16) GL(bf16, bf16);
  else if (dtype == B200SEG_F32 && (mask || out_dtype == B200SEG_F32)) GL(float, float);
  else if (dtype == B200SEG_F32 && out_dtype == B200SEG_BF16) GL(float, bf16);
  else return set_error(-1, "upsample2x_ac_generic: bad dtypes");
#undef GL
  return check_launch("upsample2x_ac_generic");
}

int b200seg_final_bwd_generic(const float* dout, void* dlogits, int dtype, int B, int h, int w, int C, int ldc,
                              b200seg_stream_t s) {
  B200_REQUIRE(C >= 1 && ldc >= C && B > 0 && h > 0 && w > 0, "final_bwd_generic: bad shape");
  const unsigned g = (unsigned)(((long long)B * h * w + 255) / 256);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == B200SEG_BF16) final_bwd_generic_kernel<bf16><<<g, 256, 0, st>>>(dout, (bf16*)dlogits, B, h, w, C, ldc);
  else if (dtype == B200SEG_F32) final_bwd_generic_kernel<float><<<g, 256, 0, st>>>(dout, (float*)dlogits, B, h, w, C, ldc);
  else return set_error(-1, "final_bwd_generic: bad dtype");
  return check_launch("final_bwd_generic");
}

int b200seg_nhwc_argmax(const void* x, int dtype, int ldc, uint8_t* mask, long long P, int C, b200seg_stream_t s) {
  B200_REQUIRE(C >= 1 && C <= 256 && ldc >= C && P > 0 && x && mask, "nhwc_argmax: bad arguments");
  const unsigned g = (unsigned)((P + 255) / 256);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == B200SEG_BF16) nhwc_argmax_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, ldc, mask, P, C);
  else if (dtype == B200SEG_F32) nhwc_argmax_kernel<float><<<g, 256, 0, st>>>((const float*)x, ldc, mask, P, C);
  else return set_error(-1, "nhwc_argmax: bad dtype");
  return check_launch("nhwc_argmax");
}

int b200seg_ce_count(const int64_t* target, float* counts, long long N, int C, long long ignore_index, b200seg_stream_t s) {
  B200_REQUIRE(target && counts && N > 0 && C >= 1, "ce_count: bad arguments");
  long long g = (N + 256 * 8 - 1) / (256 * 8);
  if (g > 148 * 8) g = 148 * 8;
  ce_count_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)s>>>(target, counts, N, C, ignore_index);
  return check_launch("ce_count");
}

}  // extern "C"
