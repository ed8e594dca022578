// Fused multi-tensor Adam step: torch.optim.Adam(model.parameters(), lr=1.5e-4) as the
// reference constructs it (main.py:100) and steps it (train.py:39), for ALL parameter tensors in one launch.
// The eager optimizer walks the 194 tensors with ~10 _foreach_ kernels (lerp_, mul_, addcmul_, sqrt, div, add, addcdiv_);
// here every element is read once (p, g, m, v) and written once (p, m, v): 185 MB per step for this model.
//
// Arithmetic follows torch/optim/adam.py (_single_tensor_adam, amsgrad=False, maximize=False, capturable=False):
//   g  = grad (+ weight_decay * p)
//   m  = m + (g - m) * (1 - beta1)                     (exp_avg.lerp_(g, 1 - beta1))
//   v  = v * beta2 + (1 - beta2) * g * g               (exp_avg_sq.mul_(beta2).addcmul_(g, g, value = 1 - beta2))
//   p  = p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)     bc1 = 1 - beta1^t, bc2 = 1 - beta2^t (computed by the host)
#include "common.cuh"

namespace b200 {

struct AdamTensor {      // one entry per parameter tensor (device table, 40 bytes)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};

constexpr int ADAM_CHUNK = 4096;      // elements per block

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamTensor* __restrict__ table, const int* __restrict__ chunk_tensor,
                  const int* __restrict__ chunk_index, float one_minus_b1, float beta2, float one_minus_b2,
                  float step_size, float bc2_sqrt, float eps, float weight_decay) {
  const AdamTensor t = table[chunk_tensor[blockIdx.x]];
  const long long base = (long long)chunk_index[blockIdx.x] * ADAM_CHUNK;
  const long long end = min(base + ADAM_CHUNK, t.n);
  const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
  if (vec) {
    for (long long i = base + threadIdx.x * 4; i + 3 < end; i += 256 * 4) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      const float4 g4 = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = *reinterpret_cast<const float4*>(t.m + i);
      float4 v = *reinterpret_cast<const float4*>(t.v + i);
      float* pp = &p.x; float* mm = &m.x; float* vv = &v.x; const float* gg = &g4.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = gg[j] + weight_decay * pp[j];
        mm[j] = mm[j] + (g - mm[j]) * one_minus_b1;
        vv[j] = vv[j] * beta2 + one_minus_b2 * g * g;
        pp[j] = pp[j] - step_size * (mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps));
      }
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
    }
  }
  // scalar path: unaligned tensors, and the < 4-element tail of an aligned chunk
  long long i0 = vec ? base + ((end - base) & ~3LL) : base;
  for (long long i = i0 + threadIdx.x; i < end; i += 256) {
    const float g = t.g[i] + weight_decay * t.p[i];
    const float m = t.m[i] + (g - t.m[i]) * one_minus_b1;
    const float v = t.v[i] * beta2 + one_minus_b2 * g * g;
    t.p[i] = t.p[i] - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
    t.m[i] = m;
    t.v[i] = v;
  }
}

}  // namespace b200

using namespace b200;

// one_minus_beta1/2 are passed as the host computes them in double (torch rounds 1 - beta to float once; 1.f - (float)beta
// differs from it by 1e-5 relative for beta2 = 0.999)
extern "C" int b200seg_adam_multi(const void* table, const int* chunk_tensor, const int* chunk_index, int n_chunks,
                                  float lr, double one_minus_beta1, float beta2, double one_minus_beta2, float eps,
                                  float weight_decay, double bias_corr1, double bias_corr2, b200seg_stream_t s) {
  B200_REQUIRE(table && chunk_tensor && chunk_index && n_chunks > 0, "adam_multi: bad arguments");
  B200_REQUIRE(bias_corr1 > 0.0 && bias_corr2 > 0.0, "adam_multi: bias corrections must be positive (step >= 1)");
  const float step_size = (float)((double)lr / bias_corr1);
  const float bc2_sqrt = (float)sqrt(bias_corr2);
  adam_multi_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)s>>>(
      (const AdamTensor*)table, chunk_tensor, chunk_index, (float)one_minus_beta1, beta2, (float)one_minus_beta2, step_size,
      bc2_sqrt, eps, weight_decay);
  return check_launch("adam_multi");
}

extern "C" int b200seg_adam_chunk(void) { return ADAM_CHUNK; }

// ---------------------------------------------------------------------------------------------------------------
// Weight repack for the training step (second half): every conv weight (fp32 OIHW master copy that
// Adam updates) -> the operand layouts the kernels read, for ALL layers in one launch:
//   kind 0  dense, tensor-core path: bf16 [Cout_pad][taps][Cin] (forward / wgrad view) and bf16 [Cin][taps flipped][Cout_pad]
//           (the transposed, tap-flipped weights of the data gradient)
//   kind 1  dense, fp32 path: the same two layouts in f32
//   kind 2  stem: f32 [kh][kw][Cin][Cout]          kind 3  depthwise: f32 [9][C] (+ the tap-flipped [9][C] for the data gradient)
//   kind 4  depthwise: bf16 [9][C] + tap-flipped bf16 [9][C]
// One thread per element of the padded OIHW tensor (coalesced reads; rows >= Cout are written as zeros).
// ---------------------------------------------------------------------------------------------------------------
namespace b200 {

struct PackEntry {       // 48 bytes
  const float* w;
  void* fwd;
  void* dgrad;
  int cout, cin, kk, cout_pad, kind, pad_;
};

constexpr int PACK_CHUNK = 2048;

__global__ void __launch_bounds__(256)
pack_weights_kernel(const PackEntry* __restrict__ table, const int* __restrict__ chunk_tensor,
                    const int* __restrict__ chunk_index) {
  const PackEntry t = table[chunk_tensor[blockIdx.x]];
  const long long per_o = (long long)t.cin * t.kk;
  const long long n = (long long)t.cout_pad * per_o;
  const long long base = (long long)chunk_index[blockIdx.x] * PACK_CHUNK;
  for (long long e = base + threadIdx.x; e < min(base + PACK_CHUNK, n); e += 256) {
    const int o = (int)(e / per_o);
    const int r = (int)(e - (long long)o * per_o);
    const int i = r / t.kk, tap = r - i * t.kk;
    const float v = o < t.cout ? __ldg(t.w + e) : 0.f;          // same linear index: the padding rows come last
    if (t.kind == 0) {
      reinterpret_cast<__nv_bfloat16*>(t.fwd)[((long long)o * t.kk + tap) * t.cin + i] = __float2bfloat16_rn(v);
      reinterpret_cast<__nv_bfloat16*>(t.dgrad)[((long long)i * t.kk + (t.kk - 1 - tap)) * t.cout_pad + o] = __float2bfloat16_rn(v);
    } else if (t.kind == 1) {
      reinterpret_cast<float*>(t.fwd)[((long long)o * t.kk + tap) * t.cin + i] = v;
      reinterpret_cast<float*>(t.dgrad)[((long long)i * t.kk + (t.kk - 1 - tap)) * t.cout_pad + o] = v;
    } else if (t.kind == 2) {
      reinterpret_cast<float*>(t.fwd)[((long long)tap * t.cin + i) * t.cout + o] = v;
    } else if (t.kind == 3) {
      reinterpret_cast<float*>(t.fwd)[(long long)tap * t.cout + o] = v;
      // tap-flipped twin: the stride-1 data gradient of a depthwise conv is the same depthwise conv with flipped taps
      if (t.dgrad) reinterpret_cast<float*>(t.dgrad)[(long long)(t.kk - 1 - tap) * t.cout + o] = v;
    } else {   // kind 4: the same two depthwise layouts in bf16 (taps of the mixed-precision-FMA kernel)
      reinterpret_cast<__nv_bfloat16*>(t.fwd)[(long long)tap * t.cout + o] = __float2bfloat16_rn(v);
      if (t.dgrad) reinterpret_cast<__nv_bfloat16*>(t.dgrad)[(long long)(t.kk - 1 - tap) * t.cout + o] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace b200

extern "C" int b200seg_pack_weights_multi(const void* table, const int* chunk_tensor, const int* chunk_index,
                                          int n_chunks, b200seg_stream_t s) {
  B200_REQUIRE(table && chunk_tensor && chunk_index && n_chunks > 0, "pack_weights_multi: bad arguments");
  b200::pack_weights_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)s>>>((const b200::PackEntry*)table, chunk_tensor,
                                                                              chunk_index);
  return b200::check_launch("pack_weights_multi");
}

extern "C" int b200seg_pack_chunk(void) { return b200::PACK_CHUNK; }

// ---------------------------------------------------------------------------------------------------------------
// Gradient finalize (train.py:38 -> p.grad): the backward kernels leave every parameter gradient as a raw partial result
// in a staging buffer -- f64 per-channel sums in `nslot` copies (BatchNorm gamma/beta, conv biases, depthwise taps) or
// f32 accumulators in the kernels' operand layout (dense OHWI, stem [kk][cin][cout]).  This kernel turns ALL tensors of
// one gradient bucket into the parameters' own layout (OIHW / [C]) inside the flat gradient arena, in ONE launch,
// pre-scaled (1/world for the data-parallel average), replacing ~90 f64->f32 launches, ~60 permute copies and ~400
// clone/scale copies per step.
//   kind 0  f64 slots -> f32 [n]                       dst[i] = scale * sum_slot src[slot*stride + i]
//   kind 1  dense f32 [cout_pad][kk][cin] -> OIHW      dst[(o*cin + i)*kk + t] = scale * src[(o*kk + t)*cin + i]
//   kind 2  stem f32 [kk][cin][cout] -> OIHW           dst[(o*cin + i)*kk + t] = scale * src[(t*cin + i)*cout + o]
//   kind 3  depthwise f64 slots [kk][C] -> [C][kk]     dst[c*kk + t] = scale * sum_slot src[slot*stride + t*C + c]
// ---------------------------------------------------------------------------------------------------------------
namespace b200 {

struct GradEntry {       // 64 bytes
  float* dst;
  const void* src;
  long long n;
  long long slot_stride;
  int kind, cout, cin, kk, nslot, pad_;
  float scale;
  int pad2_;
};

constexpr int GRAD_CHUNK = 2048;

__global__ void __launch_bounds__(256)
grad_finalize_kernel(const GradEntry* __restrict__ table, const int* __restrict__ chunk_tensor,
                     const int* __restrict__ chunk_index) {
  const GradEntry t = table[chunk_tensor[blockIdx.x]];
  const long long base = (long long)chunk_index[blockIdx.x] * GRAD_CHUNK;
  const long long end = min(base + GRAD_CHUNK, t.n);
  for (long long e = base + threadIdx.x; e < end; e += 256) {
    float v;
    if (t.kind == 0) {
      const double* s = reinterpret_cast<const double*>(t.src);
      double a = 0.0;
      for (int sl = 0; sl < t.nslot; ++sl) a += s[sl * t.slot_stride + e];
      v = (float)a;
    } else if (t.kind == 3) {
      const double* s = reinterpret_cast<const double*>(t.src);
      const int c = (int)(e / t.kk), tap = (int)(e - (long long)c * t.kk);
      double a = 0.0;
      for (int sl = 0; sl < t.nslot; ++sl) a += s[sl * t.slot_stride + (long long)tap * t.cout + c];
      v = (float)a;
    } else {
      const float* s = reinterpret_cast<const float*>(t.src);
      const long long per_o = (long long)t.cin * t.kk;
      const int o = (int)(e / per_o);
      const int r = (int)(e - (long long)o * per_o);
      const int i = r / t.kk, tap = r - i * t.kk;
      v = t.kind == 1 ? s[((long long)o * t.kk + tap) * t.cin + i] : s[((long long)tap * t.cin + i) * t.cout + o];
    }
    t.dst[e] = v * t.scale;
  }
}

}  // namespace b200

extern "C" int b200seg_grad_finalize_multi(const void* table, const int* chunk_tensor, const int* chunk_index,
                                           int n_chunks, b200seg_stream_t s) {
  B200_REQUIRE(table && chunk_tensor && chunk_index && n_chunks > 0, "grad_finalize_multi: bad arguments");
  b200::grad_finalize_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)s>>>((const b200::GradEntry*)table, chunk_tensor,
                                                                               chunk_index);
  return b200::check_launch("grad_finalize_multi");
}

extern "C" int b200seg_grad_chunk(void) { return b200::GRAD_CHUNK; }

// ---------------------------------------------------------------------------------------------------------------
// Eval-mode weight preparation (model.eval(), inference.py:25): fold every BatchNorm into its convolution in fp32
// (w' = w * g / sqrt(rv + eps); b' = beta + (b - rm) * g / sqrt(rv + eps)) and write the operand
// layout the forward kernels read -- ALL layers in one launch (before: ~20 ATen kernels per layer, ~1300 per model).
// One table row per OUTPUT operand (a layer may have several: depthwise taps are needed as [9][C] f32, zero padded to
// 64-channel chunks for the fused block, and block-diagonal bf16 for the tensor-core depthwise):
//   kind 0  dense bf16 [cout_pad][kk][cin]        kind 1  dense f32 [cout_pad][kk][cin]
//   kind 2  stem f32 [kk][cin][cout]              kind 3  depthwise f32 [kk][ldw]  (ldw >= C, padding stays zero)
//   kind 4  depthwise block-diagonal bf16 [C][kk][64]: out[c][t][c % 64] = w'[c][t]      (b200seg_dwconv3x3_tc)
//   kind 5  depthwise bf16 [kk][ldw]
// Every row may also write the folded shift to out_b (f32, caller-zeroed beyond cout).  Buffers are zero-initialised by the
// caller; only real elements are written.
// ---------------------------------------------------------------------------------------------------------------
namespace b200 {

struct FoldEntry {       // 96 bytes
  const float* w;        // OIHW master weights
  const float* cbias;    // conv bias or NULL
  const float* gamma;    // BatchNorm weight or NULL (no BN: plain copy)
  const float* beta;
  const float* rmean;
  const float* rvar;
  void* out_w;
  float* out_b;          // or NULL
  int cout, cin, kk, ldw, kind;
  float eps;
  long long pad_;
};

constexpr int FOLD_CHUNK = 2048;

__global__ void __launch_bounds__(256)
fold_pack_eval_kernel(const FoldEntry* __restrict__ table, const int* __restrict__ chunk_tensor,
                      const int* __restrict__ chunk_index) {
  const FoldEntry t = table[chunk_tensor[blockIdx.x]];
  const long long per_o = (long long)t.cin * t.kk;
  const long long n = (long long)t.cout * per_o;
  const long long base = (long long)chunk_index[blockIdx.x] * FOLD_CHUNK;
  const long long end = min(base + FOLD_CHUNK, n + (t.out_b ? t.cout : 0));
  for (long long e = base + threadIdx.x; e < end; e += 256) {
    const int o = e < n ? (int)(e / per_o) : (int)(e - n);
    float scale = 1.f;
    if (t.gamma) scale = __ldg(t.gamma + o) * rsqrtf(__ldg(t.rvar + o) + t.eps);
    if (e >= n) {                                   // folded shift
      const float b = t.cbias ? __ldg(t.cbias + o) : 0.f;
      t.out_b[o] = t.gamma ? __ldg(t.beta + o) + (b - __ldg(t.rmean + o)) * scale : b;
      continue;
    }
    const int r = (int)(e - (long long)o * per_o);
    const int i = r / t.kk, tap = r - i * t.kk;
    const float v = __ldg(t.w + e) * scale;
    switch (t.kind) {
      case 0: reinterpret_cast<__nv_bfloat16*>(t.out_w)[((long long)o * t.kk + tap) * t.cin + i] = __float2bfloat16_rn(v); break;
      case 1: reinterpret_cast<float*>(t.out_w)[((long long)o * t.kk + tap) * t.cin + i] = v; break;
      case 2: reinterpret_cast<float*>(t.out_w)[((long long)tap * t.cin + i) * t.cout + o] = v; break;
      case 3: reinterpret_cast<float*>(t.out_w)[(long long)tap * t.ldw + o] = v; break;
      case 4: reinterpret_cast<__nv_bfloat16*>(t.out_w)[((long long)o * t.kk + tap) * 64 + (o & 63)] = __float2bfloat16_rn(v); break;
      default: reinterpret_cast<__nv_bfloat16*>(t.out_w)[(long long)tap * t.ldw + o] = __float2bfloat16_rn(v); break;
    }
  }
}

}  // namespace b200

extern "C" int b200seg_fold_pack_eval_multi(const void* table, const int* chunk_tensor, const int* chunk_index,
                                            int n_chunks, b200seg_stream_t s) {
  B200_REQUIRE(table && chunk_tensor && chunk_index && n_chunks > 0, "fold_pack_eval_multi: bad arguments");
  b200::fold_pack_eval_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)s>>>((const b200::FoldEntry*)table, chunk_tensor,
                                                                                chunk_index);
  return b200::check_launch("fold_pack_eval_multi");
}

extern "C" int b200seg_fold_chunk(void) { return b200::FOLD_CHUNK; }
