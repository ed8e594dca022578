// train_ops.cu -- kernels that exist only on the training path (train.py:35-39): train-mode BatchNorm
// (batch statistics, apply, backward), weight gradients, depthwise backward, the adjoints of the two
// bilinear upsamplings / concat / maxpool.  NHWC activations in storage type T (f32 or bf16), all
// arithmetic and all reductions in f32, cross-CTA accumulation of per-channel statistics in f64.
//
// Dense data gradients (dgrad) need no kernel of their own: dgrad of a stride-1 "same" convolution is
// the same convolution with the weights transposed (and the 3x3 taps flipped), so the engine calls
// b200seg_conv_simt / b200seg_conv_tc with a re-packed weight tensor; its "+residual" epilogue doubles
// as the gradient accumulation for tensors with two consumers.
#include <stdlib.h>
#include "common.cuh"

namespace b200 {

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// per-channel reductions over all pixels of an NHWC tensor viewed as [P, C]
//   block = TX channel-vectors (x) x 256/TX pixel lanes (y), TX = the power of two covering min(C/VN, 32):
//   narrow tensors (C = 16..32, which have the MOST pixels) still keep all 256 threads busy and every warp
//   reads a contiguous run of pixels.  grid = (pixel chunks, channel-vector tiles).  Per-thread partials in
//   f32 over RED_PIX/TY pixels, tree over y in shared memory, one f64 atomic per channel per block.
// ---------------------------------------------------------------------------------------------
static inline dim3 red_block(int cvecs) {
  // x = exactly the channel vectors of one pixel (any count up to 256, not rounded to a power of two: MobileNetV2's
  // widths 24, 96, 144, 160, 192, 320, 576 left 25-45 % of the lanes idle), y = as many pixels as fit in 256 threads.
  // Consecutive threads read consecutive 16-byte vectors whatever the warp/pixel alignment.
  const int tx = cvecs < 256 ? cvecs : 256;
  return dim3(tx, 256 / tx);
}
// the depthwise backward kernels stage their taps for at most 32 channel vectors and assume 256 threads
static inline dim3 red_block_pow2(int cvecs) {
  int tx = 1;
  while (tx < cvecs && tx < 32) tx <<= 1;
  return dim3(tx, 256 / tx);
}
// pixels per block: enough blocks (~8 per SM) to cover the memory latency even for the small deep layers, at least
// 4 pixels per thread so the f64 atomics stay a minor cost
static inline int red_blocks_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("B200SEG_RED_BPS");
    v = e ? atoi(e) : 4;
    if (v < 1) v = 1;
  }
  return v;
}
static inline int red_ppb(long long P, const dim3& block, int cvecs) {
  const long long gy = (cvecs + block.x - 1) / block.x;
  const long long nb = (long long)sm_count() * red_blocks_per_sm();
  long long ppb = (P * gy + nb - 1) / nb;
  const long long lo = 4LL * block.y;
  if (ppb < lo) ppb = lo;
  if (ppb > 8192) ppb = 8192;
  return (int)ppb;
}

// ptxas sinks independent loads into the arithmetic that consumes them when that saves registers, whatever order the
// source (or the PTX) has them in, volatile or not, across warp barriers too.  What does pin a batch is a data dependence:
// one word of every loaded vector is OR-ed together, masked with a value that is zero at run time but that the compiler
// cannot prove zero (batch size >> 30), and OR-ed into an operand of the first arithmetic instruction.  Cost: N/2 LOP3.
__device__ __forceinline__ uint32_t opaque_zero(int positive_below_2_30) { return (uint32_t)positive_below_2_30 >> 30; }
template <int N>
__device__ __forceinline__ uint32_t all_loaded(const uint4 (&r)[N], uint32_t zmask) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) t |= r[i].x;
  return t & zmask;
}

// "Batched streaming": the pixel loop first issues U rows of 16-byte loads per input (raw, not yet unpacked), then does
// the arithmetic.  With the loads written inside the per-pixel body the compiler kept them behind the body's branches,
// i.e. 32 bytes in flight per thread and ~16 KB per SM: the large-layer BatchNorm passes ran at 2.3-2.9 TB/s, latency-bound.
// pixels per block of the apply kernels that derive their per-channel constants themselves (BnFin / SLOTS): enough blocks to
// stream at full rate, few enough that the repeated prologue (<= 32 f64 loads per channel and block) stays negligible
static inline int fat_ppb(long long P, const dim3& block, int cvecs) {
  static int bps = 0;
  if (bps == 0) { const char* e = getenv("B200SEG_BN_FAT_BPS"); bps = e ? atoi(e) : 4; if (bps < 1) bps = 1; }     // measured: 4 -> 8.20 ms, 8 / 16 / 32 -> 8.42 / 8.42 / 8.47
  const long long gy = (cvecs + block.x - 1) / block.x;
  const long long nb = (long long)sm_count() * bps;
  long long ppb = (P * gy + nb - 1) / nb;
  const long long lo = 8LL * block.y;
  if (ppb < lo) ppb = lo;
  if (ppb > 8192) ppb = 8192;
  return (int)ppb;
}

template <typename T, int NOUT, int NIN, int U = 4, bool PIN = false, typename F>
__device__ __forceinline__ void channel_reduce(const T* const (&in)[NIN], long long P, int C, int ppb, double* const* out,
                                               int nslot, long long slot_stride, F&& f) {
  using V = Vec16<T>;
  using Raw = typename V::Raw;
  constexpr int VN = V::N;
  const int TX = blockDim.x, TY = blockDim.y;
  const int cvec = blockIdx.y * TX + threadIdx.x;
  const int c0 = cvec * VN;
  const bool active = c0 < C;
  float acc[NOUT][VN];
#pragma unroll
  for (int o = 0; o < NOUT; ++o)
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[o][j] = 0.f;
  if (active) {
    const long long p0 = (long long)blockIdx.x * ppb, p1 = min(p0 + ppb, P);
    long long p = p0 + threadIdx.y;
    const long long rowstep = (long long)TY * C;
    for (; p + (long long)(U - 1) * TY < p1; p += (long long)U * TY) {
      Raw raw[U][NIN];
      const long long off = p * C + c0;
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < NIN; ++i) raw[u][i] = V::ldraw(in[i] + off + u * rowstep);
      uint32_t dep = 0;
      if constexpr (PIN && sizeof(T) == 2) {          // pin the batch (see all_loaded); measured slower on bn_stats
                                                      // (94 registers, half the resident blocks), so opt-in
        const uint32_t zm = opaque_zero(ppb);
#pragma unroll
        for (int u = 0; u < U; ++u) dep |= all_loaded(raw[u], zm);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        V v[NIN];
#pragma unroll
        for (int i = 0; i < NIN; ++i) v[i].unpack_dep(raw[u][i], dep);
        f(v, acc);
      }
    }
    for (; p < p1; p += TY) {
      V v[NIN];
#pragma unroll
      for (int i = 0; i < NIN; ++i) v[i].load(in[i] + p * C + c0);
      f(v, acc);
    }
  }
  __shared__ float red[256][VN + 1];
  const int tid = threadIdx.y * TX + threadIdx.x;
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VN; ++j) red[tid][j] = acc[o][j];
    __syncthreads();
    int top = 1;
    while (top < TY) top <<= 1;
    for (int stride = top >> 1; stride >= 1; stride >>= 1) {     // tree over the pixel lanes (TY need not be 2^k)
      if (threadIdx.y < stride && threadIdx.y + stride < TY) {
#pragma unroll
        for (int j = 0; j < VN; ++j) red[tid][j] += red[tid + stride * TX][j];
      }
      __syncthreads();
    }
    if (threadIdx.y == 0 && active) {
      // blocks spread their atomics over `nslot` copies of the output (summed by the consumer): with one copy
      // every channel address received one f64 atomic per block (~1200), serialised in L2 -- 50 us per call on
      // the mid-size layers, 10x the time of reading the tensor
      double* dst = out[o] + (long long)(blockIdx.x % nslot) * slot_stride + c0;
#pragma unroll
      for (int j = 0; j < VN; ++j) atomicAdd(dst + j, (double)red[threadIdx.x][j]);
    }
  }
}

// sum and sum of squares (BatchNorm batch statistics)
template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ z, long long P, int C, int ppb, double* sum, double* sumsq, int nslot,
                long long slot_stride) {
  using V = Vec16<T>;
  double* const outs[2] = {sum, sumsq};
  // shifted sums: every channel is offset by its own value at pixel 0, so that sum((z-k)^2) - sum(z-k)^2/n does not
  // cancel catastrophically for channels whose spread is tiny compared with their mean (a nearly constant channel
  // otherwise gets a variance -- hence a d(gamma) -- that is pure rounding noise).
  V k;                                        // this thread's channels at pixel 0 (loop invariant)
  {
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * V::N;
    if (c0 < C) k.load(z + c0);
  }
  const T* const ins[1] = {z};
  channel_reduce<T, 2, 1, 8>(ins, P, C, ppb, outs, nslot, slot_stride, [&](const V (&v)[1], float (&acc)[2][V::N]) {
#pragma unroll
    for (int j = 0; j < V::N; ++j) { const float d = v[0].v[j] - k.v[j]; acc[0][j] += d; acc[1][j] = fmaf(d, d, acc[1][j]); }
  });
}

// one thread per channel: mean, biased var -> invstd, fused scale/shift, running-stat update
// (momentum 0.1, unbiased variance), exactly nn.BatchNorm2d in training mode.
template <typename T>
__global__ void bn_finalize_kernel(const T* __restrict__ z0, const double* __restrict__ sum, const double* __restrict__ sumsq,
                                   int nslot, long long slot_stride, long long n,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean_out,
                                   float* invstd_out, float* scale_out, float* shift_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int sl = 0; sl < nslot; ++sl) { s1 += sum[sl * slot_stride + c]; s2 += sumsq[sl * slot_stride + c]; }
  const double ms = s1 / (double)n;                     // mean of (z - k), k = z at pixel 0
  double var = s2 / (double)n - ms * ms;
  if (var < 0.0) var = 0.0;
  const double m = ms + (double)to_f32<T>(z0[c]);
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  mean_out[c] = (float)m;
  invstd_out[c] = invstd;
  scale_out[c] = sc;
  shift_out[c] = beta[c] - (float)m * sc;
  if (running_mean) {
    const double unbiased = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// "Last block done" epilogue of a cross-CTA reduction: every block fences its atomics and takes a ticket; the block that
// draws the last ticket sees all partial sums (read through L2) and finishes the reduction in the same launch -- the tiny
// finalize / f64->f32 kernels that used to follow each of the 61 BatchNorm reductions (5-6 us of launch + dependency latency
// each, twice per layer per step) disappear.  `counter` is a zeroed 32-bit word per reduction.
// a = act(z * scale + shift) (+ residual).  Thread = one channel vector x several pixels (same (tx, ty) layout as the
// reductions): the per-channel constants are loaded once per thread instead of once per element.
// FIN: the per-channel constants are not read from `scale` / `shift` but derived by every block from the slot sums of
// bn_stats (bn_finalize_kernel's arithmetic, redundantly per block; the blocks with blockIdx.x == 0 also publish mean /
// invstd / scale / shift for the backward pass and update the running statistics).  With ~4 blocks per SM that is a few
// hundred f64 loads per block from L2 -- and one launch plus one dependency edge less per layer in the step graph
// (the 25 one-block finalize launches cost 5 us each on the critical path).
struct BnFin {
  const double* sum; const double* sumsq; int nslot; long long slot_stride; long long n;
  const float* gamma; const float* beta; float eps, momentum;
  float* running_mean; float* running_var; float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
};

template <typename T, bool HAS_RES, bool FIN>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                const T* __restrict__ res, T* __restrict__ a, long long P, int C, int ppb, int act, const BnFin f) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  const int TY = blockDim.y;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * VN;
  float ksc[VN], ksh[VN];
  if constexpr (FIN) {
    __shared__ float s_sc[256 * VN], s_sh[256 * VN];
    const int cb = blockIdx.y * blockDim.x * VN;
    const int nch = min((int)blockDim.x * VN, C - cb);
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < nch; i += blockDim.x * blockDim.y) {
      const int c = cb + i;
      double s1 = 0.0, s2 = 0.0;
      for (int sl = 0; sl < f.nslot; ++sl) { s1 += f.sum[sl * f.slot_stride + c]; s2 += f.sumsq[sl * f.slot_stride + c]; }
      const double ms = s1 / (double)f.n;                     // mean of (z - k), k = z at pixel 0
      double var = s2 / (double)f.n - ms * ms;
      if (var < 0.0) var = 0.0;
      const double m = ms + (double)to_f32<T>(z[c]);
      const float invstd = (float)(1.0 / sqrt(var + (double)f.eps));
      const float sc = f.gamma[c] * invstd;
      const float sh = f.beta[c] - (float)m * sc;
      s_sc[i] = sc;
      s_sh[i] = sh;
      if (blockIdx.x == 0) {
        f.mean_out[c] = (float)m; f.invstd_out[c] = invstd; f.scale_out[c] = sc; f.shift_out[c] = sh;
        if (f.running_mean) {
          const double unbiased = f.n > 1 ? var * ((double)f.n / (double)(f.n - 1)) : var;
          f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * (float)m;
          f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
        }
      }
    }
    __syncthreads();
    if (c0 >= C) return;
#pragma unroll
    for (int j = 0; j < VN; ++j) { ksc[j] = s_sc[threadIdx.x * VN + j]; ksh[j] = s_sh[threadIdx.x * VN + j]; }
  } else {
    if (c0 >= C) return;
#pragma unroll
    for (int j = 0; j < VN; ++j) { ksc[j] = __ldg(scale + c0 + j); ksh[j] = __ldg(shift + c0 + j); }
  }
  const float lo = act != B200SEG_ACT_NONE ? 0.f : -INFINITY, hi = act == B200SEG_ACT_RELU6 ? 6.f : INFINITY;
  const long long p0 = (long long)blockIdx.x * ppb, p1 = min(p0 + ppb, P);
  auto body = [&](const V& v, const V& r, long long off) {
    V o;
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      const float u = fminf(fmaxf(fmaf(v.v[j], ksc[j], ksh[j]), lo), hi);
      o.v[j] = HAS_RES ? u + r.v[j] : u;
    }
    o.store(a + off);
  };
  // batched streaming (see channel_reduce): all loads of U rows first
  constexpr int U = HAS_RES ? 4 : 8;
  using Raw = typename V::Raw;
  long long p = p0 + threadIdx.y;
  const long long rowstep = (long long)TY * C;
  for (; p + (long long)(U - 1) * TY < p1; p += (long long)U * TY) {
    const long long off = p * C + c0;
    Raw rz[U], rr[HAS_RES ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      rz[u] = V::ldraw(z + off + u * rowstep);
      if (HAS_RES) rr[u] = V::ldraw(res + off + u * rowstep);
    }
    uint32_t dep = 0;
    if constexpr (FIN && sizeof(T) == 2) dep = all_loaded(rz, opaque_zero(ppb));   // pin the batch (the finalize prologue's
                                                                                 // registers made ptxas split it 4 + 4)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      V v, r;
      v.unpack_dep(rz[u], dep);
      if (HAS_RES) r.unpack(rr[u]);
      body(v, r, off + u * rowstep);
    }
  }
  for (; p < p1; p += TY) {
    const long long off = p * C + c0;
    V v, r;
    v.load(z + off);
    if (HAS_RES) r.load(res + off);
    body(v, r, off);
  }
}

// Activation gradient with loop-invariant predicates: act'(u) = none || (u > 0 && (nohi || u < 6)); no branch on `act`
// inside the streaming loops (the branches kept the compiler from hoisting the loads of the next rows).
struct ActMask {
  bool none, nohi;
  __device__ __forceinline__ explicit ActMask(int act) : none(act == B200SEG_ACT_NONE), nohi(act != B200SEG_ACT_RELU6) {}
  __device__ __forceinline__ float operator()(float u, float g) const {
    return (none || (u > 0.f && (nohi || u < 6.f))) ? g : 0.f;
  }
};
// gradient of the activation evaluated at u = z*scale + shift
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == B200SEG_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == B200SEG_ACT_RELU6) return (u > 0.f && u < 6.f) ? 1.f : 0.f;
  return 1.f;
}

// sum(g) and sum(g * xhat) with g = da * act'(u), xhat = (z - mean) * invstd
template <typename T, int U = 4, int MINB = 2>
__global__ void __launch_bounds__(256, MINB)
bn_bwd_reduce_kernel(const T* __restrict__ da, const T* __restrict__ z, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                     long long P, int C, int ppb, int act, double* sg, double* sgx, int nslot, long long slot_stride) {
  using V = Vec16<T>;
  double* const outs[2] = {sg, sgx};
  // per-channel constants live in registers for the whole pixel loop (the channel vector of a thread is fixed)
  float ksc[V::N], ksh[V::N], kmu[V::N], kis[V::N];
  {
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * V::N;
#pragma unroll
    for (int j = 0; j < V::N; ++j) {
      const bool ok = c0 + j < C;
      ksc[j] = ok ? __ldg(scale + c0 + j) : 0.f; ksh[j] = ok ? __ldg(shift + c0 + j) : 0.f;
      kmu[j] = ok ? __ldg(mean + c0 + j) : 0.f;  kis[j] = ok ? __ldg(invstd + c0 + j) : 0.f;
    }
  }
  const ActMask am(act);
  const T* const ins[2] = {da, z};
  channel_reduce<T, 2, 2, U>(ins, P, C, ppb, outs, nslot, slot_stride, [&](const V (&v)[2], float (&acc)[2][V::N]) {
#pragma unroll
    for (int j = 0; j < V::N; ++j) {
      const float u = fmaf(v[1].v[j], ksc[j], ksh[j]);
      const float g = am(u, v[0].v[j]);
      acc[0][j] += g;
      acc[1][j] = fmaf(g, (v[1].v[j] - kmu[j]) * kis[j], acc[1][j]);
    }
  });
}

// dz = scale * (g - mean(g) - xhat * mean(g*xhat))        (scale = gamma * invstd; sums passed as f32, mean = sum * inv_n)
// SLOTS: sum(g) / sum(g*xhat) are not read as f32 vectors but summed by every block from the f64 slot copies bn_bwd_reduce
// left in the gradient staging (sg64 / sgx64, `nslot` copies `slot_stride` apart): the f64 -> f32 launch between the two
// BatchNorm-backward kernels disappears from the step graph.
template <typename T, bool SLOTS>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ da, const T* __restrict__ z, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ sg, const float* __restrict__ sgx, float inv_n, T* __restrict__ dz,
                    long long P, int C, int ppb, int act, const double* __restrict__ sg64, const double* __restrict__ sgx64,
                    int nslot, long long slot_stride) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  const int TY = blockDim.y;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * VN;
  float ksc[VN], ksh[VN], kmu[VN], kis[VN], kmg[VN], kmx[VN];
  if constexpr (SLOTS) {
    __shared__ float s_mg[256 * VN], s_mx[256 * VN];
    const int cb = blockIdx.y * blockDim.x * VN;
    const int nch = min((int)blockDim.x * VN, C - cb);
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 2 * nch; i += blockDim.x * blockDim.y) {
      const int which = i >= nch, ci = i - which * nch;
      const double* src = (which ? sgx64 : sg64) + cb + ci;
      double t = 0.0;
      for (int sl = 0; sl < nslot; ++sl) t += src[sl * slot_stride];
      (which ? s_mx : s_mg)[ci] = (float)t * inv_n;            // same rounding as f64_to_f32 followed by * inv_n
    }
    __syncthreads();
    if (c0 >= C) return;
#pragma unroll
    for (int j = 0; j < VN; ++j) { kmg[j] = s_mg[threadIdx.x * VN + j]; kmx[j] = s_mx[threadIdx.x * VN + j]; }
  } else {
    if (c0 >= C) return;
#pragma unroll
    for (int j = 0; j < VN; ++j) { kmg[j] = __ldg(sg + c0 + j) * inv_n; kmx[j] = __ldg(sgx + c0 + j) * inv_n; }
  }
#pragma unroll
  for (int j = 0; j < VN; ++j) {
    ksc[j] = __ldg(scale + c0 + j); ksh[j] = __ldg(shift + c0 + j);
    kmu[j] = __ldg(mean + c0 + j);  kis[j] = __ldg(invstd + c0 + j);
  }
  const long long p0 = (long long)blockIdx.x * ppb, p1 = min(p0 + ppb, P);
  const ActMask am(act);
  auto body = [&](const V& d, const V& v, long long off) {
    V o;
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      const float u = fmaf(v.v[j], ksc[j], ksh[j]);
      const float g = am(u, d.v[j]);
      const float xh = (v.v[j] - kmu[j]) * kis[j];
      o.v[j] = ksc[j] * (g - kmg[j] - xh * kmx[j]);
    }
    o.store(dz + off);
  };
  // batched streaming (see channel_reduce): all loads of U rows first
  constexpr int U = 4;
  using Raw = typename V::Raw;
  long long p = p0 + threadIdx.y;
  const long long rowstep = (long long)TY * C;
  for (; p + (long long)(U - 1) * TY < p1; p += (long long)U * TY) {
    const long long off = p * C + c0;
    Raw rd[U], rz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { rd[u] = V::ldraw(da + off + u * rowstep); rz[u] = V::ldraw(z + off + u * rowstep); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      V d, v;
      d.unpack(rd[u]); v.unpack(rz[u]);
      body(d, v, off + u * rowstep);
    }
  }
  for (; p < p1; p += TY) {
    const long long off = p * C + c0;
    V d, v;
    d.load(da + off); v.load(z + off);
    body(d, v, off);
  }
}

// dz = da * act'(z + bias) for a conv+bias(+act) layer without BatchNorm; also used with act = NONE as a cast/copy
template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* __restrict__ da, const T* __restrict__ a_out, T* __restrict__ dz, long long N, int act) {
  using V = Vec16<T>;
  const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V::N;
  if (idx >= N) return;
  V d, o, v;
  d.load(da + idx);
  v.load(a_out + idx);
#pragma unroll
  for (int j = 0; j < V::N; ++j) o.v[j] = d.v[j] * act_grad(v.v[j], act);
  o.store(dz + idx);
}

// per-channel column sum (bias gradients)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long P, int C, int ppb, double* out, int nslot, long long slot_stride) {
  using V = Vec16<T>;
  double* const outs[1] = {out};
  const T* const ins[1] = {x};
  channel_reduce<T, 1, 1, 8>(ins, P, C, ppb, outs, nslot, slot_stride, [&](const V (&v)[1], float (&acc)[1][V::N]) {
#pragma unroll
    for (int j = 0; j < V::N; ++j) acc[0][j] += v[0].v[j];
  });
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n, int nslot,
                                  long long slot_stride, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = 0.0;
  for (int sl = 0; sl < nslot; ++sl) v += in[sl * slot_stride + i];
  out[i] = (float)v * scale;
}

// ---------------------------------------------------------------------------------------------
// dense weight gradient: dW[co][tap][ci] += sum_p dz[p][co] * x[p shifted by tap][ci]
// implicit GEMM with the PIXELS as the reduction axis, split over blockIdx.z, f32 atomics at the end.
// tile 64 (co) x 64 (k = tap*Cin+ci), 16 pixels per smem step, 4x4 outputs per thread.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dz, float* __restrict__ dw, int B, int H, int W,
                  int Cin, int Cout, int taps, long long pix_per_split) {
  __shared__ float Ds[16][64 + 4];
  __shared__ float Xs[16][64 + 4];
  const int K = taps * Cin;
  const long long P = (long long)B * H * W;
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const long long p_begin = (long long)blockIdx.z * pix_per_split, p_end = min(p_begin + pix_per_split, P);
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int lp = tid / 16, l4 = (tid % 16) * 4;     // loader: pixel lane, 4 consecutive columns
  const int kk = k0 + l4;
  int tap = 0, ci = 0, dh = 0, dwv = 0;
  const bool k_ok = kk < K;
  if (k_ok) {
    tap = kk / Cin; ci = kk - tap * Cin;
    if (taps == 9) { dh = tap / 3 - 1; dwv = tap % 3 - 1; }
  }
  const int co = n0 + l4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pb = p_begin; pb < p_end; pb += 16) {
    const long long p = pb + lp;
    float dv[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p_end) {
      if (co < Cout) {
        const T* dp = dz + p * Cout + co;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (co + i < Cout) dv[i] = to_f32<T>(dp[i]);
      }
      if (k_ok) {
        const int w_ = (int)(p % W);
        const long long t = p / W;
        const int h_ = (int)(t % H);
        const int b_ = (int)(t / H);
        const int hi = h_ + dh, wi = w_ + dwv;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
          const T* xp = x + (((long long)b_ * H + hi) * W + wi) * Cin + ci;
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = to_f32<T>(xp[i]);   // Cin % 4 == 0: the 4 columns share the tap
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { Ds[lp][l4 + i] = dv[i]; Xs[lp][l4 + i] = xv[i]; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Ds[q][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Xs[q][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dw + (long long)n * K + k, acc[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 backward
// ---------------------------------------------------------------------------------------------
// dx[hi][wi][c] = sum over taps with (hi+1-kh) = S*ho, (wi+1-kw) = S*wo of dz[ho][wo][c] * w[kh*3+kw][c]
// Block = TX channel vectors x TY pixel lanes; the 9 x (TX*VN) tap weights of the block's channels are staged in
// shared memory once, every thread then walks `ppb/TY` input pixels (incremental (b,h,w) bookkeeping, no divisions).
template <typename T>
__global__ void __launch_bounds__(256)
dw_dgrad_kernel(const T* __restrict__ dz, const float* __restrict__ w, const T* __restrict__ acc_in, T* __restrict__ dx,
                int B, int H, int W, int C, int Ho, int Wo, int S, int ppb) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  __shared__ __align__(16) float sw[9][32 * VN];
  const int TX = blockDim.x, TY = blockDim.y;
  const int cbase = blockIdx.y * TX * VN;
  for (int i = threadIdx.y * TX + threadIdx.x; i < 9 * TX * VN; i += 256) {
    const int t = i / (TX * VN), c = i - t * (TX * VN);
    sw[t][c] = (cbase + c < C) ? w[t * C + cbase + c] : 0.f;
  }
  __syncthreads();
  const int c0 = cbase + threadIdx.x * VN;
  if (c0 >= C) return;
  const long long P = (long long)B * H * W;
  const long long p0 = (long long)blockIdx.x * ppb, p1 = min(p0 + ppb, P);
  long long p = p0 + threadIdx.y;
  if (p >= p1) return;
  int wi = (int)(p % W);
  long long t_ = p / W;
  int hi = (int)(t_ % H), b = (int)(t_ / H);
  for (; p < p1; p += TY) {
    V o;
    if (acc_in) o.load(acc_in + p * C + c0);
    else {
#pragma unroll
      for (int j = 0; j < VN; ++j) o.v[j] = 0.f;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = hi + 1 - kh;
      const int ho = S == 1 ? th : th >> 1;
      const bool hok = th >= 0 && (S == 1 || !(th & 1)) && ho < Ho;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = wi + 1 - kw;
        const int wo = S == 1 ? tw : tw >> 1;
        const bool ok = hok && tw >= 0 && (S == 1 || !(tw & 1)) && wo < Wo;
        if (ok) {
          V d;
          d.load(dz + (((long long)b * Ho + ho) * Wo + wo) * C + c0);
          const float* wp = &sw[kh * 3 + kw][threadIdx.x * VN];
#pragma unroll
          for (int j = 0; j < VN; j += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + j);
            o.v[j] = fmaf(d.v[j], w4.x, o.v[j]); o.v[j + 1] = fmaf(d.v[j + 1], w4.y, o.v[j + 1]);
            o.v[j + 2] = fmaf(d.v[j + 2], w4.z, o.v[j + 2]); o.v[j + 3] = fmaf(d.v[j + 3], w4.w, o.v[j + 3]);
          }
        }
      }
    }
    o.store(dx + p * C + c0);
    wi += TY;
    while (wi >= W) { wi -= W; if (++hi == H) { hi = 0; ++b; } }
  }
}

// Stride-2 data gradient, bf16 (the four stride-2 depthwise layers of the encoder; the generic kernel above spent 348 us per
// training step on them, 5x the time of their bytes).  With stride 2 the tap that links an input pixel to an output pixel is
// fixed by the pixel's parity, so a thread owns a 2 x 2 block of INPUT pixels (rows 2a, 2a+1; columns 2b, 2b+1) x 8 channels:
//   dx[2a  ][2b  ] = dz[a][b] w11                       dx[2a  ][2b+1] = dz[a][b+1] w10 + dz[a][b] w12
//   dx[2a+1][2b  ] = dz[a+1][b] w01 + dz[a][b] w21      dx[2a+1][2b+1] = dz[a+1][b+1] w00 + dz[a+1][b] w02 + dz[a][b+1] w20 + dz[a][b] w22
// four 16-byte loads, nine mixed-precision FMAs per channel (taps rounded to bf16 like the forward kernel's), four stores;
// the taps stay in registers over the thread's blocks.  Block = TX channel vectors x TY block lanes as everywhere here.
__global__ void __launch_bounds__(256, 3)
dw_dgrad_s2_bf16_kernel(const __nv_bfloat16* __restrict__ dz, const float* __restrict__ w, const __nv_bfloat16* __restrict__ acc_in,
                        __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo, int bpb) {
  const int TY = blockDim.y;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= C) return;
  uint4 wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + t * C + c0)), b = __ldg(reinterpret_cast<const float4*>(w + t * C + c0 + 4));
    wt[t] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
  const long long NB = (long long)B * Ho * Wo;                 // one block per OUTPUT pixel (a, b)
  const long long q0 = (long long)blockIdx.x * bpb, q1 = min(q0 + bpb, NB);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (long long q = q0 + threadIdx.y; q < q1; q += TY) {
    const int bo = (int)(q % Wo);
    const long long t_ = q / Wo;
    const int ao = (int)(t_ % Ho), bb = (int)(t_ / Ho);
    const __nv_bfloat16* zp = dz + (((long long)bb * Ho + ao) * Wo + bo) * C + c0;
    const bool a1 = ao + 1 < Ho, b1 = bo + 1 < Wo;
    const uint4 d00 = __ldg(reinterpret_cast<const uint4*>(zp));
    const uint4 d01 = b1 ? __ldg(reinterpret_cast<const uint4*>(zp + C)) : zero;
    const uint4 d10 = a1 ? __ldg(reinterpret_cast<const uint4*>(zp + (long long)Wo * C)) : zero;
    const uint4 d11 = (a1 && b1) ? __ldg(reinterpret_cast<const uint4*>(zp + (long long)Wo * C + C)) : zero;
    const int hi = 2 * ao, wi = 2 * bo;
    const long long o00 = (((long long)bb * H + hi) * W + wi) * C + c0;
    const bool r1 = hi + 1 < H, cc1 = wi + 1 < W;
    auto emit = [&](const long long off, const float (&v)[8]) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = v[j];
      if (acc_in) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(acc_in + off));
        f[0] += bf16lo(u.x); f[1] += bf16hi(u.x); f[2] += bf16lo(u.y); f[3] += bf16hi(u.y);
        f[4] += bf16lo(u.z); f[5] += bf16hi(u.z); f[6] += bf16lo(u.w); f[7] += bf16hi(u.w);
      }
      *reinterpret_cast<uint4*>(dx + off) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    };
#define DW_MAC(acc, d, wv)                                                                                   \
    acc[0] = fma_bf16_lo(d.x, wv.x, acc[0]); acc[1] = fma_bf16_hi(d.x, wv.x, acc[1]);                        \
    acc[2] = fma_bf16_lo(d.y, wv.y, acc[2]); acc[3] = fma_bf16_hi(d.y, wv.y, acc[3]);                        \
    acc[4] = fma_bf16_lo(d.z, wv.z, acc[4]); acc[5] = fma_bf16_hi(d.z, wv.z, acc[5]);                        \
    acc[6] = fma_bf16_lo(d.w, wv.w, acc[6]); acc[7] = fma_bf16_hi(d.w, wv.w, acc[7]);
    {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      DW_MAC(v, d00, wt[4])
      emit(o00, v);
    }
    if (cc1) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      DW_MAC(v, d01, wt[3]) DW_MAC(v, d00, wt[5])
      emit(o00 + C, v);
    }
    if (r1) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      DW_MAC(v, d10, wt[1]) DW_MAC(v, d00, wt[7])
      emit(o00 + (long long)W * C, v);
    }
    if (r1 && cc1) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      DW_MAC(v, d11, wt[0]) DW_MAC(v, d10, wt[2]) DW_MAC(v, d01, wt[6]) DW_MAC(v, d00, wt[8])
      emit(o00 + (long long)W * C + C, v);
    }
#undef DW_MAC
  }
}

// dw[tap][c] += sum_p dz[p][c] * x[p*S + tap - 1][c].  Same block shape; predicated (zero-filled) tap loads instead
// of branches, incremental pixel bookkeeping; the 9 partial sums per channel go through the shared tree reduction.
template <typename T>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dz, int B, int H, int W, int C, int Ho, int Wo, int S,
                int ppb, double* dw /* [nslot][9][C] */, int nslot) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  const int TX = blockDim.x, TY = blockDim.y;
  const int c0 = (blockIdx.y * TX + threadIdx.x) * VN;
  const bool active = c0 < C;
  const long long P = (long long)B * Ho * Wo;
  float acc[9][VN];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[t][j] = 0.f;
  const long long p0 = (long long)blockIdx.x * ppb, p1 = min(p0 + ppb, P);
  long long p = p0 + threadIdx.y;
  if (active && p < p1) {
    int wo = (int)(p % Wo);
    long long t_ = p / Wo;
    int ho = (int)(t_ % Ho), b = (int)(t_ / Ho);
    const long long rowC = (long long)W * C;
    for (; p < p1; p += TY) {
      V d;
      d.load(dz + p * C + c0);
      const int hi0 = ho * S - 1, wi0 = wo * S - 1;
      const T* xb = x + (((long long)b * H + hi0) * W + wi0) * C + c0;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const bool hok = (unsigned)(hi0 + kh) < (unsigned)H;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const bool ok = hok && (unsigned)(wi0 + kw) < (unsigned)W;
          V xv;
          if (ok) xv.load(xb + kh * rowC + kw * C);
          else {
#pragma unroll
            for (int j = 0; j < VN; ++j) xv.v[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[kh * 3 + kw][j] = fmaf(d.v[j], xv.v[j], acc[kh * 3 + kw][j]);
        }
      }
      wo += TY;
      while (wo >= Wo) { wo -= Wo; if (++ho == Ho) { ho = 0; ++b; } }
    }
  }
  __shared__ float red[256][VN + 1];
  const int tid = threadIdx.y * TX + threadIdx.x;
#pragma unroll
  for (int o = 0; o < 9; ++o) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < VN; ++j) red[tid][j] = acc[o][j];
    __syncthreads();
    int top = 1;
    while (top < TY) top <<= 1;
    for (int stride = top >> 1; stride >= 1; stride >>= 1) {
      if (threadIdx.y < stride && threadIdx.y + stride < TY) {
#pragma unroll
        for (int j = 0; j < VN; ++j) red[tid][j] += red[tid + stride * TX][j];
      }
      __syncthreads();
    }
    if (threadIdx.y == 0 && active) {
      double* dst = dw + ((long long)(blockIdx.x % nslot) * 9 + o) * C + c0;      // one of nslot copies (see channel_reduce)
#pragma unroll
      for (int j = 0; j < VN; ++j) atomicAdd(dst + j, (double)red[threadIdx.x][j]);
    }
  }
}

// bf16 specialisation: a thread owns PW consecutive output pixels of one row x 8 channels.  The 3 x ((PW-1)S+3) input
// vectors they touch are loaded once (instead of 9 per pixel) and both operands stay packed bf16: every product goes through
// the mixed-precision FMA (f32 += bf16 * bf16, exact product), so the loop has no unpack instructions at all.
// One block per SM (up to 255 registers): with the 72 accumulators capped at 128 registers ptxas serialised the loads two
// at a time; now all PW + 3*NCOL loads of a pixel group are in flight together.  Index arithmetic is 32-bit (the launcher
// checks the tensors have < 2^31 elements): the four 64-bit divisions per group cost more instructions than the FMAs.
// Epilogue: every thread parks its 72 partial sums in shared memory ([pixel lane][tap][channel]), ONE barrier, then each
// (tap, channel) is summed over the pixel lanes by one thread and leaves as one f64 atomic -- the 9 tree reductions
// (45 barriers) per block were most of the run time of the small deep layers.
template <int S, int PW>
__global__ void __launch_bounds__(256, 1)
dw_wgrad_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz, int B, int H, int W, int C,
                     int Ho, int Wo, int gpb /* pixel groups per block */, double* dw /* [nslot][9][C] */, int nslot) {
  constexpr int NCOL = (PW - 1) * S + 3;
  extern __shared__ __align__(16) float dww_red[];         // [TY][9][TX * 8]
  const int TX = blockDim.x, TY = blockDim.y;
  const int c0 = (blockIdx.y * TX + threadIdx.x) * 8;
  const bool active = c0 < C;
  const unsigned gw = (Wo + PW - 1) / PW;                  // pixel groups per output row
  const unsigned G = (unsigned)B * Ho * gw;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const unsigned g0 = blockIdx.x * (unsigned)gpb, g1 = min(g0 + (unsigned)gpb, G);
  if (active) {
    const unsigned rowC = (unsigned)W * C;
    for (unsigned gi = g0 + threadIdx.y; gi < g1; gi += TY) {
      const unsigned row = gi / gw, gx = gi - row * gw;    // row = b * Ho + ho
      const unsigned b = row / (unsigned)Ho, ho = row - b * Ho;
      const int wo0 = gx * PW;
      const int hi0 = (int)ho * S - 1, wi0 = wo0 * S - 1;
      const __nv_bfloat16* dp = dz + ((size_t)row * Wo + wo0) * C + c0;
      const __nv_bfloat16* xb = x + (size_t)b * H * rowC + c0;
      uint4 d[PW], xr[3][NCOL];
#pragma unroll
      for (int o = 0; o < PW; ++o)
        d[o] = (wo0 + o < Wo) ? __ldg(reinterpret_cast<const uint4*>(dp + o * C)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const bool hok = (unsigned)(hi0 + kh) < (unsigned)H;
        const unsigned ro = (unsigned)(hi0 + kh) * rowC;
#pragma unroll
        for (int i = 0; i < NCOL; ++i)
          xr[kh][i] = (hok && (unsigned)(wi0 + i) < (unsigned)W)
                          ? __ldg(reinterpret_cast<const uint4*>(xb + ro + (unsigned)(wi0 + i) * C))
                          : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float* a = acc[kh * 3 + kw];
#pragma unroll
          for (int o = 0; o < PW; ++o) {
            const uint4 xv = xr[kh][o * S + kw], dv = d[o];
            a[0] = fma_bf16_lo(xv.x, dv.x, a[0]); a[1] = fma_bf16_hi(xv.x, dv.x, a[1]);
            a[2] = fma_bf16_lo(xv.y, dv.y, a[2]); a[3] = fma_bf16_hi(xv.y, dv.y, a[3]);
            a[4] = fma_bf16_lo(xv.z, dv.z, a[4]); a[5] = fma_bf16_hi(xv.z, dv.z, a[5]);
            a[6] = fma_bf16_lo(xv.w, dv.w, a[6]); a[7] = fma_bf16_hi(xv.w, dv.w, a[7]);
          }
        }
    }
  }
  const int CT = TX * 8, nout = 9 * CT;
  float* mine = dww_red + (size_t)threadIdx.y * nout + threadIdx.x * 8;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    *reinterpret_cast<float4*>(mine + t * CT) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    *reinterpret_cast<float4*>(mine + t * CT + 4) = make_float4(acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
  }
  __syncthreads();
  double* dst = dw + (size_t)(blockIdx.x % nslot) * 9 * C;              // one of nslot copies (see channel_reduce)
  for (int o = threadIdx.y * TX + threadIdx.x; o < nout; o += TX * TY) {
    const int t = o / CT, cl = o - t * CT;
    const int c = blockIdx.y * CT + cl;
    if (c >= C) continue;
    float sum = 0.f;
    for (int l = 0; l < TY; ++l) sum += dww_red[l * nout + o];
    atomicAdd(dst + t * C + c, (double)sum);
  }
}

// stem / first conv weight gradient: x NCHW [B,Cin<=4,H,W], dz NHWC [B,Ho,Wo,Cout]; dw f32 [3][3][Cin][Cout]
template <typename TI, typename T>
__global__ void __launch_bounds__(256)
smallcin_wgrad_kernel(const TI* __restrict__ x, const T* __restrict__ dz, float* __restrict__ dw, int B, int Cin, int H,
                      int W, int Cout, int Ho, int Wo, int S, int pix_per_block) {
  extern __shared__ float sm[];            // [PIXB][Cout] dz, then [PIXB][9*Cin] patches
  constexpr int PIXB = 32;
  float* sd = sm;
  float* sx = sm + PIXB * Cout;
  const int KK = 9 * Cin;
  const long long P = (long long)B * Ho * Wo;
  const long long p_begin = (long long)blockIdx.x * pix_per_block, p_end = min(p_begin + pix_per_block, P);
  // a thread owns up to 2 items = (k, 4 consecutive output channels): per pixel one broadcast read of the patch
  // value and one 16-byte read of dz feed 4 FMAs (the scalar-pair version needed 2 shared loads per FMA)
  const int c4n = Cout >> 2;
  const int nitems = KK * c4n;               // <= 512
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pb = p_begin; pb < p_end; pb += PIXB) {
    __syncthreads();
    for (int i = threadIdx.x; i < PIXB * Cout; i += blockDim.x) {
      const long long p = pb + i / Cout;
      sd[i] = p < p_end ? to_f32<T>(dz[p * Cout + i % Cout]) : 0.f;
    }
    for (int i = threadIdx.x; i < PIXB * KK; i += blockDim.x) {
      const long long p = pb + i / KK;
      const int k = i % KK;
      float v = 0.f;
      if (p < p_end) {
        const int wo = (int)(p % Wo);
        const long long t = p / Wo;
        const int ho = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int tap = k / Cin, c = k % Cin;
        const int hi = ho * S - 1 + tap / 3, wi = wo * S - 1 + tap % 3;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = to_f32<TI>(x[(((long long)b * Cin + c) * H + hi) * W + wi]);
      }
      sx[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int it = threadIdx.x + i * 256;
      if (it < nitems) {
        const int k = it / c4n, c4 = (it - k * c4n) * 4;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
        for (int q = 0; q < PIXB; ++q) {
          const float a = sx[q * KK + k];
          const float4 d = *reinterpret_cast<const float4*>(sd + q * Cout + c4);
          s0 = fmaf(a, d.x, s0); s1 = fmaf(a, d.y, s1); s2 = fmaf(a, d.z, s2); s3 = fmaf(a, d.w, s3);
        }
        acc[i][0] += s0; acc[i][1] += s1; acc[i][2] += s2; acc[i][3] += s3;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int it = threadIdx.x + i * 256;
    if (it < nitems) {
      const int k = it / c4n, c4 = (it - k * c4n) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dw + k * Cout + c4 + j, acc[i][j]);     // layout [tap][c][co]
    }
  }
}

// Row-tile version for Cin == 3 (the stems): a job = one output row segment of 128 pixels.  The 9 (channel, kh) input
// rows it touches are staged in shared memory with coalesced loads (as the forward stem kernel does), dz of the segment
// as f32 [128][Cout].  Thread = (pixel quarter, 4 consecutive k = (tap, c) x 4 consecutive output channels): per pixel
// 4 patch reads + one 16-byte dz read feed 16 FMAs; the 16 partial sums stay in registers across all jobs of the block
// and are reduced over the four pixel quarters and added to dw once at the end.
template <typename TI, typename T, int S>
__global__ void __launch_bounds__(256)
smallcin3_wgrad_tile_kernel(const TI* __restrict__ x, const T* __restrict__ dz, float* __restrict__ dw, int B, int H, int W,
                            int Cout, int Ho, int Wo) {
  constexpr int NCOL = 127 * S + 3;
  constexpr int PITCH = NCOL + 3;
  extern __shared__ float sm[];
  float* tile = sm;                               // [9][PITCH]   row = c*3 + kh, column j <-> input column wo0*S - 1 + j
  float* sd = sm + ((9 * PITCH + 3) & ~3);        // [128][Cout], 16-byte aligned for the float4 reads
  const int c4n = Cout >> 2;
  const int nitems = 7 * c4n;                     // 7 groups of 4 k (27 -> 28)
  const int pq = threadIdx.x >> 6, r = threadIdx.x & 63;
  int toff[2][4], c4v[2];
  bool act[2];
  float acc[2][4][4];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int item = r + it * 64;
    act[it] = item < nitems;
    const int kq = item / c4n;
    c4v[it] = (item - kq * c4n) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kq * 4 + j;                   // k = tap*3 + c  (matches dw's [kh][kw][c][co] layout)
      const int tap = k / 3, c = k - tap * 3, kh = tap / 3, kw = tap - kh * 3;
      toff[it][j] = (k < 27) ? (c * 3 + kh) * PITCH + kw : -1;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[it][j][q] = 0.f;
    }
  }
  const int segs = (Wo + 127) / 128;
  const long long total = (long long)B * Ho * segs;
  const long long HW = (long long)H * W;
  for (long long job = blockIdx.x; job < total; job += gridDim.x) {
    const int sg = (int)(job % segs);
    long long p = job / segs;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int wo0 = sg * 128;
    const TI* xb = x + (long long)b * 3 * HW;
    const int wi0 = wo0 * S - 1, hi0 = ho * S - 1;
    __syncthreads();                              // previous job's reads are done
    for (int i = threadIdx.x; i < 9 * NCOL; i += 256) {
      const int rowi = i / NCOL, col = i - rowi * NCOL;
      const int c = rowi / 3, kh = rowi - c * 3;
      const int hi = hi0 + kh, wi = wi0 + col;
      float v = 0.f;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = to_f32<TI>(xb[(long long)c * HW + (long long)hi * W + wi]);
      tile[rowi * PITCH + col] = v;
    }
    const T* dzr = dz + (((long long)b * Ho + ho) * Wo + wo0) * Cout;
    for (int i = threadIdx.x; i < 128 * Cout; i += 256) {
      const int px = i / Cout;
      sd[i] = (wo0 + px < Wo) ? to_f32<T>(dzr[i]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      if (!act[it]) continue;
      const float* sdp = sd + c4v[it];
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const int px = pq * 32 + i;
        const float4 d = *reinterpret_cast<const float4*>(sdp + px * Cout);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = toff[it][j] >= 0 ? tile[toff[it][j] + px * S] : 0.f;
          acc[it][j][0] = fmaf(a, d.x, acc[it][j][0]); acc[it][j][1] = fmaf(a, d.y, acc[it][j][1]);
          acc[it][j][2] = fmaf(a, d.z, acc[it][j][2]); acc[it][j][3] = fmaf(a, d.w, acc[it][j][3]);
        }
      }
    }
  }
  // reduce the four pixel quarters through shared memory, then one atomic per (k, co) per block
  __syncthreads();
  float* red = sm;                                // [4][28 * Cout] <= 4 * 28 * 64 floats = 28 KB (tile + sd are >= that)
  const int stride_q = 28 * Cout;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    if (!act[it]) continue;
    const int item = r + it * 64, kq = item / c4n;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) red[pq * stride_q + (kq * 4 + j) * Cout + c4v[it] + q] = acc[it][j][q];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * Cout; i += 256)
    atomicAdd(dw + i, red[i] + red[stride_q + i] + red[2 * stride_q + i] + red[3 * stride_q + i]);
}

// Stem weight gradient for bf16 dz (the training precision of the path), Cin = 3, Cout = 32 / 64.  Same job as the row-tile
// kernel above (one 128-pixel output row segment, its 9 input rows and its dz staged in shared memory), rebuilt around what
// the profile of that kernel showed (252 us for 117 MB: 25 scalar global loads per thread between two barriers, every job):
//   * the NEXT job's global loads (<= 10 x values, 2-4 16-byte dz vectors per thread) are issued before the current job's
//     arithmetic and parked in registers, so their latency is hidden by ~1000 cycles of FMAs instead of being exposed
//     between two barriers; dz is read with 16-byte loads; the per-thread staging coordinates are computed once;
//   * warp = 16 pixels of the segment, lane = (4 consecutive k = (tap, c)) x (8 consecutive output channels): per pixel
//     two 16-byte dz reads + 4 patch reads feed 32 FMAs (was 1 + 4 for 16);
//   * partial sums stay in registers across all jobs of the block; one shared-memory reduction over the 8 warps and
//     27 * Cout atomics per block at the end.
template <typename TI, int S, int COUT>
__global__ void __launch_bounds__(256, 2)
stem_wgrad_bf16_kernel(const TI* __restrict__ x, const __nv_bfloat16* __restrict__ dz, float* __restrict__ dw, int B, int H,
                       int W, int Ho, int Wo) {
  constexpr int NCOL = 127 * S + 3, PITCH = NCOL + 3;
  constexpr int NX = (9 * NCOL + 255) / 256;            // x values staged per thread and job
  constexpr int C8N = COUT / 8;
  constexpr int ND = 128 * C8N / 256;                   // 16-byte dz vectors per thread and job
  constexpr int NITEMS = 7 * C8N, NIT = (NITEMS + 31) / 32;
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                                     // [9][PITCH]   row = c*3 + kh, column j <-> input column wo0*S - 1 + j
  float* sd = sm + ((9 * PITCH + 3) & ~3);              // [128][COUT] f32
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = H * W;
  // staging coordinates of this thread (the same for every job), packed: tile row (c*3 + kh) << 16 | column; -1 = none
  int rc[NX];
#pragma unroll
  for (int k = 0; k < NX; ++k) {
    const int e = tid + 256 * k;
    const int rowi = e / NCOL, col = e - rowi * NCOL;
    rc[k] = e < 9 * NCOL ? (rowi << 16) | col : -1;
  }
  // arithmetic coordinates
  int toff[NIT][4], c8v[NIT];
  bool act[NIT];
  float acc[NIT][4][8];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int item = lane + 32 * it;
    act[it] = item < NITEMS;
    const int kq = act[it] ? item / C8N : 0;
    c8v[it] = (item - (item / C8N) * C8N) * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kq * 4 + j;                         // k = tap*3 + c  (dw layout [kh][kw][c][co]); k = 27 is padding
      const int tap = k / 3, c = k - tap * 3, kh = tap / 3, kw = tap - kh * 3;
      toff[it][j] = (k < 27) ? (c * 3 + kh) * PITCH + kw : 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[it][j][q] = 0.f;
    }
  }
  const int segs = (Wo + 127) / 128;
  const int total = B * Ho * segs;                      // jobs (launcher: < 2^31)
  float xr[NX];
  uint4 dr[ND];
  auto fetch = [&](int job) {
    const int sg = job % segs, rowj = job / segs;       // rowj = b * Ho + ho
    const int ho = rowj % Ho, b = rowj / Ho;
    const int wo0 = sg * 128, wi0 = wo0 * S - 1, hi0 = ho * S - 1;
    const TI* xb = x + (long long)b * 3 * HW + (long long)hi0 * W + wi0;
#pragma unroll
    for (int k = 0; k < NX; ++k) {
      const int rowi = rc[k] >> 16, col = rc[k] & 0xffff;
      const int c = (rowi * 11) >> 5, kh = rowi - c * 3;          // rowi / 3 for rowi < 9
      const int hi = hi0 + kh, wi = wi0 + col;
      const bool ok = rc[k] >= 0 && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
      xr[k] = ok ? to_f32<TI>(__ldg(xb + (c * HW + kh * W + col))) : 0.f;
    }
    const __nv_bfloat16* dzr = dz + ((long long)rowj * Wo + wo0) * COUT;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const int q = tid + 256 * k, px = q / C8N;
      dr[k] = (wo0 + px < Wo) ? __ldg(reinterpret_cast<const uint4*>(dzr) + q) : make_uint4(0, 0, 0, 0);
    }
  };
  int job = blockIdx.x;
  if (job < total) fetch(job);
  for (; job < total; job += gridDim.x) {
    __syncthreads();                                    // the previous job's reads of tile / sd are done
#pragma unroll
    for (int k = 0; k < NX; ++k)
      if (rc[k] >= 0) tile[(rc[k] >> 16) * PITCH + (rc[k] & 0xffff)] = xr[k];
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const int q = tid + 256 * k;
      const uint4 u = dr[k];
      *reinterpret_cast<float4*>(sd + q * 8) = make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
      *reinterpret_cast<float4*>(sd + q * 8 + 4) = make_float4(bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w));
    }
    __syncthreads();
    if (job + (int)gridDim.x < total) fetch(job + gridDim.x);      // in flight during the arithmetic below
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      if (!act[it]) continue;
      const float* sdp = sd + c8v[it];
      const float* t0 = tile + toff[it][0];
      const float* t1 = tile + toff[it][1];
      const float* t2 = tile + toff[it][2];
      const float* t3 = tile + toff[it][3];
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int px = warp * 16 + i;
        const float4 d0 = *reinterpret_cast<const float4*>(sdp + px * COUT);
        const float4 d1 = *reinterpret_cast<const float4*>(sdp + px * COUT + 4);
        const float a[4] = {t0[px * S], t1[px * S], t2[px * S], t3[px * S]};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* r = acc[it][j];
          r[0] = fmaf(a[j], d0.x, r[0]); r[1] = fmaf(a[j], d0.y, r[1]); r[2] = fmaf(a[j], d0.z, r[2]); r[3] = fmaf(a[j], d0.w, r[3]);
          r[4] = fmaf(a[j], d1.x, r[4]); r[5] = fmaf(a[j], d1.y, r[5]); r[6] = fmaf(a[j], d1.z, r[6]); r[7] = fmaf(a[j], d1.w, r[7]);
        }
      }
    }
  }
  // reduce the eight pixel groups (warps) through shared memory, then one atomic per (k, co) per block
  __syncthreads();
  float* red = sm;                                      // [8][28 * COUT]
  constexpr int STRIDE_W = 28 * COUT;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    if (!act[it]) continue;
    const int item = lane + 32 * it, kq = item / C8N;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* dst = red + warp * STRIDE_W + (kq * 4 + j) * COUT + c8v[it];
      *reinterpret_cast<float4*>(dst) = make_float4(acc[it][j][0], acc[it][j][1], acc[it][j][2], acc[it][j][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[it][j][4], acc[it][j][5], acc[it][j][6], acc[it][j][7]);
    }
  }
  __syncthreads();
  for (int i = tid; i < 27 * COUT; i += 256) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w * STRIDE_W + i];
    atomicAdd(dw + i, sum);
  }
}

// ---------------------------------------------------------------------------------------------
// adjoints of the bilinear upsamplings.  Gather form: every source pixel enumerates the few output
// pixels that read it and re-evaluates PyTorch's forward index/weight formula for each, so clamped
// borders are exact by construction.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bil_weight_ac_false(int r, int i, int n_in) {
  // weight with which output index r reads input index i (scale 2, align_corners=False)
  const float s = fmaxf(0.5f * (r + 0.5f) - 0.5f, 0.f);
  const int i0 = (int)s;
  const int i1 = min(i0 + 1, n_in - 1);
  const float l = s - i0;
  return (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
}
__device__ __forceinline__ float bil_weight_ac_true(int r, int i, int n_in, float scale) {
  const float s = scale * r;
  const int i0 = (int)s;
  const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  const float l = s - i0;
  return (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
}

// dcat [B,2h,2w,Cs+Cu] -> dskip [B,2h,2w,Cs] (+= acc_skip) and dx [B,h,w,Cu].
// Threads [0, n_skip) copy the skip slice, the rest gather the upsampled slice: warps are uniform (one thread per
// low-resolution pixel x channel vector of ITS slice), and both paths issue all their loads before any arithmetic --
// the 4 x 4 gather reads clamped coordinates with zero weights outside the image instead of skipping taps (the
// `continue`s kept one load in flight per thread; skipped and zero-weight terms give the same sum for finite gradients).
template <typename T>
__global__ void __launch_bounds__(256)
upcat_bwd_kernel(const T* __restrict__ dcat, const T* __restrict__ acc_skip, T* __restrict__ dskip, T* __restrict__ dx,
                 int B, int h, int w, int Cs, int Cu) {
  using V = Vec16<T>;
  using Raw = typename V::Raw;
  constexpr int VN = V::N;
  const int C = Cs + Cu, cvs = Cs / VN, cvu = Cu / VN, Ho = 2 * h, Wo = 2 * w;
  const unsigned npix = (unsigned)B * h * w;               // < 2^31 threads (checked by the launcher): 32-bit div/mod
  const unsigned n_skip = npix * cvs, total = n_skip + npix * cvu;
  unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const bool skip = idx < n_skip;
  if (!skip) idx -= n_skip;
  const unsigned cv = skip ? cvs : cvu;
  const int c = (int)(idx % cv) * VN;
  unsigned p = idx / cv;
  const int j = (int)(p % w); p /= w;
  const int i = (int)(p % h);
  const int b = (int)(p / h);
  if (skip) {            // plain slice (+ accumulation of the skip tensor's other gradient)
    Raw rv[4], ru[4];
    long long pix[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      pix[a] = ((long long)b * Ho + 2 * i + (a >> 1)) * Wo + 2 * j + (a & 1);
      rv[a] = V::ldraw(dcat + pix[a] * C + c);
      if (acc_skip) ru[a] = V::ldraw(acc_skip + pix[a] * Cs + c);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      V v, u;
      v.unpack(rv[a]);
      if (acc_skip) {
        u.unpack(ru[a]);
#pragma unroll
        for (int k = 0; k < VN; ++k) v.v[k] += u.v[k];
      }
      v.store(dskip + pix[a] * Cs + c);
    }
    return;
  }
  // the 4 + 4 separable weights of this low-resolution pixel; every product wy*wx is a multiple of 1/16 <= 1, i.e. exact
  // in bf16, so bf16 storage takes the mixed-precision FMA (f32 += bf16 * bf16) on the packed vectors as loaded.
  float wy[4], wx[4];
  int rr[4], qq[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = 2 * i - 1 + k, q = 2 * j - 1 + k;
    wy[k] = (r >= 0 && r < Ho) ? bil_weight_ac_false(r, i, h) : 0.f;
    wx[k] = (q >= 0 && q < Wo) ? bil_weight_ac_false(q, j, w) : 0.f;
    rr[k] = min(max(r, 0), Ho - 1);
    qq[k] = min(max(q, 0), Wo - 1);
  }
  const T* base = dcat + (long long)b * Ho * Wo * C + Cs + c;
  Raw raw[4][4];
#pragma unroll
  for (int kr = 0; kr < 4; ++kr)
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) raw[kr][kq] = V::ldraw(base + ((long long)rr[kr] * Wo + qq[kq]) * C);
  uint32_t dep = 0;
  if constexpr (sizeof(T) == 2) {
    const uint32_t zm = opaque_zero(B);
#pragma unroll
    for (int kr = 0; kr < 4; ++kr) dep |= all_loaded(raw[kr], zm);
  }
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] = 0.f;
#pragma unroll
  for (int kr = 0; kr < 4; ++kr)
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const float ww = wy[kr] * wx[kq];
      if constexpr (sizeof(T) == 2) {
        const uint4 u = raw[kr][kq];
        const uint32_t wp = pack_bf16x2(ww, ww) | dep;
        acc[0] = fma_bf16_lo(u.x, wp, acc[0]); acc[1] = fma_bf16_hi(u.x, wp, acc[1]);
        acc[2] = fma_bf16_lo(u.y, wp, acc[2]); acc[3] = fma_bf16_hi(u.y, wp, acc[3]);
        acc[4] = fma_bf16_lo(u.z, wp, acc[4]); acc[5] = fma_bf16_hi(u.z, wp, acc[5]);
        acc[6] = fma_bf16_lo(u.w, wp, acc[6]); acc[7] = fma_bf16_hi(u.w, wp, acc[7]);
      } else {
        V v;
        v.unpack(raw[kr][kq]);
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[k] = fmaf(ww, v.v[k], acc[k]);
      }
    }
  V o;
#pragma unroll
  for (int k = 0; k < VN; ++k) o.v[k] = acc[k];
  o.store(dx + (((long long)b * h + i) * w + j) * Cu + c);
}

// dout NCHW f32 [B,C,2h,2w] -> dlogits NHWC T [B,h,w,16] (channels >= C zero), align_corners=True
template <typename T>
__global__ void __launch_bounds__(256)
final_bwd_kernel(const float* __restrict__ dout, T* __restrict__ dl, int B, int h, int w, int C) {
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
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  const int r_lo = max(2 * i - 2, 0), r_hi = min(2 * i + 3, Ho - 1);
  const int q_lo = max(2 * j - 2, 0), q_hi = min(2 * j + 3, Wo - 1);
  for (int r = r_lo; r <= r_hi; ++r) {
    const float wy = bil_weight_ac_true(r, i, h, sch);
    if (wy == 0.f) continue;
    for (int q = q_lo; q <= q_hi; ++q) {
      const float wx = bil_weight_ac_true(q, j, w, scw);
      if (wx == 0.f) continue;
      const float ww = wy * wx;
      const float* dp = dout + ((long long)b * C * Ho + r) * Wo + q;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < C) acc[c] = fmaf(ww, __ldg(dp + (long long)c * Ho * Wo), acc[c]);
    }
  }
  T* op = dl + idx * 16;
#pragma unroll
  for (int c = 0; c < 16; ++c) op[c] = from_f32<T>(acc[c]);
}

// Tiled version of final_bwd_kernel (w even): a CTA owns an 8 x 32 tile of low-resolution pixels, stages the 20 x 72 window
// of dout it needs -- 4 class planes at a time -- in shared memory with coalesced 16-byte loads (each dout element is read
// once per CTA instead of ~9 times through L1 with stride-2 scalar loads), and every thread applies its 6 + 6 separable
// align-corners weights (computed once) to all classes: 281 -> ~60 us at B=32, 10 classes, 256 x 512.
constexpr int FB_TH = 8, FB_TW = 32, FB_R = 2 * FB_TH + 4, FB_C = 2 * FB_TW + 8, FB_CH = 4;
template <typename T>
__global__ void __launch_bounds__(256)
final_bwd_tiled_kernel(const float* __restrict__ dout, T* __restrict__ dl, int B, int h, int w, int C, int tiles_h, int tiles_w) {
  __shared__ __align__(16) float tile[FB_CH][FB_R][FB_C];
  const int Ho = 2 * h, Wo = 2 * w;
  int t = blockIdx.x;
  const int j0 = (t % tiles_w) * FB_TW; t /= tiles_w;
  const int i0 = (t % tiles_h) * FB_TH;
  const int b = t / tiles_h;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = i0 + ty, j = j0 + tx;
  const float sch = (Ho > 1) ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float scw = (Wo > 1) ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  float wy[6], wx[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int r = 2 * i - 2 + k, q = 2 * j - 2 + k;
    wy[k] = (i < h && r >= 0 && r < Ho) ? bil_weight_ac_true(r, i, h, sch) : 0.f;
    wx[k] = (j < w && q >= 0 && q < Wo) ? bil_weight_ac_true(q, j, w, scw) : 0.f;
  }
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  const int r0 = 2 * i0 - 2, q0 = 2 * j0 - 4;           // dout coordinates of tile[.][0][0]; q0 % 4 == 0
  const long long plane = (long long)Ho * Wo;
#pragma unroll
  for (int cbi = 0; cbi < 16 / FB_CH; ++cbi) {
    const int cb = cbi * FB_CH;               // compile-time after unrolling: acc[] stays in registers
    if (cb >= C) break;
    __syncthreads();
    for (int idx = threadIdx.x; idx < FB_CH * FB_R * (FB_C / 4); idx += 256) {
      const int v = idx % (FB_C / 4), rr = (idx / (FB_C / 4)) % FB_R, cc = idx / ((FB_C / 4) * FB_R);
      const int r = r0 + rr, q = q0 + 4 * v, c = cb + cc;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C && r >= 0 && r < Ho && q >= 0 && q < Wo)       // Wo % 4 == 0: a vector is entirely inside or outside
        val = __ldg(reinterpret_cast<const float4*>(dout + ((long long)b * C + c) * plane + (long long)r * Wo + q));
      *reinterpret_cast<float4*>(&tile[cc][rr][4 * v]) = val;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < FB_CH; ++cc) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float* tr = &tile[cc][2 * ty + k][2 * tx + 2];
        const float2 p0 = *reinterpret_cast<const float2*>(tr), p1 = *reinterpret_cast<const float2*>(tr + 2),
                     p2 = *reinterpret_cast<const float2*>(tr + 4);
        float rs = wx[0] * p0.x;
        rs = fmaf(wx[1], p0.y, rs); rs = fmaf(wx[2], p1.x, rs); rs = fmaf(wx[3], p1.y, rs);
        rs = fmaf(wx[4], p2.x, rs); rs = fmaf(wx[5], p2.y, rs);
        a = fmaf(wy[k], rs, a);
      }
      acc[cb + cc] = a;
    }
  }
  if (i < h && j < w) {
    T* op = dl + (((long long)b * h + i) * w + j) * 16;
    Vec16<T> o;
    constexpr int VN = Vec16<T>::N;
#pragma unroll
    for (int v = 0; v < 16 / VN; ++v) {
#pragma unroll
      for (int e = 0; e < VN; ++e) o.v[e] = acc[v * VN + e];
      o.store(op + v * VN);
    }
  }
}

// NCHW f32 [B,C,H,W] -> NHWC T [B,H,W,ldc] zero padded (plain UNet: gradient of nhwc_to_nchw)
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_pad_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int C, long long HW, int ldc) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int b = (int)(idx / HW);
  const long long pix = idx % HW;
  for (int c = 0; c < ldc; ++c)
    y[idx * ldc + c] = from_f32<T>(c < C ? x[((long long)b * C + c) * HW + pix] : 0.f);
}

// MaxPool2d(2) backward: the gradient goes to the first maximum of the window (PyTorch rule)
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ acc_in, T* __restrict__ dx,
                   int B, int H, int W, int C) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  const int Ho = H / 2, Wo = W / 2, cv = C / VN;
  const long long total = (long long)B * Ho * Wo * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cv) * VN;
  long long p = idx / cv;
  const int wo = (int)(p % Wo); p /= Wo;
  const int ho = (int)(p % Ho);
  const int b = (int)(p / Ho);
  const long long o00 = (((long long)b * H + 2 * ho) * W + 2 * wo) * C + c;
  const long long offs[4] = {o00, o00 + C, o00 + (long long)W * C, o00 + (long long)W * C + C};
  V xv[4], g, out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) xv[q].load(x + offs[q]);
  g.load(dy + (((long long)b * Ho + ho) * Wo + wo) * C + c);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (acc_in) out[q].load(acc_in + offs[q]);
    else {
#pragma unroll
      for (int k = 0; k < VN; ++k) out[q].v[k] = 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    int best = 0;
    float bv = xv[0].v[k];
#pragma unroll
    for (int q = 1; q < 4; ++q)
      if (xv[q].v[k] > bv) { bv = xv[q].v[k]; best = q; }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q == best) out[q].v[k] += g.v[k];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) out[q].store(dx + offs[q]);
}

}  // namespace b200

using namespace b200;
typedef __nv_bfloat16 bf16;

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16, name)                    \
  if ((dtype) == B200SEG_F32) { CALL_F32; }                            \
  else if ((dtype) == B200SEG_BF16) { CALL_BF16; }                     \
  else return set_error(-1, name ": bad dtype %d", (int)(dtype));

extern "C" {

int b200seg_bn_stats(const void* z, int dtype, long long P, int C, double* sum, double* sumsq, int nslot,
                     long long slot_stride, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0, "bn_stats: P=%lld C=%d (C must be a multiple of %d)", P, C, vn);
  B200_REQUIRE(nslot >= 1 && (nslot == 1 || slot_stride >= C), "bn_stats: nslot=%d slot_stride=%lld", nslot, slot_stride);
  const dim3 block = red_block(C / vn);
  const int ppb = red_ppb(P, block, C / vn);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (bn_stats_kernel<float><<<grid, block, 0, st>>>((const float*)z, P, C, ppb, sum, sumsq, nslot, slot_stride)),
             (bn_stats_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)z, P, C, ppb, sum, sumsq, nslot, slot_stride)), "bn_stats")
  return check_launch("bn_stats");
}

int b200seg_bn_finalize(const void* z, int dtype, const double* sum, const double* sumsq, int nslot,
                        long long slot_stride, long long n, const float* gamma,
                        const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                        float* mean, float* invstd, float* scale, float* shift, int C, b200seg_stream_t s) {
  B200_REQUIRE(C > 0 && n > 0 && z && nslot >= 1, "bn_finalize: C=%d n=%lld nslot=%d", C, n, nslot);
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (bn_finalize_kernel<float><<<cdiv(C, 128), 128, 0, st>>>((const float*)z, sum, sumsq, nslot, slot_stride, n, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale, shift, C)),
             (bn_finalize_kernel<bf16><<<cdiv(C, 128), 128, 0, st>>>((const bf16*)z, sum, sumsq, nslot, slot_stride, n, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale, shift, C)), "bn_finalize")
  return check_launch("bn_finalize");
}

// elementwise passes: same (channel-vector, pixel-lane) block shape as the reductions, 8 pixels per thread
static inline int ew_ppb(const dim3& block) { return 8 * (int)block.y; }

int b200seg_bn_apply(const void* z, const float* scale, const float* shift, const void* res, void* a, int dtype,
                     long long P, int C, int act, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0, "bn_apply: P=%lld C=%d", P, C);
  const dim3 block = red_block(C / vn);
  const int ppb = ew_ppb(block);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  const BnFin nofin = {};
#define BN_APPLY(T, R) bn_apply_kernel<T, R, false><<<grid, block, 0, st>>>((const T*)z, scale, shift, (const T*)res, (T*)a, P, C, ppb, act, nofin)
  DISPATCH_T(dtype, (res ? BN_APPLY(float, true) : BN_APPLY(float, false)), (res ? BN_APPLY(bf16, true) : BN_APPLY(bf16, false)), "bn_apply")
#undef BN_APPLY
  return check_launch("bn_apply");
}

// bn_finalize + bn_apply in one launch (see BnFin): st = slot sums of b200seg_bn_stats ([nslot][2][C]: sum | sumsq),
// sv = [4][C] out (mean, invstd, scale, shift) for the backward pass; running statistics updated in place (may be NULL).
int b200seg_bn_finalize_apply(const void* z, const double* st_sums, int nslot, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, float* sv, const void* res, void* a,
                              int dtype, long long P, int C, int act, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0 && nslot >= 1, "bn_finalize_apply: P=%lld C=%d nslot=%d", P, C, nslot);
  B200_REQUIRE(z && st_sums && gamma && beta && sv && a, "bn_finalize_apply: null pointer");
  const dim3 block = red_block(C / vn);
  const int ppb = fat_ppb(P, block, C / vn);          // fat blocks: each one repeats the finalize arithmetic
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  const BnFin f = {st_sums, st_sums + C, nslot, 2LL * C, P, gamma, beta, eps, momentum, running_mean, running_var,
                   sv, sv + C, sv + 2 * C, sv + 3 * C};
#define BN_APPLY(T, R) bn_apply_kernel<T, R, true><<<grid, block, 0, st>>>((const T*)z, nullptr, nullptr, (const T*)res, (T*)a, P, C, ppb, act, f)
  DISPATCH_T(dtype, (res ? BN_APPLY(float, true) : BN_APPLY(float, false)), (res ? BN_APPLY(bf16, true) : BN_APPLY(bf16, false)), "bn_finalize_apply")
#undef BN_APPLY
  return check_launch("bn_finalize_apply");
}

int b200seg_bn_bwd_reduce(const void* da, const void* z, const float* scale, const float* shift, const float* mean,
                          const float* invstd, int dtype, long long P, int C, int act, double* sg, double* sgx,
                          int nslot, long long slot_stride, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0, "bn_bwd_reduce: P=%lld C=%d", P, C);
  B200_REQUIRE(nslot >= 1 && (nslot == 1 || slot_stride >= C), "bn_bwd_reduce: nslot=%d slot_stride=%lld", nslot, slot_stride);
  const dim3 block = red_block(C / vn);
  const int ppb = red_ppb(P, block, C / vn);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  // B200SEG_BNR_VARIANT (tools only): 1 = 3 blocks/SM (80 registers, spills), 2 = 2 rows in flight at 4 blocks/SM.  Measured on
  // the B = 32 step: default 8.33 ms, variant 1 8.62, variant 2 8.43 -- four rows in flight at 2 blocks/SM stays.
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("B200SEG_BNR_VARIANT"); variant = e ? atoi(e) : 0; }
  if (dtype == B200SEG_BF16 && variant == 1) {
    bn_bwd_reduce_kernel<bf16, 4, 3><<<grid, block, 0, st>>>((const bf16*)da, (const bf16*)z, scale, shift, mean, invstd, P, C, ppb, act, sg, sgx, nslot, slot_stride);
    return check_launch("bn_bwd_reduce");
  }
  if (dtype == B200SEG_BF16 && variant == 2) {
    bn_bwd_reduce_kernel<bf16, 2, 4><<<grid, block, 0, st>>>((const bf16*)da, (const bf16*)z, scale, shift, mean, invstd, P, C, ppb, act, sg, sgx, nslot, slot_stride);
    return check_launch("bn_bwd_reduce");
  }
  DISPATCH_T(dtype, (bn_bwd_reduce_kernel<float><<<grid, block, 0, st>>>((const float*)da, (const float*)z, scale, shift, mean, invstd, P, C, ppb, act, sg, sgx, nslot, slot_stride)),
             (bn_bwd_reduce_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)da, (const bf16*)z, scale, shift, mean, invstd, P, C, ppb, act, sg, sgx, nslot, slot_stride)), "bn_bwd_reduce")
  return check_launch("bn_bwd_reduce");
}

int b200seg_bn_bwd_apply(const void* da, const void* z, const float* scale, const float* shift, const float* mean,
                         const float* invstd, const float* sg, const float* sgx, void* dz, int dtype, long long P,
                         int C, int act, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0, "bn_bwd_apply: P=%lld C=%d", P, C);
  const dim3 block = red_block(C / vn);
  const int ppb = ew_ppb(block);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  const float inv_n = 1.f / (float)P;
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (bn_bwd_apply_kernel<float, false><<<grid, block, 0, st>>>((const float*)da, (const float*)z, scale, shift, mean, invstd, sg, sgx, inv_n, (float*)dz, P, C, ppb, act, nullptr, nullptr, 0, 0)),
             (bn_bwd_apply_kernel<bf16, false><<<grid, block, 0, st>>>((const bf16*)da, (const bf16*)z, scale, shift, mean, invstd, sg, sgx, inv_n, (bf16*)dz, P, C, ppb, act, nullptr, nullptr, 0, 0)), "bn_bwd_apply")
  return check_launch("bn_bwd_apply");
}

// bn_bwd_apply reading the f64 slot sums of b200seg_bn_bwd_reduce directly (red = [nslot][2][C]: sum g | sum g*xhat);
// sv = [4][C] (mean, invstd, scale, shift) as written by the forward pass.
int b200seg_bn_bwd_apply_slots(const void* da, const void* z, const float* sv, const double* red, int nslot, void* dz,
                               int dtype, long long P, int C, int act, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0 && nslot >= 1, "bn_bwd_apply_slots: P=%lld C=%d nslot=%d", P, C, nslot);
  B200_REQUIRE(da && z && sv && red && dz, "bn_bwd_apply_slots: null pointer");
  const dim3 block = red_block(C / vn);
  const int ppb = fat_ppb(P, block, C / vn);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  const float inv_n = 1.f / (float)P;
  cudaStream_t st = (cudaStream_t)s;
  const float *mean = sv, *invstd = sv + C, *scale = sv + 2 * C, *shift = sv + 3 * C;
  DISPATCH_T(dtype, (bn_bwd_apply_kernel<float, true><<<grid, block, 0, st>>>((const float*)da, (const float*)z, scale, shift, mean, invstd, nullptr, nullptr, inv_n, (float*)dz, P, C, ppb, act, red, red + C, nslot, 2LL * C)),
             (bn_bwd_apply_kernel<bf16, true><<<grid, block, 0, st>>>((const bf16*)da, (const bf16*)z, scale, shift, mean, invstd, nullptr, nullptr, inv_n, (bf16*)dz, P, C, ppb, act, red, red + C, nslot, 2LL * C)), "bn_bwd_apply_slots")
  return check_launch("bn_bwd_apply_slots");
}

int b200seg_act_bwd(const void* da, const void* a_out, void* dz, int dtype, long long N, int act, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(N > 0 && N % vn == 0, "act_bwd: N=%lld", N);
  const unsigned g = cdiv(N / vn, 256);
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (act_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)da, (const float*)a_out, (float*)dz, N, act)),
             (act_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)da, (const bf16*)a_out, (bf16*)dz, N, act)), "act_bwd")
  return check_launch("act_bwd");
}

int b200seg_colsum(const void* x, int dtype, long long P, int C, double* out, int nslot, long long slot_stride,
                   b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(P > 0 && C > 0 && C % vn == 0, "colsum: P=%lld C=%d", P, C);
  B200_REQUIRE(nslot >= 1 && (nslot == 1 || slot_stride >= C), "colsum: nslot=%d slot_stride=%lld", nslot, slot_stride);
  const dim3 block = red_block(C / vn);
  const int ppb = red_ppb(P, block, C / vn);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (colsum_kernel<float><<<grid, block, 0, st>>>((const float*)x, P, C, ppb, out, nslot, slot_stride)),
             (colsum_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)x, P, C, ppb, out, nslot, slot_stride)), "colsum")
  return check_launch("colsum");
}

int b200seg_f64_to_f32(const double* in, float* out, int n, int nslot, long long slot_stride, float scale,
                       b200seg_stream_t s) {
  B200_REQUIRE(n > 0 && nslot >= 1 && (nslot == 1 || slot_stride >= n), "f64_to_f32: n=%d nslot=%d", n, nslot);
  f64_to_f32_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)s>>>(in, out, n, nslot, slot_stride, scale);
  return check_launch("f64_to_f32");
}

int b200seg_conv_wgrad(const void* x, const void* dz, float* dw, int dtype, int B, int H, int W, int Cin, int Cout,
                       int taps, b200seg_stream_t s) {
  B200_REQUIRE(taps == 1 || taps == 9, "conv_wgrad: taps=%d", taps);
  B200_REQUIRE(Cin > 0 && Cin % 4 == 0 && Cout > 0, "conv_wgrad: Cin=%d must be a multiple of 4", Cin);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "conv_wgrad: empty tensor");
  const long long P = (long long)B * H * W;
  const int K = taps * Cin;
  const unsigned gx = cdiv(K, 64), gy = cdiv(Cout, 64);
  // enough pixel splits for ~4 waves of CTAs, at least 256 pixels each
  long long splits = ((long long)sm_count() * 8 + (long long)gx * gy - 1) / ((long long)gx * gy);
  if (splits > P / 256) splits = P / 256;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long pps = (P + splits - 1) / splits;
  pps = (pps + 15) / 16 * 16;
  splits = (P + pps - 1) / pps;
  dim3 grid(gx, gy, (unsigned)splits);
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (conv_wgrad_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)dz, dw, B, H, W, Cin, Cout, taps, pps)),
             (conv_wgrad_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dz, dw, B, H, W, Cin, Cout, taps, pps)), "conv_wgrad")
  return check_launch("conv_wgrad");
}

int b200seg_dw_dgrad(const void* dz, const float* w, const void* acc_in, void* dx, int dtype, int B, int H, int W,
                     int C, int stride, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(C > 0 && C % vn == 0 && (stride == 1 || stride == 2), "dw_dgrad: C=%d stride=%d", C, stride);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "dw_dgrad: empty tensor");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long P = (long long)B * H * W;
  const dim3 block = red_block_pow2(C / vn);
  const int ppb = 8 * (int)block.y;
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == B200SEG_BF16 && stride == 2) {
    const dim3 blk = red_block(C / vn);
    const long long NB = (long long)B * Ho * Wo;
    const int bpb = 4 * (int)blk.y;
    dim3 g2(cdiv(NB, bpb), cdiv(C / vn, blk.x));
    dw_dgrad_s2_bf16_kernel<<<g2, blk, 0, st>>>((const bf16*)dz, w, (const bf16*)acc_in, (bf16*)dx, B, H, W, C, Ho, Wo, bpb);
    return check_launch("dw_dgrad");
  }
  DISPATCH_T(dtype, (dw_dgrad_kernel<float><<<grid, block, 0, st>>>((const float*)dz, w, (const float*)acc_in, (float*)dx, B, H, W, C, Ho, Wo, stride, ppb)),
             (dw_dgrad_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)dz, w, (const bf16*)acc_in, (bf16*)dx, B, H, W, C, Ho, Wo, stride, ppb)), "dw_dgrad")
  return check_launch("dw_dgrad");
}

int b200seg_dw_wgrad(const void* x, const void* dz, double* dw, int nslot, int dtype, int B, int H, int W, int C,
                     int stride, b200seg_stream_t s) {
  B200_REQUIRE(nslot >= 1, "dw_wgrad: nslot=%d", nslot);
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(C > 0 && C % vn == 0 && (stride == 1 || stride == 2), "dw_wgrad: C=%d stride=%d", C, stride);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "dw_wgrad: empty tensor");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long P = (long long)B * Ho * Wo;
  const dim3 block = red_block(C / vn);
  const int ppb = red_ppb(P, block, C / vn);
  dim3 grid(cdiv(P, ppb), cdiv(C / vn, block.x));
  cudaStream_t st = (cudaStream_t)s;
  static int variant = -1;      // B200SEG_DW_WGRAD: 0 = generic kernel, 2 / 4 = pixels per thread of the bf16 kernel (default 4)
  if (variant < 0) { const char* e = getenv("B200SEG_DW_WGRAD"); variant = e ? atoi(e) : 4; }
  const bool small_index = (long long)B * H * W * C < (1LL << 31);
  if (dtype == B200SEG_BF16 && variant > 0 && small_index) {
    const int pw = variant == 4 ? 4 : 2;
    const long long G = (long long)B * Ho * ((Wo + pw - 1) / pw);
    static int bps = -1;          // blocks per SM (one is resident at a time; more = shorter tail, more atomics)
    if (bps < 0) { const char* e = getenv("B200SEG_DWW_BPS"); bps = e ? max(1, atoi(e)) : 1; }
    // L2-resident layers (the 13 deep ones): blocks of 2 channel vectors (one 32-byte sector per pixel) x 128 pixel lanes,
    // the channel axis on grid.y -- a block then emits 9*16 atomics instead of 9*C.  With full-width blocks the f64
    // atomics (148 blocks x 9 x 384 channels = 511 k, ~20 per ns chip-wide) were 25 of the 33 us of a 384-channel layer.
    static long long narrow_bytes = -1;
    if (narrow_bytes < 0) { const char* e = getenv("B200SEG_DWW_NARROW_MB"); narrow_bytes = (e ? atoll(e) : 48) << 20; }
    dim3 blk = block;
    if ((long long)B * H * W * C * 2 <= narrow_bytes && (C / vn) % 2 == 0 && C / vn > 2) blk = dim3(2, 128);
    const long long gy = cdiv(C / vn, blk.x);
    long long gpb = cdiv(G * gy, (long long)sm_count() * bps);
    if (gpb < (long long)blk.y) gpb = blk.y;
    dim3 g2(cdiv(G, gpb), (unsigned)gy);
    const size_t smem = (size_t)blk.y * 9 * blk.x * 8 * sizeof(float);
#define DWW(S, PW)                                                                                                     \
  {                                                                                                                    \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      cudaFuncSetAttribute(dw_wgrad_bf16_kernel<S, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 72 * 4);    \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    dw_wgrad_bf16_kernel<S, PW><<<g2, blk, smem, st>>>((const bf16*)x, (const bf16*)dz, B, H, W, C, Ho, Wo, (int)gpb, dw, nslot); \
  }
    if (stride == 1) { if (pw == 4) DWW(1, 4) else DWW(1, 2) }
    else { if (pw == 4) DWW(2, 4) else DWW(2, 2) }
#undef DWW
    return check_launch("dw_wgrad");
  }
  DISPATCH_T(dtype, (dw_wgrad_kernel<float><<<grid, block, 0, st>>>((const float*)x, (const float*)dz, B, H, W, C, Ho, Wo, stride, ppb, dw, nslot)),
             (dw_wgrad_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)x, (const bf16*)dz, B, H, W, C, Ho, Wo, stride, ppb, dw, nslot)), "dw_wgrad")
  return check_launch("dw_wgrad");
}

int b200seg_smallcin_wgrad(const void* x, int x_dtype, const void* dz, int dtype, float* dw, int B, int Cin, int H,
                           int W, int Cout, int stride, b200seg_stream_t s) {
  B200_REQUIRE(Cin >= 1 && Cin <= 4 && Cout > 0 && Cout % 4 == 0 && 9 * Cin * Cout <= 2048, "smallcin_wgrad: Cin=%d Cout=%d (Cout % 4 == 0, 9*Cin*Cout <= 2048)", Cin, Cout);
  B200_REQUIRE(stride == 1 || stride == 2, "smallcin_wgrad: stride=%d", stride);
  B200_REQUIRE(B > 0 && H > 0 && W > 0, "smallcin_wgrad: empty tensor");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long P = (long long)B * Ho * Wo;
  cudaStream_t st0 = (cudaStream_t)s;
  if (Cin == 3 && (Cout == 32 || Cout == 64) && dtype == B200SEG_BF16 && (long long)3 * H * W < (1LL << 30)) {
    // the stems of the two models in the training precision of the path: software-pipelined kernel
    const long long jobs = (long long)B * Ho * ((Wo + 127) / 128);
    B200_REQUIRE(jobs < (1LL << 31), "smallcin_wgrad: too many row segments");
    long long gb = (long long)sm_count() * 2;
    if (gb > jobs) gb = jobs;
    const int ncol = 127 * stride + 3;
    size_t smem_t = (size_t)(((9 * (ncol + 3) + 3) & ~3) + 128 * Cout) * sizeof(float);
    const size_t red_b = (size_t)8 * 28 * Cout * sizeof(float);
    if (smem_t < red_b) smem_t = red_b;
#define LAUNCH_S(TI, SS, CO)                                                                                             \
  {                                                                                                                     \
    static bool attr_set = false;                                                                                       \
    if (!attr_set) {                                                                                                    \
      cudaFuncSetAttribute(stem_wgrad_bf16_kernel<TI, SS, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
      attr_set = true;                                                                                                  \
    }                                                                                                                   \
    stem_wgrad_bf16_kernel<TI, SS, CO><<<(unsigned)gb, 256, smem_t, st0>>>((const TI*)x, (const bf16*)dz, dw, B, H, W, Ho, Wo); \
  }
#define LAUNCH_SC(TI)                                                                                                    \
  {                                                                                                                     \
    if (stride == 2) { if (Cout == 32) LAUNCH_S(TI, 2, 32) else LAUNCH_S(TI, 2, 64) }                                   \
    else { if (Cout == 32) LAUNCH_S(TI, 1, 32) else LAUNCH_S(TI, 1, 64) }                                               \
  }
    if (x_dtype == B200SEG_F32) LAUNCH_SC(float)
    else if (x_dtype == B200SEG_BF16) LAUNCH_SC(bf16)
    else return set_error(-1, "smallcin_wgrad: bad dtypes");
#undef LAUNCH_SC
#undef LAUNCH_S
    return check_launch("smallcin_wgrad");
  }
  if (Cin == 3 && Cout <= 64) {                  // other stems / fp32 training: row-tile kernel
    const long long jobs = (long long)B * Ho * ((Wo + 127) / 128);
    long long gb = (long long)sm_count() * 4;
    if (gb > jobs) gb = jobs;
    const int ncol = 127 * stride + 3;
    size_t smem_t = (size_t)(((9 * (ncol + 3) + 3) & ~3) + 128 * Cout) * sizeof(float);
    const size_t red_b = (size_t)4 * 28 * Cout * sizeof(float);
    if (smem_t < red_b) smem_t = red_b;
#define LAUNCH_T(TI, T)                                                                                                   \
  {                                                                                                                     \
    if (stride == 2) smallcin3_wgrad_tile_kernel<TI, T, 2><<<(unsigned)gb, 256, smem_t, st0>>>((const TI*)x, (const T*)dz, dw, B, H, W, Cout, Ho, Wo); \
    else smallcin3_wgrad_tile_kernel<TI, T, 1><<<(unsigned)gb, 256, smem_t, st0>>>((const TI*)x, (const T*)dz, dw, B, H, W, Cout, Ho, Wo);             \
  }
    if (x_dtype == B200SEG_F32 && dtype == B200SEG_F32) LAUNCH_T(float, float)
    else if (x_dtype == B200SEG_F32 && dtype == B200SEG_BF16) LAUNCH_T(float, bf16)
    else if (x_dtype == B200SEG_BF16 && dtype == B200SEG_BF16) LAUNCH_T(bf16, bf16)
    else if (x_dtype == B200SEG_BF16 && dtype == B200SEG_F32) LAUNCH_T(bf16, float)
    else return set_error(-1, "smallcin_wgrad: bad dtypes");
#undef LAUNCH_T
    return check_launch("smallcin_wgrad");
  }
  long long blocks = (long long)sm_count() * 4;
  long long ppb = (P + blocks - 1) / blocks;
  ppb = (ppb + 31) / 32 * 32;
  blocks = (P + ppb - 1) / ppb;
  const size_t smem = (size_t)32 * (Cout + 9 * Cin) * sizeof(float);
  cudaStream_t st = (cudaStream_t)s;
#define LAUNCH(TI, T) smallcin_wgrad_kernel<TI, T><<<(unsigned)blocks, 256, smem, st>>>((const TI*)x, (const T*)dz, dw, B, Cin, H, W, Cout, Ho, Wo, stride, (int)ppb)
  if (x_dtype == B200SEG_F32 && dtype == B200SEG_F32) LAUNCH(float, float);
  else if (x_dtype == B200SEG_F32 && dtype == B200SEG_BF16) LAUNCH(float, bf16);
  else if (x_dtype == B200SEG_BF16 && dtype == B200SEG_BF16) LAUNCH(bf16, bf16);
  else if (x_dtype == B200SEG_BF16 && dtype == B200SEG_F32) LAUNCH(bf16, float);
  else return set_error(-1, "smallcin_wgrad: bad dtypes");
#undef LAUNCH
  return check_launch("smallcin_wgrad");
}

int b200seg_upcat_bwd(const void* dcat, const void* acc_skip, void* dskip, void* dx, int dtype, int B, int h, int w,
                      int Cs, int Cu, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(Cs % vn == 0 && Cu % vn == 0 && Cu > 0 && Cs >= 0, "upcat_bwd: Cs=%d Cu=%d", Cs, Cu);
  B200_REQUIRE(B > 0 && h > 0 && w > 0, "upcat_bwd: empty tensor");
  B200_REQUIRE((long long)B * h * w * ((Cs + Cu) / vn) < (1LL << 31), "upcat_bwd: tensor too large for 32-bit thread indices");
  const unsigned g = cdiv((long long)B * h * w * ((Cs + Cu) / vn), 256);
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (upcat_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)dcat, (const float*)acc_skip, (float*)dskip, (float*)dx, B, h, w, Cs, Cu)),
             (upcat_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)dcat, (const bf16*)acc_skip, (bf16*)dskip, (bf16*)dx, B, h, w, Cs, Cu)), "upcat_bwd")
  return check_launch("upcat_bwd");
}

int b200seg_final_bwd(const float* dout, void* dlogits, int dtype, int B, int h, int w, int C, b200seg_stream_t s) {
  B200_REQUIRE(C >= 1 && C <= 16 && B > 0 && h > 0 && w > 0, "final_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)s;
  if (w % 2 == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0) {
    const int th = (h + FB_TH - 1) / FB_TH, tw = (w + FB_TW - 1) / FB_TW;
    const unsigned gt = (unsigned)((long long)B * th * tw);
    DISPATCH_T(dtype, (final_bwd_tiled_kernel<float><<<gt, 256, 0, st>>>(dout, (float*)dlogits, B, h, w, C, th, tw)),
               (final_bwd_tiled_kernel<bf16><<<gt, 256, 0, st>>>(dout, (bf16*)dlogits, B, h, w, C, th, tw)), "final_bwd")
    return check_launch("final_bwd");
  }
  const unsigned g = cdiv((long long)B * h * w, 256);
  DISPATCH_T(dtype, (final_bwd_kernel<float><<<g, 256, 0, st>>>(dout, (float*)dlogits, B, h, w, C)),
             (final_bwd_kernel<bf16><<<g, 256, 0, st>>>(dout, (bf16*)dlogits, B, h, w, C)), "final_bwd")
  return check_launch("final_bwd");
}

int b200seg_nchw_to_nhwc_pad(const float* x, void* y, int dtype, int B, int C, int H, int W, int ldc,
                             b200seg_stream_t s) {
  B200_REQUIRE(C >= 1 && ldc >= C && B > 0 && H > 0 && W > 0, "nchw_to_nhwc_pad: bad shape");
  const long long HW = (long long)H * W;
  const unsigned g = cdiv((long long)B * HW, 256);
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (nchw_to_nhwc_pad_kernel<float><<<g, 256, 0, st>>>(x, (float*)y, B, C, HW, ldc)),
             (nchw_to_nhwc_pad_kernel<bf16><<<g, 256, 0, st>>>(x, (bf16*)y, B, C, HW, ldc)), "nchw_to_nhwc_pad")
  return check_launch("nchw_to_nhwc_pad");
}

int b200seg_maxpool_bwd(const void* x, const void* dy, const void* acc_in, void* dx, int dtype, int B, int H, int W,
                        int C, b200seg_stream_t s) {
  const int vn = dtype == B200SEG_BF16 ? 8 : 4;
  B200_REQUIRE(C % vn == 0 && H % 2 == 0 && W % 2 == 0 && B > 0 && H > 0 && W > 0, "maxpool_bwd: bad shape");
  const unsigned g = cdiv((long long)B * (H / 2) * (W / 2) * (C / vn), 256);
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, (maxpool_bwd_kernel<float><<<g, 256, 0, st>>>((const float*)x, (const float*)dy, (const float*)acc_in, (float*)dx, B, H, W, C)),
             (maxpool_bwd_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, (const bf16*)acc_in, (bf16*)dx, B, H, W, C)), "maxpool_bwd")
  return check_launch("maxpool_bwd");
}

}  // extern "C"
