// bn_cluster.cu -- train-mode BatchNorm forward / backward of the small and mid-size layers as ONE launch per direction,
// synchronised through thread-block clusters instead of kernel boundaries (nn.BatchNorm2d in model.train(): train.py:24,36;
// its autograd backward: train.py:38).
//
// Why: BatchNorm statistics are per channel.  The grid-wide versions (train_ops.cu: statistics -> finalize -> apply, and
// reduce -> slot sums -> apply) need every pixel of a channel before anything can be normalised, hence three launches per
// direction with f64 atomics into 16 slot copies and a tree reduction per block.  On the ~45 layers whose tensors are a few
// MB the fixed cost of those launches dominates: 24 us forward and 30 us backward per layer for 3-8 us worth of bytes
// (profiles/r02_train_step_kernel_profile.txt).  Here a CLUSTER of up to 16 CTAs owns a slice of 16 channels (32 bytes of every
// NHWC pixel): the CTAs split the pixels, reduce through distributed shared memory (one hardware cluster barrier), and every
// CTA finishes the per-channel constants itself -- no cross-cluster communication at all, because no other cluster touches
// these channels.  The second pass over the slice (apply) re-reads it from L2, where a tensor of this size still lives.
//
//   forward : sum(z-k), sum((z-k)^2) -> cluster -> mean, invstd, scale, shift (+ running stats) -> a = act(z*scale+shift) (+res)
//   backward: sum(g), sum(g*xhat)     -> cluster -> dz = scale * (g - mean(g) - xhat * mean(g*xhat)); sums -> gradient staging
//
// Thread = one 16-byte channel vector (8 bf16) of the slice x a strided set of pixels: a warp load covers 16 pixels x 32 B
// (whole 32-byte sectors).  Partial sums: f32 per thread over <= a few hundred pixels, warp shuffles, then f64 across warps
// and across the cluster (same order in every CTA, so all CTAs of a cluster derive bit-identical constants).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {

namespace {

constexpr int BC_THREADS = 256;
constexpr int BC_SLICE = 16;        // channels per cluster
constexpr int BC_LANES = BC_THREADS / 2;   // pixel lanes per CTA (two 8-channel vectors per pixel)

__device__ __forceinline__ float act_grad_bc(float u, int act) {
  if (act == B200SEG_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == B200SEG_ACT_RELU6) return (u > 0.f && u < 6.f) ? 1.f : 0.f;
  return 1.f;
}

// Sum of v[0..NV) over the CTA's pixel lanes and then over the cluster, in f64.  `part` is this CTA's shared staging
// [8 warps][2 vectors][NV]; `tot` its shared totals [2][NV] (read by the other CTAs of the cluster through DSMEM); `fin` the
// cluster totals [2][NV].
// Returns in out[0..NV) the cluster totals of this thread's vector (tx).  Every thread of every CTA gets the same numbers.
template <int NV>
__device__ __forceinline__ void cluster_sum(const float (&v)[NV], double (&out)[NV], float* part, double* tot, double* fin, int tx) {
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float r[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = v[i];
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);     // lanes with the same parity = same vector
    r[i] = s;
  }
  if (lane < 2) {
#pragma unroll
    for (int i = 0; i < NV; ++i) part[(warp * 2 + lane) * NV + i] = r[i];
  }
  __syncthreads();
  if (threadIdx.x < 2 * NV) {
    const int vec = threadIdx.x / NV, i = threadIdx.x - vec * NV;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < BC_THREADS / 32; ++w) s += (double)part[(w * 2 + vec) * NV + i];
    tot[vec * NV + i] = s;
  }
  cluster.sync();                                    // every CTA's totals are visible cluster-wide
  // one thread per (vector, value) gathers the peers' totals through distributed shared memory (2*NV x cluster-size remote
  // loads per CTA -- every thread reading every peer was 65 k remote loads per CTA and dominated the kernel), in rank order so
  // that all CTAs of the cluster derive bit-identical sums; the result is broadcast through local shared memory
  if (threadIdx.x < 2 * NV) {
    const unsigned nr = cluster.num_blocks();
    double s = 0.0;
    for (unsigned rk = 0; rk < nr; ++rk) s += cluster.map_shared_rank(tot, rk)[threadIdx.x];
    fin[threadIdx.x] = s;
  }
  cluster.sync();                                    // remote reads done (peers may overwrite tot / exit), fin visible
#pragma unroll
  for (int i = 0; i < NV; ++i) out[i] = fin[tx * NV + i];
}

__device__ __forceinline__ void bc_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void bc_cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bc_unpack(const uint4 u, float (&v)[8]) {
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
  v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
}
constexpr int BC_RESIDENT_BYTES = 72 * 1024;      // per-CTA slice kept in shared memory between the two passes (3 CTAs per SM)

struct BnFwdArgs {
  const __nv_bfloat16* z; const __nv_bfloat16* res; __nv_bfloat16* a;
  const float* gamma; const float* beta; float* running_mean; float* running_var; float* sv;   // sv: [4][C]
  long long P; int C; int act; float eps, momentum;
};

// RESIDENT: the CTA's part of the slice (<= 72 KB) is copied to shared memory ONCE with cp.async -- every load of the pass in
// flight at the same time, no registers tied up -- and both passes read it from there: one trip to L2/HBM instead of two.
template <bool RESIDENT>
__global__ void __launch_bounds__(BC_THREADS)
bn_fwd_cluster_kernel(const BnFwdArgs g) {
  extern __shared__ __align__(16) uint8_t dyn[];
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float part[(BC_THREADS / 32) * 2 * 16];
  __shared__ double tot[2 * 16], fin[2 * 16];
  const int slice = blockIdx.x / cluster.num_blocks();
  const unsigned rank = cluster.block_rank(), nr = cluster.num_blocks();
  const int tx = threadIdx.x & 1, ty = threadIdx.x >> 1;
  const int c0 = slice * BC_SLICE + tx * 8;
  const long long p0 = g.P * rank / nr, p1 = g.P * (rank + 1) / nr;
  const int C = g.C;
  // shift k = z at pixel 0 (see bn_stats_kernel: plain sum(z^2) cancels catastrophically when |mean| >> sigma)
  float k[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(g.z + c0));
    k[0] = bf16lo(u.x); k[1] = bf16hi(u.x); k[2] = bf16lo(u.y); k[3] = bf16hi(u.y);
    k[4] = bf16lo(u.z); k[5] = bf16hi(u.z); k[6] = bf16lo(u.w); k[7] = bf16hi(u.w);
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (RESIDENT) {
    for (long long p = p0 + ty; p < p1; p += BC_LANES) bc_cp_async16(dyn + (p - p0) * 32 + tx * 16, g.z + p * C + c0);
    bc_cp_async_wait_all();                           // a thread only ever reads back the 16 bytes it copied itself
  }
#pragma unroll 4
  for (long long p = p0 + ty; p < p1; p += BC_LANES) {
    const uint4 u = RESIDENT ? *reinterpret_cast<const uint4*>(dyn + (p - p0) * 32 + tx * 16)
                             : __ldg(reinterpret_cast<const uint4*>(g.z + p * C + c0));
    float v[8];
    bc_unpack(u, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - k[j]; acc[j] += d; acc[8 + j] = fmaf(d, d, acc[8 + j]); }
  }
  double s[16];
  cluster_sum<16>(acc, s, part, tot, fin, tx);
  // per-channel constants (bn_finalize_kernel's arithmetic), redundantly in every thread; rank 0 publishes them
  float sc[8], sh[8];
  const double n = (double)g.P;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double ms = s[j] / n;
    double var = s[8 + j] / n - ms * ms;
    if (var < 0.0) var = 0.0;
    const double m = ms + (double)k[j];
    const float invstd = (float)(1.0 / sqrt(var + (double)g.eps));
    sc[j] = __ldg(g.gamma + c0 + j) * invstd;
    sh[j] = __ldg(g.beta + c0 + j) - (float)m * sc[j];
    if (rank == 0 && ty == 0) {
      g.sv[c0 + j] = (float)m;
      g.sv[C + c0 + j] = invstd;
      g.sv[2 * C + c0 + j] = sc[j];
      g.sv[3 * C + c0 + j] = sh[j];
      if (g.running_mean) {
        const double unbiased = g.P > 1 ? var * (n / (n - 1.0)) : var;
        g.running_mean[c0 + j] = (1.f - g.momentum) * g.running_mean[c0 + j] + g.momentum * (float)m;
        g.running_var[c0 + j] = (1.f - g.momentum) * g.running_var[c0 + j] + g.momentum * (float)unbiased;
      }
    }
  }
  const float lo = g.act != B200SEG_ACT_NONE ? 0.f : -INFINITY, hi = g.act == B200SEG_ACT_RELU6 ? 6.f : INFINITY;
#pragma unroll 4
  for (long long p = p0 + ty; p < p1; p += BC_LANES) {
    const long long off = p * C + c0;
    const uint4 u = RESIDENT ? *reinterpret_cast<const uint4*>(dyn + (p - p0) * 32 + tx * 16)
                             : __ldg(reinterpret_cast<const uint4*>(g.z + off));
    float v[8];
    bc_unpack(u, v);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fminf(fmaxf(fmaf(v[j], sc[j], sh[j]), lo), hi);
    if (g.res) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(g.res + off));
      o[0] += bf16lo(r.x); o[1] += bf16hi(r.x); o[2] += bf16lo(r.y); o[3] += bf16hi(r.y);
      o[4] += bf16lo(r.z); o[5] += bf16hi(r.z); o[6] += bf16lo(r.w); o[7] += bf16hi(r.w);
    }
    *reinterpret_cast<uint4*>(g.a + off) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
}

struct BnBwdArgs {
  const __nv_bfloat16* da; const __nv_bfloat16* z; __nv_bfloat16* dz;
  const float* sv;          // [4][C]: mean, invstd, scale, shift
  double* red;              // gradient staging [nslot][2][C] (zeroed): slot 0 receives sum g (d beta) and sum g*xhat (d gamma)
  long long P; int C; int act;
};

template <bool RESIDENT>
__global__ void __launch_bounds__(BC_THREADS)
bn_bwd_cluster_kernel(const BnBwdArgs g) {
  extern __shared__ __align__(16) uint8_t dyn[];    // RESIDENT: [pixels][da 32 B | z 32 B]
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float part[(BC_THREADS / 32) * 2 * 16];
  __shared__ double tot[2 * 16], fin[2 * 16];
  const int slice = blockIdx.x / cluster.num_blocks();
  const unsigned rank = cluster.block_rank(), nr = cluster.num_blocks();
  const int tx = threadIdx.x & 1, ty = threadIdx.x >> 1;
  const int C = g.C;
  const int c0 = slice * BC_SLICE + tx * 8;
  const long long p0 = g.P * rank / nr, p1 = g.P * (rank + 1) / nr;
  float ksc[8], ksh[8], kmu[8], kis[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    kmu[j] = __ldg(g.sv + c0 + j); kis[j] = __ldg(g.sv + C + c0 + j);
    ksc[j] = __ldg(g.sv + 2 * C + c0 + j); ksh[j] = __ldg(g.sv + 3 * C + c0 + j);
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  if (RESIDENT) {
    for (long long p = p0 + ty; p < p1; p += BC_LANES) {
      bc_cp_async16(dyn + (p - p0) * 64 + tx * 16, g.da + p * C + c0);
      bc_cp_async16(dyn + (p - p0) * 64 + 32 + tx * 16, g.z + p * C + c0);
    }
    bc_cp_async_wait_all();
  }
#pragma unroll 2
  for (long long p = p0 + ty; p < p1; p += BC_LANES) {
    const long long off = p * C + c0;
    const uint4 ud = RESIDENT ? *reinterpret_cast<const uint4*>(dyn + (p - p0) * 64 + tx * 16) : __ldg(reinterpret_cast<const uint4*>(g.da + off));
    const uint4 uz = RESIDENT ? *reinterpret_cast<const uint4*>(dyn + (p - p0) * 64 + 32 + tx * 16) : __ldg(reinterpret_cast<const uint4*>(g.z + off));
    float d[8], v[8];
    bc_unpack(ud, d);
    bc_unpack(uz, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = fmaf(v[j], ksc[j], ksh[j]);
      const float gg = d[j] * act_grad_bc(u, g.act);
      acc[j] += gg;
      acc[8 + j] = fmaf(gg, (v[j] - kmu[j]) * kis[j], acc[8 + j]);
    }
  }
  double s[16];
  cluster_sum<16>(acc, s, part, tot, fin, tx);
  float kmg[8], kmx[8];
  const float inv_n = 1.f / (float)g.P;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // the three-launch path hands the sums to the apply pass as f32 (f64_to_f32_kernel) and multiplies by 1/P there
    kmg[j] = (float)s[j] * inv_n;
    kmx[j] = (float)s[8 + j] * inv_n;
    if (rank == 0 && ty == 0) { g.red[c0 + j] = s[j]; g.red[C + c0 + j] = s[8 + j]; }
  }
#pragma unroll 2
  for (long long p = p0 + ty; p < p1; p += BC_LANES) {
    const long long off = p * C + c0;
    const uint4 ud = RESIDENT ? *reinterpret_cast<const uint4*>(dyn + (p - p0) * 64 + tx * 16) : __ldg(reinterpret_cast<const uint4*>(g.da + off));
    const uint4 uz = RESIDENT ? *reinterpret_cast<const uint4*>(dyn + (p - p0) * 64 + 32 + tx * 16) : __ldg(reinterpret_cast<const uint4*>(g.z + off));
    float d[8], v[8];
    bc_unpack(ud, d);
    bc_unpack(uz, v);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = fmaf(v[j], ksc[j], ksh[j]);
      const float gg = d[j] * act_grad_bc(u, g.act);
      const float xh = (v[j] - kmu[j]) * kis[j];
      o[j] = ksc[j] * (gg - kmg[j] - xh * kmx[j]);
    }
    *reinterpret_cast<uint4*>(g.dz + off) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
}

// cluster size for a layer: enough CTAs to fill the machine, at least two pixels per lane and CTA
int pick_cluster(long long P, int C) {
  const int slices = C / BC_SLICE;
  int cs = 1;
  while (cs < 16 && (long long)slices * cs < (long long)sm_count() && P / (cs * 2) >= 2 * BC_LANES) cs <<= 1;
  return cs;
}

// Cluster size and residency of a layer with `px_bytes` bytes per pixel and slice kept between the passes (32 forward: z; 64
// backward: da + z): grow the cluster until a CTA's part fits in BC_RESIDENT_BYTES (a few waves of CTAs are fine, the
// clusters are independent).
void pick_layout(long long P, int C, int px_bytes, int* cs, bool* resident) {
  int c = pick_cluster(P, C);
  auto bytes = [&](int cc) { return ((P + cc - 1) / cc) * px_bytes; };
  while (c < 16 && bytes(c) > BC_RESIDENT_BYTES && P / (c * 2) >= 2 * BC_LANES) c <<= 1;
  *cs = c;
  *resident = bytes(c) <= BC_RESIDENT_BYTES;
}

template <typename Args>
int launch_cluster(void (*kern_res)(const Args), void (*kern_stream)(const Args), const Args& a, long long P, int C, int px_bytes,
                   cudaStream_t st, const char* what) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern_res, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_stream, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern_res, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_RESIDENT_BYTES);
    if (e != cudaSuccess) return set_error((int)e, "%s: kernel attributes: %s", what, cudaGetErrorString(e));
    attr = true;
  }
  int cs = 1;
  bool resident = false;
  pick_layout(P, C, px_bytes, &cs, &resident);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((C / BC_SLICE) * cs));
  cfg.blockDim = dim3(BC_THREADS);
  cfg.dynamicSmemBytes = resident ? (size_t)(((P + cs - 1) / cs) * px_bytes) : 0;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, resident ? kern_res : kern_stream, a);
  if (e != cudaSuccess) return set_error((int)e, "%s: cluster launch (%d CTAs x cluster %d): %s", what, (C / BC_SLICE) * cs, cs, cudaGetErrorString(e));
  return check_launch(what);
}

}  // namespace

}  // namespace b200

using namespace b200;

// 1 when the cluster kernels take this layer: bf16, 16-channel slices, enough slices x cluster size to fill the machine
// (C >= 128: a 16- or 64-channel layer would stream through 16 or 64 CTAs), and a tensor small enough that the second pass
// over it is served by L2 (larger layers stream better through the grid-wide kernels).  Measured per layer at B = 32
// (tools/train_profile.py, forward / backward us): 384 ch @ 16x32 24 -> 17 / 32 -> 27, 576 ch 26 -> 28 / 50 -> 32,
// 960 ch @ 8x16 23 -> 13 / 33 -> 15.
extern "C" int b200seg_bn_cluster_supported(int dtype, long long P, int C) {
  static long long limit = -1;
  static int min_ctas = -1;
  if (limit < 0) { const char* e = getenv("B200SEG_BN_CLUSTER_MB"); limit = (e ? atoll(e) : 24) << 20; }
  if (min_ctas < 0) { const char* e = getenv("B200SEG_BN_CLUSTER_MIN_CTAS"); min_ctas = e ? atoi(e) : 120; }
  if (!(dtype == B200SEG_BF16 && C % BC_SLICE == 0 && P >= 2 * BC_LANES && P * C * 2 <= limit)) return 0;
  return (C / BC_SLICE) * pick_cluster(P, C) >= min_ctas;
}
// backward additionally wants its (da, z) slice resident in shared memory: streaming it twice was no faster than the
// grid-wide kernels (48 vs 46 us on 192 channels @ 32x64)
extern "C" int b200seg_bn_cluster_bwd_supported(int dtype, long long P, int C) {
  if (!b200seg_bn_cluster_supported(dtype, P, C)) return 0;
  int cs = 1;
  bool resident = false;
  pick_layout(P, C, 64, &cs, &resident);
  return resident ? 1 : 0;
}

// Train-mode BatchNorm2d forward (statistics, running-stat update, normalise + activation (+ residual)) in one launch.
// sv: f32 [4][C] receives mean, invstd, scale, shift (the backward pass reads them).
extern "C" int b200seg_bn_cluster_fwd(const void* z, long long P, int C, const float* gamma, const float* beta, float eps,
                                      float momentum, float* running_mean, float* running_var, float* sv, const void* res,
                                      void* a, int act, b200seg_stream_t s) {
  B200_REQUIRE(z && gamma && beta && sv && a, "bn_cluster_fwd: null pointer");
  B200_REQUIRE(b200seg_bn_cluster_supported(B200SEG_BF16, P, C), "bn_cluster_fwd: unsupported layer P=%lld C=%d", P, C);
  BnFwdArgs g;
  g.z = (const __nv_bfloat16*)z; g.res = (const __nv_bfloat16*)res; g.a = (__nv_bfloat16*)a;
  g.gamma = gamma; g.beta = beta; g.running_mean = running_mean; g.running_var = running_var; g.sv = sv;
  g.P = P; g.C = C; g.act = act; g.eps = eps; g.momentum = momentum;
  return launch_cluster(bn_fwd_cluster_kernel<true>, bn_fwd_cluster_kernel<false>, g, P, C, 32, (cudaStream_t)s, "bn_cluster_fwd");
}

// Backward of the above: dz, and the two per-channel sums into slot 0 of the gradient staging block red [nslot][2][C].
extern "C" int b200seg_bn_cluster_bwd(const void* da, const void* z, const float* sv, long long P, int C, int act, double* red,
                                      void* dz, b200seg_stream_t s) {
  B200_REQUIRE(da && z && sv && red && dz, "bn_cluster_bwd: null pointer");
  B200_REQUIRE(b200seg_bn_cluster_bwd_supported(B200SEG_BF16, P, C), "bn_cluster_bwd: unsupported layer P=%lld C=%d", P, C);
  BnBwdArgs g;
  g.da = (const __nv_bfloat16*)da; g.z = (const __nv_bfloat16*)z; g.dz = (__nv_bfloat16*)dz;
  g.sv = sv; g.red = red; g.P = P; g.C = C; g.act = act;
  return launch_cluster(bn_bwd_cluster_kernel<true>, bn_bwd_cluster_kernel<false>, g, P, C, 64, (cudaStream_t)s, "bn_cluster_bwd");
}
