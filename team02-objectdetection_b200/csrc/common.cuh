// common.cuh -- sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM), vector helpers, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200seg.h"

namespace b200 {

// ------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message; C ABI returns ints)
// ------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);          // cudaPeekAtLastError -> code
int sm_count();

// Programmatic dependent launch: a kernel launched through launch_pdl may start (prologue: barrier init, TMEM
// allocation, weight loads) while the previous kernel of the stream drains.  Every such kernel executes
// pdl_wait() before it touches memory a previous kernel wrote (or that a previous kernel still reads and this
// one writes), and pdl_trigger() as early as possible.  B200SEG_PDL=0 turns the launch attribute off.
int pdl_mode();      // 0 off, 1 every forward kernel + early trigger, 2 conv_tc only, 3 all + late trigger in conv_tc
template <typename... KA, typename... A>
inline cudaError_t launch_pdl_if(bool on, void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}
template <typename... KA, typename... A>
inline cudaError_t launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  const int m = pdl_mode();
  return launch_pdl_if(m == 1 || m == 3, kern, grid, block, smem, st, args...);
}

#define B200_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::b200::set_error(-1, __VA_ARGS__);       \
  } while (0)

// TMA descriptor factory (cuTensorMapEncodeTiled resolved through the runtime, no -lcuda)
// dims/strides innermost first; strides[i] is the byte stride of dim i+1.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle128,
                   const uint32_t* elem_strides = nullptr);   // box[i] is the TRAVERSED extent: loads ceil(box/es)

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trapped kernel (an error the host sees), never a
// hung GPU.  2 s is ~1000x the longest legitimate wait in any kernel here.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
      printf("b200seg: mbarrier wait timeout tag=%d block=(%d,%d) thread=%d parity=%u\n", tag, blockIdx.x,
             blockIdx.y, threadIdx.x, parity);
      __trap();
    }
  }
}

// one lane of the (converged) warp; ptxas keeps the guarded code in the uniform datapath
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- proxies / fences ----
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors passed as 32-bit halves (the high word is a per-kernel constant)
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i = lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 8 columns into the first 8 registers of r
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major operand tile, 128-byte swizzle, rows of 64 bf16 (128 B),
// 8-row groups 1024 B apart (SBO), LBO unused for swizzled K-major (set to 1), version 1 (sm_100).
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, 16-byte units      [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored)     [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B       [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)    [46,48)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B                      [61,64)
  return d;
}
// UMMA shared-memory descriptor of a K-major operand WITHOUT swizzle: 8-row x 16-byte core matrices (128 contiguous bytes
// each); LBO = byte distance between core matrices adjacent along K, SBO = between 8-row groups along M/N.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell); layout type 0 = SWIZZLE_NONE
  return d;
}
// Instruction descriptor for kind::f16, A=B=bf16 K-major, D=f32, shape M x N (cute InstrDescriptor)
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- small math / vector helpers ----
template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
  if (ACT == B200SEG_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == B200SEG_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
__device__ __forceinline__ float apply_act_rt(float v, int act) {
  if (act == B200SEG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == B200SEG_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// sm_100a mixed-precision FMA (SASS FHFMA.BF16, full FP32-pipe rate, measured in tools/fhfma_probe.cu): c + a.half * b.half
// with both bf16 operands taken straight from either half of a packed register -- no bf16 -> f32 unpack instructions.
__device__ __forceinline__ float fma_bf16_lo(uint32_t a, uint32_t b, float c) {
  float d;
  asm("{.reg .b16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.bf16 %0, al, bl, %3;}"
      : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fma_bf16_hi(uint32_t a, uint32_t b, float c) {
  float d;
  asm("{.reg .b16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2; fma.rn.f32.bf16 %0, ah, bh, %3;}"
      : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}

// A 16-byte vector of activations in storage type T, viewed as floats.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  typedef float4 Raw;                          // the 16 bytes as loaded: lets a kernel issue a batch of loads first and
  float v[4];                                  // unpack later (see train_ops.cu, "batched streaming")
  static __device__ __forceinline__ Raw ldraw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void unpack(const Raw& t) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ void unpack_dep(const Raw& t, uint32_t) { unpack(t); }
  __device__ __forceinline__ void load(const float* p) { unpack(ldraw(p)); }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  typedef uint4 Raw;
  float v[8];
  // volatile asm: a batch of these stays a batch (ptxas sank plain __ldg loads back into the arithmetic to save registers)
  static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) {
    uint4 t;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    return t;
  }
  __device__ __forceinline__ void unpack(const Raw& t) {
    v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
    v[4] = bf16lo(t.z); v[5] = bf16hi(t.z); v[6] = bf16lo(t.w); v[7] = bf16hi(t.w);
  }
  // unpack with a run-time-zero word folded into every element ((x << 16) + dep is one LEA, (x & 0xffff0000) | dep one
  // LOP3: no extra instructions): makes the arithmetic on this vector depend on whatever `dep` was computed from --
  // see all_loaded() in train_ops.cu
  __device__ __forceinline__ void unpack_dep(const Raw& t, uint32_t dep) {
    v[0] = __uint_as_float((t.x << 16) + dep); v[1] = __uint_as_float((t.x & 0xffff0000u) | dep);
    v[2] = __uint_as_float((t.y << 16) + dep); v[3] = __uint_as_float((t.y & 0xffff0000u) | dep);
    v[4] = __uint_as_float((t.z << 16) + dep); v[5] = __uint_as_float((t.z & 0xffff0000u) | dep);
    v[6] = __uint_as_float((t.w << 16) + dep); v[7] = __uint_as_float((t.w & 0xffff0000u) | dep);
  }
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { unpack(ldraw(p)); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
    t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

#endif  // __CUDACC__

}  // namespace b200
