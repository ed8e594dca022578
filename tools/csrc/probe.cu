// Diagnostic entry point (not on the product path): issue rate of tcgen05.mma kind::f16 for a given M x N x 16 shape.
// Every CTA zeroes one A tile (128 rows x 128 B) and one B tile (256 rows x 128 B) in shared memory, then one thread issues
// `iters` accumulating MMAs back to back on the same operands and measures the cycles until the commit barrier fires.
// Used for DESIGN.md finding 8 (cost per MMA as a function of M and N, with one or two CTAs per SM).
#include "common.cuh"

extern "C" int b200seg_probe_mma(int M, int N, int iters, int ctas_per_sm, long long* cycles_out, b200seg_stream_t s);

namespace b200 {

__global__ void __launch_bounds__(128)
mma_probe_kernel(int M, int N, int iters, int tmem_cols, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // 16 KB
  uint8_t* sB = smem + 16384;               // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, (uint32_t)tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(M, N);
      const uint64_t da = umma_desc_k128(smem_u32(sA)), db = umma_desc_k128(smem_u32(sB));
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) umma_bf16(tmem, da + (uint64_t)(2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), idesc, i > 0 ? 1u : 0u);
      umma_commit(bar);
      mbar_wait(bar, 0, 99);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, (uint32_t)tmem_cols); }
}

}  // namespace b200

using namespace b200;

extern "C" int b200seg_probe_mma(int M, int N, int iters, int ctas_per_sm, long long* cycles_out, b200seg_stream_t s) {
  B200_REQUIRE((M == 64 || M == 128) && N >= 8 && N <= 256 && N % (M == 64 ? 8 : 16) == 0, "probe_mma: M=%d N=%d", M, N);
  B200_REQUIRE(iters > 0 && (ctas_per_sm == 1 || ctas_per_sm == 2) && cycles_out, "probe_mma: bad arguments");
  int cols = 32;
  while (cols < N) cols <<= 1;
  // shared memory request decides the residency: > half of the SM for one CTA per SM
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 52 * 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "probe_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  mma_probe_kernel<<<(unsigned)(sm_count() * ctas_per_sm), 128, smem, (cudaStream_t)s>>>(M, N, iters, cols, cycles_out);
  return check_launch("probe_mma");
}
