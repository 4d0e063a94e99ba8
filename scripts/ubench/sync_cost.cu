// Micro-benchmark (debug aid, not part of the library): thread-side cost of the mbarrier / tcgen05 primitives the
// GEMM pipelines are built from.  One warp, one CTA; cycles per operation via clock64.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void k(long long* out, int N, int nmma_n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), b2 = smem_u32(&bars[2]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b1), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b2), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  __syncthreads();
  const uint32_t tmem = tmem_slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 32) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  long long t0, t1;
  int sink = 0;
  // 0: try_wait on an already completed phase (parity 1 on a fresh barrier)
  t0 = clock64();
  for (int i = 0; i < N; ++i) sink += try_wait(b0, 1);
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  // 1: test_wait likewise
  t0 = clock64();
  for (int i = 0; i < N; ++i) sink += test_wait(b0, 1);
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;
  // 2: elect + local arrive then try_wait round trip (phase flips every iteration)
  uint32_t ph = 0;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b0) : "memory");
    while (!try_wait(b0, ph)) {}
    ph ^= 1;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = t1 - t0;
  // 3: elect + tcgen05.commit (nothing outstanding) then try_wait round trip
  ph = 0;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    if (elect_one()) commit(b1);
    while (!try_wait(b1, ph)) {}
    ph ^= 1;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = t1 - t0;
  // 4: elect only
  t0 = clock64();
  for (int i = 0; i < N; ++i) sink += elect_one();
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = t1 - t0;
  // 5: MMA issue: 4 x (M=128, N=nmma_n, K=16) + commit per iteration, wait only every 8th iteration (queue depth probe)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)nmma_n >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem) + 16384);
  ph = 0;
  long long issue = 0;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    long long a0 = clock64();
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) mma(tmem, ad + 2 * kk, bd + 2 * kk, idesc, 1);
      commit(b2);
    }
    issue += clock64() - a0;
    while (!try_wait(b2, ph)) {}
    ph ^= 1;
  }
  t1 = clock64();
  if (threadIdx.x == 0) { out[5] = t1 - t0; out[6] = issue; }
  // 7: MMA issue only, 4*N MMAs back to back, one commit at the end
  t0 = clock64();
  if (elect_one()) {
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) mma(tmem, ad + 2 * kk, bd + 2 * kk, idesc, 1);
    }
    commit(b2);
  }
  long long t_issue = clock64();
  while (!try_wait(b2, ph)) {}
  t1 = clock64();
  if (threadIdx.x == 0) { out[7] = t_issue - t0; out[8] = t1 - t0; }
  if (sink == 123456789) out[15] = sink;
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int N = 1000;
  for (int nn : {208, 256, 64}) {
    cudaMemset(d, 0, 16 * sizeof(long long));
    k<<<1, 32, 64 * 1024>>>(d, N, nn);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N_mma=%d (%s): try_wait(done) %.1f  test_wait(done) %.1f  arrive+wait %.1f  commit+wait %.1f  elect %.1f | "
           "4mma+commit+wait %.1f (issue part %.1f) | 4mma stream: issue %.1f total %.1f cycles per 4 MMAs\n",
           nn, cudaGetErrorString(e), h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N,
           h[4] / (double)N, h[5] / (double)N, h[6] / (double)N, h[7] / (double)N, h[8] / (double)N);
  }
  return 0;
}
