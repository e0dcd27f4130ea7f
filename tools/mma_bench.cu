// Microbenchmark: issue rate of tcgen05.mma kind::tf32 / kind::f16 for several shapes,
// SS (A from smem) vs TS (A from TMEM).  One CTA per SM, one issuing lane.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND>  // 0 tf32, 1 f16
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = all SS, 1 = all TS, 2 = pattern TS,SS,SS (the 3xTF32 K step), 3 = SS with two alternating accumulators
template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 1023);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = __shfl_sync(0xffffffffu, tslot, 0);
  if (warp == 0) {
    const uint64_t hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    const uint32_t fmt = KIND == 0 ? 2u : 1u;   // tf32 / bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t sa = smem_u32(smem);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const int ua = mode == 4 ? (u >> 1) : u;       // mode 4: consecutive pairs share the A tile
          const uint64_t a = hi | (uint64_t)((sa + (ua & 3) * 32 + ((ua >> 2) & 1) * 16384) >> 4);
          const uint64_t b = hi | (uint64_t)((sa + 65536 + (u & 3) * 32 + (u % 3) * 32768) >> 4);
          const uint32_t d = tb + ((mode == 3 && (u & 1)) ? 256 : 0);
          const bool ts = mode == 1 || (mode == 2 && (u % 3) == 0);
          if (mode == 5) {   // stacked 3xTF32 K step: SS(A_hi, [B_hi;B_lo], 2N) then TS(A_lo, B_hi, N) into the upper half
            const uint32_t idesc2 = (idesc & ~(0x3fu << 17)) | ((uint32_t)((2 * N) >> 3) << 17);
            if (u & 1) mma_ts<KIND>(tb + N, tb + 256 + 8 * (u & 3), b, idesc, 1);
            else mma_ss<KIND>(tb, a, b, idesc2, 1);
          } else if (ts) mma_ts<KIND>(d, tb + 256 + 8 * (u & 3), b, idesc, 1);
          else mma_ss<KIND>(d, a, b, idesc, 1);
        }
      }
      commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) {
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int KIND>
void run(const char* name, int N, int mode, int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int iters = 200;
  cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<KIND><<<grid, 128, 200 * 1024>>>(N, mode, iters, d);
  bench<KIND><<<grid, 128, 200 * 1024>>>(N, mode, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
  double cyc = (double)h[0] / (iters * 12.0);
  const double kdepth = KIND == 0 ? 8 : 16;
  printf("%-5s N=%3d mode=%d grid=%3d : %7.1f cycles/MMA  -> %7.1f MAC/clk/SM   (%s)\n", name, N, mode, grid, cyc,
         128.0 * N * kdepth / cyc, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {148}) {
    for (int N : {64, 128, 256}) {
      for (int mode : {0, 1, 2, 4, 5}) {
        if (mode == 5 && N == 256) continue;
        run<0>("tf32", N, mode, grid);
      }
    }
    for (int N : {128, 256}) for (int mode : {0, 1}) run<1>("bf16", N, mode, grid);
  }
  return 0;
}
