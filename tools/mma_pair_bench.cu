// Microbenchmark: issue rate of tcgen05.mma kind::f16, TS form (A from tensor memory), for
//   cta_group::1 (M = 128, one CTA per SM) and cta_group::2 (M = 256, a CTA pair, each CTA holding half of B),
// with the operand pattern of K1's K step (hi*hi, hi*lo, lo*hi into one accumulator) or one repeated MMA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_pair_bench tools/mma_pair_bench.cu
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
template <int NCTA>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (NCTA == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (NCTA == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: one repeated MMA; mode 1: K1's K step (A_hi x B_hi, A_hi x B_lo, A_lo x B_hi), walking 16 K steps of a stage;
// mode 2: as 1 but the stacked form (A_hi x [B_hi;B_lo] with 2N columns, then A_lo x B_hi with N) -- cta_group::1 only
template <int NCTA>
__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (NCTA == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u + (i & 255);   // small fp16 pairs
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if (NCTA == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (NCTA == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = __shfl_sync(0xffffffffu, tslot, 0);
  if (warp == 0 && rank == 0) {
    const uint64_t hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * NCTA) >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)((128 * NCTA) >> 4) << 24);
    const uint32_t sa = smem_u32(smem);
    const uint32_t NC = N / NCTA;                        // B rows held by this CTA
    const uint64_t lo_off = (uint64_t)((NC * 128) >> 4);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {                 // 16 K steps: 4 chunks x 4 x 32 B
          const uint64_t b = hi | (uint64_t)((sa + (u >> 2) * (2 * NC * 128) + (u & 3) * 32) >> 4);
          const uint32_t a_hi = tb + 256 + (u >> 2) * 32 + (u & 3) * 8, a_lo = a_hi + 128;
          if (mode == 0) {
            mma_ts<NCTA>(tb, tb + 256, hi | (uint64_t)(sa >> 4), idesc, 1);
            mma_ts<NCTA>(tb, tb + 256, hi | (uint64_t)(sa >> 4), idesc, 1);
            mma_ts<NCTA>(tb, tb + 256, hi | (uint64_t)(sa >> 4), idesc, 1);
          } else if (mode == 1) {
            mma_ts<NCTA>(tb, a_hi, b, idesc, 1);
            mma_ts<NCTA>(tb, a_hi, b + lo_off, idesc, 1);
            mma_ts<NCTA>(tb, a_lo, b, idesc, 1);
          } else {
            mma_ts<NCTA>(tb, a_hi, b, idesc2, 1);
            mma_ts<NCTA>(tb + N, a_lo, b, idesc, 1);
          }
        }
      }
      commit<NCTA>(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) {
      t1 = clock64();
      out[blockIdx.x / NCTA] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (NCTA == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __syncthreads();
  }
  if (warp == 0) {
    if (NCTA == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

template <int NCTA>
void run(int N, int mode, int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int iters = 100;
  auto kern = bench<NCTA>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, kern, N, mode, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * (grid / NCTA), cudaMemcpyDeviceToHost);
  const int per_step = mode == 2 ? 2 : 3;
  const double cyc_step = (double)h[0] / (iters * 16.0);
  // per K step and SM: 128 rows x N keys x 16 channels x 3 products
  printf("cta_group::%d M=%3d N=%3d mode=%d : %7.1f cycles/K-step (%5.1f per MMA) -> %7.1f MAC/clk/SM   (%s)\n", NCTA,
         128 * NCTA, N, mode, cyc_step, cyc_step / per_step, 128.0 * N * 16 * 3 / cyc_step, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int N : {32, 64, 128}) for (int mode : {0, 1, 2}) { if (mode == 2 && N > 64) continue; run<1>(N, mode, 148); }
  for (int N : {32, 64, 96, 128}) for (int mode : {0, 1}) run<2>(N, mode, 148);
  return 0;
}
