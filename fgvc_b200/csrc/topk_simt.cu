// K1 (generic engine): exact-fp32 affinity + radius mask + running top-K on CUDA cores.
// Used for every shape the tcgen05 engine does not take (C % 32 != 0, tiny maps) and as the
// fp32-exact cross-check of the 3xTF32 tensor path in the tests.
//
// One CTA = one 8x8 query tile of one job and one memory group.  For every memory frame
// it walks the 8x8 key tiles that intersect the radius halo of the query tile (all tiles
// for an unmasked frame), forms the 64x64 affinity tile with a register-blocked smem GEMM
// (4x4 per thread, float4 LDS along the contiguous channel dimension) and lets one owner
// thread per query fold the tile into its sorted top-K list held in registers.
// The 64 x (T*H*W) affinity never leaves the SM.
#include "common.cuh"

namespace fgvc {

constexpr int TQ = 8;        // query / key tile edge
constexpr int TP = TQ * TQ;  // 64 pixels

template <int K, int FMT>
__global__ void __launch_bounds__(256, 1)
affinity_topk_simt_kernel(const void* __restrict__ bank, int H, int W, int C,
                          const fgvc_job* __restrict__ jobs, const int32_t* __restrict__ mem_feat,
                          int radius, int mode, int groups, int k_out, float* __restrict__ tv,
                          int32_t* __restrict__ ti) {
  extern __shared__ __align__(16) float smem[];
  const int ld = C + 4;                 // row stride (floats); (C+4)/4 odd-ish => conflict-free LDS.128
  float* Qp = smem;                     // [64][ld]
  float* Kp = Qp + TP * ld;             // [64][ld]
  float* S = Kp + TP * ld;              // [64][65]

  const int tiles_x = (W + TQ - 1) / TQ;
  const int qy0 = (blockIdx.x / tiles_x) * TQ, qx0 = (blockIdx.x % tiles_x) * TQ;
  const int g = blockIdx.y;
  const fgvc_job job = jobs[blockIdx.z];
  const int n_pix = H * W;
  const int tid = threadIdx.x;
  const int c4n = C / 4;

  // memory entries of this group
  const int n_mem = job.mem_end - job.mem_begin;
  const int per = (n_mem + groups - 1) / groups;
  const int e_lo = job.mem_begin + g * per;
  const int e_hi = min(job.mem_end, e_lo + per);

  // stage the query tile: x = hi + lo (the fp32 value the split encodes)
  for (int i = tid; i < TP * c4n; i += 256) {
    int p = i / c4n, c4 = i - p * c4n;
    int y = qy0 + p / TQ, x = qx0 + p % TQ;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && x < W) v = bank_load4<FMT>(bank, job.q_slot, n_pix, C, y * W + x, c4);
    *reinterpret_cast<float4*>(Qp + p * ld + c4 * 4) = v;
  }

  TopK<K> top;
  top.init();
  const int my_qy = qy0 + tid / TQ, my_qx = qx0 + tid % TQ;   // owner threads: tid < 64
  const bool owner = tid < TP && my_qy < H && my_qx < W;
  const int reach = mask_reach(radius, mode);
  const int ty = tid >> 4, tx = tid & 15;

  for (int e = e_lo; e < e_hi; ++e) {
    const int raw = mem_feat[e];
    const bool masked = !(raw & FGVC_MEM_UNMASKED);
    const int slot = raw & ~FGVC_MEM_UNMASKED;
    const int pos_base = (e - job.mem_begin) * n_pix;
    int ky_lo = 0, ky_hi = H - 1, kx_lo = 0, kx_hi = W - 1;
    if (masked) {
      ky_lo = max(0, qy0 - reach); ky_hi = min(H - 1, qy0 + TQ - 1 + reach);
      kx_lo = max(0, qx0 - reach); kx_hi = min(W - 1, qx0 + TQ - 1 + reach);
    }
    for (int ty0 = ky_lo; ty0 <= ky_hi; ty0 += TQ) {
      for (int tx0 = kx_lo; tx0 <= kx_hi; tx0 += TQ) {
        if (masked) {   // closest approach of the two 8x8 rectangles
          int dy = max(0, max(ty0 - (qy0 + TQ - 1), qy0 - (ty0 + TQ - 1)));
          int dx = max(0, max(tx0 - (qx0 + TQ - 1), qx0 - (tx0 + TQ - 1)));
          if (!in_mask(dy, dx, radius, mode)) continue;   // block-uniform
        }
        for (int i = tid; i < TP * c4n; i += 256) {
          int p = i / c4n, c4 = i - p * c4n;
          int y = ty0 + p / TQ, x = tx0 + p % TQ;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (y < H && x < W) v = bank_load4<FMT>(bank, slot, n_pix, C, y * W + x, c4);
          *reinterpret_cast<float4*>(Kp + p * ld + c4 * 4) = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int c4 = 0; c4 < c4n; ++c4) {
          float4 a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(Qp + (ty + 16 * i) * ld + c4 * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Kp + (tx + 16 * j) * ld + c4 * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
              acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
              acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
              acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) S[(ty + 16 * i) * 65 + tx + 16 * j] = acc[i][j];
        __syncthreads();
        if (owner) {
          for (int n = 0; n < TP; ++n) {
            int ky = ty0 + n / TQ, kx = tx0 + n % TQ;
            if (ky >= H || kx >= W) continue;
            if (masked && !in_mask(ky - my_qy, kx - my_qx, radius, mode)) continue;
            float v = S[tid * 65 + n];
            if (v > top.thr()) top.push(v, pos_base + ky * W + kx);
          }
        }
      }
    }
  }
  if (owner) {
    int q = my_qy * W + my_qx;
    int64_t o = (((int64_t)blockIdx.z * groups + g) * n_pix + q) * k_out;
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i < k_out) {
        tv[o + i] = top.v[i];
        ti[o + i] = top.id[i];
      }
  }
}

template <int K, int FMT>
static int launch_k(const void* bank, int H, int W, int C, const fgvc_job* jobs, int n_jobs,
                    const int32_t* mem_feat, int radius, int mode, int k_out, int groups, float* tv,
                    int32_t* ti, cudaStream_t st) {
  size_t smem = (size_t)(2 * TP * (C + 4) + TP * 65) * sizeof(float);
  FGVC_CHECK_ARG(smem <= 227 * 1024, "simt engine: C=%d needs %zu B of shared memory", C, smem);
  FGVC_CUDA(cudaFuncSetAttribute(affinity_topk_simt_kernel<K, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  dim3 grid(cdiv(H, TQ) * cdiv(W, TQ), groups, n_jobs);
  affinity_topk_simt_kernel<K, FMT><<<grid, 256, smem, st>>>(bank, H, W, C, jobs, mem_feat, radius, mode, groups,
                                                       k_out, tv, ti);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

int launch_affinity_topk_simt(const void* bank, int fmt, int H, int W, int C, const fgvc_job* jobs, int n_jobs,
                              const int32_t* mem_feat, int radius, int mode, int K, int groups, float* tv,
                              int32_t* ti, cudaStream_t st) {
  FGVC_CHECK_ARG(C % 4 == 0, "simt engine: C=%d must be a multiple of 4", C);
#define FGVC_SIMT(KK)                                                                                            \
  return fmt == FGVC_BANK_TF32                                                                                   \
             ? launch_k<KK, FGVC_BANK_TF32>(bank, H, W, C, jobs, n_jobs, mem_feat, radius, mode, K, groups, tv, ti, st) \
             : launch_k<KK, FGVC_BANK_F16>(bank, H, W, C, jobs, n_jobs, mem_feat, radius, mode, K, groups, tv, ti, st)
  if (K <= 4) { FGVC_SIMT(4); }
  if (K <= 10) { FGVC_SIMT(10); }
  FGVC_SIMT(16);
#undef FGVC_SIMT
}

}  // namespace fgvc
