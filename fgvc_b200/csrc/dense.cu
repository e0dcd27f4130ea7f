// Dense propagation (topk = None, local_attention.py:376-383): the weights are a soft-max (or clamp(a,0)^2)
// over ALL allowed candidates instead of the k best, so there is nothing to select -- this is plain
// flash-attention with the radius mask: per 8x8 query tile, stream the 8x8 key tiles of the halo, form the
// 64x64 affinity tile (fp32 on the CUDA cores, the values the bank split encodes), fold it into a running
// (max, sum) per query and accumulate  P x labels  for a chunk of <= 64 label channels.  Nothing of size
// Nq x T*Nk ever exists.  No shipped config uses this mode (every test_cfg sets topk = 10), so it is
// built for correctness and API completeness, not tuned: CUDA cores, one CTA per (query tile, label chunk, job).
#include "common.cuh"

namespace fgvc {

constexpr int DQ = 8;            // tile edge
constexpr int DP = DQ * DQ;      // 64 pixels
constexpr int DLC = 64;          // label channels per CTA

template <int FMT>
__global__ void __launch_bounds__(256, 1)
dense_propagate_kernel(const void* __restrict__ bank, int H, int W, int C, const fgvc_job* __restrict__ jobs,
                       const int32_t* __restrict__ mem_feat, const int32_t* __restrict__ mem_label, int radius, int mode,
                       float temperature, int flags, float* __restrict__ lab, int Lp) {
  extern __shared__ __align__(16) float smem[];
  const int ld = C + 4;
  float* Qp = smem;                     // [64][ld]
  float* Kp = Qp + DP * ld;             // [64][ld]
  float* S = Kp + DP * ld;              // [64][65] affinity tile, then P
  float* Vt = S + DP * 65;              // [64 keys][DLC]
  float* row_m = Vt + DP * DLC;         // [64] running max
  float* row_s = row_m + DP;            // [64] running sum
  float* row_c = row_s + DP;            // [64] rescale factor of this tile

  const int tiles_x = (W + DQ - 1) / DQ;
  const int qy0 = (blockIdx.x / tiles_x) * DQ, qx0 = (blockIdx.x % tiles_x) * DQ;
  const int l0 = blockIdx.y * DLC;
  const int nl = min(DLC, Lp - l0);
  const fgvc_job job = jobs[blockIdx.z];
  const int n_pix = H * W;
  const int tid = threadIdx.x;
  const int c4n = C / 4;
  const bool cosine = (flags & FGVC_WEIGHT_COSINE) != 0;

  for (int i = tid; i < DP * c4n; i += 256) {
    int p = i / c4n, c4 = i - p * c4n;
    int y = qy0 + p / DQ, x = qx0 + p % DQ;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && x < W) v = bank_load4<FMT>(bank, job.q_slot, n_pix, C, y * W + x, c4);
    *reinterpret_cast<float4*>(Qp + p * ld + c4 * 4) = v;
  }
  if (tid < DP) { row_m[tid] = -INFINITY; row_s[tid] = 0.f; }

  const int my_qy = qy0 + tid / DQ, my_qx = qx0 + tid % DQ;   // owner threads: tid < 64
  const bool owner = tid < DP && my_qy < H && my_qx < W;
  const int reach = mask_reach(radius, mode);
  const int ty = tid >> 4, tx = tid & 15;
  float out[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[i][j] = 0.f;

  for (int e = job.mem_begin; e < job.mem_end; ++e) {
    const int raw = mem_feat[e];
    const bool masked = !(raw & FGVC_MEM_UNMASKED);
    const int slot = raw & ~FGVC_MEM_UNMASKED;
    const int lslot = mem_label[e];
    int ky_lo = 0, ky_hi = H - 1, kx_lo = 0, kx_hi = W - 1;
    if (masked) {
      ky_lo = max(0, qy0 - reach); ky_hi = min(H - 1, qy0 + DQ - 1 + reach);
      kx_lo = max(0, qx0 - reach); kx_hi = min(W - 1, qx0 + DQ - 1 + reach);
    }
    for (int ty0 = ky_lo; ty0 <= ky_hi; ty0 += DQ) {
      for (int tx0 = kx_lo; tx0 <= kx_hi; tx0 += DQ) {
        if (masked) {   // closest approach of the two 8x8 rectangles
          int dy = max(0, max(ty0 - (qy0 + DQ - 1), qy0 - (ty0 + DQ - 1)));
          int dx = max(0, max(tx0 - (qx0 + DQ - 1), qx0 - (tx0 + DQ - 1)));
          if (!in_mask(dy, dx, radius, mode)) continue;   // block-uniform
        }
        __syncthreads();                                   // previous tile's S / Vt / row_c are consumed
        for (int i = tid; i < DP * c4n; i += 256) {
          int p = i / c4n, c4 = i - p * c4n;
          int y = ty0 + p / DQ, x = tx0 + p % DQ;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (y < H && x < W) v = bank_load4<FMT>(bank, slot, n_pix, C, y * W + x, c4);
          *reinterpret_cast<float4*>(Kp + p * ld + c4 * 4) = v;
        }
        for (int i = tid; i < DP * DLC; i += 256) {        // label rows of the key tile, this CTA's channel chunk
          int n = i / DLC, l = i - n * DLC;
          int y = ty0 + n / DQ, x = tx0 + n % DQ;
          float v = 0.f;
          if (y < H && x < W && l < nl) v = __ldg(lab + ((int64_t)lslot * n_pix + y * W + x) * Lp + l0 + l);
          Vt[i] = v;
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
        // per query: similarity -> weights, folded into the running (max, sum)
        if (tid < DP) {
          float m_tile = -INFINITY;
          if (owner) {
            for (int n = 0; n < DP; ++n) {
              const int ky = ty0 + n / DQ, kx = tx0 + n % DQ;
              float a = -INFINITY;
              if (ky < H && kx < W && (!masked || in_mask(ky - my_qy, kx - my_qx, radius, mode))) {
                const float v = S[tid * 65 + n];
                a = (flags & FGVC_SIM_L2) ? __fdiv_rn(2.f * v - 1.f, temperature) : __fdiv_rn(v, temperature);
              }
              S[tid * 65 + n] = a;
              m_tile = fmaxf(m_tile, a);
            }
          }
          float scale = 1.f;
          if (!owner) {
            for (int n = 0; n < DP; ++n) S[tid * 65 + n] = 0.f;
          } else if (cosine) {
            for (int n = 0; n < DP; ++n) {
              const float c = fmaxf(S[tid * 65 + n], 0.f);      // -inf (not allowed) -> 0
              S[tid * 65 + n] = c * c;
            }
          } else {
            const float m_old = row_m[tid];
            const float m_new = fmaxf(m_old, m_tile);
            float sum = 0.f;
            if (m_new == -INFINITY) {
              for (int n = 0; n < DP; ++n) S[tid * 65 + n] = 0.f;
            } else {
              scale = expf(m_old - m_new);                        // m_old = -inf -> 0
              for (int n = 0; n < DP; ++n) {
                const float pv = expf(S[tid * 65 + n] - m_new);   // -inf -> 0
                S[tid * 65 + n] = pv;
                sum += pv;
              }
            }
            row_m[tid] = m_new;
            row_s[tid] = row_s[tid] * scale + sum;
          }
          row_c[tid] = scale;
        }
        __syncthreads();
        // out[q][l] = out * scale[q] + sum_n P[q][n] * V[n][l]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float sc = row_c[ty + 16 * i];
#pragma unroll
          for (int j = 0; j < 4; ++j) out[i][j] *= sc;
        }
        for (int n = 0; n < DP; ++n) {
          float pq[4], vl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pq[i] = S[(ty + 16 * i) * 65 + n];
#pragma unroll
          for (int j = 0; j < 4; ++j) vl[j] = Vt[n * DLC + tx + 16 * j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) out[i][j] = fmaf(pq[i], vl[j], out[i][j]);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ty + 16 * i;
    const int y = qy0 + p / DQ, x = qx0 + p % DQ;
    if (y >= H || x >= W) continue;
    const float s = row_s[p];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = tx + 16 * j;
      if (l < nl) {
        float v = out[i][j];
        if (!cosine) v = s > 0.f ? __fdiv_rn(v, s) : 0.f;
        lab[((int64_t)job.out_slot * n_pix + y * W + x) * Lp + l0 + l] = v;
      }
    }
  }
}

template <int FMT>
static int launch_dense(const void* bank, int H, int W, int C, const fgvc_job* jobs, int n_jobs, const int32_t* mem_feat,
                        const int32_t* mem_label, int radius, int mode, float temperature, int flags, float* lab, int Lp,
                        cudaStream_t st) {
  size_t smem = (size_t)(2 * DP * (C + 4) + DP * 65 + DP * DLC + 3 * DP) * sizeof(float);
  FGVC_CHECK_ARG(smem <= 227 * 1024, "dense propagation: C=%d needs %zu B of shared memory", C, smem);
  FGVC_CUDA(cudaFuncSetAttribute(dense_propagate_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(H, DQ) * cdiv(W, DQ), cdiv(Lp, DLC), n_jobs);
  dense_propagate_kernel<FMT><<<grid, 256, smem, st>>>(bank, H, W, C, jobs, mem_feat, mem_label, radius, mode,
                                                      temperature, flags, lab, Lp);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

}  // namespace fgvc

using namespace fgvc;

extern "C" int fgvc_dense_propagate(const void* feat_bank, int32_t bank_format, int32_t H, int32_t W, int32_t C,
                                    const fgvc_job* jobs, int32_t n_jobs, const int32_t* mem_feat_slot,
                                    const int32_t* mem_label_slot, int32_t radius, int32_t mask_mode, float temperature,
                                    int32_t flags, float* lab_bank, int32_t Lp, void* stream) {
  FGVC_CHECK_ARG(feat_bank && jobs && mem_feat_slot && mem_label_slot && lab_bank, "fgvc_dense_propagate: null pointer");
  FGVC_CHECK_ARG(H > 0 && W > 0 && C > 0 && C % 4 == 0 && n_jobs > 0, "fgvc_dense_propagate: bad shape (C %% 4 == 0)");
  FGVC_CHECK_ARG(Lp > 0 && Lp % 4 == 0, "fgvc_dense_propagate: Lp=%d must be a positive multiple of 4", Lp);
  FGVC_CHECK_ARG(radius >= 1, "fgvc_dense_propagate: radius=%d must be >= 1", radius);
  FGVC_CHECK_ARG(mask_mode == FGVC_MASK_CIRCLE || mask_mode == FGVC_MASK_SQUARE, "fgvc_dense_propagate: bad mask mode");
  FGVC_CHECK_ARG(temperature > 0.f, "fgvc_dense_propagate: temperature must be > 0");
  FGVC_CHECK_ARG(bank_format == FGVC_BANK_TF32 || bank_format == FGVC_BANK_F16, "fgvc_dense_propagate: bad bank format");
  cudaStream_t st = (cudaStream_t)stream;
  if (bank_format == FGVC_BANK_TF32)
    return launch_dense<FGVC_BANK_TF32>(feat_bank, H, W, C, jobs, n_jobs, mem_feat_slot, mem_label_slot, radius, mask_mode,
                                        temperature, flags, lab_bank, Lp, st);
  return launch_dense<FGVC_BANK_F16>(feat_bank, H, W, C, jobs, n_jobs, mem_feat_slot, mem_label_slot, radius, mask_mode,
                                     temperature, flags, lab_bank, Lp, st);
}
