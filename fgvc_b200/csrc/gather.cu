// K1b: merge the per-group top-K lists, divide by the temperature, softmax over the K
// winners and gather + weighted-sum the label rows (local_attention.py:356-374).
// HBM/L2-bound: K label rows of Lp floats per query, read as float4, coalesced along L.
#include "common.cuh"

namespace fgvc {

constexpr int QB = 64;  // queries per CTA

template <int K>
__global__ void __launch_bounds__(256)
gather_labels_kernel(const float* __restrict__ tv, const int32_t* __restrict__ ti, int k_in, int groups,
                     const fgvc_job* __restrict__ jobs, int job_begin, const int32_t* __restrict__ mem_label,
                     int n_pix, float temperature, int flags, float* __restrict__ lab, int Lp) {
  __shared__ float sw[QB][K];
  __shared__ int srow[QB][K];
  const int jidx = job_begin + blockIdx.y;
  const fgvc_job job = jobs[jidx];
  const int q0 = blockIdx.x * QB;
  const int tid = threadIdx.x;
  if (tid < QB) {
    const int q = q0 + tid;
    TopK<K> top;
    top.init();
    if (q < n_pix) {
      for (int g = 0; g < groups; ++g) {
        int64_t o = (((int64_t)jidx * groups + g) * n_pix + q) * k_in;
        for (int i = 0; i < k_in; ++i) {
          float v = __ldg(tv + o + i);
          int id = __ldg(ti + o + i);
          if (id >= 0 && v > top.thr()) top.push(v, id);
        }
      }
    }
    // weights of the winners; empty entries (-inf) weigh 0.
    //   similarity  a = cos / temperature                      (dot_product, local_attention.py:321-323)
    //               a = (2 cos - 1) / sqrt(C) [= temperature]  (l2-distance on unit vectors, :324-327)
    //   weights     softmax(a) (:369)   or   clamp(a, 0)^2 ('cosine', :371)
    float a[K];
#pragma unroll
    for (int i = 0; i < K; ++i)
      a[i] = (flags & FGVC_SIM_L2) ? __fdiv_rn(2.f * top.v[i] - 1.f, temperature) : __fdiv_rn(top.v[i], temperature);
    float m = a[0], sum = 0.f;
    if (flags & FGVC_WEIGHT_COSINE) {
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float c = fmaxf(a[i], 0.f);
        a[i] = (i < k_in && top.id[i] >= 0) ? c * c : 0.f;
      }
      sum = 1.f;
    } else {
#pragma unroll
      for (int i = 0; i < K; ++i) {
        a[i] = (i < k_in && top.id[i] >= 0) ? expf(a[i] - m) : 0.f;
        sum += a[i];
      }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
      int id = top.id[i];
      bool ok = i < k_in && id >= 0;
      sw[tid][i] = ok ? ((flags & FGVC_WEIGHT_COSINE) ? a[i] : __fdiv_rn(a[i], sum)) : 0.f;
      int row = 0;
      if (ok) {
        int pos = id / n_pix;
        row = __ldg(mem_label + job.mem_begin + pos) * n_pix + (id - pos * n_pix);
      }
      srow[tid][i] = row;
    }
  }
  __syncthreads();
  const int l4n = Lp / 4;
  const float4* src = reinterpret_cast<const float4*>(lab);
  float4* dst = reinterpret_cast<float4*>(lab) + ((int64_t)job.out_slot * n_pix + q0) * l4n;
  const int nq = min(QB, n_pix - q0);
  for (int i = tid; i < nq * l4n; i += 256) {
    int q = i / l4n, c = i - q * l4n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float w = sw[q][j];
      if (w != 0.f) {
        float4 v = __ldg(src + (int64_t)srow[q][j] * l4n + c);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
    }
    dst[i] = acc;
  }
}

template <int K>
static int launch_g(const float* tv, const int32_t* ti, int k_in, int groups, const fgvc_job* jobs,
                    int job_begin, int n, const int32_t* mem_label, int n_pix, float temperature, int flags,
                    float* lab, int Lp, cudaStream_t st) {
  dim3 grid(cdiv(n_pix, QB), n);
  gather_labels_kernel<K><<<grid, 256, 0, st>>>(tv, ti, k_in, groups, jobs, job_begin, mem_label, n_pix,
                                                temperature, flags, lab, Lp);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

}  // namespace fgvc

using namespace fgvc;

extern "C" int fgvc_gather_labels(const float* topk_val, const int32_t* topk_idx, int32_t K, int32_t groups,
                                  const fgvc_job* jobs, int32_t job_begin, int32_t job_end,
                                  const int32_t* mem_label_slot, int32_t n_pix, float temperature, int32_t flags,
                                  float* lab_bank, int32_t Lp, void* stream) {
  FGVC_CHECK_ARG(topk_val && topk_idx && jobs && mem_label_slot && lab_bank, "fgvc_gather_labels: null pointer");
  FGVC_CHECK_ARG(K >= 1 && K <= 16, "fgvc_gather_labels: topk=%d not in [1,16]", K);
  FGVC_CHECK_ARG(groups >= 1 && job_end > job_begin && n_pix > 0, "fgvc_gather_labels: bad sizes");
  FGVC_CHECK_ARG(Lp > 0 && Lp % 4 == 0, "fgvc_gather_labels: Lp=%d must be a positive multiple of 4", Lp);
  FGVC_CHECK_ARG(temperature > 0.f, "fgvc_gather_labels: temperature must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  int n = job_end - job_begin;
  if (K <= 4) return launch_g<4>(topk_val, topk_idx, K, groups, jobs, job_begin, n, mem_label_slot, n_pix, temperature, flags, lab_bank, Lp, st);
  if (K <= 10) return launch_g<10>(topk_val, topk_idx, K, groups, jobs, job_begin, n, mem_label_slot, n_pix, temperature, flags, lab_bank, Lp, st);
  return launch_g<16>(topk_val, topk_idx, K, groups, jobs, job_begin, n, mem_label_slot, n_pix, temperature, flags, lab_bank, Lp, st);
}
