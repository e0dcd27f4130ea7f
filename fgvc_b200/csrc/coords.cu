// K3: bilinear up-sampling fused with img2coord (top-5 soft-argmax), the analytic
// frame-0 gaussian variant, and the VOS-style mask decode.  HBM-bound on the feature-res
// maps only: the full-resolution [T,P,h,w] tensor the reference ships to the host
// (vanilla_tracker.py:396-406) is never materialised.
#include "common.cuh"

namespace fgvc {

// F.interpolate(mode='bilinear', align_corners=False) source coordinates
struct Lerp {
  int i0, i1;
  float w1;
};
__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in_size) {
  float s = fmaxf(((float)dst + 0.5f) * scale - 0.5f, 0.f);
  int i0 = min((int)s, in_size - 1);
  Lerp l;
  l.i0 = i0;
  l.i1 = min(i0 + 1, in_size - 1);
  l.w1 = s - (float)i0;
  return l;
}

struct BilinearSrc {
  const float* m;  // [H][W] (shared or global)
  int H, W;
  float sy, sx;
  __device__ __forceinline__ float at(int oy, int ox) const {
    Lerp ly = lerp_coord(oy, sy, H), lx = lerp_coord(ox, sx, W);
    float w0x = 1.f - lx.w1, w0y = 1.f - ly.w1;
    const float* r0 = m + ly.i0 * W;
    const float* r1 = m + ly.i1 * W;
    return w0y * (w0x * r0[lx.i0] + lx.w1 * r0[lx.i1]) + ly.w1 * (w0x * r1[lx.i0] + lx.w1 * r1[lx.i1]);
  }
};

struct GaussSrc {
  float cx, cy, denom;
  __device__ __forceinline__ float at(int oy, int ox) const {
    float dx = (float)ox - cx, dy = (float)oy - cy;
    return expf(__fdiv_rn(-(dx * dx + dy * dy), denom));
  }
};

constexpr int CK = 8;  // candidates kept per thread (>= img2coord topk)

// block-wide soft-argmax over an implicit out_h x out_w map
template <class Src>
__device__ void soft_argmax_block(const Src& src, int out_h, int out_w, int topk, float* out_xy) {
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ int red_t[8];
  __shared__ double red_s[8];
  __shared__ float win_v[CK];
  __shared__ int win_i[CK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  TopK<CK> top;
  top.init();
  double sum = 0.0;
  const int total = out_h * out_w;
  for (int o = tid; o < total; o += 256) {
    int oy = o / out_w, ox = o - oy * out_w;
    float v = src.at(oy, ox);
    sum += (double)v;
    if (v >= top.thr()) top.push_ge(v, o);   // ties: the higher index wins
  }
  // all-zero test: np.sum(map) == 0
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red_s[warp] = sum;
  // topk rounds of block arg-max over the heads of the per-thread sorted lists
  int head = 0;
  for (int r = 0; r < topk; ++r) {
    float v = -INFINITY;
    int i = -1;
#pragma unroll
    for (int j = 0; j < CK; ++j)
      if (j == head) { v = top.v[j]; i = top.id[j]; }
    int t = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, i, o);
      int ot = __shfl_xor_sync(0xffffffffu, t, o);
      if (ov > v || (ov == v && oi > i)) { v = ov; i = oi; t = ot; }
    }
    __syncthreads();
    if (lane == 0) { red_v[warp] = v; red_i[warp] = i; red_t[warp] = t; }
    __syncthreads();
    float bv = red_v[0]; int bi = red_i[0], bt = red_t[0];
#pragma unroll
    for (int w = 1; w < 8; ++w)
      if (red_v[w] > bv || (red_v[w] == bv && red_i[w] > bi)) { bv = red_v[w]; bi = red_i[w]; bt = red_t[w]; }
    if (tid == bt) ++head;
    if (tid == 0) { win_v[r] = bv; win_i[r] = bi; }
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red_s[w];
    float x = -1.f, y = -1.f;
    if (tot != 0.0) {
      // np: ascending order, fp32 sum, + 1e-9, fp32 divide, fp64 weighted mean
      float s = 0.f;
      for (int r = topk - 1; r >= 0; --r) s += win_v[r];
      s += 1e-9f;
      double ax = 0.0, ay = 0.0;
      for (int r = topk - 1; r >= 0; --r) {
        double w = (double)__fdiv_rn(win_v[r], s);
        ax += (double)(win_i[r] % out_w) * w;
        ay += (double)(win_i[r] / out_w) * w;
      }
      x = (float)ax; y = (float)ay;
    }
    out_xy[0] = x; out_xy[1] = y;
  }
}

__global__ void __launch_bounds__(256)
heatmap_coords_kernel(const float* __restrict__ maps, int H, int W, int out_h, int out_w, int topk,
                      int use_smem, float* __restrict__ out_xy) {
  extern __shared__ float smap[];
  const float* m = maps + (int64_t)blockIdx.x * H * W;
  if (use_smem) {
    for (int i = threadIdx.x; i < H * W; i += 256) smap[i] = __ldg(m + i);
    __syncthreads();
    m = smap;
  }
  BilinearSrc src{m, H, W, (float)H / (float)out_h, (float)W / (float)out_w};
  soft_argmax_block(src, out_h, out_w, topk, out_xy + 2 * blockIdx.x);
}

__global__ void __launch_bounds__(256)
gaussian_coords_kernel(const float* __restrict__ pts, int out_h, int out_w, float denom, int topk,
                       float* __restrict__ out_xy) {
  GaussSrc src{__ldg(pts + 2 * blockIdx.x), __ldg(pts + 2 * blockIdx.x + 1), denom};
  soft_argmax_block(src, out_h, out_w, topk, out_xy + 2 * blockIdx.x);
}

// ---- VOS decode ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decode_minmax_kernel(const float* __restrict__ maps, int H, int W, int out_h, int out_w,
                     float* __restrict__ minmax) {
  __shared__ float smn[8], smx[8];
  const float* m = maps + (int64_t)blockIdx.x * H * W;
  BilinearSrc src{m, H, W, (float)H / (float)out_h, (float)W / (float)out_w};
  float mn = INFINITY, mx = -INFINITY;
  for (int o = threadIdx.x; o < out_h * out_w; o += 256) {
    int oy = o / out_w, ox = o - oy * out_w;
    float v = src.at(oy, ox);
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    minmax[2 * blockIdx.x] = mn; minmax[2 * blockIdx.x + 1] = mx;
  }
}

__global__ void __launch_bounds__(256)
decode_argmax_kernel(const float* __restrict__ maps, int L, int H, int W, int out_h, int out_w,
                     const float* __restrict__ minmax, uint8_t* __restrict__ out) {
  int o = blockIdx.x * 256 + threadIdx.x;
  if (o >= out_h * out_w) return;
  int oy = o / out_w, ox = o - oy * out_w;
  float best = -INFINITY;
  int arg = 0;
  for (int l = 0; l < L; ++l) {
    BilinearSrc src{maps + (int64_t)l * H * W, H, W, (float)H / (float)out_h, (float)W / (float)out_w};
    float v = src.at(oy, ox);
    float mn = __ldg(minmax + 2 * l), mx = __ldg(minmax + 2 * l + 1);
    if (mx > 0.f) v = __fdiv_rn(v - mn, (mx - mn) + 1e-12f);
    if (v > best) { best = v; arg = l; }
  }
  out[o] = (uint8_t)arg;
}

}  // namespace fgvc

using namespace fgvc;

extern "C" int fgvc_heatmap_coords(const float* maps, int32_t n_maps, int32_t H, int32_t W, int32_t out_h,
                                   int32_t out_w, int32_t topk, float* out_xy, void* stream) {
  FGVC_CHECK_ARG(maps && out_xy && n_maps > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0,
                 "fgvc_heatmap_coords: bad arguments");
  FGVC_CHECK_ARG(topk >= 1 && topk <= CK, "fgvc_heatmap_coords: topk=%d not in [1,%d]", topk, CK);
  size_t smem = (size_t)H * W * sizeof(float);
  int use_smem = smem <= 160 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    FGVC_CUDA(cudaFuncSetAttribute(heatmap_coords_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  heatmap_coords_kernel<<<n_maps, 256, use_smem ? smem : 0, (cudaStream_t)stream>>>(maps, H, W, out_h, out_w,
                                                                                     topk, use_smem, out_xy);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

extern "C" int fgvc_gaussian_coords(const float* points_xy, int32_t P, int32_t out_h, int32_t out_w, float sigma,
                                    int32_t topk, float* out_xy, void* stream) {
  FGVC_CHECK_ARG(points_xy && out_xy && P > 0 && out_h > 0 && out_w > 0 && sigma > 0,
                 "fgvc_gaussian_coords: bad arguments");
  FGVC_CHECK_ARG(topk >= 1 && topk <= CK, "fgvc_gaussian_coords: topk=%d not in [1,%d]", topk, CK);
  gaussian_coords_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(points_xy, out_h, out_w, 2.f * sigma * sigma, topk,
                                                              out_xy);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

extern "C" int fgvc_decode_masks(const float* maps, int32_t L, int32_t H, int32_t W, int32_t out_h, int32_t out_w,
                                 float* scratch_minmax, uint8_t* out_mask, void* stream) {
  FGVC_CHECK_ARG(maps && scratch_minmax && out_mask && L > 0 && L <= 255 && H > 0 && W > 0 && out_h > 0 && out_w > 0,
                 "fgvc_decode_masks: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  decode_minmax_kernel<<<L, 256, 0, st>>>(maps, H, W, out_h, out_w, scratch_minmax);
  FGVC_LAUNCH_CHECK();
  decode_argmax_kernel<<<cdiv(out_h * out_w, 256), 256, 0, st>>>(maps, L, H, W, out_h, out_w, scratch_minmax,
                                                                 out_mask);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}
