// K1 (tcgen05 engine): fused affinity + radius mask + running top-K on the 5th-gen tensor
// cores.  The (T*H*W) x (H*W) affinity of local_attention.py:321-356 never reaches HBM.
//
// One CTA = one 128-query tile (8x16 or 16x8 pixels, M = 128) of one job and one memory
// group.  For every memory frame it walks the key boxes (BH rows x 16 columns, N = 16*BH)
// covering the radius halo of the tile -- image-clipped, corner boxes outside every
// query's mask skipped -- and for each box:
//   warp 0      TMA producer: per 32-channel chunk, cp.async.bulk.tensor (5-D map over
//               feat[slot][part][H][W][C], 128B swizzle) of the query tile (hi, lo) and the
//               key box (hi, lo) into a 3-stage shared-memory ring; out-of-image pixels are
//               zero-filled by the TMA unit, so halos need no branches.
//   warp 1      MMA issuer (one elected lane): tcgen05.mma kind::tf32, M=128 x N x K=8,
//               three products per K step -- lo*hi + hi*lo + hi*hi (3xTF32: fp32-faithful
//               selection) -- accumulating in one TMEM tile; two TMEM tiles ping-pong
//               between the tensor pipe and the epilogue.
//   warps 2..5  epilogue: thread = query (TMEM lane).  tcgen05.ld pulls 16 columns (one key
//               row) at a time; the analytic circle / square mask and the image bounds are a
//               16-bit interval mask per key row; survivors above the running K-th value are
//               inserted into the thread's sorted top-K list held in registers.
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers
// (MMA <-> epilogue).  A 128 x N fp32 tile costs N*128*4 B of TMEM reads and no HBM.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fgvc {

// Two operand-staging modes:
//  RES = true  (C <= 256, the normal case): the query tile's hi part stays RESIDENT in shared
//               memory for the whole CTA (C*512 B) and its lo part lives in TENSOR MEMORY
//               (C columns, written once with tcgen05.st; the lo*hi product is a TS-form MMA),
//               so only key boxes stream through the ring: 256 B per key per 32 channels, half
//               the L2->SM traffic of the streaming mode, which was the measured limiter.
//  RES = false (C > 256): query hi/lo chunks are re-streamed with every key box.
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_MAX_BH = 8;                     // N <= 128 (two accumulators = 256 TMEM columns)
constexpr int TC_TMEM_COLS = 512;                // [0,256) accumulators, [256,256+C) query lo (RES)
constexpr int TC_ALO_COL = 256;
constexpr int TC_EPI_WG = 4;                     // epilogue warpgroups (4 warps each)
constexpr int TC_THREADS = 64 + 128 * TC_EPI_WG;
constexpr int TC_SMEM_LIMIT = 227 * 1024;
constexpr int TC_SMEM_AUX = 1024;                // barriers, tables

struct TcParams {
  int H, W, C, n_pix;
  int radius, mode, reach;
  int QH, QW, qw_shift;      // query tile (8x16 or 16x8)
  int BH;                    // key box rows; N = 16 * BH
  int groups, k_out;
  int tiles_x;
  const fgvc_job* jobs;
  const int32_t* mem_feat;
  float* tv;
  int32_t* ti;
  int n_stages, stage_bytes, a_bytes;   // ring geometry (host-chosen)
  float* dbg;                // optional raw affinity dump [box][128][128]
  int32_t* dbg_meta;         // [box][4] = (mem entry, by, bx, N)
  int dbg_max_boxes;
};

template <int K, bool RES>
__global__ void __launch_bounds__(TC_THREADS, 1)
affinity_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const float* __restrict__ bank, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: [query hi, RES only: a_bytes][ring: n_stages * stage_bytes][aux]
  uint8_t* ring = smem + p.a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + p.n_stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + TC_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + TC_MAX_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2]
  uint64_t* a_bar = tempty_bar + 2;                  // query hi landed (RES)
  uint64_t* alo_bar = a_bar + 1;                     // query lo written to TMEM (RES)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(alo_bar + 1);
  int* halfw = reinterpret_cast<int*>(tmem_slot + 4);   // [reach+1] <= 128 entries
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();   // 128B swizzle needs a 1024-aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qy0 = (blockIdx.x / p.tiles_x) * p.QH, qx0 = (blockIdx.x % p.tiles_x) * p.QW;
  const int g = blockIdx.y;
  const fgvc_job job = p.jobs[blockIdx.z];
  const int n_mem = job.mem_end - job.mem_begin;
  const int per = (n_mem + p.groups - 1) / p.groups;
  const int e_lo = job.mem_begin + g * per;
  const int e_hi = min(job.mem_end, e_lo + per);
  const int N = 16 * p.BH;
  const int n_kc = p.C / 32;
  const uint32_t b_bytes = (uint32_t)N * 128u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar + b, 1); mbar_init(tempty_bar + b, 4 * TC_EPI_WG); }
    mbar_init(a_bar, 1);
    mbar_init(alo_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int d = threadIdx.x; d <= p.reach; d += TC_THREADS) {
    int hw = -1;
    if (p.mode == FGVC_MASK_CIRCLE) {
      while (hw + 1 <= p.reach && (hw + 1) * (hw + 1) + d * d < p.radius * p.radius) ++hw;
    } else {
      hw = p.radius;
    }
    halfw[d] = hw;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform

  if (warp == 0) {
    // ================================ TMA producer ====================================
    // the whole warp walks the boxes (uniform control flow); one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    if (RES && e_lo < e_hi) {      // the query tile's hi part: loaded once, resident
      if (elect_one()) {
        mbar_expect_tx(a_bar, (uint32_t)p.a_bytes);
        for (int kc = 0; kc < n_kc; ++kc)
          tma_load_5d(&tmap_q, a_bar, smem + kc * 16384, kc * 32, qx0, qy0, 0, job.q_slot);
      }
      __syncwarp();
    }
    for (int e = e_hi - 1; e >= e_lo; --e) {   // newest memory frame first: thresholds rise early
      const int raw = p.mem_feat[e];
      const int slot = raw & ~FGVC_MEM_UNMASKED;
      const Walk w = make_walk(p, raw, qy0, qx0);
      for (int by = w.y_lo; by <= w.y_hi; by += p.BH)
        for (int bx = w.x_lo; bx <= w.x_hi; bx += 16) {
          if (box_skipped(p, w, by, bx, qy0, qx0)) continue;
          for (int kc = 0; kc < n_kc; ++kc) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            if (elect_one()) {
              uint8_t* st = ring + stage * p.stage_bytes;
              if (RES) {
                mbar_expect_tx(full_bar + stage, 2u * b_bytes);
                tma_load_5d(&tmap_k, full_bar + stage, st, kc * 32, bx, by, 0, slot);
                tma_load_5d(&tmap_k, full_bar + stage, st + 16384, kc * 32, bx, by, 1, slot);
              } else {
                mbar_expect_tx(full_bar + stage, 2u * 16384u + 2u * b_bytes);
                tma_load_5d(&tmap_q, full_bar + stage, st, kc * 32, qx0, qy0, 0, job.q_slot);
                tma_load_5d(&tmap_q, full_bar + stage, st + 16384, kc * 32, qx0, qy0, 1, job.q_slot);
                tma_load_5d(&tmap_k, full_bar + stage, st + 32768, kc * 32, bx, by, 0, slot);
                tma_load_5d(&tmap_k, full_bar + stage, st + 49152, kc * 32, bx, by, 1, slot);
              }
            }
            __syncwarp();
            if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
          }
        }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =====================================
    // converged warp; one elected lane issues the 12 MMAs of a stage and the commits
    const uint32_t idesc = make_idesc(128, N);
    int stage = 0, buf = 0;
    uint32_t phase = 0, tphase0 = 0, tphase1 = 0;
    if (RES && e_lo < e_hi) {
      mbar_wait(a_bar, 0);
      mbar_wait(alo_bar, 0);
      tc_fence_after();
    }
    const uint32_t a_res = smem_u32(smem);
    const uint32_t ring_u32 = smem_u32(ring);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    for (int e = e_hi - 1; e >= e_lo; --e) {   // newest memory frame first: thresholds rise early
      const int raw = p.mem_feat[e];
      const Walk w = make_walk(p, raw, qy0, qx0);
      for (int by = w.y_lo; by <= w.y_hi; by += p.BH)
        for (int bx = w.x_lo; bx <= w.x_hi; bx += 16) {
          if (box_skipped(p, w, by, bx, qy0, qx0)) continue;
          mbar_wait(tempty_bar + buf, (buf ? tphase1 : tphase0) ^ 1);     // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
          for (int kc = 0; kc < n_kc; ++kc) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = ring_u32 + (uint32_t)(stage * p.stage_bytes);
              if (RES) {
                const uint64_t a_hi = desc_hi | (uint64_t)((a_res + kc * 16384) >> 4);
                const uint64_t b_hi = desc_hi | (uint64_t)(sa >> 4), b_lo = desc_hi | (uint64_t)((sa + 16384) >> 4);
                const uint32_t a_lo = tmem_base + TC_ALO_COL + kc * 32;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {     // 4 x (K = 8 tf32 = 32 B) per 128 B swizzle row
                  const uint64_t o = (uint64_t)(ks * 2);
                  umma_tf32_ts(d_tmem, a_lo + ks * 8, b_hi + o, idesc, (kc | ks) != 0);
                  umma_tf32(d_tmem, a_hi + o, b_lo + o, idesc, 1);
                  umma_tf32(d_tmem, a_hi + o, b_hi + o, idesc, 1);
                }
              } else {
                const uint64_t a_hi = desc_hi | (uint64_t)(sa >> 4), a_lo = desc_hi | (uint64_t)((sa + 16384) >> 4);
                const uint64_t b_hi = desc_hi | (uint64_t)((sa + 32768) >> 4), b_lo = desc_hi | (uint64_t)((sa + 49152) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t o = (uint64_t)(ks * 2);  // +32 B in the >>4 start-address field
                  umma_tf32(d_tmem, a_lo + o, b_hi + o, idesc, (kc | ks) != 0);
                  umma_tf32(d_tmem, a_hi + o, b_lo + o, idesc, 1);
                  umma_tf32(d_tmem, a_hi + o, b_hi + o, idesc, 1);
                }
              }
              umma_commit(empty_bar + stage);         // smem slot free once these MMAs retire
              if (kc == n_kc - 1) umma_commit(tfull_bar + buf);   // accumulator complete
            }
            __syncwarp();
            if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
          }
          if (buf) tphase1 ^= 1; else tphase0 ^= 1;
          buf ^= 1;
        }
    }
  } else {
    // ================================== epilogue ======================================
    // TC_EPI_WG warpgroups; warpgroup wg owns key rows wg, wg + TC_EPI_WG of every box, so the
    // four schedulers of the SM each interleave TC_EPI_WG warps.  Lists are merged at the end.
    const int wg = (warp - 2) >> 2;
    const int lg = warp & 3;                          // TMEM lane group this warp may access
    const int m = lg * 32 + lane;                     // query row in the tile
    const int qy = qy0 + (m >> p.qw_shift), qx = qx0 + (m & (p.QW - 1));
    const bool qvalid = qy < p.H && qx < p.W;
    if (RES && wg == 0 && e_lo < e_hi) {
      // query lo part -> tensor memory: lane = query row, column TC_ALO_COL + channel
      const float4* src = reinterpret_cast<const float4*>(
          bank + ((int64_t)job.q_slot * 2 + 1) * p.n_pix * p.C + (int64_t)(qvalid ? qy * p.W + qx : 0) * p.C);
      const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + TC_ALO_COL;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < p.C; c += 16) {
        float4 a = z, b = z, c4 = z, d = z;
        if (qvalid) { a = __ldg(src + c / 4); b = __ldg(src + c / 4 + 1); c4 = __ldg(src + c / 4 + 2); d = __ldg(src + c / 4 + 3); }
        tmem_st16(ta + c, a, b, c4, d);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(alo_bar);
    }
    TopK<K> top;
    top.init();
    int buf = 0;
    uint32_t tphase[2] = {0, 0};
    int box_seq = 0;
    // 16-bit interval mask of the in-mask, in-image keys of key row ky for this thread's query
    auto row_bits = [&](const Walk& w, int ky, int bx, bool row_ok) -> uint32_t {
      if (!(qvalid && row_ok && ky < p.H)) return 0u;
      int lo, hi;
      if (w.masked) {
        int ady = abs(ky - qy);
        int hw = ady <= p.reach ? halfw[ady] : -1;
        lo = hw < 0 ? 1 : max(qx - hw, 0);
        hi = hw < 0 ? 0 : min(qx + hw, p.W - 1);
      } else {
        lo = 0; hi = p.W - 1;
      }
      lo = max(lo - bx, 0);
      hi = min(hi - bx, 15);
      return hi >= lo ? (2u << hi) - (1u << lo) : 0u;
    };
    auto scan16 = [&](const uint32_t* r, uint32_t bits, int kbase, bool active) {
      // candidates = in-mask elements above the running K-th value; inserted in warp-wide rounds
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      const float thr0 = top.thr();
      uint32_t cand = 0;
      if (active) {
#pragma unroll
        for (int j = 0; j < 16; ++j) cand |= (v[j] > thr0) ? (1u << j) : 0u;
        cand &= bits;
      }
      while (__any_sync(0xffffffffu, cand != 0)) {
        if (cand) {
          const int j = __ffs(cand) - 1;
          cand &= cand - 1;
          const float x = select16(v, j);
          if (x > top.thr()) top.push(x, kbase + j);
        }
      }
    };
    for (int e = e_hi - 1; e >= e_lo; --e) {   // newest memory frame first: thresholds rise early
      const int raw = p.mem_feat[e];
      const Walk w = make_walk(p, raw, qy0, qx0);
      const int pos_base = (e - job.mem_begin) * p.n_pix;
      for (int by = w.y_lo; by <= w.y_hi; by += p.BH)
        for (int bx = w.x_lo; bx <= w.x_hi; bx += 16) {
          if (box_skipped(p, w, by, bx, qy0, qx0)) continue;
          const int rowA = wg, rowB = wg + TC_EPI_WG;
          const uint32_t bitsA = row_bits(w, by + rowA, bx, rowA < p.BH);
          const uint32_t bitsB = row_bits(w, by + rowB, bx, rowB < p.BH);
          const bool dump = p.dbg != nullptr && box_seq < p.dbg_max_boxes;
          const bool doA = (__any_sync(0xffffffffu, bitsA != 0) || dump) && rowA < p.BH;   // warp-uniform
          const bool doB = (__any_sync(0xffffffffu, bitsB != 0) || dump) && rowB < p.BH;
          mbar_wait(tfull_bar + buf, tphase[buf]);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 128);
          uint32_t ra[16], rb[16];
          if (doA) tmem_ld16_issue(taddr + (uint32_t)(rowA * 16), ra);
          if (doB) tmem_ld16_issue(taddr + (uint32_t)(rowB * 16), rb);
          if (doA) tmem_ld_wait(ra); else if (doB) tmem_ld_wait(rb);
          if (doA && doB) reg_fence16(rb);
          // the accumulator is in registers: hand the TMEM tile back before the scan
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + buf);
          if (dump) {
            float* d = p.dbg + ((int64_t)box_seq * 128 + m) * 128;
            if (doA) for (int j = 0; j < 16; ++j) d[rowA * 16 + j] = __uint_as_float(ra[j]);
            if (doB) for (int j = 0; j < 16; ++j) d[rowB * 16 + j] = __uint_as_float(rb[j]);
            if (p.dbg_meta != nullptr && m == 0 && wg == 0) {
              p.dbg_meta[4 * box_seq + 0] = e; p.dbg_meta[4 * box_seq + 1] = by;
              p.dbg_meta[4 * box_seq + 2] = bx; p.dbg_meta[4 * box_seq + 3] = N;
            }
          }
          if (doA) scan16(ra, bitsA, pos_base + (by + rowA) * p.W + bx, bitsA != 0);   // doA/doB are warp-uniform
          if (doB) scan16(rb, bitsB, pos_base + (by + rowB) * p.W + bx, bitsB != 0);
          ++box_seq;
          tphase[buf] ^= 1;
          buf ^= 1;
        }
    }
    // ---- merge the TC_EPI_WG partial lists of every query through the (now idle) ring
    asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_WG) : "memory");
    float* mv = reinterpret_cast<float*>(ring);
    int* mi = reinterpret_cast<int*>(ring + TC_EPI_WG * 128 * K * 4);
    if (wg > 0) {
#pragma unroll
      for (int i = 0; i < K; ++i) { mv[(wg * 128 + m) * K + i] = top.v[i]; mi[(wg * 128 + m) * K + i] = top.id[i]; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(128 * TC_EPI_WG) : "memory");
    if (wg == 0 && qvalid) {
      for (int w2 = 1; w2 < TC_EPI_WG; ++w2)
        for (int i = 0; i < K; ++i) {
          const float v = mv[(w2 * 128 + m) * K + i];
          if (!(v > top.thr())) break;               // lists are sorted descending
          top.push(v, mi[(w2 * 128 + m) * K + i]);
        }
      const int q = qy * p.W + qx;
      const int64_t o = (((int64_t)blockIdx.z * p.groups + g) * p.n_pix + q) * p.k_out;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (i < p.k_out) { p.tv[o + i] = top.v[i]; p.ti[o + i] = top.id[i]; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host
// 5-D map over feat[slot][part][H][W][C]; box = (32 channels, bw, bh, 1, 1), 128B swizzle
static int make_map(CUtensorMap* map, const float* bank, int n_slots, int H, int W, int C, int bw, int bh) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return FGVC_ERR_CUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 2, (cuuint64_t)n_slots};
  cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4,
                           (cuuint64_t)2 * H * W * C * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)bank, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (H=%d W=%d C=%d box=%dx%d)", (int)r, H, W, C, bw, bh);
    return FGVC_ERR_CUDA;
  }
  return FGVC_OK;
}

bool tc_supported(int H, int W, int C, int K) {
  return C % 32 == 0 && C >= 32 && C <= 1024 && K >= 1 && K <= 16 && H >= 1 && W >= 1;
}

// key-box rows (<= 8) for a halo of `rows` rows.  Per box and 32-channel chunk the CTA streams
// 128 query rows + N key rows and issues MMAs for N columns, so cost ~ n_boxes * (128 + N).
static int box_cost(int rows, int bh) { return cdiv(rows, bh) * (128 + 16 * bh); }
static int pick_bh(int rows) {
  int best = TC_MAX_BH;
  for (int bh = TC_MAX_BH - 1; bh >= 1; --bh)
    if (box_cost(rows, bh) < box_cost(rows, best)) best = bh;
  return best;
}

template <int K, bool RES>
static int launch_tc(const CUtensorMap& mq, const CUtensorMap& mk, const float* bank, const TcParams& p, dim3 grid,
                     cudaStream_t st) {
  const int smem = p.a_bytes + p.n_stages * p.stage_bytes + TC_SMEM_AUX;
  FGVC_CUDA(cudaFuncSetAttribute(affinity_topk_tc_kernel<K, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  affinity_topk_tc_kernel<K, RES><<<grid, TC_THREADS, smem, st>>>(mq, mk, bank, p);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

int launch_affinity_topk_tc(const float* bank, int n_slots, int H, int W, int C, const fgvc_job* jobs, int n_jobs,
                            const int32_t* mem_feat, int radius, int mode, int K, int groups, float* tv, int32_t* ti,
                            float* dbg, int32_t* dbg_meta, int dbg_max_boxes, cudaStream_t st) {
  TcParams p;
  p.H = H; p.W = W; p.C = C; p.n_pix = H * W;
  p.radius = radius; p.mode = mode; p.reach = mask_reach(radius, mode);
  // orientation of the 128-query tile: the one whose padded halo is smaller
  const int reach = p.reach;
  auto halo_cost = [&](int qh, int qw) {
    int rows = min(H, qh + 2 * reach), cols = min(W, qw + 2 * reach);
    double tiles = (double)cdiv(H, qh) * cdiv(W, qw);
    return tiles * box_cost(rows, pick_bh(rows)) * cdiv(cols, 16);
  };
  if (halo_cost(16, 8) < halo_cost(8, 16)) { p.QH = 16; p.QW = 8; p.qw_shift = 3; }
  else { p.QH = 8; p.QW = 16; p.qw_shift = 4; }
  p.BH = pick_bh(min(H, p.QH + 2 * reach));
  p.groups = groups; p.k_out = K;
  p.tiles_x = cdiv(W, p.QW);
  p.jobs = jobs; p.mem_feat = mem_feat; p.tv = tv; p.ti = ti;
  p.dbg = dbg; p.dbg_meta = dbg_meta; p.dbg_max_boxes = dbg_max_boxes;
  FGVC_CHECK_ARG(p.reach + 1 <= 128, "tcgen05 engine: radius %d too large", radius);
  CUtensorMap mq, mk;
  int rc = make_map(&mq, bank, n_slots, H, W, C, p.QW, p.QH);
  if (rc) return rc;
  rc = make_map(&mk, bank, n_slots, H, W, C, 16, p.BH);
  if (rc) return rc;
  dim3 grid(cdiv(H, p.QH) * p.tiles_x, groups, n_jobs);
  static const bool force_stream = getenv("FGVC_TC_STREAM") != nullptr;   // debugging aid
  const bool res = C <= 256 && !force_stream;
  if (res) {
    p.a_bytes = C * 512;
    p.stage_bytes = 32 * 1024;
  } else {
    p.a_bytes = 0;
    p.stage_bytes = 64 * 1024;
  }
  p.n_stages = min(TC_MAX_STAGES, (TC_SMEM_LIMIT - TC_SMEM_AUX - p.a_bytes) / p.stage_bytes);
  if (res) {
    if (K <= 4) return launch_tc<4, true>(mq, mk, bank, p, grid, st);
    if (K <= 10) return launch_tc<10, true>(mq, mk, bank, p, grid, st);
    return launch_tc<16, true>(mq, mk, bank, p, grid, st);
  }
  if (K <= 4) return launch_tc<4, false>(mq, mk, bank, p, grid, st);
  if (K <= 10) return launch_tc<10, false>(mq, mk, bank, p, grid, st);
  return launch_tc<16, false>(mq, mk, bank, p, grid, st);
}

}  // namespace fgvc
