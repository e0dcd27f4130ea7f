// K2 fine stage on the tensor cores (tcgen05, fp16 three-term split): WINDOW mode of the K1 engine.
//
// masked_attention_efficient_c2f (local_attention.py:721-880): every coarse query looks, in every memory frame, at
// the (2 rf + 1)^2 window of the FINE key map centred at scale * (coarse arg-max key) -- a data-dependent centre
// per (query, frame).  One warp per candidate on the CUDA cores (c2f.cu) is bound by instruction issue
// (profiles/r1_i_c2f_fine.md).  Here the stage is the K1 machinery of topk_tc16.cu (read that header first) with
// three changes of geometry:
//   * a tile = 128 COARSE queries (8x16 block of the coarse grid); their operand rows are the fine query
//     features at the strided positions (scale * qy, scale * qx);
//   * one CTA = (tile, memory entry, chunk of that entry's boxes) -- a coarse grid has few tiles, so the entries
//     and their box lists are spread over the grid and the tail merges the partial lists;
//   * the key boxes of a memory entry cover the bounding rectangle of the tile's 128 window centres, grown by
//     rf (one box list PER ENTRY, built in the prologue from the coarse arg-max table); neighbouring coarse
//     queries have neighbouring arg-max keys, so the rectangle is a few windows wide -- and a scattered one only
//     costs tensor MACs, which this stage has to spare (1.97 GFLOP algorithmic);
//   * the mask of a lane is its own window: |ky - cy| <= rf and |kx - cx| <= rf around ITS centre of this entry.
// Zero-padded window positions (affinity 0, value 0 in the reference's F.unfold) are counted analytically and
// inserted by the tail kernel (c2f.cu), which also does the soft-max and the gather of fine labels.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fgvc {
namespace window {

constexpr int T16_STAGE_BYTES = 16 * 1024;     // 2 parts x 64 keys x 128 B (64 channels of fp16)
constexpr int T16_STAGES = 12;
constexpr int T16_MAX_BH = 4;                  // N <= 64, 2N <= 128 accumulator columns
constexpr int T16_AHI_COL = 256, T16_ALO_COL = 384;
constexpr int T16_EPI_WG = 4;
constexpr int T16_THREADS = 64 + 128 * T16_EPI_WG;
constexpr int T16_MAX_BOXES = 4096;             // x 2: all per-entry box lists of a CTA together
constexpr int TW_MAX_MEM = 64;                  // memory entries (the per-entry rectangles live in the aux block)
constexpr int T16_AUX_BYTES = 1024 + 2 * T16_MAX_BOXES * 4;
constexpr int T16_SMEM_BYTES = T16_STAGES * T16_STAGE_BYTES + T16_AUX_BYTES;

struct Tc16Params {
  int H, W, C, n_pix;          // FINE key / query-feature grid
  int HQ, WQ, scale;           // coarse query grid, fine = scale * coarse
  int rf;                      // window radius on the fine grid
  int QH, QW, qw_shift;        // tile of coarse queries
  int BH;
  int k_out;
  int tiles_x;
  fgvc_job job;                // one query frame (fine bank slots)
  const int32_t* mem_feat;
  const int32_t* best;         // [n_mem][HQ * WQ] coarse arg-max key pixel per memory entry
  int chunks;                  // grid.z: the box list of an entry is split into this many CTAs (parallelism: a
                               // 32x32 coarse grid is only 8 tiles)
  float* tv;
  int32_t* ti;
  float* dbg;
  int32_t* dbg_meta;
  int dbg_max_boxes;
  int exp_flags;             // experiments (FGVC_TC16_EXP env): 1 = skip the candidate scan, 2 = skip TMEM loads too
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c,
                                           const uint4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y),
        "r"(c.z), "r"(c.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
      : "memory");
}
// kind::f16 instruction descriptor: c_format F32 = 1 [4,6), a/b_format F16 = 0, K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int K>
__global__ void __launch_bounds__(T16_THREADS, 1)
affinity_window_tc16_kernel(const __grid_constant__ CUtensorMap tmap_k, const __half* __restrict__ bank,
                          const Tc16Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + T16_STAGES * T16_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + T16_STAGES;
  uint64_t* tfull_bar = empty_bar + T16_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint64_t* a_bar = tempty_bar + 2;               // query operand written to TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
  int* rect = reinterpret_cast<int*>(tmem_slot + 4);      // [TW_MAX_MEM][4] y_min, y_max, x_min, x_max of the window centres
  int* boff = rect + 4 * TW_MAX_MEM;                      // [TW_MAX_MEM + 1] first box of each entry's list
  // per-entry box lists (by | bx << 16), back to back
  uint32_t* boxes = reinterpret_cast<uint32_t*>(ring + T16_STAGES * T16_STAGE_BYTES + 1024);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qy0 = (blockIdx.x / p.tiles_x) * p.QH, qx0 = (blockIdx.x % p.tiles_x) * p.QW;
  const fgvc_job job = p.job;
  const int n_mem = job.mem_end - job.mem_begin;
  const int e_lo = job.mem_begin + blockIdx.y, e_hi = e_lo + 1;      // this CTA's memory entry
  const int nq = p.HQ * p.WQ;
  const int N = 16 * p.BH;
  const int n_kc = p.C / 64;
  // one stage = one whole key box (all C channels: n_kc chunks of 16 KB)
  const int stage_bytes = n_kc * T16_STAGE_BYTES;
  const int n_stages = (T16_STAGES * T16_STAGE_BYTES) / stage_bytes;
  const uint32_t stage_tx = (uint32_t)(2 * N * 128 * n_kc);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
    for (int s = 0; s < n_stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar + b, 1); mbar_init(tempty_bar + b, 4 * T16_EPI_WG); }
    mbar_init(a_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- bounding rectangle of the tile's window centres, per memory entry (warps 2-5: lane = query)
  if (warp >= 2 && warp < 6) {
    const int m = (warp - 2) * 32 + lane;
    const int qy = qy0 + (m >> p.qw_shift), qx = qx0 + (m & (p.QW - 1));
    const bool qvalid = qy < p.HQ && qx < p.WQ;
    const int t = blockIdx.y;
    {
      int cy_lo = 1 << 30, cy_hi = -1, cx_lo = 1 << 30, cx_hi = -1;
      if (qvalid) {
        const int b = max(__ldg(p.best + (int64_t)t * nq + qy * p.WQ + qx), 0) % nq;
        cy_lo = cy_hi = (b / p.WQ) * p.scale;
        cx_lo = cx_hi = (b % p.WQ) * p.scale;
      }
      cy_lo = __reduce_min_sync(0xffffffffu, cy_lo); cy_hi = __reduce_max_sync(0xffffffffu, cy_hi);
      cx_lo = __reduce_min_sync(0xffffffffu, cx_lo); cx_hi = __reduce_max_sync(0xffffffffu, cx_hi);
      if (warp == 2 && lane < 4) rect[lane] = lane == 0 ? cy_lo : (lane == 1 ? cy_hi : (lane == 2 ? cx_lo : cx_hi));
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (warp != 2 && lane == 0) {
        atomicMin(rect + 0, cy_lo); atomicMax(rect + 1, cy_hi);
        atomicMin(rect + 2, cx_lo); atomicMax(rect + 3, cx_hi);
      }
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");
    // ---- box list: this CTA's chunk of the boxes of the rectangle grown by rf, clipped to the map (warp 2)
    if (warp == 2) {
      int cnt = 0;
      if (rect[1] >= 0) {                                       // (else: no valid query in the tile)
        const int y_lo = max(0, rect[0] - p.rf), y_hi = min(p.H - 1, rect[1] + p.rf);
        const int x_lo = max(0, rect[2] - p.rf), x_hi = min(p.W - 1, rect[3] + p.rf);
        const int ncols = (x_hi - x_lo) / 16 + 1, nrows = (y_hi - y_lo) / p.BH + 1;
        const int total = nrows * ncols;
        const int i_lo = (int)((int64_t)total * blockIdx.z / p.chunks), i_hi = (int)((int64_t)total * (blockIdx.z + 1) / p.chunks);
        for (int i = i_lo + lane; i < i_hi; i += 32) {
          const int by = y_lo + (i / ncols) * p.BH, bx = x_lo + (i % ncols) * 16;
          if (i - i_lo < 2 * T16_MAX_BOXES) boxes[i - i_lo] = (uint32_t)by | ((uint32_t)bx << 16);
        }
        cnt = min(i_hi - i_lo, 2 * T16_MAX_BOXES);
      }
      if (lane == 0) { boff[0] = 0; boff[1] = cnt; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ================================ TMA producer ====================================
    // one elected lane runs the whole loop (the compiler then keeps everything in uniform registers)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int e = e_hi - 1; e >= e_lo; --e) {       // newest memory frame first: thresholds rise early
        const int slot = p.mem_feat[e] & ~FGVC_MEM_UNMASKED;
        const int b_lo = boff[e - e_lo], b_hi = boff[e - e_lo + 1];
        for (int b = b_lo; b < b_hi; ++b) {
          const uint32_t bb = boxes[b];
          const int by = (int)(bb & 0xffffu), bx = (int)(bb >> 16);
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_expect_tx(full_bar + stage, stage_tx);
          // per 64-channel chunk one TMA box = (64 channels, 16 x BH pixels, both parts): hi rows then lo rows
          for (int kc = 0; kc < n_kc; ++kc)
            tma_load_5d(&tmap_k, full_bar + stage, ring + stage * stage_bytes + kc * T16_STAGE_BYTES, kc * 64, bx, by,
                        0, slot);
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================= MMA issuer =====================================
    if (e_lo < e_hi) {
      mbar_wait(a_bar, 0);
      tc_fence_after();
    }
    const int n_total = boff[1];                     // boxes this CTA processes
    if (elect_one()) {
      const uint32_t idesc2 = make_idesc_f16(128, 2 * N), idesc1 = make_idesc_f16(128, N);
      int stage = 0, buf = 0;
      uint32_t phase = 0, tphase0 = 0, tphase1 = 0;
      const uint32_t ring_u32 = smem_u32(ring);
      const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
      // The barrier probes of box it+1 are issued while the last MMAs of box it are still queued in
      // the tensor pipe, so the pipe does not drain during the ~100-cycle try_wait round trips.
      if (n_total > 0) {
        mbar_wait(tempty_bar + 0, tphase0 ^ 1);
        mbar_wait(full_bar + 0, phase);
        tc_fence_after();
      }
      for (int it = 0; it < n_total; ++it) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
        const uint32_t sa = ring_u32 + (uint32_t)(stage * stage_bytes);
        for (int kc = 0; kc < n_kc; ++kc) {
          const uint64_t b = desc_hi | (uint64_t)((sa + (uint32_t)(kc * T16_STAGE_BYTES)) >> 4);
          const uint32_t a_hi = tmem_base + T16_AHI_COL + kc * 32, a_lo = tmem_base + T16_ALO_COL + kc * 32;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {       // 4 x (K = 16 fp16 = 32 B) per 128 B swizzle row
            if (ks == 3 && kc == n_kc - 1) break;  // the last K step is issued after the probes below
            const uint64_t o = (uint64_t)(ks * 2);
            umma_f16_ts(d_tmem, a_hi + ks * 8, b + o, idesc2, (kc | ks) != 0);   // [hi*hi | hi*lo]
            umma_f16_ts(d_tmem + N, a_lo + ks * 8, b + o, idesc1, 1);            // += lo*hi
          }
        }
        const int nstage = (stage + 1 == n_stages) ? 0 : stage + 1;
        const uint32_t nphase = phase ^ (nstage == 0 ? 1u : 0u);
        const int nbuf = buf ^ 1;
        if (it + 1 < n_total) {
          mbar_wait(tempty_bar + nbuf, (nbuf ? tphase1 : tphase0) ^ 1);   // epilogue drained the other accumulator
          mbar_wait(full_bar + nstage, nphase);                           // next key box landed
          tc_fence_after();
        }
        {
          const int kc = n_kc - 1;
          const uint64_t b = desc_hi | (uint64_t)((sa + (uint32_t)(kc * T16_STAGE_BYTES)) >> 4);
          umma_f16_ts(d_tmem, tmem_base + T16_AHI_COL + kc * 32 + 24, b + 6, idesc2, 1);
          umma_f16_ts(d_tmem + N, tmem_base + T16_ALO_COL + kc * 32 + 24, b + 6, idesc1, 1);
        }
        umma_commit(empty_bar + stage);     // smem stage free once these MMAs retire
        umma_commit(tfull_bar + buf);       // accumulator complete
        if (buf) tphase1 ^= 1; else tphase0 ^= 1;
        buf = nbuf; stage = nstage; phase = nphase;
      }
    }
    __syncwarp();
  } else {
    // ================================== epilogue ======================================
    const int wg = (warp - 2) >> 2;
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const int qy = qy0 + (m >> p.qw_shift), qx = qx0 + (m & (p.QW - 1));      // COARSE query position
    const bool qvalid = qy < p.HQ && qx < p.WQ;
    if (wg == 0 && e_lo < e_hi) {
      // both query parts -> tensor memory, two fp16 channels per 32-bit cell (lower channel in the low half);
      // the row is the FINE query feature at (scale * qy, scale * qx)   (local_attention.py:785)
      const int64_t part = (int64_t)p.n_pix * p.C;
      const __half* row = bank + (int64_t)job.q_slot * 2 * part +
                          (int64_t)(qvalid ? (qy * p.scale) * p.W + qx * p.scale : 0) * p.C;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
      for (int prt = 0; prt < 2; ++prt) {
        const uint4* src = reinterpret_cast<const uint4*>(row + prt * part);
        const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + (prt ? T16_ALO_COL : T16_AHI_COL);
        for (int c = 0; c < p.C / 2; c += 16) {          // 16 cells = 32 channels = 4 x uint4
          uint4 a = z, b = z, c4 = z, d = z;
          if (qvalid) { a = __ldg(src + c / 4); b = __ldg(src + c / 4 + 1); c4 = __ldg(src + c / 4 + 2); d = __ldg(src + c / 4 + 3); }
          tmem_st16u(ta + c, a, b, c4, d);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_bar);
    }
    TopK<K> top;
    top.init();
    int buf = 0;
    uint32_t tph0 = 0, tph1 = 0;
    int box_seq = 0;
    const int row = wg;                                  // the key row of every box this warpgroup owns
    const bool row_ok = qvalid && row < p.BH;
    const uint32_t lane_base = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(row * 16);
    for (int e = e_hi - 1; e >= e_lo; --e) {           // newest memory frame first: thresholds rise early
      const int pos_base = (e - job.mem_begin) * p.n_pix;
      // this lane's window centre in this memory entry: scale * (coarse arg-max key)   (local_attention.py:835-845)
      int cy = 0, cx = 0;
      if (qvalid) {
        const int bq = max(__ldg(p.best + (int64_t)(e - job.mem_begin) * nq + qy * p.WQ + qx), 0) % nq;
        cy = (bq / p.WQ) * p.scale;
        cx = (bq % p.WQ) * p.scale;
      }
      const int b_lo = boff[e - e_lo], b_hi = boff[e - e_lo + 1];
      for (int b = b_lo; b < b_hi; ++b) {
        const uint32_t bb = boxes[b];
        const int by = (int)(bb & 0xffffu), bx = (int)(bb >> 16);
        const int ky = by + row;
        // 16-bit interval mask of the in-window, in-image keys of this key row
        uint32_t bits = 0;
        if (row_ok && ky < p.H && abs(ky - cy) <= p.rf) {
          const int lo = max(max(cx - p.rf, 0) - bx, 0);
          const int hi = min(min(cx + p.rf, p.W - 1) - bx, 15);
          if (hi >= lo) bits = (2u << hi) - (1u << lo);
        }
        const bool dump = p.dbg != nullptr && box_seq < p.dbg_max_boxes && row < p.BH;
        const bool doit = (__any_sync(0xffffffffu, bits != 0) || dump) && !(p.exp_flags & 2);    // warp-uniform
        mbar_wait_sleep(tfull_bar + buf, buf ? tph1 : tph0);
        tc_fence_after();
        uint32_t r1[16], r2[16];
        if (doit) {
          const uint32_t taddr = lane_base + (uint32_t)(buf * 128);
          tmem_ld16_issue(taddr, r1);
          tmem_ld16_issue(taddr + (uint32_t)N, r2);
          tmem_ld_wait(r1);
          reg_fence16(r2);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + buf);    // accumulator is in registers: hand the tile back
        if (buf) tph1 ^= 1; else tph0 ^= 1;
        buf ^= 1;
        if (doit) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r2[j]), FGVC_F16_LO_INV, __uint_as_float(r1[j]));
          if (dump) {
            float* d = p.dbg + ((int64_t)box_seq * 128 + m) * 128 + row * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) d[j] = v[j];
            if (p.dbg_meta != nullptr && m == 0 && wg == 0) {
              p.dbg_meta[4 * box_seq + 0] = e; p.dbg_meta[4 * box_seq + 1] = by;
              p.dbg_meta[4 * box_seq + 2] = bx; p.dbg_meta[4 * box_seq + 3] = N;
            }
          }
          // candidates = in-mask elements above the running K-th value
          const float thr0 = (p.exp_flags & 1) ? INFINITY : top.thr();
          uint32_t cand = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) cand |= (v[j] > thr0) ? (1u << j) : 0u;
          cand &= bits;
          // Insert candidates in warp-wide rounds: in every round each lane that still has a
          // candidate takes its next one, so a round serves ~4 lanes at once instead of one
          // divergent insertion per (lane, element).
          const int kbase = pos_base + ky * p.W + bx;
          while (__any_sync(0xffffffffu, cand != 0)) {
            if (cand) {
              const int j = __ffs(cand) - 1;
              cand &= cand - 1;
              const float x = select16(v, j);
              if (x > top.thr()) top.push(x, kbase + j);
            }
          }
        }
        ++box_seq;
      }
    }
    // ---- merge the partial lists of the warpgroups through the (now idle) ring
    asm volatile("bar.sync 1, %0;" ::"n"(128 * T16_EPI_WG) : "memory");
    float* mv = reinterpret_cast<float*>(ring);
    int* mi = reinterpret_cast<int*>(ring + T16_EPI_WG * 128 * K * 4);
    if (wg > 0) {
#pragma unroll
      for (int i = 0; i < K; ++i) { mv[(wg * 128 + m) * K + i] = top.v[i]; mi[(wg * 128 + m) * K + i] = top.id[i]; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(128 * T16_EPI_WG) : "memory");
    if (wg == 0 && qvalid) {
      for (int w2 = 1; w2 < T16_EPI_WG; ++w2)
        for (int i = 0; i < K; ++i) {
          const float v = mv[(w2 * 128 + m) * K + i];
          if (!(v > top.thr())) break;
          top.push(v, mi[(w2 * 128 + m) * K + i]);
        }
      const int q = qy * p.WQ + qx;
      const int n_lists = n_mem * p.chunks;
      const int64_t o = ((int64_t)q * n_lists + (blockIdx.y * p.chunks + blockIdx.z)) * p.k_out;
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (i < p.k_out) { p.tv[o + i] = top.v[i]; p.ti[o + i] = top.id[i]; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------ host
// 5-D map over feat16[slot][part][H][W][C]; box = (64 channels, 16, bh, both parts, 1), 128B swizzle
static int make_map16(CUtensorMap* map, const void* bank, int n_slots, int H, int W, int C, int bh) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return FGVC_ERR_CUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 2, (cuuint64_t)n_slots};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)2 * H * W * C * 2};
  cuuint32_t box[5] = {64, 16, (cuuint32_t)bh, 2, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(bank), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f16) failed with %d (H=%d W=%d C=%d bh=%d)", (int)r, H, W, C, bh);
    return FGVC_ERR_CUDA;
  }
  return FGVC_OK;
}

// all MMAs are TS-form: a box costs ~N plus a small fixed hand-shake
static int box_cost16(int rows, int bh) { return cdiv(rows, bh) * (16 * bh + 24); }
static int pick_bh16(int rows) {
  int best = T16_MAX_BH;
  for (int bh = T16_MAX_BH - 1; bh >= 1; --bh)
    if (box_cost16(rows, bh) < box_cost16(rows, best)) best = bh;
  return best;
}

template <int K>
static int launch_tc16(const CUtensorMap& mk, const void* bank, const Tc16Params& p, dim3 grid, cudaStream_t st) {
  FGVC_CUDA(cudaFuncSetAttribute(affinity_window_tc16_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 T16_SMEM_BYTES));
  affinity_window_tc16_kernel<K><<<grid, T16_THREADS, T16_SMEM_BYTES, st>>>(mk, reinterpret_cast<const __half*>(bank), p);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

}  // namespace window

bool c2f_window_supported(int Hf, int Wf, int Cf, int K, int n_mem) {
  using namespace window;
  return tc16_supported(Hf, Wf, Cf, K) && n_mem >= 1 && n_mem <= TW_MAX_MEM && Hf < 65536 && Wf < 65536 &&
         (int64_t)n_mem * Hf * Wf < (1ll << 31);
}

// how many CTAs share one (tile, entry): fill the chip, at most 4 (the tail merges n_mem * chunks lists per query)
int c2f_window_chunks(int Hc, int Wc, int n_mem) {
  const long tiles = (long)cdiv(Hc, 8) * cdiv(Wc, 16);
  long c = 148 / (tiles * n_mem > 0 ? tiles * n_mem : 1);
  return (int)(c < 1 ? 1 : (c > 4 ? 4 : c));
}

// fine stage of c2f as a window-mode K1: top-K lists tv / ti [Hc * Wc][K] over the in-window, in-image fine keys
// (idx = memory position * Hf * Wf + fine key pixel).  best: [n_mem][Hc * Wc] coarse arg-max keys.
int launch_c2f_window_tc16(const void* fine_bank, int n_slots, int Hc, int Wc, int Hf, int Wf, int Cf, int scale,
                           const fgvc_job& job, const int32_t* mem_feat, const int32_t* best, int rf, int K, int chunks,
                           float* tv, int32_t* ti, cudaStream_t st) {
  using namespace window;
  Tc16Params p;
  p.H = Hf; p.W = Wf; p.C = Cf; p.n_pix = Hf * Wf;
  p.HQ = Hc; p.WQ = Wc; p.scale = scale; p.rf = rf;
  // coarse tile 8 x 16 or 16 x 8: whichever leaves fewer idle lanes on this grid
  const long waste_a = (long)cdiv(Hc, 8) * cdiv(Wc, 16), waste_b = (long)cdiv(Hc, 16) * cdiv(Wc, 8);
  if (waste_b < waste_a) { p.QH = 16; p.QW = 8; p.qw_shift = 3; }
  else { p.QH = 8; p.QW = 16; p.qw_shift = 4; }
  p.BH = T16_MAX_BH;
  p.k_out = K;
  p.chunks = chunks;
  p.tiles_x = cdiv(Wc, p.QW);
  p.job = job; p.mem_feat = mem_feat; p.best = best; p.tv = tv; p.ti = ti;
  p.dbg = nullptr; p.dbg_meta = nullptr; p.dbg_max_boxes = 0;
  p.exp_flags = 0;
  // worst case (scattered arg-max keys): every entry lists every box of the frame
  const int n_mem = job.mem_end - job.mem_begin;
  if ((int64_t)cdiv(Hf, p.BH) * cdiv(Wf, 16) > 2 * T16_MAX_BOXES) {
    set_error("c2f window engine: a %dx%d map exceeds %d key boxes", Hf, Wf, 2 * T16_MAX_BOXES);
    return FGVC_ERR_UNSUPPORTED;
  }
  CUtensorMap mk;
  int rc = make_map16(&mk, fine_bank, n_slots, Hf, Wf, Cf, p.BH);
  if (rc) return rc;
  dim3 grid(cdiv(Hc, p.QH) * p.tiles_x, n_mem, chunks);
  if (K <= 4) return launch_tc16<4>(mk, fine_bank, p, grid, st);
  if (K <= 10) return launch_tc16<10>(mk, fine_bank, p, grid, st);
  return launch_tc16<16>(mk, fine_bank, p, grid, st);
}

}  // namespace fgvc
