// K1 (tcgen05 engine, fp16 PREFILTER + exact rescoring): EXPERIMENTAL engine of the F16 bank, explicit only
// (FGVC_ENGINE_PREFILTER).  Same results as the exact engines; measured on B200 it is bound by the instruction
// issue of its epilogue, not by the tensor pipe, and is not faster yet (profiles/r1_f_prefilter_engine.md).
//
// The three-term engine (topk_tc16.cu) spends three tensor MACs per fp32-faithful MAC on
// every (query, key) pair, although only ~k of the ~10^4 pairs of a query matter.  This
// engine spends ONE fp16 MAC per pair and repairs the few pairs that matter:
//
//   stage A (this kernel, tensor pipe): a'(q,j) = <hi_q, hi_j>, hi = fp16(x), fp32 accumulate.
//       Rows are unit vectors, so |a' - a| <= sum |x_q x_j| (2 u + u^2) <= 2^-10 (u = 2^-11,
//       Cauchy-Schwarz) plus the fp32 accumulation error of 256 products (< 6e-5):
//       eps = FGVC_PREFILTER_EPS = 1.25e-3 is a rigorous bound with margin.
//   superset: if theta' is the k-th largest a' of a query, every member j of the exact top-k has
//       a'(j) >= theta' - 2 eps.  (The k keys with a' >= theta' have a >= theta' - eps, so the exact
//       k-th value is >= theta' - eps, so a(j) >= theta' - eps, so a'(j) >= theta' - 2 eps.)
//       Every epilogue thread keeps the KP best a' of its share of the keys (MinList: unsorted slots tagged in
//       the low mantissa bits, sorted once at the end); nothing is merged here -- the 4 partial lists per
//       (query, group) go to the workspace.  The 4 threads of a query share lower bounds of theta' through
//       shared memory (same-position seeds + running (k/4)-th values), so that nothing at or below
//       bound - 2 eps is inserted; key rows rotate between the warpgroups from entry to entry, or the best rows
//       of every frame would pile up in one list.
//   stage B (rescore_kernel, one warp per query): theta' from the partial lists, the superset
//       {a' >= theta' - 2 eps} (12 candidates on average for k = 10), the exact value
//       <hi_q + 2^-11 lo_q, hi_j + 2^-11 lo_j> in fp32 for each of them, the exact top-k of those.
//   stage C (exact_scan_kernel): a partial list that is full AND whose last entry is still inside
//       the 2 eps band may have dropped a superset member; such queries (30 of 404 460 on the bench
//       clip at k = 10) are queued by stage B and re-done here by a plain fp32 scan of their keys.
// The result is the exact fp32-faithful top-k -- selection and values -- at a third of the tensor
// work and half the key bytes (only the hi part of a key box is ever staged).
//
// CTA anatomy as in topk_tc16.cu: warp 0 TMA producer, warp 1 MMA issuer (TS-form, hi_q resident in
// tensor memory), warps 2-3 set-up, then 4 epilogue warpgroups (640 threads; setmaxnreg moves registers from
// warpgroup 0 to the epilogue), thread = query.  Key boxes are 16 x BH pixels with BH <= 8
// (N <= 128): one MMA per 16 channels, two 128-column accumulators ping-pong.
// TMEM: [0,128) accumulator 0, [128,256) accumulator 1, [256, 256 + C/2) hi_q.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fgvc {

constexpr int TP_RING_BYTES = 192 * 1024;
constexpr int TP_MAX_STAGES = 12;
constexpr int TP_MAX_BH = 8;                   // N <= 128
constexpr int TP_A_COL = 256;
constexpr int TP_EPI_WG = 4;
constexpr int TP_THREADS = 128 + 128 * TP_EPI_WG;   // warpgroup 0 = {TMA producer, MMA issuer, 2 set-up warps}, then 4 epilogue warpgroups
// Registers: the launch gives every thread 96 (640 x 96 <= 64 K); warpgroup 0 then hands most of its share
// back (setmaxnreg.dec) and the epilogue warpgroups, which hold 32 accumulator values + the list, grow to 112:
// 128 x 32 + 512 x 112 = 61440 = 640 x 96 (the pool setmaxnreg.inc draws from is what the launch allocated).
constexpr int TP_REGS_CTRL = 32, TP_REGS_EPI = 112;
constexpr int TP_MAX_BOXES = 3072;
constexpr int TP_THR_BYTES = 2 * TP_EPI_WG * 128 * 4;   // seed / running threshold exchange: [2][4 warpgroups][128 queries]
constexpr int TP_AUX_BYTES = 1024 + TP_THR_BYTES + 2 * TP_MAX_BOXES * 4;
constexpr int TP_SMEM_BYTES = TP_RING_BYTES + TP_AUX_BYTES;
constexpr int TP_MAX_CAND = 64;                // superset capacity of stage B (per query)

struct TcpParams {
  int H, W, C, n_pix;
  int radius, mode, reach;
  int QH, QW, qw_shift;
  int BH;
  int groups;
  int tiles_x;
  const fgvc_job* jobs;
  const int32_t* mem_feat;
  float* cv;                 // [job][group][n_pix][4][KP] approximate values
  int32_t* ci;               // same, candidate index (position * n_pix + key pixel), -1 = empty
  unsigned long long* stats;  // optional (exp_flags & 8): row scans, hot row scans, insertion rounds, insertions
  int exp_flags;             // experiments (FGVC_TCP_EXP env): 1 = skip the candidate scan, 2 = skip TMEM loads too, 4 = no seeds, 8 = statistics, 16 = entry-major box order
};

__device__ __forceinline__ void umma_f16_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16u_p(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c,
                                             const uint4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y),
        "r"(c.z), "r"(c.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
      : "memory");
}
// kind::f16 instruction descriptor: c_format F32 = 1 [4,6), a/b_format F16 = 0, K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_f16_p(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max16(const float* v) {
  const float a = fmax3(v[0], v[1], v[2]), b = fmax3(v[3], v[4], v[5]), c = fmax3(v[6], v[7], v[8]);
  const float d = fmax3(v[9], v[10], v[11]), e = fmax3(v[12], v[13], v[14]);
  return fmaxf(fmax3(a, b, c), fmax3(d, e, v[15]));
}

// The epilogue's per-thread candidate list.  The epilogue is instruction-bound (every key row is "hot" for some of
// the 32 queries of a warp), so the list is built for the cheapest possible insertion, not for order:
//   * KP slots, UNSORTED; the low 4 mantissa bits of every stored value hold its slot number, so the minimum of the
//     KP values (a 3-input min tree) is at once the K-th value and the slot the next insertion overwrites;
//   * the 4 best values seen so far are kept on the side (a min/max cascade) for the bound shared between warpgroups;
//   * one odd-even sort when the CTA is done.
// Dropping 4 mantissa bits moves a' by <= 2^-19 relative, far inside the margin of FGVC_PREFILTER_EPS.
constexpr uint32_t TP_NEG_BITS = 0xFF7FFFF0u;          // most negative finite float with the low 4 bits clear

template <int KP>
struct MinList {
  float v[KP];
  int id[KP];
  float thr;                 // min of v[]: the list's KP-th value (slot in the low 4 bits)
  float t0, t1, t2, t3;      // best four values, descending
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KP; ++s) { v[s] = __uint_as_float(TP_NEG_BITS | (uint32_t)s); id[s] = -1; }
    thr = v[0];
    t0 = t1 = t2 = t3 = __uint_as_float(TP_NEG_BITS);
  }
  // caller guarantees x > thr
  __device__ __forceinline__ void push(float x, int idx) {
    const uint32_t slot = __float_as_uint(thr) & 15u;
    const uint32_t xe = __float_as_uint(x) & ~15u;
#pragma unroll
    for (int s = 0; s < KP; ++s) {
      const bool h = slot == (uint32_t)s;
      v[s] = h ? __uint_as_float(xe | (uint32_t)s) : v[s];
      id[s] = h ? idx : id[s];
    }
    float m = v[0];
#pragma unroll
    for (int s = 1; s + 1 < KP; s += 2) m = fmin3(m, v[s], v[s + 1]);
    if ((KP & 1) == 0) m = fminf(m, v[KP - 1]);
    thr = m;
    float y, z = __uint_as_float(xe);
    y = fminf(t0, z); t0 = fmaxf(t0, z); z = y;
    y = fminf(t1, z); t1 = fmaxf(t1, z); z = y;
    y = fminf(t2, z); t2 = fmaxf(t2, z); z = y;
    t3 = fmaxf(t3, z);
  }
  __device__ __forceinline__ void sort_desc() {          // odd-even transposition, KP rounds
#pragma unroll
    for (int r = 0; r < KP; ++r) {
#pragma unroll
      for (int s = (r & 1); s + 1 < KP; s += 2) {
        const bool sw = v[s] < v[s + 1];
        const float a = v[s], b = v[s + 1];
        const int ia = id[s], ib = id[s + 1];
        v[s] = sw ? b : a; v[s + 1] = sw ? a : b;
        id[s] = sw ? ib : ia; id[s + 1] = sw ? ia : ib;
      }
    }
  }
};

// Fold one key row (16 accumulator columns) into the thread's list.  The column number goes into the low 4 bits of
// every value, so a 3-input max tree yields the row's best value AND where it is; further candidates of the same
// row are found by repeating the tree over the values strictly below the last one.  `floor` = the query's prune
// level (a lower bound of theta' minus the band): nothing at or below it can belong to the superset.  `bits` = the
// in-mask, in-image columns of this key row.  Returns true when this lane inserted something.
template <int KP>
__device__ __forceinline__ bool scan_row(MinList<KP>& top, const uint32_t* r, uint32_t bits, int kbase, float floor,
                                         bool qvalid, int& n_rounds, int& n_ins) {
  float enc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) enc[j] = __uint_as_float((r[j] & ~15u) | (uint32_t)j);
  float mx = max16(enc);
  bool changed = false;
  float cur = INFINITY;
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {                          // a row has 16 values: at most 16 extractions
    const bool hot = qvalid && mx > fmaxf(top.thr, floor);
    if (!__any_sync(0xffffffffu, hot)) break;
    ++n_rounds;
    if (hot) {
      const int j = (int)(__float_as_uint(mx) & 15u);
      if ((bits >> j) & 1u) { top.push(mx, kbase + j); changed = true; ++n_ins; }
      cur = mx;
    }
    float e2[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) e2[j] = enc[j] < cur ? enc[j] : __uint_as_float(TP_NEG_BITS);
    mx = max16(e2);
  }
  return changed;
}

// The order in which the three warp roles walk the (memory entry, key box) pairs of a CTA.
//   masked entries  : BOX-major -- box 0 (the one nearest to the query tile) of every entry, newest entry first, then
//                     box 1 of every entry, ...  The best matches of a query sit near its own position in EVERY
//                     frame, so after the first ring of boxes the lists already hold the best candidates of all
//                     frames, the shared bounds are tight, and the remaining ~3/4 of the boxes insert next to nothing.
//                     (Entry-major order only converges after about half the frames.)
//   unmasked entries: entry-major over the whole-frame list, afterwards.
// box_major = false gives the plain entry-major order (experiments).
template <class F>
__device__ __forceinline__ void walk_boxes(const int32_t* __restrict__ mem_feat, int e_lo, int e_hi, const int* nbox,
                                           const uint32_t* boxes, bool box_major, F&& body) {
  if (box_major) {
    const int nb0 = nbox[0];
    for (int b = 0; b < nb0; ++b) {
      const uint32_t bb = boxes[b];
      for (int e = e_hi - 1; e >= e_lo; --e) {
        const int raw = mem_feat[e];
        if (raw & FGVC_MEM_UNMASKED) continue;
        body(e, raw, bb);
      }
    }
    const int nb1 = nbox[1];
    for (int e = e_hi - 1; e >= e_lo; --e) {
      const int raw = mem_feat[e];
      if (!(raw & FGVC_MEM_UNMASKED)) continue;
      for (int b = 0; b < nb1; ++b) body(e, raw, boxes[TP_MAX_BOXES + b]);
    }
  } else {
    for (int e = e_hi - 1; e >= e_lo; --e) {
      const int raw = mem_feat[e];
      const int li = (raw & FGVC_MEM_UNMASKED) ? 1 : 0;
      const int nb = nbox[li];
      for (int b = 0; b < nb; ++b) body(e, raw, boxes[li * TP_MAX_BOXES + b]);
    }
  }
}

template <int KP>
__global__ void __launch_bounds__(TP_THREADS, 1)
affinity_prefilter_tc16_kernel(const __grid_constant__ CUtensorMap tmap_k, const __half* __restrict__ bank,
                               const TcpParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + TP_RING_BYTES);
  uint64_t* empty_bar = full_bar + TP_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + TP_MAX_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2]
  uint64_t* a_bar = tempty_bar + 2;                  // query operand written to TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
  int* nbox = reinterpret_cast<int*>(tmem_slot + 2);      // [2] number of boxes in each list
  int* halfw = reinterpret_cast<int*>(tmem_slot + 4);     // [reach+1] <= 128 entries
  // box lists (by | bx << 16): [0] = radius halo of this query tile minus boxes no query can see,
  // [1] = every box of the frame (unmasked memory entries)
  float* s_seed = reinterpret_cast<float*>(ring + TP_RING_BYTES + 1024);   // [4][128] per-warpgroup seed bounds
  float* s_run = s_seed + TP_EPI_WG * 128;                                 // [4][128] per-warpgroup running bounds
  uint32_t* boxes = reinterpret_cast<uint32_t*>(ring + TP_RING_BYTES + 1024 + TP_THR_BYTES);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qy0 = (blockIdx.x / p.tiles_x) * p.QH, qx0 = (blockIdx.x % p.tiles_x) * p.QW;
  const int g = blockIdx.y;
  const fgvc_job job = p.jobs[blockIdx.z];
  const int n_mem = job.mem_end - job.mem_begin;
  const int per = (n_mem + p.groups - 1) / p.groups;
  const int e_lo = job.mem_begin + g * per;
  const int e_hi = min(job.mem_end, e_lo + per);
  const int N = 16 * p.BH;
  const int n_kc = p.C / 64;
  const int chunk_bytes = N * 128;                     // N keys x 64 channels of fp16
  const int stage_bytes = n_kc * chunk_bytes;          // one stage = one whole key box (all C channels)
  const int n_stages = min(TP_MAX_STAGES, TP_RING_BYTES / stage_bytes);
  const uint32_t stage_tx = (uint32_t)stage_bytes;
  const bool box_major = !(p.exp_flags & 16);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_k) : "memory");
    for (int s = 0; s < n_stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar + b, 1); mbar_init(tempty_bar + b, 4 * TP_EPI_WG); }
    mbar_init(a_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int d = threadIdx.x; d <= p.reach; d += TP_THREADS) {
    int hw = -1;
    if (p.mode == FGVC_MASK_CIRCLE) {
      while (hw + 1 <= p.reach && (hw + 1) * (hw + 1) + d * d < p.radius * p.radius) ++hw;
    } else {
      hw = p.radius;
    }
    halfw[d] = hw;
  }
  if (warp == 2 || warp == 3) {          // one warp per list
    const int li = warp - 2;
    const Walk w = make_walk(p, li ? FGVC_MEM_UNMASKED : 0, qy0, qx0);
    const int ncols = (w.x_hi - w.x_lo) / 16 + 1, nrows = (w.y_hi - w.y_lo) / p.BH + 1;
    uint32_t* list = boxes + li * TP_MAX_BOXES;
    int cnt = 0;
    for (int base = 0; base < nrows * ncols; base += 32) {
      const int i = base + lane;
      const int by = w.y_lo + (i / ncols) * p.BH, bx = w.x_lo + (i % ncols) * 16;
      const bool keep = i < nrows * ncols && !box_skipped(p, w, by, bx, qy0, qx0);
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
      if (keep && pos < TP_MAX_BOXES) list[pos] = (uint32_t)by | ((uint32_t)bx << 16);
      cnt += __popc(bal);
    }
    cnt = min(cnt, TP_MAX_BOXES);
    __syncwarp();
    // centre-out order for the halo list: the best matches of a query sit near its own position, so
    // the running thresholds rise early and list insertions become rare.  Rank sort, n is a few dozen.
    if (li == 0 && cnt > 1 && cnt <= 128) {
      const int cy2 = 2 * qy0 + p.QH, cx2 = 2 * qx0 + p.QW;          // twice the tile centre
      uint32_t mine[4]; int rank[4];
      for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * t;
        mine[t] = i < cnt ? list[i] : 0u;
        rank[t] = 0;
      }
      for (int j = 0; j < cnt; ++j) {
        const uint32_t o = list[j];
        const int oy = 2 * (int)(o & 0xffffu) + p.BH - cy2, ox = 2 * (int)(o >> 16) + 16 - cx2;
        const int od = oy * oy + ox * ox;
        for (int t = 0; t < 4; ++t) {
          const int i = lane + 32 * t;
          const int my = 2 * (int)(mine[t] & 0xffffu) + p.BH - cy2, mx = 2 * (int)(mine[t] >> 16) + 16 - cx2;
          const int md = my * my + mx * mx;
          rank[t] += (od < md || (od == md && j < i)) ? 1 : 0;
        }
      }
      __syncwarp();
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < cnt) list[rank[t]] = mine[t];
    }
    if (lane == 0) nbox[li] = cnt;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TP_REGS_CTRL));
  if (warp == 0) {
    // ================================ TMA producer ====================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      walk_boxes(p.mem_feat, e_lo, e_hi, nbox, boxes, box_major, [&](int e, int raw, uint32_t bb) {
        const int slot = raw & ~FGVC_MEM_UNMASKED;
        const int by = (int)(bb & 0xffffu), bx = (int)(bb >> 16);
        mbar_wait(empty_bar + stage, phase ^ 1);
        mbar_expect_tx(full_bar + stage, stage_tx);
        // per 64-channel chunk one TMA box = (64 channels, 16 x BH pixels) of the hi part only
        for (int kc = 0; kc < n_kc; ++kc)
          tma_load_5d(&tmap_k, full_bar + stage, ring + stage * stage_bytes + kc * chunk_bytes, kc * 64, bx, by, 0,
                      slot);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      });
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================= MMA issuer =====================================
    if (e_lo < e_hi) {
      mbar_wait(a_bar, 0);
      tc_fence_after();
    }
    int n_total = 0;                                 // boxes this CTA processes
    for (int e = e_lo; e < e_hi; ++e) n_total += nbox[(p.mem_feat[e] & FGVC_MEM_UNMASKED) ? 1 : 0];
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16_p(128, N);
      int stage = 0, buf = 0;
      uint32_t phase = 0, tphase0 = 0, tphase1 = 0;
      const uint32_t ring_u32 = smem_u32(ring);
      const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
      // the barrier probes of box it+1 are issued while the last MMA of box it is still queued
      if (n_total > 0) {
        mbar_wait(tempty_bar + 0, tphase0 ^ 1);
        mbar_wait(full_bar + 0, phase);
        tc_fence_after();
      }
      for (int it = 0; it < n_total; ++it) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
        const uint32_t sa = ring_u32 + (uint32_t)(stage * stage_bytes);
        for (int kc = 0; kc < n_kc; ++kc) {
          const uint64_t b = desc_hi | (uint64_t)((sa + (uint32_t)(kc * chunk_bytes)) >> 4);
          const uint32_t a_hi = tmem_base + TP_A_COL + kc * 32;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {       // 4 x (K = 16 fp16 = 32 B) per 128 B swizzle row
            if (ks == 3 && kc == n_kc - 1) break;  // the last K step is issued after the probes below
            umma_f16_ts_p(d_tmem, a_hi + ks * 8, b + (uint64_t)(ks * 2), idesc, (kc | ks) != 0);
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
          const uint64_t b = desc_hi | (uint64_t)((sa + (uint32_t)(kc * chunk_bytes)) >> 4);
          umma_f16_ts_p(d_tmem, tmem_base + TP_A_COL + kc * 32 + 24, b + 6, idesc, 1);
        }
        umma_commit(empty_bar + stage);     // smem stage free once these MMAs retire
        umma_commit(tfull_bar + buf);       // accumulator complete
        if (buf) tphase1 ^= 1; else tphase0 ^= 1;
        buf = nbuf; stage = nstage; phase = nphase;
      }
    }
    __syncwarp();
  }
  } else {
    // ================================== epilogue ======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TP_REGS_EPI));
    const int wg = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int m = lg * 32 + lane;
    const int qy = qy0 + (m >> p.qw_shift), qx = qx0 + (m & (p.QW - 1));
    const bool qvalid = qy < p.H && qx < p.W;
    if (wg == 0 && e_lo < e_hi) {
      // hi part of the query tile -> tensor memory, two fp16 channels per 32-bit cell
      const __half* row = bank + (int64_t)job.q_slot * 2 * p.n_pix * p.C + (int64_t)(qvalid ? qy * p.W + qx : 0) * p.C;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      const uint4* src = reinterpret_cast<const uint4*>(row);
      const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + TP_A_COL;
      for (int c = 0; c < p.C / 2; c += 16) {          // 16 cells = 32 channels = 4 x uint4
        uint4 a = z, b = z, c4 = z, d = z;
        if (qvalid) { a = __ldg(src + c / 4); b = __ldg(src + c / 4 + 1); c4 = __ldg(src + c / 4 + 2); d = __ldg(src + c / 4 + 3); }
        tmem_st16u_p(ta + c, a, b, c4, d);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_bar);
    }
    MinList<KP> top;
    top.init();
    // ---- thresholds shared by the 4 threads of a query.  If warpgroup w publishes a value b_w such that
    // at least SH + 1 distinct candidates of ITS OWN share have a' >= b_w, then min_w b_w has
    // 4 (SH + 1) >= k candidates above it, i.e. it is a lower bound of theta'.  Two such families:
    //   seeds   : a' at the query's own position in the memory entries e = w (mod 4), computed here on the
    //             CUDA cores before the first box arrives (temporal coherence makes this bound tight at once);
    //   running : the (SH + 1)-th entry of the warpgroup's list, republished whenever the list changes.
    // Everything at or below  max(min_w seed_w, min_w run_w) - 2 eps  is outside the superset and is not inserted.
    constexpr int SH = (KP == 4) ? 0 : (KP == 10 ? 2 : 3);
    float seed_b = -INFINITY;
    if (qvalid && !(p.exp_flags & 4)) {
      const int64_t part2 = 2 * (int64_t)p.n_pix * p.C;
      const int64_t pix_off = (int64_t)(qy * p.W + qx) * p.C;
      const uint4* qrow = reinterpret_cast<const uint4*>(bank + (int64_t)job.q_slot * part2 + pix_off);
      float s0 = -INFINITY, s1 = -INFINITY, s2 = -INFINITY, s3 = -INFINITY;       // the thread's 4 best seeds, descending
      for (int e0 = e_hi - 1 - wg; e0 >= e_lo; e0 -= 4 * 3) {                       // three entries at a time
        const uint4* k0 = reinterpret_cast<const uint4*>(bank + (int64_t)(p.mem_feat[e0] & ~FGVC_MEM_UNMASKED) * part2 + pix_off);
        const bool h1 = e0 - 4 >= e_lo, h2 = e0 - 8 >= e_lo;
        const uint4* k1 = h1 ? reinterpret_cast<const uint4*>(bank + (int64_t)(p.mem_feat[e0 - 4] & ~FGVC_MEM_UNMASKED) * part2 + pix_off) : k0;
        const uint4* k2 = h2 ? reinterpret_cast<const uint4*>(bank + (int64_t)(p.mem_feat[e0 - 8] & ~FGVC_MEM_UNMASKED) * part2 + pix_off) : k0;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int c = 0; c < p.C / 8; ++c) {
          const uint4 qv = __ldg(qrow + c), x0 = __ldg(k0 + c), x1 = __ldg(k1 + c), x2 = __ldg(k2 + c);
          const __half2* qh = reinterpret_cast<const __half2*>(&qv);
          const __half2* h0 = reinterpret_cast<const __half2*>(&x0);
          const __half2* h1p = reinterpret_cast<const __half2*>(&x1);
          const __half2* h2p = reinterpret_cast<const __half2*>(&x2);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 qf = __half22float2(qh[i]);
            const float2 f0 = __half22float2(h0[i]), f1 = __half22float2(h1p[i]), f2 = __half22float2(h2p[i]);
            a0 = fmaf(qf.x, f0.x, a0); a0 = fmaf(qf.y, f0.y, a0);
            a1 = fmaf(qf.x, f1.x, a1); a1 = fmaf(qf.y, f1.y, a1);
            a2 = fmaf(qf.x, f2.x, a2); a2 = fmaf(qf.y, f2.y, a2);
          }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          float x = t == 0 ? a0 : (t == 1 ? (h1 ? a1 : -INFINITY) : (h2 ? a2 : -INFINITY));
          float y;                                         // sorted insert into (s0 >= s1 >= s2 >= s3)
          y = fminf(s0, x); s0 = fmaxf(s0, x); x = y;
          y = fminf(s1, x); s1 = fmaxf(s1, x); x = y;
          y = fminf(s2, x); s2 = fmaxf(s2, x); x = y;
          s3 = fmaxf(s3, x);
        }
      }
      seed_b = (SH == 0 ? s0 : (SH == 2 ? s2 : s3)) - 1e-5f;   // summation order differs from the tensor pipe's
    }
    s_seed[wg * 128 + m] = seed_b;
    s_run[wg * 128 + m] = -INFINITY;
    asm volatile("bar.sync 1, %0;" ::"n"(128 * TP_EPI_WG) : "memory");
    const float band = 2.f * FGVC_PREFILTER_EPS;
    const float seed_floor = fminf(fminf(s_seed[m], s_seed[128 + m]), fminf(s_seed[256 + m], s_seed[384 + m])) - band;
    float floor_q = seed_floor;
    int st_rows = 0, st_hot = 0, st_rounds = 0, st_ins = 0;
    uint32_t cached_bb = 0xffffffffu, mpk[4] = {0u, 0u, 0u, 0u};
    int buf = 0;
    uint32_t tph0 = 0, tph1 = 0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(lg * 32) << 16);
    const bool skip_scan = (p.exp_flags & 1) != 0;
    const bool skip_ld = (p.exp_flags & 2) != 0;
    // the warp's queries span the tile rows [wqy_lo, wqy_hi]
    const int wqy_lo = qy0 + ((lg * 32) >> p.qw_shift), wqy_hi = qy0 + ((lg * 32 + 31) >> p.qw_shift);
    const bool warp_valid = wqy_lo < p.H;
    walk_boxes(p.mem_feat, e_lo, e_hi, nbox, boxes, box_major, [&](int e, int raw, uint32_t bb) {
      {
        const bool masked = !(raw & FGVC_MEM_UNMASKED);
        const int pos_base = (e - job.mem_begin) * p.n_pix;
        // Warpgroup w owns key rows (w + e) % 4 and that + 4 of every box of entry e.  The rotation matters: the best
        // matches of a query sit on the same few key rows in every frame, and without it they would all pile up in
        // one warpgroup's list (and overflow it into the exact scan).
        const int row0 = (wg + e) & 3, row1 = row0 + 4;
        const int by = (int)(bb & 0xffffu), bx = (int)(bb >> 16);
        const int ky0 = by + row0, ky1 = by + row1;
        // warp-uniform: can any query of this warp see the key row at all?
        const bool do0 = !skip_ld && warp_valid && row0 < p.BH && ky0 < p.H &&
                         (!masked || (ky0 >= wqy_lo - p.reach && ky0 <= wqy_hi + p.reach));
        const bool do1 = !skip_ld && warp_valid && row1 < p.BH && ky1 < p.H &&
                         (!masked || (ky1 >= wqy_lo - p.reach && ky1 <= wqy_hi + p.reach));
        mbar_wait_sleep(tfull_bar + buf, buf ? tph1 : tph0);
        tc_fence_after();
        uint32_t r0[16], r1[16];
        const uint32_t taddr = lane_base + (uint32_t)(buf * 128);
        if (do0) tmem_ld16_issue(taddr + (uint32_t)(row0 * 16), r0);
        if (do1) tmem_ld16_issue(taddr + (uint32_t)(row1 * 16), r1);
        if (do0) tmem_ld_wait(r0);
        if (do1) tmem_ld_wait(r1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + buf);    // accumulator is in registers: hand the tile back
        if (buf) tph1 ^= 1; else tph0 ^= 1;
        buf ^= 1;
        if (skip_scan) return;
        // refresh the query's prune level from the running bounds of the 4 warpgroups (stale values are
        // merely conservative: the bounds only rise)
        {
          const float run = fminf(fminf(s_run[m], s_run[128 + m]), fminf(s_run[256 + m], s_run[384 + m])) - band;
          floor_q = fmaxf(seed_floor, run);
        }
        // in-mask, in-image columns of the thread's two key rows.  For masked entries the 8 row masks of a box
        // depend on the box only, and in box-major order a box is visited once per entry: computed on box change.
        uint32_t bits0, bits1;
        if (masked) {
          if (bb != cached_bb) {
            cached_bb = bb;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t pk = 0;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int ky = by + 2 * i + h;
                uint32_t bits = 0;
                if (qvalid && 2 * i + h < p.BH && ky < p.H) {
                  const int ady = abs(ky - qy);
                  const int hw = ady <= p.reach ? halfw[ady] : -1;
                  int lo = hw < 0 ? 1 : max(qx - hw, 0);
                  int hi = hw < 0 ? 0 : min(qx + hw, p.W - 1);
                  lo = max(lo - bx, 0);
                  hi = min(hi - bx, 15);
                  if (hi >= lo) bits = (2u << hi) - (1u << lo);
                }
                pk |= bits << (16 * h);
              }
              mpk[i] = pk;
            }
          }
          const uint32_t lo_pair = (row0 & 2) ? mpk[1] : mpk[0], hi_pair = (row0 & 2) ? mpk[3] : mpk[2];
          bits0 = (lo_pair >> (16 * (row0 & 1))) & 0xffffu;
          bits1 = (hi_pair >> (16 * (row0 & 1))) & 0xffffu;
        } else {
          const int wcols = min(16, p.W - bx);
          const uint32_t rowbits = qvalid ? ((1u << wcols) - 1u) : 0u;
          bits0 = ky0 < p.H ? rowbits : 0u;
          bits1 = ky1 < p.H ? rowbits : 0u;
        }
        bool changed = false;
        int rounds_before = st_rounds;
        if (do0) { ++st_rows; changed |= scan_row<KP>(top, r0, bits0, pos_base + ky0 * p.W + bx, floor_q, qvalid, st_rounds, st_ins); }
        st_hot += st_rounds > rounds_before ? 1 : 0;
        rounds_before = st_rounds;
        if (do1) { ++st_rows; changed |= scan_row<KP>(top, r1, bits1, pos_base + ky1 * p.W + bx, floor_q, qvalid, st_rounds, st_ins); }
        st_hot += st_rounds > rounds_before ? 1 : 0;
        if (changed) s_run[wg * 128 + m] = (SH == 0 ? top.t0 : (SH == 2 ? top.t2 : top.t3));
      }
    });
    if (p.exp_flags & 8) {
      const int ins = __reduce_add_sync(0xffffffffu, st_ins);
      if (lane == 0) {
        atomicAdd(p.stats + 0, (unsigned long long)st_rows);
        atomicAdd(p.stats + 1, (unsigned long long)st_hot);
        atomicAdd(p.stats + 2, (unsigned long long)st_rounds);
        atomicAdd(p.stats + 3, (unsigned long long)ins);
      }
    }
    // ---- the partial lists go out through the (now idle) ring so that the stores are coalesced:
    // workspace layout [job][group][query][warpgroup][KP]; a tile row of QW queries is contiguous
    top.sort_desc();
    asm volatile("bar.sync 1, %0;" ::"n"(128 * TP_EPI_WG) : "memory");
    float* mv = reinterpret_cast<float*>(ring);
    int* mi = reinterpret_cast<int*>(ring + 128 * TP_EPI_WG * KP * 4);
#pragma unroll
    for (int i = 0; i < KP; ++i) { mv[(m * TP_EPI_WG + wg) * KP + i] = top.v[i]; mi[(m * TP_EPI_WG + wg) * KP + i] = top.id[i]; }
    asm volatile("bar.sync 1, %0;" ::"n"(128 * TP_EPI_WG) : "memory");
    const int et = threadIdx.x - 128;                       // 0 .. 511
    const int per_q = TP_EPI_WG * KP;
    const int qw_valid = min(p.QW, p.W - qx0);
    const int row_elems = qw_valid * per_q;
    const int64_t jg = ((int64_t)blockIdx.z * p.groups + g) * p.n_pix;
    for (int ty = 0; ty < p.QH; ++ty) {
      if (qy0 + ty >= p.H) break;
      const int64_t o = (jg + (int64_t)(qy0 + ty) * p.W + qx0) * per_q;
      const int s = ty * p.QW * per_q;
      for (int i = et; i < row_elems; i += 128 * TP_EPI_WG) {
        p.cv[o + i] = mv[s + i];
        p.ci[o + i] = mi[s + i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ stage B: rescoring
struct RescoreParams {
  int C, n_pix, groups, k_out, n_jobs;
  const fgvc_job* jobs;
  const int32_t* mem_feat;
  const float* cv;
  const int32_t* ci;
  float* tv;                 // [job][group][n_pix][k_out]; the result goes to group 0, the others are emptied
  int32_t* ti;
  float band;                // 2 eps
  int32_t* ovf_count;        // queue of queries for stage C
  int32_t* ovf_list;
  int ovf_cap;
};

// NH2 half2 cells per lane (C = 64 * NH2): lane l holds channels [2 NH2 l, 2 NH2 (l + 1))
template <int NH2>
__device__ __forceinline__ void load_row_f32(const __half* __restrict__ hi, const __half* __restrict__ lo, int lane,
                                             float* x) {
  __half2 h[NH2], l[NH2];
  if (NH2 == 4) {
    *reinterpret_cast<uint4*>(h) = __ldg(reinterpret_cast<const uint4*>(hi) + lane);
    *reinterpret_cast<uint4*>(l) = __ldg(reinterpret_cast<const uint4*>(lo) + lane);
  } else if (NH2 == 2) {
    *reinterpret_cast<uint2*>(h) = __ldg(reinterpret_cast<const uint2*>(hi) + lane);
    *reinterpret_cast<uint2*>(l) = __ldg(reinterpret_cast<const uint2*>(lo) + lane);
  } else {
#pragma unroll
    for (int i = 0; i < NH2; ++i) {
      *reinterpret_cast<uint32_t*>(&h[i]) = __ldg(reinterpret_cast<const uint32_t*>(hi) + lane * NH2 + i);
      *reinterpret_cast<uint32_t*>(&l[i]) = __ldg(reinterpret_cast<const uint32_t*>(lo) + lane * NH2 + i);
    }
  }
#pragma unroll
  for (int i = 0; i < NH2; ++i) {
    const float2 a = __half22float2(h[i]), b = __half22float2(l[i]);
    x[2 * i] = fmaf(b.x, FGVC_F16_LO_INV, a.x);
    x[2 * i + 1] = fmaf(b.y, FGVC_F16_LO_INV, a.y);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// this lane's share of the exact fp32 value of <query row, key row>
template <int NH2>
__device__ __forceinline__ float partial_dot(const __half* __restrict__ bank, const float* xq, int slot, int pix,
                                             int n_pix, int C, int lane) {
  const __half* hi = bank + ((int64_t)slot * 2 * n_pix + pix) * C;
  float xk[2 * NH2];
  load_row_f32<NH2>(hi, hi + (int64_t)n_pix * C, lane, xk);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * NH2; ++i) s = fmaf(xq[i], xk[i], s);
  return s;
}
// exact fp32 value of <query row, key row idx> (all lanes return the same sum)
template <int NH2>
__device__ __forceinline__ float exact_dot(const __half* __restrict__ bank, const float* xq, int slot, int pix,
                                           int n_pix, int C, int lane) {
  return warp_sum(partial_dot<NH2>(bank, xq, slot, pix, n_pix, C, lane));
}

constexpr int RS_WARPS = 8;

template <int KP, int NH2>
__global__ void __launch_bounds__(32 * RS_WARPS)
rescore_kernel(const __half* __restrict__ bank, const RescoreParams p) {
  __shared__ float s_v[RS_WARPS][32 * KP];
  __shared__ int s_i[RS_WARPS][32 * KP];
  __shared__ int s_c[RS_WARPS][TP_MAX_CAND];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * RS_WARPS + warp;
  if (gw >= (int64_t)p.n_jobs * p.n_pix) return;
  const int jidx = (int)(gw / p.n_pix), q = (int)(gw - (int64_t)jidx * p.n_pix);
  const fgvc_job job = p.jobs[jidx];
  const int n_lists = p.groups * TP_EPI_WG;                 // <= 32: lane l owns partial list l
  float* lv = s_v[warp];
  int* li = s_i[warp];
  // lists of group g, warpgroup w: ws[((job * groups + g) * n_pix + q) * 4 + w][KP]; entries of a group are contiguous
  for (int g = 0; g < p.groups; ++g) {
    const int64_t o = (((int64_t)jidx * p.groups + g) * p.n_pix + q) * (TP_EPI_WG * KP);
    for (int i = lane; i < TP_EPI_WG * KP; i += 32) {
      lv[g * TP_EPI_WG * KP + i] = __ldg(p.cv + o + i);
      li[g * TP_EPI_WG * KP + i] = __ldg(p.ci + o + i);
    }
  }
  __syncwarp();
  // theta' = k-th largest approximate value of the union: k rounds over the heads of the sorted lists
  int head = 0;
  float theta = -INFINITY;
  for (int r = 0; r < p.k_out; ++r) {
    const bool have = lane < n_lists && head < KP && li[lane * KP + head] >= 0;
    const float x = have ? lv[lane * KP + head] : -INFINITY;
    const float mx = warp_max(x);
    theta = mx;
    if (mx == -INFINITY) break;
    const uint32_t who = __ballot_sync(0xffffffffu, have && x == mx);
    if (lane == __ffs(who) - 1) ++head;
  }
  const float thr = theta - p.band;                          // -inf when the query has fewer than k candidates
  int cnt = 0;
  if (lane < n_lists) {
#pragma unroll
    for (int i = 0; i < KP; ++i) cnt += (li[lane * KP + i] >= 0 && lv[lane * KP + i] >= thr) ? 1 : 0;
  }
  bool overflow = cnt == KP;                                 // a full list whose tail is still inside the band
  int pre = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += t;
  }
  const int total = __shfl_sync(0xffffffffu, pre, 31);
  pre -= cnt;
  overflow = overflow || total > TP_MAX_CAND;
  for (int i = 0; i < cnt; ++i)
    if (pre + i < TP_MAX_CAND) s_c[warp][pre + i] = li[lane * KP + i];
  __syncwarp();
  const int n_c = min(total, TP_MAX_CAND);
  if (__any_sync(0xffffffffu, overflow) && lane == 0) {
    const int at = atomicAdd(p.ovf_count, 1);
    if (at < p.ovf_cap) p.ovf_list[at] = (int)gw;
  }
  // exact values of the candidates; candidate i lives in lane i % 32, register i / 32
  float xq[2 * NH2];
  {
    const __half* qhi = bank + ((int64_t)job.q_slot * 2 * p.n_pix + q) * p.C;
    load_row_f32<NH2>(qhi, qhi + (int64_t)p.n_pix * p.C, lane, xq);
  }
  float ev0 = -INFINITY, ev1 = -INFINITY;
  int ei0 = -1, ei1 = -1;
  for (int i0 = 0; i0 < n_c; i0 += 2) {
    const int ia = s_c[warp][i0];
    const int ib = i0 + 1 < n_c ? s_c[warp][i0 + 1] : ia;
    const int pa = ia / p.n_pix, pb = ib / p.n_pix;
    const int sa = __ldg(p.mem_feat + job.mem_begin + pa) & ~FGVC_MEM_UNMASKED;
    const int sb = __ldg(p.mem_feat + job.mem_begin + pb) & ~FGVC_MEM_UNMASKED;
    const float va = exact_dot<NH2>(bank, xq, sa, ia - pa * p.n_pix, p.n_pix, p.C, lane);
    const float vb = exact_dot<NH2>(bank, xq, sb, ib - pb * p.n_pix, p.n_pix, p.C, lane);
    if ((i0 & 31) == lane) { if (i0 < 32) { ev0 = va; ei0 = ia; } else { ev1 = va; ei1 = ia; } }
    if (i0 + 1 < n_c && ((i0 + 1) & 31) == lane) { if (i0 + 1 < 32) { ev0 = vb; ei0 = ib; } else { ev1 = vb; ei1 = ib; } }
  }
  // exact top-k of the candidates: k rounds of warp arg-max.  Exact ties (the same frame twice in the memory
  // list) go to the smaller candidate index, so the result does not depend on the order of the lists.
  float out_v = -INFINITY;
  int out_i = -1;
  for (int r = 0; r < p.k_out; ++r) {
    const bool first = ev0 > ev1 || (ev0 == ev1 && (uint32_t)ei0 <= (uint32_t)ei1);   // id -1 sorts last
    const float x = first ? ev0 : ev1;
    const int xi = first ? ei0 : ei1;
    const bool have = xi >= 0;
    const float mx = warp_max(have ? x : -INFINITY);
    const bool tied = have && x == mx;
    if (!__any_sync(0xffffffffu, tied)) break;
    uint32_t best = tied ? (uint32_t)xi : 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == r) { out_v = mx; out_i = (int)best; }
    if (tied && (uint32_t)xi == best) { if (first) { ev0 = -INFINITY; ei0 = -1; } else { ev1 = -INFINITY; ei1 = -1; } }
  }
  const int64_t o0 = ((int64_t)jidx * p.groups * p.n_pix + q) * p.k_out;
  if (lane < p.k_out) { p.tv[o0 + lane] = out_v; p.ti[o0 + lane] = out_i; }
  for (int g = 1; g < p.groups; ++g) {
    const int64_t o = (((int64_t)jidx * p.groups + g) * p.n_pix + q) * p.k_out;
    if (lane < p.k_out) { p.tv[o + lane] = -INFINITY; p.ti[o + lane] = -1; }
  }
}

// ------------------------------------------------------------------ stage C: exact scan
struct ScanParams {
  int H, W, C, n_pix, groups, k_out, radius, mode, reach;
  const fgvc_job* jobs;
  const int32_t* mem_feat;
  float* tv;
  int32_t* ti;
  const int32_t* ovf_count;
  const int32_t* ovf_list;
  int ovf_cap;
};

constexpr int SC_WARPS = 8;

template <int K, int NH2>
__global__ void __launch_bounds__(32 * SC_WARPS)
exact_scan_kernel(const __half* __restrict__ bank, const ScanParams p) {
  __shared__ float s_v[SC_WARPS][K];
  __shared__ int s_i[SC_WARPS][K];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = min(*p.ovf_count, p.ovf_cap);
  for (int it = blockIdx.x; it < n; it += gridDim.x) {
    const int gw = p.ovf_list[it];
    const int jidx = gw / p.n_pix, q = gw - jidx * p.n_pix;
    const int qy = q / p.W, qx = q - qy * p.W;
    const fgvc_job job = p.jobs[jidx];
    float xq[2 * NH2];
    {
      const __half* qhi = bank + ((int64_t)job.q_slot * 2 * p.n_pix + q) * p.C;
      load_row_f32<NH2>(qhi, qhi + (int64_t)p.n_pix * p.C, lane, xq);
    }
    TopK<K> top;                                   // identical in every lane of the warp
    top.init();
    int rowctr = 0;
    for (int e = job.mem_begin; e < job.mem_end; ++e) {
      const int raw = __ldg(p.mem_feat + e);
      const bool masked = !(raw & FGVC_MEM_UNMASKED);
      const int slot = raw & ~FGVC_MEM_UNMASKED;
      const int pos_base = (e - job.mem_begin) * p.n_pix;
      const int y_lo = masked ? max(0, qy - p.reach) : 0, y_hi = masked ? min(p.H - 1, qy + p.reach) : p.H - 1;
      const int x_lo = masked ? max(0, qx - p.reach) : 0, x_hi = masked ? min(p.W - 1, qx + p.reach) : p.W - 1;
      for (int y = y_lo; y <= y_hi; ++y, ++rowctr) {
        if (rowctr % SC_WARPS != warp) continue;
        for (int x = x_lo; x <= x_hi; x += 4) {             // four keys in flight per warp (latency-bound otherwise)
          float part[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            ok[u] = x + u <= x_hi && (!masked || in_mask(y - qy, x + u - qx, p.radius, p.mode));
            part[u] = ok[u] ? partial_dot<NH2>(bank, xq, slot, y * p.W + x + u, p.n_pix, p.C, lane) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float v = warp_sum(part[u]);
            if (ok[u] && v > top.thr()) top.push(v, pos_base + y * p.W + x + u);
          }
        }
      }
    }
    __syncthreads();                               // previous iteration's merge is done with s_v / s_i
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < K; ++i) { s_v[warp][i] = top.v[i]; s_i[warp][i] = top.id[i]; }
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < SC_WARPS; ++w)
        for (int i = 0; i < K; ++i) {
          const float v = s_v[w][i];
          if (!(v > top.thr())) break;
          top.push(v, s_i[w][i]);
        }
      const int64_t o0 = ((int64_t)jidx * p.groups * p.n_pix + q) * p.k_out;
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < K; ++i)
          if (i < p.k_out) { p.tv[o0 + i] = top.v[i]; p.ti[o0 + i] = top.id[i]; }
      }
    }
  }
}

// ------------------------------------------------------------------------------ host
// 5-D map over feat16[slot][part][H][W][C]; box = (64 channels, 16, bh, ONE part, 1), 128B swizzle
static int make_map16p(CUtensorMap* map, const void* bank, int n_slots, int H, int W, int C, int bh) {
  EncodeTiledFn enc = get_tensormap_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return FGVC_ERR_CUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 2, (cuuint64_t)n_slots};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)2 * H * W * C * 2};
  cuuint32_t box[5] = {64, 16, (cuuint32_t)bh, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(bank), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f16 prefilter) failed with %d (H=%d W=%d C=%d bh=%d)", (int)r, H, W, C, bh);
    return FGVC_ERR_CUDA;
  }
  return FGVC_OK;
}

static int prefilter_kp(int K) { return K <= 2 ? 4 : (K <= 10 ? 10 : 16); }

bool tc16p_supported(int H, int W, int C, int K, int groups) {
  return C % 64 == 0 && C >= 64 && C <= 256 && K >= 1 && K <= 16 && H >= 1 && W >= 1 && groups >= 1 &&
         groups * TP_EPI_WG <= 32;
}

// workspace: candidate values + indices, the stage-C queue (counter + one entry per query)
int64_t tc16p_workspace_bytes(int n_jobs, int groups, int n_pix, int K) {
  const int64_t cand = (int64_t)n_jobs * groups * n_pix * TP_EPI_WG * prefilter_kp(K);
  return cand * 8 + 256 + (int64_t)n_jobs * n_pix * 4;
}

// one TS-form MMA of N columns per 16 channels: a box costs ~N/2 per K step plus a fixed hand-shake
static int box_costp(int rows, int bh) { return cdiv(rows, bh) * (16 * bh + 24); }
static int pick_bhp(int rows) {
  int best = TP_MAX_BH;
  for (int bh = TP_MAX_BH - 1; bh >= 1; --bh)
    if (box_costp(rows, bh) < box_costp(rows, best)) best = bh;
  return best;
}

template <int KP>
static int launch_prefilter(const CUtensorMap& mk, const void* bank, const TcpParams& p, dim3 grid, cudaStream_t st) {
  FGVC_CUDA(cudaFuncSetAttribute(affinity_prefilter_tc16_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 TP_SMEM_BYTES));
  affinity_prefilter_tc16_kernel<KP><<<grid, TP_THREADS, TP_SMEM_BYTES, st>>>(mk, reinterpret_cast<const __half*>(bank), p);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

template <int KP, int NH2>
static int launch_rescore(const void* bank, const RescoreParams& rp, cudaStream_t st) {
  const int64_t warps = (int64_t)rp.n_jobs * rp.n_pix;
  rescore_kernel<KP, NH2><<<(unsigned)((warps + RS_WARPS - 1) / RS_WARPS), 32 * RS_WARPS, 0, st>>>(
      reinterpret_cast<const __half*>(bank), rp);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}
template <int KP>
static int launch_rescore_c(const void* bank, const RescoreParams& rp, cudaStream_t st) {
  switch (rp.C / 64) {
    case 1: return launch_rescore<KP, 1>(bank, rp, st);
    case 2: return launch_rescore<KP, 2>(bank, rp, st);
    case 3: return launch_rescore<KP, 3>(bank, rp, st);
    default: return launch_rescore<KP, 4>(bank, rp, st);
  }
}
template <int K, int NH2>
static int launch_scan(const void* bank, const ScanParams& sp, cudaStream_t st) {
  exact_scan_kernel<K, NH2><<<296, 32 * SC_WARPS, 0, st>>>(reinterpret_cast<const __half*>(bank), sp);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}
template <int K>
static int launch_scan_c(const void* bank, const ScanParams& sp, cudaStream_t st) {
  switch (sp.C / 64) {
    case 1: return launch_scan<K, 1>(bank, sp, st);
    case 2: return launch_scan<K, 2>(bank, sp, st);
    case 3: return launch_scan<K, 3>(bank, sp, st);
    default: return launch_scan<K, 4>(bank, sp, st);
  }
}

int launch_affinity_topk_tc16p(const void* bank, int n_slots, int H, int W, int C, const fgvc_job* jobs, int n_jobs,
                               const int32_t* mem_feat, int radius, int mode, int K, int groups, float* tv,
                               int32_t* ti, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int n_pix = H * W;
  FGVC_CHECK_ARG(workspace != nullptr && workspace_bytes >= tc16p_workspace_bytes(n_jobs, groups, n_pix, K),
                 "prefilter engine: workspace of %lld bytes needed (fgvc_affinity_topk_workspace_bytes)",
                 (long long)tc16p_workspace_bytes(n_jobs, groups, n_pix, K));
  FGVC_CHECK_ARG((int64_t)n_jobs * n_pix < (1ll << 31), "prefilter engine: too many queries in one launch");
  TcpParams p;
  p.H = H; p.W = W; p.C = C; p.n_pix = n_pix;
  p.radius = radius; p.mode = mode; p.reach = mask_reach(radius, mode);
  const int reach = p.reach;
  auto halo_cost = [&](int qh, int qw) {
    int rows = min(H, qh + 2 * reach), cols = min(W, qw + 2 * reach);
    double tiles = (double)cdiv(H, qh) * cdiv(W, qw);
    return tiles * box_costp(rows, pick_bhp(rows)) * cdiv(cols, 16);
  };
  if (halo_cost(16, 8) < halo_cost(8, 16)) { p.QH = 16; p.QW = 8; p.qw_shift = 3; }
  else { p.QH = 8; p.QW = 16; p.qw_shift = 4; }
  p.BH = pick_bhp(min(H, p.QH + 2 * reach));
  static const int force_bh = getenv("FGVC_TCP_BH") ? atoi(getenv("FGVC_TCP_BH")) : 0;          // perf experiments only
  if (force_bh >= 1 && force_bh <= TP_MAX_BH) p.BH = force_bh;
  p.groups = groups;
  p.tiles_x = cdiv(W, p.QW);
  p.jobs = jobs; p.mem_feat = mem_feat;
  static const int exp_flags = getenv("FGVC_TCP_EXP") ? atoi(getenv("FGVC_TCP_EXP")) : 0;     // perf experiments only
  p.exp_flags = exp_flags;
  FGVC_CHECK_ARG(p.reach + 1 <= 128, "prefilter engine: radius %d too large", radius);
  if (cdiv(H, p.BH) * cdiv(W, 16) > TP_MAX_BOXES || H >= 65536 || W >= 65536) {
    set_error("prefilter engine: a %dx%d map has more than %d key boxes", H, W, TP_MAX_BOXES);
    return FGVC_ERR_UNSUPPORTED;
  }
  const int KP = prefilter_kp(K);
  const int64_t cand = (int64_t)n_jobs * groups * n_pix * TP_EPI_WG * KP;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  p.cv = reinterpret_cast<float*>(ws);
  p.ci = reinterpret_cast<int32_t*>(ws + cand * 4);
  int32_t* ovf_count = reinterpret_cast<int32_t*>(ws + cand * 8);
  int32_t* ovf_list = reinterpret_cast<int32_t*>(ws + cand * 8 + 256);
  FGVC_CUDA(cudaMemsetAsync(ovf_count, 0, 256, st));
  p.stats = reinterpret_cast<unsigned long long*>(ws + cand * 8 + 64);
  CUtensorMap mk;
  int rc = make_map16p(&mk, bank, n_slots, H, W, C, p.BH);
  if (rc) return rc;
  dim3 grid(cdiv(H, p.QH) * p.tiles_x, groups, n_jobs);
  if (KP == 4) rc = launch_prefilter<4>(mk, bank, p, grid, st);
  else if (KP == 10) rc = launch_prefilter<10>(mk, bank, p, grid, st);
  else rc = launch_prefilter<16>(mk, bank, p, grid, st);
  if (rc) return rc;

  RescoreParams rp;
  rp.C = C; rp.n_pix = n_pix; rp.groups = groups; rp.k_out = K; rp.n_jobs = n_jobs;
  rp.jobs = jobs; rp.mem_feat = mem_feat; rp.cv = p.cv; rp.ci = p.ci; rp.tv = tv; rp.ti = ti;
  rp.band = 2.f * FGVC_PREFILTER_EPS;
  rp.ovf_count = ovf_count; rp.ovf_list = ovf_list; rp.ovf_cap = n_jobs * n_pix;
  if (KP == 4) rc = launch_rescore_c<4>(bank, rp, st);
  else if (KP == 10) rc = launch_rescore_c<10>(bank, rp, st);
  else rc = launch_rescore_c<16>(bank, rp, st);
  if (rc) return rc;

  ScanParams sp;
  sp.H = H; sp.W = W; sp.C = C; sp.n_pix = n_pix; sp.groups = groups; sp.k_out = K;
  sp.radius = radius; sp.mode = mode; sp.reach = reach;
  sp.jobs = jobs; sp.mem_feat = mem_feat; sp.tv = tv; sp.ti = ti;
  sp.ovf_count = ovf_count; sp.ovf_list = ovf_list; sp.ovf_cap = n_jobs * n_pix;
  if (K <= 4) return launch_scan_c<4>(bank, sp, st);
  if (K <= 10) return launch_scan_c<10>(bank, sp, st);
  return launch_scan_c<16>(bank, sp, st);
}

}  // namespace fgvc
