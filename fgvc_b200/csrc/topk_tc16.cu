// K1 (tcgen05 engine, fp16 three-term split): fused affinity + radius mask + running top-K.
//
// Same contract and the same CTA anatomy as topk_tc.cu (TMA producer warp, MMA issuer warp,
// 4 epilogue warpgroups, thread = query), but the operands are the F16 bank:
//     x = hi + 2^-11 * lo,   hi = fp16(x),  lo = fp16((x - hi) * 2^11)
// -- 11 + 11 significant bits per operand, exactly what the 3xTF32 split carries (rows are
// unit vectors, so fp16's exponent range is enough).  Three products per K step as before:
//     D1 = hi_q * hi_k,     D2 = hi_q * lo_k + lo_q * hi_k,     affinity = D1 + 2^-11 * D2
// Why it is faster (profiles/r1_mma_issue_rates.md): an M=128 tcgen05.mma costs N/2 cycles
// when A comes from tensor memory and N/2 + 43 when A comes from shared memory.  With fp16
// BOTH query parts fit in TMEM (2 x C/2 columns, packed two channels per 32-bit cell) next
// to two accumulators, so every MMA is TS-form and runs at the nominal rate; kind::f16 also
// contracts 16 channels per instruction instead of 8, and a key costs 4 B per channel
// instead of 8.  Per K step (16 channels) and key box of N keys:
//     TS(A = hi_q, B = [hi_k ; lo_k] stacked along N, 2N columns) -> D[0,2N)   = [D1 | hi_q*lo_k]
//     TS(A = lo_q, B = hi_k,                               N columns) -> D[N,2N) += lo_q*hi_k
// The hi/lo key rows of a box are one TMA box (the "part" dimension of the 5-D map has
// extent 2), landing stacked in one 128B-swizzled stage.  No query operand lives in shared
// memory, so the whole 192 KB ring streams key boxes (3 boxes of 64 KB in flight at C = 256).
// TMEM: [0,128) accumulator 0, [128,256) accumulator 1, [256,384) hi_q, [384,512) lo_q.
#include <stdlib.h>

#include "tc_common.cuh"

namespace fgvc {

constexpr int T16_STAGE_BYTES = 16 * 1024;     // 2 parts x 64 keys x 128 B (64 channels of fp16)
constexpr int T16_STAGES = 12;
constexpr int T16_MAX_BH = 4;                  // N <= 64, 2N <= 128 accumulator columns
constexpr int T16_AHI_COL = 256, T16_ALO_COL = 384;
constexpr int T16_EPI_WG = 4;
constexpr int T16_THREADS = 64 + 128 * T16_EPI_WG;
constexpr int T16_MAX_BOXES = 4096;             // per box list (masked halo / whole frame)
constexpr int T16_AUX_BYTES = 1024 + 2 * T16_MAX_BOXES * 4;
constexpr int T16_SMEM_BYTES = T16_STAGES * T16_STAGE_BYTES + T16_AUX_BYTES;

struct Tc16Params {
  int H, W, C, n_pix;
  int radius, mode, reach;
  int QH, QW, qw_shift;
  int BH;
  int groups, k_out;
  int tiles_x;
  const fgvc_job* jobs;
  const int32_t* mem_feat;
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
affinity_topk_tc16_kernel(const __grid_constant__ CUtensorMap tmap_k, const __half* __restrict__ bank,
                          const Tc16Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + T16_STAGES * T16_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + T16_STAGES;
  uint64_t* tfull_bar = empty_bar + T16_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint64_t* a_bar = tempty_bar + 2;               // query operand written to TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
  int* nbox = reinterpret_cast<int*>(tmem_slot + 2);      // [2] number of boxes in each list
  int* halfw = reinterpret_cast<int*>(tmem_slot + 4);     // [reach+1] <= 128 entries
  // box lists (by | bx << 16): [0] = radius halo of this query tile minus boxes no query can see,
  // [1] = every box of the frame (unmasked memory entries).  Identical for all memory entries, so
  // the three warp roles just walk a list instead of re-deriving the geometry per box.
  uint32_t* boxes = reinterpret_cast<uint32_t*>(ring + T16_STAGES * T16_STAGE_BYTES + 1024);
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qy0 = (blockIdx.x / p.tiles_x) * p.QH, qx0 = (blockIdx.x % p.tiles_x) * p.QW;
  const int g = blockIdx.y;
  const fgvc_job job = p.jobs[blockIdx.z];
  const int n_mem = job.mem_end - job.mem_begin;
  const int per = (n_mem + p.groups - 1) / p.groups;
  const int e_lo = job.mem_begin + g * per;
  const int e_hi = min(job.mem_end, e_lo + per);
  if (e_lo >= e_hi) {
    // This (job, group) has no memory entry (more groups than entries: per-entry lists of a short memory).  Block
    // uniform and before any barrier / tensor-memory allocation: write the empty lists and leave.
    if (threadIdx.x < 128) {
      const int m = threadIdx.x;
      const int qy = qy0 + (m >> p.qw_shift), qx = qx0 + (m & (p.QW - 1));
      if (qy < p.H && qx < p.W) {
        const int64_t o = (((int64_t)blockIdx.z * p.groups + g) * p.n_pix + qy * p.W + qx) * p.k_out;
        for (int i = 0; i < p.k_out; ++i) { p.tv[o + i] = -INFINITY; p.ti[o + i] = -1; }
      }
    }
    return;
  }
  const int N = 16 * p.BH;
  const int n_kc = p.C / 64;
  // one stage = one whole key box (all C channels: n_kc chunks of 16 KB), so the single issuing
  // threads pay one barrier round trip per box instead of one per 64 channels
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
  for (int d = threadIdx.x; d <= p.reach; d += T16_THREADS) {
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
    uint32_t* list = boxes + li * T16_MAX_BOXES;
    int cnt = 0;
    for (int base = 0; base < nrows * ncols; base += 32) {
      const int i = base + lane;
      const int by = w.y_lo + (i / ncols) * p.BH, bx = w.x_lo + (i % ncols) * 16;
      const bool keep = i < nrows * ncols && !box_skipped(p, w, by, bx, qy0, qx0);
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
      if (keep && pos < T16_MAX_BOXES) list[pos] = (uint32_t)by | ((uint32_t)bx << 16);
      cnt += __popc(bal);
    }
    cnt = min(cnt, T16_MAX_BOXES);
    __syncwarp();
    // Centre-out order for the halo list: the best matches of a query sit near its own position, so
    // walking the boxes nearest to the tile first raises the running K-th values early and the
    // (divergent, ~80-instruction) list insertions become rare.  Rank sort, n is a few dozen.
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

  if (warp == 0) {
    // ================================ TMA producer ====================================
    // one elected lane runs the whole loop (the compiler then keeps everything in uniform registers)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int e = e_hi - 1; e >= e_lo; --e) {       // newest memory frame first: thresholds rise early
        const int raw = p.mem_feat[e];
        const int slot = raw & ~FGVC_MEM_UNMASKED;
        const int li = (raw & FGVC_MEM_UNMASKED) ? 1 : 0;
        const int nb = nbox[li];
        for (int b = 0; b < nb; ++b) {
          const uint32_t bb = boxes[li * T16_MAX_BOXES + b];
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
    int n_total = 0;                                 // boxes this CTA processes
    for (int e = e_lo; e < e_hi; ++e) n_total += nbox[(p.mem_feat[e] & FGVC_MEM_UNMASKED) ? 1 : 0];
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
    const int qy = qy0 + (m >> p.qw_shift), qx = qx0 + (m & (p.QW - 1));
    const bool qvalid = qy < p.H && qx < p.W;
    if (wg == 0 && e_lo < e_hi) {
      // both query parts -> tensor memory, two fp16 channels per 32-bit cell (lower channel in the low half)
      const int64_t part = (int64_t)p.n_pix * p.C;
      const __half* row = bank + (int64_t)job.q_slot * 2 * part + (int64_t)(qvalid ? qy * p.W + qx : 0) * p.C;
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
      const int raw = p.mem_feat[e];
      const bool masked = !(raw & FGVC_MEM_UNMASKED);
      const int li = masked ? 0 : 1;
      const int nb = nbox[li];
      const int pos_base = (e - job.mem_begin) * p.n_pix;
      for (int b = 0; b < nb; ++b) {
        const uint32_t bb = boxes[li * T16_MAX_BOXES + b];
        const int by = (int)(bb & 0xffffu), bx = (int)(bb >> 16);
        const int ky = by + row;
        // 16-bit interval mask of the in-mask, in-image keys of this key row
        uint32_t bits = 0;
        if (row_ok && ky < p.H) {
          int lo = 0, hi = p.W - 1;
          if (masked) {
            const int ady = abs(ky - qy);
            const int hw = ady <= p.reach ? halfw[ady] : -1;
            lo = hw < 0 ? 1 : max(qx - hw, 0);
            hi = hw < 0 ? 0 : min(qx + hw, p.W - 1);
          }
          lo = max(lo - bx, 0);
          hi = min(hi - bx, 15);
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

bool tc16_supported(int H, int W, int C, int K) {
  return C % 64 == 0 && C >= 64 && C <= 256 && K >= 1 && K <= 16 && H >= 1 && W >= 1;
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
  FGVC_CUDA(cudaFuncSetAttribute(affinity_topk_tc16_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 T16_SMEM_BYTES));
  affinity_topk_tc16_kernel<K><<<grid, T16_THREADS, T16_SMEM_BYTES, st>>>(mk, reinterpret_cast<const __half*>(bank), p);
  FGVC_LAUNCH_CHECK();
  return FGVC_OK;
}

int launch_affinity_topk_tc16(const void* bank, int n_slots, int H, int W, int C, const fgvc_job* jobs, int n_jobs,
                              const int32_t* mem_feat, int radius, int mode, int K, int groups, float* tv, int32_t* ti,
                              float* dbg, int32_t* dbg_meta, int dbg_max_boxes, cudaStream_t st) {
  Tc16Params p;
  p.H = H; p.W = W; p.C = C; p.n_pix = H * W;
  p.radius = radius; p.mode = mode; p.reach = mask_reach(radius, mode);
  const int reach = p.reach;
  auto halo_cost = [&](int qh, int qw) {
    int rows = min(H, qh + 2 * reach), cols = min(W, qw + 2 * reach);
    double tiles = (double)cdiv(H, qh) * cdiv(W, qw);
    return tiles * box_cost16(rows, pick_bh16(rows)) * cdiv(cols, 16);
  };
  if (halo_cost(16, 8) < halo_cost(8, 16)) { p.QH = 16; p.QW = 8; p.qw_shift = 3; }
  else { p.QH = 8; p.QW = 16; p.qw_shift = 4; }
  p.BH = pick_bh16(min(H, p.QH + 2 * reach));
  p.groups = groups; p.k_out = K;
  p.tiles_x = cdiv(W, p.QW);
  p.jobs = jobs; p.mem_feat = mem_feat; p.tv = tv; p.ti = ti;
  p.dbg = dbg; p.dbg_meta = dbg_meta; p.dbg_max_boxes = dbg_max_boxes;
  static const int exp_flags = getenv("FGVC_TC16_EXP") ? atoi(getenv("FGVC_TC16_EXP")) : 0;   // perf experiments only
  p.exp_flags = exp_flags;
  FGVC_CHECK_ARG(p.reach + 1 <= 128, "tcgen05 f16 engine: radius %d too large", radius);
  if (cdiv(H, p.BH) * cdiv(W, 16) > T16_MAX_BOXES || H >= 65536 || W >= 65536) {
    set_error("tcgen05 f16 engine: a %dx%d map has more than %d key boxes", H, W, T16_MAX_BOXES);
    return FGVC_ERR_UNSUPPORTED;      // AUTO falls back to the CUDA-core engine
  }
  CUtensorMap mk;
  int rc = make_map16(&mk, bank, n_slots, H, W, C, p.BH);
  if (rc) return rc;
  dim3 grid(cdiv(H, p.QH) * p.tiles_x, groups, n_jobs);
  if (K <= 4) return launch_tc16<4>(mk, bank, p, grid, st);
  if (K <= 10) return launch_tc16<10>(mk, bank, p, grid, st);
  return launch_tc16<16>(mk, bank, p, grid, st);
}

}  // namespace fgvc
