// extern "C" entry points that are pure plumbing: error state, engine dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fgvc {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace fgvc

using namespace fgvc;

extern "C" const char* fgvc_last_error(void) { return g_err; }
extern "C" int fgvc_version(void) { return FGVC_VERSION; }
extern "C" int64_t fgvc_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int fgvc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static bool tc_ok(int fmt, int H, int W, int C, int K) {
  return fmt == FGVC_BANK_F16 ? tc16_supported(H, W, C, K) : tc_supported(H, W, C, K);
}
extern "C" int fgvc_tc_supported(int32_t bank_format, int32_t H, int32_t W, int32_t C, int32_t K) {
  return tc_ok(bank_format, H, W, C, K) ? 1 : 0;
}

extern "C" int64_t fgvc_topk_bytes(int32_t n_jobs, int32_t groups, int32_t n_query, int32_t K) {
  return (int64_t)n_jobs * groups * n_query * K * 4;
}

static int pick_engine(int engine, int fmt, int H, int W, int C, int K, bool* use_tc) {
  bool ok = tc_ok(fmt, H, W, C, K);
  if (engine == FGVC_ENGINE_TCGEN05) {
    FGVC_CHECK_ARG(ok, "tcgen05 engine needs C %% 32 == 0 (TF32 bank) / C %% 64 == 0 (F16 bank) and K <= 16 (C=%d K=%d)", C, K);
    *use_tc = true;
  } else if (engine == FGVC_ENGINE_SIMT) {
    *use_tc = false;
  } else {
    FGVC_CHECK_ARG(engine == FGVC_ENGINE_AUTO, "unknown engine %d", engine);
    *use_tc = ok;
  }
  return FGVC_OK;
}

static int affinity_topk_impl(const void* feat_bank, int32_t fmt, int32_t n_slots, int32_t H, int32_t W, int32_t C,
                              const fgvc_job* jobs, int32_t n_jobs, const int32_t* mem_feat_slot, int32_t radius,
                              int32_t mask_mode, int32_t K, int32_t groups, float* topk_val, int32_t* topk_idx,
                              int32_t engine, float* dbg, int32_t* dbg_meta, int32_t dbg_max_boxes, void* stream,
                              const float* floor = nullptr) {
  FGVC_CHECK_ARG(feat_bank && jobs && mem_feat_slot && topk_val && topk_idx, "fgvc_affinity_topk: null pointer");
  FGVC_CHECK_ARG(H > 0 && W > 0 && C > 0 && n_jobs > 0 && n_slots > 0, "fgvc_affinity_topk: bad shape");
  FGVC_CHECK_ARG(K >= 1 && K <= 16, "fgvc_affinity_topk: topk=%d not in [1,16]", K);
  FGVC_CHECK_ARG(groups >= 1 && groups <= 64, "fgvc_affinity_topk: groups=%d not in [1,64]", groups);
  FGVC_CHECK_ARG(radius >= 1, "fgvc_affinity_topk: radius=%d must be >= 1", radius);
  FGVC_CHECK_ARG(mask_mode == FGVC_MASK_CIRCLE || mask_mode == FGVC_MASK_SQUARE, "fgvc_affinity_topk: bad mask mode");
  FGVC_CHECK_ARG(fmt == FGVC_BANK_TF32 || fmt == FGVC_BANK_F16, "fgvc_affinity_topk: bad bank format %d", fmt);
  bool use_tc = false;
  int rc = pick_engine(engine, fmt, H, W, C, K, &use_tc);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc && fmt == FGVC_BANK_F16) {
    // (the floor is an optional accelerator of the fp16 tensor engine; the other engines ignore it)
    rc = launch_affinity_topk_tc16(feat_bank, n_slots, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode, K,
                                   groups, topk_val, topk_idx, dbg, dbg_meta, dbg_max_boxes, st, floor);
    if (rc != FGVC_ERR_UNSUPPORTED || engine != FGVC_ENGINE_AUTO) return rc;
    // a shape the tensor kernel does not take (e.g. a huge map with unmasked frames): CUDA-core engine
    return launch_affinity_topk_simt(feat_bank, fmt, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode, K,
                                     groups, topk_val, topk_idx, st);
  }
  if (use_tc)
    return launch_affinity_topk_tc(reinterpret_cast<const float*>(feat_bank), n_slots, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode, K,
                                   groups, topk_val, topk_idx, dbg, dbg_meta, dbg_max_boxes, st);
  return launch_affinity_topk_simt(feat_bank, fmt, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode, K, groups,
                                   topk_val, topk_idx, st);
}

extern "C" int fgvc_affinity_topk(const void* feat_bank, int32_t bank_format, int32_t n_slots, int32_t H, int32_t W, int32_t C,
                                  const fgvc_job* jobs, int32_t n_jobs, const int32_t* mem_feat_slot, int32_t radius,
                                  int32_t mask_mode, int32_t K, int32_t groups, float* topk_val, int32_t* topk_idx,
                                  int32_t engine, void* stream) {
  return affinity_topk_impl(feat_bank, bank_format, n_slots, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode,
                            K, groups, topk_val, topk_idx, engine, nullptr, nullptr, 0, stream);
}

extern "C" int fgvc_affinity_topk_seeded(const void* feat_bank, int32_t bank_format, int32_t n_slots, int32_t H, int32_t W,
                                         int32_t C, const fgvc_job* jobs, int32_t n_jobs, const int32_t* mem_feat_slot,
                                         int32_t radius, int32_t mask_mode, int32_t K, int32_t groups,
                                         const float* floor, float* topk_val, int32_t* topk_idx, int32_t engine,
                                         void* stream) {
  return affinity_topk_impl(feat_bank, bank_format, n_slots, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode,
                            K, groups, topk_val, topk_idx, engine, nullptr, nullptr, 0, stream, floor);
}

extern "C" int fgvc_topk_floor(const void* feat_bank, int32_t bank_format, int32_t H, int32_t W, int32_t C,
                               const fgvc_job* jobs, int32_t n_jobs, const int32_t* seed_feat_slot, int32_t radius,
                               int32_t mask_mode, int32_t K, float* floor_out, void* stream) {
  FGVC_CHECK_ARG(feat_bank && jobs && seed_feat_slot && floor_out, "fgvc_topk_floor: null pointer");
  FGVC_CHECK_ARG(H > 0 && W > 0 && C > 0 && n_jobs > 0 && K >= 1 && K <= 16 && radius >= 1, "fgvc_topk_floor: bad arguments");
  FGVC_CHECK_ARG(mask_mode == FGVC_MASK_CIRCLE || mask_mode == FGVC_MASK_SQUARE, "fgvc_topk_floor: bad mask mode");
  if (bank_format != FGVC_BANK_F16) {
    set_error("fgvc_topk_floor: F16 bank only");
    return FGVC_ERR_UNSUPPORTED;
  }
  return launch_topk_floor16(feat_bank, H, W, C, jobs, n_jobs, seed_feat_slot, radius, mask_mode, K, floor_out,
                             (cudaStream_t)stream);
}

extern "C" int fgvc_packed_tile_shape(int32_t H, int32_t W, int32_t radius, int32_t mask_mode, int32_t jobs_per_tile,
                                      int32_t* tile_h, int32_t* tile_w, int32_t* box_h, int32_t* ctas_per_tile) {
  FGVC_CHECK_ARG(tile_h && tile_w && box_h && ctas_per_tile && H > 0 && W > 0 && radius >= 1,
                 "fgvc_packed_tile_shape: bad arguments");
  FGVC_CHECK_ARG(jobs_per_tile == 1 || jobs_per_tile == 2 || jobs_per_tile == 4, "fgvc_packed_tile_shape: jobs_per_tile");
  int a, b, c, d;
  packed_tile_shape(H, W, mask_reach(radius, mask_mode), jobs_per_tile, &a, &b, &c, &d);
  *tile_h = a; *tile_w = b; *box_h = c; *ctas_per_tile = d;
  return FGVC_OK;
}

extern "C" int fgvc_affinity_topk_packed(const void* feat_bank, int32_t n_slots, int32_t H, int32_t W, int32_t C,
                                         const fgvc_job* jobs, const fgvc_tile_group* tile_groups,
                                         int32_t n_tile_groups, const int32_t* union_feat_slot, const int32_t* union_pos,
                                         int32_t jobs_per_tile, int32_t radius, int32_t mask_mode, int32_t K,
                                         int32_t groups, int32_t split, float* topk_val, int32_t* topk_idx,
                                         void* stream) {
  FGVC_CHECK_ARG(feat_bank && jobs && tile_groups && union_feat_slot && union_pos && topk_val && topk_idx,
                 "fgvc_affinity_topk_packed: null pointer");
  FGVC_CHECK_ARG(H > 0 && W > 0 && n_tile_groups > 0 && n_slots > 0, "fgvc_affinity_topk_packed: bad shape");
  FGVC_CHECK_ARG(tc16_supported(H, W, C, K), "fgvc_affinity_topk_packed: needs C %% 64 == 0, C <= 256, K <= 16 (C=%d K=%d)", C, K);
  FGVC_CHECK_ARG(groups >= 1 && groups <= 64, "fgvc_affinity_topk_packed: groups=%d not in [1,64]", groups);
  FGVC_CHECK_ARG(radius >= 1, "fgvc_affinity_topk_packed: radius=%d must be >= 1", radius);
  FGVC_CHECK_ARG(mask_mode == FGVC_MASK_CIRCLE || mask_mode == FGVC_MASK_SQUARE, "fgvc_affinity_topk_packed: bad mask mode");
  return launch_affinity_topk_tc16_packed(feat_bank, n_slots, H, W, C, jobs, tile_groups, n_tile_groups, union_feat_slot,
                                          union_pos, jobs_per_tile, radius, mask_mode, K, groups, split, topk_val, topk_idx,
                                          (cudaStream_t)stream);
}

extern "C" int fgvc_debug_affinity_boxes(const void* feat_bank, int32_t bank_format, int32_t n_slots, int32_t H, int32_t W, int32_t C,
                                         const fgvc_job* jobs, int32_t n_jobs, const int32_t* mem_feat_slot,
                                         int32_t radius, int32_t mask_mode, int32_t K, float* topk_val,
                                         int32_t* topk_idx, float* dbg, int32_t* dbg_meta, int32_t dbg_max_boxes,
                                         void* stream) {
  FGVC_CHECK_ARG(dbg && dbg_meta && dbg_max_boxes > 0, "fgvc_debug_affinity_boxes: null debug buffers");
  return affinity_topk_impl(feat_bank, bank_format, n_slots, H, W, C, jobs, n_jobs, mem_feat_slot, radius, mask_mode, K,
                            1, topk_val, topk_idx, FGVC_ENGINE_TCGEN05, dbg, dbg_meta, dbg_max_boxes, stream);
}
