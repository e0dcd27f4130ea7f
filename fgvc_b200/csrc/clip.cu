// Clip-level drivers: the sequential part of the reference loop
// (vanilla_tracker.py:345-412) enqueued by ONE C call, so that the host costs a few
// microseconds per frame instead of a Python round trip per kernel.
#include "common.cuh"

using namespace fgvc;

// After K0 + K1 ran for the whole clip: for every job j (in order) gather the labels of its
// query frame (K1b) and decode them to a uint8 mask of the image size.
extern "C" int fgvc_mask_clip_tail(const float* topk_val, const int32_t* topk_idx, int32_t K, int32_t groups,
                                   const fgvc_job* jobs_dev, const fgvc_job* jobs_host, int32_t job_begin,
                                   int32_t job_end, const int32_t* mem_label_slot, int32_t H, int32_t W,
                                   float temperature, int32_t flags,
                                   float* lab_bank, int32_t Lp, int32_t L, int32_t out_h, int32_t out_w,
                                   float* scratch_minmax, uint8_t* masks, float* maps_nchw, void* chain_ws,
                                   int64_t chain_ws_bytes, void* stream) {
  FGVC_CHECK_ARG(jobs_dev && jobs_host && masks && scratch_minmax && lab_bank, "fgvc_mask_clip_tail: null pointer");
  FGVC_CHECK_ARG(L > 0 && L <= 255 && Lp >= L && Lp % 4 == 0, "fgvc_mask_clip_tail: bad label sizes");
  const int n_pix = H * W;
  const int64_t mask_elems = (int64_t)out_h * out_w;
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & FGVC_HARD_PROP) {
    // hard propagation: the memory keeps one_hot(argmax) of every propagated frame, the prediction is
    // decoded from the soft labels first (vanilla_tracker.py:762-798) -- so decode per frame
    for (int j = job_begin; j < job_end; ++j) {
      int rc = fgvc_gather_labels(topk_val, topk_idx, K, groups, jobs_dev, j, j + 1, mem_label_slot, n_pix,
                                  temperature, flags, lab_bank, Lp, stream);
      if (rc) return rc;
      const int slot = jobs_host[j].out_slot;
      float* lab_slot = lab_bank + (int64_t)slot * n_pix * Lp;
      if (maps_nchw) {
        rc = fgvc_labels_to_nchw(lab_bank, slot, Lp, L, n_pix, maps_nchw + (int64_t)slot * L * n_pix, stream);
        if (rc) return rc;
      }
      rc = launch_decode(lab_slot, true, L, Lp, H, W, out_h, out_w, reinterpret_cast<uint32_t*>(scratch_minmax),
                         masks + slot * mask_elems, st);
      if (rc) return rc;
      rc = launch_labels_harden(lab_slot, n_pix, L, Lp, st);
      if (rc) return rc;
    }
    return FGVC_OK;
  }
  // the recurrence lives only in the gather: run the chain first ...
  if (chain_ws != nullptr && job_end - job_begin > 1) {
    // ... as ONE persistent kernel with a grid barrier per frame (gather.cu)
    int rc = launch_gather_chain(topk_val, topk_idx, K, groups, jobs_dev, job_begin, job_end, mem_label_slot, nullptr,
                                 n_pix, temperature, flags, lab_bank, Lp, chain_ws, chain_ws_bytes, st);
    if (rc) return rc;
    if (maps_nchw) {
      rc = launch_labels_to_nchw_jobs(lab_bank, jobs_dev, job_begin, job_end, Lp, L, n_pix, maps_nchw, st);
      if (rc) return rc;
    }
  } else {
    for (int j = job_begin; j < job_end; ++j) {
      int rc = fgvc_gather_labels(topk_val, topk_idx, K, groups, jobs_dev, j, j + 1, mem_label_slot, n_pix,
                                  temperature, flags, lab_bank, Lp, stream);
      if (rc) return rc;
      if (maps_nchw) {
        const int slot = jobs_host[j].out_slot;
        rc = fgvc_labels_to_nchw(lab_bank, slot, Lp, L, n_pix, maps_nchw + (int64_t)slot * L * n_pix, stream);
        if (rc) return rc;
      }
    }
  }
  // ... then decode every frame of the range in two batched launches (scratch: [n_jobs][2L] words)
  return launch_decode_jobs(lab_bank, jobs_dev, job_begin, job_end, L, Lp, H, W, out_h, out_w,
                            reinterpret_cast<uint32_t*>(scratch_minmax), masks, st);
}

// Point tracking tail: the gather chain over jobs [job_begin, job_end) with an NCHW copy of every
// propagated frame into maps_nchw[slot][L][H*W], then ONE K3 launch (fused up-sample / soft-argmax)
// over all (frame, point) maps of the range: coords[slot][L][2].  The out_slots of the range must
// be consecutive (they are frame indices).
static int point_clip_tail_impl(const float* topk_val, const int32_t* topk_idx, int32_t K, int32_t groups,
                                    const fgvc_job* jobs_dev, const fgvc_job* jobs_host, int32_t job_begin,
                                    int32_t job_end, const int32_t* mem_label_slot, int32_t H, int32_t W,
                                    float temperature, int32_t flags, float* lab_bank, int32_t Lp, int32_t L, int32_t out_h,
                                    int32_t out_w, int32_t coord_topk, float* maps_nchw, float* coords,
                                    void* chain_ws, int64_t chain_ws_bytes, const int32_t* pair_ref, void* stream) {
  FGVC_CHECK_ARG(jobs_dev && jobs_host && maps_nchw && coords && lab_bank, "fgvc_point_clip_tail: null pointer");
  FGVC_CHECK_ARG(job_end > job_begin, "fgvc_point_clip_tail: empty job range");
  const int n_pix = H * W;
  const int slot0 = jobs_host[job_begin].out_slot;
  for (int j = job_begin; j < job_end; ++j)
    FGVC_CHECK_ARG(jobs_host[j].out_slot == slot0 + (j - job_begin), "fgvc_point_clip_tail: out_slots must be consecutive");
  FGVC_CHECK_ARG(pair_ref == nullptr || chain_ws != nullptr, "fgvc_point_clip_tail_shared: needs the chain workspace");
  if (chain_ws != nullptr && (job_end - job_begin > 1 || pair_ref != nullptr)) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_gather_chain(topk_val, topk_idx, K, groups, jobs_dev, job_begin, job_end, mem_label_slot, pair_ref,
                                 n_pix, temperature, flags, lab_bank, Lp, chain_ws, chain_ws_bytes, st);
    if (rc) return rc;
    rc = launch_labels_to_nchw_jobs(lab_bank, jobs_dev, job_begin, job_end, Lp, L, n_pix, maps_nchw, st);
    if (rc) return rc;
  } else {
    for (int j = job_begin; j < job_end; ++j) {
      const int slot = jobs_host[j].out_slot;
      int rc = fgvc_gather_labels(topk_val, topk_idx, K, groups, jobs_dev, j, j + 1, mem_label_slot, n_pix,
                                  temperature, flags, lab_bank, Lp, stream);
      if (rc) return rc;
      rc = fgvc_labels_to_nchw(lab_bank, slot, Lp, L, n_pix, maps_nchw + (int64_t)slot * L * n_pix, stream);
      if (rc) return rc;
    }
  }
  return fgvc_heatmap_coords(maps_nchw + (int64_t)slot0 * L * n_pix, (job_end - job_begin) * L, H, W, out_h, out_w,
                             coord_topk, coords + (int64_t)slot0 * L * 2, stream);
}

extern "C" int fgvc_point_clip_tail(const float* topk_val, const int32_t* topk_idx, int32_t K, int32_t groups,
                                    const fgvc_job* jobs_dev, const fgvc_job* jobs_host, int32_t job_begin,
                                    int32_t job_end, const int32_t* mem_label_slot, int32_t H, int32_t W,
                                    float temperature, int32_t flags, float* lab_bank, int32_t Lp, int32_t L, int32_t out_h,
                                    int32_t out_w, int32_t coord_topk, float* maps_nchw, float* coords,
                                    void* chain_ws, int64_t chain_ws_bytes, void* stream) {
  return point_clip_tail_impl(topk_val, topk_idx, K, groups, jobs_dev, jobs_host, job_begin, job_end, mem_label_slot, H, W,
                              temperature, flags, lab_bank, Lp, L, out_h, out_w, coord_topk, maps_nchw, coords, chain_ws,
                              chain_ws_bytes, nullptr, stream);
}

// Same with SHARED top-k lists: the jobs of several with_first groups that have the same query frame look at the
// same memory frames except their first one, so K1 ran once per (query frame, memory frame) pair
// (one list per pair) and pair_ref[e] names the list of memory entry e of these jobs (indices into
// topk_val / topk_idx in units of H*W*K).  Needs the chain workspace.
extern "C" int fgvc_point_clip_tail_shared(const float* topk_val, const int32_t* topk_idx, int32_t K,
                                           const int32_t* pair_ref, const fgvc_job* jobs_dev, const fgvc_job* jobs_host,
                                           int32_t job_begin, int32_t job_end, const int32_t* mem_label_slot, int32_t H,
                                           int32_t W, float temperature, int32_t flags, float* lab_bank, int32_t Lp,
                                           int32_t L, int32_t out_h, int32_t out_w, int32_t coord_topk, float* maps_nchw,
                                           float* coords, void* chain_ws, int64_t chain_ws_bytes, void* stream) {
  FGVC_CHECK_ARG(pair_ref != nullptr, "fgvc_point_clip_tail_shared: null pair_ref");
  return point_clip_tail_impl(topk_val, topk_idx, K, 1, jobs_dev, jobs_host, job_begin, job_end, mem_label_slot, H, W,
                              temperature, flags, lab_bank, Lp, L, out_h, out_w, coord_topk, maps_nchw, coords, chain_ws,
                              chain_ws_bytes, pair_ref, stream);
}
