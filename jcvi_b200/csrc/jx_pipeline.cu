// The three stages as ONE call: blastfilter -> scan -> liftover on one genome pair, with everything that flows between
// the stages left on the device.
//
// What `jcvi.compara.catalog ortholog` does with two processes' worth of file traffic (src/jcvi/compara/catalog.py:720-750:
// blastfilter_main writes `.filtered`, scan reads it back and writes `.anchors`, liftover reads `.anchors` and the raw
// table again) is here: the survivors of blastfilter go to scan as device arrays (the "%.3g" text round trip of the
// scores is applied on the device), the anchors scan finds go to liftover as device arrays while their copy to the host
// is still in flight, and liftover re-uses the records blastfilter tokenized from the same resident table.  Results are
// those of jx_blastfilter + jx_scan_filtered + jx_liftover_text (tests/test_gpu_pipeline.py).
#include "jx_internal.cuh"

extern "C" int jx_pipeline(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, const jx_filter_opts* fopts,
                           const jx_scan_opts* sopts, int liftover_dist, jx_filter_result* fout, jx_scan_result* sout,
                           jx_liftover_result* lout) {
  if (!ctx || !fopts || !sopts || !fout || !sout || !lout || liftover_dist < 0) return JX_EARG;
  memset(fout, 0, sizeof *fout); memset(sout, 0, sizeof *sout); memset(lout, 0, sizeof *lout);
  if (!ctx->pair_ready) { ctx->err = "jx_bed_load / jx_pair_setup first"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  int rc;
  // the records must outlive the first stage: the text is held by the context (uploaded here when it is host memory)
  uint64_t dptr = (uint64_t)(uintptr_t)text;
  const bool held = text_on_device && ctx->held_text && (const uint8_t*)text == ctx->held_text && nbytes == ctx->held_n;
  if (!text_on_device) { if ((rc = jx_text_upload(ctx, text, nbytes, &dptr))) return rc; }
  else if (!held) { ctx->err = "jx_pipeline: device text must come from jx_text_upload"; return JX_EARG; }
  const bool reuse0 = ctx->record_reuse;
  ctx->record_reuse = true;
  // the `.filtered` text crosses PCIe on the copy stream while scan and liftover run; whatever path leaves this
  // function waits for it first (the result's host buffer must be complete -- and quiet -- when the caller sees it)
  struct TextWait {
    jx_ctx* c;
    ~TextWait() { if (c->text_in_flight) { cudaStreamSynchronize(c->copy_stream); c->text_in_flight = false; } c->defer_text = false; }
  } text_wait{ctx};
  ctx->defer_text = true;
  rc = jx_blastfilter(ctx, (const char*)(uintptr_t)dptr, nbytes, 1, fopts, fout);
  ctx->defer_text = false;
  ctx->record_reuse = reuse0;
  if (rc) return rc;
  Arena arena(ctx);
  ScanDev dev;
  dev.defer_copy = true;
  if ((rc = jx_scan_survivors(ctx, arena, sopts, sout, &dev))) return rc;
  auto& c = ctx->cache;
  const bool have_recs = c.valid && c.plain && !ctx->is_self;
  if (have_recs) {
    c.valid = false;                              // one-shot, like jx_liftover_text
    rc = jx_liftover_core(ctx, arena, c.recs, c.n_lines, c.stats, liftover_dist, dev.q, dev.s, dev.block, dev.n, nullptr, nullptr,
                          nullptr, lout);
  }
  // the scan result's host half (its copy ran on the copy stream next to liftover)
  if (dev.stage) {
    cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
    if (e != cudaSuccess) { ctx->err = std::string("jx_pipeline: ") + cudaGetErrorString(e); return JX_ECUDA; }
    scan_finish_host(sout, dev.stage, dev.n);
  }
  if (rc) return rc;
  if (!have_recs) {
    // self comparisons / exclusion lists: liftover tokenizes with its own rules (near-diagonal drop, no exclusions)
    std::vector<int32_t> blk((size_t)sout->n_anchors);
    for (int64_t b = 0; b < sout->n_blocks; ++b)
      for (int64_t t = sout->block_start[b]; t < sout->block_start[b + 1]; ++t) blk[(size_t)t] = (int32_t)b;
    rc = jx_liftover_text(ctx, (const char*)(uintptr_t)dptr, nbytes, 1, fopts->strip_names, liftover_dist, sout->q_rank, sout->s_rank,
                          blk.data(), sout->n_anchors, lout);
  }
  return rc;
}
