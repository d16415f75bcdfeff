/* jcvi_b200 -- C ABI of the B200-native MCscan anchor-detection hot path.
 *
 * Drop-in boundary for tanghaibao/jcvi's
 *     jcvi.compara.blastfilter  ->  jcvi.compara.synteny scan  ->  jcvi.compara.synteny liftover
 * The reference has no FFI for this path (it is plain Python + one Cython class); the entry
 * points below are what a maintainer would bind with ctypes/cffi from the reference's own
 * modules (INTEGRATION.md shows the stubs).  Each entry point cites the reference code it
 * replaces, file:line relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures;
 *   - every function returns 0 on success, a negative JX_E* code on failure;
 *     jx_last_error(ctx) then holds a message (valid until the next call on ctx);
 *   - input buffers are caller-owned; `*_on_device != 0` means the pointer is a CUDA device
 *     pointer on the context's device (16-byte aligned), otherwise host memory
 *     (pinned memory gives full PCIe speed);
 *   - result structs hold library-owned host arrays, released with jx_*_result_free;
 *   - one context per GPU; a context is not thread-safe; calls release nothing of the GIL's
 *     business (ctypes drops the GIL around each call).
 *   - gene "ranks" are indices into the BED sorted by (natsort_key(seqid), start, accn)
 *     exactly as src/jcvi/formats/bed.py:147,166-167,196-198 defines Bed.order.
 */
#ifndef JCVI_B200_H
#define JCVI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jx_ctx jx_ctx;

enum {
  JX_OK = 0,
  JX_ECUDA = -1,        /* CUDA runtime failure (no device, OOM, launch error)              */
  JX_EARG = -2,         /* bad argument                                                     */
  JX_EZERODIV = -3,     /* blastfilter.py:235 ZeroDivisionError (score <= 0 with cscore on) */
  JX_EUNSUPPORTED = -4, /* input outside what the device tokenizer converts exactly
                           (inf/nan/hex floats, >19 significant digits, names >= 128 bytes) */
  JX_EDUPANCHOR = -5,   /* synteny.py:397 AssertionError: duplicate (qi,si) anchors         */
  JX_EINTERNAL = -6
};

/* ---- context -------------------------------------------------------------------------- */
int jx_ctx_create(int device, jx_ctx** out);
void jx_ctx_destroy(jx_ctx* ctx);
const char* jx_last_error(jx_ctx* ctx);
const char* jx_version(void);
/* sizeof() of jx_bed, jx_filter_opts, jx_filter_result, jx_scan_opts, jx_scan_result,
 * jx_liftover_result as this library was compiled (binding self-check) */
void jx_abi_sizes(int32_t out[6]);
/* sizeof(jx_synfind_result) as compiled (the same check for the struct added with SURVEY 8(f) N3) */
int32_t jx_abi_size_synfind(void);
/* number of kernels this library launched on ctx since creation (bench.py "gpu_launches") */
int64_t jx_launch_count(jx_ctx* ctx);
/* the context's CUDA stream as an integer handle (for event timing on the launching stream) */
uint64_t jx_stream_handle(jx_ctx* ctx);
/* accumulated device time (ms, CUDA events on the context's stream) of the K1/K2 tokenizer
 * launches since the last call with reset != 0, the number of launches, the text bytes they read
 * and the lines (16-byte records) they wrote */
int jx_tokenizer_timing(jx_ctx* ctx, int reset, double* ms, int64_t* launches, int64_t* bytes, int64_t* lines);

/* ---- text residency ------------------------------------------------------------------------
 * blastfilter and liftover both read the raw hit table (blastfilter.py:48, synteny.py:1939);
 * upload it once: the copy runs asynchronously in 64 MiB chunks and the stage calls that are
 * handed the returned device pointer (text_on_device = 1, same nbytes) start tokenizing each
 * chunk as soon as it has landed.  One held text per context; released by the next upload,
 * jx_text_release or jx_ctx_destroy.  Pinned host memory gives full PCIe speed. */
int jx_text_upload(jx_ctx* ctx, const char* host_text, int64_t nbytes, uint64_t* device_ptr);
int jx_text_release(jx_ctx* ctx);
/* The same from a FILE (replaces must_open, src/jcvi/formats/base.py:473-550, for the path's input
 * table): plain files are read by several threads straight into pinned pieces and copied as they complete; blocked gzip
 * (BGZF, as bgzip writes) is inflated block-parallel by the same threads; any other gzip is streamed through one
 * inflating thread with the copies running behind it.  *kind: 0 plain, 1 gzip stream, 2 BGZF.  JX_EUNSUPPORTED for
 * bzip2 (the caller keeps its own reader).  The result is the context's held text, exactly as after jx_text_upload. */
int jx_file_upload(jx_ctx* ctx, const char* path, uint64_t* device_ptr, int64_t* nbytes, int32_t* kind);
/* jx_liftover_text on the SAME resident text re-uses the records jx_blastfilter just tokenized (the
 * reference parses the raw table twice: blastfilter.py:48 and synteny.py:1939).  on = 0 turns that
 * off (every call tokenizes); default 1.  Results are identical either way. */
int jx_set_record_reuse(jx_ctx* ctx, int on);

/* ---- BED side tables -------------------------------------------------------------------
 * Replaces the name -> (rank, BedLine) dicts of Bed.order (formats/bed.py:196-198) and the
 * per-gene attributes the path reads (seqid, start, end, accn).  All arrays are in rank order.
 */
typedef struct {
  int64_t n_genes;
  const char* names;            /* concatenated accn bytes                                      */
  const uint32_t* name_off;     /* n_genes + 1 offsets into names                               */
  const uint8_t* indexable;     /* 1 iff Bed.order[accn] == this rank (later duplicate wins)    */
  const int32_t* chr_lex;       /* byte-order rank of the gene's seqid among this BED's seqids  */
  const int32_t* chr_begin;     /* first rank of the gene's seqid run                           */
  const int32_t* chr_end;       /* one past the last rank of the run                            */
  const int32_t* length;        /* abs(end - start), start = col2 + 1 (blastfilter.py:171)      */
  const uint32_t* accn_lexrank; /* rank of accn in byte-order sort of this BED's accns          */
  const uint32_t* name_id;      /* id of accn in the union of both BEDs' accns (filter_cscore's
                                   single dict keyed by name, blastfilter.py:227-232)          */
  int32_t n_chr;
} jx_bed;

/* side: 0 = query (--qbed), 1 = subject (--sbed) */
int jx_bed_load(jx_ctx* ctx, int side, const jx_bed* bed);
/* is_self: qbed and sbed are the same file (synteny.py:494); q_same_s[qi] = subject rank whose
 * accn equals the query gene's accn, or -1 (filter_tandem's `b.query == b.subject`,
 * blastfilter.py:253); n_name_ids = size of the union namespace */
int jx_pair_setup(jx_ctx* ctx, int is_self, const int32_t* q_same_s, uint32_t n_name_ids);

/* ---- stage 1: blastfilter  (src/jcvi/compara/blastfilter.py:35-161) -------------------- */
typedef struct {
  int32_t strip_names;          /* set_stripnames(): gene_name() on both names (cbook.py:308)   */
  double cscore;                /* --cscore, 0 disables (blastfilter.py:105)                    */
  int32_t tandem_nmax;          /* --tandem_Nmax, 0 disables (blastfilter.py:113)               */
  const uint64_t* exclude_keys; /* sorted (qi << 32 | si) pairs of --exclude (205-222), or NULL */
  int64_t n_exclude;
  int32_t want_text;            /* also produce the `.filtered` text (write_new_blast 200-202)  */
  int32_t want_mothers;         /* also return the tandem mother maps (write_localdups 164-185) */
  int32_t skip_arrays;          /* do not copy q_rank/s_rank/score/line_off to the host (they stay NULL);
                                   the survivors remain on the device for jx_scan_filtered            */
} jx_filter_opts;

typedef struct {
  int64_t n;                    /* surviving hits, in output order (score desc, file order)     */
  int32_t* q_rank;              /* after tandem substitution                                    */
  int32_t* s_rank;
  float* score;                 /* float32 as parsed (cblast.pyx:88)                            */
  uint64_t* line_off;           /* byte offset of the hit's line in the input text              */
  char* text;                   /* `.filtered` bytes when want_text; owned by this result and valid
                                   until jx_filter_result_free (text_pinned = 2: a pinned buffer
                                   borrowed from the context's pool and returned to it by the free;
                                   the context must outlive the result)                         */
  int64_t text_len;
  int32_t* q_mother;            /* [n_genes(q)] when want_mothers: mother rank of every gene    */
  int32_t* s_mother;            /* [n_genes(s)]                                                 */
  /* counters: 0 lines (non-'#'), 1 mapped records, 2 unmapped names ("not in bed" warnings),
   * 3 after exclude, 4 after cscore (pre-dedupe), 5 after dedupe, 6 after tandem, 7 '#' lines */
  int64_t stats[8];
  int32_t text_pinned;          /* 0: malloc'ed; 2: pinned, borrowed from the context's pool (see text) */
  int64_t device_token;         /* identifies the device-resident copy of the survivors kept by the
                                   context (valid until the next jx_blastfilter on it)               */
  void* owner;                  /* the context whose pinned pool `text` came from (text_pinned = 2)  */
} jx_filter_result;

int jx_blastfilter(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device,
                   const jx_filter_opts* opts, jx_filter_result* out);
void jx_filter_result_free(jx_filter_result* r);

/* ---- stage 2: synteny scan  (src/jcvi/compara/synteny.py:1849-1904) -------------------- */
typedef struct {
  int32_t xdist, ydist;         /* --dist (add_arguments, synteny.py:506-523)                   */
  int32_t min_size;             /* -n / --min_size (_score >= N, synteny.py:432)                */
  int32_t intrabound;           /* --intrabound, self comparison only (synteny.py:422-427)      */
} jx_scan_opts;

typedef struct {
  int64_t n_anchors;
  int64_t n_blocks;
  int32_t* q_rank;              /* members, cluster by cluster, each cluster ascending          */
  int32_t* s_rank;
  float* score;
  int64_t* block_start;         /* n_blocks + 1 offsets                                         */
  int64_t* index;               /* scan_points only: input index of every member                */
  /* 0 lines, 1 hits imported (read_blast), 2 points in clusters before min_size, 3 '#' lines */
  int64_t stats[8];
} jx_scan_result;

/* read_blast (synteny.py:255-297) + batch_scan (437-452) on the text of a `.filtered` file */
int jx_scan_text(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names,
                 const jx_scan_opts* opts, jx_scan_result* out);
/* batch_scan / synteny_scan (synteny.py:402-452) on caller-supplied points; group[] is the
 * chromosome-pair id of each point (ascending id == sorted(chr_pair) order); is_self as in
 * synteny_scan's argument.  Points must be distinct within a group. */
int jx_scan_points(jx_ctx* ctx, const int32_t* qi, const int32_t* si, const float* score,
                   const int64_t* group, int64_t n, int is_self, const jx_scan_opts* opts,
                   jx_scan_result* out);
/* the in-memory hand-over of a blastfilter result to scan: equivalent to writing `.filtered`
 * with "%.3g" scores (cblast.pyx:17) and reading it back with read_blast */
int jx_scan_filtered(jx_ctx* ctx, const jx_filter_result* filtered, const jx_scan_opts* opts,
                     jx_scan_result* out);
void jx_scan_result_free(jx_scan_result* r);

/* ---- stage 3: synteny liftover  (src/jcvi/compara/synteny.py:1919-1971) ---------------- */
typedef struct {
  int64_t n_lifted;             /* new pairs, ordered by (block, sorted chr pair, file order)   */
  int32_t* q_rank;
  int32_t* s_rank;
  float* score;                 /* raw float32 score; the writer prints str(int(score)) + "L"   */
  int32_t* block;
  /* per input anchor line (AnchorFile.filter_blocks, compara/base.py:52-83) */
  uint8_t* accepted;            /* 1 iff the (qi,si) pair is among the de-duplicated raw hits   */
  float* raw_score;             /* its raw score when accepted                                  */
  /* 0 lines, 1 hits imported (process_blast_for_liftover), 2 anchors in BED, 3 lifted,
   * 4 hits resolved by the KD-tree walk (cross-block ties), 5 '#' lines */
  int64_t stats[8];
} jx_liftover_result;

/* a_qi/a_si: ranks of every anchor line in file order (-1 when the name is not in the BED);
 * a_block: its block index (AnchorFile.iter_pairs, compara/base.py:21-27) */
int jx_liftover_text(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device,
                     int strip_names, int dist, const int32_t* a_qi, const int32_t* a_si,
                     const int32_t* a_block, int64_t n_anchor_lines, jx_liftover_result* out);
void jx_liftover_result_free(jx_liftover_result* r);

/* ---- the three stages as one call (what catalog.ortholog runs back to back, src/jcvi/compara/catalog.py:720-750) ----
 * jx_blastfilter -> jx_scan_filtered -> jx_liftover_text(same raw text, the anchors just found, dist = liftover_dist)
 * with everything that flows between the stages left on the device: the survivors go to scan as device arrays ("%.3g"
 * score round trip applied there), the anchors go to liftover as device arrays while their copy to the host is in
 * flight, liftover re-uses the records blastfilter tokenized, the `.filtered` text (want_text) crosses PCIe on the copy
 * stream while scan and liftover run and is complete when the call returns.  Host text is uploaded (jx_text_upload: 64 MiB
 * chunks, tokenized as they land) first; device text must come from jx_text_upload / jx_file_upload.  The three results
 * are exactly those of the three separate calls. */
int jx_pipeline(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, const jx_filter_opts* fopts,
                const jx_scan_opts* sopts, int liftover_dist, jx_filter_result* fout, jx_scan_result* sout,
                jx_liftover_result* lout);

/* synteny_liftover (synteny.py:455-473) on caller-supplied integer points: nearest[i] = index
 * of the anchor cKDTree(anchors, leafsize=16).query(p=1, distance_upper_bound=dist) returns for
 * point i, or n_anchors when none is strictly closer than dist; dist_out[i] its L1 distance. */
int jx_nearest_anchor(jx_ctx* ctx, const int64_t* points_xy, int64_t n_points,
                      const int64_t* anchors_xy, int64_t n_anchors, int dist, int64_t* nearest,
                      int64_t* dist_out);

/* ---- one genome pair sharded over several GPUs ---------------------------------
 * Every rank owns a slice of the raw table cut at line boundaries.  The only value blastfilter
 * needs globally is filter_cscore's per-name best score (blastfilter.py:227-232):
 *   1. jx_cscore_begin   tokenizes the slice and accumulates the rank-local best table; it returns
 *                        the DEVICE address of that table (n_name_ids float32 values >= 0 stored
 *                        as their bit patterns, which order like integers);
 *   2. the caller max-reduces the tables in place across ranks (ncclAllReduce(ncclMax) on int32 /
 *      torch.distributed.all_reduce(MAX) over NVLink);
 *   3. jx_cscore_finish  applies the C-score predicate with the global table and returns the raw
 *                        text lines of the slice that pass, in file order.
 * Concatenating the returned lines in rank order and running jx_blastfilter on them with
 * cscore = 0 gives exactly the unsharded `.filtered` (the later stages only see survivors).
 * jx_near_lines is the same idea for liftover: the lines of a slice whose hit lies within `dist`
 * (L1, same chromosome pair) of an anchor, anchors included; jx_liftover_text on the concatenation
 * equals jx_liftover_text on the whole table.  Returned buffers are released with jx_free. */
int jx_cscore_begin(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names,
                    const uint64_t* exclude_keys, int64_t n_exclude, uint64_t* best_table_device_ptr,
                    uint32_t* n_name_ids);
int jx_cscore_finish(jx_ctx* ctx, double cscore, char** lines, int64_t* lines_len, int64_t* n_lines);
int jx_near_lines(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                  const int32_t* a_qi, const int32_t* a_si, int64_t n_anchor_lines, char** lines,
                  int64_t* lines_len, int64_t* n_lines);
/* The same two calls with the packed lines LEFT ON THE DEVICE (for ranks that exchange them GPU to GPU over
 * NVLink instead of through host memory): *lines_dev is a device pointer into a context-owned buffer, valid
 * until the next jx_*_device call on the context; nothing to free.  jx_near_lines(_device) on the slice that
 * jx_cscore_begin just tokenized (same held text, no exclusions, not a self comparison) re-uses its records. */
int jx_cscore_finish_device(jx_ctx* ctx, double cscore, uint64_t* lines_dev, int64_t* lines_len, int64_t* n_lines);
int jx_near_lines_device(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                         const int32_t* a_qi, const int32_t* a_si, int64_t n_anchor_lines, uint64_t* lines_dev,
                         int64_t* lines_len, int64_t* n_out);
void jx_free(void* p);

/* ---- one genome pair sharded over ranks BY CHROMOSOME PAIR (SURVEY.md 8e stages B / C) ----------------------------
 * From the pair de-duplication on, the reference's loops are independent per (qseqid, sseqid): (qi, si) determines the
 * pair of seqids, tandem families never leave a seqid (blastfilter.py:262-284), synteny_scan and synteny_liftover run per
 * chromosome pair (synteny.py:444-450, 1946-1956).  Global: the per-name best score (all-reduce max, as above) and the
 * tandem families (each owner finds the edges of its pairs, the edge lists are all-gathered, every rank runs the same
 * union-find).  The caller moves 16-byte CANDIDATES {qi, si, float32 score bits, (global order << 1) | (evalue < 1e-10)}
 * GPU to GPU; global order = ord_base + record index, monotone in file order over the ranks' slices, < 2^31.
 *   jx_cscore_begin -> [all-reduce max] -> jx_shard_candidates -> [all-reduce the histogram, choose owners]
 *   -> jx_shard_partition -> [all-to-all] -> jx_shard_pairs -> [all-gather edges] -> jx_shard_survivors
 *   -> jx_scan_cands (per owner) ... jx_shard_near -> jx_shard_partition -> [all-to-all] -> jx_shard_lift (per owner)
 * Device arrays returned by these calls belong to the context's session and live until jx_shard_reset. */
int jx_shard_reset(jx_ctx* ctx);
/* record slots of the slice jx_cscore_begin retained: ord_base of rank r = sum of the slots of the ranks before it */
int jx_shard_slots(jx_ctx* ctx, int64_t* n_slots);
/* candidates of this rank's slice that pass the C-score filter + their count per chromosome pair
 * (cp = chr_lex(q) * n_chr(s) + chr_lex(s); cp_hist[n_cp] is caller-owned host memory) */
int jx_shard_candidates(jx_ctx* ctx, double cscore, uint32_t ord_base, uint64_t* cand_dev, int64_t* n, uint32_t* cp_hist, int64_t n_cp);
/* candidates grouped by destination rank owner[cp]: device array + counts[world] */
int jx_shard_partition(jx_ctx* ctx, uint64_t cand_dev, int64_t n, const int32_t* owner, int64_t n_cp, int world, uint64_t* out_dev,
                       int64_t* counts, const uint32_t* cp_hist_in /* [n_cp] as jx_shard_candidates returned it, or NULL: counted here */,
                       uint32_t* cp_hist_out /* [n_cp] or NULL: the candidates' count per chromosome pair */);
/* owner: per-pair best record of the received candidates (blastfilter.py:86-89) + the tandem edges of its pairs as
 * device arrays of (rank a, rank b) uint32 pairs, query side and subject side */
int jx_shard_pairs(jx_ctx* ctx, uint64_t recv_dev, int64_t n, int tandem_nmax, uint64_t* qedges_dev, int64_t* nq, uint64_t* sedges_dev,
                   int64_t* ns);
/* owner: ALL ranks' edges -> families, mothers (identical on every rank; copied to q_mother / s_mother when given),
 * mother substitution and second per-pair best (blastfilter.py:240-259) -> this rank's survivors (device candidates) */
int jx_shard_survivors(jx_ctx* ctx, uint64_t qedges_dev, int64_t nq, uint64_t sedges_dev, int64_t ns, int tandem_nmax,
                       uint64_t* surv_dev, int64_t* n_surv, int32_t* q_mother, int32_t* s_mother);
/* `.filtered` order (score descending, global order ascending) of a set of survivors -> host columns of `out`
 * (line_off receives the global order) */
int jx_shard_order(jx_ctx* ctx, uint64_t surv_dev, int64_t n, jx_filter_result* out);
/* scan of survivors given as device candidates (the hand-over of jx_scan_filtered).  dev_rows != NULL: the anchors are
 * also left on the device as int32 rows {qi, si, score bits, block}; host_result = 0 then skips the host arrays of `out` */
int jx_scan_cands(jx_ctx* ctx, uint64_t surv_dev, int64_t n, const jx_scan_opts* opts, jx_scan_result* out, uint64_t* dev_rows,
                  int host_result);
/* this rank's raw records within dist of any anchor (records retained by jx_cscore_begin, or `text`):
 * device array of {qi, si, score bits, ord_base + record index} */
int jx_shard_near(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist, const int32_t* a_qi,
                  const int32_t* a_si, int64_t n_anchor_lines, int anchors_on_device, uint32_t ord_base, uint64_t* cand_dev, int64_t* n);
/* owner: the near candidates received + all anchor lines (global block ids) -> jx_liftover_result for this rank's
 * chromosome pairs (lifted pairs ordered by (block, chromosome pair, global order)) */
int jx_shard_lift(jx_ctx* ctx, uint64_t recv_dev, int64_t n, int64_t n_slots_total, int dist, const int32_t* a_qi, const int32_t* a_si,
                  const int32_t* a_block, int64_t n_anchor_lines, int anchors_on_device, jx_liftover_result* out,
                  uint64_t* dev_out /* NULL, or [2]: the result stays on the device (session memory): [0] int32 columns
                                       {q | s | score bits | block} x n_lifted, [1] int64 per anchor line accepted << 32 | raw score bits;
                                       the host arrays of `out` are then not filled */);

/* ---- jcvi.formats.blast cscore (src/jcvi/formats/blast.py:1098-1189) ------------
 * `cscore` has no BED: its dictionaries are keyed by whatever (gene_name-stripped) names the table
 * holds.  jx_distinct_names returns one representative raw line per distinct name of the table
 * (`\n`-terminated, first occurrence) and, per representative, which column the name is in
 * (side[i]: 0 = query, 1 = subject).  The caller turns them into a name list, loads it as the name
 * table of both sides (jx_bed_load, jx_pair_setup) and runs jx_cscore_begin / jx_cscore_finish with
 * the cutoff; jx_session_counts then must report 0 unmapped lines -- anything else means two
 * different names shared the 64-bit hash of this `seed` (retry with another seed) or a line with
 * fewer than two fields (counts[2] of jx_distinct_names' stats).  stats: 0 non-'#' lines, 1 lines
 * with two names, 2 lines with fewer than two fields, 3 '#' lines.  Buffers: jx_free. */
int jx_distinct_names(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names,
                      uint32_t seed, char** lines, int64_t* lines_len, uint8_t** side, int64_t* n_names,
                      int64_t stats[4]);
/* The whole of `cscore` in one call: the cscore output file as text (`query\tsubject\t{:.2f}[\t{:.1f}]\n`,
 * sorted by (query, subject)) and, with want_blast, the `--writeblast` file (str(BlastLine) per pair).
 * JX_EUNSUPPORTED when a line has fewer than two fields (the reference raises ZeroDivisionError or keys
 * its dictionaries with ""), JX_EZERODIV like the reference when a line's names never scored.  Replaces
 * the name tables loaded with jx_bed_load.  stats as jx_distinct_names.  Buffers: jx_free. */
int jx_blast_cscore(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names,
                    double cutoff, int want_pct, int want_blast, char** out, int64_t* out_len,
                    char** blast_out, int64_t* blast_len, int64_t stats[4]);
/* `jcvi.formats.blast filter` (src/jcvi/formats/blast.py:620-709): the rows whose BlastLine passes
 *   remove = score < o.score or pctid < o.pctid or hitlen < o.hitlen or evalue > o.evalue or noids
 *   if inverse: remove = not remove;   remove = remove or (noself and query == subject)
 * in file order, each as row.rstrip() + "\n" ('#' rows are skipped).  use_ids: `query in ids and subject
 * in ids` on the raw names, against the name list loaded on both sides with jx_bed_load.  Comparisons are
 * the reference's: float32 score / pctid widened to double, C int hitlen, strtod(evalue) > cutoff decided
 * exactly (JX_EUNSUPPORTED when a row would need more than the exact routes).  stats: 0 rows, 1 kept,
 * 2 removed, 3 '#' rows.  Buffer: jx_free. */
int jx_blast_filter(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, double score,
                    double pctid, int64_t hitlen, double evalue, int noself, int inverse, int use_ids,
                    char** out, int64_t* out_len, int64_t stats[4]);
/* `jcvi.compara.synteny mcscan` (src/jcvi/compara/synteny.py:1538-1652): the chains of the block ranges
 * (range_chain, src/jcvi/utils/range.py:412-463, repeated on the blocks that are left until max_iter chains,
 * synteny.py:1601-1620) and the gene x track table of synteny.py:1622-1637.
 *   start / end / score / block_id [n_ranges]: AnchorFile.make_ranges' Range list (compara/base.py:29-50), in its order
 *   pair_gene / pair_partner / pair_block [n_pairs]: the block_pairs dictionaries flattened, one entry per
 *       (block, gene): gene = caller's dense id (< n_genes) of the reference-BED accn, partner = caller's id of its partner
 *   -> n_tracks; track_len[n_tracks] and track_blocks (block ids of every chain, in chain order, concatenated);
 *      track_score[n_tracks] (the chain scores the reference logs); table[n_genes * n_tracks] = partner id of the
 *      first block of the track that holds the gene, -1 for ".".  All four arrays: jx_free. */
int jx_mcscan(jx_ctx* ctx, const double* start, const double* end, const int64_t* score, const int32_t* block_id,
              int64_t n_ranges, int32_t max_iter, const int32_t* pair_gene, const int32_t* pair_partner,
              const int32_t* pair_block, int64_t n_pairs, int32_t n_genes, int32_t n_blocks, int32_t* n_tracks,
              int32_t** track_len, int32_t** track_blocks, int64_t** track_score, int32_t** table);
/* `jcvi.compara.synteny simple` (src/jcvi/compara/synteny.py:1137-1316), the per-block reductions of its loop body
 * (1206-1228): for the anchors (block, query rank, subject rank) [n] of an anchors file with n_blocks blocks ->
 * per block: min / max rank on both sides (astarti, aendi, bstarti, bendi), size (= len(block)), the number of
 * distinct ranks per side (sizeA, sizeB), and sum(x'), sum(y'), sum(x'y') of the ranks centred at the block minima --
 * get_orientation's slope (222-236) has the sign of  size * sum_qs - sum_q * sum_s.  Caller-owned arrays of n_blocks. */
int jx_block_stats(jx_ctx* ctx, const int32_t* block, const int32_t* qrank, const int32_t* srank, int64_t n,
                   int32_t n_blocks, int32_t* qmin, int32_t* qmax, int32_t* smin, int32_t* smax, uint32_t* size,
                   uint32_t* distinct_q, uint32_t* distinct_s, uint64_t* sum_q, uint64_t* sum_s, uint64_t* sum_qs);
/* ---- jcvi.compara.synfind (src/jcvi/compara/synfind.py:68-240) ---------------------------------
 * For every gene of the query BED: the syntenic regions find_synteny_region (68-127, "collinear" scoring) returns for
 * the hits of the window // 2 genes around it, in batch_query's order (130-202: genes in BED order, regions by
 * decreasing score, equal scores in the order of `sorted(Grouper)`), then the same from the subject side ("mirror",
 * synfind.py:229-232; not for a self comparison).  Rows [0, n_forward) are the forward pass, [n_forward, n_forward +
 * n_mirror) the mirror pass, where `query` is a rank of the SUBJECT BED and syntelog / left / right are ranks of the
 * query BED.  The caller prints accn(query), accn(syntelog), gray, score, the flank distance from left / right
 * (181-185), orientation and the notes. */
typedef struct {
  int64_t n_forward, n_mirror;
  int64_t n_points;             /* distinct (qi, si) pairs imported (read_blast, synteny.py:255-297)  */
  int32_t* query;               /* rank of the gene the region was found for                          */
  int32_t* syntelog;            /* rank of the closest flanker's partner (the anchor column)          */
  int32_t* left;                /* subject rank of the first / last point of the chosen track         */
  int32_t* right;
  int32_t* score;
  uint8_t* gray;                /* 0 "S" syntelog, 1 "G" gray (query outside the group's x range), 2 "F" flanked */
  uint8_t* minus;               /* 1: orientation "-" (the LDS was strictly longer)                   */
  /* 0 lines, 1 distinct pairs, 2 '#' lines */
  int64_t stats[8];
} jx_synfind_result;
/* window_half = opts.window // 2, cutoff = int(opts.cutoff * opts.window) (synfind.py:132-133) */
int jx_synfind_text(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int window_half,
                    int cutoff, jx_synfind_result* out);
/* batch_query (synfind.py:130-202) on caller-supplied DISTINCT (qi, si) pairs; passes: 1 = as given, 2 = transposed
 * (transpose=True: BEDs swapped), 3 = both */
int jx_synfind_points(jx_ctx* ctx, const int32_t* qi, const int32_t* si, int64_t n, int window_half, int cutoff, int passes,
                      jx_synfind_result* out);
void jx_synfind_result_free(jx_synfind_result* r);

/* float32 best score per name id of the open jx_cscore_begin session (first n entries) */
int jx_best_table_copy(jx_ctx* ctx, float* out, uint32_t n);
/* counters of the records jx_cscore_begin retained: 0 non-'#' lines, 1 mapped, 2 unmapped */
int jx_session_counts(jx_ctx* ctx, int64_t out[3]);
/* HOST helper: do the lines of a table come in runs of equal queries (the order lastal / BLAST write)?  Looks at the
 * first 256 KiB; 1 when at least half of the lines repeat the first field of the line before.  This is the test that
 * picks the tokenizer instantiation which merges a warp's equal-query best-score reductions (results are the same
 * either way). */
int jx_host_query_runs(const char* text, int64_t nbytes);
/* HOST helper for the few lines that survive a filter: BlastLine(line) and str(BlastLine) with the
 * very glibc calls cblast.pyx makes (sscanf format cblast.pyx:15, start<=stop swap 114-120, sprintf
 * format cblast.pyx:17).  query128 / subject128: char[128]; returns the number of converted fields. */
int jx_host_blastline(const char* line, char* query128, char* subject128, float* pctid, float* score,
                      char* str_out, int str_cap);

#ifdef __cplusplus
}
#endif
#endif /* JCVI_B200_H */
