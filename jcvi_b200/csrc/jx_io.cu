// the reader in front of the tokenizer.  Replaces `must_open` (src/jcvi/formats/base.py:473-550: plain
// files, `.gz` through the gzip module -- one Python thread inflating into pageable memory) for the path's input table:
//
//   jx_file_upload(path) -> the table resident on the device (same contract as jx_text_upload: one held text per context,
//   one event per 64 MiB chunk, so the tokenizer starts on chunk c while chunk c + 1 is still being read / inflated).
//
//   * plain file: reader threads pread() 16 MiB pieces straight into a ring of PINNED buffers, the calling thread issues
//     the host->device copies in file order as the pieces complete (no pageable staging copy inside the driver, several
//     page-cache readers in flight);
//   * blocked gzip (BGZF, what bgzip / htslib write: every member carries its compressed size in a 'BC' extra field and
//     its uncompressed size in ISIZE): the block table is walked once, uncompressed offsets follow from a prefix sum, and
//     the threads inflate whole blocks independently into the pinned ring;
//   * any other gzip (one member, or members without the BGZF field): one thread streams zlib's inflate into the ring while
//     the copies of finished pieces run behind it; the device buffer grows when the size hint (ISIZE) was too small;
//   * bzip2 and everything else: JX_EUNSUPPORTED -- the caller keeps its own reader.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "jx_internal.cuh"

namespace {

constexpr int64_t IO_SLOT = 16ll << 20;          // one pinned piece
constexpr int IO_NSLOT = 8;                      // ring of pinned pieces (128 MiB)
constexpr int64_t IO_CHUNK = 64ll << 20;         // event granularity of a held text (jx_tokenize.cu)

struct Piece { int64_t dev_off, len; };

// ring of pinned pieces + the in-order copy loop of the calling thread
struct Ring {
  jx_ctx* ctx;
  uint8_t* dev;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<int> state;                        // per piece: 0 not started, 1 filled, 2 copy issued, -1 failed
  std::vector<Piece> pieces;
  cudaEvent_t done[IO_NSLOT];
  bool have_event[IO_NSLOT];
  std::string err;
  bool failed = false;
  uint8_t* slot(int64_t k) const { return (uint8_t*)ctx->io_pin + (k % IO_NSLOT) * IO_SLOT; }
};

int ring_reserve(jx_ctx* ctx) {
  if (ctx->io_pin) return JX_OK;
  if (cudaMallocHost((void**)&ctx->io_pin, (size_t)IO_SLOT * IO_NSLOT) != cudaSuccess) {
    cudaGetLastError();
    ctx->err = "cudaMallocHost of the reader's pinned ring failed";
    return JX_ECUDA;
  }
  return JX_OK;
}

// the held text of the context with room for nbytes (+ 16); earlier work on it has finished
int held_reserve(jx_ctx* ctx, int64_t nbytes) {
  ctx->cache.valid = false;
  if (ctx->held_text && ctx->held_cap >= nbytes + 16) {
    JX_CK(cudaStreamSynchronize(ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->copy_stream));
    for (cudaEvent_t ev : ctx->held_events) cudaEventDestroy(ev);
    ctx->held_events.clear();
  } else {
    jx_text_release(ctx);
    uint8_t* buf = nullptr;
    JX_CK(cudaMalloc((void**)&buf, (size_t)nbytes + 16));
    ctx->held_text = buf;
    ctx->held_cap = nbytes + 16;
  }
  return JX_OK;
}

// a worker may overwrite ring slot k % NSLOT once the copy of piece k - NSLOT has finished
bool ring_wait_free(Ring& R, int64_t k) {
  if (k < IO_NSLOT) return true;
  const int64_t prev = k - IO_NSLOT;
  {
    std::unique_lock<std::mutex> lk(R.mu);
    R.cv.wait(lk, [&] { return R.state[(size_t)prev] == 2 || R.failed; });
    if (R.failed) return false;
  }
  return cudaEventSynchronize(R.done[prev % IO_NSLOT]) == cudaSuccess;
}
void ring_filled(Ring& R, int64_t k, bool ok, const char* why = nullptr) {
  std::lock_guard<std::mutex> lk(R.mu);
  R.state[(size_t)k] = ok ? 1 : -1;
  if (!ok) { R.failed = true; if (why && R.err.empty()) R.err = why; }
  R.cv.notify_all();
}

// calling thread: copies in piece order; one held event per 64 MiB chunk of device text
int ring_copy_loop(Ring& R, int64_t npieces, int64_t total) {
  jx_ctx* ctx = R.ctx;
  const int64_t nchunks = (total + IO_CHUNK - 1) / IO_CHUNK;
  for (int64_t k = 0; k < npieces; ++k) {
    {
      std::unique_lock<std::mutex> lk(R.mu);
      R.cv.wait(lk, [&] { return R.state[(size_t)k] != 0 || R.failed; });
      if (R.failed) { ctx->err = R.err.empty() ? "reader failed" : R.err; return JX_EARG; }
    }
    const Piece& p = R.pieces[(size_t)k];
    if (p.len) JX_CK(cudaMemcpyAsync(R.dev + p.dev_off, R.slot(k), (size_t)p.len, cudaMemcpyHostToDevice, ctx->copy_stream));
    const int s = (int)(k % IO_NSLOT);
    if (!R.have_event[s]) { JX_CK(cudaEventCreateWithFlags(&R.done[s], cudaEventDisableTiming)); R.have_event[s] = true; }
    JX_CK(cudaEventRecord(R.done[s], ctx->copy_stream));
    const int64_t end = p.dev_off + p.len;
    while ((int64_t)ctx->held_events.size() < nchunks) {       // chunk c is complete once the pieces up to its end are issued
      const int64_t cend = std::min<int64_t>(((int64_t)ctx->held_events.size() + 1) * IO_CHUNK, total);
      if (end < cend) break;
      cudaEvent_t ev;
      JX_CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      JX_CK(cudaEventRecord(ev, ctx->copy_stream));
      ctx->held_events.push_back(ev);
    }
    {
      std::lock_guard<std::mutex> lk(R.mu);
      R.state[(size_t)k] = 2;
      R.cv.notify_all();
    }
  }
  return JX_OK;
}

void ring_finish(Ring& R) {
  // the last pieces may still be in flight: the ring is only reused by the next jx_file_upload, which waits for the copy
  // stream first; destroying a recorded event releases it once it has completed
  for (int s = 0; s < IO_NSLOT; ++s) if (R.have_event[s]) cudaEventDestroy(R.done[s]);
}

int n_threads() {
  const char* e = getenv("JX_IO_THREADS");
  int t = e ? atoi(e) : 0;
  if (t <= 0) { t = (int)std::thread::hardware_concurrency(); if (t <= 0) t = 4; if (t > 8) t = 8; }
  return std::min(t, IO_NSLOT - 1);
}

// ---- plain file ----------------------------------------------------------------------------------------------------
int upload_plain(jx_ctx* ctx, int fd, int64_t size) {
  int rc;
  if ((rc = held_reserve(ctx, size))) return rc;
  Ring R;
  R.ctx = ctx; R.dev = ctx->held_text;
  memset(R.have_event, 0, sizeof R.have_event);
  const int64_t np = (size + IO_SLOT - 1) / IO_SLOT;
  R.state.assign((size_t)np, 0);
  for (int64_t k = 0; k < np; ++k) R.pieces.push_back({k * IO_SLOT, std::min(IO_SLOT, size - k * IO_SLOT)});
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    cudaSetDevice(ctx->device);
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= np) return;
      if (!ring_wait_free(R, k)) { ring_filled(R, k, false, "copy failed"); return; }
      const Piece& p = R.pieces[(size_t)k];
      uint8_t* dst = R.slot(k);
      int64_t got = 0;
      while (got < p.len) {
        const ssize_t r = pread(fd, dst + got, (size_t)(p.len - got), (off_t)(p.dev_off + got));
        if (r <= 0) { ring_filled(R, k, false, "short read (file changed while it was read?)"); return; }
        got += r;
      }
      if (k == 0) ctx->tok_grouped = jx_sample_query_runs(dst, (size_t)p.len);
      ring_filled(R, k, true);
    }
  };
  std::vector<std::thread> th;
  const int T = (int)std::min<int64_t>(n_threads(), std::max<int64_t>(np, 1));
  for (int t = 0; t < T; ++t) th.emplace_back(worker);
  rc = ring_copy_loop(R, np, size);
  if (rc) { std::lock_guard<std::mutex> lk(R.mu); R.failed = true; R.cv.notify_all(); }
  for (auto& t : th) t.join();
  ring_finish(R);
  if (rc) return rc;
  ctx->held_n = size;
  return JX_OK;
}

// ---- gzip ----------------------------------------------------------------------------------------------------------
struct GzBlock { int64_t off, clen, ulen; uint32_t crc; };   // deflate stream of the member, its length, uncompressed length, CRC32

// BGZF: every member is `1f 8b 08 04 .. XLEN [.. 'B' 'C' 02 00 BSIZE ..] deflate CRC32 ISIZE` with BSIZE = member size - 1
bool bgzf_table(const uint8_t* z, int64_t n, std::vector<GzBlock>& blocks) {
  int64_t off = 0;
  while (off < n) {
    if (n - off < 28 || z[off] != 0x1f || z[off + 1] != 0x8b || z[off + 2] != 8 || !(z[off + 3] & 4)) return false;
    const int xlen = z[off + 10] | (z[off + 11] << 8);
    if (off + 12 + xlen > n) return false;
    int64_t bsize = -1;
    for (int x = 0; x + 4 <= xlen;) {
      const uint8_t* f = z + off + 12 + x;
      const int slen = f[2] | (f[3] << 8);
      if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (f[4] | (f[5] << 8)) + 1;
      x += 4 + slen;
    }
    if (bsize < 0 || off + bsize > n || bsize < 12 + xlen + 8) return false;
    const uint8_t* t = z + off + bsize - 4;
    const int64_t isize = (int64_t)t[0] | ((int64_t)t[1] << 8) | ((int64_t)t[2] << 16) | ((int64_t)t[3] << 24);
    if (isize > 65536) return false;
    const uint32_t crc = (uint32_t)t[-4] | ((uint32_t)t[-3] << 8) | ((uint32_t)t[-2] << 16) | ((uint32_t)t[-1] << 24);
    blocks.push_back({off + 12 + xlen, bsize - 12 - xlen - 8, isize, crc});   // the raw deflate stream of the member
    off += bsize;
  }
  return true;
}

int upload_bgzf(jx_ctx* ctx, const uint8_t* z, const std::vector<GzBlock>& blocks) {
  int rc;
  // pieces = runs of whole blocks of <= IO_SLOT uncompressed bytes
  std::vector<Piece> pieces;
  std::vector<size_t> first;                      // first block of every piece (+ sentinel)
  int64_t total = 0;
  {
    int64_t cur_off = 0, cur_len = 0;
    first.push_back(0);
    for (size_t b = 0; b < blocks.size(); ++b) {
      const int64_t u = blocks[b].ulen;
      if (cur_len && cur_len + u > IO_SLOT) {
        pieces.push_back({cur_off, cur_len});
        first.push_back(b);
        cur_off += cur_len; cur_len = 0;
      }
      cur_len += u;
    }
    pieces.push_back({cur_off, cur_len});
    first.push_back(blocks.size());
    total = cur_off + cur_len;
  }
  if ((rc = held_reserve(ctx, total))) return rc;
  Ring R;
  R.ctx = ctx; R.dev = ctx->held_text;
  memset(R.have_event, 0, sizeof R.have_event);
  R.pieces = pieces;
  const int64_t np = (int64_t)pieces.size();
  R.state.assign((size_t)np, 0);
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    cudaSetDevice(ctx->device);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) {
      std::lock_guard<std::mutex> lk(R.mu);
      R.failed = true; R.err = "inflateInit2 failed"; R.cv.notify_all();
      return;
    }
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= np) break;
      if (!ring_wait_free(R, k)) { ring_filled(R, k, false, "copy failed"); break; }
      uint8_t* dst = R.slot(k);
      bool ok = true;
      for (size_t b = first[(size_t)k]; b < first[(size_t)k + 1] && ok; ++b) {
        const GzBlock& B = blocks[b];
        if (B.ulen == 0) continue;                    // the empty end-of-file block
        inflateReset(&zs);
        zs.next_in = const_cast<Bytef*>(z + B.off); zs.avail_in = (uInt)B.clen;
        zs.next_out = dst; zs.avail_out = (uInt)B.ulen;
        const int r = inflate(&zs, Z_FINISH);
        ok = (r == Z_STREAM_END) && zs.avail_out == 0 && (uint32_t)crc32(0L, dst, (uInt)B.ulen) == B.crc;   // gzip.open checks it too
        dst += B.ulen;
      }
      if (ok && k == 0) ctx->tok_grouped = jx_sample_query_runs(R.slot(k), (size_t)R.pieces[0].len);
      ring_filled(R, k, ok, "corrupt BGZF block");
      if (!ok) break;
    }
    inflateEnd(&zs);
  };
  std::vector<std::thread> th;
  const int T = (int)std::min<int64_t>(n_threads(), std::max<int64_t>(np, 1));
  for (int t = 0; t < T; ++t) th.emplace_back(worker);
  rc = ring_copy_loop(R, np, total);
  if (rc) { std::lock_guard<std::mutex> lk(R.mu); R.failed = true; R.cv.notify_all(); }
  for (auto& t : th) t.join();
  ring_finish(R);
  if (rc) return rc;
  ctx->held_n = total;
  return JX_OK;
}

// any gzip: one inflating thread, pieces copied behind it; concatenated members like gzip.open()
int upload_gzip_stream(jx_ctx* ctx, const uint8_t* z, int64_t n) {
  int rc;
  // size hint: ISIZE of the last member (exact for one member < 4 GiB), at least 3 x the compressed size
  int64_t hint = 0;
  if (n >= 4) { const uint8_t* t = z + n - 4; hint = (int64_t)t[0] | ((int64_t)t[1] << 8) | ((int64_t)t[2] << 16) | ((int64_t)t[3] << 24); }
  int64_t cap = std::max<int64_t>(hint, 3 * n) + IO_SLOT;
  if ((rc = held_reserve(ctx, cap))) return rc;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 15 + 32) != Z_OK) { ctx->err = "inflateInit2 failed"; return JX_EINTERNAL; }
  zs.next_in = const_cast<Bytef*>(z);
  int64_t in_left = n, total = 0, k = 0;
  cudaEvent_t done[IO_NSLOT];
  bool have[IO_NSLOT] = {false};
  bool end = false;
  rc = JX_OK;
  while (!end && rc == JX_OK) {
    const int s = (int)(k % IO_NSLOT);
    if (have[s]) cudaEventSynchronize(done[s]);
    uint8_t* dst = (uint8_t*)ctx->io_pin + (size_t)s * IO_SLOT;
    int64_t got = 0;
    while (got < IO_SLOT) {
      if (zs.avail_in == 0) { const int64_t take = std::min<int64_t>(in_left, 1ll << 30); zs.avail_in = (uInt)take; in_left -= take; }
      zs.next_out = dst + got;
      zs.avail_out = (uInt)(IO_SLOT - got);
      const int r = inflate(&zs, Z_NO_FLUSH);
      got = IO_SLOT - zs.avail_out;
      if (r == Z_STREAM_END) {
        if (zs.avail_in == 0 && in_left == 0) { end = true; break; }
        // another member follows (gzip.open concatenates them); trailing zero padding ends the file
        if (zs.next_in[0] != 0x1f) { end = true; break; }
        if (inflateReset(&zs) != Z_OK) { ctx->err = "inflateReset failed"; rc = JX_EINTERNAL; break; }
      } else if (r != Z_OK) {
        if (r == Z_BUF_ERROR && zs.avail_in == 0 && in_left == 0) { ctx->err = "gzip stream ends early"; }
        else ctx->err = std::string("corrupt gzip stream: ") + (zs.msg ? zs.msg : "inflate failed");
        rc = JX_EARG; break;
      }
    }
    if (rc) break;
    if (total + got + 16 > ctx->held_cap) {              // the hint was too small: grow the device text
      const int64_t ncap = std::max<int64_t>(ctx->held_cap * 2, total + got + IO_SLOT);
      uint8_t* nb = nullptr;
      if (cudaMalloc((void**)&nb, (size_t)ncap) != cudaSuccess) { cudaGetLastError(); ctx->err = "device text allocation failed"; rc = JX_ECUDA; break; }
      cudaMemcpyAsync(nb, ctx->held_text, (size_t)total, cudaMemcpyDeviceToDevice, ctx->copy_stream);
      cudaStreamSynchronize(ctx->copy_stream);
      cudaFree(ctx->held_text);
      ctx->held_text = nb; ctx->held_cap = ncap;
    }
    if (k == 0) ctx->tok_grouped = jx_sample_query_runs(dst, (size_t)got);
    if (got) {
      if (cudaMemcpyAsync(ctx->held_text + total, dst, (size_t)got, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess) {
        ctx->err = "host->device copy failed"; rc = JX_ECUDA; break;
      }
      if (!have[s]) { cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming); have[s] = true; }
      cudaEventRecord(done[s], ctx->copy_stream);
    }
    total += got;
    ++k;
  }
  inflateEnd(&zs);
  cudaStreamSynchronize(ctx->copy_stream);
  for (int s = 0; s < IO_NSLOT; ++s) if (have[s]) cudaEventDestroy(done[s]);
  if (rc) return rc;
  // the size was not known while the copies were issued: one event per chunk, all already complete
  for (int64_t off = 0; off < total; off += IO_CHUNK) {
    cudaEvent_t ev;
    JX_CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    JX_CK(cudaEventRecord(ev, ctx->copy_stream));
    ctx->held_events.push_back(ev);
  }
  ctx->held_n = total;
  return JX_OK;
}

}  // namespace

extern "C" int jx_file_upload(jx_ctx* ctx, const char* path, uint64_t* device_ptr, int64_t* nbytes, int32_t* kind) {
  if (!ctx || !path || !device_ptr || !nbytes) return JX_EARG;
  cudaSetDevice(ctx->device);
  const int fd = open(path, O_RDONLY);
  if (fd < 0) { ctx->err = std::string("cannot open ") + path; return JX_EARG; }
  struct stat st;
  if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); ctx->err = std::string("not a regular file: ") + path; return JX_EARG; }
  const int64_t size = (int64_t)st.st_size;
  uint8_t magic[4] = {0, 0, 0, 0};
  if (size >= 4 && pread(fd, magic, 4, 0) != 4) { close(fd); ctx->err = "read failed"; return JX_EARG; }
  cudaStreamSynchronize(ctx->copy_stream);        // copies of an earlier call out of the pinned ring
  int rc = ring_reserve(ctx);
  int32_t k = 0;
  if (rc == JX_OK) {
    if (size >= 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') {
      ctx->err = "bzip2 input: not inflated by the library"; rc = JX_EUNSUPPORTED;
    } else if (size >= 18 && magic[0] == 0x1f && magic[1] == 0x8b) {
      void* m = mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { ctx->err = "mmap failed"; rc = JX_EARG; }
      else {
        madvise(m, (size_t)size, MADV_SEQUENTIAL);
        std::vector<GzBlock> blocks;
        if (bgzf_table((const uint8_t*)m, size, blocks)) { k = 2; rc = upload_bgzf(ctx, (const uint8_t*)m, blocks); }
        else { k = 1; rc = upload_gzip_stream(ctx, (const uint8_t*)m, size); }
        munmap(m, (size_t)size);
      }
    } else {
      rc = upload_plain(ctx, fd, size);
    }
  }
  close(fd);
  if (rc) return rc;
  *device_ptr = (uint64_t)(uintptr_t)ctx->held_text;
  *nbytes = ctx->held_n;
  if (kind) *kind = k;
  return JX_OK;
}
