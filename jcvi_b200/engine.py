"""Per-GPU engine: owns one jx_ctx (include/jcvi_b200.h) and the BED pair loaded on it."""
import ctypes

import numpy as np

from . import _cabi as C

_ERR = {
    C.JX_ECUDA: RuntimeError,
    C.JX_EARG: ValueError,
    C.JX_EZERODIV: ZeroDivisionError,      # blastfilter.py:235
    C.JX_EUNSUPPORTED: NotImplementedError,
    C.JX_EDUPANCHOR: AssertionError,       # synteny.py:397
    C.JX_EINTERNAL: RuntimeError,
}


class DeviceText(object):
    """Text held on the device by Engine.upload (jx_text_upload)."""

    def __init__(self, ptr, nbytes, keep):
        self.ptr, self.nbytes, self._keep = ptr, nbytes, keep


def _text_ptr(text):
    """-> (pointer, nbytes, on_device, keepalive)."""
    if isinstance(text, DeviceText):
        return text.ptr, text.nbytes, 1, text
    if hasattr(text, "data_ptr") and hasattr(text, "is_cuda"):          # torch tensor
        t = text.contiguous()
        return t.data_ptr(), t.numel() * t.element_size(), int(t.is_cuda), t
    if hasattr(text, "__cuda_array_interface__"):
        cai = text.__cuda_array_interface__
        n = int(np.prod(cai["shape"])) * np.dtype(cai["typestr"]).itemsize
        return cai["data"][0], n, 1, text
    if isinstance(text, (bytes, bytearray, memoryview)):
        a = np.frombuffer(text, dtype=np.uint8)
        return a.ctypes.data, a.size, 0, a
    a = np.ascontiguousarray(text).view(np.uint8).reshape(-1)
    return a.ctypes.data, a.size, 0, a


class FilterResult(object):
    def __init__(self, r, lib, Gq, Gs, want_mothers, engine=None):
        self._engine = engine          # the text may be borrowed from the context's pinned pool: keep the context alive
        n = r.n
        self.n = n
        self.q_rank = C.as_np(r.q_rank, n, np.int32) if r.q_rank else None
        self.s_rank = C.as_np(r.s_rank, n, np.int32) if r.s_rank else None
        self.score = C.as_np(r.score, n, np.float32) if r.score else None
        self.line_off = C.as_np(r.line_off, n, np.uint64) if r.line_off else None
        # zero-copy view of the `.filtered` bytes; they belong to this result (jx_filter_result_free returns a pinned
        # buffer to the context's pool) and stay valid however many calls follow on the same engine
        self._raw, self._lib = r, lib
        self.text = (np.ctypeslib.as_array(ctypes.cast(r.text, C.c_u8p), shape=(max(r.text_len, 1),))[: r.text_len]
                     if r.text else None)
        self.q_mother = C.as_np(r.q_mother, Gq, np.int32) if want_mothers else None
        self.s_mother = C.as_np(r.s_mother, Gs, np.int32) if want_mothers else None
        self.stats = dict(zip(("lines", "mapped", "unmapped", "after_exclude", "after_cscore", "after_dedupe",
                               "after_tandem", "comments"), list(r.stats)))


    def __del__(self):
        raw = getattr(self, "_raw", None)
        if raw is not None:
            self.text = None
            self._lib.jx_filter_result_free(ctypes.byref(raw))
            self._raw = None


class ScanResult(object):
    def __init__(self, r, with_index=False):
        n, nb = r.n_anchors, r.n_blocks
        self.n_anchors, self.n_blocks = n, nb
        self.q_rank = C.as_np(r.q_rank, n, np.int32)
        self.s_rank = C.as_np(r.s_rank, n, np.int32)
        self.score = C.as_np(r.score, n, np.float32)
        self.block_start = C.as_np(r.block_start, nb + 1, np.int64)
        self.index = C.as_np(r.index, n, np.int64) if with_index else None
        self.stats = dict(zip(("lines", "hits", "clustered", "comments"), list(r.stats)[:4]))


class LiftoverResult(object):
    """Views (no copies) of the library's result arrays -- 1.7 M lifted pairs at C3 size are 27 MB that need not
    be copied and page-faulted again; the result is freed when the last array is unreachable."""

    def __init__(self, r, n_lines, lib):
        own = C._Owner(lib.jx_liftover_result_free, r)
        n = r.n_lifted
        self.n_lifted = n
        self.q_rank = C.view_np(r.q_rank, n, np.int32, own)
        self.s_rank = C.view_np(r.s_rank, n, np.int32, own)
        self.score = C.view_np(r.score, n, np.float32, own)
        self.block = C.view_np(r.block, n, np.int32, own)
        self.accepted = C.view_np(r.accepted, n_lines, np.uint8, own)
        self.raw_score = C.view_np(r.raw_score, n_lines, np.float32, own)
        self.stats = dict(zip(("lines", "hits", "anchors", "lifted", "tree_walked", "comments"), list(r.stats)[:6]))


class SynfindResult(object):
    """Rows of jx_synfind_*: [0, n_forward) the forward pass, the rest the mirror pass (ranks of the other BED)."""

    def __init__(self, r):
        n = r.n_forward + r.n_mirror
        self.n_forward, self.n_mirror, self.n_points = r.n_forward, r.n_mirror, r.n_points
        self.query = C.as_np(r.query, n, np.int32)
        self.syntelog = C.as_np(r.syntelog, n, np.int32)
        self.left = C.as_np(r.left, n, np.int32)
        self.right = C.as_np(r.right, n, np.int32)
        self.score = C.as_np(r.score, n, np.int32)
        self.gray = C.as_np(r.gray, n, np.uint8)
        self.minus = C.as_np(r.minus, n, np.uint8)
        self.stats = dict(zip(("lines", "hits", "comments"), list(r.stats)[:3]))


class Engine(object):
    """One per GPU.  Not thread-safe (like a jx_ctx)."""

    def __init__(self, device=0):
        self.lib = C.lib()
        ctx = ctypes.c_void_p()
        rc = self.lib.jx_ctx_create(int(device), ctypes.byref(ctx))
        if rc != 0:
            raise RuntimeError("jcvi_b200: no usable CUDA device %d (jx_ctx_create -> %d); there is no CPU fallback"
                               % (device, rc))
        self.ctx = ctx
        self.device = device
        self._pair = None
        self.Gq = self.Gs = 0
        self.n_name_ids = 0

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.jx_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.jx_last_error(self.ctx).decode()
            raise _ERR.get(rc, RuntimeError)(msg)

    # ---- BED pair -----------------------------------------------------------------------------
    def load_pair(self, qbed, sbed, is_self):
        key = (id(qbed), id(sbed), bool(is_self))
        if self._pair == key:
            return
        tq = qbed.tables()
        ts = tq if (is_self or sbed is qbed) else sbed.tables()
        ids = {}
        for a in tq["accns"]:
            ids.setdefault(a, len(ids))
        for a in ts["accns"]:
            ids.setdefault(a, len(ids))
        nid_q = np.fromiter((ids[a] for a in tq["accns"]), dtype=np.uint32, count=tq["G"])
        nid_s = np.fromiter((ids[a] for a in ts["accns"]), dtype=np.uint32, count=ts["G"])
        slast = ts["last"]
        same = np.fromiter((slast.get(a, -1) for a in tq["accns"]), dtype=np.int32, count=tq["G"])
        keep = []
        for side, t, nid in ((0, tq, nid_q), (1, ts, nid_s)):
            b = C.JxBed()
            b.n_genes = t["G"]
            arrs = [t["names"], t["name_off"], t["indexable"], t["chr_lex"], t["chr_begin"], t["chr_end"], t["length"],
                    t["accn_lexrank"], nid]
            keep += arrs
            (b.names, b.name_off, b.indexable, b.chr_lex, b.chr_begin, b.chr_end, b.length, b.accn_lexrank,
             b.name_id) = [C.ptr(a) for a in arrs]
            b.n_chr = t["n_chr"]
            self._check(self.lib.jx_bed_load(self.ctx, side, ctypes.byref(b)))
        self._check(self.lib.jx_pair_setup(self.ctx, int(bool(is_self)), C.ptr(same), len(ids)))
        self._pair = key
        self._beds = (qbed, sbed)          # keep alive: the cache key uses id()
        self.Gq, self.Gs = tq["G"], ts["G"]
        self.n_name_ids = len(ids)

    def load_names(self, names):
        """No BED (formats.blast cscore): `names` (sorted, distinct byte strings) become the name table
        of BOTH sides; rank == name id == position in the list."""
        G = len(names)
        blob = np.frombuffer(b"".join(names) or b"\0", dtype=np.uint8)
        off = np.zeros(G + 1, dtype=np.uint32)
        np.cumsum([len(x) for x in names], out=off[1:])
        ones = np.ones(max(G, 1), dtype=np.uint8)
        zeros = np.zeros(max(G, 1), dtype=np.int32)
        ends = np.full(max(G, 1), G, dtype=np.int32)
        ids = np.arange(max(G, 1), dtype=np.uint32)
        b = C.JxBed()
        b.n_genes = G
        keep = [blob, off, ones, zeros, ends, ids]
        (b.names, b.name_off, b.indexable, b.chr_lex, b.chr_begin, b.chr_end, b.length, b.accn_lexrank, b.name_id) = [
            C.ptr(a) for a in (blob, off, ones, zeros, zeros, ends, zeros, ids, ids)]
        b.n_chr = 1
        for side in (0, 1):
            self._check(self.lib.jx_bed_load(self.ctx, side, ctypes.byref(b)))
        same = np.full(max(G, 1), -1, dtype=np.int32)          # no `query == subject` rule in cscore
        self._check(self.lib.jx_pair_setup(self.ctx, 0, C.ptr(same), G))
        self._pair = None
        self._beds = None
        self.Gq = self.Gs = G
        del keep

    def distinct_names(self, text, strip_names=True, seed=0):
        """-> (list of representative raw lines (bytes, no terminator), uint8 side per line, stats dict)"""
        self._held_key = None
        p, n, dev, keep = _text_ptr(text)
        buf = ctypes.c_void_p(); ln = ctypes.c_int64(); side = ctypes.c_void_p(); cnt = ctypes.c_int64()
        st = (ctypes.c_int64 * 4)()
        self._check(self.lib.jx_distinct_names(self.ctx, p, n, dev, int(bool(strip_names)), int(seed), ctypes.byref(buf),
                                               ctypes.byref(ln), ctypes.byref(side), ctypes.byref(cnt), st))
        try:
            raw = ctypes.string_at(buf, ln.value) if ln.value else b""
            sides = np.frombuffer(ctypes.string_at(side, cnt.value), dtype=np.uint8) if cnt.value else np.zeros(0, np.uint8)
        finally:
            self.lib.jx_free(buf); self.lib.jx_free(side)
        lines = raw.split(b"\n")[:cnt.value]
        return lines, sides, dict(zip(("lines", "named", "short", "comments"), list(st)))

    def blast_cscore(self, text, cutoff=0.9999, strip_names=True, pct=False, writeblast=False):
        """`jcvi.formats.blast cscore` in one call -> (cscore file bytes, --writeblast file bytes or None, stats)."""
        self._held_key = None
        p, n, dev, keep = _text_ptr(text)
        out = ctypes.c_void_p(); ln = ctypes.c_int64(); bout = ctypes.c_void_p(); bln = ctypes.c_int64()
        st = (ctypes.c_int64 * 4)()
        rc = self.lib.jx_blast_cscore(self.ctx, p, n, dev, int(bool(strip_names)), float(cutoff), int(bool(pct)),
                                      int(bool(writeblast)), ctypes.byref(out), ctypes.byref(ln), ctypes.byref(bout),
                                      ctypes.byref(bln), st)
        self._pair = None                      # the name tables now hold the table's own names
        self._beds = None
        try:
            self._check(rc)
            a = ctypes.string_at(out, ln.value) if ln.value else b""
            b = (ctypes.string_at(bout, bln.value) if bln.value else b"") if writeblast else None
        finally:
            self.lib.jx_free(out); self.lib.jx_free(bout)
        return a, b, dict(zip(("lines", "named", "short", "comments"), list(st)))

    def blast_filter(self, text, score=0, pctid=95.0, hitlen=100, evalue=0.01, noself=False, inverse=False, ids=None,
                     sink=None):
        """`jcvi.formats.blast filter` -> (bytes of the kept rows, stats).  ids: iterable of names (bytes) or None.
        sink: a binary file object; the rows are then written to it straight from the library's buffer (no Python
        copy of a multi-GB result) and b"" is returned in their place."""
        self._held_key = None
        use_ids = bool(ids)
        if use_ids:
            self.load_names(sorted(set(ids)))
        p, n, dev, keep = _text_ptr(text)
        out = ctypes.c_void_p(); ln = ctypes.c_int64(); st = (ctypes.c_int64 * 4)()
        rc = self.lib.jx_blast_filter(self.ctx, p, n, dev, float(score), float(pctid), int(hitlen), float(evalue),
                                      int(bool(noself)), int(bool(inverse)), int(use_ids), ctypes.byref(out), ctypes.byref(ln), st)
        try:
            self._check(rc)
            if sink is not None:
                if ln.value:
                    sink.write(memoryview((ctypes.c_char * ln.value).from_address(out.value)))
                a = b""
            else:
                a = ctypes.string_at(out, ln.value) if ln.value else b""
        finally:
            self.lib.jx_free(out)
        return a, dict(zip(("rows", "kept", "removed", "comments"), list(st)))

    def best_table(self, n):
        out = np.zeros(max(int(n), 1), dtype=np.float32)
        self._check(self.lib.jx_best_table_copy(self.ctx, C.ptr(out), int(n)))
        return out[:n]

    def session_counts(self):
        st = (ctypes.c_int64 * 3)()
        self._check(self.lib.jx_session_counts(self.ctx, st))
        return dict(zip(("lines", "mapped", "unmapped"), list(st)))

    def host_blastline(self, line):
        """BlastLine(line) the way cblast does it (glibc): -> (nconv, query, subject, pctid, score, str(b))"""
        q = ctypes.create_string_buffer(128); s = ctypes.create_string_buffer(128); out = ctypes.create_string_buffer(1024)
        pct = ctypes.c_float(); sc = ctypes.c_float()
        n = self.lib.jx_host_blastline(line, q, s, ctypes.byref(pct), ctypes.byref(sc), out, 1024)
        return n, q.value, s.value, pct.value, sc.value, out.value

    def upload(self, text):
        """Start the asynchronous upload of a host text; the returned DeviceText can be handed to
        blastfilter / scan_text / liftover_text, which overlap tokenizing with the copy."""
        p, n, dev, keep = _text_ptr(text)
        if dev:
            raise ValueError("text is already on the device")
        out = ctypes.c_uint64()
        self._held_key = None
        self._check(self.lib.jx_text_upload(self.ctx, p, n, ctypes.byref(out)))
        return DeviceText(out.value, n, keep)

    def upload_file(self, path):
        """jx_file_upload: the table of `path` resident on the device -- plain files through reader threads and pinned
        pieces, `.gz` inflated by the library (block-parallel for BGZF); bzip2 falls back to Python's reader."""
        import os
        out = ctypes.c_uint64()
        n = ctypes.c_int64()
        kind = ctypes.c_int32()
        self._held_key = None
        rc = self.lib.jx_file_upload(self.ctx, os.fsencode(path), ctypes.byref(out), ctypes.byref(n), ctypes.byref(kind))
        if rc == C.JX_EUNSUPPORTED:
            from .apps_base import read_text_file
            return self.upload(read_text_file(path))
        if rc == C.JX_EARG and not os.path.exists(path):
            raise FileNotFoundError(path)
        self._check(rc)
        dev = DeviceText(out.value, n.value, None)
        dev.kind = ("plain", "gzip", "bgzf")[kind.value]
        return dev

    def release_text(self):
        self._held_key = None
        self._check(self.lib.jx_text_release(self.ctx))

    def remember_text(self, key, dev):
        """Name the text Engine.upload returned last (e.g. apps_base.file_key of the file it came from)."""
        self._held_key = (key, dev)

    def recall_text(self, key):
        """The resident DeviceText remembered under `key`, or None.  Every call that may replace the library's held
        text (upload, release_text, cscore_begin, blast_cscore, blast_filter, distinct_names, near_lines) forgets it."""
        hk = getattr(self, "_held_key", None)
        return hk[1] if hk is not None and hk[0] == key else None

    def set_record_reuse(self, on=True):
        """liftover re-uses the records blastfilter tokenized from the same resident text (default);
        off: every call tokenizes, like the reference's two passes over the raw table."""
        self._check(self.lib.jx_set_record_reuse(self.ctx, int(bool(on))))

    # ---- stages ---------------------------------------------------------------------------------
    def blastfilter(self, text, strip_names=True, cscore=0.7, tandem_Nmax=10, exclude_keys=None, want_text=True,
                    want_mothers=False, raw=False, skip_arrays=False):
        p, n, dev, keep = _text_ptr(text)
        o = C.JxFilterOpts()
        o.strip_names = int(bool(strip_names))
        o.cscore = float(cscore or 0.0)
        o.tandem_nmax = int(tandem_Nmax or 0)
        ek = None
        if exclude_keys is not None and len(exclude_keys):
            ek = np.unique(np.asarray(exclude_keys, dtype=np.uint64))
            o.exclude_keys = C.ptr(ek)
            o.n_exclude = len(ek)
        o.want_text = int(bool(want_text))
        o.want_mothers = int(bool(want_mothers))
        o.skip_arrays = int(bool(skip_arrays))
        r = C.JxFilterResult()
        rc = self.lib.jx_blastfilter(self.ctx, p, n, dev, ctypes.byref(o), ctypes.byref(r))
        if rc != 0:
            self.lib.jx_filter_result_free(ctypes.byref(r))
            self._check(rc)
        if raw:
            return r          # caller frees with free_filter()
        return FilterResult(r, self.lib, self.Gq, self.Gs, want_mothers, engine=self)   # owns and frees `r`

    def pipeline(self, text, strip_names=True, cscore=0.7, tandem_Nmax=10, exclude_keys=None, want_text=True,
                 skip_arrays=True, xdist=20, ydist=20, N=4, intrabound=300, liftover_dist=10):
        """blastfilter -> scan -> liftover in ONE library call (jx_pipeline): the survivors and the anchors stay on the
        device between the stages.  -> (FilterResult, ScanResult, LiftoverResult), identical to the three separate calls."""
        p, n, dev, keep = _text_ptr(text)
        o = C.JxFilterOpts()
        o.strip_names = int(bool(strip_names))
        o.cscore = float(cscore or 0.0)
        o.tandem_nmax = int(tandem_Nmax or 0)
        ek = None
        if exclude_keys is not None and len(exclude_keys):
            ek = np.unique(np.asarray(exclude_keys, dtype=np.uint64))
            o.exclude_keys = C.ptr(ek)
            o.n_exclude = len(ek)
        o.want_text = int(bool(want_text))
        o.skip_arrays = int(bool(skip_arrays))
        so = self._scan_opts(xdist, ydist, N, intrabound)
        fr = C.JxFilterResult(); sr = C.JxScanResult(); lr = C.JxLiftoverResult()
        if not dev:
            self._held_key = None                      # the library uploads (and holds) the text
        rc = self.lib.jx_pipeline(self.ctx, p, n, dev, ctypes.byref(o), ctypes.byref(so), int(liftover_dist), ctypes.byref(fr),
                                  ctypes.byref(sr), ctypes.byref(lr))
        if rc != 0:
            self.lib.jx_filter_result_free(ctypes.byref(fr))
            self.lib.jx_scan_result_free(ctypes.byref(sr))
            self.lib.jx_liftover_result_free(ctypes.byref(lr))
            self._check(rc)
        f = FilterResult(fr, self.lib, self.Gq, self.Gs, False, engine=self)
        try:
            s = ScanResult(sr)
        finally:
            self.lib.jx_scan_result_free(ctypes.byref(sr))
        return f, s, LiftoverResult(lr, s.n_anchors, self.lib)

    def free_filter(self, r):
        self.lib.jx_filter_result_free(ctypes.byref(r))

    def _scan_opts(self, xdist, ydist, N, intrabound):
        o = C.JxScanOpts()
        o.xdist, o.ydist, o.min_size, o.intrabound = int(xdist), int(ydist), int(N), int(intrabound)
        return o

    def scan_text(self, text, strip_names=False, xdist=20, ydist=20, N=4, intrabound=300):
        p, n, dev, keep = _text_ptr(text)
        o = self._scan_opts(xdist, ydist, N, intrabound)
        r = C.JxScanResult()
        rc = self.lib.jx_scan_text(self.ctx, p, n, dev, int(bool(strip_names)), ctypes.byref(o), ctypes.byref(r))
        try:
            self._check(rc)
            return ScanResult(r)
        finally:
            self.lib.jx_scan_result_free(ctypes.byref(r))

    def scan_filtered(self, raw_filter_result, xdist=20, ydist=20, N=4, intrabound=300):
        o = self._scan_opts(xdist, ydist, N, intrabound)
        r = C.JxScanResult()
        rc = self.lib.jx_scan_filtered(self.ctx, ctypes.byref(raw_filter_result), ctypes.byref(o), ctypes.byref(r))
        try:
            self._check(rc)
            return ScanResult(r)
        finally:
            self.lib.jx_scan_result_free(ctypes.byref(r))

    def scan_points(self, qi, si, score, group=None, xdist=20, ydist=20, N=4, is_self=False, intrabound=300):
        qi = np.ascontiguousarray(qi, dtype=np.int32)
        si = np.ascontiguousarray(si, dtype=np.int32)
        score = np.ascontiguousarray(score, dtype=np.float32)
        g = None if group is None else np.ascontiguousarray(group, dtype=np.int64)
        o = self._scan_opts(xdist, ydist, N, intrabound)
        r = C.JxScanResult()
        rc = self.lib.jx_scan_points(self.ctx, C.ptr(qi), C.ptr(si), C.ptr(score), C.ptr(g), len(qi), int(bool(is_self)),
                                     ctypes.byref(o), ctypes.byref(r))
        try:
            self._check(rc)
            return ScanResult(r, with_index=True)
        finally:
            self.lib.jx_scan_result_free(ctypes.byref(r))

    def liftover_text(self, text, a_qi, a_si, a_block, strip_names=True, dist=10):
        p, n, dev, keep = _text_ptr(text)
        a_qi = np.ascontiguousarray(a_qi, dtype=np.int32)
        a_si = np.ascontiguousarray(a_si, dtype=np.int32)
        a_block = np.ascontiguousarray(a_block, dtype=np.int32)
        r = C.JxLiftoverResult()
        rc = self.lib.jx_liftover_text(self.ctx, p, n, dev, int(bool(strip_names)), int(dist), C.ptr(a_qi), C.ptr(a_si),
                                       C.ptr(a_block), len(a_qi), ctypes.byref(r))
        if rc != 0:
            self.lib.jx_liftover_result_free(ctypes.byref(r))
            self._check(rc)
        return LiftoverResult(r, len(a_qi), self.lib)          # owns `r`

    def nearest_anchor(self, points, anchors, dist):
        pts = np.ascontiguousarray(points, dtype=np.int64).reshape(-1, 2)
        anc = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1, 2)
        near = np.zeros(len(pts), dtype=np.int64)
        dd = np.zeros(len(pts), dtype=np.int64)
        self._check(self.lib.jx_nearest_anchor(self.ctx, C.ptr(pts), len(pts), C.ptr(anc), len(anc), int(dist),
                                               C.ptr(near), C.ptr(dd)))
        return near, dd

    # ---- one genome pair sharded over ranks (see jcvi_b200/sharded.py) --------------------------------
    def cscore_begin(self, text, strip_names=True, exclude_keys=None):
        """Tokenize this rank's slice, accumulate the local per-name best scores.  Returns
        (device pointer of the table, number of entries): int32-orderable float bits."""
        self._held_key = None
        p, n, dev, keep = _text_ptr(text)
        ek = None
        if exclude_keys is not None and len(exclude_keys):
            ek = np.unique(np.asarray(exclude_keys, dtype=np.uint64))
        ptr = ctypes.c_uint64(); nids = ctypes.c_uint32()
        self._check(self.lib.jx_cscore_begin(self.ctx, p, n, dev, int(bool(strip_names)), C.ptr(ek), 0 if ek is None else len(ek),
                                             ctypes.byref(ptr), ctypes.byref(nids)))
        return ptr.value, nids.value

    def best_table_tensor(self, ptr, n):
        """torch int32 view of the device table (positive float bit patterns order like ints)."""
        import torch

        class _Dev(object):
            pass
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(d, device="cuda:%d" % self.device)

    def _take_lines(self, buf, ln):
        try:
            return np.ctypeslib.as_array(ctypes.cast(buf, C.c_u8p), shape=(max(ln.value, 1),))[: ln.value].copy()
        finally:
            self.lib.jx_free(buf)

    def _device_lines(self, ptr, nbytes):
        """torch uint8 view of a context-owned device result (valid until the next *_device call)."""
        import torch

        class _Dev(object):
            pass
        if not nbytes:
            return torch.empty(0, dtype=torch.uint8, device="cuda:%d" % self.device)
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(d, device="cuda:%d" % self.device)

    def cscore_finish(self, cscore, device=False):
        """-> (uint8 array of the slice's raw lines that pass the C-score filter, file order; count).
        device=True: a torch CUDA tensor viewing the library's device buffer instead (no host copy)."""
        ln = ctypes.c_int64(); cnt = ctypes.c_int64()
        if device:
            dptr = ctypes.c_uint64()
            self._check(self.lib.jx_cscore_finish_device(self.ctx, float(cscore or 0.0), ctypes.byref(dptr), ctypes.byref(ln), ctypes.byref(cnt)))
            return self._device_lines(dptr.value, ln.value), cnt.value
        buf = ctypes.c_void_p()
        self._check(self.lib.jx_cscore_finish(self.ctx, float(cscore or 0.0), ctypes.byref(buf), ctypes.byref(ln), ctypes.byref(cnt)))
        return self._take_lines(buf, ln), cnt.value

    def near_lines(self, text, a_qi, a_si, strip_names=True, dist=10, device=False):
        """-> (uint8 array of the raw lines whose hit lies within `dist` of an anchor; count).
        device=True: a torch CUDA tensor viewing the library's device buffer instead."""
        self._held_key = None
        p, n, dev, keep = _text_ptr(text)
        a_qi = np.ascontiguousarray(a_qi, dtype=np.int32)
        a_si = np.ascontiguousarray(a_si, dtype=np.int32)
        ln = ctypes.c_int64(); cnt = ctypes.c_int64()
        if device:
            dptr = ctypes.c_uint64()
            self._check(self.lib.jx_near_lines_device(self.ctx, p, n, dev, int(bool(strip_names)), int(dist), C.ptr(a_qi), C.ptr(a_si),
                                                      len(a_qi), ctypes.byref(dptr), ctypes.byref(ln), ctypes.byref(cnt)))
            return self._device_lines(dptr.value, ln.value), cnt.value
        buf = ctypes.c_void_p()
        self._check(self.lib.jx_near_lines(self.ctx, p, n, dev, int(bool(strip_names)), int(dist), C.ptr(a_qi), C.ptr(a_si),
                                           len(a_qi), ctypes.byref(buf), ctypes.byref(ln), ctypes.byref(cnt)))
        return self._take_lines(buf, ln), cnt.value

    # ---- one genome pair sharded by chromosome pair (jcvi_b200/sharded.py drives these) -----------------------------
    def _dev_records(self, ptr, n):
        """torch int32 [n, 4] view of n 16-byte device records of the context's session (valid until shard_reset)."""
        import torch

        class _Dev(object):
            pass
        if not n:
            return torch.empty((0, 4), dtype=torch.int32, device="cuda:%d" % self.device)
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (int(n), 4), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(d, device="cuda:%d" % self.device)

    def shard_reset(self):
        self._check(self.lib.jx_shard_reset(self.ctx))

    def shard_slots(self):
        n = ctypes.c_int64()
        self._check(self.lib.jx_shard_slots(self.ctx, ctypes.byref(n)))
        return n.value

    def shard_candidates(self, cscore, ord_base, n_cp):
        """-> (int32 [n, 4] device candidates, uint32 histogram per chromosome pair)"""
        hist = np.zeros(max(int(n_cp), 1), dtype=np.uint32)
        ptr = ctypes.c_uint64(); n = ctypes.c_int64()
        self._check(self.lib.jx_shard_candidates(self.ctx, float(cscore or 0.0), int(ord_base), ctypes.byref(ptr), ctypes.byref(n),
                                                 C.ptr(hist), int(n_cp)))
        return self._dev_records(ptr.value, n.value), hist[:n_cp]

    def shard_partition(self, cand, owner, world, want_hist=False, hist_in=None):
        """candidates grouped by destination rank -> (int32 [n, 4] device array, int64 counts[world][, uint32 hist per pair]);
        hist_in: the candidates' histogram when the caller already has it (shard_candidates) -- saves a pass and a round trip"""
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        counts = np.zeros(int(world), dtype=np.int64)
        hist = np.zeros(len(owner), dtype=np.uint32) if want_hist else None
        hin = None if hist_in is None else np.ascontiguousarray(hist_in, dtype=np.uint32)
        out = ctypes.c_uint64()
        self._check(self.lib.jx_shard_partition(self.ctx, int(cand.data_ptr()) if cand.numel() else 0, int(cand.shape[0]), C.ptr(owner),
                                                len(owner), int(world), ctypes.byref(out), C.ptr(counts), C.ptr(hin), C.ptr(hist)))
        part = self._dev_records(out.value, cand.shape[0])
        return (part, counts, hist) if want_hist else (part, counts)

    def _dev_pairs(self, ptr, n):
        import torch

        class _Dev(object):
            pass
        if not n:
            return torch.empty((0, 2), dtype=torch.int32, device="cuda:%d" % self.device)
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (int(n), 2), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(d, device="cuda:%d" % self.device)

    def shard_pairs(self, recv, tandem_Nmax):
        """-> (query-side edges, subject-side edges): int32 [n, 2] device arrays of gene ranks"""
        q = ctypes.c_uint64(); nq = ctypes.c_int64(); s = ctypes.c_uint64(); ns = ctypes.c_int64()
        self._keep_recv = recv
        self._check(self.lib.jx_shard_pairs(self.ctx, int(recv.data_ptr()) if recv.numel() else 0, int(recv.shape[0]), int(tandem_Nmax or 0),
                                            ctypes.byref(q), ctypes.byref(nq), ctypes.byref(s), ctypes.byref(ns)))
        return self._dev_pairs(q.value, nq.value), self._dev_pairs(s.value, ns.value)

    def shard_survivors(self, qedges, sedges, tandem_Nmax, want_mothers=False):
        """all ranks' edges -> (this rank's survivors: int32 [n, 4] device candidates, q_mother, s_mother)"""
        out = ctypes.c_uint64(); n = ctypes.c_int64()
        mq = np.zeros(max(self.Gq, 1), dtype=np.int32) if want_mothers else None
        ms = np.zeros(max(self.Gs, 1), dtype=np.int32) if want_mothers else None
        self._check(self.lib.jx_shard_survivors(self.ctx, int(qedges.data_ptr()) if qedges.numel() else 0, int(qedges.shape[0]),
                                                int(sedges.data_ptr()) if sedges.numel() else 0, int(sedges.shape[0]), int(tandem_Nmax or 0),
                                                ctypes.byref(out), ctypes.byref(n), C.ptr(mq), C.ptr(ms)))
        return self._dev_records(out.value, n.value), mq, ms

    def shard_order(self, surv):
        r = C.JxFilterResult()
        rc = self.lib.jx_shard_order(self.ctx, int(surv.data_ptr()) if surv.numel() else 0, int(surv.shape[0]), ctypes.byref(r))
        if rc != 0:
            self.lib.jx_filter_result_free(ctypes.byref(r))
            self._check(rc)
        return FilterResult(r, self.lib, self.Gq, self.Gs, False, engine=self)

    def scan_cands(self, surv, xdist=20, ydist=20, N=4, intrabound=300, device=False):
        """scan of device survivors.  device=True: -> int32 [m, 4] CUDA tensor of {qi, si, score bits, block} (the anchors
        never visit the host); else a ScanResult"""
        o = self._scan_opts(xdist, ydist, N, intrabound)
        r = C.JxScanResult()
        rows = ctypes.c_uint64()
        rc = self.lib.jx_scan_cands(self.ctx, int(surv.data_ptr()) if surv.numel() else 0, int(surv.shape[0]), ctypes.byref(o), ctypes.byref(r),
                                    ctypes.byref(rows) if device else None, 0 if device else 1)
        try:
            self._check(rc)
            if device:
                return self._dev_records(rows.value, r.n_anchors)
            return ScanResult(r)
        finally:
            self.lib.jx_scan_result_free(ctypes.byref(r))

    @staticmethod
    def _anchor_ptr(a):
        """-> (pointer, on_device, keepalive) of an int32 anchor column (numpy array or CUDA tensor)"""
        if hasattr(a, "is_cuda"):
            t = a.contiguous()
            return ctypes.c_void_p(t.data_ptr() if t.numel() else 0), 1, t
        h = np.ascontiguousarray(a, dtype=np.int32)
        return C.ptr(h), 0, h

    def shard_near(self, a_qi, a_si, ord_base, strip_names=True, dist=10, text=None):
        """this rank's raw records within dist of an anchor -> int32 [n, 4] device candidates (anchors: host arrays or
        int32 CUDA tensors)"""
        pq, dq, kq = self._anchor_ptr(a_qi); ps, ds, ks = self._anchor_ptr(a_si)
        assert dq == ds
        p, nb, dev, keep = _text_ptr(text) if text is not None else (None, 0, 1, None)
        ptr = ctypes.c_uint64(); n = ctypes.c_int64()
        self._check(self.lib.jx_shard_near(self.ctx, p, nb, dev, int(bool(strip_names)), int(dist), pq, ps, len(kq), dq,
                                           int(ord_base), ctypes.byref(ptr), ctypes.byref(n)))
        return self._dev_records(ptr.value, n.value)

    def shard_lift(self, recv, n_slots_total, a_qi, a_si, a_block, dist=10, device=False):
        """device=True: -> (int32 [4, n_lifted] CUDA tensor of the columns q / s / score bits / block, int64 [n_anchor_lines]
        CUDA tensor accepted << 32 | raw score bits); else a LiftoverResult"""
        import torch
        pq, dq, kq = self._anchor_ptr(a_qi); ps, ds, ks = self._anchor_ptr(a_si); pb, db, kb = self._anchor_ptr(a_block)
        assert dq == ds == db
        r = C.JxLiftoverResult()
        dv = (ctypes.c_uint64 * 2)()
        rc = self.lib.jx_shard_lift(self.ctx, int(recv.data_ptr()) if recv.numel() else 0, int(recv.shape[0]), int(n_slots_total), int(dist),
                                    pq, ps, pb, len(kq), dq, ctypes.byref(r), dv if device else None)
        if rc != 0:
            self.lib.jx_liftover_result_free(ctypes.byref(r))
            self._check(rc)
        if not device:
            return LiftoverResult(r, len(kq), self.lib)
        nl, nal = int(r.n_lifted), len(kq)
        self.lib.jx_liftover_result_free(ctypes.byref(r))
        dname = "cuda:%d" % self.device

        class _Dev(object):
            pass
        if nl and dv[0]:
            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (4, nl), "typestr": "<i4", "data": (int(dv[0]), False), "version": 2}
            cols = torch.as_tensor(d, device=dname)
        else:
            cols = torch.zeros((4, 0), dtype=torch.int32, device=dname)
        if nal and dv[1]:
            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (nal,), "typestr": "<i8", "data": (int(dv[1]), False), "version": 2}
            acc = torch.as_tensor(d, device=dname)
        else:
            acc = torch.zeros((nal,), dtype=torch.int64, device=dname)
        return cols, acc

    def mcscan(self, start, end, score, block_id, max_iter, pair_gene, pair_partner, pair_block, n_genes, n_blocks):
        """jx_mcscan -> (tracks: list of block-id lists in chain order, chain scores, table[n_genes, n_tracks] of partner ids, -1 = '.')."""
        start = np.ascontiguousarray(start, dtype=np.float64); end = np.ascontiguousarray(end, dtype=np.float64)
        score = np.ascontiguousarray(score, dtype=np.int64); block_id = np.ascontiguousarray(block_id, dtype=np.int32)
        pg = np.ascontiguousarray(pair_gene, dtype=np.int32); pp = np.ascontiguousarray(pair_partner, dtype=np.int32)
        pb = np.ascontiguousarray(pair_block, dtype=np.int32)
        nt = ctypes.c_int32()
        tl = ctypes.c_void_p(); tb = ctypes.c_void_p(); ts = ctypes.c_void_p(); tab = ctypes.c_void_p()
        rc = self.lib.jx_mcscan(self.ctx, C.ptr(start), C.ptr(end), C.ptr(score), C.ptr(block_id), len(start), int(max_iter),
                                C.ptr(pg), C.ptr(pp), C.ptr(pb), len(pg), int(n_genes), int(n_blocks), ctypes.byref(nt),
                                ctypes.byref(tl), ctypes.byref(tb), ctypes.byref(ts), ctypes.byref(tab))
        try:
            self._check(rc)
            T = nt.value
            lens = C.as_np(ctypes.cast(tl, C.c_i32p), T, np.int32)
            blocks = C.as_np(ctypes.cast(tb, C.c_i32p), int(lens.sum()), np.int32)
            scores = C.as_np(ctypes.cast(ts, C.c_i64p), T, np.int64)
            table = C.as_np(ctypes.cast(tab, C.c_i32p), int(n_genes) * T, np.int32).reshape(int(n_genes), T)
        finally:
            for p in (tl, tb, ts, tab):
                self.lib.jx_free(p)
        tracks = np.split(blocks, np.cumsum(lens)[:-1]) if T else []
        return [t.tolist() for t in tracks], scores.tolist(), table

    def block_stats(self, block, qrank, srank, n_blocks):
        """jx_block_stats -> dict of per-block arrays: qmin qmax smin smax size distinct_q distinct_s sum_q sum_s sum_qs."""
        b = np.ascontiguousarray(block, dtype=np.int32); x = np.ascontiguousarray(qrank, dtype=np.int32)
        y = np.ascontiguousarray(srank, dtype=np.int32)
        nb = int(n_blocks)
        out = dict(qmin=np.zeros(nb, np.int32), qmax=np.zeros(nb, np.int32), smin=np.zeros(nb, np.int32), smax=np.zeros(nb, np.int32),
                   size=np.zeros(nb, np.uint32), distinct_q=np.zeros(nb, np.uint32), distinct_s=np.zeros(nb, np.uint32),
                   sum_q=np.zeros(nb, np.uint64), sum_s=np.zeros(nb, np.uint64), sum_qs=np.zeros(nb, np.uint64))
        self._check(self.lib.jx_block_stats(self.ctx, C.ptr(b), C.ptr(x), C.ptr(y), len(b), nb,
                                            *[C.ptr(out[k]) for k in ("qmin", "qmax", "smin", "smax", "size", "distinct_q",
                                                                      "distinct_s", "sum_q", "sum_s", "sum_qs")]))
        return out

    def synfind_text(self, text, strip_names=True, window=40, cutoff=0.1):
        """jx_synfind_text (compara/synfind.py:130-133 derive the two integers)."""
        p, n, dev, keep = _text_ptr(text)
        r = C.JxSynfindResult()
        rc = self.lib.jx_synfind_text(self.ctx, p, n, dev, int(bool(strip_names)), int(window) // 2, int(cutoff * window),
                                      ctypes.byref(r))
        try:
            self._check(rc)
            return SynfindResult(r)
        finally:
            self.lib.jx_synfind_result_free(ctypes.byref(r))

    def synfind_points(self, qi, si, window_half, cutoff, passes=1):
        qi = np.ascontiguousarray(qi, dtype=np.int32)
        si = np.ascontiguousarray(si, dtype=np.int32)
        r = C.JxSynfindResult()
        rc = self.lib.jx_synfind_points(self.ctx, C.ptr(qi), C.ptr(si), len(qi), int(window_half), int(cutoff), int(passes),
                                        ctypes.byref(r))
        try:
            self._check(rc)
            return SynfindResult(r)
        finally:
            self.lib.jx_synfind_result_free(ctypes.byref(r))

    # ---- instrumentation ----------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.jx_launch_count(self.ctx))

    def tokenizer_timing(self, reset=True):
        ms = ctypes.c_double(); n = ctypes.c_int64(); b = ctypes.c_int64(); ln = ctypes.c_int64()
        self._check(self.lib.jx_tokenizer_timing(self.ctx, int(reset), ctypes.byref(ms), ctypes.byref(n), ctypes.byref(b),
                                                 ctypes.byref(ln)))
        return ms.value, n.value, b.value, ln.value


_ENGINES = {}


def get_engine(device=None):
    """Process-wide engine for `device` (default: LOCAL_RANK or 0)."""
    import os
    if device is None:
        device = int(os.environ.get("JCVI_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    e = _ENGINES.get(device)
    if e is None:
        e = _ENGINES[device] = Engine(device)
    return e
