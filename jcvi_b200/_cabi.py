"""ctypes binding of libjcvi_b200.so (C ABI declared in include/jcvi_b200.h).

There is no CPU fallback: importing this module without the built library, or creating a
context without a CUDA device, raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JCVI_B200_LIB") or os.path.join(_HERE, "libjcvi_b200.so")

JX_OK, JX_ECUDA, JX_EARG, JX_EZERODIV, JX_EUNSUPPORTED, JX_EDUPANCHOR, JX_EINTERNAL = 0, -1, -2, -3, -4, -5, -6

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


class JxBed(ctypes.Structure):
    _fields_ = [("n_genes", ctypes.c_int64), ("names", ctypes.c_void_p), ("name_off", ctypes.c_void_p),
                ("indexable", ctypes.c_void_p), ("chr_lex", ctypes.c_void_p), ("chr_begin", ctypes.c_void_p),
                ("chr_end", ctypes.c_void_p), ("length", ctypes.c_void_p), ("accn_lexrank", ctypes.c_void_p),
                ("name_id", ctypes.c_void_p), ("n_chr", ctypes.c_int32)]


class JxFilterOpts(ctypes.Structure):
    _fields_ = [("strip_names", ctypes.c_int32), ("cscore", ctypes.c_double), ("tandem_nmax", ctypes.c_int32),
                ("exclude_keys", ctypes.c_void_p), ("n_exclude", ctypes.c_int64), ("want_text", ctypes.c_int32),
                ("want_mothers", ctypes.c_int32), ("skip_arrays", ctypes.c_int32)]


class JxFilterResult(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("q_rank", c_i32p), ("s_rank", c_i32p), ("score", c_f32p),
                ("line_off", c_u64p), ("text", ctypes.c_void_p), ("text_len", ctypes.c_int64),
                ("q_mother", c_i32p), ("s_mother", c_i32p), ("stats", ctypes.c_int64 * 8),
                ("text_pinned", ctypes.c_int32), ("device_token", ctypes.c_int64), ("owner", ctypes.c_void_p)]


class JxScanOpts(ctypes.Structure):
    _fields_ = [("xdist", ctypes.c_int32), ("ydist", ctypes.c_int32), ("min_size", ctypes.c_int32),
                ("intrabound", ctypes.c_int32)]


class JxScanResult(ctypes.Structure):
    _fields_ = [("n_anchors", ctypes.c_int64), ("n_blocks", ctypes.c_int64), ("q_rank", c_i32p), ("s_rank", c_i32p),
                ("score", c_f32p), ("block_start", c_i64p), ("index", c_i64p), ("stats", ctypes.c_int64 * 8)]


class JxLiftoverResult(ctypes.Structure):
    _fields_ = [("n_lifted", ctypes.c_int64), ("q_rank", c_i32p), ("s_rank", c_i32p), ("score", c_f32p),
                ("block", c_i32p), ("accepted", c_u8p), ("raw_score", c_f32p), ("stats", ctypes.c_int64 * 8)]


class JxSynfindResult(ctypes.Structure):
    _fields_ = [("n_forward", ctypes.c_int64), ("n_mirror", ctypes.c_int64), ("n_points", ctypes.c_int64), ("query", c_i32p),
                ("syntelog", c_i32p), ("left", c_i32p), ("right", c_i32p), ("score", c_i32p), ("gray", c_u8p), ("minus", c_u8p),
                ("stats", ctypes.c_int64 * 8)]


EXPORTS = [
    "jx_ctx_create", "jx_ctx_destroy", "jx_last_error", "jx_version", "jx_launch_count", "jx_stream_handle",
    "jx_tokenizer_timing", "jx_bed_load", "jx_pair_setup", "jx_blastfilter", "jx_filter_result_free",
    "jx_scan_text", "jx_scan_points", "jx_scan_filtered", "jx_scan_result_free", "jx_liftover_text",
    "jx_liftover_result_free", "jx_nearest_anchor", "jx_text_upload", "jx_text_release", "jx_set_record_reuse", "jx_distinct_names", "jx_best_table_copy", "jx_session_counts",
    "jx_host_blastline", "jx_blast_cscore", "jx_blast_filter", "jx_mcscan", "jx_block_stats",
    "jx_pipeline", "jx_shard_reset", "jx_shard_slots", "jx_shard_candidates", "jx_shard_partition", "jx_shard_pairs",
    "jx_shard_survivors", "jx_shard_order", "jx_scan_cands", "jx_shard_near", "jx_shard_lift", "jx_cscore_begin", "jx_cscore_finish", "jx_near_lines", "jx_cscore_finish_device", "jx_near_lines_device", "jx_free", "jx_abi_sizes",
    "jx_synfind_text", "jx_synfind_points", "jx_synfind_result_free", "jx_file_upload", "jx_abi_size_synfind", "jx_host_query_runs",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "jcvi_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.jx_last_error.restype = ctypes.c_char_p
        L.jx_last_error.argtypes = [ctypes.c_void_p]
        L.jx_version.restype = ctypes.c_char_p
        L.jx_launch_count.restype = ctypes.c_int64
        L.jx_launch_count.argtypes = [ctypes.c_void_p]
        L.jx_stream_handle.restype = ctypes.c_uint64
        L.jx_stream_handle.argtypes = [ctypes.c_void_p]
        L.jx_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.jx_ctx_destroy.argtypes = [ctypes.c_void_p]
        L.jx_ctx_destroy.restype = None
        L.jx_tokenizer_timing.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double), c_i64p, c_i64p, c_i64p]
        L.jx_bed_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(JxBed)]
        L.jx_pair_setup.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32]
        L.jx_blastfilter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                     ctypes.POINTER(JxFilterOpts), ctypes.POINTER(JxFilterResult)]
        L.jx_filter_result_free.argtypes = [ctypes.POINTER(JxFilterResult)]
        L.jx_filter_result_free.restype = None
        L.jx_scan_text.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                   ctypes.POINTER(JxScanOpts), ctypes.POINTER(JxScanResult)]
        L.jx_scan_points.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_int, ctypes.POINTER(JxScanOpts), ctypes.POINTER(JxScanResult)]
        L.jx_scan_filtered.argtypes = [ctypes.c_void_p, ctypes.POINTER(JxFilterResult), ctypes.POINTER(JxScanOpts),
                                       ctypes.POINTER(JxScanResult)]
        L.jx_scan_result_free.argtypes = [ctypes.POINTER(JxScanResult)]
        L.jx_scan_result_free.restype = None
        L.jx_liftover_text.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.POINTER(JxLiftoverResult)]
        L.jx_pipeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(JxFilterOpts),
                                  ctypes.POINTER(JxScanOpts), ctypes.c_int, ctypes.POINTER(JxFilterResult),
                                  ctypes.POINTER(JxScanResult), ctypes.POINTER(JxLiftoverResult)]
        L.jx_shard_reset.argtypes = [ctypes.c_void_p]
        L.jx_shard_slots.argtypes = [ctypes.c_void_p, c_i64p]
        L.jx_shard_candidates.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_uint32, c_u64p, c_i64p, ctypes.c_void_p, ctypes.c_int64]
        L.jx_shard_partition.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                         c_u64p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.jx_shard_pairs.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, c_u64p, c_i64p, c_u64p, c_i64p]
        L.jx_shard_survivors.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int,
                                         c_u64p, c_i64p, ctypes.c_void_p, ctypes.c_void_p]
        L.jx_shard_order.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.POINTER(JxFilterResult)]
        L.jx_scan_cands.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.POINTER(JxScanOpts), ctypes.POINTER(JxScanResult),
                                    c_u64p, ctypes.c_int]
        L.jx_shard_near.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32, c_u64p, c_i64p]
        L.jx_shard_lift.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(JxLiftoverResult), c_u64p]
        L.jx_liftover_result_free.argtypes = [ctypes.POINTER(JxLiftoverResult)]
        L.jx_liftover_result_free.restype = None
        L.jx_nearest_anchor.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.jx_text_upload.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64)]
        L.jx_text_release.argtypes = [ctypes.c_void_p]
        L.jx_file_upload.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64), c_i64p,
                                     ctypes.POINTER(ctypes.c_int32)]
        L.jx_set_record_reuse.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.jx_distinct_names.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                        ctypes.POINTER(ctypes.c_int64)]
        L.jx_best_table_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
        L.jx_blast_filter.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                      ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        L.jx_blast_cscore.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        L.jx_session_counts.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
        L.jx_host_blastline.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_float),
                                        ctypes.POINTER(ctypes.c_float), ctypes.c_char_p, ctypes.c_int]
        L.jx_cscore_begin.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]
        L.jx_cscore_finish.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p), c_i64p, c_i64p]
        L.jx_near_lines.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), c_i64p, c_i64p]
        L.jx_cscore_finish_device.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_uint64), c_i64p, c_i64p]
        L.jx_near_lines_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64), c_i64p, c_i64p]
        L.jx_mcscan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_void_p),
                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]
        L.jx_block_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32] + [ctypes.c_void_p] * 10
        L.jx_free.argtypes = [ctypes.c_void_p]
        L.jx_free.restype = None
        L.jx_synfind_text.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.POINTER(JxSynfindResult)]
        L.jx_synfind_points.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.POINTER(JxSynfindResult)]
        L.jx_synfind_result_free.argtypes = [ctypes.POINTER(JxSynfindResult)]
        L.jx_synfind_result_free.restype = None
        sizes = (ctypes.c_int32 * 6)()
        L.jx_abi_sizes(sizes)
        mine = [ctypes.sizeof(x) for x in (JxBed, JxFilterOpts, JxFilterResult, JxScanOpts, JxScanResult, JxLiftoverResult)]
        if list(sizes) != mine:
            raise ImportError("jcvi_b200: struct layout mismatch between _cabi.py %s and libjcvi_b200.so %s" % (mine, list(sizes)))
        L.jx_abi_size_synfind.restype = ctypes.c_int32
        L.jx_host_query_runs.argtypes = [ctypes.c_char_p, ctypes.c_int64]
        if L.jx_abi_size_synfind() != ctypes.sizeof(JxSynfindResult):
            raise ImportError("jcvi_b200: jx_synfind_result layout mismatch between _cabi.py and libjcvi_b200.so")
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def as_np(p, n, dtype):
    """Copy a library-owned array into numpy."""
    if n == 0 or not p:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


class _Owner(object):
    """Keeps a library result alive for the numpy views taken from it; frees it when the last view is gone."""

    def __init__(self, free, struct):
        self._free, self._struct = free, struct

    def __del__(self):
        try:
            self._free(ctypes.byref(self._struct))
        except Exception:
            pass


def view_np(p, n, dtype, owner):
    """numpy VIEW of a library-owned array (no copy): the ctypes buffer object the ndarray is based on holds a
    reference to `owner`, so the memory lives exactly as long as somebody can reach it."""
    if n == 0 or not p:
        return np.zeros(0, dtype=dtype)
    dt = np.dtype(dtype)
    buf = (ctypes.c_char * (int(n) * dt.itemsize)).from_address(ctypes.cast(p, ctypes.c_void_p).value)
    buf._owner = owner
    return np.frombuffer(buf, dtype=dt, count=int(n))

