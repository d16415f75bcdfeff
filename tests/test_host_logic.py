"""CPU tests of the host side: the C ABI library loads and exports what include/jcvi_b200.h
declares, BED ranks agree with the oracle's restatement, AnchorFile/read_block semantics, argv
handling, and the multi-rank sharding (gloo, world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_header_symbols():
    import ctypes
    from jcvi_b200 import _cabi
    L = _cabi.lib()
    header = open(os.path.join(ROOT, "include", "jcvi_b200.h")).read()
    declared = set(re.findall(r"\b(jx_[a-z_0-9]+)\s*\(", header))
    assert {"jx_ctx_create", "jx_blastfilter", "jx_scan_text", "jx_liftover_text", "jx_text_upload"} <= declared
    for name in declared:
        assert hasattr(L, name), "libjcvi_b200.so does not export %s" % name
    assert set(_cabi.EXPORTS) == declared
    assert b"sm_100a" in L.jx_version()
    # struct layouts the Python side mirrors (sizes as the C compiler sees them)
    assert ctypes.sizeof(_cabi.JxScanOpts) == 16
    sizes = (ctypes.c_int32 * 6)()
    L.jx_abi_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(x) for x in (_cabi.JxBed, _cabi.JxFilterOpts, _cabi.JxFilterResult, _cabi.JxScanOpts,
                                                     _cabi.JxScanResult, _cabi.JxLiftoverResult)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jcvi_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)


def test_bed_ranks_match_oracle(tmp_path):
    import oracle
    from jcvi_b200.formats.bed import Bed, natsort_key
    p = tmp_path / "x.bed"
    rng = np.random.default_rng(3)
    rows = []
    for k, seqid in enumerate(["chr10", "chr2", "chr1", "scaffold_9", "scaffold_10", "1", "02", "X", "chrUn.3"]):
        pos = 0
        for g in range(40):
            pos += int(rng.integers(1, 500))
            rows.append("%s\t%d\t%d\tg%02d_%03d\t0\t+" % (seqid, pos, pos + int(rng.integers(1, 900)), k, g))
    rows.append("chr2\t5\t50\tdup_name\t0\t+")
    rows.append("chr10\t7\t70\tdup_name\t0\t-")
    rng.shuffle(rows)
    p.write_text("# header\ntrack name=foo\n\n" + "\n".join(rows) + "\n")
    bed = Bed(str(p))
    out = tmp_path / "order.txt"
    oracle.bed_order(str(p), str(out))
    assert [b.accn for b in bed] == out.read_text().split()
    assert natsort_key("chr10") > natsort_key("chr2") and natsort_key("1") < natsort_key("02") < natsort_key("X")
    t = bed.tables()
    assert t["G"] == len(rows) and t["n_chr"] == 9
    assert t["indexable"].sum() == len(rows) - 1            # later duplicate wins (bed.py:196-198)
    assert bed.order["dup_name"][0] == t["last"]["dup_name"]
    # chr_lex is the byte order of the seqids (sorted(chr_pair) in batch_scan, synteny.py:444)
    lex_of = {b.seqid: int(t["chr_lex"][i]) for i, b in enumerate(bed)}
    assert [s_ for s_, _ in sorted(lex_of.items(), key=lambda kv: kv[1])] == sorted(lex_of, key=lambda s_: s_.encode())
    assert (np.diff(t["chr_begin"]) >= 0).all()


def test_bed_with_colliding_natsort_keys_is_refused(tmp_path):
    from jcvi_b200.formats.bed import Bed
    p = tmp_path / "y.bed"
    p.write_text("chr1\t10\t20\ta\nchr01\t15\t25\tb\nchr1\t30\t40\tc\n")
    with pytest.raises(NotImplementedError):
        Bed(str(p)).tables()


def test_anchorfile_matches_reference_fixture(golden):
    # tests/compara/test_base.py:7-14 of the reference: 3 blocks, 14 pairs
    from jcvi_b200.compara.base import AnchorFile
    d, _ = golden("ref_fixtures")
    ac = AnchorFile(os.path.join(d, "anchorfile_test.anchors"))
    assert len(ac.blocks) == 3
    assert len(list(ac.iter_pairs())) == 14
    sizes = [len(b) for b in ac.blocks]
    assert len(list(ac.iter_pairs(minsize=5))) == sum(n for n in sizes if n >= 5)


def test_read_block_quirks(tmp_path):
    from jcvi_b200.compara.base import AnchorFile
    p = tmp_path / "q.anchors"
    p.write_text("###\n###\na\tb\t1\n# note\nc\td\t2\n")
    assert AnchorFile(str(p)).blocks == [[], [["a", "b", "1"]], [["c", "d", "2"]]]
    p.write_text("a\tb\t1\nc\td\t2\n")                    # no header at all: one block
    assert AnchorFile(str(p)).blocks == [[["a", "b", "1"], ["c", "d", "2"]]]
    p.write_text("")
    assert AnchorFile(str(p)).blocks == [[]]


def test_get_bed_filenames_inference():
    # tests/compara/test_synteny.py:23-40 of the reference
    from argparse import Namespace
    from jcvi_b200.compara.synteny import get_bed_filenames
    o = Namespace(qbed=None, sbed=None)
    assert get_bed_filenames("dir/grape.peach.last", None, o) == ("dir/grape.bed", "dir/peach.bed")
    o = Namespace(qbed="x.bed", sbed="y.bed")
    assert get_bed_filenames("a.b.last", None, o) == ("x.bed", "y.bed")


def test_cli_misuse_exits_like_the_reference():
    from jcvi_b200.compara import blastfilter, synteny
    with pytest.raises(SystemExit):
        blastfilter.main([])
    with pytest.raises(SystemExit):
        synteny.scan(["only_one_arg"])


def test_assign_pairs_is_a_balanced_partition():
    from jcvi_b200.parallel import assign_pairs
    sizes = [90, 10, 50, 50, 40, 5, 5, 70, 30, 1]
    for world in (1, 2, 4, 8):
        parts = assign_pairs(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) <= sum(sizes) / world + max(sizes)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from jcvi_b200.parallel import assign_pairs, max_over_ranks, rank_world
dist.init_process_group("gloo")
rank, world = rank_world()
sizes = [90, 10, 50, 50, 40, 5, 5, 70, 30, 1]
mine = assign_pairs(sizes, world)[rank]
t = max_over_ranks(1.0 + rank)
got = [None] * world
dist.all_gather_object(got, mine)
# a failure only one rank sees stops every rank with the same exception class instead of leaving the rest in a collective
from jcvi_b200.sharded import _agree, _raise_agreed
_agree(None)
try:
    _agree(ZeroDivisionError("float division by zero") if rank == 1 else None)
    agreed = "no exception"
except ZeroDivisionError as e:
    agreed = "ZeroDivisionError"
flags = [None] * world
dist.all_gather_object(flags, agreed)
if rank == 0:
    assert sorted(i for p in got for i in p) == list(range(len(sizes))), got
    assert t == float(world), t
    assert flags == ["ZeroDivisionError"] * world, flags
    print("GLOO_OK", got)
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_query_run_sample_tells_grouped_from_shuffled_tables(tmp_path):
    """The host test that picks the tokenizer's MERGE_Q instantiation (jx_sample_query_runs)."""
    import numpy as np
    from jcvi_b200 import _cabi, synth
    L = _cabi.lib()
    qbed, sbed, last = synth.write_dataset(str(tmp_path), "A", "B", 5, 2000, 2000, 40000, Kq=3, Ks=3)
    text = open(last, "rb").read()
    assert L.jx_host_query_runs(text, len(text)) == 0                      # the generator shuffles the lines
    rows = [r for r in text.split(b"\n") if r and not r.startswith(b"#")]
    rows.sort(key=lambda r: r.split(b"\t", 1)[0])
    grouped = b"# header\n" + b"\n".join(rows) + b"\n"
    assert L.jx_host_query_runs(grouped, len(grouped)) == 1               # lastal / BLAST order
    assert L.jx_host_query_runs(b"", 0) == 0
    assert L.jx_host_query_runs(b"a\tb\n" * 8, 32) == 0                    # too few lines to tell
    assert L.jx_host_query_runs(b"a\tb\n" * 64, 256) == 1
    assert L.jx_host_query_runs(b"\n\n#x\n" * 100, 500) == 0              # blank and comment lines only


def test_sharded_status_agreement_picks_one_exception_class():
    import numpy as np
    import pytest
    from jcvi_b200.sharded import _raise_agreed, _status_code
    _raise_agreed(np.zeros(4, dtype=np.uint32), [None])
    mine = NotImplementedError("unsupported line 17")
    with pytest.raises(NotImplementedError, match="line 17"):
        _raise_agreed([0, _status_code(mine), 0, 2], [mine])           # the failing rank re-raises its own exception
    with pytest.raises(NotImplementedError, match="rank 1"):
        _raise_agreed([0, _status_code(mine), 0, 2], [None])           # the others raise the class of the lowest failing rank
    with pytest.raises(ZeroDivisionError):
        _raise_agreed([0, 0, 0, _status_code(ZeroDivisionError())], [None])
    with pytest.raises(RuntimeError):
        _raise_agreed([_status_code(KeyError("x"))], [None])


def test_install_rebinds_the_reference_names(tmp_path, monkeypatch):
    """jcvi_b200.install() against a stand-in `jcvi` package with the import shape of
    compara/catalog.py:21,23 (`from ..compara.blastfilter import main as blastfilter_main`,
    `from ..compara.synteny import liftover, mcscan, scan`)."""
    import importlib
    pkg = tmp_path / "jcvi" / "compara"
    pkg.mkdir(parents=True)
    (tmp_path / "jcvi" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("")
    (pkg / "blastfilter.py").write_text("def main(args):\n    return 'ref'\ndef blastfilter_main(a, b, c):\n    return 'ref'\n")
    (pkg / "synteny.py").write_text("def scan(a):\n    return 'ref'\ndef liftover(a):\n    return 'ref'\ndef mcscan(a):\n    return 'ref'\n"
                                    "def batch_scan(*a, **k):\n    return 'ref'\ndef synteny_scan(*a, **k):\n    return 'ref'\n"
                                    "def synteny_liftover(*a, **k):\n    return 'ref'\ndef simple(a):\n    return 'ref'\n")
    fm = tmp_path / "jcvi" / "formats"
    fm.mkdir()
    (fm / "__init__.py").write_text("")
    (fm / "blast.py").write_text("def cscore(args):\n    return 'ref'\ndef filter(args):\n    return 'ref'\n")
    (pkg / "catalog.py").write_text("from ..compara.blastfilter import main as blastfilter_main\n"
                                    "from ..compara.synteny import liftover, mcscan, scan\n"
                                    "from ..formats.blast import cscore, filter as blast_filter\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    for m in [k for k in sys.modules if k == "jcvi" or k.startswith("jcvi.")]:
        monkeypatch.delitem(sys.modules, m)
    import jcvi_b200
    from jcvi_b200.compara import blastfilter as gbf, synteny as gsyn
    jcvi_b200.install()
    cat = importlib.import_module("jcvi.compara.catalog")
    assert cat.blastfilter_main is gbf.main and cat.scan is gsyn.scan and cat.liftover is gsyn.liftover
    assert cat.mcscan is gsyn.mcscan
    from jcvi_b200.formats import blast as gblast
    assert cat.cscore is gblast.cscore and importlib.import_module("jcvi.formats.blast").cscore is gblast.cscore
    assert cat.blast_filter is gblast.filter and importlib.import_module("jcvi.formats.blast").filter is gblast.filter
    assert importlib.import_module("jcvi.compara.synteny").batch_scan is gsyn.batch_scan
    for m in [k for k in sys.modules if k == "jcvi" or k.startswith("jcvi.")]:
        monkeypatch.delitem(sys.modules, m)


def test_result_views_keep_the_library_result_alive():
    """_cabi.view_np: numpy views of a library-owned array free the result exactly when the last view (or slice
    of one) is unreachable -- LiftoverResult hands out such views instead of copies."""
    import ctypes
    import gc
    import numpy as np
    from jcvi_b200 import _cabi as C
    libc = ctypes.CDLL("libc.so.6")
    libc.malloc.restype = ctypes.c_void_p
    libc.free.argtypes = [ctypes.c_void_p]

    class S(ctypes.Structure):
        _fields_ = [("p", ctypes.POINTER(ctypes.c_int32))]

    freed = []

    def free(ref):
        libc.free(ctypes.cast(ref._obj.p, ctypes.c_void_p))
        freed.append(1)

    s = S()
    s.p = ctypes.cast(libc.malloc(40), ctypes.POINTER(ctypes.c_int32))
    for i in range(10):
        s.p[i] = i * i
    own = C._Owner(free, s)
    a = C.view_np(s.p, 10, np.int32, own)
    b = C.view_np(s.p, 4, np.int32, own)
    del own, s
    gc.collect()
    assert a.tolist() == [i * i for i in range(10)] and not freed
    c = a[2:5]
    del a, b
    gc.collect()
    assert c.tolist() == [4, 9, 16] and not freed
    del c
    gc.collect()
    assert freed == [1]
    assert len(C.view_np(None, 0, np.int32, None)) == 0


def test_in_memory_hand_over_equals_reading_the_files_back(tmp_path, capsys):
    """scan() hands its freshly written blocks to summary / liftover instead of parsing the file again:
    AnchorFile.from_blocks must equal AnchorFile(file) for what write_anchors / print_to_file emit, and the
    in-memory summary must print what summary([file]) prints (and raise the same ValueError on nothing)."""
    import pytest
    from jcvi_b200.compara.base import AnchorFile
    from jcvi_b200.compara import synteny as gsyn
    blocks = [[["a1", "b1", "120"], ["a2", "b2", "98"], ["a2", "b3", "55L"]], [["a9", "b7", "300"]]]
    ac = AnchorFile.from_blocks(str(tmp_path / "x.anchors"), blocks)
    ac.print_to_file(str(tmp_path / "x.anchors"))
    back = AnchorFile(str(tmp_path / "x.anchors"))
    assert [[list(r) for r in b] for b in back.blocks] == blocks == [[list(r) for r in b] for b in ac.blocks]
    gsyn._summary_blocks(ac.blocks)
    mem = capsys.readouterr().err
    gsyn.summary([str(tmp_path / "x.anchors")])
    assert capsys.readouterr().err == mem and "A total of 4 (NR:3) anchors found in 2 clusters." in mem
    (tmp_path / "empty.anchors").write_text("")
    assert AnchorFile(str(tmp_path / "empty.anchors")).blocks == AnchorFile.from_blocks("e", []).blocks == [[]]
    for call in (lambda: gsyn._summary_blocks([]), lambda: gsyn._summary_blocks([[]]), lambda: gsyn.summary([str(tmp_path / "empty.anchors")])):
        with pytest.raises(ValueError, match="A total of 0 anchor was found. Aborted."):
            call()


def test_read_text_file_maps_large_tables_and_file_key_tracks_changes(tmp_path):
    import os
    import numpy as np
    from jcvi_b200.apps_base import file_key, read_text_file
    small = tmp_path / "s.last"
    small.write_bytes(b"a\tb\n" * 10)
    assert isinstance(read_text_file(str(small)), np.ndarray) and not isinstance(read_text_file(str(small)), np.memmap)
    big = tmp_path / "b.last"
    line = b"Ag000001.1\tBg000002.1\t99.0\t100\t0\t0\t1\t100\t1\t100\t1e-50\t200.0\n"
    with open(big, "wb") as fw:
        for _ in range(70):
            fw.write(line * 16384)
    m = read_text_file(str(big))
    assert isinstance(m, np.memmap) and len(m) == os.path.getsize(big) and m[:len(line)].tobytes() == line
    k0 = file_key(str(big))
    assert k0 == file_key(str(big))
    with open(big, "ab") as fw:
        fw.write(line)
    assert file_key(str(big)) != k0
    import gzip
    gz = tmp_path / "g.last.gz"
    with gzip.open(gz, "wb") as fw:
        fw.write(line * 3)
    assert read_text_file(str(gz)).tobytes() == line * 3


def test_bed_order_cache_follows_a_resort(tmp_path):
    from jcvi_b200.formats.bed import Bed
    p = tmp_path / "x.bed"
    p.write_text("chr2\t10\t20\tg3\nchr1\t30\t40\tg2\nchr1\t5\t9\tg1\n")
    b = Bed(str(p))
    assert [f.accn for f in b] == ["g1", "g2", "g3"] and b.order["g3"][0] == 2 and b.order is b.order
    b.sort(key=lambda f: -f.start)
    assert b.order["g2"][0] == 0 and b.order["g1"][0] == 2
