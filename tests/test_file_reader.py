"""jx_file_upload (jcvi_b200/csrc/jx_io.cu) in front of the tokenizer -- plain files through pinned
pieces, `.gz` inflated by the library (one stream; members; block-parallel BGZF), bzip2 through the Python fallback --
gives the same `.filtered` / `.lifted.anchors` bytes as the in-memory table (must_open, src/jcvi/formats/base.py:473-550)."""
import bz2
import gzip
import os
import shutil

import numpy as np
import pytest

from conftest import read


def test_bgzf_writer_is_a_gzip_file(tmp_path):
    from jcvi_b200 import bgzf
    rng = np.random.default_rng(3)
    data = (b"geneA.1\tgeneB.2\t98.5\t100\t1\t0\t1\t100\t5\t104\t1e-50\t200\n" * 3000) + rng.integers(0, 256, 200000, dtype=np.uint8).tobytes()
    p = str(tmp_path / "t.gz")
    bgzf.write_bgzf(data, p)
    assert gzip.open(p, "rb").read() == data                 # any gzip reader (the reference's must_open) reads it
    raw = read(p)
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC" and raw.endswith(bgzf._EOF)


@pytest.mark.gpu
def test_file_reader_variants_give_identical_outputs(tmp_path):
    from jcvi_b200 import bgzf, synth
    from jcvi_b200.compara import blastfilter, synteny
    from jcvi_b200.engine import get_engine
    d = str(tmp_path)
    qbed, sbed, last = synth.write_dataset(d, "A", "B", seed=21, Gq=5000, Gs=5000, n_lines=700000, Kq=6, Ks=6)   # ~ 50 MB: several pieces
    text = read(last)
    from jcvi_b200.compara.synteny import _load_bed
    get_engine().load_pair(_load_bed(qbed), _load_bed(sbed), False)
    want = get_engine().blastfilter(np.frombuffer(text, dtype=np.uint8), cscore=0.7, want_text=True)
    want_text = bytes(memoryview(want.text))
    variants = {}
    for tag in ("plain", "gz1", "gzmulti", "bgzf", "bz2"):
        sub = os.path.join(d, tag)
        os.makedirs(sub)
        shutil.copy(qbed, sub); shutil.copy(sbed, sub)
        f = os.path.join(sub, "A.B.last")
        if tag == "plain":
            open(f, "wb").write(text)
        elif tag == "gz1":
            f += ".gz"
            with gzip.open(f, "wb", compresslevel=4) as fh:
                fh.write(text)
        elif tag == "gzmulti":                                # concatenated ordinary members, not BGZF
            f += ".gz"
            cut = [0, len(text) // 3, len(text) // 3 + 1, len(text)]
            with open(f, "wb") as fh:
                for a, b in zip(cut, cut[1:]):
                    fh.write(gzip.compress(text[a:b], 1))
        elif tag == "bgzf":
            f += ".gz"
            bgzf.write_bgzf(text, f, level=4)
        else:
            f += ".bz2"
            with bz2.open(f, "wb", compresslevel=1) as fh:
                fh.write(text[: len(text) // 8 + text[len(text) // 8:].index(b"\n") + 1])
        variants[tag] = f
    eng = get_engine()
    for tag, f in variants.items():
        dev = eng.upload_file(f)
        if tag == "bz2":
            continue
        assert dev.nbytes == len(text), tag
        assert dev.kind == {"plain": "plain", "gz1": "gzip", "gzmulti": "gzip", "bgzf": "bgzf"}[tag]
        blastfilter.main([f, "--cscore=0.7", "--qbed=" + qbed, "--sbed=" + sbed])
        assert read(f + ".filtered") == want_text, tag
    # the whole CLI chain on the BGZF table == on the plain table
    outs = {}
    for tag in ("plain", "bgzf"):
        f = variants[tag]
        blastfilter.main([f, "--cscore=0.7", "--qbed=" + qbed, "--sbed=" + sbed])
        anchors = os.path.join(os.path.dirname(f), "A.B.anchors")
        outs[tag] = read(synteny.scan([f + ".filtered", anchors, "--liftover=" + f, "--qbed=" + qbed, "--sbed=" + sbed]))
    assert outs["plain"] == outs["bgzf"] and len(outs["plain"]) > 1000
    # bz2 goes through the Python reader
    f = variants["bz2"]
    blastfilter.main([f, "--cscore=0.7", "--qbed=" + qbed, "--sbed=" + sbed])
    assert os.path.getsize(f + ".filtered") > 1000
    # errors: a missing file, a truncated gzip stream, a corrupted BGZF block
    with pytest.raises(FileNotFoundError):
        eng.upload_file(os.path.join(d, "nope.last"))
    bad = os.path.join(d, "trunc.gz")
    open(bad, "wb").write(read(variants["gz1"])[:100000])
    with pytest.raises(ValueError):
        eng.upload_file(bad)
    raw = bytearray(read(variants["bgzf"]))
    raw[40000] ^= 0xFF; raw[40001] ^= 0xFF; raw[40002] ^= 0x55
    open(bad, "wb").write(raw)
    with pytest.raises(ValueError):
        eng.upload_file(bad)
    # a stream that inflates to far more than its size hint (last member small, ratio ~ 100): the device text grows while
    # the pieces are copied
    line = next(x for x in text.split(b"\n") if x and not x.startswith(b"#")) + b"\n"
    rep = line * 600000                                       # ~ 40 MB of one line
    sub = os.path.join(d, "grow"); os.makedirs(sub)
    shutil.copy(qbed, sub); shutil.copy(sbed, sub)
    fplain, fgz = os.path.join(sub, "A.B.last"), os.path.join(sub, "gz", "A.B.last.gz")
    os.makedirs(os.path.dirname(fgz)); shutil.copy(qbed, os.path.dirname(fgz)); shutil.copy(sbed, os.path.dirname(fgz))
    open(fplain, "wb").write(rep)
    with open(fgz, "wb") as fh:
        fh.write(gzip.compress(rep[: len(line) * 590000], 6)); fh.write(gzip.compress(rep[len(line) * 590000:], 6))
    assert os.path.getsize(fgz) * 20 < len(rep)
    dev = eng.upload_file(fgz)
    assert dev.nbytes == len(rep) and dev.kind == "gzip"
    blastfilter.main([fplain, "--cscore=0.7", "--qbed=" + qbed, "--sbed=" + sbed])
    blastfilter.main([fgz, "--cscore=0.7", "--qbed=" + qbed, "--sbed=" + sbed])
    assert read(fgz + ".filtered") == read(fplain + ".filtered") and len(read(fplain + ".filtered")) > 0
    # empty file
    open(os.path.join(d, "empty.last"), "wb").close()
    assert eng.upload_file(os.path.join(d, "empty.last")).nbytes == 0
