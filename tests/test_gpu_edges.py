"""Tokenizer edge cases against the CPU oracle: rare paths of k_tokenize (lines longer than the
shared-memory window, tiles holding more than 1024 line starts, 127-byte names, line starts on
tile boundaries, CR LF and lone-CR terminators, no trailing newline)."""
import os

import numpy as np
import pytest

from conftest import read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from jcvi_b200.compara import blastfilter, synteny
    return blastfilter, synteny


def _beds(d, names_q, names_s):
    qbed, sbed = os.path.join(d, "A.bed"), os.path.join(d, "B.bed")
    with open(qbed, "w") as fw:
        for i, n in enumerate(names_q):
            fw.write("Achr%d\t%d\t%d\t%s\n" % (1 + i // 400, 1000 * i, 1000 * i + 500 + i % 7, n))
    with open(sbed, "w") as fw:
        for i, n in enumerate(names_s):
            fw.write("Bchr%d\t%d\t%d\t%s\n" % (1 + i // 400, 1000 * i, 1000 * i + 400 + i % 5, n))
    return qbed, sbed


def _check_all(d, mods, last, qbed, sbed, cscore=0.3, lift_dist=10):
    import oracle
    bf, syn = mods
    beds = ["--qbed=" + qbed, "--sbed=" + sbed]
    oracle.blastfilter(last, qbed, sbed, out_path=last + ".orc", cscore=cscore)
    bf.main([last] + beds + ["--cscore=%g" % cscore])
    assert read(last + ".filtered") == read(last + ".orc")
    try:
        oracle.scan(last + ".filtered", last + ".orc.anchors", qbed, sbed, min_size=2)
        err = None
    except oracle.ZeroAnchors as e:
        err = str(e)
    try:
        syn.scan([last + ".filtered", last + ".anchors"] + beds + ["--min_size=2"])
        gerr = None
    except ValueError as e:
        gerr = str(e)
    assert gerr == err
    assert read(last + ".anchors") == read(last + ".orc.anchors")
    if err is None:
        oracle.liftover(last, last + ".orc.anchors", qbed, sbed, out_path=last + ".orc.lifted", dist=lift_dist)
        syn.liftover([last, last + ".anchors"] + beds + ["--dist=%d" % lift_dist, "--outfile=" + last + ".lifted"])
        assert read(last + ".lifted") == read(last + ".orc.lifted")


def _lines(rng, nq, ns, n, qname, sname):
    out = []
    for _ in range(n):
        q = int(rng.integers(0, nq)); s = min(ns - 1, max(0, q + int(rng.integers(-3, 4)))) if rng.random() < 0.7 else int(rng.integers(0, ns))
        sc = "%.1f" % rng.uniform(50, 900)
        out.append("%s\t%s\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.1e\t%s" % (
            qname(q), sname(s), rng.uniform(30, 100), rng.integers(50, 900), rng.integers(0, 90), rng.integers(0, 9),
            rng.integers(1, 900), rng.integers(900, 1800), rng.integers(1, 900), rng.integers(900, 1800),
            10.0 ** -rng.uniform(3, 60), sc))
    return out


def test_long_names_and_long_lines(tmp_path, mods):
    """127-byte names (the cblast char[128] limit) and lines of 1.5-9 KB (longer than the 1 KiB
    look-ahead, some longer than a whole 8 KiB tile) mixed into ordinary lines."""
    rng = np.random.default_rng(51)
    nq = ns = 1200
    qn = lambda i: ("Q%05d_" % i).ljust(125, "x") + ".1" if i % 3 == 0 else "Ag%06d.1" % i
    sn = lambda i: ("S%05d_" % i).ljust(127, "y") if i % 4 == 0 else "Bg%06d.1" % i
    d = str(tmp_path)
    qbed, sbed = _beds(d, [qn(i)[:-2] if qn(i).endswith(".1") else qn(i) for i in range(nq)],
                       [sn(i)[:-2] if sn(i).endswith(".1") else sn(i) for i in range(ns)])
    lines = _lines(rng, nq, ns, 30000, qn, sn)
    for k in range(0, len(lines), 137):
        lines[k] = lines[k] + "\t" + "z" * int(rng.integers(1500, 9000))      # extra columns: ignored by sscanf
    for k in range(50, len(lines), 301):
        f = lines[k].split("\t"); f[11] = f[11] + "0" * 1200; lines[k] = "\t".join(f)   # >19 digits but trailing zeros only
    last = os.path.join(d, "A.B.last")
    open(last, "w").write("\n".join(lines) + "\n")
    _check_all(d, mods, last, qbed, sbed)


def test_many_short_lines_and_boundaries(tmp_path, mods):
    """Runs of blank / 2-byte lines put > 1024 line starts into one 8 KiB tile (chunked line table),
    and padding comments move real line starts onto every offset around the tile boundaries."""
    rng = np.random.default_rng(52)
    nq = ns = 900
    qn = lambda i: "Ag%06d.1" % i
    sn = lambda i: "Bg%06d.1" % i
    d = str(tmp_path)
    qbed, sbed = _beds(d, ["Ag%06d" % i for i in range(nq)], ["Bg%06d" % i for i in range(ns)])
    lines = _lines(rng, nq, ns, 20000, qn, sn)
    out = []
    for k, ln in enumerate(lines):
        out.append(ln)
        if k % 1000 == 500:
            out.extend([""] * 6000)                      # 6000 empty lines: ~0.7 tiles of pure '\n'
        if k % 1000 == 700:
            out.extend(["x"] * 3000)                     # 2-byte lines
        if k % 97 == 0:
            out.append("#" + "p" * int(rng.integers(0, 64)))   # jitter the alignment of what follows
    last = os.path.join(d, "A.B.last")
    open(last, "w").write("\n".join(out))               # no trailing newline
    _check_all(d, mods, last, qbed, sbed)


def test_crlf_and_lone_cr(tmp_path, mods):
    """Python universal newlines: CR LF and a lone CR both end a line."""
    rng = np.random.default_rng(53)
    nq = ns = 700
    qn = lambda i: "Ag%06d.1" % i
    sn = lambda i: "Bg%06d.1" % i
    d = str(tmp_path)
    qbed, sbed = _beds(d, ["Ag%06d" % i for i in range(nq)], ["Bg%06d" % i for i in range(ns)])
    lines = _lines(rng, nq, ns, 12000, qn, sn)
    body = ""
    for k, ln in enumerate(lines):
        body += ln + ("\r\n" if k % 3 == 0 else "\r" if k % 3 == 1 else "\n")
    last = os.path.join(d, "A.B.last")
    open(last, "w", newline="").write(body)
    _check_all(d, mods, last, qbed, sbed)


def test_name_table_boundaries(tmp_path, mods):
    """Names of every length 1..40 around the 24-byte slot of the perfect-hash table, names that
    share their first 24 bytes, '|' cuts and isoform suffixes, hits against names that are not in
    the BED (lookups must miss), a BED accn repeated on two lines (the last one wins), and number
    formats that take the word-arithmetic and the generic routes."""
    rng = np.random.default_rng(77)
    nq = ns = 1000

    def base(prefix, i):
        L = 1 + (i * 7) % 40
        core = "%s%d" % (prefix, i)
        if L > len(core):
            core = core + "_" + "k" * (L - len(core) - 1)
        if i % 10 == 3:
            core = "SHAREDPREFIX_OF_24_BYTES" + core          # equal first 24 bytes, differ after
        return core
    qb = [base("q", i) for i in range(nq)]
    sb = [base("s", i) for i in range(ns)]
    assert len(set(qb)) == nq and len(set(sb)) == ns
    d = str(tmp_path)
    qbed, sbed = _beds(d, qb, sb)
    with open(qbed, "a") as fw:                                # repeated accn: Bed.order keeps the last
        fw.write("Achr9\t5\t900\t%s\n" % qb[17])

    def qn(i):
        r = i % 6
        return qb[i] + (".1" if r == 0 else "|tail.2" if r == 1 else ".2|x" if r == 2 else "")
    sn = lambda i: sb[i] + (".3" if i % 2 else "")
    lines = _lines(rng, nq, ns, 40000, qn, sn)
    for k in range(0, len(lines), 23):                         # unknown names
        f = lines[k].split("\t"); f[k % 2] = f[k % 2] + "Z"; lines[k] = "\t".join(f)
    for k in range(5, len(lines), 31):                         # scores / evalues in other notations
        f = lines[k].split("\t")
        f[11] = ["1.06e+03", "1234", "0.5", "12345678", "7.", "1e2"][k % 6]
        f[10] = ["0.0", "1E-11", "3", "1.5e-180", "0.00000000001", "5e-010"][k % 6]
        lines[k] = "\t".join(f)
    last = os.path.join(d, "A.B.last")
    open(last, "w").write("\n".join(lines) + "\n")
    _check_all(d, mods, last, qbed, sbed)


def test_names_with_equal_hashes(tmp_path, mods):
    """Two BED names with the same 32-bit table hash cannot both live in the perfect-hash table: the
    second goes to the overflow list that lookups consult after a miss.  Colliding names are constructed
    from the hash's definition and confirmed with the library's own hash."""
    import ctypes
    probe = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jcvi_b200",
                                     "libjx_hostprobe.so"))
    # the name hash is a multiply-add chain over the name's 32-bit words (jx_fast5.cuh slot_hash_*): two 8-byte names
    # collide iff  w0 * C + w1  agree mod 2^32 -- solve for the second word and keep the solutions that are printable
    C = 0x9E3779B1
    ok = np.zeros(256, dtype=bool)
    ok[[ord(c) for c in "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_-+:=@~^"]] = True
    srng = np.random.default_rng(77)
    alpha = np.flatnonzero(ok).astype(np.uint8)
    pairs = []
    while len(pairs) < 12:
        a8 = srng.choice(alpha, size=(200000, 8))
        b4 = srng.choice(alpha, size=(200000, 4))
        w0a = a8[:, :4].copy().view("<u4")[:, 0].astype(np.uint64); w1a = a8[:, 4:].copy().view("<u4")[:, 0].astype(np.uint64)
        w0b = b4.copy().view("<u4")[:, 0].astype(np.uint64)
        w1b = ((w0a * C + w1a - w0b * C) & 0xFFFFFFFF).astype("<u4")
        byt = w1b.view(np.uint8).reshape(-1, 4)
        good = ok[byt].all(axis=1) & (w0a != w0b)
        for i in np.flatnonzero(good)[:16]:
            pairs.append((bytes(a8[i]).decode(), (bytes(b4[i]) + bytes(byt[i])).decode()))
    names = ["Ag%06d" % (7 * i + 3) for i in range(3000)] + [x for pr in pairs for x in pr]
    n = len(names)
    blob = "".join(names).encode()
    off = np.concatenate([[0], np.cumsum([len(x) for x in names])]).astype(np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    probe.probe_slot_hash_many(blob, off.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n), out.ctypes.data_as(ctypes.c_void_p))
    order = np.argsort(out, kind="stable")
    same = np.nonzero(out[order][1:] == out[order][:-1])[0]
    assert len(same) >= 12, "the constructed names do not collide"
    coll = sorted(set(order[same].tolist()) | set(order[same + 1].tolist()))
    rng = np.random.default_rng(5)
    others = rng.choice(n, 1500, replace=False).tolist()
    qset = sorted(set(coll) | set(others[:700]))
    sset = sorted(set(coll) | set(others[700:]))
    d = str(tmp_path)
    qbed, sbed = _beds(d, [names[i] for i in qset], [names[i] for i in sset])
    nq, ns = len(qset), len(sset)
    qn = lambda i: names[qset[i]] + ".1"
    sn = lambda i: names[sset[i]]
    lines = _lines(rng, nq, ns, 20000, qn, sn)
    cq = [k for k, i in enumerate(qset) if i in set(coll)]
    cs = [k for k, i in enumerate(sset) if i in set(coll)]
    for k in range(0, len(lines), 7):                          # make sure the colliding names are hit
        f = lines[k].split("\t"); f[0] = qn(cq[k % len(cq)]); f[1] = sn(cs[(k // 7) % len(cs)]); lines[k] = "\t".join(f)
    last = os.path.join(d, "A.B.last")
    open(last, "w").write("\n".join(lines) + "\n")
    _check_all(d, mods, last, qbed, sbed)


def test_tiny_and_truncated_tables(tmp_path, mods):
    """Texts shorter than one 16-byte bulk-copy granule, ending mid-line, or ending exactly on a
    granule boundary: the last tile is staged partly by the bulk copy and partly by hand."""
    import oracle
    bf, _ = mods
    rng = np.random.default_rng(9)
    nq = ns = 50
    d = str(tmp_path)
    qbed, sbed = _beds(d, ["Ag%06d" % i for i in range(nq)], ["Bg%06d" % i for i in range(ns)])
    body = "\n".join(_lines(rng, nq, ns, 40, lambda i: "Ag%06d.1" % i, lambda i: "Bg%06d.1" % i)) + "\n"
    cut_at = [0, 1, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, len(body) - 1, len(body)]
    first = body.index("\n") + 1
    cut_at += [first - 1, first, first + 1, first + 16, 8 * first]
    beds = ["--qbed=" + qbed, "--sbed=" + sbed]
    for n in sorted(set(cut_at)):
        last = os.path.join(d, "t%d.last" % n)
        open(last, "w").write(body[:n])
        oracle.blastfilter(last, qbed, sbed, out_path=last + ".orc", cscore=0)
        bf.main([last] + beds + ["--cscore=0"])
        assert read(last + ".filtered") == read(last + ".orc"), n


def test_writer_lines_longer_than_the_staged_window(tmp_path):
    """Survivors whose lines exceed the 208 bytes the `.filtered` writer stages per line (names of ~100 bytes): the reader
    continues in global memory; bytes identical to the oracle (and the tokenizer takes its generic route for such names)."""
    import oracle
    from jcvi_b200.compara import blastfilter
    d = str(tmp_path)
    rng = np.random.default_rng(4)
    nq = ns = 60
    qn = ["Q" + "q" * (70 + (i % 37)) + "%03d" % i for i in range(nq)]
    sn = ["S" + "s" * (75 + (i % 41)) + "%03d" % i for i in range(ns)]
    with open(os.path.join(d, "A.bed"), "w") as fw:
        for i, n in enumerate(qn):
            fw.write("chr%d\t%d\t%d\t%s\n" % (1 + i // 30, 1000 * i, 1000 * i + 500, n))
    with open(os.path.join(d, "B.bed"), "w") as fw:
        for i, n in enumerate(sn):
            fw.write("chr%d\t%d\t%d\t%s\n" % (1 + i // 30, 1000 * i, 1000 * i + 500, n))
    last = os.path.join(d, "A.B.last")
    with open(last, "w") as fw:
        for k in range(4000):
            i, j = int(rng.integers(0, nq)), int(rng.integers(0, ns))
            if k % 3 == 0:
                j = i
            fw.write("%s\t%s\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.1e\t%d\n" % (
                qn[i], sn[j], 70 + 30 * rng.random(), rng.integers(50, 2000), rng.integers(0, 50), rng.integers(0, 9),
                rng.integers(1, 900), rng.integers(900, 3000), rng.integers(1, 900), rng.integers(900, 3000),
                10.0 ** -float(rng.integers(3, 150)), rng.integers(100, 9000)))
    oracle.blastfilter(last, os.path.join(d, "A.bed"), os.path.join(d, "B.bed"), out_path=last + ".orc", cscore=0.5, tandem_Nmax=0)
    blastfilter.main([last, "--cscore=0.5", "--tandem_Nmax=0"])
    got, want = read(last + ".filtered"), read(last + ".orc")
    assert got == want and want.count(b"\n") > 100 and max(len(x) for x in want.split(b"\n")) > 208
