"""Host-side BED handling of the hot path (kept on the host: it is O(genes), not O(hits)).

Mirrors the part of the reference that DEFINES gene ranks:
  BedLine  src/jcvi/formats/bed.py:40-73     (start = col2 + 1, accn = col4)
  Bed      src/jcvi/formats/bed.py:141-198   (skip blank/#/browser/track; sort by
           (natsort_key(seqid), start, accn); order[accn] = (rank, BedLine), later duplicate wins)
natsort is an unpinned third-party dependency of the reference (pyproject.toml:49) that is
not installed here; `natsort_key` restates its default algorithm (split on runs of decimal
digits, ints for digit runs, leading "" when the first chunk is numeric).
"""
import re

import numpy as np

_num = re.compile(r"([0-9]+)")


def natsort_key(s):
    parts = [p for p in _num.split(s) if p != ""]
    out = [int(p) if (p[0] in "0123456789" and p.isascii() and p.isdigit()) else p for p in parts]
    if out and isinstance(out[0], int):
        out.insert(0, "")
    return tuple(out)


class BedLine(object):
    """One BED feature with the attribute names the reference's callers use (formats/bed.py:40-73):
    `start` is 1-based (column 2 + 1), `accn` is column 4, optional score / strand / extra."""

    __slots__ = ("seqid", "start", "end", "accn", "extra", "score", "strand", "args", "nargs")

    def __init__(self, sline):
        cols = sline.strip().split("\t")
        n = len(cols)
        self.args, self.nargs = cols, n
        self.seqid = cols[0]
        self.start, self.end = int(cols[1]) + 1, int(cols[2])
        assert self.start <= self.end, "start={0} end={1}".format(self.start, self.end)
        self.accn = cols[3] if n > 3 else None
        self.score = cols[4] if n > 4 else None
        self.strand = cols[5] if n > 5 else None
        self.extra = cols[6:] if n > 6 else None

    def __str__(self):
        out = [self.seqid, str(self.start - 1), str(self.end)]
        out += [x for x in (self.accn, self.score, self.strand) if x is not None]
        out += [str(x) for x in (self.extra or [])]
        return "\t".join(out)

    __repr__ = __str__

    def __getitem__(self, key):
        return getattr(self, key)


def _is_feature_line(line):
    """Bed.__init__ (formats/bed.py:154-160) skips blank lines, '#' comments and browser / track
    declarations."""
    return bool(line.strip()) and line[0] != "#" and not line.startswith(("browser ", "track name"))


class Bed(list):
    """Features sorted by (natural-sort key of seqid, start, accn): the index in this list is the
    gene RANK every kernel works with (formats/bed.py:147,166-167,196-198)."""

    def __init__(self, filename=None, key=None, sorted=True):
        super().__init__()
        self.filename = filename
        seqid_keys = {}                    # a BED has few distinct seqids: one natural-sort key per seqid, not per row

        def nullkey(x):
            k = seqid_keys.get(x.seqid)
            if k is None:
                k = seqid_keys[x.seqid] = natsort_key(x.seqid)
            return (k, x.start, x.accn)

        self.nullkey = nullkey
        self.key = key or self.nullkey
        self._tables = None
        if not filename:
            return
        with open(filename) as fp:
            self.extend(BedLine(line) for line in fp if _is_feature_line(line))
        if sorted:
            self.sort(key=self.key)

    def sort(self, *args, **kwargs):
        self._order = None                 # ranks change: drop what was derived from the old order
        self._tables = None
        super().sort(*args, **kwargs)

    @property
    def order(self):
        """accn -> (rank, BedLine); for duplicated accns the later rank wins, like the reference's
        dict comprehension."""
        cached = getattr(self, "_order", None)
        if cached is None or cached[0] != len(self):           # built once per (unchanged) list, not per access
            cached = self._order = (len(self), {f.accn: (i, f) for i, f in enumerate(self)})
        return cached[1]

    @property
    def seqids(self):
        return sorted({b.seqid for b in self}, key=natsort_key)

    # ---- device tables ---------------------------------------------------------------------
    def tables(self):
        """Arrays in rank order for jx_bed_load (see include/jcvi_b200.h, struct jx_bed)."""
        if getattr(self, "_tables", None) is not None:
            return self._tables
        G = len(self)
        accns = [b.accn for b in self]
        if any(a is None for a in accns):
            raise AssertionError("BED needs >= 4 columns")
        enc = [a.encode() for a in accns]
        lens = np.fromiter((len(e) for e in enc), dtype=np.int64, count=G)
        off = np.zeros(G + 1, dtype=np.uint32)
        np.cumsum(lens, out=off[1:])
        blob = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if G else np.zeros(0, dtype=np.uint8)
        last = {}
        for i, a in enumerate(accns):
            last[a] = i
        indexable = np.zeros(G, dtype=np.uint8)
        indexable[list(last.values())] = 1
        # seqid runs; ranks of one seqid must be contiguous (equal natsort keys of distinct
        # seqids, e.g. chr1 / chr01, would interleave them)
        seqid_of = [b.seqid for b in self]
        chr_begin = np.zeros(G, dtype=np.int32)
        chr_end = np.zeros(G, dtype=np.int32)
        runs = []
        i = 0
        while i < G:
            j = i
            while j < G and seqid_of[j] == seqid_of[i]:
                j += 1
            runs.append((seqid_of[i], i, j))
            chr_begin[i:j] = i
            chr_end[i:j] = j
            i = j
        names = [r[0] for r in runs]
        if len(set(names)) != len(names):
            raise NotImplementedError(
                "seqids with equal natural-sort keys interleave in the BED order (%s); not supported" % self.filename)
        lex = {s: k for k, s in enumerate(sorted(names, key=lambda s: s.encode()))}
        chr_lex = np.zeros(G, dtype=np.int32)
        for s, a, b in runs:
            chr_lex[a:b] = lex[s]
        length = np.fromiter((abs(b.end - b.start) for b in self), dtype=np.int64, count=G).astype(np.int32)
        order = sorted(range(G), key=lambda i: (enc[i], i))
        lexrank = np.zeros(G, dtype=np.uint32)
        lexrank[order] = np.arange(G, dtype=np.uint32)
        self._tables = dict(G=G, names=blob, name_off=off, indexable=indexable, chr_lex=chr_lex, chr_begin=chr_begin,
                            chr_end=chr_end, length=length, accn_lexrank=lexrank, n_chr=len(runs), accns=accns,
                            seqids=names, last=last)
        return self._tables
