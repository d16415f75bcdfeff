"""Seeded synthetic genomes + BLAST/LAST tab12 hit tables (bench/test tooling).

Follows the recipe validated on the reference in SURVEY.md section 8(d): rearranged
ortholog blocks (30 % inverted), gene loss, multi-HSP duplicates, adjacent tandem copies,
mid-score "family" hits, noise, isoform-suffixed BLAST names (exercise gene_name), names
absent from the BED, '#' header lines, shuffled line order, natural-vs-lexical seqid
ordering (chr1..chrK with K >= 10).

The per-line text formatting is done by a small threaded C++ helper
(csrc/synth_format.cpp -> libjcvi_synth.so) so that 20 M+ line tables are produced in
seconds; everything that decides *content* is numpy with a fixed seed.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjcvi_synth.so")
_SRC = os.path.join(_HERE, "csrc", "synth_format.cpp")
_LIB = None


def build_synth(force=False):
    if force or not os.path.exists(_SO) or (
        os.path.exists(_SRC) and os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    ):
        cxx = os.environ.get("CXX", "g++")
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", _SO, _SRC])
    return _SO


def _lib():
    global _LIB
    if _LIB is None:
        build_synth()
        _LIB = ctypes.CDLL(_SO)
        _LIB.synth_format_lines.restype = ctypes.c_int64
    return _LIB


class Genome:
    """G genes on K chromosomes; rank order == BED sort order (natsort seqid, start, accn)."""

    def __init__(self, prefix, G, K, rng, digits=6):
        self.prefix, self.G, self.K, self.digits = prefix, int(G), int(K), digits
        # chromosome sizes roughly equal, in natural order <P>chr1..<P>chrK
        bounds = np.linspace(0, G, K + 1).astype(np.int64)
        self.chr_of = np.repeat(np.arange(K), np.diff(bounds)).astype(np.int32)
        self.chr_start = bounds[:-1]
        gaps = rng.integers(2000, 20000, size=G)
        lens = rng.integers(300, 5000, size=G)
        start0 = np.zeros(G, dtype=np.int64)
        for k in range(K):
            a, b = bounds[k], bounds[k + 1]
            pos = np.cumsum(gaps[a:b] + np.concatenate(([0], lens[a:b][:-1])))
            start0[a:b] = pos
        self.start0 = start0            # BED col2
        self.end = start0 + lens        # BED col3
        self.seqids = ["%schr%d" % (prefix, k + 1) for k in range(K)]

    def accn(self, i):
        return "%sg%0*d" % (self.prefix, self.digits, i)

    def write_bed(self, path, rng=None):
        """Lines are written in shuffled order so that the BED sort matters."""
        order = np.arange(self.G)
        if rng is not None:
            order = rng.permutation(self.G)
        with open(path, "w") as fw:
            fw.write("# synthetic genome %s\n" % self.prefix)
            out = []
            for i in order:
                out.append("%s\t%d\t%d\t%s\t0\t+\n" % (self.seqids[self.chr_of[i]], self.start0[i], self.end[i], self.accn(i)))
            fw.write("".join(out))


def _block_map(Gsrc, Gdst, rng, lo=50, hi=300, inv=0.3):
    """Map src ranks onto dst ranks by permuting blocks (30 % inverted)."""
    sizes = []
    tot = 0
    while tot < Gsrc:
        s = int(rng.integers(lo, hi + 1))
        sizes.append(min(s, Gsrc - tot))
        tot += sizes[-1]
    sizes = np.array(sizes)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    perm = rng.permutation(len(sizes))
    dst = np.empty(Gsrc, dtype=np.int64)
    pos = 0
    for b in perm:
        n = sizes[b]
        tgt = np.arange(pos, pos + n)
        if rng.random() < inv:
            tgt = tgt[::-1]
        dst[starts[b]:starts[b] + n] = tgt
        pos += n
    # squeeze onto the destination genome
    return (dst * Gdst // max(Gsrc, 1)).astype(np.int64)


def make_hits(gq, gs, n_lines, rng, self_mode=False, p_loss=0.25, p_hsp2=0.30, p_tandem=0.05,
              frac_mid=0.35, p_dup=0.2, p_absent=0.001, p_iso=0.05):
    """Return per-line arrays (qgene, sgene, qiso, siso, score10, pct100, ints7)."""
    Gq, Gs = gq.G, gs.G
    copies = max(1, int(round(Gq / Gs)))
    per = -(-Gq // copies)
    ortho = np.empty(Gq, dtype=np.int64)
    for c in range(copies):
        a, b = c * per, min(Gq, (c + 1) * per)
        ortho[a:b] = _block_map(b - a, Gs, rng)
    if self_mode:
        # WGD-like: paralog of gene i sits >= 1000 ranks away
        ortho = (ortho + Gs // 3) % Gs
    keep = rng.random(Gq) > p_loss
    q0 = np.nonzero(keep)[0]
    s0 = ortho[q0]
    sc0 = rng.integers(3000, 25000, size=len(q0))
    Q, S, SC = [q0], [s0], [sc0]
    m = rng.random(len(q0)) < p_hsp2                      # second, weaker HSP of the same pair
    Q.append(q0[m]); S.append(s0[m]); SC.append((sc0[m] * rng.uniform(0.2, 0.9, m.sum())).astype(np.int64))
    t = np.nonzero(rng.random(len(q0)) < p_tandem)[0]      # adjacent subject copies (tandem arrays)
    for k in (1, 2, 3):
        tk = t[rng.random(len(t)) < (1.0 / k)]
        Q.append(q0[tk]); S.append(np.minimum(s0[tk] + k, Gs - 1)); SC.append((sc0[tk] * rng.uniform(0.75, 0.98, len(tk))).astype(np.int64))
    t2 = np.nonzero(rng.random(len(q0)) < p_tandem)[0]     # adjacent query copies
    for k in (1, 2):
        tk = t2[rng.random(len(t2)) < (1.0 / k)]
        Q.append(np.minimum(q0[tk] + k, Gq - 1)); S.append(s0[tk]); SC.append((sc0[tk] * rng.uniform(0.75, 0.98, len(tk))).astype(np.int64))
    if self_mode:
        # self hits (dropped), mirrored orientation, near-diagonal local duplicates
        sh = rng.integers(0, Gq, size=max(1, n_lines // 50))
        Q.append(sh); S.append(sh); SC.append(rng.integers(5000, 30000, size=len(sh)))
        Q.append(s0.copy()); S.append(q0.copy()); SC.append((sc0 * rng.uniform(0.9, 1.0, len(q0))).astype(np.int64))
        nd = rng.integers(0, Gq - 8, size=max(1, n_lines // 40))
        Q.append(nd); S.append(nd + rng.integers(1, 8, size=len(nd))); SC.append(rng.integers(800, 9000, size=len(nd)))
    Q = np.concatenate(Q); S = np.concatenate(S); SC = np.concatenate(SC)
    n_struct = len(Q)
    n_rand = max(0, n_lines - n_struct)
    n_mid = int(n_rand * frac_mid / max(frac_mid + (1 - frac_mid), 1e-9)) if n_rand else 0
    rq = rng.integers(0, Gq, size=n_rand)
    rs = rng.integers(0, Gs, size=n_rand)
    rsc = np.concatenate((rng.integers(600, 9000, size=n_mid), rng.integers(300, 3000, size=n_rand - n_mid)))
    rng.shuffle(rsc)
    if n_rand:
        d = np.nonzero(rng.random(n_rand) < p_dup)[0]      # multi-HSP duplicates of other random pairs
        src = rng.integers(0, n_rand, size=len(d))
        rq[d] = rq[src]; rs[d] = rs[src]
    Q = np.concatenate((Q, rq))[:max(n_lines, 0)] if n_lines < n_struct + n_rand else np.concatenate((Q, rq))
    S = np.concatenate((S, rs))[:len(Q)]
    SC = np.concatenate((SC, rsc))[:len(Q)]
    n = len(Q)
    perm = rng.permutation(n)
    Q, S, SC = Q[perm], S[perm], SC[perm]
    qiso = np.ones(n, dtype=np.uint8); siso = np.ones(n, dtype=np.uint8)
    m = rng.random(n) < p_iso; qiso[m] = rng.integers(2, 4, size=m.sum())
    m = rng.random(n) < p_iso; siso[m] = rng.integers(2, 4, size=m.sum())
    m = rng.random(n) < p_absent; Q = Q.copy(); Q[m] = Gq + rng.integers(1, 1000, size=m.sum())
    m = rng.random(n) < p_absent; S = S.copy(); S[m] = Gs + rng.integers(1, 1000, size=m.sum())
    pct = rng.integers(2500, 10001, size=n)
    ints7 = np.empty((n, 7), dtype=np.int32)
    ints7[:, 0] = rng.integers(30, 1500, size=n)      # hitlen
    ints7[:, 1] = rng.integers(0, 300, size=n)        # nmismatch
    ints7[:, 2] = rng.integers(0, 20, size=n)         # ngaps
    qa = rng.integers(1, 3000, size=n); qb = qa + ints7[:, 0]
    sa = rng.integers(1, 3000, size=n); sb = sa + ints7[:, 0]
    rev = rng.random(n) < 0.3                          # minus-strand hits: sstart > sstop
    ints7[:, 3] = qa; ints7[:, 4] = qb
    ints7[:, 5] = np.where(rev, sb, sa); ints7[:, 6] = np.where(rev, sa, sb)
    return (Q.astype(np.int32), S.astype(np.int32), qiso, siso, SC.astype(np.int32), pct.astype(np.int32), ints7)


def make_hits_chunk(gq, gs, n_lines, chunk, nchunks, seed, p_loss=0.25, p_hsp2=0.30, p_tandem=0.05, frac_mid=0.35,
                    p_dup=0.2, p_absent=0.001, p_iso=0.05):
    """Chunk `chunk` of a table that is generated in `nchunks` independent pieces (ranks of a sharded run generate
    only their own slice): the structured hits (orthologs, second HSPs, tandem copies) come from ONE seeded map shared by
    all chunks and are dealt round-robin over them; the random hits are the chunk's own.  Same columns as make_hits."""
    rs = np.random.default_rng(seed)
    Gq, Gs = gq.G, gs.G
    copies = max(1, int(round(Gq / Gs)))
    per = -(-Gq // copies)
    ortho = np.empty(Gq, dtype=np.int64)
    for c in range(copies):
        a, b = c * per, min(Gq, (c + 1) * per)
        ortho[a:b] = _block_map(b - a, Gs, rs)
    keep = rs.random(Gq) > p_loss
    q0 = np.nonzero(keep)[0]
    s0 = ortho[q0]
    sc0 = rs.integers(3000, 25000, size=len(q0))
    Q, S, SC = [q0], [s0], [sc0]
    m = rs.random(len(q0)) < p_hsp2
    Q.append(q0[m]); S.append(s0[m]); SC.append((sc0[m] * rs.uniform(0.2, 0.9, m.sum())).astype(np.int64))
    t = np.nonzero(rs.random(len(q0)) < p_tandem)[0]
    for k in (1, 2, 3):
        tk = t[rs.random(len(t)) < (1.0 / k)]
        Q.append(q0[tk]); S.append(np.minimum(s0[tk] + k, Gs - 1)); SC.append((sc0[tk] * rs.uniform(0.75, 0.98, len(tk))).astype(np.int64))
    t2 = np.nonzero(rs.random(len(q0)) < p_tandem)[0]
    for k in (1, 2):
        tk = t2[rs.random(len(t2)) < (1.0 / k)]
        Q.append(np.minimum(q0[tk] + k, Gq - 1)); S.append(s0[tk]); SC.append((sc0[tk] * rs.uniform(0.75, 0.98, len(tk))).astype(np.int64))
    Q = np.concatenate(Q)[chunk::nchunks]; S = np.concatenate(S)[chunk::nchunks]; SC = np.concatenate(SC)[chunk::nchunks]
    rng = np.random.default_rng(seed * 1000003 + 17 * chunk + 1)
    n_rand = max(0, n_lines - len(Q))
    n_mid = int(n_rand * frac_mid)
    rq = rng.integers(0, Gq, size=n_rand)
    rsb = rng.integers(0, Gs, size=n_rand)
    rsc = np.concatenate((rng.integers(600, 9000, size=n_mid), rng.integers(300, 3000, size=n_rand - n_mid)))
    rng.shuffle(rsc)
    if n_rand:
        d = np.nonzero(rng.random(n_rand) < p_dup)[0]
        src = rng.integers(0, n_rand, size=len(d))
        rq[d] = rq[src]; rsb[d] = rsb[src]
    Q = np.concatenate((Q, rq))[:n_lines]; S = np.concatenate((S, rsb))[:n_lines]; SC = np.concatenate((SC, rsc))[:n_lines]
    n = len(Q)
    perm = rng.permutation(n)
    Q, S, SC = Q[perm], S[perm], SC[perm]
    qiso = np.ones(n, dtype=np.uint8); siso = np.ones(n, dtype=np.uint8)
    m = rng.random(n) < p_iso; qiso[m] = rng.integers(2, 4, size=m.sum())
    m = rng.random(n) < p_iso; siso[m] = rng.integers(2, 4, size=m.sum())
    m = rng.random(n) < p_absent; Q = Q.copy(); Q[m] = Gq + rng.integers(1, 1000, size=m.sum())
    m = rng.random(n) < p_absent; S = S.copy(); S[m] = Gs + rng.integers(1, 1000, size=m.sum())
    pct = rng.integers(2500, 10001, size=n)
    ints7 = np.empty((n, 7), dtype=np.int32)
    ints7[:, 0] = rng.integers(30, 1500, size=n)
    ints7[:, 1] = rng.integers(0, 300, size=n)
    ints7[:, 2] = rng.integers(0, 20, size=n)
    qa = rng.integers(1, 3000, size=n); qb = qa + ints7[:, 0]
    sa = rng.integers(1, 3000, size=n); sb = sa + ints7[:, 0]
    rev = rng.random(n) < 0.3
    ints7[:, 3] = qa; ints7[:, 4] = qb
    ints7[:, 5] = np.where(rev, sb, sa); ints7[:, 6] = np.where(rev, sa, sb)
    return (Q.astype(np.int32), S.astype(np.int32), qiso, siso, SC.astype(np.int32), pct.astype(np.int32), ints7)


def format_hits(gq, gs, arrays, ediv=8.0, score_style=0, nthreads=None, header=True):
    """-> bytes of the .last text (with '#' header lines)."""
    Q, S, qiso, siso, SC, pct, ints7 = [np.ascontiguousarray(a) for a in arrays]
    n = len(Q)
    cap = n * 160 + 1024
    buf = np.empty(cap, dtype=np.uint8)
    nthreads = nthreads or min(16, os.cpu_count() or 1)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    w = _lib().synth_format_lines(ctypes.c_int64(n), p(Q), p(S), p(qiso), p(siso), p(SC), p(pct), p(ints7),
                                  (gq.prefix + "g").encode(), (gs.prefix + "g").encode(), int(gq.digits),
                                  ctypes.c_double(ediv), int(score_style), p(buf), ctypes.c_int64(cap), int(nthreads))
    if w < 0:
        raise RuntimeError("synth buffer too small")
    body = buf[:w]
    if header:
        head = np.frombuffer(b"# LAST version synthetic\n# fields: query subject pctid ...\n#\n", dtype=np.uint8)
        return np.concatenate((head, body))
    return body


def make_pair(seed, Gq, Gs, n_lines, Kq=12, Ks=12, self_mode=False, score_style=0, **kw):
    """Convenience: -> (genome_q, genome_s, text uint8 array)."""
    rng = np.random.default_rng(seed)
    gq = Genome("A", Gq, Kq, rng)
    gs = gq if self_mode else Genome("B", Gs, Ks, rng)
    arrays = make_hits(gq, gs, n_lines, rng, self_mode=self_mode, **kw)
    return gq, gs, format_hits(gq, gs, arrays, score_style=score_style)


def write_dataset(outdir, name_q, name_s, seed, Gq, Gs, n_lines, **kw):
    """Write <q>.bed, <s>.bed, <q>.<s>.last into outdir; returns the three paths."""
    os.makedirs(outdir, exist_ok=True)
    gq, gs, text = make_pair(seed, Gq, Gs, n_lines, **kw)
    rng = np.random.default_rng(seed + 1000003)
    qbed = os.path.join(outdir, name_q + ".bed")
    sbed = os.path.join(outdir, name_s + ".bed")
    gq.write_bed(qbed, rng)
    if gs is not gq:
        gs.write_bed(sbed, rng)
    last = os.path.join(outdir, "%s.%s.last" % (name_q, name_s))
    text.tofile(last)
    return qbed, sbed, last
