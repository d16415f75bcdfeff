#!/usr/bin/env python
"""Cross-check of the CPU oracle against the LIVE reference (only where /root/reference exists: the build
container) on randomized small inputs beyond the committed golden vectors: `formats.blast filter`,
`formats.blast cscore` and `synteny mcscan` with tie-heavy hand-made blocks.  Run by tests/test_live_reference.py
in a subprocess (the reference package must not leak into the test process)."""
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, ROOT)
import refenv  # noqa: E402


def rd(p):
    with open(p, "rb") as fh:
        return fh.read()


def main():
    try:
        refenv.activate()
    except Exception as e:                       # no Cython / compiler here: nothing to compare with
        print("LIVE_REFERENCE_UNAVAILABLE %s" % e)
        sys.exit(77)
    from jcvi.formats import blast as ref_blast
    from jcvi.compara import synteny as ref_syn
    import oracle
    d = tempfile.mkdtemp()
    os.chdir(d)
    n_checked = 0
    for seed in range(6):
        rng = random.Random(seed)
        # ---- a small table with awkward numbers ----------------------------------------------------------
        rows = []
        for i in range(400):
            q = "g%03d.%d" % (rng.randint(0, 40), rng.randint(1, 2))
            s = q if rng.random() < 0.05 else "h%03d" % rng.randint(0, 40)
            pct = rng.choice(["95", "95.0", "94.999999", "%.2f" % rng.uniform(20, 100), "+99.5"])
            ev = rng.choice(["0", "0.01", "1e-2", "0.010000000000000002", "1e-400", "%.1e" % 10 ** rng.uniform(-50, 1)])
            sc = rng.choice(["%.1f" % rng.uniform(10, 3000), "%.3g" % rng.uniform(10, 3000), "5e2"])
            f = [q, s, pct] + [str(rng.choice([0, 99, 100, 101, rng.randint(0, 3000)])) for _ in range(7)] + [ev, sc]
            r = rng.random()
            if r < 0.03:
                f = f[:rng.randint(2, 11)]
            line = ("\t" if r > 0.06 else " ").join(f)
            if rng.random() < 0.03:
                line += " \t"
            rows.append(line)
            if rng.random() < 0.02:
                rows.append("# comment")
        name = "t%d.last" % seed
        with open(name, "w") as fw:
            fw.write("\n".join(rows) + "\n")
        for tag, argv, kw in (("a", [], {}),
                              ("b", ["--pctid=60", "--hitlen=100", "--score=400", "--evalue=1e-2", "--noself"],
                               dict(pctid=60.0, hitlen=100, score=400, evalue=1e-2, noself=True)),
                              ("c", ["--pctid=94.999999", "--inverse", "--evalue=0"], dict(pctid=94.999999, inverse=True, evalue=0.0))):
            got = ref_blast.filter([name, "-o", name + "." + tag] + argv)
            okw = dict(score=0, pctid=95.0, hitlen=100, evalue=0.01, noself=False, inverse=False)
            okw.update(kw)
            oracle.blast_filter(name, got + ".orc", **okw)
            assert rd(got) == rd(got + ".orc"), ("filter", seed, tag)
            n_checked += 1
        full = [r for r in rows if not r.startswith("#") and len(r.split()) >= 12]
        with open("c%d.last" % seed, "w") as fw:
            fw.write("\n".join(full) + "\n")
        for tag, argv, kw in (("rbh", [], {}), ("c5", ["--cutoff=0.5", "--pct"], dict(cutoff=0.5, pct=True))):
            out = "c%d.%s" % (seed, tag)
            ref_blast.cscore(["c%d.last" % seed, "-o", out] + argv)
            oracle.cscore("c%d.last" % seed, out + ".orc", **kw)
            assert rd(out) == rd(out + ".orc"), ("cscore", seed, tag)
            n_checked += 1
        # ---- mcscan: hand-made blocks with equal end points, equal scores, repeated genes, 'L' scores ----------
        G = 60
        with open("m%d.bed" % seed, "w") as fw:
            for i in range(G):
                fw.write("chr%d\t%d\t%d\tq%02d\n" % (1 + i // 30, 100 * i, 100 * i + 50, i))
        with open("m%d.anchors" % seed, "w") as fw:
            for b in range(rng.randint(4, 14)):
                fw.write("###\n")
                a = rng.randint(0, G - 8)
                ln = rng.choice([1, 2, 3, 5, 8, 21, 30])
                for k in range(rng.randint(1, 6)):
                    g = min(G - 1, a + rng.randint(0, ln))
                    sc = rng.choice([10, 10, 20, 300])
                    fw.write("q%02d\tz%d_%d\t%d%s\n" % (g, b, rng.randint(0, 3), sc, rng.choice(["", "L"])))
        for tag, argv, kw in (("d", ["--trackids"], dict(trackids=True)), ("n0", ["--Nm=0", "--iter=2"], dict(Nm=0, iter=2)),
                              ("n2", ["--Nm=2"], dict(Nm=2))):
            out = "m%d.%s.blocks" % (seed, tag)
            ref_syn.mcscan(["m%d.bed" % seed, "m%d.anchors" % seed, "-o", out] + argv)
            oracle.mcscan("m%d.bed" % seed, "m%d.anchors" % seed, out + ".orc", **kw)
            assert rd(out) == rd(out + ".orc"), ("mcscan", seed, tag)
            if kw.get("trackids"):
                assert rd(out + ".tracks") == rd(out + ".orc.tracks"), ("mcscan tracks", seed)
            n_checked += 1
    # ---- simple: random small blocks; rows of blocks whose rank covariance is exactly 0 are compared without the
    # orientation column (np.polyfit returns rounding noise of either sign there; the oracle says "+")
    os.chdir(d)
    flat_total = flat_diff = 0
    for seed in range(6):
        rng = random.Random(100 + seed)
        G = 60
        for nm in ("x", "y"):
            with open("%s%d.bed" % (nm, seed), "w") as fw:
                for i in range(G):
                    fw.write("%schr%d\t%d\t%d\t%s%02d\n" % (nm, 1 + i // 30, 100 * i, 100 * i + 50 + 7 * (i % 3), nm, i))
        flat = []
        with open("x%d.y%d.anchors" % (seed, seed), "w") as fw:
            for b in range(25):
                fw.write("###\n")
                lo = rng.choice([0, 30])
                n = rng.randint(1, 6)
                xs = [lo + rng.randint(0, 29) for _ in range(n)]
                ys = [lo + rng.randint(0, 29) for _ in range(n)]
                if rng.random() < 0.3:
                    ys = [ys[0]] * n                                 # a flat block
                for x, y in zip(xs, ys):
                    fw.write("x%02d\ty%02d\t%d\n" % (x, y, rng.randint(10, 500)))
                flat.append(n * sum(a * b for a, b in zip(xs, ys)) - sum(xs) * sum(ys) == 0 and n >= 2)
        anchors = "x%d.y%d.anchors" % (seed, seed)
        for tag, argv, kw in (("r", ["--rich"], dict(rich=True)), ("c", ["--coords"], dict(coords=True))):
            ref_syn.simple([anchors, "--qbed=x%d.bed" % seed, "--sbed=y%d.bed" % seed] + argv)
            ref_out = anchors.rsplit(".", 1)[0] + ".simple"
            oracle.simple(anchors, "x%d.bed" % seed, "y%d.bed" % seed, ref_out + ".orc", **kw)
            rl, ol = rd(ref_out).decode().split("\n"), rd(ref_out + ".orc").decode().split("\n")
            assert len(rl) == len(ol), ("simple", seed, tag)
            per = 2 if kw.get("coords") else 1
            for k, (a, b) in enumerate(zip(rl[1:], ol[1:])):
                blk = k // per
                if a == b or blk >= len(flat):
                    continue
                assert flat[blk], ("simple", seed, tag, a, b)
                col = 8 if kw.get("coords") else 5
                fa, fb = a.split("\t"), b.split("\t")
                assert fa[:col] + fa[col + 1:] == fb[:col] + fb[col + 1:], ("simple flat", a, b)
                flat_diff += 1
            flat_total += sum(flat)
            n_checked += 1
    print("simple: %d flat blocks compared without orientation, %d of them differ in it" % (flat_total, flat_diff))
    # ---- the path itself on fresh seeds (the committed goldens pin fixed ones) ----------------------------------
    from jcvi.compara import blastfilter as ref_bf
    from jcvi_b200 import synth
    for seed, self_mode in ((201, False), (202, False), (203, True)):
        sub = os.path.join(d, "p%d" % seed)
        os.makedirs(sub)
        os.chdir(sub)
        names = ("S", "S") if self_mode else ("A", "B")
        qbed, sbed, last = synth.write_dataset(sub, names[0], names[1], seed, 1500, 1500, 25000, Kq=4, Ks=4, self_mode=self_mode)
        last = os.path.basename(last)
        ref_bf.main([last, "--cscore=0.6", "--tandem_Nmax=8"])
        os.rename(last + ".filtered", "ref.filtered")
        oracle.blastfilter(last, os.path.basename(qbed), os.path.basename(sbed), out_path="orc.filtered", cscore=0.6, tandem_Nmax=8)
        assert rd("ref.filtered") == rd("orc.filtered"), ("blastfilter", seed)
        os.rename("ref.filtered", last + ".filtered")
        stem = last[:-len(".last")]
        ref_out = ref_syn.scan([last + ".filtered", stem + ".anchors", "--liftover=" + last, "--min_size=3", "--dist=15"])
        oracle.scan(last + ".filtered", "orc.anchors", os.path.basename(qbed), os.path.basename(sbed), dist=15, min_size=3)
        assert rd(stem + ".anchors") == rd("orc.anchors"), ("scan", seed)
        oracle.liftover(last, "orc.anchors", os.path.basename(qbed), os.path.basename(sbed), out_path="orc.lifted", dist=7)
        assert rd(ref_out) == rd("orc.lifted"), ("liftover", seed)
        n_checked += 3
    # ---- synfind (oracle first) on small pairs, a self comparison, two window sizes ----------------
    from jcvi.compara import synfind as ref_sf
    for seed, self_mode, argv, kw in ((301, False, [], {}), (302, False, ["--window=20", "--cutoff=0.2"], dict(window=20, cutoff=0.2)),
                                      (303, True, [], {}), (304, False, ["--window=60", "--cutoff=0.05"], dict(window=60, cutoff=0.05))):
        sub = os.path.join(d, "s%d" % seed)
        os.makedirs(sub)
        os.chdir(sub)
        names = ("S", "S") if self_mode else ("A", "B")
        qbed, sbed, last = synth.write_dataset(sub, names[0], names[1], seed, 900, 700, 14000, Kq=3, Ks=3, self_mode=self_mode)
        last, qbed, sbed = os.path.basename(last), os.path.basename(qbed), os.path.basename(sbed)
        ref_sf.main([last, "--qbed=" + qbed, "--sbed=" + sbed, "-o", "ref.synfind"] + argv)
        oracle.synfind(last, qbed, sbed, "orc.synfind", **kw)
        assert rd("ref.synfind") == rd("orc.synfind"), ("synfind", seed)
        assert len(rd("ref.synfind")) > 1000, ("synfind output suspiciously small", seed)
        n_checked += 1
        if seed == 301:                       # nothing maps onto the BEDs: the reference dies in transposed([])
            for fn in (lambda: ref_sf.main([last, "--qbed=" + qbed, "--sbed=" + sbed, "-o", "x", "--no_strip_names"]),
                       lambda: oracle.synfind(last, qbed, sbed, "y", strip_names=False)):
                try:
                    fn()
                    raise AssertionError("expected ValueError")
                except ValueError:
                    pass
    print("LIVE_REFERENCE_OK %d comparisons" % n_checked)


if __name__ == "__main__":
    main()
