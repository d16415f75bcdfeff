"""jcvi_b200 -- B200-native (sm_100a) MCscan anchor detection: a drop-in for
jcvi.compara.blastfilter -> jcvi.compara.synteny scan -> jcvi.compara.synteny liftover.

    from jcvi_b200.compara import blastfilter, synteny
    blastfilter.main(["A.B.last", "--cscore=0.7"])
    synteny.scan(["A.B.last.filtered", "A.B.anchors", "--liftover=A.B.last"])

`install()` rebinds the same names inside an importable `jcvi` so that callers such as
`jcvi.compara.catalog.ortholog` pick up the GPU path unchanged.
"""
__version__ = "0.1.0"


def install():
    """Rebind blastfilter.main / synteny.scan / synteny.liftover (and the names catalog imported
    with `from ... import`, catalog.py:21,23) inside the reference package."""
    import importlib
    from .compara import blastfilter as bf, synteny as syn

    jbf = importlib.import_module("jcvi.compara.blastfilter")
    jsyn = importlib.import_module("jcvi.compara.synteny")
    jbf.main = bf.main
    jbf.blastfilter_main = bf.blastfilter_main
    for name in ("scan", "liftover", "batch_scan", "synteny_scan", "synteny_liftover", "mcscan", "simple"):
        setattr(jsyn, name, getattr(syn, name))
    try:                                   # formats.blast cscore
        from .formats import blast as gblast
        jblast = importlib.import_module("jcvi.formats.blast")
        jblast.cscore = gblast.cscore
        jblast.filter = gblast.filter
    except ImportError:
        pass
    try:                                   # compara.synfind
        from .compara import synfind as gsf
        jsf = importlib.import_module("jcvi.compara.synfind")
        for name in ("main", "synfind", "batch_query"):
            setattr(jsf, name, getattr(gsf, name))
    except ImportError:
        pass
    try:
        jcat = importlib.import_module("jcvi.compara.catalog")
        if hasattr(jcat, "blastfilter_main"):
            jcat.blastfilter_main = bf.main
        for name in ("scan", "liftover", "mcscan"):
            if hasattr(jcat, name):
                setattr(jcat, name, getattr(syn, name))
        if hasattr(jcat, "cscore"):            # `from jcvi.formats.blast import cscore` (catalog.py:20)
            from .formats import blast as gblast
            jcat.cscore = gblast.cscore
        if hasattr(jcat, "blast_filter"):      # `from ..formats.blast import filter as blast_filter` (catalog.py:33)
            from .formats import blast as gblast
            jcat.blast_filter = gblast.filter
    except ImportError:
        pass
