"""Parity oracle (TEST INFRASTRUCTURE ONLY -- see jcvi_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (jcvi_b200) never does.
"""
from .oracle import (  # noqa: F401
    build,
    blastfilter,
    scan,
    liftover,
    cscore,
    blast_filter,
    mcscan,
    simple,
    synfind,
    kdtree,
    gene_name,
    natcmp,
    bed_order,
    blastline_str,
    OracleError,
    ZeroAnchors,
)
