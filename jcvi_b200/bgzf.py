"""Blocked gzip (BGZF, the container bgzip / htslib write) for the path's input tables: a valid multi-member `.gz` that
`gzip.open` -- hence the reference's must_open (src/jcvi/formats/base.py:473-550) -- reads like any other, whose members
(<= 64 KiB of text each) carry their compressed size in a 'BC' extra field, so that jx_file_upload can inflate them
independently on all host threads.  Writer only; it exists for tests and measurements (bgzip itself is not in this image)."""
import struct
import zlib

BLOCK = 0xFF00          # uncompressed bytes per member, as bgzip chooses
_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _member(chunk, level):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(chunk) + co.flush()
    bsize = 12 + 6 + len(body) + 8
    if bsize > 65536:                       # incompressible: store
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(body) + 8
    head = struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
    return head + body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))


def write_bgzf(data, path, level=6, threads=8):
    """data: bytes-like (or a uint8 numpy array) -> BGZF file at `path`; returns the compressed size."""
    from concurrent.futures import ThreadPoolExecutor
    mv = memoryview(data).cast("B")
    n = len(mv)
    total = 0
    with open(path, "wb") as fh, ThreadPoolExecutor(threads) as ex:       # zlib releases the GIL
        step = BLOCK * 256
        for a in range(0, n, step):
            parts = list(ex.map(lambda o: _member(bytes(mv[o:min(o + BLOCK, n, a + step)]), level),
                                range(a, min(a + step, n), BLOCK)))
            for p in parts:
                fh.write(p)
                total += len(p)
        fh.write(_EOF)
    return total + len(_EOF)
