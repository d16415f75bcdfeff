"""AnchorFile: `.anchors` reader/writer and the accept/correct pass of liftover.

Mirrors src/jcvi/compara/base.py:9-27,52-98 and read_block (src/jcvi/formats/base.py:610-639).
Host-side by design: it touches anchors (10^4..10^6 lines), never the hit table.
"""
from itertools import groupby

from ..apps_base import logger


def read_block(handle, signal):
    """formats/base.py:610-639 -- blocks split at lines whose stripped text starts with `signal`."""
    signal_len = len(signal)
    it = (x[1] for x in groupby(handle, key=lambda row: row.strip()[:signal_len] == signal))
    found_signal = False
    for header in it:
        header = list(header)
        for h in header[:-1]:
            h = h.strip()
            if h[:signal_len] != signal:
                continue
            yield h, []
        header = header[-1].strip()
        if header[:signal_len] != signal:
            continue
        found_signal = True
        seq = list(s.strip() for s in next(it))
        yield header, seq
    if not found_signal:
        handle.seek(0)
        seq = list(s.strip() for s in handle)
        yield None, seq


class AnchorFile(object):
    def __init__(self, filename, minsize=0):
        self.filename = filename
        self.blocks = list(self.iter_blocks(minsize=minsize))

    def iter_blocks(self, minsize=0):
        fp = open(self.filename)
        for _, lines in read_block(fp, "#"):
            lines = [x.split() for x in lines]
            if len(lines) >= minsize:
                yield lines

    def iter_pairs(self, minsize=0):
        block_id = -1
        for rows in self.iter_blocks(minsize=minsize):
            block_id += 1
            for row in rows:
                a, b = row[:2]
                yield a, b, block_id

    def filter_blocks(self, accepted):
        """compara/base.py:52-83; `accepted` maps (a, b) -> str(int(score)) of the raw hit."""
        new_blocks = []
        nremoved = ncorrected = nblocks_removed = 0
        for block in self.blocks:
            new_block = []
            for line in block:
                a, b, score = line
                pair = (a, b)
                if pair not in accepted:
                    nremoved += 1
                    continue
                av = accepted[pair]
                if score != av and score != av + "L":
                    score = av
                    ncorrected += 1
                new_block.append((a, b, score))
            if new_block:
                new_blocks.append(new_block)
            else:
                nblocks_removed += 1
        logger.debug("Removed %d existing anchors", nremoved)
        if nblocks_removed:
            logger.debug("Removed %d empty blocks", nblocks_removed)
        logger.debug("Corrected scores for %d anchors", ncorrected)
        self.blocks = new_blocks

    def print_to_file(self, filename="stdout"):
        import sys
        fw = sys.stdout if filename == "stdout" else open(filename, "w")
        out = []
        for block in self.blocks:
            out.append("###\n")
            for line in block:
                a, b, score = line
                out.append("\t".join((a, b, score)) + "\n")
        fw.write("".join(out))
        if fw is not sys.stdout:
            fw.close()
        logger.debug("Anchors written to `%s`", filename)

    @property
    def is_empty(self):
        blocks = self.blocks
        return not blocks or not blocks[0]
