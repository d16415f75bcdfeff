"""`.anchors` files on the host: block reader/writer and the accept/correct pass of liftover.

Behaviour follows AnchorFile (src/jcvi/compara/base.py:9-27,52-98) and the block splitter it is
built on, read_block(handle, "#") (src/jcvi/formats/base.py:610-639), including the splitter's
corner cases (tests/test_host_logic.py::test_read_block_quirks):

  * a line whose stripped text starts with '#' is a header; text before the first header is ignored;
  * in a run of consecutive headers every header but the last opens an EMPTY block;
  * a header at the very end of the file is an error in the reference (StopIteration escaping a
    generator -> RuntimeError) and is one here;
  * a file without any header is a single block holding every line (stripped).

This is O(anchors) host work by design; the hit table never comes through here.
"""
import sys

from ..apps_base import logger


def _is_header(text):
    return text.strip()[:1] == "#"


def split_blocks(lines, sep_test=_is_header):
    """-> list of blocks, each a list of stripped lines (see module docstring)."""
    flags = [sep_test(x) for x in lines]
    if True not in flags:
        return [[x.strip() for x in lines]]
    blocks = []
    i, n = flags.index(True), len(lines)
    while i < n:
        run_end = i
        while run_end < n and flags[run_end]:
            run_end += 1
        blocks.extend([] for _ in range(run_end - i - 1))
        if run_end == n:
            raise RuntimeError("anchors file ends with a '#' line (generator raised StopIteration)")
        body_end = run_end
        while body_end < n and not flags[body_end]:
            body_end += 1
        blocks.append([x.strip() for x in lines[run_end:body_end]])
        i = body_end
    return blocks


def read_block(handle, signal):
    """Compatibility shim with the reference's generator signature: yields (header, lines)."""
    lines = list(handle)
    k = len(signal)
    test = lambda row: row.strip()[:k] == signal
    if not any(test(x) for x in lines):
        yield None, [x.strip() for x in lines]
        return
    for blk in split_blocks(lines, test):
        yield signal, blk


class AnchorFile(object):
    """blocks: list of blocks; a block is a list of rows; a row is the whitespace-split line."""

    def __init__(self, filename, minsize=0):
        self.filename = filename
        self.blocks = list(self.iter_blocks(minsize=minsize))

    @classmethod
    def from_blocks(cls, filename, blocks):
        """The AnchorFile of a file whose blocks the caller already holds (rows = [a, b, score]); what parsing
        `filename` would give, without reading it again."""
        self = cls.__new__(cls)
        self.filename = filename
        self.blocks = [list(b) for b in blocks] if blocks else [[]]
        return self

    def _raw_blocks(self):
        with open(self.filename) as fp:
            return split_blocks(fp.readlines())

    def iter_blocks(self, minsize=0):
        for body in self._raw_blocks():
            rows = [text.split() for text in body]
            if len(rows) >= minsize:
                yield rows

    def iter_pairs(self, minsize=0):
        for block_id, rows in enumerate(self.iter_blocks(minsize=minsize)):
            for row in rows:
                a, b = row[:2]
                yield a, b, block_id

    def filter_blocks(self, accepted):
        """Keep the anchors whose (a, b) is a key of `accepted` (pairs present in the raw table);
        a score that is neither accepted[(a, b)] nor that value + "L" is replaced by it; blocks
        left empty disappear (compara/base.py:52-83)."""
        kept_blocks = []
        dropped = fixed = emptied = 0
        for block in self.blocks:
            kept = []
            for a, b, score in block:
                raw = accepted.get((a, b))
                if raw is None:
                    dropped += 1
                    continue
                if score not in (raw, raw + "L"):
                    score, fixed = raw, fixed + 1
                kept.append((a, b, score))
            if kept:
                kept_blocks.append(kept)
            else:
                emptied += 1
        logger.debug("Removed %d existing anchors", dropped)
        if emptied:
            logger.debug("Removed %d empty blocks", emptied)
        logger.debug("Corrected scores for %d anchors", fixed)
        self.blocks = kept_blocks

    def print_to_file(self, filename="stdout"):
        chunks = []
        for block in self.blocks:
            chunks.append("###\n")
            chunks.extend("%s\t%s\t%s\n" % (a, b, score) for a, b, score in block)
        text = "".join(chunks)
        if filename == "stdout":
            sys.stdout.write(text)
        else:
            with open(filename, "w") as fw:
                fw.write(text)
        logger.debug("Anchors written to `%s`", filename)

    @property
    def is_empty(self):
        return not self.blocks or not self.blocks[0]
