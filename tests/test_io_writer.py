"""The multi-member gzip writer behind <species>.snps.gz / <species>.genes.gz (midas_b200/run/io.py): whatever the
zlib level, the member layout and the thread count, `gzip.open()` - what the reference's readers and merge_midas.py
use (midas/utility.py:194-206) - sees the bytes that were written, in order; an empty file is a valid gzip file."""
import gzip

import numpy as np
import pytest

from midas_b200.run import io as mio


@pytest.mark.parametrize("level", [None, 1, 6, 9])
@pytest.mark.parametrize("threads", [1, 4])
def test_members_round_trip(tmp_path, level, threads):
    rng = np.random.default_rng(3)
    chunks = [bytes(rng.integers(32, 127, size=int(n), dtype=np.uint8)) for n in rng.integers(0, 200_000, size=23)]
    p = str(tmp_path / "x.gz")
    w = mio.GzipMembersWriter(p, level=level, threads=threads)
    for i, c in enumerate(chunks):
        if i % 3 == 0:
            w.write(c)
        else:
            w.write_deferred(lambda c=c: c)
    w.close()
    with gzip.open(p, "rb") as fh:
        assert fh.read() == b"".join(chunks)
    with mio.iopen(p) as fh:                               # the reference-style text reader
        assert fh.read() == b"".join(chunks).decode()


def test_empty_file_is_valid_gzip(tmp_path):
    p = str(tmp_path / "e.gz")
    mio.GzipMembersWriter(p, threads=2).close()
    with gzip.open(p, "rb") as fh:
        assert fh.read() == b""


def test_default_level_is_the_fast_one():
    assert mio.DEFAULT_GZIP_LEVEL == 1
