"""CPU: libmidas_b200.so loads, exports every symbol include/midas_b200.h declares, and its compute
entry points fail loudly (MB_ECUDA) instead of falling back when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from midas_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "midas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/midas_b200.h but not exported"
    assert sorted(lib.EXPORTS) == names, "lib.EXPORTS is out of date with the header"


def test_version_and_struct_sizes():
    assert "sm_100a" in lib.version()
    assert lib.HDR_DTYPE.itemsize == 16 and lib.STATS_DTYPE.itemsize == 40 and lib.GENE_OUT_DTYPE.itemsize == 24
    assert C.sizeof(lib.CPackedReads) == 8 * 4 + 4 * 8 + 8 * 6


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.MidasB200Error) as e:
        lib.Context(0, use_torch_stream=False)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_layout_geometry():
    lay = lib.Layout([5000, 100, 4096, 1], [0, 0, 1, 2], 3)
    assert lay.gpad % lib.MB_TILE == 0 and lay.n_tiles == lay.gpad // lib.MB_TILE
    assert list(lay.species_goff[:3]) == [0, 8192, 12288] and lay.gpad == 16384
    assert list(lay.contig_goff) == [0, 5000, 8192, 12288]
    assert list(lay.species_len) == [5100, 4096, 1]
    with pytest.raises(lib.MidasB200Error):
        lib.Layout([10, 0], [0, 0], 1)            # zero-length contig: pysam raises ValueError
    with pytest.raises(lib.MidasB200Error):
        lib.Layout([10], [3], 1)


def test_format_rows_matches_python_str():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2**32 - 1, (4, 50), dtype=np.uint64).astype(np.uint32)
    c[:, :5] = 0
    bases = np.frombuffer(b"ACGTNRYKM-" * 5, np.uint8)
    got = lib.format_snps_rows("contig|1", bases, 123, c).decode()
    exp = "".join("\t".join(str(x) for x in ["contig|1", 124 + i, chr(bases[i]), int(c[:, i].astype(np.uint64).sum()),
                                             c[0, i], c[1, i], c[2, i], c[3, i]]) + "\n" for i in range(50))
    assert got == exp
