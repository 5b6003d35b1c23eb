"""ctypes binding of include/midas_b200.h (libmidas_b200.so).

There is no CPU fallback: if the shared library is missing, or a compute entry point is
called without a B200, this module raises (MidasB200Error / OSError) instead of degrading.
Device buffers handed to the *_dev entry points are torch CUDA tensors (torch is plumbing
for device memory and streams only).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .soa import CFilters, Filters, PlainReads, RefSet

_HERE = os.path.dirname(os.path.abspath(__file__))
# MB_LIB_PATH: another build of the same library (the phase-clock build of `make prof`), for scripts/phase_prof.py
LIB_PATH = os.environ.get("MB_LIB_PATH") or os.path.join(_HERE, "libmidas_b200.so")

MB_TILE = 4096
MB_BLOCK = 32
STATUS = {0: "MB_OK", -1: "MB_EINVAL", -2: "MB_ECUDA", -3: "MB_ENOMEM", -4: "MB_EOVERFLOW",
          -5: "MB_EMALFORMED", -6: "MB_EUNSORTED"}


class MidasB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class CRef(C.Structure):
    _fields_ = [("n_contigs", C.c_uint32), ("n_species", C.c_uint32),
                ("contig_len", C.c_void_p), ("contig_species", C.c_void_p)]


class CPlainReads(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("ref_id", C.c_void_p), ("pos", C.c_void_p), ("mapq", C.c_void_p),
                ("nm", C.c_void_p), ("l_qseq", C.c_void_p), ("cigar_off", C.c_void_p), ("cigar", C.c_void_p),
                ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("qual", C.c_void_p),
                ("seq_nibbles", C.c_uint32), ("reserved", C.c_uint32)]

    @staticmethod
    def of(reads: PlainReads) -> "CPlainReads":
        """View of host numpy arrays (the caller keeps `reads` alive for the duration of the call)."""
        return CPlainReads(reads.n_reads, _ptr(reads.ref_id), _ptr(reads.pos), _ptr(reads.mapq), _ptr(reads.nm),
                           _ptr(reads.l_qseq), _ptr(reads.cigar_off), _ptr(reads.cigar), _ptr(reads.seq_off),
                           _ptr(reads.seq), _ptr(reads.qual), int(bool(reads.seq_nibbles)), 0)


class CMergeParams(C.Structure):
    _fields_ = [("snp_freq", C.c_double), ("site_ratio", C.c_double), ("site_depth", C.c_int32),
                ("reserved", C.c_int32)]


class CPackedReads(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_vec", C.c_uint64), ("n_cigar_words", C.c_uint64),
                ("total_bases", C.c_uint64), ("max_ref_span", C.c_uint32), ("sorted", C.c_uint32),
                ("row_words", C.c_uint32), ("n_long", C.c_uint32), ("q_planes", C.c_uint32),
                ("n_planes", C.c_uint32), ("max_units", C.c_uint32), ("reserved", C.c_uint32),
                ("long_idx", C.c_void_p),
                ("hdr", C.c_void_p), ("vec_off", C.c_void_p), ("cplx", C.c_void_p), ("cigar", C.c_void_p),
                ("planes", C.c_void_p)]


HDR_DTYPE = np.dtype([("loc", "<u4"), ("l_qseq", "<u2"), ("alen", "<u2"), ("q_start", "<u2"), ("nm", "<u2"),
                      ("ref_span", "<u2"), ("mapq", "u1"), ("flags", "u1")])
STATS_DTYPE = np.dtype([("genome_length", "<u8"), ("covered_bases", "<u8"), ("total_depth", "<u8"),
                        ("aligned_reads", "<u8"), ("mapped_reads", "<u8")])
GENE_OUT_DTYPE = np.dtype([("aligned_reads", "<u4"), ("mapped_reads", "<u4"), ("mapped_bp", "<u8"),
                           ("depth", "<f8")])

# every symbol include/midas_b200.h declares (tests/test_abi.py checks the .so exports them all)
EXPORTS = [
    "mb_version", "mb_last_error",
    "mb_layout_create", "mb_layout_create_genes", "mb_layout_destroy", "mb_layout_gpad", "mb_layout_n_tiles", "mb_layout_contig_goff",
    "mb_layout_species_goff", "mb_layout_species_len",
    "mb_pack_reads", "mb_packed_free", "mb_host_alloc", "mb_host_free",
    "mb_ctx_create", "mb_ctx_destroy", "mb_ctx_set_stream", "mb_ctx_sync", "mb_ctx_device",
    "mb_ctx_launch_count", "mb_ctx_last_kernel_ms", "mb_ctx_timing_reset", "mb_ctx_timing_read",
    "mb_reads_upload", "mb_reads_free", "mb_reads_bytes", "mb_reads_total_bases", "mb_reads_n",
    "mb_reads_pack", "mb_reads_pack_dev", "mb_reads_download", "mb_ctx_last_pack_ms",
    "mb_snps_pileup_dev", "mb_snps_pileup", "mb_snps_pileup_plain", "mb_ctx_overflowed", "mb_ctx_set_stats_mirror",
    "mb_genes_coverage_dev", "mb_genes_coverage", "mb_genes_coverage_plain",
    "mb_format_snps_bound", "mb_format_snps_rows",
    "mb_merge_sites_dev", "mb_format_merge_bound", "mb_format_merge_freq_rows", "mb_format_merge_depth_rows",
    "mb_bam_open", "mb_bam_close", "mb_bam_n_ref", "mb_bam_ref_name", "mb_bam_ref_len", "mb_bam_reads", "mb_bam_flags",
    "mb_bam_stream_open", "mb_bam_stream_open_dev", "mb_bam_stream_close", "mb_bam_stream_n_ref", "mb_bam_stream_ref_name", "mb_bam_stream_ref_len",
    "mb_bam_stream_bytes", "mb_snps_pileup_bam", "mb_genes_coverage_bam",
]

_lib = None


def load():
    """Load libmidas_b200.so or raise — the product path never silently degrades."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      f"or `make -C midas_b200/csrc`; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.mb_version.restype = C.c_char_p
    L.mb_last_error.restype = C.c_char_p
    L.mb_layout_gpad.restype = C.c_uint64
    L.mb_layout_gpad.argtypes = [C.c_void_p]
    L.mb_layout_n_tiles.restype = C.c_uint32
    L.mb_layout_n_tiles.argtypes = [C.c_void_p]
    for name in ("mb_layout_contig_goff", "mb_layout_species_goff", "mb_layout_species_len"):
        getattr(L, name).restype = C.c_void_p
        getattr(L, name).argtypes = [C.c_void_p]
    L.mb_layout_create.argtypes = [C.POINTER(CRef), C.POINTER(C.c_void_p)]
    L.mb_layout_create_genes.argtypes = [C.c_uint32, C.POINTER(C.c_void_p)]
    L.mb_layout_destroy.argtypes = [C.c_void_p]
    L.mb_layout_destroy.restype = None
    L.mb_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.mb_host_free.argtypes = [C.c_void_p]
    L.mb_host_free.restype = None
    for name in ("mb_reads_total_bases", "mb_reads_n"):
        getattr(L, name).argtypes = [C.c_void_p]
        getattr(L, name).restype = C.c_uint64
    L.mb_reads_pack.argtypes = [C.c_void_p, C.POINTER(CPlainReads), C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.mb_reads_pack_dev.argtypes = [C.c_void_p, C.POINTER(CPlainReads), C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.mb_reads_download.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(CPackedReads))]
    L.mb_ctx_last_pack_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.mb_snps_pileup_plain.argtypes = [C.c_void_p, C.POINTER(CPlainReads), C.c_void_p, C.POINTER(CFilters),
                                       C.c_void_p, C.c_void_p]
    L.mb_ctx_overflowed.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.mb_ctx_set_stats_mirror.argtypes = [C.c_void_p, C.c_void_p]
    L.mb_genes_coverage_plain.argtypes = [C.c_void_p, C.POINTER(CPlainReads), C.c_void_p, C.c_uint32,
                                          C.POINTER(CFilters), C.c_int, C.c_void_p]
    L.mb_pack_reads.argtypes = [C.POINTER(CPlainReads), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(C.POINTER(CPackedReads))]
    L.mb_packed_free.argtypes = [C.POINTER(CPackedReads)]
    L.mb_packed_free.restype = None
    L.mb_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.mb_ctx_destroy.argtypes = [C.c_void_p]
    L.mb_ctx_destroy.restype = None
    L.mb_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.mb_ctx_sync.argtypes = [C.c_void_p]
    L.mb_ctx_device.argtypes = [C.c_void_p]
    L.mb_ctx_launch_count.argtypes = [C.c_void_p]
    L.mb_ctx_launch_count.restype = C.c_uint64
    L.mb_ctx_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.mb_ctx_timing_reset.argtypes = [C.c_void_p]
    L.mb_ctx_timing_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.mb_reads_upload.argtypes = [C.c_void_p, C.POINTER(CPackedReads), C.POINTER(C.c_void_p)]
    L.mb_reads_free.argtypes = [C.c_void_p, C.c_void_p]
    L.mb_reads_free.restype = None
    L.mb_reads_bytes.argtypes = [C.c_void_p]
    L.mb_reads_bytes.restype = C.c_uint64
    L.mb_snps_pileup_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CFilters), C.c_void_p, C.c_void_p]
    L.mb_snps_pileup.argtypes = [C.c_void_p, C.POINTER(CPackedReads), C.c_void_p, C.POINTER(CFilters),
                                 C.c_void_p, C.c_void_p]
    L.mb_genes_coverage_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(CFilters),
                                        C.c_int, C.c_void_p]
    L.mb_genes_coverage.argtypes = [C.c_void_p, C.POINTER(CPackedReads), C.c_void_p, C.c_uint32,
                                    C.POINTER(CFilters), C.c_int, C.c_void_p]
    L.mb_format_snps_bound.restype = C.c_uint64
    L.mb_format_snps_bound.argtypes = [C.c_uint64, C.c_uint64]
    L.mb_format_merge_bound.restype = C.c_uint64
    L.mb_format_merge_bound.argtypes = [C.c_uint64, C.c_uint32]
    for name in ("mb_format_merge_freq_rows", "mb_format_merge_depth_rows"):
        getattr(L, name).restype = C.c_int64
        getattr(L, name).argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p,
                                     C.c_uint64]
    L.mb_format_snps_rows.restype = C.c_int64
    L.mb_format_snps_rows.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.mb_merge_sites_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p,
                                     C.POINTER(CMergeParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mb_bam_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.mb_bam_close.argtypes = [C.c_void_p]
    L.mb_bam_close.restype = None
    L.mb_bam_n_ref.argtypes = [C.c_void_p]
    L.mb_bam_n_ref.restype = C.c_uint32
    L.mb_bam_ref_name.argtypes = [C.c_void_p, C.c_uint32]
    L.mb_bam_ref_name.restype = C.c_char_p
    L.mb_bam_ref_len.argtypes = [C.c_void_p, C.c_uint32]
    L.mb_bam_ref_len.restype = C.c_uint32
    L.mb_bam_reads.argtypes = [C.c_void_p]
    L.mb_bam_reads.restype = C.POINTER(CPlainReads)
    L.mb_bam_flags.argtypes = [C.c_void_p]
    L.mb_bam_flags.restype = C.c_void_p
    L.mb_bam_stream_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.mb_bam_stream_open_dev.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.mb_bam_stream_close.argtypes = [C.c_void_p]
    L.mb_bam_stream_close.restype = None
    L.mb_bam_stream_n_ref.argtypes = [C.c_void_p]
    L.mb_bam_stream_n_ref.restype = C.c_uint32
    L.mb_bam_stream_ref_name.argtypes = [C.c_void_p, C.c_uint32]
    L.mb_bam_stream_ref_name.restype = C.c_char_p
    L.mb_bam_stream_ref_len.argtypes = [C.c_void_p, C.c_uint32]
    L.mb_bam_stream_ref_len.restype = C.c_uint32
    L.mb_bam_stream_bytes.argtypes = [C.c_void_p]
    L.mb_bam_stream_bytes.restype = C.c_uint64
    L.mb_genes_coverage_bam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(CFilters),
                                        C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    L.mb_snps_pileup_bam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CFilters), C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise MidasB200Error(rc, load().mb_last_error().decode(errors="replace"))


def version() -> str:
    return load().mb_version().decode()


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Layout:
    """Padded, species-major global coordinate space (mb_layout)."""

    def __init__(self, contig_len, contig_species, n_species: int):
        L = load()
        self.contig_len = np.ascontiguousarray(contig_len, np.uint32)
        self.contig_species = np.ascontiguousarray(contig_species, np.int32)
        self.n_contigs = int(self.contig_len.shape[0])
        self.n_species = int(n_species)
        cref = CRef(self.n_contigs, self.n_species, _ptr(self.contig_len), _ptr(self.contig_species))
        h = C.c_void_p()
        _check(L.mb_layout_create(C.byref(cref), C.byref(h)))
        self._h = h
        self.gpad = int(L.mb_layout_gpad(h))
        self.n_tiles = int(L.mb_layout_n_tiles(h))

        def arr(fn, n, dt):
            p = fn(h)
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), (n,)).copy()

        self.contig_goff = arr(L.mb_layout_contig_goff, self.n_contigs, np.dtype("<u4"))
        self.species_goff = arr(L.mb_layout_species_goff, self.n_species + 1, np.dtype("<u4"))
        self.species_len = arr(L.mb_layout_species_len, self.n_species, np.dtype("<u8"))

    @staticmethod
    def of(ref: RefSet) -> "Layout":
        return Layout(ref.contig_len, ref.contig_species, ref.n_species)

    @staticmethod
    def genes(n_genes: int) -> "Layout":
        """Genes mode needs no coordinates (loc = gene index): no 32-bit padded-genome limit."""
        self = Layout.__new__(Layout)
        h = C.c_void_p()
        _check(load().mb_layout_create_genes(int(n_genes), C.byref(h)))
        self._h, self.n_contigs, self.n_species, self.gpad, self.n_tiles = h, int(n_genes), 0, 0, 0
        self.contig_len = np.zeros(0, np.uint32)
        self.contig_species = np.zeros(0, np.int32)
        self.contig_goff = np.zeros(0, np.uint32)
        self.species_goff = np.zeros(1, np.uint32)
        self.species_len = np.zeros(0, np.uint64)
        return self

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mb_layout_destroy(h)

    def contig_counts(self, counts: np.ndarray, c: int) -> np.ndarray:
        """View [4][contig_len[c]] of a [4][gpad] count array."""
        o = int(self.contig_goff[c])
        return counts[:, o:o + int(self.contig_len[c])]

    def unpad(self, counts: np.ndarray) -> np.ndarray:
        """[4][gpad] -> [4][sum(contig_len)] in the given contig order (the oracle's layout)."""
        return np.concatenate([self.contig_counts(counts, c) for c in range(self.n_contigs)], axis=1) \
            if self.n_contigs else counts[:, :0]


class PackedReads:
    """Library-owned packed struct-of-arrays (mb_packed_reads), host side."""

    def __init__(self, reads: PlainReads, layout: Layout, mode: str = "snps", sort: bool = False,
                 pinned: bool = False, threads: int = 0):
        L = load()
        self._keep = reads  # keep the source arrays alive during the call
        cin = CPlainReads.of(reads)
        out = C.POINTER(CPackedReads)()
        _check(L.mb_pack_reads(C.byref(cin), layout._h, {"snps": 0, "genes": 1}[mode], int(sort), int(pinned),
                               int(threads), C.byref(out)))
        self._p = out
        self._keep = None
        self.mode = mode

    @staticmethod
    def _from_pointer(p, mode: str) -> "PackedReads":
        self = PackedReads.__new__(PackedReads)
        self._p, self._keep, self.mode = p, None, mode
        return self

    def __del__(self):
        p, self._p = getattr(self, "_p", None), None
        if p and _lib is not None:
            _lib.mb_packed_free(p)

    @property
    def c(self) -> CPackedReads:
        return self._p.contents

    n_reads = property(lambda s: int(s.c.n_reads))
    n_vec = property(lambda s: int(s.c.n_vec))
    q_planes = property(lambda s: int(s.c.q_planes))
    n_planes = property(lambda s: int(s.c.n_planes))
    total_bases = property(lambda s: int(s.c.total_bases))
    max_ref_span = property(lambda s: int(s.c.max_ref_span))
    row_words = property(lambda s: int(s.c.row_words))
    n_long = property(lambda s: int(s.c.n_long))
    max_units = property(lambda s: int(s.c.max_units))

    @property
    def nbytes(self) -> int:
        c = self.c
        return int(16 * c.n_reads + 4 * (c.n_reads + 1) + 4 * c.n_reads + 4 * c.n_cigar_words + 32 * c.n_vec
                   + 4 * c.n_long)

    def _view(self, ptr, n, dt):
        if n == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=n)

    # numpy views of the packed arrays (valid while this object lives)
    hdr = property(lambda s: s._view(s.c.hdr, s.n_reads, HDR_DTYPE))
    vec_off = property(lambda s: s._view(s.c.vec_off, s.n_reads + 1, np.dtype("<u4")))
    cplx = property(lambda s: s._view(s.c.cplx, s.n_reads, np.dtype("<u4")))
    cigar = property(lambda s: s._view(s.c.cigar, int(s.c.n_cigar_words), np.dtype("<u4")))
    planes = property(lambda s: s._view(s.c.planes, 8 * s.n_vec, np.dtype("<u4")))
    long_idx = property(lambda s: s._view(s.c.long_idx, s.n_long, np.dtype("<u4")))


class DevReads:
    """A packed batch resident in HBM (mb_dev_reads): uploaded from a host PackedReads, or packed on the
    GPU from plain records (Context.pack / Context.pack_dev)."""

    def __init__(self, ctx: "Context", packed: PackedReads | None = None, handle=None, mode: str | None = None):
        if handle is None:
            handle = C.c_void_p()
            _check(load().mb_reads_upload(ctx._h, packed._p, C.byref(handle)))
            mode = packed.mode
        self._h, self._ctx, self.mode = handle, ctx, mode
        L = load()
        self.n_reads, self.total_bases = int(L.mb_reads_n(handle)), int(L.mb_reads_total_bases(handle))
        self.nbytes = int(L.mb_reads_bytes(handle))

    def download(self) -> PackedReads:
        """Host copy of the resident arrays (to check the device packer against mb_pack_reads)."""
        out = C.POINTER(CPackedReads)()
        _check(load().mb_reads_download(self._ctx._h, self._h, C.byref(out)))
        return PackedReads._from_pointer(out, self.mode)

    def free(self):
        h, self._h = self._h, None
        if h and _lib is not None and self._ctx._h:
            _lib.mb_reads_free(self._ctx._h, h)

    __del__ = free


class Context:
    """One CUDA context/stream on one B200 (mb_ctx). Not re-entrant."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        L = load()
        h = C.c_void_p()
        _check(L.mb_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        if use_torch_stream:
            import torch
            # torch's default stream has handle 0, which this ABI reads as "context-owned stream":
            # name it explicitly as cudaStreamLegacy (0x1) so torch events bracket our kernels
            self.set_stream(torch.cuda.current_stream(device).cuda_stream or 1)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mb_ctx_destroy(h)

    __del__ = close

    def set_stream(self, cuda_stream: int | None):
        _check(load().mb_ctx_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        _check(load().mb_ctx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(load().mb_ctx_launch_count(self._h))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _check(load().mb_ctx_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def timing_reset(self):
        _check(load().mb_ctx_timing_reset(self._h))

    def timing_read(self):
        """(total_ms, n_calls) of the dominant kernel since timing_reset(); synchronises."""
        ms, n = C.c_double(), C.c_uint64()
        _check(load().mb_ctx_timing_read(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def upload(self, packed: PackedReads) -> DevReads:
        return DevReads(self, packed)

    def pack(self, reads: PlainReads, layout: Layout, mode: str = "snps") -> DevReads:
        """Decoded records in host memory -> packed batch resident in HBM; the re-layout runs on the GPU
        (mb_reads_pack). snps mode wants BAM coordinate order (MB_EUNSORTED otherwise)."""
        cin = CPlainReads.of(reads)
        h = C.c_void_p()
        _check(load().mb_reads_pack(self._h, C.byref(cin), layout._h, {"snps": 0, "genes": 1}[mode], C.byref(h)))
        return DevReads(self, handle=h, mode=mode)

    def pack_dev(self, t: dict, layout: Layout, mode: str = "snps", seq_nibbles: bool = False) -> DevReads:
        """Same, from records already on the device: `t` maps the mb_plain_reads field names to torch CUDA
        tensors (ref_id/pos/nm/l_qseq int32, mapq/seq/qual uint8, cigar_off/seq_off int64, cigar int32)."""
        n = int(t["ref_id"].shape[0])
        cin = CPlainReads(n, *[C.c_void_p(t[k].data_ptr()) for k in
                               ("ref_id", "pos", "mapq", "nm", "l_qseq", "cigar_off", "cigar", "seq_off", "seq", "qual")],
                          int(bool(seq_nibbles)), 0)
        h = C.c_void_p()
        _check(load().mb_reads_pack_dev(self._h, C.byref(cin), layout._h, {"snps": 0, "genes": 1}[mode], C.byref(h)))
        return DevReads(self, handle=h, mode=mode)

    def last_pack_ms(self) -> float:
        ms = C.c_float()
        _check(load().mb_ctx_last_pack_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def overflowed(self) -> bool:
        flag = C.c_int()
        _check(load().mb_ctx_overflowed(self._h, C.byref(flag)))
        return bool(flag.value)

    def set_stats_mirror(self, ptr: int | None):
        """Device pointer in a peer GPU's memory that also receives the stats rows (None: off)."""
        _check(load().mb_ctx_set_stats_mirror(self._h, C.c_void_p(ptr or 0)))

    # ---- snps
    def snps_pileup_dev(self, reads: DevReads, layout: Layout, filters: Filters, counts, stats):
        """counts: torch int32/uint32 CUDA tensor [4, gpad]; stats: torch int64 CUDA tensor [n_species, 5]."""
        f = CFilters.of(filters)
        _check(load().mb_snps_pileup_dev(self._h, reads._h, layout._h, C.byref(f),
                                         C.c_void_p(counts.data_ptr()), C.c_void_p(stats.data_ptr())))

    def snps_pileup(self, packed: PackedReads, layout: Layout, filters: Filters,
                    counts_out: np.ndarray | None = None):
        """Host buffers in, host buffers out. Returns (counts uint32 [4][gpad], stats STATS_DTYPE [n_species])."""
        counts = counts_out if counts_out is not None else np.empty((4, layout.gpad), np.uint32)
        assert counts.dtype == np.uint32 and counts.shape == (4, layout.gpad) and counts.flags.c_contiguous
        stats = np.zeros(layout.n_species, STATS_DTYPE)
        f = CFilters.of(filters)
        _check(load().mb_snps_pileup(self._h, packed._p, layout._h, C.byref(f), _ptr(counts), _ptr(stats)))
        return counts, stats

    def snps_pileup_plain(self, reads: PlainReads, layout: Layout, filters: Filters,
                          counts_out: np.ndarray | None = None):
        """Decoded records in host memory in, host buffers out: H2D + device re-layout + pileup + D2H
        (mb_snps_pileup_plain). Returns (counts uint32 [4][gpad], stats STATS_DTYPE [n_species])."""
        counts = counts_out if counts_out is not None else np.empty((4, layout.gpad), np.uint32)
        assert counts.dtype == np.uint32 and counts.shape == (4, layout.gpad) and counts.flags.c_contiguous
        stats = np.zeros(layout.n_species, STATS_DTYPE)
        f = CFilters.of(filters)
        cin = CPlainReads.of(reads)
        _check(load().mb_snps_pileup_plain(self._h, C.byref(cin), layout._h, C.byref(f), _ptr(counts), _ptr(stats)))
        return counts, stats

    def snps_pileup_bam(self, stream: "BamStream", tid_map, layout: Layout, filters: Filters,
                        counts_out: np.ndarray | None = None):
        """From the inflated BAM record stream (mb_snps_pileup_bam): the bytes go to the GPU as they are, the device
        parses the records, re-lays them and counts. tid_map: int32 [n_ref] BAM reference -> layout contig (-1 = drop).
        Returns (counts, stats, n_reads, n_bases)."""
        counts = counts_out if counts_out is not None else np.empty((4, layout.gpad), np.uint32)
        assert counts.dtype == np.uint32 and counts.shape == (4, layout.gpad) and counts.flags.c_contiguous
        stats = np.zeros(layout.n_species, STATS_DTYPE)
        tid_map = np.ascontiguousarray(tid_map, np.int32)
        assert len(tid_map) == len(stream.names)
        f = CFilters.of(filters)
        nr, nb = C.c_uint64(), C.c_uint64()
        _check(load().mb_snps_pileup_bam(self._h, stream._h, _ptr(tid_map), layout._h, C.byref(f), _ptr(counts),
                                         _ptr(stats), C.byref(nr), C.byref(nb)))
        return counts, stats, int(nr.value), int(nb.value)

    # ---- merge (snps)
    def merge_sites(self, counts, mean_depth, snp_freq=0.01, site_depth=1, site_ratio=2.0):
        """Pooled allele call across samples (reference midas/merge/snps.py GenomicSite).
        counts: torch CUDA int32/uint32 [S, 4, G] (per-sample pileup outputs stacked); mean_depth: [S] float64.
        Returns dict of torch CUDA tensors: pooled [4, G] int32 (bits of u32), call [G] uint8,
        count_samples [G] int32, sample_depth [S, G] int32, sample_maf [S, G] float64."""
        import torch
        S, four, G = counts.shape
        assert four == 4 and counts.is_contiguous() and counts.dtype in (torch.int32, torch.uint32)
        dev = counts.device
        md = torch.as_tensor(mean_depth, dtype=torch.float64, device=dev).contiguous()
        if bool((md <= 0).any()):
            raise MidasB200Error(-1, "mean_depth must be > 0 for every sample (reference: ZeroDivisionError snps.py:99)")
        out = dict(pooled=torch.empty((4, G), dtype=torch.int32, device=dev),
                   call=torch.empty(G, dtype=torch.uint8, device=dev),
                   count_samples=torch.empty(G, dtype=torch.int32, device=dev),
                   sample_depth=torch.empty((S, G), dtype=torch.int32, device=dev),
                   sample_maf=torch.empty((S, G), dtype=torch.float64, device=dev))
        mp = CMergeParams(float(snp_freq), float(site_ratio), int(site_depth), 0)
        _check(load().mb_merge_sites_dev(self._h, C.c_void_p(counts.data_ptr()), int(S), int(G),
                                         C.c_void_p(md.data_ptr()), C.byref(mp),
                                         C.c_void_p(out["pooled"].data_ptr()), C.c_void_p(out["call"].data_ptr()),
                                         C.c_void_p(out["count_samples"].data_ptr()),
                                         C.c_void_p(out["sample_depth"].data_ptr()),
                                         C.c_void_p(out["sample_maf"].data_ptr())))
        return out

    # ---- genes
    def genes_coverage_dev(self, reads: DevReads, gene_len, n_genes: int, filters: Filters, out, exact_order=False):
        f = CFilters.of(filters)
        _check(load().mb_genes_coverage_dev(self._h, reads._h, C.c_void_p(gene_len.data_ptr()), int(n_genes),
                                            C.byref(f), int(exact_order), C.c_void_p(out.data_ptr())))

    def genes_coverage_plain(self, reads: PlainReads, gene_len: np.ndarray, filters: Filters, exact_order=False):
        """Decoded records in host memory -> per-gene outputs (mb_genes_coverage_plain)."""
        gene_len = np.ascontiguousarray(gene_len, np.uint32)
        out = np.zeros(len(gene_len), GENE_OUT_DTYPE)
        f = CFilters.of(filters)
        cin = CPlainReads.of(reads)
        _check(load().mb_genes_coverage_plain(self._h, C.byref(cin), _ptr(gene_len), len(gene_len), C.byref(f),
                                              int(exact_order), _ptr(out)))
        return out

    def genes_coverage_bam(self, stream: "BamStream", tid_map, gene_len: np.ndarray, filters: Filters, exact_order=False):
        """From the inflated BAM record stream (mb_genes_coverage_bam); tid_map: int32 [n_ref] BAM reference -> gene
        index (-1 = drop). Returns (per-gene outputs, n_reads)."""
        gene_len = np.ascontiguousarray(gene_len, np.uint32)
        tid_map = np.ascontiguousarray(tid_map, np.int32)
        assert len(tid_map) == len(stream.names)
        out = np.zeros(len(gene_len), GENE_OUT_DTYPE)
        f = CFilters.of(filters)
        nr = C.c_uint64()
        _check(load().mb_genes_coverage_bam(self._h, stream._h, _ptr(tid_map), _ptr(gene_len), len(gene_len), C.byref(f),
                                            int(exact_order), _ptr(out), C.byref(nr)))
        return out, int(nr.value)

    def genes_coverage(self, packed: PackedReads, gene_len: np.ndarray, filters: Filters, exact_order=False):
        gene_len = np.ascontiguousarray(gene_len, np.uint32)
        out = np.zeros(len(gene_len), GENE_OUT_DTYPE)
        f = CFilters.of(filters)
        _check(load().mb_genes_coverage(self._h, packed._p, _ptr(gene_len), len(gene_len), C.byref(f),
                                        int(exact_order), _ptr(out)))
        return out


def format_snps_rows(contig_name: str, ref_bases: np.ndarray, first_pos0: int, counts4: np.ndarray) -> bytes:
    """Text rows of <species>.snps.gz (reference snps.py:201-210) for counts4 = uint32 [4][n] of one
    contig starting at 0-based position first_pos0; ref_bases = uint8 [n] upper-cased characters."""
    L = load()
    n = int(counts4.shape[1])
    assert counts4.dtype == np.uint32 and counts4.shape[0] == 4 and ref_bases.dtype == np.uint8 and len(ref_bases) == n
    planes = [np.ascontiguousarray(counts4[k]) for k in range(4)]
    ref_bases = np.ascontiguousarray(ref_bases)
    name = contig_name.encode()
    cap = int(L.mb_format_snps_bound(len(name), n))
    buf = C.create_string_buffer(max(cap, 1))
    w = L.mb_format_snps_rows(name, _ptr(ref_bases), int(first_pos0), n, _ptr(planes[0]), _ptr(planes[1]),
                              _ptr(planes[2]), _ptr(planes[3]), buf, cap)
    if w < 0:
        _check(int(w))
    return buf.raw[:w]


class _BamHandle:
    """Keeps an mb_bam alive for as long as numpy views of its arrays exist."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        h, self.h = self.h, None
        if h and _lib is not None:
            _lib.mb_bam_close(h)


def format_merge_rows(site_index: np.ndarray, maf: np.ndarray, depth: np.ndarray, first_site: int = 0):
    """Text rows of snps_freq.txt and snps_depth.txt (reference midas/merge/snps.py:196-201) for the kept sites
    `site_index` (0-based within the chunk, uint64): maf f64 [S][G] and depth u32 [S][G] as mb_merge_sites_dev wrote
    them; the printed site id is first_site + index + 1. Returns (freq_bytes, depth_bytes)."""
    L = load()
    site_index = np.ascontiguousarray(site_index, np.uint64)
    maf = np.ascontiguousarray(maf, np.float64)
    depth = np.ascontiguousarray(depth, np.uint32)
    S, G = maf.shape
    assert depth.shape == (S, G)
    n = len(site_index)
    cap = int(L.mb_format_merge_bound(n, S))
    out = []
    for fn, arr in ((L.mb_format_merge_freq_rows, maf), (L.mb_format_merge_depth_rows, depth)):
        buf = C.create_string_buffer(max(cap, 1))
        w = fn(_ptr(site_index), n, S, G, _ptr(arr), int(first_site), buf, cap)
        if w < 0:
            _check(int(w))
        out.append(buf.raw[:w])
    return out[0], out[1]


MB_BAM_PINNED, MB_BAM_SEQ_BYTES = 1, 2


def read_bam(path: str, threads: int = 0, pinned: bool = False, seq_bytes: bool = False):
    """Native BAM decode (mb_bam_open): returns (ref_names, ref_lens, PlainReads) in file order, like
    midas_b200.bam.read_bam (the pure-Python reader it must agree with). seq stays as in the file (two
    codes per byte, PlainReads.seq_nibbles) unless seq_bytes; pinned = page-locked arrays (they are about
    to cross PCIe). The arrays are views of the handle's buffers (no copy); the handle is closed when the
    last of them goes away."""
    L = load()
    h = C.c_void_p()
    _check(L.mb_bam_open(os.fsencode(path), int(threads),
                         (MB_BAM_PINNED if pinned else 0) | (MB_BAM_SEQ_BYTES if seq_bytes else 0), C.byref(h)))
    owner = _BamHandle(h)
    n_ref = int(L.mb_bam_n_ref(h))
    names = [L.mb_bam_ref_name(h, i).decode() for i in range(n_ref)]
    lens = [int(L.mb_bam_ref_len(h, i)) for i in range(n_ref)]
    r = L.mb_bam_reads(h).contents
    n = int(r.n_reads)

    def arr(ptr, count, dt):
        if count == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr)
        buf._owner = owner                      # numpy array -> ctypes buffer -> handle
        return np.frombuffer(buf, dtype=dt, count=count)

    coff = arr(r.cigar_off, n + 1, np.uint64)
    soff = arr(r.seq_off, n + 1, np.uint64)
    reads = PlainReads(arr(r.ref_id, n, np.int32), arr(r.pos, n, np.int32), arr(r.mapq, n, np.uint8),
                       arr(r.nm, n, np.int32), arr(r.l_qseq, n, np.int32), coff,
                       arr(r.cigar, int(coff[-1]), np.uint32), soff,
                       arr(r.seq, (int(soff[-1]) + 1) // 2 if r.seq_nibbles else int(soff[-1]), np.uint8),
                       arr(r.qual, int(soff[-1]), np.uint8), arr(L.mb_bam_flags(h), n, np.uint16),
                       bool(r.seq_nibbles))
    return names, lens, reads


class BamStream:
    """A BAM file inflated on the host threads (mb_bam_stream_open): header + the raw record stream, which the GPU
    parses (Context.snps_pileup_bam). No host-side record decode."""

    def __init__(self, path: str, threads: int = 0, ctx: "Context" = None):
        """With `ctx` (mb_bam_stream_open_dev) the inflated stream crosses PCIe while the host threads are still
        inflating, and ctx.snps_pileup_bam / ctx.genes_coverage_bam find it on the device."""
        L = load()
        h = C.c_void_p()
        if ctx is not None:
            _check(L.mb_bam_stream_open_dev(ctx._h, os.fsencode(path), int(threads), C.byref(h)))
        else:
            _check(L.mb_bam_stream_open(os.fsencode(path), int(threads), C.byref(h)))
        self._h = h
        n = int(L.mb_bam_stream_n_ref(h))
        self.names = [L.mb_bam_stream_ref_name(h, i).decode() for i in range(n)]
        self.lens = [int(L.mb_bam_stream_ref_len(h, i)) for i in range(n)]
        self.nbytes = int(L.mb_bam_stream_bytes(h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mb_bam_stream_close(h)

    __del__ = close


class _Pinned:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        p, self.ptr = self.ptr, None
        if p and _lib is not None:
            _lib.mb_host_free(p)


def pinned_empty(n: int, dtype) -> np.ndarray:
    """Page-locked host array (mb_host_alloc); freed when the array goes away."""
    dt = np.dtype(dtype)
    p = C.c_void_p()
    _check(load().mb_host_alloc(max(1, int(n)) * dt.itemsize, C.byref(p)))
    owner = _Pinned(p)
    if n == 0:
        return np.zeros(0, dt)
    buf = (C.c_char * (int(n) * dt.itemsize)).from_address(p.value)
    buf._owner = owner
    return np.frombuffer(buf, dtype=dt, count=int(n))


def pinned_reads(reads: PlainReads) -> PlainReads:
    """Copy of `reads` whose arrays are page-locked (what the BAM reader hands out with pinned=True)."""
    def pin(a):
        out = pinned_empty(a.shape[0], a.dtype)
        out[...] = a
        return out
    return PlainReads(pin(reads.ref_id), pin(reads.pos), pin(reads.mapq), pin(reads.nm), pin(reads.l_qseq),
                      pin(reads.cigar_off), pin(reads.cigar), pin(reads.seq_off), pin(reads.seq), pin(reads.qual),
                      reads.flag, reads.seq_nibbles)
