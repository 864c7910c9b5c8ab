"""ctypes binding of libe2fm_b200.so — the C ABI declared in include/e2fm_b200.h.

This is plumbing for the tests and bench.py: every call goes straight to the CUDA library. There is no
fallback: if the library has not been built, or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libe2fm_b200.so")

E2FM_OK = 0
MODE_COUNT = 0
MODE_LOCATE = 1
OPT_CHUNK = 2
OPT_WINDOW = 3
OPT_DENSE = 4
OPT_PIPE = 5
OPT_DENSITY = 6
OPT_TAPER = 7
OPT_PHASE_TIMERS = 8
OPT_SEED = 9
OPT_HITS_STAGE = 10

# every symbol include/e2fm_b200.h declares (tests check that the built library exports all of them)
EXPORTS = [
    "e2fm_index_open", "e2fm_parse_header", "e2fm_index_close", "e2fm_index_info", "e2fm_index_terminators", "e2fm_index_fasta_headers",
    "e2fm_index_stream", "e2fm_search_batch", "e2fm_search_batch_fixed", "e2fm_search_batch_device",
    "e2fm_search_batch_device_fixed", "e2fm_search_batch_stream", "e2fm_host_alloc", "e2fm_host_free", "e2fm_device_count",
    "e2fm_extract", "e2fm_result_patterns", "e2fm_result_total_positions", "e2fm_result_fetch", "e2fm_result_device_ptrs",
    "e2fm_result_free", "e2fm_last_batch_ms", "e2fm_last_batch_counters", "e2fm_set_option",
    "e2fm_debug_gather_bench", "e2fm_debug_keystream", "e2fm_debug_salsa20", "e2fm_debug_bwt", "e2fm_debug_marks", "e2fm_debug_lf",
    "e2fm_debug_rank", "e2fm_last_error", "e2fm_version",
    "e2fm_packed_stride", "e2fm_pack_patterns", "e2fm_search_batch_packed", "e2fm_search_batch_packed_device",
    "e2fm_result_fetch_compact", "e2fm_debug_seed",
    "e2fm_shardset_unique_id", "e2fm_shardset_create", "e2fm_shardset_destroy", "e2fm_shardset_slice", "e2fm_shardset_search",
    "e2fm_shardset_last_ms", "e2fm_search_hits", "e2fm_debug_gather_bench_ex", "e2fm_packed_tight_stride", "e2fm_pack_patterns_tight",
    "e2fm_encode_buckets", "e2fm_debug_hits_stages",
]


class E2FMError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Info(C.Structure):
    _fields_ = [("text_length", C.c_uint64), ("bucket_size", C.c_uint64), ("superbucket_size", C.c_uint64),
                ("n_buckets", C.c_uint64), ("n_superbuckets", C.c_uint64), ("n_sequences", C.c_uint64),
                ("initial_n", C.c_uint64), ("device_bytes", C.c_uint64), ("k", C.c_uint32), ("n_symbols", C.c_uint32),
                ("compact_alphabet", C.c_uint32), ("marked_rows_percentage", C.c_uint32), ("marking_rate", C.c_uint32),
                ("rank_block", C.c_uint32), ("quirks", C.c_uint32), ("dense", C.c_uint32), ("seed_chars", C.c_uint32),
                ("seed_table_mb", C.c_uint32), ("vrec_mb", C.c_uint32), ("reserved", C.c_uint32), ("load_ms", C.c_float * 8)]


class Hits(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("occ_offsets", C.c_void_p), ("seq_index", C.c_void_p), ("seq_offset", C.c_void_p),
                ("capacity", C.c_uint64), ("total", C.c_uint64)]


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C e2fm_b200/csrc). e2fm_b200 has no CPU search path.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.e2fm_last_error.restype = C.c_char_p
    L.e2fm_version.restype = C.c_char_p
    L.e2fm_device_count.restype = i32
    L.e2fm_host_alloc.restype = vp
    L.e2fm_host_alloc.argtypes = [C.c_size_t]
    L.e2fm_host_free.argtypes = [vp]
    L.e2fm_index_open.argtypes = [vp, C.c_size_t, C.c_char_p, i32, C.POINTER(vp)]
    L.e2fm_parse_header.argtypes = [vp, C.c_size_t, C.c_char_p, C.POINTER(Info), vp, u64, vp, u64]
    L.e2fm_index_close.argtypes = [vp]
    L.e2fm_index_close.restype = None
    L.e2fm_index_info.argtypes = [vp, C.POINTER(Info)]
    L.e2fm_index_terminators.argtypes = [vp, vp, u64]
    L.e2fm_index_fasta_headers.restype = u64
    L.e2fm_index_fasta_headers.argtypes = [vp, vp, u64]
    L.e2fm_index_stream.restype = vp
    L.e2fm_index_stream.argtypes = [vp]
    L.e2fm_search_batch.argtypes = [vp, vp, vp, u64, i32, C.POINTER(vp)]
    L.e2fm_search_batch_fixed.argtypes = [vp, vp, u64, u32, i32, C.POINTER(vp)]
    L.e2fm_search_batch_stream.argtypes = [vp, vp, vp, u64, u32, i32, vp, C.POINTER(vp)]
    L.e2fm_search_batch_device.argtypes = [vp, vp, vp, u64, u64, i32, C.POINTER(vp)]
    L.e2fm_search_batch_device_fixed.argtypes = [vp, vp, u64, u32, i32, C.POINTER(vp)]
    L.e2fm_packed_stride.restype = u32
    L.e2fm_packed_stride.argtypes = [u32]
    L.e2fm_pack_patterns.argtypes = [vp, u64, u32, vp, C.POINTER(u64)]
    L.e2fm_packed_tight_stride.argtypes = [u32]
    L.e2fm_debug_hits_stages.argtypes = [u64, u64, vp, u64, C.POINTER(u64)]
    L.e2fm_encode_buckets.argtypes = [i32, C.c_char_p, vp, u64, u64, u64, u64, u64, u64, vp, vp, vp, vp, u64, C.POINTER(u64), vp,
                                      C.POINTER(C.c_float)]
    L.e2fm_packed_tight_stride.restype = u32
    L.e2fm_pack_patterns_tight.argtypes = [vp, u64, u32, vp, C.POINTER(u64)]
    L.e2fm_search_batch_packed.argtypes = [vp, vp, u64, u32, i32, vp, C.POINTER(vp)]
    L.e2fm_search_batch_packed_device.argtypes = [vp, vp, u64, u32, i32, C.POINTER(vp)]
    L.e2fm_result_fetch_compact.argtypes = [vp, vp, vp, vp, vp]
    L.e2fm_debug_seed.argtypes = [vp, vp, u64, vp]
    L.e2fm_search_hits.argtypes = [vp, vp, u64, u32, i32, i32, C.POINTER(Hits)]
    L.e2fm_shardset_unique_id.argtypes = [vp]
    L.e2fm_shardset_create.argtypes = [vp, vp, i32, i32, u64, u64, C.POINTER(vp)]
    L.e2fm_shardset_destroy.argtypes = [vp]
    L.e2fm_shardset_destroy.restype = None
    L.e2fm_shardset_slice.restype = u64
    L.e2fm_shardset_slice.argtypes = [u64, i32, i32, C.POINTER(u64)]
    L.e2fm_shardset_search.argtypes = [vp, vp, u64, u32, i32, i32, C.POINTER(vp)]
    L.e2fm_shardset_last_ms.argtypes = [vp, C.POINTER(C.c_float * 8)]
    L.e2fm_extract.argtypes = [vp, u64, u64, vp, u64, C.POINTER(u64)]
    L.e2fm_result_patterns.restype = u64
    L.e2fm_result_patterns.argtypes = [vp]
    L.e2fm_result_total_positions.restype = u64
    L.e2fm_result_total_positions.argtypes = [vp]
    L.e2fm_result_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
    L.e2fm_result_device_ptrs.argtypes = [vp] + [C.POINTER(vp)] * 5
    L.e2fm_result_free.argtypes = [vp]
    L.e2fm_result_free.restype = None
    L.e2fm_last_batch_ms.argtypes = [vp, C.POINTER(C.c_float * 8)]
    L.e2fm_last_batch_counters.argtypes = [vp, C.POINTER(C.c_uint64 * 16)]
    L.e2fm_set_option.argtypes = [vp, i32, u64]
    L.e2fm_debug_keystream.argtypes = [i32, C.c_char_p, u64, u32, u64, u64, vp]
    L.e2fm_debug_gather_bench.argtypes = [i32, u64, u64, C.POINTER(C.c_double)]
    L.e2fm_debug_salsa20.argtypes = [i32, C.c_char_p, u64, u64, u64, vp]
    L.e2fm_debug_bwt.argtypes = [vp, u64, u64, vp]
    L.e2fm_debug_marks.argtypes = [vp, u64, u64, vp]
    L.e2fm_debug_lf.argtypes = [vp, vp, u64, vp]
    L.e2fm_debug_rank.argtypes = [vp, vp, vp, u64, vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != E2FM_OK:
        raise E2FMError(rc, lib().e2fm_last_error().decode(errors="replace"))


def parse_header(index_bytes: bytes, key64: bytes):
    """host-only header parse (no GPU): -> (Info, C[] as uint64 array, terminators as int64 array)"""
    info = Info()
    buf = np.frombuffer(index_bytes, dtype=np.uint8)
    c = np.zeros(1 << 16, dtype=np.uint64)
    t = np.zeros(1 << 16, dtype=np.int64)
    check(lib().e2fm_parse_header(buf.ctypes.data, len(buf), key64, C.byref(info), c.ctypes.data, len(c), t.ctypes.data, len(t)))
    return info, c[:info.compact_alphabet + 1].copy(), t[:min(int(info.n_sequences), len(t))].copy()


def device_count() -> int:
    return lib().e2fm_device_count()


class PinnedArray:
    """A numpy view over page-locked host memory from e2fm_host_alloc."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._ptr = lib().e2fm_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise MemoryError("e2fm_host_alloc failed: " + lib().e2fm_last_error().decode())
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().e2fm_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Result:
    """Device-resident result of one batch (e2fm_result)."""

    def __init__(self, handle, locate: bool):
        self._h = handle
        self.locate = locate
        self.n = int(lib().e2fm_result_patterns(handle))
        self.total = int(lib().e2fm_result_total_positions(handle))

    def fetch(self, counts=None, pos_offsets=None, positions=None, seq_index=None, seq_offset=None, want_seq=True):
        """D2H copy. Pass preallocated (pinned) arrays or leave None to allocate. Returns a dict."""
        if counts is None:
            counts = np.empty(self.n, dtype=np.uint64)
        out = {"counts": counts}
        p = lambda a: a.ctypes.data if a is not None and a.size else None  # noqa: E731
        if self.locate:
            if pos_offsets is None:
                pos_offsets = np.empty(self.n + 1, dtype=np.uint64)
            if positions is None:
                positions = np.empty(self.total, dtype=np.uint64)
            if want_seq:
                if seq_index is None:
                    seq_index = np.empty(self.total, dtype=np.uint32)
                if seq_offset is None:
                    seq_offset = np.empty(self.total, dtype=np.uint64)
            out.update(pos_offsets=pos_offsets, positions=positions, seq=seq_index, off=seq_offset)
        check(lib().e2fm_result_fetch(self._h, p(counts), pos_offsets.ctypes.data if pos_offsets is not None else None,
                                      p(positions), p(seq_index), p(seq_offset)))
        return out

    def fetch_positions(self, pos_offsets, positions):
        """D2H of the CSR offsets and positions only (counts already streamed back)"""
        check(lib().e2fm_result_fetch(self._h, None, pos_offsets.ctypes.data, positions.ctypes.data if positions.size else None, None, None))

    def fetch_occurrences(self, pos_offsets, seq_index, seq_offset):
        """D2H of what EFMCollection::search returns per pattern (FMCollection.h:23-44): CSR offsets, sequence index and
        offset in the sequence of every occurrence (counts already streamed back)"""
        check(lib().e2fm_result_fetch(self._h, None, pos_offsets.ctypes.data, None,
                                      seq_index.ctypes.data if seq_index.size else None,
                                      seq_offset.ctypes.data if seq_offset.size else None))

    def fetch_compact(self, counts=None, pos_offsets=None, seq_index=None, seq_offset=None):
        """32-bit D2H of whichever arrays are given (e2fm_result_fetch_compact)"""
        p = lambda a: a.ctypes.data if a is not None and a.size else None  # noqa: E731
        check(lib().e2fm_result_fetch_compact(self._h, p(counts), pos_offsets.ctypes.data if pos_offsets is not None else None,
                                              p(seq_index), p(seq_offset)))

    def device_ptrs(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        check(lib().e2fm_result_device_ptrs(self._h, *[C.byref(x) for x in ptrs]))
        return {k: (v.value or 0) for k, v in zip(["counts", "pos_offsets", "positions", "seq", "off"], ptrs)}

    def free(self):
        if self._h:
            lib().e2fm_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceIndex:
    """An E2FM index flattened into the HBM of one GPU (e2fm_index)."""

    def __init__(self, index_bytes: bytes, key64: bytes, device: int = 0):
        if len(key64) != 64:
            raise ValueError("the encryption key must be 64 bytes")   # EncryptionManager.cpp:25-36
        h = C.c_void_p()
        buf = np.frombuffer(index_bytes, dtype=np.uint8)
        check(lib().e2fm_index_open(buf.ctypes.data, len(buf), key64, device, C.byref(h)))
        self._h = h
        self.device = device
        self.info = Info()
        check(lib().e2fm_index_info(self._h, C.byref(self.info)))

    @classmethod
    def from_file(cls, path: str, key64: bytes, device: int = 0):
        with open(path, "rb") as f:
            return cls(f.read(), key64, device)

    def close(self):
        if getattr(self, "_h", None):
            lib().e2fm_index_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- metadata
    def terminators(self) -> np.ndarray:
        out = np.zeros(int(self.info.n_sequences), dtype=np.int64)
        check(lib().e2fm_index_terminators(self._h, out.ctypes.data, len(out)))
        return out

    def fasta_headers(self) -> bytes:
        n = lib().e2fm_index_fasta_headers(self._h, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        lib().e2fm_index_fasta_headers(self._h, buf, n)
        return buf.raw[:n]

    def stream(self) -> int:
        return lib().e2fm_index_stream(self._h) or 0

    def set_option(self, opt: int, value: int):
        check(lib().e2fm_set_option(self._h, opt, value))

    # ---- search
    def search_fixed(self, pats: np.ndarray, locate: bool) -> Result:
        """pats: (n, L) uint8 host array (pinned or not)."""
        assert pats.dtype == np.uint8 and pats.ndim == 2 and pats.flags.c_contiguous
        h = C.c_void_p()
        n, L = pats.shape
        check(lib().e2fm_search_batch_fixed(self._h, pats.ctypes.data if pats.size else None, n, L,
                                            MODE_LOCATE if locate else MODE_COUNT, C.byref(h)))
        return Result(h, locate)

    def search(self, data: np.ndarray, offsets: np.ndarray, locate: bool) -> Result:
        """Ragged batch: pattern i = data[offsets[i]:offsets[i+1]]."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h = C.c_void_p()
        n = len(offsets) - 1
        check(lib().e2fm_search_batch(self._h, data.ctypes.data if data.size else None, offsets.ctypes.data, n,
                                      MODE_LOCATE if locate else MODE_COUNT, C.byref(h)))
        return Result(h, locate)

    def search_list(self, patterns, locate: bool) -> Result:
        """patterns: iterable of bytes objects."""
        patterns = [bytes(p) for p in patterns]
        offs = np.zeros(len(patterns) + 1, dtype=np.uint64)
        if patterns:
            offs[1:] = np.cumsum([len(p) for p in patterns])
        data = np.frombuffer(b"".join(patterns), dtype=np.uint8)
        return self.search(data, offs, locate)

    def search_stream(self, pats: np.ndarray, locate: bool, counts_out: np.ndarray) -> Result:
        """fixed-length host batch, pipelined; counts arrive in counts_out (uint64[n], pinned for full overlap)"""
        assert pats.dtype == np.uint8 and pats.ndim == 2 and pats.flags.c_contiguous
        assert counts_out.dtype == np.uint64 and counts_out.size >= pats.shape[0]
        h = C.c_void_p()
        n, L = pats.shape
        check(lib().e2fm_search_batch_stream(self._h, pats.ctypes.data if pats.size else None, None, n, L,
                                             MODE_LOCATE if locate else MODE_COUNT, counts_out.ctypes.data, C.byref(h)))
        return Result(h, locate)

    def search_packed(self, packed: np.ndarray, n: int, length: int, locate: bool, counts_out: np.ndarray = None) -> Result:
        """packed: host uint8 array of n * packed_stride(length) bytes (pinned for full overlap)"""
        assert packed.dtype == np.uint8 and packed.flags.c_contiguous and packed.size >= n * packed_stride(length)
        h = C.c_void_p()
        check(lib().e2fm_search_batch_packed(self._h, packed.ctypes.data if packed.size else None, n, length,
                                             MODE_LOCATE if locate else MODE_COUNT,
                                             counts_out.ctypes.data if counts_out is not None else None, C.byref(h)))
        return Result(h, locate)

    def search_hits(self, patterns: np.ndarray, n: int, length: int, packed: bool, locate: bool, counts: np.ndarray = None,
                    occ_offsets: np.ndarray = None, seq_index: np.ndarray = None, seq_offset: np.ndarray = None) -> int:
        """e2fm_search_hits: host patterns (packed: 0/False ASCII n x length, 1/True 16-byte words, 2 tight) -> 32-bit results in
        the caller's host arrays, pipelined. Returns the number of occurrences written (locate)."""
        assert patterns.dtype == np.uint8 and patterns.flags.c_contiguous
        for a in (counts, occ_offsets, seq_index, seq_offset):
            assert a is None or (a.dtype == np.uint32 and a.flags.c_contiguous)
        h = Hits()
        h.counts = counts.ctypes.data if counts is not None else None
        h.occ_offsets = occ_offsets.ctypes.data if occ_offsets is not None else None
        h.seq_index = seq_index.ctypes.data if seq_index is not None else None
        h.seq_offset = seq_offset.ctypes.data if seq_offset is not None else None
        h.capacity = min(seq_index.size, seq_offset.size) if (seq_index is not None and seq_offset is not None) else 0
        self.last_hits_total = 0
        rc = lib().e2fm_search_hits(self._h, patterns.ctypes.data if patterns.size else None, n, length, int(packed),
                                    MODE_LOCATE if locate else MODE_COUNT, C.byref(h))
        self.last_hits_total = int(h.total)
        check(rc)
        return int(h.total)

    def search_packed_device(self, d_ptr: int, n: int, length: int, locate: bool) -> Result:
        h = C.c_void_p()
        check(lib().e2fm_search_batch_packed_device(self._h, d_ptr, n, length, MODE_LOCATE if locate else MODE_COUNT, C.byref(h)))
        return Result(h, locate)

    def debug_seed(self, codes) -> np.ndarray:
        """codes: (count, seed_chars) compact codes -> (count, 3) [start, rows, exact]"""
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        out = np.zeros((codes.shape[0], 3), dtype=np.uint32)
        check(lib().e2fm_debug_seed(self._h, codes.ctypes.data, codes.shape[0], out.ctypes.data))
        return out

    def search_device_fixed(self, d_ptr: int, n: int, length: int, locate: bool) -> Result:
        h = C.c_void_p()
        check(lib().e2fm_search_batch_device_fixed(self._h, d_ptr, n, length, MODE_LOCATE if locate else MODE_COUNT, C.byref(h)))
        return Result(h, locate)

    def extract_snippet(self, frm: int, to: int) -> bytes:
        """EFMIndex::extractSnippet(from, to)"""
        cap = max(0, to - frm + 1) + 64
        buf = C.create_string_buffer(cap)
        n = C.c_uint64(0)
        check(lib().e2fm_extract(self._h, frm & (2**64 - 1), to & (2**64 - 1), buf, cap, C.byref(n)))
        return buf.raw[:n.value]

    def last_batch_ms(self):
        a = (C.c_float * 8)()
        check(lib().e2fm_last_batch_ms(self._h, C.byref(a)))
        return dict(zip(["h2d", "backward_search", "expand", "walk", "encode", "results", "total", "firstcheck"], [float(x) for x in a]))

    def last_batch_counters(self):
        a = (C.c_uint64 * 16)()
        check(lib().e2fm_last_batch_counters(self._h, C.byref(a)))
        return dict(zip(["rank", "lf", "mark", "tasks", "rank_sectors", "launches", "occurrences", "walk_gathers", "scan_rows",
                         "chunks_redone", "complex_patterns", "seeded", "hits_stages", "hits_gave_up"], [int(x) for x in a]))

    # ---- test hooks
    def debug_bwt(self, row0: int, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint32)
        check(lib().e2fm_debug_bwt(self._h, row0, count, out.ctypes.data))
        return out

    def debug_marks(self, row0: int, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint64)
        check(lib().e2fm_debug_marks(self._h, row0, count, out.ctypes.data))
        return out

    def debug_lf(self, rows) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(len(rows), dtype=np.uint64)
        check(lib().e2fm_debug_lf(self._h, rows.ctypes.data, len(rows), out.ctypes.data))
        return out

    def debug_rank(self, codes, rows) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(len(rows), dtype=np.uint64)
        check(lib().e2fm_debug_rank(self._h, codes.ctypes.data, rows.ctypes.data, len(rows), out.ctypes.data))
        return out


def shardset_unique_id() -> bytes:
    """128 bytes of a fresh NCCL unique id (rank 0 calls this and hands them to the other ranks)"""
    buf = (C.c_uint8 * 128)()
    check(lib().e2fm_shardset_unique_id(buf))
    return bytes(buf)


def shardset_slice(n_total: int, world: int, rank: int):
    """(first pattern, number of patterns) of the batch that `rank` supplies to e2fm_shardset_search"""
    cnt = C.c_uint64(0)
    lo = lib().e2fm_shardset_slice(n_total, world, rank, C.byref(cnt))
    return int(lo), int(cnt.value)


class ShardSet:
    """one rank's shard of a collection sharded by sequence + the NCCL communicator (e2fm_shardset)"""

    def __init__(self, index: "DeviceIndex", unique_id: bytes, rank: int, world: int, first_sequence: int, text_base: int):
        assert len(unique_id) == 128
        self.index = index                      # borrowed: keep it alive
        self.rank, self.world = rank, world
        h = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().e2fm_shardset_create(index._h, idbuf, rank, world, first_sequence, text_base, C.byref(h)))
        self._h = h

    def search(self, slice_host: np.ndarray, n_total: int, length: int, packed: bool, locate: bool) -> "Result":
        """slice_host: this rank's patterns [lo, lo+count) of the batch (shardset_slice), uint8, ASCII or packed; the result
        holds the merged answers for those patterns"""
        h = C.c_void_p()
        ptr = slice_host.ctypes.data if slice_host is not None and slice_host.size else None
        check(lib().e2fm_shardset_search(self._h, ptr, n_total, length, 1 if packed else 0, MODE_LOCATE if locate else MODE_COUNT,
                                         C.byref(h)))
        return Result(h, locate)

    def last_ms(self):
        a = (C.c_float * 8)()
        check(lib().e2fm_shardset_last_ms(self._h, C.byref(a)))
        return dict(zip(["h2d_allgather", "search", "merge_comm", "merge_kernels", "total"], [float(x) for x in a[:5]]))

    def close(self):
        if getattr(self, "_h", None):
            lib().e2fm_shardset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def packed_stride(length: int) -> int:
    return int(lib().e2fm_packed_stride(length))


def pack_patterns(pats: np.ndarray, out: np.ndarray = None):
    """(n, L) ASCII uint8 -> (packed uint8 array of n * packed_stride(L) bytes, number of patterns with a byte outside ACGT)"""
    assert pats.dtype == np.uint8 and pats.ndim == 2 and pats.flags.c_contiguous
    n, L = pats.shape
    if out is None:
        out = np.zeros(n * packed_stride(L), dtype=np.uint8)
    bad = C.c_uint64(0)
    check(lib().e2fm_pack_patterns(pats.ctypes.data if pats.size else None, n, L, out.ctypes.data if out.size else None, C.byref(bad)))
    return out, int(bad.value)


def packed_tight_stride(length: int) -> int:
    return int(lib().e2fm_packed_tight_stride(length))


def pack_patterns_tight(pats: np.ndarray, out: np.ndarray = None):
    """(n, L) ASCII uint8 -> (n * ceil(L / 4) bytes without padding, number of patterns with a byte outside ACGT)"""
    assert pats.dtype == np.uint8 and pats.ndim == 2 and pats.flags.c_contiguous
    n, L = pats.shape
    if out is None:
        out = np.zeros(n * packed_tight_stride(L), dtype=np.uint8)
    bad = C.c_uint64(0)
    check(lib().e2fm_pack_patterns_tight(pats.ctypes.data if pats.size else None, n, L, out.ctypes.data if out.size else None, C.byref(bad)))
    return out, int(bad.value)


def encode_buckets(key64: bytes, codes: np.ndarray, bucket_size: int, first_bucket: int, per_super_bucket: int, total_buckets: int,
                   alpha: np.ndarray, want_sequence: bool = False, payload_cap: int = None, device: int = 0) -> dict:
    """e2fm_encode_buckets: Bucket::encodeText + the symbol loop of Bucket::writeText for consecutive buckets.
    -> {"clen": [n_buckets], "offsets": [n_buckets + 1], "payload": bytes-like, "sequence": [n_rows] or None, "device_ms": float}"""
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    alpha = np.ascontiguousarray(alpha, dtype=np.uint32)
    nb = len(alpha)
    clen = np.zeros(max(nb, 1), dtype=np.uint64)
    off = np.zeros(nb + 1, dtype=np.uint64)
    cap = payload_cap if payload_cap is not None else 2 * len(codes) + 8 * nb + 64    # <= 16 bits per symbol, 8-byte slots
    payload = np.zeros(max(cap, 8), dtype=np.uint8)
    seq = np.zeros(max(len(codes), 1), dtype=np.uint32) if want_sequence else None
    need = C.c_uint64(0)
    ms = C.c_float(0)
    rc = lib().e2fm_encode_buckets(device, key64, codes.ctypes.data if codes.size else None, len(codes), bucket_size, first_bucket, nb,
                                   per_super_bucket, total_buckets, alpha.ctypes.data if nb else None, clen.ctypes.data, off.ctypes.data,
                                   payload.ctypes.data, cap, C.byref(need), seq.ctypes.data if want_sequence else None, C.byref(ms))
    if rc != 0:
        e = E2FMError(rc, lib().e2fm_last_error().decode(errors="replace"))
        e.payload_bytes = int(need.value)
        raise e
    return {"clen": clen[:nb], "offsets": off, "payload": payload[:int(need.value)], "sequence": seq[:len(codes)] if want_sequence else None,
            "device_ms": float(ms.value), "payload_bytes": int(need.value)}


def hits_stages(n: int, stage_option: int = 0) -> np.ndarray:
    """stage bounds of an e2fm_search_hits batch of n patterns (host-only helper)"""
    out = np.zeros(n // max(1, stage_option or 131072) + 64, dtype=np.uint64)
    cnt = C.c_uint64(0)
    check(lib().e2fm_debug_hits_stages(n, stage_option, out.ctypes.data, len(out), C.byref(cnt)))
    return out[:int(cnt.value)]


def debug_keystream(key64: bytes, nonce: int, max_value: int, start: int, count: int, device: int = 0) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint32)
    check(lib().e2fm_debug_keystream(device, key64, nonce, max_value, start, count, out.ctypes.data))
    return out


def gather_bench(buffer_bytes: int = 8 << 30, n_gathers: int = 1 << 30, device: int = 0, variant: int = 0) -> float:
    """HBM random-gather roofline in 10^9 gathers/s (one 32-byte sector each); variant = the load instruction (e2fm_b200.h)"""
    out = C.c_double(0)
    L = lib()
    L.e2fm_debug_gather_bench_ex.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    check(L.e2fm_debug_gather_bench_ex(device, buffer_bytes, n_gathers, variant, C.byref(out)))
    return float(out.value)


def debug_salsa20(key32: bytes, nonce: int, first_block: int, n_blocks: int, device: int = 0) -> bytes:
    out = np.zeros(n_blocks * 64, dtype=np.uint8)
    check(lib().e2fm_debug_salsa20(device, key32, nonce, first_block, n_blocks, out.ctypes.data))
    return out.tobytes()
