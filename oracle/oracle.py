"""ctypes binding of oracle/liboracle.so (the plain-C restatement) and a thin runner for the
compiled reference (oracle/_ref/e2fm_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package (e2fm_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "e2fm_ref")


def build(quiet: bool = True) -> None:
    """Compile the C restatement, and the reference itself when /root/reference is present."""
    subprocess.run(["make", "-C", HERE, "oracle", "ref"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class OrcInfo(C.Structure):
    _fields_ = [("n", C.c_uint64), ("k", C.c_uint32), ("n_sym", C.c_uint32), ("mrp", C.c_uint32),
                ("rate", C.c_uint32), ("bucket_size", C.c_uint64), ("superbucket_size", C.c_uint64),
                ("n_buckets", C.c_uint64), ("n_superbuckets", C.c_uint64), ("compact_alphabet", C.c_uint32),
                ("n_sequences", C.c_uint64), ("initial_n", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_int_log2.restype = C.c_uint32
        L.orc_int_log2.argtypes = [C.c_uint64]
        L.orc_salsa20_block.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_keystream_ints.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_scrambling_key.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p]
        L.orc_freq_permutation.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p]
        L.orc_freq_keystream.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.c_void_p]
        L.orc_rle_decompress.restype = C.c_size_t
        L.orc_rle_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_dump.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_info.argtypes = [C.c_void_p, C.POINTER(OrcInfo)]
        L.orc_search.restype = C.c_int64
        L.orc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint64),
                                 C.c_void_p, C.c_uint64]
        L.orc_map_position.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_get_C.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_terminators.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_extract.restype = C.c_int64
        L.orc_extract.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_get_bwt.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_get_marks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_lf_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_rank_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_get_counters.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3
        L.orc_reset_counters.argtypes = [C.c_void_p]
        L.orc_bucket_info.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_bucket_codes.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_encode_bucket.restype = C.c_uint64
        L.orc_encode_bucket.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        L.orc_pack_sequence.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_decode_sequence.restype = C.c_uint64
        L.orc_decode_sequence.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def int_log2(u: int) -> int:
    return lib().orc_int_log2(u)


def salsa20_block(key32: bytes, nonce: int, counter: int) -> bytes:
    out = C.create_string_buffer(64)
    lib().orc_salsa20_block(key32, nonce, counter, out)
    return out.raw


def keystream_ints(key32: bytes, nonce: int, max_value: int, start: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint32)
    lib().orc_keystream_ints(key32, nonce, max_value, start, count, out.ctypes.data)
    return out


def scrambling_key(key64: bytes, card: int) -> np.ndarray:
    out = np.zeros(card, dtype=np.uint32)
    lib().orc_scrambling_key(key64, card, out.ctypes.data)
    return out


def freq_permutation(key64: bytes, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_freq_permutation(key64, n, out.ctypes.data)
    return out


def freq_keystream(key64: bytes, n: int, max_freq: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_freq_keystream(key64, n, max_freq, out.ctypes.data)
    return out


def rle_decompress(bits) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(bits, dtype=np.uint8))
    n = lib().orc_rle_decompress(a.ctypes.data, len(a), None, 0)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    lib().orc_rle_decompress(a.ctypes.data, len(a), out.ctypes.data, n)
    return out[:n]


class OracleIndex:
    """The C restatement of EFMCollection load + search for one index file."""

    def __init__(self, index_bytes: bytes, key64: bytes):
        err = C.create_string_buffer(256)
        self._h = lib().orc_open(index_bytes, len(index_bytes), key64, err, 256)
        if not self._h:
            raise RuntimeError(err.value.decode())
        self.info = OrcInfo()
        lib().orc_get_info(self._h, C.byref(self.info))

    @classmethod
    def from_file(cls, path: str, key64: bytes):
        with open(path, "rb") as f:
            return cls(f.read(), key64)

    def close(self):
        if self._h:
            lib().orc_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dump(self, path: str) -> None:
        if lib().orc_dump(self._h, path.encode()) != 0:
            raise OSError("cannot write " + path)

    def search(self, pattern: bytes, locate: bool = True):
        """-> (count, sorted unique global positions as uint64 array)"""
        cnt = C.c_uint64(0)
        cap = 64
        buf = np.zeros(cap, dtype=np.uint64)
        pat = np.frombuffer(bytes(pattern), dtype=np.uint8)
        p = pat.ctypes.data if len(pat) else None
        n = lib().orc_search(self._h, p, len(pat), 1 if locate else 0, C.byref(cnt), buf.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("The index doesn't contain marked rows: locate operation not supported")
        if n > cap:
            buf = np.zeros(n, dtype=np.uint64)
            n = lib().orc_search(self._h, p, len(pat), 1, C.byref(cnt), buf.ctypes.data, n)
        return int(cnt.value), buf[:n].copy()

    def search_batch(self, pats: np.ndarray, locate: bool = True):
        """pats: (n, L) uint8. -> counts u64[n], pos_offsets u64[n+1], positions u64[total]"""
        n = len(pats)
        counts = np.zeros(n, dtype=np.uint64)
        offs = np.zeros(n + 1, dtype=np.uint64)
        chunks = []
        for i in range(n):
            c, p = self.search(pats[i].tobytes(), locate)
            counts[i] = c
            offs[i + 1] = offs[i] + len(p)
            chunks.append(p)
        pos = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint64)
        return counts, offs, pos.astype(np.uint64)

    def map_position(self, pos: int):
        s, o = C.c_uint32(0), C.c_uint64(0)
        lib().orc_map_position(self._h, int(pos), C.byref(s), C.byref(o))
        return int(s.value), int(o.value)

    def header_C(self) -> np.ndarray:
        out = np.zeros(int(self.info.compact_alphabet) + 1, dtype=np.uint64)
        lib().orc_get_C(self._h, out.ctypes.data)
        return out

    def header_terminators(self) -> np.ndarray:
        out = np.zeros(int(self.info.n_sequences), dtype=np.int64)
        lib().orc_get_terminators(self._h, out.ctypes.data)
        return out

    def extract_snippet(self, frm: int, to: int) -> bytes:
        """EFMIndex::extractSnippet(from, to)"""
        cap = max(0, to - frm + 1) + 64
        buf = C.create_string_buffer(cap)
        n = lib().orc_extract(self._h, frm & (2**64 - 1), to & (2**64 - 1), buf, cap)
        if n < 0:
            raise RuntimeError("The index doesn't contain marked rows: the extraction operation is not supported")
        return buf.raw[:n]

    def bwt(self, row0: int, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint32)
        lib().orc_get_bwt(self._h, row0, count, out.ctypes.data)
        return out

    def marks(self, row0: int, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint64)
        lib().orc_get_marks(self._h, row0, count, out.ctypes.data)
        return out

    def lf_rows(self, rows) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(len(rows), dtype=np.uint64)
        lib().orc_lf_rows(self._h, rows.ctypes.data, len(rows), out.ctypes.data)
        return out

    def rank_rows(self, codes, rows) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(len(rows), dtype=np.uint64)
        lib().orc_rank_rows(self._h, codes.ctypes.data, rows.ctypes.data, len(rows), out.ctypes.data)
        return out

    # ---- build side (bucket text encoder)
    def bucket_info(self, bucket: int) -> dict:
        out = np.zeros(8, dtype=np.uint64)
        lib().orc_bucket_info(self._h, bucket, out.ctypes.data)
        return dict(zip(["length", "alpha", "is_odd", "is_first", "clen", "text_bit", "tbits"], [int(x) for x in out[:7]]))

    def bucket_codes(self, bucket: int) -> np.ndarray:
        """rows of the bucket as bucket codes (sbwt[] after both remappings, EFMIndex.cpp:841-858)"""
        out = np.zeros(self.bucket_info(bucket)["length"], dtype=np.uint32)
        lib().orc_bucket_codes(self._h, bucket, out.ctypes.data)
        return out

    def counters(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        lib().orc_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def reset_counters(self):
        lib().orc_reset_counters(self._h)


# ------------------------------------------------------------------ the compiled reference

def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def ref_build(fasta: str, index: str, keyfile: str, k: int = 4, bucket_size: int = 4096, mrp: int = 2,
              threads: int = 8) -> None:
    """Index construction by the unmodified reference (EFMCollection::build)."""
    r = subprocess.run([REF_BIN, "build", fasta, index, keyfile, str(k), str(bucket_size), str(mrp), str(threads)],
                       capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(index):
        raise RuntimeError("reference build failed: " + r.stderr[-500:])


def read_result_file(path: str):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:32], dtype=np.uint64)
    n, mode, tot = int(hdr[1]), int(hdr[2]), int(hdr[3])
    o = 32
    counts = np.frombuffer(raw, dtype=np.uint64, count=n, offset=o); o += 8 * n
    nocc = np.frombuffer(raw, dtype=np.uint64, count=n, offset=o); o += 8 * n
    gpos = np.frombuffer(raw, dtype=np.uint64, count=tot, offset=o); o += 8 * tot
    seq = np.frombuffer(raw, dtype=np.uint32, count=tot, offset=o); o += 4 * tot
    off = np.frombuffer(raw, dtype=np.uint64, count=tot, offset=o)
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(nocc, out=offs[1:])
    return {"mode": mode, "counts": counts, "pos_offsets": offs, "positions": gpos, "seq": seq, "off": off}


def ref_search(index: str, keyfile: str, patterns_file: str, out_file: str, locate: bool = True,
               threads: int = 1, passes: int = 1):
    """Run the reference's own search over a patterns file; returns (results dict, timing dict)."""
    r = subprocess.run([REF_BIN, "search", index, keyfile, patterns_file, out_file, "1" if locate else "0",
                        str(threads), str(passes)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference search failed: " + r.stderr[-500:])
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return read_result_file(out_file), json.loads(line)


def ref_dump(index: str, keyfile: str, out_file: str) -> None:
    r = subprocess.run([REF_BIN, "dump", index, keyfile, out_file], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference dump failed: " + r.stderr[-500:])


def ref_extract(index: str, keyfile: str, queries_file: str, out_file: str):
    r = subprocess.run([REF_BIN, "extract", index, keyfile, queries_file, out_file], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference extract failed: " + r.stderr[-500:])
    with open(out_file, "rb") as f:
        return f.read().split(b"\n")[:-1]


def ref_kat() -> dict:
    r = subprocess.run([REF_BIN, "kat"], capture_output=True, text=True, check=True)
    return dict(l.split("=", 1) for l in r.stdout.splitlines() if "=" in l)


def encode_bucket(key64: bytes, bucket_number: int, codes: np.ndarray, alpha: int, is_odd: bool) -> np.ndarray:
    """Bucket::encodeText (Bucket.cpp:442-536): the encrypted MTF/RLE0 sequence of one bucket"""
    codes = np.ascontiguousarray(codes, dtype=np.uint32)
    out = np.zeros(max(len(codes), 1), dtype=np.uint32)
    n = lib().orc_encode_bucket(key64[32:64], bucket_number, codes.ctypes.data, len(codes), alpha, 1 if is_odd else 0, out.ctypes.data)
    return out[:int(n)]


def pack_sequence(seq: np.ndarray, bits: int) -> bytes:
    """the symbol loop of Bucket::writeText (Bucket.cpp:232-246) + flush"""
    seq = np.ascontiguousarray(seq, dtype=np.uint32)
    out = np.zeros((len(seq) * bits + 7) // 8 + 1, dtype=np.uint8)
    lib().orc_pack_sequence(seq.ctypes.data, len(seq), bits, out.ctypes.data)
    return out[:(len(seq) * bits + 7) // 8].tobytes()


def decode_sequence(key64: bytes, bucket_number: int, seq: np.ndarray, length: int, alpha: int, is_odd: bool) -> np.ndarray:
    """Bucket::readTextUntilTo (Bucket.cpp:288-403) over a sequence in memory: the rows (bucket codes) an encoded bucket decodes to"""
    seq = np.ascontiguousarray(seq, dtype=np.uint32)
    out = np.full(max(length, 1), 0xFFFFFFFF, dtype=np.uint32)
    n = lib().orc_decode_sequence(key64[32:64], bucket_number, seq.ctypes.data, len(seq), length, alpha, 1 if is_odd else 0, out.ctypes.data)
    assert int(n) == length, f"decoded {n} of {length} rows"
    return out[:length]

