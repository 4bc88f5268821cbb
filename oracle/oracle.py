"""ctypes binding of oracle/liboracle.so plus the numeric (decimal) helpers.

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never
from citus_b200/.
"""
from __future__ import annotations

import ctypes as C
import decimal
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

COMP_NONE, COMP_PGLZ, COMP_LZ4, COMP_ZSTD = 0, 1, 2, 3
T_INT, T_FLOAT = 0, 1
OP_LT, OP_LE, OP_EQ, OP_GE, OP_GT, OP_NE = range(6)
AGG_COUNT_STAR, AGG_COUNT, AGG_SUM, AGG_MIN, AGG_MAX = range(5)
OPS = {"<": OP_LT, "<=": OP_LE, "=": OP_EQ, ">=": OP_GE, ">": OP_GT, "<>": OP_NE}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


class SkipNode(C.Structure):
    _fields_ = [
        ("has_minmax", C.c_int32), ("compression_type", C.c_int32),
        ("min_value", C.c_int64), ("max_value", C.c_int64),
        ("row_count", C.c_uint64), ("value_offset", C.c_uint64), ("value_length", C.c_uint64),
        ("exists_offset", C.c_uint64), ("exists_length", C.c_uint64),
        ("decompressed_size", C.c_uint64), ("compression_level", C.c_int32), ("pad", C.c_int32),
    ]


class Stripe(C.Structure):
    _fields_ = [
        ("id", C.c_uint64), ("file_offset", C.c_uint64), ("data_length", C.c_uint64),
        ("row_count", C.c_uint64), ("first_row_number", C.c_uint64),
        ("column_count", C.c_uint32), ("chunk_row_count", C.c_uint32),
        ("chunk_count", C.c_uint32), ("skipnode_base", C.c_uint32),
    ]


class Qual(C.Structure):
    _fields_ = [("col", C.c_int32), ("op", C.c_int32), ("konst", C.c_int64)]


class AggSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nfactors", C.c_int32), ("col", C.c_int32 * 3),
                ("is_float", C.c_int32), ("a", C.c_int64 * 3), ("b", C.c_int64 * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_hash_bytes_uint32.restype = C.c_uint32
        L.orc_hash_bytes_uint32.argtypes = [C.c_uint32]
        L.orc_hashint4.restype = C.c_int32
        L.orc_hashint4.argtypes = [C.c_int32]
        L.orc_hashint8.restype = C.c_int32
        L.orc_hashint8.argtypes = [C.c_int64]
        L.orc_synthetic_intervals.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_search_interval.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_uniform_hash_range_index.argtypes = [C.c_int32, C.c_int]
        L.orc_partition_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_char,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_copy_file_bytes.restype = C.c_int64
        L.orc_copy_file_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int, C.c_int64, C.c_int]
        L.orc_codec_compress.restype = C.c_int64
        L.orc_codec_compress.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_codec_decompress.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_table_create.restype = C.c_void_p
        L.orc_table_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int]
        L.orc_table_free.argtypes = [C.c_void_p]
        L.orc_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_table_pages.restype = C.c_void_p
        L.orc_table_pages.argtypes = [C.c_void_p]
        L.orc_table_nblocks.restype = C.c_uint64
        L.orc_table_nblocks.argtypes = [C.c_void_p]
        L.orc_table_nstripes.argtypes = [C.c_void_p]
        L.orc_table_stripes.restype = C.POINTER(Stripe)
        L.orc_table_stripes.argtypes = [C.c_void_p]
        L.orc_table_nnodes.argtypes = [C.c_void_p]
        L.orc_table_nodes.restype = C.POINTER(SkipNode)
        L.orc_table_nodes.argtypes = [C.c_void_p]
        L.orc_result_create.restype = C.c_void_p
        L.orc_result_create.argtypes = [C.c_int]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_result_ngroups.restype = C.c_int64
        L.orc_result_ngroups.argtypes = [C.c_void_p]
        L.orc_result_counter.restype = C.c_int64
        L.orc_result_counter.argtypes = [C.c_void_p, C.c_int]
        L.orc_result_group.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_result_agg.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 8
        L.orc_set_qual_expr.restype = None
        L.orc_set_qual_expr.argtypes = [C.c_void_p, C.c_int]
        L.orc_scan_aggregate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int, C.c_void_p]
        L.orc_combine.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_decode_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_storage_read.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_table_attach.restype = C.c_void_p
        L.orc_table_attach.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_splitmix64.argtypes = [C.c_uint64]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def hashint4(v: int) -> int:
    return lib().orc_hashint4(v)


def hashint8(v: int) -> int:
    return lib().orc_hashint8(v)


def synthetic_intervals(P: int):
    mins = np.zeros(P, np.int32)
    maxs = np.zeros(P, np.int32)
    lib().orc_synthetic_intervals(P, mins.ctypes.data, maxs.ctypes.data)
    return mins, maxs


def partition_rows(keys, nulls, keytype, method, mins, maxs):
    keys = np.ascontiguousarray(keys, np.int64)
    mins = np.ascontiguousarray(mins, np.int32)
    maxs = np.ascontiguousarray(maxs, np.int32)
    n = keys.shape[0]
    idx = np.zeros(n, np.int32)
    rows = np.zeros(len(mins), np.int64)
    nptr = None
    if nulls is not None:
        nulls = np.ascontiguousarray(nulls, np.uint8)
        nptr = nulls.ctypes.data
    _check(lib().orc_partition_rows(keys.ctypes.data, nptr, n, keytype, method.encode(),
                                    mins.ctypes.data, maxs.ctypes.data, len(mins),
                                    idx.ctypes.data, rows.ctypes.data))
    return idx, rows


def copy_files(index, cols, col_lens, P, binary, nulls=None, generate_empty_results=False):
    """the partition files TaskFileDestReceiver writes (worker/worker_sql_task_protocol.c:91-251; [PG] COPY text /
    binary row encoding, commands/multi_copy.c AppendCopyRowData / AppendCopyBinaryHeaders / -Footers), row at a time:
    text = decimal fields joined by TAB, NULL as \\N, LF per row; binary = 19-byte header, per row int16 field count +
    per field int32 length (-1 NULL) + big-endian value, int16 -1 trailer.  Returns P bytes objects."""
    import struct
    files = [bytearray() for _ in range(P)]
    started = [False] * P
    n = len(index)
    for r in range(n):
        p = int(index[r])
        f = files[p]
        if binary and not started[p]:
            f += b"PGCOPY\n\377\r\n\0" + struct.pack(">ii", 0, 0)
        started[p] = True
        if binary:
            f += struct.pack(">h", len(cols))
            for c, col in enumerate(cols):
                if nulls is not None and nulls[c] is not None and nulls[c][r]:
                    f += struct.pack(">i", -1)
                else:
                    f += struct.pack(">i", col_lens[c]) + int(col[r]).to_bytes(col_lens[c], "big", signed=True)
        else:
            fields = ["\\N" if (nulls is not None and nulls[c] is not None and nulls[c][r]) else str(int(col[r]))
                      for c, col in enumerate(cols)]
            f += ("\t".join(fields) + "\n").encode()
    for p in range(P):
        if binary and (started[p] or generate_empty_results):
            if not started[p]:
                files[p] += b"PGCOPY\n\377\r\n\0" + struct.pack(">ii", 0, 0)
            files[p] += struct.pack(">h", -1)
    return [bytes(f) for f in files]


def join_count_sum(bkeys, bpay, pkeys, ppay, bnulls=None, pnulls=None):
    """(joined rows, exact sum(b.payload + p.payload)) of the MERGE task's hash join, row at a time"""
    arrs = [np.ascontiguousarray(a, np.int64) for a in (bkeys, bpay, pkeys, ppay)]
    nl = [None if a is None else np.ascontiguousarray(a, np.uint8) for a in (bnulls, pnulls)]
    joined, hi, lo = C.c_int64(), C.c_int64(), C.c_uint64()
    L = lib()
    L.orc_join_count_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    _check(L.orc_join_count_sum(arrs[0].ctypes.data, None if nl[0] is None else nl[0].ctypes.data, arrs[1].ctypes.data, len(arrs[0]),
                                arrs[2].ctypes.data, None if nl[1] is None else nl[1].ctypes.data, arrs[3].ctypes.data, len(arrs[2]),
                                C.byref(joined), C.byref(hi), C.byref(lo)))
    return joined.value, (hi.value << 64) + lo.value


def copy_file_bytes(cols, collens, row_partition, partition, binary, colnulls=None):
    cols = [np.ascontiguousarray(c, np.int64) for c in cols]
    n = cols[0].shape[0]
    arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    lens = (C.c_int * len(cols))(*collens)
    nptr = None
    keep = []
    if colnulls is not None:
        keep = [None if x is None else np.ascontiguousarray(x, np.uint8) for x in colnulls]
        nptr = (C.c_void_p * len(cols))(*[None if x is None else x.ctypes.data for x in keep])
    rp = None
    if row_partition is not None:
        row_partition = np.ascontiguousarray(row_partition, np.int32)
        rp = row_partition.ctypes.data
    return lib().orc_copy_file_bytes(arr, nptr, lens, len(cols), rp, partition, n, 1 if binary else 0)


@dataclass
class Agg:
    """term = prod_i (a_i + b_i * col_i); kind in AGG_*"""
    kind: int
    factors: list = field(default_factory=list)  # [(col, a, b)]
    is_float: bool = False

    def spec(self) -> AggSpec:
        s = AggSpec()
        s.kind = self.kind
        s.nfactors = len(self.factors)
        s.is_float = 1 if self.is_float else 0
        for i, (col, a, b) in enumerate(self.factors):
            s.col[i] = col
            if self.is_float:
                s.a[i] = int(np.float64(a).view(np.int64))
                s.b[i] = int(np.float64(b).view(np.int64))
            else:
                s.a[i] = a
                s.b[i] = b
        return s


def count_star():
    return Agg(AGG_COUNT_STAR)


def count(col):
    return Agg(AGG_COUNT, [(col, 0, 1)])


def sum_(col, is_float=False):
    return Agg(AGG_SUM, [(col, 0, 1)], is_float)


def min_(col, is_float=False):
    return Agg(AGG_MIN, [(col, 0, 1)], is_float)


def max_(col, is_float=False):
    return Agg(AGG_MAX, [(col, 0, 1)], is_float)


def codec_compress(comp, data: bytes, level=3):
    """CompressBuffer of one value stream; None when the codec declines (stored raw)"""
    src = np.frombuffer(data, np.uint8)
    out = np.empty(len(data) * 2 + 1024, np.uint8)
    n = lib().orc_codec_compress(comp, level, src.ctypes.data if len(data) else None, len(data), out.ctypes.data, out.shape[0])
    if n < 0:
        raise OracleError("compressed output too large")
    return None if n == 0 else out[:n].tobytes()


def codec_decompress(comp, data: bytes, rawlen: int) -> bytes:
    """DecompressBuffer of one value stream"""
    src = np.frombuffer(data, np.uint8)
    out = np.zeros(max(rawlen, 1), np.uint8)
    if lib().orc_codec_decompress(comp, src.ctypes.data, len(data), rawlen, out.ctypes.data):
        raise OracleError(lib().orc_last_error().decode())
    return out[:rawlen].tobytes()


class Table:
    """A columnar relation image written by the oracle's row-at-a-time writer."""

    def __init__(self, attlen, atttype=None, stripe_row_limit=150000, chunk_row_limit=10000,
                 compression=COMP_NONE, compression_level=3):
        self.natts = len(attlen)
        self.attlen = list(attlen)
        self.atttype = list(atttype) if atttype is not None else [T_INT] * self.natts
        al = (C.c_int * self.natts)(*self.attlen)
        at = (C.c_int * self.natts)(*self.atttype)
        self.stripe_row_limit = stripe_row_limit
        self.chunk_row_limit = chunk_row_limit
        self.h = lib().orc_table_create(self.natts, al, at, stripe_row_limit, chunk_row_limit,
                                        compression, compression_level)

    @classmethod
    def attach(cls, pages, stripes_bytes, nodes_bytes, attlen, atttype=None, chunk_row_limit=10000):
        """read an image written by another writer (same struct layouts as the C-ABI)"""
        self = cls.__new__(cls)
        self.natts = len(attlen)
        self.attlen = list(attlen)
        self.atttype = list(atttype) if atttype is not None else [T_INT] * self.natts
        self.chunk_row_limit = chunk_row_limit
        pages = np.ascontiguousarray(pages, np.uint8)
        sb = np.ascontiguousarray(stripes_bytes, np.uint8)
        nb = np.ascontiguousarray(nodes_bytes, np.uint8)
        al = (C.c_int * self.natts)(*self.attlen)
        at = (C.c_int * self.natts)(*self.atttype)
        self.h = lib().orc_table_attach(pages.ctypes.data, pages.shape[0] // 8192, sb.ctypes.data,
                                        sb.shape[0] // C.sizeof(Stripe), nb.ctypes.data,
                                        nb.shape[0] // C.sizeof(SkipNode), self.natts, al, at, chunk_row_limit)
        return self

    @classmethod
    def attach_view(cls, pages_ptr, nblocks, stripes_ptr, nstripes, nodes_ptr, nnodes, attlen, atttype=None,
                    chunk_row_limit=10000):
        """borrow a resident image by raw pointers (no copy); nstripes may be a bounded sample"""
        self = cls.__new__(cls)
        self.natts = len(attlen)
        self.attlen = list(attlen)
        self.atttype = list(atttype) if atttype is not None else [T_INT] * self.natts
        self.chunk_row_limit = chunk_row_limit
        al = (C.c_int * self.natts)(*self.attlen)
        at = (C.c_int * self.natts)(*self.atttype)
        L = lib()
        L.orc_table_attach_view.restype = C.c_void_p
        L.orc_table_attach_view.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_uint32]
        self.h = L.orc_table_attach_view(pages_ptr, nblocks, stripes_ptr, nstripes, nodes_ptr, nnodes, self.natts,
                                         al, at, chunk_row_limit)
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_table_free(self.h)
            self.h = None

    def insert(self, cols, nulls=None):
        """cols: list of arrays (int64, or float64 for float columns); nulls: list of bool arrays or None"""
        arrs = []
        for c, a in enumerate(cols):
            if self.atttype[c] == T_FLOAT:
                arrs.append(np.ascontiguousarray(a, np.float64).view(np.int64))
            else:
                arrs.append(np.ascontiguousarray(a, np.int64))
        n = arrs[0].shape[0]
        cp = (C.c_void_p * self.natts)(*[a.ctypes.data for a in arrs])
        np_ = None
        keep = []
        if nulls is not None:
            keep = [None if x is None else np.ascontiguousarray(x, np.uint8) for x in nulls]
            np_ = (C.c_void_p * self.natts)(*[None if x is None else x.ctypes.data for x in keep])
        lib().orc_insert(self.h, cp, np_, n)

    def generate(self, columns, nrows, seed, first_row=0):
        """synthetic rows through the row-at-a-time writer; columns = [(attlen, kind, lo, hi, null_ppm)], the same
        counter-based formula as the benchmark's generator (cg_gen_relation)"""
        arr = (GenColumn * len(columns))()
        for i, (attlen, kind, lo, hi, null_ppm) in enumerate(columns):
            arr[i].attlen, arr[i].kind, arr[i].lo, arr[i].hi, arr[i].null_ppm = attlen, kind, lo, hi, null_ppm
        L = lib()
        L.orc_gen_insert.restype = None
        L.orc_gen_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64]
        L.orc_gen_insert(self.h, arr, nrows, first_row, seed)

    # --- image accessors -------------------------------------------------
    def pages(self) -> np.ndarray:
        nb = lib().orc_table_nblocks(self.h)
        p = lib().orc_table_pages(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb * 8192,)).copy()

    def stripes(self):
        n = lib().orc_table_nstripes(self.h)
        p = lib().orc_table_stripes(self.h)
        return [p[i] for i in range(n)]

    def nodes(self):
        n = lib().orc_table_nnodes(self.h)
        p = lib().orc_table_nodes(self.h)
        return [p[i] for i in range(n)]

    def stripes_array(self) -> np.ndarray:
        n = lib().orc_table_nstripes(self.h)
        p = lib().orc_table_stripes(self.h)
        buf = C.string_at(p, n * C.sizeof(Stripe))
        return np.frombuffer(buf, dtype=np.uint8).copy()

    def nodes_array(self) -> np.ndarray:
        n = lib().orc_table_nnodes(self.h)
        p = lib().orc_table_nodes(self.h)
        buf = C.string_at(p, n * C.sizeof(SkipNode))
        return np.frombuffer(buf, dtype=np.uint8).copy()

    def decode_all(self):
        total = sum(s.row_count for s in self.stripes())
        vals = [np.zeros(total, np.int64) for _ in range(self.natts)]
        nulls = [np.zeros(total, np.uint8) for _ in range(self.natts)]
        vp = (C.c_void_p * self.natts)(*[v.ctypes.data for v in vals])
        npp = (C.c_void_p * self.natts)(*[v.ctypes.data for v in nulls])
        rows = C.c_int64(0)
        _check(lib().orc_decode_all(self.h, vp, npp, C.byref(rows)))
        assert rows.value == total
        for c in range(self.natts):
            if self.atttype[c] == T_FLOAT:
                vals[c] = vals[c].view(np.float64)
        return vals, nulls

    # --- scan ------------------------------------------------------------
    def scan(self, quals=(), group_cols=(), aggs=(), qual_pushdown=True, into: "Result | None" = None):
        """quals: [(col, op_str, const)] (an AND-list) or a nested tuple tree ("and" | "or", arm, ...) over such
        atoms; returns Result (accumulating into `into` if given)."""
        quals, tokens = flatten_where(quals)
        tk = (C.c_int8 * max(1, len(tokens)))(*tokens)
        lib().orc_set_qual_expr(tk, len(tokens))        # per thread, read by the scan below
        qa = (Qual * max(1, len(quals)))()
        for i, (col, op, k) in enumerate(quals):
            qa[i].col = col
            qa[i].op = OPS[op] if isinstance(op, str) else op
            if self.atttype[col] == T_FLOAT:
                qa[i].konst = int(np.float64(k).view(np.int64))
            else:
                qa[i].konst = int(k)
        ga = (C.c_int * max(1, len(group_cols)))(*group_cols)
        sa = (AggSpec * max(1, len(aggs)))(*[a.spec() for a in aggs])
        res = into if into is not None else Result(list(aggs))
        _check(lib().orc_scan_aggregate(self.h, qa, len(quals), 1 if qual_pushdown else 0,
                                        ga, len(group_cols), sa, len(aggs), res.h))
        return res


def flatten_where(where):
    """nested ("and" | "or", arm, ...) tree over (col, op, const) atoms -> (atoms, postfix tokens: >= 0 atom,
    -1 AND, -2 OR); a plain list of atoms is their AND and has no tokens"""
    if not (isinstance(where, tuple) and where and where[0] in ("and", "or")):
        return list(where), []
    atoms, tokens = [], []

    def walk(node):
        if isinstance(node, tuple) and node and node[0] in ("and", "or"):
            walk(node[1])
            for arm in node[2:]:
                walk(arm)
                tokens.append(-1 if node[0] == "and" else -2)
        else:
            tokens.append(len(atoms))
            atoms.append(tuple(node))

    walk(where)
    return atoms, tokens


class GenColumn(C.Structure):
    _fields_ = [("attlen", C.c_int32), ("kind", C.c_int32), ("lo", C.c_int64), ("hi", C.c_int64),
                ("null_ppm", C.c_uint32), ("reserved", C.c_uint32)]


class Result:
    def __init__(self, aggs):
        self.aggs = aggs
        self.h = lib().orc_result_create(len(aggs))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_result_free(self.h)
            self.h = None

    @property
    def rows_scanned(self):
        return lib().orc_result_counter(self.h, 0)

    @property
    def rows_removed_by_filter(self):
        return lib().orc_result_counter(self.h, 1)

    @property
    def chunk_groups_filtered(self):
        return lib().orc_result_counter(self.h, 2)

    @property
    def rows_passed(self):
        return lib().orc_result_counter(self.h, 3)

    def combine(self, other: "Result"):
        sa = (AggSpec * max(1, len(self.aggs)))(*[a.spec() for a in self.aggs])
        _check(lib().orc_combine(self.h, other.h, sa))

    def export_arrays(self):
        """all groups as numpy arrays: keys[n], key_nulls[n], sum_hi / sum_lo / count [n][naggs]"""
        n = lib().orc_result_ngroups(self.h)
        na = max(len(self.aggs), 1)
        keys, kn = np.zeros(n, np.int64), np.zeros(n, np.uint8)
        hi, lo, cnt = np.zeros((n, na), np.int64), np.zeros((n, na), np.uint64), np.zeros((n, na), np.int64)
        L = lib()
        L.orc_result_export.restype = None
        L.orc_result_export.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_result_export(self.h, keys.ctypes.data, kn.ctypes.data, hi.ctypes.data, lo.ctypes.data, cnt.ctypes.data)
        return dict(n=n, keys=keys, key_nulls=kn, sum_hi=hi, sum_lo=lo, count=cnt)

    def groups(self):
        """dict: key (int or None for the NULL group) -> list per aggregate of dicts
        {sum (python int), count, min, max, fsum, fmin, fmax}"""
        out = {}
        n = lib().orc_result_ngroups(self.h)
        key = C.c_int64()
        kn = C.c_int32()
        hi = C.c_int64()
        lo = C.c_uint64()
        cnt = C.c_int64()
        imin = C.c_int64()
        imax = C.c_int64()
        fsum = C.c_double()
        fmin = C.c_double()
        fmax = C.c_double()
        for g in range(n):
            lib().orc_result_group(self.h, g, C.byref(key), C.byref(kn))
            per = []
            for a in range(len(self.aggs)):
                lib().orc_result_agg(self.h, g, a, C.byref(hi), C.byref(lo), C.byref(cnt), C.byref(imin),
                                     C.byref(imax), C.byref(fsum), C.byref(fmin), C.byref(fmax))
                per.append(dict(sum=(hi.value << 64) + lo.value, count=cnt.value, min=imin.value,
                                max=imax.value, fsum=fsum.value, fmin=fmin.value, fmax=fmax.value))
            out[None if kn.value else key.value] = per
        return out

    def arrays(self):
        """sorted-by-key arrays: keys, per aggregate (sum_hi int64, sum_lo uint64, count)"""
        g = self.groups()
        keys = sorted(k for k in g if k is not None)
        return keys, g


# ----------------------------------------------------------------------------
# [PG] numeric helpers (numeric.c): output of sum/avg exactly as PostgreSQL prints
# them.  Pinned by expected/multi_tpch_query1.out, multi_tpch_query6.out,
# columnar_query.out:9-29, multi_agg_type_conversion.out.
# ----------------------------------------------------------------------------
NUMERIC_MIN_SIG_DIGITS = 16
DEC_DIGITS = 4
NUMERIC_MIN_DISPLAY_SCALE = 0
NUMERIC_MAX_DISPLAY_SCALE = 1000


def numeric_str(unscaled: int, scale: int) -> str:
    """text of the numeric unscaled * 10^-scale with display scale `scale`"""
    sign = "-" if unscaled < 0 else ""
    digits = str(abs(unscaled))
    if scale == 0:
        return sign + digits
    digits = digits.rjust(scale + 1, "0")
    return f"{sign}{digits[:-scale]}.{digits[-scale:]}"


def _weight_firstdigit(unscaled: int, scale: int):
    """weight (base-10000 exponent of the first nonzero digit) and that digit, like
    the loop at the top of select_div_scale"""
    if unscaled == 0:
        return 0, 0
    a = abs(unscaled)
    # align so that the decimal point sits on a base-10000 digit boundary
    pad = (-scale) % DEC_DIGITS
    a *= 10 ** pad
    frac_digits = (scale + pad) // DEC_DIGITS
    ndig = 0
    first = 0
    t = a
    while t:
        first = t % 10000
        t //= 10000
        ndig += 1
    return ndig - 1 - frac_digits, first


def select_div_scale(n1: int, s1: int, n2: int, s2: int) -> int:
    w1, f1 = _weight_firstdigit(n1, s1)
    w2, f2 = _weight_firstdigit(n2, s2)
    qweight = w1 - w2
    if f1 <= f2:
        qweight -= 1
    rscale = NUMERIC_MIN_SIG_DIGITS - qweight * DEC_DIGITS
    rscale = max(rscale, s1, s2, NUMERIC_MIN_DISPLAY_SCALE)
    return min(rscale, NUMERIC_MAX_DISPLAY_SCALE)


def numeric_div_str(n1: int, s1: int, n2: int, s2: int) -> str:
    """numeric_div: n1*10^-s1 / n2*10^-s2 rounded (half away from zero) at select_div_scale"""
    rscale = select_div_scale(n1, s1, n2, s2)
    with decimal.localcontext() as ctx:
        ctx.prec = 2000
        q = (decimal.Decimal(n1).scaleb(-s1)) / (decimal.Decimal(n2).scaleb(-s2))
        q = q.quantize(decimal.Decimal(1).scaleb(-rscale), rounding=decimal.ROUND_HALF_UP)
    return f"{q:f}"
