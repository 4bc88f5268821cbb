"""Host-side handles over the C-ABI, named after the reference interfaces they stand for.

  Relation           a columnar relation image: pages + columnar.stripe / columnar.chunk rows
  ScanDesc helpers   ColumnarBeginRead's projectedColumnList / qualConditions + the worker
                     half of the aggregates (columnar.h:251-258, multi_logical_optimizer.c:3160)
  GpuColumnarAgg     the  Agg <- ColumnarScan  subtree of one worker task,
                     executed by the fused kernel; also the coordinator-side combine
  worker_partition_query_result   executor/partitioned_intermediate_results.c:115-298 (map side)

Everything computes on the GPU through libcitus_gpu.so; nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import (CG_AGG_COUNT, CG_AGG_COUNT_STAR, CG_AGG_MAX, CG_AGG_MIN, CG_AGG_SUM, CG_TYPE_FLOAT,
                   CG_TYPE_INT, CgAggSpec, CgColumnDesc, CgGenColumn, CgQual, CgRelation, CgScanDesc,
                   CgScanStats, CgSkipNode, CgStripe, check, lib)

_initialised = False
_on_torch_stream = None      # handle of the torch stream the library was put on (use_torch_stream)


def init(device: int = 0):
    global _initialised
    check(lib().cg_init(device))
    _initialised = True


def set_option(name: str, value: int):
    """cg_set_option: "jit" 0/1/2, "force_general" 0/1"""
    check(lib().cg_set_option(name.encode(), value))


def numa_bind() -> int:
    """pin this thread (and the staging/generator threads created after it) to the CPUs of the
    device's NUMA node; returns the node or -1"""
    node = C.c_int32(-1)
    check(lib().cg_numa_bind(C.byref(node)))
    return node.value


def numa_unbind():
    check(lib().cg_numa_unbind())


def use_torch_stream():
    """run the library's kernels on torch's current stream, so that torch ops, NCCL collectives and
    the library's kernels are ordered in one queue (torch's default stream is CUDA's legacy default
    stream, whose handle is 0: it is passed as cudaStreamLegacy = 1)"""
    import torch
    global _on_torch_stream
    h = torch.cuda.current_stream().cuda_stream
    check(lib().cg_set_stream(C.c_void_p(h if h else 1)))
    _on_torch_stream = h


class Relation:
    """Owns (or borrows) a relation image and exposes it as a CgRelation."""

    def __init__(self):
        self._gen = None
        self._keep = []
        self.view = CgRelation()

    # -- constructors -------------------------------------------------------
    @classmethod
    def generate(cls, columns, nrows, seed, first_row=0, stripe_row_limit=150000, chunk_row_limit=10000,
                 nthreads=8):
        """columns: [(attlen, kind, lo, hi, null_ppm)] -- splitmix64 counter-based synthetic data"""
        self = cls()
        arr = (CgGenColumn * len(columns))()
        for i, (attlen, kind, lo, hi, null_ppm) in enumerate(columns):
            arr[i].attlen, arr[i].kind, arr[i].lo, arr[i].hi, arr[i].null_ppm = attlen, kind, lo, hi, null_ppm
        h = C.c_void_p()
        check(lib().cg_gen_relation(arr, len(columns), nrows, first_row, seed, stripe_row_limit,
                                    chunk_row_limit, nthreads, C.byref(h)))
        self._gen = h
        check(lib().cg_gen_relation_view(h, C.byref(self.view)))
        return self

    @classmethod
    def write(cls, attlens, values, nulls=None, type_classes=None, stripe_row_limit=150000, chunk_row_limit=10000):
        """encode caller-supplied columns (int64 arrays; float64 arrays for float columns)"""
        self = cls()
        natts = len(attlens)
        type_classes = type_classes or [CG_TYPE_INT] * natts
        cols = (CgColumnDesc * natts)()
        for i in range(natts):
            cols[i].attlen, cols[i].type_class = attlens[i], type_classes[i]
        vals = []
        for i, v in enumerate(values):
            if type_classes[i] == CG_TYPE_FLOAT:
                vals.append(np.ascontiguousarray(v, np.float64).view(np.int64))
            else:
                vals.append(np.ascontiguousarray(v, np.int64))
        n = vals[0].shape[0] if vals else 0
        vp = (C.c_void_p * natts)(*[v.ctypes.data for v in vals])
        np_ = None
        nk = []
        if nulls is not None:
            nk = [None if x is None else np.ascontiguousarray(x, np.uint8) for x in nulls]
            np_ = (C.c_void_p * natts)(*[None if x is None else x.ctypes.data for x in nk])
        h = C.c_void_p()
        check(lib().cg_write_relation(cols, natts, vp, np_, n, stripe_row_limit, chunk_row_limit, C.byref(h)))
        self._gen = h
        check(lib().cg_gen_relation_view(h, C.byref(self.view)))
        return self

    @classmethod
    def from_image(cls, pages: np.ndarray, stripes_bytes: np.ndarray, nodes_bytes: np.ndarray, attlens,
                   type_classes=None):
        """borrow an image produced elsewhere (e.g. by the oracle's writer in tests)"""
        self = cls()
        natts = len(attlens)
        type_classes = type_classes or [CG_TYPE_INT] * natts
        cols = (CgColumnDesc * natts)()
        for i in range(natts):
            cols[i].attlen, cols[i].type_class = attlens[i], type_classes[i]
        pages = np.ascontiguousarray(pages, np.uint8)
        stripes_bytes = np.ascontiguousarray(stripes_bytes, np.uint8)
        nodes_bytes = np.ascontiguousarray(nodes_bytes, np.uint8)
        self._keep = [pages, stripes_bytes, nodes_bytes, cols]
        v = self.view
        v.pages = pages.ctypes.data
        v.nblocks = pages.shape[0] // 8192
        v.stripes = C.cast(stripes_bytes.ctypes.data, C.POINTER(CgStripe))
        v.nstripes = stripes_bytes.shape[0] // C.sizeof(CgStripe)
        v.nodes = C.cast(nodes_bytes.ctypes.data, C.POINTER(CgSkipNode))
        v.nnodes = nodes_bytes.shape[0] // C.sizeof(CgSkipNode)
        v.columns = cols
        v.natts = natts
        return self

    def prefix(self, nstripes: int) -> "Relation":
        """the first `nstripes` stripes of this relation as a relation of its own (shares the image)"""
        r = Relation()
        r._keep = [self]
        C.memmove(C.byref(r.view), C.byref(self.view), C.sizeof(CgRelation))
        r.view.nstripes = max(0, min(nstripes, self.view.nstripes))
        return r

    def register(self):
        """pin the page image (cudaHostRegister): cg_scan_relation then uses DMA staging"""
        check(lib().cg_relation_register(C.byref(self.view)))
        self._registered = True

    def unregister(self):
        if getattr(self, "_registered", False):
            lib().cg_relation_unregister(C.byref(self.view))
            self._registered = False

    def __del__(self):
        try:
            self.unregister()
            if getattr(self, "_gen", None):
                lib().cg_gen_relation_free(self._gen)
                self._gen = None
        except TypeError:           # interpreter shutdown: module globals are already gone
            pass

    # -- accessors ----------------------------------------------------------
    @property
    def natts(self):
        return self.view.natts

    @property
    def rows(self):
        return sum(self.view.stripes[i].row_count for i in range(self.view.nstripes))

    def pages(self) -> np.ndarray:
        n = self.view.nblocks * 8192
        return np.ctypeslib.as_array(C.cast(self.view.pages, C.POINTER(C.c_uint8)), shape=(n,))

    def stripes_bytes(self) -> np.ndarray:
        n = self.view.nstripes * C.sizeof(CgStripe)
        return np.frombuffer(C.string_at(self.view.stripes, n), np.uint8).copy()

    def nodes_bytes(self) -> np.ndarray:
        n = self.view.nnodes * C.sizeof(CgSkipNode)
        return np.frombuffer(C.string_at(self.view.nodes, n), np.uint8).copy()

    def column_descs(self):
        return [(self.view.columns[i].attlen, self.view.columns[i].type_class) for i in range(self.view.natts)]


# ---------------------------------------------------------------------------
# query description helpers
# ---------------------------------------------------------------------------
@dataclass
class Agg:
    kind: int
    factors: list = field(default_factory=list)    # [(column, a, b)]
    is_float: bool = False
    term_abs_bound: int = 0


def count_star():
    return Agg(CG_AGG_COUNT_STAR)


def count(col):
    return Agg(CG_AGG_COUNT, [(col, 0, 1)])


def sum_(col, is_float=False):
    return Agg(CG_AGG_SUM, [(col, 0, 1)], is_float)


def min_(col, is_float=False):
    return Agg(CG_AGG_MIN, [(col, 0, 1)], is_float)


def max_(col, is_float=False):
    return Agg(CG_AGG_MAX, [(col, 0, 1)], is_float)


def flatten_where(where):
    """WHERE as a nested tuple tree -- ("and" | "or", arm, arm, ...) over atoms (column, op, constant); a plain
    list of atoms is their AND -- to (atoms, postfix tokens).  The tokens are empty for a plain AND-list."""
    if not (isinstance(where, tuple) and where and where[0] in ("and", "or")):
        return list(where), []
    atoms, tokens = [], []

    def walk(node):
        if isinstance(node, tuple) and node and node[0] in ("and", "or"):
            if len(node) < 2:
                raise ValueError("empty boolean node")
            walk(node[1])
            for arm in node[2:]:
                walk(arm)
                tokens.append(capi.CG_QX_AND if node[0] == "and" else capi.CG_QX_OR)
        else:
            tokens.append(len(atoms))
            atoms.append(tuple(node))

    walk(where)
    return atoms, tokens


def make_desc(quals=(), group_cols=(), aggs=(), qual_pushdown=True, expected_groups=0, float_cols=()):
    d = CgScanDesc()
    quals, tokens = flatten_where(quals)
    if len(quals) > capi.CG_MAX_QUALS or len(tokens) > capi.CG_MAX_QEXPR:
        raise capi.CitusGpuError(capi.CG_EUNSUPPORTED, "WHERE tree too large")
    d.nqual_expr = len(tokens)
    for i, t in enumerate(tokens):
        d.qual_expr[i] = t
    d.nquals = len(quals)
    for i, (col, op, k) in enumerate(quals):
        d.quals[i].column = col
        d.quals[i].op = capi.CG_OP[op]
        d.quals[i].konst = int(np.float64(k).view(np.int64)) if col in float_cols else int(k)
    d.enable_qual_pushdown = 1 if qual_pushdown else 0
    d.ngroup_cols = len(group_cols)
    for i, c in enumerate(group_cols):
        d.group_cols[i] = c
    d.naggs = len(aggs)
    for i, a in enumerate(aggs):
        s = d.aggs[i]
        s.kind = a.kind
        s.nfactors = len(a.factors)
        s.is_float = 1 if a.is_float else 0
        s.term_abs_bound = a.term_abs_bound
        for f, (col, aa, bb) in enumerate(a.factors):
            s.column[f] = col
            if a.is_float:
                s.a[f] = int(np.float64(aa).view(np.int64))
                s.b[f] = int(np.float64(bb).view(np.int64))
            else:
                s.a[f], s.b[f] = aa, bb
    d.expected_groups = expected_groups
    return d


def relation_bounds(rel: Relation, desc: CgScanDesc):
    """exact key range / |sum argument| bounds / row count from the skip lists"""
    kmin, kmax, rows = C.c_int64(), C.c_int64(), C.c_int64()
    bounds = (C.c_int64 * capi.CG_MAX_AGGS)()
    check(lib().cg_relation_bounds(C.byref(rel.view), C.byref(desc), C.byref(kmin), C.byref(kmax), bounds, C.byref(rows)))
    return kmin.value, kmax.value, [bounds[i] for i in range(desc.naggs)], rows.value


class Shard:
    """A relation's projected columns staged in HBM (cg_shard_stage)."""

    def __init__(self, rel: Relation, columns=None):
        self.rel = rel
        h = C.c_void_p()
        if columns is None:
            check(lib().cg_shard_stage(C.byref(rel.view), None, 0, C.byref(h)))
        else:
            arr = (C.c_int32 * len(columns))(*columns)
            check(lib().cg_shard_stage(C.byref(rel.view), arr, len(columns), C.byref(h)))
        self.h = h

    def free(self):
        if getattr(self, "h", None):
            lib().cg_shard_free(self.h)
            self.h = None

    __del__ = free

    @property
    def device_bytes(self):
        return lib().cg_shard_device_bytes(self.h)

    @property
    def rows(self):
        return lib().cg_shard_rows(self.h)


class GpuColumnarAgg:
    """Partial aggregate of one worker task (or of all the tasks of one GPU), and the combine."""

    def __init__(self, desc: CgScanDesc, column_descs, key_min=0, key_max=-1, max_rows=0):
        self.desc = desc
        natts = len(column_descs)
        self.cols = (CgColumnDesc * natts)()
        for i, (l, t) in enumerate(column_descs):
            self.cols[i].attlen, self.cols[i].type_class = l, t
        h = C.c_void_p()
        check(lib().cg_partial_create(C.byref(desc), self.cols, natts, key_min, key_max, max_rows, C.byref(h)))
        self.h = h
        self.naggs = desc.naggs

    def free(self):
        if getattr(self, "h", None):
            lib().cg_partial_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except TypeError:           # interpreter shutdown
            pass

    def reset(self):
        check(lib().cg_partial_reset(self.h))

    def set_packing(self, enable: bool):
        check(lib().cg_partial_set_packing(self.h, 1 if enable else 0))

    def scan_shard(self, shard: Shard, want_stats=True):
        st = CgScanStats()
        check(lib().cg_scan_shard(shard.h, C.byref(self.desc), self.h, C.byref(st) if want_stats else None))
        return st

    def scan_relation(self, rel: Relation, want_stats=True):
        st = CgScanStats()
        check(lib().cg_scan_relation(C.byref(rel.view), C.byref(self.desc), self.h, C.byref(st) if want_stats else None))
        return st

    def ngroups(self):
        n = C.c_int64()
        check(lib().cg_partial_ngroups(self.h, C.byref(n)))
        return n.value

    def layout(self):
        nw, dense, cap = C.c_int32(), C.c_int32(), C.c_int64()
        ops = (C.c_int32 * 32)()
        check(lib().cg_partial_layout(self.h, C.byref(nw), ops, C.byref(dense), C.byref(cap)))
        return nw.value, [ops[i] for i in range(nw.value)], bool(dense.value), cap.value

    def fetch(self, reuse=False):
        """dict: keys, key_nulls, sum (python ints [ngroups][naggs]), count, minmax, fsum.
        reuse=True returns views of buffers owned by this object (overwritten by the next fetch):
        what a caller with its own result memory -- a tuplestore -- does"""
        na = max(self.naggs, 1)
        bufs = getattr(self, "_fetch_bufs", None) if reuse else None
        cap = bufs[0].shape[0] if bufs else 0
        if not bufs:
            cap = max(self.ngroups(), 1)
        while True:
            if not bufs or bufs[0].shape[0] < cap:
                bufs = (np.zeros(cap, np.int64), np.zeros(cap, np.uint8), np.zeros(cap * na, np.int64),
                        np.zeros(cap * na, np.uint64), np.zeros(cap * na, np.int64), np.zeros(cap * na, np.int64),
                        np.zeros(cap * na, np.float64))
            keys, kn, hi, lo, cnt, mm, fs = bufs
            got = C.c_int64()
            rc = lib().cg_partial_fetch(self.h, keys.shape[0], keys.ctypes.data, kn.ctypes.data, hi.ctypes.data, lo.ctypes.data,
                                        cnt.ctypes.data, mm.ctypes.data, fs.ctypes.data, C.byref(got))
            if rc != 0 and reuse and got.value > keys.shape[0]:      # cached buffers too small for this result: grow once
                cap, bufs = got.value, None
                continue
            check(rc)
            break
        if reuse:
            self._fetch_bufs = bufs
        n = got.value
        return dict(n=n, keys=keys[:n], key_nulls=kn[:n], sum_hi=hi[:n * na].reshape(n, na),
                    sum_lo=lo[:n * na].reshape(n, na), count=cnt[:n * na].reshape(n, na),
                    minmax=mm[:n * na].reshape(n, na), fsum=fs[:n * na].reshape(n, na))

    def groups(self):
        """same shape as oracle.Result.groups(): key (None = NULL group) -> per-aggregate dicts"""
        f = self.fetch()
        out = {}
        for i in range(f["n"]):
            per = []
            for a in range(self.naggs):
                s = (int(f["sum_hi"][i, a]) << 64) + int(f["sum_lo"][i, a])
                per.append(dict(sum=s, count=int(f["count"][i, a]), minmax=int(f["minmax"][i, a]),
                                fsum=float(f["fsum"][i, a])))
            out[None if f["key_nulls"][i] else int(f["keys"][i])] = per
        return out

    # -- combine (device arrays are torch tensors or any object with data_ptr()) ---------
    def export_device(self, d_keys_ptr, d_nulls_ptr, d_words_ptr, capacity):
        n = C.c_int64()
        check(lib().cg_partial_export_device(self.h, capacity, d_keys_ptr, d_nulls_ptr, d_words_ptr, C.byref(n)))
        return n.value

    def merge_rows(self, d_keys_ptr, d_nulls_ptr, d_words_ptr, nrows):
        check(lib().cg_partial_merge_rows(self.h, d_keys_ptr, d_nulls_ptr, d_words_ptr, nrows))

    def dense_words_enqueue(self):
        """like dense_words, but only enqueues the table maintenance (no host sync, no error check)"""
        p, total, stride = C.c_void_p(), C.c_int64(), C.c_int32()
        check(lib().cg_partial_dense_words_enqueue(self.h, C.byref(p), C.byref(total), C.byref(stride)))
        return p.value, total.value, stride.value

    def check(self):
        """wait for the partial's pending work and raise what its kernels flagged"""
        check(lib().cg_partial_check(self.h))

    def dense_words(self):
        p, total, stride = C.c_void_p(), C.c_int64(), C.c_int32()
        check(lib().cg_partial_dense_words(self.h, C.byref(p), C.byref(total), C.byref(stride)))
        return p.value, total.value, stride.value


def numeric_out(value: int, scale: int) -> str:
    buf = C.create_string_buffer(128)
    check(lib().cg_numeric_out((value >> 64), value & ((1 << 64) - 1), scale, buf, 128))
    return buf.value.decode()


def numeric_div_out(value: int, scale: int, count: int) -> str:
    buf = C.create_string_buffer(160)
    check(lib().cg_numeric_div_out((value >> 64), value & ((1 << 64) - 1), scale, count, buf, 160))
    return buf.value.decode()


def worker_partition_query_result(d_keys_ptr, d_nulls_ptr, n, key_len, method, mins, maxs, d_index_ptr, d_counts_ptr):
    """partition index of every row (device arrays).  method 'hash' | 'range'."""
    if method not in ("hash", "range"):
        raise capi.CitusGpuError(capi.CG_EINVAL, "only hash and range partitiong schemes are supported")
    mins = np.ascontiguousarray(mins, np.int32)
    maxs = np.ascontiguousarray(maxs, np.int32)
    if len(mins) != len(maxs):
        raise capi.CitusGpuError(capi.CG_EINVAL, "min values and max values must have the same number of elements")
    check(lib().cg_partition_index(d_keys_ptr, d_nulls_ptr, n, key_len, 1 if method == "hash" else 0,
                                   mins.ctypes.data, maxs.ctypes.data, len(mins), d_index_ptr, d_counts_ptr))


def join_count_sum(d_build_keys, d_build_payload, nbuild, d_probe_keys, d_probe_payload, nprobe, d_build_nulls=None,
                   d_probe_nulls=None):
    """(joined rows, exact sum(b.payload + p.payload)) of the equi-join of two device-resident partitions"""
    rows, hi, lo = C.c_int64(), C.c_int64(), C.c_uint64()
    check(lib().cg_join_count_sum(d_build_keys, d_build_nulls, d_build_payload, nbuild, d_probe_keys, d_probe_nulls,
                                  d_probe_payload, nprobe, C.byref(rows), C.byref(hi), C.byref(lo)))
    return rows.value, (hi.value << 64) + lo.value


def join_rows(d_build_keys, d_build_payload, nbuild, d_probe_keys, d_probe_payload, nprobe, capacity=0, d_out=(None, None, None),
              d_build_nulls=None, d_probe_nulls=None):
    """rows of the equi-join into the caller's device arrays (key, build payload, probe payload); returns the row count
    (capacity 0: count only)"""
    n = C.c_int64()
    check(lib().cg_join_rows(d_build_keys, d_build_nulls, d_build_payload, nbuild, d_probe_keys, d_probe_nulls, d_probe_payload, nprobe,
                             capacity, d_out[0], d_out[1], d_out[2], C.byref(n)))
    return n.value


def set_writer_compression(name: str):
    """columnar.compression for Relation.generate / Relation.write: "none", "lz4" or "zstd" (level 3)"""
    check(lib().cg_gen_set_compression({"none": 0, "lz4": 2, "zstd": 3}[name]))


def partition_copy_bytes(d_index_ptr, n, P, d_col_ptrs, col_lens, binary, d_null_ptrs=None, generate_empty_results=False):
    """(rows_written[P], bytes_written[P]) of worker_partition_query_result's return rows"""
    cols = (C.c_void_p * len(d_col_ptrs))(*d_col_ptrs)
    nulls = (C.c_void_p * len(d_col_ptrs))(*(d_null_ptrs or [None] * len(d_col_ptrs)))
    lens = (C.c_int32 * len(d_col_ptrs))(*col_lens)
    rows = np.zeros(P, np.int64)
    nbytes = np.zeros(P, np.int64)
    check(lib().cg_partition_copy_bytes(d_index_ptr, n, P, cols, nulls, lens, len(d_col_ptrs), 1 if binary else 0,
                                        1 if generate_empty_results else 0, rows.ctypes.data, nbytes.ctypes.data))
    return rows, nbytes


def partition_copy_serialize(d_index_ptr, n, P, d_col_ptrs, col_lens, binary, d_out_ptr, out_capacity, d_null_ptrs=None,
                             generate_empty_results=False):
    """the P partition files (COPY text / binary) back to back in the device buffer; returns file offsets [P + 1]"""
    cols = (C.c_void_p * len(d_col_ptrs))(*d_col_ptrs)
    nulls = (C.c_void_p * len(d_col_ptrs))(*(d_null_ptrs or [None] * len(d_col_ptrs)))
    lens = (C.c_int32 * len(d_col_ptrs))(*col_lens)
    offs = np.zeros(P + 1, np.int64)
    check(lib().cg_partition_copy_serialize(d_index_ptr, n, P, cols, nulls, lens, len(d_col_ptrs), 1 if binary else 0,
                                            1 if generate_empty_results else 0, d_out_ptr, out_capacity, offs.ctypes.data))
    return offs


def partition_scatter(d_index_ptr, n, P, d_col_ptrs, d_out_ptrs, order=None):
    """stable scatter into partition-contiguous order; order[p] = output position of partition p"""
    cols = (C.c_void_p * len(d_col_ptrs))(*d_col_ptrs)
    outs = (C.c_void_p * len(d_out_ptrs))(*d_out_ptrs)
    offs = np.zeros(P + 1, np.int64)
    if order is None:
        check(lib().cg_partition_scatter(d_index_ptr, n, P, cols, len(d_col_ptrs), outs, offs.ctypes.data))
    else:
        order = np.ascontiguousarray(order, np.int32)
        check(lib().cg_partition_scatter_ordered(d_index_ptr, n, P, order.ctypes.data, cols, len(d_col_ptrs), outs,
                                                 offs.ctypes.data))
    return offs
