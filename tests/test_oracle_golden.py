"""Pins the CPU oracle (oracle/oracle.c, oracle/oracle.py) against the reference's own
golden vectors; fixtures built by tests/golden/make_golden.py."""
import numpy as np
import pytest

Q1_CUTOFF = "1998-09-02"   # date '1998-12-01' - interval '90 days'


def pgdate(s):
    import datetime
    y, m, d = map(int, s.split("-"))
    return (datetime.date(y, m, d) - datetime.date(2000, 1, 1)).days


# --------------------------------------------------------------------------- hash
def test_hashint4_goldens(oracle, expected):
    for k, v in expected["hashint4"].items():
        assert oracle.hashint4(int(k)) == v
    assert oracle.hashint4(123) == expected["worker_hash_123"]
    # worker_hash(date) hashes the date Datum (int4 days since 2000-01-01)
    assert oracle.hashint4(expected["date_1997_08_08"]) == expected["worker_hash_date_1997_08_08"]


def test_hashint8_goldens(oracle, expected):
    got = sorted(oracle.hashint8(x) for x in expected["hashint8_inputs"])
    assert got == expected["hashint8_sorted"]
    assert oracle.hashint8(2608474032) == 2147483647
    assert oracle.hashint8(963809240) == -2147483648
    # small non-negative int8 hashes like int4
    for k, v in expected["hashint4"].items():
        assert oracle.hashint8(int(k)) == v


# --------------------------------------------------------------------------- partition
def test_synthetic_intervals_match_reference_literals(oracle, expected):
    mins, maxs = oracle.synthetic_intervals(4)
    assert mins.tolist() == expected["squares_hash_mins"]
    assert maxs.tolist() == expected["squares_hash_maxs"]
    mins, maxs = oracle.synthetic_intervals(3)      # last range widened to INT32_MAX
    assert maxs[-1] == 2147483647 and mins[0] == -2147483648
    assert mins[1] == -2147483648 + 1431655765 and maxs[0] == mins[1] - 1


def test_squares_hash_partition_membership_and_bytes(oracle, expected):
    i = np.arange(1, 11, dtype=np.int64)
    idx, rows = oracle.partition_rows(i, None, 4, "h", expected["squares_hash_mins"], expected["squares_hash_maxs"])
    for p in range(4):
        assert i[idx == p].tolist() == expected["squares_hash_members"][p]
        part, nrows, nbytes = expected["squares_hash_text"][p]
        assert rows[p] == nrows
        assert oracle.copy_file_bytes([i, i * i], [4, 4], idx, p, binary=False) == nbytes


def test_squares_range_partition_binary_bytes(oracle, expected):
    i = np.arange(1, 11, dtype=np.int64)
    idx, rows = oracle.partition_rows(i * i, None, 4, "r", [0, 21, 41, 61], [20, 40, 60, 100])
    for p in range(4):
        part, nrows, nbytes = expected["squares_range_binary"][p]
        assert rows[p] == nrows
        assert oracle.copy_file_bytes([i, i * i], [4, 4], idx, p, binary=True) == nbytes


def test_million_row_partitions(oracle, expected):
    i = np.arange(1, 1000001, dtype=np.int64)
    idx, rows = oracle.partition_rows(i, None, 4, "h", expected["squares_hash_mins"], expected["squares_hash_maxs"])
    for p in range(4):
        part, nrows, nbytes = expected["doubles_hash_text"][p]
        assert rows[p] == nrows
        assert oracle.copy_file_bytes([i, i * 2], [4, 4], idx, p, binary=False) == nbytes
    idx, rows = oracle.partition_rows(i, None, 4, "r", [0, 250001, 500001, 750001], [250000, 500000, 750000, 1000000])
    for p in range(4):
        part, nrows, nbytes = expected["doubles_range_binary"][p]
        assert rows[p] == nrows
        assert oracle.copy_file_bytes([i, i * 2], [4, 4], idx, p, binary=True) == nbytes


def test_null_key_goes_to_partition_zero_and_uniform_index_agrees(oracle):
    keys = np.array([5, 6, 7], np.int64)
    mins, maxs = oracle.synthetic_intervals(32)
    idx, rows = oracle.partition_rows(keys, np.array([0, 1, 0], np.uint8), 8, "h", mins, maxs)
    assert idx[1] == 0
    rng = np.random.default_rng(1)
    ks = rng.integers(-2**62, 2**62, 5000)
    for P in (1, 3, 4, 7, 32):
        mins, maxs = oracle.synthetic_intervals(P)
        idx, _ = oracle.partition_rows(ks, None, 8, "h", mins, maxs)
        for k, i in zip(ks[:500], idx[:500]):
            assert oracle.lib().orc_uniform_hash_range_index(oracle.hashint8(int(k)), P) == i


# --------------------------------------------------------------------------- columnar format + skip list
def _filter_quals(where):
    where = where.replace("WHERE", "").strip()
    if not where:
        return []
    if "BETWEEN" in where:
        lo, hi = where.split("BETWEEN")[1].split("AND")
        return [(0, ">=", int(lo)), (0, "<=", int(hi))]
    col, op, k = where.split()
    return [(0, op, int(k))]


def test_chunk_filtering_counts(oracle, expected):
    # expected/columnar_chunk_filtering.out: stripe_row_limit 2000, chunk_group_row_limit 1000, a int
    t = oracle.Table([4], stripe_row_limit=2000, chunk_row_limit=1000)
    t.insert([np.arange(1, 10001)])
    cases = expected["chunk_filtering"]
    for where, want in cases[:9]:
        r = t.scan(_filter_quals(where), aggs=[oracle.count_star()])
        assert r.rows_removed_by_filter == want, where
    t.insert([np.arange(1, 10001)])     # "Load data for second time": a new stripe sequence
    assert len(t.stripes()) == 10
    for where, want in cases[9:]:
        r = t.scan(_filter_quals(where), aggs=[oracle.count_star()])
        assert r.rows_removed_by_filter == want, where


def test_simple_chunk_filtering_explain_counters(oracle, expected):
    for case in expected["simple_chunk_filtering"]:
        t = oracle.Table([4])          # defaults 150000 / 10000 (columnar.c:29-30)
        t.insert([np.arange(0, case["max"] + 1)])
        r = t.scan([(0, ">", case["gt"])], aggs=[oracle.count_star()])
        assert r.chunk_groups_filtered == case["groups_removed"]
        assert r.rows_removed_by_filter == case["rows_removed"]
        assert r.rows_passed == case["actual_rows"]
        if case["max"] == 234567:
            # SET columnar.enable_qual_pushdown = false -> Rows Removed by Filter: 123457
            r = t.scan([(0, ">", case["gt"])], aggs=[oracle.count_star()], qual_pushdown=False)
            assert r.rows_removed_by_filter == 123457 and r.chunk_groups_filtered == 0


def test_multi_column_chunk_filtering(oracle, expected):
    case = expected["multi_column_chunk_filtering"]
    i = np.arange(0, case["max"] + 1)
    t = oracle.Table([4, 4])
    t.insert([i, i + 1])
    for quals in ([(0, ">", 50000)], [(0, ">", 50000), (1, ">", 50000)]):
        r = t.scan(quals, aggs=[oracle.count_star()])
        assert (r.chunk_groups_filtered, r.rows_removed_by_filter, r.rows_passed) == \
            (case["groups_removed"], case["rows_removed"], case["actual_rows"])
    # INSERT ... SELECT generate_series(0,5): b is all NULL -> no min/max on b
    t = oracle.Table([4, 4])
    t.insert([np.arange(6), np.zeros(6)], nulls=[None, np.ones(6)])
    r = t.scan([(0, ">", 50000), (1, ">", 50000)], aggs=[oracle.count_star()])
    assert r.chunk_groups_filtered == 1 and r.rows_passed == 0
    r = t.scan([(1, ">", 50000)], aggs=[oracle.count_star()])
    assert r.chunk_groups_filtered == 0 and r.rows_removed_by_filter == 6 and r.rows_passed == 0


def test_stripe_layout_and_page_framing(oracle):
    rng = np.random.default_rng(7)
    n = 23456
    a = rng.integers(-2**40, 2**40, n)
    b = rng.integers(-100, 100, n)
    c = rng.integers(0, 2, n)
    nulls_b = rng.random(n) < 0.2
    t = oracle.Table([8, 4, 2], stripe_row_limit=5000, chunk_row_limit=1000)
    t.insert([a, b, c], nulls=[None, nulls_b, None])
    stripes = t.stripes()
    nodes = t.nodes()
    assert sum(s.row_count for s in stripes) == n
    assert stripes[0].file_offset == 2 * 8168 and stripes[0].first_row_number == 1
    for s in stripes:
        assert s.file_offset % 8168 == 0            # AlignReservation
        off = 0
        for col in range(3):
            for k in range(s.chunk_count):           # all exists buffers first ...
                nd = nodes[s.skipnode_base + col * s.chunk_count + k]
                assert nd.exists_offset == off and nd.exists_length == (nd.row_count + 7) // 8
                off += nd.exists_length
            for k in range(s.chunk_count):           # ... then all value buffers
                nd = nodes[s.skipnode_base + col * s.chunk_count + k]
                assert nd.value_offset == off
                off += nd.value_length
        assert off == s.data_length
    vals, nulls = t.decode_all()
    assert np.array_equal(vals[0], a) and np.array_equal(vals[2], c)
    assert np.array_equal(nulls[1].astype(bool), nulls_b)
    assert np.array_equal(vals[1][~nulls_b], b[~nulls_b])
    # NULLs occupy no bytes in the value stream
    nd = nodes[stripes[0].skipnode_base + 1 * stripes[0].chunk_count + 0]
    assert nd.decompressed_size == 4 * int((~nulls_b[:1000]).sum())


def test_codec_stream_formats(oracle):
    """hand-assembled streams of the two LZ77 formats the GPU decoder implements (the reference holds no
    golden compressed buffers, so these pin the FORMAT, not an encoder)"""
    # pglz: control byte 0b10 = [literal 'a'][tag len 9 off 1]; 8-byte ColumnarCompressHeader in front
    body = bytes([0b00000010, 0x61, (0 << 4) | (9 - 3), 0x01])
    total = len(body) + 8
    hdr = ((total << 2) | 2).to_bytes(4, "little") + (10).to_bytes(4, "little")
    assert oracle.codec_decompress(oracle.COMP_PGLZ, hdr + body, 10) == b"a" * 10
    # a tag with the extension byte: len 18 + 7 = 25 copies of "ab" pattern (off 2)
    body = bytes([0b00000100, 0x61, 0x62, (0 << 4) | 0x0f, 0x02, 7])
    total = len(body) + 8
    hdr = ((total << 2) | 2).to_bytes(4, "little") + (27).to_bytes(4, "little")
    assert oracle.codec_decompress(oracle.COMP_PGLZ, hdr + body, 27) == (b"ab" * 14)[:27]
    with pytest.raises(oracle.OracleError):     # VARSIZE must equal the buffer length
        oracle.codec_decompress(oracle.COMP_PGLZ, hdr + body + b"\0", 27)
    with pytest.raises(oracle.OracleError):     # check_complete: the stream must produce exactly rawsize bytes
        oracle.codec_decompress(oracle.COMP_PGLZ, hdr[:4] + (28).to_bytes(4, "little") + body, 28)
    if oracle.lib().orc_have_lz4():
        # lz4 block: [token 1 literal | match 10-4]['a'][offset 1] [token 5 literals]['aaaaa']
        blk = bytes([0x16, 0x61, 0x01, 0x00, 0x50]) + b"a" * 5
        assert oracle.codec_decompress(oracle.COMP_LZ4, blk, 16) == b"a" * 16
        # 300 literals need two length-extension bytes (15 + 255 + 30), then the stream ends
        lit = bytes(range(256)) + bytes(range(44))
        blk = bytes([0xF0, 255, 30]) + lit
        assert oracle.codec_decompress(oracle.COMP_LZ4, blk, 300) == lit
        with pytest.raises(oracle.OracleError):
            oracle.codec_decompress(oracle.COMP_LZ4, blk, 299)


@pytest.mark.parametrize("comp", ["lz4", "zstd", "pglz"])
def test_compressed_roundtrip(oracle, comp):
    if comp == "lz4" and not oracle.lib().orc_have_lz4():
        pytest.skip("liblz4 missing")
    if comp == "zstd" and not oracle.lib().orc_have_zstd():
        pytest.skip("libzstd missing")
    rng = np.random.default_rng(3)
    n = 30000
    a = rng.integers(0, 50, n)
    b = np.arange(n)
    t = oracle.Table([8, 8], compression={"lz4": oracle.COMP_LZ4, "zstd": oracle.COMP_ZSTD, "pglz": oracle.COMP_PGLZ}[comp])
    t.insert([a, b])
    nodes = t.nodes()
    assert any(nd.compression_type != 0 and nd.value_length < nd.decompressed_size for nd in nodes)
    vals, _ = t.decode_all()
    assert np.array_equal(vals[0], a) and np.array_equal(vals[1], b)
    r = t.scan([(0, "<", 10)], aggs=[oracle.sum_(1), oracle.count_star()])
    g = r.groups()[0]
    assert g[0]["sum"] == int(b[a < 10].sum()) and g[1]["count"] == int((a < 10).sum())


# --------------------------------------------------------------------------- aggregates
def test_contestant_aggregates(oracle, expected):
    rating = np.array(expected["contestant_rating"])
    country = expected["contestant_country"]
    codes = sorted(set(country))
    ccode = np.array([codes.index(c) for c in country])
    t = oracle.Table([4, 4])
    t.insert([rating[:5], ccode[:5]])      # contestants.1.csv then contestants.2.csv
    t.insert([rating[5:], ccode[5:]])
    r = t.scan(aggs=[oracle.count_star(), oracle.sum_(0)])
    g = r.groups()[0]
    assert g[0]["count"] == expected["contestant_count"]
    assert oracle.numeric_div_str(g[1]["sum"], 0, g[1]["count"], 0) == expected["contestant_avg"]
    r = t.scan([(0, ">", 2200)], group_cols=[1], aggs=[oracle.sum_(0)])
    got = [[codes[k], oracle.numeric_div_str(v[0]["sum"], 0, v[0]["count"], 0)] for k, v in sorted(r.groups().items())]
    assert got == expected["contestant_group_avg"]


def test_sum_is_null_when_every_input_is_null(oracle, expected):
    """expected/aggregate_support.out:121-128, 312-322: aggdata(id, key, val, valf) hash-distributed on id, NULLs in val;
    SELECT key, sum(val) ... GROUP BY key is NULL (not 0) for keys 5 and 6, whose val is NULL in every row -- through
    worker partials per shard and the coordinator's sum(sum) like the reference plans it"""
    rows = expected["aggdata_rows"]
    ids = np.array([r[0] for r in rows]); keys = np.array([r[1] for r in rows])
    val = np.array([0 if r[2] is None else r[2] for r in rows]); val_null = np.array([r[2] is None for r in rows], np.uint8)
    mins, maxs = oracle.synthetic_intervals(4)
    shard_of, _ = oracle.partition_rows(ids, None, 4, "h", mins, maxs)          # create_distributed_table('aggdata', 'id')
    aggs = [oracle.sum_(2), oracle.count(2), oracle.count_star()]
    total = oracle.Result(aggs)
    for s in range(4):
        sel = shard_of == s
        if not sel.any():
            continue
        t = oracle.Table([4, 4, 4])
        t.insert([ids[sel], keys[sel], val[sel]], [None, None, val_null[sel]])
        total.combine(t.scan(group_cols=[1], aggs=aggs))
    got = [[int(k), (None if g[0]["count"] == 0 else int(g[0]["sum"]))] for k, g in sorted(total.groups().items())]
    assert got == expected["aggdata_sum_val_by_key"]
    assert {int(k): g[1]["count"] for k, g in total.groups().items()} == {1: 1, 2: 3, 3: 1, 5: 0, 6: 0, 7: 1, 9: 1}      # count(val)
    assert sum(g[2]["count"] for g in total.groups().values()) == len(rows)                                           # count(*)


def _lineitem_shards(oracle, li):
    """lineitem is hash-distributed on l_orderkey with shard_count 2 (sql/multi_create_table.sql:30)"""
    mins, maxs = oracle.synthetic_intervals(2)
    idx, _ = oracle.partition_rows(li["l_orderkey"], None, 8, "h", mins, maxs)
    cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus",
            "l_shipdate", "l_suppkey"]
    shards = []
    for s in range(2):
        t = oracle.Table([8, 8, 8, 8, 1, 1, 4, 4], stripe_row_limit=2000, chunk_row_limit=1000)
        sel = idx == s
        t.insert([li[c][sel] for c in cols])
        shards.append(t)
    return shards


def test_tpch_q6(oracle, expected, lineitem):
    # sum(l_extendedprice * l_discount) where shipdate in [1994-01-01, 1995-01-01), discount
    # between 0.05 and 0.07, quantity < 24 -- decimals as scaled int64 (cents)
    quals = [(6, ">=", pgdate("1994-01-01")), (6, "<", pgdate("1995-01-01")),
             (2, ">=", 5), (2, "<=", 7), (0, "<", 2400)]
    aggs = [oracle.Agg(oracle.AGG_SUM, [(1, 0, 1), (2, 0, 1)])]
    total = oracle.Result(aggs)
    for t in _lineitem_shards(oracle, lineitem):
        total.combine(t.scan(quals, aggs=aggs))          # worker partial -> coordinator sum(sum)
    assert oracle.numeric_str(total.groups()[0][0]["sum"], 4) == expected["tpch_q6"]


def test_tpch_q1(oracle, expected, lineitem):
    quals = [(6, "<=", pgdate(Q1_CUTOFF))]
    aggs = [oracle.sum_(0), oracle.sum_(1),
            oracle.Agg(oracle.AGG_SUM, [(1, 0, 1), (2, 100, -1)]),
            oracle.Agg(oracle.AGG_SUM, [(1, 0, 1), (2, 100, -1), (3, 100, 1)]),
            oracle.sum_(2), oracle.count_star()]
    total = oracle.Result(aggs)
    for t in _lineitem_shards(oracle, lineitem):
        total.combine(t.scan(quals, group_cols=[4, 5], aggs=aggs))
    rows = []
    for key, v in sorted(total.groups().items(), key=lambda kv: (kv[0] & 0xffffffff, kv[0] >> 32)):
        n = v[5]["count"]
        rows.append([chr(key & 0xff), chr((key >> 32) & 0xff),
                     oracle.numeric_str(v[0]["sum"], 2), oracle.numeric_str(v[1]["sum"], 2),
                     oracle.numeric_str(v[2]["sum"], 4), oracle.numeric_str(v[3]["sum"], 6),
                     oracle.numeric_div_str(v[0]["sum"], 2, v[0]["count"], 0),
                     oracle.numeric_div_str(v[1]["sum"], 2, v[1]["count"], 0),
                     oracle.numeric_div_str(v[4]["sum"], 2, v[4]["count"], 0),
                     str(n)])
    assert rows == expected["tpch_q1"]


def test_sum_int4_and_float_aggregates(oracle, expected, lineitem):
    shards = _lineitem_shards(oracle, lineitem)
    aggs = [oracle.sum_(7)]
    total = oracle.Result(aggs)
    for t in shards:
        total.combine(t.scan(aggs=aggs))
    assert str(total.groups()[0][0]["sum"]) == expected["sum_l_suppkey"]
    # multi_agg_type_conversion: float(20) = float4, float(40) = float8
    f = np.array([float(r[0]) for r in expected["agg_type_data"]])
    d = np.array([float(r[1]) for r in expected["agg_type_data"]])
    t = oracle.Table([4, 8], atttype=[oracle.T_FLOAT, oracle.T_FLOAT])
    t.insert([f, d])
    for col, exp in ((0, expected["agg_type_float"]), (1, expected["agg_type_double"])):
        r = t.scan(aggs=[oracle.min_(col, True), oracle.max_(col, True), oracle.sum_(col, True), oracle.count(col)])
        g = r.groups()[0]
        assert g[0]["fmin"] == pytest.approx(float(exp[0]), rel=1e-12)
        assert g[1]["fmax"] == pytest.approx(float(exp[1]), rel=1e-12)
        assert g[2]["fsum"] == pytest.approx(float(exp[2]), rel=1e-12)
        assert g[3]["count"] == int(exp[3])
        assert g[2]["fsum"] / g[3]["count"] == pytest.approx(float(exp[4]), rel=1e-12)


def test_null_semantics(oracle):
    # count(x) skips NULL; sum over all-NULL input has count 0 (SQL NULL); NULL group key
    # forms its own group; a NULL comparison drops the row
    key = np.array([1, 1, 2, 2, 0, 0])
    keyn = np.array([0, 0, 0, 0, 1, 1])
    x = np.array([10, 0, 0, 0, 5, 7])
    xn = np.array([0, 1, 1, 1, 0, 0])
    t = oracle.Table([8, 8])
    t.insert([key, x], nulls=[keyn, xn])
    r = t.scan(group_cols=[0], aggs=[oracle.count_star(), oracle.count(1), oracle.sum_(1)])
    g = r.groups()
    assert g[1][0]["count"] == 2 and g[1][1]["count"] == 1 and g[1][2]["sum"] == 10
    assert g[2][0]["count"] == 2 and g[2][1]["count"] == 0 and g[2][2]["count"] == 0
    assert g[None][0]["count"] == 2 and g[None][2]["sum"] == 12
    r = t.scan([(1, "<", 100)], aggs=[oracle.count_star()])
    assert r.rows_passed == 3 and r.rows_removed_by_filter == 3


def test_int128_sum(oracle):
    big = np.full(1000, 2**62, dtype=np.int64)
    t = oracle.Table([8])
    t.insert([big])
    r = t.scan(aggs=[oracle.sum_(0)])
    assert r.groups()[0][0]["sum"] == 1000 * 2**62
    t = oracle.Table([8])
    t.insert([-big])
    assert t.scan(aggs=[oracle.sum_(0)]).groups()[0][0]["sum"] == -1000 * 2**62


def test_join_restatement_against_brute_force(oracle):
    rng = np.random.default_rng(2)
    bk, bx = rng.integers(0, 40, 300), rng.integers(-2**62, 2**62, 300)
    pk, py = rng.integers(0, 50, 400), rng.integers(-2**62, 2**62, 400)
    bn, pn = (rng.random(300) < 0.1), (rng.random(400) < 0.1)
    joined, total = oracle.join_count_sum(bk, bx, pk, py, bn, pn)
    cnt, tot = 0, 0
    for k, y, isnull in zip(pk, py, pn):
        if isnull:
            continue
        m = bx[(bk == k) & ~bn]
        cnt += len(m)
        tot += sum(int(x) + int(y) for x in m)
    assert joined == cnt and total == tot


def test_or_clauses_pushdown_goldens(oracle, expected):
    """expected/columnar_chunk_filtering.out, pushdown_test: OR clauses reach the chunk-group filter
    (predicate_refuted_by: an OR clause is refuted when every arm is) and ExecQual re-checks every row"""
    t = oracle.Table([4, 4], stripe_row_limit=2000, chunk_row_limit=1000)
    a = np.arange(1, 200001)
    t.insert([a, np.zeros_like(a)], nulls=[None, np.ones(a.shape[0], np.uint8)])
    trees = {
        "a = 204356 or a = 104356 or a = 76556": ("or", (0, "=", 204356), (0, "=", 104356), (0, "=", 76556)),
        "a = 194356 or a = 104356 or a = 76556": ("or", (0, "=", 194356), (0, "=", 104356), (0, "=", 76556)),
        "(a > 1000 and a < 10000) or (a > 20000 and a < 50000)":
            ("or", ("and", (0, ">", 1000), (0, "<", 10000)), ("and", (0, ">", 20000), (0, "<", 50000))),
    }
    assert len(expected["pushdown_or"]) == 3
    for g in expected["pushdown_or"]:
        r = t.scan(trees[g["where"]], [], [oracle.sum_(0)])
        assert r.chunk_groups_filtered == g["groups_removed"]
        assert r.rows_removed_by_filter == g["rows_removed"]
        assert r.groups()[0][0]["sum"] == g["sum"]
    # an OR over two columns is never refuted by one column's range; a NULL arm is not TRUE
    r = t.scan(("or", (0, "=", 5), (1, "=", 0)), [], [oracle.count_star()])
    assert r.chunk_groups_filtered == 0 and r.rows_passed == 1


def test_copy_files_match_the_golden_byte_counts(oracle, expected):
    """expected/partitioned_intermediate_results.out: bytes_written of the text files (21/14/5/9) and the binary
    files (93/57/39/75) of the squares table -- now from the oracle's actual file contents"""
    i = np.arange(1, 11)
    idx, _ = oracle.partition_rows(i, None, 4, "h", expected["squares_hash_mins"], expected["squares_hash_maxs"])
    files = oracle.copy_files(idx, [i, i * i], [4, 4], 4, binary=False)
    assert [[p, int((idx == p).sum()), len(files[p])] for p in range(4)] == expected["squares_hash_text"]
    assert files[int(idx[0])].startswith(b"1\t1\n") or b"1\t1\n" in files[int(idx[0])]
    idx, _ = oracle.partition_rows(i * i, None, 4, "r", [0, 21, 41, 61], [20, 40, 60, 100])
    files = oracle.copy_files(idx, [i, i * i], [4, 4], 4, binary=True)
    assert [[p, int((idx == p).sum()), len(files[p])] for p in range(4)] == expected["squares_range_binary"]
    assert files[0][:11] == b"PGCOPY\n\xff\r\n\0" and files[0][-2:] == b"\xff\xff"
