#!/usr/bin/env python
"""Builds the committed golden fixtures from the reference's own test inputs/outputs.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
Outputs (committed): tests/golden/lineitem.npz, tests/golden/expected.json
The GPU box has no /root/reference; tests only read the committed outputs.
"""
import datetime
import json
import os
import re

import numpy as np

REF = "/root/reference/src/test/regress"
HERE = os.path.dirname(os.path.abspath(__file__))


def cents(s: str) -> int:
    """decimal(15,2) text -> scaled int64"""
    neg = s.startswith("-")
    if neg:
        s = s[1:]
    whole, _, frac = s.partition(".")
    frac = (frac + "00")[:2]
    v = int(whole) * 100 + int(frac)
    return -v if neg else v


def pgdate(s: str) -> int:
    """PostgreSQL date Datum: days since 2000-01-01"""
    y, m, d = map(int, s.split("-"))
    return (datetime.date(y, m, d) - datetime.date(2000, 1, 1)).days


def lineitem():
    cols = {k: [] for k in ("l_orderkey", "l_partkey", "l_suppkey", "l_linenumber", "l_quantity",
                            "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus",
                            "l_shipdate", "l_commitdate", "l_receiptdate")}
    file_of_row = []
    for fi in (1, 2):
        with open(f"{REF}/data/lineitem.{fi}.data") as f:
            for line in f:
                p = line.rstrip("\n").split("|")
                cols["l_orderkey"].append(int(p[0]))
                cols["l_partkey"].append(int(p[1]))
                cols["l_suppkey"].append(int(p[2]))
                cols["l_linenumber"].append(int(p[3]))
                cols["l_quantity"].append(cents(p[4]))
                cols["l_extendedprice"].append(cents(p[5]))
                cols["l_discount"].append(cents(p[6]))
                cols["l_tax"].append(cents(p[7]))
                cols["l_returnflag"].append(ord(p[8]))
                cols["l_linestatus"].append(ord(p[9]))
                cols["l_shipdate"].append(pgdate(p[10]))
                cols["l_commitdate"].append(pgdate(p[11]))
                cols["l_receiptdate"].append(pgdate(p[12]))
                file_of_row.append(fi)
    out = {k: np.asarray(v, np.int64) for k, v in cols.items()}
    out["file_of_row"] = np.asarray(file_of_row, np.int8)
    np.savez_compressed(os.path.join(HERE, "lineitem.npz"), **out)
    return len(file_of_row)


def table_rows(path, after_regex, ncols_hint=None):
    """rows of the first psql result table that follows a line matching after_regex"""
    lines = open(path).read().split("\n")
    i = next(k for k, l in enumerate(lines) if re.search(after_regex, l))
    while not lines[i].startswith("----"):
        i += 1
    i += 1
    rows = []
    while not re.match(r"^\(\d+ rows?\)", lines[i]):
        rows.append([c.strip() for c in lines[i].split("|")])
        i += 1
    return rows


def main():
    n = lineitem()
    exp = {"lineitem_rows": n}

    # TPC-H Q1 / Q6 (expected/multi_tpch_query1.out, multi_tpch_query6.out)
    exp["tpch_q1"] = table_rows(f"{REF}/expected/multi_tpch_query1.out", r"^\s+l_linestatus;")
    exp["tpch_q6"] = table_rows(f"{REF}/expected/multi_tpch_query6.out", r"and l_quantity < 24;")[0][0]
    # sum(l_suppkey) (expected/multi_agg_type_conversion.out:5-9)
    exp["sum_l_suppkey"] = table_rows(f"{REF}/expected/multi_agg_type_conversion.out",
                                      r"^SELECT sum\(l_suppkey\) FROM lineitem;")[0][0]
    exp["agg_type_float"] = table_rows(f"{REF}/expected/multi_agg_type_conversion.out",
                                       r"sum\(float_value\), count\(float_value\), avg\(float_value\)")[0]
    exp["agg_type_double"] = table_rows(f"{REF}/expected/multi_agg_type_conversion.out",
                                        r"sum\(double_value\), count\(double_value\), avg\(double_value\)")[0]
    exp["agg_type_data"] = [l.split("\t")[:2] for l in open(f"{REF}/data/agg_type.data").read().split("\n") if l]

    # hash goldens
    pir = f"{REF}/expected/partitioned_intermediate_results.out"
    h4 = {}
    members = []
    for part in range(4):
        rows = table_rows(pir, rf"read_intermediate_result\('squares_hash_{part}', 'text'\)")
        members.append([int(r[1]) for r in rows])
        for r in rows:
            h4[int(r[1])] = int(r[0])
    exp["hashint4"] = {str(k): v for k, v in sorted(h4.items())}
    exp["squares_hash_members"] = members
    exp["squares_hash_mins"] = [-2147483648, -1073741824, 0, 1073741824]
    exp["squares_hash_maxs"] = [-1073741825, -1, 1073741823, 2147483647]
    exp["squares_hash_text"] = [[int(x) for x in r] for r in table_rows(pir, r"^SELECT \* FROM worker_partition_query_result\('squares_hash'")]
    exp["squares_range_binary"] = [[int(x) for x in r] for r in table_rows(pir, r"^SELECT \* FROM worker_partition_query_result\('squares_range'")]
    exp["doubles_hash_text"] = [[int(x) for x in r] for r in table_rows(pir, r"^SELECT \* FROM worker_partition_query_result\('doubles_hash'")]
    exp["doubles_range_binary"] = [[int(x) for x in r] for r in table_rows(pir, r"^SELECT \* FROM worker_partition_query_result\('doubles_range'")]

    # hashint8 (expected/distributed_planning.out:22-33): the inserted x values and the
    # sorted hash list
    exp["hashint8_inputs"] = [1, 3, 5, 2, 4, 6, 2608474032, 963809240]
    exp["hashint8_sorted"] = [int(r[0]) for r in table_rows(f"{REF}/expected/distributed_planning.out",
                                                            r"^SELECT hashint8\(x\) FROM test ORDER BY 1;")]
    exp["worker_hash_123"] = int(table_rows(f"{REF}/expected/multi_utilities.out", r"^SELECT worker_hash\(123\);")[0][0])
    exp["worker_hash_date_1997_08_08"] = int(table_rows(f"{REF}/expected/multi_utilities.out",
                                                        r"^SELECT worker_hash\('1997-08-08'::date\);")[0][0])
    exp["date_1997_08_08"] = pgdate("1997-08-08")

    # chunk filtering (expected/columnar_chunk_filtering.out)
    cf = f"{REF}/expected/columnar_chunk_filtering.out"
    text = open(cf).read()
    frc = re.findall(r"SELECT filtered_row_count\('SELECT count\(\*\) FROM test_chunk_filtering(.*?)'\);\n filtered_row_count\n-+\n\s+(\d+)", text)
    exp["chunk_filtering"] = [[w.strip(), int(v)] for w, v in frc]
    exp["simple_chunk_filtering"] = [
        # (rows 0..N inclusive, qual const, rows removed by filter, chunk groups removed)
        {"max": 234567, "gt": 123456, "rows_removed": 3457, "groups_removed": 12, "actual_rows": 111111},
        {"max": 200000, "gt": 180000, "rows_removed": 1, "groups_removed": 18, "actual_rows": 20000},
    ]
    assert "Rows Removed by Filter: 3457" in text and "Columnar Chunk Groups Removed by Filter: 12" in text
    assert "Columnar Chunk Groups Removed by Filter: 18" in text
    exp["multi_column_chunk_filtering"] = {"max": 234567, "gt": 50000, "rows_removed": 1,
                                           "groups_removed": 5, "actual_rows": 184567}
    assert "(actual rows=184567 loops=1)" in text and "Columnar Chunk Groups Removed by Filter: 5" in text

    # OR clauses pushed down to the chunk-group filter (pushdown_test: a = 1..200000, b NULL, stripe 2000, chunk
    # group 1000): WHERE text, "Rows Removed by Filter", "Columnar Chunk Groups Removed by Filter", sum(a)
    pd = []
    for where in ("a = 204356 or a = 104356 or a = 76556", "a = 194356 or a = 104356 or a = 76556",
                  "(a > 1000 and a < 10000) or (a > 20000 and a < 50000)"):
        m = re.search(r"SELECT sum\(a\) FROM pushdown_test (?:WHERE|where) " + re.escape(where) + r";\n.*?Rows Removed by Filter: (\d+)\n.*?"
                      r"Columnar Chunk Groups Removed by Filter: (\d+)\n.*?\n\s+sum\s*\n-+\n\s+(\d+)", text, re.S)
        pd.append({"where": where, "rows_removed": int(m.group(1)), "groups_removed": int(m.group(2)), "sum": int(m.group(3))})
    exp["pushdown_or"] = pd

    # columnar_query (expected/columnar_query.out:9-29) + data/contestants.{1,2}.csv
    ratings, countries = [], []
    for fi in (1, 2):
        for line in open(f"{REF}/data/contestants.{fi}.csv"):
            p = line.rstrip("\n").split(",")
            ratings.append(int(p[2]))
            countries.append(p[4].strip())
    exp["contestant_rating"] = ratings
    exp["contestant_country"] = countries
    cq = f"{REF}/expected/columnar_query.out"
    exp["contestant_count"] = int(table_rows(cq, r"^SELECT count\(\*\) FROM contestant;")[0][0])
    exp["contestant_avg"] = table_rows(cq, r"^SELECT avg\(rating\), stddev_samp\(rating\) FROM contestant;")[0][0]
    exp["contestant_group_avg"] = table_rows(cq, r"^SELECT country, avg\(rating\) FROM contestant WHERE rating > 2200")

    # aggregate_support (expected/aggregate_support.out:121-128, 312-322): aggdata(id, key, val, valf) with NULLs in val;
    # sum(val) GROUP BY key is NULL for the groups whose val is NULL in every row
    asup = f"{REF}/expected/aggregate_support.out"
    ins = re.search(r"insert into aggdata \(id, key, val, valf\) values (.*?);\n", open(asup).read()).group(1)
    rows = [[None if x.strip() == "NULL" else float(x) if "." in x else int(x) for x in t.split(",")]
            for t in re.findall(r"\(([^)]*)\)", ins)]
    exp["aggdata_rows"] = rows
    exp["aggdata_sum_val_by_key"] = [[int(r[0]), None if r[2] == "" else int(r[2])]
                                     for r in table_rows(asup, r"^SELECT key, internalsum\(val\), sum\(val\) from aggdata group by key order by key;")]

    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump(exp, f, indent=1, sort_keys=True)
    print("wrote", n, "lineitem rows and", len(exp), "golden entries")


if __name__ == "__main__":
    main()
