/*
 * tests/c/driver.c -- a plain C caller of libcitus_gpu.so, the way the PostgreSQL-side glue (pg_glue/) calls it:
 * dlopen, the C-ABI of include/citus_gpu.h only, no Python, no torch.  Runs BASELINE configs[0] (C1) end to end:
 *
 *     4 shards x 2.5 M rows x 4 int8 columns,  SELECT sum(a), count(*) FROM t WHERE b < 250000
 *
 * shard by shard through cg_scan_relation (host page images -> pinned staging -> fused kernel), combines the four
 * per-shard partials the way the coordinator does (cg_partial_fetch -> cg_partial_merge_values), prints the result as
 * PostgreSQL would (sum(int8) is numeric: cg_numeric_out) and checks it against the oracle's row-at-a-time scan of the
 * same images (liboracle.so is test infrastructure and is only loaded here, by the test).
 *
 *     cc -O2 -I include tests/c/driver.c -ldl -o driver && ./driver citus_b200/lib/libcitus_gpu.so oracle/liboracle.so
 * exit code 0 = bit-exact.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "citus_gpu.h"

#define SYM(handle, name) __typeof__(&name) p_##name = (__typeof__(&name)) dlsym(handle, #name); \
	if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
#define CHECK(call) do { int rc__ = (call); if (rc__ != CG_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, p_cg_last_error()); return 1; } } while (0)

/* the oracle's view of a relation image (oracle/oracle.c orc_table_attach_view / orc_scan_aggregate) */
typedef struct OrcQual { int32_t col; int32_t op; int64_t konst; } OrcQual;
typedef struct OrcAggSpec { int32_t kind; int32_t nfactors; int32_t col[3]; int32_t is_float; int64_t a[3]; int64_t b[3]; } OrcAggSpec;
void *orc_table_attach_view(const uint8_t *, uint64_t, const void *, int, const void *, int, int, const int *, const int *, uint32_t);
void *orc_result_create(int);
int orc_scan_aggregate(const void *, const OrcQual *, int, int, const int *, int, const OrcAggSpec *, int, void *);
void orc_result_agg(const void *, int64_t, int, int64_t *, uint64_t *, int64_t *, int64_t *, int64_t *, double *, double *, double *);
int64_t orc_result_ngroups(const void *);

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: driver libcitus_gpu.so liboracle.so [rows]\n"); return 2; }
	void *lib = dlopen(argv[1], RTLD_NOW);
	if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
	void *orc = dlopen(argv[2], RTLD_NOW);
	if (!orc) { fprintf(stderr, "%s\n", dlerror()); return 2; }
	const int64_t total_rows = argc > 3 ? atoll(argv[3]) : 10000000;
	SYM(lib, cg_last_error); SYM(lib, cg_init); SYM(lib, cg_gen_relation); SYM(lib, cg_gen_relation_view); SYM(lib, cg_gen_relation_free);
	SYM(lib, cg_relation_bounds); SYM(lib, cg_partial_create); SYM(lib, cg_partial_free); SYM(lib, cg_scan_relation); SYM(lib, cg_partial_fetch);
	SYM(lib, cg_partial_merge_values); SYM(lib, cg_numeric_out); SYM(lib, cg_kernel_launches); SYM(lib, cg_shutdown);
	SYM(orc, orc_table_attach_view); SYM(orc, orc_result_create); SYM(orc, orc_scan_aggregate); SYM(orc, orc_result_agg); SYM(orc, orc_result_ngroups);

	CHECK(p_cg_init(0));
	const int nshards = 4;
	const int64_t per = total_rows / nshards;
	const CgGenColumn cols[4] = {{8, CG_GEN_UNIFORM, -(1ll << 31), 1ll << 31, 0, 0}, {8, CG_GEN_UNIFORM, 0, 1000000, 0, 0},
								 {8, CG_GEN_UNIFORM, 0, 1ll << 40, 0, 0}, {8, CG_GEN_SEQUENCE, 0, 0, 0, 0}};
	CgScanDesc desc;
	memset(&desc, 0, sizeof desc);
	desc.nquals = 1; desc.quals[0].column = 1; desc.quals[0].op = CG_OP_LT; desc.quals[0].konst = 250000;
	desc.enable_qual_pushdown = 1;
	desc.naggs = 2;
	desc.aggs[0].kind = CG_AGG_SUM; desc.aggs[0].nfactors = 1; desc.aggs[0].column[0] = 0; desc.aggs[0].b[0] = 1;
	desc.aggs[1].kind = CG_AGG_COUNT_STAR;
	const CgColumnDesc coldesc[4] = {{8, CG_TYPE_INT}, {8, CG_TYPE_INT}, {8, CG_TYPE_INT}, {8, CG_TYPE_INT}};

	/* the coordinator's table: a plain aggregate, filled from the per-shard partial rows */
	CgPartial *combined = NULL;
	CHECK(p_cg_partial_create(&desc, coldesc, 4, 0, -1, total_rows, &combined));
	__int128 want_sum = 0;
	int64_t want_count = 0;
	for (int s = 0; s < nshards; s++)
	{
		CgGenRelation *gen = NULL;
		CgRelation rel;
		CHECK(p_cg_gen_relation(cols, 4, (uint64_t) per, (uint64_t) (s * per), 20260921, 150000, 10000, 4, &gen));
		CHECK(p_cg_gen_relation_view(gen, &rel));
		/* one worker task: bounds from the skip lists, the fused scan, its partial row */
		CgScanDesc task = desc;
		int64_t kmin, kmax, bounds[CG_MAX_AGGS], rows;
		CHECK(p_cg_relation_bounds(&rel, &task, &kmin, &kmax, bounds, &rows));
		task.aggs[0].term_abs_bound = bounds[0];
		CgPartial *partial = NULL;
		CHECK(p_cg_partial_create(&task, coldesc, 4, kmin, kmax, rows, &partial));
		CgScanStats st;
		CHECK(p_cg_scan_relation(&rel, &task, partial, &st));
		int64_t key, hi[2], cnt[2], mm[2], n;
		uint64_t lo[2];
		uint8_t kn;
		double fs[2];
		CHECK(p_cg_partial_fetch(partial, 1, &key, &kn, hi, lo, cnt, mm, fs, &n));
		printf("shard %d: %lld rows scanned, %lld removed by filter, kernel %.3f ms, partial count %lld\n", s, (long long) st.rows_scanned,
			   (long long) st.rows_removed_by_filter, st.kernel_ms, (long long) cnt[1]);
		/* ... which the coordinator folds into its table (the combine kernel) */
		CHECK(p_cg_partial_merge_values(combined, 1, &key, &kn, hi, lo, cnt, mm, fs));
		p_cg_partial_free(partial);
		/* the oracle on the same image */
		const int attlen[4] = {8, 8, 8, 8}, atttype[4] = {0, 0, 0, 0};
		void *t = p_orc_table_attach_view(rel.pages, rel.nblocks, rel.stripes, rel.nstripes, rel.nodes, rel.nnodes, 4, attlen, atttype, 10000);
		void *res = p_orc_result_create(2);
		const OrcQual q = {1, 0 /* < */, 250000};
		const OrcAggSpec aggs[2] = {{2 /* sum */, 1, {0, 0, 0}, 0, {0, 0, 0}, {1, 0, 0}}, {0 /* count(*) */, 0, {0, 0, 0}, 0, {0, 0, 0}, {0, 0, 0}}};
		if (p_orc_scan_aggregate(t, &q, 1, 1, NULL, 0, aggs, 2, res) != 0) { fprintf(stderr, "oracle scan failed\n"); return 1; }
		if (p_orc_result_ngroups(res) > 0)
		{
			int64_t ohi, oc, omin, omax; uint64_t olo; double of, ofmin, ofmax;
			p_orc_result_agg(res, 0, 0, &ohi, &olo, &oc, &omin, &omax, &of, &ofmin, &ofmax);
			want_sum += ((__int128) ohi << 64) | (__int128) (unsigned __int128) olo;
			p_orc_result_agg(res, 0, 1, &ohi, &olo, &oc, &omin, &omax, &of, &ofmin, &ofmax);
			want_count += oc;
		}
		p_cg_gen_relation_free(gen);
	}
	int64_t key, hi[2], cnt[2], mm[2], n;
	uint64_t lo[2];
	uint8_t kn;
	double fs[2];
	CHECK(p_cg_partial_fetch(combined, 1, &key, &kn, hi, lo, cnt, mm, fs, &n));
	char text[64], want_text[64];
	CHECK(p_cg_numeric_out(hi[0], lo[0], 0, text, sizeof text));
	CHECK(p_cg_numeric_out((int64_t) (want_sum >> 64), (uint64_t) want_sum, 0, want_text, sizeof want_text));
	printf("sum(a) = %s, count(*) = %lld   (oracle: %s, %lld)   %llu kernel launches\n", text, (long long) cnt[1], want_text, (long long) want_count,
		   (unsigned long long) p_cg_kernel_launches());
	int ok = strcmp(text, want_text) == 0 && cnt[1] == want_count && cnt[1] > 0;
	p_cg_partial_free(combined);
	p_cg_shutdown();
	printf(ok ? "C1 through the C-ABI: bit-exact\n" : "MISMATCH\n");
	return ok ? 0 : 1;
}
