/*
 * gpu_combine_dest.c -- coordinator-side glue: a TupleDestination that folds the per-shard partial-aggregate rows
 * into a device-side group table as they arrive, so that the combine query's HashAggregate has nothing left to do.
 *
 * The reference (file:line under /root/reference/src):
 *   include/distributed/tuple_destination.h:45-62   TupleDestination { putTuple, tupleDescForQuery, stats };
 *       executor/adaptive_executor.c:3974-3976 -- a task's tupleDest overrides the default: the supported extension point
 *   executor/adaptive_executor.c:3964-4189 ReceiveResults builds one HeapTuple per result row and hands it to putTuple;
 *       executor/tuple_destination.c:97 TupleStoreTupleDestPutTuple stores it for the combine query
 *   planner/multi_logical_optimizer.c:1807-1885, 2231-2275  the combine: sum(sum), COALESCE(sum(count)::int8, 0),
 *       min(min), max(max), executed by PostgreSQL's HashAggregate over the tuplestore
 * Here putTuple appends (key, partial values) to a batch; a full batch goes to cg_partial_merge_values (the combine
 * kernel, cg_merge_kernel); GpuCombineFinish() reads the combined groups back and stores the FINAL rows into the
 * tuplestore the Citus Adaptive scan returns from (executor/citus_custom_scan.c:263-289), in the worker query's column
 * order, so the combine query above it sees one row per group.
 * GPU workers that keep their partials in HBM skip this path entirely: cg_comm_combine reduces them over NVLink.
 */
#include "postgres.h"

#include "access/htup_details.h"
#include "catalog/pg_type.h"
#include "fmgr.h"
#include "utils/builtins.h"
#include "utils/memutils.h"
#include "utils/numeric.h"
#include "utils/tuplestore.h"

#include "distributed/tuple_destination.h"

#include "citus_gpu.h"

#define GPU_COMBINE_BATCH 65536

/* how column i of the worker query's result row maps onto the aggregate list (filled by the planner-side caller) */
typedef enum GpuPartialColumnKind { GPU_PCOL_KEY, GPU_PCOL_COUNT, GPU_PCOL_SUM, GPU_PCOL_MINMAX, GPU_PCOL_FSUM } GpuPartialColumnKind;
typedef struct GpuPartialColumn { int32 kind; int32 index; Oid type; } GpuPartialColumn;

typedef struct GpuCombineTupleDest
{
	TupleDestination pub;            /* must be first: the executor only knows this part */
	TupleDesc tupleDesc;             /* of the worker query's rows */
	Tuplestorestate *tupleStore;     /* where the combined rows go at the end */
	CgScanDesc desc;
	CgPartial *partial;
	int ncols;
	GpuPartialColumn cols[CG_MAX_AGGS + CG_MAX_GROUP_COLS];
	/* batch */
	int64 nbatch;
	int64 *keys;
	uint8 *key_nulls;
	int64 *sum_hi;
	uint64 *sum_lo;
	int64 *count;
	int64 *minmax;
	double *fsum;
	Datum *values;
	bool *isnull;
	TupleDestinationStats stats;
} GpuCombineTupleDest;

static void
combine_check(int rc)
{
	if (rc != CG_OK)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("citus_gpu: %s", cg_last_error())));
}

/* numeric partial sum -> 128-bit integer (the partial sums of integer columns are integral) */
static void
numeric_to_i128(Datum numeric, int64 *hi, uint64 *lo)
{
	char *s = DatumGetCString(DirectFunctionCall1(numeric_out, numeric));
	bool neg = false;
	unsigned __int128 v = 0;
	const char *p = s;
	if (*p == '-') { neg = true; p++; }
	for (; *p >= '0' && *p <= '9'; p++)
		v = v * 10 + (unsigned) (*p - '0');
	if (*p != '\0')
		ereport(ERROR, (errcode(ERRCODE_FEATURE_NOT_SUPPORTED), errmsg("citus_gpu: partial sum \"%s\" is not an integer", s)));
	__int128 sv = neg ? -(__int128) v : (__int128) v;
	*hi = (int64) (sv >> 64);
	*lo = (uint64) sv;
	pfree(s);
}

static void
flush_batch(GpuCombineTupleDest *dest)
{
	if (dest->nbatch == 0)
		return;
	combine_check(cg_partial_merge_values(dest->partial, dest->nbatch, dest->keys, dest->key_nulls, dest->sum_hi, dest->sum_lo, dest->count,
										  dest->minmax, dest->fsum));
	dest->nbatch = 0;
}

static void
GpuCombinePutTuple(TupleDestination *self, Task *task, int placementIndex, int queryNumber, HeapTuple tuple, uint64 tupleLibpqSize)
{
	GpuCombineTupleDest *dest = (GpuCombineTupleDest *) self;
	(void) task; (void) placementIndex; (void) queryNumber;
	dest->stats.totalIntermediateResultSize += tupleLibpqSize;
	heap_deform_tuple(tuple, dest->tupleDesc, dest->values, dest->isnull);
	const int64 i = dest->nbatch;
	const int na = dest->desc.naggs;
	dest->keys[i] = 0;
	dest->key_nulls[i] = 0;
	for (int a = 0; a < na; a++)
	{
		dest->sum_hi[i * na + a] = 0; dest->sum_lo[i * na + a] = 0; dest->count[i * na + a] = 0; dest->minmax[i * na + a] = 0;
		dest->fsum[i * na + a] = 0;
	}
	for (int c = 0; c < dest->ncols; c++)
	{
		const GpuPartialColumn *col = &dest->cols[c];
		const int64 cell = i * na + col->index;
		const Datum v = dest->values[c];
		const bool isnull = dest->isnull[c];
		switch ((GpuPartialColumnKind) col->kind)
		{
			case GPU_PCOL_KEY:
			{
				int64 k = isnull ? 0 : col->type == INT8OID ? DatumGetInt64(v) : col->type == INT4OID ? (int64) DatumGetInt32(v) : (int64) DatumGetInt16(v);
				if (dest->desc.ngroup_cols == 2)        /* packed key: low 32 bits the first group column */
					dest->keys[i] |= col->index == 0 ? (int64) (uint32) k : (int64) ((uint64) (uint32) k << 32);
				else
					dest->keys[i] = k;
				if (isnull)
				{
					if (dest->desc.ngroup_cols == 2)
						ereport(ERROR, (errcode(ERRCODE_FEATURE_NOT_SUPPORTED), errmsg("citus_gpu: NULL in a two-column group key")));
					dest->key_nulls[i] = 1;
				}
				break;
			}
			case GPU_PCOL_COUNT:                         /* count(*) / count(x) partial: int8, never NULL */
				dest->count[cell] = isnull ? 0 : DatumGetInt64(v);
				break;
			case GPU_PCOL_SUM:                           /* NULL partial = no non-NULL input on that shard */
				if (isnull) break;
				if (col->type == NUMERICOID) numeric_to_i128(v, &dest->sum_hi[cell], &dest->sum_lo[cell]);
				else
				{
					int64 s = DatumGetInt64(v);
					dest->sum_hi[cell] = s < 0 ? -1 : 0; dest->sum_lo[cell] = (uint64) s;
				}
				dest->count[cell] = 1;                   /* "had input": the exact count rides in its own count column when the query has one */
				break;
			case GPU_PCOL_FSUM:
				if (isnull) break;
				dest->fsum[cell] = DatumGetFloat8(v);
				dest->count[cell] = 1;
				break;
			case GPU_PCOL_MINMAX:
				if (isnull) break;
				if (col->type == FLOAT8OID) { double d = DatumGetFloat8(v); memcpy(&dest->minmax[cell], &d, sizeof d); }
				else dest->minmax[cell] = col->type == INT8OID ? DatumGetInt64(v) : col->type == INT4OID ? (int64) DatumGetInt32(v) : (int64) DatumGetInt16(v);
				dest->count[cell] = 1;
				break;
		}
	}
	if (++dest->nbatch == GPU_COMBINE_BATCH)
		flush_batch(dest);
}

static TupleDesc
GpuCombineTupleDescForQuery(TupleDestination *self, int queryNumber)
{
	(void) queryNumber;
	return ((GpuCombineTupleDest *) self)->tupleDesc;
}

/*
 * desc: the worker query's aggregate list as a CgScanDesc (count(*) / count / sum / min / max over the partial
 * columns' source types); cols[i]: what column i of the worker query's result is.  key range unknown on the
 * coordinator: a hash table sized by the planner's group estimate.
 */
TupleDestination *
CreateGpuCombineTupleDest(Tuplestorestate *tupleStore, TupleDesc tupleDescriptor, const CgScanDesc *desc, const CgColumnDesc *sourceColumns,
						  int natts, const GpuPartialColumn *cols, int ncols)
{
	GpuCombineTupleDest *dest = palloc0(sizeof(GpuCombineTupleDest));
	dest->pub.putTuple = GpuCombinePutTuple;
	dest->pub.tupleDescForQuery = GpuCombineTupleDescForQuery;
	dest->pub.tupleDestinationStats = &dest->stats;
	dest->tupleDesc = tupleDescriptor;
	dest->tupleStore = tupleStore;
	dest->desc = *desc;
	dest->ncols = ncols;
	memcpy(dest->cols, cols, sizeof(GpuPartialColumn) * ncols);
	combine_check(cg_init(0));
	combine_check(cg_partial_create(desc, sourceColumns, natts, 0, -1, 0, &dest->partial));
	const Size cells = (Size) GPU_COMBINE_BATCH * Max(desc->naggs, 1);
	dest->keys = palloc(sizeof(int64) * GPU_COMBINE_BATCH); dest->key_nulls = palloc(GPU_COMBINE_BATCH);
	dest->sum_hi = palloc(sizeof(int64) * cells); dest->sum_lo = palloc(sizeof(uint64) * cells);
	dest->count = palloc(sizeof(int64) * cells); dest->minmax = palloc(sizeof(int64) * cells); dest->fsum = palloc(sizeof(double) * cells);
	dest->values = palloc(sizeof(Datum) * tupleDescriptor->natts); dest->isnull = palloc(sizeof(bool) * tupleDescriptor->natts);
	return &dest->pub;
}

/* after the last task has finished: the combined groups, one row each, into the tuplestore (same columns as the worker rows) */
void
GpuCombineFinish(TupleDestination *self)
{
	GpuCombineTupleDest *dest = (GpuCombineTupleDest *) self;
	flush_batch(dest);
	int64 n = 0;
	combine_check(cg_partial_ngroups(dest->partial, &n));
	const int na = dest->desc.naggs;
	const Size cells = (Size) Max(n, 1) * Max(na, 1);
	int64 *keys = palloc(sizeof(int64) * Max(n, 1)); uint8 *kn = palloc(Max(n, 1));
	int64 *hi = palloc(sizeof(int64) * cells); uint64 *lo = palloc(sizeof(uint64) * cells);
	int64 *cnt = palloc(sizeof(int64) * cells); int64 *mm = palloc(sizeof(int64) * cells); double *fs = palloc(sizeof(double) * cells);
	combine_check(cg_partial_fetch(dest->partial, Max(n, 1), keys, kn, hi, lo, cnt, mm, fs, &n));
	for (int64 g = 0; g < n; g++)
	{
		for (int c = 0; c < dest->ncols; c++)
		{
			const GpuPartialColumn *col = &dest->cols[c];
			const int64 cell = g * na + col->index;
			dest->isnull[c] = false;
			switch ((GpuPartialColumnKind) col->kind)
			{
				case GPU_PCOL_KEY:
				{
					int64 k = keys[g];
					if (dest->desc.ngroup_cols == 2) k = col->index == 0 ? (int64) (int32) (uint32) k : (int64) (int32) ((uint64) k >> 32);
					dest->isnull[c] = kn[g] != 0;
					dest->values[c] = col->type == INT8OID ? Int64GetDatum(k) : col->type == INT4OID ? Int32GetDatum((int32) k) : Int16GetDatum((int16) k);
					break;
				}
				case GPU_PCOL_COUNT:
					dest->values[c] = Int64GetDatum(cnt[cell]);
					break;
				case GPU_PCOL_SUM:
					if (cnt[cell] == 0) { dest->isnull[c] = true; break; }
					if (col->type == NUMERICOID)
					{
						char buf[64];
						combine_check(cg_numeric_out(hi[cell], lo[cell], 0, buf, sizeof buf));
						dest->values[c] = DirectFunctionCall3(numeric_in, CStringGetDatum(buf), ObjectIdGetDatum(InvalidOid), Int32GetDatum(-1));
					}
					else dest->values[c] = Int64GetDatum((int64) lo[cell]);
					break;
				case GPU_PCOL_FSUM:
					if (cnt[cell] == 0) { dest->isnull[c] = true; break; }
					dest->values[c] = Float8GetDatum(fs[cell]);
					break;
				case GPU_PCOL_MINMAX:
					if (cnt[cell] == 0) { dest->isnull[c] = true; break; }
					if (col->type == FLOAT8OID) { double d; memcpy(&d, &mm[cell], sizeof d); dest->values[c] = Float8GetDatum(d); }
					else dest->values[c] = col->type == INT8OID ? Int64GetDatum(mm[cell]) : col->type == INT4OID ? Int32GetDatum((int32) mm[cell]) : Int16GetDatum((int16) mm[cell]);
					break;
			}
		}
		tuplestore_putvalues(dest->tupleStore, dest->tupleDesc, dest->values, dest->isnull);
	}
	cg_partial_free(dest->partial);
	dest->partial = NULL;
}
