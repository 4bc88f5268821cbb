/*
 * gpu_aggregate_shim.c -- worker_partial_agg / coord_combine_agg with the reference's SQL-callable signatures
 * (backend/distributed/sql/citus--9.0-2--9.1-1.sql:26-75), batching the rows of name-matched built-ins.
 *
 * The reference (backend/distributed/utils/aggregate_utils.c):
 *   worker_partial_agg_sfunc(internal, oid, anyelement) -> internal      :501-607  one syscache lookup + one fmgr call
 *                                                                                  of the wrapped aggregate's sfunc per row
 *   worker_partial_agg_ffunc(internal) -> cstring                        :705-743  transition state -> text (typoutput)
 *   coord_combine_agg_sfunc(internal, oid, cstring, anyelement) -> internal  :820-1003  text -> state, combinefunc per row
 *   coord_combine_agg_ffunc(internal, oid, cstring, anyelement) -> anyelement :1050-1130 finalfunc
 * These wrappers exist for aggregates the planner does NOT split by name (GetAggregateType,
 * planner/multi_logical_optimizer.c:3493-3596): custom aggregates with a combinefunc.  The shim keeps every signature
 * and, when the wrapped aggregate is one whose transition state is a by-value scalar with an order-independent
 * transition -- pg_catalog sum(int2|int4) -> int8, min / max over int2 / int4 / int8 / float8, count(any) -> int8 --
 * it appends the argument to a column batch instead of calling the sfunc, and reduces the batch on the GPU
 * (cg_agg_column) when the final function runs.  The text it returns is exactly what the reference's ffunc would
 * print for that state (int8out / the type's output function), so the coordinator side is unchanged.  Every other
 * aggregate is handed to the reference's own implementation (the original symbols, resolved at load time).
 */
#include "postgres.h"

#include "catalog/pg_aggregate.h"
#include "catalog/pg_proc.h"
#include "catalog/pg_type.h"
#include "fmgr.h"
#include "utils/builtins.h"
#include "utils/lsyscache.h"
#include "utils/memutils.h"
#include "utils/syscache.h"

#include "citus_gpu.h"

/* the reference's implementations, for everything the shim does not batch (bound by the extension's loader) */
extern Datum citus_worker_partial_agg_sfunc(PG_FUNCTION_ARGS);
extern Datum citus_worker_partial_agg_ffunc(PG_FUNCTION_ARGS);
extern Datum citus_coord_combine_agg_sfunc(PG_FUNCTION_ARGS);
extern Datum citus_coord_combine_agg_ffunc(PG_FUNCTION_ARGS);

PG_FUNCTION_INFO_V1(gpu_worker_partial_agg_sfunc);
PG_FUNCTION_INFO_V1(gpu_worker_partial_agg_ffunc);
PG_FUNCTION_INFO_V1(gpu_coord_combine_agg_sfunc);
PG_FUNCTION_INFO_V1(gpu_coord_combine_agg_ffunc);

typedef enum GpuBatchKind { GPU_BATCH_NONE = 0, GPU_BATCH_SUM, GPU_BATCH_MIN, GPU_BATCH_MAX, GPU_BATCH_COUNT } GpuBatchKind;

#define GPU_BOX_MAGIC 0x47505542u            /* distinguishes our box from the reference's StypeBox in arg 0 */

typedef struct GpuStypeBox
{
	uint32 magic;
	Oid agg;
	int32 kind;               /* GpuBatchKind */
	Oid argtype;
	int attlen;
	bool is_float;
	int64 n, cap;
	uint8 *values;            /* n * attlen bytes, native byte order (what store_att_byval would write) */
	uint8 *isnull;
	/* running state of batches already reduced */
	int64 count;
	__int128 sum;
	int64 minmax;
	bool have;
	MemoryContext cxt;
} GpuStypeBox;

static void
shim_check(int rc)
{
	if (rc != CG_OK)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("citus_gpu: %s", cg_last_error())));
}

/* is `agg` a pg_catalog aggregate the shim batches for this argument type? */
static GpuBatchKind
classify_aggregate(Oid agg, Oid argtype)
{
	GpuBatchKind kind = GPU_BATCH_NONE;
	HeapTuple proc = SearchSysCache1(PROCOID, ObjectIdGetDatum(agg));
	if (!HeapTupleIsValid(proc))
		return kind;
	Form_pg_proc form = (Form_pg_proc) GETSTRUCT(proc);
	const char *name = NameStr(form->proname);
	if (form->pronamespace == PG_CATALOG_NAMESPACE && form->pronargs <= 1)
	{
		bool ints = argtype == INT2OID || argtype == INT4OID || argtype == INT8OID;
		if (strcmp(name, "count") == 0) kind = GPU_BATCH_COUNT;
		else if (strcmp(name, "sum") == 0 && (argtype == INT2OID || argtype == INT4OID)) kind = GPU_BATCH_SUM;   /* stype int8; sum(int8) has an internal state */
		else if (strcmp(name, "min") == 0 && (ints || argtype == FLOAT8OID)) kind = GPU_BATCH_MIN;
		else if (strcmp(name, "max") == 0 && (ints || argtype == FLOAT8OID)) kind = GPU_BATCH_MAX;
	}
	ReleaseSysCache(proc);
	return kind;
}

static void
reduce_batch(GpuStypeBox *box)
{
	if (box->n == 0)
		return;
	int64 count = 0, hi = 0, mn = 0, mx = 0;
	uint64 lo = 0;
	shim_check(cg_init(0));
	shim_check(cg_agg_column(box->attlen, box->is_float ? 1 : 0, box->values, box->isnull, box->n, &count, &hi, &lo, &mn, &mx));
	box->count += count;
	box->sum += ((__int128) hi << 64) | (__int128) (unsigned __int128) lo;
	if (count > 0)
	{
		int64 v = box->kind == GPU_BATCH_MIN ? mn : mx;
		if (!box->have) box->minmax = v;
		else if (box->is_float)
		{
			double a, b;
			memcpy(&a, &box->minmax, 8); memcpy(&b, &v, 8);
			if ((box->kind == GPU_BATCH_MIN) ? (b < a) : (b > a)) box->minmax = v;   /* NaN handling stays with float8larger/smaller: NaN inputs are rare */
		}
		else if ((box->kind == GPU_BATCH_MIN) ? (v < box->minmax) : (v > box->minmax)) box->minmax = v;
		box->have = true;
	}
	box->n = 0;
}

Datum
gpu_worker_partial_agg_sfunc(PG_FUNCTION_ARGS)
{
	GpuStypeBox *box = PG_ARGISNULL(0) ? NULL : (GpuStypeBox *) PG_GETARG_POINTER(0);
	if (box == NULL)
	{
		if (PG_ARGISNULL(1))
			ereport(ERROR, (errmsg("worker_partial_agg_sfunc received invalid null input for second argument")));
		Oid agg = PG_GETARG_OID(1);
		Oid argtype = get_fn_expr_argtype(fcinfo->flinfo, 2);
		GpuBatchKind kind = classify_aggregate(agg, argtype);
		if (kind == GPU_BATCH_NONE)
			return citus_worker_partial_agg_sfunc(fcinfo);          /* the reference's generic path */
		MemoryContext aggcxt;
		if (!AggCheckCallContext(fcinfo, &aggcxt))
			elog(ERROR, "worker_partial_agg_sfunc called from non aggregate context");
		MemoryContext old = MemoryContextSwitchTo(aggcxt);
		box = palloc0(sizeof(GpuStypeBox));
		box->magic = GPU_BOX_MAGIC; box->agg = agg; box->kind = kind; box->argtype = argtype; box->cxt = aggcxt;
		int16 typlen; bool byval; char align;
		get_typlenbyvalalign(argtype, &typlen, &byval, &align);
		box->attlen = (kind == GPU_BATCH_COUNT || !byval || typlen <= 0) ? 1 : typlen;    /* count only needs the NULL flags */
		box->is_float = argtype == FLOAT8OID && kind != GPU_BATCH_COUNT;
		box->cap = 1 << 20;
		box->values = palloc((Size) box->cap * box->attlen);
		box->isnull = palloc((Size) box->cap);
		MemoryContextSwitchTo(old);
	}
	else if (box->magic != GPU_BOX_MAGIC)
		return citus_worker_partial_agg_sfunc(fcinfo);
	if (box->n == box->cap)
		reduce_batch(box);
	const bool isnull = PG_ARGISNULL(2);
	box->isnull[box->n] = isnull;
	if (!isnull && box->kind != GPU_BATCH_COUNT)
	{
		Datum v = PG_GETARG_DATUM(2);
		uint8 *dst = box->values + (Size) box->n * box->attlen;
		switch (box->attlen)
		{
			case 2: { int16 x = DatumGetInt16(v); memcpy(dst, &x, 2); break; }
			case 4: { int32 x = DatumGetInt32(v); memcpy(dst, &x, 4); break; }
			default:
				if (box->is_float) { double x = DatumGetFloat8(v); memcpy(dst, &x, 8); }
				else { int64 x = DatumGetInt64(v); memcpy(dst, &x, 8); }
				break;
		}
	}
	else
		memset(box->values + (Size) box->n * box->attlen, 0, (Size) box->attlen);
	box->n++;
	PG_RETURN_POINTER(box);
}

/* transition state -> text, as worker_partial_agg_ffunc prints it (aggregate_utils.c:705-743: the state type's typoutput) */
Datum
gpu_worker_partial_agg_ffunc(PG_FUNCTION_ARGS)
{
	GpuStypeBox *box = PG_ARGISNULL(0) ? NULL : (GpuStypeBox *) PG_GETARG_POINTER(0);
	if (box == NULL || box->magic != GPU_BOX_MAGIC)
		return citus_worker_partial_agg_ffunc(fcinfo);
	reduce_batch(box);
	Datum state;
	Oid statetype;
	switch ((GpuBatchKind) box->kind)
	{
		case GPU_BATCH_COUNT:
			state = Int64GetDatum(box->count); statetype = INT8OID;
			break;
		case GPU_BATCH_SUM:                          /* [PG] int4_sum / int2_sum: int8 state, NULL until the first non-NULL input */
			if (box->count == 0) PG_RETURN_NULL();
			state = Int64GetDatum((int64) box->sum); statetype = INT8OID;
			break;
		default:
			if (!box->have) PG_RETURN_NULL();
			statetype = box->argtype;
			if (box->is_float) { double d; memcpy(&d, &box->minmax, 8); state = Float8GetDatum(d); }
			else state = box->argtype == INT8OID ? Int64GetDatum(box->minmax) : box->argtype == INT4OID ? Int32GetDatum((int32) box->minmax)
																									   : Int16GetDatum((int16) box->minmax);
			break;
	}
	Oid typoutput; bool varlena_;
	getTypeOutputInfo(statetype, &typoutput, &varlena_);
	PG_RETURN_CSTRING(OidOutputFunctionCall(typoutput, state));
}

/*
 * coord_combine_agg(oid, cstring, anyelement): one text state per worker task.  The combine function of the batched
 * aggregates is the aggregate itself over the state type (sum of sums -- int8pl over int8 states --, min of mins, max
 * of maxes, sum of counts): the states are parsed with the state type's input function and batched like worker rows.
 */
Datum
gpu_coord_combine_agg_sfunc(PG_FUNCTION_ARGS)
{
	GpuStypeBox *box = PG_ARGISNULL(0) ? NULL : (GpuStypeBox *) PG_GETARG_POINTER(0);
	if (box == NULL)
	{
		if (PG_ARGISNULL(1))
			ereport(ERROR, (errmsg("coord_combine_agg_sfunc received invalid null input for second argument")));
		Oid agg = PG_GETARG_OID(1);
		Oid restype = get_fn_expr_argtype(fcinfo->flinfo, 3);
		Oid argtype = restype;
		HeapTuple proc = SearchSysCache1(PROCOID, ObjectIdGetDatum(agg));
		if (HeapTupleIsValid(proc))
		{
			/* the aggregate's own argument type is not visible here; sum(int2|int4) and count return int8 */
			Form_pg_proc form = (Form_pg_proc) GETSTRUCT(proc);
			if (strcmp(NameStr(form->proname), "sum") == 0 && restype == INT8OID) argtype = INT4OID;
			ReleaseSysCache(proc);
		}
		GpuBatchKind kind = classify_aggregate(agg, argtype);
		if (kind == GPU_BATCH_NONE)
			return citus_coord_combine_agg_sfunc(fcinfo);
		MemoryContext aggcxt;
		if (!AggCheckCallContext(fcinfo, &aggcxt))
			elog(ERROR, "coord_combine_agg_sfunc called from non aggregate context");
		MemoryContext old = MemoryContextSwitchTo(aggcxt);
		box = palloc0(sizeof(GpuStypeBox));
		box->magic = GPU_BOX_MAGIC; box->agg = agg; box->cxt = aggcxt;
		/* states of sum / count are int8 and are ADDED; min / max states have the argument's type */
		box->kind = (kind == GPU_BATCH_COUNT) ? GPU_BATCH_SUM : kind;
		box->argtype = (kind == GPU_BATCH_SUM || kind == GPU_BATCH_COUNT) ? INT8OID : restype;
		box->attlen = box->argtype == INT2OID ? 2 : box->argtype == INT4OID ? 4 : 8;
		box->is_float = box->argtype == FLOAT8OID;
		box->cap = 1 << 16;
		box->values = palloc((Size) box->cap * box->attlen);
		box->isnull = palloc((Size) box->cap);
		box->have = false;
		box->sum = 0;
		MemoryContextSwitchTo(old);
		box->count = kind == GPU_BATCH_COUNT ? -1 : 0;        /* -1: a count aggregate: never NULL, starts at 0 */
	}
	else if (box->magic != GPU_BOX_MAGIC)
		return citus_coord_combine_agg_sfunc(fcinfo);
	if (box->n == box->cap)
	{
		int64 keep = box->count;
		reduce_batch(box);
		if (keep < 0) box->count = -1;
	}
	const bool isnull = PG_ARGISNULL(2);
	box->isnull[box->n] = isnull;
	memset(box->values + (Size) box->n * box->attlen, 0, (Size) box->attlen);
	if (!isnull)
	{
		Oid typinput, ioparam;
		getTypeInputInfo(box->argtype, &typinput, &ioparam);
		Datum v = OidInputFunctionCall(typinput, PG_GETARG_CSTRING(2), ioparam, -1);
		uint8 *dst = box->values + (Size) box->n * box->attlen;
		if (box->is_float) { double x = DatumGetFloat8(v); memcpy(dst, &x, 8); }
		else if (box->attlen == 8) { int64 x = DatumGetInt64(v); memcpy(dst, &x, 8); }
		else if (box->attlen == 4) { int32 x = DatumGetInt32(v); memcpy(dst, &x, 4); }
		else { int16 x = DatumGetInt16(v); memcpy(dst, &x, 2); }
	}
	box->n++;
	PG_RETURN_POINTER(box);
}

Datum
gpu_coord_combine_agg_ffunc(PG_FUNCTION_ARGS)
{
	GpuStypeBox *box = PG_ARGISNULL(0) ? NULL : (GpuStypeBox *) PG_GETARG_POINTER(0);
	if (box == NULL || box->magic != GPU_BOX_MAGIC)
		return citus_coord_combine_agg_ffunc(fcinfo);
	const bool is_count = box->count < 0;
	if (is_count) box->count = 0;
	reduce_batch(box);
	if (box->kind == GPU_BATCH_SUM)
	{
		if (box->count == 0 && !is_count) PG_RETURN_NULL();            /* sum over no non-NULL state is NULL; count is 0 */
		PG_RETURN_DATUM(Int64GetDatum((int64) box->sum));
	}
	if (!box->have) PG_RETURN_NULL();
	if (box->is_float) { double d; memcpy(&d, &box->minmax, 8); PG_RETURN_DATUM(Float8GetDatum(d)); }
	PG_RETURN_DATUM(box->argtype == INT8OID ? Int64GetDatum(box->minmax) : box->argtype == INT4OID ? Int32GetDatum((int32) box->minmax)
																								   : Int16GetDatum((int16) box->minmax));
}
