/*
 * gpu_partition_udf.c -- worker_partition_query_result with the reference's signature and return rows, routing and
 * serialising the rows on the GPU.
 *
 * The reference (file:line under /root/reference/src):
 *   backend/distributed/sql/udfs/worker_partition_query_result/latest.sql    the DDL: (result_prefix text, query text,
 *       partition_column_index int, partition_method citus.distribution_type, partition_min_values text[],
 *       partition_max_values text[], binary_copy bool, allow_null_partition_column bool, generate_empty_results bool)
 *       -> SETOF (partition_index int, rows_written bigint, bytes_written bigint)
 *   executor/partitioned_intermediate_results.c:115-298   argument decoding :122-149, must run in a transaction block,
 *       one PartitionedResultDestReceiver over the query; :493-553 per row: NULL key -> partition 0, FindShardInterval,
 *       forward to that partition's file receiver; :270-291 the return rows
 *   worker/worker_sql_task_protocol.c:91-251   TaskFileDestReceiver: the COPY text / binary files
 * Here the query's result columns (fixed-width integers: that is what a repartition join's map query produces for
 * int keys and payloads) are collected column-wise by the caller, uploaded once, and
 *   cg_partition_index          routes every row (hashint4/8 + interval search, or the raw value for range)
 *   cg_partition_copy_bytes     gives rows_written / bytes_written
 *   cg_partition_copy_serialize formats all P files on the device; they are written out with one write() each
 * When the consumer of the partitions is a GPU task too, nothing is serialised at all: cg_comm_repartition_exchange
 * moves the rows over NVLink (citus_gpu.h).
 */
#include "postgres.h"

#include "catalog/pg_type.h"
#include "fmgr.h"
#include "funcapi.h"
#include "utils/array.h"
#include "utils/builtins.h"
#include "utils/tuplestore.h"

#include "distributed/intermediate_results.h"

#include "citus_gpu.h"

PG_FUNCTION_INFO_V1(gpu_worker_partition_query_result);

/* device-side helpers the library deliberately does not expose as allocation API: the glue owns these buffers through
 * the CUDA runtime it links (cudaMalloc / cudaMemcpy); declared here to keep this file free of CUDA headers */
extern int cudaMalloc(void **ptr, size_t size);
extern int cudaFree(void *ptr);
extern int cudaMemcpy(void *dst, const void *src, size_t count, int kind);
#define GLUE_H2D 1
#define GLUE_D2H 2

/* the query's result, column-wise (filled by a DestReceiver over the map query) */
typedef struct GpuColumnBatch
{
	int ncols;
	int64 nrows;
	Oid type[8];
	int64 *values[8];
	uint8 *nulls[8];
} GpuColumnBatch;

extern GpuColumnBatch *RunQueryIntoColumnBatch(const char *query);          /* gpu_query_dest.c: DestReceiver collecting by-value columns */

static void
udf_check(int rc)
{
	if (rc != CG_OK)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("citus_gpu: %s", cg_last_error())));
}

static int32 *
text_array_to_int32(ArrayType *array, int *count)
{
	Datum *elems; bool *nulls; int n;
	deconstruct_array(array, TEXTOID, -1, false, 'i', &elems, &nulls, &n);
	int32 *out = palloc(sizeof(int32) * Max(n, 1));
	for (int i = 0; i < n; i++)
	{
		if (nulls[i])
			ereport(ERROR, (errmsg("unexpected NULL partition boundary")));
		char *s = text_to_cstring((text *) DatumGetPointer(elems[i]));
		out[i] = (int32) strtol(s, NULL, 10);
	}
	*count = n;
	return out;
}

Datum
gpu_worker_partition_query_result(PG_FUNCTION_ARGS)
{
	ReturnSetInfo *resultInfo = (ReturnSetInfo *) fcinfo->resultinfo;
	char *resultIdPrefix = text_to_cstring(PG_GETARG_TEXT_PP(0));
	char *queryString = text_to_cstring(PG_GETARG_TEXT_PP(1));
	int partitionColumnIndex = PG_GETARG_INT32(2);
	Oid partitionMethodOid = PG_GETARG_OID(3);
	int nmin = 0, nmax = 0;
	int32 *mins = text_array_to_int32(PG_GETARG_ARRAYTYPE_P(4), &nmin);
	int32 *maxs = text_array_to_int32(PG_GETARG_ARRAYTYPE_P(5), &nmax);
	bool binaryCopy = PG_GETARG_BOOL(6);
	bool allowNullPartitionColumnValues = PG_GETARG_BOOL(7);
	bool generateEmptyResults = PG_GETARG_BOOL(8);
	extern char DistributionMethodFromOid(Oid oid);         /* 'h' hash / 'r' range (metadata_utility.h) */
	char method = DistributionMethodFromOid(partitionMethodOid);

	if (method != 'h' && method != 'r')
		ereport(ERROR, (errmsg("only hash and range partitiong schemes are supported")));      /* sic: the reference's message */
	if (nmin != nmax)
		ereport(ERROR, (errmsg("min values and max values must have the same number of elements")));
	if (nmin == 0)
		ereport(ERROR, (errmsg("number of partitions cannot be 0")));
	const int P = nmin;

	GpuColumnBatch *batch = RunQueryIntoColumnBatch(queryString);
	if (partitionColumnIndex < 0 || partitionColumnIndex >= batch->ncols)
		ereport(ERROR, (errmsg("partition column index must be between 0 and %d", batch->ncols - 1)));
	const int64 n = batch->nrows;
	const uint8 *keyNulls = batch->nulls[partitionColumnIndex];
	if (!allowNullPartitionColumnValues && keyNulls)
		for (int64 i = 0; i < n; i++)
			if (keyNulls[i])
				ereport(ERROR, (errmsg("the partition column value cannot be NULL")));

	udf_check(cg_init(0));
	/* upload the columns once */
	int64 *d_cols[8]; uint8 *d_nulls[8]; int32 col_len[8];
	for (int c = 0; c < batch->ncols; c++)
	{
		col_len[c] = batch->type[c] == INT8OID ? 8 : batch->type[c] == INT4OID ? 4 : 2;
		d_nulls[c] = NULL;
		if (cudaMalloc((void **) &d_cols[c], sizeof(int64) * Max(n, 1)) != 0 ||
			cudaMemcpy(d_cols[c], batch->values[c], sizeof(int64) * n, GLUE_H2D) != 0)
			ereport(ERROR, (errcode(ERRCODE_OUT_OF_MEMORY), errmsg("citus_gpu: cannot stage %lld rows", (long long) n)));
		if (batch->nulls[c])
		{
			if (cudaMalloc((void **) &d_nulls[c], Max(n, 1)) != 0 || cudaMemcpy(d_nulls[c], batch->nulls[c], n, GLUE_H2D) != 0)
				ereport(ERROR, (errcode(ERRCODE_OUT_OF_MEMORY), errmsg("citus_gpu: cannot stage %lld rows", (long long) n)));
		}
	}
	int32 *d_index; int64 *d_counts;
	if (cudaMalloc((void **) &d_index, sizeof(int32) * Max(n, 1)) != 0 || cudaMalloc((void **) &d_counts, sizeof(int64) * (P + 1)) != 0)
		ereport(ERROR, (errcode(ERRCODE_OUT_OF_MEMORY), errmsg("citus_gpu: out of device memory")));
	const int keyLen = batch->type[partitionColumnIndex] == INT8OID ? 8 : 4;
	udf_check(cg_partition_index(d_cols[partitionColumnIndex], d_nulls[partitionColumnIndex], n, keyLen, method == 'h', mins, maxs, P, d_index,
								 d_counts));
	int64 *rowsWritten = palloc0(sizeof(int64) * P), *bytesWritten = palloc0(sizeof(int64) * P), *fileOffsets = palloc0(sizeof(int64) * (P + 1));
	udf_check(cg_partition_copy_bytes(d_index, n, P, (const int64_t *const *) d_cols, (const uint8_t *const *) d_nulls, col_len, batch->ncols,
									  binaryCopy, generateEmptyResults, rowsWritten, bytesWritten));
	int64 total = 0;
	for (int p = 0; p < P; p++) total += bytesWritten[p];
	uint8 *d_files = NULL, *files = palloc(Max(total, 1));
	if (cudaMalloc((void **) &d_files, Max(total, 1)) != 0)
		ereport(ERROR, (errcode(ERRCODE_OUT_OF_MEMORY), errmsg("citus_gpu: out of device memory")));
	udf_check(cg_partition_copy_serialize(d_index, n, P, (const int64_t *const *) d_cols, (const uint8_t *const *) d_nulls, col_len, batch->ncols,
										  binaryCopy, generateEmptyResults, d_files, total, fileOffsets));
	if (total > 0 && cudaMemcpy(files, d_files, total, GLUE_D2H) != 0)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("citus_gpu: cannot read the partition files back")));

	/* <prefix>_<i> under the per-transaction intermediate-results directory (partitioned_intermediate_results.c:165, :300-330) */
	CreateIntermediateResultsDirectory();
	for (int p = 0; p < P; p++)
	{
		if (rowsWritten[p] == 0 && !generateEmptyResults)
			continue;                                       /* lazy start-up: the file is never created (:234) */
		char resultId[256];
		snprintf(resultId, sizeof resultId, "%s_%d", resultIdPrefix, p);
		int fd = pg_glue_open_result_file(QueryResultFileName(resultId));
		pg_glue_write_result_file(fd, files + fileOffsets[p], (size_t) (fileOffsets[p + 1] - fileOffsets[p]));
		pg_glue_close_result_file(fd);
	}
	for (int c = 0; c < batch->ncols; c++) { cudaFree(d_cols[c]); if (d_nulls[c]) cudaFree(d_nulls[c]); }
	cudaFree(d_index); cudaFree(d_counts); cudaFree(d_files);

	/* the return rows (:270-291): (partition_index, rows_written, bytes_written) for every partition */
	InitMaterializedSRF(fcinfo, 0);
	for (int p = 0; p < P; p++)
	{
		Datum values[3] = {Int32GetDatum(p), Int64GetDatum(rowsWritten[p]), Int64GetDatum(bytesWritten[p])};
		bool nulls[3] = {false, false, false};
		tuplestore_putvalues(resultInfo->setResult, resultInfo->setDesc, values, nulls);
	}
	PG_RETURN_VOID();
}
