/*
 * gpu_comm_bootstrap.c -- how the backends of one query (one per GPU of the box) become the ranks of a
 * cg_comm communicator, over the connections the coordinator already holds.
 *
 * The reference has no such step: partial results travel back over libpq and are merged by the combine query
 * (executor/adaptive_executor.c:3964-4189 ReceiveResults; planner/multi_logical_optimizer.c:1807-1885).  When the worker
 * tasks of a query run on the GPUs of one box, the coordinator instead
 *   1. SELECT citus_gpu_comm_id()                       on the backend that runs the task placed on GPU 0
 *   2. SELECT citus_gpu_comm_init(id, rank, nranks)     on every backend (the task's GPU ordinal is its rank), in
 *                                                       parallel -- the call is a collective
 * with the same SendRemoteCommand / adaptive-executor machinery it uses for any other per-placement command
 * (executor/adaptive_executor.c:1901 StartDistributedExecution).  After that cg_comm_combine and
 * cg_comm_repartition_exchange (citus_gpu.h) move partial aggregates and repartitioned rows between the GPUs; at
 * transaction end every backend calls citus_gpu_comm_destroy() (also a collective).
 *
 * cg_comm_init maps the other backends' exchange buffers with CUDA IPC -- separate processes on one box, which is what
 * CUDA IPC is for; when the mapping is refused (containers without a shared IPC namespace) the library keeps the data
 * on NCCL and citus_gpu_comm_peer_window() says so.
 */
#include "postgres.h"

#include "fmgr.h"
#include "utils/builtins.h"

#include "citus_gpu.h"

PG_FUNCTION_INFO_V1(citus_gpu_comm_id);
PG_FUNCTION_INFO_V1(citus_gpu_comm_init);
PG_FUNCTION_INFO_V1(citus_gpu_comm_destroy);
PG_FUNCTION_INFO_V1(citus_gpu_comm_peer_window);

static void
comm_check(int rc)
{
	if (rc != CG_OK)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("citus_gpu: %s", cg_last_error())));
}

/* citus_gpu_comm_id() RETURNS bytea: the 128 bytes every rank passes to citus_gpu_comm_init */
Datum
citus_gpu_comm_id(PG_FUNCTION_ARGS)
{
	bytea *id = (bytea *) palloc(VARHDRSZ + CG_COMM_ID_BYTES);
	SET_VARSIZE(id, VARHDRSZ + CG_COMM_ID_BYTES);
	comm_check(cg_comm_unique_id((uint8_t *) VARDATA(id)));
	PG_RETURN_POINTER(id);
}

/* citus_gpu_comm_init(id bytea, rank int, nranks int) RETURNS void; the backend's GPU is its rank */
Datum
citus_gpu_comm_init(PG_FUNCTION_ARGS)
{
	bytea *id = (bytea *) PG_GETARG_POINTER(0);
	int32 rank = PG_GETARG_INT32(1), nranks = PG_GETARG_INT32(2);

	if (VARSIZE_ANY_EXHDR(id) != CG_COMM_ID_BYTES)
		ereport(ERROR, (errcode(ERRCODE_INVALID_PARAMETER_VALUE), errmsg("a communicator id is %d bytes", CG_COMM_ID_BYTES)));
	comm_check(cg_init(rank));                   /* lazily, in the backend: never in the postmaster (a forked child cannot inherit a CUDA context) */
	comm_check(cg_comm_init((const uint8_t *) VARDATA_ANY(id), rank, nranks));
	PG_RETURN_VOID();
}

Datum
citus_gpu_comm_destroy(PG_FUNCTION_ARGS)
{
	comm_check(cg_comm_destroy());
	PG_RETURN_VOID();
}

/* true: the library's kernels move the exchange data over NVLink through IPC-mapped buffers; false: NCCL does */
Datum
citus_gpu_comm_peer_window(PG_FUNCTION_ARGS)
{
	PG_RETURN_DATUM(BoolGetDatum(cg_comm_peer_window() != 0));
}
