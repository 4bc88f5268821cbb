/*
 * gpu_columnar_agg.c -- PostgreSQL-side glue: a "GpuColumnarAgg" CustomScan that replaces the worker task's
 *     Agg <- Custom Scan (ColumnarScan)
 * subtree with one call into libcitus_gpu.so (include/citus_gpu.h).
 *
 * What it mirrors in the reference (file:line under /root/reference/src):
 *   registration     backend/columnar/columnar_customscan.c:155-175 (CustomScanMethods / CustomExecMethods),
 *                    :199-200, :265 (hook chaining, RegisterCustomScanMethods)
 *   path injection   the reference injects at set_rel_pathlist_hook (:269-351) because it replaces a SCAN; an
 *                    aggregate replacement has to be offered one level up: create_upper_paths_hook, UPPERREL_GROUP_AGG
 *   stripe listing   backend/columnar/columnar_reader.c:687-734 AdvanceStripeRead: stripes visible to the snapshot,
 *                    unflushed ones skipped (StripeWriteState), skip lists by ReadStripeSkipList
 *                    (backend/columnar/columnar_metadata.c:717-832)
 *   page access      backend/columnar/columnar_storage.c:463-492 ColumnarStorageRead / :669-689 ReadFromBlock: here the
 *                    8 KB pages are copied out of shared_buffers once per scan into one image the library DMAs from
 *   execution        ColumnarScan_ExecCustomScan :1907-1913 returns one slot per call; so does this node: the first call
 *                    runs the fused kernel, later calls stream the partial-aggregate rows
 *   errors           the library returns codes, never longjmps; every non-zero code becomes ereport(ERROR) here
 *
 * A query shape the library does not cover (CG_EUNSUPPORTED at plan time) simply gets no GpuColumnarAgg path: the
 * reference's own Agg <- ColumnarScan plan runs.  Built against the real headers with `make -C pg_glue`; type-checked in
 * this repository against pg_glue/stub/ (tests/test_pg_glue.py).
 */
#include "postgres.h"

#include "access/table.h"
#include "catalog/pg_type.h"
#include "commands/explain.h"
#include "executor/tuptable.h"
#include "fmgr.h"
#include "nodes/extensible.h"
#include "nodes/makefuncs.h"
#include "optimizer/pathnode.h"
#include "optimizer/planner.h"
#include "storage/bufmgr.h"
#include "utils/builtins.h"
#include "utils/lsyscache.h"
#include "utils/memutils.h"
#include "utils/numeric.h"
#include "utils/rel.h"
#include "utils/snapmgr.h"

#include "columnar/columnar.h"
#include "columnar/columnar_metadata.h"

#include "citus_gpu.h"

PG_MODULE_MAGIC;

void _PG_init(void);

/* what one output column of the node is made of */
typedef enum GpuOutKind { GPU_OUT_KEY, GPU_OUT_COUNT, GPU_OUT_SUM_NUMERIC, GPU_OUT_SUM_INT8, GPU_OUT_SUM_FLOAT8, GPU_OUT_MINMAX } GpuOutKind;

typedef struct GpuOutColumn
{
	int32 kind;          /* GpuOutKind */
	int32 index;         /* group column number / aggregate number in the CgScanDesc */
	Oid type;            /* result type */
	int32 scale;         /* numeric display scale of a sum over scaled integers (0 for plain int8) */
} GpuOutColumn;

/* the plan-time product: travels in CustomScan.custom_private as one bytea Const (so that copyObject works) */
typedef struct GpuAggPlan
{
	Oid relid;
	CgScanDesc desc;
	int32 nout;
	GpuOutColumn out[CG_MAX_AGGS + CG_MAX_GROUP_COLS];
} GpuAggPlan;

typedef struct GpuColumnarAggState
{
	CustomScanState css;         /* must be first (columnar_customscan.c:65-71 does the same) */
	GpuAggPlan plan;
	CgPartial *partial;
	CgScanStats stats;
	/* result rows fetched from the device */
	int64 ngroups, next;
	int64 *keys;
	uint8 *key_nulls;
	int64 *sum_hi;
	uint64 *sum_lo;
	int64 *count;
	int64 *minmax;
	double *fsum;
	bool done_scan;
} GpuColumnarAggState;

static create_upper_paths_hook_type prev_create_upper_paths_hook = NULL;
static bool gpu_device_ready = false;

/* ------------------------------------------------------------------------------------ *
 *  library errors -> ereport
 * ------------------------------------------------------------------------------------ */
static void
gpu_check(int rc)
{
	if (rc == CG_OK)
		return;
	int sqlstate = rc == CG_EUNSUPPORTED ? ERRCODE_FEATURE_NOT_SUPPORTED :
				   rc == CG_ECORRUPT ? ERRCODE_DATA_CORRUPTED :
				   rc == CG_ENOMEM ? ERRCODE_OUT_OF_MEMORY : ERRCODE_INTERNAL_ERROR;
	ereport(ERROR, (errcode(sqlstate), errmsg("citus_gpu: %s", cg_last_error())));
}

/* CUDA is initialised lazily in the backend that needs it, never in the postmaster (a forked CUDA context is unusable) */
static void
gpu_ensure_device(void)
{
	if (gpu_device_ready)
		return;
	gpu_check(cg_init(0));
	gpu_device_ready = true;
}

/* ------------------------------------------------------------------------------------ *
 *  plan time: does the worker query fit the library's plan language?
 * ------------------------------------------------------------------------------------ */
static bool
type_is_supported(Oid type, int *attlen, int *type_class)
{
	switch (type)
	{
		case INT8OID: case TIMESTAMPOID: case TIMESTAMPTZOID: *attlen = 8; *type_class = CG_TYPE_INT; return true;
		case INT4OID: case DATEOID: *attlen = 4; *type_class = CG_TYPE_INT; return true;
		case INT2OID: *attlen = 2; *type_class = CG_TYPE_INT; return true;
		case FLOAT8OID: *attlen = 8; *type_class = CG_TYPE_FLOAT; return true;
		case FLOAT4OID: *attlen = 4; *type_class = CG_TYPE_FLOAT; return true;
		default: return false;
	}
}

static Var *
strip_to_var(Node *node)
{
	while (node && IsA(node, RelabelType))
		node = (Node *) ((RelabelType *) node)->arg;
	return (node && IsA(node, Var)) ? (Var *) node : NULL;
}

/* btree strategy of an operator on the column's default operator class, 0 if it has none ("<>" is its negator) */
static int
btree_strategy(Oid opno, Oid type)
{
	Oid opclass = GetDefaultOpClass(type, BTREE_AM_OID);
	if (opclass == InvalidOid)
		return 0;
	return get_op_opfamily_strategy(opno, get_opclass_family(opclass));
}

/* one WHERE atom "Var op Const" (or commuted) -> CgQual; false when it is anything else */
static bool
translate_atom(OpExpr *op, TupleDesc tupdesc, CgQual *qual)
{
	if (list_length(op->args) != 2)
		return false;
	Node *l = (Node *) linitial(op->args), *r = (Node *) lsecond(op->args);
	Var *var = strip_to_var(l);
	Const *konst = (r && IsA(r, Const)) ? (Const *) r : NULL;
	Oid opno = op->opno;
	if (var == NULL || konst == NULL)
	{
		var = strip_to_var(r);
		konst = (l && IsA(l, Const)) ? (Const *) l : NULL;
		opno = get_commutator(op->opno);
		if (var == NULL || konst == NULL || opno == InvalidOid)
			return false;
	}
	if (konst->constisnull || var->varattno <= 0 || var->varattno > tupdesc->natts)
		return false;
	int attlen, type_class;
	if (!type_is_supported(var->vartype, &attlen, &type_class) || konst->consttype != var->vartype)
		return false;
	int strategy = btree_strategy(opno, var->vartype);
	int cgop;
	switch (strategy)
	{
		case BTLessStrategyNumber: cgop = CG_OP_LT; break;
		case BTLessEqualStrategyNumber: cgop = CG_OP_LE; break;
		case BTEqualStrategyNumber: cgop = CG_OP_EQ; break;
		case BTGreaterEqualStrategyNumber: cgop = CG_OP_GE; break;
		case BTGreaterStrategyNumber: cgop = CG_OP_GT; break;
		default:
			if (strcmp(get_opname(opno), "<>") != 0)
				return false;
			cgop = CG_OP_NE;
			break;
	}
	qual->column = var->varattno - 1;
	qual->op = cgop;
	if (type_class == CG_TYPE_FLOAT)
	{
		double d = var->vartype == FLOAT4OID ? (double) DatumGetFloat4(konst->constvalue) : DatumGetFloat8(konst->constvalue);
		memcpy(&qual->konst, &d, sizeof d);
	}
	else
		qual->konst = attlen == 8 ? DatumGetInt64(konst->constvalue) :
					  attlen == 4 ? (int64) DatumGetInt32(konst->constvalue) : (int64) DatumGetInt16(konst->constvalue);
	return true;
}

/* WHERE tree -> atoms + postfix tokens (CgScanDesc.qual_expr); false when a node is not AND / OR / atom or the tree is too big */
static bool
translate_where(Node *node, TupleDesc tupdesc, CgScanDesc *desc)
{
	if (IsA(node, BoolExpr))
	{
		BoolExpr *b = (BoolExpr *) node;
		if (b->boolop == NOT_EXPR)
			return false;
		ListCell *lc;
		int n = 0;
		foreach(lc, b->args)
		{
			if (!translate_where((Node *) lfirst(lc), tupdesc, desc))
				return false;
			if (n++ > 0)
			{
				if (desc->nqual_expr >= CG_MAX_QEXPR)
					return false;
				desc->qual_expr[desc->nqual_expr++] = b->boolop == AND_EXPR ? CG_QX_AND : CG_QX_OR;
			}
		}
		return n > 0;
	}
	if (!IsA(node, OpExpr) || desc->nquals >= CG_MAX_QUALS || desc->nqual_expr >= CG_MAX_QEXPR)
		return false;
	if (!translate_atom((OpExpr *) node, tupdesc, &desc->quals[desc->nquals]))
		return false;
	desc->qual_expr[desc->nqual_expr++] = (int8_t) desc->nquals++;
	return true;
}

/* the worker half of the aggregate split names plain built-ins (planner/multi_logical_optimizer.c:3160-3484) */
static bool
translate_aggref(Aggref *agg, TupleDesc tupdesc, CgScanDesc *desc, GpuOutColumn *out)
{
	if (agg->aggdistinct != NIL || agg->aggorder != NIL || agg->aggfilter != NULL || agg->aggkind != 'n' || desc->naggs >= CG_MAX_AGGS)
		return false;
	const char *name = get_func_name(agg->aggfnoid);
	CgAggSpec *spec = &desc->aggs[desc->naggs];
	memset(spec, 0, sizeof *spec);
	out->index = desc->naggs;
	out->scale = 0;
	if (agg->aggstar && strcmp(name, "count") == 0)
	{
		spec->kind = CG_AGG_COUNT_STAR;
		out->kind = GPU_OUT_COUNT; out->type = INT8OID;
		desc->naggs++;
		return true;
	}
	if (list_length(agg->args) != 1)
		return false;
	Var *var = strip_to_var((Node *) ((TargetEntry *) linitial(agg->args))->expr);
	int attlen, type_class;
	if (var == NULL || var->varattno <= 0 || var->varattno > tupdesc->natts || !type_is_supported(var->vartype, &attlen, &type_class))
		return false;
	spec->nfactors = 1;
	spec->column[0] = var->varattno - 1;
	spec->is_float = type_class == CG_TYPE_FLOAT;
	spec->a[0] = 0;
	if (spec->is_float) { double one = 1.0; memcpy(&spec->b[0], &one, sizeof one); }
	else spec->b[0] = 1;
	if (strcmp(name, "count") == 0)
	{
		spec->kind = CG_AGG_COUNT; spec->is_float = 0; spec->b[0] = 1;
		out->kind = GPU_OUT_COUNT; out->type = INT8OID;
	}
	else if (strcmp(name, "sum") == 0)
	{
		spec->kind = CG_AGG_SUM;
		if (spec->is_float) { out->kind = GPU_OUT_SUM_FLOAT8; out->type = var->vartype == FLOAT4OID ? FLOAT4OID : FLOAT8OID; }
		else if (var->vartype == INT8OID) { out->kind = GPU_OUT_SUM_NUMERIC; out->type = NUMERICOID; }   /* [PG] sum(int8) is numeric */
		else if (var->vartype == INT4OID || var->vartype == INT2OID) { out->kind = GPU_OUT_SUM_INT8; out->type = INT8OID; }
		else return false;
		if (var->vartype == FLOAT4OID) return false;          /* [PG] sum(float4) accumulates in float4: not what the kernels do */
	}
	else if (strcmp(name, "min") == 0 || strcmp(name, "max") == 0)
	{
		spec->kind = name[1] == 'i' ? CG_AGG_MIN : CG_AGG_MAX;
		out->kind = GPU_OUT_MINMAX; out->type = var->vartype;
	}
	else
		return false;                                           /* anything else keeps the reference's executor */
	desc->naggs++;
	return true;
}

static bool
translate_query(PlannerInfo *root, RelOptInfo *input_rel, GpuAggPlan *plan)
{
	Query *q = root->parse;
	if (!q->hasAggs || q->hasWindowFuncs || q->hasDistinctOn || q->distinctClause != NIL || q->groupingSets != NIL || q->havingQual != NULL)
		return false;
	if (input_rel->reloptkind != RELOPT_BASEREL)
		return false;
	RangeTblEntry *rte = planner_rt_fetch(input_rel->relid, root);
	if (rte->rtekind != RTE_RELATION || !IsColumnarTableAmTable(rte->relid))
		return false;
	memset(plan, 0, sizeof *plan);
	plan->relid = rte->relid;
	plan->desc.enable_qual_pushdown = EnableColumnarQualPushdown ? 1 : 0;
	Relation rel = table_open(rte->relid, AccessShareLock);
	TupleDesc tupdesc = RelationGetDescr(rel);
	bool ok = true;
	ListCell *lc;
	/* WHERE: the implicit AND-list of restriction clauses */
	int nclauses = 0;
	foreach(lc, input_rel->baserestrictinfo)
	{
		RestrictInfo *ri = (RestrictInfo *) lfirst(lc);
		if (!(ok = translate_where((Node *) ri->clause, tupdesc, &plan->desc)))
			break;
		if (nclauses++ > 0)
		{
			if (plan->desc.nqual_expr >= CG_MAX_QEXPR) { ok = false; break; }
			plan->desc.qual_expr[plan->desc.nqual_expr++] = CG_QX_AND;
		}
	}
	/* GROUP BY plain integer columns */
	foreach(lc, q->groupClause)
	{
		if (!ok) break;
		SortGroupClause *sgc = (SortGroupClause *) lfirst(lc);
		TargetEntry *tle = get_sortgroupref_tle(sgc->tleSortGroupRef, q->targetList);
		Var *var = strip_to_var((Node *) tle->expr);
		int attlen, type_class;
		if (var == NULL || plan->desc.ngroup_cols >= CG_MAX_GROUP_COLS || !type_is_supported(var->vartype, &attlen, &type_class) ||
			type_class != CG_TYPE_INT)
			ok = false;
		else
			plan->desc.group_cols[plan->desc.ngroup_cols++] = var->varattno - 1;
	}
	/* target list: group columns and aggregates, in output order */
	foreach(lc, q->targetList)
	{
		if (!ok) break;
		TargetEntry *tle = (TargetEntry *) lfirst(lc);
		if (tle->resjunk || plan->nout >= (int) lengthof(plan->out)) { ok = false; break; }
		GpuOutColumn *out = &plan->out[plan->nout];
		if (IsA(tle->expr, Aggref))
			ok = translate_aggref((Aggref *) tle->expr, tupdesc, &plan->desc, out);
		else
		{
			Var *var = strip_to_var((Node *) tle->expr);
			ok = false;
			for (int g = 0; var != NULL && g < plan->desc.ngroup_cols; g++)
				if (plan->desc.group_cols[g] == var->varattno - 1)
				{
					out->kind = GPU_OUT_KEY; out->index = g; out->type = var->vartype; ok = true;
				}
		}
		plan->nout++;
	}
	table_close(rel, AccessShareLock);
	return ok && plan->desc.naggs > 0;
}

/* ------------------------------------------------------------------------------------ *
 *  path / plan / state creation
 * ------------------------------------------------------------------------------------ */
static Plan *GpuColumnarAgg_PlanCustomPath(PlannerInfo *root, RelOptInfo *rel, CustomPath *best_path, List *tlist, List *clauses, List *custom_plans);
static Node *GpuColumnarAgg_CreateCustomScanState(CustomScan *cscan);
static void GpuColumnarAgg_Begin(CustomScanState *node, EState *estate, int eflags);
static TupleTableSlot *GpuColumnarAgg_Exec(CustomScanState *node);
static void GpuColumnarAgg_End(CustomScanState *node);
static void GpuColumnarAgg_ReScan(CustomScanState *node);
static void GpuColumnarAgg_Explain(CustomScanState *node, List *ancestors, ExplainState *es);

static const CustomPathMethods GpuColumnarAggPathMethods = {
	.CustomName = "GpuColumnarAgg",
	.PlanCustomPath = GpuColumnarAgg_PlanCustomPath,
};

static const CustomScanMethods GpuColumnarAggScanMethods = {
	.CustomName = "GpuColumnarAgg",
	.CreateCustomScanState = GpuColumnarAgg_CreateCustomScanState,
};

static const CustomExecMethods GpuColumnarAggExecMethods = {
	.CustomName = "GpuColumnarAgg",
	.BeginCustomScan = GpuColumnarAgg_Begin,
	.ExecCustomScan = GpuColumnarAgg_Exec,
	.EndCustomScan = GpuColumnarAgg_End,
	.ReScanCustomScan = GpuColumnarAgg_ReScan,
	.ExplainCustomScan = GpuColumnarAgg_Explain,
};

static void
GpuColumnarAggUpperPaths(PlannerInfo *root, UpperRelationKind stage, RelOptInfo *input_rel, RelOptInfo *output_rel, void *extra)
{
	if (prev_create_upper_paths_hook)
		prev_create_upper_paths_hook(root, stage, input_rel, output_rel, extra);
	if (stage != UPPERREL_GROUP_AGG)
		return;
	GpuAggPlan *plan = palloc0(sizeof(GpuAggPlan));
	if (!translate_query(root, input_rel, plan))
		return;
	CustomPath *cpath = makeNode(CustomPath);
	cpath->path.pathtype = T_CustomScan;
	cpath->path.parent = output_rel;
	cpath->path.pathtarget = output_rel->reltarget;
	cpath->path.rows = output_rel->rows;
	/* one kernel pass at HBM speed: cheaper than any row-at-a-time plan over the same relation */
	cpath->path.startup_cost = 0;
	cpath->path.total_cost = input_rel->rows * 1e-6;
	cpath->methods = &GpuColumnarAggPathMethods;
	bytea *blob = palloc(VARHDRSZ + sizeof(GpuAggPlan));
	SET_VARSIZE(blob, VARHDRSZ + sizeof(GpuAggPlan));
	memcpy(VARDATA(blob), plan, sizeof(GpuAggPlan));
	cpath->custom_private = list_make1(makeConst(BYTEAOID, -1, InvalidOid, -1, PointerGetDatum(blob), false, false));
	add_path(output_rel, (Path *) cpath);
}

static Plan *
GpuColumnarAgg_PlanCustomPath(PlannerInfo *root, RelOptInfo *rel, CustomPath *best_path, List *tlist, List *clauses, List *custom_plans)
{
	CustomScan *cscan = makeNode(CustomScan);
	cscan->scan.plan.targetlist = tlist;
	cscan->scan.scanrelid = 0;                       /* not a scan of one base relation's tuples: it emits aggregate rows */
	cscan->custom_scan_tlist = tlist;
	cscan->custom_private = best_path->custom_private;
	cscan->methods = &GpuColumnarAggScanMethods;
	(void) root; (void) rel; (void) clauses; (void) custom_plans;
	return (Plan *) cscan;
}

static Node *
GpuColumnarAgg_CreateCustomScanState(CustomScan *cscan)
{
	GpuColumnarAggState *state = (GpuColumnarAggState *) newNodeImpl(sizeof(GpuColumnarAggState), T_CustomScanState);
	state->css.methods = &GpuColumnarAggExecMethods;
	Const *blob = (Const *) linitial(cscan->custom_private);
	memcpy(&state->plan, VARDATA(DatumGetPointer(blob->constvalue)), sizeof(GpuAggPlan));
	return (Node *) state;
}

/* ------------------------------------------------------------------------------------ *
 *  execution
 * ------------------------------------------------------------------------------------ */
static int64
datum_to_i64(Datum d, Oid type)
{
	switch (type)
	{
		case INT8OID: case TIMESTAMPOID: case TIMESTAMPTZOID: return DatumGetInt64(d);
		case INT4OID: case DATEOID: return (int64) DatumGetInt32(d);
		case INT2OID: return (int64) DatumGetInt16(d);
		case FLOAT8OID: { double x = DatumGetFloat8(d); int64 v; memcpy(&v, &x, sizeof v); return v; }
		case FLOAT4OID: { double x = (double) DatumGetFloat4(d); int64 v; memcpy(&v, &x, sizeof v); return v; }
		default: return 0;
	}
}

/*
 * The relation as the library wants it (CgRelation): the main fork's pages as one image + the stripes visible to the
 * snapshot + their skip lists.  CgSkipNode / CgStripe are field-for-field conversions of ColumnChunkSkipNode /
 * StripeMetadata (different order and widths: this function IS the mapping).
 */
static void
build_relation_image(Relation rel, Snapshot snapshot, CgRelation *out)
{
	TupleDesc tupdesc = RelationGetDescr(rel);
	int natts = tupdesc->natts;
	CgColumnDesc *cols = palloc0(sizeof(CgColumnDesc) * natts);
	for (int c = 0; c < natts; c++)
	{
		int attlen = 8, type_class = CG_TYPE_INT;
		Form_pg_attribute att = TupleDescAttr(tupdesc, c);
		if (att->attisdropped || !type_is_supported(att->atttypid, &attlen, &type_class))
			attlen = att->attlen > 0 && att->attlen <= 8 && att->attbyval ? att->attlen : 8;   /* never read: only planned columns are staged */
		cols[c].attlen = attlen;
		cols[c].type_class = type_class;
	}
	List *stripes = StripesForRelfilelocator(rel);           /* snapshot-aware catalog scan of columnar.stripe */
	int nstripes = 0, nnodes = 0;
	ListCell *lc;
	foreach(lc, stripes)
	{
		StripeMetadata *s = (StripeMetadata *) lfirst(lc);
		if (StripeWriteState(s) != STRIPE_WRITE_FLUSHED)     /* columnar_reader.c:721-728 */
			continue;
		nstripes++;
		nnodes += (int) (s->columnCount * s->chunkCount);
	}
	CgStripe *cs = palloc0(sizeof(CgStripe) * Max(nstripes, 1));
	CgSkipNode *cn = palloc0(sizeof(CgSkipNode) * Max(nnodes, 1));
	int si = 0, base = 0;
	foreach(lc, stripes)
	{
		StripeMetadata *s = (StripeMetadata *) lfirst(lc);
		if (StripeWriteState(s) != STRIPE_WRITE_FLUSHED)
			continue;
		StripeSkipList *skip = ReadStripeSkipList(rel, s->id, tupdesc, s->chunkCount, snapshot);
		CgStripe *d = &cs[si++];
		d->id = s->id; d->file_offset = s->fileOffset; d->data_length = s->dataLength; d->row_count = s->rowCount;
		d->first_row_number = s->firstRowNumber; d->column_count = s->columnCount; d->chunk_row_count = s->chunkGroupRowCount;
		d->chunk_count = s->chunkCount; d->skipnode_base = (uint32) base;
		for (uint32 c = 0; c < s->columnCount; c++)
			for (uint32 k = 0; k < s->chunkCount; k++)
			{
				const ColumnChunkSkipNode *n = &skip->chunkSkipNodeArray[c][k];
				CgSkipNode *o = &cn[base + c * s->chunkCount + k];
				Oid type = TupleDescAttr(tupdesc, c)->atttypid;
				int attlen, type_class;
				o->has_minmax = n->hasMinMax && type_is_supported(type, &attlen, &type_class);
				o->min_value = o->has_minmax ? datum_to_i64(n->minimumValue, type) : 0;
				o->max_value = o->has_minmax ? datum_to_i64(n->maximumValue, type) : 0;
				o->row_count = n->rowCount;
				o->value_offset = n->valueChunkOffset; o->value_length = n->valueLength;
				o->exists_offset = n->existsChunkOffset; o->exists_length = n->existsLength;
				o->decompressed_size = n->decompressedValueSize;
				o->compression_type = (int32) n->valueCompressionType;
				o->compression_level = n->valueCompressionLevel;
			}
		base += (int) (s->columnCount * s->chunkCount);
	}
	/* the pages, as they sit in shared_buffers: one share-locked copy per block */
	BlockNumber nblocks = RelationGetNumberOfBlocks(rel);
	uint8 *pages = palloc((Size) Max(nblocks, 1) * BLCKSZ);
	for (BlockNumber b = 0; b < nblocks; b++)
	{
		Buffer buf = ReadBufferExtended(rel, MAIN_FORKNUM, b, RBM_NORMAL, NULL);
		LockBuffer(buf, BUFFER_LOCK_SHARE);
		memcpy(pages + (Size) b * BLCKSZ, BufferGetPage(buf), BLCKSZ);
		UnlockReleaseBuffer(buf);
	}
	out->pages = pages; out->nblocks = nblocks;
	out->stripes = cs; out->nstripes = nstripes;
	out->nodes = cn; out->nnodes = nnodes;
	out->columns = cols; out->natts = natts;
}

static void
GpuColumnarAgg_Begin(CustomScanState *node, EState *estate, int eflags)
{
	(void) node; (void) estate; (void) eflags;      /* everything happens at the first Exec call: EXPLAIN without ANALYZE must not touch the GPU */
}

static void
run_scan(GpuColumnarAggState *state)
{
	EState *estate = state->css.ss.ps.state;
	MemoryContext old = MemoryContextSwitchTo(estate->es_query_cxt);
	gpu_ensure_device();
	Relation rel = table_open(state->plan.relid, AccessShareLock);
	CgRelation image;
	build_relation_image(rel, estate->es_snapshot ? estate->es_snapshot : GetActiveSnapshot(), &image);
	CgScanDesc *desc = &state->plan.desc;
	int64 key_min = 0, key_max = -1, rows = 0;
	int64 bounds[CG_MAX_AGGS];
	gpu_check(cg_relation_bounds(&image, desc, &key_min, &key_max, bounds, &rows));
	for (int a = 0; a < desc->naggs; a++)
		desc->aggs[a].term_abs_bound = bounds[a];
	gpu_check(cg_partial_create(desc, image.columns, image.natts, key_min, key_max, Max(rows, 1), &state->partial));
	int rc = cg_scan_relation(&image, desc, state->partial, &state->stats);
	if (rc == CG_ERETRY_UNPACKED)
	{
		/* an optimistic packed accumulator overflowed: nothing wrong was produced; rescan with wide accumulators */
		gpu_check(cg_partial_set_packing(state->partial, 0));
		gpu_check(cg_partial_reset(state->partial));
		rc = cg_scan_relation(&image, desc, state->partial, &state->stats);
	}
	gpu_check(rc);
	int64 n = 0;
	gpu_check(cg_partial_ngroups(state->partial, &n));
	Size cells = (Size) Max(n, 1) * Max(desc->naggs, 1);
	state->keys = palloc(sizeof(int64) * Max(n, 1)); state->key_nulls = palloc(Max(n, 1));
	state->sum_hi = palloc(sizeof(int64) * cells); state->sum_lo = palloc(sizeof(uint64) * cells);
	state->count = palloc(sizeof(int64) * cells); state->minmax = palloc(sizeof(int64) * cells);
	state->fsum = palloc(sizeof(double) * cells);
	gpu_check(cg_partial_fetch(state->partial, Max(n, 1), state->keys, state->key_nulls, state->sum_hi, state->sum_lo, state->count,
							   state->minmax, state->fsum, &state->ngroups));
	table_close(rel, AccessShareLock);
	state->next = 0;
	state->done_scan = true;
	MemoryContextSwitchTo(old);
}

static Datum
i64_to_datum(int64 v, Oid type)
{
	switch (type)
	{
		case INT4OID: case DATEOID: return Int32GetDatum((int32) v);
		case INT2OID: return Int16GetDatum((int16) v);
		case FLOAT8OID: { double x; memcpy(&x, &v, sizeof x); return Float8GetDatum(x); }
		default: return Int64GetDatum(v);
	}
}

static TupleTableSlot *
GpuColumnarAgg_Exec(CustomScanState *node)
{
	GpuColumnarAggState *state = (GpuColumnarAggState *) node;
	if (!state->done_scan)
		run_scan(state);
	TupleTableSlot *slot = node->ss.ps.ps_ResultTupleSlot;
	ExecClearTuple(slot);
	if (state->next >= state->ngroups)
		return slot;                                   /* empty slot = end of scan */
	const int64 g = state->next++;
	const int na = state->plan.desc.naggs;
	for (int i = 0; i < state->plan.nout; i++)
	{
		const GpuOutColumn *out = &state->plan.out[i];
		const int64 cell = g * na + out->index;
		slot->tts_isnull[i] = false;
		switch ((GpuOutKind) out->kind)
		{
			case GPU_OUT_KEY:
			{
				/* two group columns travel packed: low 32 bits the first, high 32 bits the second */
				int64 k = state->keys[g];
				if (state->plan.desc.ngroup_cols == 2)
					k = out->index == 0 ? (int64) (int32) (uint32) k : (int64) (int32) ((uint64) k >> 32);
				slot->tts_isnull[i] = state->key_nulls[g] != 0;
				slot->tts_values[i] = i64_to_datum(k, out->type);
				break;
			}
			case GPU_OUT_COUNT:
				slot->tts_values[i] = Int64GetDatum(state->count[cell]);
				break;
			case GPU_OUT_SUM_NUMERIC:
			{
				char buf[64];
				if (state->count[cell] == 0) { slot->tts_isnull[i] = true; break; }      /* sum over no non-NULL input is NULL */
				gpu_check(cg_numeric_out(state->sum_hi[cell], state->sum_lo[cell], out->scale, buf, sizeof buf));
				slot->tts_values[i] = DirectFunctionCall3(numeric_in, CStringGetDatum(buf), ObjectIdGetDatum(InvalidOid), Int32GetDatum(-1));
				break;
			}
			case GPU_OUT_SUM_INT8:
				if (state->count[cell] == 0) { slot->tts_isnull[i] = true; break; }
				slot->tts_values[i] = Int64GetDatum((int64) state->sum_lo[cell]);
				break;
			case GPU_OUT_SUM_FLOAT8:
				if (state->count[cell] == 0) { slot->tts_isnull[i] = true; break; }
				slot->tts_values[i] = Float8GetDatum(state->fsum[cell]);
				break;
			case GPU_OUT_MINMAX:
				if (state->count[cell] == 0) { slot->tts_isnull[i] = true; break; }
				slot->tts_values[i] = i64_to_datum(state->minmax[cell], out->type);
				break;
		}
	}
	return ExecStoreVirtualTuple(slot);
}

static void
GpuColumnarAgg_End(CustomScanState *node)
{
	GpuColumnarAggState *state = (GpuColumnarAggState *) node;
	if (state->partial)
		cg_partial_free(state->partial);
	state->partial = NULL;
}

static void
GpuColumnarAgg_ReScan(CustomScanState *node)
{
	GpuColumnarAggState *state = (GpuColumnarAggState *) node;
	state->next = 0;                                   /* the aggregate rows are still there */
}

/* the counters ColumnarScan's EXPLAIN ANALYZE prints (columnar_customscan.c:1966-1999), plus the device's */
static void
GpuColumnarAgg_Explain(CustomScanState *node, List *ancestors, ExplainState *es)
{
	GpuColumnarAggState *state = (GpuColumnarAggState *) node;
	(void) ancestors;
	if (!state->done_scan)
		return;
	ExplainPropertyInteger("Rows Removed by Filter", NULL, state->stats.rows_removed_by_filter, es);
	ExplainPropertyInteger("Columnar Chunk Groups Removed by Filter", NULL, state->stats.chunk_groups_filtered, es);
	ExplainPropertyInteger("GPU Bytes Scanned", "B", state->stats.bytes_scanned, es);
	ExplainPropertyFloat("GPU Kernel Time", "ms", state->stats.kernel_ms, 3, es);
}

void
_PG_init(void)
{
	RegisterCustomScanMethods(&GpuColumnarAggScanMethods);
	prev_create_upper_paths_hook = create_upper_paths_hook;
	create_upper_paths_hook = GpuColumnarAggUpperPaths;
}
