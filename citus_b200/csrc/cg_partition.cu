/*
 * cg_partition.cu -- K6: the map side of the hash repartition.
 *
 * Per row the reference does (executor/partitioned_intermediate_results.c:493-553
 * PartitionedResultDestReceiverReceive):
 *    NULL key              -> partition 0
 *    FindShardInterval     -> hashint4 / hashint8 of the key through fmgr
 *                             (utils/shardinterval_utils.c:260-282), then a binary search
 *                             of the int4 token ranges (SearchCachedShardInterval :373-414;
 *                             the synthetic intervals of planner/multi_physical_planner.c:
 *                             4667-4701 are passed as min/max arrays)
 *    receiveSlot           -> append the row to that partition's file
 * Here: one pass computes the partition index of every row and a histogram (per-block
 * counts in shared memory, one global atomic per block and partition), a second pass
 * scatters the payload columns into partition-contiguous order -- stable inside a
 * partition, like the reference's append order -- using per-block offsets.
 *
 * [PG] hash_bytes_uint32 / hashint4 / hashint8: src/common/hashfn.c,
 * src/backend/access/hash/hashfunc.c (Jenkins lookup3 final mix); pinned by the
 * reference goldens through the oracle (tests/test_oracle_golden.py).
 */
#include "cg_internal.h"

#define CGP_THREADS 256
#define CGP_ROWS_PER_BLOCK 4096
#define CGP_MAX_P 1024

__device__ __forceinline__ uint32_t rot32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

__device__ __forceinline__ uint32_t hash_bytes_uint32(uint32_t k)
{
	uint32_t a, b, c;
	a = b = c = 0x9e3779b9u + 4u + 3923095u;
	a += k;
	c ^= b; c -= rot32(b, 14);
	a ^= c; a -= rot32(c, 11);
	b ^= a; b -= rot32(a, 25);
	c ^= b; c -= rot32(b, 16);
	a ^= c; a -= rot32(c, 4);
	b ^= a; b -= rot32(a, 14);
	c ^= b; c -= rot32(b, 24);
	return c;
}

__device__ __forceinline__ int32_t hashint8_dev(int64_t v)
{
	uint32_t lo = (uint32_t) v, hi = (uint32_t) ((uint64_t) v >> 32);
	lo ^= (v >= 0) ? hi : ~hi;
	return (int32_t) hash_bytes_uint32(lo);
}

struct PartParams
{
	const int64_t *keys;
	const uint8_t *nulls;
	int64_t n;
	int32_t key_len;
	int32_t by_hash;
	int32_t P;
	const int32_t *mins;   /* device */
	const int32_t *maxs;
	int32_t *index;
	unsigned long long *counts;        /* [P] */
	unsigned long long *block_counts;  /* [nblocks][P] (may be NULL) */
	unsigned long long *errors;
};

__global__ void __launch_bounds__(CGP_THREADS)
cg_partition_index_kernel(const __grid_constant__ PartParams A)
{
	extern __shared__ unsigned int s_count[];   /* [P] counts, then [P] mins, [P] maxs */
	int32_t *s_min = (int32_t *) (s_count + A.P);
	int32_t *s_max = s_min + A.P;
	for (int p = threadIdx.x; p < A.P; p += CGP_THREADS)
	{
		s_count[p] = 0;
		s_min[p] = A.mins[p];
		s_max[p] = A.maxs[p];
	}
	__syncthreads();
	int64_t base = (int64_t) blockIdx.x * CGP_ROWS_PER_BLOCK;
	for (int i = threadIdx.x; i < CGP_ROWS_PER_BLOCK; i += CGP_THREADS)
	{
		int64_t r = base + i;
		if (r >= A.n) break;
		int idx;
		if (A.nulls && A.nulls[r]) idx = 0;
		else
		{
			int64_t k = A.keys[r];
			int32_t searched = A.by_hash ? (A.key_len == 4 ? (int32_t) hash_bytes_uint32((uint32_t) (int32_t) k) : hashint8_dev(k))
										 : (int32_t) k;
			/* SearchCachedShardInterval */
			int lower = 0, upper = A.P;
			idx = -1;
			while (lower < upper)
			{
				int middle = (lower + upper) / 2;
				if (searched < s_min[middle]) { upper = middle; continue; }
				if (searched <= s_max[middle]) { idx = middle; break; }
				lower = middle + 1;
			}
			if (idx < 0)
			{
				atomicAdd(A.errors, 1ull);   /* "could not find shard for partition column value" */
				idx = 0;
			}
		}
		A.index[r] = idx;
		atomicAdd(&s_count[idx], 1u);
	}
	__syncthreads();
	for (int p = threadIdx.x; p < A.P; p += CGP_THREADS)
	{
		unsigned int c = s_count[p];
		if (A.block_counts) A.block_counts[(uint64_t) blockIdx.x * A.P + p] = c;
		if (c) atomicAdd(A.counts + p, (unsigned long long) c);
	}
}

/* exclusive scan over blocks, per partition: block_counts[b][p] -> start offset of block b inside
 * partition p (plus the partition's base).  P threads, sequential over blocks (nblocks ~ n/4096). */
__global__ void cg_partition_scan_kernel(unsigned long long *block_counts, const unsigned long long *part_base,
										 int64_t nblocks, int P)
{
	int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= P) return;
	unsigned long long run = part_base[p];
	for (int64_t b = 0; b < nblocks; b++)
	{
		unsigned long long c = block_counts[b * P + p];
		block_counts[b * P + p] = run;
		run += c;
	}
}

struct ScatterParams
{
	const int32_t *index;
	int64_t n;
	int32_t P;
	int32_t ncols;
	const unsigned long long *block_offsets;   /* [nblocks][P] */
	const int64_t *cols[8];
	int64_t *out[8];
};

/* stable scatter: inside a block, rows of the same partition keep their order (rank by a
 * warp-ordered count over the block's rows) */
__global__ void __launch_bounds__(CGP_THREADS)
cg_partition_scatter_kernel(const __grid_constant__ ScatterParams A)
{
	extern __shared__ unsigned int s_run[];    /* [P] running count inside this block */
	for (int p = threadIdx.x; p < A.P; p += CGP_THREADS) s_run[p] = 0;
	__syncthreads();
	int64_t base = (int64_t) blockIdx.x * CGP_ROWS_PER_BLOCK;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	/* process the block's rows in order, 256 at a time; inside a step warps are serialised
	 * through shared memory counters so that the output order equals the input order */
	for (int step = 0; step < CGP_ROWS_PER_BLOCK; step += CGP_THREADS)
	{
		int64_t r = base + step + threadIdx.x;
		bool valid = r < A.n;
		int idx = valid ? A.index[r] : -1;
		/* rank among the lanes of my warp with the same partition and a lower lane id */
		unsigned peers = __match_any_sync(0xffffffffu, idx);
		unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
		unsigned leader = __ffs(peers) - 1;
		unsigned warp_count = __popc(peers);
		for (unsigned w = 0; w < CGP_THREADS / 32; w++)
		{
			unsigned mybase = 0;
			if (w == warp && valid && lane == leader)
			{
				mybase = s_run[idx];
				s_run[idx] = mybase + warp_count;
			}
			__syncthreads();
			if (w == warp)
			{
				mybase = __shfl_sync(0xffffffffu, mybase, leader);
				if (valid)
				{
					unsigned long long dst = A.block_offsets[(uint64_t) blockIdx.x * A.P + idx] + mybase + rank_in_warp;
					for (int c = 0; c < A.ncols; c++) A.out[c][dst] = A.cols[c][r];
				}
			}
		}
		if (base + step + CGP_THREADS >= A.n) break;
	}
}

static int32_t *g_d_bounds = nullptr;
static int g_bounds_cap = 0;

extern "C" int cg_partition_index(const int64_t *d_keys, const uint8_t *d_nulls, int64_t n, int32_t key_len,
								  int32_t by_hash, const int32_t *mins, const int32_t *maxs, int32_t P,
								  int32_t *d_index, int64_t *d_counts)
{
	CgContext *ctx = cg_ctx();
	if (!ctx) return CG_EINVAL;
	if (P <= 0) return cg_set_error(CG_EINVAL, "number of partitions cannot be 0");
	if (P > CGP_MAX_P) return cg_set_error(CG_EUNSUPPORTED, "more than %d partitions", CGP_MAX_P);
	if (key_len != 4 && key_len != 8) return cg_set_error(CG_EINVAL, "key_len must be 4 or 8");
	if (n < 0 || !d_keys || !d_index || !d_counts || !mins || !maxs) return cg_set_error(CG_EINVAL, "bad argument");
	if (g_bounds_cap < 2 * P + 2)
	{
		cudaFree(g_d_bounds);
		CG_CUDA(cudaMalloc(&g_d_bounds, sizeof(int32_t) * (2 * CGP_MAX_P + 2)));
		g_bounds_cap = 2 * CGP_MAX_P + 2;
	}
	CG_CUDA(cudaMemcpyAsync(g_d_bounds, mins, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->compute));
	CG_CUDA(cudaMemcpyAsync(g_d_bounds + CGP_MAX_P, maxs, sizeof(int32_t) * P, cudaMemcpyHostToDevice, ctx->compute));
	unsigned long long *d_err = nullptr;
	CG_CUDA(cudaMallocAsync((void **) &d_err, sizeof(unsigned long long), ctx->compute));
	CG_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), ctx->compute));
	CG_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int64_t) * P, ctx->compute));
	if (n > 0)
	{
		PartParams A;
		A.keys = d_keys; A.nulls = d_nulls; A.n = n; A.key_len = key_len; A.by_hash = by_hash; A.P = P;
		A.mins = g_d_bounds; A.maxs = g_d_bounds + CGP_MAX_P; A.index = d_index;
		A.counts = (unsigned long long *) d_counts; A.block_counts = nullptr; A.errors = d_err;
		int64_t nblocks = (n + CGP_ROWS_PER_BLOCK - 1) / CGP_ROWS_PER_BLOCK;
		cg_partition_index_kernel<<<(unsigned) nblocks, CGP_THREADS, 3 * P * sizeof(int32_t), ctx->compute>>>(A);
		CG_CUDA(cudaGetLastError());
	}
	unsigned long long err = 0;
	CG_CUDA(cudaMemcpyAsync(&err, d_err, sizeof err, cudaMemcpyDeviceToHost, ctx->compute));
	CG_CUDA(cudaStreamSynchronize(ctx->compute));
	CG_CUDA(cudaFreeAsync(d_err, ctx->compute));
	if (err) return cg_set_error(CG_EINVAL, "could not find shard for partition column value (%llu rows)", err);
	return CG_OK;
}

__global__ void cg_block_count_kernel(const int32_t *index, int64_t n, int P, unsigned long long *block_counts)
{
	extern __shared__ unsigned int s_count[];
	for (int p = threadIdx.x; p < P; p += CGP_THREADS) s_count[p] = 0;
	__syncthreads();
	int64_t base = (int64_t) blockIdx.x * CGP_ROWS_PER_BLOCK;
	for (int i = threadIdx.x; i < CGP_ROWS_PER_BLOCK; i += CGP_THREADS)
	{
		int64_t r = base + i;
		if (r >= n) break;
		atomicAdd(&s_count[index[r]], 1u);
	}
	__syncthreads();
	for (int p = threadIdx.x; p < P; p += CGP_THREADS) block_counts[(uint64_t) blockIdx.x * P + p] = s_count[p];
}

extern "C" int cg_partition_scatter(const int32_t *d_index, int64_t n, int32_t P, const int64_t *const *d_cols,
									int32_t ncols, int64_t *const *d_out, int64_t *h_offsets)
{
	CgContext *ctx = cg_ctx();
	if (!ctx) return CG_EINVAL;
	if (P <= 0 || P > CGP_MAX_P) return cg_set_error(CG_EINVAL, "bad partition count %d", P);
	if (ncols < 0 || ncols > 8) return cg_set_error(CG_EUNSUPPORTED, "at most 8 payload columns per call");
	if (n < 0) return cg_set_error(CG_EINVAL, "negative row count");
	int64_t nblocks = (n + CGP_ROWS_PER_BLOCK - 1) / CGP_ROWS_PER_BLOCK;
	unsigned long long *d_block = nullptr, *d_base = nullptr;
	CG_CUDA(cudaMallocAsync((void **) &d_block, sizeof(unsigned long long) * (size_t) std::max<int64_t>(nblocks, 1) * P, ctx->compute));
	CG_CUDA(cudaMallocAsync((void **) &d_base, sizeof(unsigned long long) * P, ctx->compute));
	std::vector<unsigned long long> totals(P, 0), base(P, 0);
	if (n > 0)
	{
		cg_block_count_kernel<<<(unsigned) nblocks, CGP_THREADS, P * sizeof(unsigned int), ctx->compute>>>(d_index, n, P, d_block);
		CG_CUDA(cudaGetLastError());
		/* totals per partition: run the scan with zero bases first, reading the last block's end */
		CG_CUDA(cudaMemsetAsync(d_base, 0, sizeof(unsigned long long) * P, ctx->compute));
		/* column sums on the host for the partition bases (P is small) */
		std::vector<unsigned long long> bc((size_t) nblocks * P);
		CG_CUDA(cudaMemcpyAsync(bc.data(), d_block, bc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
		CG_CUDA(cudaStreamSynchronize(ctx->compute));
		for (int64_t b = 0; b < nblocks; b++)
			for (int p = 0; p < P; p++) totals[p] += bc[(size_t) b * P + p];
	}
	unsigned long long run = 0;
	for (int p = 0; p < P; p++) { base[p] = run; h_offsets[p] = (int64_t) run; run += totals[p]; }
	h_offsets[P] = (int64_t) run;
	if (n > 0)
	{
		CG_CUDA(cudaMemcpyAsync(d_base, base.data(), sizeof(unsigned long long) * P, cudaMemcpyHostToDevice, ctx->compute));
		cg_partition_scan_kernel<<<(P + 127) / 128, 128, 0, ctx->compute>>>(d_block, d_base, nblocks, P);
		CG_CUDA(cudaGetLastError());
		ScatterParams S;
		memset(&S, 0, sizeof S);
		S.index = d_index; S.n = n; S.P = P; S.ncols = ncols; S.block_offsets = d_block;
		for (int c = 0; c < ncols; c++) { S.cols[c] = d_cols[c]; S.out[c] = d_out[c]; }
		cg_partition_scatter_kernel<<<(unsigned) nblocks, CGP_THREADS, P * sizeof(unsigned int), ctx->compute>>>(S);
		CG_CUDA(cudaGetLastError());
	}
	CG_CUDA(cudaStreamSynchronize(ctx->compute));
	CG_CUDA(cudaFreeAsync(d_block, ctx->compute));
	CG_CUDA(cudaFreeAsync(d_base, ctx->compute));
	return CG_OK;
}
