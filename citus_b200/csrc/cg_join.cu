/*
 * cg_join.cu -- the merge side of a dual-repartition join, for aggregate queries (K8).
 *
 * After worker_partition_query_result has routed both inputs by hash(join key), a MERGE task
 * reads the co-located partitions back with read_intermediate_results() and joins them with
 * PostgreSQL's executor (planner/multi_physical_planner.c:4304-4328 builds the function RTEs of
 * the merge query; executor/intermediate_results.c:789-1045 feeds them), one hash-table insert
 * per build row and one probe per probe row, then aggregates the joined rows.  With the
 * partitions already in HBM (the all-to-all delivered them), this file computes
 *
 *     SELECT count(*), sum(b.payload + p.payload) FROM build b JOIN probe p USING (key)
 *
 * -- the C4 check query of SURVEY.md 8(d) -- without materialising the joined rows: the build
 * side is aggregated per key while it is inserted (rows, sum of payload), and a probe row with
 * payload y that finds (c, s) contributes c joined rows and s + y * c to the sum, which is exactly
 * the sum over the c rows PostgreSQL's HashJoin would emit for it.  SQL semantics: a NULL key
 * joins nothing.  Sums are exact (128-bit).
 *
 * Open addressing over a separate key array (like the GROUP BY table in cg_scan.cu), load factor
 * <= 1/2, 64-bit CAS to claim a slot; values are one 32-byte sector per slot
 * [rows, sum of low 32 bits, sum of payload >> 32, unused].
 */
#include <vector>

#include "cg_internal.h"

#define CGJ_THREADS 256
#define CGJ_EMPTY ((long long) 0x8000000000000000ull)

struct JoinTable
{
	long long *keys;              /* [cap + 1]; slot cap = the key equal to the EMPTY sentinel */
	unsigned long long *vals;     /* [cap + 1][4] */
	unsigned long long cap;       /* power of two */
	int shift;                    /* 64 - log2(cap) */
	unsigned long long *errors;
};

__global__ void cg_join_init_kernel(long long *keys, unsigned long long n)
{
	for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long) gridDim.x * blockDim.x)
		keys[i] = CGJ_EMPTY;
}

__device__ __forceinline__ unsigned long long join_home(long long key, int shift)
{
	return ((unsigned long long) key * 0x9E3779B97F4A7C15ull) >> shift;
}

__global__ void __launch_bounds__(CGJ_THREADS)
cg_join_build_kernel(JoinTable T, const long long *keys, const uint8_t *nulls, const long long *payload, long long n)
{
	for (long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long) gridDim.x * blockDim.x)
	{
		if (nulls && nulls[r]) continue;
		const long long key = keys[r];
		unsigned long long slot = T.cap;
		if (key != CGJ_EMPTY)
		{
			unsigned long long h = join_home(key, T.shift);
			slot = ~0ull;
			for (unsigned long long probes = 0; probes <= T.cap; probes++)
			{
				long long cur = (long long) __ldcg((unsigned long long *) T.keys + h);
				if (cur == key) { slot = h; break; }
				if (cur == CGJ_EMPTY)
				{
					long long old = (long long) atomicCAS((unsigned long long *) T.keys + h, (unsigned long long) CGJ_EMPTY, (unsigned long long) key);
					if (old == CGJ_EMPTY || old == key) { slot = h; break; }
				}
				h = (h + 1) & (T.cap - 1);
			}
			if (slot == ~0ull) { atomicAdd(T.errors, 1ull); continue; }
		}
		const long long x = payload[r];
		unsigned long long *v = T.vals + slot * 4;
		atomicAdd(v + 0, 1ull);
		atomicAdd(v + 1, (unsigned long long) (unsigned int) x);
		atomicAdd(v + 2, (unsigned long long) (x >> 32));
	}
}

__device__ __forceinline__ unsigned long long join_wsum(unsigned long long x)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
	return x;
}

/* out[0] = joined rows; out[1..4] = the sum in 32-bit chunks (chunk 3 sign-extended), added without carries */
__global__ void __launch_bounds__(CGJ_THREADS)
cg_join_probe_kernel(JoinTable T, const long long *keys, const uint8_t *nulls, const long long *payload, long long n,
					 unsigned long long *out)
{
	unsigned long long matches = 0;
	unsigned __int128 sum = 0;
	for (long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long) gridDim.x * blockDim.x)
	{
		if (nulls && nulls[r]) continue;
		const long long key = keys[r];
		unsigned long long slot = T.cap;
		if (key != CGJ_EMPTY)
		{
			unsigned long long h = join_home(key, T.shift);
			slot = ~0ull;
			for (unsigned long long probes = 0; probes <= T.cap; probes++)
			{
				long long cur = (long long) __ldg((const unsigned long long *) T.keys + h);
				if (cur == key) { slot = h; break; }
				if (cur == CGJ_EMPTY) break;
				h = (h + 1) & (T.cap - 1);
			}
			if (slot == ~0ull) continue;
		}
		const ulonglong2 v01 = __ldg((const ulonglong2 *) (T.vals + slot * 4));
		const unsigned long long c = v01.x;
		if (c == 0) continue;
		const long long hi = (long long) __ldg(T.vals + slot * 4 + 2);
		matches += c;
		/* s + y * c with s = hi * 2^32 + lo, two's complement in 128 bits */
		__int128 s = ((__int128) hi << 32) + (__int128) (unsigned __int128) v01.y;
		s += (__int128) payload[r] * (__int128) (unsigned __int128) c;
		sum += (unsigned __int128) s;
	}
	unsigned long long c0 = (unsigned long long) (unsigned int) sum, c1 = (unsigned long long) (unsigned int) (sum >> 32),
					   c2 = (unsigned long long) (unsigned int) (sum >> 64),
					   c3 = (unsigned long long) (long long) (int) (unsigned int) (sum >> 96);
	matches = join_wsum(matches);
	c0 = join_wsum(c0); c1 = join_wsum(c1); c2 = join_wsum(c2); c3 = join_wsum(c3);
	if ((threadIdx.x & 31) == 0)
	{
		if (matches) atomicAdd(out + 0, matches);
		if (c0) atomicAdd(out + 1, c0);
		if (c1) atomicAdd(out + 2, c1);
		if (c2) atomicAdd(out + 3, c2);
		if (c3) atomicAdd(out + 4, c3);
	}
}

extern "C" int cg_join_count_sum(const int64_t *d_build_keys, const uint8_t *d_build_nulls, const int64_t *d_build_payload,
								 int64_t nbuild, const int64_t *d_probe_keys, const uint8_t *d_probe_nulls,
								 const int64_t *d_probe_payload, int64_t nprobe, int64_t *joined_rows, int64_t *sum_hi,
								 uint64_t *sum_lo)
{
	CgContext *ctx = cg_ctx();
	if (!ctx) return CG_EINVAL;
	if (nbuild < 0 || nprobe < 0 || !joined_rows || !sum_hi || !sum_lo) return cg_set_error(CG_EINVAL, "bad argument");
	if ((nbuild > 0 && (!d_build_keys || !d_build_payload)) || (nprobe > 0 && (!d_probe_keys || !d_probe_payload)))
		return cg_set_error(CG_EINVAL, "NULL column");
	*joined_rows = 0; *sum_hi = 0; *sum_lo = 0;
	if (nbuild == 0 || nprobe == 0) return CG_OK;
	JoinTable T;
	T.cap = 1024;
	while (T.cap < 2ull * (uint64_t) nbuild) T.cap <<= 1;
	T.shift = 64;
	for (unsigned long long c = T.cap; c > 1; c >>= 1) T.shift--;
	const size_t key_bytes = (T.cap + 1) * sizeof(long long), val_bytes = (T.cap + 1) * 4 * sizeof(unsigned long long);
	uint8_t *d_mem = nullptr;
	if (cudaMallocAsync((void **) &d_mem, key_bytes + 64 + val_bytes + 64, ctx->compute) != cudaSuccess)
	{
		cudaGetLastError();
		return cg_set_error(CG_ENOMEM, "cudaMallocAsync of %zu bytes for the join table failed", key_bytes + val_bytes);
	}
	T.keys = (long long *) d_mem;
	T.vals = (unsigned long long *) (d_mem + ((key_bytes + 63) & ~(size_t) 63));
	unsigned long long *d_out = (unsigned long long *) ((uint8_t *) T.vals + val_bytes);   /* 6 words: result + errors */
	T.errors = d_out + 5;
	CG_CUDA(cudaMemsetAsync(T.vals, 0, val_bytes + 64, ctx->compute));
	const unsigned blocks = (unsigned) (ctx->sm_count * 8);
	cg_join_init_kernel<<<blocks, CGJ_THREADS, 0, ctx->compute>>>(T.keys, T.cap + 1);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	cg_join_build_kernel<<<blocks, CGJ_THREADS, 0, ctx->compute>>>(T, (const long long *) d_build_keys, d_build_nulls,
																	  (const long long *) d_build_payload, (long long) nbuild);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	cg_join_probe_kernel<<<blocks, CGJ_THREADS, 0, ctx->compute>>>(T, (const long long *) d_probe_keys, d_probe_nulls,
																	  (const long long *) d_probe_payload, (long long) nprobe, d_out);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	unsigned long long h[6] = {0, 0, 0, 0, 0, 0};
	CG_CUDA(cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, ctx->compute));
	CG_CUDA(cudaStreamSynchronize(ctx->compute));
	CG_CUDA(cudaFreeAsync(d_mem, ctx->compute));
	if (h[5]) return cg_set_error(CG_ETABLEFULL, "join table overflow (%llu rows)", h[5]);
	unsigned __int128 total = (unsigned __int128) h[1] + ((unsigned __int128) h[2] << 32) + ((unsigned __int128) h[3] << 64) +
							  ((unsigned __int128) h[4] << 96);
	*joined_rows = (int64_t) h[0];
	*sum_lo = (uint64_t) total;
	*sum_hi = (int64_t) (uint64_t) (total >> 64);
	return CG_OK;
}

/* ------------------------------------------------------------------------------ *
 *  The same join emitting its rows:
 *      SELECT b.key, b.payload, p.payload FROM build b JOIN probe p USING (key)
 *  for plans whose join output is not aggregated right away (the MERGE task of planner/multi_physical_planner.c:
 *  4304-4328 hands PostgreSQL's HashJoin output to whatever sits above it).  Build: every build row is linked into the
 *  chain of its key's slot (head[slot] <- row, next[row] <- old head: one atomic exchange per row) and the slot counts its
 *  rows.  Probe pass 1 looks the key up and records how many rows it joins; an exclusive prefix sum over the probe rows
 *  gives every probe row its place in the output; pass 2 walks the chain and writes the rows.  Row order is unspecified,
 *  as for any hash join.
 * ------------------------------------------------------------------------------ */
struct JoinRowsTable
{
	long long *keys;              /* [cap + 1] */
	unsigned int *count;          /* [cap + 1] build rows with this key */
	long long *head;              /* [cap + 1] first build row of the chain, -1 = none */
	long long *next;              /* [nbuild] */
	unsigned long long cap;
	int shift;
	unsigned long long *errors;
};

__device__ __forceinline__ unsigned long long join_rows_find(const JoinRowsTable &T, long long key, bool insert)
{
	if (key == CGJ_EMPTY) return T.cap;
	unsigned long long h = join_home(key, T.shift);
	for (unsigned long long probes = 0; probes <= T.cap; probes++)
	{
		long long cur = (long long) __ldcg((unsigned long long *) T.keys + h);
		if (cur == key) return h;
		if (cur == CGJ_EMPTY)
		{
			if (!insert) return ~0ull;
			long long old = (long long) atomicCAS((unsigned long long *) T.keys + h, (unsigned long long) CGJ_EMPTY, (unsigned long long) key);
			if (old == CGJ_EMPTY || old == key) return h;
		}
		h = (h + 1) & (T.cap - 1);
	}
	return ~0ull;
}

__global__ void cg_join_rows_init_kernel(JoinRowsTable T)
{
	for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i <= T.cap; i += (unsigned long long) gridDim.x * blockDim.x)
	{
		T.keys[i] = CGJ_EMPTY; T.count[i] = 0; T.head[i] = -1;
	}
}

__global__ void __launch_bounds__(CGJ_THREADS)
cg_join_rows_build_kernel(JoinRowsTable T, const long long *keys, const uint8_t *nulls, long long n)
{
	for (long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long) gridDim.x * blockDim.x)
	{
		if (nulls && nulls[r]) continue;
		unsigned long long slot = join_rows_find(T, keys[r], true);
		if (slot == ~0ull) { atomicAdd(T.errors, 1ull); continue; }
		atomicAdd(T.count + slot, 1u);
		T.next[r] = (long long) atomicExch((unsigned long long *) T.head + slot, (unsigned long long) r);
	}
}

/* matches[r] = rows probe row r joins; block sums for the scan */
__global__ void __launch_bounds__(CGJ_THREADS)
cg_join_rows_count_kernel(JoinRowsTable T, const long long *keys, const uint8_t *nulls, long long n, unsigned int *matches,
						  unsigned long long *block_sums)
{
	__shared__ unsigned long long s_sum[CGJ_THREADS / 32];
	const long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	unsigned int m = 0;
	if (r < n && !(nulls && nulls[r]))
	{
		unsigned long long slot = join_rows_find(T, keys[r], false);
		if (slot != ~0ull) m = T.count[slot];
	}
	if (r < n) matches[r] = m;
	unsigned long long w = join_wsum((unsigned long long) m);
	if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = w;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		unsigned long long t = 0;
		for (int i = 0; i < CGJ_THREADS / 32; i++) t += s_sum[i];
		block_sums[blockIdx.x] = t;
	}
}

/* exclusive scan of the block sums by one block (nblocks is n / 256: up to ~1 M entries, a few hundred microseconds) */
__global__ void __launch_bounds__(1024)
cg_join_rows_scan_kernel(unsigned long long *block_sums, long long nblocks, unsigned long long *total)
{
	__shared__ unsigned long long s[1024];
	const long long per = (nblocks + 1023) / 1024;
	const long long b0 = (long long) threadIdx.x * per, b1 = b0 + per < nblocks ? b0 + per : nblocks;
	unsigned long long local = 0;
	for (long long b = b0; b < b1; b++) local += block_sums[b];
	s[threadIdx.x] = local;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1)
	{
		unsigned long long v = threadIdx.x >= (unsigned) off ? s[threadIdx.x - off] : 0;
		__syncthreads();
		s[threadIdx.x] += v;
		__syncthreads();
	}
	unsigned long long run = s[threadIdx.x] - local;
	for (long long b = b0; b < b1; b++) { unsigned long long c = block_sums[b]; block_sums[b] = run; run += c; }
	if (threadIdx.x == 1023) *total = s[1023];
}

__global__ void __launch_bounds__(CGJ_THREADS)
cg_join_rows_emit_kernel(JoinRowsTable T, const long long *bpay, const long long *keys, const uint8_t *nulls, const long long *ppay, long long n,
						 const unsigned int *matches, const unsigned long long *block_offsets, unsigned long long capacity,
						 long long *out_key, long long *out_bpay, long long *out_ppay)
{
	__shared__ unsigned long long s_warp[CGJ_THREADS / 32];
	const long long r = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned int m = r < n ? matches[r] : 0u;
	/* place of this row's first output row: block offset + exclusive scan inside the block */
	unsigned long long incl = m;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (unsigned) o) incl += y;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	unsigned long long pos = block_offsets[blockIdx.x] + incl - m;
	for (unsigned w = 0; w < warp; w++) pos += s_warp[w];
	if (m == 0) return;
	const long long key = keys[r];
	const unsigned long long slot = join_rows_find(T, key, false);
	const long long y = ppay[r];
	for (long long b = T.head[slot]; b >= 0; b = T.next[b], pos++)
		if (pos < capacity)
		{
			out_key[pos] = key; out_bpay[pos] = bpay[b]; out_ppay[pos] = y;
		}
}

extern "C" int cg_join_rows(const int64_t *d_build_keys, const uint8_t *d_build_nulls, const int64_t *d_build_payload, int64_t nbuild,
							const int64_t *d_probe_keys, const uint8_t *d_probe_nulls, const int64_t *d_probe_payload, int64_t nprobe,
							int64_t capacity, int64_t *d_out_key, int64_t *d_out_build_payload, int64_t *d_out_probe_payload, int64_t *nrows)
{
	CgContext *ctx = cg_ctx();
	if (!ctx) return CG_EINVAL;
	if (nbuild < 0 || nprobe < 0 || capacity < 0 || !nrows) return cg_set_error(CG_EINVAL, "bad argument");
	if ((nbuild > 0 && (!d_build_keys || !d_build_payload)) || (nprobe > 0 && (!d_probe_keys || !d_probe_payload)))
		return cg_set_error(CG_EINVAL, "NULL column");
	if (capacity > 0 && (!d_out_key || !d_out_build_payload || !d_out_probe_payload)) return cg_set_error(CG_EINVAL, "NULL output column");
	*nrows = 0;
	if (nbuild == 0 || nprobe == 0) return CG_OK;
	JoinRowsTable T;
	T.cap = 1024;
	while (T.cap < 2ull * (uint64_t) nbuild) T.cap <<= 1;
	T.shift = 64;
	for (unsigned long long c = T.cap; c > 1; c >>= 1) T.shift--;
	const long long nblocks = (nprobe + CGJ_THREADS - 1) / CGJ_THREADS;
	auto a16 = [](size_t x) { return (x + 63) & ~(size_t) 63; };
	const size_t kb = a16((T.cap + 1) * 8), cb = a16((T.cap + 1) * 4), hb = a16((T.cap + 1) * 8), nb = a16((size_t) nbuild * 8),
				 mb = a16((size_t) nprobe * 4), bb = a16((size_t) nblocks * 8);
	CgAsyncBuf mem;
	if (mem.alloc(kb + cb + hb + nb + mb + bb + 64, ctx->compute) != cudaSuccess)
	{
		cudaGetLastError();
		return cg_set_error(CG_ENOMEM, "cudaMallocAsync of %zu bytes for the join failed", kb + cb + hb + nb + mb + bb);
	}
	uint8_t *d = mem.as<uint8_t>();
	T.keys = (long long *) d; T.count = (unsigned int *) (d + kb); T.head = (long long *) (d + kb + cb); T.next = (long long *) (d + kb + cb + hb);
	unsigned int *d_matches = (unsigned int *) (d + kb + cb + hb + nb);
	unsigned long long *d_block = (unsigned long long *) (d + kb + cb + hb + nb + mb);
	unsigned long long *d_tot = (unsigned long long *) (d + kb + cb + hb + nb + mb + bb);     /* [0] total rows, [1] errors */
	T.errors = d_tot + 1;
	CG_CUDA(cudaMemsetAsync(d_tot, 0, 16, ctx->compute));
	const unsigned blocks = (unsigned) (ctx->sm_count * 8);
	cg_join_rows_init_kernel<<<blocks, CGJ_THREADS, 0, ctx->compute>>>(T);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	cg_join_rows_build_kernel<<<blocks, CGJ_THREADS, 0, ctx->compute>>>(T, (const long long *) d_build_keys, d_build_nulls, (long long) nbuild);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	cg_join_rows_count_kernel<<<(unsigned) nblocks, CGJ_THREADS, 0, ctx->compute>>>(T, (const long long *) d_probe_keys, d_probe_nulls, (long long) nprobe,
																				  d_matches, d_block);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	cg_join_rows_scan_kernel<<<1, 1024, 0, ctx->compute>>>(d_block, nblocks, d_tot);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	unsigned long long h[2] = {0, 0};
	CG_CUDA(cudaMemcpyAsync(h, d_tot, sizeof h, cudaMemcpyDeviceToHost, ctx->compute));
	CG_CUDA(cudaStreamSynchronize(ctx->compute));
	if (h[1]) return cg_set_error(CG_ETABLEFULL, "join table overflow (%llu rows)", h[1]);
	*nrows = (int64_t) h[0];
	if ((int64_t) h[0] > capacity)
		return capacity == 0 ? CG_OK : cg_set_error(CG_EINVAL, "%llu joined rows exceed the caller's capacity %lld", h[0], (long long) capacity);
	if (h[0] == 0) return CG_OK;
	cg_join_rows_emit_kernel<<<(unsigned) nblocks, CGJ_THREADS, 0, ctx->compute>>>(T, (const long long *) d_build_payload, (const long long *) d_probe_keys,
																				 d_probe_nulls, (const long long *) d_probe_payload, (long long) nprobe,
																				 d_matches, d_block, (unsigned long long) capacity,
																				 (long long *) d_out_key, (long long *) d_out_build_payload,
																				 (long long *) d_out_probe_payload);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	CG_CUDA(cudaStreamSynchronize(ctx->compute));
	return CG_OK;
}
