/* Test helper (NOT part of libcitus_gpu.so): the Zstandard decoder source the GPU kernel runs on one lane
 * (citus_b200/csrc/cg_zstd.cuh, every function __host__ __device__) compiled for the host, so that the
 * -m "not gpu" tests can check its bit-level format logic against libzstd-compressed streams. */
#include <stdlib.h>

#include "cg_zstd.cuh"

extern "C" long long zstd_decode_host(const unsigned char *src, unsigned len, unsigned char *dst, unsigned cap)
{
	ZstdTables *T = (ZstdTables *) malloc(sizeof(ZstdTables));
	T->huf = (HufEntry *) malloc(sizeof(HufEntry) * (1u << ZSTD_HUF_LOG_MAX));
	unsigned char *lit = (unsigned char *) malloc(ZSTD_BLOCK_MAX + 64);
	long long n = zs_decode_frame(*T, src, len, dst, cap, lit);
	free(lit); free(T->huf); free(T);
	return n;
}
