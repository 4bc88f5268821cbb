/* Host-side fuzz of the Zstandard decoder the GPU runs (citus_b200/csrc/cg_zstd.cuh): libzstd-compressed random
 * buffers, pristine (must round-trip) and mutated / truncated (must be rejected or decode to some size, never touch
 * memory outside its buffers).  Build with the sanitizers:
 *   g++ -O1 -g -fsanitize=address,undefined -I citus_b200/csrc -o /tmp/zstd_fuzz tools/zstd_fuzz.cpp -ldl && /tmp/zstd_fuzz 1 6000
 * Round-1 result: seeds 1-3 x 6000 cases, no sanitizer report, no mismatch. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <vector>
#include "cg_zstd.cuh"
typedef size_t (*comp_fn)(void *, size_t, const void *, size_t, int);
typedef size_t (*bound_fn)(size_t);
int main(int argc, char **argv)
{
	void *h = dlopen("libzstd.so.1", RTLD_NOW);
	comp_fn comp = (comp_fn) dlsym(h, "ZSTD_compress");
	bound_fn bound = (bound_fn) dlsym(h, "ZSTD_compressBound");
	unsigned seed = argc > 1 ? atoi(argv[1]) : 1;
	int iters = argc > 2 ? atoi(argv[2]) : 20000;
	srand(seed);
	long ok = 0, rejected = 0, wrong_size = 0;
	for (int it = 0; it < iters; it++)
	{
		size_t n = 1 + rand() % (it % 50 == 0 ? 300000 : 4000);
		std::vector<unsigned char> raw(n);
		int mode = rand() % 4;
		for (size_t i = 0; i < n; i++)
			raw[i] = mode == 0 ? rand() : mode == 1 ? (rand() % 4) : mode == 2 ? (unsigned char) (i / 7) : ((i % 8) ? 0 : rand() % 100);
		size_t cap = bound(n);
		/* exact-size heap blocks so that ASAN sees any over-read / over-write */
		unsigned char *tmp = (unsigned char *) malloc(cap);
		size_t clen = comp(tmp, cap, raw.data(), n, 1 + rand() % 19);
		unsigned char *src = (unsigned char *) malloc(clen + 16);      /* arena slots are padded by >= 16 bytes */
		memcpy(src, tmp, clen); memset(src + clen, 0, 16);
		free(tmp);
		/* mutate */
		int nmut = rand() % 4;          /* 0 = pristine */
		size_t len = clen;
		for (int m = 0; m < nmut; m++)
		{
			int kind = rand() % 3;
			if (kind == 0) src[rand() % clen] ^= 1u << (rand() % 8);
			else if (kind == 1) src[rand() % clen] = rand();
			else len = 1 + rand() % clen;    /* truncate */
		}
		unsigned char *dst = (unsigned char *) malloc(n);
		ZstdTables *T = (ZstdTables *) malloc(sizeof(ZstdTables));
		T->huf = (HufEntry *) malloc(sizeof(HufEntry) * 2048);
		unsigned char *lit = (unsigned char *) malloc(ZSTD_BLOCK_MAX + 64);
		long long r = zs_decode_frame(*T, src, (uint32_t) len, dst, (uint32_t) n, lit);
		if (nmut == 0)
		{
			if (r != (long long) n || memcmp(dst, raw.data(), n)) { printf("MISMATCH on pristine input it=%d\n", it); return 1; }
			ok++;
		}
		else if (r < 0) rejected++;
		else if (r != (long long) n) wrong_size++;
		else ok++;
		free(lit); free(T->huf); free(T); free(dst); free(src);
	}
	printf("seed %u: %ld decoded, %ld rejected, %ld other size\n", seed, ok, rejected, wrong_size);
	return 0;
}
