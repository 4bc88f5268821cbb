/* Host-side fuzz of the lane-per-stream LZ4 decoder the GPU runs (citus_b200/csrc/cg_lz4_lane.cuh): liblz4-compressed
 * random buffers, pristine (must round-trip) and mutated / truncated (must be rejected or decode exactly as liblz4's
 * LZ4_decompress_safe does, never touch memory outside the stream and the slot).  Build with the sanitizers:
 *   g++ -O1 -g -fsanitize=address,undefined -I citus_b200/csrc -o /tmp/lz4_lane_fuzz tools/lz4_lane_fuzz.cpp -ldl && /tmp/lz4_lane_fuzz 1 20000 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <vector>
#include "cg_lz4_lane.cuh"
typedef int (*comp_fn)(const char *, char *, int, int);
typedef int (*bound_fn)(int);
typedef int (*dec_fn)(const char *, char *, int, int);
int main(int argc, char **argv)
{
	void *h = dlopen("liblz4.so.1", RTLD_NOW);
	if (!h) { printf("liblz4.so.1 missing\n"); return 0; }
	comp_fn comp = (comp_fn) dlsym(h, "LZ4_compress_default");
	bound_fn bound = (bound_fn) dlsym(h, "LZ4_compressBound");
	dec_fn dec = (dec_fn) dlsym(h, "LZ4_decompress_safe");
	unsigned seed = argc > 1 ? atoi(argv[1]) : 1;
	int iters = argc > 2 ? atoi(argv[2]) : 20000;
	srand(seed);
	long ok = 0, rejected = 0, stricter = 0;
	unsigned char *win = (unsigned char *) malloc(CGL_WIN * CGL_LANES);
	for (int it = 0; it < iters; it++)
	{
		int n = rand() % (it % 50 == 0 ? 200000 : 5000);
		std::vector<unsigned char> raw(n + 1);
		int mode = rand() % 6;
		int period = 1 + rand() % 3000;
		for (int i = 0; i < n; i++)
			raw[i] = mode == 0 ? rand() : mode == 1 ? (rand() % 4) : mode == 2 ? (unsigned char) (i / 7) : mode == 3 ? ((i % 8) ? 0 : rand() % 100)
					 : mode == 4 ? (i >= period ? raw[i - period] : rand()) : ((rand() % 64) ? 0 : rand());
		int cap = bound(n);
		unsigned char *tmp = (unsigned char *) malloc(cap + 1);
		int clen = comp((const char *) raw.data(), (char *) tmp, n, cap);
		/* exact-size heap blocks so that ASAN sees any over-read / over-write; the slot is padded like an arena slot */
		unsigned char *src = (unsigned char *) malloc(clen);
		memcpy(src, tmp, clen);
		free(tmp);
		int nmut = rand() % 4;          /* 0 = pristine */
		int len = clen;
		for (int m = 0; m < nmut; m++)
		{
			int kind = rand() % 3;
			if (kind == 0) src[rand() % clen] ^= 1u << (rand() % 8);
			else if (kind == 1) src[rand() % clen] = rand();
			else len = 1 + rand() % clen;
		}
		int expect = n;
		if (nmut && rand() % 8 == 0) { expect = n + (rand() % 5) - 2; if (expect < 0) expect = 0; }
		unsigned padded = ((unsigned) expect + 15u) / 16u * 16u + 16u;
		unsigned char *dst = (unsigned char *) aligned_alloc(16, padded);
		unsigned char *ref = (unsigned char *) malloc(expect + 1);
		memset(win, 0xA5, CGL_WIN * CGL_LANES);
		Lz4Lane L;
		L.src = src; L.clen = (uint32_t) len; L.dst = dst; L.rawlen = (uint32_t) expect; { const unsigned lanes = 1u << (rand() % 6); L.wstride = 4u * lanes; L.wb = win + 4u * (rand() % lanes); }
		bool good = cgl_decode(L, padded);
		int r = dec((const char *) src, (char *) ref, len, expect);
		bool ref_good = r == expect;
		if (!good && ref_good && L.zero_offset && nmut) { stricter++; free(ref); free(dst); free(src); continue; }   /* offset 0: see cgl_decode */
		if (good != ref_good) { printf("DISAGREE with liblz4 it=%d nmut=%d (ours %d, liblz4 %d of %d)\n", it, nmut, good, r, expect); return 1; }
		if (good && memcmp(dst, ref, expect)) { printf("BYTES differ it=%d\n", it); return 1; }
		if (nmut == 0 && expect == n && (!good || memcmp(dst, raw.data(), n))) { printf("MISMATCH on pristine input it=%d\n", it); return 1; }
		if (good) { for (unsigned i = expect; i < padded; i++) if (dst[i]) { printf("padding not zero it=%d\n", it); return 1; } ok++; }
		else rejected++;
		free(ref); free(dst); free(src);
	}
	free(win);
	printf("seed %u: %ld decoded, %ld rejected, %ld refused for a zero offset that liblz4 lets through\n", seed, ok, rejected, stricter);
	return 0;
}
