/* Test helper (NOT part of libcitus_gpu.so): the lane-per-stream LZ4 decoder source the GPU kernel runs
 * (citus_b200/csrc/cg_lz4_lane.cuh, every function __host__ __device__) compiled for the host, so that the
 * -m "not gpu" tests can check its format logic, window handling and bounds against liblz4-compressed streams. */
#include <stdlib.h>

#include "cg_lz4_lane.cuh"

/* returns 1 when the stream decoded to exactly rawlen bytes, 0 / -1 when it was refused; `lane` picks the column of the interleaved window */
extern "C" int lz4_lane_decode_host(const unsigned char *src, unsigned len, unsigned char *dst, unsigned rawlen, unsigned padded,
									unsigned lane)
{
	unsigned char *win = (unsigned char *) malloc(CGL_WIN * CGL_LANES);
	for (unsigned i = 0; i < CGL_WIN * CGL_LANES; i++) win[i] = 0xA5;       /* stale bytes must never be read */
	Lz4Lane L;
	L.src = src; L.clen = len; L.dst = dst; L.rawlen = rawlen; L.wb = win + 4u * (lane % CGL_LANES); L.wstride = 4u * CGL_LANES;
	const bool ok = cgl_decode(L, padded);
	free(win);
	return ok ? 1 : (L.zero_offset ? -1 : 0);       /* -1: refused for a match with offset 0 */
}
