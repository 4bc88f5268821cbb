/*
 * cg_lz4_lane.cuh -- an LZ4 block decoder in which ONE GPU lane decodes one value stream from start to end
 * (and, for the CPU-side format tests, the host runs the same source: every function is __host__ __device__).
 *
 * Replaces the COMPRESSION_LZ4 arm of DecompressBuffer, backend/columnar/columnar_compression.c:186-208
 * (LZ4_decompress_safe(buffer, out, len, decompressedSize) must return exactly decompressedSize); the block
 * format is the published one, restated at the top of cg_decompress.cu.
 *
 * Why a lane per stream: the eight-lanes-per-stream kernel in cg_decompress.cu is bound by instruction issue --
 * about 150 warp instructions per sequence, and a sequence of a columnar value stream is ~8 bytes (a few literal
 * bytes of a value and a match for the bytes it shares with its neighbours), so most of those instructions
 * coordinate lanes that have nothing to copy.  Here a warp instruction advances 32 streams, nothing is
 * coordinated, and an overlapping match is an ordinary sequential copy.
 *
 * The last CGL_WIN decoded bytes of a stream live in a window in shared memory; the windows of the lanes of a warp
 * that decode are interleaved word by word (word w of lane l at [w * lanes + l]): lanes that work at the same position
 * of their streams hit different banks.  How many lanes of a warp decode is the launch's choice (cg_decompress.cu):
 * lanes of one warp that are in different phases of a sequence execute one after the other, so when a launch has
 * fewer streams than the GPU has warp slots, one stream per warp is fastest.  A match that reaches
 * back further than the window reads the arena (behind a flush).  The window leaves for the arena in aligned
 * 16-byte stores.
 */
#ifndef CG_LZ4_LANE_CUH
#define CG_LZ4_LANE_CUH

#include <stdint.h>
#include <string.h>

#ifndef CG_HD
#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD inline
#endif
#endif

#define CGL_WIN 2048u           /* bytes of window per lane */
#define CGL_PIECE 256u          /* a copy advances in pieces of at most this many bytes; < CGL_WIN - 16 */
#define CGL_LANES 32u
#define CGL_AHEAD 48u           /* the 8-byte copies may leave up to this many bytes of garbage beyond `op`: they land
                                 * where later bytes go, and in the oldest bytes of the window, which therefore serves
                                 * matches up to CGL_WIN - CGL_AHEAD back */

struct Lz4Lane
{
	const uint8_t *src;         /* the compressed stream */
	uint32_t clen;
	uint8_t *dst;               /* 16-byte aligned slot of `padded` bytes */
	uint32_t rawlen;
	uint8_t *wb;                /* window: byte b of the lane's word w is wb[w * wstride + b] */
	uint32_t wstride;           /* 4 * (number of lanes whose windows are interleaved) */
	uint32_t ip;                /* bytes of the stream consumed */
	uint32_t op;                /* bytes decoded */
	uint32_t flushed;           /* bytes already in dst (multiple of 16) */
	uint32_t zero_offset;       /* out: the stream was refused because a match has offset 0 (see cgl_decode) */
};

#define cgl_at(pos) ((((pos) & (CGL_WIN - 1u)) >> 2) * L.wstride + ((pos) & 3u))
#define cgl_word_at(pos) ((((pos) & (CGL_WIN - 1u)) >> 2) * L.wstride)

/* window -> dst for [flushed, upto rounded down to 16) */
CG_HD void cgl_flush(Lz4Lane &L, uint32_t upto)
{
	upto &= ~15u;
	for (uint32_t pos = L.flushed; pos < upto; pos += 16u)
	{
		uint32_t w[4];
		for (int k = 0; k < 4; k++) w[k] = *(const uint32_t *) (L.wb + cgl_at(pos + 4u * k));
#ifdef __CUDA_ARCH__
		*(uint4 *) (L.dst + pos) = make_uint4(w[0], w[1], w[2], w[3]);
#else
		memcpy(L.dst + pos, w, 16);
#endif
	}
	if (upto > L.flushed) L.flushed = upto;
}

/* before m <= CGL_PIECE more bytes are written: nothing unflushed may be overwritten */
CG_HD void cgl_room(Lz4Lane &L, uint32_t m)
{
	if (L.op + m - L.flushed > CGL_WIN) cgl_flush(L, L.op);
}

/* ---- eight bytes at a time ---- */
CG_HD uint32_t cgl_funnel(uint32_t lo, uint32_t hi, uint32_t bits)      /* low word of (hi:lo) >> bits, bits in 0, 8, 16, 24 */
{
#ifdef __CUDA_ARCH__
	return __funnelshift_r(lo, hi, bits);
#else
	return bits ? (lo >> bits) | (hi << (32u - bits)) : lo;
#endif
}

/* the 8 bytes at p; [p rounded down to 8, + 16) must be readable */
CG_HD uint64_t cgl_load8(const uint8_t *p)
{
#ifdef __CUDA_ARCH__
	const unsigned long long a = (unsigned long long) p;
	const uint64_t *q = (const uint64_t *) (a & ~7ull);
	const uint32_t sh = (uint32_t) (a & 7ull) * 8u;
	const uint64_t lo = q[0], hi = q[1];
	return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
#else
	uint64_t v;
	memcpy(&v, p, 8);
	return v;
#endif
}

/* the same from the output slot: bytes this lane stored earlier (as 16-byte vectors), hence volatile -- the loads
 * must neither be moved above those stores nor be served from a stale L1 line */
CG_HD uint64_t cgl_load8_slot(const uint8_t *p)
{
#ifdef __CUDA_ARCH__
	const unsigned long long a = (unsigned long long) p;
	const volatile uint64_t *q = (const volatile uint64_t *) (a & ~7ull);
	const uint32_t sh = (uint32_t) (a & 7ull) * 8u;
	const uint64_t lo = q[0], hi = q[1];
	return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
#else
	uint64_t v;
	memcpy(&v, p, 8);
	return v;
#endif
}


/* the 8 window bytes at position pos (any alignment) */
CG_HD uint64_t cgl_win_read8(const Lz4Lane &L, uint32_t pos)
{
	const uint32_t w0 = *(const uint32_t *) (L.wb + cgl_word_at(pos)), w1 = *(const uint32_t *) (L.wb + cgl_word_at(pos + 4u));
	const uint32_t sh = (pos & 3u) * 8u;
	if (sh == 0) return (uint64_t) w0 | ((uint64_t) w1 << 32);
	const uint32_t w2 = *(const uint32_t *) (L.wb + cgl_word_at(pos + 8u));
	return (uint64_t) cgl_funnel(w0, w1, sh) | ((uint64_t) cgl_funnel(w1, w2, sh) << 32);
}

/* v's 8 bytes to window positions [pos, pos + 8); the bytes below pos in pos's word are kept, up to 3 bytes beyond
 * pos + 8 are overwritten with zeros */
CG_HD void cgl_win_write8(Lz4Lane &L, uint32_t pos, uint64_t v)
{
	const uint32_t sh = (pos & 3u) * 8u, lo = (uint32_t) v, hi = (uint32_t) (v >> 32);
	if (sh == 0)
	{
		*(uint32_t *) (L.wb + cgl_word_at(pos)) = lo;
		*(uint32_t *) (L.wb + cgl_word_at(pos + 4u)) = hi;
		return;
	}
	uint32_t *p0 = (uint32_t *) (L.wb + cgl_word_at(pos));
	*p0 = (*p0 & ((1u << sh) - 1u)) | (lo << sh);
	*(uint32_t *) (L.wb + cgl_word_at(pos + 4u)) = cgl_funnel(lo, hi, 32u - sh);
	*(uint32_t *) (L.wb + cgl_word_at(pos + 8u)) = hi >> (32u - sh);
}

CG_HD void cgl_room8(Lz4Lane &L)
{
	if (L.op + CGL_AHEAD - L.flushed > CGL_WIN) cgl_flush(L, L.op);
}

CG_HD void cgl_literals(Lz4Lane &L, uint32_t ip, uint32_t n)
{
	/* 8 bytes at a time while the loads stay inside the stream, then byte by byte */
	while (n >= 8u && (ip & ~7u) + 16u <= L.clen)
	{
		cgl_room8(L);
		cgl_win_write8(L, L.op, cgl_load8(L.src + ip));
		L.op += 8u; ip += 8u; n -= 8u;
	}
	while (n)
	{
		const uint32_t m = n < CGL_PIECE ? n : CGL_PIECE;
		cgl_room(L, m);
		for (uint32_t i = 0; i < m; i++) L.wb[cgl_at(L.op + i)] = L.src[ip + i];
		L.op += m; ip += m; n -= m;
	}
}

/* a match whose source is in the window (off <= CGL_WIN - CGL_AHEAD), 8 bytes per step.  A step never reads a byte it
 * writes when the distance is >= 8.  A shorter distance means the output is periodic with period `off`: the first 8
 * bytes are the period replicated, and from then on the same bytes are found D = off * ceil(8 / off) >= 8 back (the
 * periodic region [op - off, op + 8) is off + 8 > D bytes long).  The last step may write beyond the match. */
CG_HD void cgl_match_win(Lz4Lane &L, uint32_t off, uint32_t n)
{
	uint32_t dist = off;
	if (off < 8u)
	{
		cgl_room8(L);
		uint64_t v = cgl_win_read8(L, L.op - off) & ((1ull << (8u * off)) - 1ull);
		uint32_t s = 8u * off;
		v |= v << s; s <<= 1;
		if (s < 64u) { v |= v << s; s <<= 1; if (s < 64u) v |= v << s; }
		cgl_win_write8(L, L.op, v);
		const uint32_t m = n < 8u ? n : 8u;
		L.op += m; n -= m;
		dist = off * ((7u + off) / off);
	}
	while (n)
	{
		const uint32_t m = n < 8u ? n : 8u;
		cgl_room8(L);
		cgl_win_write8(L, L.op, cgl_win_read8(L, L.op - dist));
		L.op += m; n -= m;
	}
}

/* n bytes from `off` bytes back (1 <= off <= op), LZ77's overlap rule: byte i is out[op - off + (i mod off)] */
CG_HD void cgl_match(Lz4Lane &L, uint32_t off, uint32_t n)
{
	if (off <= CGL_WIN - CGL_AHEAD)
	{
		/* the source of byte w is w - off: its window slot is overwritten by byte w - off + CGL_WIN > w (or by what an
		 * 8-byte copy left up to CGL_AHEAD bytes beyond its end, hence the margin) */
		cgl_match_win(L, off, n);
		return;
	}
	/* far match: its source left the window, but it is in the slot once the window is flushed: 8 bytes per step from
	 * dst (this lane's own earlier stores; one L2 round trip per step).  A step reads [op - off, op - off + 8) with
	 * off > CGL_WIN - CGL_AHEAD, which lies below the flushed mark as long as op - flushed stays small. */
	while (n)
	{
		const uint32_t m = n < 8u ? n : 8u;
		if (L.op - L.flushed >= 16u) cgl_flush(L, L.op);
		cgl_win_write8(L, L.op, cgl_load8_slot(L.dst + (L.op - off)));
		L.op += m; n -= m;
	}
}

/*
 * The decoder as a state machine, one sequence per step, so that the lanes of a warp can take their steps together
 * (cg_lz4_lane_kernel re-converges them after every step; the host tests just loop).
 *   cgl_begin   CGL_MORE, or CGL_BAD for a stream that cannot be valid
 *   cgl_step    one sequence: CGL_MORE, CGL_DONE after the last one, CGL_BAD
 *   cgl_finish  after CGL_DONE: true when exactly rawlen bytes came out; the window's rest and the padding are written
 * Agrees with LZ4_decompress_safe on every stream (tools/lz4_lane_fuzz.cpp) except one kind of damage: a match with
 * offset 0, which liblz4 1.9 "decodes" to whatever the output buffer held before the call and which is refused here.
 */
#define CGL_MORE 0
#define CGL_DONE 1
#define CGL_BAD (-1)

CG_HD int cgl_begin(Lz4Lane &L)
{
	L.ip = 0; L.op = 0; L.flushed = 0; L.zero_offset = 0;
	if (L.clen == 0) return CGL_BAD;
	if (L.rawlen == 0 && !(L.clen == 1 && L.src[0] == 0)) return CGL_BAD;      /* liblz4's rule for an empty output */
	return CGL_MORE;
}

CG_HD int cgl_step(Lz4Lane &L)
{
	uint32_t ip = L.ip;
	const uint32_t clen = L.clen, rawlen = L.rawlen;
	/*
	 * The common sequence of a columnar value stream -- up to 5 literal bytes and a match of 4..18 bytes -- in a
	 * handful of word operations: ONE 8-byte load holds the token, the literals and the offset; literals and match move
	 * 8 bytes at a time (what a copy writes beyond its end is overwritten by the next one).  Everything else, and the
	 * last 16 bytes of the stream, take the general path below.
	 */
	if ((ip & ~7u) + 16u <= clen)
	{
		const uint64_t v = cgl_load8(L.src + ip);
		const uint32_t token = (uint32_t) v & 0xffu, lit = token >> 4, mlc = token & 15u;
		const uint32_t off = (uint32_t) (v >> (8u * (1u + (lit <= 5u ? lit : 0u)))) & 0xffffu, ml = mlc + 4u;
		const uint32_t mp = L.op + lit;                         /* where the match starts */
		if (lit <= 5u && mlc != 15u && off != 0 && off <= mp && lit + ml <= rawlen - L.op)
		{
			cgl_room8(L);
			cgl_win_write8(L, L.op, v >> 8);
			L.op = mp;
			cgl_match(L, off, ml);
			L.ip = ip + 3u + lit;
			return CGL_MORE;
		}
	}
	if (ip >= clen) return CGL_BAD;
	const uint32_t token = L.src[ip++];
	uint32_t lit = token >> 4;
	if (lit == 15u)
	{
		uint32_t b;
		do
		{
			if (ip >= clen) return CGL_BAD;
			b = L.src[ip++];
			lit += b;
			if (lit > rawlen) return CGL_BAD;
		} while (b == 255u);
	}
	if (lit > clen - ip || lit > rawlen - L.op) return CGL_BAD;
	cgl_literals(L, ip, lit);
	ip += lit;
	if (ip == clen) { L.ip = ip; return CGL_DONE; }     /* the last sequence stops after its literals */
	if (clen - ip < 2u) return CGL_BAD;
	const uint32_t off = (uint32_t) L.src[ip] | ((uint32_t) L.src[ip + 1] << 8);
	ip += 2;
	uint32_t ml = token & 15u;
	if (ml == 15u)
	{
		uint32_t b;
		do
		{
			if (ip >= clen) return CGL_BAD;
			b = L.src[ip++];
			ml += b;
			if (ml > rawlen) return CGL_BAD;
		} while (b == 255u);
	}
	ml += 4u;
	if (off == 0) L.zero_offset = 1;
	if (off == 0 || off > L.op || ml > rawlen - L.op) return CGL_BAD;
	cgl_match(L, off, ml);
	L.ip = ip;
	return CGL_MORE;
}

CG_HD bool cgl_finish(Lz4Lane &L, uint32_t padded)
{
	if (L.op != L.rawlen) return false;
	cgl_flush(L, L.op);
	for (uint32_t i = L.flushed; i < L.rawlen; i++) L.dst[i] = L.wb[cgl_at(i)];
	for (uint32_t i = L.rawlen; i < padded; i++) L.dst[i] = 0;
	return true;
}

/* the whole stream on one lane.  true: the slot holds exactly rawlen decoded bytes followed by zeros up to `padded`;
 * false: malformed stream (nothing outside the slot was written; the caller zero-fills it) */
CG_HD bool cgl_decode(Lz4Lane &L, uint32_t padded)
{
	int st = cgl_begin(L);
	while (st == CGL_MORE) st = cgl_step(L);
	return st == CGL_DONE && cgl_finish(L, padded);
}

#endif
