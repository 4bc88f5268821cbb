/*
 * cg_zstd.cuh -- a Zstandard frame decoder that runs as ordinary sequential code on one GPU
 * lane (and, for the CPU-side format tests, on the host: every function is __host__ __device__).
 *
 * Replaces the COMPRESSION_ZSTD arm of DecompressBuffer, backend/columnar/columnar_compression.c:
 * 210-238: ZSTD_decompress(out, decompressedSize, buffer, len) must succeed and return exactly
 * decompressedSize.  libzstd is not part of the reference tree; the frame format is the published
 * one (RFC 8878): frame header, raw / RLE / compressed blocks, literals section (raw, RLE,
 * Huffman with 1 or 4 streams, tree described by direct or FSE-compressed weights, "treeless"
 * reuse), sequences section (predefined / RLE / FSE / repeat tables for literal lengths, offsets
 * and match lengths, backward bitstream, repeat-offset history).  Not supported (the columnar
 * writer never produces them, columnar_compression.c:103-123 calls plain ZSTD_compress): dictionaries
 * and skippable frames; a content checksum, if present, is skipped, not verified.
 *
 * Work memory: the FSE tables (ZstdTables, ~7 KB, shared memory on the GPU), the Huffman table
 * (4 KB) and one literals buffer of up to 128 KB per stream being decoded (global memory).
 */
#ifndef CG_ZSTD_CUH
#define CG_ZSTD_CUH

#include <stdint.h>

#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#define CG_HDN __host__ __device__
#else
#define CG_HD inline
#define CG_HDN
#endif

#define ZSTD_BLOCK_MAX (128u * 1024u)
#define ZSTD_LL_LOG_MAX 9
#define ZSTD_OF_LOG_MAX 8
#define ZSTD_ML_LOG_MAX 9
#define ZSTD_HUF_LOG_MAX 11
#define ZSTD_ERR (-1)

struct FseEntry { uint8_t symbol; uint8_t nbits; uint16_t base; };     /* next state = base + read(nbits) */
struct FseTable { int log; FseEntry e[1 << ZSTD_LL_LOG_MAX]; };
struct FseTableOf { int log; FseEntry e[1 << ZSTD_OF_LOG_MAX]; };
struct HufEntry { uint8_t symbol; uint8_t nbits; };

struct ZstdTables
{
	FseTable ll, ml;
	FseTableOf of;
	HufEntry *huf;              /* [1 << ZSTD_HUF_LOG_MAX], caller-provided (global memory on the GPU: keeps the
								 * shared-memory footprint of a decoder at ~7 KB, 31 decoders per SM instead of 20) */
	int huf_log;                /* 0 = no Huffman table yet */
	int have_ll, have_of, have_ml;
	uint32_t rep[3];
	int16_t norm[256];          /* scratch: normalized counts / Huffman weights */
	uint16_t next[256];         /* scratch: per-symbol state counters / rank starts */
};

CG_HD int zs_highbit(uint32_t v)
{
#ifdef __CUDA_ARCH__
	return 31 - __clz((int) v);      /* -1 for 0 */
#else
	int r = -1; while (v) { v >>= 1; r++; } return r;
#endif
}

/* ---- forward bit reader (FSE table descriptions): LSB first ---- */
struct ZsFwd { const uint8_t *p; uint32_t len; uint32_t bit; };
CG_HD uint32_t zs_fwd_read(ZsFwd &r, int n)
{
	uint32_t v = 0;
	for (int i = 0; i < n; i++)
	{
		uint32_t b = r.bit + (uint32_t) i, byte = b >> 3;
		uint32_t x = byte < r.len ? (uint32_t) (r.p[byte] >> (b & 7u)) & 1u : 0u;
		v |= x << i;
	}
	r.bit += (uint32_t) n;
	return v;
}

/* ---- backward bit reader: the stream is a little-endian bit array whose highest set bit (in the
 * last byte) is the end mark; read(n) returns the n bits just below the cursor, top bit first ---- */
struct ZsBack
{
	const uint8_t *p;
	uint32_t len;
	int32_t pos;        /* number of unread bits; may go negative (zeros).  Streams are < 2^24 bytes. */
	uint32_t hi, lo;    /* the next unread bits, left aligned: bit 31 of hi is the bit just below the cursor */
	int32_t avail;      /* how many of them are loaded (bits below the start of the stream count: they are zeros) */
};
/* (hi:lo << sh) >> 32, 0 <= sh <= 31 */
CG_HD uint32_t zs_shl_hi(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
	return __funnelshift_l(lo, hi, sh);
#else
	return (uint32_t) (((((uint64_t) hi << 32) | lo) << sh) >> 32);
#endif
}
/* window := bits [pos - 64, pos), left aligned, zeros below bit 0 */
CG_HD void zs_back_refill(ZsBack &r)
{
	const int32_t top = r.pos;
	if (top <= 0) { r.hi = r.lo = 0; r.avail = 64; return; }
	const uint32_t byte_hi = (uint32_t) (top + 7) >> 3;         /* bytes [byte_hi - 8, byte_hi) hold the window */
	const uint32_t slack = byte_hi * 8u - (uint32_t) top;        /* 0..7 bits above the cursor */
	uint32_t w0 = 0, w1 = 0;
	if (byte_hi >= 8)
	{
		const uint8_t *q = r.p + byte_hi - 8;
#ifdef __CUDA_ARCH__
		/* 8 bytes at any alignment from three aligned words (arena slots are padded, the over-read stays inside) */
		const uintptr_t a = (uintptr_t) q;
		const uint32_t *aw = (const uint32_t *) (a & ~(uintptr_t) 3);
		const uint32_t sh = (uint32_t) (a & 3u) * 8u;
		const uint32_t x0 = aw[0], x1 = aw[1], x2 = aw[2];
		w0 = __funnelshift_r(x0, x1, sh);
		w1 = __funnelshift_r(x1, x2, sh);
#else
		for (uint32_t i = 0; i < 4; i++) { w0 |= (uint32_t) q[i] << (8 * i); w1 |= (uint32_t) q[4 + i] << (8 * i); }
#endif
	}
	else
	{
		/* near the start of the stream: byte (byte_hi - 1) is the top byte, what is below byte 0 is zero */
		for (uint32_t i = 0; i < byte_hi; i++)
		{
			const uint32_t at = 8u - byte_hi + i;                /* position of stream byte i inside the 8-byte window */
			if (at < 4) w0 |= (uint32_t) r.p[i] << (8 * at); else w1 |= (uint32_t) r.p[i] << (8 * (at - 4));
		}
	}
	r.hi = zs_shl_hi(w0, w1, slack);
	r.lo = w0 << slack;
	r.avail = 64 - (int32_t) slack;
}
CG_HD bool zs_back_init(ZsBack &r, const uint8_t *p, uint32_t len)
{
	if (len == 0 || len > (1u << 24) || p[len - 1] == 0) return false;
	r.p = p;
	r.len = len;
	r.pos = (int32_t) (len - 1) * 8 + zs_highbit(p[len - 1]);
	zs_back_refill(r);
	return true;
}
/* the n bits below the cursor, top bit first (1 <= n <= 31; bits below the start of the stream are zeros) */
CG_HD uint32_t zs_back_peek(ZsBack &r, int n)
{
	if (r.avail < n) zs_back_refill(r);
	return r.hi >> (32 - n);
}
CG_HD void zs_back_skip(ZsBack &r, int n)      /* 0 <= n <= 31, n <= avail */
{
	r.hi = zs_shl_hi(r.lo, r.hi, (uint32_t) n);
	r.lo <<= n;
	r.avail -= n;
	r.pos -= n;
}
CG_HD uint32_t zs_back_read(ZsBack &r, int n)
{
	if (n == 0) return 0;
	const uint32_t v = zs_back_peek(r, n);
	zs_back_skip(r, n);
	return v;
}

/* ---- FSE ---- */
/* reads a normalized-count description; returns bytes consumed or ZSTD_ERR.  norm[] gets the counts (-1 = "less than 1") */
CG_HDN inline int zs_read_ncount(const uint8_t *src, uint32_t len, int max_log, int max_symbol, int16_t *norm, int *log_out, int *nsym_out)
{
	ZsFwd r{src, len, 0};
	int log = 5 + (int) zs_fwd_read(r, 4);
	if (log > max_log) return ZSTD_ERR;
	int remaining = 1 << log, symb = 0;
	while (remaining > 0 && symb <= max_symbol)
	{
		int bits = zs_highbit((uint32_t) (remaining + 1)) + 1;
		uint32_t val = zs_fwd_read(r, bits);
		uint32_t lower_mask = (1u << (bits - 1)) - 1u;
		uint32_t threshold = (1u << bits) - 1u - (uint32_t) (remaining + 1);
		if ((val & lower_mask) < threshold) { r.bit--; val &= lower_mask; }
		else if (val > lower_mask) val -= threshold;
		int proba = (int) val - 1;
		remaining -= proba < 0 ? -proba : proba;
		norm[symb++] = (int16_t) proba;
		if (proba == 0)
		{
			uint32_t repeat = zs_fwd_read(r, 2);
			for (;;)
			{
				for (uint32_t i = 0; i < repeat && symb <= max_symbol; i++) norm[symb++] = 0;
				if (repeat == 3) repeat = zs_fwd_read(r, 2); else break;
			}
		}
		if ((r.bit >> 3) > len) return ZSTD_ERR;
	}
	if (remaining != 0 || symb > max_symbol + 1) return ZSTD_ERR;
	*log_out = log;
	*nsym_out = symb;
	uint32_t used = (r.bit + 7) >> 3;
	if (used > len) return ZSTD_ERR;
	return (int) used;
}

/* decoding table from normalized counts (spread with step (size>>1)+(size>>3)+3, "less than 1" symbols at the top) */
CG_HDN inline int zs_build_fse(FseEntry *e, int log, const int16_t *norm, int nsym, uint16_t *next)
{
	const uint32_t size = 1u << log;
	uint32_t high = size;
	for (int s = 0; s < nsym; s++)
		if (norm[s] == -1) { e[--high].symbol = (uint8_t) s; next[s] = 1; }
	const uint32_t step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
	uint32_t pos = 0;
	for (int s = 0; s < nsym; s++)
	{
		if (norm[s] <= 0) continue;
		next[s] = (uint16_t) norm[s];
		for (int i = 0; i < norm[s]; i++)
		{
			e[pos].symbol = (uint8_t) s;
			do { pos = (pos + step) & mask; } while (pos >= high);
		}
	}
	if (pos != 0) return ZSTD_ERR;
	for (uint32_t i = 0; i < size; i++)
	{
		uint32_t ns = next[e[i].symbol]++;
		int nb = log - zs_highbit(ns);
		e[i].nbits = (uint8_t) nb;
		e[i].base = (uint16_t) ((ns << nb) - size);
	}
	return 0;
}

CG_HDN inline void zs_build_rle(FseEntry *e, int *log, uint8_t symbol)
{
	*log = 0;
	e[0].symbol = symbol; e[0].nbits = 0; e[0].base = 0;
}

/* ---- Huffman ---- */
/* weights[0..n) given (n < 256): derive the last weight, build the single-symbol decoding table */
CG_HDN inline int zs_build_huf(ZstdTables &T, int n)
{
	int16_t *w = T.norm;
	uint32_t total = 0;
	for (int i = 0; i < n; i++)
	{
		if (w[i] > ZSTD_HUF_LOG_MAX) return ZSTD_ERR;
		if (w[i] > 0) total += 1u << (w[i] - 1);
	}
	if (total == 0) return ZSTD_ERR;
	int log = zs_highbit(total) + 1;
	if (log > ZSTD_HUF_LOG_MAX) return ZSTD_ERR;
	uint32_t rest = (1u << log) - total;
	if (rest == 0 || (rest & (rest - 1)) != 0) return ZSTD_ERR;
	w[n] = (int16_t) (zs_highbit(rest) + 1);
	n++;
	/* entries in order of increasing weight, symbols in natural order inside one weight */
	uint32_t pos = 0;
	for (int weight = 1; weight <= log; weight++)
		for (int s = 0; s < n; s++)
			if (w[s] == weight)
			{
				uint32_t span = 1u << (weight - 1);
				for (uint32_t i = 0; i < span; i++) { T.huf[pos + i].symbol = (uint8_t) s; T.huf[pos + i].nbits = (uint8_t) (log + 1 - weight); }
				pos += span;
			}
	if (pos != (1u << log)) return ZSTD_ERR;
	T.huf_log = log;
	return 0;
}

/* Huffman tree description; returns bytes consumed or ZSTD_ERR */
CG_HDN inline int zs_read_huf_tree(ZstdTables &T, const uint8_t *src, uint32_t len)
{
	if (len < 1) return ZSTD_ERR;
	uint32_t hb = src[0];
	int n = 0;
	uint32_t used;
	if (hb >= 128)
	{
		n = (int) hb - 127;
		uint32_t bytes = (uint32_t) (n + 1) / 2;
		if (1 + bytes > len) return ZSTD_ERR;
		for (int i = 0; i < n; i++)
		{
			uint8_t b = src[1 + i / 2];
			T.norm[i] = (int16_t) ((i & 1) ? (b & 15) : (b >> 4));
		}
		used = 1 + bytes;
	}
	else
	{
		/* FSE-compressed weights: table (accuracy <= 6) + two interleaved states over a backward bitstream */
		if (hb == 0 || 1 + hb > len) return ZSTD_ERR;
		const uint8_t *p = src + 1;
		int16_t norm[16];
		uint16_t next[16];
		int log = 0, nsym = 0;
		int hdr = zs_read_ncount(p, hb, 6, 12, norm, &log, &nsym);
		if (hdr < 0) return ZSTD_ERR;
		FseEntry e[64];
		if (zs_build_fse(e, log, norm, nsym, next) < 0) return ZSTD_ERR;
		ZsBack r;
		if ((uint32_t) hdr >= hb || !zs_back_init(r, p + hdr, hb - (uint32_t) hdr)) return ZSTD_ERR;
		uint32_t s1 = zs_back_read(r, log), s2 = zs_back_read(r, log);
		if (r.pos < 0) return ZSTD_ERR;
		for (;;)
		{
			if (n >= 254) return ZSTD_ERR;
			T.norm[n++] = e[s1].symbol;
			s1 = e[s1].base + zs_back_read(r, e[s1].nbits);
			if (r.pos < 0) { T.norm[n++] = e[s2].symbol; break; }
			if (n >= 254) return ZSTD_ERR;
			T.norm[n++] = e[s2].symbol;
			s2 = e[s2].base + zs_back_read(r, e[s2].nbits);
			if (r.pos < 0) { T.norm[n++] = e[s1].symbol; break; }
		}
		used = 1 + hb;
	}
	if (n < 1 || n > 255) return ZSTD_ERR;
	if (zs_build_huf(T, n) < 0) return ZSTD_ERR;
	return (int) used;
}

/* one Huffman stream: exactly `count` symbols, every bit consumed */
CG_HDN inline int zs_huf_stream(const ZstdTables &T, const uint8_t *src, uint32_t len, uint8_t *out, uint32_t count)
{
	ZsBack r;
	if (!zs_back_init(r, src, len)) return ZSTD_ERR;
	const int log = T.huf_log;
	uint32_t i = 0;
	/* bytes up to a 4-byte boundary of `out`, then four symbols per store */
	while (i < count && (((uintptr_t) (out + i)) & 3u) != 0)
	{
		const HufEntry h = T.huf[zs_back_peek(r, log)];
		out[i++] = h.symbol;
		zs_back_skip(r, h.nbits);
	}
	for (; i + 4 <= count; i += 4)
	{
		uint32_t w = 0;
		for (int k = 0; k < 4; k++)
		{
			const HufEntry h = T.huf[zs_back_peek(r, log)];
			w |= (uint32_t) h.symbol << (8 * k);
			zs_back_skip(r, h.nbits);
		}
		*(uint32_t *) (out + i) = w;
		if (r.pos < 0) return ZSTD_ERR;
	}
	for (; i < count; i++)
	{
		const HufEntry h = T.huf[zs_back_peek(r, log)];
		out[i] = h.symbol;
		zs_back_skip(r, h.nbits);
	}
	return r.pos == 0 ? 0 : ZSTD_ERR;
}

/* a literals section after its headers (and Huffman tree) have been read: what is left to do is copy,
 * fill, or decode 1 / 4 independent Huffman streams -- work that can be split over lanes */
struct ZsLitPlan
{
	int kind;                   /* 0 raw copy, 1 RLE fill, 2 Huffman streams */
	uint32_t regen;             /* literals produced */
	uint32_t consumed;          /* bytes of the block taken by the whole literals section */
	const uint8_t *raw;         /* kind 0: the bytes; kind 1: the byte */
	int nstreams;
	const uint8_t *sptr[4];
	uint32_t slen[4], scount[4], sout[4];
};

/* literals section headers (+ Huffman tree description into T); returns 0 or ZSTD_ERR */
CG_HDN inline int zs_literals_prepare(ZstdTables &T, const uint8_t *src, uint32_t len, ZsLitPlan *L)
{
	if (len < 1) return ZSTD_ERR;
	const uint32_t type = src[0] & 3u, fmt = (src[0] >> 2) & 3u;
	uint32_t regen, comp = 0, hdr;
	L->nstreams = 0;
	if (type < 2)
	{
		if ((fmt & 1u) == 0) { regen = src[0] >> 3; hdr = 1; }
		else if (fmt == 1) { if (len < 2) return ZSTD_ERR; regen = (src[0] >> 4) + ((uint32_t) src[1] << 4); hdr = 2; }
		else { if (len < 3) return ZSTD_ERR; regen = (src[0] >> 4) + ((uint32_t) src[1] << 4) + ((uint32_t) src[2] << 12); hdr = 3; }
		if (regen > ZSTD_BLOCK_MAX) return ZSTD_ERR;
		L->regen = regen;
		L->raw = src + hdr;
		if (type == 0)
		{
			if (hdr + regen > len) return ZSTD_ERR;
			L->kind = 0; L->consumed = hdr + regen;
			return 0;
		}
		if (hdr + 1 > len) return ZSTD_ERR;
		L->kind = 1; L->consumed = hdr + 1;
		return 0;
	}
	int streams = 4;
	if (fmt < 2)
	{
		if (len < 3) return ZSTD_ERR;
		uint32_t h = src[0] | ((uint32_t) src[1] << 8) | ((uint32_t) src[2] << 16);
		regen = (h >> 4) & 0x3FFu; comp = (h >> 14) & 0x3FFu; hdr = 3;
		if (fmt == 0) streams = 1;
	}
	else if (fmt == 2)
	{
		if (len < 4) return ZSTD_ERR;
		uint32_t h = src[0] | ((uint32_t) src[1] << 8) | ((uint32_t) src[2] << 16) | ((uint32_t) src[3] << 24);
		regen = (h >> 4) & 0x3FFFu; comp = (h >> 18) & 0x3FFFu; hdr = 4;
	}
	else
	{
		if (len < 5) return ZSTD_ERR;
		uint64_t h = (uint64_t) src[0] | ((uint64_t) src[1] << 8) | ((uint64_t) src[2] << 16) | ((uint64_t) src[3] << 24) | ((uint64_t) src[4] << 32);
		regen = (uint32_t) ((h >> 4) & 0x3FFFFu); comp = (uint32_t) ((h >> 22) & 0x3FFFFu); hdr = 5;
	}
	if (regen > ZSTD_BLOCK_MAX || hdr + comp > len) return ZSTD_ERR;
	const uint8_t *p = src + hdr;
	uint32_t left = comp;
	if (type == 2)
	{
		int used = zs_read_huf_tree(T, p, left);
		if (used < 0) return ZSTD_ERR;
		p += used; left -= (uint32_t) used;
	}
	else if (T.huf_log == 0) return ZSTD_ERR;       /* treeless without a previous table */
	L->kind = 2; L->regen = regen; L->consumed = hdr + comp; L->nstreams = streams;
	if (streams == 1)
	{
		L->sptr[0] = p; L->slen[0] = left; L->scount[0] = regen; L->sout[0] = 0;
		return 0;
	}
	if (left < 6) return ZSTD_ERR;
	const uint32_t s1 = p[0] | ((uint32_t) p[1] << 8), s2 = p[2] | ((uint32_t) p[3] << 8), s3 = p[4] | ((uint32_t) p[5] << 8);
	if (6 + s1 + s2 + s3 > left) return ZSTD_ERR;
	const uint32_t per = (regen + 3) / 4;
	if (3 * per > regen) return ZSTD_ERR;
	const uint8_t *q = p + 6;
	L->sptr[0] = q; L->slen[0] = s1; L->scount[0] = per; L->sout[0] = 0;
	L->sptr[1] = q + s1; L->slen[1] = s2; L->scount[1] = per; L->sout[1] = per;
	L->sptr[2] = q + s1 + s2; L->slen[2] = s3; L->scount[2] = per; L->sout[2] = 2 * per;
	L->sptr[3] = q + s1 + s2 + s3; L->slen[3] = left - 6 - s1 - s2 - s3; L->scount[3] = regen - 3 * per; L->sout[3] = 3 * per;
	return 0;
}

/* literals section, sequentially: fills lit[0..*nlit); returns bytes consumed or ZSTD_ERR */
CG_HDN inline int zs_literals(ZstdTables &T, const uint8_t *src, uint32_t len, uint8_t *lit, uint32_t *nlit)
{
	ZsLitPlan L;
	if (zs_literals_prepare(T, src, len, &L) < 0) return ZSTD_ERR;
	if (L.kind == 0) for (uint32_t i = 0; i < L.regen; i++) lit[i] = L.raw[i];
	else if (L.kind == 1) for (uint32_t i = 0; i < L.regen; i++) lit[i] = L.raw[0];
	else
		for (int k = 0; k < L.nstreams; k++)
			if (zs_huf_stream(T, L.sptr[k], L.slen[k], lit + L.sout[k], L.scount[k]) < 0) return ZSTD_ERR;
	*nlit = L.regen;
	return (int) L.consumed;
}

/* ---- sequences ---- */
/* code -> baseline (low 24 bits) and number of extra bits (high 8 bits) */
CG_HD void zs_ll_code(uint32_t code, uint32_t *base, int *bits)
{
	static const uint32_t t[36] = {
		0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
		16 | 1u << 24, 18 | 1u << 24, 20 | 1u << 24, 22 | 1u << 24, 24 | 2u << 24, 28 | 2u << 24, 32 | 3u << 24, 40 | 3u << 24,
		48 | 4u << 24, 64 | 6u << 24, 128 | 7u << 24, 256 | 8u << 24, 512 | 9u << 24, 1024 | 10u << 24, 2048 | 11u << 24,
		4096 | 12u << 24, 8192 | 13u << 24, 16384 | 14u << 24, 32768 | 15u << 24, 65536 | 16u << 24};
	const uint32_t v = t[code];
	*base = v & 0xFFFFFFu; *bits = (int) (v >> 24);
}
CG_HD void zs_ml_code(uint32_t code, uint32_t *base, int *bits)
{
	static const uint32_t t[53] = {
		3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
		35 | 1u << 24, 37 | 1u << 24, 39 | 1u << 24, 41 | 1u << 24, 43 | 2u << 24, 47 | 2u << 24, 51 | 3u << 24, 59 | 3u << 24,
		67 | 4u << 24, 83 | 4u << 24, 99 | 5u << 24, 131 | 7u << 24, 259 | 8u << 24, 515 | 9u << 24, 1027 | 10u << 24,
		2051 | 11u << 24, 4099 | 12u << 24, 8195 | 13u << 24, 16387 | 14u << 24, 32771 | 15u << 24, 65539 | 16u << 24};
	const uint32_t v = t[code];
	*base = v & 0xFFFFFFu; *bits = (int) (v >> 24);
}

/* one of the three symbol tables: mode 0 predefined, 1 RLE, 2 FSE description, 3 repeat; returns bytes consumed */
CG_HDN inline int zs_seq_table(ZstdTables &T, int which, int mode, const uint8_t *src, uint32_t len)
{
	FseEntry *e = which == 0 ? T.ll.e : which == 1 ? T.of.e : T.ml.e;
	int *log = which == 0 ? &T.ll.log : which == 1 ? &T.of.log : &T.ml.log;
	int *have = which == 0 ? &T.have_ll : which == 1 ? &T.have_of : &T.have_ml;
	const int max_log = which == 0 ? ZSTD_LL_LOG_MAX : which == 1 ? ZSTD_OF_LOG_MAX : ZSTD_ML_LOG_MAX;
	const int max_sym = which == 0 ? 35 : which == 1 ? 31 : 52;
	if (mode == 3) return *have ? 0 : ZSTD_ERR;
	*have = 1;
	if (mode == 1)
	{
		if (len < 1 || src[0] > max_sym) return ZSTD_ERR;
		zs_build_rle(e, log, src[0]);
		return 1;
	}
	int nsym = 0, used = 0;
	if (mode == 0)
	{
		static const int8_t ll_def[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
		static const int8_t of_def[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
		static const int8_t ml_def[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
								   1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
		if (which == 0) { nsym = 36; *log = 6; for (int i = 0; i < nsym; i++) T.norm[i] = ll_def[i]; }
		else if (which == 1) { nsym = 29; *log = 5; for (int i = 0; i < nsym; i++) T.norm[i] = of_def[i]; }
		else { nsym = 53; *log = 6; for (int i = 0; i < nsym; i++) T.norm[i] = ml_def[i]; }
	}
	else
	{
		used = zs_read_ncount(src, len, max_log, max_sym, T.norm, log, &nsym);
		if (used < 0) return ZSTD_ERR;
	}
	if (zs_build_fse(e, *log, T.norm, nsym, T.next) < 0) return ZSTD_ERR;
	return used;
}

CG_HD uint32_t zs_entry(const FseEntry *e, uint32_t state)
{
	const uint8_t *q = (const uint8_t *) (e + state);
#ifdef __CUDA_ARCH__
	return *(const uint32_t *) q;
#else
	return (uint32_t) q[0] | ((uint32_t) q[1] << 8) | ((uint32_t) q[2] << 16) | ((uint32_t) q[3] << 24);
#endif
}

/* overlapping forward copy (LZ77 semantics) */
CG_HD void zs_copy_match(uint8_t *dst, uint32_t op, uint32_t off, uint32_t n)
{
	const uint8_t *m = dst + op - off;
	uint8_t *d = dst + op;
	for (uint32_t i = 0; i < n; i++) d[i] = m[i];
}

/* the sequences section as an iterator: begin() reads the header and the three tables and primes the
 * backward bitstream, next() decodes one (literal length, match length, offset) applying the
 * repeat-offset history, end() checks that the bitstream was used up exactly */
struct ZsSeq
{
	ZsBack r;
	uint32_t sl, so, sm;
	uint32_t nseq, done;
};

CG_HDN inline int zs_sequences_begin(ZstdTables &T, const uint8_t *p, uint32_t left, ZsSeq *S)
{
	S->done = 0;
	if (left < 1) return ZSTD_ERR;
	uint32_t nseq;
	if (p[0] == 0) { nseq = 0; p += 1; left -= 1; }
	else if (p[0] < 128) { nseq = p[0]; p += 1; left -= 1; }
	else if (p[0] < 255) { if (left < 2) return ZSTD_ERR; nseq = ((uint32_t) (p[0] - 128) << 8) + p[1]; p += 2; left -= 2; }
	else { if (left < 3) return ZSTD_ERR; nseq = (uint32_t) p[1] + ((uint32_t) p[2] << 8) + 0x7F00u; p += 3; left -= 3; }
	S->nseq = nseq;
	if (nseq == 0) return left == 0 ? 0 : ZSTD_ERR;
	if (left < 1) return ZSTD_ERR;
	const uint32_t modes = p[0];
	if (modes & 3u) return ZSTD_ERR;
	p += 1; left -= 1;
	for (int which = 0; which < 3; which++)
	{
		int mode = (int) (modes >> (6 - 2 * which)) & 3;
		int u = zs_seq_table(T, which, mode, p, left);
		if (u < 0) return ZSTD_ERR;
		p += u; left -= (uint32_t) u;
	}
	if (!zs_back_init(S->r, p, left)) return ZSTD_ERR;
	S->sl = zs_back_read(S->r, T.ll.log);
	S->so = zs_back_read(S->r, T.of.log);
	S->sm = zs_back_read(S->r, T.ml.log);
	return S->r.pos < 0 ? ZSTD_ERR : 0;
}

CG_HD int zs_sequences_next(ZstdTables &T, ZsSeq *S, uint32_t *llen_out, uint32_t *mlen_out, uint32_t *offset_out)
{
	ZsBack &r = S->r;
	/* one 32-bit load per table entry: symbol | nbits << 8 | base << 16 */
	const uint32_t eo = zs_entry(T.of.e, S->so), em = zs_entry(T.ml.e, S->sm), el = zs_entry(T.ll.e, S->sl);
	const uint32_t of_code = eo & 0xffu, ml_sym = em & 0xffu, ll_sym = el & 0xffu;
	if (of_code > 31 || ml_sym > 52 || ll_sym > 35) return ZSTD_ERR;
	const uint32_t ofv = (1u << of_code) + zs_back_read(r, (int) of_code);
	uint32_t mbase, lbase; int mbits, lbits;
	zs_ml_code(ml_sym, &mbase, &mbits);
	zs_ll_code(ll_sym, &lbase, &lbits);
	const uint32_t mlen = mbase + zs_back_read(r, mbits);
	const uint32_t llen = lbase + zs_back_read(r, lbits);
	if (++S->done < S->nseq)
	{
		S->sl = (el >> 16) + zs_back_read(r, (int) ((el >> 8) & 0xffu));
		S->sm = (em >> 16) + zs_back_read(r, (int) ((em >> 8) & 0xffu));
		S->so = (eo >> 16) + zs_back_read(r, (int) ((eo >> 8) & 0xffu));
	}
	if (r.pos < 0) return ZSTD_ERR;
	/* repeat-offset history */
	uint32_t offset;
	if (ofv > 3)
	{
		offset = ofv - 3;
		T.rep[2] = T.rep[1]; T.rep[1] = T.rep[0]; T.rep[0] = offset;
	}
	else
	{
		const uint32_t idx = ofv - 1 + (llen == 0 ? 1u : 0u);       /* 0..3 */
		if (idx == 0) offset = T.rep[0];
		else
		{
			offset = idx < 3 ? T.rep[idx] : T.rep[0] - 1;
			if (offset == 0) return ZSTD_ERR;
			if (idx > 1) T.rep[2] = T.rep[1];
			T.rep[1] = T.rep[0];
			T.rep[0] = offset;
		}
	}
	*llen_out = llen; *mlen_out = mlen; *offset_out = offset;
	return 0;
}

CG_HD int zs_sequences_end(const ZsSeq *S) { return (S->nseq == 0 || S->r.pos == 0) ? 0 : ZSTD_ERR; }

/* one compressed block, sequentially; dst/op: the frame's output so far; returns the new op or ZSTD_ERR */
CG_HDN inline int64_t zs_compressed_block(ZstdTables &T, const uint8_t *src, uint32_t len, uint8_t *dst, uint32_t op, uint32_t cap,
										  uint8_t *lit)
{
	uint32_t nlit = 0;
	int used = zs_literals(T, src, len, lit, &nlit);
	if (used < 0) return ZSTD_ERR;
	ZsSeq S;
	if (zs_sequences_begin(T, src + used, len - (uint32_t) used, &S) < 0) return ZSTD_ERR;
	uint32_t litpos = 0;
	for (uint32_t s = 0; s < S.nseq; s++)
	{
		uint32_t llen, mlen, offset;
		if (zs_sequences_next(T, &S, &llen, &mlen, &offset) < 0) return ZSTD_ERR;
		if (llen > nlit - litpos || llen > cap - op) return ZSTD_ERR;
		for (uint32_t i = 0; i < llen; i++) dst[op + i] = lit[litpos + i];
		op += llen; litpos += llen;
		if (offset > op || mlen > cap - op) return ZSTD_ERR;
		zs_copy_match(dst, op, offset, mlen);
		op += mlen;
	}
	if (zs_sequences_end(&S) < 0) return ZSTD_ERR;
	const uint32_t rest = nlit - litpos;
	if (rest > cap - op) return ZSTD_ERR;
	for (uint32_t i = 0; i < rest; i++) dst[op + i] = lit[litpos + i];
	return (int64_t) op + rest;
}

/* frame header; returns the offset of the first block or ZSTD_ERR; *fcs = content size or ~0 when absent */
CG_HDN inline int zs_frame_header(const uint8_t *src, uint32_t len, uint32_t cap, uint64_t *fcs_out, int *checksum_out)
{
	if (len < 6) return ZSTD_ERR;
	if (src[0] != 0x28 || src[1] != 0xB5 || src[2] != 0x2F || src[3] != 0xFD) return ZSTD_ERR;
	const uint32_t fhd = src[4];
	const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1u, dict_flag = fhd & 3u;
	if (fhd & 0x08u) return ZSTD_ERR;                 /* reserved bit */
	if (dict_flag) return ZSTD_ERR;                   /* dictionaries: never used by the columnar writer */
	uint32_t pos = 5;
	if (!single) pos += 1;                            /* window descriptor: the whole output is addressable here */
	const uint32_t fcs_bytes = fcs_flag == 0 ? (single ? 1u : 0u) : fcs_flag == 1 ? 2u : fcs_flag == 2 ? 4u : 8u;
	if (pos + fcs_bytes > len) return ZSTD_ERR;
	uint64_t fcs = 0;
	for (uint32_t i = 0; i < fcs_bytes; i++) fcs |= (uint64_t) src[pos + i] << (8 * i);
	if (fcs_bytes == 2) fcs += 256;
	pos += fcs_bytes;
	if (fcs_bytes && fcs > cap) return ZSTD_ERR;
	*fcs_out = fcs_bytes ? fcs : ~0ull;
	*checksum_out = (int) ((fhd >> 2) & 1u);
	return (int) pos;
}

/*
 * One frame (what ZSTD_compress emits): returns the number of bytes produced (the caller compares it
 * with decompressedValueSize, like columnar_compression.c:226-232) or ZSTD_ERR.  lit: scratch of
 * ZSTD_BLOCK_MAX bytes.
 */
CG_HDN inline int64_t zs_decode_frame(ZstdTables &T, const uint8_t *src, uint32_t len, uint8_t *dst, uint32_t cap, uint8_t *lit)
{
	uint64_t fcs;
	int checksum;
	int hp = zs_frame_header(src, len, cap, &fcs, &checksum);
	if (hp < 0) return ZSTD_ERR;
	uint32_t pos = (uint32_t) hp;
	const uint32_t fcs_bytes = fcs != ~0ull;
	T.huf_log = 0; T.have_ll = T.have_of = T.have_ml = 0;
	T.rep[0] = 1; T.rep[1] = 4; T.rep[2] = 8;
	uint32_t op = 0;
	for (;;)
	{
		if (pos + 3 > len) return ZSTD_ERR;
		const uint32_t bh = src[pos] | ((uint32_t) src[pos + 1] << 8) | ((uint32_t) src[pos + 2] << 16);
		pos += 3;
		const uint32_t last = bh & 1u, type = (bh >> 1) & 3u, bsize = bh >> 3;
		if (type == 0)
		{
			if (bsize > ZSTD_BLOCK_MAX || pos + bsize > len || bsize > cap - op) return ZSTD_ERR;
			for (uint32_t i = 0; i < bsize; i++) dst[op + i] = src[pos + i];
			op += bsize; pos += bsize;
		}
		else if (type == 1)
		{
			if (bsize > ZSTD_BLOCK_MAX || pos + 1 > len || bsize > cap - op) return ZSTD_ERR;
			for (uint32_t i = 0; i < bsize; i++) dst[op + i] = src[pos];
			op += bsize; pos += 1;
		}
		else if (type == 2)
		{
			if (bsize > ZSTD_BLOCK_MAX || pos + bsize > len) return ZSTD_ERR;
			int64_t nop = zs_compressed_block(T, src + pos, bsize, dst, op, cap, lit);
			if (nop < 0 || nop - (int64_t) op > (int64_t) ZSTD_BLOCK_MAX) return ZSTD_ERR;
			op = (uint32_t) nop; pos += bsize;
		}
		else
			return ZSTD_ERR;
		if (last) break;
	}
	if (checksum) { if (pos + 4 > len) return ZSTD_ERR; pos += 4; }
	if (pos != len) return ZSTD_ERR;                   /* one frame, nothing behind it */
	if (fcs_bytes && fcs != op) return ZSTD_ERR;
	return (int64_t) op;
}

#endif
