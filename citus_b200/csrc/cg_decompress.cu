/*
 * cg_decompress.cu -- value-stream decompression on the GPU (K7).
 *
 * Replaces DecompressBuffer, backend/columnar/columnar_compression.c:165-270, for the two
 * byte-oriented LZ77 codecs a columnar chunk can carry:
 *
 *   COMPRESSION_LZ4   :186-208  LZ4_decompress_safe(buffer, out, len, decompressedSize), and the
 *                               result must be exactly decompressedSize bytes.  liblz4 is not part
 *                               of the reference tree; the block format is the published one
 *                               (token = literal length << 4 | match length - 4, 255-continued
 *                               lengths, literals, 16-bit little-endian offset; the last sequence
 *                               has literals only).
 *   COMPRESSION_PG_LZ :240-267  8-byte ColumnarCompressHeader {vl_len_, rawsize} (:38-42), VARSIZE
 *                               must equal the buffer length, then PostgreSQL's
 *                               pglz_decompress(..., check_complete = true) (src/common/
 *                               pg_lzcompress.c: control byte = 8 items, LSB first; bit set = tag
 *                               [len-3 | off>>8 <<4][off & 0xff][+1 length byte when len == 18],
 *                               bit clear = one literal byte).
 *
 * The exists bitmap is never compressed (columnar_writer.c:606-653); zstd chunks are refused at
 * staging (CG_EUNSUPPORTED).
 *
 * One warp decodes one chunk buffer (a C2 relation has ~10^5 of them per column, so the grid
 * is wide).  The compressed stream is pulled through a 256-byte shared-memory ring per warp
 * (aligned 4-byte loads, one refill per 128 bytes consumed), so tokens, lengths and offsets are
 * shared-memory broadcasts instead of dependent global loads.  A match of any length and any
 * overlap is copied in parallel: byte i of the match is out[pos - off + (i mod off)], which
 * only reads bytes written before the match began.  Lanes communicate through the output buffer,
 * ordered by __syncwarp().
 *
 * A malformed stream never writes outside the item's slot; it raises `flag` in *err and the
 * host reports CG_ECORRUPT ("cannot decompress the buffer") at its next synchronisation.
 */
#include "cg_internal.h"

#define CGD_WARPS 4
#define CGD_RING 256u

struct Stream
{
	const uint8_t *src;     /* 16-byte aligned */
	uint32_t len;           /* bytes in the stream */
	uint32_t loadable;      /* bytes that may be read (the slot, a multiple of 16) */
	uint32_t base;          /* ring holds [base, base + CGD_RING) */
	volatile uint8_t *ring;
	int lane;

	__device__ __forceinline__ void fill(uint32_t from)   /* 128 bytes at stream position `from` */
	{
		uint32_t p = from + 4u * lane;
		uint32_t w = p < loadable ? *(const uint32_t *) (src + p) : 0u;
		*(volatile uint32_t *) (ring + (p & (CGD_RING - 1))) = w;
	}
	__device__ __forceinline__ void open()
	{
		base = 0;
		fill(0);
		fill(128);
		__syncwarp();
	}
	/* after the call [ip, ip + 128) is in the ring */
	__device__ __forceinline__ void ensure(uint32_t ip)
	{
		while (ip >= base + 128u)
		{
			__syncwarp();                 /* every lane is done reading the half being replaced */
			fill(base + CGD_RING);
			base += 128u;
			__syncwarp();
		}
	}
	__device__ __forceinline__ uint32_t at(uint32_t p) const { return ring[p & (CGD_RING - 1)]; }
};

/* out[op .. op + n) = the n bytes at stream position ip (n may be large) */
__device__ __forceinline__ void copy_literals(Stream &s, uint32_t ip, uint8_t *out, uint32_t op, uint32_t n)
{
	while (n)
	{
		s.ensure(ip);
		uint32_t m = min(n, s.base + CGD_RING - ip);
		for (uint32_t i = s.lane; i < m; i += 32) out[op + i] = (uint8_t) s.at(ip + i);
		ip += m; op += m; n -= m;
	}
}

/* out[op .. op + n) = out[op - off ...] with LZ77 overlap semantics */
__device__ __forceinline__ void copy_match(uint8_t *out, uint32_t op, uint32_t off, uint32_t n, int lane)
{
	const uint8_t *m = out + op - off;
	if (off >= n)
		for (uint32_t i = lane; i < n; i += 32) out[op + i] = m[i];
	else
		for (uint32_t i = lane; i < n; i += 32) out[op + i] = m[i % off];
}

static __device__ bool lz4_block(Stream &s, uint8_t *out, uint32_t rawlen)
{
	uint32_t ip = 0, op = 0;
	const uint32_t clen = s.len;
	if (clen == 0) return false;
	for (;;)
	{
		s.ensure(ip);
		if (ip >= clen) return false;
		const uint32_t token = s.at(ip++);
		uint32_t lit = token >> 4;
		if (lit == 15)
		{
			uint32_t b;
			do
			{
				s.ensure(ip);
				if (ip >= clen) return false;
				b = s.at(ip++);
				lit += b;
				if (lit > rawlen) return false;
			} while (b == 255);
		}
		if (lit > clen - ip || lit > rawlen - op) return false;
		copy_literals(s, ip, out, op, lit);
		ip += lit; op += lit;
		if (ip == clen) break;                       /* the last sequence stops after its literals */
		s.ensure(ip);
		if (clen - ip < 2) return false;
		const uint32_t off = s.at(ip) | (s.at(ip + 1) << 8);
		ip += 2;
		uint32_t ml = token & 15;
		if (ml == 15)
		{
			uint32_t b;
			do
			{
				s.ensure(ip);
				if (ip >= clen) return false;
				b = s.at(ip++);
				ml += b;
				if (ml > rawlen) return false;
			} while (b == 255);
		}
		ml += 4;
		if (off == 0 || off > op || ml > rawlen - op) return false;
		__syncwarp();                                /* the bytes the match reads are visible */
		copy_match(out, op, off, ml, s.lane);
		op += ml;
		__syncwarp();
	}
	return op == rawlen;
}

static __device__ bool pglz_stream(Stream &s, uint8_t *out, uint32_t rawlen)
{
	/* ColumnarCompressHeader: int32 vl_len_ (4-byte varlena header, little endian: length << 2 | flags), int32 rawsize */
	if (s.len < 8) return false;
	uint32_t hdr = s.at(0) | (s.at(1) << 8) | (s.at(2) << 16) | (s.at(3) << 24);
	uint32_t raw = s.at(4) | (s.at(5) << 8) | (s.at(6) << 16) | (s.at(7) << 24);
	if (((hdr >> 2) & 0x3FFFFFFFu) != s.len) return false;     /* compressedDataSize + HDRSZ != buffer->len */
	if (raw != rawlen) return false;
	uint32_t sp = 8, dp = 0;
	const uint32_t srcend = s.len;
	while (sp < srcend && dp < rawlen)
	{
		s.ensure(sp);
		uint32_t ctrl = s.at(sp++);
		for (int c = 0; c < 8 && sp < srcend && dp < rawlen; c++, ctrl >>= 1)
		{
			s.ensure(sp);
			if (ctrl & 1)
			{
				uint32_t b0 = s.at(sp), b1 = s.at(sp + 1);
				uint32_t len = (b0 & 0x0f) + 3;
				uint32_t off = ((b0 & 0xf0) << 4) | b1;
				sp += 2;
				if (len == 18) len += s.at(sp++);
				if (sp > srcend || off == 0 || off > dp) return false;
				len = min(len, rawlen - dp);
				__syncwarp();
				copy_match(out, dp, off, len, s.lane);
				dp += len;
				__syncwarp();
			}
			else
			{
				if (s.lane == 0) out[dp] = (uint8_t) s.at(sp);
				sp++; dp++;
			}
		}
	}
	return dp == rawlen && sp == srcend;             /* check_complete */
}

__global__ void __launch_bounds__(CGD_WARPS * 32)
cg_decompress_kernel(uint8_t *arena, const DecodeItem *items, uint32_t nitems, unsigned long long *err,
					 unsigned long long flag)
{
	__shared__ __align__(16) uint8_t rings[CGD_WARPS][CGD_RING];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t idx = blockIdx.x * CGD_WARPS + warp;
	if (idx >= nitems) return;
	const DecodeItem it = items[idx];
	Stream s;
	s.src = arena + it.src;
	s.len = it.comp_len;
	s.loadable = (it.comp_len + 15u) & ~15u;
	s.ring = rings[warp];
	s.lane = lane;
	s.open();
	uint8_t *out = arena + it.dst;
	bool ok;
	if (it.kind == CG_COMPRESSION_LZ4) ok = lz4_block(s, out, it.raw_len);
	else if (it.kind == CG_COMPRESSION_PGLZ) ok = pglz_stream(s, out, it.raw_len);
	else ok = false;
	/* the slot's padding is zero, like every other arena slot */
	for (uint32_t i = it.raw_len + lane; i < it.padded; i += 32) out[i] = 0;
	if (!ok && lane == 0) atomicOr(err, flag);
}

int cg_launch_decompress(uint8_t *arena, const DecodeItem *items, uint64_t nitems, unsigned long long *err,
						 unsigned long long flag, cudaStream_t stream)
{
	if (nitems == 0) return CG_OK;
	unsigned blocks = (unsigned) ((nitems + CGD_WARPS - 1) / CGD_WARPS);
	cg_decompress_kernel<<<blocks, CGD_WARPS * 32, 0, stream>>>(arena, items, (uint32_t) nitems, err, flag);
	CG_CUDA(cudaGetLastError()); g_cg_launches++;
	return CG_OK;
}
