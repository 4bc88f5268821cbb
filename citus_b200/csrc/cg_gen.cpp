/*
 * cg_gen.cpp -- writes columnar relation images (bench / test tooling).
 *
 * Produces exactly the bytes the reference's writer would for compression = none, and for
 * compression = lz4 when liblz4 is the library the reference links (value streams go through
 * the same LZ4_compress_default, columnar_compression.c:75-97):
 *   exists bitmap   SerializeBoolArray       backend/columnar/columnar_writer.c:523-545
 *   value stream    SerializeSingleDatum     :555-585 (store_att_byval, NULL rows take no bytes)
 *   stripe layout   FlushStripe              :425-502 (per column: all exists buffers, then all
 *                                             value buffers; offsets relative to the stripe)
 *   min/max         UpdateChunkSkipNodeMinMax :663-718
 *   page framing    ColumnarStorageWrite / LogicalToPhysical  backend/columnar/columnar_storage.c:117-126,
 *                   stripes start on a page boundary (AlignReservation :756-772), metapage in
 *                   block 0 (:57-89), data from logical offset 2 * 8168
 * but column-at-a-time and with one thread per stripe, instead of the reference's
 * row-at-a-time ColumnarWriteRow.  tests/test_writer.py decodes these images with the
 * oracle's independent row-at-a-time reader and compares them byte for byte with the
 * oracle's writer.
 */
#include <dlfcn.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cg_internal.h"

/* liblz4 through dlopen (its header is not installed): the same LZ4_compress_default the reference's
 * CompressBuffer calls (columnar_compression.c:75-97), so the bytes are the reference writer's */
typedef int (*lz4_bound_fn)(int);
typedef int (*lz4_comp_fn)(const char *, char *, int, int);
static lz4_bound_fn g_lz4_bound;
static lz4_comp_fn g_lz4_compress;
/* libzstd likewise: ZSTD_compress(dst, cap, src, len, level) as columnar_compression.c:103-123 */
typedef size_t (*zstd_bound_fn)(size_t);
typedef size_t (*zstd_comp_fn)(void *, size_t, const void *, size_t, int);
typedef unsigned (*zstd_iserr_fn)(size_t);
static zstd_bound_fn g_zstd_bound;
static zstd_comp_fn g_zstd_compress;
static zstd_iserr_fn g_zstd_iserr;
static bool load_zstd()
{
	static int state = 0;
	if (state == 0)
	{
		void *h = dlopen("libzstd.so.1", RTLD_NOW);
		if (h)
		{
			g_zstd_bound = (zstd_bound_fn) dlsym(h, "ZSTD_compressBound");
			g_zstd_compress = (zstd_comp_fn) dlsym(h, "ZSTD_compress");
			g_zstd_iserr = (zstd_iserr_fn) dlsym(h, "ZSTD_isError");
		}
		state = (g_zstd_bound && g_zstd_compress && g_zstd_iserr) ? 1 : -1;
	}
	return state == 1;
}

static bool load_lz4()
{
	static int state = 0;
	if (state == 0)
	{
		void *h = dlopen("liblz4.so.1", RTLD_NOW);
		if (h)
		{
			g_lz4_bound = (lz4_bound_fn) dlsym(h, "LZ4_compressBound");
			g_lz4_compress = (lz4_comp_fn) dlsym(h, "LZ4_compress_default");
		}
		state = (g_lz4_bound && g_lz4_compress) ? 1 : -1;
	}
	return state == 1;
}

struct CgGenRelation
{
	std::vector<CgColumnDesc> cols;
	uint8_t *pages = nullptr;
	uint64_t nblocks = 0;
	std::vector<CgStripe> stripes;
	std::vector<CgSkipNode> nodes;
};

static int g_gen_compression = CG_COMPRESSION_NONE;

static inline uint64_t splitmix64(uint64_t x)
{
	x += 0x9e3779b97f4a7c15ull;
	x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
	x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
	return x ^ (x >> 31);
}

struct SynthSource
{
	const CgGenColumn *cols;
	uint64_t seed, first_row;
	inline bool isnull(int c, uint64_t row) const
	{
		if (cols[c].null_ppm == 0) return false;
		return splitmix64(~seed ^ ((uint64_t) c << 56) ^ (first_row + row)) % 1000000ull < cols[c].null_ppm;
	}
	inline int64_t value(int c, uint64_t row) const
	{
		const CgGenColumn &g = cols[c];
		if (g.kind == CG_GEN_SEQUENCE) return g.lo + (int64_t) (first_row + row);
		uint64_t span = (uint64_t) (g.hi - g.lo);
		uint64_t h = splitmix64(seed ^ ((uint64_t) c << 56) ^ (first_row + row));
		return g.lo + (int64_t) (span ? h % span : 0);
	}
	inline bool nullable(int c) const { return cols[c].null_ppm != 0; }
};

struct ArraySource
{
	const int64_t *const *values;
	const uint8_t *const *nulls;
	inline bool isnull(int c, uint64_t row) const { return nulls && nulls[c] && nulls[c][row]; }
	inline int64_t value(int c, uint64_t row) const { return values[c][row]; }
	inline bool nullable(int c) const { return nulls && nulls[c]; }
};

static void storage_write(uint8_t *pages, uint64_t logical, const uint8_t *data, uint64_t amount)
{
	uint64_t done = 0;
	while (done < amount)
	{
		uint64_t L = logical + done;
		uint64_t blockno = L / CG_BYTES_PER_PAGE;
		uint32_t offset = CG_PAGE_HEADER + (uint32_t) (L % CG_BYTES_PER_PAGE);
		uint64_t n = std::min<uint64_t>(amount - done, CG_BLCKSZ - offset);
		memcpy(pages + blockno * CG_BLCKSZ + offset, data + done, n);
		done += n;
	}
}

static void init_page_header(uint8_t *page, uint16_t pd_lower)
{
	/* [PG] PageHeaderData: pd_lsn(8) pd_checksum(2) pd_flags(2) pd_lower(2) pd_upper(2)
	 * pd_special(2) pd_pagesize_version(2) pd_prune_xid(4) */
	memset(page, 0, CG_PAGE_HEADER);
	uint16_t upper = CG_BLCKSZ, special = CG_BLCKSZ, psv = CG_BLCKSZ | 4;
	memcpy(page + 12, &pd_lower, 2);
	memcpy(page + 14, &upper, 2);
	memcpy(page + 16, &special, 2);
	memcpy(page + 18, &psv, 2);
}

/* one column chunk: exists bitmap (SerializeBoolArray :523-545), NULL-compacted value stream
 * (SerializeSingleDatum :555-585) and the skip node's min/max (UpdateChunkSkipNodeMinMax :663-718).
 * n.row_count is set by the caller; returns the number of value bytes written to vbuf. */
template <typename Source>
static uint64_t encode_chunk(const Source &src, int c, int len, bool isf, uint64_t c0, CgSkipNode &n, uint8_t *ebuf, uint8_t *vbuf)
{
	uint32_t rows = (uint32_t) n.row_count;
	memset(ebuf, 0, n.exists_length);
	uint8_t *vp = vbuf;
	bool has = false;
	int64_t mn = 0, mx = 0;
	double fmn = 0, fmx = 0;
	const bool nullable = src.nullable(c);
	for (uint32_t i = 0; i < rows; i++)
	{
		if (nullable && src.isnull(c, c0 + i)) continue;
		ebuf[i >> 3] |= (uint8_t) (1u << (i & 7));
		int64_t v = src.value(c, c0 + i);
		if (isf)
		{
			double d;
			memcpy(&d, &v, 8);
			if (len == 4) { float f = (float) d; memcpy(vp, &f, 4); d = f; }
			else memcpy(vp, &d, 8);
			/* float8 btree order: NaN above everything */
			auto fcmp = [](double x, double y) {
				bool xn = x != x, yn = y != y;
				if (xn || yn) return (int) xn - (int) yn;
				return (int) (x > y) - (int) (x < y);
			};
			if (!has) { fmn = fmx = d; has = true; }
			else
			{
				if (fcmp(d, fmn) < 0) fmn = d;
				if (fcmp(d, fmx) > 0) fmx = d;
			}
		}
		else
		{
			memcpy(vp, &v, (size_t) len);
			if (!has) { mn = mx = v; has = true; }
			else { if (v < mn) mn = v; if (v > mx) mx = v; }
		}
		vp += len;
	}
	n.has_minmax = has ? 1 : 0;
	if (has)
	{
		if (isf) { memcpy(&n.min_value, &fmn, 8); memcpy(&n.max_value, &fmx, 8); }
		else { n.min_value = mn; n.max_value = mx; }
	}
	return (uint64_t) (vp - vbuf);
}

template <typename Source>
static int write_relation(const Source &src, const CgColumnDesc *cols, int natts, uint64_t nrows,
						  uint64_t stripe_row_limit, uint32_t chunk_row_limit, int nthreads, int compression, CgGenRelation **out)
{
	if (compression != CG_COMPRESSION_NONE && compression != CG_COMPRESSION_LZ4 && compression != CG_COMPRESSION_ZSTD)
		return cg_set_error(CG_EUNSUPPORTED, "the shard writer compresses with none, lz4 or zstd");
	if (compression == CG_COMPRESSION_LZ4 && !load_lz4())
		return cg_set_error(CG_EUNSUPPORTED, "liblz4.so.1 is not available");
	if (compression == CG_COMPRESSION_ZSTD && !load_zstd())
		return cg_set_error(CG_EUNSUPPORTED, "libzstd.so.1 is not available");
	if (natts <= 0 || natts > 256) return cg_set_error(CG_EINVAL, "natts %d", natts);
	/* include/columnar/columnar.h:42-45 limits */
	if (stripe_row_limit < 1000 || stripe_row_limit > 10000000) return cg_set_error(CG_EINVAL, "stripe_row_limit out of range");
	if (chunk_row_limit < 1000 || chunk_row_limit > 100000) return cg_set_error(CG_EINVAL, "chunk_group_row_limit out of range");
	for (int c = 0; c < natts; c++)
	{
		int l = cols[c].attlen;
		if (l != 1 && l != 2 && l != 4 && l != 8) return cg_set_error(CG_EUNSUPPORTED, "attlen %d", l);
	}
	if (nthreads < 1) nthreads = 1;
	CgGenRelation *g = new CgGenRelation();
	g->cols.assign(cols, cols + natts);

	uint64_t nstripes = (nrows + stripe_row_limit - 1) / stripe_row_limit;
	g->stripes.resize(nstripes);
	std::vector<uint32_t> chunk_counts(nstripes);
	uint64_t total_nodes = 0;
	for (uint64_t s = 0; s < nstripes; s++)
	{
		uint64_t r0 = s * stripe_row_limit, r1 = std::min(nrows, r0 + stripe_row_limit);
		uint32_t cc = (uint32_t) ((r1 - r0 + chunk_row_limit - 1) / chunk_row_limit);
		chunk_counts[s] = cc;
		CgStripe &st = g->stripes[s];
		memset(&st, 0, sizeof st);
		st.id = s + 1;
		st.row_count = r1 - r0;
		st.first_row_number = 1 + r0;
		st.column_count = (uint32_t) natts;
		st.chunk_row_count = chunk_row_limit;
		st.chunk_count = cc;
		st.skipnode_base = (uint32_t) total_nodes;
		total_nodes += (uint64_t) natts * cc;
	}
	if (total_nodes > 0x7fffffffull) { delete g; return cg_set_error(CG_EUNSUPPORTED, "too many chunks"); }
	g->nodes.resize(total_nodes);

	/* compressed relations: every chunk is encoded and compressed first (its size decides the
	 * layout, columnar_writer.c:590-654: the value buffer is replaced by the compressor's output
	 * whenever CompressBuffer succeeds, even if that is larger); the stripe's bytes wait in
	 * `staged` until the offsets are known */
	std::vector<std::vector<uint8_t>> staged(compression != CG_COMPRESSION_NONE ? nstripes : 0);
	if (compression != CG_COMPRESSION_NONE)
	{
#pragma omp parallel num_threads(nthreads)
		{
			std::vector<uint8_t> vbuf((size_t) chunk_row_limit * 8 + 16), ebuf(chunk_row_limit / 8 + 16);
			const size_t raw_max = (size_t) chunk_row_limit * 8;
			std::vector<uint8_t> cbuf((compression == CG_COMPRESSION_LZ4 ? (size_t) g_lz4_bound((int) raw_max) : g_zstd_bound(raw_max)) + 16);
#pragma omp for schedule(dynamic, 1)
			for (int64_t s = 0; s < (int64_t) nstripes; s++)
			{
				CgStripe &st = g->stripes[s];
				uint64_t r0 = (uint64_t) s * stripe_row_limit;
				std::vector<uint8_t> &buf = staged[s];
				for (int c = 0; c < natts; c++)
				{
					const int len = cols[c].attlen;
					const bool isf = cols[c].type_class == CG_TYPE_FLOAT;
					std::vector<uint8_t> values;     /* this column's value buffers, behind its exists buffers */
					for (uint32_t k = 0; k < st.chunk_count; k++)
					{
						CgSkipNode &n = g->nodes[st.skipnode_base + (uint32_t) c * st.chunk_count + k];
						memset(&n, 0, sizeof n);
						uint64_t c0 = r0 + (uint64_t) k * chunk_row_limit;
						uint64_t c1 = std::min(r0 + st.row_count, c0 + chunk_row_limit);
						n.row_count = c1 - c0;
						n.exists_length = (n.row_count + 7) / 8;
						n.compression_level = 3;
						uint64_t raw = encode_chunk(src, c, len, isf, c0, n, ebuf.data(), vbuf.data());
						n.decompressed_size = raw;
						n.exists_offset = buf.size();
						buf.insert(buf.end(), ebuf.data(), ebuf.data() + n.exists_length);
						int clen;
					if (compression == CG_COMPRESSION_LZ4)
						clen = g_lz4_compress((const char *) vbuf.data(), (char *) cbuf.data(), (int) raw, (int) cbuf.size());
					else
					{
						size_t zn = g_zstd_compress(cbuf.data(), cbuf.size(), vbuf.data(), raw, 3);   /* columnar.compression_level default */
						clen = g_zstd_iserr(zn) ? 0 : (int) zn;
					}
						n.value_offset = values.size();             /* rebased below */
						if (clen > 0)
						{
							n.compression_type = compression;
							n.value_length = (uint64_t) clen;
							values.insert(values.end(), cbuf.data(), cbuf.data() + clen);
						}
						else
						{
							n.compression_type = CG_COMPRESSION_NONE;
							n.value_length = raw;
							values.insert(values.end(), vbuf.data(), vbuf.data() + raw);
						}
					}
					for (uint32_t k = 0; k < st.chunk_count; k++)
						g->nodes[st.skipnode_base + (uint32_t) c * st.chunk_count + k].value_offset += buf.size();
					buf.insert(buf.end(), values.begin(), values.end());
				}
				st.data_length = buf.size();
			}
		}
	}
	else
	{
		/* pass 1: value counts per (stripe, column, chunk) -> sizes and offsets */
	#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
		for (int64_t s = 0; s < (int64_t) nstripes; s++)
		{
			CgStripe &st = g->stripes[s];
			uint64_t r0 = (uint64_t) s * stripe_row_limit;
			uint64_t off = 0;
			for (int c = 0; c < natts; c++)
			{
				for (uint32_t k = 0; k < st.chunk_count; k++)
				{
					CgSkipNode &n = g->nodes[st.skipnode_base + (uint32_t) c * st.chunk_count + k];
					memset(&n, 0, sizeof n);
					uint64_t c0 = r0 + (uint64_t) k * chunk_row_limit;
					uint64_t c1 = std::min(r0 + st.row_count, c0 + chunk_row_limit);
					n.row_count = c1 - c0;
					uint64_t nonnull = n.row_count;
					if (src.nullable(c))
					{
						nonnull = 0;
						for (uint64_t r = c0; r < c1; r++) nonnull += !src.isnull(c, r);
					}
					n.decompressed_size = nonnull * cols[c].attlen;
					n.value_length = n.decompressed_size;
					n.exists_length = (n.row_count + 7) / 8;
					n.compression_type = CG_COMPRESSION_NONE;
					n.compression_level = 3;     /* columnar.c:43 default, recorded even for none */
					n.exists_offset = off;
					off += n.exists_length;
				}
				for (uint32_t k = 0; k < st.chunk_count; k++)
				{
					CgSkipNode &n = g->nodes[st.skipnode_base + (uint32_t) c * st.chunk_count + k];
					n.value_offset = off;
					off += n.value_length;
				}
			}
			st.data_length = off;
		}
	}
	uint64_t reserved = CG_FIRST_LOGICAL_OFFSET;
	for (uint64_t s = 0; s < nstripes; s++)
	{
		if (reserved % CG_BYTES_PER_PAGE) reserved = (reserved / CG_BYTES_PER_PAGE + 1) * CG_BYTES_PER_PAGE;
		g->stripes[s].file_offset = reserved;
		reserved += g->stripes[s].data_length;
	}
	g->nblocks = std::max<uint64_t>(2, (reserved + CG_BYTES_PER_PAGE - 1) / CG_BYTES_PER_PAGE);
	g->pages = (uint8_t *) aligned_alloc(4096, g->nblocks * CG_BLCKSZ);
	if (!g->pages)
	{
		const unsigned long long nblocks = g->nblocks;
		delete g;
		return cg_set_error(CG_ENOMEM, "cannot allocate %llu pages", nblocks);
	}

	/* metapage + empty block (columnar_storage.c:57-89) */
	memset(g->pages, 0, 2 * CG_BLCKSZ);
	{
		struct { uint32_t versionMajor, versionMinor; uint64_t storageId, reservedStripeId, reservedRowNumber, reservedOffset; uint8_t unloggedReset; } meta;
		memset(&meta, 0, sizeof meta);
		meta.versionMajor = 2; meta.versionMinor = 0; meta.storageId = 10000000;
		meta.reservedStripeId = nstripes + 1; meta.reservedRowNumber = nrows + 1; meta.reservedOffset = reserved;
		init_page_header(g->pages, (uint16_t) (CG_PAGE_HEADER + sizeof meta));
		memcpy(g->pages + CG_PAGE_HEADER, &meta, sizeof meta);
		init_page_header(g->pages + CG_BLCKSZ, CG_PAGE_HEADER);
	}

	/* pass 2: encode chunks straight into the pages */
#pragma omp parallel num_threads(nthreads)
	{
		std::vector<uint8_t> vbuf((size_t) chunk_row_limit * 8 + 16), ebuf(chunk_row_limit / 8 + 16);
#pragma omp for schedule(dynamic, 1)
		for (int64_t s = 0; s < (int64_t) nstripes; s++)
		{
			const CgStripe &st = g->stripes[s];
			uint64_t r0 = (uint64_t) s * stripe_row_limit;
			/* page headers of the stripe's blocks */
			uint64_t b0 = st.file_offset / CG_BYTES_PER_PAGE;
			uint64_t left = st.data_length;
			for (uint64_t b = b0; left > 0; b++)
			{
				uint64_t n = std::min<uint64_t>(left, CG_BYTES_PER_PAGE);
				init_page_header(g->pages + b * CG_BLCKSZ, (uint16_t) (CG_PAGE_HEADER + n));
				if (n < CG_BYTES_PER_PAGE)
					memset(g->pages + b * CG_BLCKSZ + CG_PAGE_HEADER + n, 0, CG_BYTES_PER_PAGE - n);
				left -= n;
			}
			if (compression != CG_COMPRESSION_NONE)
			{
				storage_write(g->pages, st.file_offset, staged[s].data(), staged[s].size());
				std::vector<uint8_t>().swap(staged[s]);
				continue;
			}
			for (int c = 0; c < natts; c++)
			{
				const int len = cols[c].attlen;
				const bool isf = cols[c].type_class == CG_TYPE_FLOAT;
				for (uint32_t k = 0; k < st.chunk_count; k++)
				{
					CgSkipNode &n = g->nodes[st.skipnode_base + (uint32_t) c * st.chunk_count + k];
					uint64_t c0 = r0 + (uint64_t) k * chunk_row_limit;
					encode_chunk(src, c, len, isf, c0, n, ebuf.data(), vbuf.data());
					storage_write(g->pages, st.file_offset + n.exists_offset, ebuf.data(), n.exists_length);
					storage_write(g->pages, st.file_offset + n.value_offset, vbuf.data(), n.value_length);
				}
			}
		}
	}
	*out = g;
	return CG_OK;
}

extern "C" int cg_gen_relation(const CgGenColumn *cols, int32_t natts, uint64_t nrows, uint64_t first_row,
							   uint64_t seed, uint64_t stripe_row_limit, uint32_t chunk_row_limit,
							   int32_t nthreads, CgGenRelation **out)
{
	if (!cols || !out) return cg_set_error(CG_EINVAL, "NULL argument");
	std::vector<CgColumnDesc> desc(natts > 0 ? natts : 0);
	for (int c = 0; c < natts; c++)
	{
		desc[c].attlen = cols[c].attlen;
		desc[c].type_class = CG_TYPE_INT;
		if (cols[c].kind == CG_GEN_UNIFORM && cols[c].hi < cols[c].lo) return cg_set_error(CG_EINVAL, "column %d: hi < lo", c);
	}
	SynthSource src{cols, seed, first_row};
	return write_relation(src, desc.data(), natts, nrows, stripe_row_limit, chunk_row_limit, nthreads, g_gen_compression, out);
}

/* columnar.compression for the relations written after the call (none or lz4; tooling, not thread safe) */
extern "C" int cg_gen_set_compression(int32_t compression)
{
	if (compression != CG_COMPRESSION_NONE && compression != CG_COMPRESSION_LZ4 && compression != CG_COMPRESSION_ZSTD)
		return cg_set_error(CG_EUNSUPPORTED, "the shard writer compresses with none, lz4 or zstd");
	g_gen_compression = compression;
	return CG_OK;
}

extern "C" int cg_write_relation(const CgColumnDesc *cols, int32_t natts, const int64_t *const *values,
								 const uint8_t *const *nulls, uint64_t nrows, uint64_t stripe_row_limit,
								 uint32_t chunk_row_limit, CgGenRelation **out)
{
	if (!cols || !values || !out) return cg_set_error(CG_EINVAL, "NULL argument");
	ArraySource src{values, nulls};
	return write_relation(src, cols, natts, nrows, stripe_row_limit, chunk_row_limit, omp_get_num_procs() > 8 ? 8 : omp_get_num_procs(), g_gen_compression, out);
}

extern "C" int cg_gen_relation_view(const CgGenRelation *g, CgRelation *view)
{
	if (!g || !view) return cg_set_error(CG_EINVAL, "NULL argument");
	view->pages = g->pages;
	view->nblocks = g->nblocks;
	view->stripes = g->stripes.data();
	view->nstripes = (int32_t) g->stripes.size();
	view->nodes = g->nodes.data();
	view->nnodes = (int32_t) g->nodes.size();
	view->columns = g->cols.data();
	view->natts = (int32_t) g->cols.size();
	return CG_OK;
}

extern "C" void cg_gen_relation_free(CgGenRelation *g)
{
	if (!g) return;
	free(g->pages);
	delete g;
}
