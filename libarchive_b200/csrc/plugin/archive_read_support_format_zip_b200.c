/*
 * archive_read_support_format_zip_b200.c — libarchive ZIP format module
 * (seekable reader) whose bodies are decoded on a B200 through
 * include/b200inflate.h instead of zlib.
 *
 * Drop-in for the reference translation unit
 * libarchive/archive_read_support_format_zip.c: it defines the same three
 * public entry points (archive_read_support_format_zip, _seekable,
 * _streamable; zip.c:3320-3328, 4359-4404, 3567-3610) and registers with the
 * unmodified read core through __archive_read_register_format
 * (archive_read_private.h:229-240).  Written from scratch against that
 * interface; behaviour it mirrors is cited inline.
 *
 * Design (not the reference's): the reference decodes one entry at a time,
 * 256 KiB per read_data call, with zlib.  Here the bid looks at the last 16 KiB
 * only (like zip.c:3720-3773); the central directory is flattened once
 * (b2i_zip_index_build / _cb) and the entries' bodies are decoded ahead of the
 * caller in device passes:
 *   - a small archive whose image the read core hands out in one piece is
 *     decoded by ONE b2i_decode_host call on the first body request;
 *   - anything larger goes through the streaming engine (b2i_pipe_*): windows of
 *     bounded output are staged into pinned buffers, decoded by all the GPUs the
 *     archive is worth, and served in order while the next windows are in flight;
 *     a file-backed source (archive_read_open_filename) is read window by window
 *     through __archive_read_seek / __archive_read_ahead on the caller's thread,
 *     so resident memory is bounded by the ring, not by the archive.
 * read_data serves slices of the decoded buffers with the reference's block
 * contract, return codes and messages.  There is no CPU inflate here: methods
 * other than 0/8 and encrypted entries are refused with the reference's own
 * messages.
 */
#include "archive_platform.h"

#ifdef HAVE_ERRNO_H
#include <errno.h>
#endif
#ifdef HAVE_STDLIB_H
#include <stdlib.h>
#endif
#ifdef HAVE_STRING_H
#include <string.h>
#endif
#include <wchar.h>

#include "archive.h"
#include "archive_entry.h"
#include "archive_entry_locale.h"
#include "archive_private.h"
#include "archive_read_private.h"
#include "archive_string.h"

#include "b200inflate.h"
#include "b200_ctx_pool.h"
#include "zip_b200_local.h"

#define ZB_MAX_GPUS        8
#define ZB_IMAGE_MAX       ((int64_t)32 << 20)   /* a file up to this size is taken as one image */
#define ZB_BATCH_MAX       ((uint64_t)24 << 20)  /* in + out bytes decoded by one call, no pipeline */
#define ZB_FETCH_WINDOW    ((size_t)4 << 20)     /* read-ahead of the index walk over a file */

struct zip_b200 {
	struct zb_common    c;             /* must be first */
	b2i_zip_index       ix;
	int                 have_index;
	b2i_stream_desc    *descs;
	b2i_stream_result  *res;           /* one-call mode: all results */
	size_t             *desc_of;       /* entry -> descriptor index, or SIZE_MAX */
	size_t              ndesc;
	unsigned char      *out;           /* one-call mode, pinned: decoded bytes of the batch */
	size_t              out_bytes, out_cap;
	unsigned char     **retry;         /* one-call mode, per descriptor: buffer after an overflow retry */
	const unsigned char *image;        /* the whole archive, as handed out by the read core */
	size_t              image_len;
	int                 file_mode;     /* the source is read in windows, there is no image */
	int64_t             file_size;
	int64_t             cache_off;     /* file mode: what the last fetch left in the read core's buffer */
	size_t              cache_len;
	const unsigned char *cache;
	struct archive_read *a;
	b2i_pipe           *pipe;          /* streaming mode */
	b2i_ctx            *ctxs[ZB_MAX_GPUS];
	int                 nctx;
	size_t              cur_desc;      /* streaming mode: the descriptor cur_* belong to */
	const unsigned char *cur_out, *cur_in;
	b2i_stream_result   cur_res;
	unsigned char      *cur_retry;
	size_t              next;          /* next entry to hand out */
	size_t              cur;
	int64_t             delivered;     /* entry_uncompressed_bytes_read */
	int                 decoded, end_of_entry;
};

/* ---- file-backed sources: byte ranges through the read core, on the caller's thread ---- */
static const uint8_t *
zf_fetch(void *user, uint64_t off, size_t len)
{
	struct zip_b200 *z = user;
	struct archive_read *a = z->a;
	ssize_t avail = 0;
	size_t want;
	const void *p;

	if (z->cache != NULL && (int64_t)off >= z->cache_off &&
	    off + len <= (uint64_t)z->cache_off + z->cache_len)
		return (z->cache + (off - (uint64_t)z->cache_off));
	z->cache = NULL;
	if (off > (uint64_t)z->file_size || (uint64_t)z->file_size - off < len)
		return (NULL);
	if (__archive_read_seek(a, (int64_t)off, SEEK_SET) < 0)
		return (NULL);
	want = len > ZB_FETCH_WINDOW ? len : ZB_FETCH_WINDOW;
	if ((uint64_t)z->file_size - off < want)
		want = (size_t)((uint64_t)z->file_size - off);
	if ((p = __archive_read_ahead(a, want, &avail)) == NULL || avail < (ssize_t)len)
		return (NULL);
	z->cache = p;
	z->cache_off = (int64_t)off;
	z->cache_len = (size_t)avail;
	return (p);
}

/* the streaming engine asks for the compressed span of a window */
static int
zf_fill(void *user, uint64_t off, uint64_t len, void *dst)
{
	struct zip_b200 *z = user;
	struct archive_read *a = z->a;
	unsigned char *d = dst;

	z->cache = NULL;
	if (__archive_read_seek(a, (int64_t)off, SEEK_SET) < 0)
		return (B2I_E_FORMAT);
	while (len > 0) {
		ssize_t avail = 0;
		size_t n;
		const void *p = __archive_read_ahead(a, 1, &avail);
		if (p == NULL || avail <= 0)
			return (B2I_E_FORMAT);
		n = (uint64_t)avail < len ? (size_t)avail : (size_t)len;
		memcpy(d, p, n);
		__archive_read_consume(a, (int64_t)n);
		d += n;
		len -= n;
	}
	return (B2I_OK);
}

/* ---- bid: is there a usable end-of-central-directory record? (zip.c:3720-3773) */
static int
zip_b200_bid(struct archive_read *a, int best_bid)
{
	struct zip_b200 *z = (struct zip_b200 *)a->format->data;
	int64_t size;
	ssize_t avail = 0;
	size_t tail;
	const void *p;
	char err[128];
	int rc;

	if (best_bid > 32)
		return (-1);
	size = __archive_read_seek(a, 0, SEEK_END);
	if (size <= 0)
		return (0);
	/* like the reference, look at the last 16 KiB only: a large file that is not a ZIP
	 * archive costs one small read, whatever other formats are registered */
	tail = size < 16384 ? (size_t)size : 16384;
	if (__archive_read_seek(a, size - (int64_t)tail, SEEK_SET) < 0 ||
	    (p = __archive_read_ahead(a, tail, NULL)) == NULL ||
	    !b2i_zip_probe_tail(p, tail, (uint64_t)size))
		return (0);
	if (z->have_index) {
		b2i_zip_index_free(&z->ix);
		z->have_index = 0;
	}
	z->a = a;
	z->file_size = size;
	z->cache = NULL;
	/* a memory source hands out its whole image in one block; a small file is taken whole
	 * too; anything else is walked through bounded fetches and later read in windows */
	if (__archive_read_seek(a, 0, SEEK_SET) < 0 || (p = __archive_read_ahead(a, 1, &avail)) == NULL)
		return (0);
	z->file_mode = !(avail >= size || size <= ZB_IMAGE_MAX);
	if (getenv("B2I_ZIP_FILE_MODE") != NULL && avail < size)
		z->file_mode = 1;                      /* tests: windows even for small files */
	if (!z->file_mode) {
		if ((p = __archive_read_ahead(a, (size_t)size, NULL)) == NULL)
			return (0);
		rc = b2i_zip_index_build(p, (size_t)size, &z->ix, err);
	} else {
		rc = b2i_zip_index_build_cb(zf_fetch, z, (uint64_t)size, &z->ix, err);
	}
	/* a directory that cannot be walked leaves the archive to the streaming reader */
	if (rc != B2I_OK)
		return (0);
	z->have_index = 1;
	z->image_len = (size_t)size;
	return (32);            /* encrypted entries are reported as their headers are read (zip.c:962-972) */
}

static int
zip_b200_options(struct archive_read *a, const char *key, const char *val)
{
	return (zb_options(a, (struct zb_common *)a->format->data, key, val));
}

/* ---- the batch: one descriptor per decodable entry, one device pass ---------- */
static int
decodable(const b2i_zip_entry *e)
{
	unsigned type = e->mode & AE_IFMT;
	if (e->warn & (B2I_ZW_BAD_LOCAL_HEADER | B2I_ZW_TRUNCATED))
		return (0);
	if (type != AE_IFREG && !(type == AE_IFLNK && e->compressed_size <= 64 * 1024))
		return (0);
	if (e->zip_flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED))
		return (0);
	if (e->method != 0 && e->method != 8)
		return (0);
	return (e->compressed_size >= 1);
}

/* how many GPUs an archive of this size is worth (contexts cost a second to create the
 * first time; later handles borrow them from the pool) */
static int
gpus_for(uint64_t total_out)
{
	const char *ev = getenv("B2I_PLUGIN_GPUS");
	int have = b2i_device_count(), want;

	if (have < 1)
		have = 1;
	if (have > ZB_MAX_GPUS)
		have = ZB_MAX_GPUS;
	if (ev != NULL && atoi(ev) >= 1)
		return (atoi(ev) < have ? atoi(ev) : have);
	want = (int)(total_out / ((uint64_t)2 << 30));      /* one more GPU per 2 GiB of output */
	if (want < 1)
		want = 1;
	return (want < have ? want : have);
}

/* overflow: a stream that outgrows its directory size is decoded again, alone, with room,
 * so that the reference's "wrong size" message can quote the true count */
static unsigned char *
retry_with_room(struct zip_b200 *z, b2i_ctx *ctx, const void *in, size_t in_bytes, const b2i_stream_desc *d0,
    b2i_stream_result *res)
{
	size_t cap = (size_t)d0->out_cap;
	unsigned char *buf = NULL;

	(void)z;
	const size_t most = (size_t)d0->in_len * 1032u + 65536u;   /* deflate's expansion limit */

	while (res->status == B2I_S_OUT_OVERFLOW && cap < most) {
		b2i_stream_desc d = *d0;
		cap = cap < 4096 ? 16384 : cap * 4;
		if (cap > most)
			cap = most;
		b200_buf_release_tagged(buf);
		if ((buf = b200_buf_acquire_tagged(cap + 16)) == NULL)
			break;
		d.out_off = 0;
		d.out_cap = cap;
		if (b2i_decode_host(ctx, in, in_bytes, &d, 1, buf, cap, res) != B2I_OK)
			break;
	}
	return (buf);
}

/* first body request: descriptors for every decodable entry, then either ONE device pass
 * (small archive) or the streaming engine */
static int
zip_b200_prepare(struct archive_read *a, struct zip_b200 *z)
{
	size_t i, n = 0, out = 0;
	uint64_t in_lo = UINT64_MAX, in_hi = 0;
	int rc, g;

	if (z->decoded)
		return (ARCHIVE_OK);
	if (z->c.ctx == NULL && (rc = b200_ctx_acquire(&z->c.ctx)) != B2I_OK) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
		    "No usable B200 device (b2i_ctx_create: %d); this build has no CPU inflate", rc);
		return (ARCHIVE_FATAL);
	}
	z->descs = calloc(z->ix.n ? z->ix.n : 1, sizeof(*z->descs));
	z->desc_of = calloc(z->ix.n ? z->ix.n : 1, sizeof(*z->desc_of));
	if (z->descs == NULL || z->desc_of == NULL) {
		archive_set_error(&a->archive, ENOMEM, "No memory for ZIP decompression");
		return (ARCHIVE_FATAL);
	}
	for (i = 0; i < z->ix.n; i++) {
		const b2i_zip_entry *e = &z->ix.entries[i];
		b2i_stream_desc *d;

		z->desc_of[i] = SIZE_MAX;
		if (!decodable(e))
			continue;
		d = &z->descs[n];
		d->in_off = e->data_offset;
		d->in_len = e->compressed_size;
		d->expect_out = e->uncompressed_size;
		d->expect_crc = e->crc32;
		d->method = (uint8_t)e->method;
		d->flags = z->c.ignore_crc32 ? B2I_F_NO_CRC : 0;
		d->out_off = out;
		if (e->method == 0) {
			d->flags |= B2I_F_NO_COPY;          /* served from the source bytes: zero copy */
		} else {
			/* deflate cannot expand by more than 1032:1 (a 258-byte match in 2 bits): a
			 * directory that claims more gets the honest bound, not gigabytes of pinned
			 * memory; the size check at the end of the entry still sees expect_out */
			uint64_t bound = e->compressed_size * 1032u + 65536u;
			d->out_cap = e->uncompressed_size < bound ? e->uncompressed_size : bound;
			out = (out + (size_t)d->out_cap + 15) & ~(size_t)15;
		}
		if (d->in_off < in_lo) in_lo = d->in_off;
		if (d->in_off + d->in_len > in_hi) in_hi = d->in_off + d->in_len;
		z->desc_of[i] = n++;
	}
	z->ndesc = n;
	z->out_bytes = out;
	z->cur_desc = SIZE_MAX;
	z->decoded = 1;
	if (n == 0)
		return (ARCHIVE_OK);

	if (!z->file_mode && (in_hi - in_lo) + (uint64_t)out <= ZB_BATCH_MAX && getenv("B2I_ZIP_PIPE") == NULL) {
		/* ---- one call: the whole archive is one small batch ---- */
		z->res = calloc(n, sizeof(*z->res));
		z->retry = calloc(n, sizeof(*z->retry));
		z->out = b200_buf_acquire(out + 16, &z->out_cap);
		if (z->res == NULL || z->retry == NULL || z->out == NULL) {
			archive_set_error(&a->archive, ENOMEM, "No memory for ZIP decompression");
			return (ARCHIVE_FATAL);
		}
		if ((rc = b2i_decode_host(z->c.ctx, z->image, z->image_len, z->descs, n, z->out, out,
		    z->res)) != B2I_OK) {
			z->c.ctx_bad = (rc == B2I_E_CUDA);       /* a rejected batch does not poison the context */
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "B200 decode failed (%d): %s", rc,
			    b2i_last_error(z->c.ctx));
			return (ARCHIVE_FATAL);
		}
		for (i = 0; i < n; i++)
			if (z->res[i].status == B2I_S_OUT_OVERFLOW)
				z->retry[i] = retry_with_room(z, z->c.ctx, z->image, z->image_len, &z->descs[i], &z->res[i]);
		return (ARCHIVE_OK);
	}

	/* ---- streaming: windows through the pinned ring, on as many GPUs as it is worth ---- */
	z->nctx = gpus_for((uint64_t)out);
	z->ctxs[0] = z->c.ctx;              /* the handle's own context drives device 0 */
	for (g = 1; g < z->nctx; g++) {
		if (b200_ctx_acquire_dev(g, &z->ctxs[g]) != B2I_OK) {
			z->nctx = g;            /* fewer devices than advertised: go on with what we have */
			break;
		}
	}
	rc = b2i_pipe_open(z->ctxs, z->nctx, z->file_mode ? NULL : z->image, z->file_mode ? 0 : z->image_len,
	    z->file_mode ? zf_fill : NULL, z, z->descs, n, NULL, &z->pipe);
	if (rc != B2I_OK) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "B200 pipeline could not be started (%d)", rc);
		return (ARCHIVE_FATAL);
	}
	return (ARCHIVE_OK);
}

/* the current entry's result and bytes; streaming mode waits for its window here */
static const b2i_stream_result *
entry_result(struct archive_read *a, struct zip_b200 *z, size_t ei)
{
	const size_t di = z->desc_of[ei];
	const void *o = NULL, *in = NULL;
	int rc;

	if (z->pipe == NULL)
		return (&z->res[di]);
	if (z->cur_desc == di)
		return (&z->cur_res);
	b200_buf_release_tagged(z->cur_retry);
	z->cur_retry = NULL;
	if ((rc = b2i_pipe_get(z->pipe, di, &o, &in, &z->cur_res)) != B2I_OK) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "B200 decode failed (%d): %s", rc,
		    b2i_pipe_error(z->pipe));
		return (NULL);
	}
	z->cur_desc = di;
	z->cur_out = o;
	z->cur_in = in;
	if (z->cur_res.status == B2I_S_OUT_OVERFLOW) {
		/* decoded again alone, from the staged input, on a context of its own (the
		 * pipeline's workers own the handle's) */
		b2i_stream_desc d = z->descs[di];
		b2i_ctx *tmp = NULL;
		d.in_off = 0;
		if (b200_ctx_acquire_dev(0, &tmp) == B2I_OK) {
			z->cur_retry = retry_with_room(z, tmp, in, (size_t)d.in_len, &d, &z->cur_res);
			b200_ctx_release(tmp, 1);
		}
	}
	return (&z->cur_res);
}

static const unsigned char *
entry_bytes(struct zip_b200 *z, size_t ei)
{
	size_t di = z->desc_of[ei];
	if (z->pipe != NULL) {
		if (z->ix.entries[ei].method == 0)
			return (z->cur_in);
		return (z->cur_retry ? z->cur_retry : z->cur_out);
	}
	if (z->ix.entries[ei].method == 0)
		return (z->image + z->descs[di].in_off);
	return (z->retry[di] ? z->retry[di] : z->out + z->descs[di].out_off);
}

/* ---- read_header (zip.c:4268-4343 + 905-1287) ----------------------------------- */
static int
zip_b200_read_header(struct archive_read *a, struct archive_entry *entry)
{
	struct zip_b200 *z = (struct zip_b200 *)a->format->data;
	const b2i_zip_entry *e;
	struct zb_meta m;
	const char *name;
	int ret = ARCHIVE_OK, r;
	unsigned mode;

	if (z->c.has_encrypted_entries == ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW)
		z->c.has_encrypted_entries = 0;
	/* the central directory is read on the first call; any encrypted record in it
	 * makes the answer 1 from then on (zip.c:3963-3967) */
	if (z->ix.has_encrypted_entries)
		z->c.has_encrypted_entries = 1;
	__archive_read_reset_passphrase(a);
	a->archive.archive_format = ARCHIVE_FORMAT_ZIP;
	if (a->archive.archive_format_name == NULL)
		a->archive.archive_format_name = "ZIP";

	if (z->image == NULL && !z->file_mode) {
		/* bidding is over: take the image once (zero copy for memory sources) */
		if (__archive_read_seek(a, 0, SEEK_SET) < 0 ||
		    (z->image = __archive_read_ahead(a, z->image_len, NULL)) == NULL) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file header");
			return (ARCHIVE_FATAL);
		}
	}
	if (z->next >= z->ix.n)
		return (ARCHIVE_EOF);
	z->cur = z->next++;
	e = &z->ix.entries[z->cur];
	/* streaming mode: everything in front of this entry is done with */
	if (z->pipe != NULL && z->desc_of != NULL) {
		size_t k = z->cur;
		while (k < z->ix.n && z->desc_of[k] == SIZE_MAX)
			k++;
		b2i_pipe_release(z->pipe, k < z->ix.n ? z->desc_of[k] : z->ndesc);
	}
	z->delivered = 0;
	z->end_of_entry = 0;

	if (e->warn & B2I_ZW_TRUNCATED && e->name_len == 0 && e->data_offset == 0) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file header");
		return (ARCHIVE_FATAL);
	}
	if (e->warn & B2I_ZW_BAD_LOCAL_HEADER) {
		archive_set_error(&a->archive, -1, "Damaged Zip archive");
		return (ARCHIVE_FATAL);
	}
	if (e->zip_flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED)) {
		z->c.has_encrypted_entries = 1;
		archive_entry_set_is_data_encrypted(entry, 1);
		if ((e->zip_flags & ZIP_CD_ENCRYPTED) && (e->zip_flags & ZIP_ENCRYPTED) &&
		    (e->zip_flags & ZIP_STRONG_ENCRYPTED)) {
			archive_entry_set_is_metadata_encrypted(entry, 1);
			return (ARCHIVE_FATAL);
		}
	}

	/* pathname, the Unicode path of the local extra field, type fix-ups (zip.c:976-1103) */
	memset(&m, 0, sizeof(m));
	m.zip_flags = e->zip_flags;
	m.system = e->system;
	m.mode = e->mode;
	m.mtime = e->mtime;
	m.atime = e->atime;
	m.ctime = e->ctime;
	m.uid = e->uid;
	m.gid = e->gid;
	name = z->ix.names + e->name_offset;
	r = zb_set_pathname(a, &z->c, entry, name, e->name_len, e->zip_flags);
	if (r == ARCHIVE_FATAL)
		return (r);
	if (r != ARCHIVE_OK)
		ret = r;
	if (e->local_extra_len != 0 && e->local_extra_offset + e->local_extra_len <= z->image_len) {
		struct zb_meta scratch = m;      /* everything but 0x7075 was applied by the index */
		const unsigned char *xp = z->file_mode ?
		    zf_fetch(z, e->local_extra_offset, e->local_extra_len) : z->image + e->local_extra_offset;
		if (xp != NULL)
			(void)zb_process_extra(a, &z->c, entry, xp, e->local_extra_len, &scratch, NULL);
	}
	zb_fix_path_and_mode(entry, &m);
	mode = m.mode;
	if (e->warn & B2I_ZW_CRC_INCONSISTENT && !z->c.ignore_crc32) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Inconsistent CRC32 values");
		ret = ARCHIVE_WARN;
	}
	if (e->warn & B2I_ZW_CSIZE_INCONSISTENT) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Inconsistent compressed size: central directory and local header disagree");
		ret = ARCHIVE_WARN;
	}
	if (e->warn & B2I_ZW_USIZE_INCONSISTENT) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Inconsistent uncompressed size: central directory and local header disagree");
		ret = ARCHIVE_WARN;
	}
	zb_populate(entry, &m);

	if ((mode & AE_IFMT) == AE_IFLNK) {
		/* the link target is the entry body (zip.c:1160-1265) */
		size_t len = (size_t)e->compressed_size;
		const unsigned char *p = NULL;

		if (e->compressed_size > 64 * 1024) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "Zip file with oversized link entry");
			return (ARCHIVE_FATAL);
		}
		archive_entry_set_size(entry, 0);
		if (e->compressed_size >= 1) {
			if (e->method != 0 && e->method != 8) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
				    "Unsupported ZIP compression method during decompression of link entry (%d: %s)",
				    e->method, zb_compression_name(e->method));
				return (ARCHIVE_FAILED);
			}
			if (zip_b200_prepare(a, z) != ARCHIVE_OK)
				return (ARCHIVE_FATAL);
			if (z->desc_of[z->cur] != SIZE_MAX) {
				const b2i_stream_result *r = entry_result(a, z, z->cur);
				if (r == NULL)
					return (ARCHIVE_FATAL);
				if (r->status != B2I_S_OK) {
					archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
					    "Unsupported ZIP compression method during decompression of link entry (%d: %s)",
					    e->method, zb_compression_name(e->method));
					return (ARCHIVE_FAILED);
				}
				p = entry_bytes(z, z->cur);
				len = (size_t)r->out_bytes;
			}
		}
		if (p == NULL && len > 0) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "Truncated Zip file");
			return (ARCHIVE_FATAL);
		}
		r = zb_set_symlink(a, &z->c, entry, p, len, e->zip_flags);
		if (r == ARCHIVE_FATAL)
			return (r);
		if (r != ARCHIVE_OK)
			ret = r;
		z->end_of_entry = 1;
	} else {
		archive_entry_set_size(entry, (int64_t)e->uncompressed_size);
		if (e->compressed_size < 1)
			z->end_of_entry = 1;                 /* no body: EOF immediately (zip.c:1275-1277) */
	}

	archive_string_empty(&z->c.format_name);
	archive_string_sprintf(&z->c.format_name, "ZIP %d.%d (%s)", e->version / 10, e->version % 10,
	    zb_compression_name(e->method));
	a->archive.archive_format_name = z->c.format_name.s;
	return (ret);
}

/* ---- read_data (zip.c:3071-3198, 2535-2690, 1592-1706) ----------------------------- */
static int
zip_b200_read_data(struct archive_read *a, const void **buff, size_t *size, int64_t *offset)
{
	struct zip_b200 *z = (struct zip_b200 *)a->format->data;
	const b2i_zip_entry *e = &z->ix.entries[z->cur];
	const b2i_stream_result *r;
	int64_t left;
	size_t n;
	int last;

	if (z->c.has_encrypted_entries == ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW)
		z->c.has_encrypted_entries = 0;
	*offset = z->delivered;
	*size = 0;
	*buff = NULL;
	if (z->end_of_entry)
		return (ARCHIVE_EOF);
	if (AE_IFREG != (e->mode & AE_IFMT))
		return (ARCHIVE_EOF);
	if (e->zip_flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED)) {
		z->c.has_encrypted_entries = 1;
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Encrypted ZIP entries are not supported by this build");
		return (ARCHIVE_FAILED);
	}
	if (e->method != 0 && e->method != 8) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Unsupported ZIP compression method (%d: %s)", e->method, zb_compression_name(e->method));
		return (ARCHIVE_FAILED);
	}
	if (e->warn & B2I_ZW_TRUNCATED) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file body");
		return (ARCHIVE_FATAL);
	}
	if (zip_b200_prepare(a, z) != ARCHIVE_OK)
		return (ARCHIVE_FATAL);
	if ((r = entry_result(a, z, z->cur)) == NULL)
		return (ARCHIVE_FATAL);
	left = (int64_t)r->out_bytes - z->delivered;

	if (e->method == 0) {
		/* zip_read_data_none hands out whatever the read core holds: the whole
		 * remaining body here; the end is noticed on the following call */
		if (left > 0) {
			*buff = entry_bytes(z, z->cur) + z->delivered;
			*size = (size_t)left;
			z->delivered += left;
			return (ARCHIVE_OK);
		}
		last = 1;
		n = 0;
	} else {
		n = left > ZIP_BLOCK ? ZIP_BLOCK : (size_t)left;
		last = ((int64_t)n == left);
		if (r->status == B2I_S_BUF_ERROR) {
			/* zlib returns what it could produce with Z_OK and reports Z_BUF_ERROR on
			 * the call that finds no input left (zip.c:2570-2657) */
			if (left == 0) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
				    "ZIP decompression failed (%d)", (int)r->status);
				return (ARCHIVE_FATAL);
			}
			last = 0;
		} else if (r->status != B2I_S_OK) {
			if (n < ZIP_BLOCK) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
				    "ZIP decompression failed (%d)", (int)r->status);
				return (ARCHIVE_FATAL);
			}
			last = 0;
		}
		*buff = entry_bytes(z, z->cur) + z->delivered;
		*size = n;
		z->delivered += (int64_t)n;
		if (!last)
			return (ARCHIVE_OK);
	}
	/* end of entry: CRC, compressed size, uncompressed size - in this order (zip.c:3164-3194) */
	z->end_of_entry = 1;
	if ((r->flags & B2I_R_CRC_MISMATCH) && !z->c.ignore_crc32) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "ZIP bad CRC: 0x%lx should be 0x%lx",
		    (unsigned long)r->crc, (unsigned long)e->crc32);
		*size = 0; *buff = NULL;
		return (ARCHIVE_FAILED);
	}
	if (r->flags & B2I_R_IN_MISMATCH) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
		    "ZIP compressed data is wrong size (read %jd, expected %jd)",
		    (intmax_t)r->in_bytes, (intmax_t)e->compressed_size);
		*size = 0; *buff = NULL;
		return (ARCHIVE_FAILED);
	}
	if (r->flags & B2I_R_OUT_MISMATCH) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
		    "ZIP uncompressed data is wrong size (read %jd, expected %jd)\n",
		    (intmax_t)r->out_bytes, (intmax_t)e->uncompressed_size);
		*size = 0; *buff = NULL;
		return (ARCHIVE_FAILED);
	}
	return (ARCHIVE_OK);
}

static int
zip_b200_read_data_skip(struct archive_read *a)
{
	(void)a;                               /* everything is addressed by offset (zip.c:4349-4357) */
	return (ARCHIVE_OK);
}

static int
zip_b200_cleanup(struct archive_read *a)
{
	struct zip_b200 *z = (struct zip_b200 *)a->format->data;
	size_t i;

	int g;

	if (z->pipe != NULL)
		b2i_pipe_close(z->pipe);            /* joins the workers before the contexts go back */
	for (g = 1; g < z->nctx; g++)
		b200_ctx_release(z->ctxs[g], 1);
	b200_buf_release_tagged(z->cur_retry);
	if (z->retry != NULL)
		for (i = 0; i < z->ndesc; i++)
			b200_buf_release_tagged(z->retry[i]);
	free(z->retry);
	free(z->descs);
	free(z->res);
	free(z->desc_of);
	b200_buf_release(z->out, z->out_cap);
	if (z->have_index)
		b2i_zip_index_free(&z->ix);
	b200_ctx_release(z->c.ctx, !z->c.ctx_bad);
	archive_string_free(&z->c.format_name);
	free(z);
	a->format->data = NULL;
	return (ARCHIVE_OK);
}

static int
zip_b200_capabilities(struct archive_read *a)
{
	(void)a;
	return (ARCHIVE_READ_FORMAT_CAPS_ENCRYPT_DATA | ARCHIVE_READ_FORMAT_CAPS_ENCRYPT_METADATA);
}

static int
zip_b200_has_encrypted_entries(struct archive_read *a)
{
	if (a && a->format && a->format->data)
		return (((struct zip_b200 *)a->format->data)->c.has_encrypted_entries);
	return (ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW);
}

int
archive_read_support_format_zip_seekable(struct archive *_a)
{
	struct archive_read *a = (struct archive_read *)_a;
	struct zip_b200 *z;
	int r;

	archive_check_magic(_a, ARCHIVE_READ_MAGIC, ARCHIVE_STATE_NEW,
	    "archive_read_support_format_zip_seekable");
	z = calloc(1, sizeof(*z));
	if (z == NULL) {
		archive_set_error(&a->archive, ENOMEM, "Can't allocate zip data");
		return (ARCHIVE_FATAL);
	}
	z->c.has_encrypted_entries = ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW;
	r = __archive_read_register_format(a, z, "zip", zip_b200_bid, zip_b200_options,
	    zip_b200_read_header, zip_b200_read_data, zip_b200_read_data_skip, NULL, zip_b200_cleanup,
	    zip_b200_capabilities, zip_b200_has_encrypted_entries);
	if (r != ARCHIVE_OK)
		free(z);
	return (ARCHIVE_OK);
}

/* archive_read_support_format_zip_streamable: archive_read_support_format_zip_stream_b200.c */
int
archive_read_support_format_zip(struct archive *a)
{
	int r = archive_read_support_format_zip_streamable(a);
	if (r != ARCHIVE_OK)
		return (r);
	return (archive_read_support_format_zip_seekable(a));
}
