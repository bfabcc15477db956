/*
 * archive_read_support_format_zip_stream_b200.c — libarchive ZIP format module,
 * STREAMING reader (no central directory), bodies decoded on a B200 through
 * include/b200inflate.h instead of zlib.
 *
 * Defines archive_read_support_format_zip_streamable (zip.c:3567-3610) and
 * registers with the unmodified read core like the seekable module next to it.
 * It is what the core falls back to when the input cannot seek or has no usable
 * end-of-central-directory record (bid 29 against the seekable reader's 32,
 * zip.c:3345-3380).
 *
 * Design (not the reference's): entries are discovered one local header at a
 * time, so there is no archive-wide batch here.  A deflate body is handed to
 * the device WHOLE the first time its data is asked for: with known sizes the
 * compressed span is requested from the read core in one piece; with
 * length-at-end entries (sizes in a trailing data descriptor) the decoder is
 * given everything the core can provide and reports where the stream ended,
 * which is exactly how the reference finds the descriptor (by decoding,
 * zip.c:3519-3532).  The decoded entry is then served in the reference's
 * 256 KiB blocks.  Stored bodies are passed through from the read-ahead buffer
 * (zero copy, zip.c:1592-1706) with their CRC-32 accumulated on the device.
 * No CPU inflate: other methods and encrypted entries are refused.
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
#include "archive_time_private.h"

#include "b200inflate.h"
#include "b200_ctx_pool.h"
#include "zip_b200_local.h"

struct zs_b200 {
	struct zb_common    c;             /* must be first */
	struct zb_meta      m;             /* the current entry */
	int64_t             unconsumed;    /* bytes handed out that the core still holds  */
	int64_t             remaining;     /* entry_bytes_remaining (zip.c:1272)          */
	int64_t             cread, uread;  /* entry_{compressed,uncompressed}_bytes_read  */
	uint32_t            computed_crc;
	int                 end_of_entry;
	/* deflate: the whole entry, decoded once */
	int                 decoded;
	unsigned char      *out;           /* pinned */
	size_t              out_cap;
	b2i_stream_result   res;
	int64_t             delivered;     /* bytes of `out` already returned             */
	const unsigned char *cur;          /* the decoded bytes being served: `out` or a slot of the batch */
	/* look-ahead batch: deflate entries with known sizes that follow the current one in
	 * the read-ahead buffer, decoded in the same device pass (zs_batch_decode) */
	struct {
		size_t             n, next;        /* entries, first one not taken yet */
		int64_t           *pos;            /* absolute input position of each entry's data */
		b2i_stream_desc   *d;
		b2i_stream_result *r;
		unsigned char     *out;
		size_t             out_cap, cap_n;
	} b;
};

#define ZS_BATCH_MAX_ENTRIES 4096
#define ZS_BATCH_MAX_IN      ((size_t)64 << 20)
#define ZS_BATCH_MAX_OUT     ((size_t)256 << 20)

/* ---- bid (zip.c:3345-3380) ------------------------------------------------------ */
static int
zs_bid(struct archive_read *a, int best_bid)
{
	const char *p;

	(void)best_bid;
	if ((p = __archive_read_ahead(a, 4, NULL)) == NULL)
		return (-1);
	if (p[0] == 'P' && p[1] == 'K') {
		if ((p[2] == '\001' && p[3] == '\002') || (p[2] == '\003' && p[3] == '\004') ||
		    (p[2] == '\005' && p[3] == '\006') || (p[2] == '\006' && p[3] == '\006') ||
		    (p[2] == '\007' && p[3] == '\010') || (p[2] == '0' && p[3] == '0'))
			return (29);
	}
	return (0);
}

static int
zs_options(struct archive_read *a, const char *key, const char *val)
{
	return (zb_options(a, (struct zb_common *)a->format->data, key, val));
}

static int
zs_need_ctx(struct archive_read *a, struct zs_b200 *z)
{
	int rc;

	if (z->c.ctx == NULL && (rc = b200_ctx_acquire(&z->c.ctx)) != B2I_OK) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
		    "No usable B200 device (b2i_ctx_create: %d); this build has no CPU inflate", rc);
		return (ARCHIVE_FATAL);
	}
	return (ARCHIVE_OK);
}

/*
 * Decode the deflate body that starts at the read position.  `known` > 0: the
 * compressed size; 0: unknown (length-at-end), the decoder finds the end.
 * Leaves the result in z->res / z->out and consumes what the stream used.
 */
static int
zs_decode_entry(struct archive_read *a, struct zs_b200 *z, int64_t known, int64_t expect_out)
{
	size_t want = known > 0 ? (size_t)known : 64 * 1024, in_len;
	size_t cap = expect_out > 0 ? (size_t)expect_out : 0;
	const void *p;
	ssize_t avail;
	int final = 0, rc;

	if (zs_need_ctx(a, z) != ARCHIVE_OK)
		return (ARCHIVE_FATAL);
	for (;;) {
		b2i_stream_desc d;

		p = __archive_read_ahead(a, want, &avail);
		if (p == NULL) {
			/* the input ends first: take what there is (zlib would be fed the same
			 * bytes and then report Z_BUF_ERROR, zip.c:2570-2657) */
			if (avail < 0) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file body");
				return (ARCHIVE_FATAL);
			}
			final = 1;
			in_len = (size_t)avail;
			p = in_len ? __archive_read_ahead(a, in_len, &avail) : "";
			if (p == NULL)
				in_len = 0, p = "";
		} else
			in_len = known > 0 ? (size_t)known : (size_t)avail;
		/* deflate cannot expand by more than 1032:1: a header that claims more gets the
		 * honest bound, not gigabytes of pinned memory (the size check still sees expect_out) */
		if (expect_out > 0 && (uint64_t)expect_out > (uint64_t)in_len * 1032u + 65536u)
			cap = in_len * 1032u + 65536u;
		if (cap < 4 * in_len && expect_out <= 0)
			cap = 4 * in_len;
		if (cap < 65536 && expect_out <= 0)
			cap = 65536;
		for (;;) {
			if (z->out == NULL || z->out_cap < cap + 16) {
				b200_buf_release(z->out, z->out_cap);
				if ((z->out = b200_buf_acquire(cap + 16, &z->out_cap)) == NULL) {
					z->out_cap = 0;
					archive_set_error(&a->archive, ENOMEM, "No memory for ZIP decompression");
					return (ARCHIVE_FATAL);
				}
			}
			memset(&d, 0, sizeof(d));
			d.in_off = 0;
			d.in_len = in_len;
			d.out_off = 0;
			d.out_cap = cap;
			d.expect_out = expect_out > 0 ? (uint64_t)expect_out : 0;
			d.expect_crc = z->m.crc32;
			d.method = 8;
			d.flags = z->c.ignore_crc32 ? B2I_F_NO_CRC : 0;
			memset(&z->res, 0, sizeof(z->res));
			if (in_len == 0) {
				z->res.status = B2I_S_BUF_ERROR;
				break;
			}
			if ((rc = b2i_decode_host(z->c.ctx, p, in_len, &d, 1, z->out, cap, &z->res)) != B2I_OK) {
				z->c.ctx_bad = (rc == B2I_E_CUDA);
				archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "B200 decode failed (%d): %s", rc,
				    b2i_last_error(z->c.ctx));
				return (ARCHIVE_FATAL);
			}
			/* output outgrew the announced size: decode again with room, so that the
			 * "wrong size" message can quote the true count */
			if (z->res.status != B2I_S_OUT_OVERFLOW || cap >= in_len * 1032u + 65536u)
				break;
			cap = cap < 4096 ? 16384 : cap * 4;
			if (cap > in_len * 1032u + 65536u)
				cap = in_len * 1032u + 65536u;
		}
		if (z->res.status == B2I_S_BUF_ERROR && !final && known <= 0) {
			want = in_len * 2 > want ? in_len * 2 : want * 2;   /* the stream is longer: get more input */
			continue;
		}
		break;
	}
	/* a failed stream may report a position past its end (the bit reader runs ahead);
	 * never consume more than it was given */
	if (z->res.in_bytes > in_len)
		z->res.in_bytes = in_len;
	__archive_read_consume(a, (int64_t)z->res.in_bytes);
	z->remaining -= (int64_t)z->res.in_bytes;
	z->cread = (int64_t)z->res.in_bytes;
	z->decoded = 1;
	z->delivered = 0;
	z->cur = z->out;
	return (ARCHIVE_OK);
}

/*
 * Entries of a streamed archive are discovered one header at a time, but when
 * the sizes are in the local headers (no length-at-end flag) the headers that
 * follow can be read ahead without decoding anything.  The first data request
 * of a deflate entry therefore walks the local headers behind it as far as the
 * read core can provide bytes (bounded), makes one descriptor per plain deflate
 * entry and decodes them all in ONE device pass; the entries that follow find
 * their result here (matched by absolute input position, sizes and CRC as their
 * own header states them) instead of paying a device round trip each.
 * Returns 1 when the current entry was decoded as slot 0 of a new batch.
 */
static int
zs_batch_decode(struct archive_read *a, struct zs_b200 *z)
{
	const int64_t here = archive_filter_bytes(&a->archive, 0);
	const unsigned char *p;
	ssize_t avail;
	size_t o, n = 0, out = 0, total;
	int rc;

	if (z->remaining <= 0 || (size_t)z->remaining > ZS_BATCH_MAX_IN ||
	    z->m.uncompressed_size > ZS_BATCH_MAX_OUT || z->m.uncompressed_size >= 0xffffffffu)
		return (0);
	if (z->b.cap_n == 0) {
		z->b.pos = calloc(ZS_BATCH_MAX_ENTRIES, sizeof(*z->b.pos));
		z->b.d = calloc(ZS_BATCH_MAX_ENTRIES, sizeof(*z->b.d));
		z->b.r = calloc(ZS_BATCH_MAX_ENTRIES, sizeof(*z->b.r));
		if (z->b.pos == NULL || z->b.d == NULL || z->b.r == NULL)
			return (0);
		z->b.cap_n = ZS_BATCH_MAX_ENTRIES;
	}
	z->b.n = z->b.next = 0;
	/* slot 0: the current entry */
	memset(&z->b.d[0], 0, sizeof(z->b.d[0]));
	z->b.d[0].in_off = 0;
	z->b.d[0].in_len = (uint64_t)z->remaining;
	z->b.d[0].out_cap = z->m.uncompressed_size;
	z->b.d[0].expect_out = z->m.uncompressed_size;
	z->b.d[0].expect_crc = z->m.crc32;
	z->b.d[0].method = 8;
	z->b.d[0].flags = z->c.ignore_crc32 ? B2I_F_NO_CRC : 0;
	z->b.d[0].out_off = 0;
	z->b.pos[0] = here;
	out = ((size_t)z->m.uncompressed_size + 15) & ~(size_t)15;
	n = 1;
	o = (size_t)z->remaining;
	total = o;
	/* the local headers that follow */
	while (n < ZS_BATCH_MAX_ENTRIES) {
		uint32_t flags, method, csize, usize, nl, xl, hcrc;

		if ((p = __archive_read_ahead(a, o + 30, &avail)) == NULL)
			break;
		p += o;
		if (memcmp(p, "PK\003\004", 4) != 0)
			break;
		flags = zb_le16(p + 6);
		method = zb_le16(p + 8);
		csize = zb_le32(p + 18);
		usize = zb_le32(p + 22);
		nl = zb_le16(p + 26);
		xl = zb_le16(p + 28);
		hcrc = zb_le32(p + 14);            /* the next look may move the buffer: p is not used below */
		if ((flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED | ZIP_LENGTH_AT_END)) ||
		    csize == 0xffffffffu || usize == 0xffffffffu || (method != 0 && method != 8))
			break;
		if (o + 30 + nl + xl + csize > ZS_BATCH_MAX_IN || out + usize > ZS_BATCH_MAX_OUT)
			break;
		if (__archive_read_ahead(a, o + 30 + nl + xl + csize, &avail) == NULL)
			break;                                  /* the body is not all there */
		if (method == 8 && csize > 0) {
			b2i_stream_desc *d = &z->b.d[n];
			memset(d, 0, sizeof(*d));
			d->in_off = o + 30 + nl + xl;
			d->in_len = csize;
			d->out_cap = usize;
			d->expect_out = usize;
			d->expect_crc = hcrc;
			d->method = 8;
			d->flags = z->c.ignore_crc32 ? B2I_F_NO_CRC : 0;
			d->out_off = out;
			z->b.pos[n] = here + (int64_t)d->in_off;
			out += ((size_t)usize + 15) & ~(size_t)15;
			n++;
		}
		o += 30 + nl + xl + csize;
		total = o;
	}
	if (n < 2)
		return (0);                                     /* nothing to gain: the plain path */
	if ((p = __archive_read_ahead(a, total, &avail)) == NULL)
		return (0);
	if (zs_need_ctx(a, z) != ARCHIVE_OK)
		return (0);
	if (z->b.out == NULL || z->b.out_cap < out + 16) {
		b200_buf_release(z->b.out, z->b.out_cap);
		if ((z->b.out = b200_buf_acquire(out + 16, &z->b.out_cap)) == NULL) {
			z->b.out_cap = 0;
			return (0);
		}
	}
	if ((rc = b2i_decode_host(z->c.ctx, p, total, z->b.d, n, z->b.out, out, z->b.r)) != B2I_OK) {
		z->c.ctx_bad = (rc == B2I_E_CUDA);
		return (0);                                     /* the plain path reports the failure */
	}
	z->b.n = n;
	z->b.next = 0;
	return (1);
}

/* the current entry's result from the batch, if it is there: 1 = taken */
static int
zs_batch_take(struct archive_read *a, struct zs_b200 *z)
{
	const int64_t here = archive_filter_bytes(&a->archive, 0);
	size_t k;

	for (k = z->b.next; k < z->b.n; k++) {
		const b2i_stream_desc *d = &z->b.d[k];
		if (z->b.pos[k] < here)
			continue;
		if (z->b.pos[k] != here)
			break;
		z->b.next = k + 1;
		if ((int64_t)d->in_len != z->remaining || d->expect_out != z->m.uncompressed_size ||
		    d->expect_crc != z->m.crc32 || ((d->flags & B2I_F_NO_CRC) != 0) != (z->c.ignore_crc32 != 0) ||
		    z->b.r[k].status == B2I_S_OUT_OVERFLOW)
			return (0);             /* not what this entry's header says (or needs room): decode it alone */
		z->res = z->b.r[k];
		if (z->res.in_bytes > d->in_len)
			z->res.in_bytes = d->in_len;
		z->cur = z->b.out + d->out_off;
		__archive_read_consume(a, (int64_t)z->res.in_bytes);
		z->remaining -= (int64_t)z->res.in_bytes;
		z->cread = (int64_t)z->res.in_bytes;
		z->decoded = 1;
		z->delivered = 0;
		return (1);
	}
	z->b.next = k;
	return (0);
}

/*
 * The record that follows a length-at-end body: optional PK\7\8, CRC, sizes in
 * 32 or 64 bits.  All four layouts are tried, longest first, against what was
 * actually read; a full match is consumed, otherwise plausible values are kept
 * for the error report and nothing is consumed (zip.c:1394-1565).
 */
static void
zs_end_of_file_marker(struct archive_read *a, struct zs_b200 *z)
{
	const uint32_t PK78 = 0x08074B50u;
	const unsigned char *p;
	uint64_t cact = (uint64_t)z->cread, uact = (uint64_t)z->uread;
	uint32_t crc = z->computed_crc;
	int ign = z->c.ignore_crc32;

	if ((z->m.zip_flags & ZIP_LENGTH_AT_END) == 0)
		return;
	if ((p = __archive_read_ahead(a, 24, NULL)) == NULL)
		return;
#define TAKE(n) do { if (!ign) z->m.crc32 = crc; z->m.compressed_size = cact; \
		z->m.uncompressed_size = uact; z->unconsumed += (n); return; } while (0)
	if (zb_le32(p) == PK78 && (zb_le32(p + 4) == crc || ign) && zb_le64(p + 8) == cact && zb_le64(p + 16) == uact)
		TAKE(24);
	if ((zb_le32(p) == crc || ign) && zb_le64(p + 4) == cact && zb_le64(p + 12) == uact)
		TAKE(20);
	if (zb_le32(p) == PK78 && (zb_le32(p + 4) == crc || ign) && zb_le32(p + 8) == cact && zb_le32(p + 12) == uact)
		TAKE(16);
	if ((zb_le32(p) == crc || ign) && zb_le32(p + 4) == cact && zb_le32(p + 8) == uact)
		TAKE(12);
#undef TAKE
	if (zb_le32(p) == PK78)
		p += 4;
	z->m.crc32 = zb_le32(p);
	p += 4;
	if (zb_le32(p) == cact && zb_le32(p + 4) == uact) {
		z->m.compressed_size = zb_le32(p);
		z->m.uncompressed_size = zb_le32(p + 4);
	} else if (zb_le64(p) == cact || zb_le64(p + 8) == uact) {
		z->m.compressed_size = zb_le64(p);
		z->m.uncompressed_size = zb_le64(p + 8);
	} else {
		z->m.compressed_size = zb_le32(p);
		z->m.uncompressed_size = zb_le32(p + 4);
	}
}

/* ---- local file header (zip.c:905-1287) ---------------------------------------- */
static int
zs_local_header(struct archive_read *a, struct archive_entry *entry, struct zs_b200 *z)
{
	struct zb_meta *m = &z->m;
	const unsigned char *p;
	const void *h;
	size_t name_len, extra_len;
	int ret = ARCHIVE_OK, r;

	memset(m, 0, sizeof(*m));
	z->end_of_entry = 0;
	z->decoded = 0;
	z->delivered = 0;
	z->cread = z->uread = 0;
	z->computed_crc = 0;

	if ((p = __archive_read_ahead(a, 30, NULL)) == NULL) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file header");
		return (ARCHIVE_FATAL);
	}
	if (memcmp(p, "PK\003\004", 4) != 0) {
		archive_set_error(&a->archive, -1, "Damaged Zip archive");
		return (ARCHIVE_FATAL);
	}
	m->version = p[4];
	m->system = p[5];
	m->zip_flags = zb_le16(p + 6);
	if (m->zip_flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED)) {
		z->c.has_encrypted_entries = 1;
		archive_entry_set_is_data_encrypted(entry, 1);
		if ((m->zip_flags & ZIP_CD_ENCRYPTED) && (m->zip_flags & ZIP_ENCRYPTED) &&
		    (m->zip_flags & ZIP_STRONG_ENCRYPTED)) {
			archive_entry_set_is_metadata_encrypted(entry, 1);
			return (ARCHIVE_FATAL);
		}
	}
	m->method = (uint16_t)(zb_le16(p + 8) & 0xff);          /* the reference keeps a char */
	m->mtime = dos_to_unix(zb_le32(p + 10));
	m->crc32 = zb_le32(p + 14);
	m->compressed_size = zb_le32(p + 18);
	m->uncompressed_size = zb_le32(p + 22);
	name_len = zb_le16(p + 26);
	extra_len = zb_le16(p + 28);
	__archive_read_consume(a, 30);

	if ((h = __archive_read_ahead(a, name_len, NULL)) == NULL) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file header");
		return (ARCHIVE_FATAL);
	}
	r = zb_set_pathname(a, &z->c, entry, h, name_len, m->zip_flags);
	if (r == ARCHIVE_FATAL)
		return (r);
	if (r != ARCHIVE_OK)
		ret = r;
	__archive_read_consume(a, name_len);

	if ((h = __archive_read_ahead(a, extra_len, NULL)) == NULL) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file header");
		return (ARCHIVE_FATAL);
	}
	if (zb_process_extra(a, &z->c, entry, h, extra_len, m, NULL) != ARCHIVE_OK)
		return (ARCHIVE_FATAL);
	__archive_read_consume(a, extra_len);

	zb_fix_path_and_mode(entry, m);
	zb_populate(entry, m);

	if ((m->mode & AE_IFMT) == AE_IFLNK) {
		/* the link target is the body (zip.c:1160-1265) */
		size_t len = (size_t)m->compressed_size, full = len;
		const unsigned char *t;

		if (m->compressed_size > 64 * 1024) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "Zip file with oversized link entry");
			return (ARCHIVE_FATAL);
		}
		archive_entry_set_size(entry, 0);
		if (m->method != 0) {
			z->remaining = (int64_t)m->compressed_size;
			if (m->method != 8 || len == 0 || zs_decode_entry(a, z, (int64_t)len, 0) != ARCHIVE_OK ||
			    (z->res.status != B2I_S_OK && z->res.out_bytes == 0)) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
				    "Unsupported ZIP compression method during decompression of link entry (%d: %s)",
				    m->method, zb_compression_name(m->method));
				return (ARCHIVE_FAILED);
			}
			t = z->cur;
			full = (size_t)z->res.out_bytes;
			len = 0;                        /* already consumed by the decode */
		} else
			t = __archive_read_ahead(a, len, NULL);
		if (t == NULL) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "Truncated Zip file");
			return (ARCHIVE_FATAL);
		}
		r = zb_set_symlink(a, &z->c, entry, t, full, m->zip_flags);
		if (r == ARCHIVE_FATAL)
			return (r);
		if (r != ARCHIVE_OK)
			ret = r;
		m->uncompressed_size = m->compressed_size = 0;
		if (__archive_read_consume(a, (int64_t)len) < 0) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "Read error skipping symlink target name");
			return (ARCHIVE_FATAL);
		}
		z->decoded = 0;
	} else if (0 == (m->zip_flags & ZIP_LENGTH_AT_END) ||
	    (m->uncompressed_size > 0 && m->uncompressed_size != 0xffffffff)) {
		archive_entry_set_size(entry, (int64_t)m->uncompressed_size);
	}
	z->remaining = (int64_t)m->compressed_size;
	if (0 == (m->zip_flags & ZIP_LENGTH_AT_END) && z->remaining < 1)
		z->end_of_entry = 1;

	archive_string_empty(&z->c.format_name);
	archive_string_sprintf(&z->c.format_name, "ZIP %d.%d (%s)", m->version / 10, m->version % 10,
	    zb_compression_name(m->method));
	a->archive.archive_format_name = z->c.format_name.s;
	return (ret);
}

/* ---- read_header (zip.c:3382-3473) ----------------------------------------------- */
static int
zs_read_header(struct archive_read *a, struct archive_entry *entry)
{
	struct zs_b200 *z = (struct zs_b200 *)a->format->data;

	a->archive.archive_format = ARCHIVE_FORMAT_ZIP;
	if (a->archive.archive_format_name == NULL)
		a->archive.archive_format_name = "ZIP";
	if (z->c.has_encrypted_entries == ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW)
		z->c.has_encrypted_entries = 0;
	__archive_read_reset_passphrase(a);

	/* search ahead for the next local file header */
	__archive_read_consume(a, z->unconsumed);
	z->unconsumed = 0;
	for (;;) {
		int64_t skipped = 0;
		const char *p, *end;
		ssize_t bytes;

		if ((p = __archive_read_ahead(a, 4, &bytes)) == NULL)
			return (ARCHIVE_FATAL);
		end = p + bytes;
		while (p + 4 <= end) {
			if (p[0] == 'P' && p[1] == 'K') {
				if (p[2] == '\003' && p[3] == '\004') {
					__archive_read_consume(a, skipped);
					return (zs_local_header(a, entry, z));
				}
				/* the central directory, or the end record of an empty archive */
				if ((p[2] == '\001' && p[3] == '\002') || (p[2] == '\005' && p[3] == '\006') ||
				    (p[2] == '\006' && p[3] == '\006'))
					return (ARCHIVE_EOF);
			}
			++p;
			++skipped;
		}
		__archive_read_consume(a, skipped);
	}
}

/* ---- stored bodies (zip.c:1592-1706, without decryption) --------------------------- */
static int
zs_read_stored(struct archive_read *a, struct zs_b200 *z, const void **buff, size_t *size)
{
	const char *b, *p;
	ssize_t avail;

	if (z->m.zip_flags & ZIP_LENGTH_AT_END) {
		b = __archive_read_ahead(a, 24, &avail);
		if (avail < 24) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file data");
			return (ARCHIVE_FATAL);
		}
		p = b;
		if (p[0] == 'P' && p[1] == 'K' && p[2] == '\007' && p[3] == '\010' &&
		    (zb_le32((const unsigned char *)p + 4) == z->computed_crc || z->c.ignore_crc32)) {
			z->end_of_entry = 1;
			return (ARCHIVE_OK);
		}
		++p;
		/* return the bytes before the next place a PK\7\8 could start */
		while (p < b + avail - 4) {
			if (p[3] == 'P') p += 3;
			else if (p[3] == 'K') p += 2;
			else if (p[3] == '\007') p += 1;
			else if (p[3] == '\010' && p[2] == '\007' && p[1] == 'K' && p[0] == 'P') break;
			else p += 4;
		}
		avail = p - b;
	} else {
		if (z->remaining == 0) {
			z->end_of_entry = 1;
			return (ARCHIVE_OK);
		}
		b = __archive_read_ahead(a, 1, &avail);
		if (avail <= 0) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file data");
			return (ARCHIVE_FATAL);
		}
		if (avail > z->remaining)
			avail = (ssize_t)z->remaining;
	}
	z->remaining -= avail;
	z->uread += avail;
	z->cread += avail;
	z->unconsumed += avail;
	*size = (size_t)avail;
	*buff = b;
	return (ARCHIVE_OK);
}

/* ---- read_data (zip.c:3071-3198) --------------------------------------------------- */
static int
zs_read_data(struct archive_read *a, const void **buff, size_t *size, int64_t *offset)
{
	struct zs_b200 *z = (struct zs_b200 *)a->format->data;
	struct zb_meta *m = &z->m;
	int r;

	if (z->c.has_encrypted_entries == ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW)
		z->c.has_encrypted_entries = 0;
	*offset = z->uread;
	*size = 0;
	*buff = NULL;
	if (z->end_of_entry)
		return (ARCHIVE_EOF);
	if (AE_IFREG != (m->mode & AE_IFMT))
		return (ARCHIVE_EOF);
	__archive_read_consume(a, z->unconsumed);
	z->unconsumed = 0;

	if (m->zip_flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED)) {
		z->c.has_encrypted_entries = 1;
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Encrypted ZIP entries are not supported by this build");
		return (ARCHIVE_FAILED);
	}
	if (m->method == 0) {
		if ((r = zs_read_stored(a, z, buff, size)) != ARCHIVE_OK)
			return (r);
		if (*size > 0 && !z->c.ignore_crc32) {
			/* the running CRC of everything delivered (zip.c:3154-3158), on the device */
			if (zs_need_ctx(a, z) != ARCHIVE_OK)
				return (ARCHIVE_FATAL);
			if (b2i_crc32(z->c.ctx, z->computed_crc, *buff, *size, &z->computed_crc) != B2I_OK) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "B200 CRC failed: %s",
				    b2i_last_error(z->c.ctx));
				return (ARCHIVE_FATAL);
			}
		}
	} else if (m->method == 8) {
		int64_t left;
		size_t n;
		int last;

		if (!z->decoded) {
			int known = 0 == (m->zip_flags & ZIP_LENGTH_AT_END);
			if (known && (zs_batch_take(a, z) || (zs_batch_decode(a, z) && zs_batch_take(a, z))))
				;                       /* decoded together with its neighbours */
			else if (zs_decode_entry(a, z, known ? z->remaining : 0,
			    known ? (int64_t)m->uncompressed_size : 0) != ARCHIVE_OK)
				return (ARCHIVE_FATAL);
			z->computed_crc = z->res.crc;
		}
		left = (int64_t)z->res.out_bytes - z->delivered;
		n = left > ZIP_BLOCK ? ZIP_BLOCK : (size_t)left;
		last = ((int64_t)n == left);
		if (z->res.status == B2I_S_BUF_ERROR) {
			/* zlib hands out what it could produce and reports Z_BUF_ERROR on the
			 * call that finds no input left */
			if (left == 0) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "ZIP decompression failed (%d)",
				    (int)z->res.status);
				return (ARCHIVE_FATAL);
			}
			last = 0;
		} else if (z->res.status != B2I_S_OK) {
			if (n < ZIP_BLOCK) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "ZIP decompression failed (%d)",
				    (int)z->res.status);
				return (ARCHIVE_FATAL);
			}
			last = 0;
		}
		*buff = z->cur + z->delivered;
		*size = n;
		z->delivered += (int64_t)n;
		z->uread = z->delivered;
		if (last)
			z->end_of_entry = 1;
	} else {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Unsupported ZIP compression method (%d: %s)", m->method, zb_compression_name(m->method));
		return (ARCHIVE_FAILED);
	}
	if (z->end_of_entry) {
		zs_end_of_file_marker(a, z);
		r = zb_end_of_entry_checks(a, &z->c, z->computed_crc, m->crc32, z->cread,
		    (int64_t)m->compressed_size, z->uread, (int64_t)m->uncompressed_size);
		if (r != ARCHIVE_OK) {
			*size = 0;
			*buff = NULL;
			return (r);
		}
	}
	return (ARCHIVE_OK);
}

/* ---- skip (zip.c:3475-3565) ------------------------------------------------------- */
static int
zs_read_data_skip(struct archive_read *a)
{
	struct zs_b200 *z = (struct zs_b200 *)a->format->data;
	int64_t n;

	n = __archive_read_consume(a, z->unconsumed);
	z->unconsumed = 0;
	if (n < 0)
		return (ARCHIVE_FATAL);
	if (z->end_of_entry)
		return (ARCHIVE_OK);
	if (0 == (z->m.zip_flags & ZIP_LENGTH_AT_END) || z->m.compressed_size > 0) {
		if (__archive_read_consume(a, z->remaining) < 0)
			return (ARCHIVE_FATAL);
		return (ARCHIVE_OK);
	}
	if (z->m.zip_flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED)) {
		z->c.has_encrypted_entries = 1;
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Encrypted ZIP entries are not supported by this build");
		return (ARCHIVE_FAILED);
	}
	if (z->m.method == 8) {
		/* the end of a body of unknown length is found by decoding it */
		while (!z->end_of_entry) {
			const void *b;
			size_t s;
			int64_t o;
			int r = zs_read_data(a, &b, &s, &o);
			if (r != ARCHIVE_OK && r != ARCHIVE_FAILED)
				return (r);
			if (r == ARCHIVE_FAILED)
				break;
		}
		return (ARCHIVE_OK);
	}
	/* stored or unknown: scan for the PK\7\8 signature of the data descriptor */
	for (;;) {
		const char *p, *b;
		ssize_t avail;

		b = __archive_read_ahead(a, 16, &avail);
		if (avail < 16) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Truncated ZIP file data");
			return (ARCHIVE_FATAL);
		}
		p = b;
		while (p <= b + avail - 16) {
			if (p[3] == 'P') p += 3;
			else if (p[3] == 'K') p += 2;
			else if (p[3] == '\007') p += 1;
			else if (p[3] == '\010' && p[2] == '\007' && p[1] == 'K' && p[0] == 'P') {
				__archive_read_consume(a, p - b + (z->m.used_zip64 ? 24 : 16));
				return (ARCHIVE_OK);
			} else p += 4;
		}
		__archive_read_consume(a, p - b);
	}
}

static int
zs_cleanup(struct archive_read *a)
{
	struct zs_b200 *z = (struct zs_b200 *)a->format->data;

	b200_buf_release(z->out, z->out_cap);
	b200_buf_release(z->b.out, z->b.out_cap);
	free(z->b.pos);
	free(z->b.d);
	free(z->b.r);
	b200_ctx_release(z->c.ctx, !z->c.ctx_bad);
	archive_string_free(&z->c.format_name);
	free(z);
	a->format->data = NULL;
	return (ARCHIVE_OK);
}

static int
zs_capabilities(struct archive_read *a)
{
	(void)a;
	return (ARCHIVE_READ_FORMAT_CAPS_ENCRYPT_DATA | ARCHIVE_READ_FORMAT_CAPS_ENCRYPT_METADATA);
}

static int
zs_has_encrypted_entries(struct archive_read *a)
{
	if (a && a->format && a->format->data)
		return (((struct zs_b200 *)a->format->data)->c.has_encrypted_entries);
	return (ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW);
}

int
archive_read_support_format_zip_streamable(struct archive *_a)
{
	struct archive_read *a = (struct archive_read *)_a;
	struct zs_b200 *z;
	int r;

	archive_check_magic(_a, ARCHIVE_READ_MAGIC, ARCHIVE_STATE_NEW, "archive_read_support_format_zip");
	if ((z = calloc(1, sizeof(*z))) == NULL) {
		archive_set_error(&a->archive, ENOMEM, "Can't allocate zip data");
		return (ARCHIVE_FATAL);
	}
	z->c.has_encrypted_entries = ARCHIVE_READ_FORMAT_ENCRYPTION_DONT_KNOW;
	r = __archive_read_register_format(a, z, "zip", zs_bid, zs_options, zs_read_header, zs_read_data,
	    zs_read_data_skip, NULL, zs_cleanup, zs_capabilities, zs_has_encrypted_entries);
	if (r != ARCHIVE_OK)
		free(z);
	return (ARCHIVE_OK);
}
