/*
 * zip_index.c — host side of the seekable ZIP reader, flattened: locate the
 * end-of-central-directory record, walk the central directory once into flat
 * arrays, sort by local-header offset, and resolve every entry's data offset
 * and definitive CRC/sizes from its local header.  The result is what the
 * batch plan is built from (one descriptor per entry).
 *
 * Written from the ZIP APPNOTE; the acceptance rules and field precedence
 * follow the reference reader so that both see the same entries:
 *   EOCD search in the last 16 KiB, last match wins, i > 0 only
 *       archive_read_support_format_zip.c:3720-3773 (seekable_bid)
 *   EOCD sanity checks and cd offset                      :3635-3669 (read_eocd)
 *   ZIP64 locator exactly 20 bytes before the EOCD        :3675-3718
 *   forward scan for the first PK\1\2 / PK\5\6 / PK\6\6 => `correction`
 *                                                         :3887-3924
 *   46-byte central records until PK\5\6 / PK\6\6         :3936-3986
 *   ZIP64 extra 0x0001: usize, csize, offset in that order, each only when
 *   the 32-bit field is 0xffffffff                        :526-577
 *   iteration in ascending local-header offset, duplicates of one offset
 *   dropped (red-black tree insert)                       :3778-3789, 4031, 4298-4303
 *   local header: flags/method/crc/sizes/name/extra, reconciliation with the
 *   central values (local wins when non-zero, WARN on mismatch),
 *   data start = offset + 30 + name + extra               :936-972, 1106-1150
 *
 * No allocation per entry, no tree: two passes over flat arrays.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200inflate.h"

#define ZIP_ENCRYPTED          (1 << 0)
#define ZIP_LENGTH_AT_END      (1 << 3)
#define ZIP_STRONG_ENCRYPTED   (1 << 6)

#define IFMT   0170000u
#define IFDIR  0040000u
#define IFREG  0100000u
#define IFIFO  0010000u

static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }
static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint64_t le64(const uint8_t *p) { return (uint64_t)le32(p) | (uint64_t)le32(p + 4) << 32; }

static int
err(char errbuf[128], const char *msg)
{
	if (errbuf)
		snprintf(errbuf, 128, "%s", msg);
	return B2I_E_FORMAT;
}

/* MS-DOS date/time -> time_t, as libarchive/archive_time.c dos_to_unix does
 * (local time zone interpretation through mktime) */
#include <time.h>
static int64_t dos_time_slow(uint32_t d);

/* mktime() costs microseconds and archives repeat time stamps: small
 * direct-mapped memo (per call tree, no shared state) */
struct dos_memo { uint32_t key[64]; int64_t val[64]; uint8_t set[64]; };
static int64_t
dos_time(struct dos_memo *m, uint32_t d)
{
	unsigned h = (d ^ (d >> 11) ^ (d >> 22)) & 63u;
	if (!m->set[h] || m->key[h] != d) {
		m->key[h] = d;
		m->val[h] = dos_time_slow(d);
		m->set[h] = 1;
	}
	return m->val[h];
}

static int64_t
dos_time_slow(uint32_t d)
{
	struct tm ts;
	uint16_t msTime = (uint16_t)(0xffff & d), msDate = (uint16_t)(d >> 16);

	memset(&ts, 0, sizeof(ts));
	ts.tm_year = ((msDate >> 9) & 0x7f) + 80;
	ts.tm_mon = ((msDate >> 5) & 0x0f) - 1;
	ts.tm_mday = msDate & 0x1f;
	ts.tm_hour = (msTime >> 11) & 0x1f;
	ts.tm_min = (msTime >> 5) & 0x3f;
	ts.tm_sec = (msTime << 1) & 0x3e;
	ts.tm_isdst = -1;
	return (int64_t)mktime(&ts);
}

struct cdrec {
	uint64_t lho, csize, usize;
	uint32_t crc, idx, mode, uid, gid;
	int64_t  mtime, atime, ctime;
	uint16_t flags;
	uint8_t  method, system;
};

/* Extra fields, with the reference's order and conditions (zip.c:474-900): ZIP64
 * sizes/offset (0x0001), times (0x5455, 0x5855), owner (0x5855, 0x7855, 0x7875)
 * and the experimental attributes field (0x6c78).  The Unicode path (0x7075)
 * needs the converted pathname and is left to the caller.  `with_lho`: central
 * directory records only.  Returns 0, or -1 with *why set. */
static int
apply_extra(const uint8_t *p, size_t n, struct cdrec *r, int with_lho, const char **why)
{
	size_t off = 0;

	if (n == 0)
		return 0;
	if (n < 4) {
		for (size_t i = 0; i < n; i++)
			if (p[i] != 0) { *why = "Too-small extra data"; return -1; }
		return 0;
	}
	while (off <= n - 4) {
		uint16_t id = le16(p + off), sz = le16(p + off + 2);
		size_t o;
		off += 4;
		if (off + sz > n) { *why = "Extra data overflow"; return -1; }
		o = off;
		if (id == 0x0001) {
			unsigned left = sz;
			if (r->usize == 0xffffffffull) {
				uint64_t t;
				if (left < 8 || (t = le64(p + o)) > INT64_MAX) { *why = "Malformed 64-bit uncompressed size"; return -1; }
				r->usize = t; o += 8; left -= 8;
			}
			if (r->csize == 0xffffffffull) {
				uint64_t t;
				if (left < 8 || (t = le64(p + o)) > INT64_MAX) { *why = "Malformed 64-bit compressed size"; return -1; }
				r->csize = t; o += 8; left -= 8;
			}
			if (with_lho && r->lho == 0xffffffffull) {
				uint64_t t;
				if (left < 8 || (t = le64(p + o)) > INT64_MAX) { *why = "Malformed 64-bit local header offset"; return -1; }
				r->lho = t; o += 8; left -= 8;
			}
		} else if (id == 0x5455) {
			unsigned left = sz;
			int fl;
			if (left == 0) { *why = "Incomplete extended time field"; return -1; }
			fl = p[o++]; left--;
			if (fl & 1) { if (left >= 4) { r->mtime = le32(p + o); o += 4; left -= 4; } else goto next; }
			if (fl & 2) { if (left >= 4) { r->atime = le32(p + o); o += 4; left -= 4; } else goto next; }
			if (fl & 4) { if (left >= 4) { r->ctime = le32(p + o); o += 4; left -= 4; } else goto next; }
		} else if (id == 0x5855) {
			if (sz >= 8) { r->atime = le32(p + o); r->mtime = le32(p + o + 4); }
			if (sz >= 12) { r->uid = le16(p + o + 8); r->gid = le16(p + o + 10); }
		} else if (id == 0x7855) {
			if (sz >= 2) r->uid = le16(p + o);
			if (sz >= 4) r->gid = le16(p + o + 2);
		} else if (id == 0x7875) {
			unsigned us = 0, gs;
			if (sz >= 1 && p[o] == 1) {
				if (sz >= 4) {
					us = p[o + 1];
					if (us == 2) r->uid = le16(p + o + 2);
					else if (us == 4 && sz >= 6) r->uid = le32(p + o + 2);
				}
				if (sz >= 2 + us + 3) {
					gs = p[o + 2 + us];
					if (gs == 2) r->gid = le16(p + o + 2 + us + 1);
					else if (gs == 4 && sz >= 2 + us + 5) r->gid = le32(p + o + 2 + us + 1);
				}
			}
		} else if (id == 0x6c78) {
			unsigned left = sz;
			int bitmap, last;
			if (left < 1) goto next;
			last = bitmap = p[o++]; left--;
			while ((last & 0x80) && left >= 1) { last = p[o++]; left--; }
			if (bitmap & 1) { if (left < 2) goto next; r->system = (uint8_t)(le16(p + o) >> 8); o += 2; left -= 2; }
			if (bitmap & 2) { if (left < 2) goto next; o += 2; left -= 2; }
			if (bitmap & 4) {
				uint32_t ext;
				if (left < 4) goto next;
				ext = le32(p + o);
				if (r->system == 3)
					r->mode = ext >> 16;
				else if (r->system == 0) {
					r->mode = (ext & 0x10) ? (IFDIR | 0775) : (IFREG | 0664);
					if (ext & 0x01)
						r->mode &= 0555;
				} else
					r->mode = 0;
			}
		} else if (id == 0x9901) {
			if (sz < 6) { *why = "Incomplete AES field"; return -1; }
		}
next:
		off += sz;
	}
	return 0;
}

static int
cmp_rec(const void *a, const void *b)
{
	const struct cdrec *x = a, *y = b;
	if (x->lho != y->lho)
		return x->lho < y->lho ? -1 : 1;
	return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* ---- source access --------------------------------------------------------------
 * The walk below never assumes the archive is one memory image: it asks for byte
 * ranges through a fetch function (valid until the next call).  A memory image answers
 * with a pointer into itself; the libarchive plugin answers from a sliding read-ahead
 * window of its seekable source (archive_read_open_filename), so that indexing a large
 * file reads the tail, the directory and the local headers - never the whole file into
 * one buffer. */
struct src {
	b2i_fetch_fn fetch;
	void *user;
	uint64_t size;
};

static const uint8_t *
mem_fetch(void *user, uint64_t off, size_t len)
{
	(void)len;
	return (const uint8_t *)user + off;
}

static const uint8_t *
get(const struct src *s, uint64_t off, size_t len)
{
	if (off > s->size || s->size - off < len)
		return NULL;
	return s->fetch(s->user, off, len);
}

/* Does the last 16 KiB of a file hold an end-of-central-directory record the reader
 * would accept (zip.c:3720-3773, 3635-3718)?  `tail` = the last tail_len bytes. */
int
b2i_zip_probe_tail(const void *tailp, size_t tail_len, uint64_t file_size)
{
	const uint8_t *tail = tailp;
	long i;

	if (tail == NULL || tail_len < 22 || tail_len > file_size)
		return 0;
	for (i = (long)tail_len - 22; i > 0; i--) {
		const uint8_t *p = tail + i;
		if (memcmp(p, "PK\005\006", 4) != 0)
			continue;
		{
			uint64_t pos = file_size - tail_len + (uint64_t)i;
			uint32_t cd_size = le32(p + 12), cd_off = le32(p + 16);
			if (le16(p + 4) == 0 && le16(p + 6) == 0 && le16(p + 10) == le16(p + 8) &&
			    (uint64_t)cd_off + cd_size <= pos)
				return 1;
			if (i >= 20 && memcmp(p - 20, "PK\006\007", 4) == 0 &&
			    le32(p - 20 + 4) == 0 && le32(p - 20 + 16) == 1)
				return 1;       /* the ZIP64 record itself is validated by the index walk */
		}
		break;
	}
	return 0;
}

int
b2i_zip_index_build_cb(b2i_fetch_fn fetch, void *user, uint64_t size, b2i_zip_index *out, char errbuf[128])
{
	struct src S = { fetch, user, size };
	int64_t cd_offset = -1, cd_adjusted = -1;
	size_t tail;
	uint64_t tail_start;
	long i;
	int found = 0;
	struct dos_memo memo;
	uint8_t *tailbuf = NULL, *cd = NULL;
	const uint8_t *p;

	memset(&memo, 0, sizeof(memo));

	if (out == NULL || fetch == NULL)
		return B2I_E_INVAL;
	memset(out, 0, sizeof(*out));
	if (errbuf)
		errbuf[0] = 0;
	if (size == 0)
		return err(errbuf, "empty file");

	/* --- end of central directory: last PK\5\6 in the final 16 KiB, i > 0 --- */
	tail = size < 16384 ? (size_t)size : 16384;
	tail_start = size - tail;
	if ((p = get(&S, tail_start, tail)) == NULL || (tailbuf = malloc(tail)) == NULL)
		return p == NULL ? err(errbuf, "cannot read the end of the file") : B2I_E_NOMEM;
	memcpy(tailbuf, p, tail);
	for (i = (long)tail - 22; i > 0; i--) {
		p = tailbuf + i;
		if (memcmp(p, "PK\005\006", 4) != 0)
			continue;
		{
			int64_t pos = (int64_t)(tail_start + (uint64_t)i);
			uint16_t disk = le16(p + 4);
			uint32_t cd_size = le32(p + 12), cd_off = le32(p + 16);
			if (disk == 0 && le16(p + 6) == 0 && le16(p + 10) == le16(p + 8) &&
			    (int64_t)cd_off + cd_size <= pos) {
				cd_offset = cd_off;
				cd_adjusted = pos - cd_size;
				found = 1;
			}
			if (i >= 20 && memcmp(p - 20, "PK\006\007", 4) == 0) {
				const uint8_t *l = p - 20;
				if (le32(l + 4) == 0 && le32(l + 16) == 1) {
					uint64_t e64 = le64(l + 8);
					const uint8_t *q;
					if (e64 <= size && size - e64 >= 56 && (q = get(&S, e64, 56)) != NULL) {
						uint64_t e64size = le64(q + 4) + 12;
						if (e64size >= 56 && e64size <= 16384 && size - e64 >= e64size &&
						    le32(q + 16) == 0 && le32(q + 20) == 0 &&
						    le64(q + 24) == le64(q + 32)) {
							cd_offset = (int64_t)le64(q + 48);
							cd_adjusted = cd_offset;
							found = 1;
						}
					}
				}
			}
		}
		break;      /* only the last EOCD signature is examined */
	}
	free(tailbuf);
	if (!found)
		return err(errbuf, "no end-of-central-directory record");
	if (cd_adjusted < 0 || (uint64_t)cd_adjusted > size)
		return err(errbuf, "central directory offset out of range");

	/* --- the directory region [cd_adjusted, size): kept in memory for the walk --- */
	const size_t cdlen = (size_t)(size - (uint64_t)cd_adjusted);
	if ((p = get(&S, (uint64_t)cd_adjusted, cdlen)) == NULL)
		return err(errbuf, "cannot read the central directory");
	if (fetch == mem_fetch) {
		cd = NULL;                    /* a memory image: walk it in place */
	} else {
		if ((cd = malloc(cdlen ? cdlen : 1)) == NULL)
			return B2I_E_NOMEM;
		memcpy(cd, p, cdlen);
		p = cd;
	}
	const uint8_t *base = p - cd_adjusted;   /* base + file offset, valid for offsets >= cd_adjusted only */

	/* --- real start of the directory => correction for prepended data --- */
	uint64_t pos = (uint64_t)cd_adjusted;
	for (found = 0; pos + 4 < size; pos++) {
		const uint8_t *c = base + pos;
		if (c[0] == 'P' && c[1] == 'K' &&
		    ((c[2] == 1 && c[3] == 2) || (c[2] == 5 && c[3] == 6) || (c[2] == 6 && c[3] == 6))) {
			found = 1;
			break;
		}
	}
	if (!found || size - pos < 20) {
		free(cd);
		return err(errbuf, "central directory not found");
	}
	out->correction = (int64_t)pos - cd_offset;

	/* --- pass 1: count records --- */
	size_t n = 0;
	uint64_t q = pos;
	for (;;) {
		const char *bad = NULL;
		if (size - q < 4)
			bad = "truncated central directory";
		else if (memcmp(base + q, "PK\006\006", 4) == 0 || memcmp(base + q, "PK\005\006", 4) == 0)
			break;
		else if (memcmp(base + q, "PK\001\002", 4) != 0)
			bad = "Invalid central directory signature";
		else if (size - q < 46)
			bad = "truncated central directory";
		else if (size - q - 46 < (size_t)le16(base + q + 28) + le16(base + q + 30))
			bad = "Truncated ZIP file header";
		if (bad != NULL) {
			free(cd);
			return err(errbuf, bad);
		}
		q += 46 + (size_t)le16(base + q + 28) + le16(base + q + 30) + le16(base + q + 32);
		if (q > size)
			q = size;       /* a long comment may run off the end; the next read fails */
		n++;
	}

	struct cdrec *recs = malloc((n ? n : 1) * sizeof(*recs));
	if (recs == NULL) {
		free(cd);
		return B2I_E_NOMEM;
	}

	/* --- pass 2: decode records --- */
	q = pos;
	for (size_t k = 0; k < n; k++) {
		const uint8_t *c = base + q;
		struct cdrec *r = &recs[k];
		uint32_t ext = le32(c + 38);
		size_t nl = le16(c + 28), xl = le16(c + 30), cl = le16(c + 32);
		const char *why = NULL;
		uint64_t lho32 = le32(c + 42);

		memset(r, 0, sizeof(*r));
		r->idx = (uint32_t)k;
		r->system = c[5];
		r->flags = le16(c + 8);
		if (r->flags & (ZIP_ENCRYPTED | ZIP_STRONG_ENCRYPTED))
			out->has_encrypted_entries = 1;
		r->method = (uint8_t)le16(c + 10);
		r->mtime = dos_time(&memo, le32(c + 12));
		r->crc = le32(c + 16);
		r->csize = le32(c + 20);
		r->usize = le32(c + 24);
		/* the reference adds the correction BEFORE testing for 0xffffffff, so a
		 * ZIP64 offset is only honoured when the correction is zero (zip.c:3985, :559) */
		r->lho = lho32 + (uint64_t)out->correction;
		if (c[5] == 3)
			r->mode = ext >> 16;
		else if (c[5] == 0) {
			r->mode = (ext & 0x10) ? (IFDIR | 0775) : (IFREG | 0664);
			if (ext & 0x01)
				r->mode &= 0555 | IFMT;
		} else
			r->mode = 0;
		if (apply_extra(c + 46 + nl, xl, r, 1, &why) != 0) {
			free(recs);
			free(cd);
			return err(errbuf, why);
		}
		q += 46 + nl + xl + cl;
	}
	free(cd);

	/* --- ascending local-header offset; first record of an offset wins --- */
	qsort(recs, n, sizeof(*recs), cmp_rec);
	size_t m = 0;
	for (size_t k = 0; k < n; k++)
		if (m == 0 || recs[k].lho != recs[m - 1].lho)
			recs[m++] = recs[k];

	b2i_zip_entry *ents = calloc(m ? m : 1, sizeof(*ents));
	size_t names_cap = 64 * (m + 1);
	char *names = malloc(names_cap);
	if (ents == NULL || names == NULL) {
		free(recs); free(ents); free(names);
		return B2I_E_NOMEM;
	}

	/* --- pass 3: local headers, in ascending offset (sequential for the source) --- */
	size_t names_len = 0;
	for (size_t k = 0; k < m; k++) {
		const struct cdrec *r = &recs[k];
		b2i_zip_entry *e = &ents[k];

		e->local_header_offset = r->lho;
		e->compressed_size = r->csize;
		e->uncompressed_size = r->usize;
		e->crc32 = r->crc;
		e->zip_flags = r->flags;
		e->method = r->method;
		e->mode = r->mode;
		e->mtime = r->mtime;
		e->atime = r->atime;
		e->ctime = r->ctime;
		e->uid = r->uid;
		e->gid = r->gid;
		if (r->lho > size || size - r->lho < 30 || (p = get(&S, r->lho, 30)) == NULL) {
			e->warn |= B2I_ZW_TRUNCATED;
			continue;
		}
		if (memcmp(p, "PK\003\004", 4) != 0) {
			e->warn |= B2I_ZW_BAD_LOCAL_HEADER;
			continue;
		}
		size_t nl = le16(p + 26), xl = le16(p + 28);
		if (size - r->lho - 30 < nl + xl || (p = get(&S, r->lho, 30 + nl + xl)) == NULL) {
			e->warn |= B2I_ZW_TRUNCATED;
			continue;
		}
		uint64_t l_usize, l_csize;
		uint32_t l_crc = le32(p + 14);
		struct cdrec l = *r;     /* what the central record said stays unless the local header overrides it */
		const char *why = NULL;

		l.usize = le32(p + 22);
		l.csize = le32(p + 18);
		l.mtime = dos_time(&memo, le32(p + 10));
		l.system = p[5];

		e->version = p[4];
		e->system = p[5];
		e->zip_flags = le16(p + 6);
		e->method = (uint8_t)le16(p + 8);
		if (names_len + nl > names_cap) {
			char *nn;
			names_cap = 2 * (names_len + nl);
			if ((nn = realloc(names, names_cap)) == NULL) {
				free(recs); free(ents); free(names);
				return B2I_E_NOMEM;
			}
			names = nn;
		}
		e->name_offset = (uint32_t)names_len;
		e->name_len = (uint16_t)nl;
		memcpy(names + names_len, p + 30, nl);
		names_len += nl;
		e->local_extra_offset = r->lho + 30 + nl;
		e->local_extra_len = (uint16_t)xl;
		if (apply_extra(p + 30 + nl, xl, &l, 0, &why) != 0) {
			e->warn |= B2I_ZW_BAD_LOCAL_HEADER;
			continue;
		}
		l_usize = l.usize;
		l_csize = l.csize;
		e->mtime = l.mtime;
		e->atime = l.atime;
		e->ctime = l.ctime;
		e->uid = l.uid;
		e->gid = l.gid;
		e->mode = l.mode;
		e->system = l.system;
		e->data_offset = r->lho + 30 + nl + xl;
		/* central values are definitive; local ones win when present (zip.c:1106-1150) */
		e->zip_flags &= (uint16_t)~ZIP_LENGTH_AT_END;
		if (l_crc != 0) {
			if (l_crc != r->crc)
				e->warn |= B2I_ZW_CRC_INCONSISTENT;
			e->crc32 = l_crc;
		}
		if (l_csize != 0 && l_csize != 0xffffffffull) {
			if (l_csize != r->csize)
				e->warn |= B2I_ZW_CSIZE_INCONSISTENT;
			e->compressed_size = l_csize;
		}
		if (l_usize != 0 && l_usize != 0xffffffffull) {
			if (l_usize != r->usize)
				e->warn |= B2I_ZW_USIZE_INCONSISTENT;
			e->uncompressed_size = l_usize;
		}
		if (e->data_offset > size || size - e->data_offset < e->compressed_size)
			e->warn |= B2I_ZW_TRUNCATED;
		/* file type fix-ups (zip.c:1029-1062) */
		if ((e->mode & IFMT) == IFIFO)
			e->mode = (e->mode & ~IFMT) | IFREG;
		if (e->mode == 0)
			e->mode |= 0664;
		if ((e->mode & IFMT) != IFDIR) {
			if (nl > 0 && p[30 + nl - 1] == '/')
				e->mode = (e->mode & ~IFMT) | IFDIR | 0111;
			else if ((e->mode & IFMT) == 0)
				e->mode |= IFREG;
		}
	}
	free(recs);
	out->n = m;
	out->entries = ents;
	out->names = names;
	out->names_len = names_len;
	return B2I_OK;
}

int
b2i_zip_index_build(const void *archive, size_t size, b2i_zip_index *out, char errbuf[128])
{
	if (archive == NULL && size)
		return B2I_E_INVAL;
	return b2i_zip_index_build_cb(mem_fetch, (void *)(uintptr_t)archive, size, out, errbuf);
}

void
b2i_zip_index_free(b2i_zip_index *ix)
{
	if (ix == NULL)
		return;
	free(ix->entries);
	free(ix->names);
	memset(ix, 0, sizeof(*ix));
}
