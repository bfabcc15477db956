/*
 * gzip_scan.c — host side of the gzip read filter: member header parsing and
 * the BGZF member chain (K4 of SURVEY.md 2.2).
 *
 * Header rules follow the reference's peek_at_header
 * (archive_read_support_filter_gzip.c:128-239): magic 1F 8B 08, no reserved
 * flag bits (FLG & 0xE0), MTIME at +4, optional FEXTRA (XLEN u16 + payload),
 * FNAME and FCOMMENT (NUL terminated, at most 1 MiB each), FHCRC (2 bytes, not
 * verified).  Returns the header length, 0 if this is not a gzip header.
 *
 * New relative to the reference (which skips the FEXTRA payload unparsed,
 * :170-176, and so must decode member N to find member N+1): the 'BC' subfield
 * of BGZF (SAM/BAM specification 4.1) gives BSIZE = member size - 1, so the
 * whole chain {deflate_offset, deflate_len, crc32, isize} is known up front and
 * every member becomes an independent descriptor for one device pass.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/b200inflate.h"

#define MAX_FIELD (1024 * 1024L)

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

size_t
b2i_gzip_peek_header(const void *buf, size_t size, size_t off, b2i_gzip_member *m)
{
	const uint8_t *p = (const uint8_t *)buf + off;
	size_t avail, len = 10;
	unsigned flags;
	uint32_t bsize = 0;
	int have_bsize = 0;

	if (m)
		memset(m, 0, sizeof(*m));
	if (buf == NULL || off >= size)
		return 0;
	avail = size - off;
	if (avail < 10)
		return 0;
	if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8)
		return 0;
	if (p[3] & 0xE0)
		return 0;
	flags = p[3];
	if (flags & 4) {
		size_t xlen, x;
		if (avail < len + 2)
			return 0;
		xlen = p[len] | (p[len + 1] << 8);
		/* subfields: SI1 SI2 SLEN(2) data; look for 'B','C',2 */
		if (avail >= len + 2 + xlen) {
			for (x = len + 2; x + 4 <= len + 2 + xlen;) {
				size_t slen = p[x + 2] | (p[x + 3] << 8);
				if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2 && x + 6 <= len + 2 + xlen) {
					bsize = p[x + 4] | (p[x + 5] << 8);
					have_bsize = 1;
				}
				x += 4 + slen;
			}
		}
		len += 2 + xlen;
	}
	if (flags & 8) {
		size_t start = len;
		do {
			++len;
			if (len > avail || len - start > (size_t)MAX_FIELD)
				return 0;
		} while (p[len - 1] != 0);
		if (m)
			m->name_offset = (uint32_t)(off + start);
	}
	if (flags & 16) {
		size_t start = len;
		do {
			++len;
			if (len > avail || len - start > (size_t)MAX_FIELD)
				return 0;
		} while (p[len - 1] != 0);
	}
	if (flags & 2) {
		if (avail < len + 2)
			return 0;
		len += 2;
	}
	if (len > avail)
		return 0;
	if (m) {
		m->header_offset = off;
		m->header_len = (uint32_t)len;
		m->deflate_offset = off + len;
		m->mtime = le32(p + 4);
		if (have_bsize) {
			/* total member = BSIZE + 1 = header + deflate + 8-byte trailer */
			size_t total = (size_t)bsize + 1;
			if (total >= len + 8 && total <= avail) {
				m->deflate_len = total - len - 8;
				m->crc32 = le32(p + total - 8);
				m->isize = le32(p + total - 4);
				if (m->deflate_len == 0)
					m->deflate_len = 0;    /* cannot happen: an empty block is >= 2 bytes */
			}
		}
	}
	return len;
}

int
b2i_gzip_scan_bgzf(const void *buf, size_t size, size_t off, b2i_gzip_member **members,
    size_t *n, size_t *end_off)
{
	size_t cap = 1024, cnt = 0;
	b2i_gzip_member *v;

	if (members == NULL || n == NULL)
		return B2I_E_INVAL;
	v = malloc(cap * sizeof(*v));
	if (v == NULL)
		return B2I_E_NOMEM;
	while (off < size) {
		b2i_gzip_member m;
		size_t hl = b2i_gzip_peek_header(buf, size, off, &m);
		if (hl == 0 || m.deflate_len == 0)
			break;          /* garbage, or a member without BSIZE: caller takes over */
		if (cnt == cap) {
			b2i_gzip_member *nv = realloc(v, cap * 2 * sizeof(*v));
			if (nv == NULL) {
				free(v);
				return B2I_E_NOMEM;
			}
			v = nv;
			cap *= 2;
		}
		v[cnt++] = m;
		off = m.deflate_offset + m.deflate_len + 8;
	}
	*members = v;
	*n = cnt;
	if (end_off)
		*end_off = off;
	return B2I_OK;
}
