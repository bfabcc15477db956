/*
 * archive_read_support_filter_gzip_b200.c — libarchive gzip read filter whose
 * members are decoded on a B200 through include/b200inflate.h instead of zlib.
 *
 * Drop-in for libarchive/archive_read_support_filter_gzip.c: defines
 * archive_read_support_filter_gzip (and the deprecated
 * archive_read_support_compression_gzip, gzip.c:85-92) and registers a bidder
 * with the unmodified read core (__archive_read_register_bidder,
 * archive_read_private.h:242-245).  Written from scratch against that
 * interface.
 *
 * The reference discovers members serially (header, zlib until Z_STREAM_END,
 * 8-byte trailer, next header; gzip.c:431-511).  Here a window of upstream
 * bytes is scanned for the BGZF member chain (BSIZE in the 'BC' extra
 * subfield), every complete member of the window becomes a descriptor, ONE
 * device pass decodes them all, and read() serves the decoded bytes in blocks
 * of at most 64 KiB (gzip.c:314).  The chain is known without decoding, so the
 * NEXT window is submitted (b2i_submit) as soon as the current one has been
 * collected: it is copied in, decoded and copied out into a second pinned
 * buffer while the current one is served.  A member without BSIZE is decoded on the
 * device from "here to the end of the window"; the number of bytes it consumed
 * locates its trailer and the next header.  So that format bidding on a large
 * member does not have to decode all of it, the FIRST block of such a member is
 * produced from the bytes already buffered (decode with the output capped at
 * one block); the member is decoded whole when more is asked for.  Like the
 * reference, trailer CRC-32 and ISIZE are not checked by default (gzip.c:423
 * leaves a TODO); the environment variable B2I_GZIP_VERIFY=1 turns the check on
 * (mismatch => "gzip decompression failed").
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

#include "archive.h"
#include "archive_entry.h"
#include "archive_private.h"
#include "archive_read_private.h"

#include "b200inflate.h"
#include "b200_ctx_pool.h"
#include <stdio.h>

#define OUT_BLOCK      (64 * 1024)            /* gzip.c:314 */
#define WINDOW_TARGET  ((size_t)32 << 20)     /* decode at most this much input per device pass */
#define GZ_JOBS        3                      /* BGZF windows in flight ahead of the one being served */
#define WINDOW_READ    ((size_t)16 << 20)     /* what a block-sized source (a file) is asked to have
                                                 buffered before a pass: one pass per 64 KiB read
                                                 block would be one pass per three members */

struct gz_b200 {
	b2i_ctx        *ctx;
	int             ctx_bad;      /* a device call failed: do not hand the context back */
	unsigned char  *out;          /* pinned: decoded bytes of the current window */
	size_t          out_cap, out_len, served;
	int             eof;          /* no more members */
	int             failed;       /* sticky fatal */
	const char     *pending_fail; /* error to raise once the full blocks before it are served */
	int             verify;
	size_t          member_skip;  /* bytes of the member at the read position already served */
	uint32_t        mtime;
	char           *name;
	int             have_meta;
	/* BGZF windows decoding ahead (b2i_submit), oldest first: each with its own pinned buffer
	 * (decoded bytes land at buf + OUT_BLOCK: room for the tail of `out`), descriptors, results */
	struct gz_job {
		b2i_job        *job;
		unsigned char  *buf;
		size_t          cap, n;
		b2i_stream_desc   *d;
		b2i_stream_result *r;
		const unsigned char *src;     /* the window's input after upstream has consumed it: the
		                               * library's staged copy, or `keep` below */
		unsigned char  *keep;
		size_t          keep_cap;
	} q[GZ_JOBS];
	int             qh, qn;       /* head and length of the queue (slots are used round-robin) */
	size_t          windows;      /* BGZF windows submitted so far */
};

static int	bgzf_submit(struct archive_read_filter *, const unsigned char *, const b2i_gzip_member *, size_t);
static int	bgzf_collect(struct archive_read_filter *);
static int	bgzf_redo_window(struct gz_b200 *, struct gz_job *);

static int	gz_bid(struct archive_read_filter_bidder *, struct archive_read_filter *);
static int	gz_init(struct archive_read_filter *);

static const struct archive_read_filter_bidder_vtable gz_bidder_vtable = {
	.bid = gz_bid,
	.init = gz_init,
};

#if ARCHIVE_VERSION_NUMBER < 4000000
int
archive_read_support_compression_gzip(struct archive *a)
{
	return (archive_read_support_filter_gzip(a));
}
#endif

int
archive_read_support_filter_gzip(struct archive *_a)
{
	struct archive_read *a = (struct archive_read *)_a;

	if (__archive_read_register_bidder(a, NULL, "gzip", &gz_bidder_vtable) != ARCHIVE_OK)
		return (ARCHIVE_FATAL);
	return (ARCHIVE_OK);
}

/* header parse over the upstream read-ahead buffer: returns the header length or 0
 * (gzip.c:128-239 semantics, implemented by b2i_gzip_peek_header) */
static size_t
peek_header(struct archive_read_filter *up, b2i_gzip_member *m)
{
	ssize_t avail;
	size_t want = 10;
	const void *p;

	for (;;) {
		p = __archive_read_filter_ahead(up, want, &avail);
		if (p == NULL) {
			/* fewer than `want` bytes left: look at what there is */
			p = __archive_read_filter_ahead(up, 1, &avail);
			if (p == NULL || avail <= 0)
				return (0);
			return (b2i_gzip_peek_header(p, (size_t)avail, 0, m));
		}
		size_t hl = b2i_gzip_peek_header(p, (size_t)avail, 0, m);
		if (hl != 0)
			return (hl);
		/* a header with long name/comment/extra may simply not fit yet */
		if ((size_t)avail < want || want > (3u << 20))
			return (0);
		if (((const unsigned char *)p)[0] != 0x1f || ((const unsigned char *)p)[1] != 0x8b ||
		    ((const unsigned char *)p)[2] != 8 || (((const unsigned char *)p)[3] & 0xE0))
			return (0);
		want = (size_t)avail >= want * 4 ? (size_t)avail + 1 : want * 4;
	}
}

static int
gz_bid(struct archive_read_filter_bidder *self, struct archive_read_filter *filter)
{
	(void)self;
	return (peek_header(filter, NULL) ? 27 : 0);       /* 24 magic bits + 3 reserved-flag bits */
}

static int
gz_read_header(struct archive_read_filter *self, struct archive_entry *entry)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;

	if (g->mtime != 0)
		archive_entry_set_mtime(entry, g->mtime, 0);
	if (g->name)
		archive_entry_set_pathname(entry, g->name);
	return (ARCHIVE_OK);
}

static int
fatal(struct archive_read_filter *self, struct gz_b200 *g, const char *msg)
{
	if (getenv("B2I_GZ_TRACE")) fprintf(stderr, "gz fatal: %s (out_len %zu served %zu skip %zu)\n", msg, g->out_len, g->served, g->member_skip);
	archive_set_error(&self->archive->archive, ARCHIVE_ERRNO_MISC, "%s", msg);
	g->failed = 1;
	return (ARCHIVE_FATAL);
}

static void
note_header(struct gz_b200 *g, const unsigned char *base, const b2i_gzip_member *m)
{
	/* like the reference, the entry's name / mtime come from a member header */
	g->mtime = m->mtime;
	free(g->name);
	g->name = m->name_offset ? strdup((const char *)base + m->name_offset) : NULL;
	g->have_meta = 1;
}

/* pending bytes move to the front; room for `need` more behind them */
static int
make_room(struct gz_b200 *g, size_t need)
{
	size_t pending = g->out_len - g->served;

	if (pending + need + 16 > g->out_cap) {
		size_t cap = 0;
		unsigned char *nb = b200_buf_acquire(pending + need + 16 + ((pending + need) >> 2), &cap);
		if (nb == NULL)
			return (-1);
		if (pending)
			memcpy(nb, g->out + g->served, pending);
		b200_buf_release(g->out, g->out_cap);
		g->out = nb;
		g->out_cap = cap;
	} else if (g->served && pending) {
		memmove(g->out, g->out + g->served, pending);
	}
	g->served = 0;
	g->out_len = pending;
	return (0);
}

/* Look at what is buffered upstream; if a BGZF chain starts there, return it.  *pp: the
 * read-ahead view.  n == 0: not a (complete) BGZF member at the read position. */
static int
bgzf_scan(struct archive_read_filter *up, const unsigned char **pp, b2i_gzip_member **mem, size_t *n)
{
	const unsigned char *p;
	ssize_t avail;
	size_t end = 0;

	*mem = NULL;
	*n = 0;
	/* a memory source answers with everything it has; a file source (64 KiB read blocks,
	 * archive_read_open_filename.c:389-461) is asked to collect a window's worth first */
	p = __archive_read_filter_ahead(up, 18, &avail);
	if (p != NULL && avail < (ssize_t)WINDOW_READ && p[0] == 0x1f && p[1] == 0x8b && (p[3] & 4) &&
	    p[12] == 'B' && p[13] == 'C')
		(void)__archive_read_filter_ahead(up, WINDOW_READ, &avail);    /* NULL: fewer bytes are left */
	p = __archive_read_filter_ahead(up, 1, &avail);
	*pp = p;
	if (p == NULL || avail <= 0)
		return (0);
	if (b2i_gzip_scan_bgzf(p, (size_t)avail, 0, mem, n, &end) != B2I_OK)
		return (-1);
	return (0);
}

/* queue one device pass over the members of the chain that fit a window; their input is
 * consumed upstream as soon as it has been copied to the device */
static int
bgzf_submit(struct archive_read_filter *self, const unsigned char *p, const b2i_gzip_member *mem, size_t n)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;
	struct gz_job *j = &g->q[(g->qh + g->qn) % GZ_JOBS];
	size_t out = 0, in_used = 0, m_used = 0, i;
	int rc;

	free(j->d);
	free(j->r);
	j->d = calloc(n, sizeof(*j->d));
	j->r = calloc(n, sizeof(*j->r));
	if (j->d == NULL || j->r == NULL)
		return (fatal(self, g, "Can't allocate data for gzip decompression"));
	for (i = 0; i < n; i++) {
		b2i_stream_desc *d = &j->d[i];
		/* the first window is small: the caller (format bidding, the first read) waits for it */
		if (i > 0 && mem[i].header_offset >= (g->windows ? WINDOW_TARGET : WINDOW_TARGET / 4))
			break;
		d->in_off = mem[i].deflate_offset;
		d->in_len = mem[i].deflate_len;
		d->expect_out = mem[i].isize;
		d->expect_crc = mem[i].crc32;
		d->method = B2I_METHOD_DEFLATE;
		d->flags = g->verify ? 0 : B2I_F_NO_CRC;
		d->out_off = out;
		d->out_cap = mem[i].isize;
		out = (out + mem[i].isize + 15) & ~(size_t)15;
		in_used = (size_t)(mem[i].deflate_offset + mem[i].deflate_len + 8);
		m_used = i + 1;
	}
	note_header(g, p, &mem[m_used - 1]);
	if (j->buf == NULL || j->cap < OUT_BLOCK + out + 16) {
		b200_buf_release(j->buf, j->cap);
		j->buf = b200_buf_acquire(OUT_BLOCK + out + 16 + (out >> 3), &j->cap);
		if (j->buf == NULL) {
			j->cap = 0;
			return (fatal(self, g, "Can't allocate data for gzip decompression"));
		}
	}
	j->n = m_used;
	rc = b2i_submit(g->ctx, p, in_used, j->d, m_used, j->buf + OUT_BLOCK, out, &j->job);
	if (rc == B2I_OK) {
		/* A member whose trailer understates its size has to be decoded again with room
		 * (bgzf_collect), after upstream has given its bytes away: large windows are
		 * found in the library's pinned staging, small ones are kept here. */
		j->src = b2i_job_staged_input(j->job);
		if (j->src == NULL && !g->verify) {
			if (j->keep_cap < in_used) {
				free(j->keep);
				j->keep = malloc(in_used);
				j->keep_cap = j->keep ? in_used : 0;
			}
			if (j->keep != NULL) {
				memcpy(j->keep, p, in_used);
				j->src = j->keep;
			}
		}
		rc = b2i_job_wait_input(j->job);
	}
	if (rc != B2I_OK) {
		g->ctx_bad = (rc == B2I_E_CUDA);
		return (fatal(self, g, b2i_last_error(g->ctx)));
	}
	g->qn++;
	g->windows++;
	__archive_read_filter_consume(self->upstream, (int64_t)in_used);
	return (ARCHIVE_OK);
}

/* keep GZ_JOBS windows decoding ahead as long as the chain goes on */
static int
bgzf_top_up(struct archive_read_filter *self)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;

	while (g->qn < GZ_JOBS && g->pending_fail == NULL) {
		const unsigned char *p;
		b2i_gzip_member *mem;
		size_t n;
		int rc;
		if (bgzf_scan(self->upstream, &p, &mem, &n) != 0)
			return (fatal(self, g, "Out of memory"));
		if (n == 0) {
			b2i_free(mem);
			break;
		}
		rc = bgzf_submit(self, p, mem, n);
		b2i_free(mem);
		if (rc != ARCHIVE_OK)
			return (rc);
	}
	return (ARCHIVE_OK);
}

/* Members of a finished window overflowed the size their trailers announced: decode those
 * again alone with room (deflate cannot expand by more than 1032:1) and lay the window out
 * anew, every member at an offset that fits what it really produced. */
static int
bgzf_redo_window(struct gz_b200 *g, struct gz_job *j)
{
	unsigned char **alt = calloc(j->n, sizeof(*alt)), *nb;
	size_t i, total = 0, ncap = 0, o = 0;
	int rc = ARCHIVE_OK;

	if (alt == NULL)
		return (ARCHIVE_FATAL);
	for (i = 0; i < j->n; i++) {
		b2i_stream_result *r = &j->r[i];
		const size_t most = (size_t)j->d[i].in_len * 1032u + 65536u;
		size_t cap = (size_t)j->d[i].out_cap;

		while (r->status == B2I_S_OUT_OVERFLOW && cap < most) {
			b2i_stream_desc d = j->d[i];
			cap = cap < 16384 ? 65536 : cap * 4;
			if (cap > most)
				cap = most;
			b200_buf_release_tagged(alt[i]);
			if ((alt[i] = b200_buf_acquire_tagged(cap + 16)) == NULL) {
				rc = ARCHIVE_FATAL;
				break;
			}
			d.in_off = 0;
			d.out_off = 0;
			d.out_cap = cap;
			if (b2i_decode_host(g->ctx, j->src + j->d[i].in_off, (size_t)d.in_len, &d, 1, alt[i], cap, r) != B2I_OK) {
				r->status = B2I_S_OUT_OVERFLOW;      /* reported as a failed member, nothing of it served */
				r->out_bytes = 0;
				break;
			}
		}
		total += ((size_t)r->out_bytes + 15) & ~(size_t)15;
	}
	nb = rc == ARCHIVE_OK ? b200_buf_acquire(OUT_BLOCK + total + 16, &ncap) : NULL;
	if (nb == NULL)
		rc = ARCHIVE_FATAL;
	for (i = 0; i < j->n; i++) {
		if (rc == ARCHIVE_OK) {
			memcpy(nb + OUT_BLOCK + o, alt[i] ? alt[i] : j->buf + OUT_BLOCK + j->d[i].out_off,
			    (size_t)j->r[i].out_bytes);
			j->d[i].out_off = o;
			o += ((size_t)j->r[i].out_bytes + 15) & ~(size_t)15;
		}
		b200_buf_release_tagged(alt[i]);
	}
	free(alt);
	if (rc == ARCHIVE_OK) {
		b200_buf_release(j->buf, j->cap);
		j->buf = nb;
		j->cap = ncap;
	}
	return (rc);
}

/* wait for the oldest window in flight, make its buffer the one being served (the few bytes
 * still pending move in front of it) and queue the windows after it */
static int
bgzf_collect(struct archive_read_filter *self)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;
	struct gz_job *j = &g->q[g->qh];
	unsigned char *dst = j->buf + OUT_BLOCK, *t;
	size_t pending = g->out_len - g->served, w = 0, i, tc;
	int rc;

	rc = b2i_wait(j->job, j->r);
	j->job = NULL;
	g->qh = (g->qh + 1) % GZ_JOBS;
	g->qn--;
	if (rc != B2I_OK) {
		g->ctx_bad = (rc == B2I_E_CUDA);
		return (fatal(self, g, b2i_last_error(g->ctx)));
	}
	/* The reference never looks at ISIZE (gzip.c:423, "XXX TODO: Verify the length and
	 * CRC"): a member that produced more than its trailer says is decoded again with
	 * room, and the window is put together anew (unless verification was asked for). */
	if (!g->verify && j->src != NULL)
		for (i = 0; i < j->n; i++)
			if (j->r[i].status == B2I_S_OUT_OVERFLOW) {
				if ((rc = bgzf_redo_window(g, j)) != ARCHIVE_OK)
					return (fatal(self, g, "Can't allocate data for gzip decompression"));
				dst = j->buf + OUT_BLOCK;
				break;
			}
	/* members are served back to back: close the 16-byte alignment gaps */
	for (i = 0; i < j->n; i++) {
		const b2i_stream_desc *d = &j->d[i];
		const b2i_stream_result *r = &j->r[i];
		int bad = r->status != B2I_S_OK || (r->flags & B2I_R_IN_MISMATCH) ||
		    (g->verify && (r->flags & (B2I_R_CRC_MISMATCH | B2I_R_OUT_MISMATCH)));
		if (w != d->out_off)
			memmove(dst + w, dst + d->out_off, (size_t)r->out_bytes);
		w += (size_t)r->out_bytes;
		if (bad) {
			/* the reference has already handed out the full blocks before the bad
			 * spot when zlib reports it; the partial block is dropped */
			g->pending_fail = "gzip decompression failed";
			break;
		}
	}
	/* next_window is only called with less than one block pending */
	if (pending > OUT_BLOCK)
		return (fatal(self, g, "internal error: gzip window"));
	if (pending)
		memcpy(dst - pending, g->out + g->served, pending);
	t = g->out; g->out = j->buf; j->buf = t;
	tc = g->out_cap; g->out_cap = j->cap; j->cap = tc;
	g->served = OUT_BLOCK - pending;
	g->out_len = OUT_BLOCK + w;
	if (g->pending_fail != NULL) {
		/* what was decoding behind a bad member is dropped */
		while (g->qn > 0) {
			struct gz_job *k = &g->q[g->qh];
			(void)b2i_wait(k->job, k->r);
			k->job = NULL;
			g->qh = (g->qh + 1) % GZ_JOBS;
			g->qn--;
		}
		return (ARCHIVE_OK);
	}
	return (bgzf_top_up(self));
}

/* decode the next window of members and append the bytes to g->out */
static int
next_window(struct archive_read_filter *self)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;
	struct archive_read_filter *up = self->upstream;
	const unsigned char *p;
	ssize_t avail;
	b2i_gzip_member *mem = NULL;
	size_t n = 0, end = 0;

	if (g->qn > 0)
		return (bgzf_collect(self));       /* the oldest of the windows that were decoding while we served */

	/* a memory source answers with everything it has; a file source (64 KiB read blocks,
	 * archive_read_open_filename.c:389-461) is asked to collect a window's worth first */
	p = __archive_read_filter_ahead(up, 18, &avail);
	if (p != NULL && avail < (ssize_t)WINDOW_READ && p[0] == 0x1f && p[1] == 0x8b && (p[3] & 4) &&
	    p[12] == 'B' && p[13] == 'C') {
		/* a BGZF chain (only then: a plain member is decoded from its head on demand) */
		(void)__archive_read_filter_ahead(up, WINDOW_READ, &avail);    /* NULL: fewer bytes are left */
	}
	p = __archive_read_filter_ahead(up, 1, &avail);
	if (p == NULL || avail <= 0) {
		g->eof = 1;
		return (ARCHIVE_OK);
	}
	/* ---- BGZF chain: every complete member in the read-ahead buffer ---- */
	if (b2i_gzip_scan_bgzf(p, (size_t)avail, 0, &mem, &n, &end) != B2I_OK)
		return (fatal(self, g, "Out of memory"));
	if (n == 0) {
		/* maybe the first member is only partly buffered: ask for all of it */
		b2i_gzip_member m;
		size_t hl = peek_header(up, &m);
		b2i_free(mem);
		mem = NULL;
		if (hl == 0) {
			g->eof = 1;           /* trailing garbage: silent end (gzip.c:450-454) */
			return (ARCHIVE_OK);
		}
		p = __archive_read_filter_ahead(up, 1, &avail);
		if (hl >= 18 && p != NULL) {
			/* BSIZE known but the member ran past the buffer? */
			const unsigned char *q = p;
			if ((q[3] & 4) && q[12] == 'B' && q[13] == 'C') {
				size_t total = (size_t)(q[16] | (q[17] << 8)) + 1;
				const void *pp = __archive_read_filter_ahead(up, total, &avail);
				if (pp == NULL)
					return (fatal(self, g, "truncated gzip input"));
				p = pp;
				if (b2i_gzip_scan_bgzf(p, (size_t)avail, 0, &mem, &n, &end) != B2I_OK)
					return (fatal(self, g, "Out of memory"));
			}
		}
	}
	if (g->ctx == NULL) {
		int rc = b200_ctx_acquire(&g->ctx);
		if (rc != B2I_OK) {
			b2i_free(mem);
			return (fatal(self, g, "No usable B200 device; this build has no CPU inflate"));
		}
	}
	if (n > 0) {
		int r = bgzf_submit(self, p, mem, n);
		b2i_free(mem);
		if (r != ARCHIVE_OK)
			return (r);
		return (bgzf_collect(self));
	}
	b2i_free(mem);

	/* ---- a member without BSIZE ----
	 * The decoder is given what is buffered and an output budget; the bytes it
	 * consumed locate the trailer when the stream ends inside the budget.  The
	 * budget starts at one block and quadruples every time more of the same member
	 * is asked for (each pass decodes from the member's start and drops what was
	 * already served), so looking at the head of a huge member is cheap, input is
	 * only requested when the decoder really ran out of it, and the total work stays
	 * within a small factor of one full decode. */
	{
		b2i_gzip_member m;
		b2i_stream_desc d;
		b2i_stream_result r;
		size_t hl = peek_header(up, &m), want, budget, fresh, in_lim, in_use;
		int rc, exhausted = 0;

		if (hl == 0) {
			g->eof = 1;
			return (ARCHIVE_OK);
		}
		p = __archive_read_filter_ahead(up, hl, &avail);
		if (p == NULL)
			return (fatal(self, g, "truncated gzip input"));
		note_header(g, p, &m);
		{
			/* a match is never written partially: one maximal match of slack fills the block */
			size_t pending = g->out_len - g->served;
			budget = g->member_skip ? g->member_skip * 4 :
			    (pending < OUT_BLOCK ? OUT_BLOCK - pending : 1) + 258;
		}
		/* deflate expands by at most a few bytes per 64 KiB: this much input fills the
		 * budget unless the stream idles in empty blocks (then the limit doubles) */
		in_lim = hl + budget + (budget >> 8) + 4096;
		want = (size_t)avail;
		for (;;) {
			if ((size_t)avail <= hl || want > (size_t)avail) {
				const void *pp = __archive_read_filter_ahead(up, want > hl ? want : hl + 1, &avail);
				if (pp == NULL) {
					pp = __archive_read_filter_ahead(up, 1, &avail);
					exhausted = 1;
					if (pp == NULL || (size_t)avail <= hl)
						return (fatal(self, g, "truncated gzip input"));
				}
				p = pp;
			}
			memset(&d, 0, sizeof(d));
			d.in_off = hl;
			in_use = (size_t)avail < in_lim ? (size_t)avail : in_lim;
			d.in_len = in_use - hl;
			d.method = B2I_METHOD_DEFLATE;
			d.flags = g->verify ? 0 : B2I_F_NO_CRC;
			d.out_cap = budget;
			if (make_room(g, budget) != 0)
				return (fatal(self, g, "Can't allocate data for gzip decompression"));
			rc = b2i_decode_host(g->ctx, p, in_use, &d, 1, g->out + g->out_len, budget, &r);
			if (rc != B2I_OK)
				{ g->ctx_bad = (rc == B2I_E_CUDA); return (fatal(self, g, b2i_last_error(g->ctx))); }
			if (r.status == B2I_S_BUF_ERROR && in_use < (size_t)avail) {
				in_lim *= 2;                       /* more of what is buffered */
				continue;
			}
			if (r.status == B2I_S_BUF_ERROR && !exhausted) {
				want = (size_t)avail * 2;          /* the member continues past what is buffered */
				in_lim *= 2;
				continue;
			}
			if (r.status == B2I_S_OUT_OVERFLOW && r.out_bytes <= g->member_skip) {
				budget *= 4;                       /* cannot happen with a growing budget; be safe */
				continue;
			}
			break;
		}
		/* what earlier passes already served is not served twice */
		fresh = (size_t)r.out_bytes > g->member_skip ? (size_t)r.out_bytes - g->member_skip : 0;
		if (g->member_skip && fresh)
			memmove(g->out + g->out_len, g->out + g->out_len + g->member_skip, fresh);
		g->out_len += fresh;
		if (r.status == B2I_S_OUT_OVERFLOW) {
			g->member_skip = (size_t)r.out_bytes;  /* nothing consumed: the member stays at the read position */
			return (ARCHIVE_OK);
		}
		g->member_skip = 0;
		if (r.status != B2I_S_OK) {
			g->pending_fail = r.status == B2I_S_BUF_ERROR ? "truncated gzip input" :
			    "gzip decompression failed";
			return (ARCHIVE_OK);
		}
		{
			size_t trailer = hl + (size_t)r.in_bytes;
			if ((size_t)avail < trailer + 8) {
				const void *pp = __archive_read_filter_ahead(up, trailer + 8, &avail);
				if (pp == NULL) {
					g->failed = 1;      /* consume_trailer: fewer than 8 bytes (gzip.c:418-420) */
					return (ARCHIVE_FATAL);
				}
				p = pp;
			}
			if (g->verify) {
				uint32_t crc = (uint32_t)p[trailer] | (uint32_t)p[trailer + 1] << 8 |
				    (uint32_t)p[trailer + 2] << 16 | (uint32_t)p[trailer + 3] << 24;
				uint32_t isz = (uint32_t)p[trailer + 4] | (uint32_t)p[trailer + 5] << 8 |
				    (uint32_t)p[trailer + 6] << 16 | (uint32_t)p[trailer + 7] << 24;
				if (crc != r.crc || isz != (uint32_t)(r.out_bytes & 0xffffffffu)) {
					g->pending_fail = "gzip decompression failed";
					return (ARCHIVE_OK);
				}
			}
			__archive_read_filter_consume(up, (int64_t)(trailer + 8));
		}
	}
	return (ARCHIVE_OK);
}

static ssize_t
gz_read(struct archive_read_filter *self, const void **p)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;
	size_t n;

	if (g->failed)
		return (ARCHIVE_FATAL);
	/* like the reference, fill a 64 KiB block across member boundaries: keep
	 * decoding windows until that much is pending or the input ends */
	while (g->out_len - g->served < OUT_BLOCK && !g->eof && g->pending_fail == NULL) {
		int r = next_window(self);
		if (r != ARCHIVE_OK)
			return (r);
	}
	n = g->out_len - g->served;
	if (n > OUT_BLOCK)
		n = OUT_BLOCK;
	if (n < OUT_BLOCK && g->pending_fail != NULL)
		return (fatal(self, g, g->pending_fail));   /* the block the error fell into is not delivered */
	*p = n ? g->out + g->served : NULL;
	g->served += n;
	return ((ssize_t)n);
}

static int
gz_close(struct archive_read_filter *self)
{
	struct gz_b200 *g = (struct gz_b200 *)self->data;

	while (g->qn > 0) {                     /* nothing may still be writing into the buffers */
		struct gz_job *k = &g->q[g->qh];
		(void)b2i_wait(k->job, k->r);
		g->qh = (g->qh + 1) % GZ_JOBS;
		g->qn--;
	}
	for (int i = 0; i < GZ_JOBS; i++) {
		b200_buf_release(g->q[i].buf, g->q[i].cap);
		free(g->q[i].d);
		free(g->q[i].r);
		free(g->q[i].keep);
	}
	b200_buf_release(g->out, g->out_cap);
	b200_ctx_release(g->ctx, !g->ctx_bad);
	free(g->name);
	free(g);
	return (ARCHIVE_OK);
}

static const struct archive_read_filter_vtable gz_reader_vtable = {
	.read = gz_read,
	.close = gz_close,
	.read_header = gz_read_header,
};

static int
gz_init(struct archive_read_filter *self)
{
	struct gz_b200 *g;

	self->code = ARCHIVE_FILTER_GZIP;
	self->name = "gzip";
	g = calloc(1, sizeof(*g));
	if (g == NULL) {
		archive_set_error(&self->archive->archive, ENOMEM,
		    "Can't allocate data for gzip decompression");
		return (ARCHIVE_FATAL);
	}
	g->verify = getenv("B2I_GZIP_VERIFY") != NULL;
	self->data = g;
	self->vtable = &gz_reader_vtable;
	return (ARCHIVE_OK);
}
