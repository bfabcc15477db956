/*
 * oracle_inflate.c — TEST INFRASTRUCTURE ONLY (the checker, never the product).
 *
 * CPU restatement of what the reference's hot path computes when it hands a
 * ZIP entry body or a gzip member to zlib:
 *     inflateInit2(&strm, -15); inflate(&strm, 0) ... until Z_STREAM_END
 *   archive_read_support_format_zip.c:2510-2533 (init), :2643 (inflate),
 *     :2647-2657 (Z_STREAM_END / error mapping), :2661-2665 (total_in/out)
 *   archive_read_support_filter_gzip.c:363 (init), :479 (inflate), :483-499
 *
 * The arithmetic itself lives in zlib, which is NOT under /root/reference
 * (system dependency, un-pinned, >= 1.2.1: CMakeLists.txt:449-453; this image
 * has zlib 1.3).  This file restates the published algorithm (RFC 1951) with
 * zlib 1.3's acceptance rules, written from the
 * specification as a deliberately naive bit-at-a-time canonical-Huffman
 * decoder so that it shares no structure with the device kernel's LUT decoder.
 *
 * Pinning: tests/test_oracle.py checks this file against (1) Python's zlib
 * (same zlib 1.3) on every generated case and on a malformed-stream zoo, and
 * (2) oracle/_ref (the unmodified reference compiled from /root/reference)
 * on the reference's own ZIP/gzip fixtures.  Only tests/, smoke() and
 * bench.py's cpu_baseline leg may call into oracle/.
 *
 * Result semantics (what zlib would report after being fed the whole input):
 *   status  0  : final block ended; in_bytes = total_in, out_bytes = total_out
 *   status -3  : Z_DATA_ERROR (detail = which zlib message)
 *   status -5  : Z_BUF_ERROR  (input exhausted first; the reference prints
 *                "ZIP decompression failed (-5)" / "truncated gzip input")
 *   status -100: output capacity exceeded (not a zlib condition; the caller
 *                under-sized the buffer)
 * A syntax element whose bits extend past the input is a BUF error even if
 * its value would also be invalid: zlib only judges bits it has.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include "oracle.h"

typedef struct {
	const uint8_t *in;
	uint64_t nbits;   /* input size in bits */
	uint64_t pos;     /* next bit */
} bitsrc;

/* returns -1 when the bits are not there */
static int
getbits(bitsrc *s, int n, uint32_t *v)
{
	uint32_t r = 0;
	int i;

	if (s->pos + (uint64_t)n > s->nbits)
		return -1;
	for (i = 0; i < n; i++) {
		uint64_t p = s->pos + i;
		r |= (uint32_t)((s->in[p >> 3] >> (p & 7)) & 1) << i;
	}
	s->pos += n;
	*v = r;
	return 0;
}

typedef struct {
	uint16_t count[16];
	uint16_t sym[320];
	int max;          /* longest length in use, 0 = no codes at all */
	int incomplete;   /* single 1-bit code (zlib tolerates it)      */
} hcode;

#define H_OK 0
#define H_BAD 1

enum { T_CODES, T_LENS, T_DISTS };

/* zlib inflate_table()'s accept/reject rule, restated */
static int
build(hcode *h, const uint8_t *lens, int n, int type)
{
	uint16_t offs[16];
	int len, i, left;

	memset(h, 0, sizeof(*h));
	for (i = 0; i < n; i++)
		h->count[lens[i]]++;
	h->count[0] = 0;
	for (len = 15; len >= 1; len--)
		if (h->count[len])
			break;
	h->max = len;
	if (h->max == 0)
		return H_OK;             /* "no symbols to code at all" is legal */
	left = 1;
	for (len = 1; len <= 15; len++) {
		left <<= 1;
		left -= h->count[len];
		if (left < 0)
			return H_BAD;        /* over-subscribed */
	}
	if (left > 0) {
		if (type == T_CODES || h->max != 1)
			return H_BAD;        /* incomplete */
		h->incomplete = 1;
	}
	offs[1] = 0;
	for (len = 1; len < 15; len++)
		offs[len + 1] = offs[len] + h->count[len];
	for (i = 0; i < n; i++)
		if (lens[i])
			h->sym[offs[lens[i]]++] = (uint16_t)i;
	return H_OK;
}

#define DEC_NEED  -1   /* ran out of input */
#define DEC_INVAL -2   /* unassigned code  */

static int
decode(bitsrc *s, const hcode *h)
{
	int code = 0, first = 0, index = 0, len;
	uint32_t b;

	if (h->max == 0 || h->incomplete) {
		/* zlib's table here has 1-bit entries: the single code (if any)
		 * is '0'; everything else is an invalid-code marker of 1 bit. */
		if (getbits(s, 1, &b))
			return DEC_NEED;
		if (h->incomplete && b == 0)
			return h->sym[0];
		return DEC_INVAL;
	}
	for (len = 1; len <= 15; len++) {
		int count;
		if (getbits(s, 1, &b))
			return DEC_NEED;
		code |= (int)b;
		count = h->count[len];
		if (code - count < first)
			return h->sym[index + (code - first)];
		index += count;
		first += count;
		first <<= 1;
		code <<= 1;
	}
	return DEC_INVAL; /* unreachable for complete codes */
}

static const uint16_t len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17,
	19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2,
	2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49,
	65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
	8193, 12289, 16385, 24577 };
static const uint8_t dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5,
	6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

#define FAIL(st, dt) do { r->status = (st); r->detail = (dt); goto done; } while (0)
#define NEED(n, v) do { if (getbits(&s, (n), &(v))) FAIL(ORC_BUF_ERROR, 0); } while (0)

int
orc_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
    orc_result *r)
{
	bitsrc s;
	uint64_t outp = 0;
	uint32_t last, type, v;
	hcode lencode, distcode;
	uint8_t lens[320];

	s.in = in;
	s.nbits = (uint64_t)in_len * 8;
	s.pos = 0;
	memset(r, 0, sizeof(*r));

	do {
		NEED(1, last);
		NEED(2, type);
		if (type == 3)
			FAIL(ORC_DATA_ERROR, ORC_D_BAD_BLOCK_TYPE);
		if (type == 0) {
			uint32_t len, nlen;
			uint64_t byte, avail;

			s.pos = (s.pos + 7) & ~(uint64_t)7;
			NEED(16, len);
			NEED(16, nlen);
			if (len != (nlen ^ 0xffff))
				FAIL(ORC_DATA_ERROR, ORC_D_BAD_STORED_LEN);
			byte = s.pos >> 3;
			avail = in_len - byte;
			if (avail > len)
				avail = len;
			if (outp + avail > out_cap)
				FAIL(ORC_OUT_OVERFLOW, 0);
			memcpy(out + outp, in + byte, (size_t)avail);
			outp += avail;
			s.pos += avail * 8;
			if (avail < len)
				FAIL(ORC_BUF_ERROR, 0);
			continue;
		}
		if (type == 1) {
			int i;
			for (i = 0; i < 144; i++) lens[i] = 8;
			for (; i < 256; i++) lens[i] = 9;
			for (; i < 280; i++) lens[i] = 7;
			for (; i < 288; i++) lens[i] = 8;
			build(&lencode, lens, 288, T_LENS);
			for (i = 0; i < 32; i++) lens[i] = 5;
			build(&distcode, lens, 32, T_DISTS);
		} else {
			static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6,
				10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
			uint32_t nlen, ndist, ncode;
			uint8_t cl[19];
			hcode clcode;
			int have = 0, i;

			NEED(5, nlen);  nlen += 257;
			NEED(5, ndist); ndist += 1;
			NEED(4, ncode); ncode += 4;
			if (nlen > 286 || ndist > 30)
				FAIL(ORC_DATA_ERROR, ORC_D_TOO_MANY_SYMS);
			memset(cl, 0, sizeof(cl));
			for (i = 0; i < (int)ncode; i++) {
				NEED(3, v);
				cl[order[i]] = (uint8_t)v;
			}
			if (build(&clcode, cl, 19, T_CODES) != H_OK)
				FAIL(ORC_DATA_ERROR, ORC_D_BAD_CODELEN_SET);
			while (have < (int)(nlen + ndist)) {
				int sym;
				uint32_t copy, fill;

				if (clcode.max == 0) {
					/* zlib quirk: with no code-length codes at all its
					 * table holds 1-bit "invalid" entries whose value
					 * field is 0, and the CODELENS state only looks at
					 * the value: each length reads as 0 using 1 bit. */
					NEED(1, v);
					sym = 0;
				} else {
					sym = decode(&s, &clcode);
					if (sym == DEC_NEED)
						FAIL(ORC_BUF_ERROR, 0);
				}
				if (sym < 16) {
					lens[have++] = (uint8_t)sym;
					continue;
				}
				if (sym == 16) {
					NEED(2, v);
					if (have == 0)
						FAIL(ORC_DATA_ERROR, ORC_D_BAD_BITLEN_REPEAT);
					fill = lens[have - 1];
					copy = 3 + v;
				} else if (sym == 17) {
					NEED(3, v);
					fill = 0;
					copy = 3 + v;
				} else {
					NEED(7, v);
					fill = 0;
					copy = 11 + v;
				}
				if (have + copy > nlen + ndist)
					FAIL(ORC_DATA_ERROR, ORC_D_BAD_BITLEN_REPEAT);
				while (copy--)
					lens[have++] = (uint8_t)fill;
			}
			if (lens[256] == 0)
				FAIL(ORC_DATA_ERROR, ORC_D_NO_EOB);
			if (build(&lencode, lens, (int)nlen, T_LENS) != H_OK)
				FAIL(ORC_DATA_ERROR, ORC_D_BAD_LITLEN_SET);
			if (build(&distcode, lens + nlen, (int)ndist, T_DISTS) != H_OK)
				FAIL(ORC_DATA_ERROR, ORC_D_BAD_DIST_SET);
		}
		for (;;) {
			int sym = decode(&s, &lencode);
			uint32_t len, dist, eb;
			uint64_t i;

			if (sym == DEC_NEED)
				FAIL(ORC_BUF_ERROR, 0);
			if (sym == DEC_INVAL || sym > 285)
				FAIL(ORC_DATA_ERROR, ORC_D_BAD_LITLEN_CODE);
			if (sym < 256) {
				if (outp >= out_cap)
					FAIL(ORC_OUT_OVERFLOW, 0);
				out[outp++] = (uint8_t)sym;
				continue;
			}
			if (sym == 256)
				break;
			sym -= 257;
			NEED(len_extra[sym], eb);
			len = len_base[sym] + eb;
			sym = decode(&s, &distcode);
			if (sym == DEC_NEED)
				FAIL(ORC_BUF_ERROR, 0);
			if (sym == DEC_INVAL || sym > 29)
				FAIL(ORC_DATA_ERROR, ORC_D_BAD_DIST_CODE);
			NEED(dist_extra[sym], eb);
			dist = dist_base[sym] + eb;
			if (dist > outp)
				FAIL(ORC_DATA_ERROR, ORC_D_DIST_TOO_FAR);
			if (outp + len > out_cap)
				FAIL(ORC_OUT_OVERFLOW, 0);
			for (i = 0; i < len; i++, outp++)   /* byte-serial: overlap semantics */
				out[outp] = out[outp - dist];
		}
	} while (!last);
	r->status = ORC_OK;
done:
	r->out_bytes = outp;
	r->in_bytes = (s.pos + 7) >> 3;
	return r->status;
}
