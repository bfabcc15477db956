/*
 * oracle_crc32.c — TEST INFRASTRUCTURE ONLY.
 *
 * orc_crc32 restates libarchive/archive_crc32.h:43-84, the reference's own
 * drop-in for zlib crc32(): reflected IEEE 802.3 polynomial 0xEDB88320,
 * 256-entry table, pre/post inversion by 0xffffffff, crc32(x, NULL, 0) == 0,
 * chaining by feeding the previous result back in.  (With zlib present the
 * reference calls zlib's crc32 through real_crc32,
 * archive_read_support_format_zip.c:405-409; both compute the same function.)
 *
 * orc_bitcrc32 restates the reference test-suite's independent bit-at-a-time
 * checker, test_utils/test_utils.c:113-139.
 *
 * orc_crc32_combine has NO reference counterpart (the reference never calls
 * crc32_combine); it is validated in tests/test_oracle.py against
 * crc(A||B) computed directly and against Python zlib.crc32.
 */
#include "oracle.h"

#define POLY 0xEDB88320u

static uint32_t table[256];
static int table_ready;

static void
make_table(void)
{
	uint32_t b, c;
	int k;

	for (b = 0; b < 256; b++) {
		c = b;
		for (k = 0; k < 8; k++)
			c = (c & 1) ? (c >> 1) ^ POLY : c >> 1;
		table[b] = c;
	}
	table_ready = 1;
}

uint32_t
orc_crc32(uint32_t crc, const void *buf, size_t len)
{
	const uint8_t *p = buf;

	if (p == NULL)
		return 0;
	if (!table_ready)
		make_table();
	crc ^= 0xffffffffu;
	while (len--)
		crc = table[(crc ^ *p++) & 0xff] ^ (crc >> 8);
	return crc ^ 0xffffffffu;
}

uint32_t
orc_bitcrc32(uint32_t crc, const void *buf, size_t len)
{
	const uint8_t *p = buf;
	int bit;

	crc ^= 0xffffffffu;
	while (len--) {
		uint8_t byte = *p++;
		for (bit = 0; bit < 8; bit++) {
			uint32_t feed = (crc ^ byte) & 1;
			crc >>= 1;
			if (feed)
				crc ^= POLY;
			byte >>= 1;
		}
	}
	return crc ^ 0xffffffffu;
}

/* a(x)*b(x) mod P(x), reflected representation: bit 31 is x^0 */
static uint32_t
mulmod(uint32_t a, uint32_t b)
{
	uint32_t p = 0;
	int i;

	for (i = 0; i < 32; i++) {
		if (a & (0x80000000u >> i))
			p ^= b;
		b = (b & 1) ? (b >> 1) ^ POLY : b >> 1;
	}
	return p;
}

/* x^(8*n) mod P */
static uint32_t
xpow8n(uint64_t n)
{
	uint32_t r = 0x80000000u;       /* x^0 */
	uint32_t sq = 0x00800000u;      /* x^8  (bit 31-8) */

	while (n) {
		if (n & 1)
			r = mulmod(r, sq);
		sq = mulmod(sq, sq);
		n >>= 1;
	}
	return r;
}

uint32_t
orc_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
	return mulmod(xpow8n(len_b), crc_a) ^ crc_b;
}

int
orc_decode_batch(const uint8_t *in, size_t in_bytes, const orc_desc *d, size_t n,
    uint8_t *out, size_t out_bytes, orc_stream_result *res)
{
	size_t i;

	for (i = 0; i < n; i++) {
		orc_stream_result *r = &res[i];
		r->status = 0; r->crc = 0; r->out_bytes = 0; r->in_bytes = 0;
		r->detail = 0; r->flags = 0;
		if (d[i].in_off + d[i].in_len > in_bytes ||
		    d[i].out_off + d[i].out_cap > out_bytes)
			return -1;
		if (d[i].method == 8) {
			orc_result o;
			orc_inflate(in + d[i].in_off, d[i].in_len, out + d[i].out_off,
			    d[i].out_cap, &o);
			r->status = o.status; r->detail = (uint32_t)o.detail;
			r->out_bytes = o.out_bytes; r->in_bytes = o.in_bytes;
			if (!(d[i].flags & 0x02))
				r->crc = orc_crc32(0, out + d[i].out_off, o.out_bytes);
		} else if (d[i].method == 0) {
			const uint8_t *src = in + d[i].in_off;
			if (!(d[i].flags & 0x01)) {
				uint64_t k, m = d[i].in_len < d[i].out_cap ? d[i].in_len : d[i].out_cap;
				if (d[i].in_len > d[i].out_cap) { r->status = ORC_OUT_OVERFLOW; continue; }
				for (k = 0; k < m; k++) out[d[i].out_off + k] = src[k];
			}
			r->out_bytes = r->in_bytes = d[i].in_len;
			if (!(d[i].flags & 0x02))
				r->crc = orc_crc32(0, src, d[i].in_len);
		} else {
			r->status = -101;
			continue;
		}
		if (r->status == 0) {
			if (!(d[i].flags & 0x02) && r->crc != d[i].expect_crc) r->flags |= 1;
			if (r->in_bytes != d[i].in_len) r->flags |= 2;
			if ((r->out_bytes & 0xffffffffu) != (d[i].expect_out & 0xffffffffu)) r->flags |= 4;
		}
	}
	return 0;
}
