/*
 * inflate_core.cuh — one warp decodes one raw deflate stream (RFC 1951).
 *
 * Replaces zlib inflate() as the reference calls it per ZIP entry
 * (archive_read_support_format_zip.c:2510-2533 init, :2643 inflate, :2647-2665
 * result handling) and per gzip member (archive_read_support_filter_gzip.c:363,
 * :479); accept/reject behaviour follows zlib 1.3.
 * From-scratch design, not a port of zlib:
 *
 *   - the compressed bytes are staged into a per-warp shared-memory ring by
 *     1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), two 128-byte
 *     halves, the next half always in flight;
 *   - a 64-bit bit buffer refilled 32 bits at a time from the ring, kept
 *     pre-shifted by 2 so that (low word & 0xFFC) is the byte offset of the
 *     primary table entry;
 *   - canonical-Huffman lookup tables built by the whole warp in shared
 *     memory: lit/len 2^10 primary + secondary (<= 1334 entries), distance 2^8
 *     primary + secondary (<= 400 entries); bounds are the exact maxima over
 *     all complete codes;
 *   - symbols are decoded warp-uniformly in batches of up to 32; lane k keeps
 *     symbol k in a register.  The batch is then resolved by all lanes, one
 *     output byte per lane: a shuffle binary search over the prefix sums finds
 *     the owning symbol, literals and bytes fetched from already flushed output
 *     (four loads in flight per lane) go into a shared-memory staging buffer,
 *     the few bytes whose source lies in the same batch are patched from the
 *     staging buffer, and the batch leaves as aligned 16-byte stores;
 *   - stored blocks are copied global->global by the warp;
 *   - the CRC-32 of the produced bytes is computed by the same warp right
 *     after the last block (crc32_core.cuh), while they are still in L2.
 *
 * All integer arithmetic; no tensor cores (not a contraction).
 */
#pragma once
#include "b2i_common.cuh"
#include "crc32_core.cuh"

/* Two builds of the single-warp kernel (b2i_kernels.cu is compiled twice):
 *   default  10-bit lit/len root, 7 CTAs x 4 warps per SM - fastest per stream;
 *   B2I_R9    9-bit root (smaller tables), 8 CTAs per SM at 64 registers - more
 *             resident warps, which pays when a launch has several waves of streams. */
#ifdef B2I_R9
#define LIT_ROOT    9
#define LIT_TABLE   856    /* >= 852 = max over complete 286-symbol codes, root 9 (the fixed code needs 512) */
#else
#define LIT_ROOT    10
#define LIT_TABLE   1336   /* >= 1334 = max over complete 288-symbol codes   */
#endif
#define DIST_ROOT   8
#define CL_ROOT     7
#define DIST_TABLE  400    /* max over complete 30/32-symbol codes, root 8   */
#define RING_HALF   128
#define RING_BYTES  (2 * RING_HALF)

/*
 * Table entries (32 bit).  Common: [4:0] bits to drop (code + extra),
 * bit 5 LIT, bit 6 SLOW.
 *   literal  (lit/len table) : LIT,  [24:8] = 0x10000 | byte   (= the batch record)
 *   length   (lit/len table) : [12:8] code length + 2, [23:16] extra-bit mask,
 *                              [31:24] base - 3
 *   distance (dist table)    : [12:8] code length + 2, [16:13] extra-bit count,
 *                              [31:17] base
 *   code-length symbol       : LIT,  [31:8] symbol
 *   SLOW kinds [13:12]       : 0 sub-table ([11:8] index bits, [31:16] offset),
 *                              1 end of block, 2 invalid code
 * "code length + 2" because the bit buffer is kept shifted left by 2.
 */
#define E_LIT        0x20u
#define E_SLOW       0x40u
#define E_TOTAL(e)   ((e) & 31u)
#define SK_SUB       0u
#define SK_EOB       1u
#define SK_INV       2u
#define E_SLOWKIND(e) (((e) >> 12) & 3u)
#define E_SUBBITS(e) (((e) >> 8) & 15u)
#define E_SUBOFF(e)  ((e) >> 16)
#define ENTRY_INV(total)      ((uint32_t)(total) | E_SLOW | (SK_INV << 12))
#define ENTRY_EOB(total)      ((uint32_t)(total) | E_SLOW | (SK_EOB << 12))
#define ENTRY_SUB(bits, off)  (E_SLOW | (SK_SUB << 12) | ((uint32_t)(bits) << 8) | ((uint32_t)(off) << 16))

/* Output staging: a batch of symbols is assembled here and leaves for global
 * memory as aligned 16-byte stores.  It shares its space with the block-header
 * scratch (code lengths, counters), which is only live while a header is parsed. */
#define STAGE_BYTES 704
#define STAGE_WORDS (STAGE_BYTES / 32)   /* words of the one-bit-per-byte owner bitmap */
#define BATCH_SOFT  431   /* keep decoding while the batch holds <= this many
                             bytes: 15 (carry) + 431 + 258 (one more match) = 704 */
struct HeaderScratch {
	uint8_t  lens[320];
	uint16_t cnt[16];
	uint16_t run[16];
	uint8_t  cl[32];
};
struct __align__(16) WarpSmem {
	uint32_t lit[LIT_TABLE];
	uint32_t dist[DIST_TABLE];
	uint8_t  ring[RING_BYTES];
	union {
		HeaderScratch h;
		uint8_t stage[STAGE_BYTES];
	} u;
	unsigned long long mbar[2];
	uint32_t bm[STAGE_WORDS];   /* batch resolution: bit s set = a symbol starts at staging byte s */
	uint8_t  wp[24];            /* symbols that start in earlier bitmap words */
};

/* ------------------------------------------------------------------------ */
/* small PTX helpers                                                        */
/* ------------------------------------------------------------------------ */
#ifndef B2I_HOST_EMUL
/* low 32 bits of (hi:lo) >> (n mod 32) */
B2I_DEV uint32_t shf_r_wrap(uint32_t lo, uint32_t hi, uint32_t n)
{
	uint32_t r;
	asm("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(n));
	return r;
}
B2I_DEV uint32_t byte2(uint32_t e) { return __byte_perm(e, 0, 0x4442); }
/* v << s for s in 0..32 (0 when s == 32) */
B2I_DEV uint32_t shl_clamp(uint32_t v, uint32_t s)
{
	uint32_t r;
	asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
	return r;
}
/* the bits of v that a left shift by s (0..32) pushes past bit 31 */
B2I_DEV uint32_t funnel_hi(uint32_t v, uint32_t s)
{
	uint32_t r;
	asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(0), "r"(s));
	return r;
}
#else
B2I_DEV uint32_t shf_r_wrap(uint32_t lo, uint32_t hi, uint32_t n)
{
	return (uint32_t)((((uint64_t)hi << 32) | lo) >> (n & 31));
}
B2I_DEV uint32_t byte2(uint32_t e) { return (e >> 16) & 0xffu; }
B2I_DEV uint32_t shl_clamp(uint32_t v, uint32_t s) { return s >= 32 ? 0 : v << s; }
B2I_DEV uint32_t funnel_hi(uint32_t v, uint32_t s) { return s == 0 ? 0 : (s >= 32 ? v : v >> (32 - s)); }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
#endif

/* ------------------------------------------------------------------------ */
/* input ring                                                               */
/* ------------------------------------------------------------------------ */
struct Ring {
	const uint8_t *gbase;   /* 16-byte aligned global address of segment 0 */
	uint64_t glimit;        /* bytes available from gbase (multiple of 16) */
	uint32_t next_issue;    /* next segment to request                     */
	uint32_t state;         /* per half h: bit h = fills issued mod 2, bit 2+h =
	                           latest fill has landed, bit 4+h = ever filled.
	                           Lives as long as the warp (mbarrier phases
	                           persist across streams). */
};
#define RING_PAR(h)    (1u << (h))
#define RING_WAITED(h) (4u << (h))
#define RING_USED(h)   (16u << (h))

#ifndef B2I_HOST_EMUL
B2I_DEV uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
B2I_DEV void mbar_init(unsigned long long *bar)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(bar)) : "memory");
}
B2I_DEV void ring_hw_issue(WarpSmem *sm, int h, const uint8_t *src, uint32_t bytes)
{
	uint32_t bar = smem_addr(&sm->mbar[h]);
	if (bytes) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		    :: "r"(bar), "r"(bytes) : "memory");
		asm volatile(
		    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		    :: "r"(smem_addr(&sm->ring[h * RING_HALF])), "l"(src), "r"(bytes), "r"(bar) : "memory");
	} else {
		asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
	}
}
B2I_DEV void ring_hw_wait(WarpSmem *sm, int h, uint32_t parity)
{
	uint32_t bar = smem_addr(&sm->mbar[h]);
	uint32_t done;
	do {
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}
#endif

/* once per warp, before the first stream */
B2I_DEV void ring_init(WarpSmem *sm, Ring &r)
{
	r.state = 0;
	r.next_issue = 0;
#ifndef B2I_HOST_EMUL
	if (b2i_lane() == 0) {
		mbar_init(&sm->mbar[0]);
		mbar_init(&sm->mbar[1]);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
#else
	(void)sm;
#endif
	__syncwarp();
}

B2I_DEV void ring_issue(WarpSmem *sm, Ring &r)
{
	const uint32_t seg = r.next_issue++;
	const int h = seg & 1;
	const uint64_t off = (uint64_t)seg * RING_HALF;
	uint32_t bytes = 0;
	if (off < r.glimit)
		bytes = (r.glimit - off) < RING_HALF ? (uint32_t)(r.glimit - off) : RING_HALF;
	__syncwarp();           /* every lane is done reading the half we overwrite */
	r.state ^= RING_PAR(h);
	r.state = (r.state & ~RING_WAITED(h)) | RING_USED(h);
#ifdef B2I_HOST_EMUL
	if (b2i_lane() == 0 && bytes)
		memcpy(&sm->ring[h * RING_HALF], r.gbase + off, bytes);
	__syncwarp();
#else
	if (b2i_lane() == 0)
		ring_hw_issue(sm, h, r.gbase + off, bytes);
#endif
}

B2I_DEV void ring_wait(WarpSmem *sm, Ring &r, int h)
{
	if ((r.state & RING_WAITED(h)) || !(r.state & RING_USED(h)))
		return;
#ifndef B2I_HOST_EMUL
	/* n fills issued -> the latest one completes phase (n-1) & 1 */
	ring_hw_wait(sm, h, ((r.state >> h) & 1u) ^ 1u);
#else
	(void)sm;
#endif
	r.state |= RING_WAITED(h);
}

/* make segment `seg` readable and keep the one after it in flight */
B2I_DEV_NOINLINE void ring_ensure(WarpSmem *sm, Ring &r, uint32_t seg)
{
	if (seg >= r.next_issue || seg + 2 < r.next_issue) {
		/* Not one of the two resident segments (next_issue-2, next_issue-1): a
		 * forward jump past everything requested, or a step back of a few bytes
		 * into a segment whose half has already been refilled (the bit reader
		 * runs up to 8 bytes ahead of the byte position a stored block resumes
		 * at).  Let in-flight copies land first: an mbarrier phase must complete
		 * before it is re-armed. */
		ring_wait(sm, r, 0);
		ring_wait(sm, r, 1);
		r.next_issue = seg;
	}
	while (r.next_issue <= seg + 1)
		ring_issue(sm, r);
	ring_wait(sm, r, seg & 1);
}

/* ------------------------------------------------------------------------ */
/* bit reader: (hi:lo) holds the unread bits shifted left by 2               */
/* ------------------------------------------------------------------------ */
struct Bits {
	uint32_t lo, hi;
	int32_t  cnt;       /* valid bits */
	uint32_t rd;        /* byte offset (from gbase) of the next word to load */
	uint32_t rd_end;    /* lead + in_len: loads at rd >= rd_end are past the stream */
	int32_t  over;      /* buffered bits that lie beyond the end of the stream */
};

/* requires cnt <= 30 */
B2I_DEV void bits_load_word(WarpSmem *sm, Ring &r, Bits &b)
{
	if ((b.rd & (RING_HALF - 1)) == 0)
		ring_ensure(sm, r, b.rd / RING_HALF);
	uint32_t w = *(const uint32_t *)&sm->ring[b.rd & (RING_BYTES - 1)];
	uint64_t W = (((uint64_t)b.hi << 32) | b.lo) | ((uint64_t)w << (b.cnt + 2));
	b.lo = (uint32_t)W;
	b.hi = (uint32_t)(W >> 32);
	b.cnt += 32;
	b.rd += 4;
	if (b.rd > b.rd_end)
		b.over = (int32_t)(b.rd - b.rd_end) * 8;
}

/* afterwards cnt >= 31: enough for any lit/len code + extra (20) or distance
 * code + extra (28) */
B2I_DEV void bits_refill(WarpSmem *sm, Ring &r, Bits &b)
{
	if (b.cnt <= 30)
		bits_load_word(sm, r, b);
}

/* n <= 31 (the shifter takes n mod 32); the two low bits may hold junk afterwards */
B2I_DEV void bits_drop(Bits &b, uint32_t n)
{
	b.lo = shf_r_wrap(b.lo, b.hi, n);
	b.hi = shf_r_wrap(b.hi, 0, n);
	b.cnt -= (int32_t)(n & 31u);
}

/* the next (up to 30) unread bits, right aligned */
B2I_DEV uint32_t bits_peek(const Bits &b) { return shf_r_wrap(b.lo, b.hi, 2); }

B2I_DEV bool bits_exhausted(const Bits &b) { return b.cnt < b.over; }

/* position the reader on absolute byte `pos` (offset from gbase) */
B2I_DEV void bits_seek(WarpSmem *sm, Ring &r, Bits &b, uint32_t pos)
{
	b.lo = b.hi = 0;
	b.cnt = 0;
	b.over = 0;
	b.rd = pos & ~3u;
	ring_ensure(sm, r, b.rd / RING_HALF);
	bits_load_word(sm, r, b);
	bits_drop(b, (pos & 3u) * 8);
	bits_refill(sm, r, b);
}

/* stream-relative position of the next unread bit */
B2I_DEV uint64_t bits_pos(const Bits &b, uint32_t lead)
{
	return ((uint64_t)b.rd - lead) * 8 - (uint64_t)(int64_t)b.cnt;
}

/* ------------------------------------------------------------------------ */
/* table construction                                                       */
/* ------------------------------------------------------------------------ */
#define TB_LIT   0
#define TB_DIST  1
#define TB_CODES 2

B2I_DEV uint32_t make_entry(int type, uint32_t sym, uint32_t L)
{
	if (type == TB_LIT) {
		if (sym < 256)
			return L | E_LIT | ((0x10000u | sym) << 8);
		if (sym == 256)
			return ENTRY_EOB(L);
		if (sym > 285)
			return ENTRY_INV(L);
		uint32_t i = sym - 257, eb, base;
		if (i < 8) { eb = 0; base = 3 + i; }
		else if (i == 28) { eb = 0; base = 258; }
		else { eb = (i >> 2) - 1; base = 3 + ((4 + (i & 3)) << eb); }
		return (L + eb) | ((L + 2) << 8) | (((1u << eb) - 1u) << 16) | ((base - 3) << 24);
	}
	if (type == TB_DIST) {
		if (sym > 29)
			return ENTRY_INV(L);
		uint32_t eb, base;
		if (sym < 4) { eb = 0; base = sym + 1; }
		else { eb = (sym >> 1) - 1; base = 1 + ((2 + (sym & 1)) << eb); }
		return (L + eb) | ((L + 2) << 8) | (eb << 13) | (base << 17);
	}
	return L | E_LIT | (sym << 8);
}

/*
 * Build the lookup table for `nsyms` code lengths at `lens` (shared memory).
 * Returns 0, or 1 for an over-subscribed / disallowed incomplete set
 * (zlib inflate_table's rule: incomplete only if the longest code is 1 bit and
 * the table is not the code-length code).
 */
B2I_DEV_NOINLINE int build_table(WarpSmem *sm, const uint8_t *lens, int nsyms, int type,
    uint32_t *table, int root, int cap)
{
	const unsigned lane = b2i_lane();
	const unsigned lt_mask = (1u << lane) - 1u;
	const int nchunk = (nsyms + 31) >> 5;

	if (lane < 16)
		sm->u.h.cnt[lane] = 0;
	__syncwarp();
	for (int c = 0; c < nchunk; c++) {
		int s = c * 32 + lane;
		unsigned L = s < nsyms ? lens[s] : 0;
		unsigned m = __match_any_sync(B2I_FULL, L);
		if (L && (m & lt_mask) == 0)
			sm->u.h.cnt[L] += (uint16_t)__popc(m);
		__syncwarp();
	}
	unsigned myc = (lane >= 1 && lane < 16) ? sm->u.h.cnt[lane] : 0;
	unsigned used = __ballot_sync(B2I_FULL, myc != 0);
	int maxlen = used ? 31 - __clz(used) : 0;
	unsigned kraft = __reduce_add_sync(B2I_FULL, myc << (15 - (lane & 15)));
	const int size = 1 << root;
	if (maxlen == 0 || (kraft < 32768u && maxlen == 1 && type != TB_CODES)) {
		/* no codes, or a single 1-bit code: unused patterns are 1-bit invalid codes */
		for (int i = lane; i < size; i += 32)
			table[i] = ENTRY_INV(1);
		__syncwarp();
		if (maxlen == 0)
			return 0;
	} else if (kraft != 32768u) {
		return 1;
	}
	/* first canonical code of each length, kept as the running next code */
	if (lane >= 1 && lane < 16) {
		unsigned code = 0;
		for (unsigned j = 1; j < lane; j++)
			code = (code + sm->u.h.cnt[j]) << 1;
		sm->u.h.run[lane] = (uint16_t)code;
	}
	__syncwarp();
	/* secondary tables: one per root-bit prefix that holds longer codes */
	if (maxlen > root) {
		/* 15-bit left-aligned code space; long codes start here */
		const unsigned x0 = (unsigned)sm->u.h.run[root + 1] << (15 - (root + 1));
		const unsigned p0 = x0 >> (15 - root);
		unsigned next_off = (unsigned)size;
		for (unsigned pb = p0; pb < (unsigned)size; pb += 32) {
			unsigned p = pb + lane;
			unsigned bits = 0;
			if (p < (unsigned)size) {
				unsigned x = ((p + 1) << (15 - root)) - 1;   /* last slot of the prefix */
				for (int L = root + 1; L <= 15; L++) {
					unsigned end = ((unsigned)sm->u.h.run[L] + sm->u.h.cnt[L]) << (15 - L);
					if (x < end) { bits = L - root; break; }
				}
			}
			unsigned sz = bits ? (1u << bits) : 0;
			unsigned incl = sz;
			for (int o = 1; o < 32; o <<= 1) {
				unsigned t = __shfl_up_sync(B2I_FULL, incl, o);
				if ((int)lane >= o) incl += t;
			}
			unsigned off = next_off + incl - sz;
			next_off += __shfl_sync(B2I_FULL, incl, 31);
			if (bits && off + sz <= (unsigned)cap)
				table[__brev(p) >> (32 - root)] = ENTRY_SUB(bits, off);
		}
		if (next_off > (unsigned)cap)
			return 1;          /* cannot happen for a complete code (bound proven) */
		__syncwarp();
	}
	/* place every symbol */
	for (int c = 0; c < nchunk; c++) {
		int s = c * 32 + lane;
		unsigned L = s < nsyms ? lens[s] : 0;
		unsigned m = __match_any_sync(B2I_FULL, L);
		unsigned code = L ? (unsigned)sm->u.h.run[L] + __popc(m & lt_mask) : 0;
		__syncwarp();
		if (L && (m & lt_mask) == 0)
			sm->u.h.run[L] += (uint16_t)__popc(m);
		__syncwarp();
		uint32_t e = make_entry(type, (uint32_t)s, L);
		/* codes so short that they own >= 32 primary slots: whole warp fills */
		unsigned wide = __ballot_sync(B2I_FULL, L != 0 && (int)L + 5 <= root);
		while (wide) {
			int src = __ffs(wide) - 1;
			wide &= wide - 1;
			unsigned wl = __shfl_sync(B2I_FULL, L, src);
			unsigned wc = __shfl_sync(B2I_FULL, code, src);
			uint32_t we = __shfl_sync(B2I_FULL, e, src);
			unsigned r = __brev(wc) >> (32 - wl);
			for (unsigned i = r + (lane << wl); i < (unsigned)size; i += 32u << wl)
				table[i] = we;
		}
		if (L != 0 && (int)L + 5 > root) {
			if ((int)L <= root) {
				unsigned r = __brev(code) >> (32 - L);
				for (unsigned i = r; i < (unsigned)size; i += 1u << L)
					table[i] = e;
			} else {
				unsigned p = code >> (L - root);
				uint32_t sub = table[__brev(p) >> (32 - root)];
				unsigned sbits = E_SUBBITS(sub), off = E_SUBOFF(sub);
				unsigned low = code & ((1u << (L - root)) - 1u);
				unsigned r = __brev(low) >> (32 - (L - root));
				for (unsigned i = r; i < (1u << sbits); i += 1u << (L - root)) {
					B2I_CHECK(off + i < (unsigned)cap);
					table[off + i] = e;
				}
			}
		}
	}
	__syncwarp();
	return 0;
}

/* ------------------------------------------------------------------------ */
/* symbol decoding: one batch                                               */
/* ------------------------------------------------------------------------ */

/* second-level lookup for a SUB entry */
B2I_DEV uint32_t lookup_sub(const uint32_t *table, const Bits &b, uint32_t e, int root)
{
	uint32_t idx = shf_r_wrap(b.lo, b.hi, root + 2) & ((1u << E_SUBBITS(e)) - 1u);
	B2I_CHECK(E_SUBOFF(e) + idx < (root == LIT_ROOT ? LIT_TABLE : DIST_TABLE));
	return table[E_SUBOFF(e) + idx];
}

/*
 * Decode symbols until 32 are held, the batch is full, the block ends or an
 * error shows.  Lane k keeps symbol k as (length << 16 | literal-or-distance).
 * CHECK = false is used while the stream has so much input left that no symbol
 * of this batch can run past its end (no per-symbol exhaustion test).
 * Returns 0 go on, 1 end of block, < 0 a B2I status (detail in *detail).
 */
template <bool CHECK>
B2I_DEV int decode_batch(WarpSmem *sm, Ring &r, Bits &b, uint32_t &my, uint32_t &n_out,
    uint32_t &detail)
{
	const unsigned lane = b2i_lane();
	const char *litp = (const char *)sm->lit;
	const char *distp = (const char *)sm->dist;
	uint32_t n = 0, cum = 0;
	int stop = 0;

	my = 0;
	while (n < 32 && cum <= BATCH_SOFT) {
		bits_refill(sm, r, b);
		uint32_t e = *(const uint32_t *)(litp + (b.lo & ((1u << (LIT_ROOT + 2)) - 4u)));
		if (e & E_SLOW) {
			if (E_SLOWKIND(e) == SK_SUB)
				e = lookup_sub(sm->lit, b, e, LIT_ROOT);
			if (e & E_SLOW) {
				bits_drop(b, e);
				if (bits_exhausted(b)) { stop = S_BUF_ERROR; break; }
				if (E_SLOWKIND(e) == SK_EOB) { stop = 1; break; }
				stop = S_DATA_ERROR; detail = D_BAD_LITLEN_CODE;
				break;
			}
		}
		if (e & E_LIT) {
			bits_drop(b, e);
			if (CHECK && bits_exhausted(b)) { stop = S_BUF_ERROR; break; }
			if (lane == n) my = e >> 8;
			n++;
			cum++;
			continue;
		}
		/* length: base - 3 in the top byte, extra bits above the code */
		uint32_t len = (e >> 24) + (shf_r_wrap(b.lo, b.hi, e >> 8) & byte2(e)) + 3u;
		bits_drop(b, e);
		if (CHECK && bits_exhausted(b)) { stop = S_BUF_ERROR; break; }
		bits_refill(sm, r, b);
		uint32_t d = *(const uint32_t *)(distp + (b.lo & ((1u << (DIST_ROOT + 2)) - 4u)));
		if (d & E_SLOW) {
			if (E_SLOWKIND(d) == SK_SUB)
				d = lookup_sub(sm->dist, b, d, DIST_ROOT);
			if (d & E_SLOW) {
				/* zlib judges the code with only its own bits */
				bits_drop(b, d);
				if (bits_exhausted(b)) { stop = S_BUF_ERROR; break; }
				stop = S_DATA_ERROR; detail = D_BAD_DIST_CODE;
				break;
			}
		}
		uint32_t dist = (d >> 17) + (shf_r_wrap(b.lo, b.hi, d >> 8) & ((1u << ((d >> 13) & 15u)) - 1u));
		bits_drop(b, d);
		if (CHECK && bits_exhausted(b)) { stop = S_BUF_ERROR; break; }
		if (lane == n) my = (len << 16) | dist;
		n++;
		cum += len;
	}
	n_out = n;
	return stop;
}

/* ------------------------------------------------------------------------ */
/* batch resolution: symbols -> bytes in global memory                       */
/* ------------------------------------------------------------------------ */
/*
 * Lane k (k < n) holds symbol k as (length << 16 | literal-or-distance).
 * Appends the batch to the output.  `stop`/`stop_detail` are only written when
 * a symbol cannot be produced (distance reaches before the start of the
 * output, or the output capacity is exceeded): the batch is then cut right
 * before that symbol, exactly where zlib would stop.
 */
#ifdef B2I_HOST_EMUL
extern long g_stat_hist[64], g_stat_instage, g_stat_overlap, g_stat_batches, g_stat_bytes, g_stat_syms;
#endif
#ifndef B2I_HOST_EMUL
B2I_DEV uint8_t load_fresh(const uint8_t *p) { return __ldcg(p); }   /* L2: other warps' stores */
B2I_DEV void fence_block() { __threadfence_block(); }
#else
B2I_DEV uint8_t load_fresh(const uint8_t *p) { return *(volatile const uint8_t *)p; }
B2I_DEV void fence_block() { __sync_synchronize(); }
#endif

B2I_DEV uint32_t resolve_batch(WarpSmem *sm, uint8_t *out, uint8_t *mir, uint32_t cap, uint32_t &outp, uint32_t &carry,
    uint32_t my, uint32_t n, int32_t &stop, uint32_t &stop_detail)
{
	const unsigned lane = b2i_lane();
	uint32_t len = lane < n ? my >> 16 : 0;
	const uint32_t val = my & 0xffffu;
	uint32_t incl = len;
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(B2I_FULL, incl, o);
		if ((int)lane >= o) incl += t;
	}
	uint32_t rel = incl - len;                /* offset inside the batch */
	/* take symbols while the batch holds <= BATCH_SOFT bytes (what decode_batch
	 * guarantees by construction; pre-decoded token runs are cut here) */
	n = (uint32_t)__popc(__ballot_sync(B2I_FULL, len != 0 && rel <= BATCH_SOFT));
	if (lane >= n) len = 0;
	const uint32_t taken = n;
	/* first symbol that cannot be written: bad distance or no room */
	bool far = len >= 3 && val > outp + rel;
	bool full = len != 0 && (outp > cap || rel + len > cap - outp);
	unsigned badmask = __ballot_sync(B2I_FULL, far || full);
	if (badmask) {
		int first = __ffs(badmask) - 1;
		bool first_far = __shfl_sync(B2I_FULL, (int)far, first) != 0;
		n = (uint32_t)first;
		if (lane >= n) len = 0;
		stop = first_far ? S_DATA_ERROR : S_OUT_OVERFLOW;
		stop_detail = first_far ? D_DIST_TOO_FAR : 0;
	}
	uint32_t T = __shfl_sync(B2I_FULL, incl, n ? n - 1 : 0);
	if (n == 0) T = 0;
	const uint32_t c = outp & 15u;
	uint8_t *stg = sm->u.stage;
	uint8_t *g16 = out + (outp - c);           /* global address of stg[0] */
	const uint32_t pk = (len << 16) | val;
#ifdef B2I_HOST_EMUL
	if (lane < n) {
		__sync_fetch_and_add(&g_stat_hist[len > 63 ? 63 : len], 1);
		if (len >= 3 && (int)(c + rel) - (int)val + (int)(len < val ? len : val) > 0) __sync_fetch_and_add(&g_stat_instage, 1);
		if (len >= 3 && val < len) __sync_fetch_and_add(&g_stat_overlap, 1);
	}
	if (lane == 0) { g_stat_batches++; g_stat_bytes += T; g_stat_syms += n; }
#endif
	/* Which symbol owns staging byte s?  Every symbol sets the bit of its first
	 * byte in a bitmap; the owner of s is (number of set bits at or below s) - 1,
	 * i.e. one popcount plus a per-word running count. */
	if (lane < STAGE_WORDS)
		sm->bm[lane] = 0;
	if (lane < c)
		stg[lane] = (uint8_t)carry;
	__syncwarp();
	B2I_CHECK(c + T <= STAGE_BYTES);
	if (len != 0)
		atomicOr(&sm->bm[(c + rel) >> 5], 1u << ((c + rel) & 31u));
	__syncwarp();
	{
		uint32_t pc = lane < STAGE_WORDS ? (uint32_t)__popc(sm->bm[lane]) : 0, inc = pc;
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(B2I_FULL, inc, o);
			if ((int)lane >= o) inc += t;
		}
		if (lane < STAGE_WORDS)
			sm->wp[lane] = (uint8_t)(inc - pc);
	}
	__syncwarp();
	/* round 1, one output byte per lane: a literal is stored as is, a match byte
	 * whose source was flushed to global memory before this batch is fetched
	 * (four loads in flight per lane); sources inside the staging buffer wait
	 * for round 2. */
	const int hs = 0;
	for (uint32_t t0 = 0; t0 < T; t0 += 128) {
		uint32_t v[4];
		bool st[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t t = t0 + 32 * k + lane;
			const bool in = t < T;
			const uint32_t si = c + (in ? t : 0);           /* staging index */
			/* set bits at or below si in its bitmap word: shift them to the top and count */
			const uint32_t own = (uint32_t)sm->wp[si >> 5] +
			    (uint32_t)__popc(sm->bm[si >> 5] << (31u - (si & 31u))) - 1u;
			const uint32_t opk = __shfl_sync(B2I_FULL, pk, own);
			const uint32_t ro = __shfl_sync(B2I_FULL, rel, own);
			const uint32_t olen = opk >> 16, oval = opk & 0xffffu;
			st[k] = in;
			v[k] = oval;
			if (in && olen != 1) {
				uint32_t off = t - ro;
				if (oval < olen)
					off %= oval;               /* overlapping copy: period = distance */
				int sidx = (int)(c + ro + off) - (int)oval;
				st[k] = sidx < hs;
				B2I_CHECK(sidx >= -(int)(outp - c) && sidx < (int)STAGE_BYTES);
				if (sidx < hs)
					v[k] = g16[sidx];
			}
		}
#pragma unroll
		for (int k = 0; k < 4; k++)
			if (st[k]) {
				B2I_CHECK(c + t0 + 32 * k + lane < STAGE_BYTES);
				stg[c + t0 + 32 * k + lane] = (uint8_t)v[k];
			}
	}
	__syncwarp();
	/* round 2: bytes whose source is still in the staging buffer (the carry
	 * or this very batch), match by match in stream order */
	unsigned mm = __ballot_sync(B2I_FULL,
	    len >= 3 && c + rel + (len < val ? len : val) > val + (uint32_t)hs);
	while (mm) {
		int src_lane = __ffs(mm) - 1;
		mm &= mm - 1;
		uint32_t mrel = __shfl_sync(B2I_FULL, rel, src_lane);
		uint32_t mpk = __shfl_sync(B2I_FULL, pk, src_lane);
		uint32_t mlen = mpk >> 16, mdist = mpk & 0xffffu;
		for (uint32_t j = lane; j < mlen; j += 32) {
			uint32_t off = mdist < mlen ? j % mdist : j;
			int sidx = (int)(c + mrel + off) - (int)mdist;
			B2I_CHECK(c + mrel + j < STAGE_BYTES && sidx < (int)(c + mrel + j));
			if (sidx >= hs)
				stg[c + mrel + j] = stg[sidx];
		}
		__syncwarp();
	}
	/* flush whole 16-byte units, keep the rest as the new carry */
	{
		const uint32_t fill = c + T, groups = fill >> 4;
		/* at most 44 units: two predicated stores, no loop */
#pragma unroll
		for (uint32_t g = lane; g < 64; g += 32) {
			if (g < groups) {
				const uint4 v16 = *(const uint4 *)(stg + 16 * g);
				*(uint4 *)(g16 + 16 * g) = v16;
				if (mir)            /* host-mapped copy of the output (B2I_MIRROR) */
					*(uint4 *)(mir + (outp - c) + 16 * g) = v16;
			}
		}
		if (lane < (fill & 15u))
			carry = stg[16 * groups + lane];
		__syncwarp();
	}
	outp += T;
	return taken;
}

#include "inflate_lp.cuh"
#include "inflate_team.cuh"

/* ------------------------------------------------------------------------ */
/* one stream                                                               */
/* ------------------------------------------------------------------------ */
struct StreamOut {
	int32_t  status;
	uint32_t detail;
	uint64_t out_bytes;
	uint64_t in_bytes;
};

#define FAIL(st, dt) do { res.status = (st); res.detail = (dt); goto done; } while (0)
#define NEEDOK() do { if (bits_exhausted(b)) FAIL(S_BUF_ERROR, 0); } while (0)

B2I_DEV StreamOut inflate_stream(WarpSmem *sm, Ring &r, uint32_t *scratch, TeamShared *team, const uint8_t *in_base,
    uint64_t in_total, uint64_t in_off, uint64_t in_len, uint8_t *out, uint8_t *mir, uint64_t out_cap)
{
	const unsigned lane = b2i_lane();
	StreamOut res;
	Bits b;
	uint32_t outp = 0;                      /* bytes produced (streams < 4 GiB) */
	/* The last (outp & 15) bytes produced are not in global memory yet: lane i
	 * (i < outp & 15) keeps byte (outp & ~15) + i in `carry`, so that every
	 * store of decoded data is a full, aligned 16-byte unit. */
	uint32_t carry = 0;
	const uint32_t cap = (uint32_t)out_cap;
	const uint32_t lead = (uint32_t)(in_off & 15);
	uint32_t last;

	res.status = S_OK;
	res.detail = 0;
	r.gbase = in_base + (in_off & ~(uint64_t)15);
	r.glimit = ((in_total + 15) & ~(uint64_t)15) - (in_off & ~(uint64_t)15);
	r.next_issue = 0;
	b.rd_end = lead + (uint32_t)in_len;
	bits_seek(sm, r, b, lead);
	if (team)
		team_stream_begin(team, out, mir, cap);
	PH_DECL();

	do {
		uint32_t hdr;
		bits_refill(sm, r, b);
		hdr = bits_peek(b) & 7u;
		bits_drop(b, 3);
		NEEDOK();
		last = hdr & 1u;
		hdr >>= 1;
		if (hdr == 3)
			FAIL(S_DATA_ERROR, D_BAD_BLOCK_TYPE);
		if (hdr == 0) {
			/* ---- stored block: LEN, NLEN at the next byte boundary ---- */
			bits_drop(b, (uint32_t)b.cnt & 7u);
			bits_refill(sm, r, b);
			uint32_t len = bits_peek(b) & 0xffffu;
			bits_drop(b, 16);
			bits_refill(sm, r, b);
			uint32_t nlen = bits_peek(b) & 0xffffu;
			bits_drop(b, 16);
			NEEDOK();
			if (len != (nlen ^ 0xffffu))
				FAIL(S_DATA_ERROR, D_BAD_STORED_LEN);
			uint32_t pos = b.rd - ((uint32_t)b.cnt >> 3);     /* byte offset from gbase */
			uint32_t avail = b.rd_end > pos ? b.rd_end - pos : 0;
			uint32_t ncopy = len < avail ? len : avail;
			if (ncopy > cap - outp)
				FAIL(S_OUT_OVERFLOW, 0);
			const uint8_t *src = r.gbase + pos;
			uint8_t *dst = out + outp;
			if (lane < (outp & 15u)) {
				out[(outp & ~15u) + lane] = (uint8_t)carry;
				if (mir) mir[(outp & ~15u) + lane] = (uint8_t)carry;
			}
			if (team && ncopy >= 4096u) {
				__syncwarp();
				fence_block();
				team_copy_stored(team, src, dst, ncopy);
				fence_block();
			} else {
				for (uint32_t i = lane; i < ncopy; i += 32) {
					const uint8_t v8 = src[i];
					dst[i] = v8;
					if (mir) mir[outp + i] = v8;
				}
			}
			__syncwarp();
			outp += ncopy;
			if (lane < (outp & 15u))
				carry = load_fresh(out + (outp & ~15u) + lane);
			bits_seek(sm, r, b, pos + ncopy);
			if (ncopy < len) {
				b.over = 1 << 20;  /* input ran out inside the block */
				FAIL(S_BUF_ERROR, 0);
			}
			PH_ADD(PH_STORED);
			continue;
		}
		if (hdr == 1) {
			/* ---- fixed Huffman: lengths from RFC 1951 3.2.6 ---- */
			for (int i = lane; i < 288; i += 32)
				sm->u.h.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
			sm->u.h.lens[288 + lane] = 5;
			__syncwarp();
			build_table(sm, sm->u.h.lens, 288, TB_LIT, sm->lit, LIT_ROOT, LIT_TABLE);
			build_table(sm, sm->u.h.lens + 288, 32, TB_DIST, sm->dist, DIST_ROOT, DIST_TABLE);
		} else {
			/* ---- dynamic Huffman header ---- */
			bits_refill(sm, r, b);
			uint32_t h14 = bits_peek(b);
			uint32_t nlen = (h14 & 31u) + 257;
			uint32_t ndist = ((h14 >> 5) & 31u) + 1;
			uint32_t ncode = ((h14 >> 10) & 15u) + 4;
			bits_drop(b, 14);
			NEEDOK();
			if (nlen > 286 || ndist > 30)
				FAIL(S_DATA_ERROR, D_TOO_MANY_SYMS);
			sm->u.h.cl[lane] = 0;
			__syncwarp();
			for (uint32_t i = 0; i < ncode; i++) {
				/* order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, packed 5 bits each */
				const uint64_t ord_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 |
				    7ull << 25 | 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
				const uint64_t ord_hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 |
				    1ull << 25 | 15ull << 30;
				uint32_t sym = i < 12 ? (uint32_t)(ord_lo >> (5 * i)) & 31u
				                      : (uint32_t)(ord_hi >> (5 * (i - 12))) & 31u;
				bits_refill(sm, r, b);
				if (lane == 0)
					sm->u.h.cl[sym] = (uint8_t)(bits_peek(b) & 7u);
				bits_drop(b, 3);
				NEEDOK();
			}
			__syncwarp();
			if (build_table(sm, sm->u.h.cl, 19, TB_CODES, sm->dist, CL_ROOT, DIST_TABLE))
				FAIL(S_DATA_ERROR, D_BAD_CODELEN_SET);
			uint32_t have = 0, prev = 0;
			const uint32_t want = nlen + ndist;
			while (have < want) {
				bits_refill(sm, r, b);
				uint32_t e = sm->dist[bits_peek(b) & ((1u << CL_ROOT) - 1u)];
				uint32_t sym = e >> 8;
				if (e & E_SLOW)
					sym = 0;    /* zlib quirk: empty code-length code reads as length 0, 1 bit */
				bits_drop(b, e);
				NEEDOK();
				if (sym < 16) {
					if (lane == 0)
						sm->u.h.lens[have] = (uint8_t)sym;
					prev = sym;
					have++;
					continue;
				}
				uint32_t copy, fill;
				uint32_t x = bits_peek(b);
				if (sym == 16) {
					copy = 3 + (x & 3u);
					bits_drop(b, 2);
					NEEDOK();
					if (have == 0)
						FAIL(S_DATA_ERROR, D_BAD_BITLEN_REPEAT);
					fill = prev;
				} else if (sym == 17) {
					copy = 3 + (x & 7u);
					bits_drop(b, 3);
					NEEDOK();
					fill = 0;
				} else {
					copy = 11 + (x & 127u);
					bits_drop(b, 7);
					NEEDOK();
					fill = 0;
				}
				if (have + copy > want)
					FAIL(S_DATA_ERROR, D_BAD_BITLEN_REPEAT);
				for (uint32_t i = lane; i < copy; i += 32)
					sm->u.h.lens[have + i] = (uint8_t)fill;
				prev = fill;
				have += copy;
			}
			__syncwarp();
			if (sm->u.h.lens[256] == 0)
				FAIL(S_DATA_ERROR, D_NO_EOB);
			if (build_table(sm, sm->u.h.lens, (int)nlen, TB_LIT, sm->lit, LIT_ROOT, LIT_TABLE))
				FAIL(S_DATA_ERROR, D_BAD_LITLEN_SET);
			if (build_table(sm, sm->u.h.lens + nlen, (int)ndist, TB_DIST, sm->dist, DIST_ROOT,
			    DIST_TABLE))
				FAIL(S_DATA_ERROR, D_BAD_DIST_SET);
		}

		PH_ADD(PH_HEADER);
		/* ---- symbols: lane-parallel while the block has enough input left ---- */
		if (scratch) {
			uint64_t P = (uint64_t)b.rd * 8 - (uint64_t)(int64_t)b.cnt;   /* bits from gbase */
			uint32_t det = 0;
			int st = 2;
			if (team)           /* a team of warps shares the rounds of a large stream */
				st = lp_block_team(team, sm, r.gbase, r.glimit, (uint64_t)b.rd_end * 8, P, out, mir,
				    cap, outp, carry, det);
			if (st == 2)
				st = lp_block(sm, r.gbase, r.glimit, (uint64_t)b.rd_end * 8, P, scratch, out, mir, cap,
				    outp, carry, det);
			bits_seek(sm, r, b, (uint32_t)(P >> 3));
			bits_drop(b, (uint32_t)P & 7u);
			PH_ADD(7);
			if (st < 0)
				FAIL(st, det);
			if (st == 0)
				continue;            /* end-of-block consumed: next block header */
		}
		/* ---- symbols, batch after batch (warp-uniform) ---- */
		for (;;) {
			uint32_t my, n, stop_detail = 0;
			int32_t stop;
			/* a batch consumes at most 32 x 48 bits = 192 bytes */
			if (b.rd + 200u <= b.rd_end)
				stop = decode_batch<false>(sm, r, b, my, n, stop_detail);
			else
				stop = decode_batch<true>(sm, r, b, my, n, stop_detail);

			resolve_batch(sm, out, mir, cap, outp, carry, my, n, stop, stop_detail);
			if (stop == 1)
				break;
			if (stop < 0)
				FAIL(stop, stop_detail);
		}
		PH_ADD(PH_UNIF);
	} while (!last);
done:
	/* the carry joins the rest of the output (error paths included: whatever was
	 * produced is in global memory when the warp reports out_bytes) */
	if (lane < (outp & 15u)) {
		out[(outp & ~15u) + lane] = (uint8_t)carry;
		if (mir) mir[(outp & ~15u) + lane] = (uint8_t)carry;
	}
	__syncwarp();
	/* no bulk copy may still be in flight into this warp's ring when the
	 * shared memory is handed to the next stream or the CTA exits */
	ring_wait(sm, r, 0);
	ring_wait(sm, r, 1);
	res.out_bytes = outp;
	res.in_bytes = (bits_pos(b, lead) + 7) >> 3;
	return res;
}
#undef FAIL
#undef NEEDOK
