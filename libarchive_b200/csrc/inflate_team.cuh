/*
 * inflate_team.cuh — one LARGE deflate stream decoded by a CTA of TEAM_WARPS warps.
 * (included by inflate_core.cuh)
 *
 * The reference streams a large entry serially through zlib, 256 KiB per call
 * (archive_read_support_format_zip.c:2535-2690); a single warp here does the
 * same at a few tens of MB/s, which makes an archive with multi-megabyte entries
 * (BASELINE config 4) as slow as its largest entry.  Here a CTA works on ONE
 * stream and both halves of the work are spread over all of its lanes:
 *
 *   warp 0 owns the stream exactly as in the single-warp kernel (block headers,
 *   table construction, the uniform tail, the final checks) and drives the other
 *   warps through shared memory and a named barrier:
 *
 *   PASS    every lane of the CTA decodes its own segment of the block
 *           (lp_pass, tables read from warp 0's shared memory); warp 0 evaluates
 *           the chain of exits and repeats for the lanes whose start moved,
 *           exactly like the single-warp convergence loop, with TEAM_LANES
 *           segments per round instead of 32.  Tokens go to the warps' scratch.
 *
 *   CHUNK   the tokens become bytes through a window kept IN SHARED MEMORY:
 *           consecutive segments' tokens whose output fits TEAM_CHUNK bytes form
 *           a chunk.  (1) expand: every lane walks the tokens of its own segment
 *           and writes one 16-bit symbol per output byte into the chunk buffer -
 *           a literal, the byte itself when a match source lies before the chunk
 *           (read from the 32 KiB history ring in shared memory), or a POINTER to
 *           the source position inside the chunk.  Pointers always point
 *           backwards.  (2) sweep: every warp walks its slice of the chunk in
 *           order and replaces each pointer by what it points at until a literal
 *           is reached (s <- sym[ptr(s)]; chains across slices are followed, the
 *           invariant "a symbol is the right byte or a pointer to an earlier
 *           position holding the same byte" makes any interleaving correct).
 *           (3) flush: symbols are packed to bytes and leave as aligned 16-byte
 *           stores to global memory and into the history ring.
 *           No step waits for L2: that round trip per 100-byte batch is what
 *           bounds the single-warp resolution.
 *
 *   COPY    stored blocks are copied by all lanes, 16 bytes per lane and step.
 *   CRC     the CRC-32 of the finished stream is computed by all warps (one slice
 *           each, crc_warp_raw0) and merged by warp 0 (crc32_combine arithmetic).
 *
 * zlib's "distance too far back" and the capacity check are applied per token in
 * the expand step; the earliest failing position (in stream order) wins and the
 * output is cut right before that symbol, so results are bit-identical to the
 * single-warp path and to the serial decode.
 */
#pragma once

#ifndef TEAM_WARPS
#define TEAM_WARPS    16                    /* one CTA per SM: all of its warps on ONE stream */
#endif
#define TEAM_LANES    (32 * TEAM_WARPS)
#define TEAM_SEG_MAX  1024u                 /* bits per lane and round, at most */
/* below this the block is finished by warp 0 alone; a round may well be shorter than
 * TEAM_LANES segments (lanes that start past the end of the input drop out at once) */
#define TEAM_MIN_BITS ((uint64_t)32 * LP_SEG_MIN)
#ifndef TEAM_CHUNK
#define TEAM_CHUNK    (3584u * TEAM_WARPS)  /* 16-bit symbols in the chunk buffer (112 KB for 16 warps) */
#endif
#define TEAM_HIST     32768u                /* history ring: the deflate window */
#ifndef TEAM_MARGIN_X2
#define TEAM_MARGIN_X2 2u                  /* pass A starts this many half-segments early */
#endif
#define TEAM_SEG_INIT 288u                  /* first round: assume 3:1 */

#define TC_PASS_A    1u
#define TC_PASS_EMIT 2u
#define TC_CHUNK     3u
#define TC_QUIT      4u
#define TC_COPY      5u
#define TC_CRC       6u

/* a symbol is a byte (< 256) or 256 + the chunk index it points at (TEAM_CHUNK + 256 <= 65536) */
#define TS_IS_PTR(s)  ((s) >= 256u)
#define TS_MAKE(q)    ((q) + 256u)
#define TS_INDEX(s)   ((s) - 256u)
#define TEAM_NOERR   0xffffffffu

struct __align__(16) TeamShared {
	uint16_t sym[TEAM_CHUNK];           /* chunk buffer: index i <-> output byte (abs0 & ~15) + i */
	uint8_t  hist[TEAM_HIST];           /* output bytes [hist_pos - 32768, hist_pos) at (pos & 32767) */
	volatile uint32_t cmd;
	uint32_t wbase, max_word, hard_end, seg, p0, cap, m;
	const uint32_t *gw;
	const WarpSmem *tables;
	uint8_t *out, *mir;
	uint32_t *scratch[TEAM_WARPS];
	uint32_t start[TEAM_LANES], exit_[TEAM_LANES], term[TEAM_LANES], nsym[TEAM_LANES], nbytes[TEAM_LANES];
	uint32_t run[TEAM_LANES];
	uint32_t roff[TEAM_LANES + 1];      /* exclusive prefix of nbytes over the round's regions */
	uint32_t tokpos[TEAM_LANES];        /* tokens of a region already turned into bytes */
	uint32_t rdone[TEAM_LANES];         /* bytes of a region already produced */
	uint32_t ck_g0, ck_g1;              /* the chunk holds regions [g0, g1) ... */
	uint32_t ck_partial;                /* ... or as much of region g0 as fits */
	uint32_t ck_abs0;                   /* output position of the chunk's first byte */
	uint32_t ck_len;                    /* bytes in the chunk */
	uint32_t ck_err;                    /* earliest failure: (chunk byte offset << 1) | overflow */
	uint32_t hist_pos;                  /* TEAM_NOERR: the ring does not mirror the output */
	uint32_t next_seg;                  /* segment size the last round suggests for the next */
	uint32_t stage_words;               /* words of the round's input staged in sym[] (0: read global memory) */
	uint32_t stage_lead;                /* staged word 0 sits this many words into the 16-byte aligned copy */
	uint32_t stage_parity;
	unsigned long long stage_bar;       /* mbarrier of the bulk copy */
	const uint8_t *cp_src;
	uint8_t *cp_dst;
	uint32_t cp_n;
	const uint8_t *crc_p;
	uint64_t crc_n;
	const uint32_t *crc_tab_g, *crc_xp8;
	uint32_t crc_part[TEAM_WARPS];
};

#ifndef B2I_HOST_EMUL
B2I_DEV void team_sync() { asm volatile("bar.sync 1, %0;" :: "n"(TEAM_LANES) : "memory"); }
B2I_DEV uint32_t team_atomic_min(uint32_t *p, uint32_t v) { return atomicMin(p, v); }
B2I_DEV uint32_t pack_even_bytes(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x6420); }
#else
static inline uint32_t team_atomic_min(uint32_t *p, uint32_t v)
{
	uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
		;
	return old;
}
static inline uint32_t pack_even_bytes(uint32_t a, uint32_t b)
{
	return (a & 0xffu) | ((a >> 8) & 0xff00u) | ((b & 0xffu) << 16) | ((b << 8) & 0xff000000u);
}
#endif

/*
 * The compressed bytes of a round are staged into shared memory (the chunk buffer is
 * idle while the lanes decode) by ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier
 * complete_tx), so that the 256 lanes stream their segments from shared memory instead
 * of chasing 4-byte words through L1/L2.  Word i of the stage is input word
 * min(wbase + i, max_word), exactly what lp_pass would have read from global memory.
 */
B2I_DEV void team_stage_input(TeamShared *ts, unsigned w)
{
	const unsigned tid = 32 * w + b2i_lane();
	const uint32_t nwords = ts->stage_words;
	if (nwords == 0)
		return;
	const uint32_t wbase = ts->wbase, max_word = ts->max_word;
	const uint32_t lead = wbase & 3u;                  /* words in front of wbase in its 16-byte unit */
	const uint32_t first = wbase - lead;               /* multiple of 4 words: 16-byte aligned */
	uint32_t avail = max_word + 1u - first;            /* words that exist from there (multiple of 4) */
	uint32_t want = (nwords + lead + 3u) & ~3u;
	const uint32_t copy = want < avail ? want : avail;
	uint32_t *dst = (uint32_t *)ts->sym;
#ifndef B2I_HOST_EMUL
	if (tid == 0) {
		const uint32_t bar = smem_addr(&ts->stage_bar);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(copy * 4u) : "memory");
		asm volatile(
		    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		    :: "r"(smem_addr(dst)), "l"(ts->gw + first), "r"(copy * 4u), "r"(bar) : "memory");
	}
	{
		const uint32_t bar = smem_addr(&ts->stage_bar);
		const uint32_t parity = ts->stage_parity;
		uint32_t done;
		do {
			asm volatile(
			    "{\n\t.reg .pred p;\n\t"
			    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			    "selp.u32 %0, 1, 0, p;\n\t}"
			    : "=r"(done) : "r"(bar), "r"(parity) : "memory");
		} while (!done);
	}
#else
	for (uint32_t i = tid; i < copy; i += TEAM_LANES)
		dst[i] = ts->gw[first + i];
	team_sync();
#endif
	/* past the end of the input: the last word repeats (lp_pass clamps its index) */
	if (copy < want) {
		team_sync();
		const uint32_t last = dst[copy - 1];
		for (uint32_t i = copy + tid; i < want; i += TEAM_LANES)
			dst[i] = last;
	}
	team_sync();
	if (tid == 0)
		ts->stage_parity ^= 1u;
}

template <bool EMIT>
B2I_DEV void team_do_pass(TeamShared *ts, unsigned w, const WarpSmem *tables)
{
	const unsigned lane = b2i_lane();
	const unsigned gid = 32 * w + lane;
	const bool run = ts->run[gid] != 0;
	LpOut o;
	o.exit = 0; o.nsym = 0; o.term = LT_NONE; o.nbytes = 0;
	const uint32_t nominal_end = ts->p0 + (gid + 1u) * ts->seg;
	uint32_t start = ts->start[gid];
	if (!EMIT && gid != 0) {
		/* pass A only looks for the place where this lane's decoder crosses into the next
		 * segment: starting one margin EARLIER gives it that much more input to fall into
		 * step with the true symbol sequence before it gets there */
		const uint32_t margin = ts->seg * TEAM_MARGIN_X2 / 2u;
		start = start - ts->p0 > margin ? start - margin : ts->p0;
	}
	if (ts->stage_words)
		lp_pass<EMIT, 32>(tables, (const uint32_t *)ts->sym + ts->stage_lead, 0u, ts->stage_words - 1u, run, start,
		    nominal_end, ts->hard_end, ts->scratch[w] + lane, o);
	else
		lp_pass<EMIT, 32>(tables, ts->gw, ts->wbase, ts->max_word, run, start,
		    nominal_end, ts->hard_end, ts->scratch[w] + lane, o);
	if (run) {
		ts->exit_[gid] = o.exit;
		ts->term[gid] = o.term;
		ts->nsym[gid] = o.nsym;
		ts->nbytes[gid] = o.nbytes;
	}
}

/* ---- CHUNK: tokens -> bytes through shared memory (all warps) ------------------- */
B2I_DEV void team_chunk(TeamShared *ts, unsigned w)
{
	const unsigned lane = b2i_lane();
	const unsigned tid = 32 * w + lane;
	const uint32_t abs0 = ts->ck_abs0;
	const uint32_t lead = abs0 & 15u;           /* chunk index of its first byte */
	const uint32_t g0 = ts->ck_g0, g1 = ts->ck_g1;
	const bool partial = ts->ck_partial != 0;
	uint8_t *out = ts->out, *mir = ts->mir;
#if defined(B2I_PHASE_CLOCKS) && !defined(B2I_HOST_EMUL)
	long long tph_ = clock64();
#define TPH(slot) do { long long n_ = clock64(); if (tid == 0) atomicAdd(&g_b2i_phase[slot], (unsigned long long)(n_ - tph_)); tph_ = n_; } while (0)
#else
#define TPH(slot) do { } while (0)
#endif

	/* 0. the history ring must hold the 32 KiB in front of the chunk (it does unless
	 * a stored block or the uniform tail wrote output since the last chunk) */
	if (ts->hist_pos != abs0) {
		const uint32_t lo = abs0 > TEAM_HIST ? abs0 - TEAM_HIST : 0;
		for (uint32_t p = lo + tid; p < abs0; p += TEAM_LANES)
			ts->hist[p & (TEAM_HIST - 1)] = load_fresh(out + p);
	}
	team_sync();
	TPH(PH_TPLAN);

	/* 1. expand: one region per lane, one symbol per lane and step; the warp stays
	 * converged (a lane that fetches a token and a lane in the middle of a match run the
	 * same few predicated instructions), so a step costs one pass over the loop body */
	{
		const uint32_t r = g0 + tid;
		bool active = r < g1;
		const uint32_t *rt = ts->scratch[0];
		uint32_t ntok = 0, j = 0, pos = lead, lim = lead, rem = 0, d = 0;
		/* the next four tokens are kept in registers (token k of a lane sits at [32 k]):
		 * a freshly loaded one is not needed before three others have been consumed */
		uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
		bool cap_bound = false;
		if (active) {
			rt = ts->scratch[r >> 5] + (r & 31u);
			ntok = ts->nsym[r];
			j = ts->tokpos[r];
			pos = lead + (ts->roff[r] + ts->rdone[r]) - (ts->roff[g0] + ts->rdone[g0]);
			/* bytes the chunk may hold: the buffer, or what is left of the output capacity */
			const uint32_t room = ts->cap - abs0;
			cap_bound = room < TEAM_CHUNK - lead;
			lim = lead + (cap_bound ? room : TEAM_CHUNK - lead);
			q0 = j < ntok ? rt[32u * j] : 0;
			q1 = j + 1 < ntok ? rt[32u * (j + 1)] : 0;
			q2 = j + 2 < ntok ? rt[32u * (j + 2)] : 0;
			q3 = j + 3 < ntok ? rt[32u * (j + 3)] : 0;
		}
#ifdef B2I_HOST_EMUL
		/* the lanes do not talk to each other in this loop: under the host emulation every
		 * lane simply runs its own (a vote per step would cost two thread barriers) */
		while (active) {
#else
		while (__any_sync(B2I_FULL, active)) {
#endif
			if (active) {
				uint32_t v = 0;
				bool emit = true;
				if (rem == 0) {
					const uint32_t t = q0;
					const uint32_t len = t >> 16;
					if (j >= ntok) {
						active = false;
						emit = false;
					} else if (pos + len > lim) {
						/* does not fit: the chunk is full (partial region) or the output is */
						if (cap_bound || !partial)
							team_atomic_min(&ts->ck_err, ((pos - lead) << 1) | 1u);
						active = false;
						emit = false;
					} else {
						j++;
						q0 = q1; q1 = q2; q2 = q3;
						q3 = rt[32u * (j + 3 < ntok ? j + 3 : j)];
						if (len == 1) {
							v = t & 0xffu;
						} else {
							d = t & 0xffffu;
							rem = len;
							if (d > abs0 + (pos - lead)) {      /* zlib: invalid distance too far back */
								team_atomic_min(&ts->ck_err, (pos - lead) << 1);
								j--;
								active = false;
								emit = false;
							}
						}
					}
				}
				if (emit) {
					if (rem != 0) {
						/* up to four symbols of a match per step: independent stores */
						const uint32_t nn = rem < 4u ? rem : 4u;
#pragma unroll
						for (uint32_t k = 0; k < 4u; k++) {
							const uint32_t p = pos + k;
							if (k < nn) {
								B2I_CHECK(p < TEAM_CHUNK);
								ts->sym[p] = (uint16_t)(p >= d + lead ? TS_MAKE(p - d)
								    : (uint32_t)ts->hist[(abs0 + (p - lead) - d) & (TEAM_HIST - 1)]);
							}
						}
						pos += nn;
						rem -= nn;
					} else {
						B2I_CHECK(pos < TEAM_CHUNK);
						ts->sym[pos++] = (uint16_t)v;
					}
				}
			}
		}
		if (r < g1 && partial) {
			ts->tokpos[r] = j;
			ts->rdone[r] += pos - lead;
			ts->ck_len = pos - lead;
		}
	}
	team_sync();
	TPH(PH_TEXP);

	/* 2. sweep: pointers -> bytes, every warp its slice, in order */
	uint32_t T = ts->ck_len;
	if (ts->ck_err != TEAM_NOERR && (ts->ck_err >> 1) < T)
		T = ts->ck_err >> 1;
	const uint32_t end = lead + T;
	{
		const uint32_t slice = (((end + TEAM_WARPS - 1) / TEAM_WARPS) + 127u) & ~127u;
		const uint32_t stop_at = (w + 1) * slice < end ? (w + 1) * slice : end;
		for (uint32_t base = w * slice; base < stop_at; base += 128) {
			const uint32_t i = base + 4u * lane;
			if (i < stop_at) {
				uint2 v = *(const uint2 *)&ts->sym[i];
				uint32_t a = v.x & 0xffffu, b = v.x >> 16, c = v.y & 0xffffu, e = v.y >> 16;
				/* symbols in front of the chunk's first byte or past its end are not ours */
				if (i < lead) a = 0;
				if (i + 1 < lead || i + 1 >= end) b = 0;
				if (i + 2 < lead || i + 2 >= end) c = 0;
				if (i + 3 < lead || i + 3 >= end) e = 0;
				while ((a | b | c | e) >= 256u) {
					if (TS_IS_PTR(a)) a = ts->sym[TS_INDEX(a)];
					if (TS_IS_PTR(b)) b = ts->sym[TS_INDEX(b)];
					if (TS_IS_PTR(c)) c = ts->sym[TS_INDEX(c)];
					if (TS_IS_PTR(e)) e = ts->sym[TS_INDEX(e)];
				}
				v.x = a | (b << 16);
				v.y = c | (e << 16);
				*(uint2 *)&ts->sym[i] = v;
			}
			/* the lanes stay together: everything below `base` is bytes by now, so a chain
			 * ends after a few hops whatever the distances are */
			__syncwarp();
		}
	}
	team_sync();
	TPH(PH_TSWEEP);

	/* 3. flush: 16 symbols -> 16 bytes per lane and step, to global memory and the ring */
	{
		const uint32_t base = abs0 - lead;          /* output position of sym[0], 16-byte aligned */
		const uint32_t units = (end + 15u) >> 4;
		for (uint32_t u = tid; u < units; u += TEAM_LANES) {
			const uint4 s0 = *(const uint4 *)&ts->sym[16 * u];
			const uint4 s1 = *(const uint4 *)&ts->sym[16 * u + 8];
			uint4 o;
			o.x = pack_even_bytes(s0.x, s0.y);
			o.y = pack_even_bytes(s0.z, s0.w);
			o.z = pack_even_bytes(s1.x, s1.y);
			o.w = pack_even_bytes(s1.z, s1.w);
			const uint32_t p = base + 16 * u;
			if (16 * u >= lead && 16 * u + 16 <= end) {
				*(uint4 *)(out + p) = o;
				if (mir) *(uint4 *)(mir + p) = o;
				*(uint4 *)&ts->hist[p & (TEAM_HIST - 1)] = o;
			} else {
				/* first / last unit: only the bytes that belong to the chunk */
				const uint32_t lo = 16 * u < lead ? lead - 16 * u : 0;
				const uint32_t hi = 16 * u + 16 <= end ? 16 : end - 16 * u;
				const uint32_t wv[4] = { o.x, o.y, o.z, o.w };
				for (uint32_t k = lo; k < hi; k++) {
					const uint8_t by = (uint8_t)(wv[k >> 2] >> (8 * (k & 3u)));
					out[p + k] = by;
					if (mir) mir[p + k] = by;
					ts->hist[(p + k) & (TEAM_HIST - 1)] = by;
				}
			}
		}
	}
	team_sync();
	TPH(PH_TFLUSH);
#undef TPH
	if (tid == 0) {
		ts->ck_len = T;
		ts->hist_pos = abs0 + T;
	}
}

/* ---- COPY: a stored block, all lanes --------------------------------------------- */
B2I_DEV void team_copy(TeamShared *ts, unsigned w)
{
	const unsigned tid = 32 * w + b2i_lane();
	const uint8_t *src = ts->cp_src;
	uint8_t *dst = ts->cp_dst;
	uint8_t *mir = ts->mir ? ts->mir + (dst - ts->out) : nullptr;
	const uint32_t n = ts->cp_n;
	uint32_t head = (uint32_t)((0 - (uintptr_t)dst) & 15u);
	if (head > n) head = n;
	for (uint32_t i = tid; i < head; i += TEAM_LANES) {
		dst[i] = src[i];
		if (mir) mir[i] = src[i];
	}
	const uint32_t units = (n - head) >> 4;
	const uint8_t *s = src + head;
	const uint32_t sh = (uint32_t)((uintptr_t)s & 3u) * 8u;
	const uint32_t *sw = (const uint32_t *)(s - ((uintptr_t)s & 3u));
	for (uint32_t u = tid; u < units; u += TEAM_LANES) {
		const uint32_t *q = sw + 4 * u;
		uint32_t a = q[0], b = q[1], c = q[2], e = q[3];
		uint4 o;
		if (sh) {
			const uint32_t f = q[4];
			o.x = shf_r_wrap(a, b, sh); o.y = shf_r_wrap(b, c, sh);
			o.z = shf_r_wrap(c, e, sh); o.w = shf_r_wrap(e, f, sh);
		} else {
			o.x = a; o.y = b; o.z = c; o.w = e;
		}
		*(uint4 *)(dst + head + 16 * u) = o;
		if (mir) *(uint4 *)(mir + head + 16 * u) = o;
	}
	for (uint32_t i = head + 16 * units + tid; i < n; i += TEAM_LANES) {
		dst[i] = src[i];
		if (mir) mir[i] = src[i];
	}
}

/* ---- CRC: every warp one slice of the finished output ---------------------------- */
B2I_DEV void team_crc(TeamShared *ts, unsigned w)
{
	/* slice tables live in the chunk buffer, which is dead once the stream is decoded */
	uint32_t *tab = (uint32_t *)ts->sym;
	for (unsigned i = 32 * w + b2i_lane(); i < 1024; i += TEAM_LANES)
		tab[i] = ts->crc_tab_g[i];
	team_sync();
	const uint64_t n = ts->crc_n;
	const uint64_t slice = (((n + TEAM_WARPS - 1) / TEAM_WARPS) + 15) & ~(uint64_t)15;     /* the slices cover n */
	const uint64_t lo = (uint64_t)w * slice < n ? (uint64_t)w * slice : n;
	const uint64_t hi = lo + slice < n ? lo + slice : n;
	uint32_t raw0 = crc_warp_raw0(ts->crc_p + lo, hi - lo, tab, ts->crc_xp8);
	/* shift by the bytes that follow this slice */
	raw0 = crc_mulmod(raw0, crc_xpow8(n - hi, ts->crc_xp8));
	if (b2i_lane() == 0)
		ts->crc_part[w] = raw0;
}

/* what the helper warps (1..TEAM_WARPS-1) do for the lifetime of the CTA */
B2I_DEV void team_serve(TeamShared *ts, unsigned w, const WarpSmem *tables)
{
	for (;;) {
		team_sync();
		const uint32_t c = ts->cmd;
		if (c == TC_QUIT)
			break;
		if (c == TC_PASS_A) {
			team_stage_input(ts, w);
			team_do_pass<false>(ts, w, tables);
		} else if (c == TC_PASS_EMIT)
			team_do_pass<true>(ts, w, tables);
		else if (c == TC_CHUNK)
			team_chunk(ts, w);
		else if (c == TC_COPY)
			team_copy(ts, w);
		else
			team_crc(ts, w);
		team_sync();
	}
}

B2I_DEV void team_command(TeamShared *ts, uint32_t c)
{
	if (b2i_lane() == 0)
		ts->cmd = c;
	fence_block();
	team_sync();
}

/* warp 0: a stored block's payload, copied by the whole team */
B2I_DEV void team_copy_stored(TeamShared *ts, const uint8_t *src, uint8_t *dst, uint32_t n)
{
	if (b2i_lane() == 0) {
		ts->cp_src = src;
		ts->cp_dst = dst;
		ts->cp_n = n;
		ts->hist_pos = TEAM_NOERR;
	}
	__syncwarp();
	team_command(ts, TC_COPY);
	team_copy(ts, 0);
	team_sync();
}

/* warp 0: raw0 of out[0, n) by the whole team */
B2I_DEV uint32_t team_crc_raw0(TeamShared *ts, const uint8_t *p, uint64_t n, const uint32_t *crc_tab_g,
    const uint32_t *xp8)
{
	if (b2i_lane() == 0) {
		ts->crc_p = p;
		ts->crc_n = n;
		ts->crc_tab_g = crc_tab_g;
		ts->crc_xp8 = xp8;
	}
	__syncwarp();
	team_command(ts, TC_CRC);
	team_crc(ts, 0);
	team_sync();
	uint32_t acc = 0;
	for (unsigned w = 0; w < TEAM_WARPS; w++)
		acc ^= ts->crc_part[w];
	return acc;
}

/* once per CTA (one thread), before the first barrier */
B2I_DEV void team_init(TeamShared *ts)
{
	ts->stage_parity = 0;
	ts->stage_words = 0;
#ifndef B2I_HOST_EMUL
	mbar_init(&ts->stage_bar);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

/* once per stream (warp 0) */
B2I_DEV void team_stream_begin(TeamShared *ts, uint8_t *out, uint8_t *mir, uint32_t cap)
{
	if (b2i_lane() == 0) {
		ts->hist_pos = TEAM_NOERR;
		ts->next_seg = TEAM_SEG_INIT;
		ts->stage_words = 0;
		ts->out = out;
		ts->mir = mir;
		ts->cap = cap;
	}
	__syncwarp();
}

/* warp 0: same contract as lp_block, the work is shared with the team.  `carry` is
 * flushed on entry and re-read on exit: the chunks write their bytes themselves. */
B2I_DEV int lp_block_team(TeamShared *ts, WarpSmem *sm, const uint8_t *gbase, uint64_t glimit,
    uint64_t end_bits, uint64_t &P, uint8_t *out, uint8_t *mir, uint32_t cap, uint32_t &outp,
    uint32_t &carry, uint32_t &detail)
{
	const unsigned lane = b2i_lane();
	int ret = 2;
	bool entered = false;

	for (;;) {
		if (end_bits < P + TEAM_MIN_BITS) {
			ret = 2;
			break;
		}
		if (!entered) {
			/* everything produced so far goes to global memory */
			if (lane < (outp & 15u)) {
				out[(outp & ~15u) + lane] = (uint8_t)carry;
				if (mir) mir[(outp & ~15u) + lane] = (uint8_t)carry;
			}
			__syncwarp();
			fence_block();
			entered = true;
		}
		const uint64_t remaining = end_bits - P;
		const uint32_t p0 = (uint32_t)P & 31u;
		const uint32_t wbase = (uint32_t)(P >> 5);
		uint32_t seg = (uint32_t)((remaining + TEAM_LANES - 1) / TEAM_LANES);
		seg = (seg + 31u) & ~31u;
		if (seg > ts->next_seg) seg = ts->next_seg;
		/* every lane is done reading the previous round's shared state */
		__syncwarp();
		if (seg > TEAM_SEG_MAX) seg = TEAM_SEG_MAX;
		if (seg < LP_SEG_MIN) seg = LP_SEG_MIN;
		if (lane == 0) {
			ts->wbase = wbase;
			ts->p0 = p0;
			ts->seg = seg;
			ts->hard_end = remaining + p0 > 0x7fffffffull ? 0x7fffffffu : (uint32_t)remaining + p0;
			ts->gw = (const uint32_t *)gbase;
			ts->max_word = (uint32_t)(glimit >> 2) - 1u;
			ts->tables = sm;
			{
				/* everything the lanes may read: the segments, one symbol past the last one
				 * and the word the bit reader keeps in flight */
				const uint32_t need = (p0 + TEAM_LANES * seg) / 32u + 6u;
				ts->stage_words = need + 4u <= (TEAM_CHUNK * 2u) / 4u ? need : 0u;
				ts->stage_lead = wbase & 3u;
			}
			ts->out = out;
			ts->mir = mir;
			ts->cap = cap;
		}
		for (unsigned w = 0; w < TEAM_WARPS; w++) {
			const unsigned gid = 32 * w + lane;
			ts->start[gid] = p0 + gid * seg;
			ts->run[gid] = 1;
			ts->term[gid] = LT_NONE;
			ts->nsym[gid] = 0;
			ts->nbytes[gid] = 0;
			ts->exit_[gid] = 0;
			ts->tokpos[gid] = 0;
			ts->rdone[gid] = 0;
		}
		__syncwarp();
		PH_DECL();
		PH_COUNT(PH_TROUNDS, 1);
		/* pass A, then emit passes until every lane starts at its predecessor's exit */
		team_command(ts, TC_PASS_A);
		team_stage_input(ts, 0);
		team_do_pass<false>(ts, 0, sm);
		team_sync();
		bool first = true;
#ifdef B2I_HOST_EMUL
		if (lane == 0) g_lp_rounds++;
#endif
		for (;;) {
			unsigned any = 0;
			for (unsigned w = 0; w < TEAM_WARPS; w++) {
				const unsigned gid = 32 * w + lane;
				const uint32_t st = ts->start[gid];
				uint32_t want = st;
				if (gid != 0 && ts->term[gid - 1] == LT_NONE)
					want = ts->exit_[gid - 1];
				const bool redo = first || want != st;
				any |= __ballot_sync(B2I_FULL, redo);
				__syncwarp();
				ts->start[gid] = want;
				ts->run[gid] = redo;
			}
			if (!any)
				break;
			first = false;
			__syncwarp();
#ifdef B2I_HOST_EMUL
			if (lane == 0) g_lp_passes++;
#endif
			PH_COUNT(PH_TPASSES, 1);
			team_command(ts, TC_PASS_EMIT);
			team_do_pass<true>(ts, 0, sm);
			team_sync();
		}
		PH_ADD(PH_TPASS);
		/* the first lane that met EOB / an invalid code / the end of input ends the round */
		unsigned m = TEAM_LANES - 1;
		for (unsigned w = TEAM_WARPS; w-- > 0;) {
			unsigned tmask = __ballot_sync(B2I_FULL, ts->term[32 * w + lane] != LT_NONE);
			if (tmask)
				m = 32 * w + (unsigned)(__ffs(tmask) - 1);
		}
		const uint32_t mterm = ts->term[m], mexit = ts->exit_[m];
		/* region offsets inside the round: exclusive prefix sums of the byte counts */
		uint32_t total = 0;
		for (unsigned w = 0; w < TEAM_WARPS; w++) {
			const unsigned gid = 32 * w + lane;
			const uint32_t nb = gid <= m ? ts->nbytes[gid] : 0;
			uint32_t incl = nb;
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t t = __shfl_up_sync(B2I_FULL, incl, o);
				if ((int)lane >= o) incl += t;
			}
			ts->roff[gid] = total + incl - nb;
			total += __shfl_sync(B2I_FULL, incl, 31);
		}
		if (lane == 0) {
			ts->roff[TEAM_LANES] = total;
			ts->m = m;
			/* next round: segments sized so that a round's output is about one chunk */
			const uint32_t bits = mexit > p0 ? mexit - p0 : 1u;
			if (total != 0) {
				uint64_t want = (uint64_t)(TEAM_CHUNK - 2048u) * bits / total / TEAM_LANES;
				want &= ~(uint64_t)31;
				ts->next_seg = want > TEAM_SEG_MAX ? TEAM_SEG_MAX : (want < LP_SEG_MIN ? LP_SEG_MIN : (uint32_t)want);
			}
		}
		__syncwarp();
		/* chunks: consecutive regions whose bytes fit the buffer */
		const uint32_t base = outp;
		int32_t fst = 0;
		uint32_t g = 0;
		while (g <= m && fst == 0) {
			const uint32_t from = ts->roff[g] + ts->rdone[g];          /* round-relative */
			const uint32_t abs0 = base + from;
			const uint32_t room = TEAM_CHUNK - (abs0 & 15u);
			/* largest k with roff[k] - from <= room (roff is monotone) */
			uint32_t cnt = 0;
			for (unsigned w = 0; w < TEAM_WARPS; w++) {
				const uint32_t k = g + 1 + 32 * w + lane;
				const uint32_t lim = k <= m ? ts->roff[k] : (k == m + 1 ? total : 0xffffffffu);
				cnt += (uint32_t)__popc(__ballot_sync(B2I_FULL, k <= m + 1 && lim - from <= room));
			}
			const bool partial = cnt == 0 || ts->rdone[g] != 0;
			const uint32_t g1 = partial ? g + 1 : g + cnt;
			if (lane == 0) {
				ts->ck_g0 = g;
				ts->ck_g1 = g1;
				ts->ck_partial = partial;
				ts->ck_abs0 = abs0;
				ts->ck_len = partial ? 0 : (g1 <= m ? ts->roff[g1] : total) - from;
				ts->ck_err = TEAM_NOERR;
			}
			__syncwarp();
			team_command(ts, TC_CHUNK);
			team_chunk(ts, 0);
			team_sync();
			PH_COUNT(PH_BATCH, 1);
			const uint32_t err = ts->ck_err;
			outp = abs0 + ts->ck_len;
			if (err != TEAM_NOERR) {
				fst = (err & 1u) ? S_OUT_OVERFLOW : S_DATA_ERROR;
				detail = (err & 1u) ? 0 : D_DIST_TOO_FAR;
				break;
			}
			if (partial) {
				if (ts->tokpos[g] >= ts->nsym[g])
					g++;
			} else {
				g = g1;
			}
		}
		P = (uint64_t)wbase * 32u + mexit;
		if (fst < 0) { ret = fst; break; }
		if (mterm == LT_EOB) { ret = 0; break; }
		if (mterm == LT_EXH) { ret = S_BUF_ERROR; break; }
		if (mterm == LT_BADLIT) { detail = D_BAD_LITLEN_CODE; ret = S_DATA_ERROR; break; }
		if (mterm == LT_BADDST) { detail = D_BAD_DIST_CODE; ret = S_DATA_ERROR; break; }
	}
	if (entered) {
		/* everything produced so far is in global memory; pick the partial unit up again */
		fence_block();
		if (lane < (outp & 15u))
			carry = load_fresh(out + (outp & ~15u) + lane);
	}
	return ret;
}
