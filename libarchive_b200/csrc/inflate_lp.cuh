/*
 * inflate_lp.cuh — lane-parallel Huffman decoding of one deflate block.
 * (included by inflate_core.cuh)
 *
 * The warp-uniform decoder spends one instruction issue per lane-step on ONE
 * symbol; this decoder lets the 32 lanes of the warp decode 32 different
 * segments of the same block at once and recovers exactness by iteration
 * (the self-synchronisation property of prefix codes):
 *
 *   round:  the next 32 x SEG bits of the block are cut into 32 segments.
 *     pass A  lane i decodes from the START of segment i (lane 0 from the true
 *             position, the others from a guess) until it crosses the end of
 *             its segment, and reports the bit position where it crossed
 *             (its exit).  A decoder started at a wrong position falls into
 *             step with the true symbol sequence after a few symbols, so most
 *             exits are true symbol boundaries.
 *     pass B  lane i restarts from the exit of lane i-1 and, this time, writes
 *             its symbols (tokens: length << 16 | literal-or-distance) to its
 *             region of a per-warp scratch buffer in global memory.
 *     pass C+ repeated for the lanes whose start changed, until every lane's
 *             start equals its predecessor's exit.  Lane 0 is always true, so
 *             by induction lane k is true after at most k + 2 passes; in
 *             practice two or three passes converge.
 *   The first lane (in stream order) that meets the end-of-block code, an
 *   invalid code, or the end of the input ends the round; later lanes are
 *   speculation and are dropped.  The tokens of the valid lanes are then
 *   turned into bytes, in stream order, by the same batch resolution the
 *   uniform decoder uses (resolve_batch), which is also where zlib's
 *   "distance too far back" and the capacity checks are applied, so results
 *   are bit-identical to the serial decode.
 *
 * Input is read straight from global memory (each lane streams through its
 * own segment; the 128-byte lines stay in L1), tables are the same shared
 * memory tables the uniform decoder builds.
 */
#pragma once

#define LP_SEG_MAX   2048u                 /* bits per lane and round, at most   */
#define LP_SEG_MIN   128u
#define LP_CAP       (LP_SEG_MAX + 2u)     /* tokens a lane can produce per round */
#define LP_MIN_BITS  (32u * LP_SEG_MIN)    /* below this, decode uniformly        */
#define LP_SCRATCH_WORDS (32u * LP_CAP)    /* per warp, 32-bit tokens             */

#ifdef B2I_HOST_EMUL
extern long g_lp_rounds, g_lp_passes;   /* emit passes per round: test statistics */
#endif

#define LT_NONE   0u
#define LT_EOB    1u
#define LT_BADLIT 2u
#define LT_BADDST 3u
#define LT_EXH    4u    /* the symbol runs past the end of the stream */

struct LpOut {
	uint32_t exit;     /* bit position (window relative) after the last symbol */
	uint32_t nsym;
	uint32_t term;     /* LT_* */
	uint32_t nbytes;   /* output bytes of the nsym symbols (EMIT passes only) */
};

/*
 * One pass: this lane decodes from bit `start` (relative to word `wbase` of
 * the input) until it reaches `nominal_end`, the end of the block, an invalid
 * code or the end of the stream (`hard_end`).  Lanes with run == false keep
 * out of it.  EMIT: write tokens to tok[0..nsym).
 */
/* TSTRIDE: distance between consecutive tokens of one lane.  1: every lane owns a
 * contiguous run (the single-warp resolution reads one lane's tokens with all 32 lanes);
 * 32: token k of lane l sits at [32 k + l], so that 32 lanes writing - or, in the team
 * kernel's expand step, reading - "their next token" touch one or two 128-byte lines. */
template <bool EMIT, int TSTRIDE = 1>
B2I_DEV void lp_pass(const WarpSmem *sm, const uint32_t *gw, uint32_t wbase, uint32_t max_word,
    bool run, uint32_t start, uint32_t nominal_end, uint32_t hard_end, uint32_t *tok, LpOut &o)
{
	const char *litp = (const char *)sm->lit;
	const char *distp = (const char *)sm->dist;
	uint32_t lo = 0, hi = 0, widx = start >> 5, ns = 0, tm = LT_NONE, ex = 0, nw = 0, nb = 0, lastb = 0;
	int32_t cnt = 0;
	bool active = run;
	const uint32_t stop_at = nominal_end <= hard_end ? nominal_end : hard_end + 1u;

/* the word after the one being consumed is always already on its way (nw);
 * merging it in needs cnt <= 30: shift by cnt + 2 <= 32 */
#define LP_FETCH(i_) gw[min(wbase + (i_), max_word)]
#define LP_LOAD() do { \
		const uint32_t s_ = (uint32_t)cnt + 2u; \
		lo |= shl_clamp(nw, s_); \
		hi |= funnel_hi(nw, s_); \
		cnt += 32; widx++; \
		nw = LP_FETCH(widx); \
	} while (0)
#define LP_DROP(n_) do { \
		uint32_t n__ = (n_); \
		lo = shf_r_wrap(lo, hi, n__); hi = shf_r_wrap(hi, 0, n__); cnt -= (int32_t)(n__ & 31u); \
	} while (0)

	if (active) {
		nw = LP_FETCH(widx);
		LP_LOAD();
		LP_DROP(start & 31u);
	}
	while (__any_sync(B2I_FULL, active)) {
		if (!active)
			continue;
		if (cnt <= 30)
			LP_LOAD();
		uint32_t e = *(const uint32_t *)(litp + (lo & ((1u << (LIT_ROOT + 2)) - 4u)));
		if ((e & E_SLOW) && E_SLOWKIND(e) == SK_SUB) {
			uint32_t idx = shf_r_wrap(lo, hi, LIT_ROOT + 2) & ((1u << E_SUBBITS(e)) - 1u);
			e = sm->lit[E_SUBOFF(e) + idx];
		}
		uint32_t token;
		if (e & E_SLOW) {
			LP_DROP(e);
			uint32_t pos = widx * 32u - (uint32_t)cnt;
			tm = pos > hard_end ? LT_EXH : (E_SLOWKIND(e) == SK_EOB ? LT_EOB : LT_BADLIT);
			ex = pos;
			active = false;
			continue;
		}
		if (e & E_LIT) {
			token = e >> 8;
			LP_DROP(e);
		} else {
			uint32_t len = (e >> 24) + (shf_r_wrap(lo, hi, e >> 8) & byte2(e)) + 3u;
			LP_DROP(e);
			if (cnt <= 30)
				LP_LOAD();
			uint32_t d = *(const uint32_t *)(distp + (lo & ((1u << (DIST_ROOT + 2)) - 4u)));
			if ((d & E_SLOW) && E_SLOWKIND(d) == SK_SUB) {
				uint32_t idx = shf_r_wrap(lo, hi, DIST_ROOT + 2) & ((1u << E_SUBBITS(d)) - 1u);
				d = sm->dist[E_SUBOFF(d) + idx];
			}
			if (d & E_SLOW) {
				LP_DROP(d);
				uint32_t pos = widx * 32u - (uint32_t)cnt;
				tm = pos > hard_end ? LT_EXH : LT_BADDST;
				ex = pos;
				active = false;
				continue;
			}
			uint32_t dist = (d >> 17) + (shf_r_wrap(lo, hi, d >> 8) & ((1u << ((d >> 13) & 15u)) - 1u));
			LP_DROP(d);
			token = (len << 16) | dist;
		}
		B2I_CHECK(ns < LP_CAP);
		if (EMIT) {
			tok[ns * TSTRIDE] = token;
			lastb = token >> 16;
			nb += lastb;
		}
		ns++;
		/* one comparison per symbol: stop at the end of the segment or right
		 * after the first symbol that runs past the end of the stream */
		uint32_t pos = widx * 32u - (uint32_t)cnt;
		if (pos >= stop_at) {
			ex = pos;
			active = false;
			if (pos > hard_end) {
				tm = LT_EXH;        /* that symbol does not count */
				ns--;
				nb -= lastb;
			}
		}
	}
#undef LP_LOAD
#undef LP_FETCH
#undef LP_DROP
	if (run) {
		o.exit = ex;
		o.nsym = ns;
		o.term = tm;
		o.nbytes = nb;
	}
}

/*
 * Decode the symbols of the current Huffman block lane-parallel, starting at
 * the true bit position P (in bits from r.gbase), and append the bytes to the
 * output.  Returns through `P` the position reached and in `status`:
 *   0  the block ended (end-of-block code consumed), P is just past it
 *   2  fewer than LP_MIN_BITS of input remain: continue uniformly from P
 *  <0  a B2I status (detail in `detail`); P / outp say how far decoding got
 */
B2I_DEV int lp_block(WarpSmem *sm, const uint8_t *gbase, uint64_t glimit, uint64_t end_bits,
    uint64_t &P, uint32_t *scratch, uint8_t *out, uint8_t *mir, uint32_t cap, uint32_t &outp, uint32_t &carry,
    uint32_t &detail)
{
	const unsigned lane = b2i_lane();
	const uint32_t *gw = (const uint32_t *)gbase;
	const uint32_t max_word = (uint32_t)(glimit >> 2) - 1u;
	uint32_t *tok = scratch + lane * LP_CAP;

	for (;;) {
		if (end_bits < P + LP_MIN_BITS)
			return 2;
		const uint64_t remaining = end_bits - P;
		const uint32_t wbase = (uint32_t)(P >> 5);
		const uint32_t p0 = (uint32_t)P & 31u;
		const uint32_t hard_end = remaining + p0 > 0x7fffffffull ? 0x7fffffffu : (uint32_t)remaining + p0;
		uint32_t seg = (uint32_t)((remaining + 31) >> 5);
		seg = (seg + 31u) & ~31u;
		if (seg > LP_SEG_MAX) seg = LP_SEG_MAX;
		if (seg < LP_SEG_MIN) seg = LP_SEG_MIN;
		const uint32_t nominal_end = p0 + (lane + 1u) * seg;
		uint32_t start = p0 + lane * seg;
		LpOut o;
		o.exit = 0; o.nsym = 0; o.term = LT_NONE; o.nbytes = 0;
		PH_DECL();

		/* pass A: where does every segment's decoder cross into the next one? */
		lp_pass<false>(sm, gw, wbase, max_word, true, start, nominal_end, hard_end, tok, o);
		bool first = true;
#ifdef B2I_HOST_EMUL
		if (lane == 0) g_lp_rounds++;
#endif
		for (;;) {
			/* a lane's true start is its predecessor's exit (if that one got there) */
			uint32_t pex = __shfl_up_sync(B2I_FULL, o.exit, 1);
			uint32_t ptm = __shfl_up_sync(B2I_FULL, o.term, 1);
			uint32_t want = (lane == 0 || ptm != LT_NONE) ? start : pex;
			bool redo = first || want != start;
			if (!__any_sync(B2I_FULL, redo))
				break;
			first = false;
			start = want;
#ifdef B2I_HOST_EMUL
			if (lane == 0) g_lp_passes++;
#endif
			lp_pass<true>(sm, gw, wbase, max_word, redo, start, nominal_end, hard_end, tok, o);
		}
		/* the first lane that met EOB / an invalid code / the end of input ends the round */
		unsigned tmask = __ballot_sync(B2I_FULL, o.term != LT_NONE);
		const int m = tmask ? __ffs(tmask) - 1 : 31;
		const uint32_t mterm = __shfl_sync(B2I_FULL, o.term, m);
		const uint32_t mexit = __shfl_sync(B2I_FULL, o.exit, m);
		__syncwarp();
		PH_ADD(PH_LPDEC);
		/* tokens -> bytes, region after region in stream order */
		for (int rgn = 0; rgn <= m; rgn++) {
			const uint32_t cnt = __shfl_sync(B2I_FULL, o.nsym, rgn);
			const uint32_t *rt = scratch + (uint32_t)rgn * LP_CAP;
			uint32_t j = 0;
			/* the tokens of the NEXT batch are requested before this one is resolved
			 * (a batch almost always takes all 32; if it takes fewer, reload) */
			uint32_t nxt = lane < cnt ? rt[lane] : 0;
			while (j < cnt) {
				const uint32_t my = nxt;
				const uint32_t avail = cnt - j < 32u ? cnt - j : 32u;
				int32_t stop = 0;
				nxt = j + 32u + lane < cnt ? rt[j + 32u + lane] : 0;
				/* resolve_batch takes as many symbols as its staging buffer holds */
				uint32_t n = resolve_batch(sm, out, mir, cap, outp, carry, my, avail, stop, detail);
				PH_COUNT(PH_BATCH, 1);
				if (stop < 0) {
					P = (uint64_t)wbase * 32u + mexit;
					return stop;
				}
				j += n;
				if (n != 32u && j < cnt)
					nxt = j + lane < cnt ? rt[j + lane] : 0;
			}
		}
		PH_ADD(PH_LPRES);
		P = (uint64_t)wbase * 32u + mexit;
		if (mterm == LT_EOB)
			return 0;
		if (mterm == LT_EXH)
			return S_BUF_ERROR;
		if (mterm == LT_BADLIT) { detail = D_BAD_LITLEN_CODE; return S_DATA_ERROR; }
		if (mterm == LT_BADDST) { detail = D_BAD_DIST_CODE; return S_DATA_ERROR; }
	}
}
