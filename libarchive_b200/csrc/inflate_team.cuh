/*
 * inflate_team.cuh — one LARGE deflate stream decoded by a team of warps.
 * (included by inflate_core.cuh)
 *
 * A single warp decodes a stream at a few tens of MB/s, which makes archives
 * with multi-megabyte entries (BASELINE config 4) latency-bound by their
 * largest entry.  Here a CTA of TEAM_WARPS warps works on ONE stream:
 *
 *   warp 0 owns the stream exactly as in the single-warp kernel (block
 *   headers, table construction, stored blocks, the uniform tail, the CRC
 *   epilogue).  For the symbols of a Huffman block it runs the lane-parallel
 *   rounds over TEAM_WARPS x 32 segments instead of 32 and hands two kinds of
 *   work to the other warps through shared memory and a named barrier:
 *
 *     PASS     every warp decodes its 32 segments (lp_pass, same code, tables
 *              read from warp 0's shared memory), exits / counts go to shared
 *              arrays; warp 0 evaluates the chain and repeats for the lanes
 *              whose start moved, exactly like the single-warp convergence loop;
 *     RESOLVE  output offsets are the prefix sums of the segments' byte counts,
 *              so every warp resolves its own 32 token regions at its own
 *              offset, concurrently.  A warp may gather match sources from an
 *              earlier warp's range only below that warp's published flush
 *              position (TeamLink, resolve_batch_t<true>); with >= 32 KiB of
 *              output per warp range the warps advance in lock-step and rarely
 *              wait.  The 16-byte unit two neighbouring ranges share is written
 *              bytewise by both.
 *
 * Results are bit-identical to the single-warp path: same tokens, same order,
 * the same checks at the same symbols; the first failing position wins.
 */
#pragma once

#define TEAM_WARPS   4
#define TEAM_LANES   (32 * TEAM_WARPS)
#define TEAM_MIN_BITS ((uint64_t)TEAM_LANES * LP_SEG_MIN)

#define TC_PASS_A    1u
#define TC_PASS_EMIT 2u
#define TC_RESOLVE   3u
#define TC_QUIT      4u

struct TeamShared {
	volatile uint32_t cmd;
	uint32_t wbase, max_word, hard_end, seg, p0, cap, m;
	const uint32_t *gw;
	const WarpSmem *tables;
	uint8_t *out, *mir;
	uint32_t *scratch[TEAM_WARPS];
	uint32_t start[TEAM_LANES], exit_[TEAM_LANES], term[TEAM_LANES], nsym[TEAM_LANES], nbytes[TEAM_LANES];
	uint32_t run[TEAM_LANES];
	uint32_t range_start[TEAM_WARPS], range_end[TEAM_WARPS];
	volatile uint32_t done_pos[TEAM_WARPS];
	int32_t  fail_status[TEAM_WARPS];
	uint32_t fail_detail[TEAM_WARPS], fail_pos[TEAM_WARPS];
};

#ifndef B2I_HOST_EMUL
B2I_DEV void team_sync() { asm volatile("bar.sync 1, %0;" :: "n"(TEAM_LANES) : "memory"); }
#endif

template <bool EMIT>
B2I_DEV void team_do_pass(TeamShared *ts, unsigned w)
{
	const unsigned lane = b2i_lane();
	const unsigned gid = 32 * w + lane;
	const bool run = ts->run[gid] != 0;
	LpOut o;
	o.exit = 0; o.nsym = 0; o.term = LT_NONE; o.nbytes = 0;
	lp_pass<EMIT>(ts->tables, ts->gw, ts->wbase, ts->max_word, run, ts->start[gid],
	    ts->p0 + (gid + 1u) * ts->seg, ts->hard_end, ts->scratch[w] + lane * LP_CAP, o);
	if (run) {
		ts->exit_[gid] = o.exit;
		ts->term[gid] = o.term;
		ts->nsym[gid] = o.nsym;
		ts->nbytes[gid] = o.nbytes;
	}
}

/* warp w turns the tokens of its regions into bytes at its own output offset */
B2I_DEV void team_do_resolve(TeamShared *ts, unsigned w, WarpSmem *sm, uint32_t carry)
{
	const unsigned lane = b2i_lane();
	TeamLink tl;
	uint32_t outp = ts->range_start[w];
	int32_t stop = 0;
	uint32_t detail = 0;

	tl.done_pos = ts->done_pos;
	tl.range_start = ts->range_start;
	tl.range_end = ts->range_end;
	tl.w = w;
	tl.head_skip = w == 0 ? 0 : (outp & 15u);
	if (w != 0)
		carry = 0;
	ts->fail_status[w] = 0;
	for (unsigned rgn = 32 * w; rgn <= ts->m && rgn < 32 * w + 32 && stop >= 0; rgn++) {
		const uint32_t cnt = ts->nsym[rgn];
		const uint32_t *rt = ts->scratch[w] + (rgn - 32 * w) * LP_CAP;
		uint32_t j = 0;
		uint32_t nxt = lane < cnt ? rt[lane] : 0;
		while (j < cnt) {
			const uint32_t my = nxt;
			const uint32_t avail = cnt - j < 32u ? cnt - j : 32u;
			nxt = j + 32u + lane < cnt ? rt[j + 32u + lane] : 0;
			uint32_t n = resolve_batch_t<true>(sm, ts->out, ts->mir, ts->cap, outp, carry, my, avail,
			    stop, detail, &tl);
			if (stop < 0)
				break;
			j += n;
			if (n != 32u && j < cnt)
				nxt = j + lane < cnt ? rt[j + lane] : 0;
		}
	}
	/* the tail (< 16 bytes) goes out bytewise: the next range continues in the same unit */
	if (lane < (outp & 15u) && lane >= tl.head_skip) {
		ts->out[(outp & ~15u) + lane] = (uint8_t)carry;
		if (ts->mir) ts->mir[(outp & ~15u) + lane] = (uint8_t)carry;
	}
	__syncwarp();
	fence_block();
	if (lane == 0) {
		if (stop < 0) {
			ts->fail_status[w] = stop;
			ts->fail_detail[w] = detail;
			ts->fail_pos[w] = outp;
		}
		ts->done_pos[w] = ts->range_end[w];     /* also on failure: nobody may wait forever */
	}
}

/* what the helper warps (1..TEAM_WARPS-1) do for the lifetime of the CTA */
B2I_DEV void team_serve(TeamShared *ts, unsigned w, WarpSmem *sm)
{
	for (;;) {
		team_sync();
		const uint32_t c = ts->cmd;
		if (c == TC_QUIT)
			break;
		if (c == TC_PASS_A)
			team_do_pass<false>(ts, w);
		else if (c == TC_PASS_EMIT)
			team_do_pass<true>(ts, w);
		else
			team_do_resolve(ts, w, sm, 0);
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

/* warp 0: same contract as lp_block, the work is shared with the team */
B2I_DEV int lp_block_team(TeamShared *ts, WarpSmem *sm, const uint8_t *gbase, uint64_t glimit,
    uint64_t end_bits, uint64_t &P, uint8_t *out, uint8_t *mir, uint32_t cap, uint32_t &outp,
    uint32_t &carry, uint32_t &detail)
{
	const unsigned lane = b2i_lane();

	for (;;) {
		if (end_bits < P + TEAM_MIN_BITS)
			return 2;
		const uint64_t remaining = end_bits - P;
		const uint32_t p0 = (uint32_t)P & 31u;
		uint32_t seg = (uint32_t)((remaining + TEAM_LANES - 1) / TEAM_LANES);
		seg = (seg + 31u) & ~31u;
		if (seg > LP_SEG_MAX) seg = LP_SEG_MAX;
		if (seg < LP_SEG_MIN) seg = LP_SEG_MIN;
		if (lane == 0) {
			ts->wbase = (uint32_t)(P >> 5);
			ts->p0 = p0;
			ts->seg = seg;
			ts->hard_end = remaining + p0 > 0x7fffffffull ? 0x7fffffffu : (uint32_t)remaining + p0;
			ts->gw = (const uint32_t *)gbase;
			ts->max_word = (uint32_t)(glimit >> 2) - 1u;
			ts->tables = sm;
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
		}
		__syncwarp();
		/* pass A, then emit passes until every lane starts at its predecessor's exit */
		team_command(ts, TC_PASS_A);
		team_do_pass<false>(ts, 0);
		team_sync();
		bool first = true;
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
			team_command(ts, TC_PASS_EMIT);
			team_do_pass<true>(ts, 0);
			team_sync();
		}
		/* the first lane that met EOB / an invalid code / the end of input ends the round */
		unsigned m = TEAM_LANES - 1;
		for (unsigned w = TEAM_WARPS; w-- > 0;) {
			unsigned tmask = __ballot_sync(B2I_FULL, ts->term[32 * w + lane] != LT_NONE);
			if (tmask)
				m = 32 * w + (unsigned)(__ffs(tmask) - 1);
		}
		const uint32_t mterm = ts->term[m], mexit = ts->exit_[m];
		/* output ranges of the warps: prefix sums of the byte counts */
		uint32_t base = outp;
		for (unsigned w = 0; w < TEAM_WARPS; w++) {
			const unsigned gid = 32 * w + lane;
			uint32_t nb = gid <= m ? ts->nbytes[gid] : 0;
			for (int o = 16; o; o >>= 1)
				nb += __shfl_xor_sync(B2I_FULL, nb, o);
			if (lane == 0) {
				ts->range_start[w] = base;
				ts->range_end[w] = base + nb;
				ts->done_pos[w] = base;
			}
			base += nb;
		}
		if (lane == 0)
			ts->m = m;
		__syncwarp();
		team_command(ts, TC_RESOLVE);
		team_do_resolve(ts, 0, sm, carry);
		team_sync();
		P = (uint64_t)ts->wbase * 32u + mexit;
		/* the earliest failure (in stream order) decides */
		int32_t fst = 0;
		uint32_t fpos = 0xffffffffu, fdet = 0;
		for (unsigned w = 0; w < TEAM_WARPS; w++)
			if (ts->fail_status[w] < 0 && ts->fail_pos[w] < fpos) {
				fst = ts->fail_status[w];
				fpos = ts->fail_pos[w];
				fdet = ts->fail_detail[w];
			}
		outp = fst < 0 ? fpos : base;
		/* everything produced so far is in global memory; pick the partial unit up again */
		if (lane < (outp & 15u))
			carry = load_fresh(out + (outp & ~15u) + lane);
		if (fst < 0) {
			detail = fdet;
			return fst;
		}
		if (mterm == LT_EOB)
			return 0;
		if (mterm == LT_EXH)
			return S_BUF_ERROR;
		if (mterm == LT_BADLIT) { detail = D_BAD_LITLEN_CODE; return S_DATA_ERROR; }
		if (mterm == LT_BADDST) { detail = D_BAD_DIST_CODE; return S_DATA_ERROR; }
	}
}
