/*
 * crc32_core.cuh — warp-level CRC-32 (reflected 0xEDB88320) building blocks.
 *
 * Replaces zlib crc32() as called through zip->crc32func
 * (archive_read_support_format_zip.c:405-409, 3154-3157); semantics are those
 * of the reference's own restatement, archive_crc32.h:43-84: init/final XOR
 * 0xffffffff, crc32(x, NULL, 0) == 0, chaining by passing the previous value.
 *
 * Decomposition (no reference counterpart — the reference is byte-serial):
 *   raw0(M)      CRC register after M starting from register 0 (linear in M)
 *   shift(v, n)  register v after n zero bytes = v * x^(8n) mod P
 *   crc(c, M)    = raw0(M) ^ shift(c ^ ~0, |M|) ^ ~0
 *   raw0(A||B)   = shift(raw0(A), |B|) ^ raw0(B)        (crc32_combine)
 * A warp cuts a region into 32 equal lane slices, each lane runs slice-by-4
 * over its slice with tables in shared memory, and the 32 partials are merged
 * by a 5-level shuffle tree whose level-k multiplier is x^(8*S*2^k).
 */
#pragma once
#include "b2i_common.cuh"

/* a(x)*b(x) mod P, reflected representation (bit 31 is x^0) */
B2I_DEV uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
	uint32_t p = 0;
#pragma unroll 8
	for (int i = 0; i < 32; i++) {
		p ^= b & (uint32_t)((int32_t)a >> 31);
		a <<= 1;
		b = (b >> 1) ^ (CRC_POLY & (0u - (b & 1u)));
	}
	return p;
}

/* x^(8n) mod P;  xp8[k] = x^(8 * 2^k) mod P, k = 0..39 */
B2I_DEV uint32_t crc_xpow8(uint64_t n, const uint32_t *xp8)
{
	uint32_t r = 0x80000000u;
	int k = 0;
	while (n) {
		if (n & 1)
			r = crc_mulmod(r, xp8[k]);
		n >>= 1;
		k++;
	}
	return r;
}

/* tab[k*256 + b] = register after byte b followed by k zero bytes */
B2I_DEV uint32_t crc_word(uint32_t s, uint32_t w, const uint32_t *tab)
{
	s ^= w;
	return tab[768 + (s & 0xff)] ^ tab[512 + ((s >> 8) & 0xff)] ^
	    tab[256 + ((s >> 16) & 0xff)] ^ tab[s >> 24];
}

B2I_DEV uint32_t crc_byte(uint32_t s, uint32_t b, const uint32_t *tab)
{
	return tab[(s ^ b) & 0xff] ^ (s >> 8);
}

/* merge 32 per-lane partials of equal slice length (multiplier X = x^(8*S));
 * lane 31 holds the last slice.  Result valid in every lane. */
B2I_DEV uint32_t crc_warp_merge(uint32_t r, uint32_t X, uint32_t *X32)
{
	const unsigned lane = b2i_lane();
#pragma unroll 1
	for (int k = 0; k < 5; k++) {
		/* pairs (lo, hi) at distance 2^k: lo' = shift(lo) ^ hi, kept in hi's slot */
		uint32_t shifted = crc_mulmod(r, X);
		uint32_t other = __shfl_xor_sync(B2I_FULL, shifted, 1u << k);
		if (lane & (1u << k))
			r ^= other;          /* hi lane: own ^ shifted lower half */
		X = crc_mulmod(X, X);
	}
	*X32 = X;                   /* x^(8*S*32): shift across the whole body */
	return __shfl_sync(B2I_FULL, r, 31);
}

/*
 * raw0 of `n` bytes at p (any alignment), computed by the whole warp.
 * tab: 4x256 slice tables in shared memory; xp8: powers table (global).
 * Three parts: a head up to 16-byte alignment and a tail (< 16 bytes) done
 * bytewise by lane 0, an aligned body of 32 equal slices of S bytes (S a
 * multiple of 16), and a remainder of 16-byte slices right-aligned on the
 * lanes.  Result valid in every lane.
 */
B2I_DEV uint32_t crc_warp_raw0(const uint8_t *p, uint64_t n, const uint32_t *tab,
    const uint32_t *xp8)
{
	const unsigned lane = b2i_lane();
	uint32_t head = (uint32_t)((0 - (uintptr_t)p) & 15);
	uint32_t acc = 0;

	if (head > n)
		head = (uint32_t)n;
	if (head) {
		uint32_t s = 0;
		if (lane == 0)
			for (uint32_t i = 0; i < head; i++)
				s = crc_byte(s, p[i], tab);
		acc = __shfl_sync(B2I_FULL, s, 0);
		p += head;
		n -= head;
	}
	/* body: 32 slices of S bytes */
	uint64_t S = (n >> 5) & ~(uint64_t)15;
	if (S) {
		const uint4 *q = (const uint4 *)(p + (uint64_t)lane * S);
		uint32_t s = 0;
		/* two 16-byte loads in flight per lane: the next one is requested before
		 * the current one is folded in */
		uint4 v = *q++;
		for (uint64_t i = 16; i < S; i += 16) {
			const uint4 nv = *q++;
			s = crc_word(s, v.x, tab);
			s = crc_word(s, v.y, tab);
			s = crc_word(s, v.z, tab);
			s = crc_word(s, v.w, tab);
			v = nv;
		}
		s = crc_word(s, v.x, tab);
		s = crc_word(s, v.y, tab);
		s = crc_word(s, v.z, tab);
		s = crc_word(s, v.w, tab);
		uint32_t Xall;
		uint32_t body = crc_warp_merge(s, crc_xpow8(S, xp8), &Xall);
		acc = crc_mulmod(acc, Xall) ^ body;
		p += S * 32;
		n -= S * 32;
	}
	/* remainder: q 16-byte slices (q <= 31), right-aligned on the lanes */
	uint32_t qn = (uint32_t)(n >> 4);
	if (qn) {
		uint32_t s = 0;
		int j = (int)lane - (int)(32 - qn);
		if (j >= 0) {
			uint4 v = ((const uint4 *)p)[j];
			s = crc_word(s, v.x, tab);
			s = crc_word(s, v.y, tab);
			s = crc_word(s, v.z, tab);
			s = crc_word(s, v.w, tab);
		}
		uint32_t X512;
		uint32_t rem = crc_warp_merge(s, xp8[4], &X512);   /* x^(8*16) per slice */
		acc = crc_mulmod(acc, crc_xpow8((uint64_t)qn * 16, xp8)) ^ rem;
		p += qn * 16;
		n -= qn * 16;
	}
	if (n) {
		uint32_t s = acc;
		if (lane == 0)
			for (uint32_t i = 0; i < (uint32_t)n; i++)
				s = crc_byte(s, p[i], tab);
		acc = __shfl_sync(B2I_FULL, s, 0);
	}
	return acc;
}

/* finish: crc(c_in, M) from raw0(M) */
B2I_DEV uint32_t crc_finish(uint32_t crc_in, uint32_t raw0, uint64_t n, const uint32_t *xp8)
{
	return raw0 ^ crc_mulmod(crc_in ^ 0xffffffffu, crc_xpow8(n, xp8)) ^ 0xffffffffu;
}

/* copy the 4x256 slice tables (global) into shared memory with the warp */
B2I_DEV void crc_load_tables(uint32_t *dst_smem, const uint32_t *src_global)
{
	const unsigned lane = b2i_lane();
	for (int i = lane; i < 1024; i += 32)
		dst_smem[i] = src_global[i];
	__syncwarp();
}
