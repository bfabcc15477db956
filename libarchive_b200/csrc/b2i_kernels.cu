/*
 * b2i_kernels.cu — sm_100a kernels behind include/b200inflate.h.
 *
 *   b2i_inflate_kernel      K1: one warp per deflate stream (+ fused CRC-32)
 *   b2i_crc_chunks_kernel   K2/K3: chunked CRC-32 (+ optional copy) of stored
 *                           entries, one warp per 16 KiB chunk
 *   b2i_crc_combine_kernel  K2: merges chunk partials per entry
 *                           (crc32_combine arithmetic) and applies the
 *                           reference's end-of-entry checks
 *   b2i_tables_kernel       one-time: slice-by-4 tables and x^(8*2^k) powers
 *
 * Grid sizing: K1 keeps 28 warps (7 CTAs x 4 warps) resident per SM — the
 * per-warp shared-memory tables (7888 B) are the limiter — and CTAs pull
 * streams from a global counter, largest streams first, so a launch never
 * runs more than 148 x 7 CTAs and has no tail of idle CTAs.
 */
#include "stream_core.cuh"
#include "b2i_kernels.h"

#include <stdio.h>
#define INFLATE_WARPS 4
#ifdef B2I_R9
/* second build of this file: only the single-warp inflate kernel and its launch
 * wrapper, under other names (see inflate_core.cuh) */
#define INFLATE_CTAS 8
#define b2i_inflate_kernel b2i_inflate_kernel_r9
#define b2i_launch_inflate b2i_launch_inflate_r9
#else
#define INFLATE_CTAS 7
#endif
#define TEAM_CTAS_PER_SM (TEAM_WARPS >= 16 ? 1 : 2)
#define INFLATE_SLOTS 64u      /* token regions per SM: twice the resident warps of the larger build */

#if defined(B2I_PHASE_CLOCKS) && !defined(B2I_R9)
__global__ void b2i_phase_dump_kernel()
{
	printf("B2I_PHASE header=%llu lpdec=%llu lpres=%llu unif=%llu crc=%llu batches=%llu stored=%llu\n",
	    g_b2i_phase[0], g_b2i_phase[1], g_b2i_phase[2], g_b2i_phase[3], g_b2i_phase[4], g_b2i_phase[5],
	    g_b2i_phase[6]);
	printf("B2I_TEAM passes=%llu expand=%llu sweep=%llu flush=%llu plan=%llu rounds=%llu emit_passes=%llu expand_warp_iters=%llu\n",
	    g_b2i_phase[8], g_b2i_phase[9], g_b2i_phase[10], g_b2i_phase[11], g_b2i_phase[12], g_b2i_phase[13],
	    g_b2i_phase[14], g_b2i_phase[15]);
	for (int i = 0; i < 16; i++) g_b2i_phase[i] = 0;
}
void b2i_phase_dump(cudaStream_t st) { b2i_phase_dump_kernel<<<1, 1, 0, st>>>(); }
#endif

extern "C" __global__ void __launch_bounds__(INFLATE_WARPS * 32, INFLATE_CTAS)
b2i_inflate_kernel(const uint8_t *__restrict__ in, uint64_t in_total, uint8_t *__restrict__ out,
    uint8_t *__restrict__ out_mirror,
    const B2iDesc *__restrict__ descs, B2iResult *__restrict__ results,
    const uint32_t *__restrict__ order, uint32_t n, unsigned int *counter,
    const uint32_t *__restrict__ crc_tab, const uint32_t *__restrict__ xp8, uint32_t *scratch,
    unsigned int *slot_busy, uint32_t nslots)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	WarpSmem *sm = reinterpret_cast<WarpSmem *>(smem_raw) + (threadIdx.x >> 5);
	const unsigned lane = threadIdx.x & 31;
	Ring ring;
	/* Token scratch for the lane-parallel decoder (NULL: uniform only).  Several
	 * launches may be resident at once (pipelined host path), so a warp claims a
	 * free region for its lifetime: there are twice as many regions as warps
	 * that fit on the device, probing starts at a per-SM, per-warp position. */
	uint32_t *my_scratch = nullptr;
	uint32_t slot = 0;
	if (scratch) {
		if (lane == 0) {
			unsigned smid;
			asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
			slot = (smid * INFLATE_SLOTS + (blockIdx.x * INFLATE_WARPS + (threadIdx.x >> 5)) % (2u * INFLATE_CTAS * INFLATE_WARPS)) % nslots;
			while (atomicCAS(&slot_busy[slot], 0u, 1u) != 0u)
				slot = slot + 1 == nslots ? 0 : slot + 1;
		}
		slot = __shfl_sync(B2I_FULL, slot, 0);
		my_scratch = scratch + (size_t)slot * LP_SCRATCH_WORDS;
	}

	ring_init(sm, ring);
	for (;;) {
		uint32_t slot = 0;
		if (lane == 0)
			slot = atomicAdd(counter, 1u);
		slot = __shfl_sync(B2I_FULL, slot, 0);
		if (slot >= n)
			break;
		const uint32_t idx = order[slot];
		const B2iDesc d = descs[idx];
		process_deflate_stream(sm, ring, my_scratch, nullptr, in, in_total, out, out_mirror, d, &results[idx], crc_tab, xp8);
		__syncwarp();
	}
	if (scratch && lane == 0) {
		__threadfence();
		atomicExch(&slot_busy[slot], 0u);
	}
}

#ifndef B2I_R9
/*
 * K1 for LARGE streams: one CTA (TEAM_WARPS warps) per stream, see inflate_team.cuh.
 * Warp 0 pulls streams from the counter and owns them; the other warps serve its
 * PASS / RESOLVE commands until it quits.
 */
extern "C" __global__ void __launch_bounds__(TEAM_LANES, TEAM_CTAS_PER_SM)
b2i_inflate_team_kernel(const uint8_t *__restrict__ in, uint64_t in_total, uint8_t *__restrict__ out,
    uint8_t *__restrict__ out_mirror,
    const B2iDesc *__restrict__ descs, B2iResult *__restrict__ results,
    const uint32_t *__restrict__ order, uint32_t n, unsigned int *counter,
    const uint32_t *__restrict__ crc_tab, const uint32_t *__restrict__ xp8, uint32_t *scratch,
    unsigned int *slot_busy, uint32_t nslots)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const unsigned w = threadIdx.x >> 5;
	const unsigned lane = threadIdx.x & 31;
	TeamShared *ts = reinterpret_cast<TeamShared *>(smem_raw);
	WarpSmem *sm = reinterpret_cast<WarpSmem *>(smem_raw + sizeof(TeamShared));   /* warp 0's tables */
	uint32_t slot = 0;

	if (lane == 0) {
		unsigned smid;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		slot = (smid * INFLATE_SLOTS + (blockIdx.x * TEAM_WARPS + w) % INFLATE_SLOTS) % nslots;
		while (atomicCAS(&slot_busy[slot], 0u, 1u) != 0u)
			slot = slot + 1 == nslots ? 0 : slot + 1;
		ts->scratch[w] = scratch + (size_t)slot * LP_SCRATCH_WORDS;
		if (w == 0)
			team_init(ts);
	}
	slot = __shfl_sync(B2I_FULL, slot, 0);
	__syncthreads();
	if (w == 0) {
		Ring ring;
		ring_init(sm, ring);
		for (;;) {
			uint32_t s = 0;
			if (lane == 0)
				s = atomicAdd(counter, 1u);
			s = __shfl_sync(B2I_FULL, s, 0);
			if (s >= n)
				break;
			const uint32_t idx = order[s];
			const B2iDesc d = descs[idx];
			process_deflate_stream(sm, ring, ts->scratch[0], ts, in, in_total, out, out_mirror, d,
			    &results[idx], crc_tab, xp8);
			__syncwarp();
		}
		team_command(ts, TC_QUIT);
	} else {
		team_serve(ts, w, sm);
	}
	if (lane == 0) {
		__threadfence();
		atomicExch(&slot_busy[slot], 0u);
	}
}

/* ---- stored entries ------------------------------------------------------ */

/*
 * K2: CRC-32 of stored entries, one warp per work item, persistent CTAs (one
 * per SM, 32 warps).  A work item whose bytes are 16-byte aligned and a
 * multiple of 512 long takes the streaming path:
 *
 *   every warp load is one coalesced 512-byte line group (16 bytes per lane);
 *   each lane runs FOUR independent word streams (the k-th word of its uint4,
 *   stride 512 bytes): u <- u * x^4096 mod P  ^  word, i.e. four table
 *   look-ups per word in tables Z0..Z3 (byte j of u advanced by 512 bytes).
 *   The tables are replicated 32 times in shared memory, [table][byte][lane],
 *   so lane l only ever touches bank l: no bank conflicts at random indices.
 *   At the end the four streams of a lane are folded with three ordinary
 *   4-byte steps, each lane's value is multiplied by its position constant
 *   x^(32 + 128*(31-lane)) and the warp XOR-reduces.
 *
 * Anything else (heads, tails, small entries) takes the generic lane-slice
 * routine crc_warp_raw0.  Both produce raw0 of the piece; the combine kernel
 * shifts every partial by the bytes that follow it.
 */
#define CRC_THREADS 1024
#define ZREP_WORDS (4 * 256 * 32)

extern "C" __global__ void __launch_bounds__(CRC_THREADS, 1)
b2i_crc_chunks_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
    const B2iDesc *__restrict__ descs, const B2iCrcWork *__restrict__ work, uint32_t nwork,
    uint32_t *__restrict__ partial, const uint32_t *__restrict__ crc_tab,
    const uint32_t *__restrict__ xp8, const uint32_t *__restrict__ ztab,
    const uint32_t *__restrict__ lane_mul)
{
	extern __shared__ __align__(16) uint32_t zrep[];       /* [4][256][32] */
	__shared__ uint32_t tab[1024];
	const unsigned lane = threadIdx.x & 31;
	const unsigned warps_per_block = blockDim.x >> 5;

	for (int i = threadIdx.x; i < 1024; i += blockDim.x)
		tab[i] = crc_tab[i];
	for (int i = threadIdx.x; i < ZREP_WORDS; i += blockDim.x)
		zrep[i] = ztab[i >> 5];
	__syncthreads();
	const char *zl = (const char *)zrep + lane * 4;            /* this lane's bank */
	const uint32_t my_mul = lane_mul[lane];

	for (uint32_t w = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < nwork;
	    w += gridDim.x * warps_per_block) {
		const B2iCrcWork k = work[w];
		const B2iDesc d = descs[k.entry];
		const uint8_t *src = in + d.in_off + k.rel;
		uint32_t raw0 = 0;
		if (!(d.flags & F_NO_CRC)) {
			if ((((uintptr_t)src & 15) | (k.len & 511)) == 0 && k.len) {
				const uint4 *q = (const uint4 *)src + lane;
				uint32_t u0 = 0, u1 = 0, u2 = 0, u3 = 0;
				const uint32_t iters = k.len >> 9;
#define ZSTEP(u, word) do { \
		uint32_t a0 = (u << 7) & 0x7f80u, a1 = (u >> 1) & 0x7f80u, \
		    a2 = (u >> 9) & 0x7f80u, a3 = (u >> 17) & 0x7f80u; \
		u = *(const uint32_t *)(zl + a0) ^ *(const uint32_t *)(zl + 32768 + a1) ^ \
		    *(const uint32_t *)(zl + 65536 + a2) ^ *(const uint32_t *)(zl + 98304 + a3) ^ (word); \
	} while (0)
				/* four 512-byte groups per trip, their loads issued before the first use */
				uint32_t it = 0;
				for (; it + 4 <= iters; it += 4) {
					const uint4 v0 = __ldg(q), v1 = __ldg(q + 32), v2 = __ldg(q + 64), v3 = __ldg(q + 96);
					q += 128;
					ZSTEP(u0, v0.x); ZSTEP(u1, v0.y); ZSTEP(u2, v0.z); ZSTEP(u3, v0.w);
					ZSTEP(u0, v1.x); ZSTEP(u1, v1.y); ZSTEP(u2, v1.z); ZSTEP(u3, v1.w);
					ZSTEP(u0, v2.x); ZSTEP(u1, v2.y); ZSTEP(u2, v2.z); ZSTEP(u3, v2.w);
					ZSTEP(u0, v3.x); ZSTEP(u1, v3.y); ZSTEP(u2, v3.z); ZSTEP(u3, v3.w);
				}
				for (; it < iters; it++) {
					const uint4 v = __ldg(q);
					q += 32;
					ZSTEP(u0, v.x);
					ZSTEP(u1, v.y);
					ZSTEP(u2, v.z);
					ZSTEP(u3, v.w);
				}
#undef ZSTEP
				/* fold the lane's four streams (4 bytes apart), then place the lane */
				uint32_t v = crc_word(u0, 0, tab) ^ u1;
				v = crc_word(v, 0, tab) ^ u2;
				v = crc_word(v, 0, tab) ^ u3;
				v = crc_mulmod(v, my_mul);
				for (int o = 16; o; o >>= 1)
					v ^= __shfl_xor_sync(B2I_FULL, v, o);
				raw0 = v;
			} else {
				raw0 = crc_warp_raw0(src, k.len, tab, xp8);
			}
		}
		if (lane == 0)
			partial[w] = raw0;
		/* an entry that does not fit its reserved output is not copied at all: the
		 * combine kernel reports S_OUT_OVERFLOW for it and nobody else's bytes are touched */
		if (!(d.flags & F_NO_COPY) && d.in_len <= d.out_cap) {
			uint8_t *dst = out + d.out_off + k.rel;
			if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
				const uint4 *s4 = (const uint4 *)src;
				uint4 *d4 = (uint4 *)dst;
				uint32_t n16 = k.len >> 4;
				for (uint32_t i = lane; i < n16; i += 32)
					d4[i] = s4[i];
				for (uint32_t i = (n16 << 4) + lane; i < k.len; i += 32)
					dst[i] = src[i];
			} else {
				for (uint32_t i = lane; i < k.len; i += 32)
					dst[i] = src[i];
			}
		}
	}
}

extern "C" __global__ void __launch_bounds__(128)
b2i_crc_combine_kernel(const B2iDesc *__restrict__ descs, B2iResult *__restrict__ results,
    const B2iCrcEntry *__restrict__ ents, uint32_t nents, const B2iCrcWork *__restrict__ work,
    const uint32_t *__restrict__ partial, const uint32_t *__restrict__ xp8)
{
	const unsigned lane = threadIdx.x & 31;
	const uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (e >= nents)
		return;
	const B2iCrcEntry ce = ents[e];
	const B2iDesc d = descs[ce.entry];
	uint32_t acc = 0;
	/* raw0(entry) = XOR_j shift(raw0(chunk_j), bytes after chunk j) */
	for (uint32_t j = lane; j < ce.nwork; j += 32) {
		const B2iCrcWork k = work[ce.first_work + j];
		uint64_t after = d.in_len - k.rel - k.len;
		acc ^= crc_mulmod(partial[ce.first_work + j], crc_xpow8(after, xp8));
	}
	for (int o = 16; o; o >>= 1)
		acc ^= __shfl_xor_sync(B2I_FULL, acc, o);
	if (lane == 0) {
		B2iResult r;
		r.status = S_OK;
		r.detail = 0;
		r.flags = 0;
		r.out_bytes = d.in_len;
		r.in_bytes = d.in_len;
		r.crc = 0;
		if (!(d.flags & F_NO_COPY) && d.in_len > d.out_cap) {
			r.status = S_OUT_OVERFLOW;     /* host sized the plan wrongly */
		} else {
			if (!(d.flags & F_NO_CRC)) {
				r.crc = crc_finish(0, acc, d.in_len, xp8);
				if (r.crc != d.expect_crc)
					r.flags |= R_CRC_MISMATCH;
			}
			if ((d.in_len & 0xffffffffull) != (d.expect_out & 0xffffffffull))
				r.flags |= R_OUT_MISMATCH;
		}
		results[ce.entry] = r;
	}
}

extern "C" __global__ void
b2i_unsupported_kernel(const B2iDesc *__restrict__ descs, B2iResult *__restrict__ results,
    const uint32_t *__restrict__ list, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	B2iResult r;
	r.status = S_UNSUPPORTED;
	r.crc = 0; r.out_bytes = 0; r.in_bytes = 0; r.detail = descs[list[i]].method; r.flags = 0;
	results[list[i]] = r;
}

/* one block of 256 threads: tab[k*256+b] and xp8[k] = x^(8*2^k) */
extern "C" __global__ void
b2i_tables_kernel(uint32_t *crc_tab, uint32_t *xp8, uint32_t *ztab, uint32_t *lane_mul)
{
	__shared__ uint32_t t0[256];
	const uint32_t b = threadIdx.x;
	uint32_t c = b;
	for (int k = 0; k < 8; k++)
		c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
	t0[b] = c;
	crc_tab[b] = c;
	__syncthreads();
	for (int k = 1; k < 4; k++) {
		c = t0[c & 0xff] ^ (c >> 8);
		crc_tab[k * 256 + b] = c;
	}
	if (b == 0) {
		uint32_t p = 0x00800000u;       /* x^8 */
		for (int k = 0; k < 40; k++) {
			xp8[k] = p;
			p = crc_mulmod(p, p);
		}
	}
	__syncthreads();
	__threadfence();
	/* ztab[j*256 + b] = (b << 8j) * x^4096 mod P: byte j of a register advanced by 512 bytes */
	const uint32_t x4096 = xp8[9];
	for (int j = 0; j < 4; j++)
		ztab[j * 256 + b] = crc_mulmod(b << (8 * j), x4096);
	/* lane_mul[l] = x^(32 + 128*(31-l)): 4 + 16*(31-l) bytes */
	if (b < 32)
		lane_mul[b] = crc_xpow8(4u + 16u * (31u - b), xp8);
}

/* ---- launch wrappers (called from b2i_api.cpp; plain C++ signatures) ------ */

size_t b2i_inflate_smem_bytes(void) { return sizeof(WarpSmem) * INFLATE_WARPS; }
/* token scratch for a launch on `num_sms` SMs (every resident warp owns one region) */
uint32_t b2i_inflate_scratch_slots(int num_sms) { return (uint32_t)num_sms * INFLATE_SLOTS; }
size_t b2i_inflate_scratch_bytes(int num_sms)
{
	return (size_t)b2i_inflate_scratch_slots(num_sms) * LP_SCRATCH_WORDS * sizeof(uint32_t);
}

cudaError_t b2i_launch_tables(uint32_t *crc_tab, uint32_t *xp8, uint32_t *ztab, uint32_t *lane_mul,
    cudaStream_t st)
{
	b2i_tables_kernel<<<1, 256, 0, st>>>(crc_tab, xp8, ztab, lane_mul);
	return cudaGetLastError();
}

#endif /* !B2I_R9 */

/* Function attributes are per device: b2i_ctx_create calls this with the context's
 * device current (not a process-wide flag: a second GPU needs its own settings). */
#ifdef B2I_R9
cudaError_t b2i_kernels_configure_r9(void)
#else
cudaError_t b2i_kernels_configure_r9(void);
cudaError_t b2i_kernels_configure(void)
#endif
{
	cudaError_t e = cudaFuncSetAttribute(b2i_inflate_kernel,
	    cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	if (e != cudaSuccess)
		return e;
	e = cudaFuncSetAttribute(b2i_inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    (int)(sizeof(WarpSmem) * INFLATE_WARPS));
	if (e != cudaSuccess)
		return e;
#ifndef B2I_R9
	e = cudaFuncSetAttribute(b2i_inflate_team_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    (int)(sizeof(WarpSmem) + sizeof(TeamShared)));
	if (e != cudaSuccess)
		return e;
	e = cudaFuncSetAttribute(b2i_inflate_team_kernel,
	    cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	if (e != cudaSuccess)
		return e;
	e = cudaFuncSetAttribute(b2i_crc_chunks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	    (int)(ZREP_WORDS * sizeof(uint32_t)));
	if (e != cudaSuccess)
		return e;
	e = b2i_kernels_configure_r9();
#endif
	return e;
}

cudaError_t b2i_launch_inflate(const uint8_t *in, uint64_t in_total, uint8_t *out, uint8_t *out_mirror,
    const B2iDesc *descs, B2iResult *results, const uint32_t *order, uint32_t n,
    unsigned int *counter, const uint32_t *crc_tab, const uint32_t *xp8, uint32_t *scratch,
    unsigned int *slot_busy, int num_sms, cudaStream_t st)
{
	const size_t smem = sizeof(WarpSmem) * INFLATE_WARPS;
	uint32_t blocks = (n + INFLATE_WARPS - 1) / INFLATE_WARPS;
	uint32_t max_blocks = (uint32_t)num_sms * INFLATE_CTAS;
	if (blocks > max_blocks)
		blocks = max_blocks;
	b2i_inflate_kernel<<<blocks, INFLATE_WARPS * 32, smem, st>>>(in, in_total, out, out_mirror, descs,
	    results, order, n, counter, crc_tab, xp8, scratch, slot_busy, b2i_inflate_scratch_slots(num_sms));
	return cudaGetLastError();
}

#ifndef B2I_R9
cudaError_t b2i_launch_inflate_team(const uint8_t *in, uint64_t in_total, uint8_t *out, uint8_t *out_mirror,
    const B2iDesc *descs, B2iResult *results, const uint32_t *order, uint32_t n,
    unsigned int *counter, const uint32_t *crc_tab, const uint32_t *xp8, uint32_t *scratch,
    unsigned int *slot_busy, int num_sms, cudaStream_t st)
{
	const size_t smem = sizeof(WarpSmem) + sizeof(TeamShared);
	uint32_t blocks = n < (uint32_t)num_sms * TEAM_CTAS_PER_SM ? n : (uint32_t)num_sms * TEAM_CTAS_PER_SM;
	b2i_inflate_team_kernel<<<blocks, TEAM_LANES, smem, st>>>(in, in_total, out, out_mirror, descs,
	    results, order, n, counter, crc_tab, xp8, scratch, slot_busy, b2i_inflate_scratch_slots(num_sms));
	return cudaGetLastError();
}

cudaError_t b2i_launch_crc_chunks(const uint8_t *in, uint8_t *out, const B2iDesc *descs,
    const B2iCrcWork *work, uint32_t nwork, uint32_t *partial, const uint32_t *crc_tab,
    const uint32_t *xp8, const uint32_t *ztab, const uint32_t *lane_mul, int num_sms, cudaStream_t st)
{
	const size_t smem = ZREP_WORDS * sizeof(uint32_t);
	uint32_t blocks = (nwork + 31) / 32;
	if (blocks > (uint32_t)num_sms)
		blocks = (uint32_t)num_sms;
	b2i_crc_chunks_kernel<<<blocks, CRC_THREADS, smem, st>>>(in, out, descs, work, nwork, partial,
	    crc_tab, xp8, ztab, lane_mul);
	return cudaGetLastError();
}

cudaError_t b2i_launch_crc_combine(const B2iDesc *descs, B2iResult *results,
    const B2iCrcEntry *ents, uint32_t nents, const B2iCrcWork *work, const uint32_t *partial,
    const uint32_t *xp8, cudaStream_t st)
{
	b2i_crc_combine_kernel<<<(nents + 3) / 4, 128, 0, st>>>(descs, results, ents, nents, work,
	    partial, xp8);
	return cudaGetLastError();
}

cudaError_t b2i_launch_unsupported(const B2iDesc *descs, B2iResult *results,
    const uint32_t *list, uint32_t n, cudaStream_t st)
{
	b2i_unsupported_kernel<<<(n + 127) / 128, 128, 0, st>>>(descs, results, list, n);
	return cudaGetLastError();
}
#endif /* !B2I_R9 */
