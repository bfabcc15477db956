/*
 * warp_emul.h — TEST INFRASTRUCTURE ONLY.
 *
 * Lets tests compile libarchive_b200/csrc/{inflate,crc32,stream}_core.cuh for
 * the host: 32 pthreads stand in for the 32 lanes of one warp, and every warp
 * collective (__shfl*, __ballot, __match_any, __reduce_add, __syncwarp) is a
 * barrier-synchronised exchange.  This checks the warp-cooperative logic
 * (table construction, batch resolution, CRC merge tree) without a GPU; it is
 * never linked into the product library, which has no CPU path.
 */
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#define B2I_DEV static inline
#define B2I_DEV_NOINLINE static
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };

struct EmulWarp {
	pthread_barrier_t bar;
	uint64_t xchg[32];
};
extern thread_local unsigned tl_lane;
extern thread_local EmulWarp *tl_warp;
extern thread_local pthread_barrier_t *tl_team;     /* all lanes of all warps of a team */
static inline void team_sync() { pthread_barrier_wait(tl_team); }

static inline unsigned b2i_lane() { return tl_lane; }
static inline void emul_sync() { pthread_barrier_wait(&tl_warp->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul_sync(); }

static inline uint64_t emul_xchg(uint64_t v, unsigned src)
{
	tl_warp->xchg[tl_lane] = v;
	emul_sync();
	uint64_t r = tl_warp->xchg[src & 31];
	emul_sync();
	return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{ return (T)emul_xchg((uint64_t)v, (unsigned)src); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{ return tl_lane >= d ? (T)emul_xchg((uint64_t)v, tl_lane - d) : (emul_xchg((uint64_t)v, tl_lane), v); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, unsigned m)
{ return (T)emul_xchg((uint64_t)v, tl_lane ^ m); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
	tl_warp->xchg[tl_lane] = pred ? 1 : 0;
	emul_sync();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r |= (unsigned)tl_warp->xchg[i] << i;
	emul_sync();
	return r;
}
static inline int __any_sync(unsigned m, int pred);
static inline unsigned __match_any_sync(unsigned, unsigned v)
{
	tl_warp->xchg[tl_lane] = v;
	emul_sync();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if ((unsigned)tl_warp->xchg[i] == v) r |= 1u << i;
	emul_sync();
	return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
	tl_warp->xchg[tl_lane] = v;
	emul_sync();
	unsigned r = 0;
	for (int i = 0; i < 32; i++) r += (unsigned)tl_warp->xchg[i];
	emul_sync();
	return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __sync_fetch_and_or(p, v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __brev(unsigned v)
{
	v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
	v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
	v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
	return __builtin_bswap32(v);
}
