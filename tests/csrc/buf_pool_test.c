/*
 * buf_pool_test.c — TEST ONLY: the plugins' shelf of pinned buffers
 * (libarchive_b200/csrc/plugin/b200_ctx_pool.c) with malloc standing in for
 * cudaHostAlloc.  Prints "ok" or the first broken expectation.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200inflate.h"
#include "b200_ctx_pool.h"

static long live, allocs, frees;
static size_t live_bytes;
struct hdr { size_t n; size_t pad; };

void *b2i_host_alloc(size_t n)
{
	struct hdr *h = malloc(sizeof(*h) + (n > (64u << 20) ? 128 : n));   /* sizes are what matters */
	h->n = n; live++; allocs++; live_bytes += n;
	return (h + 1);
}
void b2i_host_free(void *p)
{
	if (p == NULL) return;
	struct hdr *h = (struct hdr *)p - 1;
	live--; frees++; live_bytes -= h->n;
	free(h);
}
/* the context half of the file is not exercised here */
int b2i_ctx_create(int d, void *s, b2i_ctx **o) { (void)d; (void)s; *o = NULL; return B2I_E_NODEVICE; }
void b2i_ctx_destroy(b2i_ctx *c) { (void)c; }
int b2i_ctx_device(const b2i_ctx *c) { (void)c; return 0; }

#define EXPECT(c) do { if (!(c)) { printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(void)
{
	size_t cap, cap2;
	void *p, *q;

	/* size classes: powers of two, 256 KiB at least; a released buffer is reused */
	p = b200_buf_acquire(1000, &cap);
	EXPECT(p != NULL && cap == (256u << 10));
	b200_buf_release(p, cap);
	q = b200_buf_acquire(200000, &cap2);
	EXPECT(q == p && cap2 == cap && allocs == 1);
	b200_buf_release(q, cap2);
	p = b200_buf_acquire((3u << 20) + 5, &cap);
	EXPECT(cap == (4u << 20) && allocs == 2);
	b200_buf_release(p, cap);

	/* small and large archives alternating: the shelf fills with small buffers, the large one
	 * must still be there when it is wanted again (the smallest make room, not the newcomer) */
	{
		void *small[12];
		size_t sc[12];
		for (int i = 0; i < 12; i++)
			small[i] = b200_buf_acquire((size_t)(300 + i) << 10, &sc[i]);   /* all 512 KiB class */
		for (int i = 0; i < 12; i++)
			b200_buf_release(small[i], sc[i]);
	}
	{
		long before = allocs;
		p = b200_buf_acquire(100u << 20, &cap);
		EXPECT(cap == (128u << 20) && allocs == before + 1);
		b200_buf_release(p, cap);
		for (int r = 0; r < 5; r++) {
			q = b200_buf_acquire(400u << 10, &cap2);
			b200_buf_release(q, cap2);
			p = b200_buf_acquire(90u << 20, &cap);
			EXPECT(cap == (128u << 20));
			b200_buf_release(p, cap);
		}
		EXPECT(allocs == before + 1);
	}
	/* the idle total is bounded (2 GiB): three 1 GiB buffers cannot all stay */
	{
		void *big[3];
		size_t bc[3];
		for (int i = 0; i < 3; i++)
			big[i] = b200_buf_acquire((size_t)1 << 30, &bc[i]);
		for (int i = 0; i < 3; i++)
			b200_buf_release(big[i], bc[i]);
		EXPECT(live_bytes <= ((size_t)2 << 30) + (1u << 20));
	}
	/* tagged buffers carry their capacity and go back to the same shelf */
	{
		long before = allocs;
		unsigned char *t = b200_buf_acquire_tagged(50000);
		EXPECT(t != NULL);
		memset(t, 0xAB, 48);       /* the stand-in allocator backs large sizes with 64 bytes only (header included) */
		b200_buf_release_tagged(t);
		t = b200_buf_acquire_tagged(60000);
		b200_buf_release_tagged(t);
		EXPECT(allocs <= before + 1);
		b200_buf_release_tagged(NULL);
	}
	printf("ok allocs=%ld frees=%ld live=%ld\n", allocs, frees, live);
	return 0;
}
