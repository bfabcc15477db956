/*
 * b200_ctx_pool.c — device contexts shared by the archive handles of a process.
 *
 * A b2i_ctx owns streams, pinned staging and the token scratch of the inflate
 * kernel; creating one costs milliseconds, which is more than decoding a small
 * archive.  Handles therefore borrow a context for their lifetime and give it
 * back on cleanup; idle contexts wait here for the next handle.  One handle
 * uses one context at a time (libarchive handles are single-threaded,
 * README.md:217-220), different handles may live on different threads, hence
 * the lock.
 *
 * The same goes for the pinned host buffers the decoded bytes land in: pinning
 * memory costs about as much per byte as decoding it (a 256 MiB cudaHostAlloc
 * takes longer than the whole device pass), so buffers are kept and handed to the
 * next handle that needs one.
 */
#include <pthread.h>
#include <stddef.h>

#include "b200inflate.h"
#include "b200_ctx_pool.h"

#define POOL_MAX 24

static pthread_mutex_t pool_lock = PTHREAD_MUTEX_INITIALIZER;
static struct { b2i_ctx *c; int dev; } pool[POOL_MAX];
static int pool_n;

/* a context on GPU `device` (the streaming engine spreads a large archive over several) */
int
b200_ctx_acquire_dev(int device, b2i_ctx **out)
{
	int i;

	pthread_mutex_lock(&pool_lock);
	for (i = 0; i < pool_n; i++)
		if (pool[i].dev == device) {
			*out = pool[i].c;
			pool[i] = pool[--pool_n];
			pthread_mutex_unlock(&pool_lock);
			return (B2I_OK);
		}
	pthread_mutex_unlock(&pool_lock);
	return (b2i_ctx_create(device, NULL, out));
}

int
b200_ctx_acquire(b2i_ctx **out)
{
	return (b200_ctx_acquire_dev(0, out));
}

void
b200_ctx_release(b2i_ctx *c, int healthy)
{
	if (c == NULL)
		return;
	if (!healthy) {                 /* after a device error the context is not reused */
		b2i_ctx_destroy(c);
		return;
	}
	pthread_mutex_lock(&pool_lock);
	if (pool_n < POOL_MAX) {
		pool[pool_n].c = c;
		pool[pool_n].dev = b2i_ctx_device(c);
		pool_n++;
		c = NULL;
	}
	pthread_mutex_unlock(&pool_lock);
	if (c != NULL)
		b2i_ctx_destroy(c);
}

/* ---- pinned output buffers ----------------------------------------------------
 * Sizes are rounded to powers of two (256 KiB at least) so that a run of archives of
 * similar size shares one buffer; when the shelf is full the SMALLEST buffers make
 * room - a process that alternates small and large archives keeps the large one,
 * which is the one that is expensive to pin again. */
#define BUF_MAX    8
#define BUF_MIN    ((size_t)256 << 10)
#define BUF_KEEP   ((size_t)2 << 30)     /* total bytes kept idle at most */

static struct { void *p; size_t cap; } bufs[BUF_MAX + 1];
static int bufs_n;

void *
b200_buf_acquire(size_t need, size_t *cap)
{
	void *p = NULL;
	int best = -1;
	size_t want = BUF_MIN;

	pthread_mutex_lock(&pool_lock);
	for (int i = 0; i < bufs_n; i++)
		if (bufs[i].cap >= need && (best < 0 || bufs[i].cap < bufs[best].cap))
			best = i;
	if (best >= 0) {
		p = bufs[best].p;
		*cap = bufs[best].cap;
		bufs[best] = bufs[--bufs_n];
	}
	pthread_mutex_unlock(&pool_lock);
	if (p != NULL)
		return (p);
	while (want < need && want < ((size_t)1 << 30))
		want <<= 1;
	if (want < need)                      /* beyond 1 GiB: an eighth of headroom */
		want = need + need / 8;
	*cap = want;
	if ((p = b2i_host_alloc(*cap)) == NULL) {
		*cap = need;
		p = b2i_host_alloc(need);
	}
	return (p);
}

void
b200_buf_release(void *p, size_t cap)
{
	void *drop[BUF_MAX + 1];
	int ndrop = 0;
	size_t idle = 0;

	if (p == NULL)
		return;
	pthread_mutex_lock(&pool_lock);
	bufs[bufs_n].p = p;                   /* the array has one slot more than the shelf */
	bufs[bufs_n].cap = cap;
	bufs_n++;
	for (int i = 0; i < bufs_n; i++)
		idle += bufs[i].cap;
	while (bufs_n > BUF_MAX || (bufs_n > 0 && idle > BUF_KEEP)) {
		int s = 0;                    /* the smallest goes first */
		for (int i = 1; i < bufs_n; i++)
			if (bufs[i].cap < bufs[s].cap)
				s = i;
		idle -= bufs[s].cap;
		drop[ndrop++] = bufs[s].p;
		bufs[s] = bufs[--bufs_n];
	}
	pthread_mutex_unlock(&pool_lock);
	for (int i = 0; i < ndrop; i++)
		b2i_host_free(drop[i]);
}

/* the same shelf for callers that do not keep the capacity: it rides in front of the bytes */
#define BUF_HDR 64

void *
b200_buf_acquire_tagged(size_t need)
{
	size_t cap;
	unsigned char *p = b200_buf_acquire(need + BUF_HDR, &cap);

	if (p == NULL)
		return (NULL);
	*(size_t *)(void *)p = cap;
	return (p + BUF_HDR);
}

void
b200_buf_release_tagged(void *q)
{
	if (q != NULL) {
		unsigned char *p = (unsigned char *)q - BUF_HDR;
		b200_buf_release(p, *(size_t *)(void *)p);
	}
}
