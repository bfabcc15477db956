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
 */
#include <pthread.h>
#include <stddef.h>

#include "b200inflate.h"
#include "b200_ctx_pool.h"

#define POOL_MAX 4

static pthread_mutex_t pool_lock = PTHREAD_MUTEX_INITIALIZER;
static b2i_ctx *pool[POOL_MAX];
static int pool_n;

int
b200_ctx_acquire(b2i_ctx **out)
{
	pthread_mutex_lock(&pool_lock);
	if (pool_n > 0) {
		*out = pool[--pool_n];
		pthread_mutex_unlock(&pool_lock);
		return (B2I_OK);
	}
	pthread_mutex_unlock(&pool_lock);
	return (b2i_ctx_create(0, NULL, out));
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
		pool[pool_n++] = c;
		c = NULL;
	}
	pthread_mutex_unlock(&pool_lock);
	if (c != NULL)
		b2i_ctx_destroy(c);
}
