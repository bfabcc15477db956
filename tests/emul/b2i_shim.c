/*
 * b2i_shim.c — TEST INFRASTRUCTURE ONLY, never part of the product.
 *
 * Answers the part of the C ABI (include/b200inflate.h) that the libarchive
 * plugin modules call with the CPU oracle (oracle/liboracle.so), so that the
 * HOST logic of the plugins — header walking, block contract, error mapping,
 * the state machine — can be run through the reference's own test programs in
 * the no-GPU suite (tests/refsuite).  The product library libb200inflate.so
 * has no such path: without a GPU b2i_ctx_create fails.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "b200inflate.h"
#include "oracle.h"
#include <stdio.h>

/* B2I_SHIM_STATS=1: report how many device calls / streams the plugins made */
static long g_calls, g_streams;
static void shim_report(void) { if (getenv("B2I_SHIM_STATS")) fprintf(stderr, "shim: %ld decode calls, %ld streams\n", g_calls, g_streams); }

struct b2i_ctx { void *keep; };   /* the last finished job's input copy (b2i_job_staged_input) */

int b2i_ctx_create(int device, void *cuda_stream, b2i_ctx **out)
{
	(void)device; (void)cuda_stream;
	*out = calloc(1, sizeof(struct b2i_ctx));
	return *out ? B2I_OK : B2I_E_NOMEM;
}
void b2i_ctx_destroy(b2i_ctx *c) { if (c) free(c->keep); free(c); }
const char *b2i_last_error(const b2i_ctx *c) { (void)c; return "shim"; }
void *b2i_host_alloc(size_t bytes) { if (getenv("B2I_SHIM_STATS") && bytes > (32u << 20)) fprintf(stderr, "shim: host_alloc %zu\n", bytes); return malloc(bytes ? bytes : 1); }
void b2i_host_free(void *p) { free(p); }
void b2i_free(void *p) { free(p); }

int b2i_crc32(b2i_ctx *c, uint32_t crc, const void *buf, size_t len, uint32_t *out)
{
	(void)c;
	*out = orc_crc32(crc, buf, len);
	return B2I_OK;
}

int b2i_decode_host(b2i_ctx *c, const void *host_in, size_t in_bytes, const b2i_stream_desc *descs,
    size_t n, void *host_out, size_t out_bytes, b2i_stream_result *results)
{
	(void)c;
	if (__atomic_fetch_add(&g_calls, 1, __ATOMIC_RELAXED) == 0)      /* several worker threads call in */
		atexit(shim_report);
	__atomic_fetch_add(&g_streams, (long)n, __ATOMIC_RELAXED);
	_Static_assert(sizeof(orc_desc) == sizeof(b2i_stream_desc), "descriptor layouts differ");
	_Static_assert(sizeof(orc_stream_result) == sizeof(b2i_stream_result), "result layouts differ");
	return orc_decode_batch(host_in, in_bytes, (const orc_desc *)descs, n, host_out, out_bytes,
	    (orc_stream_result *)results) == 0 ? B2I_OK : B2I_E_INVAL;
}

/* ---- what the streaming engine (csrc/b2i_pipe.cpp) and the plugins' pool need ---- */
int b2i_device_count(void) { const char *e = getenv("B2I_SHIM_GPUS"); return e ? atoi(e) : 1; }
int b2i_ctx_device(const b2i_ctx *c) { (void)c; return 0; }

struct b2i_job {
	b2i_ctx *c;
	const void *in; size_t in_bytes;
	b2i_stream_desc *descs; size_t n;
	void *out; size_t out_bytes;
	void *owned;
};

int b2i_submit(b2i_ctx *c, const void *host_in, size_t in_bytes, const b2i_stream_desc *descs, size_t n,
    void *host_out, size_t out_bytes, b2i_job **job)
{
	struct b2i_job *j = calloc(1, sizeof(*j));
	if (j == NULL)
		return B2I_E_NOMEM;
	free(c->keep);                 /* "until the next b2i_submit on the context" */
	c->keep = NULL;
	j->c = c; j->in = host_in; j->in_bytes = in_bytes; j->n = n; j->out = host_out; j->out_bytes = out_bytes;
	j->descs = malloc((n ? n : 1) * sizeof(*descs));
	if (n) memcpy(j->descs, descs, n * sizeof(*descs));
	/* like the library, inputs of 4 MiB and more are "staged" at once (b2i_job_staged_input) */
	if (n && in_bytes >= ((size_t)4 << 20) && (j->owned = malloc(in_bytes)) != NULL) {
		memcpy(j->owned, host_in, in_bytes);
		j->in = j->owned;
	}
	*job = j;
	return B2I_OK;
}

/* the shim decodes at wait time, from the caller's memory: keep a copy of the input */
int b2i_job_wait_input(b2i_job *j)
{
	if (j->n && j->owned == NULL) {
		j->owned = malloc(j->in_bytes ? j->in_bytes : 1);
		if (j->owned == NULL)
			return B2I_E_NOMEM;
		memcpy(j->owned, j->in, j->in_bytes);
		j->in = j->owned;
	}
	return B2I_OK;
}

const void *b2i_job_staged_input(const b2i_job *j) { return j->owned; }

int b2i_wait(b2i_job *j, b2i_stream_result *res)
{
	int rc = j->n ? b2i_decode_host(j->c, j->in, j->in_bytes, j->descs, j->n, j->out, j->out_bytes, res) : B2I_OK;
	free(j->descs);
	free(j->c->keep);
	j->c->keep = j->owned;
	free(j);
	return rc;
}
