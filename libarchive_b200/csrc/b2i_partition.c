/*
 * b2i_partition.c — host-side partition of independent deflate streams over the
 * GPUs of one box.  ZIP entries and BGZF members never share a
 * window (inflateReset per entry, archive_read_support_format_zip.c:2517-2518;
 * inflateInit2 per member, archive_read_support_filter_gzip.c:363), so the
 * partition is pure host arithmetic and no collective follows it.
 *
 *   b2i_partition_contiguous  G contiguous descriptor ranges of about equal
 *                             weight: one compressed byte range in, one output
 *                             range back per GPU;
 *   b2i_partition_lpt         longest-processing-time greedy for batches whose
 *                             largest streams are a large share of the total
 *                             (BASELINE config 4): streams by falling weight,
 *                             each to the least loaded GPU.
 * Weight of a stream = in_len + max(expect_out, out_cap): the bytes it moves.
 */
#include <stdlib.h>
#include <string.h>
#include "../../include/b200inflate.h"

static uint64_t
weight_of(const b2i_stream_desc *d)
{
	uint64_t o = d->expect_out > d->out_cap ? d->expect_out : d->out_cap;
	return d->in_len + o + 1;       /* +1: empty streams still cost a visit */
}

int
b2i_partition_contiguous(const b2i_stream_desc *descs, size_t n, int parts, size_t *cuts)
{
	uint64_t total = 0, acc = 0;
	size_t i;
	int k = 1;

	if (parts < 1 || cuts == NULL || (n && descs == NULL))
		return B2I_E_INVAL;
	for (i = 0; i < n; i++)
		total += weight_of(&descs[i]);
	cuts[0] = 0;
	for (i = 0; i < n && k < parts; i++) {
		const uint64_t w = weight_of(&descs[i]);
		/* close range k-1 in front of the stream whose middle passes k/parts of the total */
		while (k < parts && (long double)acc + (long double)w / 2 >=
		    (long double)total * k / parts)
			cuts[k++] = i;
		acc += w;
	}
	while (k <= parts)
		cuts[k++] = n;
	return B2I_OK;
}

struct wi { uint64_t w; uint32_t i; };

static int
by_weight_desc(const void *a, const void *b)
{
	const struct wi *x = a, *y = b;
	if (x->w != y->w)
		return x->w < y->w ? 1 : -1;
	return x->i < y->i ? -1 : (x->i > y->i);
}

int
b2i_partition_lpt(const b2i_stream_desc *descs, size_t n, int parts, uint32_t *owner,
    uint64_t *load)
{
	struct wi *v;
	uint64_t *ld;
	size_t i;
	int p;

	if (parts < 1 || (n && (descs == NULL || owner == NULL)) || n > 0xffffffffu)
		return B2I_E_INVAL;
	v = malloc((n ? n : 1) * sizeof(*v));
	ld = calloc((size_t)parts, sizeof(*ld));
	if (v == NULL || ld == NULL) {
		free(v);
		free(ld);
		return B2I_E_NOMEM;
	}
	for (i = 0; i < n; i++) {
		v[i].w = weight_of(&descs[i]);
		v[i].i = (uint32_t)i;
	}
	qsort(v, n, sizeof(*v), by_weight_desc);
	for (i = 0; i < n; i++) {
		int best = 0;
		for (p = 1; p < parts; p++)
			if (ld[p] < ld[best])
				best = p;
		owner[v[i].i] = (uint32_t)best;
		ld[best] += v[i].w;
	}
	if (load != NULL)
		memcpy(load, ld, (size_t)parts * sizeof(*ld));
	free(v);
	free(ld);
	return B2I_OK;
}
