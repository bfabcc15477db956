/*
 * mt_read.c — TEST ONLY: several threads, each with its own archive handle, read the same
 * in-memory ZIP (or gzip with --raw) through libarchive's public API at the same time;
 * every thread must see the same per-archive CRC-32 of all bytes.  Linked against the
 * host-logic (or the drop-in) libarchive: exercises the context pool, the shelf of pinned
 * buffers, the copy threads and the streaming engine under concurrency (ThreadSanitizer
 * build: make -C tests/refsuite tsan).
 * usage: mt_read <file> <threads> <rounds> [--raw]
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <archive.h>
#include <archive_entry.h>
#include <zlib.h>

static unsigned char *image;
static size_t image_len;
static int rounds, raw;

struct result { uint32_t crc; uint64_t bytes; int entries, errors; };

static void *
worker(void *arg)
{
	struct result *res = arg;
	for (int r = 0; r < rounds; r++) {
		struct archive *a = archive_read_new();
		struct archive_entry *e;
		uint32_t crc = 0;
		uint64_t bytes = 0;
		int entries = 0;

		archive_read_support_filter_gzip(a);
		if (raw) archive_read_support_format_raw(a); else archive_read_support_format_zip(a);
		if (archive_read_open_memory(a, image, image_len) != ARCHIVE_OK) { res->errors++; archive_read_free(a); continue; }
		while (archive_read_next_header(a, &e) == ARCHIVE_OK) {
			const void *b; size_t n; int64_t off; int rc;
			entries++;
			while ((rc = archive_read_data_block(a, &b, &n, &off)) == ARCHIVE_OK) {
				crc = (uint32_t)crc32(crc, b, (unsigned)n);
				bytes += n;
			}
			if (rc != ARCHIVE_EOF) res->errors++;
		}
		archive_read_free(a);
		if (r == 0) { res->crc = crc; res->bytes = bytes; res->entries = entries; }
		else if (res->crc != crc || res->bytes != bytes) res->errors++;
	}
	return NULL;
}

int
main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: mt_read <file> <threads> <rounds> [--raw]\n"); return 2; }
	FILE *f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	fseek(f, 0, SEEK_END); image_len = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
	image = malloc(image_len + 1);
	if (fread(image, 1, image_len, f) != image_len) return 2;
	fclose(f);
	int nt = atoi(argv[2]);
	rounds = atoi(argv[3]);
	raw = argc > 4 && strcmp(argv[4], "--raw") == 0;
	pthread_t th[64];
	struct result res[64];
	memset(res, 0, sizeof(res));
	if (nt > 64) nt = 64;
	for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, worker, &res[i]);
	int bad = 0;
	for (int i = 0; i < nt; i++) {
		pthread_join(th[i], NULL);
		if (res[i].errors || res[i].crc != res[0].crc || res[i].bytes != res[0].bytes) bad++;
	}
	printf("{\"threads\":%d,\"rounds\":%d,\"entries\":%d,\"bytes\":%llu,\"crc\":\"%08x\",\"bad_threads\":%d}\n",
	    nt, rounds, res[0].entries, (unsigned long long)res[0].bytes, res[0].crc, bad);
	return bad ? 1 : 0;
}
