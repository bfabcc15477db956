/*
 * api_bench.c — throughput of an archive through libarchive's PUBLIC read API
 * (archive_read_open_memory / archive_read_open_filename, archive_read_next_header,
 * archive_read_data_block or archive_read_data into a 64 KiB buffer): the call
 * sequence BASELINE config 1 names.  Linked against the drop-in library it measures
 * the B200 path end to end as a caller of libarchive sees it (source bytes in
 * pageable host memory or in a file, decoded bytes delivered to the caller); no
 * symbol of this repo is used directly.
 *
 *   api_bench <file> [--raw] [--mode block|data] [--steps K] [--warmup W] [--file]
 *             [--check]    (--check: zlib crc32 over everything delivered, per step)
 * Prints one JSON line.
 */
#include <archive.h>
#include <archive_entry.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

static double
now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void *
slurp(const char *path, size_t *len)
{
	FILE *f = fopen(path, "rb");
	void *buf;
	long n;

	if (f == NULL) { perror(path); exit(2); }
	fseek(f, 0, SEEK_END);
	n = ftell(f);
	fseek(f, 0, SEEK_SET);
	buf = malloc((size_t)n + 64);
	if (buf == NULL || fread(buf, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "read failed\n"); exit(2); }
	fclose(f);
	*len = (size_t)n;
	return buf;
}

struct pass { uint64_t bytes, entries, errors; uint32_t crc; double seconds, t_open, t_first_header, t_first_block; };

static struct pass
one_pass(const char *path, const void *buf, size_t len, int raw, int use_file, int mode_data, int check)
{
	struct pass r = { 0, 0, 0, 0, 0 };
	struct archive *a = archive_read_new();
	struct archive_entry *e;
	static char out[65536];
	double t0 = now();
	int rc;

	if (raw) {
		archive_read_support_filter_gzip(a);
		archive_read_support_format_raw(a);
	} else {
		archive_read_support_format_zip(a);
	}
	rc = use_file ? archive_read_open_filename(a, path, 1 << 20) : archive_read_open_memory(a, buf, len);
	if (rc != ARCHIVE_OK) {
		fprintf(stderr, "open: %s\n", archive_error_string(a));
		r.errors++;
		archive_read_free(a);
		return r;
	}
	r.t_open = now() - t0;
	while ((rc = archive_read_next_header(a, &e)) == ARCHIVE_OK || rc == ARCHIVE_WARN) {
		if (r.entries == 0)
			r.t_first_header = now() - t0;
		r.entries++;
		if (mode_data == 2) {
			archive_read_data_skip(a);           /* headers only: the per-entry cost of the read core */
			continue;
		}
		if (mode_data) {
			la_ssize_t n;
			while ((n = archive_read_data(a, out, sizeof(out))) > 0) {
				r.bytes += (uint64_t)n;
				if (check) r.crc = (uint32_t)crc32(r.crc, (const Bytef *)out, (uInt)n);
			}
			if (n < 0) {
				if (r.errors++ == 0) fprintf(stderr, "read_data: %s\n", archive_error_string(a));
			}
		} else {
			const void *p;
			size_t n;
			la_int64_t off;
			while ((rc = archive_read_data_block(a, &p, &n, &off)) == ARCHIVE_OK) {
				if (r.bytes == 0)
					r.t_first_block = now() - t0;
				r.bytes += n;
				if (check) r.crc = (uint32_t)crc32(r.crc, p, (uInt)n);
			}
			if (rc != ARCHIVE_EOF) {
				if (r.errors++ == 0) fprintf(stderr, "read_data_block: %s\n", archive_error_string(a));
			}
		}
	}
	if (rc != ARCHIVE_EOF)
		r.errors++;
	archive_read_free(a);
	r.seconds = now() - t0;
	return r;
}

/* peak and current resident memory of this process, from /proc/self/status (kB) */
static void
mem_status(long *hwm, long *anon, long *file, long *shmem)
{
	FILE *f = fopen("/proc/self/status", "r");
	char line[256];

	*hwm = *anon = *file = *shmem = -1;
	if (f == NULL)
		return;
	while (fgets(line, sizeof(line), f) != NULL) {
		sscanf(line, "VmHWM: %ld", hwm);
		sscanf(line, "RssAnon: %ld", anon);
		sscanf(line, "RssFile: %ld", file);
		sscanf(line, "RssShmem: %ld", shmem);
	}
	fclose(f);
}

int
main(int argc, char **argv)
{
	int raw = 0, use_file = 0, steps = 5, warmup = 2, mode_data = 0, check = 0, i;
	size_t len = 0;
	void *buf = NULL;
	double sum = 0, best = 1e30;
	struct pass last = { 0, 0, 0, 0, 0 };
	uint64_t errors = 0;

	if (argc < 2) {
		fprintf(stderr, "usage: %s <file> [--raw] [--mode block|data] [--steps K] [--warmup W] [--file] [--check]\n", argv[0]);
		return 2;
	}
	for (i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--raw")) raw = 1;
		else if (!strcmp(argv[i], "--file")) use_file = 1;
		else if (!strcmp(argv[i], "--check")) check = 1;
		else if (!strcmp(argv[i], "--mode") && i + 1 < argc) {
			i++;
			mode_data = !strcmp(argv[i], "data") ? 1 : !strcmp(argv[i], "headers") ? 2 : 0;
		}
		else if (!strcmp(argv[i], "--steps") && i + 1 < argc) steps = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warmup = atoi(argv[++i]);
	}
	if (!use_file)
		buf = slurp(argv[1], &len);
	for (i = 0; i < warmup + steps; i++) {
		struct pass r = one_pass(argv[1], buf, len, raw, use_file, mode_data, check);
		errors += r.errors;
		if (i >= warmup) {
			sum += r.seconds;
			if (r.seconds < best) best = r.seconds;
		}
		last = r;
	}
	long hwm, anon, filek, shmem;
	mem_status(&hwm, &anon, &filek, &shmem);
	printf("{\"vm_hwm_kb\":%ld,\"rss_anon_kb\":%ld,\"rss_file_kb\":%ld,\"rss_shmem_kb\":%ld,", hwm, anon, filek, shmem);
	printf("\"api\":\"%s\",\"source\":\"%s\",\"steps\":%d,\"warmup\":%d,\"bytes\":%llu,\"entries\":%llu,"
	    "\"errors\":%llu,\"seconds_mean\":%.6f,\"seconds_best\":%.6f,\"gbps_mean\":%.4f,\"crc\":\"%08x\","
	    "\"last_pass\":{\"open_s\":%.6f,\"first_header_s\":%.6f,\"first_block_s\":%.6f,\"total_s\":%.6f},"
	    "\"libarchive\":\"%s\"}\n",
	    mode_data == 1 ? "archive_read_data(64KiB)" : mode_data == 2 ? "headers only" : "archive_read_data_block",
	    use_file ? "open_filename" : "open_memory",
	    steps, warmup, (unsigned long long)last.bytes, (unsigned long long)last.entries,
	    (unsigned long long)errors, steps ? sum / steps : 0, best, steps && sum > 0 ? last.bytes * steps / sum / 1e9 : 0,
	    last.crc, last.t_open, last.t_first_header, last.t_first_block, last.seconds, archive_version_string());
	return errors ? 1 : 0;
}
