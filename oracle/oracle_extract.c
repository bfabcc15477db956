/*
 * oracle_extract.c — TEST INFRASTRUCTURE ONLY.
 *
 * Driver around the UNMODIFIED reference (oracle/_ref/libarchive_ref.so, built
 * by oracle/Makefile from the sources where they lie under /root/reference)
 * using only its public API (archive.h): archive_read_new,
 * archive_read_support_format_zip / _raw, archive_read_support_filter_gzip,
 * archive_read_open_memory, archive_read_next_header, archive_read_data_block.
 *
 *   oracle_extract list  <file> [--raw] [--opt zip:ignorecrc32] [--dump out.bin] [--stream BLOCK] [--meta]
 *       one JSON line per entry: name, size, header/read return codes, bytes
 *       read, CRC-32 (zlib) of the bytes read, block sizes, error string
 *   oracle_extract bench <file> [--raw] --procs P [--reps R]
 *       CPU baseline: P forked processes, each reading a disjoint contiguous
 *       shard of the entries (ZIP) or of the BGZF members (--raw) with
 *       archive_read_data into a 64 KiB buffer, CRC check on; prints one JSON
 *       line with wall seconds of the slowest process and bytes produced.
 */
#include <archive.h>
#include <archive_entry.h>
#include <zlib.h>

#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>

static double
now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + tv.tv_usec * 1e-6;
}

static void *
slurp(const char *path, size_t *len)
{
	struct stat st;
	int fd = open(path, O_RDONLY);
	char *buf;
	size_t got = 0;

	if (fd < 0 || fstat(fd, &st) != 0) {
		perror(path);
		exit(2);
	}
	buf = malloc(st.st_size ? st.st_size : 1);
	while (got < (size_t)st.st_size) {
		ssize_t r = read(fd, buf + got, st.st_size - got);
		if (r <= 0)
			break;
		got += r;
	}
	close(fd);
	*len = got;
	return buf;
}

static void
json_str(FILE *f, const char *s)
{
	fputc('"', f);
	for (; s && *s; s++) {
		unsigned char c = (unsigned char)*s;
		if (c == '"' || c == '\\')
			fprintf(f, "\\%c", c);
		else if (c < 0x20 || c >= 0x7f)
			fprintf(f, "\\u%04x", c);
		else
			fputc(c, f);
	}
	fputc('"', f);
}

/* --stream N: hand the bytes out in N-byte blocks through a read callback only
 * (no seek, no skip), so that the ZIP STREAMING reader is the one that runs */
static struct { const unsigned char *p; size_t left, blk; } g_src;
static int g_stream_blk;
static const char *g_file_path;     /* --file: archive_read_open_filename instead of open_memory */
static int g_meta;      /* --meta: owner, access/change times, link target, encryption flags as well */

static la_ssize_t
stream_read(struct archive *a, void *cd, const void **buff)
{
	size_t n = g_src.left < g_src.blk ? g_src.left : g_src.blk;
	(void)a; (void)cd;
	*buff = g_src.p;
	g_src.p += n;
	g_src.left -= n;
	return (la_ssize_t)n;
}

static struct archive *
open_reader(const void *buf, size_t len, int raw, const char *opt)
{
	struct archive *a = archive_read_new();

	if (raw) {
		archive_read_support_filter_gzip(a);
		archive_read_support_format_raw(a);
	} else {
		archive_read_support_format_zip(a);
	}
	if (opt && archive_read_set_options(a, opt) != ARCHIVE_OK) {
		fprintf(stderr, "set_options: %s\n", archive_error_string(a));
		exit(2);
	}
	if (g_stream_blk > 0) {
		g_src.p = buf;
		g_src.left = len;
		g_src.blk = (size_t)g_stream_blk;
	}
	if ((g_file_path != NULL ? archive_read_open_filename(a, g_file_path, 65536) :
	    g_stream_blk > 0 ? archive_read_open(a, NULL, NULL, stream_read, NULL) :
	    archive_read_open_memory(a, buf, len)) != ARCHIVE_OK) {
		printf("{\"open\":%d,\"err\":", -30);
		json_str(stdout, archive_error_string(a));
		printf("}\n");
		archive_read_free(a);
		return NULL;
	}
	return a;
}

static int
cmd_list(const void *buf, size_t len, int raw, const char *opt, const char *dump)
{
	struct archive *a = open_reader(buf, len, raw, opt);
	struct archive_entry *e;
	FILE *df = dump ? fopen(dump, "wb") : NULL;
	int idx = 0, hr;

	if (a == NULL)
		return 0;
	for (;;) {
		const void *blk;
		size_t bsz;
		int64_t off;
		uint64_t nbytes = 0;
		uLong crc = crc32(0L, NULL, 0);
		int rr, nblk = 0;
		size_t first_blocks[8];
		char herr[512] = "", rerr[512] = "";

		hr = archive_read_next_header(a, &e);
		if (hr == ARCHIVE_EOF)
			break;
		if (hr != ARCHIVE_OK && archive_error_string(a))
			snprintf(herr, sizeof(herr), "%s", archive_error_string(a));
		if (hr == ARCHIVE_FATAL) {
			printf("{\"i\":%d,\"hdr\":%d,\"herr\":", idx, hr);
			json_str(stdout, herr);
			printf("}\n");
			break;
		}
		for (;;) {
			rr = archive_read_data_block(a, &blk, &bsz, &off);
			if (rr != ARCHIVE_OK)
				break;
			if (nblk < 8)
				first_blocks[nblk] = bsz;
			nblk++;
			if (bsz > 0)
				crc = crc32(crc, blk, (uInt)bsz);
			nbytes += bsz;
			if (df)
				fwrite(blk, 1, bsz, df);
		}
		if (rr != ARCHIVE_EOF && archive_error_string(a))
			snprintf(rerr, sizeof(rerr), "%s", archive_error_string(a));
		printf("{\"i\":%d,\"name\":", idx);
		json_str(stdout, archive_entry_pathname(e));
		printf(",\"size\":%lld,\"size_set\":%d,\"mode\":%u,\"mtime\":%lld,\"hdr\":%d,\"herr\":",
		    (long long)archive_entry_size(e), archive_entry_size_is_set(e),
		    (unsigned)archive_entry_mode(e), (long long)archive_entry_mtime(e), hr);
		json_str(stdout, herr);
		if (g_meta) {
			printf(",\"uid\":%lld,\"gid\":%lld,\"atime\":%lld,\"ctime\":%lld,\"enc\":%d,\"menc\":%d,\"has_enc\":%d,\"link\":",
			    (long long)archive_entry_uid(e), (long long)archive_entry_gid(e),
			    (long long)archive_entry_atime(e), (long long)archive_entry_ctime(e),
			    archive_entry_is_data_encrypted(e), archive_entry_is_metadata_encrypted(e),
			    archive_read_has_encrypted_entries(a));
			json_str(stdout, archive_entry_symlink(e));
		}
		printf(",\"format\":");
		json_str(stdout, archive_format_name(a));
		printf(",\"rd\":%d,\"nbytes\":%llu,\"crc\":\"%08lx\",\"nblk\":%d,\"blocks\":[",
		    rr, (unsigned long long)nbytes, (unsigned long)crc, nblk);
		for (int k = 0; k < nblk && k < 8; k++)
			printf("%s%zu", k ? "," : "", first_blocks[k]);
		printf("],\"err\":");
		json_str(stdout, rerr);
		printf("}\n");
		idx++;
	}
	printf("{\"eof\":1,\"file_count\":%d,\"filter0\":", archive_file_count(a));
	json_str(stdout, archive_filter_name(a, 0));
	printf("}\n");
	archive_read_free(a);
	if (df)
		fclose(df);
	return 0;
}

/* count entries (ZIP) so that shards can be cut by index */
static int
count_entries(const void *buf, size_t len)
{
	struct archive *a = open_reader(buf, len, 0, NULL);
	struct archive_entry *e;
	int n = 0;

	if (a == NULL)
		return 0;
	while (archive_read_next_header(a, &e) == ARCHIVE_OK)
		n++;
	archive_read_free(a);
	return n;
}

/* BGZF member boundaries via BSIZE (to cut the file into per-process ranges;
 * the reference itself never reads BSIZE) */
static size_t
bgzf_boundaries(const unsigned char *p, size_t len, size_t **offs)
{
	size_t cap = 1024, n = 0, off = 0;
	size_t *o = malloc(cap * sizeof(*o));

	while (off + 18 <= len && p[off] == 0x1f && p[off + 1] == 0x8b &&
	    (p[off + 3] & 4) && p[off + 12] == 'B' && p[off + 13] == 'C') {
		size_t bsize = p[off + 16] | (p[off + 17] << 8);
		if (n + 2 > cap) {
			cap *= 2;
			o = realloc(o, cap * sizeof(*o));
		}
		o[n++] = off;
		off += bsize + 1;
	}
	o[n] = off;
	*offs = o;
	return n;
}

static uint64_t
read_shard_zip(const void *buf, size_t len, int lo, int hi, int *bad)
{
	struct archive *a = open_reader(buf, len, 0, NULL);
	struct archive_entry *e;
	static char out[65536];
	uint64_t total = 0;
	int i = 0;

	if (a == NULL)
		return 0;
	while (i < hi && archive_read_next_header(a, &e) == ARCHIVE_OK) {
		if (i >= lo) {
			la_ssize_t r;
			while ((r = archive_read_data(a, out, sizeof(out))) > 0)
				total += r;
			if (r < 0)
				(*bad)++;
		}
		i++;
	}
	archive_read_free(a);
	return total;
}

static uint64_t
read_shard_raw(const unsigned char *buf, size_t lo, size_t hi, int *bad)
{
	struct archive *a = open_reader(buf + lo, hi - lo, 1, NULL);
	struct archive_entry *e;
	static char out[65536];
	uint64_t total = 0;
	la_ssize_t r;

	if (a == NULL)
		return 0;
	if (archive_read_next_header(a, &e) == ARCHIVE_OK) {
		while ((r = archive_read_data(a, out, sizeof(out))) > 0)
			total += r;
		if (r < 0)
			(*bad)++;
	}
	archive_read_free(a);
	return total;
}

static int
cmd_bench(const void *buf, size_t len, int raw, int procs, int reps,
    int limit_entries)
{
	size_t *offs = NULL, nmem = 0;
	int nent = 0, p, rep;
	double best = 1e30;
	uint64_t bytes = 0;
	int bad_total = 0;

	if (raw)
		nmem = bgzf_boundaries(buf, len, &offs);
	else
		nent = count_entries(buf, len);
	if (limit_entries > 0) {
		if (raw && (size_t)limit_entries < nmem) nmem = limit_entries;
		if (!raw && limit_entries < nent) nent = limit_entries;
	}
	for (rep = 0; rep < reps; rep++) {
		int (*pipes)[2] = calloc(procs, sizeof(*pipes));
		double t0 = now(), t1;
		uint64_t total = 0;

		for (p = 0; p < procs; p++) {
			pid_t pid;
			if (pipe(pipes[p]) != 0)
				exit(2);
			pid = fork();
			if (pid == 0) {
				uint64_t v[2];
				int bad = 0;
				if (raw) {
					size_t lo = nmem * p / procs, hi = nmem * (p + 1) / procs;
					v[0] = hi > lo ? read_shard_raw(buf, offs[lo], offs[hi], &bad) : 0;
				} else {
					int lo = (int)((long long)nent * p / procs);
					int hi = (int)((long long)nent * (p + 1) / procs);
					v[0] = read_shard_zip(buf, len, lo, hi, &bad);
				}
				v[1] = bad;
				if (write(pipes[p][1], v, sizeof(v)) != sizeof(v))
					_exit(3);
				_exit(0);
			}
			close(pipes[p][1]);
		}
		for (p = 0; p < procs; p++) {
			uint64_t v[2] = { 0, 0 };
			if (read(pipes[p][0], v, sizeof(v)) != sizeof(v))
				bad_total++;
			total += v[0];
			bad_total += (int)v[1];
			close(pipes[p][0]);
		}
		while (wait(NULL) > 0)
			;
		t1 = now();
		if (t1 - t0 < best)
			best = t1 - t0;
		bytes = total;
		free(pipes);
	}
	printf("{\"bench\":1,\"procs\":%d,\"reps\":%d,\"seconds\":%.6f,\"out_bytes\":%llu,"
	    "\"gbps\":%.6f,\"bad\":%d,\"units\":%llu,\"zlib\":\"%s\"}\n",
	    procs, reps, best, (unsigned long long)bytes, bytes / best / 1e9, bad_total,
	    (unsigned long long)(raw ? nmem : (size_t)nent), zlibVersion());
	return 0;
}

int
main(int argc, char **argv)
{
	const char *opt = NULL, *dump = NULL;
	int raw = 0, procs = 1, reps = 1, limit = 0, i;
	size_t len;
	void *buf;

	if (argc < 3) {
		fprintf(stderr, "usage: %s list|bench <file> [--raw] [--opt o] [--dump f] [--procs P] [--reps R] [--limit N]\n", argv[0]);
		return 2;
	}
	for (i = 3; i < argc; i++) {
		if (!strcmp(argv[i], "--raw")) raw = 1;
		else if (!strcmp(argv[i], "--opt") && i + 1 < argc) opt = argv[++i];
		else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];
		else if (!strcmp(argv[i], "--procs") && i + 1 < argc) procs = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--limit") && i + 1 < argc) limit = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--stream") && i + 1 < argc) g_stream_blk = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--meta")) g_meta = 1;
		else if (!strcmp(argv[i], "--file")) g_file_path = argv[2];
	}
	buf = slurp(argv[2], &len);
	if (!strcmp(argv[1], "list"))
		return cmd_list(buf, len, raw, opt, dump);
	if (!strcmp(argv[1], "bench"))
		return cmd_bench(buf, len, raw, procs, reps, limit);
	return 2;
}
