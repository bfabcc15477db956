/*
 * b200inflate.h — thin C ABI between libarchive's C read plugins and the
 * sm_100a CUDA hot path (batched raw-deflate decode + CRC-32).
 *
 * This header is the drop-in boundary.  Every entry point names the
 * reference call (file:line under antekone/libarchive) it replaces.  Nothing
 * here exposes CUDA or torch types: plain pointers, sizes and integers.
 *
 *   reference call site                                    replaced by
 *   ----------------------------------------------------   --------------------------
 *   inflateInit2/inflateReset/inflate/inflateEnd           b2i_plan_* / b2i_decode_host
 *     archive_read_support_format_zip.c:2510-2533, 2643    (one descriptor per entry,
 *     archive_read_support_filter_gzip.c:357-363, 479      one device pass per batch)
 *   crc32() through zip->crc32func (real_crc32)            fused in the inflate pass;
 *     archive_read_support_format_zip.c:405-409, 3154-3157 b2i_crc32 for the scalar
 *     archive_crc32.h:43-84 (semantics)                    drop-in
 *   zip_read_data_none                                     method 0 descriptors
 *     archive_read_support_format_zip.c:1592-1706          (CRC in place, optional copy)
 *   slurp_central_directory / zip_read_local_file_header   b2i_zip_index_build
 *     archive_read_support_format_zip.c:3867-4092, 905-972 (host, flat arrays)
 *   peek_at_header / consume_header / consume_trailer      b2i_gzip_scan_bgzf
 *     archive_read_support_filter_gzip.c:128-239, 340-429  (host; BGZF BSIZE chain)
 *
 * Error model: infrastructure failures (bad arguments, CUDA errors, OOM) are
 * negative B2I_E_* return values and a message in b2i_last_error(); per-stream
 * data errors travel in b2i_stream_result.status using zlib's numeric codes so
 * the plugin can print the reference's "ZIP decompression failed (%d)".
 */
#ifndef B200INFLATE_H
#define B200INFLATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2I_ABI_VERSION 4

/* ---- return codes of the API functions -------------------------------- */
#define B2I_OK            0
#define B2I_E_INVAL      -1   /* bad argument (alignment, NULL, overlap)      */
#define B2I_E_CUDA       -2   /* a CUDA runtime call failed                    */
#define B2I_E_NOMEM      -3   /* host or device allocation failed              */
#define B2I_E_NODEVICE   -4   /* no sm_100 device / extension not usable       */
#define B2I_E_FORMAT     -5   /* container framing could not be parsed         */

/* ---- per-stream status (b2i_stream_result.status) ---------------------- */
#define B2I_S_OK             0
#define B2I_S_DATA_ERROR    -3   /* == Z_DATA_ERROR: malformed deflate data     */
#define B2I_S_BUF_ERROR     -5   /* == Z_BUF_ERROR: input ended before the
                                    final block did (zip.c:2654-2657 prints it) */
#define B2I_S_OUT_OVERFLOW  -100 /* stream wants more than out_cap bytes; the
                                    host re-submits it with a larger capacity  */
#define B2I_S_UNSUPPORTED   -101 /* method is neither 0 nor 8                   */

/* why a stream got B2I_S_DATA_ERROR (zlib's msg strings, by number) */
#define B2I_D_NONE               0
#define B2I_D_BAD_BLOCK_TYPE     1  /* "invalid block type"                      */
#define B2I_D_BAD_STORED_LEN     2  /* "invalid stored block lengths"            */
#define B2I_D_TOO_MANY_SYMS      3  /* "too many length or distance symbols"     */
#define B2I_D_BAD_CODELEN_SET    4  /* "invalid code lengths set"                */
#define B2I_D_BAD_BITLEN_REPEAT  5  /* "invalid bit length repeat"               */
#define B2I_D_NO_EOB             6  /* "invalid code -- missing end-of-block"    */
#define B2I_D_BAD_LITLEN_SET     7  /* "invalid literal/lengths set"             */
#define B2I_D_BAD_DIST_SET       8  /* "invalid distances set"                   */
#define B2I_D_BAD_LITLEN_CODE    9  /* "invalid literal/length code"             */
#define B2I_D_BAD_DIST_CODE     10  /* "invalid distance code"                   */
#define B2I_D_DIST_TOO_FAR      11  /* "invalid distance too far back"           */

/* ---- descriptor flags ---------------------------------------------------- */
#define B2I_F_NO_COPY      0x01  /* method 0 only: CRC in place, write nothing
                                    (zero-copy like zip_read_data_none)          */
#define B2I_F_NO_CRC       0x02  /* skip CRC-32 (zip "ignorecrc32" option)       */

/* ---- result flags (the reference's end-of-entry checks, zip.c:3164-3194) -- */
#define B2I_R_CRC_MISMATCH   0x01  /* crc != expect_crc                          */
#define B2I_R_IN_MISMATCH    0x02  /* in_bytes != in_len ("compressed data is
                                      wrong size")                               */
#define B2I_R_OUT_MISMATCH   0x04  /* low 32 bits of out_bytes != expect_out     */

#define B2I_METHOD_STORED   0
#define B2I_METHOD_DEFLATE  8

/* One independent stream: a ZIP entry body or one gzip/BGZF member's raw
 * deflate payload.  Offsets are relative to the input / output buffers handed
 * to the launch call.  out_off must be a multiple of 16. */
typedef struct b2i_stream_desc {
	uint64_t in_off;      /* first byte of the payload                        */
	uint64_t in_len;      /* bytes that belong to it (compressed_size)        */
	uint64_t out_off;     /* where its decoded bytes go                       */
	uint64_t out_cap;     /* capacity reserved there (>= expected size)       */
	uint64_t expect_out;  /* uncompressed_size from the directory / ISIZE     */
	uint32_t expect_crc;  /* CRC-32 from the directory / gzip trailer         */
	uint8_t  method;      /* B2I_METHOD_*                                     */
	uint8_t  flags;       /* B2I_F_*                                          */
	uint16_t reserved;
} b2i_stream_desc;

typedef struct b2i_stream_result {
	int32_t  status;      /* B2I_S_*                                          */
	uint32_t crc;         /* CRC-32 of the bytes produced (0 if B2I_F_NO_CRC) */
	uint64_t out_bytes;   /* zlib total_out                                   */
	uint64_t in_bytes;    /* zlib total_in (whole bytes consumed)             */
	uint32_t detail;      /* B2I_D_*                                          */
	uint32_t flags;       /* B2I_R_*                                          */
} b2i_stream_result;

typedef struct b2i_ctx  b2i_ctx;
typedef struct b2i_plan b2i_plan;

/* ---- context --------------------------------------------------------------
 * One context per (host thread, GPU).  `cuda_stream` may be NULL (the context
 * creates its own non-blocking stream) or an existing cudaStream_t cast to
 * void* (all work is then enqueued there, so the caller's events time it). */
int  b2i_ctx_create(int device, void *cuda_stream, b2i_ctx **out);
void b2i_ctx_destroy(b2i_ctx *);
const char *b2i_last_error(const b2i_ctx *);      /* "" when none            */
int  b2i_abi_version(void);
int  b2i_device_count(void);
int  b2i_ctx_sync(b2i_ctx *);
int  b2i_ctx_device(const b2i_ctx *);            /* the device it was created on */
/* how many kernels this context has launched so far (bench's gpu_launches) */
uint64_t b2i_ctx_launch_count(const b2i_ctx *);

/* pinned host memory / device memory owned by the library (plain pointers) */
void *b2i_host_alloc(size_t bytes);
void  b2i_host_free(void *);
void *b2i_device_alloc(b2i_ctx *, size_t bytes);
void  b2i_device_free(b2i_ctx *, void *);
int   b2i_memcpy_h2d(b2i_ctx *, void *dst_dev, const void *src_host, size_t bytes); /* async */
int   b2i_memcpy_d2h(b2i_ctx *, void *dst_host, const void *src_dev, size_t bytes); /* async */

/* ---- batch plan: "the host batches entry offsets up front" ---------------
 * b2i_plan_create validates and uploads the descriptors and the derived
 * schedule; b2i_plan_launch enqueues ONE device pass over device-resident
 * input/output (inflate + fused CRC for method 8, chunked CRC (+copy) for
 * method 0); b2i_plan_results synchronises and copies the per-stream results
 * back.  d_in must be 16-byte aligned with the allocation padded to a multiple
 * of 16 bytes; d_out 16-byte aligned. */
int  b2i_plan_create(b2i_ctx *, const b2i_stream_desc *descs, size_t n, b2i_plan **out);
int  b2i_plan_launch(b2i_plan *, const void *d_in, size_t in_bytes,
                     void *d_out, size_t out_bytes);
int  b2i_plan_results(b2i_plan *, b2i_stream_result *res /* n entries */);
void b2i_plan_destroy(b2i_plan *);

/* ---- host-buffer convenience (the end-to-end path the plugins use) -------
 * Copies host_in[0..in_bytes) to the device, runs the plan, copies
 * out[0..out_bytes) back into host_out (may be NULL: results only, e.g. CRC
 * verification of stored entries) and fills res[].  Synchronous. */
int  b2i_decode_host(b2i_ctx *, const void *host_in, size_t in_bytes,
                     const b2i_stream_desc *descs, size_t n,
                     void *host_out, size_t out_bytes, b2i_stream_result *res);

/* The same, split in two so that a reader working through several archives (or
 * one archive in pieces) keeps the device and both directions of the host link
 * busy: b2i_submit queues copy-in, kernels and copy-out and returns; b2i_wait
 * blocks until the job's last byte has landed in host_out, fills res[n] and
 * releases the job.  Up to five jobs of one context may be in flight, so the
 * copy-out of one overlaps the copy-in and decode of the next ones; a sixth submit
 * fails with B2I_E_INVAL.  The descriptors are copied by b2i_submit; host_in and
 * host_out must stay valid (and should be pinned, b2i_host_alloc) until b2i_wait
 * returns.  Replaces the reference's one-entry-at-a-time inflate loop
 * (archive_read_support_format_zip.c:2535-2690) the same way b2i_decode_host does. */
typedef struct b2i_job b2i_job;
int  b2i_submit(b2i_ctx *, const void *host_in, size_t in_bytes,
                const b2i_stream_desc *descs, size_t n,
                void *host_out, size_t out_bytes, b2i_job **job);
int  b2i_wait(b2i_job *job, b2i_stream_result *res /* n entries */);
/* blocks until the job's input has been copied to the device: host_in may then be
 * released (a read filter consumes its upstream bytes) while the job decodes on */
int  b2i_job_wait_input(b2i_job *job);
/* Where the library keeps its own pinned copy of the job's input (pageable inputs of a few
 * MiB and more are staged), or NULL if it reads host_in directly.  base + in_off holds the
 * bytes of every stream of the job.  Ask before b2i_wait; the bytes stay until the next
 * b2i_submit on the context.  (The gzip filter releases its upstream bytes after
 * b2i_job_wait_input and comes back here for the one member whose trailer lied.) */
const void *b2i_job_staged_input(const b2i_job *job);

/* The same over several GPUs of one box: the batch is partitioned on the
 * host (b2i_partition_contiguous, or b2i_partition_lpt when a few streams dominate),
 * every context decodes its share on its own thread with its own streams and pinned
 * copies, nothing is exchanged between the devices.  One context per device is the
 * intended use (several on one device also work); host_in / host_out should be pinned. */
int  b2i_decode_host_multi(b2i_ctx *const *ctxs, int nctx, const void *host_in, size_t in_bytes,
                           const b2i_stream_desc *descs, size_t n,
                           void *host_out, size_t out_bytes, b2i_stream_result *res);

/* ---- streaming pipeline: bounded memory, file-backed sources, all GPUs ----------
 * Replaces the reference's serial "one block of the source, one inflate() call"
 * loop (archive_read_open_filename.c:389-461 feeding
 * archive_read_support_format_zip.c:2535-2690 / archive_read_support_filter_gzip.c:
 * 431-511) for a whole archive: the descriptors (in_off = offset in the SOURCE,
 * out_off ignored, archive order) are cut into windows of bounded output; windows are
 * staged through a ring of pinned buffers, decoded round-robin by the given contexts
 * (one worker thread and four jobs in flight per GPU) and handed out in order.
 * Resident memory is windows_per_device x nctx x (window input + output), whatever
 * the archive size.  Source: `mem` (the archive image in host memory: staged by copy
 * threads, or used in place when it is pinned) or `fill` (called ONLY on the thread
 * that calls b2i_pipe_get, so it may use libarchive's read filters: reads
 * [offset, offset + len) of the source into dst, returns B2I_OK or an error). */
typedef int (*b2i_fill_fn)(void *user, uint64_t offset, uint64_t len, void *dst);
typedef struct b2i_pipe b2i_pipe;
typedef struct b2i_pipe_opts {
	size_t window_out_bytes;        /* 0: a quarter of the batch's output per device, 16..256 MiB
	                                   (callback sources: at most 64 MiB) */
	size_t first_window_out_bytes;  /* 0: a quarter of that (first bytes arrive sooner) */
	int    windows_per_device;      /* ring depth, 0: 5 (four of them in flight) */
	int    copy_threads;            /* reserved (the staging threads are one process-wide pool of 6,
	                                   B2I_COPY_THREADS) */
} b2i_pipe_opts;
int  b2i_pipe_open(b2i_ctx *const *ctxs, int nctx, const void *mem, uint64_t mem_size,
                   b2i_fill_fn fill, void *user, const b2i_stream_desc *descs, size_t n,
                   const b2i_pipe_opts *opts, b2i_pipe **out);
/* Blocks until stream idx is decoded.  *out_data: its bytes (NULL for B2I_F_NO_COPY
 * stored streams), *in_data: its compressed bytes as staged; both stay valid until
 * b2i_pipe_release moves past idx or a stream of a LATER window is asked for (which
 * gives up all earlier windows and drops those not yet started: read_data_skip). */
int  b2i_pipe_get(b2i_pipe *, size_t idx, const void **out_data, const void **in_data,
                  b2i_stream_result *res);
void b2i_pipe_release(b2i_pipe *, size_t idx);    /* streams < idx are done with */
size_t b2i_pipe_window_count(const b2i_pipe *);
const char *b2i_pipe_error(const b2i_pipe *);
void b2i_pipe_close(b2i_pipe *);

/* ---- scalar drop-ins --------------------------------------------------------
 * b2i_crc32: same contract as zlib crc32()/archive_crc32.h:43-84:
 * crc32(x, NULL, 0) == 0, chaining by passing the previous value.  Runs the
 * chunked CRC kernel on `ctx` (host buffer is copied to the device). */
int      b2i_crc32(b2i_ctx *, uint32_t crc, const void *host_buf, size_t len, uint32_t *out);
/* CRC of device-resident bytes (no copy); device pointer, any alignment */
int      b2i_crc32_device(b2i_ctx *, uint32_t crc, const void *d_buf, size_t len, uint32_t *out);
/* crc32_combine: CRC of A||B from crc(A), crc(B), len(B) (host arithmetic) */
uint32_t b2i_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* ---- host-side framing (no GPU involved) -----------------------------------
 * ZIP: the seekable reader's directory walk flattened into arrays. */
typedef struct b2i_zip_entry {
	uint64_t local_header_offset;  /* after `correction` (zip.c:3985-3986)  */
	uint64_t data_offset;          /* header + 30 + name + extra            */
	uint64_t compressed_size;      /* after CD/local reconciliation          */
	uint64_t uncompressed_size;    /*   (zip.c:1106-1150)                    */
	uint32_t crc32;
	uint32_t name_offset;          /* into b2i_zip_index.names               */
	uint16_t name_len;
	uint16_t zip_flags;            /* local header general purpose flags     */
	uint16_t method;               /* local header compression method        */
	uint8_t  version;              /* local header "version needed" low byte */
	uint8_t  system;
	uint32_t mode;                 /* as zip.c:3991-4006 derives it          */
	uint32_t warn;                 /* B2I_ZW_* inconsistencies (WARN)        */
	int64_t  mtime;                /* DOS time, then extras 0x5455 / 0x5855 (zip.c:597-650) */
	int64_t  atime, ctime;         /* 0 when the archive carries none        */
	uint32_t uid, gid;             /* extras 0x5855 / 0x7855 / 0x7875         */
	uint64_t local_extra_offset;   /* the local header's extra field (for 0x7075) */
	uint16_t local_extra_len;
	uint16_t reserved[3];
} b2i_zip_entry;

#define B2I_ZW_CRC_INCONSISTENT   0x1
#define B2I_ZW_CSIZE_INCONSISTENT 0x2
#define B2I_ZW_USIZE_INCONSISTENT 0x4
#define B2I_ZW_BAD_LOCAL_HEADER   0x8  /* no PK\3\4 at the offset: "Damaged Zip archive" */
#define B2I_ZW_TRUNCATED          0x10 /* header or body extends past the file  */

typedef struct b2i_zip_index {
	size_t         n;
	b2i_zip_entry *entries;   /* ascending local_header_offset (zip.c:3778-3789) */
	char          *names;     /* concatenated raw names                           */
	size_t         names_len;
	int64_t        correction;
	int            has_encrypted_entries;
} b2i_zip_index;

int  b2i_zip_index_build(const void *archive, size_t size, b2i_zip_index *out, char errbuf[128]);
/* The same walk over a source that is not one memory image (a seekable file behind
 * libarchive's read filters, archive_read_open_filename.c:389-461): `fetch` returns a
 * pointer to [off, off + len) that stays valid until its next call, or NULL.  Reads the
 * tail, the directory and the local headers (in ascending offset), nothing else. */
typedef const uint8_t *(*b2i_fetch_fn)(void *user, uint64_t off, size_t len);
int  b2i_zip_index_build_cb(b2i_fetch_fn fetch, void *user, uint64_t size, b2i_zip_index *out,
                            char errbuf[128]);
/* seekable bid (zip.c:3720-3773): is there an acceptable end record in the last
 * tail_len (<= 16 KiB) bytes of a file of file_size bytes?  1 / 0 */
int  b2i_zip_probe_tail(const void *tail, size_t tail_len, uint64_t file_size);
void b2i_zip_index_free(b2i_zip_index *);

/* gzip / BGZF: member chain.  With a BGZF 'BC' extra subfield the whole chain
 * is found without decoding; a plain member has deflate_len == 0 (unknown). */
typedef struct b2i_gzip_member {
	uint64_t header_offset;
	uint32_t header_len;      /* peek_at_header's return value               */
	uint64_t deflate_offset;
	uint64_t deflate_len;     /* 0 = unknown (no BSIZE)                      */
	uint32_t crc32;           /* from the trailer when deflate_len known     */
	uint32_t isize;
	uint32_t mtime;
	uint32_t name_offset;     /* offset of FNAME inside the file, 0 if none  */
} b2i_gzip_member;

/* parse one member header at `off`; returns header length, 0 if not a gzip header */
size_t b2i_gzip_peek_header(const void *buf, size_t size, size_t off, b2i_gzip_member *m);
/* walk the BSIZE chain from `off`; stops at the first non-BGZF header, garbage
 * or end of file; *end_off = offset where the walk stopped */
int  b2i_gzip_scan_bgzf(const void *buf, size_t size, size_t off,
                        b2i_gzip_member **members, size_t *n, size_t *end_off);
void b2i_free(void *);

/* ---- multi-GPU partition (host arithmetic, no GPU involved) ---------
 * Streams are independent (inflateReset per entry, archive_read_support_format_zip.c:
 * 2517-2518; inflateInit2 per member, archive_read_support_filter_gzip.c:363), so one
 * archive is split across the GPUs of a box on the host and nothing is exchanged.
 * contiguous: cuts[parts + 1], part p = descriptors [cuts[p], cuts[p+1]) in input
 * order, about equal in_len + out bytes: one compressed range in, one output range
 * out per GPU.  lpt: owner[i] = part of stream i, largest streams placed first on the
 * least loaded part (batches dominated by a few huge entries); load[parts] optional. */
int  b2i_partition_contiguous(const b2i_stream_desc *descs, size_t n, int parts, size_t *cuts);
int  b2i_partition_lpt(const b2i_stream_desc *descs, size_t n, int parts, uint32_t *owner,
                       uint64_t *load);

#ifdef __cplusplus
}
#endif
#endif /* B200INFLATE_H */
