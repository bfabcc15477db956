/*
 * b2i_api.cpp — the C ABI of include/b200inflate.h over the sm_100a kernels.
 *
 * Host side of "the host batches entry offsets up front, launches one device
 * pass and serves archive_read_data from the decoded buffers": a plan is the
 * batch (descriptors + schedule, uploaded once), a launch is the device pass.
 * There is no CPU decode path in this library: without a usable sm_100 device
 * every entry point fails with B2I_E_NODEVICE.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/b200inflate.h"
#include "b2i_kernels.h"

static_assert(sizeof(b2i_stream_desc) == sizeof(B2iDesc), "descriptor layout");
static_assert(sizeof(b2i_stream_result) == sizeof(B2iResult), "result layout");
static_assert(offsetof(b2i_stream_desc, expect_crc) == offsetof(B2iDesc, expect_crc), "descriptor layout");
static_assert(offsetof(b2i_stream_result, detail) == offsetof(B2iResult, detail), "result layout");

/* per-stream limits of this build: 32-bit positions inside one stream */
#define B2I_MAX_STREAM_BYTES 0xFFFF0000ull
#define B2I_MAX_JOBS_ 5
#define B2I_PIPE_STREAMS (B2I_MAX_JOBS_ * B2I_PIPE_SLICES)   /* compute streams: every slice of every job in flight its own */
#define B2I_PIPE_SLICES  12
#define B2I_TEAM_STREAMS (B2I_MAX_JOBS_ * B2I_PIPE_SLICES)

struct b2i_plan;
#define B2I_MAX_JOBS B2I_MAX_JOBS_
/* one host-buffer decode in flight: its own device staging, plan arena and events,
 * so that the copy-out of one job overlaps the copy-in and kernels of the next */
struct b2i_job {
	struct b2i_ctx *ctx;
	bool busy, ev_ready;
	uint8_t *d_in;  size_t d_in_cap;
	uint8_t *d_out; size_t d_out_cap;
	uint8_t *h_stage; size_t h_stage_cap;   /* pinned staging for pageable input */
	bool staged;                            /* this job reads its input from h_stage */
	uint8_t *arena_d; uint8_t *arena_h; size_t arena_cap;
	cudaEvent_t ev_in[B2I_PIPE_SLICES], ev_k[B2I_PIPE_SLICES], ev_done;
	b2i_plan *plans[B2I_PIPE_SLICES];
	size_t cut[B2I_PIPE_SLICES + 1];
	size_t K, n;
};

struct b2i_ctx {
	int device;
	int num_sms;
	cudaStream_t stream;
	bool own_stream;
	uint32_t *d_crc_tab;   /* 1024 */
	uint32_t *d_xp8;       /* 40 */
	uint32_t *d_ztab;      /* 1024: advance-by-512-bytes tables */
	uint32_t *d_lane_mul;  /* 32 */
	unsigned int *d_slot_busy;   /* one flag per token region */
	uint32_t *d_scratch;   /* token regions of the lane-parallel decoder, one per resident warp */
	uint64_t launches;
	/* grow-only staging for b2i_crc32 (host buffers) */
	uint8_t *d_in;  size_t d_in_cap;
	/* pipelined host path: copy-in / compute / copy-out streams, and a reusable
	 * (grow-only) arena for the slice plans so that no call allocates */
	cudaStream_t s_in, s_out, s_cmp[B2I_PIPE_STREAMS];
	cudaStream_t s_team[B2I_TEAM_STREAMS];   /* large streams (one CTA each) run beside the single-warp kernel;
	                                           every slice of every job in flight gets its own, so that a
	                                           slice's long CTAs never queue behind another slice's */
	cudaEvent_t ev_free;
	bool pipe_ready;
	size_t high_in, high_out, high_arena, high_crc, high_stage;   /* largest staging requests so far (ensure_dev) */
	b2i_job jobs[B2I_MAX_JOBS];     /* host-buffer decodes in flight (b2i_submit / b2i_wait) */
	char err[256];
};

struct b2i_plan {
	b2i_ctx *ctx;
	size_t n;
	uint32_t n_deflate, n_big, n_stored, n_work, n_unsup;   /* n_deflate: single-warp streams; n_big: team streams */
	cudaEvent_t ev_fork, ev_join;                            /* team kernel runs beside the rest */
	bool need_aligned_in;
	uint64_t max_in_end, max_out_end;
	/* device */
	uint8_t *d_block;
	B2iDesc *d_descs;
	B2iResult *d_results;
	uint32_t *d_order;
	B2iCrcWork *d_work;
	B2iCrcEntry *d_ents;
	uint32_t *d_partial;
	uint32_t *d_unsup;
	unsigned int *d_counter;
	/* pinned host mirror used for upload and result download */
	uint8_t *h_block;
	size_t block_bytes, results_off;
	cudaStream_t stream;     /* where this plan's upload, kernels and result copy run */
	cudaStream_t team_stream;
	bool owns_memory;        /* false: d_block / h_block live in the context's arena */
	uint8_t *out_mirror;     /* host-mapped twin of the output (inflated bytes are stored to both) */
	size_t batch_streams;    /* deflate streams of the whole batch this plan is a slice of (0: just this plan) */
	bool defer_join;         /* the caller orders its consumers behind ev_join itself (pipelined slices) */
};

static int fail(b2i_ctx *c, int code, const char *fmt, ...)
{
	if (c) {
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(c->err, sizeof(c->err), fmt, ap);
		va_end(ap);
	}
	return code;
}

#define CU(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail((c), B2I_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

extern "C" int b2i_abi_version(void) { return B2I_ABI_VERSION; }

extern "C" int b2i_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess)
		return 0;
	return n;
}

/* B2I_CALL_STATS=1: wall time and count of the entry points a plugin calls per archive,
 * printed at exit (diagnostic for the small-archive path) */
struct CallStat { const char *name; std::atomic<uint64_t> ns{0}, calls{0}; };
static CallStat g_cs_ctx{"b2i_ctx_create"}, g_cs_decode{"b2i_decode_host"}, g_cs_halloc{"b2i_host_alloc"},
    g_cs_hfree{"b2i_host_free"}, g_cs_submit{"b2i_submit"}, g_cs_wait{"b2i_wait"};
static bool call_stats_on()
{
	static int on = -1;
	if (on < 0) {
		on = getenv("B2I_CALL_STATS") != NULL;
		if (on)
			atexit([] {
				for (CallStat *c : {&g_cs_ctx, &g_cs_decode, &g_cs_submit, &g_cs_wait, &g_cs_halloc, &g_cs_hfree})
					if (c->calls)
						fprintf(stderr, "B2I_CALL %-16s calls %8llu  total %9.3f ms  mean %8.1f us\n", c->name,
						    (unsigned long long)c->calls.load(), c->ns.load() / 1e6,
						    c->ns.load() / 1e3 / (double)c->calls.load());
			});
	}
	return on != 0;
}
struct CallTimer {
	CallStat *s; std::chrono::steady_clock::time_point t0;
	explicit CallTimer(CallStat &st) : s(call_stats_on() ? &st : nullptr) { if (s) t0 = std::chrono::steady_clock::now(); }
	~CallTimer() { if (s) { s->ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); s->calls++; } }
};

extern "C" int b2i_ctx_create(int device, void *cuda_stream, b2i_ctx **out)
{
	if (out == NULL)
		return B2I_E_INVAL;
	CallTimer ct_(g_cs_ctx);
	*out = NULL;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
		return B2I_E_NODEVICE;
	int major = 0, sms = 0;
	if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
	    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
		return B2I_E_NODEVICE;
	if (major != 10)
		return B2I_E_NODEVICE;      /* kernels are sm_100a only; no fallback */
	b2i_ctx *c = new (std::nothrow) b2i_ctx();
	if (c == NULL)
		return B2I_E_NOMEM;
	memset(c, 0, sizeof(*c));
	c->device = device;
	c->num_sms = sms;
	if (cudaSetDevice(device) != cudaSuccess) {
		delete c;
		return B2I_E_CUDA;
	}
	if (cuda_stream) {
		c->stream = (cudaStream_t)cuda_stream;
	} else {
		if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
			delete c;
			return B2I_E_CUDA;
		}
		c->own_stream = true;
	}
	if (b2i_kernels_configure() != cudaSuccess ||
	    cudaMalloc(&c->d_crc_tab, 1024 * 4) != cudaSuccess ||
	    cudaMalloc(&c->d_xp8, 40 * 4) != cudaSuccess ||
	    cudaMalloc(&c->d_ztab, 1024 * 4) != cudaSuccess ||
	    cudaMalloc(&c->d_lane_mul, 32 * 4) != cudaSuccess ||
	    cudaMalloc(&c->d_scratch, b2i_inflate_scratch_bytes(sms)) != cudaSuccess ||
	    cudaMalloc(&c->d_slot_busy, b2i_inflate_scratch_slots(sms) * 4) != cudaSuccess ||
	    cudaMemset(c->d_slot_busy, 0, b2i_inflate_scratch_slots(sms) * 4) != cudaSuccess ||
	    b2i_launch_tables(c->d_crc_tab, c->d_xp8, c->d_ztab, c->d_lane_mul, c->stream) != cudaSuccess ||
	    cudaStreamSynchronize(c->stream) != cudaSuccess) {
		b2i_ctx_destroy(c);
		return B2I_E_CUDA;
	}
	c->launches = 1;
	*out = c;
	return B2I_OK;
}

extern "C" void b2i_ctx_destroy(b2i_ctx *c)
{
	if (c == NULL)
		return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	cudaFree(c->d_crc_tab);
	cudaFree(c->d_xp8);
	cudaFree(c->d_scratch);
	cudaFree(c->d_slot_busy);
	cudaFree(c->d_ztab);
	cudaFree(c->d_lane_mul);
	cudaFree(c->d_in);
	for (int i = 0; i < B2I_MAX_JOBS; i++) {
		b2i_job *J = &c->jobs[i];
		for (size_t k = 0; k < B2I_PIPE_SLICES; k++)
			if (J->plans[k])
				b2i_plan_destroy(J->plans[k]);
		cudaFree(J->d_in);
		cudaFree(J->d_out);
		cudaFreeHost(J->h_stage);
		cudaFree(J->arena_d);
		cudaFreeHost(J->arena_h);
		if (J->ev_ready) {
			for (int k = 0; k < B2I_PIPE_SLICES; k++) { cudaEventDestroy(J->ev_in[k]); cudaEventDestroy(J->ev_k[k]); }
			cudaEventDestroy(J->ev_done);
		}
	}
	for (int i = 0; i < B2I_TEAM_STREAMS; i++)
		if (c->s_team[i]) cudaStreamDestroy(c->s_team[i]);
	if (c->pipe_ready) {
		cudaStreamDestroy(c->s_in);
		cudaStreamDestroy(c->s_out);
		for (int i = 0; i < B2I_PIPE_STREAMS; i++) cudaStreamDestroy(c->s_cmp[i]);
		cudaEventDestroy(c->ev_free);
	}
	if (c->own_stream)
		cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" const char *b2i_last_error(const b2i_ctx *c) { return c ? c->err : "no context"; }
extern "C" uint64_t b2i_ctx_launch_count(const b2i_ctx *c) { return c ? c->launches : 0; }
extern "C" int b2i_ctx_device(const b2i_ctx *c) { return c ? c->device : -1; }

#ifdef B2I_PHASE_CLOCKS
void b2i_phase_dump(cudaStream_t st);
#endif

extern "C" int b2i_ctx_sync(b2i_ctx *c)
{
	if (c == NULL)
		return B2I_E_INVAL;
#ifdef B2I_PHASE_CLOCKS
	b2i_phase_dump(c->stream);
#endif
	CU(c, cudaStreamSynchronize(c->stream));
	return B2I_OK;
}

extern "C" void *b2i_host_alloc(size_t bytes)
{
	void *p = NULL;
	CallTimer ct_(g_cs_halloc);
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess)
		return NULL;
	return p;
}
extern "C" void b2i_host_free(void *p)
{
	if (p) {
		CallTimer ct_(g_cs_hfree);
		cudaFreeHost(p);
	}
}

extern "C" void *b2i_device_alloc(b2i_ctx *c, size_t bytes)
{
	void *p = NULL;
	if (c == NULL || cudaSetDevice(c->device) != cudaSuccess)
		return NULL;
	/* +16: the input ring reads whole 16-byte units */
	if (cudaMalloc(&p, ((bytes + 15) & ~(size_t)15) + 16) != cudaSuccess)
		return NULL;
	return p;
}
extern "C" void b2i_device_free(b2i_ctx *c, void *p) { if (c && p) { cudaSetDevice(c->device); cudaFree(p); } }

extern "C" int b2i_memcpy_h2d(b2i_ctx *c, void *dst, const void *src, size_t bytes)
{
	if (c == NULL)
		return B2I_E_INVAL;
	CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
	return B2I_OK;
}
extern "C" int b2i_memcpy_d2h(b2i_ctx *c, void *dst, const void *src, size_t bytes)
{
	if (c == NULL)
		return B2I_E_INVAL;
	CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
	return B2I_OK;
}

/* ---- plan ------------------------------------------------------------------ */

static size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }
void b2i_parallel_copy(void *dst, const void *src, size_t len);      /* b2i_pipe.cpp */

static size_t plan_block_bound(size_t n, size_t stored_bytes)
{
	/* generous upper bound of the block b2i_plan_build lays out for n streams */
	size_t works = stored_bytes / 512 + 3 * n + 8;
	return n * (sizeof(B2iDesc) + sizeof(B2iResult) + 8 + sizeof(B2iCrcEntry)) +
	    works * (sizeof(B2iCrcWork) + 4) + 16 * 256;
}

/* mem_d / mem_h: caller-provided (arena) memory of mem_cap bytes, or NULL to allocate */
static int b2i_plan_build(b2i_ctx *c, const b2i_stream_desc *descs, size_t n, cudaStream_t stream,
    cudaStream_t upload_stream, uint8_t *mem_d, uint8_t *mem_h, size_t mem_cap, b2i_plan **out, int team_lane = 0)
{
	if (c == NULL || out == NULL || (n && descs == NULL) || n > 0x7fffffffu)
		return fail(c, B2I_E_INVAL, "b2i_plan_create: bad arguments");
	*out = NULL;
	CU(c, cudaSetDevice(c->device));

	std::vector<uint32_t> deflate, unsup;
	std::vector<B2iCrcWork> work;
	std::vector<B2iCrcEntry> ents;
	uint64_t max_in = 0, max_out = 0;
	for (size_t i = 0; i < n; i++) {
		const b2i_stream_desc &d = descs[i];
		if (d.out_off & 15)
			return fail(c, B2I_E_INVAL, "stream %zu: out_off must be a multiple of 16", i);
		if (d.in_off + d.in_len < d.in_off || d.out_off + d.out_cap < d.out_off)
			return fail(c, B2I_E_INVAL, "stream %zu: offset overflow", i);
		if (d.method == B2I_METHOD_DEFLATE) {
			if (d.in_len > B2I_MAX_STREAM_BYTES || d.out_cap > B2I_MAX_STREAM_BYTES)
				return fail(c, B2I_E_INVAL, "stream %zu: larger than this build's 4 GiB per-stream limit", i);
			deflate.push_back((uint32_t)i);
			max_in = std::max<uint64_t>(max_in, d.in_off + d.in_len);
			max_out = std::max<uint64_t>(max_out, d.out_off + d.out_cap);
		} else if (d.method == B2I_METHOD_STORED) {
			B2iCrcEntry e;
			e.entry = (uint32_t)i;
			e.first_work = (uint32_t)work.size();
			e.pad = 0;
			/* pieces: a head up to 16-byte alignment of the input offset, streaming
			 * pieces of up to 32 KiB in multiples of 512 bytes, and a tail */
			auto push = [&](uint64_t rel, uint64_t len) {
				B2iCrcWork w;
				w.rel = rel;
				w.len = (uint32_t)len;
				w.entry = (uint32_t)i;
				work.push_back(w);
			};
			uint64_t rel = 0, left = d.in_len;
			uint64_t head = std::min<uint64_t>((0 - d.in_off) & 15, left);
			if (head) { push(rel, head); rel += head; left -= head; }
			while (left >= 512) {
				uint64_t len = std::min<uint64_t>(B2I_CRC_CHUNK, left & ~(uint64_t)511);
				push(rel, len); rel += len; left -= len;
			}
			if (left) push(rel, left);
			e.nwork = (uint32_t)work.size() - e.first_work;
			ents.push_back(e);
			max_in = std::max<uint64_t>(max_in, d.in_off + d.in_len);
			/* an entry larger than its reserved output is never copied (S_OUT_OVERFLOW) */
			if (!(d.flags & B2I_F_NO_COPY) && d.in_len <= d.out_cap)
				max_out = std::max<uint64_t>(max_out, d.out_off + d.in_len);
		} else {
			unsup.push_back((uint32_t)i);
		}
	}
	/* largest streams first: the slowest warp starts earliest */
	std::stable_sort(deflate.begin(), deflate.end(), [&](uint32_t a, uint32_t b) {
		return descs[a].in_len + descs[a].out_cap > descs[b].in_len + descs[b].out_cap;
	});
	/* Streams that move at least B2I_TEAM_MIN_BYTES (compressed + uncompressed, default
	 * 1 MiB; 0 = never) get a whole CTA each (inflate_team.cuh): a lone warp decodes one
	 * stream at a few tens of MB/s, a CTA with its window in shared memory an order of
	 * magnitude faster, which is what bounds batches with multi-megabyte entries. */
	uint64_t team_min = 1u << 20;
	if (const char *ev = getenv("B2I_TEAM_MIN_BYTES"))
		team_min = strtoull(ev, NULL, 10);
	std::vector<uint32_t> big;
	if (team_min != 0) {
		/* a batch with fewer streams than two per SM cannot fill the GPU with single warps
		 * anyway: its medium streams (>= 48 KiB) take a CTA each too, which cuts their
		 * latency from ~1.5 ms to ~0.3 ms (the first window of a pipelined archive, small
		 * archives) */
		if (deflate.size() <= 2u * (size_t)c->num_sms && getenv("B2I_TEAM_MIN_BYTES") == NULL)
			team_min = 48u << 10;
		std::vector<uint32_t> small;
		for (uint32_t i : deflate)
			(descs[i].in_len + descs[i].out_cap >= team_min ? big : small).push_back(i);
		deflate.swap(small);
	}

	b2i_plan *p = new (std::nothrow) b2i_plan();
	if (p == NULL)
		return fail(c, B2I_E_NOMEM, "out of memory");
	memset(p, 0, sizeof(*p));
	p->ctx = c;
	p->n = n;
	p->n_deflate = (uint32_t)deflate.size();
	p->n_big = (uint32_t)big.size();
	p->n_stored = (uint32_t)ents.size();
	p->n_work = (uint32_t)work.size();
	p->n_unsup = (uint32_t)unsup.size();
	p->need_aligned_in = !deflate.empty() || !big.empty();
	p->max_in_end = max_in;
	p->max_out_end = max_out;
	p->stream = stream;
	p->owns_memory = (mem_d == NULL);

	size_t off = 0;
	const size_t o_descs = off;   off = align_up(off + n * sizeof(B2iDesc), 256);
	const size_t o_order = off;   off = align_up(off + (deflate.size() + big.size()) * 4, 256);
	const size_t o_work = off;    off = align_up(off + work.size() * sizeof(B2iCrcWork), 256);
	const size_t o_ents = off;    off = align_up(off + ents.size() * sizeof(B2iCrcEntry), 256);
	const size_t o_unsup = off;   off = align_up(off + unsup.size() * 4, 256);
	const size_t upload_bytes = off;
	const size_t o_partial = off; off = align_up(off + work.size() * 4, 256);
	const size_t o_counter = off; off = align_up(off + 8, 256);
	const size_t o_results = off; off = align_up(off + n * sizeof(B2iResult), 256);
	p->block_bytes = off;
	p->results_off = o_results;

	if (mem_d != NULL) {
		if (off > mem_cap) {
			delete p;
			return fail(c, B2I_E_NOMEM, "plan arena too small (%zu > %zu)", off, mem_cap);
		}
		p->d_block = mem_d;
		p->h_block = mem_h;
	} else if (cudaMalloc(&p->d_block, off ? off : 256) != cudaSuccess ||
	    cudaHostAlloc((void **)&p->h_block, off ? off : 256, cudaHostAllocDefault) != cudaSuccess) {
		b2i_plan_destroy(p);
		return fail(c, B2I_E_NOMEM, "plan allocation of %zu bytes failed", off);
	}
	p->d_descs = (B2iDesc *)(p->d_block + o_descs);
	p->d_order = (uint32_t *)(p->d_block + o_order);
	p->d_work = (B2iCrcWork *)(p->d_block + o_work);
	p->d_ents = (B2iCrcEntry *)(p->d_block + o_ents);
	p->d_unsup = (uint32_t *)(p->d_block + o_unsup);
	p->d_partial = (uint32_t *)(p->d_block + o_partial);
	p->d_counter = (unsigned int *)(p->d_block + o_counter);
	p->d_results = (B2iResult *)(p->d_block + o_results);

	if (!big.empty()) {
		team_lane = (int)((unsigned)team_lane % B2I_TEAM_STREAMS);
		if (c->s_team[team_lane] == NULL &&
		    cudaStreamCreateWithFlags(&c->s_team[team_lane], cudaStreamNonBlocking) != cudaSuccess) {
			b2i_plan_destroy(p);
			return fail(c, B2I_E_CUDA, "team stream");
		}
		p->team_stream = c->s_team[team_lane];
		if (cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess) {
			b2i_plan_destroy(p);
			return fail(c, B2I_E_CUDA, "team events");
		}
	}
	if (n) memcpy(p->h_block + o_descs, descs, n * sizeof(B2iDesc));
	if (!deflate.empty()) memcpy(p->h_block + o_order, deflate.data(), deflate.size() * 4);
	if (!big.empty()) memcpy(p->h_block + o_order + deflate.size() * 4, big.data(), big.size() * 4);
	if (!work.empty()) memcpy(p->h_block + o_work, work.data(), work.size() * sizeof(B2iCrcWork));
	if (!ents.empty()) memcpy(p->h_block + o_ents, ents.data(), ents.size() * sizeof(B2iCrcEntry));
	if (!unsup.empty()) memcpy(p->h_block + o_unsup, unsup.data(), unsup.size() * 4);
	cudaError_t e = cudaMemcpyAsync(p->d_block, p->h_block, upload_bytes, cudaMemcpyHostToDevice, upload_stream);
	if (e != cudaSuccess) {
		b2i_plan_destroy(p);
		return fail(c, B2I_E_CUDA, "plan upload: %s", cudaGetErrorString(e));
	}
	*out = p;
	return B2I_OK;
}

extern "C" int b2i_plan_create(b2i_ctx *c, const b2i_stream_desc *descs, size_t n, b2i_plan **out)
{
	if (c == NULL)
		return B2I_E_INVAL;
	return b2i_plan_build(c, descs, n, c->stream, c->stream, NULL, NULL, 0, out);
}

extern "C" int b2i_plan_launch(b2i_plan *p, const void *d_in, size_t in_bytes, void *d_out, size_t out_bytes)
{
	if (p == NULL)
		return B2I_E_INVAL;
	b2i_ctx *c = p->ctx;
	if (p->n == 0)
		return B2I_OK;
	if (d_in == NULL || p->max_in_end > in_bytes)
		return fail(c, B2I_E_INVAL, "input buffer too small: plan reads up to %llu, have %zu",
		    (unsigned long long)p->max_in_end, in_bytes);
	if (p->max_out_end > out_bytes || (p->max_out_end && d_out == NULL))
		return fail(c, B2I_E_INVAL, "output buffer too small: plan writes up to %llu, have %zu",
		    (unsigned long long)p->max_out_end, out_bytes);
	if (p->need_aligned_in && ((uintptr_t)d_in & 15))
		return fail(c, B2I_E_INVAL, "d_in must be 16-byte aligned");
	if (p->max_out_end && ((uintptr_t)d_out & 15))
		return fail(c, B2I_E_INVAL, "d_out must be 16-byte aligned");
	CU(c, cudaSetDevice(c->device));
	if (p->n_big) {
		/* fork: the team kernel (one CTA per large stream) runs on its own stream next to
		 * everything else; the single-warp kernel is launched first so that the short
		 * streams occupy the SMs while they last and the team CTAs fill in behind them */
		CU(c, cudaMemsetAsync(p->d_counter, 0, 8, p->stream));
		CU(c, cudaEventRecord(p->ev_fork, p->stream));
		CU(c, cudaStreamWaitEvent(p->team_stream, p->ev_fork, 0));
	}
	if (p->n_deflate) {
		if (!p->n_big)
			CU(c, cudaMemsetAsync(p->d_counter, 0, 4, p->stream));
		/* Two builds of the kernel (inflate_core.cuh): with at least two waves of streams
		 * in the batch the one with more resident warps per SM wins (+5..7 % measured on
		 * 16 384 and 500 000 streams), below that the one that is faster per stream (config 1). */
		size_t streams = p->batch_streams > p->n_deflate ? p->batch_streams : p->n_deflate;
		bool r9 = streams >= (size_t)c->num_sms * 28u * 2u;
		if (const char *ev = getenv("B2I_KERNEL"))
			r9 = strcmp(ev, "r9") == 0;
		CU(c, (r9 ? b2i_launch_inflate_r9 : b2i_launch_inflate)((const uint8_t *)d_in, in_bytes, (uint8_t *)d_out,
		    p->out_mirror, p->d_descs, p->d_results, p->d_order, p->n_deflate, p->d_counter, c->d_crc_tab, c->d_xp8,
		    getenv("B2I_UNIFORM_ONLY") ? NULL : c->d_scratch, c->d_slot_busy, c->num_sms, p->stream));
		c->launches++;
	}
	if (p->n_big) {
		CU(c, b2i_launch_inflate_team((const uint8_t *)d_in, in_bytes, (uint8_t *)d_out, p->out_mirror, p->d_descs,
		    p->d_results, p->d_order + p->n_deflate, p->n_big, p->d_counter + 1, c->d_crc_tab, c->d_xp8,
		    c->d_scratch, c->d_slot_busy, c->num_sms, p->team_stream));
		CU(c, cudaEventRecord(p->ev_join, p->team_stream));
		c->launches++;
	}
	if (p->n_stored) {
		if (p->n_work) {
			CU(c, b2i_launch_crc_chunks((const uint8_t *)d_in, (uint8_t *)d_out, p->d_descs, p->d_work,
			    p->n_work, p->d_partial, c->d_crc_tab, c->d_xp8, c->d_ztab, c->d_lane_mul, c->num_sms,
			    p->stream));
			c->launches++;
		}
		CU(c, b2i_launch_crc_combine(p->d_descs, p->d_results, p->d_ents, p->n_stored, p->d_work,
		    p->d_partial, c->d_xp8, p->stream));
		c->launches++;
	}
	if (p->n_unsup) {
		CU(c, b2i_launch_unsupported(p->d_descs, p->d_results, p->d_unsup, p->n_unsup, p->stream));
		c->launches++;
	}
	/* join: the plan's stream waits for the team kernel - unless the caller takes care of
	 * that (a pipelined slice orders only its copy-out behind it, so that the compute
	 * stream is not held up for as long as the slice's largest stream takes) */
	if (p->n_big && !p->defer_join)
		CU(c, cudaStreamWaitEvent(p->stream, p->ev_join, 0));
	return B2I_OK;
}

extern "C" int b2i_plan_results(b2i_plan *p, b2i_stream_result *res)
{
	if (p == NULL || (p->n && res == NULL))
		return B2I_E_INVAL;
	b2i_ctx *c = p->ctx;
	if (p->n == 0)
		return B2I_OK;
	CU(c, cudaMemcpyAsync(p->h_block + p->results_off, p->d_results, p->n * sizeof(B2iResult),
	    cudaMemcpyDeviceToHost, p->stream));
	CU(c, cudaStreamSynchronize(p->stream));
	memcpy(res, p->h_block + p->results_off, p->n * sizeof(B2iResult));
	return B2I_OK;
}

extern "C" void b2i_plan_destroy(b2i_plan *p)
{
	if (p == NULL)
		return;
	cudaSetDevice(p->ctx->device);
	if (p->owns_memory) {
		cudaStreamSynchronize(p->stream);
		cudaFree(p->d_block);
		cudaFreeHost(p->h_block);
	}
	/* arena-backed plans (slices of a job) are destroyed after the job's last event has
	 * been waited for: synchronising their compute stream here would also wait for the
	 * OTHER jobs in flight on it and undo the overlap */
	if (p->ev_fork) cudaEventDestroy(p->ev_fork);
	if (p->ev_join) cudaEventDestroy(p->ev_join);
	delete p;
}

/* ---- host-buffer path --------------------------------------------------------- */

/* Growing a job's device staging frees the old block, and cudaFree waits for EVERYTHING in
 * flight on the device - with other jobs running that stalls the submitting thread for
 * their whole duration (3 ms measured inside the streaming engine, where the job slots
 * meet windows of different sizes in a different order every pass).  So a block that has
 * to grow grows to the largest request the context has seen (*high), and every slot
 * settles after at most one reallocation. */
static int ensure_dev(b2i_ctx *c, uint8_t **buf, size_t *cap, size_t need, size_t *high)
{
	need = align_up(need, 16) + 16;
	if (need > *high)
		*high = need;
	if (*cap >= need)
		return B2I_OK;
	if (*buf) {
		cudaStreamSynchronize(c->stream);
		cudaFree(*buf);
		*buf = NULL;
		*cap = 0;
	}
	size_t want = std::max(*high, *cap + *cap / 2);
	if (cudaMalloc(buf, want) != cudaSuccess) {
		if (cudaMalloc(buf, need) != cudaSuccess)
			return fail(c, B2I_E_NOMEM, "device allocation of %zu bytes failed", need);
		want = need;
	}
	*cap = want;
	return B2I_OK;
}

static int ensure_pipe(b2i_ctx *c)
{
	if (c->pipe_ready)
		return B2I_OK;
	CU(c, cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
	CU(c, cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
	for (int i = 0; i < B2I_PIPE_STREAMS; i++)
		CU(c, cudaStreamCreateWithFlags(&c->s_cmp[i], cudaStreamNonBlocking));
	CU(c, cudaEventCreateWithFlags(&c->ev_free, cudaEventDisableTiming));
	c->pipe_ready = true;
	return B2I_OK;
}

static int ensure_arena(b2i_ctx *c, b2i_job *J, size_t need)
{
	if (need > c->high_arena)
		c->high_arena = need;
	if (J->arena_cap >= need)
		return B2I_OK;
	cudaDeviceSynchronize();
	cudaFree(J->arena_d);
	cudaFreeHost(J->arena_h);
	J->arena_d = NULL; J->arena_h = NULL; J->arena_cap = 0;
	need = c->high_arena + c->high_arena / 2;          /* see ensure_dev */
	if (cudaMalloc(&J->arena_d, need) != cudaSuccess ||
	    cudaHostAlloc((void **)&J->arena_h, need, cudaHostAllocDefault) != cudaSuccess)
		return fail(c, B2I_E_NOMEM, "plan arena of %zu bytes", need);
	J->arena_cap = need;
	return B2I_OK;
}

/*
 * Host buffers in, host buffers out.  The descriptors are cut into up to
 * B2I_PIPE_SLICES contiguous slices of equal weight; slice s is copied in on
 * the copy-in stream, decoded on a compute stream as soon as its bytes have
 * landed (the slice kernels are small enough to be co-resident), and copied
 * out on the copy-out stream as soon as its kernel is done - so H2D of slice
 * s+1, the kernel of slice s and D2H of slice s-1 overlap.
 *
 * b2i_submit queues all of that and returns; b2i_wait blocks until the job's
 * last copy has landed and hands out the results.  Up to B2I_MAX_JOBS jobs of
 * one context may be in flight: they share the three streams (so copies of one
 * direction stay in submission order) and own their device staging, so the
 * copy-out of job k overlaps the copy-in and kernels of job k+1.
 * b2i_decode_host is submit + wait.
 */
static void job_drain(b2i_ctx *c, b2i_job *J)
{
	cudaStreamSynchronize(c->s_out);
	cudaStreamSynchronize(c->s_in);
	for (int i = 0; i < B2I_PIPE_STREAMS; i++)
		cudaStreamSynchronize(c->s_cmp[i]);
	for (size_t s = 0; s < B2I_PIPE_SLICES; s++) {
		if (J->plans[s])
			b2i_plan_destroy(J->plans[s]);
		J->plans[s] = NULL;
	}
	J->busy = false;
}

extern "C" int b2i_submit(b2i_ctx *c, const void *host_in, size_t in_bytes,
    const b2i_stream_desc *descs, size_t n, void *host_out, size_t out_bytes, b2i_job **job)
{
	if (c == NULL || job == NULL)
		return B2I_E_INVAL;
	*job = NULL;
	if (n != 0 && (host_in == NULL || descs == NULL))
		return fail(c, B2I_E_INVAL, "b2i_submit: NULL argument");
	CallTimer ct_(g_cs_submit);
	b2i_job *J = NULL;
	for (int i = 0; i < B2I_MAX_JOBS && J == NULL; i++)
		if (!c->jobs[i].busy)
			J = &c->jobs[i];
	if (J == NULL)
		return fail(c, B2I_E_INVAL, "b2i_submit: %d jobs already in flight on this context", B2I_MAX_JOBS);
	J->ctx = c;
	J->n = n;
	J->K = 0;
	J->staged = false;
	for (size_t s = 0; s < B2I_PIPE_SLICES; s++)
		J->plans[s] = NULL;
	CU(c, cudaSetDevice(c->device));
	int rc;
	if ((rc = ensure_pipe(c)) != B2I_OK)
		return rc;
	if (!J->ev_ready) {
		for (int i = 0; i < B2I_PIPE_SLICES; i++) {
			CU(c, cudaEventCreateWithFlags(&J->ev_in[i], cudaEventDisableTiming));
			CU(c, cudaEventCreateWithFlags(&J->ev_k[i], cudaEventDisableTiming));
		}
		CU(c, cudaEventCreateWithFlags(&J->ev_done, cudaEventDisableTiming));
		J->ev_ready = true;
	}
	if (n == 0) {
		J->busy = true;
		CU(c, cudaEventRecord(J->ev_done, c->s_out));
		*job = J;
		return B2I_OK;
	}
	if ((rc = ensure_dev(c, &J->d_in, &J->d_in_cap, in_bytes, &c->high_in)) != B2I_OK)
		return rc;
	if ((rc = ensure_dev(c, &J->d_out, &J->d_out_cap, out_bytes, &c->high_out)) != B2I_OK)
		return rc;

	/* slices: contiguous descriptor ranges of about equal csize + usize */
	uint64_t total_w = 0, stored_bytes = 0;
	for (size_t i = 0; i < n; i++) {
		total_w += descs[i].in_len + descs[i].out_cap;
		if (descs[i].method == B2I_METHOD_STORED)
			stored_bytes += descs[i].in_len;
		if ((descs[i].method == B2I_METHOD_DEFLATE || descs[i].method == B2I_METHOD_STORED) &&
		    descs[i].in_off + descs[i].in_len > in_bytes)
			return fail(c, B2I_E_INVAL, "a stream extends past the input buffer");
	}
	/* Equal slices.  (A schedule with small first slices, meant to start the
	 * copy-out stream earlier, measured slower: 37.5 vs 40.0 GB/s on config 1.) */
	double frac[B2I_PIPE_SLICES];
	size_t K = (size_t)std::min<uint64_t>(8, std::max<uint64_t>(1, total_w / (32u << 20)));
	if (n < 16 * K)
		K = 1;
	for (size_t k = 0; k < K; k++)
		frac[k] = (double)(k + 1) / (double)K;
	if (const char *ek = getenv("B2I_PIPE_SLICES")) {        /* tuning knob: equal slices */
		int v = atoi(ek);
		if (v >= 1 && v <= B2I_PIPE_SLICES && n >= 16u * (size_t)v) {
			K = (size_t)v;
			for (size_t k = 0; k < K; k++)
				frac[k] = (double)(k + 1) / (double)K;
		}
	}
	size_t *cut = J->cut;
	{
		uint64_t acc = 0;
		size_t k = 1;
		cut[0] = 0;
		for (size_t i = 0; i < n && k < K; i++) {
			acc += descs[i].in_len + descs[i].out_cap;
			while (k < K && (double)acc >= (double)total_w * frac[k - 1])
				cut[k++] = i + 1;
		}
		while (k <= K)
			cut[k++] = n;
	}
	const size_t per_slice = plan_block_bound(n, stored_bytes);      /* each slice <= whole */
	size_t arena_need = 0;
	size_t arena_off[B2I_PIPE_SLICES];
	for (size_t s = 0; s < K; s++) {
		arena_off[s] = arena_need;
		arena_need += align_up(plan_block_bound(cut[s + 1] - cut[s], stored_bytes), 256);
	}
	(void)per_slice;
	if ((rc = ensure_arena(c, J, arena_need)) != B2I_OK)
		return rc;
	J->K = K;
	/* work of an earlier call on the caller's stream (if any) comes first */
	CU(c, cudaEventRecord(c->ev_free, c->stream));
	CU(c, cudaStreamWaitEvent(c->s_in, c->ev_free, 0));
	J->busy = true;

	/* Experimental (B2I_MIRROR=1): when host_out is pinned the inflate kernel can
	 * store every 16-byte unit to it as well, so decoded bytes cross the host link
	 * while the kernel runs.  Measured on B200: SM stores to system memory throttle
	 * the kernel (22 GB/s end to end vs 31 GB/s with the copy pass), so it is off. */
	uint8_t *mirror = NULL;
	if (host_out != NULL && getenv("B2I_MIRROR") != NULL) {       /* opt-in: measured slower than the copy pass */
		cudaPointerAttributes pa;
		if (cudaPointerGetAttributes(&pa, host_out) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
		    pa.devicePointer != NULL)
			mirror = (uint8_t *)pa.devicePointer;
		else
			cudaGetLastError();
		for (size_t i = 0; i < n && mirror; i++)
			if (descs[i].method == B2I_METHOD_STORED && !(descs[i].flags & B2I_F_NO_COPY))
				mirror = NULL;
	}
	/* pageable input of some size is staged through the job's own pinned buffer */
	uint8_t *stage = NULL;
	if (in_bytes >= ((size_t)4 << 20) && in_bytes <= ((size_t)1 << 30) && getenv("B2I_NO_STAGE") == NULL) {
		cudaPointerAttributes pa;
		bool pageable = true;
		if (cudaPointerGetAttributes(&pa, host_in) == cudaSuccess)
			pageable = pa.type == cudaMemoryTypeUnregistered;
		else
			cudaGetLastError();
		if (pageable) {
			if (J->h_stage_cap < in_bytes + 64) {
				cudaFreeHost(J->h_stage);
				J->h_stage = NULL;
				J->h_stage_cap = 0;
				size_t want = in_bytes + in_bytes / 4 + 64;
				if (want > c->high_stage)
					c->high_stage = want;
				want = c->high_stage;          /* see ensure_dev */
				if (cudaHostAlloc((void **)&J->h_stage, want, cudaHostAllocDefault) == cudaSuccess)
					J->h_stage_cap = want;
				else
					cudaGetLastError();
			}
			stage = J->h_stage;          /* NULL: no pinned memory to be had, the driver copies */
			J->staged = stage != NULL;
		}
	}
	b2i_plan **plans = J->plans;
	rc = B2I_OK;
	for (size_t s = 0; s < K && rc == B2I_OK; s++) {
		const b2i_stream_desc *sd = descs + cut[s];
		const size_t sn = cut[s + 1] - cut[s];
		cudaStream_t cs = c->s_cmp[((size_t)(J - c->jobs) * B2I_PIPE_SLICES + s) % B2I_PIPE_STREAMS];
		if (sn == 0)
			continue;
		/* copy-in stream, in order: this slice's plan block, then its bytes.  (Copies
		 * of one direction execute in submission order, so the small plan upload
		 * must not queue behind later slices' data.) */
		rc = b2i_plan_build(c, sd, sn, cs, c->s_in, J->arena_d + arena_off[s], J->arena_h + arena_off[s],
		    align_up(plan_block_bound(sn, stored_bytes), 256), &plans[s], (int)((J - c->jobs) * B2I_PIPE_SLICES + s));
		if (rc != B2I_OK)
			break;
		/* the slice's input: runs of streams that follow each other in the source (one run
		 * for a contiguous range of an archive; several when the caller handed us a
		 * subset, e.g. one GPU's share of an LPT partition) */
		{
			uint64_t lo = ~0ull, hi = 0;
			bool bad = false;
			auto flush_in = [&]() {
				if (lo < hi) {
					lo &= ~(uint64_t)15;
					const uint8_t *src = (const uint8_t *)host_in + lo;
					if (stage != NULL) {
						/* pageable input: several threads copy the run into pinned staging and
						 * the DMA runs from there (the driver's own pageable path is a single
						 * blocking copy at about 10 GB/s) */
						b2i_parallel_copy(stage + lo, src, hi - lo);
						src = stage + lo;
					}
					cudaError_t e = cudaMemcpyAsync(J->d_in + lo, src, hi - lo,
					    cudaMemcpyHostToDevice, c->s_in);
					if (e != cudaSuccess) { rc = fail(c, B2I_E_CUDA, "H2D: %s", cudaGetErrorString(e)); bad = true; }
				}
				lo = ~0ull; hi = 0;
			};
			for (size_t i = 0; i < sn && !bad; i++) {
				if (sd[i].method != B2I_METHOD_DEFLATE && sd[i].method != B2I_METHOD_STORED)
					continue;
				const uint64_t a = sd[i].in_off, b = sd[i].in_off + sd[i].in_len;
				if (lo < hi && (a + (256u << 10) < lo || a > hi + (256u << 10)))
					flush_in();
				lo = std::min<uint64_t>(lo, a);
				hi = std::max<uint64_t>(hi, b);
			}
			if (!bad)
				flush_in();
			if (bad)
				break;
		}
		if (cudaEventRecord(J->ev_in[s], c->s_in) != cudaSuccess ||
		    cudaStreamWaitEvent(cs, J->ev_in[s], 0) != cudaSuccess) {
			rc = fail(c, B2I_E_CUDA, "event"); break;
		}
		plans[s]->out_mirror = mirror;
		plans[s]->batch_streams = n;
		plans[s]->defer_join = true;
		rc = b2i_plan_launch(plans[s], J->d_in, in_bytes, J->d_out, out_bytes);
		if (rc != B2I_OK)
			break;
		if (cudaEventRecord(J->ev_k[s], cs) != cudaSuccess ||
		    cudaStreamWaitEvent(c->s_out, J->ev_k[s], 0) != cudaSuccess ||
		    (plans[s]->n_big && cudaStreamWaitEvent(c->s_out, plans[s]->ev_join, 0) != cudaSuccess)) {
			rc = fail(c, B2I_E_CUDA, "event"); break;
		}
		/* copy-out stream, in order: the slice's decoded bytes, then its results */
		if (host_out != NULL && mirror == NULL && plans[s]->max_out_end) {
			/* runs of streams whose outputs follow each other (gaps are alignment padding
			 * only): bytes between two runs belong to somebody else and are not touched */
			uint64_t olo = ~0ull, ohi = 0;
			bool bad = false;
			auto flush_out = [&]() {
				if (olo < ohi && ohi <= out_bytes) {
					cudaError_t e = cudaMemcpyAsync((uint8_t *)host_out + olo, J->d_out + olo, ohi - olo,
					    cudaMemcpyDeviceToHost, c->s_out);
					if (e != cudaSuccess) { rc = fail(c, B2I_E_CUDA, "D2H: %s", cudaGetErrorString(e)); bad = true; }
				}
				olo = ~0ull; ohi = 0;
			};
			for (size_t i = 0; i < sn && !bad; i++) {
				if (!(sd[i].method == B2I_METHOD_DEFLATE ||
				    (sd[i].method == B2I_METHOD_STORED && !(sd[i].flags & B2I_F_NO_COPY))))
					continue;
				const uint64_t a = sd[i].out_off, b = sd[i].out_off + sd[i].out_cap;
				if (olo < ohi && (a < ohi || a > ohi + 15))
					flush_out();
				olo = std::min<uint64_t>(olo, a);
				ohi = std::max<uint64_t>(ohi, b);
			}
			if (!bad)
				flush_out();
			if (bad)
				break;
		}
		if (cudaMemcpyAsync(plans[s]->h_block + plans[s]->results_off, plans[s]->d_results,
		    sn * sizeof(B2iResult), cudaMemcpyDeviceToHost, c->s_out) != cudaSuccess) {
			rc = fail(c, B2I_E_CUDA, "results D2H"); break;
		}
	}
	/* everything funnels into the copy-out stream (it waited on every slice's kernel) */
	if (rc == B2I_OK && cudaEventRecord(J->ev_done, c->s_out) != cudaSuccess)
		rc = fail(c, B2I_E_CUDA, "event");
	if (rc != B2I_OK) {
		job_drain(c, J);
		return rc;
	}
	*job = J;
	return B2I_OK;
}

extern "C" const void *b2i_job_staged_input(const b2i_job *J)
{
	return (J != NULL && J->busy && J->staged) ? J->h_stage : NULL;
}

/* the job's input has been copied to the device: host_in may be reused or released */
extern "C" int b2i_job_wait_input(b2i_job *J)
{
	if (J == NULL || !J->busy)
		return B2I_E_INVAL;
	if (J->K == 0)
		return B2I_OK;
	b2i_ctx *c = J->ctx;
	CU(c, cudaEventSynchronize(J->ev_in[J->K - 1]));     /* copies of one direction run in submission order */
	return B2I_OK;
}

extern "C" int b2i_wait(b2i_job *J, b2i_stream_result *res)
{
	if (J == NULL || !J->busy)
		return B2I_E_INVAL;
	CallTimer ct_(g_cs_wait);
	b2i_ctx *c = J->ctx;
	int rc = B2I_OK;
	cudaError_t e = cudaEventSynchronize(J->ev_done);
	if (e != cudaSuccess)
		rc = fail(c, B2I_E_CUDA, "pipeline: %s", cudaGetErrorString(e));
	if (rc == B2I_OK && J->n != 0 && res == NULL)
		rc = fail(c, B2I_E_INVAL, "b2i_wait: NULL results");
	for (size_t s = 0; s < J->K; s++) {
		if (J->plans[s] == NULL)
			continue;
		if (rc == B2I_OK)
			memcpy(res + J->cut[s], J->plans[s]->h_block + J->plans[s]->results_off,
			    (J->cut[s + 1] - J->cut[s]) * sizeof(B2iResult));
		b2i_plan_destroy(J->plans[s]);
		J->plans[s] = NULL;
	}
	J->busy = false;
	return rc;
}

extern "C" int b2i_decode_host(b2i_ctx *c, const void *host_in, size_t in_bytes,
    const b2i_stream_desc *descs, size_t n, void *host_out, size_t out_bytes,
    b2i_stream_result *res)
{
	if (c == NULL)
		return B2I_E_INVAL;
	if (n == 0)
		return B2I_OK;
	if (host_in == NULL || descs == NULL || res == NULL)
		return fail(c, B2I_E_INVAL, "b2i_decode_host: NULL argument");
	CallTimer ct_(g_cs_decode);
	b2i_job *J = NULL;
	int rc = b2i_submit(c, host_in, in_bytes, descs, n, host_out, out_bytes, &J);
	if (rc != B2I_OK)
		return rc;
	return b2i_wait(J, res);
}

/* ---- scalar CRC drop-ins ---------------------------------------------------- */

static uint32_t host_mulmod(uint32_t a, uint32_t b)
{
	uint32_t p = 0;
	for (int i = 0; i < 32; i++) {
		if (a & (0x80000000u >> i))
			p ^= b;
		b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
	}
	return p;
}

extern "C" uint32_t b2i_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b)
{
	uint32_t r = 0x80000000u, sq = 0x00800000u;   /* x^0, x^8 */
	while (len_b) {
		if (len_b & 1)
			r = host_mulmod(r, sq);
		sq = host_mulmod(sq, sq);
		len_b >>= 1;
	}
	return host_mulmod(r, crc_a) ^ crc_b;
}

static int crc_of_device_span(b2i_ctx *c, uint32_t crc, const void *d_buf, size_t len, uint32_t *out)
{
	b2i_stream_desc d;
	b2i_stream_result r;
	b2i_plan *p = NULL;
	int rc;

	memset(&d, 0, sizeof(d));
	d.in_len = len;
	d.expect_out = len;
	d.method = B2I_METHOD_STORED;
	d.flags = B2I_F_NO_COPY;
	if ((rc = b2i_plan_create(c, &d, 1, &p)) != B2I_OK)
		return rc;
	rc = b2i_plan_launch(p, d_buf, len, NULL, 0);
	if (rc == B2I_OK)
		rc = b2i_plan_results(p, &r);
	b2i_plan_destroy(p);
	if (rc != B2I_OK)
		return rc;
	/* chaining: crc(c, M) = combine(c, crc(0, M), |M|)   (archive_crc32.h:43-84 contract) */
	*out = b2i_crc32_combine(crc, r.crc, len);
	return B2I_OK;
}

extern "C" int b2i_crc32_device(b2i_ctx *c, uint32_t crc, const void *d_buf, size_t len, uint32_t *out)
{
	if (c == NULL || out == NULL)
		return B2I_E_INVAL;
	if (d_buf == NULL) {          /* crc32(x, NULL, 0) == 0 */
		*out = 0;
		return B2I_OK;
	}
	if (len == 0) {
		*out = crc;
		return B2I_OK;
	}
	return crc_of_device_span(c, crc, d_buf, len, out);
}

extern "C" int b2i_crc32(b2i_ctx *c, uint32_t crc, const void *host_buf, size_t len, uint32_t *out)
{
	if (c == NULL || out == NULL)
		return B2I_E_INVAL;
	if (host_buf == NULL) {
		*out = 0;
		return B2I_OK;
	}
	if (len == 0) {
		*out = crc;
		return B2I_OK;
	}
	CU(c, cudaSetDevice(c->device));
	int rc = ensure_dev(c, &c->d_in, &c->d_in_cap, len, &c->high_crc);
	if (rc != B2I_OK)
		return rc;
	CU(c, cudaMemcpyAsync(c->d_in, host_buf, len, cudaMemcpyHostToDevice, c->stream));
	return crc_of_device_span(c, crc, c->d_in, len, out);
}

extern "C" void b2i_free(void *p) { free(p); }
