/*
 * b2i_pipe.cpp — the streaming host engine behind b2i_pipe_* (include/b200inflate.h).
 *
 * The reference serves an archive entry by entry from one inflate stream
 * (archive_read_support_format_zip.c:2535-2690) fed by block-sized slices of the
 * source (archive_read_open_filename.c:389-461).  Here ONE archive's descriptors are
 * cut into windows of bounded output size; windows travel through a ring of pinned
 * staging buffers and are decoded by the GPUs of the box round-robin (window k on
 * device k mod G: the partition of SURVEY 8e, taken dynamically so that every GPU
 * always has two windows in flight), while the caller consumes window after window
 * in archive order:
 *
 *      fill      compressed span of window k -> pinned input buffer of its slot:
 *                memory sources by the device's worker thread (a small pool of copy
 *                threads, or no copy at all when the memory is already pinned),
 *                callback sources by the CALLER's thread from inside b2i_pipe_get
 *                (libarchive's read filters may only be used on the caller's thread)
 *      decode    b2i_submit on the slot's device (H2D, kernels, D2H into the slot's
 *                pinned output buffer), four jobs in flight per device
 *      serve     b2i_pipe_get(i) blocks until stream i's window has landed and hands
 *                out pointers into the slot; b2i_pipe_release recycles slots
 *
 * Memory is bounded by slots x (window input + window output) however large the
 * archive; a stream larger than the window gets a window of its own.
 * No CPU decode path: every byte is produced by the kernels.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/b200inflate.h"

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }

/* ---- process-wide cache of pinned buffers (pinning costs more than decoding) ---- */
struct PinnedPool {
	std::mutex mu;
	struct Buf { void *p; size_t cap; };
	std::vector<Buf> free_;
	size_t cached = 0;
	/* idle bytes kept: one ring of the largest default windows (5 x 512 MiB out plus their
	 * input, about 4 GiB for text) must fit, or every pass pins its buffers again - 0.3 s
	 * per GiB, more than the decode (measured: config 4 at 3 GB/s instead of 14) */
	size_t kMaxCached = getenv("B2I_PINNED_CACHE_MB") ? (size_t)atol(getenv("B2I_PINNED_CACHE_MB")) << 20 : (size_t)5 << 30;

	void *get(size_t need, size_t *cap)
	{
		{
			std::lock_guard<std::mutex> g(mu);
			size_t best = free_.size();
			for (size_t i = 0; i < free_.size(); i++)
				if (free_[i].cap >= need && (best == free_.size() || free_[i].cap < free_[best].cap))
					best = i;
			if (best != free_.size() && free_[best].cap <= 4 * need + ((size_t)1 << 20)) {
				Buf b = free_[best];
				free_.erase(free_.begin() + (long)best);
				cached -= b.cap;
				*cap = b.cap;
				return b.p;
			}
		}
		void *p = NULL;
		size_t want = align_up(need ? need : 1, (size_t)1 << 20);
		if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
			cudaGetLastError();
			return NULL;
		}
		*cap = want;
		return p;
	}
	void put(void *p, size_t cap)
	{
		if (p == NULL)
			return;
		std::vector<Buf> drop;
		{
			/* the newest buffers stay (the next archive most likely wants the same sizes);
			 * the oldest ones go when the cache is over its limit */
			std::lock_guard<std::mutex> g(mu);
			free_.push_back({p, cap});
			cached += cap;
			while (cached > kMaxCached && !free_.empty()) {
				drop.push_back(free_.front());
				cached -= free_.front().cap;
				free_.erase(free_.begin());
			}
		}
		for (const Buf &b : drop)
			cudaFreeHost(b.p);
	}
};
PinnedPool g_pinned;

/* ---- a few threads that copy pageable memory into pinned staging ------------------ */
struct CopyPool {
	struct Task { void *dst; const void *src; size_t len; std::atomic<int> *left; };
	std::mutex mu;
	std::condition_variable cv, cv_done;
	std::deque<Task> q;
	std::vector<std::thread> th;
	bool quit = false;

	explicit CopyPool(int n)
	{
		for (int i = 0; i < n; i++)
			th.emplace_back([this] { run(); });
	}
	~CopyPool()
	{
		{
			std::lock_guard<std::mutex> g(mu);
			quit = true;
		}
		cv.notify_all();
		for (auto &t : th)
			t.join();
	}
	void run()
	{
		for (;;) {
			Task t;
			{
				std::unique_lock<std::mutex> g(mu);
				cv.wait(g, [this] { return quit || !q.empty(); });
				if (q.empty())
					return;
				t = q.front();
				q.pop_front();
			}
			memcpy(t.dst, t.src, t.len);
			if (t.left->fetch_sub(1) == 1) {
				std::lock_guard<std::mutex> g(mu);
				cv_done.notify_all();
			}
		}
	}
	/* copy [src, src+len) to dst with the pool's threads and the calling thread */
	void copy(void *dst, const void *src, size_t len)
	{
		const size_t nt = th.size() + 1;
		if (len < ((size_t)1 << 20) || nt == 1) {
			memcpy(dst, src, len);
			return;
		}
		const size_t piece = align_up((len + nt - 1) / nt, 4096);
		std::atomic<int> left{0};
		size_t off = piece;         /* piece 0 is ours */
		int queued = 0;
		{
			std::lock_guard<std::mutex> g(mu);
			while (off < len) {
				const size_t l = std::min(piece, len - off);
				q.push_back({(uint8_t *)dst + off, (const uint8_t *)src + off, l, &left});
				off += l;
				queued++;
			}
			left.store(queued);
		}
		cv.notify_all();
		memcpy(dst, src, std::min(piece, len));
		std::unique_lock<std::mutex> g(mu);
		cv_done.wait(g, [&] { return left.load() == 0; });
	}
};

/* one pool for the whole process, started on first use */
CopyPool *shared_copiers()
{
	static std::mutex mu;
	static CopyPool *pool = NULL;
	std::lock_guard<std::mutex> g(mu);
	if (pool == NULL) {
		int n = 6;
		if (const char *ev = getenv("B2I_COPY_THREADS"))
			n = std::max(0, atoi(ev));
		pool = new CopyPool(n);
	}
	return pool;
}

enum SlotState { EMPTY = 0, FILLING, FILLED, INFLIGHT, READY };

struct Window {
	size_t first, count;
	uint64_t in_lo, in_hi;       /* source span, in_lo 16-byte aligned */
	size_t out_bytes;
};

struct Slot {
	SlotState state = EMPTY;
	size_t window = (size_t)-1;
	uint8_t *h_in = NULL;  size_t h_in_cap = 0;
	uint8_t *h_out = NULL; size_t h_out_cap = 0;
	const uint8_t *in_base = NULL;       /* where the window's span starts (h_in, or the pinned source itself) */
	std::vector<b2i_stream_desc> descs;  /* window-local offsets */
	std::vector<b2i_stream_result> res;
};

} // namespace

/* pageable -> pinned with several threads (the driver's own pageable path runs at about
 * 10 GB/s and blocks the caller): used by the pipeline and by b2i_submit's staging */
void b2i_parallel_copy(void *dst, const void *src, size_t len)
{
	shared_copiers()->copy(dst, src, len);
}

struct b2i_pipe {
	/* B2I_PIPE_TRACE=1: one line per window event on stderr, microseconds since open */
	bool trace = getenv("B2I_PIPE_TRACE") != NULL;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void note(const char *what, size_t k) const
	{
		if (trace)
			fprintf(stderr, "B2I_PIPE %9.1f us  window %3zu  %s\n",
			    std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(), k, what);
	}
	std::vector<b2i_ctx *> ctxs;
	const uint8_t *mem = NULL;
	uint64_t mem_size = 0;
	bool mem_pinned = false;
	b2i_fill_fn fill = NULL;
	void *user = NULL;
	std::vector<b2i_stream_desc> descs;
	std::vector<Window> win;
	std::vector<Slot> slots;
	std::vector<std::thread> workers;
	CopyPool *copiers = NULL;

	std::mutex mu;
	std::condition_variable cv_work, cv_ready;
	size_t fill_cursor = 0;      /* next window the caller's thread fills (callback sources) */
	size_t released_upto = 0;    /* streams below this index are no longer needed */
	size_t skip_floor = 0;       /* windows below this one are not decoded any more */
	bool closing = false;
	int error = B2I_OK;
	char err[256] = {0};
	uint64_t stat_windows = 0, stat_fill_ns = 0;
	size_t max_jobs = 4;         /* device passes in flight per GPU */
	size_t window_out = 0;       /* output bytes a window is cut at */

	size_t window_of(size_t idx) const
	{
		size_t lo = 0, hi = win.size();
		while (lo + 1 < hi) {
			size_t mid = (lo + hi) / 2;
			if (win[mid].first <= idx) lo = mid; else hi = mid;
		}
		return lo;
	}
};

static void pipe_fail(b2i_pipe *p, int code, const char *fmt, ...)
{
	/* caller holds p->mu */
	if (p->error != B2I_OK)
		return;
	p->error = code;
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(p->err, sizeof(p->err), fmt, ap);
	va_end(ap);
}

static bool slot_buffers(b2i_pipe *p, Slot &s, const Window &w, bool need_in)
{
	/* input spans differ from window to window with the compression ratio: sizes are rounded
	 * up to a coarse step so that a slot does not trade its buffer for a slightly larger one */
	const size_t step = std::max<size_t>((size_t)1 << 20, align_up(p->window_out / 8, (size_t)1 << 20));
	const size_t in_need = align_up((size_t)(w.in_hi - w.in_lo) + 64, step);
	if (need_in && s.h_in_cap < in_need) {
		g_pinned.put(s.h_in, s.h_in_cap);
		s.h_in = (uint8_t *)g_pinned.get(in_need, &s.h_in_cap);
		if (s.h_in == NULL) { s.h_in_cap = 0; return false; }
	}
	const size_t out_need = w.out_bytes + 64;
	if (s.h_out_cap < out_need) {
		g_pinned.put(s.h_out, s.h_out_cap);
		s.h_out = (uint8_t *)g_pinned.get(out_need, &s.h_out_cap);
		if (s.h_out == NULL) { s.h_out_cap = 0; return false; }
	}
	(void)p;
	return true;
}

/* window-local descriptors: input offsets relative to the span, outputs packed */
static void slot_descs(b2i_pipe *p, Slot &s, const Window &w)
{
	s.descs.resize(w.count);
	s.res.resize(w.count);
	size_t out = 0;
	for (size_t i = 0; i < w.count; i++) {
		b2i_stream_desc d = p->descs[w.first + i];
		if (d.method == B2I_METHOD_DEFLATE || d.method == B2I_METHOD_STORED)
			d.in_off -= w.in_lo;
		d.out_off = out;
		if (!(d.method == B2I_METHOD_STORED && (d.flags & B2I_F_NO_COPY)))
			out = align_up(out + d.out_cap, 16);
		s.descs[i] = d;
	}
}

static void worker_main(b2i_pipe *p, size_t dev)
{
	const size_t G = p->ctxs.size();
	b2i_ctx *ctx = p->ctxs[dev];
	struct Pending { size_t k; b2i_job *job; };
	std::deque<Pending> pending;
	size_t k = dev;
	const size_t max_jobs = p->max_jobs;

	for (;;) {
		bool started = false;
		std::unique_lock<std::mutex> lk(p->mu);
		while (k < p->win.size() && k < p->skip_floor && !p->closing)
			k += G;                                   /* the caller jumped past these */
		if (k < p->win.size() && !p->closing && p->error == B2I_OK && pending.size() < max_jobs) {
			Slot &s = p->slots[k % p->slots.size()];
			const Window &w = p->win[k];
			if (p->fill == NULL && s.state == EMPTY) {
				/* memory source: stage the span ourselves */
				s.state = FILLING;
				s.window = k;
				lk.unlock();
				p->note("stage", k);
				bool ok = slot_buffers(p, s, w, !p->mem_pinned);
				if (ok) {
					if (p->mem_pinned) {
						s.in_base = p->mem + w.in_lo;
					} else {
						p->copiers->copy(s.h_in, p->mem + w.in_lo, (size_t)(w.in_hi - w.in_lo));
						s.in_base = s.h_in;
					}
					slot_descs(p, s, w);
				}
				lk.lock();
				if (!ok)
					pipe_fail(p, B2I_E_NOMEM, "pinned staging for window %zu", k);
				s.state = k < p->skip_floor ? EMPTY : FILLED;       /* skipped while we staged it */
			}
			if (s.window == k && s.state == FILLED && p->error == B2I_OK) {
				lk.unlock();
				p->note("submit", k);
				b2i_job *job = NULL;
				int rc = b2i_submit(ctx, s.in_base, (size_t)(w.in_hi - w.in_lo), s.descs.data(), w.count,
				    s.h_out, w.out_bytes, &job);
				lk.lock();
				if (rc != B2I_OK) {
					pipe_fail(p, rc, "window %zu: %s", k, b2i_last_error(ctx));
					p->cv_ready.notify_all();
				} else {
					s.state = INFLIGHT;
					pending.push_back({k, job});
					p->stat_windows++;
					k += G;
					started = true;
				}
			}
		}
		const bool done = (k >= p->win.size() || p->closing || p->error != B2I_OK);
		if (!pending.empty() && (pending.size() >= max_jobs || !started)) {
			/* nothing more to start right now: collect the oldest job */
			Pending pd = pending.front();
			pending.pop_front();
			Slot &s = p->slots[pd.k % p->slots.size()];
			lk.unlock();
			int rc = b2i_wait(pd.job, s.res.data());
			p->note("done", pd.k);
			lk.lock();
			if (rc != B2I_OK)
				pipe_fail(p, rc, "window %zu: %s", pd.k, b2i_last_error(ctx));
			const Window &w = p->win[pd.k];
			if (w.first + w.count <= p->released_upto) {
				s.state = EMPTY;                      /* nobody will ask for it */
				p->cv_work.notify_all();
			} else {
				s.state = READY;
			}
			p->cv_ready.notify_all();
			continue;
		}
		if (started)
			continue;
		if (done && pending.empty())
			break;
		p->cv_work.wait(lk);
	}
}

extern "C" int b2i_pipe_open(b2i_ctx *const *ctxs, int nctx, const void *mem, uint64_t mem_size,
    b2i_fill_fn fill, void *user, const b2i_stream_desc *descs, size_t n, const b2i_pipe_opts *opts,
    b2i_pipe **out)
{
	if (out == NULL)
		return B2I_E_INVAL;
	*out = NULL;
	if (ctxs == NULL || nctx < 1 || (mem == NULL) == (fill == NULL) || (n && descs == NULL))
		return B2I_E_INVAL;
	/* Windows: large enough that a device pass fills the GPU, small enough that the first
	 * bytes arrive early and three passes overlap (copy-in, kernels, copy-out): a sixth of
	 * the batch per device, within 16..256 MiB, unless the caller says otherwise. */
	size_t window_out = opts && opts->window_out_bytes ? opts->window_out_bytes : 0;
	if (window_out == 0) {
		uint64_t total = 0;
		for (size_t i = 0; i < n; i++)
			total += descs[i].out_cap;
		/* memory sources: a quarter of the batch per device, 16..256 MiB (four passes in
		 * flight fill the GPU: measured on config 1 through the public API, 22 -> 28 GB/s
		 * against 43 MiB windows and three passes).  Callback sources are bounded by the
		 * caller's thread reading and copying (a few GB/s): 64 MiB windows keep the pinned
		 * ring - which costs seconds to pin per GiB in a cold process - small. */
		window_out = (size_t)std::min<uint64_t>((uint64_t)256 << 20,
		    std::max<uint64_t>((uint64_t)16 << 20, total / (4u * (unsigned)nctx)));
		/* A window is done when its LONGEST stream is (a 16 MiB entry takes 40 ms on its CTA,
		 * whatever else the window holds), so archives with large entries need more bytes in
		 * flight to keep the device busy: 32 largest-entries per window, up to 512 MiB
		 * (config 4 through the public API: 12.4 GB/s with 256 MiB windows, 21.0 with 512). */
		uint64_t largest = 0;
		for (size_t i = 0; i < n; i++)
			largest = std::max<uint64_t>(largest, descs[i].out_cap);
		if (largest >= ((uint64_t)1 << 20))
			window_out = (size_t)std::max<uint64_t>(window_out,
			    std::min<uint64_t>({(uint64_t)512 << 20, 32 * largest, std::max<uint64_t>(total / 2, (uint64_t)16 << 20)}));
		/* callback sources keep the ring small (64 MiB windows) unless the entries are large:
		 * 16 largest-entries per window, 256 MiB at most (config 4 from a file: 3 GB/s with
		 * 64 MiB windows, which hold four 16 MiB entries each) */
		if (fill != NULL) {
			const size_t small = (size_t)std::max<uint64_t>((uint64_t)64 << 20,
			    std::min<uint64_t>((uint64_t)256 << 20, 16 * largest));
			if (window_out > small)
				window_out = small;
		}
	}
	size_t first_out = opts && opts->first_window_out_bytes ? opts->first_window_out_bytes
	    : std::min<size_t>(window_out / 4, (size_t)16 << 20);
	int depth = opts && opts->windows_per_device > 0 ? opts->windows_per_device : 5;
	if (const char *ev = getenv("B2I_PIPE_WINDOW_MB"))
		window_out = (size_t)std::max(1, atoi(ev)) << 20, first_out = window_out / 4;
	if (const char *ev = getenv("B2I_PIPE_FIRST_MB"))
		first_out = (size_t)std::max(1, atoi(ev)) << 20;
	if (const char *ev = getenv("B2I_PIPE_DEPTH"))
		depth = std::max(2, atoi(ev));
	size_t jobs = 4;
	if (const char *ev = getenv("B2I_PIPE_JOBS"))
		jobs = (size_t)std::min(4, std::max(1, atoi(ev)));

	b2i_pipe *p = new (std::nothrow) b2i_pipe();
	if (p == NULL)
		return B2I_E_NOMEM;
	p->ctxs.assign(ctxs, ctxs + nctx);
	p->max_jobs = jobs;
	p->window_out = window_out;
	p->mem = (const uint8_t *)mem;
	p->mem_size = mem_size;
	p->fill = fill;
	p->user = user;
	p->descs.assign(descs, descs + n);
	if (mem != NULL) {
		cudaPointerAttributes pa;
		if (cudaPointerGetAttributes(&pa, mem) == cudaSuccess && pa.type == cudaMemoryTypeHost)
			p->mem_pinned = true;
		else
			cudaGetLastError();
		if (!p->mem_pinned)
			p->copiers = shared_copiers();
	}

	/* windows: consecutive streams while the output (and the input span) stays bounded */
	const size_t in_limit = 2 * window_out;
	Window cur = {0, 0, ~0ull, 0, 0};
	for (size_t i = 0; i < n; i++) {
		const b2i_stream_desc &d = descs[i];
		const bool data = d.method == B2I_METHOD_DEFLATE || d.method == B2I_METHOD_STORED;
		const size_t o = (d.method == B2I_METHOD_STORED && (d.flags & B2I_F_NO_COPY)) ? 0 : align_up(d.out_cap, 16);
		uint64_t lo = cur.in_lo, hi = cur.in_hi;
		if (data) {
			if (mem != NULL && d.in_off + d.in_len > mem_size) {
				delete p;
				return B2I_E_INVAL;
			}
			lo = std::min<uint64_t>(lo, d.in_off & ~(uint64_t)15);
			hi = std::max<uint64_t>(hi, d.in_off + d.in_len);
		}
		const size_t limit = p->win.empty() ? first_out : window_out;
		if (cur.count && (cur.out_bytes + o > limit || (hi > lo && hi - lo > in_limit))) {
			if (cur.in_lo > cur.in_hi) cur.in_lo = cur.in_hi = 0;
			p->win.push_back(cur);
			cur = {i, 0, ~0ull, 0, 0};
			lo = data ? (d.in_off & ~(uint64_t)15) : ~0ull;
			hi = data ? d.in_off + d.in_len : 0;
		}
		cur.count++;
		cur.out_bytes += o;
		cur.in_lo = lo;
		cur.in_hi = hi;
	}
	if (cur.count) {
		if (cur.in_lo > cur.in_hi) cur.in_lo = cur.in_hi = 0;
		p->win.push_back(cur);
	}
	p->slots.resize((size_t)nctx * (size_t)depth);
	for (int d = 0; d < nctx; d++)
		p->workers.emplace_back(worker_main, p, (size_t)d);
	*out = p;
	return B2I_OK;
}

/* callback sources: the caller's thread fills the next window whose slot is free */
static bool fill_one(b2i_pipe *p, std::unique_lock<std::mutex> &lk)
{
	while (p->fill_cursor < p->win.size() && p->fill_cursor < p->skip_floor)
		p->fill_cursor++;
	if (p->fill == NULL || p->fill_cursor >= p->win.size() || p->error != B2I_OK)
		return false;
	const size_t k = p->fill_cursor;
	Slot &s = p->slots[k % p->slots.size()];
	if (s.state != EMPTY)
		return false;
	const Window &w = p->win[k];
	s.state = FILLING;
	s.window = k;
	p->fill_cursor++;
	lk.unlock();
	bool ok = slot_buffers(p, s, w, true);
	int rc = B2I_OK;
	if (ok) {
		slot_descs(p, s, w);
		s.in_base = s.h_in;
		if (w.in_hi > w.in_lo)
			rc = p->fill(p->user, w.in_lo, w.in_hi - w.in_lo, s.h_in);
	}
	lk.lock();
	if (!ok)
		pipe_fail(p, B2I_E_NOMEM, "pinned staging for window %zu", k);
	else if (rc != B2I_OK)
		pipe_fail(p, rc, "source read failed for window %zu", k);
	s.state = FILLED;
	p->cv_work.notify_all();
	return true;
}

extern "C" int b2i_pipe_get(b2i_pipe *p, size_t idx, const void **out_data, const void **in_data,
    b2i_stream_result *res)
{
	if (p == NULL || idx >= p->descs.size())
		return B2I_E_INVAL;
	const size_t wi = p->window_of(idx);
	std::unique_lock<std::mutex> lk(p->mu);
	if (wi < p->skip_floor && !(p->slots[wi % p->slots.size()].window == wi))
		return B2I_E_INVAL;                 /* skipped earlier */
	if (wi > p->skip_floor) {
		/* the caller moved on: windows in front of this one that have not been started are
		 * dropped, staged input that nobody will decode gives its slot back */
		p->skip_floor = wi;
		for (Slot &t : p->slots)
			if (t.state == FILLED && t.window != (size_t)-1 && t.window < wi)
				t.state = EMPTY;
		p->cv_work.notify_all();
	}
	/* asking for a stream gives up every window in front of its own: their slots go back
	 * to the ring (the pointers of an earlier get stay valid only within one window) */
	if (p->win[wi].first > p->released_upto) {
		p->released_upto = p->win[wi].first;
		for (Slot &t : p->slots)
			if (t.state == READY && t.window != (size_t)-1 && t.window < wi)
				t.state = EMPTY;
		p->cv_work.notify_all();
	}
	Slot &s = p->slots[wi % p->slots.size()];
	for (;;) {
		if (p->error != B2I_OK)
			return p->error;
		if (s.window == wi && s.state == READY)
			break;
		if (fill_one(p, lk))
			continue;
		p->cv_ready.wait(lk);
	}
	if (idx == p->win[wi].first)
		p->note("handed out", wi);
	/* keep the ring full while the caller chews on this window */
	while (fill_one(p, lk))
		;
	if (p->error != B2I_OK)
		return p->error;
	const Window &w = p->win[wi];
	const size_t j = idx - w.first;
	const b2i_stream_desc &d = s.descs[j];
	if (res) *res = s.res[j];
	if (out_data)
		*out_data = (d.method == B2I_METHOD_STORED && (d.flags & B2I_F_NO_COPY)) ? NULL : s.h_out + d.out_off;
	if (in_data)
		*in_data = (d.method == B2I_METHOD_DEFLATE || d.method == B2I_METHOD_STORED) ? s.in_base + d.in_off : NULL;
	return B2I_OK;
}

extern "C" void b2i_pipe_release(b2i_pipe *p, size_t idx)
{
	if (p == NULL)
		return;
	std::lock_guard<std::mutex> g(p->mu);
	if (idx <= p->released_upto)
		return;
	p->released_upto = idx;
	for (Slot &s : p->slots) {
		if (s.state != READY || s.window == (size_t)-1)
			continue;
		const Window &w = p->win[s.window];
		if (w.first + w.count <= idx)
			s.state = EMPTY;
	}
	p->cv_work.notify_all();
}

extern "C" const char *b2i_pipe_error(const b2i_pipe *p) { return p ? p->err : "no pipe"; }

extern "C" size_t b2i_pipe_window_count(const b2i_pipe *p) { return p ? p->win.size() : 0; }

extern "C" void b2i_pipe_close(b2i_pipe *p)
{
	if (p == NULL)
		return;
	{
		std::lock_guard<std::mutex> g(p->mu);
		p->closing = true;
	}
	p->cv_work.notify_all();
	for (auto &t : p->workers)
		t.join();
	for (Slot &s : p->slots) {
		g_pinned.put(s.h_in, s.h_in_cap);
		g_pinned.put(s.h_out, s.h_out_cap);
	}
	delete p;
}

/* ---- one call over a whole batch on several GPUs ---------------------------------- */
extern "C" int b2i_decode_host_multi(b2i_ctx *const *ctxs, int nctx, const void *host_in, size_t in_bytes,
    const b2i_stream_desc *descs, size_t n, void *host_out, size_t out_bytes, b2i_stream_result *res)
{
	if (ctxs == NULL || nctx < 1 || (n && (host_in == NULL || descs == NULL || res == NULL)))
		return B2I_E_INVAL;
	if (n == 0)
		return B2I_OK;
	/* LPT when a few streams dominate, contiguous ranges otherwise: per GPU one list of
	 * streams, one thread, one pipelined b2i_decode_host */
	uint64_t total = 0, biggest = 0;
	for (size_t i = 0; i < n; i++) {
		uint64_t w = descs[i].in_len + descs[i].out_cap;
		total += w;
		biggest = std::max(biggest, w);
	}
	std::vector<uint32_t> owner(n);
	if (biggest * (uint64_t)nctx * 8 > total) {
		int rc = b2i_partition_lpt(descs, n, nctx, owner.data(), NULL);
		if (rc != B2I_OK)
			return rc;
	} else {
		std::vector<size_t> cuts((size_t)nctx + 1);
		int rc = b2i_partition_contiguous(descs, n, nctx, cuts.data());
		if (rc != B2I_OK)
			return rc;
		for (int g = 0; g < nctx; g++)
			for (size_t i = cuts[(size_t)g]; i < cuts[(size_t)g + 1]; i++)
				owner[i] = (uint32_t)g;
	}
	std::vector<int> rcs((size_t)nctx, B2I_OK);
	std::vector<std::thread> th;
	for (int g = 0; g < nctx; g++) {
		th.emplace_back([&, g] {
			std::vector<b2i_stream_desc> sub;
			std::vector<size_t> idx;
			for (size_t i = 0; i < n; i++)
				if (owner[i] == (uint32_t)g) {
					sub.push_back(descs[i]);
					idx.push_back(i);
				}
			if (sub.empty())
				return;
			std::vector<b2i_stream_result> r(sub.size());
			rcs[(size_t)g] = b2i_decode_host(ctxs[g], host_in, in_bytes, sub.data(), sub.size(), host_out,
			    out_bytes, r.data());
			if (rcs[(size_t)g] == B2I_OK)
				for (size_t k = 0; k < idx.size(); k++)
					res[idx[k]] = r[k];
		});
	}
	for (auto &t : th)
		t.join();
	for (int g = 0; g < nctx; g++)
		if (rcs[(size_t)g] != B2I_OK)
			return rcs[(size_t)g];
	return B2I_OK;
}
