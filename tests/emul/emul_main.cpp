/*
 * emul_main.cpp — TEST INFRASTRUCTURE ONLY: runs the device per-stream code of
 * libarchive_b200/csrc on the host under warp_emul.h (32 pthreads = 1 warp).
 */
#define B2I_HOST_EMUL 1
#include "../../libarchive_b200/csrc/stream_core.cuh"
#include <stdlib.h>
#include <vector>

thread_local unsigned tl_lane;
thread_local EmulWarp *tl_warp;
thread_local pthread_barrier_t *tl_team;

static uint32_t g_crc_tab[1024];
static uint32_t g_xp8[40];
static int g_tab_ready;

static void make_tables()
{
	for (uint32_t b = 0; b < 256; b++) {
		uint32_t c = b;
		for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
		g_crc_tab[b] = c;
	}
	for (int k = 1; k < 4; k++)
		for (uint32_t b = 0; b < 256; b++) {
			uint32_t c = g_crc_tab[(k - 1) * 256 + b];
			g_crc_tab[k * 256 + b] = g_crc_tab[c & 0xff] ^ (c >> 8);
		}
	uint32_t p = 0x00800000u; /* x^8 */
	for (int k = 0; k < 40; k++) { g_xp8[k] = p; p = crc_mulmod(p, p); }
	g_tab_ready = 1;
}

struct Job {
	EmulWarp warp;
	WarpSmem sm;
	const uint8_t *in; uint64_t in_total; uint8_t *out;
	B2iDesc d; B2iResult res;
	int mode; /* 0 inflate+crc, 1 crc of in[in_off, +in_len) */
	uint32_t *scratch;
	uint32_t crc_out;
};
struct LaneArg { Job *job; unsigned lane; };

static void *lane_main(void *p)
{
	LaneArg *a = (LaneArg *)p;
	Job *j = a->job;
	tl_lane = a->lane;
	tl_warp = &j->warp;
	if (j->mode == 0) {
		Ring ring;
		ring_init(&j->sm, ring);
		process_deflate_stream(&j->sm, ring, j->scratch, nullptr, j->in, j->in_total, j->out, nullptr, j->d, &j->res, g_crc_tab, g_xp8);
	} else {
		crc_load_tables(j->sm.lit, g_crc_tab);
		uint32_t raw0 = crc_warp_raw0(j->in + j->d.in_off, j->d.in_len, j->sm.lit, g_xp8);
		uint32_t c = crc_finish(j->d.expect_crc, raw0, j->d.in_len, g_xp8);
		if (tl_lane == 0) j->crc_out = c;
	}
	return NULL;
}

static void run(Job *j)
{
	pthread_t th[32];
	LaneArg args[32];
	pthread_barrier_init(&j->warp.bar, NULL, 32);
	for (unsigned i = 0; i < 32; i++) { args[i].job = j; args[i].lane = i; pthread_create(&th[i], NULL, lane_main, &args[i]); }
	for (unsigned i = 0; i < 32; i++) pthread_join(th[i], NULL);
	pthread_barrier_destroy(&j->warp.bar);
}

long g_lp_rounds, g_lp_passes;
long g_stat_hist[64], g_stat_instage, g_stat_overlap, g_stat_batches, g_stat_bytes, g_stat_syms;
extern "C" void emul_stats(long *h, long *o) { for (int i=0;i<64;i++) h[i]=g_stat_hist[i]; o[0]=g_stat_instage;o[1]=g_stat_overlap;o[2]=g_stat_batches;o[3]=g_stat_bytes;o[4]=g_stat_syms; }
extern "C" void emul_lp_stats(long *r, long *p) { *r = g_lp_rounds; *p = g_lp_passes; }
static int g_use_lp = 1;
extern "C" void emul_set_lane_parallel(int on) { g_use_lp = on; }

extern "C" int emul_inflate(const uint8_t *in, uint64_t in_total, uint8_t *out, const B2iDesc *d, B2iResult *res)
{
	if (!g_tab_ready) make_tables();
	Job *j = new Job();
	j->in = in; j->in_total = in_total; j->out = out; j->d = *d; j->mode = 0;
	j->scratch = g_use_lp ? (uint32_t *)malloc(LP_SCRATCH_WORDS * 4) : NULL;
	run(j);
	*res = j->res;
	free(j->scratch);
	delete j;
	return 0;
}

extern "C" uint32_t emul_crc32(uint32_t crc, const uint8_t *buf, uint64_t off, uint64_t len)
{
	if (!g_tab_ready) make_tables();
	Job *j = new Job();
	j->in = buf; j->d.in_off = off; j->d.in_len = len; j->d.expect_crc = crc; j->mode = 1;
	run(j);
	uint32_t c = j->crc_out;
	delete j;
	return c;
}


/* ---- a team of TEAM_WARPS warps on one stream (inflate_team.cuh) ---------------- */
struct TeamJob {
	EmulWarp warp[TEAM_WARPS];
	WarpSmem sm[1];
	TeamShared ts;
	pthread_barrier_t bar;
	uint32_t *scratch[TEAM_WARPS];
	const uint8_t *in; uint64_t in_total; uint8_t *out;
	B2iDesc d; B2iResult res;
};
struct TeamLaneArg { TeamJob *job; unsigned w, lane; };

static void *team_lane_main(void *p)
{
	TeamLaneArg *a = (TeamLaneArg *)p;
	TeamJob *j = a->job;
	tl_lane = a->lane;
	tl_warp = &j->warp[a->w];
	tl_team = &j->bar;
	if (a->lane == 0) {
		j->ts.scratch[a->w] = j->scratch[a->w];
		if (a->w == 0)
			team_init(&j->ts);
	}
	team_sync();
	if (a->w == 0) {
		Ring ring;
		ring_init(&j->sm[0], ring);
		process_deflate_stream(&j->sm[0], ring, j->scratch[0], &j->ts, j->in, j->in_total, j->out, nullptr,
		    j->d, &j->res, g_crc_tab, g_xp8);
		team_command(&j->ts, TC_QUIT);
	} else {
		team_serve(&j->ts, a->w, &j->sm[0]);
	}
	return NULL;
}

extern "C" int emul_inflate_team(const uint8_t *in, uint64_t in_total, uint8_t *out, const B2iDesc *d, B2iResult *res)
{
	if (!g_tab_ready) make_tables();
	TeamJob *j = new TeamJob();
	pthread_t th[TEAM_LANES];
	TeamLaneArg args[TEAM_LANES];
	j->in = in; j->in_total = in_total; j->out = out; j->d = *d;
	pthread_barrier_init(&j->bar, NULL, TEAM_LANES);
	for (unsigned w = 0; w < TEAM_WARPS; w++) {
		pthread_barrier_init(&j->warp[w].bar, NULL, 32);
		j->scratch[w] = (uint32_t *)malloc(LP_SCRATCH_WORDS * 4);
	}
	for (unsigned i = 0; i < TEAM_LANES; i++) {
		args[i].job = j; args[i].w = i / 32; args[i].lane = i % 32;
		pthread_create(&th[i], NULL, team_lane_main, &args[i]);
	}
	for (unsigned i = 0; i < TEAM_LANES; i++) pthread_join(th[i], NULL);
	*res = j->res;
	for (unsigned w = 0; w < TEAM_WARPS; w++) free(j->scratch[w]);
	delete j;
	return 0;
}
