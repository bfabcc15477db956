/*
 * b2i_common.cuh — shared definitions for the sm_100a kernels.
 *
 * The warp-level algorithms in inflate_core.cuh / crc32_core.cuh are written
 * against a tiny portability shim so that tests/emul can compile the very same
 * source for the host with 32 pthreads standing in for the 32 lanes of a warp
 * (B2I_HOST_EMUL; test infrastructure only — the product build never defines
 * it and there is no CPU path in the library).
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef B2I_HOST_EMUL
#include "warp_emul.h"
#else
#include <cuda_runtime.h>
#define B2I_DEV __device__ __forceinline__
#define B2I_DEV_NOINLINE static __device__ __noinline__
B2I_DEV unsigned b2i_lane() { unsigned l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
#endif

#define B2I_FULL 0xffffffffu

/* -DB2I_DEBUG_BOUNDS: every staging / table / scratch / window index is checked and
 * a violation traps the kernel (compute-sanitizer is not available on the pool;
 * tests/test_gpu_inflate.py can be run against such a build via B2I_LIB) */
#if defined(B2I_DEBUG_BOUNDS) && !defined(B2I_HOST_EMUL)
#define B2I_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#elif defined(B2I_HOST_EMUL)
#include <assert.h>
#define B2I_CHECK(cond) assert(cond)
#else
#define B2I_CHECK(cond) do { } while (0)
#endif

/* must match include/b200inflate.h (static_asserts in b2i_api.cu) */
struct B2iDesc {
	uint64_t in_off, in_len, out_off, out_cap, expect_out;
	uint32_t expect_crc;
	uint8_t  method, flags;
	uint16_t reserved;
};
struct B2iResult {
	int32_t  status;
	uint32_t crc;
	uint64_t out_bytes, in_bytes;
	uint32_t detail, flags;
};

#define S_OK            0
#define S_DATA_ERROR   -3
#define S_BUF_ERROR    -5
#define S_OUT_OVERFLOW -100
#define S_UNSUPPORTED  -101

#define D_BAD_BLOCK_TYPE     1
#define D_BAD_STORED_LEN     2
#define D_TOO_MANY_SYMS      3
#define D_BAD_CODELEN_SET    4
#define D_BAD_BITLEN_REPEAT  5
#define D_NO_EOB             6
#define D_BAD_LITLEN_SET     7
#define D_BAD_DIST_SET       8
#define D_BAD_LITLEN_CODE    9
#define D_BAD_DIST_CODE     10
#define D_DIST_TOO_FAR      11

#define F_NO_COPY 0x01
#define F_NO_CRC  0x02
#define R_CRC_MISMATCH 0x01
#define R_IN_MISMATCH  0x02
#define R_OUT_MISMATCH 0x04

#define CRC_POLY 0xEDB88320u

/* -DB2I_PHASE_CLOCKS (make prof): per-phase cycle counters, summed over all warps,
 * printed and cleared by b2i_ctx_sync.  Profiling builds only. */
#if defined(B2I_PHASE_CLOCKS) && !defined(B2I_HOST_EMUL)
static __device__ unsigned long long g_b2i_phase[16];
#define PH_DECL()      long long ph_t_ = clock64()
#define PH_ADD(slot)   do { long long n_ = clock64(); if (b2i_lane() == 0) atomicAdd(&g_b2i_phase[slot], (unsigned long long)(n_ - ph_t_)); ph_t_ = n_; } while (0)
#define PH_COUNT(slot, v) do { if (b2i_lane() == 0) atomicAdd(&g_b2i_phase[slot], (unsigned long long)(v)); } while (0)
#else
#define PH_DECL()      do { } while (0)
#define PH_ADD(slot)   do { } while (0)
#define PH_COUNT(slot, v) do { } while (0)
#endif
#define PH_HEADER 0
#define PH_LPDEC  1
#define PH_LPRES  2
#define PH_UNIF   3
#define PH_CRC    4
#define PH_BATCH  5   /* count of resolve batches in LP */
#define PH_STORED 6
#define PH_TPASS  8   /* team: decode passes */
#define PH_TEXP   9   /* team: expand */
#define PH_TSWEEP 10  /* team: pointer sweep */
#define PH_TFLUSH 11  /* team: flush */
#define PH_TPLAN  12  /* team: hist reload + chunk planning */
#define PH_TROUNDS 13
#define PH_TPASSES 14
