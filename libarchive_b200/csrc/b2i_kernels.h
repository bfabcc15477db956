/* b2i_kernels.h — launch wrappers of b2i_kernels.cu, used by b2i_api.cpp */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b2i_common.cuh"

/* one 16 KiB (or shorter, last) piece of a stored entry */
struct B2iCrcWork {
	uint64_t rel;     /* offset of the piece inside the entry */
	uint32_t len;
	uint32_t entry;   /* descriptor index */
};
struct B2iCrcEntry {
	uint32_t entry;       /* descriptor index */
	uint32_t first_work;
	uint32_t nwork;
	uint32_t pad;
};
#define B2I_CRC_CHUNK (32u * 1024u)   /* streaming pieces: 16-byte aligned, multiples of 512 bytes */

/* per-device function attributes (dynamic shared memory sizes, carveout) */
cudaError_t b2i_kernels_configure(void);
size_t b2i_inflate_smem_bytes(void);
size_t b2i_inflate_scratch_bytes(int num_sms);
uint32_t b2i_inflate_scratch_slots(int num_sms);
cudaError_t b2i_launch_tables(uint32_t *crc_tab, uint32_t *xp8, uint32_t *ztab, uint32_t *lane_mul,
    cudaStream_t st);
cudaError_t b2i_launch_inflate(const uint8_t *in, uint64_t in_total, uint8_t *out, uint8_t *out_mirror,
    const B2iDesc *descs, B2iResult *results, const uint32_t *order, uint32_t n,
    unsigned int *counter, const uint32_t *crc_tab, const uint32_t *xp8, uint32_t *scratch,
    unsigned int *slot_busy, int num_sms, cudaStream_t st);
/* the 9-bit-root / 8-CTA build of the same kernel (launches with several waves of streams) */
cudaError_t b2i_launch_inflate_r9(const uint8_t *in, uint64_t in_total, uint8_t *out, uint8_t *out_mirror,
    const B2iDesc *descs, B2iResult *results, const uint32_t *order, uint32_t n,
    unsigned int *counter, const uint32_t *crc_tab, const uint32_t *xp8, uint32_t *scratch,
    unsigned int *slot_busy, int num_sms, cudaStream_t st);
cudaError_t b2i_launch_inflate_team(const uint8_t *in, uint64_t in_total, uint8_t *out, uint8_t *out_mirror,
    const B2iDesc *descs, B2iResult *results, const uint32_t *order, uint32_t n,
    unsigned int *counter, const uint32_t *crc_tab, const uint32_t *xp8, uint32_t *scratch,
    unsigned int *slot_busy, int num_sms, cudaStream_t st);
cudaError_t b2i_launch_crc_chunks(const uint8_t *in, uint8_t *out, const B2iDesc *descs,
    const B2iCrcWork *work, uint32_t nwork, uint32_t *partial, const uint32_t *crc_tab,
    const uint32_t *xp8, const uint32_t *ztab, const uint32_t *lane_mul, int num_sms, cudaStream_t st);
cudaError_t b2i_launch_crc_combine(const B2iDesc *descs, B2iResult *results,
    const B2iCrcEntry *ents, uint32_t nents, const B2iCrcWork *work, const uint32_t *partial,
    const uint32_t *xp8, cudaStream_t st);
cudaError_t b2i_launch_unsupported(const B2iDesc *descs, B2iResult *results,
    const uint32_t *list, uint32_t n, cudaStream_t st);
