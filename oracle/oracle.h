/*
 * oracle.h — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * inflate + CRC-32 hot path (see oracle_inflate.c / oracle_crc32.c headers
 * for the reference file:line each function follows and how it is pinned).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use anything under oracle/.
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define ORC_OK            0
#define ORC_DATA_ERROR   -3
#define ORC_BUF_ERROR    -5
#define ORC_OUT_OVERFLOW -100

#define ORC_D_BAD_BLOCK_TYPE     1
#define ORC_D_BAD_STORED_LEN     2
#define ORC_D_TOO_MANY_SYMS      3
#define ORC_D_BAD_CODELEN_SET    4
#define ORC_D_BAD_BITLEN_REPEAT  5
#define ORC_D_NO_EOB             6
#define ORC_D_BAD_LITLEN_SET     7
#define ORC_D_BAD_DIST_SET       8
#define ORC_D_BAD_LITLEN_CODE    9
#define ORC_D_BAD_DIST_CODE     10
#define ORC_D_DIST_TOO_FAR      11

typedef struct orc_result {
	int32_t  status;
	int32_t  detail;
	uint64_t out_bytes;
	uint64_t in_bytes;
} orc_result;

int      orc_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap, orc_result *r);
uint32_t orc_crc32(uint32_t crc, const void *buf, size_t len);          /* archive_crc32.h:43-84 */
uint32_t orc_bitcrc32(uint32_t crc, const void *buf, size_t len);       /* test_utils/test_utils.c:113-139 */
uint32_t orc_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* batch helper for the Python tests and the CPU baseline: decode n streams
 * described like b2i_stream_desc (see include/b200inflate.h) */
typedef struct orc_desc {
	uint64_t in_off, in_len, out_off, out_cap, expect_out;
	uint32_t expect_crc;
	uint8_t  method, flags;
	uint16_t reserved;
} orc_desc;
typedef struct orc_stream_result {
	int32_t  status;
	uint32_t crc;
	uint64_t out_bytes, in_bytes;
	uint32_t detail, flags;
} orc_stream_result;
int orc_decode_batch(const uint8_t *in, size_t in_bytes, const orc_desc *d, size_t n,
    uint8_t *out, size_t out_bytes, orc_stream_result *res);
#endif
