/*
 * stream_core.cuh — what one warp does with one descriptor: inflate (method
 * 8) and the reference's end-of-entry checks, CRC included.
 *
 * Mirrors the epilogue of archive_read_format_zip_read_data
 * (archive_read_support_format_zip.c:3154-3195): running CRC over everything
 * delivered, then CRC / compressed-size / uncompressed-size (low 32 bits)
 * comparisons — recorded as flags for the host to turn into the reference's
 * messages at the moment the entry's last block is requested.
 */
#pragma once
#include "inflate_core.cuh"

B2I_DEV void process_deflate_stream(WarpSmem *sm, Ring &ring, uint32_t *scratch, TeamShared *team, const uint8_t *in,
    uint64_t in_total, uint8_t *out, uint8_t *out_mirror, const B2iDesc &d, B2iResult *res,
    const uint32_t *crc_tab_g, const uint32_t *xp8)
{
	StreamOut so = inflate_stream(sm, ring, scratch, team, in, in_total, d.in_off, d.in_len,
	    out + d.out_off, out_mirror ? out_mirror + d.out_off : nullptr, d.out_cap);
	uint32_t crc = 0;

	PH_DECL();
	if (so.status == S_OK && !(d.flags & F_NO_CRC)) {
		uint32_t raw0;
		if (team && so.out_bytes >= 65536u) {
			/* a large stream: every warp of the team takes a slice */
			__syncwarp();
			fence_block();
			raw0 = team_crc_raw0(team, out + d.out_off, so.out_bytes, crc_tab_g, xp8);
		} else {
			/* the lit/len table is dead now: its space holds the slice tables */
			crc_load_tables(sm->lit, crc_tab_g);
			raw0 = crc_warp_raw0(out + d.out_off, so.out_bytes, sm->lit, xp8);
		}
		crc = crc_finish(0, raw0, so.out_bytes, xp8);
		__syncwarp();
	}
	PH_ADD(PH_CRC);
	if (b2i_lane() == 0) {
		uint32_t flags = 0;
		if (so.status == S_OK) {
			if (!(d.flags & F_NO_CRC) && crc != d.expect_crc)
				flags |= R_CRC_MISMATCH;
			if (so.in_bytes != d.in_len)
				flags |= R_IN_MISMATCH;
			if ((so.out_bytes & 0xffffffffull) != (d.expect_out & 0xffffffffull))
				flags |= R_OUT_MISMATCH;
		}
		res->status = so.status;
		res->crc = crc;
		res->out_bytes = so.out_bytes;
		res->in_bytes = so.in_bytes;
		res->detail = so.detail;
		res->flags = flags;
	}
}
