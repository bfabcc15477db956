/*
 * zip_b200_local.h — what the two ZIP reader modules of this directory share:
 * options, the metadata carried by "extra" fields, pathname / mode fix-ups and
 * the end-of-entry verdict.  Host-side header logic only; entry BODIES are
 * decoded on the device (include/b200inflate.h).
 *
 * Behaviour follows the reference reader (archive_read_support_format_zip.c);
 * the lines each piece mirrors are cited at the piece.
 */
#ifndef ZIP_B200_LOCAL_H
#define ZIP_B200_LOCAL_H

#define ZIP_ENCRYPTED            (1 << 0)
#define ZIP_LENGTH_AT_END        (1 << 3)
#define ZIP_STRONG_ENCRYPTED     (1 << 6)
#define ZIP_UTF8_NAME            (1 << 11)
#define ZIP_CD_ENCRYPTED         (1 << 13)
#define ZIP_BLOCK                (256 * 1024)      /* zip.c:2550 */

#define ZB_UNUSED __attribute__((unused))

/* options and conversion state both readers keep (zip.c:3271-3318, 930-1003) */
struct zb_common {
	b2i_ctx   *ctx;
	int        ctx_bad;               /* a device call failed: do not hand the context back */
	int        ignore_crc32;
	int        has_encrypted_entries;
	int        init_default_conversion;
	struct archive_string_conv *sconv, *sconv_default, *sconv_utf8;
	struct archive_string format_name;
};

/* what the local header and the extra fields say about one entry */
struct zb_meta {
	uint64_t compressed_size, uncompressed_size;
	uint32_t crc32;
	uint16_t zip_flags, method;
	uint8_t  version, system, used_zip64;
	uint32_t mode;
	int64_t  mtime, atime, ctime;
	int64_t  uid, gid;
};

static inline uint16_t zb_le16(const unsigned char *p) { return (uint16_t)(p[0] | p[1] << 8); }
static inline uint32_t zb_le32(const unsigned char *p)
{
	return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}
static inline uint64_t zb_le64(const unsigned char *p) { return (uint64_t)zb_le32(p) | (uint64_t)zb_le32(p + 4) << 32; }

ZB_UNUSED static const char *
zb_compression_name(int m)          /* the names zip.c:362-403 prints */
{
	static const struct { int id; const char *name; } t[] = {
		{ 0, "uncompressed" }, { 1, "shrinking" }, { 2, "reduced-1" }, { 3, "reduced-2" },
		{ 4, "reduced-3" }, { 5, "reduced-4" }, { 6, "imploded" }, { 7, "reserved" },
		{ 8, "deflation" }, { 9, "deflation-64-bit" }, { 10, "ibm-terse" }, { 11, "reserved" },
		{ 12, "bzip" }, { 13, "reserved" }, { 14, "lzma" }, { 15, "reserved" }, { 16, "reserved" },
		{ 17, "reserved" }, { 18, "ibm-terse-new" }, { 19, "ibm-lz777" }, { 93, "zstd" },
		{ 95, "xz" }, { 96, "jpeg" }, { 97, "wav-pack" }, { 98, "ppmd-1" }, { 99, "aes" }
	};
	for (size_t i = 0; i < sizeof(t) / sizeof(t[0]); i++)
		if (t[i].id == m)
			return (t[i].name);
	return ("??");
}

ZB_UNUSED static int
zb_options(struct archive_read *a, struct zb_common *c, const char *key, const char *val)
{                                                              /* zip.c:3271-3318 */
	int ret = ARCHIVE_FAILED;

	if (strcmp(key, "compat-2x") == 0) {
		c->init_default_conversion = (val != NULL) ? 1 : 0;
		return (ARCHIVE_OK);
	} else if (strcmp(key, "hdrcharset") == 0) {
		if (val == NULL || val[0] == 0)
			archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
			    "zip: hdrcharset option needs a character-set name");
		else {
			c->sconv = archive_string_conversion_from_charset(&a->archive, val, 0);
			if (c->sconv != NULL) {
				if (strcmp(val, "UTF-8") == 0)
					c->sconv_utf8 = c->sconv;
				ret = ARCHIVE_OK;
			} else
				ret = ARCHIVE_FATAL;
		}
		return (ret);
	} else if (strcmp(key, "ignorecrc32") == 0) {
		c->ignore_crc32 = !(val == NULL || val[0] == 0);
		return (ARCHIVE_OK);
	} else if (strcmp(key, "mac-ext") == 0) {
		return (ARCHIVE_OK);             /* resource-fork folding: not provided (off by default off macOS) */
	}
	return (ARCHIVE_WARN);
}

/* CRC-32 of a file NAME for the 0x7075 check below (a few bytes of header
 * metadata; entry data never goes through this) */
ZB_UNUSED static uint32_t
zb_name_crc(const char *s, size_t n)
{
	uint32_t c = 0xffffffffu;
	for (size_t i = 0; i < n; i++) {
		c ^= (unsigned char)s[i];
		for (int k = 0; k < 8; k++)
			c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
	}
	return (~c);
}

/*
 * Extra fields (zip.c:474-900): ZIP64 sizes (0x0001, in the reference's order and
 * only for fields that read 0xffffffff), times (0x5455 "UT", 0x5855 "UX"), owner
 * (0x5855, 0x7855 "Ux", 0x7875 "ux"), the experimental 0x6c78 attributes and the
 * Unicode path 0x7075 (only with an entry to rename).  `lho` may be NULL.
 */
ZB_UNUSED static int
zb_process_extra(struct archive_read *a, struct zb_common *c, struct archive_entry *entry,
    const unsigned char *p, size_t n, struct zb_meta *m, uint64_t *lho)
{
	size_t off = 0;

	if (n == 0)
		return (ARCHIVE_OK);
	if (n < 4) {
		for (size_t i = 0; i < n; i++)
			if (p[i] != 0) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
				    "Too-small extra data: Need at least 4 bytes, but only found %d bytes", (int)n);
				return (ARCHIVE_FAILED);
			}
		return (ARCHIVE_OK);
	}
	while (off <= n - 4) {
		unsigned id = zb_le16(p + off), sz = zb_le16(p + off + 2);
		size_t o;

		off += 4;
		if (off + sz > n) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
			    "Extra data overflow: Need %d bytes but only found %d bytes", (int)sz, (int)(n - off));
			return (ARCHIVE_FAILED);
		}
		o = off;
		switch (id) {
		case 0x0001: {
			unsigned left = sz;
			const char *what = NULL;
			m->used_zip64 = 1;
			if (m->uncompressed_size == 0xffffffffull) {
				uint64_t t = 0;
				if (left < 8 || (t = zb_le64(p + o)) > INT64_MAX)
					what = "uncompressed size";
				else { m->uncompressed_size = t; o += 8; left -= 8; }
			}
			if (what == NULL && m->compressed_size == 0xffffffffull) {
				uint64_t t = 0;
				if (left < 8 || (t = zb_le64(p + o)) > INT64_MAX)
					what = "compressed size";
				else { m->compressed_size = t; o += 8; left -= 8; }
			}
			if (what == NULL && lho != NULL && *lho == 0xffffffffull) {
				uint64_t t = 0;
				if (left < 8 || (t = zb_le64(p + o)) > INT64_MAX)
					what = "header offset";
				else { *lho = t; o += 8; left -= 8; }
			}
			if (what != NULL) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Malformed 64-bit %s", what);
				return (ARCHIVE_FAILED);
			}
			break;
		}
		case 0x5455: {
			unsigned left = sz;
			int flags;
			if (left == 0) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Incomplete extended time field");
				return (ARCHIVE_FAILED);
			}
			flags = p[o++];
			left--;
			if (flags & 1) {
				if (left < 4) break;
				m->mtime = zb_le32(p + o); o += 4; left -= 4;
			}
			if (flags & 2) {
				if (left < 4) break;
				m->atime = zb_le32(p + o); o += 4; left -= 4;
			}
			if (flags & 4) {
				if (left < 4) break;
				m->ctime = zb_le32(p + o); o += 4; left -= 4;
			}
			break;
		}
		case 0x5855:
			if (sz >= 8) {
				m->atime = zb_le32(p + o);
				m->mtime = zb_le32(p + o + 4);
			}
			if (sz >= 12) {
				m->uid = zb_le16(p + o + 8);
				m->gid = zb_le16(p + o + 10);
			}
			break;
		case 0x6c78: {
			unsigned left = sz;
			int bitmap, last;
			if (left < 1) break;
			last = bitmap = p[o++];
			left--;
			while ((last & 0x80) != 0 && left >= 1) { last = p[o++]; left--; }
			if (bitmap & 1) {
				if (left < 2) break;
				m->system = (uint8_t)(zb_le16(p + o) >> 8); o += 2; left -= 2;
			}
			if (bitmap & 2) {
				if (left < 2) break;
				o += 2; left -= 2;
			}
			if (bitmap & 4) {
				uint32_t ext;
				if (left < 4) break;
				ext = zb_le32(p + o);
				if (m->system == 3)
					m->mode = ext >> 16;
				else if (m->system == 0) {
					m->mode = (ext & 0x10) ? (AE_IFDIR | 0775) : (AE_IFREG | 0664);
					if (ext & 0x01)
						m->mode &= 0555;
				} else
					m->mode = 0;
				o += 4; left -= 4;
			}
			break;
		}
		case 0x7075: {
			if (sz < 5 || entry == NULL)
				break;
			if (c->sconv_utf8 == NULL) {
				c->sconv_utf8 = archive_string_conversion_from_charset(&a->archive, "UTF-8", 1);
				if (c->sconv_utf8 == NULL)
					break;
			}
			if (!c->ignore_crc32) {
				const char *cp = archive_entry_pathname(entry);
				if (cp != NULL && zb_name_crc(cp, strlen(cp)) != zb_le32(p + o + 1))
					break;
			}
			(void)archive_entry_copy_pathname_l(entry, (const char *)p + o + 5, sz - 5, c->sconv_utf8);
			break;
		}
		case 0x7855:
			if (sz >= 2) m->uid = zb_le16(p + o);
			if (sz >= 4) m->gid = zb_le16(p + o + 2);
			break;
		case 0x7875: {
			int us = 0, gs;
			if (sz >= 1 && p[o] == 1) {
				if (sz >= 4) {
					us = p[o + 1];
					if (us == 2) m->uid = zb_le16(p + o + 2);
					else if (us == 4 && sz >= 6) m->uid = zb_le32(p + o + 2);
				}
				if (sz >= (unsigned)(2 + us + 3)) {
					gs = p[o + 2 + us];
					if (gs == 2) m->gid = zb_le16(p + o + 2 + us + 1);
					else if (gs == 4 && sz >= (unsigned)(2 + us + 5)) m->gid = zb_le32(p + o + 2 + us + 1);
				}
			}
			break;
		}
		case 0x9901:
			if (sz < 6) {
				archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT, "Incomplete AES field");
				return (ARCHIVE_FAILED);
			}
			break;
		default:
			break;
		}
		off += sz;
	}
	return (ARCHIVE_OK);
}

/* pathname with the reference's choice of conversion (zip.c:976-1003); returns
 * ARCHIVE_OK / WARN / FATAL */
ZB_UNUSED static int
zb_set_pathname(struct archive_read *a, struct zb_common *c, struct archive_entry *entry,
    const char *name, size_t len, unsigned zip_flags)
{
	struct archive_string_conv *sconv;

	if (c->sconv == NULL && !c->init_default_conversion) {
		c->sconv_default = archive_string_default_conversion_for_read(&a->archive);
		c->init_default_conversion = 1;
	}
	if (zip_flags & ZIP_UTF8_NAME) {
		if (c->sconv_utf8 == NULL) {
			c->sconv_utf8 = archive_string_conversion_from_charset(&a->archive, "UTF-8", 1);
			if (c->sconv_utf8 == NULL)
				return (ARCHIVE_FATAL);
		}
		sconv = c->sconv_utf8;
	} else if (c->sconv != NULL)
		sconv = c->sconv;
	else
		sconv = c->sconv_default;
	if (archive_entry_copy_pathname_l(entry, name, len, sconv) != 0) {
		if (errno == ENOMEM) {
			archive_set_error(&a->archive, ENOMEM, "Can't allocate memory for Pathname");
			return (ARCHIVE_FATAL);
		}
		archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
		    "Pathname cannot be converted from %s to current locale.",
		    archive_string_conversion_charset_name(sconv));
		return (ARCHIVE_WARN);
	}
	return (ARCHIVE_OK);
}

/* mode defaults, backslash separators, directories and their trailing slash
 * (zip.c:1031-1103) */
ZB_UNUSED static void
zb_fix_path_and_mode(struct archive_entry *entry, struct zb_meta *m)
{
	const wchar_t *wp;
	const char *cp;
	size_t len;

	if ((m->mode & AE_IFMT) == AE_IFIFO)
		m->mode = (m->mode & ~AE_IFMT) | AE_IFREG;
	if (m->mode == 0)
		m->mode |= 0664;
	if (m->system == 0 && (wp = archive_entry_pathname_w(entry)) != NULL &&
	    wcschr(wp, L'/') == NULL && wcschr(wp, L'\\') != NULL) {
		size_t k;
		wchar_t *w;
		len = wcslen(wp);
		if ((w = malloc((len + 1) * sizeof(*w))) != NULL) {
			for (k = 0; k <= len; k++)
				w[k] = wp[k] == L'\\' ? L'/' : wp[k];
			archive_entry_copy_pathname_w(entry, w);
			free(w);
		}
	}
	if ((m->mode & AE_IFMT) != AE_IFDIR) {
		int has_slash;
		if ((wp = archive_entry_pathname_w(entry)) != NULL) {
			len = wcslen(wp);
			has_slash = len > 0 && wp[len - 1] == L'/';
		} else {
			cp = archive_entry_pathname(entry);
			len = cp != NULL ? strlen(cp) : 0;
			has_slash = len > 0 && cp[len - 1] == '/';
		}
		if (has_slash)
			m->mode = (m->mode & ~AE_IFMT) | AE_IFDIR | 0111;
		else if ((m->mode & AE_IFMT) == 0)
			m->mode |= AE_IFREG;
	}
	if ((m->mode & AE_IFMT) == AE_IFDIR) {
		if ((wp = archive_entry_pathname_w(entry)) != NULL) {
			len = wcslen(wp);
			if (len > 0 && wp[len - 1] != L'/') {
				wchar_t *w = malloc((len + 2) * sizeof(*w));
				if (w != NULL) {
					wmemcpy(w, wp, len);
					w[len] = L'/';
					w[len + 1] = 0;
					archive_entry_copy_pathname_w(entry, w);
					free(w);
				}
			}
		} else {
			cp = archive_entry_pathname(entry);
			len = cp != NULL ? strlen(cp) : 0;
			if (len > 0 && cp[len - 1] != '/') {
				char *s = malloc(len + 2);
				if (s != NULL) {
					memcpy(s, cp, len);
					s[len] = '/';
					s[len + 1] = 0;
					archive_entry_set_pathname(entry, s);
					free(s);
				}
			}
		}
	}
}

ZB_UNUSED static void
zb_populate(struct archive_entry *entry, const struct zb_meta *m)   /* zip.c:1152-1158 */
{
	archive_entry_set_mode(entry, m->mode);
	archive_entry_set_uid(entry, m->uid);
	archive_entry_set_gid(entry, m->gid);
	archive_entry_set_mtime(entry, m->mtime, 0);
	archive_entry_set_ctime(entry, m->ctime, 0);
	archive_entry_set_atime(entry, m->atime, 0);
}

/* symlink target with the reference's conversion rules (zip.c:1226-1257) */
ZB_UNUSED static int
zb_set_symlink(struct archive_read *a, struct zb_common *c, struct archive_entry *entry,
    const unsigned char *p, size_t len, unsigned zip_flags)
{
	struct archive_string_conv *sconv = c->sconv;

	if (sconv == NULL && (zip_flags & ZIP_UTF8_NAME))
		sconv = c->sconv_utf8;
	if (sconv == NULL)
		sconv = c->sconv_default;
	if (archive_entry_copy_symlink_l(entry, (const char *)p, len, sconv) != 0) {
		if (errno != ENOMEM && sconv == c->sconv_utf8 && (zip_flags & ZIP_UTF8_NAME))
			archive_entry_copy_symlink_l(entry, (const char *)p, len, NULL);
		if (errno == ENOMEM) {
			archive_set_error(&a->archive, ENOMEM, "Can't allocate memory for Symlink");
			return (ARCHIVE_FATAL);
		}
		if (sconv != c->sconv_utf8 || (zip_flags & ZIP_UTF8_NAME) == 0) {
			archive_set_error(&a->archive, ARCHIVE_ERRNO_FILE_FORMAT,
			    "Symlink cannot be converted from %s to current locale.",
			    archive_string_conversion_charset_name(sconv));
			return (ARCHIVE_WARN);
		}
	}
	return (ARCHIVE_OK);
}

/* the three end-of-entry checks, in the reference's order (zip.c:3164-3194) */
ZB_UNUSED static int
zb_end_of_entry_checks(struct archive_read *a, const struct zb_common *c, uint32_t computed_crc,
    uint32_t expect_crc, int64_t cread, int64_t csize, int64_t uread, int64_t usize)
{
	if (expect_crc != computed_crc && !c->ignore_crc32) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC, "ZIP bad CRC: 0x%lx should be 0x%lx",
		    (unsigned long)computed_crc, (unsigned long)expect_crc);
		return (ARCHIVE_FAILED);
	}
	if (csize != cread) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
		    "ZIP compressed data is wrong size (read %jd, expected %jd)", (intmax_t)cread, (intmax_t)csize);
		return (ARCHIVE_FAILED);
	}
	if ((usize & 0xffffffffll) != (uread & 0xffffffffll)) {
		archive_set_error(&a->archive, ARCHIVE_ERRNO_MISC,
		    "ZIP uncompressed data is wrong size (read %jd, expected %jd)\n", (intmax_t)uread, (intmax_t)usize);
		return (ARCHIVE_FAILED);
	}
	return (ARCHIVE_OK);
}

#endif
