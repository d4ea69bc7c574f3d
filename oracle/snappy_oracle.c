/*
 * oracle/snappy_oracle.c -- CPU restatement of nim-snappy's framed / raw Snappy
 * hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file.  The product (nim-snappy_b200/) never
 * links, imports or calls it; the product fails loudly without its CUDA library.
 *
 * What it restates (paths relative to /root/reference, commit 8ab5cf0):
 *   snappy/encoder.nim:27-37   tableSize, hash
 *   snappy/encoder.nim:44-73   emitLiteral
 *   snappy/encoder.nim:81-125  emitCopy
 *   snappy/encoder.nim:130-182 findMatchLength
 *   snappy/encoder.nim:184-383 encodeBlock
 *   snappy/encoder.nim:385-426 encodeFrame
 *   snappy/decoder.nim:20-155  decodeAllTags
 *   snappy.nim:27-64, 84-110, 130-155, 169-267   compress / uncompress /
 *                               compressFramed / uncompressFramed
 *   snappy/codec.nim:92-120, 129-214  bounds, header-only length walks
 *   snappy/crc32c.c:204-214, 676-763   masked CRC-32C (restated bytewise here;
 *                               oracle/_ref builds the reference's own file to
 *                               check this restatement against)
 * Varints follow stew/leb128 (un-vendored nimble dependency, version unpinned
 * in snappy.nimble:10-14): little-endian base-128, at most 5 bytes for u32 /
 * 10 for u64, overflow and truncation rejected (pinned by
 * tests/test_snappy.nim:168-174).
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   decoder, CRC, framing walk: pinned by the reference's committed fixtures
 *     (the .sz files of tests/stream_compressed, Mark.Twain-Tom.Sawyer.txt.rawsnappy, the
 *     19 malformed vectors and hand-built streams of tests/test_snappy.nim).
 *   encoder: pinned byte-for-byte on incompressible input by house.jpg.sz and
 *     structurally by orc_encode_block_ex(hash_bits=15, exit_ge=1) ==
 *     bundled libsnappy on the whole corpus; the two nim-specific parameters
 *     (14-bit table encoder.nim:11, `ip > ipLimit` encoder.nim:362) are
 *     one-token changes.  No fixture produced by the Nim encoder itself
 *     exists for compressible data (no Nim toolchain here): that part of the
 *     encoder parity is UNPINNED by reference output.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_OK 0
#define ORC_BUFFER_TOO_SMALL 1 /* 1 + ord(CodecError.bufferTooSmall / FrameError.bufferTooSmall) */
#define ORC_INVALID_INPUT 2
#define ORC_CRC_MISMATCH 3
#define ORC_UNKNOWN_CHUNK 4

#define MAX_BLOCK 65536u
#define INPUT_MARGIN 15u
#define MIN_NON_LITERAL 17u

static const uint8_t kStreamHeader[10] = {0xff, 0x06, 0x00, 0x00, 0x73, 0x4e, 0x61, 0x50, 0x70, 0x59};

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* ------------------------------------------------------------------ CRC */
/* crc32c.c:747 calculate_crc32c -> :676-731 crc32c_sb8_64_bit (slicing by 8: eight table lookups per 8 input
 * bytes), :204-214 (bytewise table loop, used here for the tail), :759-763 (mask).  Tables are generated instead of
 * spelled out; table k maps a byte that is followed by k more bytes of zeros. */
static uint32_t crc_tab[8][256];
__attribute__((constructor)) static void crc_init(void) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0);
    crc_tab[0][i] = c;
  }
  for (int k = 1; k < 8; k++)
    for (uint32_t i = 0; i < 256; i++) crc_tab[k][i] = (crc_tab[k - 1][i] >> 8) ^ crc_tab[0][crc_tab[k - 1][i] & 0xff];
}
uint32_t orc_masked_crc32c(const uint8_t *p, size_t n) {
  uint32_t c = 0xffffffffu;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w = ld64(p + i) ^ c; /* little-endian host, as the reference's fast path assumes */
    c = crc_tab[7][w & 0xff] ^ crc_tab[6][(w >> 8) & 0xff] ^ crc_tab[5][(w >> 16) & 0xff] ^ crc_tab[4][(w >> 24) & 0xff] ^
        crc_tab[3][(w >> 32) & 0xff] ^ crc_tab[2][(w >> 40) & 0xff] ^ crc_tab[1][(w >> 48) & 0xff] ^ crc_tab[0][w >> 56];
  }
  for (; i < n; i++) c = crc_tab[0][(c ^ p[i]) & 0xff] ^ (c >> 8);
  c = ~c;
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

/* --------------------------------------------------------------- bounds */
/* codec.nim:92-120 */
uint64_t orc_max_compressed_len(uint64_t n) { return 32 + n + n / 6; }
/* codec.nim:140-164 */
uint64_t orc_max_compressed_len_framed(int64_t n) {
  if (n <= 0) return 10;
  uint64_t frames = ((uint64_t)n + MAX_BLOCK - 1) / MAX_BLOCK;
  return (frames - 1) * (MAX_BLOCK + 8) + orc_max_compressed_len(MAX_BLOCK) + 8 + 10;
}

/* ------------------------------------------------------------- varints */
static int put_varint32(uint8_t *o, uint32_t v) {
  int n = 0;
  while (v >= 0x80) { o[n++] = (uint8_t)(v | 0x80); v >>= 7; }
  o[n++] = (uint8_t)v;
  return n;
}
/* Returns bytes consumed, 0 if truncated, -1 on overflow (stew/leb128 fromBytes). */
static int get_varint(const uint8_t *p, size_t n, int bits, uint64_t *out) {
  uint64_t v = 0;
  int max_bytes = (bits + 6) / 7;
  for (int i = 0; i < max_bytes; i++) {
    if ((size_t)i >= n) return 0;
    uint64_t b = p[i];
    int shift = 7 * i;
    if (i == max_bytes - 1) {
      /* last permitted byte: no continuation, payload must fit */
      if (b & 0x80) return -1;
      if (bits - shift < 7 && (b >> (bits - shift)) != 0) return -1;
    }
    v |= (b & 0x7f) << shift;
    if (!(b & 0x80)) { *out = v; return i + 1; }
  }
  return -1;
}
/* codec.nim:129-138: u64 varint; -1 when absent/invalid. */
int64_t orc_uncompressed_len(const uint8_t *p, size_t n, uint64_t *out) {
  int r = get_varint(p, n, 64, out);
  return r > 0 ? r : -1;
}

/* -------------------------------------------------------------- encoder */
/* encoder.nim:44-73 (the `fast` 16-byte over-copy only touches scratch). */
static size_t emit_literal(uint8_t *o, const uint8_t *s, size_t n) {
  uint32_t m = (uint32_t)((uint16_t)n - 1) & 0xffff; /* uint16(slen)-1 */
  size_t w;
  if (m < 60) { o[0] = (uint8_t)(m << 2); w = 1; }
  else if (m < 256) { o[0] = 60 << 2; o[1] = (uint8_t)m; w = 2; }
  else { o[0] = 61 << 2; o[1] = (uint8_t)m; o[2] = (uint8_t)(m >> 8); w = 3; }
  memcpy(o + w, s, n);
  return w + n;
}
/* encoder.nim:81-125 */
static size_t emit_copy(uint8_t *o, uint32_t off, uint32_t len) {
  size_t w = 0;
  while (len >= 68) { o[w] = (63 << 2) | 2; o[w + 1] = (uint8_t)off; o[w + 2] = (uint8_t)(off >> 8); w += 3; len -= 64; }
  if (len > 64) { o[w] = (59 << 2) | 2; o[w + 1] = (uint8_t)off; o[w + 2] = (uint8_t)(off >> 8); w += 3; len -= 60; }
  if (len >= 12 || off >= 2048) {
    o[w] = (uint8_t)(((len - 1) << 2) | 2); o[w + 1] = (uint8_t)off; o[w + 2] = (uint8_t)(off >> 8);
    return w + 3;
  }
  o[w] = (uint8_t)(((off >> 8) << 5) | ((len - 4) << 2) | 1);
  o[w + 1] = (uint8_t)off;
  return w + 2;
}
/* encoder.nim:130-182: longest common prefix of in[s1..], in[s2..], s2 side bounded by len. */
static uint32_t match_len(const uint8_t *in, size_t len, size_t s1, size_t s2) {
  uint32_t m = 0;
  while (s2 + 8 <= len) {
    uint64_t x = ld64(in + s1 + m) ^ ld64(in + s2);
    if (x) return m + (uint32_t)(__builtin_ctzll(x) >> 3);
    s2 += 8; m += 8;
  }
  while (s2 < len && in[s1 + m] == in[s2]) { s2++; m++; }
  return m;
}

/*
 * encoder.nim:184-383.  hash_bits / exit_ge exist only for the structural
 * self-test against bundled libsnappy (15, 1); the reference is (14, 0).
 */
size_t orc_encode_block_ex(const uint8_t *in, size_t len, uint8_t *out, int hash_bits, int exit_ge) {
  if (len < MIN_NON_LITERAL) return emit_literal(out, in, len); /* :227-229 */

  static __thread uint16_t table[1 << 15];
  uint32_t tsize = 256;                                        /* :27-34 */
  while (tsize < (1u << hash_bits) && tsize < len) tsize *= 2;
  const uint32_t mask = tsize - 1;
  const int shift = 32 - hash_bits;                            /* :36-37 */
  memset(table, 0, tsize * sizeof(uint16_t));
#define HASH(u) ((((uint32_t)(u)) * 0x1e35a7bdu >> shift) & mask)

  const size_t ip_limit = len - INPUT_MARGIN;                  /* :245 */
  size_t ip = 0, op = 0;

  for (;;) {
    size_t next_emit = ip++;
    uint32_t skip = 32, cand = 0;
    int found = 0;
    if (ip_limit >= ip + 16) {                                 /* :281-309 */
      for (uint32_t i = 0; i < 16; i++) {
        uint32_t d = ld32(in + ip + i), h = HASH(d);
        cand = table[h];
        table[h] = (uint16_t)(ip + i);
        if (ld32(in + cand) == d) { ip += i; found = 1; break; }
      }
      if (!found) { ip += 16; skip += 16; }
    }
    while (!found) {                                           /* :311-331 */
      uint32_t d = ld32(in + ip), h = HASH(d);
      uint32_t step = skip >> 5;
      skip += step;
      size_t next_ip = ip + step;
      if (next_ip > ip_limit) goto remainder_from_next_emit;   /* :319-321 */
      cand = table[h];
      table[h] = (uint16_t)ip;
      if (ld32(in + cand) == d) break;
      ip = next_ip;
    }
    op += emit_literal(out + op, in + next_emit, ip - next_emit); /* :336-338 */

    for (;;) {                                                 /* :350-380 */
      size_t base = ip;
      uint32_t matched = 4 + match_len(in, len, cand + 4, ip + 4);
      ip += matched;
      op += emit_copy(out + op, (uint32_t)(base - cand), matched);
      if (exit_ge ? ip >= ip_limit : ip > ip_limit) {          /* :362 */
        next_emit = ip;
        goto remainder_from_next_emit;
      }
      table[HASH(ld32(in + ip - 1))] = (uint16_t)(ip - 1);     /* :371 */
      uint32_t d = ld32(in + ip), h = HASH(d);
      cand = table[h];
      table[h] = (uint16_t)ip;
      if (ld32(in + cand) != d) break;
    }
    continue;

  remainder_from_next_emit:                                    /* :249-253 */
    if (next_emit < len) op += emit_literal(out + op, in + next_emit, len - next_emit);
    return op;
  }
#undef HASH
}
size_t orc_encode_block(const uint8_t *in, size_t len, uint8_t *out) {
  return orc_encode_block_ex(in, len, out, 14, 0);
}

/* encoder.nim:385-426.  out needs 32+len+len/6 bytes. */
size_t orc_encode_frame(const uint8_t *in, size_t len, uint8_t *out) {
  uint32_t crc = orc_masked_crc32c(in, len);
  memcpy(out + 4, &crc, 4);
  if (len >= MIN_NON_LITERAL) {
    uint8_t hdr[5];
    int hl = put_varint32(hdr, (uint32_t)len);
    size_t bl = orc_encode_block(in, len, out + 8 + hl);
    if (bl <= len - len / 8) {
      memcpy(out + 8, hdr, hl);
      uint32_t fl = (uint32_t)(hl + bl + 4);
      out[0] = 0x00; out[1] = (uint8_t)fl; out[2] = (uint8_t)(fl >> 8); out[3] = (uint8_t)(fl >> 16);
      return fl + 4;
    }
  }
  uint32_t fl = (uint32_t)len + 4;
  out[0] = 0x01; out[1] = (uint8_t)fl; out[2] = (uint8_t)(fl >> 8); out[3] = (uint8_t)(fl >> 16);
  memcpy(out + 8, in, len);
  return fl + 4;
}

/* snappy.nim:27-64 */
int orc_compress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *written) {
  if (n > 0xffffffffull) return ORC_INVALID_INPUT;
  if (cap < orc_max_compressed_len(n)) return ORC_BUFFER_TOO_SMALL;
  uint64_t w = put_varint32(out, (uint32_t)n), r = 0;
  while (r < n) {
    uint64_t b = n - r < MAX_BLOCK ? n - r : MAX_BLOCK;
    w += orc_encode_block(in + r, b, out + w);
    r += b;
  }
  *written = w;
  return ORC_OK;
}

/* snappy.nim:130-155 */
int orc_compress_framed(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, uint64_t *written) {
  if (cap < orc_max_compressed_len_framed((int64_t)n)) return ORC_BUFFER_TOO_SMALL;
  memcpy(out, kStreamHeader, 10);
  uint64_t w = 10, r = 0;
  while (r < n) {
    uint64_t b = n - r < MAX_BLOCK ? n - r : MAX_BLOCK;
    w += orc_encode_frame(in + r, b, out + w);
    r += b;
  }
  *written = w;
  return ORC_OK;
}

/* Framed compress of chunks [first, first+count) only, no stream header: the
 * unit of work bench.py's multi-threaded CPU baseline partitions over cores. */
uint64_t orc_encode_frames_range(const uint8_t *in, uint64_t n, uint64_t first, uint64_t count, uint8_t *out) {
  uint64_t w = 0;
  for (uint64_t c = first; c < first + count; c++) {
    uint64_t r = c * MAX_BLOCK;
    if (r >= n) break;
    uint64_t b = n - r < MAX_BLOCK ? n - r : MAX_BLOCK;
    w += orc_encode_frame(in + r, b, out + w);
  }
  return w;
}

/* -------------------------------------------------------------- decoder */
/* decoder.nim:20-155.  Returns status; *written = bytes produced on success. */
int orc_decode_all_tags(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written) {
  *written = 0;
  if (n == 0) return ORC_OK;                     /* :26-27 */
  if (cap == 0) return ORC_BUFFER_TOO_SMALL;     /* :29-30 */
  size_t ip = 0, op = 0;
  while (ip < n) {
    uint8_t tag = in[ip];
    size_t length;
    uint32_t offset;
    switch (tag & 3) {
    case 0: {
      ip += 1;
      length = (size_t)(tag >> 2) + 1;
      if (length >= 61) {                        /* :54-75 */
        if (n - ip < 61) return ORC_INVALID_INPUT;
        size_t ll = length - 60;
        uint32_t v = ld32(in + ip);
        if (ll < 4) v &= (1u << (8 * ll)) - 1;
        uint32_t l32 = v + 1;
        if (l32 == 0) return ORC_INVALID_INPUT;  /* :67-68 */
        length = l32;
        ip += ll;
      }
      if (cap - op < length || n - ip < length) return ORC_INVALID_INPUT; /* :77-79 */
      memcpy(out + op, in + ip, length);
      op += length; ip += length;
      continue;
    }
    case 1:
      if (n - ip < 2) return ORC_INVALID_INPUT;
      length = 4 + ((tag >> 2) & 7);
      offset = ((uint32_t)(tag & 0xe0) << 3) | in[ip + 1];
      ip += 2;
      break;
    case 2:
      if (n - ip < 3) return ORC_INVALID_INPUT;
      length = 1 + (tag >> 2);
      offset = in[ip + 1] | ((uint32_t)in[ip + 2] << 8);
      ip += 3;
      break;
    default:
      if (n - ip < 5) return ORC_INVALID_INPUT;
      length = 1 + (tag >> 2);
      offset = ld32(in + ip + 1);
      ip += 5;
      break;
    }
    if ((uint32_t)op <= offset - 1u) return ORC_INVALID_INPUT;   /* :112-113 */
    if (cap - op < length) return ORC_INVALID_INPUT;             /* :127 */
    for (size_t i = 0; i < length; i++) out[op + i] = out[op + i - offset]; /* :130-151 */
    op += length;
  }
  *written = op;
  return ORC_OK;
}

/* snappy.nim:84-110 */
int orc_uncompress(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *written) {
  uint64_t len;
  *written = 0;
  int br = get_varint(in, n, 32, &len);
  if (br <= 0) return ORC_INVALID_INPUT;
  if (cap < len) return ORC_BUFFER_TOO_SMALL;
  if (len == 0) return (size_t)br != n ? ORC_INVALID_INPUT : ORC_OK;
  size_t w;
  int st = orc_decode_all_tags(in + br, n - br, out, cap, &w);
  if (st) return st;
  if (w != len) return ORC_INVALID_INPUT;
  *written = w;
  return ORC_OK;
}

/* snappy.nim:169-267.  On ORC_OK, (*read,*written) is the reference's tuple. */
int orc_uncompress_framed(const uint8_t *in, size_t n, uint8_t *out, size_t cap, int check_header,
                          int check_integrity, size_t *read_out, size_t *written_out) {
  size_t rd = 0, wr = 0;
  *read_out = 0; *written_out = 0;
  if (check_header) {
    if (n < 10 || memcmp(in, kStreamHeader, 10) != 0) return ORC_INVALID_INPUT;
    rd = 10;
  }
  while (rd < n) {
    size_t remaining = n - rd;
    if (remaining < 4) return ORC_INVALID_INPUT;
    uint32_t hdr = ld32(in + rd);
    uint8_t id = (uint8_t)hdr;
    size_t dlen = hdr >> 8;
    rd += 4;
    if (remaining - 4 < dlen) return ORC_INVALID_INPUT;
    if (id == 0x00) {
      if (dlen < 4) return ORC_INVALID_INPUT;
      uint32_t crc = ld32(in + rd);
      size_t room = cap - wr < MAX_BLOCK ? cap - wr : MAX_BLOCK, got;
      int st = orc_uncompress(in + rd + 4, dlen - 4, out + wr, room, &got);
      if (st == ORC_BUFFER_TOO_SMALL) {                       /* :219-229 */
        uint64_t ul;
        if (orc_uncompressed_len(in + rd + 4, dlen - 4, &ul) < 0 || ul > MAX_BLOCK) return ORC_INVALID_INPUT;
        *read_out = rd - 4; *written_out = wr;
        return ORC_OK;
      }
      if (st) return ORC_INVALID_INPUT;
      if (check_integrity && orc_masked_crc32c(out + wr, got) != crc) return ORC_CRC_MISMATCH;
      wr += got;
    } else if (id == 0x01) {
      if (dlen < 4) return ORC_INVALID_INPUT;
      uint32_t crc = ld32(in + rd);
      if (check_integrity && orc_masked_crc32c(in + rd + 4, dlen - 4) != crc) return ORC_CRC_MISMATCH;
      size_t ul = dlen - 4;
      if (ul > MAX_BLOCK) return ORC_INVALID_INPUT;
      if (ul > cap - wr) { *read_out = rd - 4; *written_out = wr; return ORC_OK; }
      memcpy(out + wr, in + rd + 4, ul);
      wr += ul;
    } else if (id < 0x80) {
      return ORC_UNKNOWN_CHUNK;
    }
    rd += dlen;
  }
  *read_out = rd; *written_out = wr;
  return ORC_OK;
}

/* codec.nim:178-214: header-only walk; -1 = none. */
int64_t orc_uncompressed_len_framed(const uint8_t *in, size_t n) {
  size_t rd = 0;
  uint64_t expected = 0;
  while (rd < n) {
    size_t remaining = n - rd;
    if (remaining < 4) return -1;
    uint32_t hdr = ld32(in + rd);
    uint8_t id = (uint8_t)hdr;
    size_t dlen = hdr >> 8;
    if (remaining < dlen + 4) return -1;
    rd += 4;
    uint64_t ul = 0;
    if (id == 0x00) {
      if (dlen < 4) return -1;
      if (orc_uncompressed_len(in + rd + 4, dlen - 4, &ul) < 0) return -1;
    } else if (id == 0x01) {
      if (dlen < 4) return -1;
      ul = dlen - 4;
    } else if (id < 0x80) {
      return -1;
    }
    if (ul > MAX_BLOCK) return -1;
    expected += ul;
    rd += dlen;
  }
  return (int64_t)expected;
}

/* Decode chunks of a framed stream starting at byte `start` (a chunk header)
 * until `max_chunks` data chunks are done; used by the threaded CPU baseline. */
int orc_uncompress_frames_range(const uint8_t *in, size_t n, size_t start, size_t max_chunks, uint8_t *out,
                                size_t cap, size_t *read_out, size_t *written_out) {
  size_t rd = start, wr = 0, done = 0;
  while (rd < n && done < max_chunks) {
    if (n - rd < 4) return ORC_INVALID_INPUT;
    uint32_t hdr = ld32(in + rd);
    uint8_t id = (uint8_t)hdr;
    size_t dlen = hdr >> 8;
    if (n - rd - 4 < dlen) return ORC_INVALID_INPUT;
    rd += 4;
    if (id <= 0x01) {
      if (dlen < 4) return ORC_INVALID_INPUT;
      uint32_t crc = ld32(in + rd);
      size_t got;
      if (id == 0x00) {
        size_t room = cap - wr < MAX_BLOCK ? cap - wr : MAX_BLOCK;
        if (orc_uncompress(in + rd + 4, dlen - 4, out + wr, room, &got)) return ORC_INVALID_INPUT;
      } else {
        got = dlen - 4;
        if (got > cap - wr) return ORC_BUFFER_TOO_SMALL;
        memcpy(out + wr, in + rd + 4, got);
      }
      if (orc_masked_crc32c(out + wr, got) != crc) return ORC_CRC_MISMATCH;
      wr += got;
      done++;
    } else if (id < 0x80) {
      return ORC_UNKNOWN_CHUNK;
    }
    rd += dlen;
  }
  *read_out = rd; *written_out = wr;
  return ORC_OK;
}
