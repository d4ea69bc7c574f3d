"""Second, independently written restatement of the reference's block encoder -- TEST INFRASTRUCTURE ONLY.

Why it exists: no byte the Nim encoder ever produced for COMPRESSIBLE input is available here (no Nim
toolchain; the reference's fixtures pin the decoder, the CRC, the framing walk and the stored-chunk encoder
only), so the C oracle's encoder (oracle/snappy_oracle.c) is pinned structurally.  This module is a second
reading of the same source, written line by line from /root/reference/snappy/encoder.nim WITHOUT looking at
the C restatement and sharing no code with it: it keeps the reference's own formulation -- the `preload` /
`data` registers (encoder.nim:247, :272, :296-305, :329, :381), the unrolled 4 x 4 first probes (:281-309),
findMatchLength's three stages and its stale-`data` corner (:130-182) -- where the C oracle uses direct loads.
tests/test_oracle.py fuzzes the two against each other; a disagreement means one of the two readings is wrong.

Only tests/ may import this file.
"""

MAX_HASH_TABLE_BITS = 14                     # encoder.nim:11
MAX_TABLE_SIZE = 1 << MAX_HASH_TABLE_BITS    # :12
INPUT_MARGIN = 15                            # codec.nim:26
MIN_NON_LITERAL_BLOCK_SIZE = 17              # codec.nim:53
TAG_LITERAL, TAG_COPY1, TAG_COPY2 = 0, 1, 2  # codec.nim:21-24
U16, U32, U64 = 0xFFFF, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF


def _load(buf, off, nbytes):
    """load32 / load64 (encoder.nim:17-25): unchecked little-endian loads.  The reference guarantees by
    construction that every load it USES lies inside the input; a load that would run past the end is a bug in
    this restatement (or a dead value), so bytes past the end read as a poison pattern, not as zeros."""
    chunk = buf[off:off + nbytes]
    if len(chunk) < nbytes:
        chunk = bytes(chunk) + b"\xa5" * (nbytes - len(chunk))
    return int.from_bytes(chunk, "little")


def table_size(n):                           # encoder.nim:27-34
    size = 1 << 8
    while size < MAX_TABLE_SIZE and size < n:
        size *= 2
    return size, size - 1


def hash_(u, mask):                          # encoder.nim:36-37
    return (((u * 0x1E35A7BD) & U32) >> (32 - MAX_HASH_TABLE_BITS)) & mask


def emit_literal(dst, src, fast=False):      # encoder.nim:44-73; dst is a bytearray we append to
    slen = len(src)
    n = (slen & U16) - 1 & U16               # uint16(slen) - 1
    if fast and slen <= 16:                  # :52-57 (the 16-byte over-copy lands in scratch beyond the result)
        dst.append(((n & 0xFF) << 2 & 0xFF) | TAG_LITERAL)
        dst += src
        return
    if n < 60:
        dst.append(((n << 2) & 0xFF) | TAG_LITERAL)
    elif n < (1 << 8):
        dst.append(((60 << 2) & 0xFF) | TAG_LITERAL)
        dst.append(n & 0xFF)
    else:
        dst.append(((61 << 2) & 0xFF) | TAG_LITERAL)
        dst.append(n & 0xFF)
        dst.append((n >> 8) & 0xFF)
    dst += src


def emit_copy(dst, offset, length):          # encoder.nim:81-125
    while length >= 68:
        dst += bytes((((63 << 2) & 0xFF) | TAG_COPY2, offset & 0xFF, (offset >> 8) & 0xFF))
        length -= 64
    if length > 64:
        dst += bytes((((59 << 2) & 0xFF) | TAG_COPY2, offset & 0xFF, (offset >> 8) & 0xFF))
        length -= 60
    if length >= 12 or offset >= 2048:
        dst += bytes(((((length - 1) & 0xFF) << 2 & 0xFF) | TAG_COPY2, offset & 0xFF, (offset >> 8) & 0xFF))
        return
    dst += bytes(((((offset >> 8) << 5) & 0xFF) | (((length - 4) << 2) & 0xFF) | TAG_COPY1, offset & 0xFF))


def first_one(x):                            # stew/bitops2 firstOne: 1-based index of the lowest set bit
    return (x & -x).bit_length()


def find_match_length(inp, s1, s2, data):    # encoder.nim:130-182 -> (matched, data)
    n = len(inp)
    matched = 0
    if s2 + 16 <= n:                                             # :138
        a1, a2 = _load(inp, s1, 8), _load(inp, s2, 8)
        if a1 != a2:
            matched_bytes = (first_one(a1 ^ a2) - 1) >> 3
            return matched_bytes, _load(inp, s2 + matched_bytes, 8)
        s2 += 8
        matched = 8
    while s2 + 16 <= n:                                          # :156
        a1, a2 = _load(inp, s1 + matched, 8), _load(inp, s2, 8)
        if a1 == a2:
            s2 += 8
            matched += 8
        else:
            matched_bytes = (first_one(a1 ^ a2) - 1) >> 3
            return (matched + matched_bytes) & U16, _load(inp, s2 + matched_bytes, 8)
    while s2 < n:                                                # :173
        if inp[s1 + matched] == inp[s2]:
            s2 += 1
            matched += 1
        else:
            if s2 + 8 <= n:
                data = _load(inp, s2, 8)
            break
    return matched & U16, data


def encode_block(inp):                       # encoder.nim:184-383 -> bytes
    inp = bytes(inp)
    n = len(inp)
    assert 0 < n <= 65536
    out = bytearray()
    if n < MIN_NON_LITERAL_BLOCK_SIZE:                           # :227-229
        emit_literal(out, inp)
        return bytes(out)
    tsize, tmask = table_size(n)                                 # :234-238
    table = [0] * tsize
    ip_limit = n - INPUT_MARGIN                                  # :245
    ip = 0
    preload = _load(inp, ip + 1, 4)                              # :247

    def remainder(frm):                                          # :249-253
        if frm < n:
            emit_literal(out, inp[frm:])
        return bytes(out)

    while True:                                                  # :255
        next_emit = ip                                           # :268
        ip += 1
        data = _load(inp, ip, 8)                                 # :272
        skip = 32
        candidate = 0
        found = False
        if ip_limit >= ip + 16:                                  # :277
            delta = ip & U16
            for j in range(4):
                for k in range(4):
                    i = 4 * j + k
                    dword = preload if i == 0 else data & U32    # :286
                    h = hash_(dword, tmask)
                    candidate = table[h]
                    table[h] = (delta + i) & U16
                    if _load(inp, candidate, 4) == dword:        # :293
                        out.append(((i << 2) & 0xFF) | TAG_LITERAL)     # :294: a literal of i + 1 bytes
                        out += inp[next_emit:next_emit + i + 1]         # (16 copied, op advances by i + 2)
                        ip += i
                        found = True
                        break
                    data >>= 8                                   # :301
                if found:
                    break
                data = _load(inp, ip + 4 * j + 4, 8)             # :303
            if not found:
                ip += 16                                         # :305-306
                skip += 16
        if not found:
            while True:                                          # :308
                h = hash_(data & U32, tmask)
                between = skip >> 5
                skip += between
                next_ip = ip + between
                if next_ip > ip_limit:                           # :316-318
                    return remainder(next_emit)
                candidate = table[h]
                table[h] = ip & U16
                if data & U32 == _load(inp, candidate, 4):       # :324
                    break
                data = _load(inp, next_ip, 4)                    # :327
                ip = next_ip
            emit_literal(out, inp[next_emit:ip], fast=True)      # :333-335
        while True:                                              # :349
            base = ip & U16
            m, data = find_match_length(inp, candidate + 4, ip + 4, data)
            matched = (m + 4) & U16
            ip += matched
            emit_copy(out, (base - candidate) & U16, matched)    # :357-358
            if ip > ip_limit:                                    # :360-361
                return remainder(ip)
            table[hash_(_load(inp, ip - 1, 4), tmask)] = (ip - 1) & U16   # :369
            h = hash_(data & U32, tmask)
            candidate = table[h]
            table[h] = ip & U16
            if data & U32 != _load(inp, candidate, 4):           # :377
                break
        preload = (data >> 8) & U32                              # :379


def compress(inp):                           # snappy.nim:27-64: Leb128(len) + one encode_block per 64 KiB
    inp = bytes(inp)
    out = bytearray()
    v = len(inp)
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    for off in range(0, len(inp), 65536):
        out += encode_block(inp[off:off + 65536])
    return bytes(out)
