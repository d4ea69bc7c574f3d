"""Known-answer vectors transcribed from the reference's own tests (file:line cited per group).
Also written to tests/golden/vectors.json by tests/golden/make_vectors.py for non-Python consumers."""

# tests/test_snappy.nim:161-218 -- every one must decode to the empty sequence
MALFORMED_RAW = [
    b"\x05\x00a",                                   # :168 fewer bytes than the header reports
    b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\x00",  # :171 varint overflows u64
    b"\x80\x80\x80\x80\x10",                        # :174 fits u64, overflows u32
    b"\x02\x00hi",                                  # :179 literal too short, then truncated copy1
    b"\x02\xechi",                                  # :182 literal length too big
    b"\x02\xf0hi",                                  # :186 literal needs 1 extra length byte, src too short
    b"\x02\xf0hi\x00\x00\x00",                      # :190 src too short for the full literal
    b"\x02\x00a\x01",                               # :194 copy1 stops at the tag
    b"\x11\x00a\x3e",                               # :198 copy2 stops at the tag
    b"\x11\x00a\x3e\x01",                           # :199 copy2 stops inside the offset
    b"\x11\x00a\x3f",                               # :202 copy4 truncated ...
    b"\x11\x00a\x3f\x00",
    b"\x11\x00a\x3f\x00\x00",
    b"\x11\x00a\x3f\x00\x00\x00",                   # :205
    b"\x11\x00a\x01\x00",                           # :208 offset zero
    b"\x11\x00a\x01\xff",                           # :211 offset too big
    b"\x05\x00a\x1d\x01",                           # :214 copy length too big
    b"\x11\x00\x00\xfc\xfe\xff\xff\xff",            # :216 4-byte literal length
    b"\x11\x00\x00\xfc\xff\xff\xff\xff",            # :218 4-byte literal length wraps
]

# tests/test_snappy.nim:136-148 -- hand-built valid streams and their exact decodings
HANDBUILT_RAW = [
    (bytes([27, 0b000010_00, 1, 2, 3, 0b000_000_10, 3, 0, 0b010110_00] + list(range(4, 27))),
     bytes([1, 2, 3, 1] + list(range(4, 27)))),
    (bytes([28, 0b000010_00, 1, 2, 3, 0b000_000_10, 3, 0, 0b010111_00] + list(range(4, 28))),
     bytes([1, 2, 3, 1] + list(range(4, 28)))),
]

# tests/test_snappy.nim:222-240 -- quickcheck failure witnesses (round-trip inputs)
RANDOM1 = bytes([
    0, 0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0, 5, 0, 0, 1, 1,
    0, 0, 1, 2, 0, 0, 2, 1, 0, 0, 2, 2, 0, 0, 0, 6, 0, 0, 3, 1, 0, 0, 0, 7, 0,
    0, 1, 3, 0, 0, 0, 8, 0, 0, 2, 3, 0, 0, 0, 9, 0, 0, 1, 4, 0, 0, 1, 0, 0, 3,
    0, 0, 1, 0, 1, 0, 0, 0, 10, 0, 0, 0, 0, 2, 4, 0, 0, 2, 0, 0, 3, 0, 1, 0, 0,
    1, 5, 0, 0, 6, 0, 0, 0, 0, 11, 0, 0, 1, 6, 0, 0, 1, 7, 0, 0, 0, 12, 0, 0,
    3, 2, 0, 0, 0, 13, 0, 0, 2, 5, 0, 0, 0, 3, 3, 0, 0, 0, 1, 8, 0, 0, 1, 0,
    1, 0, 0, 0, 4, 1, 0, 0, 0, 0, 14, 0, 0, 0, 1, 9, 0, 0, 0, 1, 10, 0, 0, 0,
    0, 1, 11, 0, 0, 0, 1, 0, 2, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 5, 1, 0, 0, 0, 1,
    2, 1, 0, 0, 0, 0, 0, 2, 6, 0, 0, 0, 0, 0, 1, 12, 0, 0, 0, 0, 0, 3, 4, 0, 0,
    0, 0, 0, 7, 0, 0, 0, 0, 0, 1, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
RANDOM2 = bytes([10, 2, 14, 13, 0, 8, 2, 10, 2, 14, 13, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
RANDOM3 = bytes([0, 0, 0, 4, 1, 4, 0, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
RANDOM4 = bytes([
    0, 0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0, 5, 0, 0, 1, 1,
    0, 0, 1, 2, 0, 0, 1, 3, 0, 0, 1, 4, 0, 0, 2, 1, 0, 0, 0, 4, 0, 1, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
WITNESSES = [RANDOM1, RANDOM2, RANDOM3, RANDOM4]

# tests/test_framed.nim:117-119 -- (fixture, source) pairs; masked CRCs of alice29.txt's 64 KiB chunks
FRAMED_FIXTURES = [("alice29.txt.sz-32k", "alice29.txt"), ("alice29.txt.sz-64k", "alice29.txt"),
                   ("house.jpg.sz", "house.jpg")]
ALICE_CHUNK_CRCS = [0x8C7A2D1B, 0x17515333, 0xC0D03CB9]

# tests/test_framed.nim:198-200 -- invalid framed payloads (with output sizes)
INVALID_FRAMED = [(bytes([3, 2, 1, 0]), 0), (bytes([0, 0, 0, 0, 42]), 0)]

# Encoder output of the validated restatement: name -> (raw len, framed len)
ENCODED_LENGTHS = {
    "alice29.txt": (88034, 88074), "asyoulik.txt": (78455, 78484), "lcet10.txt": (234653, 234737),
    "plrabn12.txt": (319262, 319357), "html": (22843, 22872), "html_x_4": (92234, 92318),
    "urls.10K": (335491, 335619), "geo.protodata": (23335, 23364), "kppkn.gtb": (69522, 69562),
    "paper-100k.pdf": (85304, 85327), "fireworks.jpeg": (123034, 123119), "house.jpg": (126806, 126984),
    "Mark.Twain-Tom.Sawyer.txt": (9977, 9995),
}
ALICE_RAW_SHA256 = "d9b27949428e5678cd7a4f00baaba000612d180d9028d28a6ab3a5e308272869"
ALICE_FRAMED_SHA256 = "e96bd7aceb34fe1db3a696d6b3f92d507872abee05093f12ac063f1a2d00ed0f"
