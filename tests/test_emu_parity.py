"""CPU-only parity of the PRODUCT sources (csrc/snb_device.cuh + snb_api.cu) compiled against the
SIMT emulator in tests/emu: the kernels' warp-level logic and the host chunk drivers are checked
against the oracle without a GPU.  Sizes are small because every warp collective costs 32 fiber
switches; the full-size parity runs on the B200 in test_gpu_parity.py."""
import numpy as np
import pytest

import parity_cases as P
import vectors as V
from conftest import read_data


def test_crc(emu_api, oracle):
    P.check_crc(emu_api, oracle, [0, 1, 3, 4, 5, 15, 16, 17, 511, 512, 513, 527, 528, 1500, 4096, 21017, 65536, 70000])
    alice = read_data("alice29.txt")
    assert [emu_api.maskedCrc(alice[i:i + 65536]) for i in range(0, len(alice), 65536)] == V.ALICE_CHUNK_CRCS


def test_crc_unaligned(emu_api, oracle):
    base = np.frombuffer(read_data("html")[:70000], dtype=np.uint8)
    for mis in (1, 3, 8, 13):
        d = base[mis:mis + 66000]
        assert emu_api.maskedCrc(d) == oracle.masked_crc32c(d)


@pytest.mark.parametrize("name,limit", [("alice29.txt", 70000), ("html", 66000), ("geo.protodata", 40000),
                                        ("fireworks.jpeg", 66000), ("kppkn.gtb", 30000), ("urls.10K", 50000),
                                        ("Mark.Twain-Tom.Sawyer.txt", None)])
def test_corpus_prefixes(emu_api, oracle, name, limit):
    d = read_data(name)[:limit]
    P.check_raw(emu_api, oracle, d, oracle.ref_snappycpp())
    P.check_framed(emu_api, oracle, d)


def test_small_and_boundary_sizes(emu_api, oracle):
    for i in range(1, 33, 3):                                  # tests/test_snappy.nim:111-114
        P.check_raw(emu_api, oracle, b"aaaa" + b"b" * i + b"aaaabbbb")
    pat = bytes((j % 10) + ord("a") for j in range(70000))
    for n in [1, 2, 15, 16, 17, 18, 31, 32, 33, 47, 64, 100, 255, 256, 257, 1000, 4095, 4096, 4097, 16385]:
        P.check_raw(emu_api, oracle, pat[:n])
        P.check_raw(emu_api, oracle, bytes(n))
    for n in (65531, 65536, 65541):                            # :124-134 (m = 1)
        P.check_raw(emu_api, oracle, pat[:n])
        P.check_raw(emu_api, oracle, bytes(n))
    for w in V.WITNESSES:                                      # :222-245
        P.check_raw(emu_api, oracle, w)
    assert len(emu_api.encode(bytes(1024))) < 512              # :150-154


def test_random_strings(emu_api, oracle):
    rng = np.random.default_rng(11)
    for _ in range(12):                                        # tests/test_snappy.nim:247-253 (seeded here)
        n = int(rng.integers(1000, 10000))
        alphabet = int(rng.integers(2, 64))
        P.check_raw(emu_api, oracle, rng.integers(0, alphabet, n, dtype=np.uint8).tobytes())
    P.check_raw(emu_api, oracle, rng.integers(0, 256, 20000, dtype=np.uint8).tobytes())


def test_malformed(emu_api, oracle):
    P.check_malformed(emu_api, oracle)


def test_framed_fixture_house(emu_api, oracle):
    P.check_framed_fixture(emu_api, oracle, "house.jpg.sz", "house.jpg")


def test_framed_fixture_alice(emu_api, oracle):
    P.check_framed_fixture(emu_api, oracle, "alice29.txt.sz-64k", "alice29.txt")


def test_framed_edges(emu_api, oracle):
    P.check_framed_edges(emu_api, oracle, full=False)


def test_decoder_differential_fuzz(emu_api, oracle):
    rng = np.random.default_rng(5)
    base = oracle.encode(read_data("alice29.txt")[:3000])
    for s in P.mutate_streams(oracle, base, rng, 60):
        P.check_raw_decode_parity(emu_api, oracle, s)
    fbase = oracle.encode_framed(read_data("html")[:2000] + bytes(rng.integers(0, 256, 300, dtype=np.uint8)))
    for s in P.mutate_streams(oracle, fbase, rng, 40):
        P.check_framed_decode_parity(emu_api, oracle, s, 4096)


@pytest.mark.parametrize("algo", [5, 6])
def test_every_block_encoder_variant(emu_api, oracle, algo):
    """The two selectable encoders for 64 KiB chunks (5: one warp per chunk, the default; 6: resolver warps + walker
    warp with the record ring) are byte-identical to the oracle, including the chunks v6 hands to v5 (short last chunk)."""
    lib = emu_api.lib
    lib.snb_set_compress_algo(algo)
    try:
        for name, cut in [("alice29.txt", 140000), ("paper-100k.pdf", 102400), ("kppkn.gtb", 70000), ("html_x_4", 140000),
                          ("fireworks.jpeg", 70000), ("urls.10K", 131072)]:
            d = read_data(name)[:cut]
            P.check_raw(emu_api, oracle, d)
            P.check_framed(emu_api, oracle, d)
        for n in [0, 1, 16, 17, 18, 300, 65535, 65536, 65537, 131072]:
            P.check_framed(emu_api, oracle, bytes(i & 0xFF for i in range(n)))
            P.check_raw(emu_api, oracle, bytes(n))
        rng = np.random.default_rng(31)
        for alphabet in (2, 4, 64):                                 # dense short matches: many same-window hash groups
            P.check_framed(emu_api, oracle, rng.integers(0, alphabet, 65536 + 777, dtype=np.uint8).tobytes())
    finally:
        lib.snb_set_compress_algo(0)


def test_raw_multiblock(emu_api, oracle):
    rng = np.random.default_rng(9)
    alice = read_data("alice29.txt")
    P.check_raw_multiblock(emu_api, oracle, oracle.ref_snappycpp(),
                           [alice[:140000], rng.integers(0, 256, 135000, dtype=np.uint8).tobytes(),
                            alice[:66000] + rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + alice[:3000]], fuzz=3)


def test_raw_random_element_streams(emu_api, oracle):
    P.check_random_element_streams(emu_api, oracle, range(4), 150000)


def test_host_message_batch(emu_api, oracle):
    """snb_compress_batch / snb_uncompress_batch (host buffers, packed output): BASELINE config 4 in miniature."""
    rng = np.random.default_rng(17)
    src = read_data("geo.protodata") + read_data("alice29.txt")
    msgs = [b"", b"a", bytes(17)] + [src[o:o + n] for o, n in zip(rng.integers(0, len(src) - 4096, 40), rng.integers(1, 4097, 40))]
    msgs.append(rng.integers(0, 256, 3000, dtype=np.uint8).tobytes())
    P.check_batch(emu_api, oracle, msgs, extra_streams=V.MALFORMED_RAW[:8] + [b"\x05", b""])


def test_host_message_batch_many(emu_api, oracle):
    """More messages than the single-CTA prefix sum has threads (every thread then owns a run of lengths): the packed
    output offsets of snb_compress_batch must still be the running sum."""
    rng = np.random.default_rng(19)
    src = read_data("alice29.txt")
    msgs = [src[o:o + n] for o, n in zip(rng.integers(0, len(src) - 64, 2500), rng.integers(0, 48, 2500))]
    P.check_batch(emu_api, oracle, msgs)


def test_device_side_header_walk(emu_api, oracle):
    """snb_dev_uncompress_framed: the header walk of snappy.nim:199-265 done on the device (candidate scan, links,
    pointer jumping) must give the reference's statuses, (read, written) and bytes -- valid streams, every partial-output
    room, skippable / unknown chunks, trailing bytes, mutated streams; with and without the compressor's frame table."""
    rng = np.random.default_rng(23)
    alice = read_data("alice29.txt")
    datas = [b"", b"abc", alice[:5000], alice[:70000] + rng.integers(0, 256, 66000, dtype=np.uint8).tobytes(), bytes(140000)]
    P.check_dev_framed(P.emu_dev_call(emu_api.lib), oracle, datas, rng, fuzz=12, lib=emu_api.lib)
