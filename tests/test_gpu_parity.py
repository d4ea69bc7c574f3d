"""Parity of the CUDA path (libsnb200.so through the C ABI) against the oracle, on a real B200.

Byte-identical compressed streams (raw and framed), bit-exact decompression of oracle-encoded
streams, identical accept/reject statuses -- at the reference tests' sizes, on every corpus file,
on the committed golden fixtures and at BASELINE.json's sizes (1 GiB corpus, message batches)."""
import hashlib

import numpy as np
import pytest

import parity_cases as P
import vectors as V
from conftest import CORPUS, read_data

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded(gpu_api):
    assert gpu_api.lib.path.endswith("libsnb200.so")
    before = gpu_api.lib.snb_kernel_launches()
    gpu_api.maskedCrc(b"123456789")
    assert gpu_api.lib.snb_kernel_launches() > before


def test_crc(gpu_api, oracle):
    assert gpu_api.maskedCrc(b"") == 0xA282EAD8
    assert gpu_api.maskedCrc(b"123456789") == 0xC78AB0E5
    P.check_crc(gpu_api, oracle, [1, 2, 3, 4, 5, 15, 16, 17, 31, 33, 511, 512, 513, 527, 528, 1500, 4096, 21017,
                                  65535, 65536, 65537, 70000, 131072, 1 << 20, (1 << 22) + 123, 70 << 20])
    alice = read_data("alice29.txt")
    assert [gpu_api.maskedCrc(alice[i:i + 65536]) for i in range(0, len(alice), 65536)] == V.ALICE_CHUNK_CRCS
    base = np.frombuffer(read_data("urls.10K"), dtype=np.uint8)
    for mis in range(1, 16):
        d = base[mis:mis + 66000 + mis * 37]
        assert gpu_api.maskedCrc(d) == oracle.masked_crc32c(d)


@pytest.mark.parametrize("name", CORPUS + ["COPYING", "baddata1.snappy", "baddata2.snappy", "baddata3.snappy"])
def test_corpus(gpu_api, oracle, name):                          # tests/test_snappy.nim:93-108, test_framed.nim:121-135
    d = read_data(name)
    P.check_raw(gpu_api, oracle, d, oracle.ref_snappycpp())
    P.check_framed(gpu_api, oracle, d)


def test_encoded_lengths_and_sha(gpu_api):
    for name, (raw, framed) in V.ENCODED_LENGTHS.items():
        d = read_data(name)
        assert (len(gpu_api.encode(d)), len(gpu_api.encodeFramed(d))) == (raw, framed)
    d = read_data("alice29.txt")
    assert hashlib.sha256(gpu_api.encode(d)).hexdigest() == V.ALICE_RAW_SHA256
    assert hashlib.sha256(gpu_api.encodeFramed(d)).hexdigest() == V.ALICE_FRAMED_SHA256


def test_misc_sizes(gpu_api, oracle):                            # tests/test_snappy.nim:110-134
    for i in range(1, 33):
        P.check_raw(gpu_api, oracle, b"aaaa" + b"b" * i + b"aaaabbbb")
    pat = bytes((j % 10) + ord("a") for j in range(5 * 65536 + 6))
    for n in range(1, 20000, 23):
        P.check_raw(gpu_api, oracle, pat[:n])
    for m in range(1, 6):
        for n in range(m * 65536 - 5, m * 65536 + 6):
            P.check_raw(gpu_api, oracle, bytes(n))
            P.check_raw(gpu_api, oracle, pat[:n])
    for n in list(range(1, 70)) + [255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097, 16383, 16384, 16385]:
        P.check_raw(gpu_api, oracle, bytes(n))
        P.check_framed(gpu_api, oracle, pat[:n])
    assert len(gpu_api.encode(bytes(1024))) < 512                # :150-154


def test_random_data(gpu_api, oracle):                           # tests/test_snappy.nim:220-253
    for w in V.WITNESSES:
        P.check_raw(gpu_api, oracle, w)
        P.check_framed(gpu_api, oracle, w)
    rng = np.random.default_rng(11)
    for _ in range(100):
        n = int(rng.integers(1000, 10000))
        alphabet = int(rng.integers(2, 256))
        P.check_raw(gpu_api, oracle, rng.integers(0, alphabet, n, dtype=np.uint8).tobytes())
    for n in (65536, 200000):
        P.check_framed(gpu_api, oracle, rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        P.check_framed(gpu_api, oracle, rng.integers(0, 4, n, dtype=np.uint8).tobytes())


def test_malformed(gpu_api, oracle):
    P.check_malformed(gpu_api, oracle)


@pytest.mark.parametrize("fixture,source", V.FRAMED_FIXTURES)
def test_framed_fixtures(gpu_api, oracle, fixture, source):      # tests/test_framed.nim:117-119
    P.check_framed_fixture(gpu_api, oracle, fixture, source)


def test_raw_golden(gpu_api, oracle):                            # tests/test_snappy.nim:71-79
    from conftest import read_golden
    raw = read_golden("Mark.Twain-Tom.Sawyer.txt.rawsnappy")
    assert gpu_api.decode(raw) == oracle.decode(raw) and len(oracle.decode(raw)) == 14168


def test_framed_edges(gpu_api, oracle):
    P.check_framed_edges(gpu_api, oracle, full=True)


def test_partial_output_sweep(gpu_api, oracle):
    """Resume contract at every chunk boundary +-1 (snappy.nim:176-183, 219-229, 253-254)."""
    d = read_data("alice29.txt") + read_data("house.jpg")[:70000]
    s = oracle.encode_framed(d)
    for cap in [0, 1, 65535, 65536, 65537, 131071, 131072, 152088, 152089, 152090, 200000, len(d) - 1, len(d), len(d) + 9]:
        P.check_framed_decode_parity(gpu_api, oracle, s, cap)


def test_decoder_differential_fuzz(gpu_api, oracle):             # tests/fuzzing/fuzz_block_format.nim:8-33
    rng = np.random.default_rng(5)
    for name, cut in [("alice29.txt", 3000), ("html", 20000), ("geo.protodata", 70000), ("fireworks.jpeg", 5000)]:
        base = oracle.encode(read_data(name)[:cut])
        for s in P.mutate_streams(oracle, base, rng, 150):
            P.check_raw_decode_parity(gpu_api, oracle, s)
    fbase = oracle.encode_framed(read_data("html")[:100000] + bytes(rng.integers(0, 256, 70000, dtype=np.uint8)))
    for s in P.mutate_streams(oracle, fbase, rng, 150):
        P.check_framed_decode_parity(gpu_api, oracle, s, 200000)
        P.check_framed_decode_parity(gpu_api, oracle, s, 200000, check_integrity=False)


def test_raw_multiblock_and_big(gpu_api, oracle):
    """Raw streams larger than one block: chunk-parallel compress, single-stream decode."""
    d = read_data("plrabn12.txt") + read_data("kppkn.gtb")
    P.check_raw(gpu_api, oracle, d)
    big = np.tile(np.frombuffer(read_data("lcet10.txt"), dtype=np.uint8), 170)[: 70 << 20].tobytes()  # > 1 host slab
    enc = oracle.encode(big)
    assert gpu_api.encode(big) == enc
    n0 = gpu_api.lib.snb_raw_block_decodes()
    assert gpu_api.decode(enc) == big                      # 1 120 blocks, found on the device and decoded in parallel
    assert gpu_api.lib.snb_raw_block_decodes() == n0 + 1


def test_raw_multiblock_decode(gpu_api, oracle):
    """block discovery + block-parallel decode of long raw streams, and the streams that must not take it."""
    import time
    from nim_snappy_b200.workloads import mixed_corpus, text_corpus
    rng = np.random.default_rng(4)
    P.check_raw_multiblock(gpu_api, oracle, oracle.ref_snappycpp(),
                           [read_data("lcet10.txt"), read_data("urls.10K"), read_data("fireworks.jpeg"),
                            read_data("alice29.txt") + rng.integers(0, 256, 300000, dtype=np.uint8).tobytes() + read_data("html")])
    for d in (text_corpus(300 << 20), mixed_corpus(300 << 20, seed=2)):     # beyond the 256 MiB single-warp bound
        enc = oracle.encode(d)
        out = np.empty(d.size, dtype=np.uint8)
        n0 = gpu_api.lib.snb_raw_block_decodes()
        t0 = time.perf_counter()
        r = gpu_api.uncompress(enc, out)
        dt = time.perf_counter() - t0
        assert r == d.size and bool((out == d).all())
        assert gpu_api.lib.snb_raw_block_decodes() == n0 + 1
        print(f"raw uncompress of {d.size >> 20} MiB through host buffers: {d.size / dt / 1e9:.1f} GB/s")


def test_host_api_multi_slab_and_pinned(gpu_api, oracle):
    """Host buffers larger than one 64 MiB slab, pageable and pinned (direct DMA) paths."""
    import torch
    from nim_snappy_b200.workloads import mixed_corpus
    d = mixed_corpus(150 << 20, seed=1)
    st, expect = oracle.compress_framed(d, as_array=True)
    assert st == 0
    out = np.empty(oracle.max_compressed_len_framed(d.size), dtype=np.uint8)
    n = gpu_api.compressFramed(d, out)
    assert n == expect.size and np.array_equal(out[:n], expect)
    back = np.empty(d.size, dtype=np.uint8)
    assert gpu_api.uncompressFramed(expect, back) == (expect.size, d.size)
    assert np.array_equal(back, d)
    pin_in = torch.from_numpy(d).pin_memory()
    pin_out = torch.empty(out.size, dtype=torch.uint8).pin_memory()
    n2 = gpu_api.compressFramed(pin_in.numpy(), pin_out.numpy())
    assert n2 == expect.size and np.array_equal(pin_out.numpy()[:n2], expect)
    pin_s = torch.from_numpy(expect.copy()).pin_memory()
    pin_back = torch.empty(d.size, dtype=torch.uint8).pin_memory()
    assert gpu_api.uncompressFramed(pin_s.numpy(), pin_back.numpy()) == (expect.size, d.size)
    assert np.array_equal(pin_back.numpy(), d)


# ------------------------------------------------------------------ device-resident API
@pytest.fixture(scope="module")
def dc(gpu_api):
    from nim_snappy_b200.device import DeviceCodec
    return DeviceCodec(gpu_api.lib)


def test_device_framed_64mib(dc, oracle):
    import torch
    from nim_snappy_b200.workloads import mixed_corpus
    d = mixed_corpus(64 << 20, seed=2)
    st, expect = oracle.compress_framed(d, as_array=True)
    d_in = torch.from_numpy(d).cuda()
    lens = torch.zeros(d.size // 65536, dtype=torch.int32, device="cuda")
    out, n = dc.compress_framed(d_in, frame_lens=lens)
    assert n == expect.size and np.array_equal(out[:n].cpu().numpy(), expect)
    assert int(lens.sum().item()) + 10 == n                       # the length table shards gather
    crc = dc.masked_crc_chunks(d_in).cpu().numpy().view(np.uint32)
    assert crc[0] == oracle.masked_crc32c(d[:65536]) and crc[-1] == oracle.masked_crc32c(d[-65536:])
    chunks, walk = dc.frame_index(expect)
    assert walk.status == 0 and walk.written == d.size and walk.read == expect.size
    d_out = torch.zeros(d.size, dtype=torch.uint8, device="cuda")
    status = dc.uncompress_chunks(torch.from_numpy(expect.copy()).cuda(), chunks, d_out)
    assert not status.any() and np.array_equal(d_out.cpu().numpy(), d)


def test_device_framed_1gib_config(dc, oracle):
    """BASELINE config 2/3: 1 GiB text-like corpus, byte-identical stream, bit-exact decode."""
    import torch
    from nim_snappy_b200.workloads import text_corpus
    d = text_corpus(1 << 30)
    st, expect = oracle.compress_framed(d, as_array=True)
    assert st == 0
    d_in = torch.from_numpy(d).cuda()
    out, n = dc.compress_framed(d_in)
    assert n == expect.size
    assert bool((out[:n] == torch.from_numpy(expect).cuda()).all())
    chunks, walk = dc.frame_index(expect)
    assert len(chunks) == 16384 and walk.written == d.size
    d_out = torch.zeros(d.size, dtype=torch.uint8, device="cuda")
    status = dc.uncompress_chunks(out[:n], chunks, d_out)
    assert not status.any() and bool((d_out == d_in).all())
    # corrupt one byte inside chunk 1000's payload: exactly that chunk reports an error
    bad = out[:n].clone()
    pos = int(chunks[1000]["in_off"]) + 100
    bad[pos] ^= 0x40
    status = dc.uncompress_chunks(bad, chunks, d_out)
    assert status[1000] != 0 and not np.delete(status, 1000).any()


def test_device_message_batch(dc, oracle):
    """BASELINE config 4 (scaled to 50 000 messages for the oracle): per-message byte equality."""
    import torch
    from nim_snappy_b200.codec import maxCompressedLen
    from nim_snappy_b200.workloads import message_batch
    packed, offs, lens = message_batch(50000, seed=0)
    stride = (maxCompressedLen(4096) + 15) // 16 * 16
    d_in = torch.from_numpy(packed).cuda()
    d_off = torch.from_numpy(offs).cuda()
    d_len = torch.from_numpy(lens).cuda()
    out_off = torch.arange(len(lens), dtype=torch.int64, device="cuda") * stride
    d_out = torch.zeros(len(lens) * stride, dtype=torch.uint8, device="cuda")
    out_len = dc.compress_batch(d_in, d_off, d_len, 4096, d_out, out_off).cpu().numpy()
    h_out = d_out.cpu().numpy()
    for i in range(0, len(lens), 1):
        exp = oracle.encode(packed[offs[i]:offs[i] + lens[i]])
        assert out_len[i] == len(exp) and h_out[i * stride:i * stride + out_len[i]].tobytes() == exp, i
    # decode the oracle-identical streams back, batched
    ulen = dc.uncompressed_lens(d_out, out_off, torch.from_numpy(out_len).cuda()).cpu().numpy()
    assert np.array_equal(ulen, lens)
    back = torch.zeros(packed.size, dtype=torch.uint8, device="cuda")
    got_len, status = dc.uncompress_batch(d_out, out_off, torch.from_numpy(out_len).cuda(), back, d_off, d_len)
    assert not status.cpu().numpy().any() and np.array_equal(got_len.cpu().numpy(), lens)
    assert np.array_equal(back.cpu().numpy(), packed)
    # too-small output room -> bufferTooSmall for exactly those messages (snappy.nim:96-97)
    small_cap = d_len.clone()
    small_cap[::7] -= 1
    _, status = dc.uncompress_batch(d_out, out_off, torch.from_numpy(out_len).cuda(), back, d_off, small_cap)
    st = status.cpu().numpy()
    assert (st[::7] == 1).all() and not np.delete(st, np.arange(0, len(st), 7)).any()


def test_device_mixed_sizes_batch(dc, oracle):
    """Batch with every table-size class (256..16384 entries) and lengths 0..65536."""
    import torch
    from nim_snappy_b200.codec import maxCompressedLen
    rng = np.random.default_rng(9)
    src = np.frombuffer(read_data("html_x_4"), dtype=np.uint8)
    lens = np.array([0, 1, 16, 17, 255, 256, 257, 511, 512, 1024, 4096, 4097, 16384, 16385, 65535, 65536] * 4, dtype=np.int32)
    offs = rng.integers(0, src.size - 65536, len(lens)).astype(np.int64)
    stride = (maxCompressedLen(65536) + 15) // 16 * 16
    out_off = torch.arange(len(lens), dtype=torch.int64, device="cuda") * stride
    d_out = torch.zeros(len(lens) * stride, dtype=torch.uint8, device="cuda")
    out_len = dc.compress_batch(torch.from_numpy(src).cuda(), torch.from_numpy(offs).cuda(),
                                torch.from_numpy(lens).cuda(), 0, d_out, out_off).cpu().numpy()
    h = d_out.cpu().numpy()
    for i in range(len(lens)):
        exp = oracle.encode(src[offs[i]:offs[i] + lens[i]])
        assert h[i * stride:i * stride + out_len[i]].tobytes() == exp, (i, lens[i])


def test_raw_random_element_streams(gpu_api, oracle):
    """Valid raw streams that are not block structured (and their corruptions): discovery must hand them to the single-warp decoder."""
    P.check_random_element_streams(gpu_api, oracle, range(12), 400000)
    P.check_random_element_streams(gpu_api, oracle, range(100, 103), 3000000)


# ------------------------------------------------------------------ BASELINE configs 4 and 5 at their sizes
def test_config4_one_million_messages_host_batch(gpu_api, oracle):
    """BASELINE configs[3]: 1 M independent 1-4 KiB messages (about 2.5 GiB) through the HOST batch calls
    (snb_compress_batch / snb_uncompress_batch, packed output).  Oracle compare on a seeded 20 000-message sample
    (byte for byte, snappy.nim:27-64 per message) + full round trip of all 1 M on the GPU path."""
    import time
    from nim_snappy_b200.workloads import message_batch
    count = 1_000_000
    packed, offs, lens = message_batch(count, seed=0)
    t0 = time.perf_counter()
    out, out_len, status = gpu_api.compressBatch(packed, offs, lens)
    t_c = time.perf_counter() - t0
    assert not status.any()
    ends = np.cumsum(out_len, dtype=np.uint64)
    rng = np.random.default_rng(1)
    sample = np.concatenate([rng.choice(count, size=19990, replace=False), np.arange(5), np.arange(count - 5, count)])
    for i in sample:
        exp = oracle.encode(packed[offs[i]:offs[i] + lens[i]])
        e = int(ends[i])
        assert out_len[i] == len(exp) and out[e - len(exp):e].tobytes() == exp, int(i)
    total = int(ends[-1])
    t0 = time.perf_counter()
    back, back_len, st2 = gpu_api.uncompressBatch(out[:total], ends - out_len, out_len)
    t_d = time.perf_counter() - t0
    assert not st2.any() and np.array_equal(back_len, lens.astype(np.uint32))
    assert back.size == packed.size and np.array_equal(back, packed)
    print(f"config 4 through host (pageable) buffers: compress {packed.size / t_c / 1e9:.2f} GB/s, "
          f"uncompress {packed.size / t_d / 1e9:.2f} GB/s, ratio {total / packed.size:.3f}")


def test_config5_mixed_8gib_slab_loop(dc, oracle):
    """BASELINE configs[4] on one GPU: 8 GiB of the mixed corpus (text | protobuf-like | random per MiB) through the
    device slab loop (8 slabs of 16 384 chunks).  Sampled oracle compare of whole frames at every slab boundary and at
    random chunks (frame bytes, CRC and placement), stream size == 10 + sum of the frame lengths, full GPU round trip."""
    import torch
    from nim_snappy_b200.workloads import mixed_corpus
    gib, slabs = 1 << 30, 8
    d_in = torch.empty(slabs * gib, dtype=torch.uint8, device="cuda")
    host = {}
    for k in range(slabs):                                  # every slab different: seed and phase move with k
        h = mixed_corpus(gib, seed=100 + k, phase_mib=k)
        d_in[k * gib:(k + 1) * gib] = torch.from_numpy(h).cuda()
        if k in (0, 3, 7):
            host[k] = h
        else:
            host[k] = np.concatenate([h[:3 * 65536], h[-3 * 65536:]])      # the chunks next to the slab boundaries
    n = slabs * gib
    nchunks = n // 65536
    lens = torch.zeros(nchunks, dtype=torch.int32, device="cuda")
    out, w = dc.compress_framed(d_in, frame_lens=lens)
    fl = lens.cpu().numpy().astype(np.int64)
    assert w == 10 + int(fl.sum())
    starts = 10 + np.concatenate([[0], np.cumsum(fl)[:-1]])
    rng = np.random.default_rng(7)
    sample = set()
    for k in range(slabs):
        sample.update(range(k * 16384, k * 16384 + 3))
        sample.update(range((k + 1) * 16384 - 3, (k + 1) * 16384))
    for k in (0, 3, 7):
        sample.update(int(k * 16384 + c) for c in rng.integers(0, 16384, 400))
    for c in sorted(sample):
        k, j = divmod(c, 16384)
        h = host[k]
        if h.size == gib:
            chunk = h[j * 65536:(j + 1) * 65536]
        else:
            chunk = h[j * 65536:(j + 1) * 65536] if j < 3 else h[(j - 16384 + 6) * 65536:(j - 16384 + 7) * 65536]
        exp = oracle.encode_framed(chunk)[10:]                 # one frame: header, masked CRC, payload
        got = out[int(starts[c]):int(starts[c]) + int(fl[c])].cpu().numpy().tobytes()
        assert got == exp, c
    # full round trip on the device path (header walk of the host copy of the stream; see snb_dev_uncompress_framed)
    stream_h = out[:w].cpu().numpy()
    chunks, walk = dc.frame_index(stream_h)
    assert len(chunks) == nchunks and walk.status == 0 and walk.written == n
    d_back = torch.empty(n, dtype=torch.uint8, device="cuda")
    status = dc.uncompress_chunks(out[:w], chunks, d_back)
    assert not status.any()
    for k in range(slabs):
        assert bool((d_back[k * gib:(k + 1) * gib] == d_in[k * gib:(k + 1) * gib]).all()), k


def test_device_side_header_walk(dc, gpu_api, oracle):
    """snb_dev_frame_index / snb_dev_uncompress_framed: the framed header walk on the device."""
    import torch
    from nim_snappy_b200.workloads import mixed_corpus, text_corpus
    rng = np.random.default_rng(23)
    alice = read_data("alice29.txt")
    datas = [b"", b"abc", alice, read_data("house.jpg") + read_data("html"), bytes(200000),
             mixed_corpus(5 << 20, seed=3).tobytes()]
    P.check_dev_framed(P.gpu_dev_call(dc), oracle, datas, rng, fuzz=40, lib=gpu_api.lib)
    # the stream a device-resident caller has: produced by snb_dev_compress_framed, never on the host
    d = text_corpus(256 << 20)
    d_in = torch.from_numpy(d).cuda()
    lens = torch.zeros(d.size // 65536, dtype=torch.int32, device="cuda")
    out, n = dc.compress_framed(d_in, frame_lens=lens)
    d_back = torch.zeros(d.size, dtype=torch.uint8, device="cuda")
    f0 = gpu_api.lib.snb_frame_index_fallbacks()
    for fl in (None, lens):
        d_back.zero_()
        assert dc.uncompress_framed(out[:n], d_back, frame_lens=fl) == (0, n, d.size)
        assert bool((d_back == d_in).all())
    d_chunks, walk = dc.dev_frame_index(out[:n])
    h_chunks, hwalk = dc.frame_index(out[:n].cpu().numpy())
    assert (walk.status, walk.read, walk.written) == (hwalk.status, hwalk.read, hwalk.written)
    assert d_chunks.cpu().numpy().tobytes() == h_chunks.tobytes()          # the same descriptors as the host walk
    assert gpu_api.lib.snb_frame_index_fallbacks() == f0
    bad = out[:n].clone()
    bad[int(h_chunks[2000]["in_off"]) + 50] ^= 1                           # payload damage: that chunk's CRC / tags fail
    st, rd, wr = dc.uncompress_framed(bad, d_back)
    assert st in (2, 3) and rd == int(h_chunks[2000]["in_off"]) - 8 and wr == 2000 * 65536
    bad = out[:n].clone()
    bad[int(h_chunks[3000]["in_off"]) - 8] = 0x7f                          # header damage: unknownChunk after 3000 chunks
    ref = oracle.uncompress_framed(bad.cpu().numpy().tobytes(), d.size)
    assert dc.uncompress_framed(bad, d_back)[0] == ref[0] == 4


def test_concurrent_callers_and_streams(gpu_api, dc, oracle):
    """Four host threads call snb_compress_framed / snb_uncompress_framed / the batch calls at once while a fifth
    issues snb_dev_* calls without synchronising on two different CUDA streams: every result byte-exact (the library
    serialises the host calls and orders the device calls' use of its shared work buffers across streams)."""
    import threading
    import torch
    from nim_snappy_b200.workloads import mixed_corpus, text_corpus
    datas = [text_corpus(20 << 20, phase=i * 1000).tobytes() for i in range(2)] + [mixed_corpus(12 << 20, seed=i).tobytes() for i in range(2)]
    expect = [oracle.encode_framed(d) for d in datas]
    errors = []

    def host_worker(i):
        try:
            for _ in range(4):
                assert gpu_api.encodeFramed(datas[i]) == expect[i]
                assert gpu_api.decodeFramed(expect[i]) == datas[i]
                msgs = [datas[i][k * 3000:(k + 1) * 3000] for k in range(50)]
                assert gpu_api.encodeBatch(msgs) == [oracle.encode(m) for m in msgs]
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    def dev_worker():
        try:
            streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            d_ins = [torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda() for d in datas]
            outs = [torch.empty(oracle.max_compressed_len_framed(len(d)), dtype=torch.uint8, device="cuda") for d in datas]
            torch.cuda.synchronize()
            for rep in range(6):
                counts = []
                for j in range(len(datas)):
                    with torch.cuda.stream(streams[(j + rep) & 1]):
                        counts.append(dc.compress_framed(d_ins[j], out=outs[j], sync=False)[1])   # no host sync between calls
                torch.cuda.synchronize()
                for j in range(len(datas)):
                    n = int(counts[j].item())
                    assert n == len(expect[j]) and outs[j][:n].cpu().numpy().tobytes() == expect[j], (rep, j)
        except Exception as e:  # noqa: BLE001
            errors.append(("dev", repr(e)))

    ts = [threading.Thread(target=host_worker, args=(i,)) for i in range(4)] + [threading.Thread(target=dev_worker)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_streaming_compress_survives_allocator_traffic(gpu_api, oracle):
    """A pageable input above 64 MiB is copied while the streaming kernel already runs; cudaMalloc / cudaFree from another
    thread waits for that kernel and can keep the copying thread out of the driver.  The kernel then gives up after its
    timeout and the call is repeated with the copies enqueued first: the result is byte-exact either way, never
    SNB_TIMEOUT (include/snb.h, snb_set_arrive_timeout_ms)."""
    import threading
    import torch
    from nim_snappy_b200.workloads import text_corpus
    data = text_corpus(96 << 20, phase=77).tobytes()
    expect = oracle.encode_framed(data)
    lib = gpu_api.lib
    stop = threading.Event()

    def churn():
        while not stop.is_set():
            t = torch.empty(48 << 20, dtype=torch.uint8, device="cuda")
            del t
            torch.cuda.empty_cache()      # cudaFree: device-wide synchronisation point

    lib.snb_set_arrive_timeout_ms(300)    # keep a recovered attempt short
    th = threading.Thread(target=churn)
    th.start()
    try:
        for _ in range(6):
            assert gpu_api.encodeFramed(data) == expect
    finally:
        stop.set()
        th.join()
        lib.snb_set_arrive_timeout_ms(2000)
    print("timeouts recovered:", lib.snb_timeouts_recovered())
