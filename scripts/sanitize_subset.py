#!/usr/bin/env python
"""The reduced parity subset that runs under compute-sanitizer (scripts/gpu_sanitize.sh): alice29 raw + framed both ways,
4 MiB of the mixed corpus, the reference's malformed vectors, one multi-slab host call, a message batch, the device-side
header walk, and the experimental v6 encoder.  Every result is compared with the oracle; exits non-zero on a mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import nim_snappy_b200 as snappy
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus, text_corpus
from oracle import oracle as o
import parity_cases as P
import vectors as V

small = len(sys.argv) > 1 and sys.argv[1] == "small"      # racecheck: shared-memory hazards only, much slower
lib = _lib.default()
api = snappy.Snappy(lib)
dc = DeviceCodec(lib)
alice = open(os.path.join(ROOT, "tests", "data", "alice29.txt"), "rb").read()
P.check_raw(api, o, alice)
P.check_framed(api, o, alice)
mixed = mixed_corpus((1 if small else 4) << 20, seed=5).tobytes()
P.check_framed(api, o, mixed)
P.check_malformed(api, o)
big = text_corpus((70 if small else 150) << 20)                       # more than one 64 MiB decode slab
enc = o.encode_framed(big.tobytes())
assert api.encodeFramed(big) == enc and api.decodeFramed(enc) == big.tobytes()
msgs = [alice[i * 1500:(i + 1) * 1500 + 700] for i in range(60)]
P.check_batch(api, o, msgs, extra_streams=V.MALFORMED_RAW[:6])
rng = np.random.default_rng(3)
P.check_dev_framed(P.gpu_dev_call(dc), o, [alice, mixed[: 1 << 20]], rng, fuzz=6, lib=lib)
lib.snb_set_compress_algo(6)                                             # resolver / walker kernel
P.check_framed(api, o, alice + mixed[: 1 << 20])
lib.snb_set_compress_algo(0)
torch.cuda.synchronize()
print("sanitize subset ok")
