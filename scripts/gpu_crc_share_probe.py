#!/usr/bin/env python
"""How much of the decode kernel is the CRC verification: snb_dev_uncompress_framed with and without checkIntegrity."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import text_corpus
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = _lib.default(); dc = DeviceCodec(lib)
d_in = torch.from_numpy(text_corpus(mib << 20)).cuda()
out, w = dc.compress_framed(d_in)
s = out[:w].clone()
d_back = torch.empty_like(d_in)
dc.set_profiling(True)
for ci in (True, False, True, False):
    ms = []
    for i in range(5):
        st, rd, wr = dc.uncompress_framed(s, d_back, check_integrity=ci)
        torch.cuda.synchronize()
        ms.append(dc.profile()["decompress_ms"])
    print(f"check_integrity={ci}: kernel min {min(ms):.3f} ms, status {st}", flush=True)
