import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus
lib=_lib.default(); dc=DeviceCodec(lib)
n=(int(sys.argv[1]) if len(sys.argv)>1 else 64)<<20
h=mixed_corpus(n, seed=0)
d_in=torch.from_numpy(h).cuda()
out,w=dc.compress_framed(d_in)
d_back=torch.empty(n,dtype=torch.uint8,device='cuda')
print(dc.uncompress_framed(out[:w], d_back), w, lib.snb_frame_index_fallbacks())
print(bool((d_back==d_in).all()))
