import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from nim_snappy_b200 import _lib
from nim_snappy_b200.device import DeviceCodec
from nim_snappy_b200.workloads import mixed_corpus
from oracle import oracle as o
import parity_cases as P
lib=_lib.default(); dc=DeviceCodec(lib); call=P.gpu_dev_call(dc)
rd=lambda n: open('/root/repo/tests/data/'+n,'rb').read()
datas={'abc':b'abc','alice':rd('alice29.txt'),'house+html':rd('house.jpg')+rd('html'),'mixed5':mixed_corpus(5<<20,seed=3).tobytes()}
for k,d in datas.items():
    s=o.encode_framed(d); n=len(d)
    for name,args in (('full',(s,n)),('n-1',(s,max(n-1,0))),('skip',(s[:10]+b"\xfe\x03\x00\x00abc"+s[10:],n))):
        f0=lib.snb_frame_index_fallbacks()
        r=call(*args)
        print(k,name,r[:3],'fallback' if lib.snb_frame_index_fallbacks()!=f0 else 'device', flush=True)
