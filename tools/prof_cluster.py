import sys
import numpy as np
sys.path.insert(0, ".")
from pyranges_b200 import engine as E, synthetic as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
iv = S.reads(n, seed=3, min_len=50, max_len=500)
q = E.DeviceIntervals.from_numpy(iv.starts, iv.ends, iv.seg_off)
q.bounds
for _ in range(2):
    E.cluster(q)
import torch
torch.cuda.synchronize()
