"""One 2^log_n forward NTT on device data, twice (the first call builds the plan's tables): target of the ncu capture of the
NTT passes.  usage: python tools/ntt_probe.py [log_n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_b200

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
be = snark_b200.Backend(curve=0, device=0)
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.randint(-(1 << 31), (1 << 31) - 1, (1 << log_n, 8), dtype=torch.int32, device="cuda", generator=g)
x[:, 7] &= 0x1FFFFFFF
for _ in range(2):
    be.ntt(x, log_n)
    be.sync()
print("launches", be.launches)
be.close()
