"""Times resident NTTs (one inverse + one forward, in place) at several sizes and batch widths.
usage: time_ntt.py [lg:batch ...]   default 20:1 20:2 22:2 24:1
Tuning hooks CS_NTT_TWS / CS_NTT_THREADS are read by the library once per process."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from co_snarks_b200 import binding as B

cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(20, 1), (20, 2), (22, 2), (24, 1)]
stream = torch.cuda.Stream()
ctx = B.Context(0, stream=stream.cuda_stream)
rng = np.random.Generator(np.random.PCG64(4))
res = {"tws": os.environ.get("CS_NTT_TWS", "default"), "threads": os.environ.get("CS_NTT_THREADS", "default")}
for lg, batch in cases:
    n = 1 << lg
    a = rng.integers(0, 2 ** 63, size=(n * batch, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    dom = ctx.domain(B.CS_BN254, lg, ctx.roots_of_unity(B.CS_BN254, lg)[0])
    d = ctx.to_device(a)
    ms = []
    for i in range(7):
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dom.ifft_in_to_out(d, batch)
            dom.fft_out_to_in(d, batch)
            e1.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1) / 2)
    back = ctx.d2h(d, (n * batch, 4))
    assert (back == a).all(), "round trip failed"
    t = sum(ms) / len(ms)
    res["2p%d_b%d_ms" % (lg, batch)] = round(t, 4)
    res["2p%d_b%d_gbs" % (lg, batch)] = round(64.0 * n * batch / (t * 1e-3) / 1e9, 1)
    res["2p%d_b%d_gmul" % (lg, batch)] = round(n * batch * lg / 2 / (t * 1e-3) / 1e9, 2)
    ctx.free(d)
    dom.free()
print(json.dumps(res))
