"""Rep3 mul_vec on large share vectors, three parties on three GPUs of one box:
  fused : cs_rep3_mul_vec_reshare storing the b-halves straight into the next party's vector (NVLink peer memory)
  nccl  : the same arithmetic kernel without the peer store, then gather + NCCL send/recv + cs_rep3_set_b
Both must produce bit-identical share vectors that open to x*y.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29517 \
        tools/time_mul_vec.py [log_n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from co_snarks_b200 import binding as B
from co_snarks_b200.rep3 import Rep3MulVec, Rep3Network

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
rank = int(os.environ["RANK"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
assert dist.get_world_size() == 3
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = B.Context(local, stream=stream.cuda_stream)
net = Rep3Network(device="cuda")
mv = Rep3MulVec(ctx, net)
nxt, prv = (rank + 1) % 3, (rank + 2) % 3
seeds = [bytes((11 * p + i) & 0xff for i in range(32)) for p in range(3)]
prf = (seeds[rank], 0, seeds[prv], 0, 12)
sizes = [int(a) for a in sys.argv[1:]] or [20, 22, 24]
REPS = 5


def timed(fn):
    ms = []
    for i in range(2 + REPS):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= 2:
            ms.append(float(t.item()))
    return sum(ms) / len(ms)


out = {"world": 3, "reps": REPS}
for lg in sizes:
    n = 1 << lg
    # canonical 253-bit values < r used directly as Montgomery limbs: every rank draws all three shares
    g = np.random.Generator(np.random.PCG64(100 + lg))
    sh = g.integers(0, 2 ** 63, size=(2, 3, n, 4), dtype=np.uint64)
    sh[..., 3] &= np.uint64((1 << 60) - 1)

    def mine(v):
        a = np.empty((n, 2, 4), dtype=np.uint64)
        a[:, 0, :] = v[rank]
        a[:, 1, :] = v[prv]
        return a
    d_a, d_b = ctx.to_device(mine(sh[0])), ctx.to_device(mine(sh[1]))
    d_out = ctx.alloc(n * 64)
    mv.connect(d_out)
    peer = mv._peers[d_out]
    out_t = torch.zeros(n * 8, dtype=torch.int64, device="cuda")
    z_t = torch.empty(n * 4, dtype=torch.int64, device="cuda")
    recv_t = torch.empty(n * 4, dtype=torch.int64, device="cuda")

    def fused():
        ctx.rep3_mul_vec_reshare(B.CS_BN254, d_a, d_b, n, prf, d_out, peer)

    def local_only():
        ctx.rep3_mul_vec_reshare(B.CS_BN254, d_a, d_b, n, prf, out_t.data_ptr(), None)

    def nccl():
        local_only()
        z_t.view(n, 4).copy_(out_t.view(n, 2, 4)[:, 0, :])
        ops = [dist.P2POp(dist.isend, z_t, nxt), dist.P2POp(dist.irecv, recv_t, prv)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        ctx.rep3_set_b(B.CS_BN254, recv_t.data_ptr(), n, out_t.data_ptr())

    t_fused = timed(fused)
    t_local = timed(local_only)
    t_nccl = timed(nccl)
    torch.cuda.synchronize()
    dist.barrier()
    a = ctx.d2h(d_out, (n, 2, 4))
    b = out_t.cpu().numpy().view(np.uint64).reshape(n, 2, 4)
    same = bool((a == b).all())
    # opening check on a sample: sum of the three a-halves == x*y
    k = 512
    t = torch.from_numpy(a[:k, 0, :].copy().view(np.int64)).cuda()
    outs = [torch.empty_like(t) for _ in range(3)]
    dist.all_gather(outs, t)
    ok = True
    if rank == 0:
        parts = [o.cpu().numpy().view(np.uint64) for o in outs]
        lim = lambda arr, i: sum(int(arr[i, j]) << (64 * j) for j in range(4))
        rinv = pow(1 << 256, -1, R)
        for i in range(k):
            x = sum(lim(sh[0, p], i) for p in range(3)) % R
            y = sum(lim(sh[1, p], i) for p in range(3)) % R
            zsum = sum(lim(parts[p], i) for p in range(3)) % R
            # limbs are Montgomery representations: z_mont = x_mont*y_mont/R
            ok &= zsum == x * y * rinv % R
    out["2p%d" % lg] = {"fused_ms": round(t_fused, 4), "kernel_without_peer_store_ms": round(t_local, 4),
                        "nccl_staged_ms": round(t_nccl, 4), "identical": same, "opens_to_xy": bool(ok),
                        "payload_MB": n * 32 / 1e6, "fused_payload_GBs": round(n * 32 / (t_fused * 1e-3) / 1e9, 1)}
    mv.disconnect(d_out)
    for d in (d_a, d_b, d_out):
        ctx.free(d)
    del out_t, z_t, recv_t
if rank == 0:
    print(json.dumps(out))
dist.barrier()
ctx.close()
dist.destroy_process_group()
