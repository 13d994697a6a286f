"""CPU-only, world_size 3 over gloo: the Rep3 co-Groth16 party driver (co_snarks_b200/rep3.py) end to end.
Three processes = three parties, each running its local phase on the CPU emulation of the kernels
(tests/emu) and exchanging the reference's four point-sized messages through torch.distributed.
Mirrors tests/tests/circom/e2e_tests/rep3.rs:36-137 of the reference: all parties return the same
proof and it verifies; additionally it must equal the plain proof for r = sum r_i.a, s = sum s_i.a."""
import os
import random
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _party(rank, port, emu_path, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from co_snarks_b200 import binding as B
    from co_snarks_b200.rep3 import Rep3CoGroth16, Rep3Network, Rep3State
    from helpers import Conv, golden_groth16, make_key
    from oracle import groth16 as OG
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=3)
    try:
        ctx = B.Context(0, lib_path=emu_path)
        cv = Conv("bn254")
        z, m, w, g = golden_groth16("multiplier2")
        ni = m["num_instance_variables"]
        pk = make_key(ctx, cv, z, m)
        wsh = OG.share_rep3(w[ni:], cv.r, random.Random(5))  # same seed everywhere -> consistent sharing
        mine = cv.fr([x for ab in wsh[rank] for x in ab]).reshape(-1, 8)
        net = Rep3Network()
        state = Rep3State(net, seed=1000 + rank)
        prover = Rep3CoGroth16(ctx, pk)
        A, Bp, Cp = prover.prove(net, state, cv.fr(w[:ni]), mine, cv.g1([z["delta_g1"]])[0])
        r_sh, s_sh = prover.last_randomness
        q.put((rank, cv.pt1(A), cv.pt2(Bp), cv.pt1(Cp), cv.fr_back(r_sh), cv.fr_back(s_sh), net.bytes_sent))
        pk.free()
        ctx.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_rep3_three_parties_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu = build_emu.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_party, args=(r, port, emu, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import golden_groth16, ih
    from oracle import groth16 as OG
    from oracle.fields import BN254
    from oracle.pairing_bn254 import groth16_verify
    proofs = [(a, b, c) for _, a, b, c, _, _, _ in res]
    assert proofs[0] == proofs[1] == proofs[2], "parties disagree on the proof"
    z, m, w, g = golden_groth16("multiplier2")
    assert groth16_verify(OG.vk_from_zkey(z), [ih(x) for x in g["public"]], proofs[0])
    # replicated randomness is consistent (party i's b == party i-1's a) and the proof is the plain one
    for i in range(3):
        assert res[i][4][1] == res[(i + 2) % 3][4][0] and res[i][5][1] == res[(i + 2) % 3][5][0]
    r_tot = sum(x[4][0] for x in res) % BN254.r
    s_tot = sum(x[5][0] for x in res) % BN254.r
    assert proofs[0] == OG.prove_plain(z, m, w, r_tot, s_tot)
    # the reference exchanges only point-sized messages on this path
    assert all(x[6] < 4096 for x in res)


def _party6(rank, port, emu_path, q):
    """world 6 = 3 parties x 2 GPUs: even ranks run {A, B1, L} + the protocol, odd ranks {witness map -> H, B2}."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from co_snarks_b200 import binding as B
    from co_snarks_b200.rep3 import PairLink, Rep3CoGroth16, Rep3Network, Rep3State
    from helpers import Conv, golden_groth16, make_key
    from oracle import groth16 as OG
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=6)
    try:
        mains = dist.new_group([0, 2, 4])
        helpers = dist.new_group([1, 3, 5])
        party, role = rank // 2, rank % 2
        ctx = B.Context(0, lib_path=emu_path)
        cv = Conv("bn254")
        z, m, w, g = golden_groth16("poseidon")
        ni = m["num_instance_variables"]
        pk = make_key(ctx, cv, z, m)
        wsh = OG.share_rep3(w[ni:], cv.r, random.Random(5))
        mine = cv.fr([x for ab in wsh[party] for x in ab]).reshape(-1, 8)
        net = Rep3Network(mains if role == 0 else helpers)
        assert net.id == party
        state = Rep3State(net, seed=1000 + party)  # both GPUs of a party derive identical streams
        prover = Rep3CoGroth16(ctx, pk)
        link = PairLink(rank + 1 if role == 0 else rank - 1)
        pub = cv.fr(w[:ni])
        if role == 0:
            A, Bp, Cp = prover.prove(net, state, pub, mine, cv.g1([z["delta_g1"]])[0],
                                     pair_recv=lambda: link.recv(4 * pk.fq + 2 * pk.fq))
            r_sh, s_sh = prover.last_randomness
            q.put((party, cv.pt1(A), cv.pt2(Bp), cv.pt1(Cp), cv.fr_back(r_sh), cv.fr_back(s_sh)))
        else:
            prover.helper_step(party, state, pub, mine, link.send)
        pk.free()
        ctx.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_rep3_two_gpus_per_party_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu = build_emu.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_party6, args=(r, port, emu, q)) for r in range(6)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(3))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import golden_groth16, ih
    from oracle import groth16 as OG
    from oracle.fields import BN254
    from oracle.pairing_bn254 import groth16_verify
    proofs = [(a, b, c) for _, a, b, c, _, _ in res]
    assert proofs[0] == proofs[1] == proofs[2]
    z, m, w, g = golden_groth16("poseidon")
    assert groth16_verify(OG.vk_from_zkey(z), [ih(x) for x in g["public"]], proofs[0])
    r_tot = sum(x[4][0] for x in res) % BN254.r
    s_tot = sum(x[5][0] for x in res) % BN254.r
    assert proofs[0] == OG.prove_plain(z, m, w, r_tot, s_tot)


def _party_mul(rank, port, emu_path, q, fused=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from co_snarks_b200 import binding as B
    from co_snarks_b200.rep3 import Rep3MulVec, Rep3Network, Rep3State
    from helpers import Conv
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=3)
    try:
        if fused:  # real GPUs: one process per party, on distinct GPUs when the box has three
            import torch
            ctx = B.Context(rank % torch.cuda.device_count())
        else:
            ctx = B.Context(0, lib_path=emu_path)
        cv = Conv("bn254")
        n = 70 if not fused else 5000
        rng = random.Random(21)  # same seed everywhere -> consistent sharings of x and y
        xs = [rng.randrange(cv.r) for _ in range(n)]
        ys = [rng.randrange(cv.r) for _ in range(n)]

        def mine(vals):
            out = []
            for v in vals:
                s0, s1 = rng.randrange(cv.r), rng.randrange(cv.r)
                sh = [s0, s1, (v - s0 - s1) % cv.r]
                out += [sh[rank], sh[(rank + 2) % 3]]
            return cv.fr(out)
        d_a, d_b = ctx.to_device(mine(xs)), ctx.to_device(mine(ys))
        d_out = ctx.alloc(n * 64)
        net = Rep3Network()
        state = Rep3State(net, seed=2000 + rank)
        pos0 = state.prf_args()[1]
        mv = Rep3MulVec(ctx, net)
        if fused:
            mv.connect(d_out)
            mv.mul_vec_fused(state, d_a, d_b, n, d_out)
        else:
            mv.mul_vec_staged(state, d_a, d_b, n, d_out)
        assert state.prf_args()[1] == pos0 + 8 * n
        q.put((rank, cv.fr_back(ctx.d2h(d_out, (2 * n, 4))), xs, ys))
        if fused:
            mv.disconnect(d_out)
        for d in (d_a, d_b, d_out):
            ctx.free(d)
        ctx.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.gpu
def test_rep3_mul_vec_fused_peer_memory():
    """The fused kernel through a real CUDA IPC mapping: three processes (three GPUs, or one shared GPU), each
    storing its products into the next process's share vector."""
    _run_mul_vec(fused=True)


def test_rep3_mul_vec_staged_gloo():
    """mul_vec (arithmetic.rs:165-176) across three processes: shares stay replicated (b_i == a_{i-1}) and
    open to x*y, i.e. the on-device ChaCha masks of the three parties cancel."""
    _run_mul_vec(fused=False)


def _run_mul_vec(fused):
    import torch.multiprocessing as mp
    emu = None
    if not fused:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        emu = build_emu.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_party_mul, args=(r, port, emu, q, fused)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle.fields import BN254
    r = BN254.r
    xs, ys = res[0][2], res[0][3]
    sh = [x[1] for x in res]
    for i in range(3):
        assert sh[i][1::2] == sh[(i + 2) % 3][0::2]
    assert [(sh[0][2 * k] + sh[1][2 * k] + sh[2][2 * k]) % r for k in range(len(xs))] == [x * y % r for x, y in zip(xs, ys)]


def _party_plonk(rank, port, emu_path, q, peer):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from co_snarks_b200 import binding as B
    from co_snarks_b200.plonk import DistRep3Comm, Rep3CoPlonk
    from co_snarks_b200.rep3 import Rep3Network, Rep3State
    from helpers import Conv, golden_plonk, make_plonk_key, plonk_proof_from_device
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=3)
    try:
        if peer:
            import torch
            ctx = B.Context(rank % torch.cuda.device_count())
        else:
            ctx = B.Context(0, lib_path=emu_path)
        cv = Conv("bn254")
        z, w, g = golden_plonk("multiplier2")
        npub = z["n_public"]
        pk = make_plonk_key(ctx, cv, z)
        rng = random.Random(41)  # same seed everywhere -> consistent sharings

        def mine(vals):
            out = []
            for v in vals:
                s0, s1 = rng.randrange(cv.r), rng.randrange(cv.r)
                sh = [s0, s1, (v - s0 - s1) % cv.r]
                out += [sh[rank], sh[(rank + 2) % 3]]
            return cv.fr(out).reshape(-1, 2, 4)
        wsh = mine(w[npub + 1:])
        bsh = mine(list(range(11)))  # the reference's deterministic blinders, secret-shared
        net = Rep3Network()
        state = Rep3State(net, seed=3000 + rank)
        prover = Rep3CoPlonk(ctx, pk, rank)
        comm = DistRep3Comm(prover, net, peer=peer)
        vkp = cv.g1([z["vk_" + k] for k in ("qm", "ql", "qr", "qo", "qc", "s1", "s2", "s3")])
        pts, evs = comm.run(prover.prove(state, cv.fr(w[:npub + 1]), wsh, vkp, z["domain_size"], bsh))
        comm.close()
        q.put((rank, plonk_proof_from_device(cv, pts, evs), net.bytes_sent))
        prover.free()
        pk.free()
        ctx.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run_plonk(peer):
    import torch.multiprocessing as mp
    emu = None
    if not peer:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        emu = build_emu.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_party_plonk, args=(r, port, emu, q, peer)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import golden_plonk
    from oracle.formats import plonk_proof_to_json
    _, _, g = golden_plonk("multiplier2")
    assert res[0][1] == res[1][1] == res[2][1], "parties disagree on the proof"
    # shared deterministic blinders -> the reference's known answers (co-plonk/src/round{1..5}.rs tests)
    assert plonk_proof_to_json(res[0][1]) == g["oracle_proof_json"]


def test_co_plonk_rep3_three_processes_gloo():
    """Rep3CoPlonk::prove across three processes (tests/tests/circom/e2e_tests/rep3.rs:36-137 for Plonk): the
    products travel through the network (staged reshare), every party opens the reference's known-answer proof."""
    _run_plonk(peer=False)


@pytest.mark.gpu
def test_co_plonk_rep3_peer_memory():
    """The same with one process per party on real GPUs and the arena of the next party mapped through CUDA IPC."""
    _run_plonk(peer=True)


def _party_shamir(rank, port, emu_path, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from co_snarks_b200 import binding as B
    from co_snarks_b200.shamir import ShamirCoGroth16, ShamirNetwork
    from helpers import Conv, golden_groth16, make_key
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=3)
    try:
        ctx = B.Context(0, lib_path=emu_path)
        cv = Conv("bn254")
        z, m, w, g = golden_groth16("multiplier2")
        ni = m["num_instance_variables"]
        pk = make_key(ctx, cv, z, m)
        rng = random.Random(77)  # same seed everywhere -> consistent degree-1 sharings of the witness
        mine = []
        for v in w[ni:]:
            a = rng.randrange(cv.r)
            mine.append((v + a * (rank + 1)) % cv.r)
        net = ShamirNetwork()
        prover = ShamirCoGroth16(ctx, pk)
        A, Bp, Cp = prover.prove(net, cv.fr(w[:ni]), cv.fr(mine), cv.g1([z["delta_g1"]])[0])
        # privacy of the dealing: from every dealer this party holds ONE evaluation per sharing (its own row);
        # two would determine the degree-1 polynomial and with it r(0), s(0)
        view = net.last_received
        assert len(view) == 3 and all(v.shape == (2, 4) for v in view)
        q.put((rank, cv.pt1(A), cv.pt2(Bp), cv.pt1(Cp), prover.last_randomness, net.bytes_sent))
        pk.free()
        ctx.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_shamir_co_groth16_three_parties_gloo():
    """ShamirCoGroth16::prove for t = 1, n = 3 (tests/tests/circom/e2e_tests/shamir.rs:37-91): all parties open the
    same proof, it verifies, and it is the plain proof for r = r(0), s = s(0) of the jointly drawn sharings."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu = build_emu.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_party_shamir, args=(r, port, emu, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from co_snarks_b200.shamir import lagrange_at_zero
    from helpers import golden_groth16, ih
    from oracle import groth16 as OG
    from oracle.fields import BN254
    from oracle.pairing_bn254 import groth16_verify
    proofs = [(a, b, c) for _, a, b, c, _, _ in res]
    assert proofs[0] == proofs[1] == proofs[2], "parties disagree on the proof"
    z, m, w, g = golden_groth16("multiplier2")
    assert groth16_verify(OG.vk_from_zkey(z), [ih(x) for x in g["public"]], proofs[0])
    lam = lagrange_at_zero(3)
    assert lam == [3, BN254.r - 3, 1]
    r0 = sum(l * x[4][0] for l, x in zip(lam, res)) % BN254.r
    s0 = sum(l * x[4][1] for l, x in zip(lam, res)) % BN254.r
    assert proofs[0] == OG.prove_plain(z, m, w, r0, s0)
    assert all(x[5] < 4096 for x in res)  # point- and scalar-sized messages only
