"""Rep3CoPlonk::prove INSIDE the library (cs_plonk_rep3_prove): step sequence, Keccak transcript and openings in C++
over cs_net (co-plonk/src/lib.rs:222-240, prove_inner :80-115).

Three party threads over in-process mailbox nets, each with its own context, key and device session:
 * peer mode   -- products stored into the next party's arena by the kernels, n-sized openings by reading the peers'
                  out-vectors (cs_plonk_rep3_connect / _connect_io): what three GPUs of one box run;
 * staged mode -- nothing connected: a-halves and opened vectors travel through the net (parties on different hosts).
With the blinder shares of b = [0..11) the opened proof is the reference's known answer (round 1-5 KATs through
oracle/plonk.py); with blinders drawn from the correlated streams it verifies.  CPU: emulation build; GPU: the real
library."""
import os
import random
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emu_factory():
    from co_snarks_b200 import binding as B
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    emu = build_emu.build()
    return lambda: B.Context(0, lib_path=emu)


def _run(mk, name="multiplier2", peer=True, fixed_blinders=True, seed=23):
    from co_snarks_b200 import binding as B
    from helpers import Conv, golden_plonk, ih, make_plonk_key, plonk_proof_from_device, plonk_vk_from_zkey
    from oracle import plonk as OP
    from oracle.fields import BN254
    from oracle.formats import plonk_proof_to_json
    from oracle.pairing_bn254 import pairing_product_is_one
    cv = Conv("bn254")
    r = cv.r
    z, w, g = golden_plonk(name)
    npub = z["n_public"]
    pub = cv.fr(w[:npub + 1])
    rng = random.Random(seed)

    def share(vals):
        out = [[], [], []]
        for v in vals:
            s0, s1 = rng.randrange(r), rng.randrange(r)
            sh = [s0, s1, (v - s0 - s1) % r]
            for p in range(3):
                out[p] += [sh[p], sh[(p + 2) % 3]]  # party p holds (x_p, x_{p-1})  rep3.rs:281-293
        return [cv.fr(o).reshape(-1, 2, 4) for o in out]
    wsh = share(w[npub + 1:])
    bsh = share(list(range(11))) if fixed_blinders else [None] * 3
    ctxs = [mk() for _ in range(3)]
    lib = ctxs[0].lib
    pks = [make_plonk_key(c, cv, z) for c in ctxs]
    sess = [B.PlonkRep3Session(ctxs[p], pks[p], p) for p in range(3)]
    nets = [B.Net.peer(ctxs[p], p, 3) for p in range(3)]
    for n in nets:
        n.connect_local(nets)
    if peer:
        for p in range(3):
            sess[p].connect(sess[(p + 1) % 3].arena)
            sess[p].connect_io(sess[(p + 2) % 3].d_out, sess[(p + 1) % 3].d_out)
    seeds = [bytes((31 * p + i) & 0xff for i in range(32)) for p in range(3)]
    states = [B.Rep3StateC.from_seeds(lib, p, seeds[p], seeds[(p + 2) % 3]) for p in range(3)]
    res, errs = {}, []

    def party(p):
        try:
            res[p] = sess[p].prove(nets[p], states[p], pub, wsh[p], bsh[p])
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=party, args=(p,)) for p in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not errs, errs
    proofs = [plonk_proof_from_device(cv, *res[p]) for p in range(3)]
    assert proofs[0] == proofs[1] == proofs[2], "parties disagree on the proof"
    if fixed_blinders:
        assert plonk_proof_to_json(proofs[0]) == g["oracle_proof_json"]
    assert OP.verify(BN254, plonk_vk_from_zkey(z, g["vk_power"]), proofs[0], [ih(x) for x in g["public"]], pairing_product_is_one)
    # consistent PRF consumption: party p's own stream is party p+1's "previous" stream
    pos = [(s.prf()[1], s.prf()[3]) for s in states]
    assert all(pos[p][0] == pos[(p + 1) % 3][1] and pos[p][0] > 0 for p in range(3))
    sent = [n.bytes_sent for n in nets]
    n_dom = z["domain_size"]
    if peer:  # only tokens, points and the pulled vectors are accounted
        assert all(s > 0 for s in sent)
    else:     # every reshare and both large openings crossed the net
        assert all(s > 32 * (7 * n_dom + 12 * 4 * n_dom) for s in sent)
    for s in sess:
        s.free()
    for pk in pks:
        pk.free()
    for n in nets:
        n.free()
    for s in states:
        s.free()
    for c in ctxs:
        c.close()


def test_plonk_rep3_native_peer_kat_emu():
    _run(_emu_factory(), peer=True, fixed_blinders=True)


def test_plonk_rep3_native_staged_drawn_blinders_emu():
    _run(_emu_factory(), peer=False, fixed_blinders=False)


@pytest.mark.gpu
def test_plonk_rep3_native_peer_kat_gpu():
    from co_snarks_b200 import binding as B
    _run(lambda: B.Context(0), name="poseidon", peer=True, fixed_blinders=True)


@pytest.mark.gpu
def test_plonk_rep3_native_staged_gpu():
    from co_snarks_b200 import binding as B
    _run(lambda: B.Context(0), peer=False, fixed_blinders=False)
