"""Rep3 co-Plonk party driver over the device sessions of the C ABI (cs_plonk_rep3_*).

Mirrors co-circom/co-plonk/src/lib.rs:80-115 (prove_inner) for Rep3PlonkDriver (mpc/rep3.rs): the five rounds,
the Keccak transcript (types.rs:140-190) and the openings (open_point_g1 / open_vec, mpc/rep3.rs:113-138).
Everything vector-sized runs on the party's GPU (co_snarks_b200/csrc/cs_plonk_rep3.cuh); this module only
sequences the steps, sums the parties' partial commitments / evaluations / masked vectors and derives the
challenges.  `Rep3CoPlonk.prove` is written as a generator that yields what must be exchanged, so the same code
runs three parties in one process (LocalRep3Comm: tests, single-GPU use) or one party per process / GPU
(DistRep3Comm over torch.distributed + NVLink peer memory).

A proof equals the plain prover's for blinders b_k = sum of the parties' blinder shares (all masks cancel), which
is how the tests pin it to oracle/plonk.py and through it to the reference's known answers.
"""
import ctypes as C

import numpy as np

from . import binding as B

R_MOD = {B.CS_BN254: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
         B.CS_BLS12_381: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001}


def _conv(lib, fn, curve, arr):
    """cs_fr_to_mont / cs_fr_from_mont / cs_fq_from_mont over the rows of a 2-D limb array."""
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    assert a.ndim == 2
    out = np.zeros_like(a)
    rc = getattr(lib, fn)(curve, B._ptr(a), B._ptr(out), a.shape[0])
    assert rc == 0, lib.cs_last_error().decode()
    return out


class Transcript:
    """Keccak256Transcript (types.rs:140-190): big-endian canonical scalars / coordinates, zero bytes for infinity."""

    def __init__(self, lib, curve):
        self.lib, self.curve, self.buf = lib, curve, bytearray()

    def add_scalar(self, mont):
        c = _conv(self.lib, "cs_fr_from_mont", self.curve, np.asarray(mont, dtype=np.uint64).reshape(1, 4))
        self.buf += c.reshape(-1).astype("<u8").tobytes()[::-1]

    def add_point(self, affine_mont):
        fq = B.limbs_of(self.curve, "fq")
        c = _conv(self.lib, "cs_fq_from_mont", self.curve, np.asarray(affine_mont, dtype=np.uint64).reshape(2, fq))
        for row in c:
            self.buf += row.astype("<u8").tobytes()[::-1]

    def get_challenge(self):
        v = int.from_bytes(B.keccak256(self.lib, bytes(self.buf)), "big") % R_MOD[self.curve]
        return _conv(self.lib, "cs_fr_to_mont", self.curve, B.ints_to_limbs([v], 4))[0]


class Rep3CoPlonk:
    """One party.  `prove` is a generator: it yields (kind, payload) requests and expects the combined value back:
         ("sum_points", [k, 2 fq])  -> the k opened points          ("sum_vec", [m, 4]) -> the m opened scalars
         ("sum_dev", m) -> None once sess.d_in holds the sum over the parties of their m-element vectors at sess.d_out
         ("reshare", (slots, count)) -> None once every party's products for these arena slots are in place."""

    def __init__(self, ctx, pk, party, curve=B.CS_BN254):
        self.ctx, self.pk, self.party, self.curve = ctx, pk, party, curve
        self.sess = B.PlonkRep3Session(ctx, pk, party)

    def free(self):
        self.sess.free()

    def prove(self, state, public_inputs, witness_shares, vk_points, domain_size, blinder_shares=None):
        lib, cv, s = self.ctx.lib, self.curve, self.sess
        n = domain_size
        fq = B.limbs_of(cv, "fq")
        pub = np.ascontiguousarray(public_inputs, dtype=np.uint64).reshape(-1, 4)
        if blinder_shares is None:  # Round1Challenges::random (round1.rs:82-92): eleven T::rand shares
            blinder_shares = np.stack([state.rand() for _ in range(11)])
        # The streams move past whatever this proof consumed even when it aborts half-way (e.g. "Cannot invert
        # zero", a transport error, the caller dropping the generator): PRF output is never used twice.
        try:
            pts = s.round1(state.prf_args(), pub, witness_shares, blinder_shares)
            abc = yield ("sum_points", pts)
            t = Transcript(lib, cv)
            for P in np.asarray(vk_points, dtype=np.uint64).reshape(8, 2 * fq):
                t.add_point(P)
            for v in pub[1:]:
                t.add_scalar(v)
            for P in abc:
                t.add_point(P)
            beta = t.get_challenge()
            t = Transcript(lib, cv)
            t.add_scalar(beta)
            gamma = t.get_challenge()
            s.step(B.R3_ROUND2_A, np.stack([beta, gamma]))
            yield ("reshare", ([0, 1], n))
            s.step(B.R3_ROUND2_B)
            yield ("reshare", ([2, 3], n))
            s.step(B.R3_ROUND2_C)
            yield ("sum_dev", 2 * n + 1)  # device-resident opening: sess.d_out summed over the parties into sess.d_in
            s.step(B.R3_ROUND2_D)
            yield ("reshare", ([4, 5], n))
            s.step(B.R3_ROUND2_E)
            yield ("reshare", ([6], n))
            s.step(B.R3_ROUND2_F)
            yield ("sum_dev", n)
            zp = s.step(B.R3_ROUND2_G, None, (1, 2 * fq))
            (Z,) = yield ("sum_points", zp)
            t = Transcript(lib, cv)
            t.add_scalar(beta)
            t.add_scalar(gamma)
            t.add_point(Z)
            alpha = t.get_challenge()
            s.step(B.R3_ROUND3_A, alpha.reshape(1, 4))
            yield ("reshare", (list(range(12)), 4 * n))
            tp = s.step(B.R3_ROUND3_B, None, (3, 2 * fq))
            T = yield ("sum_points", tp)
            t = Transcript(lib, cv)
            t.add_scalar(alpha)
            for P in T:
                t.add_point(P)
            xi = t.get_challenge()
            ev = s.step(B.R3_ROUND4, xi.reshape(1, 4), (6, 4))
            opened = yield ("sum_vec", ev[:4])
            ea, eb, ec, ezw = opened
            es1, es2 = ev[4], ev[5]
            t = Transcript(lib, cv)
            for v in (xi, ea, eb, ec, es1, es2, ezw):
                t.add_scalar(v)
            v0 = t.get_challenge()
            wp = s.step(B.R3_ROUND5, np.stack([xi, v0, ea, eb, ec, es1, es2, ezw]), (2, 2 * fq))
            W = yield ("sum_points", wp)
        finally:
            state.advance(s.prf_words())
        points = np.concatenate([abc, Z.reshape(1, -1), T, W])  # A B C Z T1 T2 T3 Wxi Wxiw
        evals = np.stack([ea, eb, ec, es1, es2, ezw])
        return points, evals


def _sum_points(lib, curve, parts):
    out = []
    for k in range(parts[0].shape[0]):
        acc = parts[0][k]
        for p in parts[1:]:
            acc = B.point_add(lib, curve, B.CS_G1, acc, p[k])
        out.append(acc)
    return np.stack(out)


def _vec_add(ctx, curve, d_a, d_b, d_out, m):
    ctx._check(ctx.lib.cs_vec_add(ctx.h, curve, C.c_void_p(d_a), C.c_void_p(d_b), C.c_void_p(d_out), m))


class _CudaView:
    """Zero-copy torch view of library-owned device memory (CUDA array interface)."""

    def __init__(self, ptr, nwords64):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": (nwords64,), "typestr": "<i8", "version": 2}


def _sum_vec(ctx, curve, parts):
    m = parts[0].shape[0]
    d = [ctx.to_device(np.ascontiguousarray(p)) for p in parts]
    ctx._check(ctx.lib.cs_vec_add(ctx.h, curve, C.c_void_p(d[0]), C.c_void_p(d[1]), C.c_void_p(d[0]), m))
    ctx._check(ctx.lib.cs_vec_add(ctx.h, curve, C.c_void_p(d[0]), C.c_void_p(d[2]), C.c_void_p(d[0]), m))
    out = ctx.d2h(d[0], (m, 4))
    for x in d:
        ctx.free(x)
    return out


class LocalRep3Comm:
    """Three parties in one process (one GPU or the test emulation): products are stored straight into the next
    party's arena, and the 'network' is this scheduler."""

    def __init__(self, provers):
        assert len(provers) == 3
        self.provers = provers
        for p in range(3):
            provers[p].sess.connect(provers[(p + 1) % 3].sess.arena)

    def run(self, gens):
        ctx, curve = self.provers[0].ctx, self.provers[0].curve
        reqs = [next(g) for g in gens]
        results = [None] * 3
        done = 0
        while done < 3:
            kind = reqs[0][0]
            assert all(r[0] == kind for r in reqs), "parties out of step"
            if kind == "sum_points":
                ans = _sum_points(ctx.lib, curve, [r[1] for r in reqs])
            elif kind == "sum_vec":
                ans = _sum_vec(ctx, curve, [r[1] for r in reqs])
            elif kind == "sum_dev":
                m, ss = reqs[0][1], [pr.sess for pr in self.provers]
                for p in range(3):
                    _vec_add(ctx, curve, ss[0].d_out, ss[1].d_out, ss[p].d_in, m)
                    _vec_add(ctx, curve, ss[p].d_in, ss[2].d_out, ss[p].d_in, m)
                ctx.synchronize()
                ans = None
            else:
                ctx.synchronize()
                ans = None
            for p in range(3):
                try:
                    reqs[p] = gens[p].send(ans)
                except StopIteration as e:
                    results[p] = e.value
                    done += 1
        return results


class DistRep3Comm:
    """One party per process over a 3-rank torch.distributed group (co_snarks_b200.rep3.Rep3Network).
    With `peer=True` the next party's arena is mapped through CUDA IPC and products cross NVLink inside the
    kernels; otherwise the a-halves travel through the network and are packed in with cs_rep3_set_b."""

    def __init__(self, prover, net, peer=True):
        self.prover, self.net, self.peer = prover, net, peer
        self._mapped = None
        if peer:
            import torch
            ctx = prover.ctx
            h = torch.from_numpy(ctx.ipc_export(prover.sess.arena).copy()).to(net.device)
            outs = [torch.empty_like(h) for _ in range(3)]
            net.dist.all_gather(outs, h, group=net.group)
            self._mapped = ctx.ipc_open(outs[net.next].cpu().numpy())
            prover.sess.connect(self._mapped)

    def close(self):
        if self._mapped:
            self.prover.ctx.synchronize()
            self.net.dist.barrier(group=self.net.group)
            self.prover.ctx.ipc_close(self._mapped)
            self._mapped = None

    def _gather(self, arr):
        prev, nxt = self.net.broadcast(arr)
        return [arr, prev, nxt]

    def _sum_dev(self, m):
        """Opens an m-element additive vector without leaving the device when the group's tensors live on the
        GPU (all_gather straight from the session buffer, two cs_vec_add); through the host otherwise."""
        ctx, curve, sess = self.prover.ctx, self.prover.curve, self.prover.sess
        ctx.synchronize()
        if str(self.net.device).startswith("cuda"):
            import torch
            mine = torch.as_tensor(_CudaView(sess.d_out, 4 * m), device="cuda")
            outs = [torch.empty(4 * m, dtype=torch.int64, device="cuda") for _ in range(3)]
            self.net.dist.all_gather(outs, mine, group=self.net.group)
            torch.cuda.synchronize()
            self.net.bytes_sent += 2 * 32 * m
            _vec_add(ctx, curve, outs[0].data_ptr(), outs[1].data_ptr(), sess.d_in, m)
            _vec_add(ctx, curve, sess.d_in, outs[2].data_ptr(), sess.d_in, m)
            ctx.synchronize()
        else:
            parts = self._gather(ctx.d2h(sess.d_out, (m, 4)))
            ctx.h2d(sess.d_in, _sum_vec(ctx, curve, parts))

    def run(self, gen, trace=None):
        """`trace` (a list) receives (request kind, ms until the party's GPU drained, ms in the exchange) per step."""
        import time
        ctx, curve, sess = self.prover.ctx, self.prover.curve, self.prover.sess
        t0 = time.perf_counter()
        req = next(gen)
        while True:
            kind, payload = req
            if trace is not None:
                ctx.synchronize()
                t1 = time.perf_counter()
            if kind == "sum_points":
                ans = _sum_points(ctx.lib, curve, self._gather(payload))
            elif kind == "sum_vec":
                ans = _sum_vec(ctx, curve, self._gather(payload))
            elif kind == "sum_dev":
                self._sum_dev(payload)
                ans = None
            else:
                slots, count = payload
                ctx.synchronize()
                if self.peer:
                    self.net.dist.barrier(group=self.net.group)
                else:
                    for k in slots:
                        base = sess.arena + k * sess.slot_bytes
                        za = ctx.d2h(base, (count, 2, 4))[:, 0, :].copy()
                        zb = self.net.reshare(za)
                        d = ctx.to_device(np.ascontiguousarray(zb))
                        ctx.rep3_set_b(curve, d, count, base)
                        ctx.synchronize()
                        ctx.free(d)
                ans = None
            if trace is not None:
                t2 = time.perf_counter()
                trace.append((kind, round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
                t0 = t2
            try:
                req = gen.send(ans)
            except StopIteration as e:
                if trace is not None:
                    ctx.synchronize()
                    trace.append(("end", round((time.perf_counter() - t0) * 1e3, 2), 0.0))
                return e.value
