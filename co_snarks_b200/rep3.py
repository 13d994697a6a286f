"""Rep3 co-Groth16 party driver: the host-side mirror of `Rep3Groth16Driver` + `Rep3CoGroth16::prove`.

Reference: co-circom/co-groth16/src/groth16.rs:125-177 (prove_inner), :207-338
(create_proof_with_assignment), :360-379 (Rep3CoGroth16::prove); driver co-groth16/src/mpc/rep3.rs;
protocol pieces mpc-core/src/protocols/rep3/{rngs.rs:103-156, arithmetic.rs:132-146,357-360,
pointshare.rs:119-155, network.rs:30-79, id.rs:31-47}.

Party i runs on rank i (GPU i of one box).  Everything vector-sized (witness map on shares, the five
MSMs) happens in ONE C-ABI call on the party's GPU (`cs_groth16_rep3_local`); what is left is the
reference's two network legs -- four point-sized messages -- carried by `torch.distributed` (NCCL on
GPUs, gloo on CPU) in place of mpc-net's TCP.  Single points are combined with the library's host
helpers (`cs_point_*`), as the reference does on the CPU.

Randomness: the reference seeds two ChaCha12 PRFs per party by a seed exchange (rep3.rs:71-110) and
draws field elements from them; here the same correlated-randomness structure (own stream / previous
party's stream) is driven by seeded PCG64 streams.  Masks cancel on opening (rngs.rs:103-106) either
way; bit-compatibility with ChaCha12 (needed only to interoperate with a CPU party) is listed under
"next" in DESIGN.md.
"""
import numpy as np

from . import binding as B

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_G1_GEN_CANON = (1, 2)


class Rep3Network:
    """Rep3NetworkExt over a 3-rank torch.distributed group: `reshare` = send to next, receive from
    previous (network.rs:30-36); `broadcast` = exchange with both (network.rs:57-79)."""

    def __init__(self, group=None, device="cpu"):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.id = dist.get_rank(group)
        assert dist.get_world_size(group) == 3, "Rep3 needs exactly three parties"
        self.next = (self.id + 1) % 3   # id.rs:31-47
        self.prev = (self.id + 2) % 3
        self.device = device
        self.bytes_sent = 0

    def _t(self, arr):
        import torch
        return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy()).to(self.device)

    def reshare(self, arr):
        """send `arr` to next, return what prev sent."""
        import torch
        out = torch.empty_like(self._t(arr))
        src = self._t(arr)
        ops = [self.dist.P2POp(self.dist.isend, src, self._global(self.next), self.group),
               self.dist.P2POp(self.dist.irecv, out, self._global(self.prev), self.group)]
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()
        self.bytes_sent += src.numel() * 8
        return out.cpu().numpy().view(np.uint64).reshape(np.shape(arr))

    def _global(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def broadcast(self, arr):
        """-> (value of prev, value of next)."""
        import torch
        src = self._t(arr)
        outs = [torch.empty_like(src) for _ in range(3)]
        self.dist.all_gather(outs, src, group=self.group)
        self.bytes_sent += 2 * src.numel() * 8
        conv = lambda t: t.cpu().numpy().view(np.uint64).reshape(np.shape(arr))
        return conv(outs[self.prev]), conv(outs[self.next])


class PairLink:
    """Point-sized link between the two GPUs (ranks) of one party."""

    def __init__(self, peer_global_rank, device="cpu"):
        import torch.distributed as dist
        self.dist, self.peer, self.device = dist, peer_global_rank, device

    def send(self, arr):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy()).to(self.device)
        self.dist.send(t, self.peer)

    def recv(self, nwords):
        import torch
        t = torch.empty(nwords, dtype=torch.int64, device=self.device)
        self.dist.recv(t, self.peer)
        return t.cpu().numpy().view(np.uint64)


class Rep3State:
    """Correlated randomness of one party: rng1 = own stream, rng2 = previous party's stream
    (Rep3Rand, rngs.rs:86-156; seeds exchanged once over the network, rep3.rs:71-110)."""

    def __init__(self, net, seed):
        own = np.array([seed & (2 ** 63 - 1), net.id], dtype=np.uint64)
        prev = net.reshare(own)
        self.id = net.id
        self.rng1 = np.random.Generator(np.random.PCG64(int(own[0]) * 4 + int(own[1])))
        self.rng2 = np.random.Generator(np.random.PCG64(int(prev[0]) * 4 + int(prev[1])))

    @staticmethod
    def _fes(rng, n):
        """n field elements as canonical limbs [n,4] (253-bit draws, < r)."""
        a = rng.integers(0, 2 ** 63, size=(n, 4), dtype=np.uint64) << np.uint64(1)
        a |= rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 61) - 1)
        return a

    def random_fes(self, n=1):
        return self._fes(self.rng1, n), self._fes(self.rng2, n)

    def rand(self, lib, curve):
        """arithmetic::rand (arithmetic.rs:357-360): share (a, b) = (F(rng1), F(rng2)) in Montgomery form."""
        a, b = self.random_fes(1)
        out = np.zeros((2, 4), dtype=np.uint64)
        lib.cs_fr_to_mont(curve, B._ptr(np.concatenate([a, b])), B._ptr(out), 2)
        return out  # [[a],[b]] Montgomery

    def masking_field_elements_vec(self, lib, curve, n):
        """rngs.rs:137-156: a_i - b_i (Montgomery limbs); sums to zero over the three parties."""
        a, b = self.random_fes(n)
        ai, bi = B.limbs_to_ints(a), B.limbs_to_ints(b)
        diff = B.ints_to_limbs([(x - y) % BN254_R for x, y in zip(ai, bi)], 4)
        out = np.zeros_like(diff)
        lib.cs_fr_to_mont(curve, B._ptr(diff), B._ptr(out), n)
        return out

    def masking_ec_element(self, lib, curve, gen_mont):
        """rngs.rs:177-186: C::rand(rng1) - C::rand(rng2), here k1*G - k2*G."""
        a, b = self.random_fes(1)
        am, bm = np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64)
        lib.cs_fr_to_mont(curve, B._ptr(a), B._ptr(am), 1)
        lib.cs_fr_to_mont(curve, B._ptr(b), B._ptr(bm), 1)
        p1 = B.point_scalar_mul(lib, curve, B.CS_G1, gen_mont, am[0])
        p2 = B.point_scalar_mul(lib, curve, B.CS_G1, gen_mont, bm[0])
        return B.point_add(lib, curve, B.CS_G1, p1, B.point_neg(lib, curve, B.CS_G1, p2))


def _fr_mul_mont(lib, curve, x_mont, y_mont):
    """product of two Montgomery Fr elements via canonical ints (single elements; latency-only)."""
    xc, yc = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    lib.cs_fr_from_mont(curve, B._ptr(np.ascontiguousarray(x_mont)), B._ptr(xc), 1)
    lib.cs_fr_from_mont(curve, B._ptr(np.ascontiguousarray(y_mont)), B._ptr(yc), 1)
    v = B.limbs_to_ints(xc.reshape(1, 4))[0] * B.limbs_to_ints(yc.reshape(1, 4))[0] % BN254_R
    return v


def _fr_from_int(lib, curve, v):
    c = B.ints_to_limbs([v % BN254_R], 4)
    out = np.zeros_like(c)
    lib.cs_fr_to_mont(curve, B._ptr(c), B._ptr(out), 1)
    return out[0]


class Rep3CoGroth16:
    """`Rep3CoGroth16::<P>::prove::<N, CircomReduction>(net0, net1, &pkey, &matrices, witness)`
    (groth16.rs:360-379): `pk` is the device-resident key (binding.Groth16Key) holding pkey + matrices."""

    def __init__(self, ctx, pk, curve=B.CS_BN254):
        self.ctx, self.pk, self.curve, self.lib = ctx, pk, curve, ctx.lib
        q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
        self.gen_mont = B.ints_to_limbs(B.to_mont_ints(list(_G1_GEN_CANON), q, 4), 4).reshape(-1)

    def draw(self, state, masks=None):
        """All correlated randomness of one proof, in the order the reference consumes it: two mask
        vectors (reduction.rs:160,182), r, s (groth16.rs:157), the rs mask (groth16.rs:297) and the EC
        mask of scalar_mul (pointshare.rs:119-125)."""
        lib, cv, n = self.lib, self.curve, self.pk.domain_size()
        if masks is None:
            m1 = state.masking_field_elements_vec(lib, cv, n)
            m2 = state.masking_field_elements_vec(lib, cv, n)
        else:
            m1, m2 = masks  # pre-drawn from the same two streams (bench.py draws them on the device)
        r_sh, s_sh = state.rand(lib, cv), state.rand(lib, cv)
        rs_mask = B.from_mont_ints(B.limbs_to_ints(state.masking_field_elements_vec(lib, cv, 1)), BN254_R, 4)[0]
        ec_mask = state.masking_ec_element(lib, cv, self.gen_mont)
        return m1, m2, r_sh, s_sh, rs_mask, ec_mask

    def helper_step(self, pid, state, public_inputs, witness_shares, pair_send, masks=None):
        """Second GPU of a party: runs {witness map -> H, B2} and hands the two points to
        the party's first GPU.  Draws the same randomness so both GPUs stay in lock-step."""
        m1, m2, r_sh, s_sh, _, _ = self.draw(state, masks)
        _, _, g2_b, _, h_acc = self.pk.rep3_local(pid, public_inputs, witness_shares, m1, m2, r_sh, s_sh,
                                                  parts=B.CS_PART_B2 | B.CS_PART_H)
        pair_send(np.concatenate([g2_b, h_acc]))

    def prove(self, net, state, public_inputs, witness_shares, delta_g1, masks=None, pair_recv=None):
        """witness_shares: [nw, 8] uint64 (a‖b per share, Montgomery); public_inputs incl. the leading 1.
        delta_g1: affine Montgomery (pkey.delta_g1).  Returns (A, B, C) affine Montgomery; all parties
        return the same proof (tests/test_dist_rep3.py).  pair_recv: when the party owns a second GPU,
        a callable returning the helper's (g2_b ‖ h_acc)."""
        lib, cv, pk = self.lib, self.curve, self.pk
        pid = net.id
        m1, m2, r_sh, s_sh, mask, ec_mask = self.draw(state, masks)
        if pair_recv is None:
            g_a, g1_b, g2_b, l_acc, h_acc = pk.rep3_local(pid, public_inputs, witness_shares, m1, m2, r_sh, s_sh)
        else:
            g_a, g1_b, _, l_acc, _ = pk.rep3_local(pid, public_inputs, witness_shares, None, None, r_sh, s_sh,
                                                   parts=B.CS_PART_A | B.CS_PART_B1 | B.CS_PART_L)
            both = pair_recv()
            g2_b, h_acc = both[:g2_b_len(pk)], both[g2_b_len(pk):]
        # rs = local_mul_vec([r],[s]) (groth16.rs:297): r.a*s.a + r.a*s.b + r.b*s.a + mask
        rs = (_fr_mul_mont(lib, cv, r_sh[0], s_sh[0]) + _fr_mul_mont(lib, cv, r_sh[0], s_sh[1]) +
              _fr_mul_mont(lib, cv, r_sh[1], s_sh[0]) + mask) % BN254_R
        r_s_delta = B.point_scalar_mul(lib, cv, B.CS_G1, delta_g1, _fr_from_int(lib, cv, rs))
        # network round 1 (groth16.rs:305-308): open_half_point(g_a) | scalar_mul(g1_b, r)
        pa, pn = net.broadcast(g_a)
        g_a_opened = B.point_add(lib, cv, B.CS_G1, B.point_add(lib, cv, B.CS_G1, g_a, pa), pn)
        g1_b_prev = net.reshare(g1_b)  # Rep3PointShare::new(a = own, b = prev's)  (mpc/rep3.rs:158-160)
        t = B.point_scalar_mul(lib, cv, B.CS_G1, g1_b, r_sh[0])                       # rhs.a * self.a
        t = B.point_add(lib, cv, B.CS_G1, t, B.point_scalar_mul(lib, cv, B.CS_G1, g1_b_prev, r_sh[0]))  # rhs.b * self.a
        t = B.point_add(lib, cv, B.CS_G1, t, B.point_scalar_mul(lib, cv, B.CS_G1, g1_b, r_sh[1]))       # rhs.a * self.b
        r_g1_b = B.point_add(lib, cv, B.CS_G1, t, ec_mask)
        # groth16.rs:314-322
        g_c = B.point_scalar_mul(lib, cv, B.CS_G1, g_a_opened, s_sh[0])
        g_c = B.point_add(lib, cv, B.CS_G1, g_c, r_g1_b)
        g_c = B.point_add(lib, cv, B.CS_G1, g_c, B.point_neg(lib, cv, B.CS_G1, r_s_delta))
        g_c = B.point_add(lib, cv, B.CS_G1, g_c, l_acc)
        g_c = B.point_add(lib, cv, B.CS_G1, g_c, h_acc)
        # network round 2 (groth16.rs:325-328)
        pa, pn = net.broadcast(g_c)
        g_c_opened = B.point_add(lib, cv, B.CS_G1, B.point_add(lib, cv, B.CS_G1, g_c, pa), pn)
        pa, pn = net.broadcast(g2_b)
        g2_b_opened = B.point_add(lib, cv, B.CS_G2, B.point_add(lib, cv, B.CS_G2, g2_b, pa), pn)
        self.last_randomness = (r_sh, s_sh)
        return g_a_opened, g2_b_opened, g_c_opened


def g2_b_len(pk):
    return 4 * pk.fq
