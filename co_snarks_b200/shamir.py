"""Shamir co-Groth16 party driver for the three-party setting (t = 1, n = 3 -- the only 3-party setting the
reference accepts: mpc-core/src/protocols/shamir.rs:41-43).

Mirrors ShamirCoGroth16::prove (co-circom/co-groth16/src/groth16.rs:420-465) with ShamirGroth16Driver
(mpc/shamir.rs:29-119): the local phase -- witness map on degree-t shares and the five MSMs -- is one call on the
party's GPU (cs_groth16_shamir_local); the rest is single points.  With 2t + 1 = n every degree-2t half share can
be opened by all three parties directly (shamir/pointshare.rs:102-111), so the proof needs exactly two opening
rounds and no degree reduction:
    A  = open(g_a)                                   (degree t)
    C  = open(s_i A + r_i g1_b_i - (r_i s_i) delta + l_i + h_i)     (degree 2t)      B = open(g2_b)
Randomness: the reference takes r, s from preprocessed double sharings (ShamirState::rand); here the parties
build the same kind of object -- a degree-t sharing of a value no party knows -- by each dealing a random value
and adding up the dealt shares (one exchange of two field elements per pair).
"""
import secrets

import numpy as np

from . import binding as B

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def lagrange_at_zero(n=3):
    """Interpolation weights at 0 for evaluation points 1..n (shamir/core: lagrange_from_coeff)."""
    out = []
    for i in range(1, n + 1):
        num, den = 1, 1
        for j in range(1, n + 1):
            if j != i:
                num = num * (-j) % R_MOD
                den = den * (i - j) % R_MOD
        out.append(num * pow(den, -1, R_MOD) % R_MOD)
    return out


def _mont(vals):
    return B.ints_to_limbs(B.to_mont_ints(vals, R_MOD, 4), 4)


def _unmont(arr):
    return B.from_mont_ints(B.limbs_to_ints(np.asarray(arr).reshape(-1, 4)), R_MOD, 4)


class ShamirNetwork:
    """All-to-all exchange over a 3-rank torch.distributed group (party i evaluates polynomials at i + 1)."""

    def __init__(self, group=None, device="cpu"):
        import torch.distributed as dist
        self.dist, self.group, self.device = dist, group, device
        self.id = dist.get_rank(group)
        assert dist.get_world_size(group) == 3, "this driver covers t = 1, n = 3"
        self.bytes_sent = 0

    def all_gather(self, arr):
        """-> list of the three parties' arrays (own included), in party order."""
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy()).to(self.device)
        outs = [torch.empty_like(t) for _ in range(3)]
        self.dist.all_gather(outs, t, group=self.group)
        self.bytes_sent += 2 * t.numel() * 8
        return [o.cpu().numpy().view(np.uint64).reshape(np.shape(arr)) for o in outs]

    def scatter_private(self, per_recipient):
        """per_recipient[j] = what this party deals to party j.  Point-to-point only: party j receives its own
        row from every dealer and nothing else (a dealer's rows for the OTHER recipients never leave it -- two
        evaluations would determine a degree-1 polynomial).  -> list over dealers of the rows dealt to me."""
        import torch
        rows = [torch.from_numpy(np.ascontiguousarray(per_recipient[j]).view(np.int64).copy()).to(self.device) for j in range(3)]
        got = [torch.empty_like(rows[self.id]) for _ in range(3)]
        ops = []
        for j in range(3):
            if j == self.id:
                continue
            peer = self.dist.get_global_rank(self.group, j) if self.group is not None else j
            ops.append(self.dist.P2POp(self.dist.isend, rows[j], peer, self.group))
            ops.append(self.dist.P2POp(self.dist.irecv, got[j], peer, self.group))
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()
        got[self.id] = rows[self.id]
        self.bytes_sent += 2 * rows[0].numel() * 8
        self.last_received = [g.cpu().numpy().view(np.uint64).reshape(np.shape(per_recipient[0])) for g in got]
        return self.last_received


class ShamirCoGroth16:
    def __init__(self, ctx, pk, curve=B.CS_BN254):
        self.ctx, self.pk, self.curve = ctx, pk, curve
        self.lam = lagrange_at_zero(3)
        self.last_randomness = None

    def _joint_random_shares(self, net, count):
        """count degree-1 sharings of unknown uniform values: every party deals one value per sharing."""
        deal = np.zeros((3, count, 4), dtype=np.uint64)  # [recipient, k] = share for party `recipient`
        for k in range(count):
            v, a = secrets.randbelow(R_MOD), secrets.randbelow(R_MOD)
            deal[:, k, :] = B.ints_to_limbs([(v + a * (j + 1)) % R_MOD for j in range(3)], 4)
        got = net.scatter_private(deal)  # got[p] = the row party p dealt to me (and only that row)
        mine = [sum(B.limbs_to_ints(got[p][k].reshape(1, 4))[0] for p in range(3)) % R_MOD for k in range(count)]
        return mine  # canonical ints

    def _open_points(self, net, point, group):
        """Lagrange-weighted sum of the three parties' point shares (shamir/pointshare.rs:86-113)."""
        lib = self.ctx.lib
        parts = net.all_gather(point)
        acc = None
        for p in range(3):
            term = B.point_scalar_mul(lib, self.curve, group, parts[p], _mont([self.lam[p]])[0])
            acc = term if acc is None else B.point_add(lib, self.curve, group, acc, term)
        return acc

    def prove(self, net, public_inputs, witness_shares, delta_g1, r_share=None, s_share=None):
        """public_inputs: ni Montgomery elements (leading 1 included); witness_shares: this party's degree-1 shares
        (Montgomery).  r_share / s_share (canonical ints) override the jointly drawn randomness (tests)."""
        lib, cv = self.ctx.lib, self.curve
        if r_share is None or s_share is None:
            r_share, s_share = self._joint_random_shares(net, 2)
        self.last_randomness = (r_share, s_share)
        r_m, s_m = _mont([r_share])[0], _mont([s_share])[0]
        ga, gb1, gb2, l_acc, h_acc = self.pk.shamir_local(public_inputs, witness_shares, r_m.reshape(1, 4), s_m.reshape(1, 4))
        A = self._open_points(net, ga, B.CS_G1)                       # groth16.rs:305-308 (open_half_point)
        rs_m = _mont([r_share * s_share % R_MOD])[0]                   # local_mul_vec([r], [s]): degree 2t
        c = B.point_scalar_mul(lib, cv, B.CS_G1, A, s_m)
        c = B.point_add(lib, cv, B.CS_G1, c, B.point_scalar_mul(lib, cv, B.CS_G1, gb1, r_m))
        c = B.point_add(lib, cv, B.CS_G1, c, B.point_neg(lib, cv, B.CS_G1, B.point_scalar_mul(lib, cv, B.CS_G1, delta_g1, rs_m)))
        c = B.point_add(lib, cv, B.CS_G1, B.point_add(lib, cv, B.CS_G1, c, l_acc), h_acc)   # groth16.rs:314-322
        C = self._open_points(net, c, B.CS_G1)                         # groth16.rs:325-328
        Bp = self._open_points(net, gb2, B.CS_G2)
        return A, Bp, C
