"""Groth16 prover restatement (oracle; test infrastructure only).

Follows co-circom/co-groth16/src/groth16.rs:125-338 (prove_inner, calculate_coeff,
create_proof_with_assignment) and groth16/reduction.rs:45-60,77-193,196-226 (CircomReduction),
with the driver semantics of mpc/plain.rs and mpc/rep3.rs and the Rep3 share algebra of
mpc-core/src/protocols/rep3/{arithmetic.rs:41-58,132-146, arithmetic/ops.rs:69-76,
pointshare.rs:119-155, id.rs:31-47}.  Randomness (r, s, Rep3 masks) is injected so results are
reproducible: the reference draws it from thread_rng / ChaCha12 PRFs (mpc/plain.rs:23-26,
rep3/rngs.rs:103-156), which is why its own tests pin validity, not bytes.
"""
from .fields import groth16_roots_of_unity, inv
from .ntt import ifft_in_to_out, fft_out_to_in, bit_reverse_perm
from .ec import g1 as _g1, g2 as _g2


def bit_reversed_coset_table(shift, size, r):
    """reduction.rs:45-60."""
    t, cur = [], 1
    for _ in range(size):
        t.append(cur)
        cur = cur * shift % r
    return bit_reverse_perm(t)


def domain_params(matrices, r, two_adicity):
    """reduction.rs:83-93."""
    n = matrices["num_constraints"] + matrices["num_instance_variables"]
    domain_size = 1 << max(0, (n - 1).bit_length())
    power = domain_size.bit_length() - 1
    if power > two_adicity:
        raise ValueError("Polynomial Degree too large")
    gen, shift = groth16_roots_of_unity(r, power)
    return domain_size, gen, shift


# ---------------------------------------------------------------- plain driver (mpc/plain.rs)
def evaluate_constraint_plain(rows, public_inputs, witness, domain_size, r):
    """reduction.rs:196-210 with mpc/plain.rs:29-43."""
    npub = len(public_inputs)
    out = []
    for row in rows:
        acc = 0
        for coeff, idx in row:
            acc += coeff * (public_inputs[idx] if idx < npub else witness[idx - npub])
        out.append(acc % r)
    out += [0] * (domain_size - len(out))
    return out


def witness_map_plain(matrices, public_inputs, witness, r, two_adicity):
    """CircomReduction::witness_map_from_matrices with T = PlainGroth16Driver -> h (natural order)."""
    nc, ni = matrices["num_constraints"], matrices["num_instance_variables"]
    n, gen, shift = domain_params(matrices, r, two_adicity)
    table = bit_reversed_coset_table(shift, n, r)
    a = evaluate_constraint_plain(matrices["a"], public_inputs, witness, n, r)
    a[nc:nc + ni] = list(public_inputs[:ni])
    b = evaluate_constraint_plain(matrices["b"], public_inputs, witness, n, r)

    def coset(v):
        v = ifft_in_to_out(v, gen, r)
        v = [x * t % r for x, t in zip(v, table)]
        return fft_out_to_in(v, gen, r)

    c = coset([x * y % r for x, y in zip(a, b)])
    a2, b2 = coset(a), coset(b)
    return [(x * y - z) % r for x, y, z in zip(a2, b2, c)]


# ---------------------------------------------------------------- LibSnarkReduction (reduction.rs:241-342)
ARK_GENERATOR = {21888242871839275222246405745257275088548364400416034343698204186575808495617: 5,
                 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001: 7,
                 # BLS12-377 Fr (ark-bls12-377): the LibSnark fixtures of the reference live on this field
                 8444461749428370424248824938781546531375899335154063827935233455917409239041: 22}


def ark_domain(n_min, r):
    """Domain::new(n): size = next power of two, generator = GENERATOR^((r-1)/size) (arkworks'
    get_root_of_unity; checked against ark-bn254 / ark-bls12-381's TWO_ADIC_ROOT_OF_UNITY)."""
    size = 1 << max(0, (n_min - 1).bit_length())
    g = ARK_GENERATOR[r]
    return size, pow(g, (r - 1) // size, r), g


def witness_map_libsnark(matrices, public_inputs, witness, r, kind="plain", pid=0, mask=None):
    """LibSnarkReduction::witness_map_from_matrices -> coefficients of H in natural order.
    kind = "plain" (witness = values) or "rep3" (witness = (a, b) shares, one mask vector).
    PINNED on the reference's own fixture: with these coefficients and the reference's BLS12-377 Penumbra proving key
    the assembled proof verifies under its circuit.vk (tests/golden/make_libsnark_bls12_377.py,
    tests/test_oracle_golden.py::test_libsnark_reduction_pinned_on_the_reference_bls12_377_fixture)."""
    nc, ni = matrices["num_constraints"], matrices["num_instance_variables"]
    n, gen, g = ark_domain(nc + ni, r)
    table = bit_reversed_coset_table(g, n, r)

    def coset1(v):
        v = ifft_in_to_out(v, gen, r)
        v = [x * t % r for x, t in zip(v, table)]
        return fft_out_to_in(v, gen, r)

    if kind == "plain":
        a = evaluate_constraint_plain(matrices["a"], public_inputs, witness, n, r)
        a[nc:nc + ni] = list(public_inputs[:ni])
        b = evaluate_constraint_plain(matrices["b"], public_inputs, witness, n, r)
        c = evaluate_constraint_plain(matrices["c"], public_inputs, witness, n, r)
        a, b = coset1(a), coset1(b)
        ab = [x * y % r for x, y in zip(a, b)]
    else:
        a = evaluate_constraint_rep3(pid, matrices["a"], public_inputs, witness, n, r)
        for k in range(ni):
            a[nc + k] = promote_to_trivial_share(pid, public_inputs[k])
        b = evaluate_constraint_rep3(pid, matrices["b"], public_inputs, witness, n, r)
        # evaluate_constraint_half_share (mpc/rep3.rs:51-74): public terms on party 0 only, witness.a
        npub = len(public_inputs)
        c = []
        for row in matrices["c"]:
            acc = 0
            for coeff, idx in row:
                if idx < npub:
                    if pid == 0:
                        acc += public_inputs[idx] * coeff
                else:
                    acc += witness[idx - npub][0] * coeff
            c.append(acc % r)
        c += [0] * (n - len(c))
        a = list(zip(coset1([x[0] for x in a]), coset1([x[1] for x in a])))
        b = list(zip(coset1([x[0] for x in b]), coset1([x[1] for x in b])))
        ab = local_mul_vec_rep3(a, b, mask if mask is not None else [0] * n, r)
    c = coset1(c)
    vinv = inv((pow(g, n, r) - 1) % r, r)
    ab = [(x - y) * vinv % r for x, y in zip(ab, c)]
    ab = bit_reverse_perm(ifft_in_to_out(ab, gen, r))
    ginv = inv(g, r)
    out, cur = [], 1
    for x in ab:
        out.append(x * cur % r)
        cur = cur * ginv % r
    return out


def _calc_coeff(G, initial, query, vk_param, inputs, aux, add_public=True):
    """groth16.rs:179-203.  `add_public` = whether this party adds public points (party 0 / plain)."""
    npub = len(inputs)
    priv = G.msm(query[1 + npub:], aux)
    res = G.to_jac(initial)
    if add_public:
        pub = G.msm(query[1:1 + npub], inputs)
        for P in (query[0], vk_param, pub):
            res = G.jadd(res, G.to_jac(P))
    res = G.jadd(res, G.to_jac(priv))
    return G.to_affine(res)


def prove_plain(z, matrices, full_witness, r_rand, s_rand):
    """Groth16::plain_prove (groth16.rs:484-490) with injected r, s.  `full_witness` = wtns values
    (w[0] = 1).  Returns affine (A, B, C)."""
    curve = z["curve"]
    r = curve.r
    G1, G2 = _g1(curve), _g2(curve)
    ni = matrices["num_instance_variables"]
    public_inputs, witness = full_witness[:ni], full_witness[ni:]
    assert len(witness) == matrices["num_witness_variables"]
    h = witness_map_plain(matrices, public_inputs, witness, r, curve.two_adicity)
    inputs = public_inputs[1:]
    A = _calc_coeff(G1, G1.mul(z["delta_g1"], r_rand), z["a_query"], z["alpha_g1"], inputs, witness)
    B1 = _calc_coeff(G1, G1.mul(z["delta_g1"], s_rand), z["b_g1_query"], z["beta_g1"], inputs, witness)
    B2 = _calc_coeff(G2, G2.mul(z["delta_g2"], s_rand), z["b_g2_query"], z["beta_g2"], inputs, witness)
    L = G1.msm(z["l_query"], witness)
    H = G1.msm(z["h_query"], h)
    rs = r_rand * s_rand % r
    C = G1.add(G1.mul(A, s_rand), G1.mul(B1, r_rand))
    C = G1.add(C, G1.neg(G1.mul(z["delta_g1"], rs)))
    C = G1.add(G1.add(C, L), H)
    return A, B2, C


# ---------------------------------------------------------------- Rep3 emulation (mpc/rep3.rs)
def share_rep3(values, r, rng):
    """mpc-core/src/protocols/rep3.rs:281-293: x = x0+x1+x2; party i holds (a, b) = (x_i, x_{i-1})."""
    shares = ([], [], [])
    for x in values:
        x0, x1 = rng.randrange(r), rng.randrange(r)
        x2 = (x - x0 - x1) % r
        xs = (x0, x1, x2)
        for i in range(3):
            shares[i].append((xs[i], xs[(i + 2) % 3]))
    return shares


def evaluate_constraint_rep3(pid, rows, public_inputs, wshares, domain_size, r):
    """mpc/rep3.rs:31-49: public terms go to party 0's `a` / party 1's `b` (arithmetic.rs:52-58)."""
    npub = len(public_inputs)
    out = []
    for row in rows:
        a = b = 0
        for coeff, idx in row:
            if idx < npub:
                m = public_inputs[idx] * coeff
                if pid == 0:
                    a += m
                elif pid == 1:
                    b += m
            else:
                wa, wb = wshares[idx - npub]
                a += wa * coeff
                b += wb * coeff
        out.append((a % r, b % r))
    out += [(0, 0)] * (domain_size - len(out))
    return out


def promote_to_trivial_share(pid, v):
    """rep3/arithmetic/types.rs promote_from_trivial: party0 (v,0), party1 (0,v), party2 (0,0)."""
    return ((v, 0), (0, v), (0, 0))[pid]


def local_mul_vec_rep3(x, y, masks, r):
    """rep3/arithmetic.rs:132-146 with ops.rs:69-76: a.a*b.a + a.a*b.b + a.b*b.a + mask."""
    return [(xa * ya + xa * yb + xb * ya + m) % r for (xa, xb), (ya, yb), m in zip(x, y, masks)]


def witness_map_rep3(pid, matrices, public_inputs, wshares, masks1, masks2, r, two_adicity):
    """CircomReduction with T = Rep3Groth16Driver for one party -> half shares of h."""
    nc, ni = matrices["num_constraints"], matrices["num_instance_variables"]
    n, gen, shift = domain_params(matrices, r, two_adicity)
    table = bit_reversed_coset_table(shift, n, r)
    a = evaluate_constraint_rep3(pid, matrices["a"], public_inputs, wshares, n, r)
    for k in range(ni):
        a[nc + k] = promote_to_trivial_share(pid, public_inputs[k])
    b = evaluate_constraint_rep3(pid, matrices["b"], public_inputs, wshares, n, r)

    def coset1(v):
        v = ifft_in_to_out(v, gen, r)
        v = [x * t % r for x, t in zip(v, table)]
        return fft_out_to_in(v, gen, r)

    def coset2(v):
        return list(zip(coset1([x[0] for x in v]), coset1([x[1] for x in v])))

    c = coset1(local_mul_vec_rep3(a, b, masks1, r))
    a2, b2 = coset2(a), coset2(b)
    ab = local_mul_vec_rep3(a2, b2, masks2, r)
    return [(x - y) % r for x, y in zip(ab, c)]


def prove_rep3(z, matrices, full_witness, rng):
    """Three-party emulation of Rep3CoGroth16::prove (groth16.rs:360-379) in one process.
    Returns (proof, r_total, s_total): masks cancel on opening (rngs.rs:103-106), so the proof equals
    prove_plain with r = sum r_i.a, s = sum s_i.a."""
    curve = z["curve"]
    r = curve.r
    G1, G2 = _g1(curve), _g2(curve)
    ni = matrices["num_instance_variables"]
    public_inputs = full_witness[:ni]
    wsh = share_rep3(full_witness[ni:], r, rng)
    n = domain_params(matrices, r, curve.two_adicity)[0]

    def zero_masks(k):
        # mask_i = prf_i - prf_{i-1}  (rngs.rs:103-106)
        prf = [[rng.randrange(r) for _ in range(k)] for _ in range(3)]
        return [[(prf[i][j] - prf[(i + 2) % 3][j]) % r for j in range(k)] for i in range(3)]

    m1, m2 = zero_masks(n), zero_masks(n)
    hs = [witness_map_rep3(i, matrices, public_inputs, wsh[i], m1[i], m2[i], r, curve.two_adicity)
          for i in range(3)]
    rsh = share_rep3([rng.randrange(r)], r, rng)
    ssh = share_rep3([rng.randrange(r)], r, rng)
    rs_mask = zero_masks(1)
    ec_mask_scalars = [rng.randrange(r) for _ in range(3)]
    inputs = public_inputs[1:]
    gA, gB1, gB2, gL, gH = [], [], [], [], []
    for i in range(3):
        aux = [w[0] for w in wsh[i]]
        ri, si = rsh[i][0][0], ssh[i][0][0]
        gA.append(_calc_coeff(G1, G1.mul(z["delta_g1"], ri), z["a_query"], z["alpha_g1"], inputs, aux, i == 0))
        gB1.append(_calc_coeff(G1, G1.mul(z["delta_g1"], si), z["b_g1_query"], z["beta_g1"], inputs, aux, i == 0))
        gB2.append(_calc_coeff(G2, G2.mul(z["delta_g2"], si), z["b_g2_query"], z["beta_g2"], inputs, aux, i == 0))
        gL.append(G1.msm(z["l_query"], aux))
        gH.append(G1.msm(z["h_query"], hs[i]))
    # open_half_point(A): broadcast + sum (pointshare.rs:152-155)
    A = None
    for P in gA:
        A = G1.add(A, P)
    # scalar_mul(B1, r): reshare then local  b*a + mask (mpc/rep3.rs:152-161, pointshare.rs:119-125)
    ec_masks = [G1.mul(curve.g1, (ec_mask_scalars[i] - ec_mask_scalars[(i + 2) % 3]) % r) for i in range(3)]
    gC = []
    for i in range(3):
        pa, pb = gB1[i], gB1[(i + 2) % 3]  # point share (a = own, b = prev's)
        ra, rb = rsh[i][0]
        t = G1.add(G1.add(G1.mul(pa, ra), G1.mul(pa, rb)), G1.mul(pb, ra))
        r_b1 = G1.add(t, ec_masks[i])
        rs_i = local_mul_vec_rep3([rsh[i][0]], [ssh[i][0]], [rs_mask[i][0]], r)[0]
        c = G1.mul(A, ssh[i][0][0])
        c = G1.add(c, r_b1)
        c = G1.add(c, G1.neg(G1.mul(z["delta_g1"], rs_i)))
        c = G1.add(G1.add(c, gL[i]), gH[i])
        gC.append(c)
    C, B = None, None
    for P in gC:
        C = G1.add(C, P)
    for P in gB2:
        B = G2.add(B, P)
    r_tot = sum(x[0][0] for x in rsh) % r
    s_tot = sum(x[0][0] for x in ssh) % r
    return (A, B, C), r_tot, s_tot


def vk_from_zkey(z):
    return dict(alpha_g1=z["alpha_g1"], beta_g2=z["beta_g2"], gamma_g2=z["gamma_g2"],
                delta_g2=z["delta_g2"], ic=z["ic"])
