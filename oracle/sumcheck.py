"""UltraHonk sumcheck pieces, restated on Python integers (ORACLE: test infrastructure only -- nothing under
co_snarks_b200/ or bench.py's timed region imports this).

PARITY UNPINNED: the reference holds no known-answer vectors at this granularity (its Honk tests compare whole proofs
produced from Noir circuits through the circuit builder, co-noir/co-ultrahonk/tests, test_vectors/noir/*/kat), and no
Rust toolchain exists here to run it.  Each function follows the cited reference lines; the tests additionally check
what the protocol needs from them: the Rep3 versions open to the plain ones, and the round univariates built from the
accumulators satisfy the sumcheck identity S_i(0) + S_i(1) = S_{i-1}(u_{i-1}).

  gate_separator         co-noir/ultrahonk/src/decider/types.rs:53-67        GateSeparatorPolynomial::new
  partially_evaluate     co-noir/co-ultrahonk/src/co_decider/co_sumcheck/co_sumcheck_prover.rs:33-97
  extend_from_2          co-noir/ultrahonk/src/decider/univariate.rs:57-68   Univariate::extend_from, length 2
  arith_round_plain      co-noir/co-ultrahonk/src/co_decider/co_sumcheck/co_sumcheck_round.rs:261-305 (edge loop) with
                         relations/ultra_arithmetic_relation.rs:87-241 on the plain driver
  arith_round_rep3       the same on Rep3 shares: local_mul_vec without masks (mpc-core/.../rep3/arithmetic.rs:132-146),
                         mul_with_public_to_half_share / add_assign_public_half_share / add_with_public
                         (co-noir/co-noir-common/src/mpc/rep3.rs:69-95,131-137; rep3/arithmetic.rs:41-48)
  batch_univariates      co_sumcheck_round.rs:84-141 (scale by the alphas, extend to the batched length, multiply by the
                         pow-factor (1 - X) + X beta_i and the partial evaluation result) for this one relation
"""

MAX_PARTIAL_RELATION_LENGTH = 7
R0_LEN, R1_LEN = 6, 5
ARITH_WITNESS = ("w_l", "w_r", "w_o", "w_4", "w_l_shift", "w_4_shift")
ARITH_SELECTORS = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_arith")


def gate_separator(betas, log_n, r):
    out = [1] * (1 << log_n)
    for i, beta in enumerate(betas[:log_n]):
        index = 1 << i
        out[index] = beta % r
        for j in range(1, index):
            out[index + j] = out[j] * beta % r
    return out


def partially_evaluate(poly, u, r):
    """Public polynomial: list of ints; Rep3 shares: list of (a, b) -- both components (mul_with_public)."""
    out = []
    for i in range(0, len(poly), 2):
        lo, hi = poly[i], poly[i + 1]
        if isinstance(lo, tuple):
            out.append(tuple((x + (y - x) * u) % r for x, y in zip(lo, hi)))
        else:
            out.append((lo + (hi - lo) * u) % r)
    if len(out) < 2:
        out.append((0, 0) if isinstance(poly[0], tuple) else 0)
    return out


def extend_from_2(v0, v1, size, r):
    out = [v0 % r, v1 % r]
    delta = (v1 - v0) % r
    for _ in range(2, size):
        out.append((out[-1] + delta) % r)
    return out


def _edge(polys, name, e, r):
    p = polys[name]
    return extend_from_2(p[2 * e], p[2 * e + 1], MAX_PARTIAL_RELATION_LENGTH, r)


def arith_round_plain(polys, round_size, beta_products, periodicity, r):
    """polys: name -> list of ints (round_size rows).  -> (r0[6], r1[5])"""
    neg_half = (-pow(2, -1, r)) % r
    r0, r1 = [0] * R0_LEN, [0] * R1_LEN
    for e in range(round_size // 2):
        sf = beta_products[e * periodicity]
        x = {n: _edge(polys, n, e, r) for n in ARITH_WITNESS + ARITH_SELECTORS}
        if all(v == 0 for v in x["q_arith"]):  # can_skip (ultra_arithmetic_relation.rs:248-250)
            continue
        for k in range(MAX_PARTIAL_RELATION_LENGTH):
            qa = x["q_arith"][k]
            if k < R0_LEN:
                tmp = x["w_l"][k] * x["w_r"][k] * x["q_m"][k] % r
                tmp = tmp * (qa - 3) % r * neg_half % r
                tmp += x["q_l"][k] * x["w_l"][k] + x["q_r"][k] * x["w_r"][k] + x["q_o"][k] * x["w_o"][k] + x["q_4"][k] * x["w_4"][k]
                tmp += x["q_c"][k]
                tmp += (qa - 1) * x["w_4_shift"][k]
                r0[k] = (r0[k] + tmp % r * qa % r * sf) % r
            if k < R1_LEN:
                tmp = (x["w_l"][k] + x["w_4"][k] - x["w_l_shift"][k] + x["q_m"][k]) % r
                tmp = tmp * (qa - 2) % r * (qa - 1) % r * qa % r
                r1[k] = (r1[k] + tmp * sf) % r
    return r0, r1


def _edge_share(polys, name, e, r):
    p = polys[name]
    a = extend_from_2(p[2 * e][0], p[2 * e + 1][0], MAX_PARTIAL_RELATION_LENGTH, r)
    b = extend_from_2(p[2 * e][1], p[2 * e + 1][1], MAX_PARTIAL_RELATION_LENGTH, r)
    return list(zip(a, b))


def arith_round_rep3(polys, party, round_size, beta_products, periodicity, r):
    """One party's view.  polys: witness names -> list of (a, b) shares, selector names -> list of ints.
    -> (r0[6] additive shares WITHOUT the zero-share masks, r1[5] Rep3 shares (a, b))"""
    neg_half = (-pow(2, -1, r)) % r
    r0, r1 = [0] * R0_LEN, [(0, 0)] * R1_LEN
    for e in range(round_size // 2):
        sf = beta_products[e * periodicity]
        w = {n: _edge_share(polys, n, e, r) for n in ARITH_WITNESS}
        q = {n: _edge(polys, n, e, r) for n in ARITH_SELECTORS}
        if all(v == 0 for v in q["q_arith"]):
            continue
        for k in range(MAX_PARTIAL_RELATION_LENGTH):
            qa = q["q_arith"][k]
            if k < R0_LEN:
                (la, lb), (ra, rb) = w["w_l"][k], w["w_r"][k]
                mul = (la * ra + la * rb + lb * ra) % r
                tmp = mul * q["q_m"][k] % r * (qa - 3) % r * neg_half % r
                tmp += q["q_l"][k] * la + q["q_r"][k] * ra + q["q_o"][k] * w["w_o"][k][0] + q["q_4"][k] * w["w_4"][k][0]
                if party == 0:
                    tmp += q["q_c"][k]
                tmp += (qa - 1) * w["w_4_shift"][k][0]
                r0[k] = (r0[k] + tmp % r * qa % r * sf) % r
            if k < R1_LEN:
                f = (qa - 2) * (qa - 1) % r * qa % r * sf % r
                ta = w["w_l"][k][0] + w["w_4"][k][0] - w["w_l_shift"][k][0]
                tb = w["w_l"][k][1] + w["w_4"][k][1] - w["w_l_shift"][k][1]
                if party == 0:
                    ta += q["q_m"][k]
                if party == 1:
                    tb += q["q_m"][k]
                r1[k] = ((r1[k][0] + ta * f) % r, (r1[k][1] + tb * f) % r)
    return r0, r1


def lagrange_extend(evals, size, r):
    """Evaluations at 0..len-1 of a polynomial of degree < len -> evaluations at 0..size-1 (what Univariate::extend_from
    computes by its barycentric / finite-difference branches, univariate.rs:57-190: the polynomial is unique)."""
    n = len(evals)
    out = list(evals)
    for x in range(n, size):
        acc = 0
        for i, v in enumerate(evals):
            num, den = 1, 1
            for j in range(n):
                if j != i:
                    num = num * (x - j) % r
                    den = den * (i - j) % r
            acc = (acc + v * num % r * pow(den, -1, r)) % r
        out.append(acc)
    return out


def batch_univariates(r0, r1, alphas2, beta_i, partial_evaluation_result, size, r):
    """The round univariate from the two arithmetic accumulators: scale r0 by 1 and r1 by the first relation separator
    (AllRelationAcc::scale with running challenge 1: the first sub-relation is not scaled, co_sumcheck_round.rs:128-131),
    extend both to `size` evaluations and multiply by the pow-factor and the partial evaluation result
    (extend_and_batch_univariates, univariates.rs:56-73; both sub-relations are linearly independent)."""
    pow_poly = extend_from_2(1, beta_i, size, r)
    e0 = lagrange_extend(r0, size, r)
    e1 = lagrange_extend([v * alphas2 % r for v in r1], size, r)
    return [((a + b) * p % r) * partial_evaluation_result % r for a, b, p in zip(e0, e1, pow_poly)]
