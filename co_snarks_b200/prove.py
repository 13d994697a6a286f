"""`python -m co_snarks_b200.prove --zkey circuit.zkey --wtns witness.wtns --out proof.json`

The GPU counterpart of `co-circom generate-proof groth16` for the plain driver
(co-circom/co-circom/src/bin/co-circom.rs:966-1066): zkey -> device-resident key, wtns -> witness,
Groth16::plain_prove with fresh (r, s), proof written in snarkjs' JSON layout (decimal strings, the layout
of test_vectors/Groth16/bn254/multiplier2/circom.proof) plus public.json.
"""
import argparse
import json
import secrets
import time

import numpy as np

from . import binding as B

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
BN254_Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def _canon(lib, arr, field):
    out = np.zeros_like(arr)
    fn = lib.cs_fq_from_mont if field == "fq" else lib.cs_fr_from_mont
    fn(B.CS_BN254, B._ptr(np.ascontiguousarray(arr)), B._ptr(out), arr.size // 4)
    return B.limbs_to_ints(out.reshape(-1, 4))


def proof_json(lib, A, Bp, C):
    a, b, c = _canon(lib, A, "fq"), _canon(lib, Bp, "fq"), _canon(lib, C, "fq")
    return {"pi_a": [str(a[0]), str(a[1]), "1"],
            "pi_b": [[str(b[0]), str(b[1])], [str(b[2]), str(b[3])], ["1", "0"]],
            "pi_c": [str(c[0]), str(c[1]), "1"], "protocol": "groth16", "curve": "bn128"}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--zkey", required=True)
    ap.add_argument("--wtns", required=True)
    ap.add_argument("--out", default="proof.json")
    ap.add_argument("--public-out", default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--lib", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args(argv)
    ctx = B.Context(args.device, lib_path=args.lib)
    t0 = time.time()
    pk = B.Groth16Key.from_zkey(ctx, args.zkey)
    wit = B.read_wtns(ctx.lib, args.wtns)
    t1 = time.time()
    rs = B.ints_to_limbs(B.to_mont_ints([secrets.randbelow(BN254_R), secrets.randbelow(BN254_R)], BN254_R, 4), 4)
    A, Bp, C = pk.prove_plain(np.ascontiguousarray(wit[:pk.ni]), np.ascontiguousarray(wit[pk.ni:]), rs[0:1], rs[1:2])
    t2 = time.time()
    with open(args.out, "w") as f:
        json.dump(proof_json(ctx.lib, A, Bp, C), f)
    if args.public_out:
        with open(args.public_out, "w") as f:
            json.dump([str(x) for x in _canon(ctx.lib, wit[1:pk.ni], "fr")], f)
    print("key+witness load %.1f ms, Generate proof took %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    pk.free()
    ctx.close()


if __name__ == "__main__":
    main()
