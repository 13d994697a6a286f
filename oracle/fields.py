"""Curve / field constants (oracle; test infrastructure only).

BN254 and BLS12-381 parameters.  q = base field, r = scalar field.  The snarkjs root-of-unity
construction restates co-groth16/src/groth16.rs:60-100.
"""
from functools import lru_cache


class Curve:
    def __init__(self, name, q, r, b, g1, b2, g2, two_adicity, nlimbs64_q):
        self.name, self.q, self.r, self.b, self.g1 = name, q, r, b, g1
        self.b2, self.g2 = b2, g2  # twist coefficient (Fq2 tuple), G2 generator ((x0,x1),(y0,y1))
        self.two_adicity = two_adicity
        self.nq = nlimbs64_q  # 64-bit limbs of Fq
        self.nr = 4


BN254 = Curve(
    "bn254",
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    3,
    (1, 2),
    # 3/(9+u)
    (19485874751759354771024239261021720505790618469301721065564631296452457478373,
     266929791119991161246907387137283842545076965332900288569378510910307636690),
    ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
      11559732032986387107991004021392285783925812861821192530917403151452391805634),
     (8495653923123431417604973247489272438418190587263600148770280649306958101930,
      4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    28, 4)

BLS12_381 = Curve(
    "bls12_381",
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    4,
    (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
     0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    (4, 4),
    ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
      0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
     (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
      0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    32, 6)

CURVES = {"bn254": BN254, "bls12_381": BLS12_381}


def curve_by_q(q):
    for c in CURVES.values():
        if c.q == q:
            return c
    raise ValueError("unknown base field")


def inv(a, p):
    return pow(a, p - 2, p)


@lru_cache(maxsize=None)
def roots_of_unity(r):
    """co-groth16/src/groth16.rs:60-73: smallest QNR q, z = q^TRACE, roots[k] = primitive 2^k-th root."""
    s, t = 0, r - 1
    while t % 2 == 0:
        t //= 2
        s += 1
    qnr = 1
    while pow(qnr, (r - 1) // 2, r) != r - 1:
        qnr += 1
    roots = [pow(qnr, t, r)]
    for _ in range(s):
        roots.append(roots[-1] * roots[-1] % r)
    roots.reverse()
    return qnr, tuple(roots)


def groth16_roots_of_unity(r, power):
    """co-groth16/src/groth16.rs:91-100 -> (group_gen, coset_shift)."""
    qnr, roots = roots_of_unity(r)
    two_adicity = len(roots) - 1
    gen = roots[power]
    shift = qnr * qnr % r if power == two_adicity else roots[power + 1]
    return gen, shift
