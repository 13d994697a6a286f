"""CPU oracle for the co-snarks hot path (TEST INFRASTRUCTURE ONLY).

Pure-Python big-int restatement of the reference algorithm (TaceoLabs/co-snarks @ 2b4592e) for
the Groth16 / Plonk-round-1 hot path: snarkjs roots of unity, radix-2 NTT, MSM, the
`CircomReduction` witness map, Groth16 proof assembly with injected (r, s), Rep3 share emulation,
file-format readers and a BN254 pairing verifier.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this package -- never the product (`co_snarks_b200/`).  Each function cites the reference
file:line it restates.  The arithmetic lives in the un-vendored crates `taceo-ark-algebra 0.1.0` and
arkworks 0.6 (Cargo.lock:4771), so parity is anchored on the reference's own fixtures and KATs:
Plonk round-1 commitments (co-plonk/src/round1.rs:351-371, 397-417) pin iNTT+MSM bit-for-bit;
the snarkjs proofs + verification keys under test_vectors/Groth16 pin the verifier and the prover.
"""
