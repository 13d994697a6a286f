"""The C++ host-side mirror of the reference interface (include/co_groth16.hpp) driven like the
reference's own tests (tests/cpp/test_co_groth16.cpp): plain prove == oracle proof for fixed (r, s),
prove_inner's error message, and a 3-thread LocalNetwork Rep3 run whose proof all parties share.
CPU: linked against the kernel emulation build; GPU (-m gpu): against libcosnarks_gpu.so."""
import os
import random
import struct
import subprocess
import sys

import numpy as np
import pytest

from helpers import Conv, golden_groth16, ih
from oracle import groth16 as OG
from oracle.fields import BN254

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_fixture(path, name="multiplier2"):
    cv = Conv("bn254")
    z, m, w, g = golden_groth16(name)
    ni = m["num_instance_variables"]
    pr = g["oracle_proofs"][-1]
    r_, s_ = ih(pr["r"]), ih(pr["s"])
    exp = OG.prove_plain(z, m, w, r_, s_)
    wsh = OG.share_rep3(w[ni:], cv.r, random.Random(5))

    def vec(f, arr, dtype):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        f.write(struct.pack("<Q", arr.size // {np.uint32: 1, np.uint64: 1}.get(dtype, 1)))
        f.write(arr.tobytes())

    with open(path, "wb") as f:
        f.write(struct.pack("<QQQ", m["num_constraints"], ni, m["num_witness_variables"]))
        for mat in (m["a"], m["b"]):
            rp, col, cf = cv.csr(mat)
            vec(f, rp, np.uint32)
            vec(f, col, np.uint32)
            vec(f, cf.reshape(-1), np.uint64)
        for arr in (cv.g1([z["alpha_g1"]]), cv.g1([z["beta_g1"]]), cv.g2([z["beta_g2"]]), cv.g1([z["delta_g1"]]),
                    cv.g2([z["delta_g2"]]), cv.g1(z["a_query"]), cv.g1(z["b_g1_query"]), cv.g2(z["b_g2_query"]),
                    cv.g1(z["l_query"]), cv.g1(z["h_query"])):
            vec(f, arr.reshape(-1), np.uint64)

        def frvec(vals):  # count = number of Fr elements
            a = cv.fr(vals)
            f.write(struct.pack("<Q", a.shape[0]))
            f.write(a.tobytes())

        frvec(w[:ni])
        frvec(w[ni:])
        frvec([r_, s_])
        vec(f, cv.g1([exp[0]]).reshape(-1), np.uint64)
        vec(f, cv.g2([exp[1]]).reshape(-1), np.uint64)
        vec(f, cv.g1([exp[2]]).reshape(-1), np.uint64)
        for i in range(3):
            a = cv.fr([x for ab in wsh[i] for x in ab])
            f.write(struct.pack("<Q", len(wsh[i])))
            f.write(a.tobytes())
        vec(f, cv.g1([BN254.g1]).reshape(-1), np.uint64)


def _build_and_run(tmp_path, lib_path):
    fx = str(tmp_path / "fixture.bin")
    _write_fixture(fx)
    exe = str(tmp_path / "test_co_groth16")
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_co_groth16.cpp"), "-o", exe,
                           "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, fx], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "checks passed" in out.stdout


def _write_plonk_fixture(tmp_path, name="multiplier2"):
    """-> (zkey path, fixture path): witness, the golden proof for b = [0..11), replicated shares of both."""
    import zkey_writer
    from helpers import golden_plonk, plonk_proof_from_json
    cv = Conv("bn254")
    z, w, g = golden_plonk(name)
    npub = z["n_public"]
    zp, fx = str(tmp_path / "plonk.zkey"), str(tmp_path / "plonk_fixture.bin")
    zkey_writer.write_plonk_zkey(zp, z)
    exp = plonk_proof_from_json(g["oracle_proof_deterministic_blinders"])
    rng = random.Random(61)

    def shares_of(vals):
        out = [[], [], []]
        for v in vals:
            s0, s1 = rng.randrange(cv.r), rng.randrange(cv.r)
            sh = [s0, s1, (v - s0 - s1) % cv.r]
            for p in range(3):
                out[p] += [sh[p], sh[(p + 2) % 3]]
        return out
    with open(fx, "wb") as f:
        def frvec(vals, per=1):
            a = cv.fr(vals)
            f.write(struct.pack("<Q", a.shape[0] // per))
            f.write(a.tobytes())
        frvec(w[:npub + 1])
        frvec(w[npub + 1:])
        pts = cv.g1([exp[k] for k in ("a", "b", "c", "z", "t1", "t2", "t3", "wxi", "wxiw")])
        f.write(struct.pack("<Q", 9))
        f.write(pts.tobytes())
        frvec([exp[k] for k in ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw")])
        for part in shares_of(w[npub + 1:]):
            frvec(part, 2)
        for part in shares_of(list(range(11))):
            frvec(part, 2)
    return zp, fx


def _build_and_run_plonk(tmp_path, lib_path, extra_args=()):
    zp, fx = _write_plonk_fixture(tmp_path)
    exe = str(tmp_path / "test_co_plonk")
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_co_plonk.cpp"), "-o", exe,
                           "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, zp, fx] + list(extra_args), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "checks passed" in out.stdout


def test_cpp_plonk_mirror_on_emulated_kernels(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    _build_and_run_plonk(tmp_path, build_emu.build(), ["library-driver"])  # + Rep3CoPlonk::prove_in_library


@pytest.mark.gpu
def test_cpp_plonk_mirror_on_gpu(tmp_path):
    from co_snarks_b200 import binding as B
    _build_and_run_plonk(tmp_path, B.DEFAULT_LIB)


def test_cpp_mirror_on_emulated_kernels(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    _build_and_run(tmp_path, build_emu.build())


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    from co_snarks_b200 import binding as B
    _build_and_run(tmp_path, B.DEFAULT_LIB)
