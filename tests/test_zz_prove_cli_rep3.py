"""The two GPU checks added after the last full GPU run of the round, kept in a file that sorts last:
 * `python -m co_snarks_b200.prove --rep3-shares`: three parties of one box from their share files (replicated and
   additive / compressed forms), opened proof accepted under the fixture's verification key;
 * the C++ mirror's Rep3CoPlonk::prove_in_library (cs_plonk_rep3_prove over the callback transport)."""
import pytest

import kernel_checks as K


@pytest.mark.gpu
def test_prove_cli_rep3_share_files_gpu(tmp_path):
    K.check_prove_cli_rep3_shares(None, tmp_path, "multiplier2")


@pytest.mark.gpu
def test_cpp_plonk_library_driver_on_gpu(tmp_path):
    from co_snarks_b200 import binding as B
    import test_cpp_mirror as M
    M._build_and_run_plonk(tmp_path, B.DEFAULT_LIB, ["library-driver"])
