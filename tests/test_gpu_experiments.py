"""GPU: the parametrised DME experiment runner emits the reference's pickle keys (ND:227-259) and both
NMSE conventions; the type-quantizer lines land where the reference's do."""
import pickle

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_runner_keys_conventions_and_levels(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dme_b200 import experiments as ex
    users = [1, 11]
    avg, mx, std_avg = ex.run("normal", dim=2048, num_users_list=users, num_instances=24, seed=7, kashin=False, out_dir=str(tmp_path))
    assert sorted(avg) == sorted(f"NMSE_{ln}_avg" for ln in ex.LINES) and sorted(mx) == sorted(f"NMSE_{ln}_max" for ln in ex.LINES)
    for k, v in avg.items():
        assert v.shape == (len(users),)
        if "Kashin" in k:
            assert np.isnan(v).all()                     # kashin=False
        else:
            assert np.isfinite(v).all() and (v > 0).all(), k
            assert (mx[k.replace("_avg", "_max")] >= v).all()
    # F9: reference convention = standard / (50 n^2)
    for ui, n in enumerate(users):
        r = avg["NMSE_Type_Unbiased_1bit_avg"][ui] / std_avg["NMSE_Type_Unbiased_1bit_avg"][ui]
        assert abs(r * 50 * n * n - 1.0) < 1e-4
    # SURVEY F9 sanity: unbiased type quantizer, R = 1, Gaussian: standard NMSE * n ~ 2.0 (reference: 0.19969 * 10 at n = 10)
    s = std_avg["NMSE_Type_Unbiased_1bit_avg"][1] * users[1]
    assert 1.7 < s < 2.3, s
    # more bits, less error; biased beats unbiased at n = 1 (no averaging to profit from unbiasedness)
    assert avg["NMSE_Type_Unbiased_2bit_avg"][1] < avg["NMSE_Type_Unbiased_1bit_avg"][1]
    assert avg["NMSE_Type_Biased_1bit_avg"][0] < avg["NMSE_Type_Unbiased_1bit_avg"][0]
    assert avg["NMSE_Scalar_4bit_avg"][1] < avg["NMSE_Scalar_1bit_avg"][1]
    with open(tmp_path / "nmse_avg_data_Normal_dist.pkl", "rb") as f:
        assert sorted(pickle.load(f)) == sorted(avg)
