"""Posterior post-processing of CarmaSample (carma_pack.py:286-315, 439-546) without a GPU: AR roots, MA coefficients
and sigma of every sample against the values the REFERENCE's own CarmaSample methods give for the same theta rows
(tests/golden/derived_params.npz, made by tests/golden/make_golden.py), and `loglik` without re-filtering."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_case_names


@pytest.fixture(scope="module")
def derived():
    return dict(np.load(os.path.join(GOLDEN, "derived_params.npz")))


def test_roots_ma_sigma_match_reference_post_processing(loglik_cases, derived):
    from carma_pack_b200 import CarmaSample
    t, y, e = loglik_cases["t60"], loglik_cases["y60"], loglik_cases["ysig60"]
    checked = 0
    for name in golden_case_names(loglik_cases):
        p, q = int(loglik_cases[name + "_p"]), int(loglik_cases[name + "_q"])
        th = loglik_cases[name + "_theta"]
        if th.shape[1] != 3 + p + q:
            continue  # ZCARMA rows carry kappa: not a CarmaSample trace
        lp = loglik_cases[name + "_logpost"]
        # no GPU is touched: the device series is only created by predict / simulate / kalman_filter
        cs = CarmaSample(t, y, e, trace=th, logpost=lp, p=p, q=q, postprocess="numpy")
        assert cs._series_obj is None
        roots = cs._samples["ar_roots"]
        want = derived[name + "_roots"]
        assert np.allclose(roots, want, rtol=1e-13, atol=0), name
        ma = np.zeros((th.shape[0], p))
        ma[:, :cs._samples["ma_coefs"].shape[1]] = cs._samples["ma_coefs"]
        assert np.allclose(ma, derived[name + "_ma"], rtol=1e-12, atol=1e-15), name
        assert np.allclose(np.ravel(cs._samples["sigma"]) ** 2, derived[name + "_sigsqr"], rtol=1e-10), name
        assert np.allclose(cs._samples["psd_centroid"], np.abs(want.imag) / (2 * np.pi)) and \
            np.allclose(cs._samples["psd_width"], -want.real / (2 * np.pi))
        # loglik: the reference re-filters every sample with SetMLE(True) (carma_pack.py:307-313), which returns
        # loglik + LogPrior = the stored log-posterior for an in-prior sample; loglik_only removes the prior
        assert np.array_equal(np.ravel(cs._samples["loglik"]), lp)
        fin = np.isfinite(lp)
        assert np.allclose(np.ravel(cs._samples["loglik_only"])[fin], loglik_cases[name + "_loglik"][fin], rtol=1e-10, atol=1e-8), name
        assert np.allclose(CarmaSample.log_prior(th)[fin], loglik_cases[name + "_logprior"][fin], rtol=1e-12)
        checked += 1
    assert checked >= 10
