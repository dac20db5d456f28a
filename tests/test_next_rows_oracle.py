"""CPU: the numpy oracle of the SURVEY §8f rows (loss, densification bookkeeping) against golden vectors produced
by the reference's own code (tests/golden/make_next_rows_golden.py -> next_rows.npz).  This row's parity is PINNED."""
import os

import numpy as np
import pytest

from oracle import next_rows as NR

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows.npz"))


def test_window_matches_reference_formula():
    g = NR.window_1d()
    assert g.dtype == np.float32 and g.shape == (11,)
    assert abs(float(g.sum()) - 1.0) < 1e-6
    assert np.allclose(g, g[::-1]) and g.argmax() == 5
    assert np.array_equal(g, GOLD["window_1d"])               # bit for bit loss_utils.gaussian(11, 1.5)
    assert np.array_equal(NR.window_2d(), GOLD["window_2d"])  # and create_window's fp32 outer product


@pytest.mark.parametrize("key", ["a64", "b64", "c64", "d64", "e64"])
def test_loss_oracle_matches_reference_fp64(key):
    img, gt = GOLD[f"{key}.img"], GOLD[f"{key}.gt"]
    lam, lam_m = float(GOLD[f"{key}.lambda"]), float(GOLD[f"{key}.lambda_mask"])
    has_mask = f"{key}.opacity" in GOLD.files
    o = NR.l1_ssim_loss(img, gt, lam, GOLD[f"{key}.opacity"][0] if has_mask else None,
                        GOLD[f"{key}.mask"][0] if has_mask else None, lam_m)
    assert abs(o["l1"] - float(GOLD[f"{key}.l1"])) < 1e-13
    if lam != 0.0:
        assert abs(o["ssim"] - float(GOLD[f"{key}.ssim"])) < 1e-12
    assert abs(o["loss"] - float(GOLD[f"{key}.loss"])) < 1e-12
    ref = GOLD[f"{key}.dL_dimg"]
    assert np.abs(o["dL_dimg"] - ref).max() <= 1e-12 + 1e-9 * np.abs(ref).max()
    if has_mask:
        assert abs(o["mask_l1"] - float(GOLD[f"{key}.mask_l1"])) < 1e-13
        assert np.abs(o["dL_dopacity"] - GOLD[f"{key}.dL_dopacity"][0]).max() < 1e-15


def test_loss_oracle_vs_reference_fp32_run():
    """What the reference computes in practice (fp32 convolutions) agrees with the fp64 oracle to fp32 noise."""
    img, gt = GOLD["a32.img"], GOLD["a32.gt"]
    o = NR.l1_ssim_loss(img, gt, 0.2)
    assert abs(o["loss"] - float(GOLD["a32.loss"])) < 2e-6
    ref = GOLD["a32.dL_dimg"]
    assert np.abs(o["dL_dimg"] - ref).max() <= 2e-3 * np.abs(ref).max()


@pytest.mark.parametrize("key", ["dn0", "dn1"])
def test_densify_oracle_matches_reference(key):
    P = GOLD[f"{key}.raw_scaling"].shape[0]
    accum, denom, mr = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros(P, np.float32)
    for v in range(3):
        # the golden file was produced by torch on the CPU: its norm kernel rounds as fma (see densify_stats)
        accum, denom, mr = NR.densify_stats(GOLD[f"{key}.view{v}.grad"], GOLD[f"{key}.view{v}.radii"], accum, denom, mr,
                                            norm="cpu")
        assert np.array_equal(accum, GOLD[f"{key}.view{v}.accum"].reshape(-1))
        assert np.array_equal(denom, GOLD[f"{key}.view{v}.denom"].reshape(-1))
        assert np.array_equal(mr, GOLD[f"{key}.view{v}.max_radii2D"])
    thr, pd, ext, mino, mss = GOLD[f"{key}.params"]
    clone, split, prune = NR.densify_masks(accum, denom, GOLD[f"{key}.raw_scaling"], GOLD[f"{key}.raw_opacity"], mr,
                                           thr, pd, ext, mino, mss, raw=True)
    assert np.array_equal(clone, GOLD[f"{key}.clone"])
    assert np.array_equal(split, GOLD[f"{key}.split"])
    assert np.array_equal(prune, GOLD[f"{key}.prune"])
    assert clone.sum() > 0 and split.sum() > 0 and prune.sum() > 0      # the case exercises every branch
