"""Kaggle submission format: the host-side C writer / parser of libpcm_b200.so against the
loop-for-loop restatement of the reference (oracle/kaggle_oracle.py), and — in the build container, where
/root/reference exists — the UNMODIFIED _climate_kaggle_metric.score on the rows our writer produced.
CPU tests call only the HOST entry points (no kernels); the device scorer is covered by the -m gpu tests."""
import json
import os

import numpy as np
import pytest

import pcm_b200  # noqa: F401
from oracle import kaggle_oracle as KO
from oracle import ref_loader
from pcm_b200 import kaggle as K

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    with open(os.path.join(GOLD, "kaggle_roundtrip.json")) as f:
        return json.load(f)


def test_writer_matches_reference_loop(tmp_path):
    pred, true, lat, lon, names = KO.synth_submission(T=4, seed=7)
    ids, vals = KO.convert_predictions_to_kaggle_format(pred, np.arange(4), lat, lon, names)
    df = K.convert_predictions_to_kaggle_format(pred, np.arange(4), lat, lon, names)
    assert list(df.columns) == ["ID", "Prediction"]
    assert df["ID"].tolist() == ids
    assert np.array_equal(df["Prediction"].to_numpy(), vals)
    g = _gold()
    assert ids[:3] == g["first_ids"] and ids[-1] == g["last_id"] and len(ids) == g["n_rows"]
    # the file: same rows, values read back to the same float32 bits
    import pandas as pd
    path = tmp_path / "submission.csv"
    K.write_submission_csv(str(path), pred, lat, lon, names)
    back = pd.read_csv(path)
    assert list(back.columns) == ["ID", "Prediction"]
    assert back["ID"].tolist() == ids
    assert np.array_equal(back["Prediction"].to_numpy().astype(np.float32), vals.astype(np.float32))
    ref_path = tmp_path / "ref.csv"
    pd.DataFrame({"ID": ids, "Prediction": vals}).to_csv(ref_path, index=False)      # what the reference saves
    ref_back = pd.read_csv(ref_path)
    assert np.array_equal(ref_back["Prediction"].to_numpy().astype(np.float32), back["Prediction"].to_numpy().astype(np.float32))


def test_writer_edge_cases():
    # empty time axis, single cell, negative / three-digit coordinates, >999 time steps (t1000: %03d grows)
    assert K.format_ids(0, [0.0], [0.0], ["tas"]) == []
    assert K.format_ids(1, [-88.586387], [356.875], ["pr"]) == ["t000_pr_-88.59_356.88"]
    ids = K.format_ids(1001, [1.0], [2.0], ["tas"])
    assert ids[-1] == "t1000_tas_1.00_2.00" and len(ids) == 1001
    lat = np.array([-0.004, 0.005, 89.995]); lon = np.array([0.125, 359.999])
    want, _ = KO.convert_predictions_to_kaggle_format(np.zeros((2, 1, 3, 2)), range(2), lat, lon, ["tas"])
    assert K.format_ids(2, lat, lon, ["tas"]) == want


def test_parser_matches_reference_regex():
    pred, _, lat, lon, names = KO.synth_submission(T=3, seed=1)
    ids, _ = KO.convert_predictions_to_kaggle_format(pred, np.arange(3), lat, lon, names)
    ids += ["t7_pr_12_-3.", "t010_tas_-0.50_10.25trailing", "t3_x_1.5_2"]           # shapes the regex accepts
    want = KO.parse_ids(ids)
    time, code, vnames, la, lo = K.parse_ids(ids)
    assert time.tolist() == [w[0] for w in want]
    assert [vnames[c] for c in code] == [w[1] for w in want]
    assert la.tolist() == [w[2] for w in want] and lo.tolist() == [w[3] for w in want]
    assert vnames == ["tas", "pr", "x"]                                             # first-appearance order
    for bad in ["x000_tas_1.0_2.0", "t_tas_1.0_2.0", "t000_TAS_1.0_2.0", "t000_tas_1.0", "t000_tas_.5_2.0", ""]:
        with pytest.raises(ValueError, match="Invalid ID format"):
            KO.parse_ids([bad])
        with pytest.raises(ValueError, match="Invalid ID format"):
            K.parse_ids(ids[:5] + [bad])
    assert K.parse_ids([])[0].size == 0


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_unmodified_reference_score_on_our_rows():
    """Round trip: our writer -> the reference's own scorer == the pinned golden."""
    R = ref_loader.load()
    pred, true, lat, lon, names = KO.synth_submission(T=4, seed=7)
    sol = K.convert_predictions_to_kaggle_format(true, np.arange(4), lat, lon, names)
    sub = K.convert_predictions_to_kaggle_format(pred, np.arange(4), lat, lon, names)
    sol, sub = sol.astype({"Prediction": np.float64}), sub.astype({"Prediction": np.float64})
    assert abs(float(R.score(sol, sub, "ID")) - _gold()["reference_score"]) < 1e-12
