"""`sampling.SmartSampler`: the reference's sampling flow (`sampling.py:69-348`) with this build's validators."""
import numpy as np

from plfem_b200.sampling import ParametricSpace, PhysicalValidator, SmartSampler, samples_to_designs
from plfem_b200 import sweep


def test_stratified_samples_are_valid_reproducible_and_diverse():
    a = SmartSampler(base_seed=42).generate_stratified_samples(60)
    b = SmartSampler(base_seed=42).generate_stratified_samples(60)
    c = SmartSampler(base_seed=43).generate_stratified_samples(60)
    assert 25 <= len(a) <= 60 and a == b and a != c                       # seeds are pure functions of their inputs
    space, val = ParametricSpace(), PhysicalValidator()
    for s in a:
        assert space.validate_sample_geometry(s)[0] and val.validate_sample_physics(s)[0]
        assert s["quality_score"] >= 0.35 and 0.5 <= s["core_radius_um"] <= 3.0 and 3.0 <= s["pitch_um"] <= 15.0
    assert len({s["n_cores"] for s in a}) >= 8                            # stratified over the layouts
    X = np.array([[(s["core_radius_um"] - 0.5) / 2.5, (s["pitch_um"] - 3.0) / 12.0] for s in a])
    d = np.linalg.norm(X[:, None] - X[None], axis=2) + np.eye(len(a))
    assert d.min() >= 0.05 - 1e-12                                        # greedy diversity filter (`sampling.py:235-288`)


def test_filters_reject_and_rank():
    smp = SmartSampler(base_seed=1)
    raw = smp._lhs_for_architecture(7, 20, False, 0.0, 1.0)
    kept = smp._lhs_for_architecture(7, 20, True, 0.5, 3.0)
    assert len(raw) <= 20 and all("quality_score" not in s for s in raw)
    scores = [s["quality_score"] for s in kept]
    assert scores == sorted(scores, reverse=True) and min(scores) >= 0.5   # ranked, thresholded
    ok, msg, m = PhysicalValidator().validate_sample_physics(dict(n_cores=7, core_radius_um=0.5, pitch_um=14.0, wavelength_nm=1650))
    assert ok and m["V_number"] > 1.2
    ok, msg, _ = PhysicalValidator().validate_sample_physics(dict(n_cores=19, core_radius_um=2.9, pitch_um=6.0, wavelength_nm=1490))
    assert not ok


def test_focused_samples_stay_near_the_reference_design():
    smp = SmartSampler()
    ref = dict(n_cores=7, core_radius_um=1.5, pitch_um=8.0, wavelength_nm=1550, sample_id="REF7")
    out = smp.generate_focused_samples(ref, 10, rel_variation=0.1)
    assert len(out) == 10 and out == SmartSampler().generate_focused_samples(ref, 10, rel_variation=0.1)
    assert all(abs(s["core_radius_um"] - 1.5) < 0.5 and abs(s["pitch_um"] - 8.0) < 2.5 for s in out)
    assert smp.get_sampling_stats()["base_seed"] == 42


def test_samples_feed_the_sweep_driver():
    designs = samples_to_designs(SmartSampler().generate_stratified_samples(24))
    for d in designs:
        g = sweep.design_geometry(d)
        assert g.validate()[0] and d["n_modes"] == min(3 * d["n_cores"], 40)
