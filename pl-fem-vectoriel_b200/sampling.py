"""Design sampler in front of the hot path: stratified LHS per layout, physics / quality / diversity filters.

Mirrors `SmartSampler` of the reference's `sampling.py` (`:33-372`): `generate_stratified_samples` (per-architecture
scrambled Latin hypercube, geometric + physical validation, quality threshold, ranking, greedy diversity filter,
truncation), `generate_focused_samples` (Gaussian perturbations of a reference design) and `get_sampling_stats`.

The reference imports `ParametricSpace`, `PhysicalValidator` and `SampleQualityScorer` from a module
(`parametric_space`) that is absent from its checkout, so only the FLOW of the filters is defined there.  The three
classes below are this build's definitions, kept deliberately plain and stated here:

* `ParametricSpace`: core radius 0.5-3.0 um, pitch 3-15 um (`README.md:242-243`), the twelve layouts of
  `geometry_unified.py:97-184`, bands 1490/1550/1600/1650 nm, Cauchy IP-Dip index; geometric validity is
  `MCFGeometry.validate()` (`geometry_unified.py:351-363`).
* `PhysicalValidator`: a core must guide (V >= 1.2), must not be grossly multimode (V <= 12), neighbouring cores must
  be separated by cladding (pitch >= 2.4 r) and the lantern must fit the printable field (cladding radius <= 62.5 um,
  the reference's `r_clad_SM`, `losses.py:968`).
* `SampleQualityScorer`: mean of three [0, 1] terms — closeness of V to the few-mode sweet spot (2.4-6), of
  pitch / diameter to 2-4 (coupling without overlap), and the packing efficiency relative to 0.35.

Seeds are pure functions of (base_seed, n_cores, n_target): the reference's `hash(str)` seeds (`sampling.py:161`) are
salted per process and not reproducible.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Tuple

import numpy as np

from .config import IPDipCauchy
from .geometry import MCFGeometry, SAMPLING_WEIGHTS

logger = logging.getLogger("pl_v17.sampling")


class ParametricSpace:
    def __init__(self, n_cores_options=None, core_radius_um=(0.5, 3.0), pitch_um=(3.0, 15.0),
                 wavelengths_nm=(1490, 1550, 1600, 1650), n_clad: float = 1.0):
        self.n_cores_options = list(n_cores_options) if n_cores_options is not None else sorted(SAMPLING_WEIGHTS)
        self._bounds = {"core_radius_um": tuple(map(float, core_radius_um)), "pitch_um": tuple(map(float, pitch_um))}
        self._discrete = {"wavelength_nm": list(wavelengths_nm), "taper_profile": ["linear", "exponential"],
                          "arrangement": ["default"]}
        self.n_clad = float(n_clad)

    def get_continuous_bounds(self) -> Dict[str, Tuple[float, float]]:
        return dict(self._bounds)

    def get_discrete_options(self) -> Dict[str, list]:
        return {k: list(v) for k, v in self._discrete.items()}

    def geometry(self, sample: Dict) -> MCFGeometry:
        wl = float(sample["wavelength_nm"])
        variant = sample.get("variant")
        if variant is None and sample.get("arrangement") == "pentagon_center":
            variant = "pentagon_center"
        return MCFGeometry(int(sample["n_cores"]), float(sample["pitch_um"]), float(sample["core_radius_um"]),
                           float(sample.get("n_core", IPDipCauchy.n(wl))), float(sample.get("n_clad", self.n_clad)),
                           wl / 1000.0, variant=variant)

    def validate_sample_geometry(self, sample: Dict) -> Tuple[bool, str]:
        for name, (lo, hi) in self._bounds.items():
            if not (lo <= sample[name] <= hi):
                return False, f"{name} outside [{lo}, {hi}]"
        try:
            return self.geometry(sample).validate()
        except ValueError as e:
            return False, str(e)


class PhysicalValidator:
    V_MIN, V_MAX, PITCH_OVER_R_MIN, CLAD_MAX_UM = 1.2, 12.0, 2.4, 62.5

    def __init__(self, space: Optional[ParametricSpace] = None):
        self.space = space or ParametricSpace()

    def validate_sample_physics(self, sample: Dict) -> Tuple[bool, str, Dict]:
        g = self.space.geometry(sample)
        m = dict(V_number=float(g.V_number), NA=float(np.sqrt(max(g.n_core ** 2 - g.n_clad ** 2, 0.0))),
                 pitch_ratio=float(g.pitch_ratio), packing_efficiency=float(g.packing_efficiency),
                 cladding_radius_um=float(g.cladding_radius), n_core=float(g.n_core))
        if m["V_number"] < self.V_MIN:
            return False, f"V = {m['V_number']:.2f}: the cores barely guide", m
        if m["V_number"] > self.V_MAX:
            return False, f"V = {m['V_number']:.2f}: grossly multimode cores", m
        if g.n_cores > 1 and g.pitch < self.PITCH_OVER_R_MIN * g.r_core:
            return False, "cores closer than 2.4 radii", m
        if m["cladding_radius_um"] > self.CLAD_MAX_UM:
            return False, "lantern larger than the 62.5 um cladding", m
        return True, "OK", m


class SampleQualityScorer:
    @staticmethod
    def _window(x: float, lo: float, hi: float, fall: float) -> float:
        """1 inside [lo, hi], linear fall-off to 0 over `fall` outside."""
        if x < lo:
            return max(0.0, 1.0 - (lo - x) / fall)
        if x > hi:
            return max(0.0, 1.0 - (x - hi) / fall)
        return 1.0

    def score_sample(self, sample: Dict, metrics: Dict) -> float:
        sv = self._window(metrics["V_number"], 2.4, 6.0, 4.0)
        sp = self._window(metrics["pitch_ratio"], 2.0, 4.0, 3.0)
        sk = min(1.0, metrics["packing_efficiency"] / 0.35)
        return float((sv + sp + sk) / 3.0)


class SmartSampler:
    def __init__(self, space: Optional[ParametricSpace] = None, config=None, base_seed: int = 42):
        self.space = space or ParametricSpace()
        self.config = config
        self.validator = PhysicalValidator(self.space)
        self.scorer = SampleQualityScorer()
        self.base_seed = int(base_seed)
        self.rng = np.random.default_rng(self.base_seed)
        self.total_generated = 0
        self.total_valid = 0
        self.generation_history: List[Dict] = []

    # -- `sampling.py:69-141` ------------------------------------------------------------------------------------
    def generate_stratified_samples(self, n_samples: int, apply_filter: bool = True, quality_threshold: float = 0.35,
                                    oversample_factor: float = 3.0, ensure_diversity: bool = True,
                                    min_distance: float = 0.05) -> List[Dict]:
        options = self.space.n_cores_options
        if not options:
            raise ValueError("ParametricSpace.n_cores_options is empty")
        per_arch = max(1, n_samples // len(options))
        samples: List[Dict] = []
        for n_cores in options:
            samples.extend(self._lhs_for_architecture(n_cores, per_arch, apply_filter, quality_threshold, oversample_factor))
        missing = n_samples - len(samples)
        if missing > 0:                       # top up from one more architecture, as the reference does
            extra = int(self.rng.choice(options))
            samples.extend(self._lhs_for_architecture(extra, missing, apply_filter, quality_threshold, oversample_factor))
        if ensure_diversity and len(samples) > 1:
            samples = self._ensure_diversity(samples, min_distance)
        samples = samples[:n_samples]
        self.total_generated += int(n_samples * oversample_factor)
        self.total_valid += len(samples)
        self.generation_history.append(dict(kind="stratified", requested=n_samples, returned=len(samples)))
        return samples

    def _seed(self, n_cores: int, n_target: int) -> int:
        return (self.base_seed * 1_000_003 + n_cores * 10_007 + n_target * 101) % (2 ** 31)

    # -- `sampling.py:143-233` -----------------------------------------------------------------------------------
    def _lhs_for_architecture(self, n_cores: int, n_target: int, apply_filter: bool, quality_threshold: float,
                              oversample_factor: float) -> List[Dict]:
        from scipy.stats import qmc
        bounds = self.space.get_continuous_bounds()
        discrete = self.space.get_discrete_options()
        n_gen = max(int(n_target * oversample_factor) if apply_filter else n_target, 1)
        seed = self._seed(n_cores, n_target)
        names = list(bounds)
        raw = qmc.LatinHypercube(d=len(names), scramble=True, seed=seed).random(n=n_gen)
        scaled = qmc.scale(raw, [bounds[n][0] for n in names], [bounds[n][1] for n in names])
        out: List[Dict] = []
        rejected = dict(geom=0, phys=0, quality=0)
        for idx, row in enumerate(scaled):
            s: Dict = {n: float(v) for n, v in zip(names, row)}
            local = np.random.default_rng(seed + idx)
            s["n_cores"] = int(n_cores)
            s["wavelength_nm"] = int(local.choice(discrete["wavelength_nm"]))
            s["taper_profile"] = str(local.choice(discrete["taper_profile"]))
            s["arrangement"] = str(local.choice(["ring", "pentagon_center"])) if n_cores == 6 else str(local.choice(discrete["arrangement"]))
            s["sample_id"] = f"S_{n_cores}C_{len(out):04d}"
            ok, _ = self.space.validate_sample_geometry(s)
            if not ok:
                rejected["geom"] += 1
                continue
            if apply_filter:
                ok, _, metrics = self.validator.validate_sample_physics(s)
                if not ok:
                    rejected["phys"] += 1
                    continue
                score = self.scorer.score_sample(s, metrics)
                if score < quality_threshold:
                    rejected["quality"] += 1
                    continue
                s.update(metrics)
                s["quality_score"] = score
            out.append(s)
            if not apply_filter and len(out) >= n_target:
                break
        logger.debug("%d cores: %d of %d kept (%s)", n_cores, len(out), n_gen, rejected)
        if apply_filter and out:
            out = sorted(out, key=lambda s: s.get("quality_score", 0.0), reverse=True)
        return out[:n_target]

    # -- `sampling.py:235-288` -----------------------------------------------------------------------------------
    def _ensure_diversity(self, samples: List[Dict], min_distance: float) -> List[Dict]:
        """Greedy: keep a sample when it is at least `min_distance` (Euclidean, bounds-normalised) from all kept ones."""
        if len(samples) < 2:
            return samples
        from scipy.spatial.distance import pdist, squareform
        bounds = self.space.get_continuous_bounds()
        X = np.array([[(s[n] - lo) / (hi - lo + 1e-12) if n in s else 0.0 for n, (lo, hi) in bounds.items()] for s in samples])
        D = squareform(pdist(X, metric="euclidean"))
        kept = [0]
        for i in range(1, len(samples)):
            if np.min(D[i, kept]) >= min_distance:
                kept.append(i)
        return [samples[i] for i in kept]

    # -- `sampling.py:290-348` -----------------------------------------------------------------------------------
    def generate_focused_samples(self, reference: Dict, n_samples: int, rel_variation: float = 0.15,
                                 min_distance: Optional[float] = 0.02) -> List[Dict]:
        bounds = self.space.get_continuous_bounds()
        key = tuple(sorted((k, repr(v)) for k, v in reference.items()))
        import zlib
        local = np.random.default_rng(self.base_seed + zlib.crc32(repr(key).encode()) % (2 ** 31))
        out: List[Dict] = []
        for i in range(n_samples * 3):
            s = dict(reference)
            for name, (lo, hi) in bounds.items():
                if name in s:
                    s[name] = float(np.clip(local.normal(s[name], rel_variation * (hi - lo) / 3.0), lo, hi))
            s["sample_id"] = f"FOCUS_{i:04d}_{reference.get('sample_id', 'REF')}"
            if not self.space.validate_sample_geometry(s)[0]:
                continue
            if min_distance and out and min(self._sample_distance(s, o) for o in out) < min_distance:
                continue
            out.append(s)
            if len(out) >= n_samples:
                break
        return out[:n_samples]

    def _sample_distance(self, s1: Dict, s2: Dict) -> float:
        d = [(s1[n] - s2[n]) / (hi - lo) for n, (lo, hi) in self.space.get_continuous_bounds().items()
             if n in s1 and n in s2 and hi > lo]
        return float(np.sqrt(np.mean(np.square(d)))) if d else 0.0

    def get_sampling_stats(self) -> Dict:
        return dict(total_generated=self.total_generated, total_valid=self.total_valid,
                    validation_rate=self.total_valid / max(self.total_generated, 1), base_seed=self.base_seed,
                    n_calls=len(self.generation_history))


def samples_to_designs(samples: List[Dict]) -> List[Dict]:
    """Sampler output -> the design dicts `sweep.run_sweep` takes (n_modes = min(3 N_cores, 40))."""
    out = []
    for s in samples:
        wl = int(s["wavelength_nm"])
        out.append(dict(n_cores=int(s["n_cores"]), core_radius_um=float(s["core_radius_um"]), pitch_um=float(s["pitch_um"]),
                        wavelength_nm=wl, n_core=float(s.get("n_core", IPDipCauchy.n(wl))), n_clad=float(s.get("n_clad", 1.0)),
                        n_modes=min(3 * int(s["n_cores"]), 40),
                        variant="pentagon_center" if s.get("arrangement") == "pentagon_center" else None,
                        sample_id=s.get("sample_id")))
    return out
