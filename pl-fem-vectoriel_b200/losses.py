"""Loss columns for vectorial H-field modes — host-side consumer of the hot path's mode records.

Mirrors the part of the reference's `losses.py` that `LossCalculator.calculate_physical_losses` runs for
modes with ``is_vectorial = True`` (`losses.py:742-826`): design parameters rebuilt from the geometry
(`:871-989`), the three sections of `VectorialLossCalculator` (`:1012-1221`), the spectral-spread crosstalk
estimate (`:546-619`), the summed-power PDL (`:445-468`) and the confinement-based radiation penalty (`:693-716`).
Everything is closed-form arithmetic on at most a few dozen mode scalars: it stays on the host.

`config.PhotonicLanternDesignParameters`, which the reference imports, is absent from its checkout
(SURVEY.md section 0); the dataclass below carries the fields `_build_design_params` sets.

The scalar route (`EnhancedLossCalculator.calculate_sectional_losses`, `losses.py:74-440`) serves the scalar
solver's records and is not part of this build: scalar modes get ``{'success': False, ...}``.

Checked against values produced by the REAL reference module (`tests/golden/make_golden.py::golden_losses`,
`tests/golden/losses.json`), including the reference's own self-check numbers XT = -25.31 dB, PDL = 0.878 dB
(`losses.py:1228-1259`).
"""
from __future__ import annotations

import logging
import math
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np

logger = logging.getLogger("pl_v18.losses")


@dataclass
class PhotonicLanternDesignParameters:
    """Fields of the (absent) reference class as `losses.py:952-985` fills them."""
    N_cores: int = 7
    has_central_core: bool = True
    config_type: str = "hexagonal"
    geometry_config: str = "7-hexagonal"
    n_peripheral_cores: int = 6
    R_ring: float = 8.0
    packing_efficiency: float = 0.2
    pitch: float = 8.0
    pitch_min: float = 8.0
    pitch_ratio: float = 2.7
    wavelength: float = 1550.0
    r_core_SM: float = 1.5
    r_clad_SM: float = 62.5
    n_core_SM: float = 1.535
    n_clad_SM: float = 1.0
    V_SM: float = 5.0
    NA_SM: float = 1.16
    MFD: float = 3.0
    n_eff_LP01: float = 1.5
    r_core_MM: float = 25.0
    V_MM: float = 13.0
    NA_MM: float = 0.22
    M_max: int = 1
    n_polymer: float = 1.535
    d_polymer: float = 2.0
    coupling_uniformity: float = 0.95
    L_mux: float = 200.0
    L_taper: float = 375.0
    L_MMF: float = 100.0
    L_total: float = 675.0
    n_taper: float = 1.0
    taper_profile: str = "exponential"


def _first(x) -> float:
    return float(np.asarray(x).flat[0])


def _clip(v, lo, hi) -> float:
    return float(np.clip(v, lo, hi))


class EnhancedLossCalculator:
    """The mode-list estimators the vectorial route uses (static methods, names as in the reference)."""

    @staticmethod
    def _calculate_pdl_vectorial(modes: Sequence[Dict]) -> float:
        """10 log10 of the ratio of the summed P_x and P_y (`losses.py:445-468`)."""
        px = float(np.sum([m.get("P_x", 1.0) for m in modes]))
        py = float(np.sum([m.get("P_y", 1.0) for m in modes]))
        tiny = 1e-30
        if px < tiny and py < tiny:
            return 0.1
        return _clip(10.0 * np.log10(max(px, py) / (min(px, py) + tiny)), 0.0, 50.0)

    @staticmethod
    def _calculate_crosstalk_vectorial(modes: Sequence[Dict]) -> float:
        """Crosstalk proxy from the spread and regularity of the n_eff ladder and the mean confinement
        (`losses.py:546-619`): -10 - 20 Q - 5 CV - 5 Gamma, clipped to [-40, -15] dB."""
        if len(modes) < 2:
            return -25.0
        ne = np.sort([float(m["n_eff"]) for m in modes])
        conf = np.array([m.get("confinement", 0.5) for m in modes])
        gaps = np.diff(ne)
        top, bottom = float(ne[-1]), float(ne[0])
        guide = max((top + 0.01) - (bottom - 0.002), 1e-6)          # estimated n_core - n_clad
        Q = _clip((top - bottom) / guide, 0.0, 1.0)
        if len(gaps) > 1:
            cv = _clip(float(np.std(gaps)) / (float(np.mean(gaps)) + 1e-12) / 2.0, 0.0, 1.0)
        else:
            cv = 0.5
        guided = conf > 0.01
        gamma = float(np.mean(conf[guided])) if np.any(guided) else 0.5
        return _clip(-10.0 - 20.0 * Q - 5.0 * cv - 5.0 * gamma, -40.0, -15.0)

    @staticmethod
    def _calculate_crosstalk_scalar(modes: Sequence[Dict]) -> float:
        """Largest normalised field overlap between scalar modes, with the reference's penalty for n_eff gaps
        below 1e-4 (`losses.py:622-663`)."""
        best = 0.0
        fields = [(m.get("field_vector"), m) for m in modes]
        for i, (ei, _) in enumerate(fields):
            if ei is None:
                continue
            pi = float(np.real(np.vdot(ei, ei)))
            if pi < 1e-12:
                continue
            for ej, _ in fields[i + 1:]:
                if ej is None:
                    continue
                pj = float(np.real(np.vdot(ej, ej)))
                if pj < 1e-12:
                    continue
                best = max(best, float(np.abs(np.vdot(ei, ej)) ** 2 / (pi * pj + 1e-16)))
        if len(modes) < 2 or best == 0.0:
            return -70.0
        xt = -10.0 * np.log10(best + 1e-15)
        ne = np.sort([float(m["n_eff"]) for m in modes])
        gap = float(np.min(np.diff(ne)))
        if gap < 1e-4:
            xt -= 15.0 + (1e-4 - gap) * 1e6
        return _clip(xt, -70.0, -15.0)

    @staticmethod
    def _calculate_crosstalk(modes: Sequence[Dict]) -> float:
        """Route on ``is_vectorial`` (`losses.py:666-686`)."""
        if not modes:
            return -70.0
        if modes[0].get("is_vectorial", False):
            return EnhancedLossCalculator._calculate_crosstalk_vectorial(modes)
        return EnhancedLossCalculator._calculate_crosstalk_scalar(modes)

    @staticmethod
    def _calculate_radiation_loss(modes: Sequence[Dict], wavelength_nm: float) -> float:
        """Mean radiation penalty in dB/m (`losses.py:693-716`): Im(beta) when beta is complex, else from 1 - confinement."""
        scale = 1550.0 / wavelength_nm
        out = []
        for m in modes:
            conf, beta = m["confinement"], m["beta"]
            if np.iscomplexobj(beta) and abs(beta.imag) > 1e-9:
                out.append(2.0 * abs(beta.imag) * 1e6 * 8.685889638 * scale)
            else:
                p = max(0.0, 1.0 - conf) * 100.0
                if conf < 0.95:
                    p += (0.95 - conf) * 250.0
                out.append(p)
        return float(np.mean(out)) if out else 0.0


class VectorialLossCalculator:
    """IL / MDL / PDL per section from vectorial mode records (`losses.py:996-1221`)."""

    @staticmethod
    def _polymer_vectorial(modes_v, design_params, wavelength_nm: float) -> Dict:
        il = 0.2 * (design_params.d_polymer * 1e-6)                     # IP-Dip, 0.2 dB/m over the polymer thickness
        conf = [m["confinement"] for m in modes_v]
        mdl = 10.0 * np.log10(max(conf) / (min(conf) + 1e-12)) if len(conf) > 1 else 0.0
        px = float(np.sum([m.get("P_x", 1.0) for m in modes_v]))
        py = float(np.sum([m.get("P_y", 1.0) for m in modes_v]))
        tiny = 1e-30
        pdl = 10.0 * np.log10(max(px, py) / (min(px, py) + tiny)) if (px > tiny and py > tiny) else 0.1
        return {"IL": _clip(il, 0.0, 1.0), "MDL": _clip(mdl, 0.0, 2.0), "PDL": _clip(pdl, 0.05, 1.0), "PDL_x": px, "PDL_y": py}

    @staticmethod
    def _taper_vectorial(modes_v, design_params, wavelength_nm: float) -> Dict:
        L, n_t = design_params.L_taper, design_params.n_taper
        eta = 1.0 - np.exp(-L / (150.0 * max(n_t, 0.5)))                # adiabaticity against a 150 um beat length
        conf = np.array([m["confinement"] for m in modes_v])
        il = (-10.0 * np.log10(max(eta, 1e-6)) + 0.5 * (L * 1e-6)
              + max(0.0, 1.0 - float(np.mean(conf))) * 0.5 + 0.05 * np.log10(len(modes_v) + 1))
        px = [m.get("P_x", 1.0) for m in modes_v]
        py = [m.get("P_y", 1.0) for m in modes_v]
        mdl = 10.0 * np.log10(1.0 + (np.var(px) + np.var(py)) / 2.0) if len(px) > 1 else 0.0
        pdl_each = [m.get("PDL_dB", 0.0) for m in modes_v]
        power = [a + b for a, b in zip(px, py)]
        pdl = float(np.average(pdl_each, weights=power)) if sum(power) > 1e-12 else float(np.mean(pdl_each))
        pdl += 4.343 * (2.0 * np.pi / (wavelength_nm * 1e-3)) * 1e-5 * L   # taper birefringence 1e-5
        return {"IL": _clip(il, 0.0, 10.0), "MDL": _clip(mdl, 0.0, 5.0), "PDL": _clip(pdl, 0.01, 3.0),
                "PDL_x": float(np.sum(px)), "PDL_y": float(np.sum(py))}

    @staticmethod
    def _mmf_vectorial(modes_v, design_params) -> Dict:
        return {"IL": 0.32, "MDL": 0.05, "PDL": 0.05,
                "PDL_x": float(np.mean([m.get("P_x", 1.0) for m in modes_v])),
                "PDL_y": float(np.mean([m.get("P_y", 1.0) for m in modes_v]))}

    @staticmethod
    def calculate_vectorial_losses(modes_vectorial: List[Dict], geometry, design_params, direction: str = "mux",
                                   wavelength_nm: float = 1550.0) -> Dict:
        if not modes_vectorial:
            return {"success": False, "error": "no modes"}
        if not modes_vectorial[0].get("is_vectorial", False):
            logger.warning("non-vectorial modes handed to VectorialLossCalculator")
            return {"success": False, "error": "modes not vectorial"}
        try:
            sec = {"polymer": VectorialLossCalculator._polymer_vectorial(modes_vectorial, design_params, wavelength_nm),
                   "taper": VectorialLossCalculator._taper_vectorial(modes_vectorial, design_params, wavelength_nm),
                   "MMF": VectorialLossCalculator._mmf_vectorial(modes_vectorial, design_params)}
        except Exception as e:                              # noqa: BLE001 — the reference reports, it does not raise
            logger.error("VectorialLossCalculator: %s", e)
            return {"success": False, "error": str(e)}
        out: Dict = {"success": True, "is_vectorial": True}
        for name, s in sec.items():
            out[f"IL_{name}"], out[f"MDL_{name}"], out[f"PDL_{name}"] = s["IL"], s["MDL"], s["PDL"]
            out[f"PDL_x_{name}"], out[f"PDL_y_{name}"] = s["PDL_x"], s["PDL_y"]
        out["IL_total"] = _clip(sum(s["IL"] for s in sec.values()), 0.0, 40.0)
        out["MDL_total"] = _clip(math.sqrt(sum(s["MDL"] ** 2 for s in sec.values())), 0.0, 10.0)
        out["PDL_total"] = _clip(sum(s["PDL"] for s in sec.values()), 0.05, 10.0)
        out.update(n_modes_used=len(modes_vectorial), direction=direction, wavelength_nm=float(wavelength_nm))
        return out


class LossCalculator(EnhancedLossCalculator):
    """`calculate_physical_losses(modes, geometry, direction, wavelength_nm)` -> the dict of `losses.py:813-825`."""

    @staticmethod
    def _build_design_params(modes: List[Dict], geometry, wavelength_nm: float) -> PhotonicLanternDesignParameters:
        """Design parameters from the geometry actually solved (`losses.py:871-989`)."""
        n_cores = int(getattr(geometry, "n_cores", 3))
        radii = getattr(geometry, "core_radii", None)
        r_core = _first(radii) if radii is not None else float(getattr(geometry, "r_core", 1.2))
        n_core = _first(getattr(geometry, "core_index", getattr(geometry, "n_core", 1.535)))
        n_clad = _first(getattr(geometry, "clad_index", getattr(geometry, "n_clad", 1.0)))
        k0 = _first(getattr(geometry, "k0", 2.0 * np.pi / (wavelength_nm / 1000.0)))
        contrast = max(n_core ** 2 - n_clad ** 2, 1e-6)
        V = getattr(geometry, "V_number", None)
        V = _first(V) if V is not None else float(k0 * r_core * np.sqrt(contrast))
        Vc = max(V, 0.5)
        mfd = float(2.0 * r_core * (0.65 + 1.619 / Vc ** 1.5 + 2.879 / Vc ** 6))        # Marcuse
        pos = getattr(geometry, "positions", getattr(geometry, "core_positions", None))
        pos = list(pos) if pos is not None else None
        if pos and len(pos) >= 2:
            P = np.array(pos, dtype=float)
            d = [float(np.linalg.norm(P[i] - P[j])) for i in range(len(P)) for j in range(i + 1, len(P))]
            pitch = float(np.min(d)) if d else 8.0
            R_ring = float(np.max(np.linalg.norm(P, axis=1)))
        else:
            pitch = R_ring = 8.0
        central = bool(pos) and bool(np.any(np.linalg.norm(np.array(pos, dtype=float), axis=1) < 0.5 * r_core))
        kind = "hexagonal" if n_cores in (7, 19) else "circular"
        taper = getattr(geometry, "taper_length", None)
        taper = _first(taper) if taper is not None else 0.0
        L_taper, L_mux = (taper, max(taper * 0.5, 100.0)) if taper > 0.0 else (375.0, 200.0)
        L_mmf = 100.0
        return PhotonicLanternDesignParameters(
            N_cores=n_cores, has_central_core=central, config_type=kind, geometry_config=f"{n_cores}-{kind}",
            n_peripheral_cores=n_cores - (1 if central else 0), R_ring=R_ring,
            packing_efficiency=_clip(n_cores * np.pi * r_core ** 2 / (np.pi * max(R_ring + r_core, 1.0) ** 2), 0.01, 0.90),
            pitch=pitch, pitch_min=pitch, pitch_ratio=float(pitch / (2.0 * r_core + 1e-9)), wavelength=float(wavelength_nm),
            r_core_SM=r_core, r_clad_SM=62.5, n_core_SM=float(n_core), n_clad_SM=float(n_clad), V_SM=float(V),
            NA_SM=float(np.sqrt(contrast)), MFD=mfd, n_eff_LP01=float(modes[0]["n_eff"]) if modes else float(n_core - 0.01),
            r_core_MM=25.0, V_MM=float(np.sqrt(n_cores) * V), NA_MM=0.22, M_max=max(int(n_cores * V ** 2 / 4), 1),
            n_polymer=float(n_core), d_polymer=2.0, coupling_uniformity=0.95, L_mux=L_mux, L_taper=L_taper, L_MMF=L_mmf,
            L_total=L_mux + L_taper + L_mmf, n_taper=1.0, taper_profile="exponential")

    @staticmethod
    def calculate_physical_losses(modes: List[Dict], geometry, direction: str = "mux", wavelength_nm: float = 1550.0) -> Dict:
        if not (modes and modes[0].get("is_vectorial", False)):
            return {"success": False, "error": "scalar-mode losses (losses.py:74-440) are outside this build: vectorial modes only"}
        params = LossCalculator._build_design_params(modes, geometry, wavelength_nm)
        res = VectorialLossCalculator.calculate_vectorial_losses(modes, geometry, params, direction, wavelength_nm)
        if not res.get("success", False):
            return {"success": False, "error": res.get("error", "unknown")}
        pdl = res["PDL_total"]
        if direction == "demux":
            # demultiplexing excites the weakly confined high-order modes first: PDL grows by 2-12 % (`losses.py:773-801`)
            each = np.array([m.get("PDL_dB", 0.0) for m in modes])
            if len(each) >= 4:
                srt = np.sort(each)
                spread = max(float(np.mean(srt[-4:])) - float(np.mean(srt[:4])), 0.0)
            else:
                spread = 0.3
            conf = np.array([m.get("confinement", 0.5) for m in modes])
            cv = float(np.std(conf) / (np.mean(conf) + 1e-9))
            pdl = pdl * (1.0 + _clip(0.04 + 0.06 * cv + 0.02 * spread, 0.02, 0.12))
        conf_all = [m.get("confinement", 0.0) for m in modes]
        return {"IL_dB": res["IL_total"], "MDL_dB": res["MDL_total"], "PDL_dB": _clip(pdl, 0.05, 10.0),
                "crosstalk_dB": EnhancedLossCalculator._calculate_crosstalk_vectorial(modes),
                "radiation_loss_dB_per_m": EnhancedLossCalculator._calculate_radiation_loss(modes, wavelength_nm),
                "avg_confinement": float(np.mean(conf_all)) if conf_all else 0.0, "n_modes_used": res["n_modes_used"],
                "direction": direction, "wavelength_nm": float(wavelength_nm), "is_vectorial": True, "success": True}
