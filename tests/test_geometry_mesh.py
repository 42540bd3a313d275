"""Host boundary vs values frozen from the real reference modules (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import plfem_b200 as P
from plfem_b200.mesh import MeshGenerator, MeshTri, signed_double_area
from plfem_b200.sweep import design_geometry, lhs_designs

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(G, "geometry.json")))


def test_layouts_bit_identical_to_reference(gold):
    xs = np.array(gold["eps_sample_points"])
    for key, ref in gold.items():
        if ":" not in key:
            continue
        n, variant = key.split(":")
        g = P.MCFGeometry(int(n), 8.0, 1.5, 1.535, 1.0, variant=None if variant == "None" else variant)
        assert np.array_equal(g.positions, np.array(ref["positions"])), key
        assert g.config_type == ref["config_type"]
        assert g.domain_radius == ref["domain_radius"] and g.cladding_radius == ref["cladding_radius"]
        assert g.hash == ref["hash"] and g.k0 == ref["k0"] and float(g.V_number) == ref["V_number"]
        eps = g.epsilon(xs[0], xs[1])
        assert digest(np.real(eps)) == ref["eps_real_digest"] and digest(np.imag(eps)) == ref["eps_imag_digest"]


def test_reference_selfcheck_values(gold):
    # geometry_unified.py:736-772: V = 5.63 at (r=1.2, n=1.53, 1.55 um), eps(0,0) = n_core^2
    g = P.MCFGeometry(7, 8.0, 1.2, 1.53, 1.0)
    assert abs(g.V_number - 5.63) < 5e-3 and float(g.V_number) == gold["selfcheck"]["V"]
    assert np.real(g.epsilon(np.array([0.0]), np.array([0.0])))[0] == 1.53 ** 2 == gold["selfcheck"]["eps00"]
    assert np.real(g.epsilon(np.array([100.0]), np.array([0.0])))[0] == 1.0
    ok, _ = g.validate()
    assert ok and len(g.hash) == 20


def test_both_constructor_surfaces_agree():
    a = P.PhotonicLanternGeometry(arrangement="hexagonal_1plus6_7", core_radius_um=1.5, pitch_um=8.0,
                                  n_core=1.535, n_clad=1.0, wavelength_nm=1550)
    pos, *_ = P.mcf_positions(7, 8.0)
    b = P.PhotonicLanternGeometry(7, "hexagonal_1plus6_7", pos, np.full(7, 1.5), 1.535, 1.0, wavelength=1.55)
    assert np.array_equal(a.positions, b.positions) and a.k0 == b.k0 and a.n_cores == b.n_cores == 7
    assert a.domain_radius == pytest.approx(b.domain_radius)
    with pytest.raises(ValueError):
        P.MCFGeometry(10, 8.0, 1.5, 1.535)
    with pytest.raises(ValueError):
        P.MCFGeometry(7, 8.0, 1.5, 1.0, 1.0)          # delta_n too small


def test_mesh_recipe_matches_reference_mesh_py():
    gold = json.load(open(os.path.join(G, "mesh_recipe.json")))
    for key, ref in gold.items():
        n, refinement = key.split(":")
        n = int(n)
        g = P.MCFGeometry(n, 8.0 if n != 3 else 6.0, 1.5 if n != 3 else 1.2, 1.535 if n != 3 else 1.53, 1.0)
        raw = MeshGenerator._delaunay_mesh(g, float(refinement))
        assert (raw.p.shape[1], raw.t.shape[1]) == (ref["V"], ref["T"])
        assert digest(raw.p) == ref["p_digest"] and digest(raw.t.astype(np.int64)) == ref["t_digest"]


def test_flat_triangles_are_dropped_and_nothing_else(cfg1):
    g, mesh = cfg1
    raw = MeshGenerator._delaunay_mesh(g, 1.0)
    det = signed_double_area(raw.p, raw.t)
    assert (det == 0).sum() == 25                      # hull-collinear grid points, see mesh.drop_flat_triangles
    assert mesh.t.shape[1] == raw.t.shape[1] - 25
    assert (signed_double_area(mesh.p, mesh.t) != 0).all()
    assert np.array_equal(mesh.p, raw.p)
    assert (np.diff(mesh.t, axis=0) > 0).all()         # columns sorted like scikit-fem's MeshTri


def test_mesh_cache_and_refine():
    MeshGenerator.clear_cache()
    g = P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0)
    m1, _ = MeshGenerator.generate(g, 0.4)
    m2, _ = MeshGenerator.generate(g, 0.4)
    assert m1 is m2 and MeshGenerator.get_cache_stats()["hits"] == 1
    # the mesh does not depend on wavelength or index: a band sweep shares it
    g2 = P.MCFGeometry(3, 6.0, 1.2, 1.529, 1.0, 1.49)
    assert MeshGenerator.generate(g2, 0.4)[0] is m1
    r = m1.refined()
    assert r.t.shape[1] == 4 * m1.t.shape[1]
    assert np.isclose(np.abs(signed_double_area(r.p, r.t)).sum(), np.abs(signed_double_area(m1.p, m1.t)).sum())
    s = MeshTri.init_structured(4, 3, 2.0)
    assert s.p.shape[1] == 20 and s.t.shape[1] == 24


def test_cauchy_indices():
    # SURVEY.md 8(d) config 3
    for lam, n in ((1490, 1.529816), (1550, 1.529516), (1600, 1.529291), (1650, 1.529087)):
        assert abs(P.IPDipCauchy.n(lam) - n) < 1e-6


def test_hub_vertices_fit_the_device_pattern_builder(cfg1, cfg2):
    """The recipe makes every core centre the hub of a whole ring of points (34 triangles around one vertex in config 1).
    The device-side pattern builder (`pattern_rows_kernel`, assembly.cu) holds the 6 x valence candidate columns of a row
    in shared memory, PATTERN_MAXCAND = 768, i.e. valence <= 128: the recipe's meshes must stay well inside that."""
    import re
    src = open(os.path.join(ROOT, "pl-fem-vectoriel_b200", "csrc", "assembly.cu")).read()
    cap = int(re.search(r"PATTERN_MAXCAND\s*=\s*(\d+)", src).group(1)) // 6
    sample = [cfg1[1], cfg2[1]]
    for d in lhs_designs(24, seed=7)[:8]:
        sample.append(P.MeshGenerator.generate(design_geometry(d))[0])
    worst = max(int(np.bincount(np.asarray(m.t).ravel()).max()) for m in sample)
    assert 16 <= worst <= cap // 2, worst


def test_mesh_cache_round_trip(tmp_path):
    """`save_cache` / `load_cache` (`mesh.py:385-416`): same keys, same meshes, counters kept; a hit after reloading."""
    import plfem_b200 as P
    P.MeshGenerator.clear_cache()
    g = P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55)
    cfg = P.SimulationConfig()
    cfg.enable_mesh_cache = True
    m1, _ = P.MeshGenerator.generate(g, 0.4, cfg)
    path = tmp_path / "mesh_cache.npz"
    P.MeshGenerator.save_cache(path)
    before = P.MeshGenerator.get_cache_stats()
    P.MeshGenerator.clear_cache()
    P.MeshGenerator.load_cache(tmp_path / "missing.npz")          # no file: nothing happens
    assert P.MeshGenerator.get_cache_stats()["size"] == 0
    P.MeshGenerator.load_cache(path)
    assert P.MeshGenerator.get_cache_stats() == before
    m2, _ = P.MeshGenerator.generate(g, 0.4, cfg)
    assert P.MeshGenerator.get_cache_stats()["hits"] == before["hits"] + 1
    assert np.array_equal(m1.p, m2.p) and np.array_equal(m1.t, m2.t)
    P.MeshGenerator.clear_cache()


def test_recipe_meshes_stay_below_the_device_valence_cap():
    """The device pattern builder serves nodes that belong to at most 128 elements (README.md, known limits).  The recipe's
    hub is a core centre: it meets one triangle per direction of the polar rings, 16 * refinement of them (`mesh.py:232-297`) —
    34 elements at refinement 1, 40 at 2; the cap is reached beyond refinement 8 (105 there, a 300 k-vertex mesh)."""
    import plfem_b200 as P
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
    for ref, bound in ((1.0, 48), (2.0, 64)):
        mesh, _ = P.MeshGenerator.generate(g, refinement=ref)
        valence = np.bincount(np.asarray(mesh.t).ravel(), minlength=mesh.p.shape[1])
        assert valence.max() <= bound < 128
