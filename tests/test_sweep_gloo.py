"""Host logic of the multi-GPU sweep on CPU: sharding, the 86-slot record, the single gather
(world_size 2, gloo).  The per-design solver is replaced by the oracle here — test infrastructure only;
on GPUs `run_sweep` calls the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from plfem_b200 import sweep


def test_record_layout_and_designs():
    assert sweep.N_RECORD == 86 and len(set(sweep.RECORD_FIELDS)) == 86
    bands = sweep.band_sweep_designs()
    assert [d["wavelength_nm"] for d in bands] == [1490, 1550, 1600, 1650]
    assert abs(bands[1]["n_core"] - 1.529516) < 1e-6
    a, b = sweep.lhs_designs(60), sweep.lhs_designs(60)
    assert a == b and len(a) == 60                                  # reproducible (no salted hash seeds)
    assert {d["n_cores"] for d in a} <= set(sweep.SAMPLING_WEIGHTS)
    assert all(0.5 <= d["core_radius_um"] <= 3.0 and 3.0 <= d["pitch_um"] <= 15.0 for d in a)
    assert all(sweep.design_geometry(d).validate()[0] for d in a)
    n7 = sum(d["n_cores"] == 7 for d in a)
    assert n7 == max(sum(d["n_cores"] == n for d in a) for n in sweep.SAMPLING_WEIGHTS)   # 7-core weight is the largest
    assert sweep.shard(10, 1, 4) == [1, 5, 9] and sorted(sum((sweep.shard(10, r, 4) for r in range(4)), [])) == list(range(10))


def _oracle_worker(d):
    import time
    from oracle import fem_oracle as O
    from plfem_b200.mesh import MeshGenerator
    if d.get("fail"):
        raise RuntimeError("injected failure")
    g = sweep.design_geometry(d)
    mesh, _ = MeshGenerator.generate(g, 0.4)
    t = time.perf_counter()
    modes = O.solve_vectorial_modes(g, mesh, d["n_modes"])
    return g, mesh, modes, dict(sigma=O.sigma_estimate(g), n_dofs=2 * len(modes[0]["Ex_dofs"])), time.perf_counter() - t


def _designs():
    ds = [dict(n_cores=n, core_radius_um=1.2, pitch_um=6.0, wavelength_nm=lam, n_core=1.53, n_clad=1.0, n_modes=2, variant=None)
          for n, lam in ((1, 1550), (2, 1550), (3, 1490), (3, 1650), (2, 1600))]
    ds[3]["fail"] = True
    return ds


def _oracle_forest(ds):
    """Forest worker stand-in: one result or Exception per design (what `solve_forest_gpu` returns)."""
    out = []
    for d in ds:
        try:
            out.append(_oracle_worker(d))
        except Exception as e:                                      # noqa: BLE001
            out.append(e)
    return out


def test_forest_mode_gives_the_same_records():
    single = sweep.run_sweep(_designs(), 0, 1, solve_fn=_oracle_worker)
    f = {k: i for i, k in enumerate(sweep.RECORD_FIELDS)}
    keep = [i for i in range(sweep.N_RECORD) if i != f["solver_time_s"]]
    for B in (1, 2, 5, 8):
        rec = sweep.run_sweep(_designs(), 0, 1, forest=B, forest_fn=_oracle_forest)
        assert np.array_equal(rec[:, keep], single[:, keep], equal_nan=True), B
    calls = []
    rec = sweep.run_sweep(_designs(), 1, 2, forest=2, forest_fn=lambda ds: calls.append(len(ds)) or _oracle_forest(ds), gather=False)
    assert calls == [2] and list(rec[:, f["success"]][[1, 3]]) == [1, 0]     # rank 1 of 2 owns designs 1 and 3: one forest

    def broken(ds):
        raise RuntimeError("whole forest failed")
    rec = sweep.run_sweep(_designs(), 0, 1, forest=3, forest_fn=broken)
    assert list(rec[:, f["success"]]) == [0, 0, 0, 0, 0]


def _rank_main(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec = sweep.run_sweep(_designs(), rank, world, solve_fn=_oracle_worker)
        np.save(os.path.join(out_dir, f"rec{rank}.npy"), rec)
    finally:
        dist.destroy_process_group()


def test_two_rank_sweep_matches_single_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rec0.npy"), np.load(tmp_path / "rec1.npy")
    single = sweep.run_sweep(_designs(), 0, 1, solve_fn=_oracle_worker)
    f = {k: i for i, k in enumerate(sweep.RECORD_FIELDS)}
    t = f["solver_time_s"]
    keep = [i for i in range(sweep.N_RECORD) if i != t]
    assert r0.shape == (5, 86)
    assert np.array_equal(r0[:, keep], r1[:, keep], equal_nan=True)          # every rank holds all records
    assert np.array_equal(r0[:, keep], single[:, keep], equal_nan=True)      # and they equal the 1-rank sweep
    assert list(r0[:, f["sample_id"]]) == [0, 1, 2, 3, 4]
    assert list(r0[:, f["success"]]) == [1, 1, 1, 0, 1]                      # the failed design poisons nothing
    assert np.isnan(r0[3, f["n_eff_mean"]]) and r0[2, f["n_modes_found"]] >= 1
    assert r0[2, f["wavelength_nm"]] == 1490 and r0[0, f["n_cores"]] == 1
    ok = r0[:, f["n_modes_found"]] >= 1                                      # loss columns: filled wherever modes were found
    assert np.isfinite(r0[ok][:, [f["loss_IL_mux_dB"], f["loss_PDL_demux_dB"], f["loss_XT_mux_dB"], f["loss_taper_dB"]]]).all()
    assert (r0[ok, f["loss_PDL_demux_dB"]] >= r0[ok, f["loss_PDL_mux_dB"]]).all() and np.isnan(r0[3, f["loss_IL_mux_dB"]])
    p = tmp_path / "records.csv"
    sweep.records_to_csv(r0, str(p))
    lines = open(p).read().splitlines()
    assert len(lines) == 6 and lines[0].split(",")[:4] == sweep.RECORD_FIELDS[:4]
