"""Host-side logic of the depth-sharded multi-GPU mode (stardis_b200/distributed.py): depth views of the plasma / model and
the all-to-all between the depth and the nu decomposition, on CPU with gloo and world_size 2."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_depth_views_slice_every_per_depth_table():
    from stardis_b200 import units as u
    from stardis_b200.distributed import DepthSlicedModel, DepthSlicedPlasma, depth_indices
    from stardis_b200.synthetic import make_workload

    w = make_workload("sim10aa", seed=3, n_lines=50, device_strengths=True)
    model, plasma = w["model"], w["plasma"]
    D = model.no_of_depth_points
    idx = depth_indices(D, 1, 4)
    assert list(idx[:3]) == [1, 6, 9] and sorted(np.concatenate([depth_indices(D, r, 4) for r in range(4)])) == list(range(D))
    m, p = DepthSlicedModel(model, idx), DepthSlicedPlasma.of(plasma, idx)
    assert m.no_of_depth_points == len(idx) and p is DepthSlicedPlasma.of(plasma, idx)
    np.testing.assert_array_equal(u.values_of(m.temperatures), u.values_of(model.temperatures)[idx])
    np.testing.assert_array_equal(p.electron_densities.values, plasma.electron_densities.values[idx])
    np.testing.assert_array_equal(p.ion_number_density.loc[1, 0].values, plasma.ion_number_density.loc[1, 0].values[idx])
    np.testing.assert_array_equal(p.level_number_density.values, plasma.level_number_density.values[:, idx])
    lt = p.line_table
    np.testing.assert_array_equal(lt.alpha_line, plasma.line_table.alpha_line[:, idx])
    assert lt.strength.tables["n_over_u"].shape[1] == len(idx) and lt.strength.per_line["gf"] is plasma.line_table.strength.per_line["gf"]
    assert p.ionization_data is plasma.ionization_data and m.composition is model.composition
    # the sliced line-strength inputs reproduce the sliced table
    np.testing.assert_allclose(lt.strength.host_alpha(u.values_of(m.temperatures), lt.nu, lt.level_energy_lower), lt.alpha_line, rtol=1e-15)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["SD_ROOT"])
    from stardis_b200.distributed import (all_shards, allgather_depth_columns, depth_indices, exchange_depth_to_nu,
                                          upload_columns_striped)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    D, N, Ln = 7, 101, 13
    full = np.arange(D * N, dtype=np.float64).reshape(D, N) + 0.25
    idx = depth_indices(D, rank, world)
    out = exchange_depth_to_nu(torch.from_numpy(full[idx].copy()), D, N)
    a, b = all_shards(N, world)[rank]
    np.testing.assert_array_equal(out.numpy(), full[:, a:b])
    bounds = [(0, 30), (30, N)]
    out = exchange_depth_to_nu(torch.from_numpy(full[idx].copy()), D, N, bounds=bounds)
    np.testing.assert_array_equal(out.numpy(), full[:, bounds[rank][0]:bounds[rank][1]])
    table = np.arange(Ln * D, dtype=np.float64).reshape(Ln, D)
    cols = allgather_depth_columns(torch.from_numpy(np.ascontiguousarray(table[:, idx])), D)
    np.testing.assert_array_equal(cols.numpy(), table)
    cols = dict(nu=np.linspace(1.0, 2.0, 37), Z=np.arange(37, dtype=np.int64) % 5, A=np.geomspace(1e6, 1e9, 37))
    got = upload_columns_striped(cols, torch.device("cpu"))
    for k, v in cols.items():
        assert got[k].dtype == (torch.float64 if v.dtype.kind == "f" else torch.int64)
        np.testing.assert_array_equal(got[k].numpy(), v)
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_depth_to_nu_exchange_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SD_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2
