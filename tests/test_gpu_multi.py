"""Sharded (multi-GPU) step vs the single-GPU step; needs >= 2 B200s on the box (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_step_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tools", "mgpu_check.py"), "256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "OK" in out.stdout


SHARD_EXE = os.path.join(ROOT, "tests", "native", "_build", "sharded_step_test")


def build_shard_exe():
    os.makedirs(os.path.dirname(SHARD_EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "dune-mmesh_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "sharded_step_test.cc"), "-o", SHARD_EXE,
                           "-L", os.path.join(ROOT, "dune-mmesh_b200", "lib"), "-lmmesh_b200", "-lpthread",
                           "-Wl,-rpath," + os.path.join(ROOT, "dune-mmesh_b200", "lib")])


def test_cpp_two_rank_driver(tmp_path):
    """A C++ caller (no Python, no torch in the rank processes): makeShard + mmb_comm_init + mmb_adapt_step_sharded on two
    GPUs; the global scalars printed by both ranks equal a single-GPU step, the owned marks add up to the same checksum."""
    import numpy as np
    import torch
    build_shard_exe()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
    from mmesh_b200 import capi, multigpu, synth
    m = synth.config3(nx=192, ny=96)
    s = synth.shifts_uniform(m, 0.3 * m.h, 30)
    new_off, new_cell, old_xy, old_idx = synth.pairs_self_and_neighbours(m, m.xy - s)
    rng = np.random.default_rng(1)
    dofs = rng.uniform(size=(m.nc, 3))
    minH, maxH, factor = multigpu.global_indicator_params(m)
    for policy in (0, 1):
        f = tmp_path / ("shard_%d.bin" % policy)
        with open(f, "wb") as out:
            np.array([m.nv, m.nc, m.edges.shape[0], m.segs.shape[0], old_idx.shape[0], policy], np.int64).tofile(out)
            for a, t in ((m.xy, np.float64), (m.cells, np.int32), (m.perm, np.uint8), (m.edges, np.int32), (m.edge_cell, np.int32), (m.segs, np.int32),
                         (s, np.float64), (new_off, np.int64), (new_cell, np.int32), (old_xy, np.float64), (old_idx, np.int32), (dofs, np.float64),
                         (np.array([minH, maxH, factor]), np.float64)):
                np.ascontiguousarray(a, t).tofile(out)
        idf = str(tmp_path / ("nccl_id_%d" % policy))
        procs = [subprocess.Popen([SHARD_EXE, str(f), str(r), "2", idf], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
        outs = [p.communicate(timeout=300) for p in procs]
        assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
        lines = [[ln for ln in o[0].splitlines() if ln.startswith("rank ")][-1] for o in outs]     # (NCCL may print its version first)
        rows = [dict(zip(ln.split()[2::2], ln.split()[3::2])) for ln in lines]
        ref = capi.Context(2, 0)
        ref.upload_mesh(m.xy, m.cells, m.perm, m.edges, m.segs)
        p2 = ref.get_params(); p2.minH, p2.maxH, p2.factor, p2.drop_area, p2.move_policy = minH, maxH, factor, 1e-12, policy; ref.set_params(p2)
        ref.upload_pairs(new_off, new_cell, old_xy, old_idx)
        ref.project(3, dofs, download=False)
        r = ref.adapt_step(s, 3).asdict()
        marks = ref.download_fields(0)["marks"]
        ref.close()
        for row in rows:
            for k in ("n_invalid", "n_orient_nonpos", "n_illegal", "n_refine", "n_coarsen", "n_pieces", "moved"):
                assert int(row[k]) == r[k], (policy, k, row[k], r[k])
            assert float(row["max_dist"]) == r["max_dist"]
            assert abs(float(row["mass"]) - r["mass_new"]) <= 1e-13 * abs(r["mass_new"])
        want = int((marks.astype(np.int64) * ((np.arange(m.nc) % 1000) + 1)).sum())
        assert sum(int(row["marksum"]) for row in rows) == want
        if policy == 0:
            assert r["n_invalid"] > 0 and r["moved"] == 0
