"""Sweep construction (restating main.py:118-142, 504-534) and the N>1 host path: cases dealt
one per rank, timing reduced as the max over ranks - exercised with world_size 2 on gloo."""
import os
import subprocess
import sys
import textwrap

from openfoam_tpp_b200 import ensemble as en

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parse_range_matlab_semantics():
    assert en.parse_range("1.6:0.2:3.0") == [1.6, 1.8, 2.0, 2.2, 2.4, 2.6, 2.8, 3.0]  # inclusive end
    assert en.parse_range("0.002:0.001:0.005") == [0.002, 0.003, 0.004, 0.005]
    assert en.parse_range("1:3") == [1.0, 2.0, 3.0]
    assert en.parse_range("0.004, 0.008") == [0.004, 0.008]


def test_sweep_zip_vs_product_and_names():
    base = {"H": 0.1, "D": 0.02, "geo": "flat", "R": 0.003, "freq": 2.0, "duration": 10.0, "mesh": 0.002}
    z = en.build_param_sets(base, {"R": [0.002, 0.003], "freq": [1.0, 2.0]})
    assert [(p["R"], p["freq"]) for p in z] == [(0.002, 1.0), (0.003, 2.0)]  # equal lengths: zip
    sweeps = {"H": en.parse_range("0.004,0.008"), "R": en.parse_range("0.002:0.001:0.005"), "freq": en.parse_range("1.6:0.2:3.0")}
    prod = en.build_param_sets(base, sweeps)
    assert len(prod) == 64  # BASELINE.json config 5
    assert en.case_name(prod[0]) == "case_H0.004_D0.02_flat_R0.002_f1.6_d10.0_m0.002"
    assert len({en.case_name(p) for p in prod}) == 64


def test_shard_covers_every_case_once():
    items = list(range(64))
    for world in (1, 2, 4, 8):
        got = sorted(x for r in range(world) for x in en.shard(items, world, r))
        assert got == items
        assert max(len(en.shard(items, world, r)) for r in range(world)) == 64 // world


def test_two_rank_ensemble_gloo(tmp_path, emu_lib):
    """world_size 2 on gloo: each rank steps its own case (host emulation of the kernels), the
    job time is the max over ranks, the cell count the sum."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, time
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        from openfoam_tpp_b200 import case as cs, ensemble as en, solver as sv
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        freqs = en.parse_range("2.0:0.5:2.5")
        f = en.shard(freqs, world, rank)[0]
        d = os.path.join({str(tmp_path)!r}, f"case{{rank}}")
        cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=f, duration=0.2, n_rings=4 + rank, n_layers=3)
        c = cs.Case(d)
        g = sv.Solver(c.mesh, c.cfg, lib_path={emu_lib!r})
        g.load_case_fields(c)
        t0 = time.perf_counter(); g.step(3); sec = time.perf_counter() - t0 + rank
        tmax = en.max_over_ranks([sec])[0]
        cells = en.sum_over_ranks([float(c.mesh.n_cells)])[0]
        assert tmax >= 1.0 and tmax >= sec
        print(f"RANK{{rank}} f={{f}} cells={{c.mesh.n_cells}} total={{int(cells)}} t={{g.info()['t']:.6e}}")
        dist.destroy_process_group()
    """))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout + r.stderr
    assert "RANK0 f=2.0" in out and "RANK1 f=2.5" in out
    n0 = 18 * 16 * 3
    n1 = 18 * 25 * 3
    assert f"total={n0 + n1}" in out


def test_run_sweep_sets_up_and_runs_every_case(tmp_path, emu_lib):
    """the sweep loop (main.py:599-608) on one rank: 2 x 2 cases, a few steps each, named as the
    reference names them, each leaving a restartable case directory"""
    base = {"H": 0.004, "D": 0.0221, "geo": "flat", "R": 0.005, "freq": 2.0, "duration": 0.02, "mesh": 0.003}
    done = en.run_sweep(str(tmp_path), base, {"R": [0.004, 0.005], "freq": [1.5, 2.0, 2.5][:2] + [3.0]}, max_steps=3, lib_path=emu_lib)
    names = [n for n, _ in done]
    assert len(done) == 6 and len(set(names)) == 6 and all(o["steps"] == 3 for _, o in done)
    assert "case_H0.004_D0.0221_flat_R0.004_f1.5_d0.02_m0.003" in names
    for n in names:
        assert os.path.isfile(os.path.join(str(tmp_path), n, "constant", "polyMesh", "owner"))
        assert os.path.isfile(os.path.join(str(tmp_path), n, "constant", "6DoF.dat"))


def test_run_sweep_concurrent_cases_match_sequential(tmp_path, emu_lib):
    """cases_per_gpu = 3: the same cases advanced by three host threads at once (one handle each) end
    in the same state as when they run one after another."""
    base = {"H": 0.004, "D": 0.0221, "geo": "flat", "R": 0.005, "freq": 2.0, "duration": 0.002, "mesh": 0.003}
    sw = {"freq": [1.5, 2.0, 2.5, 3.0]}
    a = en.run_sweep(str(tmp_path / "seq"), base, sw, max_steps=4, lib_path=emu_lib)
    b = en.run_sweep(str(tmp_path / "par"), base, sw, max_steps=4, lib_path=emu_lib, cases_per_gpu=3)
    assert [n for n, _ in a] == [n for n, _ in b] and len(b) == 4
    for (n, oa), (_, ob) in zip(a, b):
        assert oa["steps"] == ob["steps"] == 4 and oa["t"] == ob["t"], (n, oa, ob)


def test_parse_range_reproduces_the_reference_loop():
    """parse_range sums the step cumulatively and applies the 1e-9 end slack and the 6-decimal rounding
    exactly like the reference's while-loop (main.py:118-142), restated here as the checker."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    def loop(start, step, end):
        vals, v = [], start
        while v <= end + 1e-9:
            vals.append(round(v, 6))
            v += step
        return vals

    @settings(max_examples=300, deadline=None)
    @given(start=st.floats(-5, 5), step=st.floats(1e-3, 2.0), n=st.integers(0, 60), frac=st.floats(0, 0.999))
    def check(start, step, n, frac):
        end = start + (n + frac) * step
        text = f"{start!r}:{step!r}:{end!r}"
        assert en.parse_range(text) == loop(float(repr(start)), float(repr(step)), float(repr(end)))

    check()
    assert en.parse_range("3:1") == [] and en.parse_range("2:2") == [2.0]
    import pytest

    with pytest.raises(ValueError):
        en.parse_range("1:0:2")
    with pytest.raises(ValueError):
        en.parse_range("1:2:3:4")
