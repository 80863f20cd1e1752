"""bench.py host logic (no GPU): byte model, level selection and the `--impl reference` arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_byte_model_reproduces_baseline_worked_examples():
    # 4M fine nodes, 6 levels at 4x coarsening, coarse n=25: 52.0 GB (p=6), 30.3 GB (p=4)
    sides = [63, 125, 250, 500, 1000, 2000]
    assert abs(bench.algorithmic_bytes_per_cycle(sides, 6) / 1e9 - 52.0) < 1.0
    assert abs(bench.algorithmic_bytes_per_cycle(sides, 4) / 1e9 - 30.3) < 1.0


def test_level_sides():
    assert bench.level_sides(2000) == [16, 32, 63, 125, 250, 500, 1000, 2000]
    assert bench.level_sides(1000, 6) == [32, 63, 125, 250, 500, 1000]
    assert bench.level_sides(100, 4) == [13, 25, 50, 100]


def _run(rank):
    env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                           "--side", "100"], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_line_on_rank0_only():
    r0, r1 = _run(0), _run(1)
    assert r0.returncode == 0 and r1.returncode == 0
    assert r1.stdout.strip() == ""                                     # other ranks exit 0 without work
    line = json.loads(r0.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "vcycles_per_s" and line["unit"] == "V-cycles/s"
    import oracle
    # the timed cycles run the reference's own objects whenever oracle/_ref/libref.so exists (it travels to the GPU box), else the port
    assert line["cpu_baseline"]["kind"] == ("reference" if oracle.ReferenceHierarchy.available() else "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["scaling"] == "strong"
    # the arm MEASURES the stated workload: the reference's stop rule to 1e-8 with its own smoother, timed steps inside that solve
    assert line["solve"]["converged"] and line["solve"]["final_residual"] < 1e-8 and line["solve"]["cycles"] >= 1
    assert "100x100" in line["config"]["workload"] and "lexicographic" in line["config"]["smoother"]
    assert abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6


def test_golden_histories_of_the_bench_workload_are_committed():
    # the in-run check of bench.py: the first cycles of the 2000^2 workload from the CPU oracle alone, polyDeg 4 and 6
    for poly in (4, 6):
        for key in ("lexicographic_omega1.4", "multicolour_omega0.8"):
            h = bench.golden_history(2000, poly, key)
            assert h is not None and len(h) >= 6 and h[0] == 1.0
            assert all(b < a for a, b in zip(h, h[1:]))               # both smoothers contract on this workload
    assert bench.golden_history(2000, 5, "multicolour_omega0.8") is None   # no fixture: bench reports check = null, it does not guess


def test_ncu_traffic_lookup_is_keyed_on_kernel_and_workload():
    t = bench.ncu_traffic("k_sor_mc_tma<8,5,2>", 2000, 4)
    assert t is not None and 0.9 < t / 9.44e9 < 1.1                  # measured DRAM bytes per launch against the algorithmic 9.44 GB
    assert bench.ncu_traffic("k_sor_mc_tma<8,5,2>", 1000, 4) is None and bench.ncu_traffic("no_such_kernel", 2000, 4) is None
