"""bench.py contract checks that need no GPU: the reference arm (`--impl reference` = the CPU oracle on a bounded
sample) prints one JSON line with the contract's keys, uses ALL host cores even when the launcher exports
OMP_NUM_THREADS=1 (torch.distributed.run does), and carries the same `config` dict, key for key, as the product arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_reference(extra_env):
    env = dict(os.environ, **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ne", "6",
                          "--cpu-sample-ne", "4", "--steps", "2", "--warmup", "1"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line_and_threads():
    line = _run_reference({"OMP_NUM_THREADS": "1"})
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "GDOF/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"]
    ncpu = len(os.sched_getaffinity(0))
    assert cb["cores"] == ncpu, f"the oracle must use all {ncpu} host cores, not OMP_NUM_THREADS=1 (got {cb['cores']})"


def test_both_arms_share_the_config_dict():
    sys.path.insert(0, ROOT)
    import bench
    class A:
        gpus, lx, ne = 4, 8, 64
    cfg = bench.config_dict(A)
    assert set(cfg) == {"workload", "lx", "elements_per_gpu", "dof_per_gpu", "rank_grid", "l2"}
    assert cfg["elements_per_gpu"] == 64 ** 3 and cfg["dof_per_gpu"] == 8 ** 3 * 64 ** 3 and cfg["rank_grid"] == [2, 2, 1]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_dict(args)') == 2, "both arms must build `config` with config_dict(args)"
