"""bench.py --impl reference on the CPU tier: the restated reference CPU path (oracle) timed on host threads.  The line
must carry the contract's keys with the SAME `config` the GPU arm emits, and the process must never load the product
library (the driver records the .so files each arm maps)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("workload", ["fir", "decim"])
def test_reference_arm_line(workload):
    env = dict(os.environ, SGPU_TEST_DUMP_MAPS="1")
    code = (
        "import sys, runpy\n"
        f"sys.argv = ['bench.py', '--impl', 'reference', '--workload', '{workload}', '--steps', '1', '--warmup', '0', '--cpu-seconds', '0.2']\n"
        "try:\n"
        "    runpy.run_path('bench.py', run_name='__main__')\n"
        "finally:\n"
        "    print('MAPS', [l.split()[-1] for l in open('/proc/self/maps') if '.so' in l and ('solid' in l or 'sgpu' in l)])\n"
    )
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    # the GPU arm's config for the same workload (pure function of the workload table)
    sys.path.insert(0, str(ROOT))
    import bench
    assert line["config"] == bench.workload_config(workload, 1, 30)
    maps = [l for l in r.stdout.splitlines() if l.startswith("MAPS")][-1]
    assert "libsolid_gpu" not in maps and "libsgpu_peakbench" not in maps and "libsolid_oracle" in maps
