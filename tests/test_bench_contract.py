"""The reference arm of bench.py is CPU code (the C oracle port on the host cores): run it here and check the JSON line
the driver parses.  The GPU arm's line is checked on the GPU box (test_gpu_bench_line below)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), *args], capture_output=True, text=True,
                         cwd=REPO, timeout=1500)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    line = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert BASE_KEYS <= set(line) and line['impl'] == 'reference'
    assert line['metric'] == 'grid_points_per_sec' and line['unit'] == 'grid points/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['steps'] == 1 and line['dtype'] == 'f64' and line['data'] == 'synthetic'
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == dict(value=line['value'], unit=line['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in line['config'] and 'model' not in line['config']


@pytest.mark.gpu
def test_gpu_bench_line():
    line = run_bench('--steps', '3', '--warmup', '3', '--members', '2')
    assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline', 'parity', 'collision_steps', 'spectroscopy'} <= set(line)
    assert line['metric'] == 'grid_points_per_sec' and line['scaling'] == 'weak' and line['config']['members_per_gpu'] == 2
    assert line['gpu_launches'] > 0 and line['value'] > 1e6 and line['e2e']['value'] > 1e6
    assert line['e2e']['h2d_bytes_per_step'] > 0 and line['e2e']['d2h_bytes_per_step'] == 2 * 3 * 128 * 16 * 128 * 16
    roof = line['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(roof) and roof['peak'] > 0
    assert abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-9 and 0 < roof['frac'] < 1
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['value'] > 0
    assert line['parity_max_rel_err_vs_oracle'] < 1e-10 and line['parity']['points'] >= 12
    assert line['collision_steps']['parity_max_abs_err_vs_oracle'] < 1e-10
    c2_roof = line['collision_steps']['roofline']
    assert 0 < c2_roof['frac'] < 1
    assert not set(line['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
