"""bench.py's one-line JSON contract, checked on the lines the GPU runs of this round actually printed (profiles/)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
            'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'roofline', 'clocks']


def _line(name):
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        pytest.skip(f'{name} not recorded yet')
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize('name', ['r01_bench_fp16x2_final.json', 'r01_bench_fp16x2_2gpu.json', 'r02_final_bench.json', 'r02_final_bench_first.json',
                                  'r02o_bench_2gpu.json', 'r02z_bench_4gpu.json'])
def test_recorded_bench_lines_carry_the_contract(name):
    d = _line(name)
    for k in REQUIRED:
        assert k in d, k
    assert d['metric'] == 'idispnet_roi_crops_per_s' and d['unit'] == 'ROIs/s' and d['higher_is_better'] is True
    assert d['scaling'] == 'weak' and d['vs_baseline'] is None and d['data'] == 'synthetic' and d['warmup'] >= 3
    assert 'workload' in d['config'] and 'model' not in d['config']
    e = d['e2e']
    assert e['unit'] == 'ROIs/s' and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] < d['value'] * 1.03   # (pipelined copies: e2e approaches the device-resident value; separate timed regions)
    r = d['roofline']
    assert r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['traffic'] is None or r['traffic']['dram_gb_per_step'] > 0
    assert d['gpu_launches'] > 0 and d['value'] > 0 and abs(d['value'] - d['config']['global_batch'] * 1e3 / d['ms_per_step']) / d['value'] < 1e-6
    c = d['clocks']
    assert c['sm_mhz'] and c['sm_max_mhz'] and not set(c['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    if d['n_gpus'] == 1:
        b = d['cpu_baseline']
        assert b['kind'] in ('port', 'reference') and b['cores'] >= 1 and b['value'] > 0 and b['sample']
