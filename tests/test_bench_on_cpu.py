"""bench.py's GPU arms end to end on the CPU: the library emulation stands in for
libb200sph.so (tests/test_library_on_cpu.py) and a few lines of fake `torch.cuda` (events
that read the wall clock, no-op synchronize / set_device, pin_memory = identity) stand in
for the device.  This checks the code NO CPU test otherwise executes -- workload set-up,
pair counting, the timed loop, the e2e loop, the roofline / cpu_baseline arithmetic and the
JSON line's keys -- at sizes that take seconds.  The numbers it produces mean nothing."""
import importlib
import io
import json
import os
import sys
import time

import pytest

from test_library_on_cpu import emulated_library  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event(object):
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max(1e3 * (other.t - self.t), 1e-3)


class _Stream(object):
    cuda_stream = 0


@pytest.fixture
def fake_cuda(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self, *a, **k: self)
    return torch


CONTRACT = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
            'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'clocks',
            'e2e', 'gpu_launches', 'roofline', 'cpu_baseline']


@pytest.mark.parametrize('argv', [
    ['--dx', '0.08'],
    ['--workload', 'taylor_green', '--nx', '10'],
    ['--workload', 'rings', '--dx', '0.0025', '--lz', '0.0075'],
], ids=['dam_break', 'taylor_green', 'rings'])
def test_bench_gpu_arm_on_the_emulated_library(emulated_library, fake_cuda, monkeypatch,  # noqa: F811
                                               capfd, argv):
    sys.path.insert(0, ROOT)
    bench = importlib.import_module('bench')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--steps', '3', '--warmup', '3',
                                      '--e2e-steps', '2', '--cpu-budget', '5'] + argv)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        monkeypatch.delenv(k, raising=False)
    saved = os.dup(1)
    try:
        bench.main()
    finally:
        os.dup2(saved, 1)           # bench points fd 1 at stderr until its line is ready
        os.close(saved)
    out = capfd.readouterr().out
    lines = [ln for ln in out.splitlines() if ln.startswith('{')]
    assert len(lines) == 1, out[-2000:]
    d = json.loads(lines[0])
    for k in CONTRACT:
        assert k in d, k
    assert d['unit'] == 'pairs/s' and d['n_gpus'] == 1 and d['steps'] == 3 and d['warmup'] == 3
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['higher_is_better'] is True
    assert d['gpu_launches'] > 0 and d['config']['pairs_per_step'] > 0
    assert 'workload' in d['config'] and 'model' not in d['config']
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and r['peak'] > 0
    c = d['cpu_baseline']
    assert c['kind'] == 'port' and c['cores'] == 1 and c['value'] > 0 and c['sample']
