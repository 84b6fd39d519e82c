"""bench.py is what the driver measures with, and its GPU arm cannot run in the build container. This drives
bench.main() end to end on the CPU with the device layer replaced by recorders (fake torch.cuda entry points, fake
JxgContext / Batch), real synthetic frames, the real host front-end and the real oracle: it checks the script's control
flow and the shape of the JSON line, not any number."""
import json
import sys
import types

import pytest


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 1.0


class _FakeStream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass


def _install_fakes(monkeypatch, fail_e2e=False):
    import torch
    import bench
    from jxl_rs_b200 import decoder
    import jxl_rs_b200 as j

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _FakeStream())
    real_empty, real_tensor = torch.empty, torch.tensor
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)

    class FakeCtx:
        def __init__(self, device=0):
            self.device = device

        def close(self):
            pass

    class FakeBatch:
        runs = 0

        def __init__(self, ctx, n=0, staging_threads=0):
            self.n = 0

        def add(self, fr, ptr, stride, fmt, out_is_device):
            assert fr.width > 0 and ptr != 0 and stride >= fr.width * 3
            self.n += 1

        def set_profile(self, on):
            pass

        def run(self, stream_ptr=0):
            FakeBatch.runs += 1
            if fail_e2e and FakeBatch.runs > 4:
                raise RuntimeError("injected failure of the end-to-end leg")

        def rerun_device(self, stream_ptr=0):
            pass

        def wait(self):
            pass

        def stage_times(self):
            return {"memset": 0.1, "entropy": 3.0, "dequant_idct": 1.0, "epf2": 1.2}

        def stats(self):
            return {"h2d_bytes": 1000, "d2h_bytes": 2000, "kernel_launches": 7, "device_ms": 5.5}

        def close(self):
            pass

    for mod in (decoder, j):
        monkeypatch.setattr(mod, "JxgContext", FakeCtx)
        monkeypatch.setattr(mod, "Batch", FakeBatch)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    return bench


@pytest.mark.parametrize("fail_e2e", [False, True])
def test_bench_main_control_flow(monkeypatch, capsys, fail_e2e):
    bench = _install_fakes(monkeypatch, fail_e2e)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "1", "--steps", "3", "--warmup", "3", "--frames", "3",
                                      "--width", "320", "--height", "200", "--cpu-sample-frames", "1", "--lf-tree", "1"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["config"]["frames_per_gpu"] == 3
    assert "weighted-predictor" in line["config"]["workload"]
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    if fail_e2e:
        assert line["e2e"]["value"] is None and "injected failure" in line["e2e"]["error"]
        assert line["value"] > 0  # the device-resident measurement survives
    else:
        assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 1000
