"""bench.py host logic that never needs a GPU: argument presets, the config dict both arms must share, the per-launch
timing wrappers / aggregation (with stand-in events), the ceiling fallback, and the reference arm's JSON contract."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def parse(bench, monkeypatch, *argv):
    monkeypatch.setattr(sys, "argv", ["bench.py", *argv])
    return bench.parse_args()


def test_presets_follow_baseline_configs(bench, monkeypatch):
    from autosmoothquant_b200.harness import CONFIGS

    a = parse(bench, monkeypatch)
    assert (a.config, a.model, a.seq, a.batch, a.gpus, a.impl) == (2, "llama-2-7b", 2048, 1, 1, "ours") and a.steps >= 1 and a.warmup >= 3
    assert bench.quant_overrides(a) == {}
    a3 = parse(bench, monkeypatch, "--config", "3")
    assert a3.model == "llama-2-13b" and a3.batch == 32 and a3.parallel == "dp"
    assert bench.quant_overrides(a3) == {"out": "per-token", "fc2": "per-token"}
    a4 = parse(bench, monkeypatch, "--config", "4", "--gpus", "8")
    assert a4.model == "mixtral-8x7b" and a4.fixed_global_batch == 1 and bench.quant_overrides(a4) == {"fc1": "per-token", "fc2": "per-token"}
    a5 = parse(bench, monkeypatch, "--config", "5")
    assert a5.model == "llama-2-70b" and bench.quant_overrides(a5)["type"] == "fp8"
    for a_ in (a, a3, a4, a5):
        assert a_.model in CONFIGS
    # both arms print the SAME config dict for the same command line (the driver compares them)
    cfg = CONFIGS[a.model]
    for world in (1, 2, 8):
        c = bench.workload_config(a, cfg, cfg.layers, world)
        assert c["parallelism"] == ("dp1" if world == 1 else f"tp{world}") and c["global_batch"] == world and c["baseline_config"] == 2
        assert "workload" in c and "model" not in c
    assert bench.workload_config(a5, CONFIGS[a5.model], 80, 8)["global_batch"] == 1  # one 2048-token sequence over 8 GPUs
    assert "Llama-2-7B INT8" in bench.metric_name(a, cfg)


class FakeEvent:
    clock = 0.0

    def record(self):
        FakeEvent.clock += 0.25  # ms
        self.t = FakeEvent.clock

    def elapsed_time(self, other):
        return other.t - self.t


def test_launch_timers_and_aggregation(bench):
    events = []
    timed, timed_prologue = bench.make_launch_timers(events, FakeEvent, max_routed_rows=4096)
    x = torch.zeros(2048, 4096)
    w = torch.zeros(11008, 4096, dtype=torch.int8)
    calls = []
    linear = timed(lambda *a, **k: calls.append("lin") or "y", "w8a8_linear_q8")
    assert linear(x, w, None, 1.0, out_dtype=torch.bfloat16) == "y"
    quant = timed_prologue(lambda x_, *a, **k: calls.append("q") or ("q8", None))
    assert quant(x, 1, 0.05) == ("q8", None)
    peer = timed(lambda self_, x_, w_, *a, **k: "z", "linear_q8_allreduce_nvls(GEMM+all-reduce, one launch)", collective=True)
    assert peer(object(), x, torch.zeros(4096, 4096, dtype=torch.int8), None, 1.0) == "z"
    grouped = timed(lambda *a, **k: "g", "w8a8_grouped_linear")
    xs = torch.zeros(6144, 4096)  # 4096 routed rows + tile padding
    grouped(xs, torch.zeros(8 * 1024, 4096, dtype=torch.int8), torch.zeros(48, dtype=torch.int32), torch.zeros(8))
    assert calls == ["lin", "q"] and len(events) == 4
    agg = bench.aggregate_launch_events(events)
    assert agg["n_lin"] == 3 and agg["n_coll"] == 1
    assert agg["lin_ops"] == 2.0 * 2048 * 4096 * 11008 + 2.0 * 4096 * 4096 * 1024  # the prologue adds time, not ops
    assert agg["lin_time"] == pytest.approx(3 * 0.25e-3) and agg["coll_time"] == pytest.approx(0.25e-3)
    assert agg["coll_bytes"] == 2048 * 4096 * 2.0
    by = {b["entry"]: b for b in agg["by_launch_shape"]}
    assert by["quantize_act (stand-alone prologue launch)"]["tops"] == 0.0 and by["quantize_act (stand-alone prologue launch)"]["M"] == 2048
    assert by["w8a8_grouped_linear"]["M"] == 4096 and by["w8a8_grouped_linear"]["N"] == 1024
    assert by["w8a8_linear_q8"]["avg_us"] == pytest.approx(250.0)
    assert bench.aggregate_launch_events([])["by_launch_shape"] == []


def test_tensor_ceiling_falls_back_to_the_committed_capture(bench):
    out, src = bench.tensor_ceiling(live=False)
    assert src == "profiles/int8_ceiling.json" and 3000 < out["best"]["i8"] < 5000 and 2500 < out["best"]["fp8"] < 5000
    for v in out["variants"]:
        assert v["rc"] == 0 and abs(v["cycles_per_mma"] - 128.0) < 0.1  # 8192 MAC / clk / SM
    live, src_live = bench.tensor_ceiling(live=True)  # no GPU here: every variant fails, the capture is used
    assert src_live == "profiles/int8_ceiling.json" and live == out


def test_reference_arm_prints_the_contract_line():
    """--impl reference: the reference's unmodified Linear classes on the host cores, one bounded sample per step."""
    if not (ROOT / "baseline" / "_ref" / "autosmoothquant" / "layers" / "nn" / "linear.py").exists():
        pytest.skip("baseline/_ref not staged (__graft_entry__.build() copies it where /root/reference exists)")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--seq", "256"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["scaled_by"] == 32
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["parallelism"] == "dp1" and line["config"]["seq_len"] == 256
    # rank != 0 under torchrun: exits 0 without work
    q = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference"], capture_output=True, text=True, timeout=120,
                       env={**__import__("os").environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert q.returncode == 0 and q.stdout.strip() == ""


def test_auto_reduction_policy(bench):
    """auto: peer stores at 2 GPUs, the in-switch kernel only at world sizes where it has run (8; at 2 FP8 has no other fused
    option and probes for a multicast address), NCCL elsewhere; an explicit choice is never overridden."""
    auto = SimpleNamespace(tp_reduce="auto")
    assert bench.pick_tp_reduce(auto, 2, None, False) == ("fused", None)
    mode, note = bench.pick_tp_reduce(auto, 4, None, False)
    assert mode == "nccl" and "world 4" in note
    assert bench.pick_tp_reduce(auto, 4, None, True)[0] == "nccl"
    for explicit in ("nccl", "nccl-int32", "fused", "fused-int32", "nvls"):
        assert bench.pick_tp_reduce(SimpleNamespace(tp_reduce=explicit), 4, None, False) == (explicit, None)
