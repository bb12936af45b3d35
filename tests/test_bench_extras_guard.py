"""bench.py runs the extra workloads in a child process under a time limit; the parent only parses and aggregates.  These tests drive
the parent side with a scripted child (no GPU): a complete result, a child killed half way (the finished workloads survive), a child
that dies without a result."""
import importlib.util
import json
import os
import subprocess
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def bench(monkeypatch):
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    import vlfm_b200.utils.dist as d

    monkeypatch.setattr(d, "max_over_ranks", lambda v, dev: v)          # world size 1: the maximum over ranks is the local value
    return b


ARGS = types.SimpleNamespace(extra_batch=32, extra_budget=5.0)


def _line(b, out, pending):
    return b.EXTRAS_MARK + json.dumps({"out": out, "pending": pending})


def test_complete_result_is_aggregated(bench, monkeypatch):
    out = {"peak_hbm_gbs": 6500.0, "configs1_b32": {"value": 0}, "configs2_full_step": {"value": 1}, "configs3_slice": {"error": "boom"},
           "configs4_slice": {"value": 2}}

    def run(cmd, **kw):
        assert "--extras-child" in cmd and kw["timeout"] == 5.0 and kw["env"]["LOCAL_RANK"] == "0"
        return types.SimpleNamespace(returncode=0, stderr="", stdout="noise\n" + _line(bench, out, [[32, 6, 0.5], [32, 4, 0.4], [8, 4, 0.2]]) + "\n")

    monkeypatch.setattr(subprocess, "run", run)
    got = bench.run_extras(ARGS, None, 1, 0, 0)
    assert got["configs1_b32"]["value"] == pytest.approx(32 * 6 / 0.5)
    assert got["configs2_full_step"]["value"] == pytest.approx(32 * 4 / 0.4)
    assert got["configs4_slice"]["value"] == pytest.approx(8 * 4 / 0.2)
    assert got["configs3_slice"] == {"error": "boom"} and "error" not in got


def test_a_killed_child_keeps_the_finished_workloads(bench, monkeypatch):
    first = _line(bench, {"configs1_b32": {"value": 0}}, [[32, 6, 0.5]])
    second = _line(bench, {"configs1_b32": {"value": 0}, "configs2_full_step": {"value": 1}}, [[32, 6, 0.5], [32, 4, 0.4]])

    def run(cmd, **kw):
        raise subprocess.TimeoutExpired(cmd, kw["timeout"], output=(first + "\n" + second + "\n").encode())

    monkeypatch.setattr(subprocess, "run", run)
    got = bench.run_extras(ARGS, None, 1, 0, 0)
    assert got["configs2_full_step"]["value"] == pytest.approx(320.0) and "killed" in got["error"]
    assert "configs3_slice" not in got


def test_a_child_without_result_is_an_error_entry(bench, monkeypatch):
    monkeypatch.setattr(subprocess, "run", lambda cmd, **kw: types.SimpleNamespace(returncode=1, stderr="Traceback ... RuntimeError: no device", stdout=""))
    got = bench.run_extras(ARGS, None, 1, 0, 0)
    assert "no device" in got["error"]
