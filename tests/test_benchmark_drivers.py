"""The command-line drivers (benchmarks/imagenet_benchmark.py, benchmarks/bert_benchmark.py) run every ``--method`` on two
CPU ranks and print the reference's log lines; the batch runner (benchmarks.py) launches a job, scrapes it and resumes.

Reference: dear/imagenet_benchmark.py:138-172, dear/bert_benchmark.py:138-175 (log format), benchmarks.py:86-151
(ledger, scrape, reports.json)."""
import contextlib
import io
import json
import os
import shutil
import subprocess
import sys

import pytest

from _mp import ROOT, run_ranks

sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

TINY = ["--batch-size", "1", "--no-cuda", "--num-warmup-batches", "1", "--num-batches-per-iter", "1", "--num-iters", "2"]


def _imagenet_worker(rank, world, method, extra):
    import imagenet_benchmark as drv
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = drv.main(["--model", "resnet18", "--fused-bn", "0", "--image-size", "64", "--method", method] + TINY + list(extra))
    return res, buf.getvalue()


IMAGENET_CASES = [
    ("dear", ()), ("dear-notf", ()), ("dear-naive", ()), ("dear-wt", ()), ("dear-rb", ()), ("dear", ("--optimizer", "adamw")),
    ("dear", ("--exclude-parts", "allgather")),
    ("wfbp", ()), ("mgwfbp", ()), ("asc", ()), ("ddp", ()), ("ddp-zero", ()), ("horovod", ()), ("bytescheduler", ()),
    ("single", ()),
    # flags of the reference's baseline drivers (horovod/, pytorch-ddp/, bytescheduler/ imagenet_benchmark.py)
    ("horovod", ("--fp16-allreduce",)), ("horovod", ("--use-adasum",)), ("ddp", ("--use-zero", "1")),
    ("bytescheduler", ("--partition", "100000")),
    ("wfbp", ("--compressor", "eftopk", "--density", "0.01")),
    ("wfbp", ("--compressor", "gtopk", "--density", "0.01", "--momentum", "0.9", "--momentum-correction")),
]


@pytest.mark.parametrize("method,extra", IMAGENET_CASES, ids=["%s%s" % (m, ("-" + "-".join(a.strip("-") for a in e)) if e else "")
                                                              for m, e in IMAGENET_CASES])
def test_imagenet_driver_runs_every_method_on_two_ranks(method, extra):
    outs = run_ranks(_imagenet_worker, world=2, backend="gloo", args=(method, extra), timeout=300)
    res, text = outs[0]
    assert res["total"] > 0 and res["iter_time_s"] > 0
    lines = text.strip().split("\n")
    assert any(l.startswith("Iter #1: ") and l.endswith("img/sec per GPU") for l in lines)
    assert lines[-1].startswith("Total img/sec on 2 CPU(s): ")
    assert outs[1][1].strip() == "" or "Total" not in outs[1][1]          # only rank 0 logs


EMU_CASES = [("dear", ()), ("dear-bo", ()), ("dear-rb", ()), ("dear", ("--dtype", "amp")), ("dear", ("--dtype", "bf16")),
             ("dear", ("--optimizer", "adam", "--exclude-parts", "reducescatter"))]


@pytest.mark.parametrize("method,extra", EMU_CASES, ids=["%s%s" % (m, ("-" + "-".join(a.strip("-") for a in e)) if e else "")
                                                         for m, e in EMU_CASES])
def test_imagenet_driver_on_the_native_runtime_host_emulation(method, extra):
    """Same command lines on DEAR_BACKEND=emu: the C++ runtime a GPU run uses (bucket sets, pack tables, flag protocol,
    direct wgrad, sharded update), with the kernels emulated on the host; autocast and bf16 parameters included."""
    outs = run_ranks(_imagenet_worker, world=2, backend="emu", args=(method, extra), timeout=300)
    res, text = outs[0]
    assert res["total"] > 0
    assert "backend: emu" in text and text.strip().split("\n")[-1].startswith("Total img/sec on 2 CPU(s): ")


def _bert_worker(rank, world, method, cfg_path):
    import bert_benchmark as drv
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = drv.main(["--config", cfg_path, "--sentence-len", "16", "--method", method] + TINY)
    return res, buf.getvalue()


@pytest.mark.parametrize("method", ["dear", "dear-bo", "ddp", "bytescheduler"])
def test_bert_driver_on_a_tiny_config(tmp_path, method):
    cfg = tmp_path / "tiny_bert.json"
    cfg.write_text(json.dumps({"vocab_size": 96, "hidden_size": 32, "num_hidden_layers": 2, "num_attention_heads": 2,
                               "intermediate_size": 64, "max_position_embeddings": 32}))
    outs = run_ranks(_bert_worker, world=2, backend="gloo", args=(method, str(cfg)), timeout=300)
    res, text = outs[0]
    assert res["total"] > 0
    assert text.strip().split("\n")[-1].startswith("Total sentences/sec on 2 CPU(s): ")


def test_batch_runner_launches_scrapes_and_resumes():
    prefix = "pytest_runner_%d" % os.getpid()
    logdir = os.path.join(ROOT, "logs", prefix)
    cmd = [sys.executable, os.path.join(ROOT, "benchmarks.py"), "--methods", "dear", "--tasks", "resnet18:1", "--gpus", "2",
           "--prefix", prefix, "--timeout", "300", "--", "--fused-bn", "0", "--image-size", "64"] + TINY[2:]
    env = dict(os.environ, DEAR_BACKEND="gloo")
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        key = "dear|resnet18|1|2|fp32"
        with open(os.path.join(logdir, "reports.json")) as f:
            rep = json.load(f)
        assert rep[key] is not None and rep[key] > 0
        with open(os.path.join(logdir, "exp.log")) as f:
            assert f.read().split() == [key]
        # second invocation: the ledger says done -> nothing is launched, the number is scraped from the existing log
        log = os.path.join(logdir, "dtype-fp32-method-dear-dnn-resnet18-bs-1-gpus-2.log")
        before = os.path.getmtime(log)
        out2 = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
        assert out2.returncode == 0 and os.path.getmtime(log) == before
        assert ("%s %s" % (key, rep[key])) in out2.stdout
    finally:
        shutil.rmtree(logdir, ignore_errors=True)


def test_batch_runner_dry_run_lists_the_reference_matrix():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks.py"), "--dry-run", "--gpus", "8"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT).stdout
    cmds = [l for l in out.split("\n") if "torch.distributed.run" in l]
    assert len(cmds) == 6 * 7                                  # tasks x methods (reference benchmarks.py:21,10-19)
    assert any("bert_benchmark.py --model bert --batch-size 32 --method dear" in c and "--sentence-len 64" in c for c in cmds)
    assert any("imagenet_benchmark.py --model vgg16 --batch-size 64 --method bytescheduler" in c for c in cmds)


def test_launch_script_env_flags():
    """scripts/launch.sh keeps the reference launcher's environment "flags" (dear/horovod_mpi_cj.sh:2-29): a compressor
    selects the sparse WFBP baseline with density 0.001 and one 64 Mi-element group."""
    env = dict(os.environ, DEAR_BACKEND="gloo", dnn="resnet18", bs="1", nworkers="2", compressor="eftopk", PY=sys.executable)
    out = subprocess.run([os.path.join(ROOT, "scripts", "launch.sh"), "--fused-bn", "0", "--image-size", "64"] + TINY[2:],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Method: wfbp" in out.stdout and "Total img/sec on 2 CPU(s): " in out.stdout
