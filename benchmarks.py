#!/usr/bin/env python
"""Batch runner: sweep methods x tasks x #GPUs, resumable, scrapes "Total ... on N GPU(s)" lines.

Counterpart of the reference's benchmarks.py:10-176 (which sweeps method x task x nworkers x rdma
over mpirun + hostfiles).  Here a job is one ``torchrun`` on this node; the "rdma" axis (10 GbE vs
InfiniBand) has no meaning on NVLink and is replaced by the dtype axis.

    python benchmarks.py                         # full sweep -> logs/<prefix>/*.log, reports.json
    python benchmarks.py --methods dear,ddp --tasks resnet50:64 --gpus 1,2,4,8 --dry-run
"""
from __future__ import annotations

import argparse
import itertools
import json
import os
import socket
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

# (task, per-GPU batch size) — the reference's matrix (benchmarks.py:21)
TASKS = [("resnet50", 64), ("densenet201", 32), ("inceptionv4", 64), ("bert_base", 64), ("bert", 32), ("vgg16", 64)]
METHODS_TF = ["horovod", "ddp", "mgwfbp", "dear"]                 # with tensor fusion
METHODS_NOTF = ["wfbp", "bytescheduler", "dear-notf"]             # without tensor fusion
NUM_OF_TRIES = 1


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def gen_cmd(method, task, bs, ngpu, dtype, extra):
    driver = "bert_benchmark.py" if task.startswith("bert") else "imagenet_benchmark.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ngpu),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "benchmarks", driver),
           "--model", task, "--batch-size", str(bs), "--method", method, "--dtype", dtype]
    if task.startswith("bert"):
        cmd += ["--sentence-len", "64"]                          # the reference launcher's value (horovod_mpi_cj.sh:6)
    return cmd + list(extra)


def log_name(prefix, method, task, bs, ngpu, dtype):
    return os.path.join(ROOT, "logs", prefix, "dtype-%s-method-%s-dnn-%s-bs-%d-gpus-%d.log" % (dtype, method, task, bs, ngpu))


def extract_log(path):
    """Return the number in the last 'Total ... on N GPU(s): X +-Y' line (reference benchmarks.py:119-128)."""
    try:
        with open(path) as f:
            lines = [l for l in f if l.startswith("Total ") and "(s):" in l]
        if not lines:
            return None
        return float(lines[-1].split("(s):")[1].split("+-")[0])
    except (OSError, ValueError, IndexError):
        return None


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--methods", default=",".join(METHODS_TF + METHODS_NOTF))
    ap.add_argument("--tasks", default=",".join("%s:%d" % t for t in TASKS))
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--dtypes", default="fp32")
    ap.add_argument("--prefix", default="b200")
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--timeout", type=int, default=1800)
    ap.add_argument("extra", nargs="*", help="extra args passed to the drivers (after --)")
    args = ap.parse_args(argv)
    methods = args.methods.split(",")
    tasks = [(t.split(":")[0], int(t.split(":")[1])) for t in args.tasks.split(",")]
    gpus = [int(g) for g in args.gpus.split(",")]
    dtypes = args.dtypes.split(",")
    if not args.dry_run:
        os.makedirs(os.path.join(ROOT, "logs", args.prefix), exist_ok=True)
    ledger = os.path.join(ROOT, "logs", args.prefix, "exp.log")        # resumable, like the reference (:86-99)
    done = set()
    if os.path.exists(ledger):
        with open(ledger) as f:
            done = set(l.strip() for l in f)
    reports = {}
    for dtype, (task, bs), ngpu, method in itertools.product(dtypes, tasks, gpus, methods):
        key = "%s|%s|%d|%d|%s" % (method, task, bs, ngpu, dtype)
        path = log_name(args.prefix, method, task, bs, ngpu, dtype)
        cmd = gen_cmd(method, task, bs, ngpu, dtype, args.extra)
        if args.dry_run:
            print(" ".join(cmd), ">", path)
            continue
        if key not in done:
            for _ in range(NUM_OF_TRIES):
                with open(path, "w") as f:
                    try:
                        rc = subprocess.call(cmd, stdout=f, stderr=subprocess.STDOUT, timeout=args.timeout, cwd=ROOT)
                    except subprocess.TimeoutExpired:
                        rc = -1
                if rc == 0 and extract_log(path) is not None:
                    break
                time.sleep(2)
            with open(ledger, "a") as f:
                f.write(key + "\n")
        reports[key] = extract_log(path)
        print(key, reports[key], flush=True)
    if not args.dry_run:
        with open(os.path.join(ROOT, "logs", args.prefix, "reports.json"), "w") as f:
            json.dump(reports, f, indent=1)


if __name__ == "__main__":
    main()
