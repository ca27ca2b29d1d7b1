"""BASELINE.json config #1: the MNIST example with dear.DistributedOptimizer, world_size=2, CPU/gloo."""
import os
import sys

import pytest

from _mp import run_ranks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mnist_worker(rank, world):
    sys.path.insert(0, os.path.join(ROOT, "examples", "mnist"))
    import pytorch_mnist
    loss, acc = pytorch_mnist.main(["--no-cuda", "--epochs", "2", "--train-size", "1500", "--test-size", "400",
                                    "--batch-size", "50", "--log-interval", "1000", "--lr", "0.05"])
    return loss, acc


def test_mnist_two_ranks_gloo_learns():
    outs = run_ranks(mnist_worker, world=2, backend="gloo", timeout=400)
    (l0, a0), (l1, a1) = outs
    assert abs(l0 - l1) < 1e-6 and abs(a0 - a1) < 1e-6       # metric averaging agrees across ranks
    assert a0 > 0.7, "accuracy %.3f: the model did not learn" % a0


def mnist_amp_worker(rank, world):
    sys.path.insert(0, os.path.join(ROOT, "examples", "mnist"))
    import pytorch_mnist
    return pytorch_mnist.main(["--no-cuda", "--epochs", "2", "--train-size", "1500", "--test-size", "400", "--batch-size", "50",
                               "--log-interval", "1000", "--lr", "0.05", "--use-mixed-precision", "--loss-scale", "64"])


@pytest.mark.parametrize("backend", ["gloo", "emu"])
def test_mnist_mixed_precision_path_with_static_loss_scale(backend):
    """The reference's train_mixed_precision loop (skip_synchronize + scaled loss): the 1/loss_scale un-scaling is folded
    into the reduce-scatter, so training behaves like the unscaled run."""
    outs = run_ranks(mnist_amp_worker, world=2, backend=backend, timeout=400)     # emu: native runtime + direct wgrad under autocast
    (l0, a0), (l1, a1) = outs
    assert abs(l0 - l1) < 1e-6 and abs(a0 - a1) < 1e-6
    assert a0 > 0.7, "accuracy %.3f: the mixed-precision run did not learn" % a0
