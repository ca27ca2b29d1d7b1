"""Small utilities: clock-sample summary (bench.py's `clocks` field), the input prefetcher on a CPU device, the ncu
launch-list aggregator."""
import pytest
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clock_sampler_summary_and_missing_nvidia_smi():
    from dear_pytorch_b200.utils.clocks import ClockSampler
    s = ClockSampler()
    now = time.time()
    s.rows = [(now - 3, 1965.0, 1965.0, 310.0, ["Not Active"] * 4),
              (now - 2, 1950.0, 1965.0, 320.0, ["Not Active", "Not Active", "Not Active", "Active"]),
              (now - 1, 1965.0, 1965.0, 300.0, ["Not Active"] * 4)]
    out = s.summary()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 3
    assert s.summary(now - 1.5, now)["samples"] == 1
    empty = ClockSampler()
    assert empty.summary()["samples"] == 0
    empty.start()          # no nvidia-smi on the CPU box: must not raise
    empty.stop()


def test_prefetcher_on_cpu_yields_batches_in_order():
    from dear_pytorch_b200.utils.data import PinnedPrefetcher, SyntheticImages
    src = SyntheticImages(2, image_size=8, num_classes=5, n_buffers=3)
    feed = PinnedPrefetcher(iter(src), torch.device("cpu"), depth=2)
    got = [next(feed) for _ in range(5)]
    for i, (x, y) in enumerate(got):
        ex, ey = src.batches[i % 3]
        assert torch.equal(x, ex) and torch.equal(y, ey)
    finite = PinnedPrefetcher(iter(src.batches[:2]), torch.device("cpu"))
    assert len(list(finite)) == 2


def test_ncu_launch_list_aggregation():
    csv = ('"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device",'
           '"CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
           '"0","1","python","h","dear::rs_kernel<float, 1, false>(dear::RSParams)","1","7","(512, 1, 1)","(128, 1, 1)","0","10.0","X","gpu__time_duration.sum","us","10.0"\n'
           '"1","1","python","h","dear::rs_kernel<float, 1, false>(dear::RSParams)","1","7","(512, 1, 1)","(128, 1, 1)","0","10.0","X","gpu__time_duration.sum","us","14.0"\n'
           '"2","1","python","h","void cudnn::conv(float*)","1","7","(256, 1, 1)","(64, 1, 1)","0","10.0","X","gpu__time_duration.sum","ns","6000"\n')
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write("==PROF== Connected\n" + csv)
        path = f.name
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path], capture_output=True, text=True,
                             check=True).stdout
    finally:
        os.unlink(path)
    assert "total 30.0 us over 3 launches" in out
    assert "dear::rs_kernel" in out and "n=   2" in out and "80.0%" in out


def test_ncu_summary_sums_instruction_counters_like_the_reference_extractor():
    """horovod/extract_profilings.py of the reference: invocations x FP32 instructions, summed over an nvprof dump."""
    hdr = ('"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device",'
           '"CC","Section Name","Metric Name","Metric Unit","Metric Value"\n')
    row = '"%d","1","python","h","%s","1","7","(256, 1, 1)","(64, 1, 1)","0","10.0","X","%s","%s","%s"\n'
    csv = hdr + row % (0, "gemm(float*)", "gpu__time_duration.sum", "us", "5.0") \
        + row % (0, "gemm(float*)", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "inst", "3,000,000,000") \
        + row % (1, "relu(float*)", "gpu__time_duration.sum", "us", "5.0") \
        + row % (1, "relu(float*)", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "inst", "1,000,000,000")
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write(csv)
        path = f.name
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path], capture_output=True, text=True,
                             check=True).stdout
    finally:
        os.unlink(path)
    assert "== gpu__time_duration.sum" in out and "total 10.0 us over 2 launches" in out
    assert "total 4e+09 inst over 2 launches (4.000 G)" in out and "75.0%" in out


def test_prefetcher_upload_delay_is_a_noop_on_cpu():
    from dear_pytorch_b200.utils.data import PinnedPrefetcher, SyntheticImages
    src = SyntheticImages(2, image_size=8, num_classes=5, n_buffers=2)
    feed = PinnedPrefetcher(iter(src), torch.device("cpu"), upload_delay_us=1500.0)
    assert feed._delay_cycles == 0
    x, y = next(feed)
    assert torch.equal(x, src.batches[0][0])



def test_reference_utils_helpers():
    import numpy as np
    from dear_pytorch_b200.utils import misc
    assert len(misc.gen_random_id()) == 64 and misc.gen_random_id() != misc.gen_random_id()
    with tempfile.TemporaryDirectory() as d:
        p = misc.create_path("a/b", base=d)
        assert os.path.isdir(p) and misc.create_path("a/b", base=d) == p
    timers = {}
    misc.force_insert_item(timers, "w", 0.1)
    misc.force_insert_item(timers, "w", 0.2)
    assert timers == {"w": [0.1, 0.2]}
    idx, vals = misc.topk(np.array([0.1, -5.0, 0.3, 4.0, -0.2]), 2)
    assert sorted(idx.tolist()) == [1, 3] and sorted(vals.tolist()) == [-5.0, 4.0]
    assert [misc.get_approximate_sigma_scale(d) for d in (0.9, 0.5, 0.03, 0.001)] == [0.5, 1.5, 2.0, 3.0]

    class Item:
        size = None
        def set_fontsize(self, s): self.size = s
    class Axis:
        def __init__(self): self.label = Item()
    class Ax:
        def __init__(self):
            self.title, self.xaxis, self.yaxis, self.ticks, self.texts = Item(), Axis(), Axis(), [Item(), Item()], []
        def get_xticklabels(self): return self.ticks[:1]
        def get_yticklabels(self): return self.ticks[1:]
        def text(self, x, y, s, **kw): self.texts.append((x, y, s))
    class Rect:
        def get_y(self): return 0.0
        def get_height(self): return 2.0
        def get_x(self): return 1.0
        def get_width(self): return 0.5
    ax = Ax()
    misc.update_fontsize(ax, 9)
    assert ax.title.size == ax.xaxis.label.size == ax.ticks[1].size == 9
    misc.autolabel([Rect()], ax, "x1.9")
    assert ax.texts == [(1.25, 2.06, "x1.9")]
