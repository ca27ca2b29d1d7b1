#!/usr/bin/env python
"""Aggregate an `ncu --metrics <m1,m2,...> --csv` launch list by kernel name, one table per metric.

The analogue of the reference's horovod/extract_profilings.py, which sums `invocations x average FP32 instructions` over
an `nvprof --metrics inst_fp_32` dump (horovod/prof.sh:1-2): with

    ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum --csv --log-file launches.csv ...

this prints the time share per kernel and the FP32 instruction total (and GFLOP-instructions) per kernel.

    python tools/ncu_summary.py launches.csv [top]
"""
import collections
import csv
import re
import sys

TIME_SCALE = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per_metric = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        metric = row.get("Metric Name", "gpu__time_duration.sum")
        unit = row.get("Metric Unit", "us")
        is_time = unit in TIME_SCALE
        if is_time:
            v *= TIME_SCALE[unit]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"<.*", "", name)[:80]
        agg = per_metric.setdefault(metric, {"unit": "us" if is_time else unit, "time": is_time,
                                             "k": collections.defaultdict(lambda: [0, 0.0])})["k"]
        agg[name][0] += 1
        agg[name][1] += v
    for metric, m in per_metric.items():
        agg, unit = m["k"], m["unit"]
        tot = sum(v[1] for v in agg.values())
        n = sum(v[0] for v in agg.values())
        if len(per_metric) > 1:
            print("== %s" % metric)
        if m["time"]:
            print("total %.1f us over %d launches (serialised, cold-cache: compare shares, not absolutes)" % (tot, n))
        else:
            print("total %.4g %s over %d launches (%.3f G)" % (tot, unit, n, tot / 1e9))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            if m["time"]:
                print("%-82s n=%4d %10.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / max(tot, 1e-30)))
            else:
                print("%-82s n=%4d %12.4g %s %5.1f%%" % (k, v[0], v[1], unit, 100 * v[1] / max(tot, 1e-30)))


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit("usage: tools/ncu_summary.py <launches.csv> [top]")
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
