"""Per-kernel share of one bench step from an ncu launch list (--metrics gpu__time_duration.sum --csv).

    python tools/launch_breakdown.py gpurun_out/launches.csv

A step starts at the device-RNG kernel (k_rng_spectrum); the last step of the usual length is reported."""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))
    starts = [i for i, (k, _) in enumerate(rows) if "k_rng_spectrum" in k]
    if len(starts) < 2:
        raise SystemExit("fewer than two steps in the launch list")
    # steps of the device-RNG throughput loop all have the same number of launches; later launches (host-buffer leg, roofline
    # legs) follow the last device-RNG kernel and are not delimited by it
    spans = [(starts[k], starts[k + 1]) for k in range(len(starts) - 1)]
    lengths = [b - a for a, b in spans]
    usual = max(set(lengths), key=lengths.count)
    a0, b0 = [sp for sp, ln in zip(spans, lengths) if ln == usual][-1]
    step = rows[a0:b0]
    step = [(k, t) for k, t in step if "k_" in k]       # this library's kernels (ncu prints some without their namespace)
    total = sum(t for _, t in step)
    agg = OrderedDict()
    for k, t in step:
        a = agg.setdefault(k, [0.0, 0])
        a[0] += t
        a[1] += 1
    print(f"one bench step = 8 realizations of config 3 (ncu --metrics gpu__time_duration.sum, cold-cache, serialised): total {total:.1f} us, "
          f"{len(step)} launches")
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {t:8.1f} us  {c:2d}x  {100 * t / total:5.1f}%  {k[:150]}")


if __name__ == "__main__":
    main(sys.argv[1])
