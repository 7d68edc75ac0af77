"""Print the headline metrics and the top stall instructions of an .ncu-rep (run where ncu is installed)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__average_warps_issue_stalled"]


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("==", row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
        for h, u, v in zip(hdr, units, row):
            if any(h == k or (k.endswith("stalled") and h.startswith(k) and h.endswith("per_issue_active.ratio")) for k in KEYS):
                print(f"  {h} [{u}] = {v}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    a, b, c = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [(int(r[b] or 0), r[a].strip(), int(r[c] or 0)) for r in rows[hi + 1:] if len(r) > b]
    tot = sum(d[0] for d in data) or 1
    print(f"-- top stall samples (of {tot}, {len(data)} SASS instructions) --")
    for s, txt, ex in sorted(data, reverse=True)[:top]:
        print(f"  {100 * s / tot:5.1f}%  exec={ex:9d}  {txt[:100]}")
    # instruction mix
    mix = {}
    for s, txt, ex in data:
        op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
        op = op.split(".")[0]
        mix[op] = mix.get(op, 0) + ex
    tot_ex = sum(mix.values()) or 1
    print("-- executed instruction mix --")
    print("  " + ", ".join(f"{k}:{100 * v / tot_ex:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:18]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
