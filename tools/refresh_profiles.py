"""Copies the outputs of tools/gpu/job_final.sh from gpurun_out/ into profiles/ and re-stamps profiles/r2_traffic.json with the
dram__bytes of the ncu captures taken by that job and the sha1 of the FFT sources they were taken from (bench.py reports
`roofline.traffic` only while the sources still match the stamp).   python tools/refresh_profiles.py"""
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def dram_bytes(summary):
    text = open(os.path.join(OUT, summary)).read()
    get = lambda key: float(re.search(key + r"\.sum \[Mbyte\] = ([0-9.]+)", text).group(1)) * 1e6
    return int(round(get("dram__bytes_read") + get("dram__bytes_write")))


traffic_path = os.path.join(PROF, "r2_traffic.json")
traffic = json.load(open(traffic_path))
traffic["k_cols"] = dram_bytes("r2_prof_cols_summary.txt")
traffic["k_rows(ifft*screen*fft)"] = dram_bytes("r2_prof_rows_summary.txt")
traffic["kernel_source_stamp"] = bench.kernel_source_stamp()
json.dump(traffic, open(traffic_path, "w"), indent=1)

for src, dst in [("r2_prof_cols_summary.txt", "r2_prof_cols_tma_summary.txt"), ("r2_prof_rows_summary.txt", "r2_prof_rows_summary.txt"),
                 ("r2_gpu_tests_final.log", "r2_gpu_tests.log"), ("r2_smoke_final.log", "r2_smoke.log"),
                 ("r2_fft_variants_final.log", "r2_fft_variants.log"), ("r2_launches_final.csv", "r2_launches.csv"),
                 ("r2_step_breakdown_final.txt", "r2_step_breakdown.txt"), ("r2_step_breakdown_chunk32.txt", "r2_step_breakdown_chunk32.txt"),
                 ("r2_c4_final.json", "r2_c4_n1.json"), ("r2_c5_final.json", "r2_c5_n1.json")]:
    shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))

# the bench line of that job ran before the re-stamp: fill in the traffic of the captures taken by the same job
line = json.load(open(os.path.join(OUT, "r2_bench_final.json")))
line["roofline"]["traffic"] = traffic["k_cols"]
line["roofline"]["traffic_source"] = "ncu --set full capture, profiles/r2_traffic.json"
json.dump(line, open(os.path.join(PROF, "r2_bench_n1.json"), "w"))
print(json.dumps(traffic))
