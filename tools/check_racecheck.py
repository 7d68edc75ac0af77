"""Asserts that a compute-sanitizer racecheck log of the CTA-pair screen kernel reports nothing but the paired tensor-memory
allocation: `tcgen05.alloc.cta_group::2` writes the allocated TMEM address into the shared-memory slot of BOTH CTAs of the pair
(the same value, the allocation sequence of every 2-SM kernel), which racecheck sees as two writers.  Any other hazard fails.

    python tools/check_racecheck.py profiles/r2_racecheck_tc_pair.log [source line of the alloc]
"""
import re
import subprocess
import sys


def alloc_line():
    src = open("pyatmosphere_b200/csrc/screen_tc.cu").read().splitlines()
    return [i + 1 for i, ln in enumerate(src) if "tcgen05.alloc.cta_group::2" in ln]


def main(path):
    text = open(path).read()
    allowed = set(alloc_line())
    hazards = re.findall(r"(?:Error|Warning): .*?(?=\n=========\s*\n|\Z)", text, flags=re.S)
    bad = []
    for h in hazards:
        lines = {int(m) for m in re.findall(r"screen_tc\.cu:(\d+)", h)}
        if not lines or not lines <= allowed:
            bad.append(h[:400])
    summary = re.search(r"RACECHECK SUMMARY: (\d+) hazards displayed \((\d+) errors?, (\d+) warnings?\)", text)
    print(f"{path}: {len(hazards)} hazard reports, {len(bad)} outside the paired TMEM allocation (screen_tc.cu:{sorted(allowed)}); "
          f"summary: {summary.group(0) if summary else 'none'}")
    if bad:
        print("\n".join(bad))
        sys.exit(1)


if __name__ == "__main__":
    main(sys.argv[1])
