python -m pytest tests/test_gpu_screen_tc.py tests/test_gpu_c3_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3
for f in 0 1; do echo "fork=$f"; PYATM_TC_FORK=$f python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline_screen']['us_per_screen'], d['roofline_screen']['frac'], d['gpu_launches'])"; done
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_screen_tc.py -m gpu -x -q -k "tc_screen_vs_oracle_256" > gpurun_out/r2_racecheck_tc_pair.log 2>&1; python tools/check_racecheck.py gpurun_out/r2_racecheck_tc_pair.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_c3_parity.py -m gpu -x -q -k "simulate_batch" 2>&1 | tail -3
