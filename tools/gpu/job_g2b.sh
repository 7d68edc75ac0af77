mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded_over_two_gpus" > gpurun_out/r2_nccl_two_gpu_test.log 2>&1; echo "2-gpu test rc $?"; tail -3 gpurun_out/r2_nccl_two_gpu_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --workload c4 > gpurun_out/r2_c4_n2.json 2> gpurun_out/r2_c4_n2.err; echo "c4 n2 rc $?"; grep '^{' gpurun_out/r2_c4_n2.json | cut -c1-250; tail -2 gpurun_out/r2_c4_n2.err
python bench.py --workload c4 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; echo "c4 n1 rc $?"; cut -c1-250 gpurun_out/r2_c4_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "c3 n2 rc $?"; grep '^{' gpurun_out/r2_bench_n2.json | cut -c1-300; tail -2 gpurun_out/r2_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>&1 | grep '^{' | cut -c1-200
