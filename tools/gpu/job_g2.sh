mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded_over_two_gpus" > gpurun_out/r2_nccl_two_gpu_test.log 2>&1; echo "2-gpu test rc $?"; tail -3 gpurun_out/r2_nccl_two_gpu_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload c4 > gpurun_out/r2_c4_n2.json 2> gpurun_out/r2_c4_n2.err; echo "c4 n2 rc $?"; cat gpurun_out/r2_c4_n2.json; tail -3 gpurun_out/r2_c4_n2.err
python bench.py --workload c4 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; echo "c4 n1 rc $?"; cat gpurun_out/r2_c4_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "c3 n2 rc $?"; cut -c1-900 gpurun_out/r2_bench_n2.json; tail -3 gpurun_out/r2_bench_n2.err
