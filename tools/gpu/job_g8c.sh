mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
run 8 29703 --workload c5 > gpurun_out/r2_c5_n8.json 2> gpurun_out/r2_c5_n8.err; echo "c5 n8 rc $? lines $(wc -l < gpurun_out/r2_c5_n8.json)"; cut -c1-300 gpurun_out/r2_c5_n8.json
