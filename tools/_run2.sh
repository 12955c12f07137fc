timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "multi_gpu or host_positions" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 --no-side-configs > gpurun_out/bench_r02_n2b.json 2> gpurun_out/bench_r02_n2b.err
tail -c 600 gpurun_out/bench_r02_n2b.err
