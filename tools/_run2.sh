export GFA_FUSED_TIMEOUT_MS=2000
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k multi_gpu 2>&1 | tail -5
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
tail -c 1200 gpurun_out/bench_r02_n2.err
