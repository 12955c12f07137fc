export GFA_FUSED_TIMEOUT_MS=2000
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench.py --steps 30 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 1500 gpurun_out/bench_r02_a.err
timeout 600 python tools/measure_traffic.py 2>&1 | tail -30
cp profiles/traffic_r02.json profiles/eval_pipe_r02.json gpurun_out/
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; tail -c 600 gpurun_out/bench_r02_ref.err
