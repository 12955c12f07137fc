#!/bin/bash
# The measurement pass a round ends with (one GPU, under gpurun): tests, smoke, the bench line, the reference arm,
# the ncu launch list and --set full captures of the two kernels of a step, and the DRAM traffic of a step.
# Usage: tools/final_run.sh <tag>      outputs under gpurun_out/
tag=${1:-r02_final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gputests_${tag}.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${tag}.txt 2>&1
python bench.py 2> gpurun_out/bench_${tag}.err | tail -1 > gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 3 --warmup 1 --no-full-size-step 2> gpurun_out/bench_${tag}_ref.err | tail -1 > gpurun_out/bench_${tag}_reference_arm.json
bash tools/profile.sh ${tag} > /dev/null 2>&1
python tools/ncu_keys.py gpurun_out/prof_eval_${tag}.ncu-rep > gpurun_out/prof_eval_${tag}_summary.csv
python tools/ncu_keys.py gpurun_out/prof_scatter_${tag}.ncu-rep > gpurun_out/prof_scatter_${tag}_summary.csv
python tools/measure_traffic.py > gpurun_out/traffic_${tag}.txt 2>&1
cat gpurun_out/gputests_${tag}.txt gpurun_out/smoke_${tag}.txt; cut -c1-600 gpurun_out/bench_${tag}.json; cat gpurun_out/traffic_${tag}.txt | tail -3
