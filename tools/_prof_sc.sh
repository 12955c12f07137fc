for u in 1 4; do
GFA_SCATTER_UNROLL=$u ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 3 -c 1 -f -o gpurun_out/prof_scatter_u$u python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
done
