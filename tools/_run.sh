export GFA_SHELL_PAIR=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_pair -s 3 -c 1 -f -o gpurun_out/prof_eval_pair_a python tools/ring_probe.py steps=1 2>&1 | grep -E "rror" | tail -3
