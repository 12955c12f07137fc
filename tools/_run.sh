export GFA_FUSED_TIMEOUT_MS=2000
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -k "shell_load or host_positions or set_dofs_refuses" 2>&1 | tail -8
