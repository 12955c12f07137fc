export GFA_FUSED_TIMEOUT_MS=20000
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ring.py -q -x -k "oracle_and_classic and (shell or mixed or unconstrained) or scrambled" 2>&1 | tail -6
echo "=== memcheck on boundary / shell load / shipped meshes"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "shell_load or host_positions or set_dofs_refuses or shipped or tutorial01 or smoke" 2>&1 | tail -6
echo "=== racecheck on the ring kernels (shared-memory staging)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ring.py -q -x -k "oracle_and_classic and beam" 2>&1 | tail -6
echo "=== host mirror"
timeout 900 python -m pytest tests/test_host_mirror.py -q -m gpu 2>&1 | tail -4
