export GFA_FUSED_TIMEOUT_MS=2000 GFA_FUSED_DEBUG=1
timeout 1500 python -m pytest tests/test_gpu_ring.py -q 2>&1 | tail -5
