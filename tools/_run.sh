export GFA_FUSED_TIMEOUT_MS=2000
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
