GFA_SETUP_TIMING=1 timeout 300 python tools/ring_probe.py steps=3 2>&1 | grep -E "gfa\]|RESULT" | cut -c1-160
export GFA_FUSED_TIMEOUT_MS=2000
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
