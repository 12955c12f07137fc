export GFA_FUSED_TIMEOUT_MS=2000
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/ring_probe.py steps=30 2>&1 | grep -E "RESULT|rror" | cut -c1-200
