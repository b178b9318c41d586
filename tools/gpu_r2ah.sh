#!/bin/bash
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -k "attention" 2>&1 | tail -6
HSIMAE_ATTN_SMALL=0 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --tb=short -k "attention" 2>&1 | tail -3
