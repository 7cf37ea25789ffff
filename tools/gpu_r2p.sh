#!/bin/bash
# 1 GPU: score / fused-tail parity tests only (default = sorted search in the general route), then the same file with the
# experimental tree search
timeout 120 python -m pytest tests/test_gpu_scores.py -x -q -m gpu 2>&1 | tail -2
ALQ_TAIL_TREE=1 timeout 120 python -m pytest tests/test_gpu_scores.py -x -q -m gpu -k "fused_tail" 2>&1 | tail -2
