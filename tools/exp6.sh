#!/usr/bin/env bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_attention.py -q 2>&1 | tail -60 > gpurun_out/exp6_attn.log
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/exp6_all.log
echo done
