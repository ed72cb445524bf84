set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sharded_final_pass or fused_kernel" 2>&1 | tail -60 | tee gpurun_out/r02_pytest_sharded.log
timeout 240 python scripts/fwd_ab.py --no-check 2>&1 | grep -E "fused|\"pair\"|cublas" | tee gpurun_out/r02_fwd_ab3.jsonl
