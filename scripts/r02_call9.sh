set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r02_pytest_gpu_try3.log
timeout 200 python scripts/ab_batch.py 4096x4096 18 2>&1 | tee gpurun_out/r02_ab_batch_permatrix.jsonl
timeout 200 python scripts/ab_batch.py 11008x4096 18 2>&1 | tee -a gpurun_out/r02_ab_batch_permatrix.jsonl
