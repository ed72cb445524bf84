set -x
for lead in lo hi; do ASVD_B200_LEAD=$lead PROF_BATCH=18 timeout 100 python scripts/time_classes.py; done 2>&1 | tee gpurun_out/r02_lead_warp_ab.log
ASVD_B200_LEAD=lo timeout 200 python scripts/ab_batch.py 4096x4096 18 2>&1 | tee -a gpurun_out/r02_lead_warp_ab.log
ASVD_B200_LEAD=hi timeout 200 python scripts/ab_batch.py 4096x4096 18 2>&1 | tee -a gpurun_out/r02_lead_warp_ab.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_cases or batched_equals or full_size_4096" 2>&1 | tail -3
