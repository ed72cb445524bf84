set -x
scripts/probes/ffma2_probe.bin | tee gpurun_out/r02_ffma2_probe.log
for b in 18; do timeout 200 python scripts/ab_batch.py 11008x4096 $b; timeout 200 python scripts/ab_batch.py 4096x11008 $b; done 2>&1 | tee gpurun_out/r02_ab_batch18.jsonl
