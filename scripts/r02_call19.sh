set -x
PROBE_GB=3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/exchange_probe.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_exchange_probe.log
