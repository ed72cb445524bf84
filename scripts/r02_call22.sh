set -x
mkdir -p gpurun_out/cli && cd gpurun_out/cli && rm -rf cache output
( time CUDA_VISIBLE_DEVICES=0 timeout 400 python ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 ) > ../r02_cli_opt125m_1gpu.log 2>&1
rm -rf cache output
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 ) > ../r02_cli_opt125m_2gpu.log 2>&1
rm -rf cache output
( time CUDA_VISIBLE_DEVICES=0 timeout 400 python ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 --eval_batch_size 1 ) > ../r02_cli_opt125m_1gpu_evalbatch1.log 2>&1
cd ../.. && rm -rf gpurun_out/cli
grep -h "phase times\|calib_ppl\|sharded final" gpurun_out/r02_cli_opt125m_1gpu.log gpurun_out/r02_cli_opt125m_2gpu.log gpurun_out/r02_cli_opt125m_1gpu_evalbatch1.log
python - <<'PY'
import re
def table(path):
    t = {}
    for l in open(path):
        m = re.match(r"^(\S+) (0\.\d+) ([\d.eE+-]+)$", l.strip())
        if m: t[(m.group(1), m.group(2))] = float(m.group(3))
    return t
a, b, c = table("gpurun_out/r02_cli_opt125m_1gpu.log"), table("gpurun_out/r02_cli_opt125m_2gpu.log"), table("gpurun_out/r02_cli_opt125m_1gpu_evalbatch1.log")
print(len(a), len(b), len(c), "rank-0 rows on 2 GPUs are a subset:", all(a[k] == v for k, v in b.items()))
print("max rel diff eval batch 8 vs 1:", max(abs(a[k] - c[k]) / c[k] for k in a))
PY
