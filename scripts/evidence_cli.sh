# End-to-end runs of the upstream CLI surface (asvd.py) on the GPU box, synthetic OPT-125m (BASELINE configs[0]).
set -x
mkdir -p gpurun_out/cli && cd gpurun_out/cli && rm -rf cache output
( time timeout 240 python ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 \
    --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 ) > ../r01_cli_opt125m_abs_mean.log 2>&1
ls -la cache >> ../r01_cli_opt125m_abs_mean.log
( time timeout 100 python ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 \
    --n_calib_samples 4 --scaling_method fisher_abs_mean --sensitivity_metric stable_rank --param_ratio_target 0.9 ) > ../r01_cli_opt125m_fisher_stable_rank.log 2>&1
ls -la cache >> ../r01_cli_opt125m_fisher_stable_rank.log
cd ../.. && rm -rf gpurun_out/cli
tail -n 12 gpurun_out/r01_cli_opt125m_abs_mean.log; tail -n 12 gpurun_out/r01_cli_opt125m_fisher_stable_rank.log
