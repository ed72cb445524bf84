set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "linear_forward_stat or fused_calibration or tiny_opt_pipeline or absstat or forward" 2>&1 | tail -25 | tee gpurun_out/r02_pytest_n3.log
mkdir -p gpurun_out/cli && cd gpurun_out/cli && rm -rf cache output
for mode in hook fused; do rm -rf cache output; ( ASVD_B200_CALIB=$mode timeout 300 python ../../asvd.py --synthetic_model opt-125m --calib_dataset synthetic --act_aware --alpha 0.5 --n_calib_samples 16 --scaling_method abs_mean --param_ratio_target 0.9 --sensitivity_metric stable_rank ) 2>&1 | grep -E "phase times|calib_ppl" | sed "s/^/$mode: /"; done | tee ../r02_cli_calibration_fused_vs_hook.log
cd ../.. && rm -rf gpurun_out/cli
