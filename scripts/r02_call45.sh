set -x
cd /root/repo
for mode in default tri oddeven; do
  if [ $mode = default ]; then unset ASVD_B200_SOLVE; else export ASVD_B200_SOLVE=$mode; fi
  echo "== $mode" | tee -a gpurun_out/r02_tri_rect.log
  MAXB=8 SHAPES=11008x4096,4096x11008,768x3072,3072x768 RATIO=0.9 timeout 600 python scripts/check_shapes.py 2>&1 | tee -a gpurun_out/r02_tri_rect.log
done
