set -x
timeout 120 python tools/prof_post_kernels.py 2>&1 | tail -2
for k in overlay_kernel contour_largest_kernel fold_average_threshold_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${k}_r1 python tools/prof_post_kernels.py > gpurun_out/ncu_${k}.log 2>&1
  tail -2 gpurun_out/ncu_${k}.log
done
