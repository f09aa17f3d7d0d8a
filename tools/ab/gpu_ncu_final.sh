set -x
for k in overlay_kernel contour_largest_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${k}_r1c python tools/prof_post_kernels.py > gpurun_out/ncu_${k}_c.log 2>&1
  tail -1 gpurun_out/ncu_${k}_c.log
done
