set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contour_largest_kernel -s 1 -c 1 -f -o gpurun_out/prof_contour_largest_kernel_r1b python tools/prof_post_kernels.py > gpurun_out/ncu_contour_b.log 2>&1
tail -2 gpurun_out/ncu_contour_b.log
