set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"render_forward_wide|render_backward_generic" -s 4 -c 2 -o gpurun_out/r1_generic_full -f python tools/generic_profile.py > gpurun_out/ncu_generic.log 2>&1
tail -2 gpurun_out/ncu_generic.log
