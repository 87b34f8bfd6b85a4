#!/bin/bash
# Everything profiles/r2_* is made from, in ONE gpurun call (tag = $1): GPU tests, both bench arms, the other BASELINE
# configs, the ncu captures of the step and of the 80-channel (tcgen05) blend kernels.
tag=${1:-r2z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_tests.log
python bench.py --steps 30 --warmup 5 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench_ours.json
python bench.py --impl reference --steps 10 --warmup 3 2>gpurun_out/${tag}_ref.err | tail -1 > gpurun_out/${tag}_bench_reference.json
for c in config3 config4 config5; do
  python bench.py --workload $c --steps 10 --warmup 3 2>gpurun_out/${tag}_$c.err | tail -1 > gpurun_out/${tag}_$c.json
done
OCRF_TC=0 python bench.py --workload config4 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${tag}_config4_cuda_cores.json
bash tools/profile_step.sh ${tag}
ncu --set full --clock-control none --import-source on -k regex:"render_forward_tc|render_backward_tc" -c 2 -f \
    -o gpurun_out/${tag}_generic python tools/generic_profile.py > gpurun_out/${tag}_generic.log 2>&1
cat gpurun_out/${tag}_tests.log
