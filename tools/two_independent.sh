#!/bin/bash
# Two INDEPENDENT single-GPU bench processes side by side on GPUs 0 and 1 (no process group, no collective): separates
# what a second busy GPU in the box costs from what the gradient exchange costs.  usage: tools/two_independent.sh <outdir>
out=${1:-gpurun_out}
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 10 --warmup 3 --profile > $out/indep_gpu0.json 2> $out/indep_gpu0.err &
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 10 --warmup 3 --profile > $out/indep_gpu1.json 2> $out/indep_gpu1.err &
wait
cat $out/indep_gpu0.json $out/indep_gpu1.json
