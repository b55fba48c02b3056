#!/bin/bash
# Runs the GPU test groups in SEPARATE processes (a faulting kernel poisons its CUDA context) and keeps head + tail of
# each log under gpurun_out/.  usage: tools/gpu_tests.sh [group ...]
mkdir -p gpurun_out
run() {  # name, pytest args...
  name=$1; shift
  timeout 900 python -m pytest "$@" -q -m gpu $PYTEST_X > gpurun_out/t_$name.full 2>&1
  { head -c 30000 gpurun_out/t_$name.full; echo; echo ...; tail -c 12000 gpurun_out/t_$name.full; } > gpurun_out/t_$name.log
  rm -f gpurun_out/t_$name.full
  echo "== $name: $(tail -1 gpurun_out/t_$name.log)"
}
groups=${@:-"wgrad bwd train fast conv masking solver loss graph"}
for g in $groups; do
  case $g in
    wgrad) run wgrad tests/test_bwd_kernels_gpu.py -k "wgrad or stride2 or convtranspose" ;;
    bwd) run bwd tests/test_bwd_kernels_gpu.py -k "not (wgrad or stride2 or convtranspose)" ;;
    train) run train tests/test_trainpath_gpu.py ;;
    fast) run fast tests/test_fastpath_gpu.py ;;
    conv) run conv tests/test_conv_gpu.py ;;
    masking) run masking tests/test_masking_gpu.py ;;
    solver) run solver tests/test_solver_gpu.py ;;
    loss) run loss tests/test_loss_gpu.py ;;
    graph) run graph tests/test_graph_gpu.py ;;
  esac
done
