#!/bin/bash
# Multi-GPU pass (under `gpurun --gpus N`): bash tools/scale_check.sh N [check]
#   check: tests/test_dp_gpu.py, tools/p2p_check.py and tools/symm_probe.py first; then the default bench line on N GPUs
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$2" = "check" ]; then
  timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q > gpurun_out/pt_dp_${N}gpu.log 2>&1; echo "pytest dp rc=$?"; tail -2 gpurun_out/pt_dp_${N}gpu.log
  timeout 600 $TR tools/p2p_check.py --quick > gpurun_out/p2p_check_${N}gpu.log 2>&1; echo "p2p_check rc=$?"; grep "sharded\|ok\|Error\|rror:" gpurun_out/p2p_check_${N}gpu.log | tail -8 | cut -c1-700
  timeout 300 $TR tools/symm_probe.py > gpurun_out/symm_probe_${N}gpu.log 2>&1; echo "symm_probe rc=$?"; grep -v "^W\|^\[\|^\*\|OMP" gpurun_out/symm_probe_${N}gpu.log | tail -6 | cut -c1-300
fi
timeout 900 $TR bench.py --gpus $N --no-cpu > gpurun_out/scale${N}_default.json 2> gpurun_out/scale${N}_default.err; echo "bench default rc=$?"; tail -2 gpurun_out/scale${N}_default.err | cut -c1-300
wc -l gpurun_out/scale${N}_default.json
