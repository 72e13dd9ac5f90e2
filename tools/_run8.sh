N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$2" = "check" ]; then
  timeout 600 $TR tools/p2p_check.py --quick > gpurun_out/p2p_check_${N}gpu.log 2>&1; echo "p2p_check rc=$?"; grep "sharded\|ok\|Error\|rror:" gpurun_out/p2p_check_${N}gpu.log | tail -8 | cut -c1-500
fi
timeout 600 $TR bench.py --gpus $N --workload train --no-cpu --sustained-s 0 --dp-optimizer sharded > gpurun_out/scale${N}_sharded.json 2> gpurun_out/scale${N}_sharded.err; echo "bench sharded rc=$?"; tail -3 gpurun_out/scale${N}_sharded.err | cut -c1-300
wc -l gpurun_out/scale${N}_sharded.json
