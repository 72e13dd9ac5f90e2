#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list + full capture of the three hot kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tests] [smoke] [bench] [kbench] [launches] [ncu] [infer] [p2p]
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
for w in $what; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log ;;
    probe)  # hardware questions answered by tiny stand-alone programs (tools/probes)
      timeout 60 tools/probes/gather4_probe > gpurun_out/gather4_probe.log 2>&1; echo "gather4 probe rc=$?"; cat gpurun_out/gather4_probe.log ;;
    tests_split)  # one pytest process per file: a trapped kernel only takes its own file down
      for f in tests/test_*.py; do
        n=$(basename $f .py)
        timeout 900 python -m pytest $f -m gpu -q -x > gpurun_out/pt_$n.log 2>&1; echo "$n rc=$?" | tee -a gpurun_out/pt_$n.log; tail -2 gpurun_out/pt_$n.log | head -1
      done ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
      timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json ;;
    kbench)
      timeout 300 python tools/kbench.py 30 > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log ;;
    ncufwd)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-^head_fwd}" -s 4 -c 2 -f -o gpurun_out/prof_k \
        python tools/kbench.py 2 > gpurun_out/ncu_k.log 2>&1; echo "ncuk rc=$?" ;;
    infer)
      timeout 300 python bench.py --workload infer --protos 1000000 --steps 10 --warmup 3 > gpurun_out/infer_1m.json 2> gpurun_out/infer_1m.err; echo "infer1m rc=$?"
      timeout 300 python bench.py --workload infer --protos 10000000 --steps 5 --warmup 3 --no-cpu > gpurun_out/infer_10m.json 2> gpurun_out/infer_10m.err; echo "infer10m rc=$?" ;;
    ncuinfer)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^(proto_retrieve|proto_refine|head_fwd|fuse_flat|fuse_rows)" -s 15 -c 5 -f -o gpurun_out/prof_infer \
        python bench.py --workload infer --protos 1000000 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_infer.log 2>&1; echo "ncuinfer rc=$?" ;;
    p2p)  # needs gpurun --gpus N (N = 2, 4 or 8): gradient exchange kernels against NCCL + data-parallel bench
      N=$(nvidia-smi -L | wc -l)
      T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540"
      timeout 300 $T tools/p2p_check.py --quick > gpurun_out/p2p_check_${N}gpu.log 2>&1; echo "p2p_check rc=$?"; tail -14 gpurun_out/p2p_check_${N}gpu.log
      timeout 600 $T bench.py --gpus $N --no-cpu > gpurun_out/scale${N}.json 2> gpurun_out/scale${N}.err; echo "bench$N rc=$?"; tail -3 gpurun_out/scale${N}.err ;;
    p2pcmp)  # the exchange variants side by side (training half only)
      N=$(nvidia-smi -L | wc -l)
      T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
      for c in fused p2p nvls nccl; do
        timeout 300 $T bench.py --gpus $N --no-cpu --workload train --sustained-s 0 --dp-comm $c > gpurun_out/scale${N}_$c.json 2> gpurun_out/scale${N}_$c.err; echo "bench$N $c rc=$?"
      done ;;
    klaunch)
      timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 30 -c 60 --csv --log-file gpurun_out/klaunch.csv \
        python tools/kbench.py 3 > gpurun_out/klaunch.log 2>&1; echo "klaunch rc=$?" ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        python bench.py --workload train --sustained-s 0 --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/bench_under_ncu.log 2>&1; echo "launches rc=$?" ;;
    ncu)
      timeout 1200 ncu --set full --clock-control none --import-source on \
        -k regex:"${NCU_KERNELS:-^(cast_weight|fuse_|hav_|head_|label_xyz)}" -s ${NCU_SKIP:-28} -c ${NCU_COUNT:-7} -f -o gpurun_out/prof_train \
        python bench.py --workload train --sustained-s 0 --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?" ;;
  esac
done
