TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python tools/dp_overlap_probe.py | grep "gg_"
timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_head_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 $TR bench.py --gpus 2 --workload train --no-cpu --sustained-s 0 --dp-optimizer sharded > gpurun_out/s2.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/s2.json').read().strip().splitlines()[-1]); print('N 2', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, d['dp_check']['ok'])"
