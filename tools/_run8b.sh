N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --workload train --no-cpu --sustained-s 0 --dp-optimizer sharded > gpurun_out/s${N}_sharded.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/s${N}_sharded.json').read().strip().splitlines()[-1]); print('N', $N, round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, d['dp_check']['ok'])"
