# strong-scaling run of the sharded bench: bash tools/scale_run.sh <N> <config> [extra bench args]
N=$1; CFG=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --config $CFG --no-cpu-baseline "$@" > gpurun_out/scale_${CFG}_n${N}.json 2> gpurun_out/scale_${CFG}_n${N}.err
tail -2 gpurun_out/scale_${CFG}_n${N}.err | cut -c1-300
python -c "
import json,sys
d=json.load(open('gpurun_out/scale_${CFG}_n${N}.json'))
print('N=$N cfg=$CFG value', round(d['value']), 'ms', round(d['ms_per_step']), 'e2e', round(d['e2e']['value']), d['scaling'], 'steps', [round(x) for x in d.get('step_wall_ms',[])])
print(' stage_ms', {k: round(v) for k,v in d['stage_ms'].items()})
"
