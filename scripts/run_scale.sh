# One rank per GPU, the command the driver uses: bash scripts/run_scale.sh N  ->  gpurun_out/r2_final_nN.json (+ a short summary)
set -x
N=$1
P=$((29500 + RANDOM % 400))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_n$N.json 2> gpurun_out/r2_final_n$N.err
tail -2 gpurun_out/r2_final_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_final_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.4g ms %.4f'%(d['value'], d['ms_per_step']), 'e2e %.3g blocking %.3g'%(d['e2e']['value'], d['e2e'].get('value_blocking_partial_fit',0)))
print('dp_check', d.get('dp_check',{}).get('ok'), d.get('dp_exchange'))
for k in ('topn','topn_c5','topn_c5_context_sharded'):
    print(k, d.get(k,{}).get('ms_per_query_batch'), '%.3g'%d.get(k,{}).get('value',0))
PY
