set -x
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_gputest_q.log
cat gpurun_out/r2_gputest_q.log
P=$((29500 + RANDOM % 400))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_ctx.json 2> gpurun_out/r2_bench_n2_ctx.err
tail -3 gpurun_out/r2_bench_n2_ctx.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_ctx.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
for k in ('topn','topn_c5','topn_c5_context_sharded'):
    print(k, d.get(k,{}).get('ms_per_query_batch'), d.get(k,{}).get('value'))
PY
