set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_gputest_full2.log
cat gpurun_out/r2_gputest_full2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2_smoke2.log 2>&1; tail -2 gpurun_out/r2_smoke2.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_v2.json 2> gpurun_out/r2_bench_n1_v2.err; tail -3 gpurun_out/r2_bench_n1_v2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_v2.json').read().strip().splitlines()[-1])
print('value %.4g ms %.4f'%(d['value'], d['ms_per_step']), 'e2e', d['e2e']['value'], d['e2e'].get('value_blocking_partial_fit'))
print('roofline', {k:d['roofline'].get(k) for k in ('bound','achieved','peak','frac','traffic')})
print('roofline_hbm', {k:d.get('roofline_hbm',{}).get(k) for k in ('achieved','frac')})
for k,v in d.get('models',{}).items(): print(' ',k, v.get('ms_per_step'), v.get('roofline',{}).get('frac'), v.get('error'))
print('topn', d['topn']['ms_per_query_batch'], d['topn']['roofline']['frac'], 'topn_c5', d['topn_c5']['ms_per_query_batch'], d['topn_c5']['roofline']['frac'])
print('bands', json.dumps(d.get('parity_bands'))[:600])
print('epoch', d.get('e2e_epoch'))
print('cpu', d.get('cpu_baseline'))
print('clocks', d.get('clocks'))
PY
