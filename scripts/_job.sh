set -x
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -4
bash scripts/_job_scale.sh 2
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_n2.json').read().strip().splitlines()[-1])
print('lists_identical', d['topn_c5_context_sharded'].get('lists_identical_to_item_sharded'))
PY
