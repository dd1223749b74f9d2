for m in 0 1 2 3; do
TB_CONV_DBG=$m timeout 200 python bench.py --steps 5 --warmup 3 --no_cpu_baseline --graph 0 > gpurun_out/conv_dbg_$m.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/conv_dbg_$m.json').read().strip().splitlines()[-1])
print($m, d['ms_per_step'], {o['op']:round(o['ms_per_step'],3) for o in d['roofline_ops'] if 'conv1' in o['op'] or 'frames' in o['op']})
PY
done
