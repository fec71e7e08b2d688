# usage: tools/quick_bench.sh [extra bench args]; prints a compact summary of the bench line
python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value',j['value'],'enc',j['encode_mpix_s'],'dec',j['decode_mpix_s'],'t1e',j['t1']['encode_ms'],'t1d',j['t1']['decode_ms'],'dwt GB/s',j['roofline']['achieved'],'frac',j['roofline']['frac'],'e2e',j['e2e']['value'],'launches',j['gpu_launches'])
    else: print(l.rstrip()[-300:])
"
