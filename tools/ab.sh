for v in "" variant_e5_d6.so variant_e6_d6.so variant_e8_d8.so; do
  if [ -n "$v" ]; then export GB200_LIB=$PWD/grokimagecompression_b200/$v; else unset GB200_LIB; fi
  echo "== ${v:-baseline}"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('value',j['value'],'enc',j['encode_mpix_s'],'dec',j['decode_mpix_s'],'t1e',j['t1']['encode_ms'],'t1d',j['t1']['decode_ms'],'e2e',j['e2e']['value'])
    else: print(l.rstrip()[-200:])
"
done
