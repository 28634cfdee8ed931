B="python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu"
run() { # name lib streams [env]
  name=$1; lib=$2; streams=$3; shift 3
  env SDRGPU_LIB=build_variants/libsdrgpu_$lib.so "$@" $B --streams $streams > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/v_$name.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$name', 'frac=%.4f'%r['frac'], 'k1_ms=%.4f'%r['k1_ms_per_launch'], 'value=%.0f'%d['value'], 'check', (d.get('parity_spot_check') or {}).get('ok'))
except Exception as e:
    print('$name ERR', e, open('gpurun_out/v_$name.err').read()[-400:])
PY
}
for v in "$@"; do
  IFS=: read name lib streams envs <<< "$v"
  run $name $lib $streams $envs
done
