# debugging aid: end-to-end throughput of bench.py under explicit host-pipeline stage patterns (FSUAE_HOST_STAGES, repeated per submission; "" = built-in policy).
# Measured (final build): built-in 24.7 k streaming / 20.0 k blocking; 16,16,16,16 24.3 / 17.7; 8,12,16,16,12 23.5 / 19.2; 8,8,... 21.9 / 18.4
for st in "" "8,8,8,8,8,8,8,8" "8,12,16,16,12" "4,8,12,16,12,8,4" "6,10,16,16,10,6" "8,16,24,16" "16,16,16,16" "8,24,32"; do
  echo "== stages [$st]"
  FSUAE_HOST_STAGES="$st" timeout 300 python bench.py --steps 10 --no-cpu-baseline --stream-frames 0 --sustain 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('device', round(d['value']), 'streaming', round(d['e2e']['value']), 'blocking', round(d['e2e']['sync_call_value']), '1-frame host p50', round(d['latency_host_1frame_p50_ms']*1e3))"
done
