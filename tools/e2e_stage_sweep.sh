# debugging aid: end-to-end throughput of bench.py under explicit host-pipeline stage sizes (FSUAE_HOST_CHUNK = largest stage)
for hc in ${SWEEP:-32 48 64}; do
  echo "== host chunk [$hc]"
  FSUAE_HOST_CHUNK="$hc" timeout 300 python bench.py --steps 10 --no-cpu-baseline --stream-frames 0 --sustain 0 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('device', round(d['value']), 'streaming', round(d['e2e']['value']), 'blocking', round(d['e2e']['sync_call_value']), 'ceiling', round(d['e2e']['copy_ceiling_fps']))"
done
