# debugging aid: end-to-end throughput of bench.py under explicit host-pipeline stage schedules ("" = built-in policy)
for st in ${SWEEP:-"" "16" "8"}; do
  echo "== stages [$st]"
  FSUAE_HOST_STAGES="$st" timeout 300 python bench.py --steps 10 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('device', round(d['value']), 'streaming', round(d['e2e']['value']), 'blocking', round(d['e2e']['sync_call_value']))"
done
