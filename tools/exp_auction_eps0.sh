for e in default 0 0.2 1.0 4.0; do
  if [ "$e" = default ]; then unset GHICP_AUCTION_EPS0; else export GHICP_AUCTION_EPS0=$e; fi
  echo "=== EPS0=$e"
  GHICP_AUCTION_DEBUG=1 timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu 2>gpurun_out/exp_$e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print([ (x['it'], round(x['ms'],1), x['rounds'], x['cor'], round(x['km_energy'],3)) for x in d['first_iterations']], d['ms_per_step'])
"
  grep -E "free-object|phase" gpurun_out/exp_$e.err | head -24 | cut -c1-150
done
