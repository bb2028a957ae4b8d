# Carve-out hypothesis for the two regimes of the alias kernel: alias-only processes with the shared-memory carve-out the
# alias kernel asks for (default), with the one the inverse-CDF kernel asks for (41 %), and a few others.
set -x
out=gpurun_out/regime
mkdir -p $out
rm -f $out/regime2.log
for pct in default 13 41 41 50 100 default; do
  if [ $pct = default ]; then unset SART_CARVEOUT_PCT; else export SART_CARVEOUT_PCT=$pct; fi
  echo "== carveout=$pct" >> $out/regime2.log
  timeout 300 python bench.py --sampler alias --steps 5 --warmup 3 --no-cpu-baseline --no-presampled --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('alias ms_per_step', d['ms_per_step'], 'kernel', d['kernel_ms_over_ranks']['median'])" >> $out/regime2.log
done
cat $out/regime2.log
