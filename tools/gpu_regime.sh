# The "two regimes" of the alias-sampler kernel: bench runs with and without the first launches
# that sart_create makes (warm_f32), three fresh processes each.
set -x
out=gpurun_out/regime
mkdir -p $out
rm -f $out/regime.log
for w in 1 0 1 0 1 0; do
  if [ $w = 0 ]; then export SART_NO_WARM=1; else unset SART_NO_WARM; fi
  echo "== warm=$w" >> $out/regime.log
  timeout 300 python bench.py --sampler ${SMP:-alias} --steps 5 --warmup 3 --no-cpu-baseline --no-presampled --no-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('headline ms_per_step', d['ms_per_step'], 'kernel', d['kernel_ms_over_ranks']['median'])" >> $out/regime.log
done
cat $out/regime.log
