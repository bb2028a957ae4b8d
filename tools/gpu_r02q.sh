set -x
out=gpurun_out/r02q
mkdir -p $out
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -3 $out/bench.err; cat $out/bench.json
