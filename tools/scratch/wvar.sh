#!/bin/bash
# compares builds of the wavefront select (warps per tree) on the headline search
ex='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "T2 nps", round(d["value"]), "tree_ms", round(d["config"]["tree_stream_ms_per_step"],2), "net_ms", round(d["config"]["net_ms_per_step"],2), "T1 nps", round(d["threads1"]["nps"]))'
for v in w4 default w8 w12; do
  if [ $v = default ]; then unset ARA_B200_LIB; else export ARA_B200_LIB=$PWD/build/libara_b200_$v.so; fi
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --trees 0 --selfplay-seconds 0 2>/dev/null | python -c "$ex" $v
done
export ARA_B200_LIB=$PWD/build/libara_b200_fine.so
timeout 200 python tools/prof_select.py 2>&1 | tail -12
