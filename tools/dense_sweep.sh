#!/bin/bash
# sweep of the second-generation dense kernel's tile widths / stages
for cfg in "GEN=1" "BN=64" "BN=128 ASTAGES=3" "BN=128 ASTAGES=4" "BN=128 ASTAGES=5" "BN=160 ASTAGES=4" "BN=160 ASTAGES=5" "BN=160 ASTAGES=6" "BN=128 ASTAGES=3 CLUSTER=2" "BN=160 ASTAGES=6 CLUSTER=2" "AUTO=1"; do
  envs=""
  for kv in $cfg; do envs="$envs FEO_DENSE_$kv"; done
  echo "== $cfg"
  env $envs timeout 120 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1
done
echo "== sizes (auto)"
for nb in "387 1000" "914 1000" "2680 1000" "813 1024" "2549 8192" "2549 100"; do
  timeout 120 python tools/time_dense.py $nb 50 2>&1 | tail -1
done
