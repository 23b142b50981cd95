for nb in "914 1000" "914 1024" "813 1000" "813 1024" "1003 1000" "387 1000"; do
 for cfg in "GEN=1" "BN=64" "BN=128" "BN=160"; do
  envs=""; for kv in $cfg; do envs="$envs FEO_DENSE_$kv"; done
  echo -n "$cfg: "; env $envs timeout 120 python tools/time_dense.py $nb 300 2>&1 | tail -1 | cut -c1-70
 done
done
