echo "== gen2"; timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-200
for cfg in "BN=160" "BN=128" "BN=160 ASTAGES=6" "BN=160 ASTAGES=4"; do
  envs="FEO_DENSE_GEN=3"; for kv in $cfg; do envs="$envs FEO_DENSE_$kv"; done
  echo "== gen3 $cfg"; env $envs timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-200
done
echo "== gen3 B=8192"; FEO_DENSE_GEN=3 timeout 25 python tools/time_dense.py 2549 8192 20 2>&1 | tail -1 | cut -c1-200
echo "== gen3 n=2680 B=1000 (odd tiles)"; FEO_DENSE_GEN=3 FEO_DENSE_BN=160 timeout 25 python tools/time_dense.py 2680 1000 20 2>&1 | tail -1 | cut -c1-200
echo "== gen3 n=300 B=100"; FEO_DENSE_GEN=3 FEO_DENSE_BN=128 timeout 25 python tools/time_dense.py 300 100 20 2>&1 | tail -1 | cut -c1-200
