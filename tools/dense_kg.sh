# CTA-pair kernel: k-blocks per stage (FEO_DENSE_KG)
for cfg in "GEN=3 KG=1" "GEN=3 KG=2" "GEN=3 KG=4" "GEN=3 KG=2 ASTAGES=3" "GEN=3 KG=2 DEBUG=2" "GEN=3 KG=4 DEBUG=2" "GEN=3 KG=2 BN=192" "GEN=3 KG=2 FLUSH=8"; do
  envs=""; for kv in $cfg; do envs="$envs FEO_DENSE_$kv"; done
  echo "== $cfg"; env $envs timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-200
done
echo "== KG=2 n=2541 (odd k-block count) B=1024"; FEO_DENSE_GEN=3 FEO_DENSE_KG=2 timeout 25 python tools/time_dense.py 2541 1024 20 2>&1 | tail -1 | cut -c1-200
echo "== KG=4 n=813 B=257 BN=128"; FEO_DENSE_GEN=3 FEO_DENSE_KG=4 FEO_DENSE_BN=128 timeout 25 python tools/time_dense.py 813 257 20 2>&1 | tail -1 | cut -c1-200
