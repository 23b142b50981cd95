# CTA-pair kernel: clusters of two pairs sharing the operator stages (FEO_DENSE_CM=2), 192-column tiles
for cfg in "GEN=3" "GEN=3 BN=192" "GEN=3 BN=192 CM=2" "GEN=3 BN=192 CM=2 ASTAGES=5" "GEN=3 BN=160 CM=2" "GEN=3 BN=128 CM=2" "GEN=3 BN=192 CM=2 DEBUG=2"; do
  envs=""; for kv in $cfg; do envs="$envs FEO_DENSE_$kv"; done
  echo "== $cfg"; env $envs timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-200
done
echo "== 813x257 CM=2 BN=128"; FEO_DENSE_GEN=3 FEO_DENSE_BN=128 FEO_DENSE_CM=2 timeout 25 python tools/time_dense.py 813 257 20 2>&1 | tail -1 | cut -c1-200
