# sweep of the refill gap (k-blocks between a stage's MMAs and its refill) of the dense kernels
for g in 2 3 4; do echo "== gen2 BN=160 S=6 gap $g"; FEO_DENSE_GAP=$g timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-150; done
for g in 2 3; do echo "== gen2 BN=160 S=6 gap $g no MMA"; FEO_DENSE_DEBUG=2 FEO_DENSE_GAP=$g timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-80; done
for g in 2 4 6; do echo "== gen3 BN=160 S=8 gap $g"; FEO_DENSE_GEN=3 FEO_DENSE_GAP=$g timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-150; done
echo "== gen3 gap 4 no MMA"; FEO_DENSE_GEN=3 FEO_DENSE_DEBUG=2 FEO_DENSE_GAP=4 timeout 25 python tools/time_dense.py 2549 1024 50 2>&1 | tail -1 | cut -c1-80
echo "== gen2 BN=64 S=4 gap 2 / 3 at n=914"; for g in 2 3; do FEO_DENSE_GAP=$g timeout 25 python tools/time_dense.py 914 1000 200 2>&1 | tail -1 | cut -c1-80; done
