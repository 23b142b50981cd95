#!/bin/bash
# Developer sweeps of the preconditioner GEMM (feo_dense_tc.cu) at N = 2549, B = 1024; runs on the GPU box, prints one line per
# configuration (tools/time_dense.py).  Every run is wrapped in a short timeout: an experimental pipeline that deadlocks must
# not hold the box.
#   usage: dense_sweeps.sh [gens|tiles|pairs|small]
run() { local cfg="$1"; shift; local envs=""; for kv in $cfg; do envs="$envs FEO_DENSE_$kv"; done
        echo "== ${cfg:-default} $*"; env $envs timeout 40 python tools/time_dense.py "$@" 2>&1 | tail -1 | cut -c1-200; }
case "${1:-gens}" in
  gens)   # the three kernel generations and the fp32 FMA comparison kernel
    for cfg in "" "GEN=3" "GEN=2" "GEN=1" "SIMT=1"; do run "$cfg" 2549 1024 100; done
    for nb in "2549 2048 50" "2549 8192 20" "2680 1000 200" "1003 10000 20"; do run "GEN=2" $nb; run "GEN=3" $nb; done ;;
  tiles)  # second generation: tile widths, stages, cluster multicast of the operator stages, refill gap, drains
    for cfg in "GEN=2 BN=64" "GEN=2 BN=128 ASTAGES=3" "GEN=2 BN=128 ASTAGES=4" "GEN=2 BN=160 ASTAGES=4" "GEN=2 BN=160 ASTAGES=6" \
               "GEN=2 BN=128 ASTAGES=3 CLUSTER=2" "GEN=2 BN=160 GAP=2" "GEN=2 BN=160 GAP=4" "GEN=2 FLUSH=2" "GEN=2 FLUSH=8" "GEN=2 FLUSH=1000" \
               "GEN=2 DEBUG=1" "GEN=2 DEBUG=2"; do run "$cfg" 2549 1024 100; done ;;
  pairs)  # third generation (CTA pairs): k-blocks per stage, stages, drains, clusters of two pairs, 192-column tiles, no-MMA floor
    for cfg in "GEN=3 KG=1" "GEN=3 KG=2" "GEN=3 KG=4" "GEN=3 KG=2 ASTAGES=3" "GEN=3 KG=2 FLUSH=8" "GEN=3 KG=2 FLUSH=1000" "GEN=3 KG=2 DEBUG=2" \
               "GEN=3 BN=128" "GEN=3 BN=192" "GEN=3 BN=192 CM=2" "GEN=3 BN=160 CM=2" "GEN=3 KG=1 GAP=2" "GEN=3 KG=1 GAP=6"; do run "$cfg" 2549 1024 100; done
    run "GEN=3 KG=2" 2541 1024 50 ;;   # an odd count of k-blocks: the last stage holds one
  small)  # launch-bound sizes
    for nb in "387 1000 300" "813 1024 300" "914 1000 300" "1003 1000 300"; do for cfg in "GEN=1" "BN=64" "BN=128" "BN=160"; do run "$cfg" $nb; done; done ;;
esac
