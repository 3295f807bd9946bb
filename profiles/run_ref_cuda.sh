#!/bin/bash
# profiles/run_ref_cuda.sh -- the reference's own Kokkos-CUDA build (oracle/_ref/athena_blast_cuda,
# built by `oracle/build_ref.sh cuda`) on the benchmark workload: 3-D MHD blast 256^3, one MeshBlock,
# PPM4+HLLD+RK3.  Prints the binary's zone-cycles/cpu_second line (= cell-updates/s).
N=${1:-256}; K=${2:-13}
D=$(mktemp -d); cd $D
$GRAFT_REPO_ROOT/oracle/_ref/athena_blast_cuda -i $GRAFT_REPO_ROOT/tests/decks/blast3d.athinput \
  mesh/nx1=$N mesh/nx2=$N mesh/nx3=$N meshblock/nx1=$N meshblock/nx2=$N meshblock/nx3=$N \
  time/nlim=$K output1/dcycle=-1 2>&1 | grep -E "elapsed=|zone-cycles|cpu time|FATAL|rror" | tail -8
