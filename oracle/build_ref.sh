#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Builds the unmodified reference into oracle/_ref/ (see oracle/ref.mk).  Two binaries are
# needed because a user problem generator replaces the built-in ones (src/pgen/pgen.cpp:911-916):
#   athena_builtin : linear wave, Orszag-Tang, shock tube ...
#   athena_blast   : PROBLEM=fluids/blast
# Skips quietly when /root/reference is absent (GPU box: prebuilt binaries travel with the repo).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-8}"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF not present, nothing to do"; exit 0; fi
export PATH="$HERE/tools:$PATH"     # bc shim for Makefile.kokkos
build() { # name problem
  mkdir -p "$OUT/build_$1"
  make -C "$OUT/build_$1" -f "$HERE/ref.mk" REF="$REF" PROBLEM="$2" NAME="$1" OUTDIR="$OUT" -j"$JOBS" -s
}
if [ "$1" = "cuda" ]; then
  # optional GPU baseline: the reference's generic Kokkos-CUDA path compiled for sm_100
  # (SURVEY.md section 0 fact 4); ~10 min.  Not needed by any test.
  mkdir -p "$OUT/build_athena_blast_cuda"
  NVCC_WRAPPER_DEFAULT_COMPILER=g++ make -C "$OUT/build_athena_blast_cuda" -f "$HERE/ref.mk" REF="$REF" \
    PROBLEM=fluids/blast NAME=athena_blast_cuda OUTDIR="$OUT" KDEV=Cuda,Serial KARCH=Blackwell100 \
    RCXX="$REF/kokkos/bin/nvcc_wrapper" ROMP= -j"$JOBS" -s
  echo "build_ref: cuda ok"; exit 0
fi
build athena_blast fluids/blast
build athena_builtin built_in_pgens
echo "build_ref: ok"
