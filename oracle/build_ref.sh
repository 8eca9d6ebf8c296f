#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own scalar host oracles (spmm_host_coo / spmm_host_csr /
# spmm_host) and partitioners IN PLACE from /root/reference into oracle/_ref/*.so,
# one library per (variant, dtype) because the reference selects val_dt with -D flags
# (backend_pim/spmm_default/support/common.h:39-60, CMakeLists.txt:13-19,37-41).
# The UPMEM SDK is absent, so <dpu.h> is replaced by the aborting stubs in
# oracle/ref_stubs/ (only host arithmetic is ever called).  No reference source is
# copied into this repository; outputs are git-ignored but travel with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${PYGIM_REFERENCE_ROOT:-/root/reference}/backend_pim"
OUT="$HERE/_ref"
STUBS="$HERE/ref_stubs"
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
# flags of the reference's default build (spmm_default/CMakeLists.txt:13-19,37-41);
# -fwrapv makes the int32/int64 overflow the DPU code relies on defined on the host.
COMMON="-std=gnu11 -O2 -fPIC -fwrapv -w -DNR_TASKLETS=16 -DPIM_SEQREAD_CACHE_SIZE=32 -I$STUBS"
# per-variant default -D sets, as in each variant's CMakeLists.txt (IF(NOT DEFINED extra_def) blocks)
DEF_DEFAULT="-DBLNC_NNZ=1 -DBLNC_NNZ_RGRN=0 -DBLNC_ROW=0 -DBLNC_TSKLT_ROW=0 -DBLNC_TSKLT_NNZ_RGRN=0 -DBLNC_TSKLT_NNZ=1 \
 -DLOCKFREEV2=1 -DCG_LOCK=0 -DROW_MERGE=1 -DBLOCK_MERGE=0 -DSYNC=1"   # spmm_default/CMakeLists.txt:13-19
DEF_GRANDE="-DBLNC_NNZ=0 -DBLNC_NNZ_RGRN=1 -DBLNC_ROW=0 -DBLNC_TSKLT_ROW=0 -DBLNC_TSKLT_NNZ_RGRN=0 -DBLNC_TSKLT_NNZ=1 \
 -DLOCKFREEV2=1 -DCG_LOCK=0 -DROW_MERGE=0 -DBLOCK_MERGE=1 -DSYNC=0"   # spmm_grande/CMakeLists.txt:14-19
DEF_SPMV="-DBLNC_NNZ=1 -DBLNC_NNZ_RGRN=0 -DBLNC_ROW=0 -DBLNC_TSKLT_ROW=0 -DBLNC_TSKLT_NNZ_RGRN=0 -DBLNC_TSKLT_NNZ=1 \
 -DROW_MERGE=1 -DBLOCK_MERGE=0 -DSYNC=1"                              # spmv_sparseP/CMakeLists.txt:12-17
for DT in INT8 INT16 INT32 INT64 FLT32 DBL64; do
  D="$REF/spmm_default"
  gcc $COMMON $DEF_DEFAULT -D$DT=1 -I"$D" -shared -o "$OUT/libref_default_$DT.so" \
      "$D/spmm_mul_coo.c" "$D/spmm_mul_csr.c" "$D/support/partition.c" "$D/support/timer.c" "$STUBS/dpu_stubs.c"
  G="$REF/spmm_grande"
  gcc $COMMON $DEF_GRANDE -D$DT=1 -I"$G" -shared -o "$OUT/libref_grande_$DT.so" \
      "$G/spmm_mul_csr.c" "$G/support/partition.c" "$G/support/timer.c" "$STUBS/dpu_stubs.c"
  V="$REF/spmv_sparseP"
  gcc $COMMON $DEF_SPMV -D$DT=1 -I"$V" -shared -o "$OUT/libref_spmv_$DT.so" \
      "$V/spmv_mul_coo.c" "$V/support/partition.c" "$V/support/timer.c" "$STUBS/dpu_stubs.c"
done
# the row-balanced partitioners are compiled out of the default build (BLNC_ROW=0): a seventh library with them
D="$REF/spmm_default"
gcc $COMMON -DINT32=1 -DBLNC_ROW=1 -DBLNC_NNZ=1 -DBLNC_NNZ_RGRN=1 -DBLNC_TSKLT_ROW=1 -DBLNC_TSKLT_NNZ=1 \
    -DBLNC_TSKLT_NNZ_RGRN=1 -I"$D" -shared -o "$OUT/libref_partition.so" "$D/support/partition.c"
echo "built $(ls "$OUT" | wc -l) reference host-oracle libraries in $OUT"
