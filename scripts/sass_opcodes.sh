#!/bin/bash
# Architectural evidence, generated in the build container (no GPU needed): per-kernel histogram of the Blackwell-only
# SASS opcodes in the shipped library.  UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store,
# UBLKCP = bulk copy, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit.   Usage: bash scripts/sass_opcodes.sh [tag]
set -eu
cd "$(dirname "$0")/.."
TAG=${1:-r2}
LIB=polyphonicformer_b200/lib/libpf_decoder.so
OUT=profiles/${TAG}_sass_opcodes.txt
{
  echo "# $(basename $LIB): cuobjdump -sass, per-kernel count of Blackwell (sm_100a) opcodes; built by polyphonicformer_b200/build.py"
  echo "# nvcc: $(nvcc --version | tail -1)"
  printf "%-46s %8s %8s %8s %7s %6s %6s %7s %6s\n" kernel UTCHMMA UTMALDG UTMASTG UBLKCP LDTM STTM UTCBAR HMMA
  cuobjdump -sass $LIB | awk '
    /Function :/ { if (name != "") emit(); name=$3; split("", c) }
    { for (op in ops) if (index($0, op)) c[op]++ }
    BEGIN { ops["UTCHMMA"]; ops["UTMALDG"]; ops["UTMASTG"]; ops["UBLKCP"]; ops["LDTM"]; ops["STTM"]; ops["UTCBAR"]; ops[" HMMA"] }
    function emit() { printf "%s %d %d %d %d %d %d %d %d\n", name, c["UTCHMMA"], c["UTMALDG"], c["UTMASTG"], c["UBLKCP"], c["LDTM"], c["STTM"], c["UTCBAR"], c[" HMMA"] }
    END { if (name != "") emit() }' | while read -r line; do
      set -- $line; n=$(echo "$1" | c++filt 2>/dev/null | sed 's/^void //; s/(.*//' | tr -d ' ' | cut -c1-46); shift; printf "%-46s %8s %8s %8s %7s %6s %6s %7s %6s\n" "$n" "$@"; done
} > $OUT
cat $OUT
