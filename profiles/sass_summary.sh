#!/bin/bash
# SASS evidence of the sm_100a code paths (regenerate after every build):  bash profiles/sass_summary.sh > profiles/sass_summary.txt
# UBLKCP.S.G = cp.async.bulk (1-D TMA) global -> shared; SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK = mbarrier expect_tx / try_wait;
# VIADDMNMX.U16x2 = the packed far test of the fold; ATOMS = shared-memory atomics; ATOMG/RED = global cursors / counters.
# No UTMALDG / UTC*MMA / LDTM is expected: the path has no 2-D tile and no contraction (integer gather / partition / fold work).
cd "$(dirname "$0")/.."
for lib in bella_b200/libbella_b200.so bella_b200/libbella_xdrop.so bella_b200/libbella_kmers.so; do
  echo "== $lib"
  cuobjdump -lelf "$lib" | sed 's/^/   /'
  cuobjdump -sass "$lib" > /tmp/sass_$$.txt
  for m in UBLKCP SYNCS.ARRIVE.TRANS64 SYNCS.PHASECHK VIADDMNMX.U16x2 "ATOMS" "ATOMG" "RED.E" "REDUX" "MATCH" "SHFL" "LDGSTS" UTMALDG UTCHMMA LDTM HMMA; do
    printf "   %-22s %6d\n" "$m" "$(grep -c "$m" /tmp/sass_$$.txt)"
  done
  echo "   kernels using UBLKCP (function, count):"
  awk '/Function :/ {f=$3} /UBLKCP/ {c[f]++} END {for (k in c) printf "      %s %d\n", k, c[k]}' /tmp/sass_$$.txt | sed 's/_ZN2bk//' | sort | head -12
  rm -f /tmp/sass_$$.txt
done
