#!/bin/bash
# SASS census of the built library (run anywhere nvcc's cuobjdump is on PATH; no GPU needed):
#   tools/sass_census.sh > profiles/r2_sass_census.txt
# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk
# (TMA bulk copies without a tensor map), UTMALDG / UTMASTG = tensor-map TMA (not used), UCGABAR = cluster barriers,
# SYNCS = mbarrier operations, LDGSTS = cp.async, MATCH = match.any (radix ranks), REDUX = warp reductions.
set -e
SO="$(dirname "$0")/../point-of-interest-recommendation_b200/libpoi_b200.so"
T=$(mktemp)
cuobjdump -sass "$SO" > "$T"
echo "library: $(basename "$SO")  $(stat -c %s "$SO") bytes  source hash $(cat "$SO.srchash" 2>/dev/null)"
echo "kernels: $(grep -c '^\s*Function :' "$T")"
for m in UTCHMMA LDTM STTM UTCBAR UBLKCP UTMALDG UTMASTG UCGABAR SYNCS LDGSTS MATCH REDUX; do
  printf "%-10s %s\n" "$m" "$(grep -c "$m" "$T")"
done
echo "--- kernels containing UTCHMMA (tcgen05.mma) ---"
awk '/Function :/{f=$3} /UTCHMMA/{c[f]++} END{for(k in c) printf "%6d  %s\n", c[k], k}' "$T" | sort -k2 | c++filt | cut -c1-150
rm -f "$T"
