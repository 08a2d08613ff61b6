#!/bin/bash
# Is the device code of the working tree identical to that of a commit?  usage: scripts/sass_compare.sh <commit>
# Builds the commit in a temporary clone and compares the SASS of every object file, instruction by
# instruction (addresses stripped). Used to prove that comment / #ifdef-guard / test-only edits of the
# CUDA sources leave the GPU-verified machine code untouched when no GPU is at hand.
set -e
C=${1:?commit}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP=$(mktemp -d)
git clone -q "$ROOT" "$TMP/r" && cd "$TMP/r" && git checkout -q "$C"
python kafka-lag-based-assignor_b200/_build.py > /dev/null
python "$ROOT/kafka-lag-based-assignor_b200/_build.py" > /dev/null
h() { cuobjdump -sass "$1" | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\///' | md5sum | cut -c1-16; }
rc=0
for f in kafka-lag-based-assignor_b200/build/*.o; do
  n=$(basename "$f"); a=$(h "$f"); b=$(h "$ROOT/kafka-lag-based-assignor_b200/build/$n")
  if [ "$a" = "$b" ]; then echo "$n SAME"; else echo "$n DIFFERENT"; rc=1; fi
done
rm -rf "$TMP"
exit $rc
