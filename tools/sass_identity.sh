#!/bin/bash
# usage: tools/sass_identity.sh <git-ref>
# Proves that a refactoring did not change the machine code of the DEFAULT kernels: builds the
# csrc/ of <git-ref> into a temp dir and compares the normalised SASS (addresses and encodings
# stripped) of every kernel entry that exists in both builds.  Differences in register naming
# are reported as DIFF - inspect them with cuobjdump before trusting the new build unvalidated.
set -e
ref=${1:?git ref of the validated build}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/csrc
for f in $(git -C $root ls-tree --name-only $ref x_clip_b200/csrc/ | grep -E "\.(cu|cuh|h)$"); do
  git -C $root show $ref:$f > $tmp/csrc/$(basename $f)
done
sed -i "s#\"../../include/xclip_b200.h\"#\"$root/include/xclip_b200.h\"#" $tmp/csrc/host.h
(cd $tmp/csrc && for f in *.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -c $f -o ${f%.cu}.o &
done; wait)
python -m x_clip_b200.build > /dev/null
norm() { cuobjdump -sass -fun "$2" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s#^\s+/\*[0-9a-f]+\*/##' | md5sum | cut -c1-12; }
for o in $tmp/csrc/*.o; do
  b=$(basename $o); n=$root/x_clip_b200/csrc/build/$b
  [ -f $n ] || { echo "MISSING $b"; continue; }
  for fn in $(cuobjdump -sass $o | grep "Function :" | sed 's/.*Function : //'); do
    a=$(norm $o $fn); c=$(norm $n $fn)
    if [ "$a" == "$c" ]; then echo "SAME $b $fn"; else echo "DIFF $b $fn"; fi
  done
done | sort | uniq -c | awk '{print $2, $3, substr($4,1,70)}' | sort | awk '{c[$1]++; print} END {for (k in c) print k": "c[k] > "/dev/stderr"}'
rm -rf $tmp
