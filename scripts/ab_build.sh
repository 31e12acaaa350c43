#!/bin/bash
# Build the library from a given git revision's csrc (or the working tree with "wt") into gpurun_out-free scratch
# build/ab/<name>/libreductive_b200.so for A/B timing on the same GPU box:  scripts/ab_build.sh <name> <rev|wt> [extra nvcc flags]
set -e
name=$1; rev=$2
root=$(cd "$(dirname "$0")/.." && pwd)
dst=$root/reductive_b200/lib_ab/$name
rm -rf /tmp/ab_$name && mkdir -p /tmp/ab_$name/reductive_b200 /tmp/ab_$name/include $dst
if [ "$rev" = "wt" ]; then cp -r $root/reductive_b200/csrc /tmp/ab_$name/reductive_b200/; cp $root/include/*.h /tmp/ab_$name/include/
else (cd $root && git archive $rev reductive_b200/csrc include | tar -x -C /tmp/ab_$name); fi
make -s -C /tmp/ab_$name/reductive_b200/csrc -j8 EXTRA="$3" >/dev/null 2>&1
cp /tmp/ab_$name/reductive_b200/lib/libreductive_b200.so $dst/
echo built $dst
