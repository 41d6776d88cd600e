#!/bin/bash
# Is the device code of the library in this tree the same as at another git revision?  Builds that revision's openlte_b200/csrc in a scratch
# directory, dumps the SASS of both libraries (cuobjdump), normalises the per-translation-unit hash that nvcc puts into the names of
# anonymous-namespace kernels, and compares.  Needs no GPU.
#   usage: tools/device_code_diff.sh <git-rev>        (run from the repository root after `make -C openlte_b200/csrc`)
set -e
rev=${1:?usage: tools/device_code_diff.sh <git-rev>}
tmp=$(mktemp -d)
git archive "$rev" openlte_b200/csrc include | tar x -C "$tmp"
make -C "$tmp/openlte_b200/csrc" -j"$(nproc)" > "$tmp/build.log" 2>&1
norm() { cuobjdump -sass "$1" | sed -E 's/_GLOBAL__N__[0-9a-f]{8}_[0-9]+_[a-z_0-9]+_cu_[0-9a-f]{8}/_ANON_/g' | grep -v '^Fatbin\|^=====\|^$'; }
norm "$tmp/openlte_b200/libopenlte_b200.so" > "$tmp/a.sass"
norm openlte_b200/libopenlte_b200.so > "$tmp/b.sass"
echo "revision $(git rev-parse --short "$rev"): $(grep -c 'Function :' "$tmp/a.sass") kernels, $(wc -l < "$tmp/a.sass") SASS lines, md5 $(md5sum < "$tmp/a.sass" | cut -d' ' -f1)"
echo "this tree ($(git rev-parse --short HEAD)$(git diff --quiet || echo '+uncommitted')): $(grep -c 'Function :' "$tmp/b.sass") kernels, $(wc -l < "$tmp/b.sass") SASS lines, md5 $(md5sum < "$tmp/b.sass" | cut -d' ' -f1)"
if cmp -s "$tmp/a.sass" "$tmp/b.sass"; then echo "device code: IDENTICAL"; else echo "device code: DIFFERS"; diff "$tmp/a.sass" "$tmp/b.sass" | head -20; fi
rm -rf "$tmp"
