#!/usr/bin/env bash
# Development aid: the CURRENT tree + scripts/sorted_variant.patch (side-sorted corruptions, four-at-a-time score pass)
# built into _variants/libkge_sorted.so for A/B runs:  KGE_B200_LIB=$PWD/_variants/libkge_sorted.so python scripts/kbench.py ...
set -euo pipefail
root=$(cd "$(dirname "$0")/.." && pwd)
wt=$(mktemp -d /tmp/kge_sorted.XXXXXX)
trap 'rm -rf "$wt"' EXIT
mkdir -p "$wt/ampligraph_b200" && cp -r "$root/ampligraph_b200/csrc" "$wt/ampligraph_b200/" && cp -r "$root/include" "$wt/"
rm -rf "$wt/ampligraph_b200/csrc/_build"
patch -s -d "$wt/ampligraph_b200/csrc" -p0 kge_train.cu < "$root/scripts/sorted_variant.patch"
sed -i 's/aux = 3 \* h->eta_pad \* 4 + 16/aux = 4 * h->eta_pad * 4 + 16/' "$wt/ampligraph_b200/csrc/kge_api.cu"
grep -q "aux = 4 \* h->eta_pad" "$wt/ampligraph_b200/csrc/kge_api.cu"
make -C "$wt/ampligraph_b200/csrc" -j"$(nproc)" > /dev/null
mkdir -p "$root/_variants" && cp "$wt/ampligraph_b200/libkge_b200.so" "$root/_variants/libkge_sorted.so"
echo "built _variants/libkge_sorted.so"
