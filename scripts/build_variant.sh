#!/usr/bin/env bash
# Development aid: build libkge_b200.so of another git ref (branch / commit) into _variants/libkge_<name>.so
# without touching the working tree, so that one GPU run can A/B kernels:
#   scripts/build_variant.sh r2-prep prep
#   KGE_B200_LIB=$PWD/_variants/libkge_prep.so python scripts/kbench.py cfg2 cfg3
# (_variants/*.so is git-ignored but travels to the GPU box with the snapshot.)
set -euo pipefail
ref=${1:?git ref}; name=${2:?variant name}
root=$(git rev-parse --show-toplevel)
wt=$(mktemp -d /tmp/kge_variant.XXXXXX)
trap 'git -C "$root" worktree remove --force "$wt" >/dev/null 2>&1 || rm -rf "$wt"' EXIT
git -C "$root" worktree add --detach "$wt" "$ref" >/dev/null
make -C "$wt/ampligraph_b200/csrc" -j"$(nproc)" >/dev/null
mkdir -p "$root/_variants"
cp "$wt/ampligraph_b200/libkge_b200.so" "$root/_variants/libkge_$name.so"
echo "built $ref -> _variants/libkge_$name.so"
