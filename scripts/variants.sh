#!/usr/bin/env bash
# tuning helper (GPU box): run a command once per built library variant (tvm_b200/lib/libtvm_b200<suffix>.so)
# usage: scripts/variants.sh "<suffix list, '-' = product build>" <command...>
sfx="$1"; shift
for s in $sfx; do
  if [ "$s" = "-" ]; then s=""; fi
  echo "== variant '${s}'"
  TVMB200_LIB_SUFFIX="$s" "$@"
done
