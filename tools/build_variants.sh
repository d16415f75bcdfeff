#!/bin/bash
# Build tokenizer variants of libjcvi_b200.so into build/variants/ (tuning aid; see tools/tok_variants.py).
#   tools/build_variants.sh name1 "flags1" name2 "flags2" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/variants
mkdir -p "$OUT"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  tmp=$(mktemp -d)
  mkdir -p "$tmp/jcvi_b200" "$tmp/include"
  cp -r "$ROOT/jcvi_b200/csrc" "$tmp/jcvi_b200/csrc"
  cp "$ROOT/include/"*.h "$tmp/include/"
  rm -f "$tmp/jcvi_b200/csrc/"*.o
  ( cd "$tmp/jcvi_b200/csrc" && make -s -j8 ../libjcvi_b200.so EXTRA="$flags" 2>&1 | grep -v "warning\|^ *\^\|detected during\|^$\|rd(\|uint32_t c = " || true )
  cp "$tmp/jcvi_b200/libjcvi_b200.so" "$OUT/lib_$name.so"
  cuobjdump --dump-resource-usage "$OUT/lib_$name.so" 2>/dev/null | grep -A1 "k_tokenizeILi0" | grep -o "REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*" | paste - - - | sed "s/^/$name: /"
  rm -rf "$tmp"
done
