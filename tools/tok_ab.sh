# A/B of tokenizer builds on the C2 table (tuning aid): bash tools/tok_ab.sh lib1.so lib2.so ...
H=${HITS:-20000000}
for lib in "$@"; do echo "$lib"; JCVI_B200_LIB=$lib python tools/tok_variants.py --child $H 2>&1 | grep RESULT; done
echo "tree"; python tools/tok_variants.py --child $H 2>&1 | grep RESULT
