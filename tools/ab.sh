# usage: tools/ab.sh variantA.so variantB.so ...  (files under grokimagecompression_b200/)
for v in "$@"; do
  export GB200_LIB=$PWD/grokimagecompression_b200/$v
  echo "== $v"; bash tools/quick_bench.sh
done
