# usage: tools/dwt_variants.sh "<suffixes>" <workload> <steps> <variant args...>   (A/B of library builds made by tools/build_variant.sh)
sfx=$1; shift
for v in $sfx; do
  if [ "$v" == "base" ]; then export GB200_LIB=$PWD/grokimagecompression_b200/libgrok_b200.so; else export GB200_LIB=$PWD/grokimagecompression_b200/libgrok_b200_$v.so; fi
  echo "== build '$v'"
  python tools/dwt_bench.py "$@" 2>&1 | grep variant | cut -c1-230
done
