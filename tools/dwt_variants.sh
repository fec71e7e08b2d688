for v in "" _m4 _m5 _m6 _m8 _w1m16; do
  export GB200_LIB=$PWD/grokimagecompression_b200/libgrok_b200$v.so
  echo "== variant '$v'"
  python tools/dwt_bench.py c2 10 rows=0,unroll=2 rows=0,unroll=1 rows=32,unroll=1 2>&1 | grep variant | cut -c1-260
  python tools/dwt_bench.py c3 5 rows=0,unroll=2 rows=0,unroll=1 2>&1 | grep variant | cut -c1-260
done
