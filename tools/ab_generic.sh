# usage: tools/ab_generic.sh "<grep regex>" lib1.so lib2.so ...   (first entry "default" = in-tree library)
pat=$1; shift
for lib in "$@"; do
  echo "== $lib"
  if [ "$lib" = default ]; then python tools/layer_times.py 6 128 | grep -E "$pat"; else SEUNET_LIB_PATH=$lib python tools/layer_times.py 6 128 | grep -E "$pat"; fi
done
