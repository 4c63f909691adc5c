# A/B the per-layer forward timing of two builds on the same box: tools/ab_layers.sh <old.so>
for r in 1 2; do
  echo "== new (run $r)"; python tools/layer_times.py 6 128 | grep -E "conv:(ec1|ec2|ec3|dc3|dc5|dc6|ec6|dc4) |total"
  echo "== old (run $r)"; SEUNET_LIB_PATH=$1 python tools/layer_times.py 6 128 | grep -E "conv:(ec1|ec2|ec3|dc3|dc5|dc6|ec6|dc4) |total"
done
