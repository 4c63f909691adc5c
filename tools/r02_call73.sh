set -x
python -c "import __graft_entry__ as g; g.build()"
for v in 4 8 16; do
  echo "== SEUNET_BWDB_VPT=$v"
  for b in 8 1; do SEUNET_BWDB_VPT=$v timeout 300 python tools/time_train.py $b 128 2 2>&1 | head -2; done
done
