python -c "import __graft_entry__ as g; g.build()"
for b in 1 2 8; do timeout 300 python tools/graph_step.py $b 128 2>&1 | tail -3; done
