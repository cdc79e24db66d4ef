mkdir -p gpurun_out
for i in 1 2; do timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3; done
python -c "import __graft_entry__ as g; g.smoke()"
