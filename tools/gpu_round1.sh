set -x
cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --small --steps 3 --warmup 1 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>&1 | tail -3
