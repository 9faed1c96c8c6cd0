#!/bin/bash
# last check of the tree as committed: every GPU test, smoke(), one short bench line
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-220
