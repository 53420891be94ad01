#!/bin/bash
cd "$(dirname "$0")/.."
echo "== full GPU suite"
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
