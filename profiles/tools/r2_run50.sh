#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
