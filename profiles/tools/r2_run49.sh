#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 300 python tests/obs_sweep.py --envs 4096 --variants r4b --steps 12 2>&1 | tail -2
