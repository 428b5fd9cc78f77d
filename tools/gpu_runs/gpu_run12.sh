cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --tb=short -x -k "backward or training" 2>&1 | tail -25
