cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short -s -k "edge_geometry" 2>&1 | grep -E "edge geometry|passed|failed|Error|assert" | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
