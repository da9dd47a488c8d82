#!/bin/bash
echo "### default B=32"; timeout 300 python scripts/debug_graph_vs_eager.py fp32 32 2>&1 | tail -12
echo "### mlp wgrad on main"; PIDM_MLP_WGRAD_MAIN=1 timeout 300 python scripts/debug_graph_vs_eager.py fp32 32 2>&1 | tail -12
echo "### default B=8"; timeout 300 python scripts/debug_graph_vs_eager.py fp32 8 2>&1 | tail -12
