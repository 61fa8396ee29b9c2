for W in 1 2 3 4 6; do echo "FNST_WGRAD_WAVES_X2=$W"; FNST_WGRAD_WAVES_X2=$W python tools/prof_train_parts.py 2>&1 | grep -E "net bwd graph|^total|backward"; done
