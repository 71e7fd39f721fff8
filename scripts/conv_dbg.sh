for d in 61 125; do PDEQX_CONV_DEBUG=$d python scripts/run_conv.py 2>&1 | sed "s/^/dbg=$d /" | cut -c1-120; done
