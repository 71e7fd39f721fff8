timeout 60 python scripts/dbg_wg.py 2>&1 | tail -12
MODE=neumann DIL=2 timeout 60 python scripts/dbg_wg.py 2>&1 | grep "gw rel"
MODE=dirichlet DIL=4 timeout 60 python scripts/dbg_wg.py 2>&1 | grep "gw rel"
