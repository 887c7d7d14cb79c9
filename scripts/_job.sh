run() { f=$(( ($2<<8) )); echo "chi=$1 rc_nt=$2"; QSP_DEBUG_FLAGS=$f timeout 200 python scripts/probe.py 50 $1 4 1024 2 2>&1 | tail -1; }
run 8 0; run 8 128; run 12 0; run 12 64; run 12 256; run 16 0; run 24 0; run 24 128; run 48 0; run 48 256
