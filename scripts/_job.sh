run() { f=$(( ($5<<20) )); echo "L=$1 chi=$2 lc_nt=$5"; QSP_DEBUG_FLAGS=$f PROBE_ENC=0 timeout 300 python scripts/probe.py $1 $2 $3 $4 2 2>&1 | tail -1; }
run 100 64 1 592 0; run 100 64 1 592 256
run 50 16 4 1024 0; run 50 16 4 1024 64; run 50 16 4 1024 32
run 100 32 6 512 0; run 100 32 6 512 256; run 100 32 6 512 512
run 50 8 4 1024 0; run 50 8 4 1024 32
