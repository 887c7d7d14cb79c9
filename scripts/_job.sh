for lib in "" scripts/_alt/libqsp_sort.so; do echo "== ${lib:-default}"; QSP_B200_LIB=$lib PROBE_PROF=1 PROBE_ENC=0 timeout 300 python scripts/probe.py 100 64 4 592 1 2>&1 | tail -2; done
