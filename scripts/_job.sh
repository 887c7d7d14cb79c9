mkdir -p gpurun_out
for lib in "" scripts/_alt/libqsp_nonc.so; do
  echo "== lib=${lib:-default(norm cache)}"
  QSP_B200_LIB=$lib timeout 200 python scripts/jacobi_bench.py 128 128 512 592 2>&1 | tail -1
  QSP_B200_LIB=$lib timeout 200 python scripts/jacobi_bench.py 64 128 512 592 2>&1 | tail -1
  QSP_B200_LIB=$lib PROBE_PROF=1 PROBE_ENC=0 timeout 300 python scripts/probe.py 100 64 1 592 2 2>&1 | tail -3
done
