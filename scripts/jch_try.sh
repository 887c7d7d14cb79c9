for ch in 2 4 8 16; do
  sed -i "s/#define QSP_JCH [0-9]* /#define QSP_JCH $ch /" qsp_b200/csrc/qsp_jacobi.cuh
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC qsp_b200/csrc/qsp_capi.cu -o qsp_b200/libqsp_b200.so
  echo "JCH=$ch"; timeout 100 python scripts/jacobi_bench.py 128 128 512 2>&1 | tail -1
  PROBE_ENC=0 timeout 200 python scripts/probe.py 100 64 1 148 2 2>&1 | tail -1
done
