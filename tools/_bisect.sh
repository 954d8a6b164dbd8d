for d in 0 1 2 4 16 32 64 128 96 224 225; do echo "DEBUG=$d"; JYUTVOICE_B200_DEBUG=$d timeout 60 python tools/bench_gemm.py 38656 "conv     N=256  K=3x256 LN1+Mish" 20 2>&1 | tail -2; done
