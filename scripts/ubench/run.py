"""Build and run the micro-benchmarks (needs a B200):  python scripts/ubench/run.py"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(here, "ubench")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                       os.path.join(here, "ubench.cu"), "-o", exe])
if "--build-only" not in sys.argv:
    sys.exit(subprocess.call([exe]))
