#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vae_gpu.py -q -m gpu -s --tb=short -p no:cacheprovider "$@" > gpurun_out/vae_tests.log 2>&1
rc=$?
tail -n 40 gpurun_out/vae_tests.log
exit $rc
