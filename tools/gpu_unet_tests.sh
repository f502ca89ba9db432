#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_unet_gpu.py -q -m gpu -s --tb=short -p no:cacheprovider "$@" > gpurun_out/unet_tests.log 2>&1
rc=$?
tail -n 40 gpurun_out/unet_tests.log
exit $rc
