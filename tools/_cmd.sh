set -u
OUT=gpurun_out; TAG=r2q
(time timeout 600 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_dropin.py -x -q) > $OUT/pytest_multidev_$TAG.log 2>&1; tail -n 15 $OUT/pytest_multidev_$TAG.log
