for i in 1 2 3; do python -m pytest tests/test_gpu_composites.py tests/test_gpu_e2e.py -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed|parity:" | tr '\n' ' '; echo " [default $i]"; done
for i in 1 2 3; do DCA_SIDE_STREAM=0 python -m pytest tests/test_gpu_composites.py tests/test_gpu_e2e.py -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed" | tr '\n' ' '; echo " [no side stream $i]"; done
for i in 1 2 3; do DCA_PDL=0 python -m pytest tests/test_gpu_composites.py tests/test_gpu_e2e.py -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed" | tr '\n' ' '; echo " [no pdl $i]"; done
for i in 1 2 3; do DCA_PDL=0 DCA_SIDE_STREAM=0 python -m pytest tests/test_gpu_composites.py tests/test_gpu_e2e.py -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed" | tr '\n' ' '; echo " [no pdl no side $i]"; done
