python -m pytest tests/test_gpu_lu_pivot.py tests/test_gpu_yago.py -m gpu -x -q 2>&1 | tail -2
python scripts/lu_profile.py 2>&1 | tail -1 | cut -c1-200
