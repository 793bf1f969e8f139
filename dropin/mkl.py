"""Stand-in for the `mkl` service module yue.py imports unconditionally (yue.py:9, 76).  Only the two
calls the reference makes are provided; BLAS threading is irrelevant once the hot path is on the GPU."""
import os


def get_max_threads():
    return os.cpu_count() or 1


def set_num_threads(n):
    return None
