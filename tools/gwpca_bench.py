"""`python tools/gwpca_bench.py` == `python bench.py --workload gwpca` (the CPU-baseline leg lives in bench.py, the only
script besides tests/ and smoke() that may execute oracle/)."""
import os
import runpy
import sys

sys.argv = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), "--workload", "gwpca"] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
