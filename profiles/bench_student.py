"""Development aid (run under gpurun): bench.py's student_t secondary entry alone."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


class Args:
    option = []
    torch_allreduce = False
    no_burst = True


b = bench.Bench(Args())
out = b.run_student(20, 5)
if b.clocks:
    b.clocks.close()
print(json.dumps(out))
